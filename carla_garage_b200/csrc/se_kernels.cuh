// Squeeze-excite bottleneck MLP (timm SEModule, oracle/regnety.py) as two tiny batched contractions, forward and
// backward.  The work is a few MFLOP; what matters is that the weights are read once (not once per sample) and that
// enough CTAs are in flight to hide latency, so both contractions tile (outputs x samples) over the grid.
#pragma once
#include "common.cuh"

namespace {

enum { SE_FWD = 0, SE_BWD = 1 };

// Long contraction over the C channels:  out[b,r] = epi(sum_c X[b,c] * W[r*w_sr + c*w_sc]),  8 r x 8 b per CTA.
//   SE_FWD: X = pool_sum * x_scale,          W = fc1.weight (R,C),  epi = relu(. + bias[r])
//   SE_BWD: X = dgate_sum * g * (1 - g),     W = fc2.weight (C,R),  epi = . * (hidden[b,r] > 0); X is also written
//           to x_out (the sigmoid-adjoint ds, needed for the fc2 parameter gradients) by the blockIdx.x == 0 column.
template <int MODE>
__global__ void __launch_bounds__(256) se_contract_c_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                            float x_scale, const float* __restrict__ w, int w_sr,
                                                            int w_sc, const float* __restrict__ bias,
                                                            const float* __restrict__ hidden, float* __restrict__ out,
                                                            float* __restrict__ x_out, int B, int C, int R) {
  __shared__ float red[8][64];
  const int r0 = blockIdx.x * 8, b0 = blockIdx.y * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    float xv[8], wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j;
      float v = 0.f;
      if (b < B) {
        const long long o = static_cast<long long>(b) * C + c;
        v = __ldg(x + o);
        if (MODE == SE_BWD) {
          const float gg = __ldg(g + o);
          v = v * gg * (1.f - gg);
          if (blockIdx.x == 0) x_out[o] = v;
        } else {
          v *= x_scale;
        }
      }
      xv[j] = v;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + i;
      wv[i] = r < R ? __ldg(w + static_cast<long long>(r) * w_sr + static_cast<long long>(c) * w_sc) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = warp_sum(acc[i][j]);
      if (lane == 0) red[warp][i * 8 + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < 64) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    const int r = r0 + (threadIdx.x >> 3), b = b0 + (threadIdx.x & 7);
    if (r < R && b < B) {
      if (MODE == SE_FWD) v = fmaxf(v + __ldg(bias + r), 0.f);
      else v = __ldg(hidden + static_cast<long long>(b) * R + r) > 0.f ? v : 0.f;
      out[static_cast<long long>(b) * R + r] = v;
    }
  }
}

// Short contraction over the R bottleneck units:  out[b,c] = epi(sum_r H[b,r] * W[c*w_sc + r*w_sr]), one c per thread,
// 8 samples per CTA (H staged in shared memory).
//   SE_FWD: H = hidden, W = fc2.weight (C,R), epi = sigmoid(. + bias[c])        -> gate
//   SE_BWD: H = dpre,   W = fc1.weight (R,C), epi = . * out_scale (= 1/HW)      -> pool_grad
template <int MODE>
__global__ void __launch_bounds__(128) se_contract_r_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                            int w_sc, int w_sr, const float* __restrict__ bias,
                                                            float out_scale, float* __restrict__ out, int B, int C,
                                                            int R) {
  extern __shared__ float hs[];  // [R][8]
  const int b0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < R * 8; i += blockDim.x) {
    const int r = i >> 3, j = i & 7;
    hs[i] = b0 + j < B ? h[static_cast<long long>(b0 + j) * R + r] : 0.f;
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* wp = w + static_cast<long long>(c) * w_sc;
  // the weight loads are the latency: 16 of them in flight per thread before the first is consumed
  for (int r0 = 0; r0 < R; r0 += 16) {
    float wv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) wv[i] = r0 + i < R ? __ldg(wp + static_cast<long long>(r0 + i) * w_sr) : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (r0 + i < R) {
        const float4 h0 = *reinterpret_cast<const float4*>(hs + (r0 + i) * 8);
        const float4 h1 = *reinterpret_cast<const float4*>(hs + (r0 + i) * 8 + 4);
        acc[0] = fmaf(wv[i], h0.x, acc[0]);
        acc[1] = fmaf(wv[i], h0.y, acc[1]);
        acc[2] = fmaf(wv[i], h0.z, acc[2]);
        acc[3] = fmaf(wv[i], h0.w, acc[3]);
        acc[4] = fmaf(wv[i], h1.x, acc[4]);
        acc[5] = fmaf(wv[i], h1.y, acc[5]);
        acc[6] = fmaf(wv[i], h1.z, acc[6]);
        acc[7] = fmaf(wv[i], h1.w, acc[7]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (b0 + j >= B) break;
    float v = acc[j];
    if (MODE == SE_FWD) v = 1.f / (1.f + __expf(-(v + __ldg(bias + c))));
    else v *= out_scale;
    out[static_cast<long long>(b0 + j) * C + c] = v;
  }
}

}  // namespace
