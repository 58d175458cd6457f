// Feature-map kernels of the TransFuser++ step that are not GEMM-shaped (all HBM-bound, NHWC bf16, 16-byte
// vector accesses along the channel dimension, fp32 statistics):
//   stem conv (+ImageNet normalisation), BatchNorm statistics/finalise/apply, squeeze-excite, stride-2 parity split,
//   adaptive average pooling into the fusion token matrix, bilinear resize(+add), layout conversion.
// Reference call sites are cited per entry point.
#include "../../include/tfpp.h"
#include "common.cuh"

#include <cstdlib>
#include "se_kernels.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ stem conv
// timm RegNet stem ConvNormAct(in,32,k3,s2,p1) (oracle/regnety.py header) fused with normalize_imagenet
// (team_code/transfuser_utils.py:542-551).  Input NCHW f32, output NHWC bf16 (B,H/2,W/2,32).
// One thread per output pixel, all 32 output channels in registers; weights in shared memory.
template <int CIN>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ in_scale,
                                                        const float* __restrict__ in_shift,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        int act, bf16* __restrict__ out, float* __restrict__ stat_sum,
                                                        float* __restrict__ stat_sq, int B, int H, int W) {
  __shared__ float sw[32 * CIN * 9];
  __shared__ float ssum[32], ssq[32];
  for (int i = threadIdx.x; i < 32 * CIN * 9; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 32) {
    ssum[threadIdx.x] = 0.f;
    ssq[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = idx < total;
  float acc[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) acc[o] = 0.f;
  if (valid) {
    const int ox = static_cast<int>(idx % Wo);
    const int oy = static_cast<int>((idx / Wo) % Ho);
    const int b = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float a = in_scale ? in_scale[c] : 1.f, sft = in_shift ? in_shift[c] : 0.f;
      const float* xp = x + (static_cast<long long>(b) * CIN + c) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 + kx - 1;
          float v = 0.f;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xp + static_cast<long long>(iy) * W + ix) * a + sft;
#pragma unroll
          for (int o = 0; o < 32; ++o) acc[o] = fmaf(v, sw[(o * CIN + c) * 9 + ky * 3 + kx], acc[o]);
        }
      }
    }
  }
  if (stat_sum != nullptr) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      const float v = valid ? acc[o] : 0.f;
      const float s = warp_sum(v), q = warp_sum(v * v);
      if (lane == 0) {
        atomicAdd(&ssum[o], s);
        atomicAdd(&ssq[o], q);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      atomicAdd(stat_sum + threadIdx.x, ssum[threadIdx.x]);
      atomicAdd(stat_sq + threadIdx.x, ssq[threadIdx.x]);
    }
  }
  if (valid && out != nullptr) {
    uint32_t pk[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      float a0 = acc[2 * o], a1 = acc[2 * o + 1];
      if (scale) {
        a0 = a0 * scale[2 * o] + shift[2 * o];
        a1 = a1 * scale[2 * o + 1] + shift[2 * o + 1];
      }
      pk[o] = pack_bf16x2(apply_act(a0, act), apply_act(a1, act));
    }
    uint4* op = reinterpret_cast<uint4*>(out + idx * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) op[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// nn.BatchNorm2d training semantics (SURVEY.md §8a'): normalise with the biased batch variance, update running
// stats with the unbiased one, momentum 0.1.  sum / sumsq come from the producing conv's epilogue.
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sq,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd, int C, float count, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = sum[c] / count;
  float var = sq[c] / count - mean * mean;
  var = fmaxf(var, 0.f);
  const float invstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - mean * g * invstd;
  if (save_mean) save_mean[c] = mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) {
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// y = act(x * scale[c] + shift[c] (+ res)); optional per-(b,c) sums of y for squeeze-excite.
// grid (chunks, B); each thread owns 8 consecutive channels of a pixel (one 16-byte access).
__global__ void __launch_bounds__(256) scale_shift_act_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const float* __restrict__ res_scale,
                                                              const float* __restrict__ res_shift, int act,
                                                              bf16* __restrict__ y, float* __restrict__ pool_sum,
                                                              int HW, int C, int pix_per_block) {
  extern __shared__ float spool[];
  const int b = blockIdx.y;
  const int c8n = C / 8;
  if (pool_sum != nullptr) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) spool[i] = 0.f;
    __syncthreads();
  }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const long long base = static_cast<long long>(b) * HW * C;
  // a thread owns one 8-channel group and walks pixels: per-thread constants, register accumulation of the SE squeeze
  const int rows_pp = blockDim.x / c8n;
  const int cg = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  if (prow < rows_pp) {
    const int c0 = cg * 8;
    float sc[8], sh[8], rsc[8], rsh[8], pl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = scale ? __ldg(scale + c0 + j) : 1.f;
      sh[j] = scale ? __ldg(shift + c0 + j) : 0.f;
      rsc[j] = res_scale ? __ldg(res_scale + c0 + j) : 1.f;
      rsh[j] = res_scale ? __ldg(res_shift + c0 + j) : 0.f;
      pl[j] = 0.f;
    }
    // U pixel rows per trip: all loads are issued before any is consumed (2*U 16-byte requests in flight per thread),
    // which is what it takes to cover HBM latency at ~512 resident threads per SM
    constexpr int U = 4;
    for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
      uint4 u[U], r[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px < p1) {
          const long long off = base + static_cast<long long>(px) * C + c0;
          u[k] = ld_stream16(x + off);
          if (res != nullptr) r[k] = ld_stream16(res + off);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px >= p1) break;
        const long long off = base + static_cast<long long>(px) * C + c0;
        const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          v[2 * j] = f.x * sc[2 * j] + sh[2 * j];
          v[2 * j + 1] = f.y * sc[2 * j + 1] + sh[2 * j + 1];
        }
        if (res != nullptr) {
          const uint32_t rw[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(rw[j]);
            v[2 * j] += f.x * rsc[2 * j] + rsh[2 * j];
            v[2 * j + 1] += f.y * rsc[2 * j + 1] + rsh[2 * j + 1];
          }
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = pack_bf16x2(apply_act(v[2 * j], act), apply_act(v[2 * j + 1], act));
          if (pool_sum != nullptr) {  // SE squeezes the tensor the next layer will actually read (bf16-rounded)
            const float2 f = unpack_bf16x2(o[j]);
            pl[2 * j] += f.x;
            pl[2 * j + 1] += f.y;
          }
        }
        *reinterpret_cast<uint4*>(y + off) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (pool_sum != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&spool[c0 + j], pl[j]);
    }
  }
  if (pool_sum != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(pool_sum + static_cast<long long>(b) * C + i, spool[i]);
  }
}

// Streaming variant (default): the per-channel constants live in shared memory instead of 32 registers per thread, so
// 8 x 16-byte loads are in flight per thread without spills; the grid is one wave of equal-work CTAs (two per SM).
__device__ __forceinline__ void lds8(const float* p, float* v) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a));
}

template <int U>
__global__ void __launch_bounds__(256, 2) scale_shift_act_s_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                                   const float* __restrict__ scale,
                                                                   const float* __restrict__ shift,
                                                                   const float* __restrict__ res_scale,
                                                                   const float* __restrict__ res_shift, int act,
                                                                   bf16* __restrict__ y, float* __restrict__ pool_sum,
                                                                   int HW, int C, int pix_per_block) {
  extern __shared__ __align__(16) float sm[];  // scale | shift | res_scale | res_shift | pool, C floats each
  float* csc = sm;
  float* csh = sm + C;
  float* crs = sm + 2 * C;
  float* crh = sm + 3 * C;
  float* spool = sm + 4 * C;
  const int b = blockIdx.y;
  const int c8n = C / 8;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    csc[i] = scale ? __ldg(scale + i) : 1.f;
    csh[i] = scale ? __ldg(shift + i) : 0.f;
    if (res != nullptr) {
      crs[i] = res_scale ? __ldg(res_scale + i) : 1.f;
      crh[i] = res_scale ? __ldg(res_shift + i) : 0.f;
    }
    if (pool_sum != nullptr) spool[i] = 0.f;
  }
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const long long base = static_cast<long long>(b) * HW * C;
  const int rows_pp = blockDim.x / c8n;
  const int cg = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  if (prow < rows_pp) {
    const int c0 = cg * 8;
    float pl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pl[j] = 0.f;
    for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
      uint4 u[U], r[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px < p1) {
          const long long off = base + static_cast<long long>(px) * C + c0;
          u[k] = ld_stream16(x + off);
          if (res != nullptr) r[k] = ld_stream16(res + off);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px >= p1) break;
        const long long off = base + static_cast<long long>(px) * C + c0;
        const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
        float v[8], t[8], h[8];
        lds8(csc + c0, t);
        lds8(csh + c0, h);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          v[2 * j] = f.x * t[2 * j] + h[2 * j];
          v[2 * j + 1] = f.y * t[2 * j + 1] + h[2 * j + 1];
        }
        if (res != nullptr) {
          const uint32_t rw[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
          lds8(crs + c0, t);
          lds8(crh + c0, h);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(rw[j]);
            v[2 * j] += f.x * t[2 * j] + h[2 * j];
            v[2 * j + 1] += f.y * t[2 * j + 1] + h[2 * j + 1];
          }
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = pack_bf16x2(apply_act(v[2 * j], act), apply_act(v[2 * j + 1], act));
          if (pool_sum != nullptr) {  // SE squeezes the tensor the next layer will actually read (bf16-rounded)
            const float2 f = unpack_bf16x2(o[j]);
            pl[2 * j] += f.x;
            pl[2 * j + 1] += f.y;
          }
        }
        *reinterpret_cast<uint4*>(y + off) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (pool_sum != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&spool[c0 + j], pl[j]);
    }
  }
  if (pool_sum != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(pool_sum + static_cast<long long>(b) * C + i, spool[i]);
  }
}

// y[b,p,c] = x[b,p,c] * gate[b,c]
__global__ void __launch_bounds__(256) channel_scale_kernel(const bf16* __restrict__ x, const float* __restrict__ gate,
                                                            bf16* __restrict__ y, long long total8, int HW, int C) {
  constexpr int U = 4;  // 16-byte groups per thread, strided by the block so every request stays coalesced
  const long long i0 = static_cast<long long>(blockIdx.x) * (blockDim.x * U) + threadIdx.x;
  const int c8n = C / 8;
  uint4 u[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const long long i = i0 + static_cast<long long>(k) * blockDim.x;
    if (i < total8) u[k] = ld_stream16(x + i * 8);
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const long long i = i0 + static_cast<long long>(k) * blockDim.x;
    if (i >= total8) break;
    const unsigned iu = static_cast<unsigned>(i);   // total8 < 2^31 (checked on the host): 32-bit divisions
    const int c0 = static_cast<int>(iu % static_cast<unsigned>(c8n)) * 8;
    const int b = static_cast<int>((iu / static_cast<unsigned>(c8n)) / static_cast<unsigned>(HW));
    const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
    uint32_t o[4];
    const float4* g = reinterpret_cast<const float4*>(gate + static_cast<long long>(b) * C + c0);
    const float4 g0 = __ldg(g), g1 = __ldg(g + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      o[j] = pack_bf16x2(f.x * gg[2 * j], f.y * gg[2 * j + 1]);
    }
    *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// (B,H,W,C) -> (4B,H/2,W/2,C): plane q = (y&1)*2 + (x&1) stored at batch q*B + b.  Turns stride-2 3x3 / 1x1
// convolutions into stride-1 tap convolutions on parity planes (DESIGN.md "stride-2").
__global__ void __launch_bounds__(256) parity_split_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                           long long total8, int B, int H, int W, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const unsigned c8n = C / 8;
  const unsigned iu = static_cast<unsigned>(i);   // total8 < 2^31 (checked on the host): 32-bit divisions
  const int c8 = static_cast<int>(iu % c8n);
  unsigned pix = iu / c8n;
  const int xx = static_cast<int>(pix % static_cast<unsigned>(W));
  pix /= static_cast<unsigned>(W);
  const int yy = static_cast<int>(pix % static_cast<unsigned>(H));
  const int b = static_cast<int>(pix / static_cast<unsigned>(H));
  const int q = (yy & 1) * 2 + (xx & 1);
  const long long dst = (((static_cast<long long>(q) * B + b) * (H / 2) + (yy >> 1)) * (W / 2) + (xx >> 1)) * C + c8 * 8;
  *reinterpret_cast<uint4*>(y + dst) = *reinterpret_cast<const uint4*>(x + i * 8);
}

// nn.AdaptiveAvgPool2d to (ph,pw) with divisible windows (transfuser.py:36,57,230-231) written straight into the
// fusion token matrix (transfuser.py:317-325): out[b, row0 + py*pw + px, c] = mean + pos_emb[row0 + ..., c].
// One warp per (token, 256-channel slab): lanes own 8 channels each.
__global__ void __launch_bounds__(256) avgpool_tokens_kernel(const bf16* __restrict__ x, const float* __restrict__ pos,
                                                             void* __restrict__ out, int out_f32, int B, int H, int W,
                                                             int C, int ph, int pw, int rows_per_batch, int row0) {
  const int c8n = C / 8;
  const long long total = static_cast<long long>(B) * ph * pw * c8n;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % c8n) * 8;
  long long t = i / c8n;
  const int px = static_cast<int>(t % pw);
  t /= pw;
  const int py = static_cast<int>(t % ph);
  const int b = static_cast<int>(t / ph);
  const int wh = H / ph, ww = W / pw;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int dy = 0; dy < wh; ++dy) {
    for (int dx = 0; dx < ww; ++dx) {
      const long long off = ((static_cast<long long>(b) * H + py * wh + dy) * W + px * ww + dx) * C + c0;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + off));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
  const float inv = 1.f / static_cast<float>(wh * ww);
  const int row = row0 + py * pw + px;
  const long long o = (static_cast<long long>(b) * rows_per_batch + row) * C + c0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] *= inv;
    if (pos) acc[j] += __ldg(pos + static_cast<long long>(row) * C + c0 + j);
  }
  if (out_f32) {
    float4* op = reinterpret_cast<float4*>(static_cast<float*>(out) + o);
    op[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    op[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else {
    *reinterpret_cast<uint4*>(static_cast<bf16*>(out) + o) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                   pack_bf16x2(acc[6], acc[7]));
  }
}

// F.interpolate(mode='bilinear', align_corners=False) (SURVEY.md §8a'): src = (dst+0.5)*in/out-0.5 clamped at 0,
// neighbour clamped at in-1.  out = (add ? add : 0) + resize(src).  src is (B, sh, sw, C) with arbitrary batch /
// row strides (so it can be a slab of the token matrix); add/out are NHWC bf16 (B, dh, dw, C).
__global__ void __launch_bounds__(256) bilinear_kernel(const void* __restrict__ src, int src_f32, long long s_sb,
                                                       long long s_srow, const bf16* __restrict__ add,
                                                       bf16* __restrict__ out, int B, int sh, int sw, int dh, int dw,
                                                       int C) {
  const unsigned c8n = C / 8;
  const unsigned total = static_cast<unsigned>(B) * dh * dw * c8n;   // < 2^31 (checked on the host): 32-bit index math
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % c8n) * 8;
  unsigned t = i / c8n;
  const int x = static_cast<int>(t % static_cast<unsigned>(dw));
  t /= static_cast<unsigned>(dw);
  const int y = static_cast<int>(t % static_cast<unsigned>(dh));
  const int b = static_cast<int>(t / static_cast<unsigned>(dh));
  const float fy = fmaxf((y + 0.5f) * (static_cast<float>(sh) / dh) - 0.5f, 0.f);
  const float fx = fmaxf((x + 0.5f) * (static_cast<float>(sw) / dw) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(fy), sh - 1), x0 = min(static_cast<int>(fx), sw - 1);
  const int y1 = min(y0 + 1, sh - 1), x1 = min(x0 + 1, sw - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  float v[8];
  auto tap = [&](int yy, int xx, float wgt, bool first) {
    const long long off = b * s_sb + (static_cast<long long>(yy) * sw + xx) * s_srow + c0;
    float f[8];
    if (src_f32) {
      const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(src) + off);
      const float4 a = __ldg(p), c = __ldg(p + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
    } else {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const bf16*>(src) + off));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 g = unpack_bf16x2(w[j]);
        f[2 * j] = g.x;
        f[2 * j + 1] = g.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = first ? wgt * f[j] : fmaf(wgt, f[j], v[j]);
  };
  tap(y0, x0, w00, true);
  tap(y0, x1, w01, false);
  tap(y1, x0, w10, false);
  tap(y1, x1, w11, false);
  const long long o = ((static_cast<long long>(b) * dh + y) * dw + x) * C + c0;
  if (add != nullptr) {
    const uint4 u = *reinterpret_cast<const uint4*>(add + o);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g = unpack_bf16x2(w[j]);
      v[2 * j] += g.x;
      v[2 * j + 1] += g.y;
    }
  }
  *reinterpret_cast<uint4*>(out + o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                   pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// bev_semantic_decoder tail (model.py:88-90,385): bilinear resize of NHWC bf16 (B,sh,sw,Cs) to NCHW f32
// (B,C,dh,dw) times the camera-frustum mask (1,1,dh,dw).
__global__ void __launch_bounds__(256) bilinear_nchw_mask_kernel(const bf16* __restrict__ src,
                                                                 const float* __restrict__ mask,
                                                                 float* __restrict__ out, int B, int sh, int sw, int Cs,
                                                                 int C, int dh, int dw) {
  const long long total = static_cast<long long>(B) * C * dh * dw;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % dw);
  long long t = i / dw;
  const int y = static_cast<int>(t % dh);
  t /= dh;
  const int c = static_cast<int>(t % C);
  const int b = static_cast<int>(t / C);
  const float fy = fmaxf((y + 0.5f) * (static_cast<float>(sh) / dh) - 0.5f, 0.f);
  const float fx = fmaxf((x + 0.5f) * (static_cast<float>(sw) / dw) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(fy), sh - 1), x0 = min(static_cast<int>(fx), sw - 1);
  const int y1 = min(y0 + 1, sh - 1), x1 = min(x0 + 1, sw - 1);
  const float ly = fy - y0, lx = fx - x0;
  auto at = [&](int yy, int xx) { return bf2f(src[((static_cast<long long>(b) * sh + yy) * sw + xx) * Cs + c]); };
  float v = (1.f - ly) * ((1.f - lx) * at(y0, x0) + lx * at(y0, x1)) + ly * ((1.f - lx) * at(y1, x0) + lx * at(y1, x1));
  if (mask) v *= __ldg(mask + static_cast<long long>(y) * dw + x);
  out[i] = v;
}

// NCHW f32 <-> NHWC bf16 (API boundary only: the reference's module surface is NCHW f32).
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                           long long total, int C, int HW) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long p = i / C;
  const long long b = p / HW, hw = p % HW;
  y[i] = f2bf(x[(b * C + c) * HW + hw]);
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const bf16* __restrict__ x, float* __restrict__ y,
                                                           long long total, int C, int HW) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long hw = i % HW;
  const long long t = i / HW;
  const int c = static_cast<int>(t % C);
  const long long b = t / C;
  y[i] = bf2f(x[(b * HW + hw) * C + c]);
}

}  // namespace

#define STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int tfpp_stem_conv(const float* x, const float* w, const float* in_scale, const float* in_shift,
                              const float* scale, const float* shift, int act, void* out, float* stat_sum,
                              float* stat_sq, int batch, int cin, int height, int width, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(cin >= 1 && cin <= 3, "stem conv supports 1..3 input channels");
  TFPP_CHECK_ARG(height % 2 == 0 && width % 2 == 0, "even input size");
  const long long total = static_cast<long long>(batch) * (height / 2) * (width / 2);
  const int blocks = static_cast<int>(ceil_div_ll(total, 128));
  bf16* o = static_cast<bf16*>(out);
  if (cin == 1)
    stem_conv_kernel<1><<<blocks, 128, 0, stream>>>(x, w, in_scale, in_shift, scale, shift, act, o, stat_sum, stat_sq,
                                                    batch, height, width);
  else if (cin == 2)
    stem_conv_kernel<2><<<blocks, 128, 0, stream>>>(x, w, in_scale, in_shift, scale, shift, act, o, stat_sum, stat_sq,
                                                    batch, height, width);
  else
    stem_conv_kernel<3><<<blocks, 128, 0, stream>>>(x, w, in_scale, in_shift, scale, shift, act, o, stat_sum, stat_sq,
                                                    batch, height, width);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bn_finalize(const float* sum, const float* sq, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float* scale, float* shift, float* save_mean,
                                float* save_invstd, int channels, float count, float eps, float momentum,
                                tfpp_stream_t stream_) {
  STREAM;
  bn_finalize_kernel<<<ceil_div(channels, 128), 128, 0, stream>>>(sum, sq, gamma, beta, running_mean, running_var, scale,
                                                                  shift, save_mean, save_invstd, channels, count, eps,
                                                                  momentum);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_scale_shift_act(const void* x, const void* res, const float* scale, const float* shift,
                                    const float* res_scale, const float* res_shift, int act, void* y, float* pool_sum,
                                    int batch, int hw, int channels, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0 && channels <= 2048, "channels must be a multiple of 8, <= 2048");
  TFPP_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  // two full waves of (up to) three resident CTAs per SM, never a partial extra wave
  int chunks = TFPP_NUM_SMS * 6 / batch;
  if (chunks < 1) chunks = 1;
  int pix_per_block = ceil_div(hw, chunks);
  if (pix_per_block < 8) pix_per_block = 8;
  chunks = ceil_div(hw, pix_per_block);
  static const bool stream_on = [] { const char* e = getenv("TFPP_BN_STREAM"); return e == nullptr || e[0] != '0'; }();
  if (stream_on) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(scale_shift_act_s_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 2048 * 4);
      cudaFuncSetAttribute(scale_shift_act_s_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 2048 * 4);
      attr_set = true;
    }
    chunks = TFPP_NUM_SMS * 2 / batch;  // one wave of two resident CTAs per SM
    if (chunks < 1) chunks = 1;
    pix_per_block = ceil_div(hw, chunks);
    if (pix_per_block < 8) pix_per_block = 8;
    chunks = ceil_div(hw, pix_per_block);
    // one input tensor: 8 pixel rows per trip keep the same 128 bytes per thread in flight as 4 rows of two tensors
    auto kern = res == nullptr ? scale_shift_act_s_kernel<8> : scale_shift_act_s_kernel<4>;
    kern<<<dim3(chunks, batch), 256, sizeof(float) * 5 * channels, stream>>>(
        static_cast<const bf16*>(x), static_cast<const bf16*>(res), scale, shift, res_scale, res_shift, act,
        static_cast<bf16*>(y), pool_sum, hw, channels, pix_per_block);
    TFPP_CHECK_LAUNCH();
    return TFPP_OK;
  }
  dim3 grid(chunks, batch);
  const size_t smem = pool_sum ? sizeof(float) * channels : 0;
  scale_shift_act_kernel<<<grid, 256, smem, stream>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(res), scale,
                                                      shift, res_scale, res_shift, act, static_cast<bf16*>(y), pool_sum,
                                                      hw, channels, pix_per_block);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_se_gate(const float* pool_sum, int hw, const float* w1, const float* b1, const float* w2,
                            const float* b2, float* gate, float* hidden, int batch, int channels, int rd,
                            tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(hidden != nullptr && rd <= 2048, "hidden workspace (B,rd) is required; rd <= 2048");
  // gate = sigmoid(fc2(relu(fc1(mean_hw(x))))): two batched contractions (se_kernels.cuh)
  se_contract_c_kernel<SE_FWD><<<dim3(ceil_div(rd, 8), ceil_div(batch, 8)), 256, 0, stream>>>(
      pool_sum, nullptr, 1.f / hw, w1, channels, 1, b1, nullptr, hidden, nullptr, batch, channels, rd);
  TFPP_CHECK_LAUNCH();
  se_contract_r_kernel<SE_FWD><<<dim3(ceil_div(channels, 128), ceil_div(batch, 8)), 128, sizeof(float) * rd * 8, stream>>>(
      hidden, w2, rd, 1, b2, 1.f, gate, batch, channels, rd);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_channel_scale(const void* x, const float* gate, void* y, int batch, int hw, int channels,
                                  tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0, "channels must be a multiple of 8");
  const long long total8 = static_cast<long long>(batch) * hw * channels / 8;
  TFPP_CHECK_ARG(total8 < (1ll << 31), "tensor too large for the 32-bit index path");
  channel_scale_kernel<<<static_cast<int>(ceil_div_ll(total8, 256 * 4)), 256, 0, stream>>>(
      static_cast<const bf16*>(x), gate, static_cast<bf16*>(y), total8, hw, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_parity_split(const void* x, void* y, int batch, int height, int width, int channels,
                                 tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0 && height % 2 == 0 && width % 2 == 0, "need C%8==0 and even H, W");
  const long long total8 = static_cast<long long>(batch) * height * width * channels / 8;
  TFPP_CHECK_ARG(total8 < (1ll << 31), "tensor too large for the 32-bit index path");
  parity_split_kernel<<<static_cast<int>(ceil_div_ll(total8, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), total8, batch, height, width, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_avgpool_tokens(const void* x, const float* pos_emb, void* out, int out_f32, int batch, int height,
                                   int width, int channels, int ph, int pw, int rows_per_batch, int row0,
                                   tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0, "channels must be a multiple of 8");
  TFPP_CHECK_ARG(height % ph == 0 && width % pw == 0, "pooling windows must divide the map");
  const long long total = static_cast<long long>(batch) * ph * pw * (channels / 8);
  avgpool_tokens_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(x), pos_emb, out, out_f32, batch, height, width, channels, ph, pw, rows_per_batch, row0);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bilinear(const void* src, int src_f32, long long src_batch_stride, long long src_row_stride,
                             const void* add, void* out, int batch, int sh, int sw, int dh, int dw, int channels,
                             tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0, "channels must be a multiple of 8");
  const long long total = static_cast<long long>(batch) * dh * dw * (channels / 8);
  TFPP_CHECK_ARG(total < (1ll << 31), "tensor too large for the 32-bit index path");
  bilinear_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      src, src_f32, src_batch_stride, src_row_stride, static_cast<const bf16*>(add), static_cast<bf16*>(out), batch, sh,
      sw, dh, dw, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bilinear_nchw_mask(const void* src, const float* mask, float* out, int batch, int sh, int sw,
                                       int src_channels, int channels, int dh, int dw, tfpp_stream_t stream_) {
  STREAM;
  const long long total = static_cast<long long>(batch) * channels * dh * dw;
  bilinear_nchw_mask_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(src), mask, out, batch, sh, sw, src_channels, channels, dh, dw);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_nchw_f32_to_nhwc_bf16(const float* x, void* y, int batch, int channels, int hw,
                                          tfpp_stream_t stream_) {
  STREAM;
  const long long total = static_cast<long long>(batch) * channels * hw;
  nchw_to_nhwc_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(x, static_cast<bf16*>(y), total,
                                                                                     channels, hw);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_nhwc_bf16_to_nchw_f32(const void* x, float* y, int batch, int channels, int hw,
                                          tfpp_stream_t stream_) {
  STREAM;
  const long long total = static_cast<long long>(batch) * channels * hw;
  nhwc_to_nchw_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(static_cast<const bf16*>(x), y,
                                                                                     total, channels, hw);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
