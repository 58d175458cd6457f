// Backward kernels of the feature-map ops (adjoints of featmap.cu): BatchNorm(+ReLU,+SE gate,+SE squeeze) backward,
// squeeze-excite backward, activation/bias backward, bilinear and average-pool adjoints, stem weight gradient.
// All HBM-bound; NHWC bf16 activations/gradients, fp32 reductions.  They replace the autograd graph torch builds for
// team_code/train.py:898 (loss.backward()) over the modules cited in featmap.cu.
#include "../../include/tfpp.h"
#include "common.cuh"
#include "se_kernels.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

namespace {

__device__ __forceinline__ void load8(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float* v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                            pack_bf16x2(v[6], v[7]));
}

// ---------------------------------------------------------------------------------------------- BatchNorm backward
// forward: y = act(xhat * gamma + beta (+ res)), xhat = (raw - mean) * invstd, then optionally out = y * gate[b,c]
// and pool[b,c] = sum_p y.  Incoming: dy (grad wrt y*gate if gate given), pool_grad[b,c] (grad wrt every y of (b,c)).
//   dz = (dy * gate + pool_grad) * act'(y);  s1[c] = sum dz;  s2[c] = sum dz * xhat
//   draw = gamma * invstd * (dz - s1/N - xhat * s2/N);  dgamma = s2;  dbeta = s1
// pass 1 (reduce) / pass 2 (apply); grid (chunks, B), 8 channels per thread.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                            const bf16* __restrict__ raw, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            const float* __restrict__ gate,
                                                            const float* __restrict__ pool_grad,
                                                            const float* __restrict__ fscale,
                                                            const float* __restrict__ fshift, int act,
                                                            float* __restrict__ s1, float* __restrict__ s2, int HW,
                                                            int C, int pix_per_block) {
  extern __shared__ float sm[];  // [2][C]
  float* a1 = sm;
  float* a2 = sm + C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, c8n = C / 8;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const long long base = static_cast<long long>(b) * HW * C;
  // a thread owns one 8-channel group and walks pixels (blockDim.x / c8n pixels per pass): register accumulation,
  // contiguous 16-byte loads across the block, one shared-memory atomic per channel per thread at the end
  const int rows_pp = blockDim.x / c8n;
  const int cg = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  if (prow < rows_pp) {
    const int c0 = cg * 8;
    // ReLU mask: from the saved output y, or (y == nullptr) recomputed from raw with the forward's folded affine —
    // one tensor less to read
    const bool mask_from_raw = (act == ACT_RELU) && (y == nullptr);
    float l1[8], l2[8], mu[8], is[8], gt[8], pg[8], fs[8], fh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      l1[j] = l2[j] = 0.f;
      mu[j] = __ldg(mean + c0 + j);
      is[j] = __ldg(invstd + c0 + j);
      fs[j] = mask_from_raw ? __ldg(fscale + c0 + j) : 0.f;
      fh[j] = mask_from_raw ? __ldg(fshift + c0 + j) : 0.f;
      gt[j] = gate ? __ldg(gate + static_cast<long long>(b) * C + c0 + j) : 1.f;
      pg[j] = pool_grad ? __ldg(pool_grad + static_cast<long long>(b) * C + c0 + j) : 0.f;
    }
    constexpr int U = 2;  // pixel rows per trip; every load of the trip is in flight before the first is consumed
    for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
      uint4 ud[U], ur[U], uy[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px < p1) {
          const long long off = base + static_cast<long long>(px) * C + c0;
          ud[k] = ld_stream16(dy + off);
          ur[k] = ld_stream16(raw + off);
          if (act == ACT_RELU && !mask_from_raw) uy[k] = ld_stream16(y + off);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (pix + k * rows_pp >= p1) break;
        float d[8], yy[8], r[8];
        unpack8(ud[k], d);
        unpack8(ur[k], r);
        if (act == ACT_RELU && !mask_from_raw) unpack8(uy[k], yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dz = d[j] * gt[j] + pg[j];
          if (mask_from_raw) yy[j] = fmaf(r[j], fs[j], fh[j]);
          if (act == ACT_RELU && !(yy[j] > 0.f)) dz = 0.f;
          l1[j] += dz;
          l2[j] = fmaf(dz, (r[j] - mu[j]) * is[j], l2[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&a1[c0 + j], l1[j]);
      atomicAdd(&a2[c0 + j], l2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(s1 + i, a1[i]);
    atomicAdd(s2 + i, a2[i]);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                           const bf16* __restrict__ raw, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ s1,
                                                           const float* __restrict__ s2, const float* __restrict__ gate,
                                                           const float* __restrict__ pool_grad,
                                                           const float* __restrict__ fscale,
                                                           const float* __restrict__ fshift, int act, float inv_n,
                                                           bf16* __restrict__ draw, bf16* __restrict__ dz_out, int HW,
                                                           int C, int pix_per_block) {
  // same walk as the reduce pass: a thread keeps its 8 channels' constants in registers and streams pixels
  //   draw = k0 * dz + k1 * raw + k2,  k0 = gamma*invstd, k1 = -k0*invstd*s2/N, k2 = -k0*s1/N - k1*mean
  const int b = blockIdx.y, c8n = C / 8;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const long long base = static_cast<long long>(b) * HW * C;
  const int rows_pp = blockDim.x / c8n;
  const int cg = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  if (prow >= rows_pp) return;
  const int c0 = cg * 8;
  const bool mask_from_raw = (act == ACT_RELU) && (y == nullptr);
  float k0[8], k1[8], k2[8], gt[8], pg[8], fs[8], fh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    fs[j] = mask_from_raw ? __ldg(fscale + c) : 0.f;
    fh[j] = mask_from_raw ? __ldg(fshift + c) : 0.f;
    const float is = __ldg(invstd + c);
    k0[j] = __ldg(gamma + c) * is;
    k1[j] = -k0[j] * is * __ldg(s2 + c) * inv_n;
    k2[j] = -k0[j] * __ldg(s1 + c) * inv_n - k1[j] * __ldg(mean + c);
    gt[j] = gate ? __ldg(gate + static_cast<long long>(b) * C + c) : 1.f;
    pg[j] = pool_grad ? __ldg(pool_grad + static_cast<long long>(b) * C + c) : 0.f;
  }
  constexpr int U = 2;
  for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
    uint4 ud[U], ur[U], uy[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int px = pix + k * rows_pp;
      if (px < p1) {
        const long long off = base + static_cast<long long>(px) * C + c0;
        ud[k] = ld_stream16(dy + off);
        ur[k] = ld_stream16(raw + off);
        if (act == ACT_RELU && !mask_from_raw) uy[k] = ld_stream16(y + off);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int px = pix + k * rows_pp;
      if (px >= p1) break;
      const long long off = base + static_cast<long long>(px) * C + c0;
      float d[8], yy[8], r[8], o[8], z[8];
      unpack8(ud[k], d);
      unpack8(ur[k], r);
      if (act == ACT_RELU && !mask_from_raw) unpack8(uy[k], yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dz = d[j] * gt[j] + pg[j];
        if (mask_from_raw) yy[j] = fmaf(r[j], fs[j], fh[j]);
        if (act == ACT_RELU && !(yy[j] > 0.f)) dz = 0.f;
        z[j] = dz;
        o[j] = fmaf(k0[j], dz, fmaf(k1[j], r[j], k2[j]));
      }
      store8(draw + off, o);
      if (dz_out) store8(dz_out + off, z);
    }
  }
}

// ---- streaming variants (default): per-channel constants live in shared memory instead of 48-64 registers per thread,
// so U = 4 pixel rows (8 x 16-byte loads in flight per thread) fit without spills at two resident CTAs per SM, and the
// grid is ONE wave of equal-work CTAs (the constant prologue and the statistics flush are paid once per SM slot).  Same math as the kernels above; the reduce pass accumulates
// sum dz*(raw - mean) and multiplies by invstd once per CTA at the flush.
__device__ __forceinline__ void lds8(const float* p, float* v) {
  // volatile: keeps the compiler from hoisting the constants back into registers
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a));
}

template <int U, int MASK, bool GATE>  // MASK: 0 none, 1 ReLU mask recomputed from raw, 2 ReLU mask from the saved output
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_s_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                                 const bf16* __restrict__ raw,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd,
                                                                 const float* __restrict__ gate,
                                                                 const float* __restrict__ pool_grad,
                                                                 const float* __restrict__ fscale,
                                                                 const float* __restrict__ fshift, int act,
                                                                 float* __restrict__ s1, float* __restrict__ s2, int HW,
                                                                 int C, int pix_per_block) {
  extern __shared__ __align__(16) float sm[];  // a1 | a2 | mean | fscale | fshift | gate | pool_grad, C floats each
  float* a1 = sm;
  float* a2 = sm + C;
  float* cmu = sm + 2 * C;
  float* cfs = sm + 3 * C;
  float* cfh = sm + 4 * C;
  float* cgt = sm + 5 * C;
  float* cpg = sm + 6 * C;
  const int b = blockIdx.y, c8n = C / 8;
  constexpr bool relu = MASK != 0;
  constexpr bool mask_from_raw = MASK == 1;
  constexpr bool has_gate = GATE;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    a1[i] = 0.f;
    a2[i] = 0.f;
    cmu[i] = __ldg(mean + i);
    if (mask_from_raw) {
      cfs[i] = __ldg(fscale + i);
      cfh[i] = __ldg(fshift + i);
    }
    if (has_gate) {
      cgt[i] = gate ? __ldg(gate + static_cast<long long>(b) * C + i) : 1.f;
      cpg[i] = pool_grad ? __ldg(pool_grad + static_cast<long long>(b) * C + i) : 0.f;
    }
  }
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const long long base = static_cast<long long>(b) * HW * C;
  const int rows_pp = blockDim.x / c8n;
  const int cg = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  if (prow < rows_pp) {
    const int c0 = cg * 8;
    float l1[8], l2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) l1[j] = l2[j] = 0.f;
    for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
      uint4 ud[U], ur[U], uy[MASK == 2 ? U : 1];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px < p1) {
          const long long off = base + static_cast<long long>(px) * C + c0;
          ud[k] = ld_stream16(dy + off);
          ur[k] = ld_stream16(raw + off);
          if (MASK == 2) uy[MASK == 2 ? k : 0] = ld_stream16(y + off);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (pix + k * rows_pp >= p1) break;
        float d[8], r[8], t[8];
        unpack8(ud[k], d);
        unpack8(ur[k], r);
        if (has_gate) {
          lds8(cgt + c0, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] *= t[j];
          lds8(cpg + c0, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] += t[j];
        }
        if (mask_from_raw) {
          float h[8];
          lds8(cfs + c0, t);
          lds8(cfh + c0, h);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (!(fmaf(r[j], t[j], h[j]) > 0.f)) d[j] = 0.f;
        } else if (relu) {
          unpack8(uy[MASK == 2 ? k : 0], t);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (!(t[j] > 0.f)) d[j] = 0.f;
        }
        lds8(cmu + c0, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          l1[j] += d[j];
          l2[j] = fmaf(d[j], r[j] - t[j], l2[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&a1[c0 + j], l1[j]);
      atomicAdd(&a2[c0 + j], l2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(s1 + i, a1[i]);
    atomicAdd(s2 + i, a2[i] * __ldg(invstd + i));
  }
}

template <int U, int MASK, bool GATE>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_s_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                                const bf16* __restrict__ raw,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ s1, const float* __restrict__ s2,
                                                                const float* __restrict__ gate,
                                                                const float* __restrict__ pool_grad,
                                                                const float* __restrict__ fscale,
                                                                const float* __restrict__ fshift, int act, float inv_n,
                                                                bf16* __restrict__ draw, bf16* __restrict__ dz_out, int HW,
                                                                int C, int pix_per_block) {
  //   draw = k0 * dz + k1 * raw + k2,  k0 = gamma*invstd, k1 = -k0*invstd*s2/N, k2 = -k0*s1/N - k1*mean
  extern __shared__ __align__(16) float sm[];  // k0 | k1 | k2 | fscale | fshift | gate | pool_grad
  float* ck0 = sm;
  float* ck1 = sm + C;
  float* ck2 = sm + 2 * C;
  float* cfs = sm + 3 * C;
  float* cfh = sm + 4 * C;
  float* cgt = sm + 5 * C;
  float* cpg = sm + 6 * C;
  const int b = blockIdx.y, c8n = C / 8;
  constexpr bool relu = MASK != 0;
  constexpr bool mask_from_raw = MASK == 1;
  constexpr bool has_gate = GATE;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float is = __ldg(invstd + i);
    const float k0 = __ldg(gamma + i) * is;
    const float k1 = -k0 * is * __ldg(s2 + i) * inv_n;
    ck0[i] = k0;
    ck1[i] = k1;
    ck2[i] = -k0 * __ldg(s1 + i) * inv_n - k1 * __ldg(mean + i);
    if (mask_from_raw) {
      cfs[i] = __ldg(fscale + i);
      cfh[i] = __ldg(fshift + i);
    }
    if (has_gate) {
      cgt[i] = gate ? __ldg(gate + static_cast<long long>(b) * C + i) : 1.f;
      cpg[i] = pool_grad ? __ldg(pool_grad + static_cast<long long>(b) * C + i) : 0.f;
    }
  }
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const long long base = static_cast<long long>(b) * HW * C;
  const int rows_pp = blockDim.x / c8n;
  const int cg = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  if (prow >= rows_pp) return;
  const int c0 = cg * 8;
  for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
    uint4 ud[U], ur[U], uy[MASK == 2 ? U : 1];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int px = pix + k * rows_pp;
      if (px < p1) {
        const long long off = base + static_cast<long long>(px) * C + c0;
        ud[k] = ld_stream16(dy + off);
        ur[k] = ld_stream16(raw + off);
        if (MASK == 2) uy[MASK == 2 ? k : 0] = ld_stream16(y + off);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int px = pix + k * rows_pp;
      if (px >= p1) break;
      const long long off = base + static_cast<long long>(px) * C + c0;
      float d[8], r[8], t[8], h[8];
      unpack8(ud[k], d);
      unpack8(ur[k], r);
      if (has_gate) {
        lds8(cgt + c0, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] *= t[j];
        lds8(cpg + c0, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += t[j];
      }
      if (mask_from_raw) {
        lds8(cfs + c0, t);
        lds8(cfh + c0, h);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(fmaf(r[j], t[j], h[j]) > 0.f)) d[j] = 0.f;
      } else if (relu) {
        unpack8(uy[MASK == 2 ? k : 0], t);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(t[j] > 0.f)) d[j] = 0.f;
      }
      if (dz_out) store8(dz_out + off, d);
      lds8(ck1 + c0, t);
      lds8(ck2 + c0, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(t[j], r[j], h[j]);
      lds8(ck0 + c0, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(t[j], d[j], r[j]);
      store8(draw + off, r);
    }
  }
}

// EXPERIMENTAL (TFPP_BN_BWD_FUSED=1, never run on a GPU yet): both passes in ONE cooperative launch for activations
// that fit the 126 MB L2 — reduce, grid barrier, apply.  The second pass then reads from L2 instead of HBM and half of
// the 272 BatchNorm-backward launches of a step disappear (they are latency-bound on the small stage-3/4 maps).
// Work items = (sample, pixel chunk) pairs, the same item -> CTA mapping in both phases.
__global__ void __launch_bounds__(256, 2) bn_bwd_fused_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                              const bf16* __restrict__ raw, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ gate,
                                                              const float* __restrict__ pool_grad,
                                                              const float* __restrict__ fscale,
                                                              const float* __restrict__ fshift, int act, float inv_n,
                                                              float* __restrict__ s1, float* __restrict__ s2,
                                                              bf16* __restrict__ draw, bf16* __restrict__ dz_out, int HW,
                                                              int C, int pix_per_block, int chunks, int B) {
  extern __shared__ float sm[];  // [2][C]
  float* a1 = sm;
  float* a2 = sm + C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int c8n = C / 8;
  const int rows_pp = blockDim.x / c8n;
  const int cgi = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  const int c0 = cgi * 8;
  const bool active = prow < rows_pp;
  const bool mask_from_raw = (act == ACT_RELU) && (y == nullptr);
  const int items = chunks * B;
  constexpr int U = 2;
  float mu[8], is[8], fs[8], fh[8];
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = __ldg(mean + c0 + j);
      is[j] = __ldg(invstd + c0 + j);
      fs[j] = mask_from_raw ? __ldg(fscale + c0 + j) : 0.f;
      fh[j] = mask_from_raw ? __ldg(fshift + c0 + j) : 0.f;
    }
  }
  // ---- phase 1: s1[c] = sum dz, s2[c] = sum dz * xhat
  if (active) {
    float l1[8], l2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) l1[j] = l2[j] = 0.f;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int b = item / chunks, chunk = item - b * chunks;
      const int p0 = chunk * pix_per_block, p1 = min(HW, p0 + pix_per_block);
      const long long base = static_cast<long long>(b) * HW * C;
      float gt[8], pg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gt[j] = gate ? __ldg(gate + static_cast<long long>(b) * C + c0 + j) : 1.f;
        pg[j] = pool_grad ? __ldg(pool_grad + static_cast<long long>(b) * C + c0 + j) : 0.f;
      }
      for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
        uint4 ud[U], ur[U], uy[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int px = pix + k * rows_pp;
          if (px < p1) {
            const long long off = base + static_cast<long long>(px) * C + c0;
            ud[k] = ld_stream16(dy + off);
            ur[k] = ld_stream16(raw + off);
            if (act == ACT_RELU && !mask_from_raw) uy[k] = ld_stream16(y + off);
          }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          if (pix + k * rows_pp >= p1) break;
          float d[8], yy[8], r[8];
          unpack8(ud[k], d);
          unpack8(ur[k], r);
          if (act == ACT_RELU && !mask_from_raw) unpack8(uy[k], yy);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float dz = d[j] * gt[j] + pg[j];
            if (mask_from_raw) yy[j] = fmaf(r[j], fs[j], fh[j]);
            if (act == ACT_RELU && !(yy[j] > 0.f)) dz = 0.f;
            l1[j] += dz;
            l2[j] = fmaf(dz, (r[j] - mu[j]) * is[j], l2[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&a1[c0 + j], l1[j]);
      atomicAdd(&a2[c0 + j], l2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(s1 + i, a1[i]);
    atomicAdd(s2 + i, a2[i]);
  }
  __threadfence();
  cooperative_groups::this_grid().sync();
  // ---- phase 2: draw = k0 * dz + k1 * raw + k2
  if (!active) return;
  float k0[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    k0[j] = __ldg(gamma + c) * is[j];
    k1[j] = -k0[j] * is[j] * __ldcg(s2 + c) * inv_n;
    k2[j] = -k0[j] * __ldcg(s1 + c) * inv_n - k1[j] * mu[j];
  }
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / chunks, chunk = item - b * chunks;
    const int p0 = chunk * pix_per_block, p1 = min(HW, p0 + pix_per_block);
    const long long base = static_cast<long long>(b) * HW * C;
    float gt[8], pg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gt[j] = gate ? __ldg(gate + static_cast<long long>(b) * C + c0 + j) : 1.f;
      pg[j] = pool_grad ? __ldg(pool_grad + static_cast<long long>(b) * C + c0 + j) : 0.f;
    }
    for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
      uint4 ud[U], ur[U], uy[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px < p1) {
          const long long off = base + static_cast<long long>(px) * C + c0;
          ud[k] = ld_stream16(dy + off);
          ur[k] = ld_stream16(raw + off);
          if (act == ACT_RELU && !mask_from_raw) uy[k] = ld_stream16(y + off);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px >= p1) break;
        const long long off = base + static_cast<long long>(px) * C + c0;
        float d[8], yy[8], r[8], o[8], z[8];
        unpack8(ud[k], d);
        unpack8(ur[k], r);
        if (act == ACT_RELU && !mask_from_raw) unpack8(uy[k], yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dz = d[j] * gt[j] + pg[j];
          if (mask_from_raw) yy[j] = fmaf(r[j], fs[j], fh[j]);
          if (act == ACT_RELU && !(yy[j] > 0.f)) dz = 0.f;
          z[j] = dz;
          o[j] = fmaf(k0[j], dz, fmaf(k1[j], r[j], k2[j]));
        }
        store8(draw + off, o);
        if (dz_out) store8(dz_out + off, z);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- SE backward
// dgate_sum[b,c] = sum_p da2s[b,p,c] * a2[b,p,c]
__global__ void __launch_bounds__(256) se_bwd_reduce_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ a2,
                                                            float* __restrict__ dgate_sum, int HW, int C,
                                                            int pix_per_block) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, c8n = C / 8;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const long long base = static_cast<long long>(b) * HW * C;
  const int rows_pp = blockDim.x / c8n;
  const int cg = threadIdx.x % c8n, prow = threadIdx.x / c8n;
  if (prow < rows_pp) {
    const int c0 = cg * 8;
    float l[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    constexpr int U = 4;
    for (int pix = p0 + prow; pix < p1; pix += rows_pp * U) {
      uint4 ud[U], ua[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int px = pix + k * rows_pp;
        if (px < p1) {
          const long long off = base + static_cast<long long>(px) * C + c0;
          ud[k] = ld_stream16(dout + off);
          ua[k] = ld_stream16(a2 + off);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (pix + k * rows_pp >= p1) break;
        float d[8], a[8];
        unpack8(ud[k], d);
        unpack8(ua[k], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) l[j] = fmaf(d[j], a[j], l[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sm[c0 + j], l[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dgate_sum + static_cast<long long>(b) * C + i, sm[i]);
}

// gate = sigmoid(W2 relu(W1 mean + b1) + b2), mean = pool_sum / hw.
// Passes 1-2 (se_kernels.cuh): ds = dgate * g(1-g), dpre = relu'(hid) * (W2^T ds), pool_grad = W1^T dpre / hw.
// Pass 3 (grid over weight elements): dW2 = ds^T hid, dW1 = dpre^T mean, db = column sums — a batch-contraction
// without atomics.
__global__ void __launch_bounds__(256) se_param_grad_kernel(const float* __restrict__ ds, const float* __restrict__ dpre,
                                                            const float* __restrict__ hidden,
                                                            const float* __restrict__ pool_sum, float inv_hw,
                                                            float* __restrict__ dw1, float* __restrict__ db1,
                                                            float* __restrict__ dw2, float* __restrict__ db2, int B,
                                                            int C, int R) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * R) {
    {  // dw2 (C, R)
      const int c = i / R, r = i % R;
      float a = 0.f;
      for (int b = 0; b < B; ++b) a = fmaf(ds[b * C + c], hidden[b * R + r], a);
      dw2[i] += a;
    }
    {  // dw1 (R, C)
      const int r = i / C, c = i % C;
      float a = 0.f;
      for (int b = 0; b < B; ++b) a = fmaf(dpre[b * R + r], pool_sum[b * C + c] * inv_hw, a);
      dw1[i] += a;
    }
  }
  if (i < C) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += ds[b * C + i];
    db2[i] += a;
  }
  if (i < R) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += dpre[b * R + i];
    db1[i] += a;
  }
}

// ---------------------------------------------------------------------------------------------- activation / bias
// dz = dy * act'(y), dbias[c] += sum dz.  dy, y: NHWC bf16 (C channels) or NCHW f32; dz: NHWC bf16 with Cp >= C
// channels (zero padded) so it can feed the TMA-based GEMMs (Cp % 8 == 0).
__global__ void __launch_bounds__(256) act_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ y,
                                                      int nchw_f32, int act, int act_n_limit, float dy_scale,
                                                      bf16* __restrict__ dz, float* __restrict__ dbias, long long npix,
                                                      int HW, int C, int Cp, const unsigned long long* drop_rng,
                                                      float drop_p, unsigned drop_site) {
  extern __shared__ float sm[];  // Cp
  const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);
  for (int i = threadIdx.x; i < Cp; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const long long total = npix * Cp;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % Cp);
    const long long pix = i / Cp;
    float v = 0.f;
    if (c < C) {
      float d, yy;
      if (nchw_f32 == 1) {
        const long long b = pix / HW, hw = pix % HW;
        const long long off = (b * C + c) * HW + hw;
        d = static_cast<const float*>(dy)[off];
        yy = y ? static_cast<const float*>(y)[off] : 0.f;
      } else if (nchw_f32 == 2) {
        d = static_cast<const float*>(dy)[pix * C + c];
        yy = y ? bf2f(static_cast<const bf16*>(y)[pix * C + c]) : 0.f;
      } else {
        d = bf2f(static_cast<const bf16*>(dy)[pix * C + c]);
        yy = y ? bf2f(static_cast<const bf16*>(y)[pix * C + c]) : 0.f;
      }
      d *= dy_scale;
      if (drop.on) d *= drop_mult(drop, static_cast<unsigned long long>(pix) * C + c);
      const int a = (act_n_limit == 0 || c < act_n_limit) ? act : ACT_NONE;
      if (a == ACT_RELU) d = yy > 0.f ? d : 0.f;
      else if (a == ACT_SIGMOID) d = d * yy * (1.f - yy);
      v = d;
      if (dbias) atomicAdd(&sm[c], v);
    }
    if (dz) dz[i] = f2bf(v);
  }
  if (dbias) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dbias + i, sm[i]);
  }
}

// 8-channel-vector variant for the common case (NHWC, C % 8 == 0, no channel padding): a thread owns one 8-channel
// group of a slab of <= 256 groups and walks rows, bias sums stay in registers until the end.
// DY_F32: dy is fp32 (layout 2) instead of bf16.
template <bool DY_F32>
__global__ void __launch_bounds__(256) act_bwd_vec_kernel(const void* __restrict__ dy, const bf16* __restrict__ y, int act,
                                                          int act_n_limit, float dy_scale, bf16* __restrict__ dz,
                                                          float* __restrict__ dbias, long long npix, int C,
                                                          int slab_groups, long long pix_per_block,
                                                          const unsigned long long* drop_rng, float drop_p,
                                                          unsigned drop_site) {
  __shared__ float sm[2048];
  const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);
  const int c8n = C / 8;
  const int g0 = blockIdx.y * slab_groups;
  const int ng = min(slab_groups, c8n - g0);
  if (dbias) {
    for (int i = threadIdx.x; i < ng * 8; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  const long long p0 = blockIdx.x * pix_per_block;
  const long long p1 = min(npix, p0 + pix_per_block);
  const int rows_pp = blockDim.x / ng;
  const int cg = threadIdx.x % ng, prow = threadIdx.x / ng;
  if (prow < rows_pp) {
    const int c0 = (g0 + cg) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    constexpr int U = 2;
    for (long long pix = p0 + prow; pix < p1; pix += rows_pp * U) {
      uint4 ud[U][DY_F32 ? 2 : 1], uy[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long px = pix + k * rows_pp;
        if (px < p1) {
          const long long off = px * C + c0;
          if (DY_F32) {
            ud[k][0] = ld_stream16(static_cast<const float*>(dy) + off);
            ud[k][DY_F32 ? 1 : 0] = ld_stream16(static_cast<const float*>(dy) + off + 4);
          } else {
            ud[k][0] = ld_stream16(static_cast<const bf16*>(dy) + off);
          }
          if (y) uy[k] = ld_stream16(y + off);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long px = pix + k * rows_pp;
        if (px >= p1) break;
        float d[8], yy[8];
        if (DY_F32) {
          const uint4 a = ud[k][0], b = ud[k][DY_F32 ? 1 : 0];
          d[0] = __uint_as_float(a.x); d[1] = __uint_as_float(a.y); d[2] = __uint_as_float(a.z); d[3] = __uint_as_float(a.w);
          d[4] = __uint_as_float(b.x); d[5] = __uint_as_float(b.y); d[6] = __uint_as_float(b.z); d[7] = __uint_as_float(b.w);
        } else {
          unpack8(ud[k][0], d);
        }
        if (y) unpack8(uy[k], yy);
        float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (drop.on) drop_mult8(drop, static_cast<unsigned long long>(px) * C + c0, dm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = d[j] * dy_scale * dm[j];
          const int a = (act_n_limit == 0 || c0 + j < act_n_limit) ? act : ACT_NONE;
          if (a == ACT_RELU) v = yy[j] > 0.f ? v : 0.f;
          else if (a == ACT_SIGMOID) v = v * yy[j] * (1.f - yy[j]);
          d[j] = v;
          acc[j] += v;
        }
        if (dz) store8(dz + px * C + c0, d);
      }
    }
    if (dbias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sm[cg * 8 + j], acc[j]);
    }
  }
  if (dbias) {
    __syncthreads();
    for (int i = threadIdx.x; i < ng * 8; i += blockDim.x) atomicAdd(dbias + g0 * 8 + i, sm[i]);
  }
}

// ---------------------------------------------------------------------------------------------- bilinear adjoint
// dsrc[b,sy,sx,c] (+)= sum over destination pixels whose interpolation stencil touches (sy,sx).
__device__ __forceinline__ void bl_coef(int d, int dn, int sn, int& i0, int& i1, float& l) {
  const float f = fmaxf((d + 0.5f) * (static_cast<float>(sn) / dn) - 0.5f, 0.f);
  i0 = min(static_cast<int>(f), sn - 1);
  i1 = min(i0 + 1, sn - 1);
  l = f - i0;
}
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const bf16* __restrict__ dout, void* __restrict__ dsrc,
                                                           int dsrc_f32, long long s_sb, long long s_srow,
                                                           int accumulate, int B, int sh, int sw, int dh, int dw, int C) {
  const unsigned c8n = C / 8;
  const unsigned total = static_cast<unsigned>(B) * sh * sw * c8n;   // < 2^31 (checked on the host): 32-bit index math
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % c8n) * 8;
  unsigned t = i / c8n;
  const int sx = static_cast<int>(t % static_cast<unsigned>(sw));
  t /= static_cast<unsigned>(sw);
  const int sy = static_cast<int>(t % static_cast<unsigned>(sh));
  const int b = static_cast<int>(t / static_cast<unsigned>(sh));
  // destination rows / columns whose 2-tap stencil can touch source row sy: |fy - sy| < 1 with fy = (y + .5) * sh/dh - .5
  const float ryf = static_cast<float>(dh) / sh, rxf = static_cast<float>(dw) / sw;
  const int ylo = max(0, static_cast<int>(floorf(ryf * (sy - 0.5f) - 0.5f))),
            yhi = min(dh - 1, static_cast<int>(ceilf(ryf * (sy + 1.5f) - 0.5f)));
  const int xlo = max(0, static_cast<int>(floorf(rxf * (sx - 0.5f) - 0.5f))),
            xhi = min(dw - 1, static_cast<int>(ceilf(rxf * (sx + 1.5f) - 0.5f)));
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int y = ylo; y <= yhi; ++y) {
    int y0, y1;
    float ly;
    bl_coef(y, dh, sh, y0, y1, ly);
    const float wy = (y0 == sy ? 1.f - ly : 0.f) + (y1 == sy ? ly : 0.f);
    if (wy == 0.f) continue;
    for (int x = xlo; x <= xhi; ++x) {
      int x0, x1;
      float lx;
      bl_coef(x, dw, sw, x0, x1, lx);
      const float wx = (x0 == sx ? 1.f - lx : 0.f) + (x1 == sx ? lx : 0.f);
      if (wx == 0.f) continue;
      float d[8];
      load8(dout + ((static_cast<long long>(b) * dh + y) * dw + x) * C + c0, d);
      const float wgt = wy * wx;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, d[j], acc[j]);
    }
  }
  const long long off = b * s_sb + (static_cast<long long>(sy) * sw + sx) * s_srow + c0;
  if (dsrc_f32) {
    float* p = static_cast<float*>(dsrc) + off;
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = accumulate ? p[j] + acc[j] : acc[j];
  } else {
    bf16* p = static_cast<bf16*>(dsrc) + off;
    if (accumulate) {
      float o[8];
      load8(p, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += o[j];
    }
    store8(p, acc);
  }
}

// adjoint of bilinear_nchw_mask: dsrc[b,sy,sx,c] = sum_dest w * mask * dout_nchw_f32[b,c,y,x]; dsrc NHWC bf16 (Cs pad)
__global__ void __launch_bounds__(256) bilinear_nchw_mask_bwd_kernel(const float* __restrict__ dout,
                                                                     const float* __restrict__ mask,
                                                                     bf16* __restrict__ dsrc, int B, int sh, int sw,
                                                                     int Cs, int C, int dh, int dw) {
  const long long total = static_cast<long long>(B) * sh * sw * Cs;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % Cs);
  long long t = i / Cs;
  const int sx = static_cast<int>(t % sw);
  t /= sw;
  const int sy = static_cast<int>(t % sh);
  const int b = static_cast<int>(t / sh);
  float acc = 0.f;
  if (c < C) {
    const int ry = (dh + sh - 1) / sh, rx = (dw + sw - 1) / sw;
    const int ylo = max(0, (sy - 1) * ry - 1), yhi = min(dh - 1, (sy + 2) * ry);
    const int xlo = max(0, (sx - 1) * rx - 1), xhi = min(dw - 1, (sx + 2) * rx);
    for (int y = ylo; y <= yhi; ++y) {
      int y0, y1;
      float ly;
      bl_coef(y, dh, sh, y0, y1, ly);
      const float wy = (y0 == sy ? 1.f - ly : 0.f) + (y1 == sy ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int x = xlo; x <= xhi; ++x) {
        int x0, x1;
        float lx;
        bl_coef(x, dw, sw, x0, x1, lx);
        const float wx = (x0 == sx ? 1.f - lx : 0.f) + (x1 == sx ? lx : 0.f);
        if (wx == 0.f) continue;
        const float m = mask ? __ldg(mask + static_cast<long long>(y) * dw + x) : 1.f;
        acc = fmaf(wy * wx * m, dout[((static_cast<long long>(b) * C + c) * dh + y) * dw + x], acc);
      }
    }
  }
  dsrc[i] = f2bf(acc);
}

// adjoint of avgpool_tokens fused with the residual pass-through: out = dout + dtok[b, token(y,x), :] / window
__global__ void __launch_bounds__(256) pool_bwd_add_kernel(const bf16* __restrict__ dout, const void* __restrict__ dtok,
                                                           int dtok_f32, bf16* __restrict__ out, int B, int H, int W,
                                                           int C, int ph, int pw, int rows_per_batch, int row0) {
  const int c8n = C / 8;
  const long long total = static_cast<long long>(B) * H * W * c8n;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c0 = static_cast<int>(i % c8n) * 8;
  long long t = i / c8n;
  const int x = static_cast<int>(t % W);
  t /= W;
  const int y = static_cast<int>(t % H);
  const int b = static_cast<int>(t / H);
  const int wh = H / ph, ww = W / pw;
  const int row = row0 + (y / wh) * pw + (x / ww);
  const float inv = 1.f / static_cast<float>(wh * ww);
  const long long toff = (static_cast<long long>(b) * rows_per_batch + row) * C + c0;
  float v[8], d[8];
  if (dtok_f32) {
    const float* p = static_cast<const float*>(dtok) + toff;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[j] * inv;
  } else {
    load8(static_cast<const bf16*>(dtok) + toff, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= inv;
  }
  if (dout) {
    load8(dout + i * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += d[j];
  }
  store8(out + i * 8, v);
}

// y = a + b (bf16 NHWC gradients meeting at a fork)
__global__ void __launch_bounds__(256) add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                       bf16* __restrict__ y, long long total8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  float u[8], v[8];
  load8(a + i * 8, u);
  load8(b + i * 8, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) u[j] += v[j];
  store8(y + i * 8, u);
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y,
                                                            long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = f2bf(x[i]);
}

// strided row gather + cast: out[g*rows + r, :] = bf16(x[g, row0 + r, :]); dbias[c] += column sums
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ x, bf16* __restrict__ out,
                                                        float* __restrict__ dbias, int groups, int group_rows, int row0,
                                                        int rows, int C) {
  extern __shared__ float sm[];
  if (dbias) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  const long long total = static_cast<long long>(groups) * rows * C;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const long long g = r / rows, rr = r % rows;
    const float v = x[(g * group_rows + row0 + rr) * C + c];
    out[i] = f2bf(v);
    if (dbias) atomicAdd(&sm[c], v);
  }
  if (dbias) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dbias + i, sm[i]);
  }
}

__global__ void __launch_bounds__(256) batch_reduce_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                           int batch, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int b = 0; b < batch; ++b) a += x[b * n + i];
  out[i] += a;
}

// adjoint of parity_split: (4B,H/2,W/2,C) planes -> (B,H,W,C)
__global__ void __launch_bounds__(256) parity_merge_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                           long long total8, int B, int H, int W, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8n = C / 8;
  const int c8 = static_cast<int>(i % c8n);
  long long pix = i / c8n;
  const int xx = static_cast<int>(pix % W);
  pix /= W;
  const int yy = static_cast<int>(pix % H);
  const int b = static_cast<int>(pix / H);
  const int q = (yy & 1) * 2 + (xx & 1);
  const long long src = (((static_cast<long long>(q) * B + b) * (H / 2) + (yy >> 1)) * (W / 2) + (xx >> 1)) * C + c8 * 8;
  *reinterpret_cast<uint4*>(y + i * 8) = *reinterpret_cast<const uint4*>(x + src);
}

// ---------------------------------------------------------------------------------------------- stem weight gradient
// dW[o][c][ky][kx] = sum_{b,oy,ox} draw[b,oy,ox,o] * xin[b,c,2oy+ky-1,2ox+kx-1]   (xin = normalised image, zero pad)
// One block per (c, ky, kx) x pixel chunk; 32 output channels per thread row.
template <int CIN>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ draw,
                                                         const float* __restrict__ in_scale,
                                                         const float* __restrict__ in_shift, float* __restrict__ dw,
                                                         int B, int H, int W) {
  __shared__ float acc[32];
  if (threadIdx.x < 32) acc[threadIdx.x] = 0.f;
  __syncthreads();
  const int tap = blockIdx.y;  // c*9 + ky*3 + kx
  const int c = tap / 9, ky = (tap % 9) / 3, kx = tap % 3;
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  const float a = in_scale ? in_scale[c] : 1.f, sft = in_shift ? in_shift[c] : 0.f;
  float l[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) l[o] = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int ox = static_cast<int>(idx % Wo), oy = static_cast<int>((idx / Wo) % Ho);
    const int b = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
    const int iy = oy * 2 + ky - 1, ix = ox * 2 + kx - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const float v = x[((static_cast<long long>(b) * CIN + c) * H + iy) * W + ix] * a + sft;
    const uint4* dp = reinterpret_cast<const uint4*>(draw + idx * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = __ldg(dp + q);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        l[q * 8 + 2 * j] = fmaf(v, f.x, l[q * 8 + 2 * j]);
        l[q * 8 + 2 * j + 1] = fmaf(v, f.y, l[q * 8 + 2 * j + 1]);
      }
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    const float s = warp_sum(l[o]);
    if (lane == 0) atomicAdd(&acc[o], s);
  }
  __syncthreads();
  if (threadIdx.x < 32) atomicAdd(dw + (static_cast<long long>(threadIdx.x) * CIN + c) * 9 + ky * 3 + kx, acc[threadIdx.x]);
}

}  // namespace

#define STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

static void chunking(int batch, int hw, int* chunks, int* pix_per_block, int ctas_per_sm = 6) {
  int ch = TFPP_NUM_SMS * ctas_per_sm / batch;  // default: two full waves of three resident CTAs per SM
  if (ch < 1) ch = 1;
  int ppb = ceil_div(hw, ch);
  if (ppb < 8) ppb = 8;
  *chunks = ceil_div(hw, ppb);
  *pix_per_block = ppb;
}

template <int MASK, bool GATE>
static int launch_bn_bwd_s2(dim3 grid, size_t smem, cudaStream_t stream, const bf16* dy, const bf16* y, const bf16* raw,
                            const float* mean, const float* invstd, const float* gamma, const float* gate,
                            const float* pool_grad, const float* fs, const float* fh, int act, float* s1, float* s2,
                            float inv_n, bf16* draw, bf16* dz_out, int hw, int channels, int ppb) {
  constexpr int U = 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(bn_bwd_reduce_s_kernel<U, MASK, GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4);
    cudaFuncSetAttribute(bn_bwd_apply_s_kernel<U, MASK, GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 2048 * 4);
    attr_set = true;
  }
  bn_bwd_reduce_s_kernel<U, MASK, GATE><<<grid, 256, smem, stream>>>(dy, y, raw, mean, invstd, gate, pool_grad, fs, fh, act,
                                                                     s1, s2, hw, channels, ppb);
  TFPP_CHECK_LAUNCH();
  bn_bwd_apply_s_kernel<U, MASK, GATE><<<grid, 256, smem, stream>>>(dy, y, raw, mean, invstd, gamma, s1, s2, gate, pool_grad,
                                                                    fs, fh, act, inv_n, draw, dz_out, hw, channels, ppb);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

template <typename... A>
static int launch_bn_bwd_s(int mask, bool gated, A... a) {
  switch (mask * 2 + (gated ? 1 : 0)) {
    case 0: return launch_bn_bwd_s2<0, false>(a...);
    case 1: return launch_bn_bwd_s2<0, true>(a...);
    case 2: return launch_bn_bwd_s2<1, false>(a...);
    case 3: return launch_bn_bwd_s2<1, true>(a...);
    case 4: return launch_bn_bwd_s2<2, false>(a...);
    default: return launch_bn_bwd_s2<2, true>(a...);
  }
}

extern "C" int tfpp_bn_bwd(const void* dy, const void* y, const void* raw, const float* mean, const float* invstd,
                           const float* gamma, const float* gate, const float* pool_grad, const float* fwd_scale,
                           const float* fwd_shift, int act, float* s1, float* s2, void* draw, void* dz_out, int batch,
                           int hw, int channels, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0 && channels <= 2048, "channels must be a multiple of 8, <= 2048");
  TFPP_CHECK_ARG(act != ACT_RELU || y != nullptr || (fwd_scale != nullptr && fwd_shift != nullptr),
                 "ReLU mask needs y, or the forward scale/shift to recompute it from raw");
  int chunks, ppb;
  chunking(batch, hw, &chunks, &ppb);
  dim3 grid(chunks, batch);
  {
    // EXPERIMENTAL single cooperative launch (reduce -> grid barrier -> apply) for tensors that stay in L2
    static const bool fused_on = [] { const char* e = getenv("TFPP_BN_BWD_FUSED"); return e != nullptr && e[0] == '1'; }();
    const size_t tensor_bytes = static_cast<size_t>(batch) * hw * channels * 2;
    if (fused_on && tensor_bytes * (y != nullptr ? 3 : 2) <= (80u << 20)) {
      static int max_ctas = 0;
      if (max_ctas == 0) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel, 256, sizeof(float) * 2 * 2048);
        max_ctas = (per_sm > 0 ? per_sm : 1) * TFPP_NUM_SMS;
      }
      int items = chunks * batch;
      int nctas = items < max_ctas ? items : max_ctas;
      const bf16* a_dy = static_cast<const bf16*>(dy);
      const bf16* a_y = static_cast<const bf16*>(y);
      const bf16* a_raw = static_cast<const bf16*>(raw);
      bf16* a_draw = static_cast<bf16*>(draw);
      bf16* a_dz = static_cast<bf16*>(dz_out);
      float inv_n = 1.f / (static_cast<float>(batch) * hw);
      void* args[] = {&a_dy, &a_y, &a_raw, &mean, &invstd, &gamma, &gate, &pool_grad, &fwd_scale, &fwd_shift, &act,
                      &inv_n, &s1, &s2, &a_draw, &a_dz, &hw, &channels, &ppb, &chunks, &batch};
      cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(bn_bwd_fused_kernel), dim3(nctas), dim3(256),
                                                  args, sizeof(float) * 2 * channels, stream);
      if (e != cudaSuccess) {
        tfpp_set_error("%s:%d: cooperative launch: %s", __FILE__, __LINE__, cudaGetErrorString(e));
        return TFPP_ERR_CUDA;
      }
      return TFPP_OK;
    }
  }
  static const bool stream_on = [] { const char* e = getenv("TFPP_BN_STREAM"); return e == nullptr || e[0] != '0'; }();
  if (stream_on) {
    chunking(batch, hw, &chunks, &ppb, 2);  // one wave, two resident CTAs per SM
    const int mask = act != ACT_RELU ? 0 : (y == nullptr ? 1 : 2);
    const bool gated = gate != nullptr || pool_grad != nullptr;
    return launch_bn_bwd_s(mask, gated, dim3(chunks, batch), sizeof(float) * 7 * channels, stream,
                           static_cast<const bf16*>(dy), static_cast<const bf16*>(y), static_cast<const bf16*>(raw), mean,
                           invstd, gamma, gate, pool_grad, fwd_scale, fwd_shift, act, s1, s2,
                           1.f / (static_cast<float>(batch) * hw), static_cast<bf16*>(draw), static_cast<bf16*>(dz_out), hw,
                           channels, ppb);
  }
  bn_bwd_reduce_kernel<<<grid, 256, sizeof(float) * 2 * channels, stream>>>(
      static_cast<const bf16*>(dy), static_cast<const bf16*>(y), static_cast<const bf16*>(raw), mean, invstd, gate,
      pool_grad, fwd_scale, fwd_shift, act, s1, s2, hw, channels, ppb);
  TFPP_CHECK_LAUNCH();
  bn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(
      static_cast<const bf16*>(dy), static_cast<const bf16*>(y), static_cast<const bf16*>(raw), mean, invstd, gamma, s1,
      s2, gate, pool_grad, fwd_scale, fwd_shift, act, 1.f / (static_cast<float>(batch) * hw), static_cast<bf16*>(draw),
      static_cast<bf16*>(dz_out), hw, channels, ppb);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_se_bwd(const void* dout, const void* a2, const float* gate, const float* hidden,
                           const float* pool_sum, int hw, const float* w1, const float* w2, float* dgate_sum, float* ws,
                           float* dw1, float* db1, float* dw2, float* db2, float* pool_grad, int batch, int channels,
                           int rd, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0, "channels must be a multiple of 8");
  int chunks, ppb;
  chunking(batch, hw, &chunks, &ppb);
  dim3 grid(chunks, batch);
  if (dout != nullptr) {  // dout == NULL: dgate_sum was reduced by the caller (fp32 parity mode, tfpp_se_bwd_reduce_f32)
    se_bwd_reduce_kernel<<<grid, 256, sizeof(float) * channels, stream>>>(static_cast<const bf16*>(dout),
                                                                          static_cast<const bf16*>(a2), dgate_sum, hw,
                                                                          channels, ppb);
    TFPP_CHECK_LAUNCH();
  }
  // workspace: ds (B,C) then dpre (B,rd)
  float* ds_ws = ws;
  float* dpre_ws = ws + static_cast<long long>(batch) * channels;
  TFPP_CHECK_ARG(rd <= 2048, "rd <= 2048");
  se_contract_c_kernel<SE_BWD><<<dim3(ceil_div(rd, 8), ceil_div(batch, 8)), 256, 0, stream>>>(
      dgate_sum, gate, 1.f, w2, 1, rd, nullptr, hidden, dpre_ws, ds_ws, batch, channels, rd);
  TFPP_CHECK_LAUNCH();
  se_contract_r_kernel<SE_BWD><<<dim3(ceil_div(channels, 128), ceil_div(batch, 8)), 128, sizeof(float) * rd * 8, stream>>>(
      dpre_ws, w1, 1, channels, nullptr, 1.f / hw, pool_grad, batch, channels, rd);
  TFPP_CHECK_LAUNCH();
  se_param_grad_kernel<<<ceil_div(channels * rd, 256), 256, 0, stream>>>(ds_ws, dpre_ws, hidden, pool_sum, 1.f / hw, dw1,
                                                                         db1, dw2, db2, batch, channels, rd);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

// In-place dropout over a flat tensor (GPT.drop on the token matrix, transfuser.py:325, and its adjoint).
namespace {
__global__ void __launch_bounds__(256) dropout_inplace_kernel(void* __restrict__ x, int f32, long long n8,
                                                              const unsigned long long* rng, float p, unsigned site) {
  const DropCtx drop = drop_ctx(rng, p, site);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8 || !drop.on) return;
  float dm[8];
  drop_mult8(drop, static_cast<unsigned long long>(i) * 8, dm);
  if (f32) {
    float4* q = reinterpret_cast<float4*>(x) + 2 * i;
    float4 a = q[0], b = q[1];
    a.x *= dm[0]; a.y *= dm[1]; a.z *= dm[2]; a.w *= dm[3];
    b.x *= dm[4]; b.y *= dm[5]; b.z *= dm[6]; b.w *= dm[7];
    q[0] = a;
    q[1] = b;
  } else {
    float v[8];
    load8(static_cast<const bf16*>(x) + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= dm[j];
    store8(static_cast<bf16*>(x) + i * 8, v);
  }
}
}  // namespace

extern "C" int tfpp_dropout(void* x, int x_f32, long long n, const unsigned long long* rng, float p, unsigned site,
                            tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(n % 8 == 0, "element count must be a multiple of 8");
  TFPP_CHECK_ARG(p >= 0.f && p < 1.f, "dropout probability must be in [0, 1)");
  if (rng == nullptr || p == 0.f || n == 0) return TFPP_OK;
  dropout_inplace_kernel<<<static_cast<unsigned>(ceil_div_ll(n / 8, 256)), 256, 0, stream>>>(x, x_f32, n / 8, rng, p, site);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_act_bwd(const void* dy, const void* y, int nchw_f32, int act, int act_n_limit, float dy_scale,
                            void* dz, float* dbias, int batch, int hw, int channels, int channels_padded,
                            tfpp_stream_t stream_) {
  return tfpp_act_bwd_dropout(dy, y, nchw_f32, act, act_n_limit, dy_scale, dz, dbias, batch, hw, channels,
                              channels_padded, nullptr, 0.f, 0u, stream_);
}

extern "C" int tfpp_act_bwd_dropout(const void* dy, const void* y, int nchw_f32, int act, int act_n_limit,
                                    float dy_scale, void* dz, float* dbias, int batch, int hw, int channels,
                                    int channels_padded, const unsigned long long* drop_rng, float drop_p,
                                    unsigned drop_site, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(drop_rng == nullptr || drop_p == 0.f || nchw_f32 != 1, "dropout adjoint: NHWC layouts only");
  if (drop_p <= 0.f) drop_rng = nullptr;
  const long long npix = static_cast<long long>(batch) * hw;
  if (nchw_f32 != 1 && channels % 8 == 0 && channels_padded == channels) {
    const int c8n = channels / 8;
    const int nslabs = ceil_div(c8n, 256);
    const int slab_groups = ceil_div(c8n, nslabs);
    long long chunks = TFPP_NUM_SMS * 6 / nslabs;
    if (chunks < 1) chunks = 1;
    long long ppb = ceil_div_ll(npix, chunks);
    if (ppb < 8) ppb = 8;
    chunks = ceil_div_ll(npix, ppb);
    dim3 grid(static_cast<unsigned>(chunks), nslabs);
    if (nchw_f32 == 2)
      act_bwd_vec_kernel<true><<<grid, 256, 0, stream>>>(dy, static_cast<const bf16*>(y), act, act_n_limit, dy_scale,
                                                         static_cast<bf16*>(dz), dbias, npix, channels, slab_groups, ppb,
                                                         drop_rng, drop_p, drop_site);
    else
      act_bwd_vec_kernel<false><<<grid, 256, 0, stream>>>(dy, static_cast<const bf16*>(y), act, act_n_limit, dy_scale,
                                                          static_cast<bf16*>(dz), dbias, npix, channels, slab_groups,
                                                          ppb, drop_rng, drop_p, drop_site);
    TFPP_CHECK_LAUNCH();
    return TFPP_OK;
  }
  long long blocks = ceil_div_ll(npix * channels_padded, 256 * 4);
  if (blocks > TFPP_NUM_SMS * 8) blocks = TFPP_NUM_SMS * 8;
  if (blocks < 1) blocks = 1;
  act_bwd_kernel<<<static_cast<int>(blocks), 256, sizeof(float) * channels_padded, stream>>>(
      dy, y, nchw_f32, act, act_n_limit, dy_scale, static_cast<bf16*>(dz), dbias, npix, hw, channels, channels_padded,
      drop_rng, drop_p, drop_site);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bilinear_bwd(const void* dout, void* dsrc, int dsrc_f32, long long src_batch_stride,
                                 long long src_row_stride, int accumulate, int batch, int sh, int sw, int dh, int dw,
                                 int channels, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0, "channels must be a multiple of 8");
  const long long total = static_cast<long long>(batch) * sh * sw * (channels / 8);
  TFPP_CHECK_ARG(total < (1ll << 31), "tensor too large for the 32-bit index path");
  bilinear_bwd_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(dout), dsrc, dsrc_f32, src_batch_stride, src_row_stride, accumulate, batch, sh, sw, dh,
      dw, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bilinear_nchw_mask_bwd(const float* dout, const float* mask, void* dsrc, int batch, int sh, int sw,
                                           int src_channels, int channels, int dh, int dw, tfpp_stream_t stream_) {
  STREAM;
  const long long total = static_cast<long long>(batch) * sh * sw * src_channels;
  bilinear_nchw_mask_bwd_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      dout, mask, static_cast<bf16*>(dsrc), batch, sh, sw, src_channels, channels, dh, dw);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_pool_bwd_add(const void* dout, const void* dtok, int dtok_f32, void* out, int batch, int height,
                                 int width, int channels, int ph, int pw, int rows_per_batch, int row0,
                                 tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0, "channels must be a multiple of 8");
  const long long total = static_cast<long long>(batch) * height * width * (channels / 8);
  pool_bwd_add_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(dout), dtok, dtok_f32, static_cast<bf16*>(out), batch, height, width, channels, ph, pw,
      rows_per_batch, row0);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_add_bf16(const void* a, const void* b, void* y, long long n, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(n % 8 == 0, "n must be a multiple of 8");
  add_bf16_kernel<<<static_cast<int>(ceil_div_ll(n / 8, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(a), static_cast<const bf16*>(b), static_cast<bf16*>(y), n / 8);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_cast_f32_bf16(const float* x, void* y, long long n, tfpp_stream_t stream_) {
  STREAM;
  cast_f32_bf16_kernel<<<static_cast<int>(ceil_div_ll(n, 256)), 256, 0, stream>>>(x, static_cast<bf16*>(y), n);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_cast_rows(const float* x, void* out, float* dbias, int groups, int group_rows, int row0, int rows,
                              int channels, tfpp_stream_t stream_) {
  STREAM;
  const long long total = static_cast<long long>(groups) * rows * channels;
  long long blocks = ceil_div_ll(total, 256 * 4);
  if (blocks > TFPP_NUM_SMS * 8) blocks = TFPP_NUM_SMS * 8;
  if (blocks < 1) blocks = 1;
  cast_rows_kernel<<<static_cast<int>(blocks), 256, dbias ? sizeof(float) * channels : 0, stream>>>(
      x, static_cast<bf16*>(out), dbias, groups, group_rows, row0, rows, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_batch_reduce(const float* x, float* out, int batch, long long n, tfpp_stream_t stream_) {
  STREAM;
  batch_reduce_kernel<<<static_cast<int>(ceil_div_ll(n, 256)), 256, 0, stream>>>(x, out, batch, n);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_parity_merge(const void* x, void* y, int batch, int height, int width, int channels,
                                 tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 8 == 0 && height % 2 == 0 && width % 2 == 0, "need C%8==0 and even H, W");
  const long long total8 = static_cast<long long>(batch) * height * width * channels / 8;
  parity_merge_kernel<<<static_cast<int>(ceil_div_ll(total8, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), total8, batch, height, width, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_stem_wgrad(const float* x, const void* draw, const float* in_scale, const float* in_shift,
                               float* dw, int batch, int cin, int height, int width, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(cin >= 1 && cin <= 3, "stem supports 1..3 input channels");
  dim3 grid(TFPP_NUM_SMS, cin * 9);
  const bf16* d = static_cast<const bf16*>(draw);
  if (cin == 1) stem_wgrad_kernel<1><<<grid, 256, 0, stream>>>(x, d, in_scale, in_shift, dw, batch, height, width);
  else if (cin == 2) stem_wgrad_kernel<2><<<grid, 256, 0, stream>>>(x, d, in_scale, in_shift, dw, batch, height, width);
  else stem_wgrad_kernel<3><<<grid, 256, 0, stream>>>(x, d, in_scale, in_shift, dw, batch, height, width);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
