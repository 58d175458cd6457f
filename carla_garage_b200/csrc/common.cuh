// Shared helpers for the TransFuser++ sm_100a kernels (carla_garage_b200/csrc).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TFPP_OK 0
#define TFPP_ERR_ARG 1
#define TFPP_ERR_CUDA 2
#define TFPP_ERR_DRIVER 3

extern "C" void tfpp_set_error(const char* fmt, ...);

#define TFPP_CHECK_ARG(cond, msg)                              \
  do {                                                         \
    if (!(cond)) {                                             \
      tfpp_set_error("%s:%d: %s", __FILE__, __LINE__, msg);    \
      return TFPP_ERR_ARG;                                     \
    }                                                          \
  } while (0)

// Launch errors only (no sync on the hot path, SURVEY.md §8b "Error convention").
#define TFPP_CHECK_LAUNCH()                                                         \
  do {                                                                              \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) {                                                        \
      tfpp_set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return TFPP_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 16-byte streaming load (no L1 allocation: every byte of these feature maps is touched exactly once per kernel)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// activation codes shared by every epilogue
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_GELU = 3 };

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    default: return v;
  }
}

// ---------------------------------------------------------------------------------------------- dropout RNG
// Counter-based Philox4x32-10 (Salmon et al. 2011).  A dropout site (one nn.Dropout / attention-probability dropout of
// one layer: transfuser.py:325,374,379,395; nn.TransformerDecoderLayer's dropouts, model.py:137-140) draws the word for
// element i of its tensor from counter (i / 4, site, step) and key = seed; word i % 4 of the block.  The element is
// dropped when word < p * 2^32.  Nothing is stored: the backward kernels regenerate the same words from (seed, step,
// site, i).  rng = device pointer to {seed, step} (uint64 each) so a CUDA-graph replay sees a fresh step.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
struct DropCtx {
  uint2 key;
  uint32_t site, step, thresh;
  float inv_keep;
  bool on;
};
__device__ __forceinline__ DropCtx drop_ctx(const unsigned long long* rng, float p, unsigned site) {
  DropCtx d;
  d.on = rng != nullptr && p > 0.f;
  d.site = site;
  d.inv_keep = 1.f;
  d.thresh = 0u;
  d.step = 0u;
  d.key = make_uint2(0u, 0u);
  if (d.on) {
    const unsigned long long seed = rng[0], step = rng[1];
    d.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    d.step = static_cast<uint32_t>(step);
    const double t = static_cast<double>(p) * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t);
    d.inv_keep = 1.f / (1.f - p);
  }
  return d;
}
// the four 32-bit words of elements [4 * idx4, 4 * idx4 + 4)
__device__ __forceinline__ uint4 drop_words(const DropCtx& d, unsigned long long idx4) {
  return philox4x32_10(make_uint4(static_cast<uint32_t>(idx4), static_cast<uint32_t>(idx4 >> 32), d.site, d.step), d.key);
}
// multiplier (0 or 1/(1-p)) of element idx
__device__ __forceinline__ float drop_mult(const DropCtx& d, unsigned long long idx) {
  const uint4 w = drop_words(d, idx >> 2);
  const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
  return ws[idx & 3] >= d.thresh ? d.inv_keep : 0.f;
}
// multipliers of the 8 consecutive elements starting at idx (idx % 4 == 0)
__device__ __forceinline__ void drop_mult8(const DropCtx& d, unsigned long long idx, float* m) {
  const uint4 a = drop_words(d, idx >> 2), b = drop_words(d, (idx >> 2) + 1);
  m[0] = a.x >= d.thresh ? d.inv_keep : 0.f; m[1] = a.y >= d.thresh ? d.inv_keep : 0.f;
  m[2] = a.z >= d.thresh ? d.inv_keep : 0.f; m[3] = a.w >= d.thresh ? d.inv_keep : 0.f;
  m[4] = b.x >= d.thresh ? d.inv_keep : 0.f; m[5] = b.y >= d.thresh ? d.inv_keep : 0.f;
  m[6] = b.z >= d.thresh ? d.inv_keep : 0.f; m[7] = b.w >= d.thresh ? d.inv_keep : 0.f;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

#define TFPP_NUM_SMS 148
