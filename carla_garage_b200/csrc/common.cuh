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

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

#define TFPP_NUM_SMS 148
