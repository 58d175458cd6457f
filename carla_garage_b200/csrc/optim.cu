// Fused AdamW(amsgrad=True) step.  Reference: optim.AdamW(lr, amsgrad=True) at team_code/train.py:527-531 (torch
// defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0.01).  One pass over (param, grad, exp_avg, exp_avg_sq,
// max_exp_avg_sq): 5 reads + 4 writes of fp32 per element, HBM-bound.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float& vmax, float lr, float beta1,
                                          float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  const float gr = g * grad_scale;
  const float pv = p * (1.f - lr * wd);        // decoupled weight decay
  m = beta1 * m + (1.f - beta1) * gr;
  v = beta2 * v + (1.f - beta2) * gr * gr;
  vmax = fmaxf(vmax, v);
  const float denom = sqrtf(vmax) / bc2_sqrt + eps;
  p = pv - (lr / bc1) * (m / denom);
}

// 4 elements per thread (16-byte accesses; the flat buffers are padded to a multiple of 4), scalar tail otherwise.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    float* __restrict__ vmax, long long n, float lr, float beta1,
                                                    float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    float grad_scale, const float* __restrict__ dev_state,
                                                    const unsigned char* __restrict__ flags) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  // per-element flags (optional): bit 0 = no weight decay (create_optimizer_groups, model.py:556-645), bit 1 = frozen
  // (requires_grad=False, train.py:495-508): the element is left untouched
  unsigned fl = 0u;
  if (flags != nullptr) {
    if (i + 4 <= n) fl = *reinterpret_cast<const unsigned*>(flags + i);
    else for (long long j = i; j < n; ++j) fl |= static_cast<unsigned>(flags[j]) << (8 * (j - i));
    if ((fl & 0x02020202u) == 0x02020202u) return;
  }
  const float wd0 = (fl & 0x00000001u) ? 0.f : wd, wd1 = (fl & 0x00000100u) ? 0.f : wd,
              wd2 = (fl & 0x00010000u) ? 0.f : wd, wd3 = (fl & 0x01000000u) ? 0.f : wd;
  if (dev_state != nullptr) {  // CUDA-graph friendly: step count, learning rate and bias corrections live on the device
    lr = dev_state[1];
    bc1 = dev_state[2];
    bc2_sqrt = dev_state[3];
  }
  if (i + 4 <= n) {
    float4 pv = *reinterpret_cast<float4*>(p + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    float4 xv = *reinterpret_cast<float4*>(vmax + i);
    if (!(fl & 0x00000002u)) adamw_one(pv.x, gv.x, mv.x, vv.x, xv.x, lr, beta1, beta2, eps, wd0, bc1, bc2_sqrt, grad_scale);
    if (!(fl & 0x00000200u)) adamw_one(pv.y, gv.y, mv.y, vv.y, xv.y, lr, beta1, beta2, eps, wd1, bc1, bc2_sqrt, grad_scale);
    if (!(fl & 0x00020000u)) adamw_one(pv.z, gv.z, mv.z, vv.z, xv.z, lr, beta1, beta2, eps, wd2, bc1, bc2_sqrt, grad_scale);
    if (!(fl & 0x02000000u)) adamw_one(pv.w, gv.w, mv.w, vv.w, xv.w, lr, beta1, beta2, eps, wd3, bc1, bc2_sqrt, grad_scale);
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    *reinterpret_cast<float4*>(vmax + i) = xv;
  } else {
    for (long long j = i; j < n; ++j) {
      const unsigned f = (fl >> (8 * (j - i))) & 0xffu;
      if (!(f & 2u)) adamw_one(p[j], g[j], m[j], v[j], vmax[j], lr, beta1, beta2, eps, (f & 1u) ? 0.f : wd, bc1, bc2_sqrt, grad_scale);
    }
  }
}
// dev_state = [step, lr, 1 - beta1^step, sqrt(1 - beta2^step)]
__global__ void adamw_tick_kernel(float* dev_state, float beta1, float beta2) {
  const float t = dev_state[0] + 1.f;
  dev_state[0] = t;
  dev_state[2] = 1.f - powf(beta1, t);
  dev_state[3] = sqrtf(1.f - powf(beta2, t));
}


// Weight-pack refresh (engine.PackPlan): out[i] = idx[i] >= 0 ? flat[idx[i]] : 0, 8 elements per thread.  Most maps are
// long contiguous runs (coalesced); the transposed dgrad packs gather with a stride but hit L2 across neighbouring CTAs.
__global__ void __launch_bounds__(256) gather_pack_kernel(const float* __restrict__ flat, const int* __restrict__ idx,
                                                          void* __restrict__ out, long long n8, int out_f32) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int4 a = __ldg(reinterpret_cast<const int4*>(idx) + 2 * i);
  const int4 b = __ldg(reinterpret_cast<const int4*>(idx) + 2 * i + 1);
  const int ix[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = ix[j] >= 0 ? __ldg(flat + ix[j]) : 0.f;
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(out) + 2 * i;
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                  pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}
}  // namespace

extern "C" int tfpp_gather_pack(const float* flat, const int* idx, void* out, long long n, int out_f32,
                                tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(n % 8 == 0, "n must be a multiple of 8");
  if (n == 0) return TFPP_OK;
  gather_pack_kernel<<<static_cast<unsigned>(ceil_div_ll(n / 8, 256)), 256, 0, stream>>>(flat, idx, out, n / 8, out_f32);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_adamw_amsgrad(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                  float* max_exp_avg_sq, long long n, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, int step, float grad_scale, float* dev_state,
                                  const unsigned char* flags, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(step >= 1 || dev_state != nullptr, "step counts from 1");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  if (dev_state != nullptr) adamw_tick_kernel<<<1, 1, 0, stream>>>(dev_state, beta1, beta2);
  TFPP_CHECK_ARG((reinterpret_cast<uintptr_t>(param) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0,
                 "parameter / gradient buffers must be 16-byte aligned");
  adamw_kernel<<<static_cast<int>(ceil_div_ll(ceil_div_ll(n, 4), 256)), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq,
                                                                          max_exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                          weight_decay, bc1, sqrtf(bc2), grad_scale,
                                                                          dev_state, flags);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
