// Fused AdamW(amsgrad=True) step.  Reference: optim.AdamW(lr, amsgrad=True) at team_code/train.py:527-531 (torch
// defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0.01).  One pass over (param, grad, exp_avg, exp_avg_sq,
// max_exp_avg_sq): 5 reads + 4 writes of fp32 per element, HBM-bound.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    float* __restrict__ vmax, long long n, float lr, float beta1,
                                                    float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    float grad_scale, const float* __restrict__ dev_state) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dev_state != nullptr) {  // CUDA-graph friendly: step count and learning rate live on the device
    const float t = dev_state[0];
    lr = dev_state[1];
    bc1 = 1.f - powf(beta1, t);
    bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  }
  const float gr = g[i] * grad_scale;
  float pv = p[i];
  pv *= (1.f - lr * wd);                       // decoupled weight decay
  const float mv = beta1 * m[i] + (1.f - beta1) * gr;
  const float vv = beta2 * v[i] + (1.f - beta2) * gr * gr;
  const float vm = fmaxf(vmax[i], vv);
  m[i] = mv;
  v[i] = vv;
  vmax[i] = vm;
  const float denom = sqrtf(vm) / bc2_sqrt + eps;
  p[i] = pv - (lr / bc1) * (mv / denom);
}
__global__ void adamw_tick_kernel(float* dev_state) { dev_state[0] += 1.f; }

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
                                  tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(step >= 1 || dev_state != nullptr, "step counts from 1");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  if (dev_state != nullptr) adamw_tick_kernel<<<1, 1, 0, stream>>>(dev_state);
  adamw_kernel<<<static_cast<int>(ceil_div_ll(n, 256)), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq,
                                                                          max_exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                          weight_decay, bc1, sqrtf(bc2), grad_scale,
                                                                          dev_state);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
