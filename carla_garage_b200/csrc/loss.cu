// Fused loss + loss-gradient kernels of the TransFuser++ training step.
// Reference: LidarCenterNet.compute_loss (team_code/model.py:394-445), LidarCenterNetHead.loss
// (team_code/center_net.py:77-123), gaussian_focal_loss (team_code/transfuser_utils.py:341-364), loss weighting at
// team_code/train.py:452-456,889-896.  Each kernel reads the prediction once, accumulates the scalar loss and writes
// d(weighted total loss)/d(pre-activation) directly in the NHWC bf16 layout the backward GEMMs consume.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

// Cross entropy over a dense map: logits NCHW f32 (B,C,H,W), labels int64 (B,H,W).  valid (H,W) f32 optional:
// pixels with valid == 0 are ignored (model.py:427-430 turns them into ignore_index=-1).  Mean over counted pixels is
// applied by the caller through grad_scale / the loss_sum normalisation (class weights are all 1 on these maps).
__global__ void __launch_bounds__(256) ce_map_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                     const float* __restrict__ valid, float grad_scale,
                                                     const float* __restrict__ w_dev,
                                                     float* __restrict__ loss_sum, bf16* __restrict__ dz_nhwc,
                                                     float* __restrict__ dz_nchw, float* __restrict__ dbias, int C,
                                                     int Cp, int HW, long long npix) {
  __shared__ float sb[32];
  __shared__ float sl;
  if (threadIdx.x < 32) sb[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) sl = 0.f;
  __syncthreads();
  if (w_dev) grad_scale *= *w_dev;  // loss weight living on the device (autograd boundary: d total / d this loss)
  float lsum = 0.f;
  float lb[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) lb[c] = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix; i += stride) {
    const long long b = i / HW, hw = i % HW;
    const float* lp = logits + b * C * HW + hw;
    const bool ok = valid == nullptr || valid[hw] != 0.f;
    float v[16];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      v[c] = c < C ? lp[static_cast<long long>(c) * HW] : -INFINITY;
      m = fmaxf(m, v[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      v[c] = c < C ? __expf(v[c] - m) : 0.f;
      s += v[c];
    }
    const int y = static_cast<int>(labels[i]);
    const float inv = 1.f / s;
    if (ok) {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c == y) lsum += -__logf(fmaxf(v[c] * inv, 1e-38f));
    }
    float g[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      g[c] = (ok && c < C) ? (v[c] * inv - (c == y ? 1.f : 0.f)) * grad_scale : 0.f;
      lb[c] += g[c];
    }
    if (dz_nhwc) {
      bf16* o = dz_nhwc + i * Cp;
#pragma unroll
      for (int c = 0; c < 16; c += 8)
        if (c < Cp)
          *reinterpret_cast<uint4*>(o + c) = make_uint4(pack_bf16x2(g[c], g[c + 1]), pack_bf16x2(g[c + 2], g[c + 3]),
                                                        pack_bf16x2(g[c + 4], g[c + 5]), pack_bf16x2(g[c + 6], g[c + 7]));
    }
    if (dz_nchw) {
      float* o = dz_nchw + b * C * HW + hw;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < C) o[static_cast<long long>(c) * HW] = g[c];
    }
  }
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sl, lsum);
  if (dbias) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float t = warp_sum(lb[c]);
      if ((threadIdx.x & 31) == 0 && c < C) atomicAdd(&sb[c], t);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(loss_sum, sl);
  if (dbias && threadIdx.x < C) atomicAdd(dbias + threadIdx.x, sb[threadIdx.x]);
}

// depth: loss = mean |sigmoid(z) - y| (model.py:379,434); p = sigmoid(z) already computed by the conv epilogue.
__global__ void __launch_bounds__(256) l1_sigmoid_kernel(const float* __restrict__ p, const float* __restrict__ y,
                                                         float grad_scale, const float* __restrict__ w_dev,
                                                         float* __restrict__ loss_sum,
                                                         bf16* __restrict__ dz_nhwc, float* __restrict__ dbias, int Cp,
                                                         long long n) {
  __shared__ float sl, sb;
  if (threadIdx.x == 0) sl = sb = 0.f;
  __syncthreads();
  if (w_dev) grad_scale *= *w_dev;
  float ls = 0.f, lb = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float pv = p[i], d = pv - y[i];
    ls += fabsf(d);
    const float gz = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * pv * (1.f - pv) * grad_scale;
    lb += gz;
    if (dz_nhwc) {
      bf16* o = dz_nhwc + i * Cp;
      *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(gz, 0.f), 0u, 0u, 0u);
      for (int c = 8; c < Cp; c += 8) *reinterpret_cast<uint4*>(o + c) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  ls = warp_sum(ls);
  lb = warp_sum(lb);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sl, ls);
    atomicAdd(&sb, lb);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum, sl);
    if (dbias) atomicAdd(dbias, sb);
  }
}

// CenterNet head losses (center_net.py:77-123) on the fused (B,21,H,W) f32 map [heat 4 | wh 2 | offset 2 | yaw_cls 12
// | yaw_res 1].  losses[0..4] (already divided by avg_factor), dz NHWC bf16 (B,H,W,Cp=24), dbias (21).
__global__ void __launch_bounds__(256) center_loss_kernel(
    const float* __restrict__ maps, const float* __restrict__ t_heat, const float* __restrict__ t_wh,
    const float* __restrict__ t_off, const long long* __restrict__ t_ycls, const float* __restrict__ t_yres,
    const float* __restrict__ pix_w, const float* __restrict__ avg_factor, const float* __restrict__ w5,
    float* __restrict__ losses, bf16* __restrict__ dz, float* __restrict__ dbias, int B, int HW, int n_cls, int n_bins,
    int Cp) {
  __shared__ float sl[5];
  __shared__ float sb[32];
  __shared__ float s_avg;
  if (threadIdx.x < 5) sl[threadIdx.x] = 0.f;
  if (threadIdx.x < 32) sb[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < B; ++i) a += avg_factor[i];
    s_avg = a + 1.1920929e-07f;  // torch.finfo(float32).eps (center_net.py:101)
  }
  __syncthreads();
  const float inv_avg = 1.f / s_avg;
  const int C = n_cls + 4 + n_bins + 1;
  const long long npix = static_cast<long long>(B) * HW;
  float l[5] = {0, 0, 0, 0, 0};
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npix; i += stride) {
    const long long b = i / HW, hw = i % HW;
    const float* mp = maps + b * C * HW + hw;
    float g[24];
#pragma unroll
    for (int c = 0; c < 24; ++c) g[c] = 0.f;
    // heat map: gaussian focal loss, alpha 2, gamma 4 (transfuser_utils.py:341-364)
    for (int c = 0; c < n_cls; ++c) {
      const float p = mp[static_cast<long long>(c) * HW];
      const float t = t_heat[(b * n_cls + c) * HW + hw];
      const float eps = 1e-12f;
      const float pos = (t == 1.f) ? 1.f : 0.f;
      const float omt = 1.f - t;
      const float negw = omt * omt * omt * omt;
      const float lp = __logf(p + eps), ln = __logf(1.f - p + eps);
      l[0] += (-lp * (1.f - p) * (1.f - p) * pos - ln * p * p * negw) * inv_avg;
      const float dldp = pos * (-(1.f - p) * (1.f - p) / (p + eps) + 2.f * (1.f - p) * lp) +
                         negw * (p * p / (1.f - p + eps) - 2.f * p * ln);
      g[c] = dldp * p * (1.f - p) * inv_avg * w5[0];
    }
    const float pw0 = pix_w[(b * 2 + 0) * HW + hw], pw1 = pix_w[(b * 2 + 1) * HW + hw];
    int ch = n_cls;
    for (int c = 0; c < 2; ++c) {  // wh: L1 * pixel_weight / (avg * 2)
      const float d = mp[static_cast<long long>(ch + c) * HW] - t_wh[(b * 2 + c) * HW + hw];
      const float pw = c == 0 ? pw0 : pw1;
      l[1] += fabsf(d) * pw * inv_avg * 0.5f;
      g[ch + c] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * pw * inv_avg * 0.5f * w5[1];
    }
    ch += 2;
    for (int c = 0; c < 2; ++c) {  // offset
      const float d = mp[static_cast<long long>(ch + c) * HW] - t_off[(b * 2 + c) * HW + hw];
      const float pw = c == 0 ? pw0 : pw1;
      l[2] += fabsf(d) * pw * inv_avg * 0.5f;
      g[ch + c] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * pw * inv_avg * 0.5f * w5[2];
    }
    ch += 2;
    {  // yaw class: CE(reduction none) * pixel_weight[:,0] / avg
      float m = -INFINITY;
      for (int c = 0; c < n_bins; ++c) m = fmaxf(m, mp[static_cast<long long>(ch + c) * HW]);
      float s = 0.f;
      for (int c = 0; c < n_bins; ++c) s += __expf(mp[static_cast<long long>(ch + c) * HW] - m);
      const int y = static_cast<int>(t_ycls[b * HW + hw]);
      const float inv = 1.f / s;
      for (int c = 0; c < n_bins; ++c) {
        const float pr = __expf(mp[static_cast<long long>(ch + c) * HW] - m) * inv;
        if (c == y) l[3] += -__logf(fmaxf(pr, 1e-38f)) * pw0 * inv_avg;
        g[ch + c] = (pr - (c == y ? 1.f : 0.f)) * pw0 * inv_avg * w5[3];
      }
    }
    ch += n_bins;
    {  // yaw residual: SmoothL1 (beta 1) * pixel_weight[:,0:1] / avg
      const float d = mp[static_cast<long long>(ch) * HW] - t_yres[b * HW + hw];
      const float ad = fabsf(d);
      l[4] += (ad < 1.f ? 0.5f * d * d : ad - 0.5f) * pw0 * inv_avg;
      g[ch] = (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * pw0 * inv_avg * w5[4];
    }
    if (dz) {
      bf16* o = dz + i * Cp;
#pragma unroll
      for (int c = 0; c < 24; c += 8)
        *reinterpret_cast<uint4*>(o + c) = make_uint4(pack_bf16x2(g[c], g[c + 1]), pack_bf16x2(g[c + 2], g[c + 3]),
                                                      pack_bf16x2(g[c + 4], g[c + 5]), pack_bf16x2(g[c + 6], g[c + 7]));
    }
    if (dbias) {
#pragma unroll
      for (int c = 0; c < 24; ++c)
        if (c < C && g[c] != 0.f) atomicAdd(&sb[c], g[c]);
    }
  }
  for (int k = 0; k < 5; ++k) {
    const float t = warp_sum(l[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sl[k], t);
  }
  __syncthreads();
  if (threadIdx.x < 5) atomicAdd(losses + threadIdx.x, sl[threadIdx.x]);
  if (dbias && threadIdx.x < C) atomicAdd(dbias + threadIdx.x, sb[threadIdx.x]);
}

// target-speed CE with class weights (model.py:416; nn.CrossEntropyLoss(weight): sum(w_y * nll) / sum(w_y)) and
// checkpoint L1 (model.py:419).  Single CTA.
__global__ void __launch_bounds__(256) planner_loss_kernel(const float* __restrict__ logits,
                                                           const long long* __restrict__ labels,
                                                           const float* __restrict__ class_w,
                                                           const float* __restrict__ cp, const float* __restrict__ cp_t,
                                                           float w_ts, float w_cp, const float* __restrict__ w2_dev,
                                                           float* __restrict__ losses,
                                                           float* __restrict__ dlogits, float* __restrict__ dcp, int B,
                                                           int n_cls, int n_cp) {
  __shared__ float s_w, s_l, s_c;
  if (threadIdx.x == 0) s_w = s_l = s_c = 0.f;
  __syncthreads();
  if (w2_dev) {
    w_ts *= w2_dev[0];
    w_cp *= w2_dev[1];
  }
  // logits == nullptr: L1 part only (loss_wp of the use_wp_gru branch, model.py:400-411); cp == nullptr: CE part only
  if (logits == nullptr) B = (cp != nullptr) ? B : 0;
  const int Bce = logits != nullptr ? B : 0;
  float lw = 0.f, ll = 0.f;
  for (int b = threadIdx.x; b < Bce; b += blockDim.x) {
    const int y = static_cast<int>(labels[b]);
    float m = -INFINITY;
    for (int c = 0; c < n_cls; ++c) m = fmaxf(m, logits[b * n_cls + c]);
    float s = 0.f;
    for (int c = 0; c < n_cls; ++c) s += __expf(logits[b * n_cls + c] - m);
    const float w = class_w[y];
    lw += w;
    ll += -w * (logits[b * n_cls + y] - m - __logf(s));
  }
  lw = warp_sum(lw);
  ll = warp_sum(ll);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_w, lw);
    atomicAdd(&s_l, ll);
  }
  float lc = 0.f;
  const int ncp_total = cp != nullptr ? B * n_cp : 0;
  for (int i = threadIdx.x; i < ncp_total; i += blockDim.x) {
    const float d = cp[i] - cp_t[i];
    lc += fabsf(d);
    dcp[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * w_cp / static_cast<float>(B * n_cp);
  }
  lc = warp_sum(lc);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_c, lc);
  __syncthreads();
  for (int b = threadIdx.x; b < Bce; b += blockDim.x) {
    const int y = static_cast<int>(labels[b]);
    float m = -INFINITY;
    for (int c = 0; c < n_cls; ++c) m = fmaxf(m, logits[b * n_cls + c]);
    float s = 0.f;
    for (int c = 0; c < n_cls; ++c) s += __expf(logits[b * n_cls + c] - m);
    const float w = class_w[y] / s_w;
    for (int c = 0; c < n_cls; ++c)
      dlogits[b * n_cls + c] = w * (__expf(logits[b * n_cls + c] - m) / s - (c == y ? 1.f : 0.f)) * w_ts;
  }
  if (threadIdx.x == 0) {
    if (logits != nullptr) losses[0] = s_l / s_w;
    if (cp != nullptr) losses[1] = s_c / static_cast<float>(B * n_cp);
  }
}

}  // namespace

#define STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int tfpp_ce_map_loss(const float* logits, const long long* labels, const float* valid, float grad_scale,
                                const float* w_dev, float* loss_sum, void* dz_nhwc, float* dz_nchw, float* dbias, int batch, int classes,
                                int channels_padded, int hw, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(classes <= 16 && channels_padded <= 16 && channels_padded % 8 == 0, "ce_map: <= 16 classes");
  const long long npix = static_cast<long long>(batch) * hw;
  long long blocks = ceil_div_ll(npix, 256);
  if (blocks > TFPP_NUM_SMS * 8) blocks = TFPP_NUM_SMS * 8;
  ce_map_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(logits, labels, valid, grad_scale, w_dev, loss_sum,
                                                              static_cast<bf16*>(dz_nhwc), dz_nchw, dbias, classes,
                                                              channels_padded, hw, npix);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_l1_sigmoid_loss(const float* p, const float* target, float grad_scale, const float* w_dev,
                                    float* loss_sum,
                                    void* dz_nhwc, float* dbias, int channels_padded, long long n,
                                    tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels_padded % 8 == 0 && channels_padded >= 8, "channels_padded must be a positive multiple of 8");
  long long blocks = ceil_div_ll(n, 256);
  if (blocks > TFPP_NUM_SMS * 8) blocks = TFPP_NUM_SMS * 8;
  l1_sigmoid_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p, target, grad_scale, w_dev, loss_sum,
                                                                  static_cast<bf16*>(dz_nhwc), dbias, channels_padded, n);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_center_head_loss(const float* maps, const float* t_heat, const float* t_wh, const float* t_off,
                                     const long long* t_ycls, const float* t_yres, const float* pix_w,
                                     const float* avg_factor, const float* w5, float* losses, void* dz, float* dbias,
                                     int batch, int hw, int n_cls, int n_bins, int channels_padded,
                                     tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(n_cls + 4 + n_bins + 1 <= 24 && channels_padded == 24, "center head layout: <= 24 channels");
  const long long npix = static_cast<long long>(batch) * hw;
  long long blocks = ceil_div_ll(npix, 256);
  if (blocks > TFPP_NUM_SMS * 4) blocks = TFPP_NUM_SMS * 4;
  center_loss_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(maps, t_heat, t_wh, t_off, t_ycls, t_yres, pix_w,
                                                                   avg_factor, w5, losses, static_cast<bf16*>(dz), dbias,
                                                                   batch, hw, n_cls, n_bins, channels_padded);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_planner_loss(const float* logits, const long long* labels, const float* class_w, const float* cp,
                                 const float* cp_t, float w_ts, float w_cp, const float* w2_dev, float* losses,
                                 float* dlogits, float* dcp, int batch, int n_cls, int n_cp, tfpp_stream_t stream_) {
  STREAM;
  planner_loss_kernel<<<1, 256, 0, stream>>>(logits, labels, class_w, cp, cp_t, w_ts, w_cp, w2_dev, losses, dlogits, dcp, batch,
                                             n_cls, n_cp);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
