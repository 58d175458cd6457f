// fp32 parity mode, backward half (see fp32_path.cu): the adjoints of the fp32 forward ops with fp32 gradient storage, so
// that loss.backward() through the engine's hand-scheduled backward (training.Backward) can be compared with the
// reference's autograd gradients at north_star's 1e-3 instead of at the bf16 noise floor.  Same contracts as the bf16
// twins (featmap_bwd.cu, tc_wgrad.cu, gconv3x3.cu, fusion_attn.cu, transformer_bwd.cu), suffix _f32; simple kernels,
// written for exactness, used by the parity tests only.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ weight gradients
// dw[co, tap_w, ci] += sum_pixels dy[pixel, co] * x[pixel + shift(tap), ci]   (contract of tfpp_conv_wgrad, dense)
struct WgP {
  const float* dy;
  const float* x;
  float* dw;
  int batch, height, width, cout, cout_valid, x_batch, x_channels, cin, ntaps;
  long long x_batch_stride, s_co, s_tap, s_ci;
  int tap_dx[9], tap_dy[9], tap_db[9], tap_w[9];
  long long pix_per_chunk;
};

__global__ void __launch_bounds__(256) conv_wgrad_f32_kernel(const WgP p) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.cout_valid) * p.ntaps * p.cin;
  if (e >= total) return;
  const int ci = static_cast<int>(e % p.cin);
  const int tap = static_cast<int>((e / p.cin) % p.ntaps);
  const int co = static_cast<int>(e / (static_cast<long long>(p.cin) * p.ntaps));
  const long long npix = static_cast<long long>(p.batch) * p.height * p.width;
  const long long p0 = blockIdx.y * p.pix_per_chunk, p1 = min(npix, p0 + p.pix_per_chunk);
  const long long img = p.x_batch_stride > 0 ? p.x_batch_stride : static_cast<long long>(p.height) * p.width * p.x_channels;
  float acc = 0.f;
  for (long long pix = p0; pix < p1; ++pix) {
    const int xx = static_cast<int>(pix % p.width);
    const int yy = static_cast<int>((pix / p.width) % p.height);
    const int bb = static_cast<int>(pix / (static_cast<long long>(p.width) * p.height));
    const int sx = xx + p.tap_dx[tap], sy = yy + p.tap_dy[tap], sb = bb + p.tap_db[tap];
    if (sx < 0 || sx >= p.width || sy < 0 || sy >= p.height || sb < 0 || sb >= p.x_batch) continue;
    acc = fmaf(p.dy[pix * p.cout + co], p.x[sb * img + (static_cast<long long>(sy) * p.width + sx) * p.x_channels + ci], acc);
  }
  atomicAdd(p.dw + co * p.s_co + p.tap_w[tap] * p.s_tap + ci * p.s_ci, acc);
}

// RegNet group conv: dw (C,24,3,3) += ..., x (B,H,W,C), dy (B,H/s,W/s,C)
__global__ void __launch_bounds__(256) gconv_wgrad_f32_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                              float* __restrict__ dw, int B, int H, int W, int C,
                                                              int stride, long long pix_per_chunk) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(C) * 24 * 9;
  if (e >= total) return;
  const int kx = static_cast<int>(e % 3), ky = static_cast<int>((e / 3) % 3);
  const int ci = static_cast<int>((e / 9) % 24);
  const int co = static_cast<int>(e / (9 * 24));
  const int g = co / 24;
  const int Ho = H / stride, Wo = W / stride;
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  const long long p0 = blockIdx.y * pix_per_chunk, p1 = min(npix, p0 + pix_per_chunk);
  float acc = 0.f;
  for (long long pix = p0; pix < p1; ++pix) {
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    const int ix = ox * stride + kx - 1, iy = oy * stride + ky - 1;
    if (ix < 0 || ix >= W || iy < 0 || iy >= H) continue;
    acc = fmaf(dy[pix * C + co], x[((static_cast<long long>(b) * H + iy) * W + ix) * C + g * 24 + ci], acc);
  }
  atomicAdd(dw + e, acc);
}

// input gradient of the stride-2 group conv; w_t (C/24, 9, 24, 24) = [g][flipped tap][ci][co] (pack_gconv_halo transpose)
__global__ void __launch_bounds__(256) gconv_dgrad_s2_f32_kernel(const float* __restrict__ dy, const float* __restrict__ wt,
                                                                 float* __restrict__ dx, int B, int Ho, int Wo, int C) {
  const int H = 2 * Ho, W = 2 * Wo;
  const long long total = static_cast<long long>(B) * H * W * C;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  long long t = i / C;
  const int ix = static_cast<int>(t % W);
  t /= W;
  const int iy = static_cast<int>(t % H);
  const int b = static_cast<int>(t / H);
  const int g = c / 24, ci = c % 24;
  float acc = 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int ny = iy + 1 - ky;
    if (ny < 0 || (ny & 1)) continue;
    const int oy = ny >> 1;
    if (oy >= Ho) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int nx = ix + 1 - kx;
      if (nx < 0 || (nx & 1)) continue;
      const int ox = nx >> 1;
      if (ox >= Wo) continue;
      const float* dp = dy + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C + g * 24;
      const float* wp = wt + ((static_cast<long long>(g) * 9 + (2 - ky) * 3 + (2 - kx)) * 24 + ci) * 24;
#pragma unroll
      for (int co = 0; co < 24; ++co) acc = fmaf(dp[co], wp[co], acc);
    }
  }
  dx[i] = acc;
}

// stem: dW[o][c][ky][kx] += sum draw[b,oy,ox,o] * (x[b,c,2oy+ky-1,2ox+kx-1] * in_scale[c] + in_shift[c])
__global__ void __launch_bounds__(256) stem_wgrad_f32_kernel(const float* __restrict__ x, const float* __restrict__ draw,
                                                             const float* __restrict__ in_scale,
                                                             const float* __restrict__ in_shift, float* __restrict__ dw,
                                                             int B, int CIN, int H, int W, long long pix_per_chunk) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 32 * CIN * 9) return;
  const int kx = e % 3, ky = (e / 3) % 3, c = (e / 9) % CIN, o = e / (9 * CIN);
  const int Ho = H / 2, Wo = W / 2;
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  const long long p0 = blockIdx.y * pix_per_chunk, p1 = min(npix, p0 + pix_per_chunk);
  const float a = in_scale ? in_scale[c] : 1.f, sft = in_shift ? in_shift[c] : 0.f;
  float acc = 0.f;
  for (long long pix = p0; pix < p1; ++pix) {
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    const int ix = ox * 2 + kx - 1, iy = oy * 2 + ky - 1;
    if (ix < 0 || ix >= W || iy < 0 || iy >= H) continue;
    acc = fmaf(draw[pix * 32 + o], x[((static_cast<long long>(b) * CIN + c) * H + iy) * W + ix] * a + sft, acc);
  }
  atomicAdd(dw + e, acc);
}

// ------------------------------------------------------------------------------------------------ BatchNorm backward
// see featmap_bwd.cu: dz = (dy * gate + pool_grad) * act'(y);  s1 = sum dz;  s2 = sum dz * xhat;
// draw = gamma * invstd * (dz - s1/N - xhat * s2/N)
__device__ __forceinline__ float bn_dz(const float* dy, const float* y, const float* raw, long long off, float gt, float pg,
                                       float fs, float fh, int act, bool from_raw) {
  float dz = dy[off] * gt + pg;
  if (act == ACT_RELU) {
    const float yy = from_raw ? fmaf(raw[off], fs, fh) : y[off];
    if (!(yy > 0.f)) dz = 0.f;
  }
  return dz;
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                const float* __restrict__ raw,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd,
                                                                const float* __restrict__ gate,
                                                                const float* __restrict__ pool_grad,
                                                                const float* __restrict__ fscale,
                                                                const float* __restrict__ fshift, int act,
                                                                float* __restrict__ s1, float* __restrict__ s2, int HW,
                                                                int C, int pix_per_block) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const bool from_raw = act == ACT_RELU && y == nullptr;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float gt = gate ? gate[static_cast<long long>(b) * C + c] : 1.f;
    const float pg = pool_grad ? pool_grad[static_cast<long long>(b) * C + c] : 0.f;
    const float fs = from_raw ? fscale[c] : 0.f, fh = from_raw ? fshift[c] : 0.f;
    const float mu = mean[c], is = invstd[c];
    float a1 = 0.f, a2 = 0.f;
    for (int px = p0; px < p1; ++px) {
      const long long off = (static_cast<long long>(b) * HW + px) * C + c;
      const float dz = bn_dz(dy, y, raw, off, gt, pg, fs, fh, act, from_raw);
      a1 += dz;
      a2 = fmaf(dz, (raw[off] - mu) * is, a2);
    }
    atomicAdd(s1 + c, a1);
    atomicAdd(s2 + c, a2);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                               const float* __restrict__ raw,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ invstd,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ s1, const float* __restrict__ s2,
                                                               const float* __restrict__ gate,
                                                               const float* __restrict__ pool_grad,
                                                               const float* __restrict__ fscale,
                                                               const float* __restrict__ fshift, int act, float inv_n,
                                                               float* __restrict__ draw, float* __restrict__ dz_out,
                                                               long long total, int HW, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long b = i / (static_cast<long long>(HW) * C);
  const bool from_raw = act == ACT_RELU && y == nullptr;
  const float gt = gate ? gate[b * C + c] : 1.f, pg = pool_grad ? pool_grad[b * C + c] : 0.f;
  const float dz = bn_dz(dy, y, raw, i, gt, pg, from_raw ? fscale[c] : 0.f, from_raw ? fshift[c] : 0.f, act, from_raw);
  const float xh = (raw[i] - mean[c]) * invstd[c];
  const float g = gamma ? gamma[c] : 1.f;
  draw[i] = g * invstd[c] * (dz - s1[c] * inv_n - xh * s2[c] * inv_n);
  if (dz_out) dz_out[i] = dz;
}

// squeeze-excite: dgate_sum[b,c] = sum_p dout[b,p,c] * a2[b,p,c]
__global__ void __launch_bounds__(256) se_bwd_reduce_f32_kernel(const float* __restrict__ dout, const float* __restrict__ a2,
                                                                float* __restrict__ dgate_sum, int HW, int C,
                                                                int pix_per_block) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int px = p0; px < p1; ++px) {
      const long long off = (static_cast<long long>(b) * HW + px) * C + c;
      a = fmaf(dout[off], a2[off], a);
    }
    atomicAdd(dgate_sum + static_cast<long long>(b) * C + c, a);
  }
}

// ------------------------------------------------------------------------------------------------ element-wise adjoints
// dz = dy * act'(y) (+ dropout mask), dbias += column sums; layout 1: dy / y NCHW f32, else NHWC f32; dz NHWC f32 (Cp)
__global__ void __launch_bounds__(256) act_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                          int nchw, int act, int act_n_limit, float dy_scale,
                                                          float* __restrict__ dz, float* __restrict__ dbias,
                                                          long long npix, int HW, int C, int Cp,
                                                          const unsigned long long* drop_rng, float drop_p,
                                                          unsigned drop_site) {
  const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);
  const long long total = npix * Cp;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % Cp);
    const long long pix = i / Cp;
    float v = 0.f;
    if (c < C) {
      long long off = pix * C + c;
      if (nchw) {
        const long long b = pix / HW, hw = pix % HW;
        off = (b * C + c) * HW + hw;
      }
      float d = dy[off] * dy_scale;
      const float yy = y ? y[off] : 0.f;
      if (drop.on) d *= drop_mult(drop, static_cast<unsigned long long>(pix) * C + c);
      const int a = (act_n_limit == 0 || c < act_n_limit) ? act : ACT_NONE;
      if (a == ACT_RELU) d = yy > 0.f ? d : 0.f;
      else if (a == ACT_SIGMOID) d = d * yy * (1.f - yy);
      v = d;
      if (dbias) atomicAdd(dbias + c, v);
    }
    if (dz) dz[i] = v;
  }
}

__device__ __forceinline__ void bl_coef(int d, int dn, int sn, int& i0, int& i1, float& l) {
  const float f = fmaxf((d + 0.5f) * (static_cast<float>(sn) / dn) - 0.5f, 0.f);
  i0 = min(static_cast<int>(f), sn - 1);
  i1 = min(i0 + 1, sn - 1);
  l = f - i0;
}

// adjoint of the bilinear resize (destination-driven scatter with atomics: exact stencils, any ratio)
__global__ void __launch_bounds__(256) bilinear_bwd_f32_kernel(const float* __restrict__ dout, float* __restrict__ dsrc,
                                                               long long s_sb, long long s_srow, int B, int sh, int sw,
                                                               int dh, int dw, int C) {
  const long long total = static_cast<long long>(B) * dh * dw * C;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  long long t = i / C;
  const int x = static_cast<int>(t % dw);
  t /= dw;
  const int y = static_cast<int>(t % dh);
  const int b = static_cast<int>(t / dh);
  int y0, y1, x0, x1;
  float ly, lx;
  bl_coef(y, dh, sh, y0, y1, ly);
  bl_coef(x, dw, sw, x0, x1, lx);
  const float d = dout[i];
  float* base = dsrc + b * s_sb + c;
  atomicAdd(base + (static_cast<long long>(y0) * sw + x0) * s_srow, (1.f - ly) * (1.f - lx) * d);
  atomicAdd(base + (static_cast<long long>(y0) * sw + x1) * s_srow, (1.f - ly) * lx * d);
  atomicAdd(base + (static_cast<long long>(y1) * sw + x0) * s_srow, ly * (1.f - lx) * d);
  atomicAdd(base + (static_cast<long long>(y1) * sw + x1) * s_srow, ly * lx * d);
}
__global__ void __launch_bounds__(256) zero_slab_f32_kernel(float* __restrict__ dsrc, long long s_sb, long long s_srow,
                                                            int B, int rows, int C) {
  const long long total = static_cast<long long>(B) * rows * C;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long r = (i / C) % rows, b = i / (static_cast<long long>(C) * rows);
  dsrc[b * s_sb + r * s_srow + c] = 0.f;
}

// adjoint of bilinear_nchw_mask: dsrc (B,sh,sw,Cs) f32 (channels >= C zero) from dout NCHW f32 (B,C,dh,dw)
__global__ void __launch_bounds__(256) bilinear_nchw_mask_bwd_f32_kernel(const float* __restrict__ dout,
                                                                         const float* __restrict__ mask,
                                                                         float* __restrict__ dsrc, int B, int sh, int sw,
                                                                         int Cs, int C, int dh, int dw) {
  const long long total = static_cast<long long>(B) * C * dh * dw;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % dw);
  long long t = i / dw;
  const int y = static_cast<int>(t % dh);
  t /= dh;
  const int c = static_cast<int>(t % C);
  const int b = static_cast<int>(t / C);
  int y0, y1, x0, x1;
  float ly, lx;
  bl_coef(y, dh, sh, y0, y1, ly);
  bl_coef(x, dw, sw, x0, x1, lx);
  const float d = dout[i] * (mask ? mask[static_cast<long long>(y) * dw + x] : 1.f);
  float* base = dsrc + static_cast<long long>(b) * sh * sw * Cs + c;
  atomicAdd(base + (static_cast<long long>(y0) * sw + x0) * Cs, (1.f - ly) * (1.f - lx) * d);
  atomicAdd(base + (static_cast<long long>(y0) * sw + x1) * Cs, (1.f - ly) * lx * d);
  atomicAdd(base + (static_cast<long long>(y1) * sw + x0) * Cs, ly * (1.f - lx) * d);
  atomicAdd(base + (static_cast<long long>(y1) * sw + x1) * Cs, ly * lx * d);
}

// out = dout + dtok[b, token(y,x), :] / window
__global__ void __launch_bounds__(256) pool_bwd_add_f32_kernel(const float* __restrict__ dout, const float* __restrict__ dtok,
                                                               float* __restrict__ out, int B, int H, int W, int C, int ph,
                                                               int pw, int rows_per_batch, int row0) {
  const long long total = static_cast<long long>(B) * H * W * C;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  long long t = i / C;
  const int x = static_cast<int>(t % W);
  t /= W;
  const int y = static_cast<int>(t % H);
  const int b = static_cast<int>(t / H);
  const int wh = H / ph, ww = W / pw;
  const int row = row0 + (y / wh) * pw + (x / ww);
  float v = dtok[(static_cast<long long>(b) * rows_per_batch + row) * C + c] * (1.f / static_cast<float>(wh * ww));
  if (dout) v += dout[i];
  out[i] = v;
}

__global__ void __launch_bounds__(256) add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      float* __restrict__ y, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

// out[g*rows + r, :] = x[g, row0 + r, :]; dbias += column sums
__global__ void __launch_bounds__(256) copy_rows_f32_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                            float* __restrict__ dbias, int groups, int group_rows,
                                                            int row0, int rows, int C) {
  const long long total = static_cast<long long>(groups) * rows * C;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const long long g = r / rows, rr = r % rows;
    const float v = x[(g * group_rows + row0 + rr) * C + c];
    out[i] = v;
    if (dbias) atomicAdd(dbias + c, v);
  }
}

// ------------------------------------------------------------------------------------------------ attention backward
// O = (P o M) V, P = softmax(Q K^T * scale), M = dropout multipliers.  Pass 1 (one warp per query row): recompute P,
// dP = dO V^T, delta = sum_k dP M P, dS = P (dP M - delta) -> workspace, Pm = P M -> workspace, dQ = dS K * scale.
// Pass 2 (one warp per key row): dK = dS^T Q * scale, dV = Pm^T dO.
__global__ void __launch_bounds__(256) mha_bwd_q_f32_kernel(const float* __restrict__ q, long long q_sb, long long q_sr,
                                                            const float* __restrict__ k, long long k_sb, long long k_sr,
                                                            const float* __restrict__ v, long long v_sb, long long v_sr,
                                                            const float* __restrict__ dout, long long o_sb, long long o_sr,
                                                            float* __restrict__ dq, long long dq_sb, long long dq_sr,
                                                            float* __restrict__ ws_ds, float* __restrict__ ws_pm,
                                                            int heads, int Tq, int Tk, int hd, float scale,
                                                            const unsigned long long* drop_rng, float drop_p,
                                                            unsigned drop_site) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* pw = sm + warp * 2 * Tk;   // P row
  float* dp = pw + Tk;              // dS row
  const int b = blockIdx.z, h = blockIdx.y;
  const int r = blockIdx.x * nwarps + warp;
  if (r >= Tq) return;
  const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);
  const float* qp = q + b * q_sb + r * q_sr + h * hd;
  const float* dop = dout + b * o_sb + r * o_sr + h * hd;
  float m = -INFINITY;
  for (int c = lane; c < Tk; c += 32) {
    const float* kp = k + b * k_sb + c * k_sr + h * hd;
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(qp[d], kp[d], a);
    a *= scale;
    pw[c] = a;
    m = fmaxf(m, a);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < Tk; c += 32) {
    const float e = expf(pw[c] - m);
    pw[c] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  const unsigned long long base = ((static_cast<unsigned long long>(b) * heads + h) * Tq + r) * Tk;
  float delta = 0.f;
  for (int c = lane; c < Tk; c += 32) {
    const float p = pw[c] * inv;
    const float* vp = v + b * v_sb + c * v_sr + h * hd;
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(dop[d], vp[d], a);
    const float mk = drop.on ? drop_mult(drop, base + c) : 1.f;
    pw[c] = p;
    dp[c] = a * mk;            // dP (through the dropout)
    delta = fmaf(a * mk, p, delta);
    ws_pm[base + c] = p * mk;
  }
  delta = warp_sum(delta);
  __syncwarp();
  for (int c = lane; c < Tk; c += 32) {
    const float ds = pw[c] * (dp[c] - delta);
    dp[c] = ds;
    ws_ds[base + c] = ds;
  }
  __syncwarp();
  for (int d = lane; d < hd; d += 32) {
    float a = 0.f;
    for (int c = 0; c < Tk; ++c) a = fmaf(dp[c], k[b * k_sb + c * k_sr + h * hd + d], a);
    dq[b * dq_sb + r * dq_sr + h * hd + d] = a * scale;
  }
}

__global__ void __launch_bounds__(256) mha_bwd_kv_f32_kernel(const float* __restrict__ q, long long q_sb, long long q_sr,
                                                             const float* __restrict__ dout, long long o_sb, long long o_sr,
                                                             float* __restrict__ dk, long long dk_sb, long long dk_sr,
                                                             float* __restrict__ dv, long long dv_sb, long long dv_sr,
                                                             const float* __restrict__ ws_ds,
                                                             const float* __restrict__ ws_pm, int accumulate, int heads,
                                                             int Tq, int Tk, int hd, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int c = blockIdx.x * nwarps + warp;
  if (c >= Tk) return;
  const unsigned long long base = (static_cast<unsigned long long>(b) * heads + h) * Tq * Tk;
  for (int d = lane; d < hd; d += 32) {
    float ak = 0.f, av = 0.f;
    for (int r = 0; r < Tq; ++r) {
      ak = fmaf(ws_ds[base + static_cast<unsigned long long>(r) * Tk + c], q[b * q_sb + r * q_sr + h * hd + d], ak);
      av = fmaf(ws_pm[base + static_cast<unsigned long long>(r) * Tk + c], dout[b * o_sb + r * o_sr + h * hd + d], av);
    }
    float* kp = dk + b * dk_sb + c * dk_sr + h * hd + d;
    float* vp = dv + b * dv_sb + c * dv_sr + h * hd + d;
    *kp = accumulate ? *kp + ak * scale : ak * scale;
    *vp = accumulate ? *vp + av : av;
  }
}

}  // namespace

#define STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

static inline long long chunks_for(long long npix, long long elems, long long* per) {
  // enough (element, pixel-chunk) work items to fill the machine, at most 256 chunks
  long long want = (148ll * 2048 * 4) / (elems > 0 ? elems : 1);
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  if (want > npix) want = npix > 0 ? npix : 1;
  *per = ceil_div_ll(npix, want);
  return ceil_div_ll(npix, *per);
}

extern "C" int tfpp_conv_wgrad_f32(const tfpp_wgrad_args* a, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(a != nullptr && a->dy != nullptr && a->x != nullptr && a->dw != nullptr, "null operand");
  TFPP_CHECK_ARG(a->group_width == 0, "the fp32 path has a dedicated group-conv weight gradient");
  TFPP_CHECK_ARG(a->ntaps >= 1 && a->ntaps <= 9, "1..9 taps");
  WgP p;
  p.dy = static_cast<const float*>(a->dy);
  p.x = static_cast<const float*>(a->x);
  p.dw = a->dw;
  p.batch = a->batch; p.height = a->height; p.width = a->width; p.cout = a->cout;
  p.cout_valid = a->cout_valid > 0 ? a->cout_valid : a->cout;
  p.x_batch = a->x_batch; p.x_channels = a->x_channels; p.cin = a->cin; p.ntaps = a->ntaps;
  p.x_batch_stride = a->x_batch_stride;
  p.s_co = a->dw_s_co; p.s_tap = a->dw_s_tap; p.s_ci = a->dw_s_ci;
  for (int i = 0; i < 9; ++i) {
    p.tap_dx[i] = a->tap_dx[i]; p.tap_dy[i] = a->tap_dy[i]; p.tap_db[i] = a->tap_db[i]; p.tap_w[i] = a->tap_w[i];
  }
  const long long elems = static_cast<long long>(p.cout_valid) * p.ntaps * p.cin;
  const long long npix = static_cast<long long>(a->batch) * a->height * a->width;
  const long long nchunks = chunks_for(npix, elems, &p.pix_per_chunk);
  dim3 grid(static_cast<unsigned>(ceil_div_ll(elems, 256)), static_cast<unsigned>(nchunks));
  conv_wgrad_f32_kernel<<<grid, 256, 0, stream>>>(p);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_gconv3x3_wgrad_f32(const float* dy, const float* x, float* dw, int batch, int height, int width,
                                       int channels, int stride, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 24 == 0 && (stride == 1 || stride == 2), "group width 24, stride 1 or 2");
  const long long elems = static_cast<long long>(channels) * 24 * 9;
  const long long npix = static_cast<long long>(batch) * (height / stride) * (width / stride);
  long long per;
  const long long nchunks = chunks_for(npix, elems, &per);
  dim3 grid(static_cast<unsigned>(ceil_div_ll(elems, 256)), static_cast<unsigned>(nchunks));
  gconv_wgrad_f32_kernel<<<grid, 256, 0, stream>>>(dy, x, dw, batch, height, width, channels, stride, per);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_gconv3x3_dgrad_s2_f32(const float* dy, const float* w_t, float* dx, int batch, int out_height,
                                          int out_width, int channels, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 24 == 0, "group width 24");
  const long long total = static_cast<long long>(batch) * 4 * out_height * out_width * channels;
  gconv_dgrad_s2_f32_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(dy, w_t, dx, batch, out_height,
                                                                                                 out_width, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_stem_wgrad_f32(const float* x, const float* draw, const float* in_scale, const float* in_shift,
                                   float* dw, int batch, int cin, int height, int width, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(cin >= 1 && cin <= 3, "stem supports 1..3 input channels");
  const long long npix = static_cast<long long>(batch) * (height / 2) * (width / 2);
  long long per;
  const long long nchunks = chunks_for(npix, 32ll * cin * 9, &per);
  dim3 grid(ceil_div(32 * cin * 9, 256), static_cast<unsigned>(nchunks));
  stem_wgrad_f32_kernel<<<grid, 256, 0, stream>>>(x, draw, in_scale, in_shift, dw, batch, cin, height, width, per);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bn_bwd_f32(const float* dy, const float* y, const float* raw, const float* mean, const float* invstd,
                               const float* gamma, const float* gate, const float* pool_grad, const float* fwd_scale,
                               const float* fwd_shift, int act, float* s1, float* s2, float* draw, float* dz_out,
                               int batch, int hw, int channels, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(act != ACT_RELU || y != nullptr || (fwd_scale != nullptr && fwd_shift != nullptr),
                 "ReLU mask needs y, or the forward scale/shift to recompute it from raw");
  int chunks = TFPP_NUM_SMS * 4 / (batch > 0 ? batch : 1);
  if (chunks < 1) chunks = 1;
  int ppb = ceil_div(hw, chunks);
  if (ppb < 4) ppb = 4;
  chunks = ceil_div(hw, ppb);
  bn_bwd_reduce_f32_kernel<<<dim3(chunks, batch), 256, 0, stream>>>(dy, y, raw, mean, invstd, gate, pool_grad, fwd_scale,
                                                                   fwd_shift, act, s1, s2, hw, channels, ppb);
  TFPP_CHECK_LAUNCH();
  const long long total = static_cast<long long>(batch) * hw * channels;
  bn_bwd_apply_f32_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      dy, y, raw, mean, invstd, gamma, s1, s2, gate, pool_grad, fwd_scale, fwd_shift, act,
      1.f / (static_cast<float>(batch) * static_cast<float>(hw)), draw, dz_out, total, hw, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_se_bwd_reduce_f32(const float* dout, const float* a2, float* dgate_sum, int batch, int hw,
                                      int channels, tfpp_stream_t stream_) {
  STREAM;
  int chunks = TFPP_NUM_SMS * 4 / (batch > 0 ? batch : 1);
  if (chunks < 1) chunks = 1;
  int ppb = ceil_div(hw, chunks);
  if (ppb < 4) ppb = 4;
  chunks = ceil_div(hw, ppb);
  se_bwd_reduce_f32_kernel<<<dim3(chunks, batch), 256, 0, stream>>>(dout, a2, dgate_sum, hw, channels, ppb);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_act_bwd_f32(const float* dy, const float* y, int nchw, int act, int act_n_limit, float dy_scale,
                                float* dz, float* dbias, int batch, int hw, int channels, int channels_padded,
                                const unsigned long long* drop_rng, float drop_p, unsigned drop_site,
                                tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(drop_rng == nullptr || drop_p == 0.f || nchw != 1, "dropout adjoint: NHWC layouts only");
  const long long npix = static_cast<long long>(batch) * hw;
  long long blocks = ceil_div_ll(npix * channels_padded, 256);
  if (blocks > TFPP_NUM_SMS * 16) blocks = TFPP_NUM_SMS * 16;
  if (blocks < 1) blocks = 1;
  act_bwd_f32_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(dy, y, nchw == 1, act, act_n_limit, dy_scale, dz, dbias,
                                                                       npix, hw, channels, channels_padded,
                                                                       drop_p > 0.f ? drop_rng : nullptr, drop_p, drop_site);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bilinear_bwd_f32(const float* dout, float* dsrc, long long src_batch_stride, long long src_row_stride,
                                     int accumulate, int batch, int sh, int sw, int dh, int dw, int channels,
                                     tfpp_stream_t stream_) {
  STREAM;
  if (!accumulate) {
    const long long tz = static_cast<long long>(batch) * sh * sw * channels;
    zero_slab_f32_kernel<<<static_cast<unsigned>(ceil_div_ll(tz, 256)), 256, 0, stream>>>(dsrc, src_batch_stride, src_row_stride,
                                                                                          batch, sh * sw, channels);
    TFPP_CHECK_LAUNCH();
  }
  const long long total = static_cast<long long>(batch) * dh * dw * channels;
  bilinear_bwd_f32_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      dout, dsrc, src_batch_stride, src_row_stride, batch, sh, sw, dh, dw, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bilinear_nchw_mask_bwd_f32(const float* dout, const float* mask, float* dsrc, int batch, int sh, int sw,
                                               int src_channels, int channels, int dh, int dw, tfpp_stream_t stream_) {
  STREAM;
  cudaError_t e = cudaMemsetAsync(dsrc, 0, sizeof(float) * static_cast<size_t>(batch) * sh * sw * src_channels, stream);
  if (e != cudaSuccess) {
    tfpp_set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return TFPP_ERR_CUDA;
  }
  const long long total = static_cast<long long>(batch) * channels * dh * dw;
  bilinear_nchw_mask_bwd_f32_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      dout, mask, dsrc, batch, sh, sw, src_channels, channels, dh, dw);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_pool_bwd_add_f32(const float* dout, const float* dtok, float* out, int batch, int height, int width,
                                     int channels, int ph, int pw, int rows_per_batch, int row0, tfpp_stream_t stream_) {
  STREAM;
  const long long total = static_cast<long long>(batch) * height * width * channels;
  pool_bwd_add_f32_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(dout, dtok, out, batch, height,
                                                                                               width, channels, ph, pw,
                                                                                               rows_per_batch, row0);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_add_f32(const float* a, const float* b, float* y, long long n, tfpp_stream_t stream_) {
  STREAM;
  add_f32_kernel<<<static_cast<unsigned>(ceil_div_ll(n, 256)), 256, 0, stream>>>(a, b, y, n);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_copy_rows_f32(const float* x, float* out, float* dbias, int groups, int group_rows, int row0,
                                  int rows, int channels, tfpp_stream_t stream_) {
  STREAM;
  long long blocks = ceil_div_ll(static_cast<long long>(groups) * rows * channels, 256);
  if (blocks > TFPP_NUM_SMS * 8) blocks = TFPP_NUM_SMS * 8;
  if (blocks < 1) blocks = 1;
  copy_rows_f32_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, out, dbias, groups, group_rows, row0, rows,
                                                                         channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_mha_bwd_f32(const float* q, long long q_sb, long long q_sr, const float* k, long long k_sb,
                                long long k_sr, const float* v, long long v_sb, long long v_sr, const float* dout,
                                long long o_sb, long long o_sr, float* dq, long long dq_sb, long long dq_sr, float* dk,
                                long long dk_sb, long long dk_sr, float* dv, long long dv_sb, long long dv_sr,
                                float* workspace, int accumulate_kv, int batch, int heads, int tq, int tk, int head_dim,
                                const unsigned long long* drop_rng, float drop_p, unsigned drop_site,
                                tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(tk >= 1 && tk <= 1024 && workspace != nullptr, "mha_bwd_f32: 1 <= tk <= 1024, workspace of 2*B*heads*tq*tk floats");
  const int warps = 8;
  const float scale = 1.0f / sqrtf(static_cast<float>(head_dim));
  float* ws_ds = workspace;
  float* ws_pm = workspace + static_cast<long long>(batch) * heads * tq * tk;
  const unsigned long long* rng = (drop_rng != nullptr && drop_p > 0.f) ? drop_rng : nullptr;
  mha_bwd_q_f32_kernel<<<dim3(ceil_div(tq, warps), heads, batch), warps * 32, sizeof(float) * warps * 2 * tk, stream>>>(
      q, q_sb, q_sr, k, k_sb, k_sr, v, v_sb, v_sr, dout, o_sb, o_sr, dq, dq_sb, dq_sr, ws_ds, ws_pm, heads, tq, tk, head_dim,
      scale, rng, drop_p, drop_site);
  TFPP_CHECK_LAUNCH();
  mha_bwd_kv_f32_kernel<<<dim3(ceil_div(tk, warps), heads, batch), warps * 32, 0, stream>>>(
      q, q_sb, q_sr, dout, o_sb, o_sr, dk, dk_sb, dk_sr, dv, dv_sb, dv_sr, ws_ds, ws_pm, accumulate_kv, heads, tq, tk, head_dim,
      scale);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
