// Backward kernels of the token-side ops (adjoints of transformer.cu): LayerNorm, fusion self-attention, decoder
// attention, extra-sensor token MLP, GRU path decoder + target-speed MLP.  They replace the autograd graph of
// team_code/train.py:898 over transfuser.py:362-402 and model.py:299-358,857-867.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// y = xhat * gamma + beta, xhat = (x - mean) * rstd.  g = dy * gamma;
// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)) (+ dres);  dgamma += dy * xhat;  dbeta += dy.
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ dy, int dy_f32,
                                                            const float* __restrict__ x,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ dres, float* __restrict__ dx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int rows, int C) {
  extern __shared__ float sm[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row < rows) {
    const long long base = static_cast<long long>(row) * C;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = dy_f32 ? static_cast<const float*>(dy)[base + c] : bf2f(static_cast<const bf16*>(dy)[base + c]);
      const float xh = (x[base + c] - mu) * rs;
      const float g = d * gamma[c];
      s1 += g;
      s2 = fmaf(g, xh, s2);
      atomicAdd(&sm[c], d * xh);
      atomicAdd(&sm[C + c], d);
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    for (int c = lane; c < C; c += 32) {
      const float d = dy_f32 ? static_cast<const float*>(dy)[base + c] : bf2f(static_cast<const bf16*>(dy)[base + c]);
      const float xh = (x[base + c] - mu) * rs;
      float v = rs * (d * gamma[c] - s1 - xh * s2);
      if (dres) v += dres[base + c];
      dx[base + c] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, sm[i]);
    atomicAdd(dbeta + i, sm[C + i]);
  }
}

// ------------------------------------------------------------------------------------------------ decoder attention bwd
// One CTA per (batch, head); everything in fp32 shared memory (Tq <= 16, Tk <= 32 * NPL, hd <= 64).  NPL = keys per lane:
// 4 for the 65-token memory of the TransFuser backbone, 16 for the 257-token memory of the bev_encoder backbone.
template <int NPL>
__global__ void __launch_bounds__(256) small_mha_bwd_kernel(const bf16* __restrict__ q, long long q_sb, long long q_sr,
                                                            const bf16* __restrict__ k, long long k_sb, long long k_sr,
                                                            const bf16* __restrict__ v, long long v_sb, long long v_sr,
                                                            const bf16* __restrict__ dout, long long o_sb, long long o_sr,
                                                            bf16* __restrict__ dq, long long dq_sb, long long dq_sr,
                                                            bf16* __restrict__ dk, long long dk_sb, long long dk_sr,
                                                            bf16* __restrict__ dv, long long dv_sb, long long dv_sr,
                                                            int accumulate_kv, int Tq, int Tk, int hd, float scale,
                                                            const unsigned long long* drop_rng, float drop_p,
                                                            unsigned drop_site) {
  // probability dropout (nn.MultiheadAttention(dropout=0.1)): dV = (P o M')^T dO, dP = (dO V^T) o M', M' = mask / (1-p)
  const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);
  extern __shared__ float sm[];
  const int P1 = hd + 1;
  float* qs = sm;                  // Tq x P1
  float* ks = qs + Tq * P1;        // Tk x P1
  float* vs = ks + Tk * P1;        // Tk x P1
  float* dos = vs + Tk * P1;       // Tq x P1
  float* ps = dos + Tq * P1;       // Tq x Tk   (P, then dS)
  const int b = blockIdx.x, h = blockIdx.y;
  for (int i = threadIdx.x; i < Tq * hd; i += blockDim.x) {
    const int r = i / hd, d = i % hd;
    qs[r * P1 + d] = bf2f(q[b * q_sb + r * q_sr + h * hd + d]);
    dos[r * P1 + d] = bf2f(dout[b * o_sb + r * o_sr + h * hd + d]);
  }
  for (int i = threadIdx.x; i < Tk * hd; i += blockDim.x) {
    const int r = i / hd, d = i % hd;
    ks[r * P1 + d] = bf2f(k[b * k_sb + r * k_sr + h * hd + d]);
    vs[r * P1 + d] = bf2f(v[b * v_sb + r * v_sr + h * hd + d]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < Tq; r += nw) {
    float* pr = ps + r * Tk;
    float m = -INFINITY;
    for (int c = lane; c < Tk; c += 32) {
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(qs[r * P1 + d], ks[c * P1 + d], a);
      a *= scale;
      pr[c] = a;
      m = fmaxf(m, a);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < Tk; c += 32) {
      const float e = __expf(pr[c] - m);
      pr[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    __syncwarp();
    // dP = dO V^T ; rowsum(dP * P)
    float rsum = 0.f;
    float dpl[NPL];  // Tk <= 32 * NPL -> up to NPL per lane
    float dml[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) dml[i] = 1.f;
    int n = 0;
    const unsigned long long dbase = ((static_cast<unsigned long long>(b) * gridDim.y + h) * Tq + r) * Tk;
#pragma unroll
    for (n = 0; n < NPL; ++n) {
      const int c = lane + 32 * n;
      if (c >= Tk) break;
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(dos[r * P1 + d], vs[c * P1 + d], a);
      const float p = pr[c] * inv;
      pr[c] = p;
      if (drop.on) {
        dml[n] = drop_mult(drop, dbase + c);
        a *= dml[n];
      }
      dpl[n] = a;
      rsum = fmaf(a, p, rsum);
    }
    rsum = warp_sum(rsum);
    n = 0;
    __syncwarp();
    // keep P in a register copy for dV: store dS in ps after use -> need P too: pack dS into a second pass
#pragma unroll
    for (n = 0; n < NPL; ++n) {
      const int c = lane + 32 * n;
      if (c >= Tk) break;
      const float p = pr[c];
      dpl[n] = p * (dpl[n] - rsum) * scale;  // dS
      if (drop.on) pr[c] = p * dml[n];       // the dV pass below multiplies the DROPPED probabilities
    }
    // dV += P^T dO is accumulated after the loop (needs all rows) -> store P in place and dS in registers -> write dQ now
    __syncwarp();
    // dQ[r] = dS[r,:] K
    // stash dS into shared by swapping with P: we still need P for dV, so use two-step: write dS to a side buffer
    float* dsr = ps + Tq * Tk + r * Tk;  // second Tq x Tk block
#pragma unroll
    for (n = 0; n < NPL; ++n)
      if (lane + 32 * n < Tk) dsr[lane + 32 * n] = dpl[n];
    __syncwarp();
    for (int d = lane; d < hd; d += 32) {
      float a = 0.f;
      for (int c = 0; c < Tk; ++c) a = fmaf(dsr[c], ks[c * P1 + d], a);
      dq[b * dq_sb + r * dq_sr + h * hd + d] = f2bf(a);
    }
  }
  __syncthreads();
  const float* ds = ps + Tq * Tk;
  for (int i = threadIdx.x; i < Tk * hd; i += blockDim.x) {
    const int c = i / hd, d = i % hd;
    float ak = 0.f, av = 0.f;
    for (int r = 0; r < Tq; ++r) {
      ak = fmaf(ds[r * Tk + c], qs[r * P1 + d], ak);
      av = fmaf(ps[r * Tk + c], dos[r * P1 + d], av);
    }
    bf16* pk = dk + b * dk_sb + c * dk_sr + h * hd + d;
    bf16* pv = dv + b * dv_sb + c * dv_sr + h * hd + d;
    if (accumulate_kv) {
      ak += bf2f(*pk);
      av += bf2f(*pv);
    }
    *pk = f2bf(ak);
    *pv = f2bf(av);
  }
}

// ------------------------------------------------------------------------------------------------ small heads bwd
__device__ __forceinline__ float wdot(const float* __restrict__ w, const float* __restrict__ x, int n, int lane) {
  float a = 0.f;
  for (int i = lane; i < n; i += 32) a = fmaf(__ldg(w + i), x[i], a);
  return warp_sum(a);
}

// extra-sensor token backward (model.py:308-319): dmem_row (B, d_model) f32 -> parameter gradients.
__global__ void __launch_bounds__(256) extra_sensor_bwd_kernel(const float* __restrict__ ego_vel,
                                                               const float* __restrict__ command, float vel_mean,
                                                               float vel_invstd, int use_batch_stats, float eps,
                                                               const float* __restrict__ w0, const float* __restrict__ b0,
                                                               const float* __restrict__ w1, const float* __restrict__ b1,
                                                               const float* __restrict__ dmem, long long dmem_stride,
                                                               float* __restrict__ dw0, float* __restrict__ db0,
                                                               float* __restrict__ dw1, float* __restrict__ db1,
                                                               float* __restrict__ dpos, int B, int n_cmd, int hidden,
                                                               int d_model) {
  extern __shared__ float sm[];
  float* in = sm;                 // 8
  float* hid = sm + 8;            // hidden
  float* dpre1 = hid + hidden;    // d_model
  float* dhid = dpre1 + d_model;  // hidden
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (warp == 0) {
    float mean = vel_mean, invstd = vel_invstd;
    if (use_batch_stats) {
      float s = 0.f;
      for (int i = lane; i < B; i += 32) s += ego_vel[i];
      mean = warp_sum(s) / B;
      float qv = 0.f;
      for (int i = lane; i < B; i += 32) {
        const float d = ego_vel[i] - mean;
        qv = fmaf(d, d, qv);
      }
      invstd = rsqrtf(warp_sum(qv) / B + eps);
    }
    if (lane == 0) in[0] = (ego_vel[b] - mean) * invstd;
    if (lane >= 1 && lane <= n_cmd) in[lane] = command[b * n_cmd + lane - 1];
  }
  __syncthreads();
  const int n_in = 1 + n_cmd;
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
    float a = b0[j];
    for (int i = 0; i < n_in; ++i) a = fmaf(w0[j * n_in + i], in[i], a);
    hid[j] = fmaxf(a, 0.f);
  }
  __syncthreads();
  for (int j = warp; j < d_model; j += nw) {
    const float a = wdot(w1 + static_cast<long long>(j) * hidden, hid, hidden, lane) + b1[j];
    if (lane == 0) {
      const float d = dmem[b * dmem_stride + j];
      atomicAdd(dpos + j, d);
      const float dp = a > 0.f ? d : 0.f;
      dpre1[j] = dp;
      atomicAdd(db1 + j, dp);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d_model * hidden; i += blockDim.x) {
    const int j = i / hidden, k = i % hidden;
    atomicAdd(dw1 + i, dpre1[j] * hid[k]);
  }
  for (int k = threadIdx.x; k < hidden; k += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < d_model; ++j) a = fmaf(dpre1[j], w1[static_cast<long long>(j) * hidden + k], a);
    const float dp = hid[k] > 0.f ? a : 0.f;
    dhid[k] = dp;
    atomicAdd(db0 + k, dp);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hidden * n_in; i += blockDim.x) atomicAdd(dw0 + i, dhid[i / n_in] * in[i % n_in]);
}

// planner head backward: GRU BPTT (model.py:857-867) + target-speed MLP (model.py:118-119).  One CTA per sample.
// Inputs: joined (B, n_wp+1, D) f32, h_all (B, n_wp, HS) f32 (hidden states saved by the forward kernel),
// dcp (B, n_wp, 2), dlogits (B, n_speed).  Outputs: djoined (B, n_wp+1, D) f32 and atomically accumulated parameter
// gradients.
__global__ void __launch_bounds__(256) planner_head_bwd_kernel(
    const float* __restrict__ joined, const float* __restrict__ target_point, const float* __restrict__ h_all,
    const float* __restrict__ w_enc, const float* __restrict__ b_enc, const float* __restrict__ w_ih,
    const float* __restrict__ w_hh, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
    const float* __restrict__ w_dec, const float* __restrict__ w_ts0, const float* __restrict__ b_ts0,
    const float* __restrict__ w_ts1, const float* __restrict__ dcp, const float* __restrict__ dlogits,
    float* __restrict__ djoined, float* __restrict__ dw_enc, float* __restrict__ db_enc, float* __restrict__ dw_ih,
    float* __restrict__ dw_hh, float* __restrict__ db_ih, float* __restrict__ db_hh, float* __restrict__ dw_dec,
    float* __restrict__ db_dec, float* __restrict__ dw_ts0, float* __restrict__ db_ts0, float* __restrict__ dw_ts1,
    float* __restrict__ db_ts1, int n_wp, int D, int HS, int n_speed) {
  extern __shared__ float sm[];
  float* x = sm;                          // (n_wp+1) * D
  float* hs = x + (n_wp + 1) * D;         // (n_wp+1) * HS : h_0 .. h_{n_wp}
  float* dgi = hs + (n_wp + 1) * HS;      // n_wp * 3HS
  float* dgh = dgi + n_wp * 3 * HS;       // n_wp * 3HS
  float* dh = dgh + n_wp * 3 * HS;        // HS
  float* gi = dh + HS;                    // 3HS
  float* gh = gi + 3 * HS;                // 3HS
  float* hid = gh + 3 * HS;               // D
  float* dhid = hid + D;                  // D
  float* dout = dhid + D;                 // n_wp * 2 (grad wrt decoder outputs)
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int n_rows = n_wp + (n_speed > 0 ? 1 : 0);   // n_speed == 0: GRU only (wp_decoder), joined has n_wp rows
  for (int i = threadIdx.x; i < n_rows * D; i += blockDim.x) x[i] = joined[static_cast<long long>(b) * n_rows * D + i];
  for (int i = threadIdx.x; i < n_wp * HS; i += blockDim.x) hs[HS + i] = h_all[static_cast<long long>(b) * n_wp * HS + i];
  if (threadIdx.x < HS) {
    const float tx = target_point[b * 2], ty = target_point[b * 2 + 1];
    hs[threadIdx.x] = w_enc[threadIdx.x * 2] * tx + w_enc[threadIdx.x * 2 + 1] * ty + b_enc[threadIdx.x];
    dh[threadIdx.x] = 0.f;
  }
  // reverse cumsum of dcp -> grad wrt the per-step decoder outputs
  if (threadIdx.x < 2) {
    float run = 0.f;
    for (int t = n_wp - 1; t >= 0; --t) {
      run += dcp[(static_cast<long long>(b) * n_wp + t) * 2 + threadIdx.x];
      dout[t * 2 + threadIdx.x] = run;
    }
  }
  __syncthreads();
  for (int t = n_wp - 1; t >= 0; --t) {
    const float* hp = hs + t * HS;        // h_{t-1}
    const float* hc = hs + (t + 1) * HS;  // h_t
    // decoder: out_t = W_dec h_t + b_dec
    if (threadIdx.x < HS) {
      const int j = threadIdx.x;
      dh[j] += w_dec[j] * dout[t * 2] + w_dec[HS + j] * dout[t * 2 + 1];
      atomicAdd(dw_dec + j, dout[t * 2] * hc[j]);
      atomicAdd(dw_dec + HS + j, dout[t * 2 + 1] * hc[j]);
    }
    if (threadIdx.x < 2) atomicAdd(db_dec + threadIdx.x, dout[t * 2 + threadIdx.x]);
    // recompute gate pre-activations
    for (int j = warp; j < 3 * HS; j += nw) {
      const float a = wdot(w_ih + static_cast<long long>(j) * D, x + t * D, D, lane);
      const float c = wdot(w_hh + static_cast<long long>(j) * HS, hp, HS, lane);
      if (lane == 0) {
        gi[j] = a + b_ih[j];
        gh[j] = c + b_hh[j];
      }
    }
    __syncthreads();
    if (threadIdx.x < HS) {
      const int j = threadIdx.x;
      const float r = 1.f / (1.f + __expf(-(gi[j] + gh[j])));
      const float z = 1.f / (1.f + __expf(-(gi[HS + j] + gh[HS + j])));
      const float n = tanhf(gi[2 * HS + j] + r * gh[2 * HS + j]);
      const float d = dh[j];
      const float dn = d * (1.f - z);
      const float dz = d * (hp[j] - n);
      const float dn_pre = dn * (1.f - n * n);
      const float dr = dn_pre * gh[2 * HS + j];
      const float dz_pre = dz * z * (1.f - z);
      const float dr_pre = dr * r * (1.f - r);
      float* gi_t = dgi + t * 3 * HS;
      float* gh_t = dgh + t * 3 * HS;
      gi_t[j] = dr_pre; gi_t[HS + j] = dz_pre; gi_t[2 * HS + j] = dn_pre;
      gh_t[j] = dr_pre; gh_t[HS + j] = dz_pre; gh_t[2 * HS + j] = dn_pre * r;
      dh[j] = d * z;  // direct path to h_{t-1}
    }
    __syncthreads();
    // dh_{t-1} += W_hh^T dgh_t ; dx_t = W_ih^T dgi_t
    if (threadIdx.x < HS) {
      const int k = threadIdx.x;
      float a = 0.f;
      const float* gh_t = dgh + t * 3 * HS;
      for (int j = 0; j < 3 * HS; ++j) a = fmaf(w_hh[static_cast<long long>(j) * HS + k], gh_t[j], a);
      dh[k] += a;
    }
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      float a = 0.f;
      const float* gi_t = dgi + t * 3 * HS;
      for (int j = 0; j < 3 * HS; ++j) a = fmaf(w_ih[static_cast<long long>(j) * D + k], gi_t[j], a);
      djoined[(static_cast<long long>(b) * n_rows + t) * D + k] = a;
    }
    __syncthreads();
  }
  // h_0 = encoder(target_point)
  if (threadIdx.x < HS) {
    const int j = threadIdx.x;
    atomicAdd(dw_enc + j * 2, dh[j] * target_point[b * 2]);
    atomicAdd(dw_enc + j * 2 + 1, dh[j] * target_point[b * 2 + 1]);
    atomicAdd(db_enc + j, dh[j]);
  }
  // parameter gradients of the GRU: sum over steps inside the CTA, one atomic per element per sample
  for (int i = threadIdx.x; i < 3 * HS * D; i += blockDim.x) {
    const int j = i / D, k = i % D;
    float a = 0.f;
    for (int t = 0; t < n_wp; ++t) a = fmaf(dgi[t * 3 * HS + j], x[t * D + k], a);
    atomicAdd(dw_ih + i, a);
  }
  for (int i = threadIdx.x; i < 3 * HS * HS; i += blockDim.x) {
    const int j = i / HS, k = i % HS;
    float a = 0.f;
    for (int t = 0; t < n_wp; ++t) a = fmaf(dgh[t * 3 * HS + j], hs[t * HS + k], a);
    atomicAdd(dw_hh + i, a);
  }
  for (int j = threadIdx.x; j < 3 * HS; j += blockDim.x) {
    float a = 0.f, c = 0.f;
    for (int t = 0; t < n_wp; ++t) {
      a += dgi[t * 3 * HS + j];
      c += dgh[t * 3 * HS + j];
    }
    atomicAdd(db_ih + j, a);
    atomicAdd(db_hh + j, c);
  }
  // target-speed MLP backward
  if (n_speed == 0) return;
  const float* ts = x + n_wp * D;
  for (int j = warp; j < D; j += nw) {
    const float a = wdot(w_ts0 + static_cast<long long>(j) * D, ts, D, lane);
    if (lane == 0) hid[j] = fmaxf(a + b_ts0[j], 0.f);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < n_speed; ++s) a = fmaf(w_ts1[static_cast<long long>(s) * D + j], dlogits[b * n_speed + s], a);
    dhid[j] = hid[j] > 0.f ? a : 0.f;
    atomicAdd(db_ts0 + j, dhid[j]);
    for (int s = 0; s < n_speed; ++s) atomicAdd(dw_ts1 + static_cast<long long>(s) * D + j, dlogits[b * n_speed + s] * hid[j]);
  }
  if (threadIdx.x < n_speed) atomicAdd(db_ts1 + threadIdx.x, dlogits[b * n_speed + threadIdx.x]);
  __syncthreads();
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) atomicAdd(dw_ts0 + i, dhid[i / D] * ts[i % D]);
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < D; ++j) a = fmaf(w_ts0[static_cast<long long>(j) * D + k], dhid[j], a);
    djoined[(static_cast<long long>(b) * (n_wp + 1) + n_wp) * D + k] = a;
  }
}

}  // namespace

#define STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int tfpp_layernorm_bwd(const void* dy, int dy_f32, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, const float* dres, float* dx, float* dgamma, float* dbeta,
                                  int rows, int channels, tfpp_stream_t stream_) {
  STREAM;
  layernorm_bwd_kernel<<<ceil_div(rows, 8), 256, sizeof(float) * 2 * channels, stream>>>(
      dy, dy_f32, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_small_mha_bwd(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb,
                                  long long k_sr, const void* v, long long v_sb, long long v_sr, const void* dout,
                                  long long o_sb, long long o_sr, void* dq, long long dq_sb, long long dq_sr, void* dk,
                                  long long dk_sb, long long dk_sr, void* dv, long long dv_sb, long long dv_sr,
                                  int accumulate_kv, int batch, int heads, int tq, int tk, int head_dim,
                                  tfpp_stream_t stream_) {
  return tfpp_small_mha_bwd_dropout(q, q_sb, q_sr, k, k_sb, k_sr, v, v_sb, v_sr, dout, o_sb, o_sr, dq, dq_sb, dq_sr, dk,
                                    dk_sb, dk_sr, dv, dv_sb, dv_sr, accumulate_kv, batch, heads, tq, tk, head_dim,
                                    nullptr, 0.f, 0u, stream_);
}

extern "C" int tfpp_small_mha_bwd_dropout(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb,
                                          long long k_sr, const void* v, long long v_sb, long long v_sr,
                                          const void* dout, long long o_sb, long long o_sr, void* dq, long long dq_sb,
                                          long long dq_sr, void* dk, long long dk_sb, long long dk_sr, void* dv,
                                          long long dv_sb, long long dv_sr, int accumulate_kv, int batch, int heads,
                                          int tq, int tk, int head_dim, const unsigned long long* drop_rng, float drop_p,
                                          unsigned drop_site, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(tk <= 512 && tq <= 16 && head_dim <= 64, "small_mha_bwd: tq <= 16, tk <= 512, head_dim <= 64");
  const size_t smem = sizeof(float) * ((2 * tq + 2 * tk) * (head_dim + 1) + 2 * tq * tk);
  TFPP_CHECK_ARG(smem <= 200 * 1024, "small_mha_bwd: shared-memory budget exceeded");
  dim3 grid(batch, heads);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(small_mha_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(small_mha_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  auto kern = tk <= 128 ? small_mha_bwd_kernel<4> : small_mha_bwd_kernel<16>;
  kern<<<grid, 256, smem, stream>>>(
      static_cast<const bf16*>(q), q_sb, q_sr, static_cast<const bf16*>(k), k_sb, k_sr, static_cast<const bf16*>(v), v_sb,
      v_sr, static_cast<const bf16*>(dout), o_sb, o_sr, static_cast<bf16*>(dq), dq_sb, dq_sr, static_cast<bf16*>(dk),
      dk_sb, dk_sr, static_cast<bf16*>(dv), dv_sb, dv_sr, accumulate_kv, tq, tk, head_dim,
      1.0f / sqrtf(static_cast<float>(head_dim)), drop_p > 0.f ? drop_rng : nullptr, drop_p, drop_site);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_extra_sensor_token_bwd(const float* ego_vel, const float* command, float vel_mean, float vel_var,
                                           int use_batch_stats, const float* w0, const float* b0, const float* w1,
                                           const float* b1, const float* dmem, long long dmem_stride, float* dw0,
                                           float* db0, float* dw1, float* db1, float* dpos, int batch, int n_cmd,
                                           int hidden, int d_model, tfpp_stream_t stream_) {
  STREAM;
  const float eps = 1e-5f;
  extra_sensor_bwd_kernel<<<batch, 256, sizeof(float) * (8 + 2 * hidden + d_model), stream>>>(
      ego_vel, command, vel_mean, rsqrtf(vel_var + eps), use_batch_stats, eps, w0, b0, w1, b1, dmem, dmem_stride, dw0,
      db0, dw1, db1, dpos, batch, n_cmd, hidden, d_model);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_planner_head_bwd(const float* joined, const float* target_point, const float* h_all,
                                     const float* w_enc, const float* b_enc, const float* w_ih, const float* w_hh,
                                     const float* b_ih, const float* b_hh, const float* w_dec, const float* w_ts0,
                                     const float* b_ts0, const float* w_ts1, const float* dcp, const float* dlogits,
                                     float* djoined, float* dw_enc, float* db_enc, float* dw_ih, float* dw_hh,
                                     float* db_ih, float* db_hh, float* dw_dec, float* db_dec, float* dw_ts0,
                                     float* db_ts0, float* dw_ts1, float* db_ts1, int batch, int n_wp, int d_model,
                                     int hidden, int n_speed, tfpp_stream_t stream_) {
  STREAM;
  const size_t smem = sizeof(float) * ((n_wp + 1) * d_model + (n_wp + 1) * hidden + 2 * n_wp * 3 * hidden + hidden +
                                       6 * hidden + 2 * d_model + 2 * n_wp);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(planner_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    attr_set = true;
  }
  planner_head_bwd_kernel<<<batch, 256, smem, stream>>>(joined, target_point, h_all, w_enc, b_enc, w_ih, w_hh, b_ih,
                                                        b_hh, w_dec, w_ts0, b_ts0, w_ts1, dcp, dlogits, djoined, dw_enc,
                                                        db_enc, dw_ih, dw_hh, db_ih, db_hh, dw_dec, db_dec, dw_ts0,
                                                        db_ts0, dw_ts1, db_ts1, n_wp, d_model, hidden, n_speed);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
