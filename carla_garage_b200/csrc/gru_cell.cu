// The original TransFuser planner head (config.transformer_decoder_join = False): GRUWaypointsPredictorTransFuser
// (team_code/model.py:870-913) — an autoregressive nn.GRUCell whose hidden state starts at the joined scene feature and
// whose input is [previous waypoint, target point]; every step adds nn.Linear(hidden, 2) of the new state to the
// waypoint — fused with the target-speed MLP on the same feature (model.py:113-118,369-376).  One CTA per sample, one
// thread per gate row; backward = BPTT with the gates recomputed from the saved hidden states, weight gradients summed
// per CTA in shared memory and added to the parameter gradients once.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

constexpr int kMaxH = 128;

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + expf(-v)); }

struct GruW {
  const float* w_ih;   // (3H, I)  rows [r | z | n]
  const float* w_hh;   // (3H, H)
  const float* b_ih;   // (3H)
  const float* b_hh;   // (3H)
  const float* w_out;  // (2, H)
  const float* b_out;  // (2)
};

// gate pre-activations of one step: gi = W_ih xin + b_ih, gh = W_hh h + b_hh (thread j < 3H owns row j)
__device__ __forceinline__ void gate_rows(const GruW& w, const float* xin, const float* h, float* gi, float* gh, int H,
                                          int I, int j) {
  float a = w.b_ih[j];
  for (int i = 0; i < I; ++i) a = fmaf(w.w_ih[j * I + i], xin[i], a);
  float c = w.b_hh[j];
  const float* wr = w.w_hh + static_cast<long long>(j) * H;
  for (int k = 0; k < H; ++k) c = fmaf(wr[k], h[k], c);
  gi[j] = a;
  gh[j] = c;
}

__global__ void __launch_bounds__(384) gru_cell_head_kernel(const float* __restrict__ joined, int jstride,
                                                            const float* __restrict__ tp, const GruW w,
                                                            const float* __restrict__ w_ts0, const float* __restrict__ b_ts0,
                                                            const float* __restrict__ w_ts1, const float* __restrict__ b_ts1,
                                                            float* __restrict__ wp_out, float* __restrict__ ts_out,
                                                            float* __restrict__ h_all, int T, int H, int I,
                                                            int learn_origin, int n_speed) {
  __shared__ float h[kMaxH], hn[kMaxH], z0[kMaxH], xin[4], gi[3 * kMaxH], gh[3 * kMaxH], hid[kMaxH];
  const int b = blockIdx.x, j = threadIdx.x;
  const float* jp = joined + static_cast<long long>(b) * jstride;
  if (j < H) {
    h[j] = jp[j];
    z0[j] = jp[j];
    if (h_all) h_all[(static_cast<long long>(b) * (T + 1)) * H + j] = jp[j];
  }
  if (j < 2) xin[j] = learn_origin ? jp[H + j] : 0.f;
  if (j >= 2 && j < I) xin[j] = tp[b * 2 + (j - 2)];
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    if (j < 3 * H) gate_rows(w, xin, h, gi, gh, H, I, j);
    __syncthreads();
    if (j < H) {
      const float r = sigm(gi[j] + gh[j]);
      const float z = sigm(gi[H + j] + gh[H + j]);
      const float n = tanhf(gi[2 * H + j] + r * gh[2 * H + j]);
      hn[j] = (1.f - z) * n + z * h[j];
    }
    __syncthreads();
    if (j < H) {
      h[j] = hn[j];
      if (h_all) h_all[(static_cast<long long>(b) * (T + 1) + t + 1) * H + j] = hn[j];
    }
    if (j < 2) {   // x += Linear(h')  (model.py:905-907)
      float a = w.b_out[j];
      for (int k = 0; k < H; ++k) a = fmaf(w.w_out[j * H + k], hn[k], a);
      xin[j] += a;
      wp_out[(static_cast<long long>(b) * T + t) * 2 + j] = xin[j];
    }
    __syncthreads();
  }
  if (ts_out != nullptr) {   // target_speed_network on the first H features (model.py:371,376)
    if (j < H) {
      float a = b_ts0[j];
      for (int k = 0; k < H; ++k) a = fmaf(w_ts0[j * H + k], z0[k], a);
      hid[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (j < n_speed) {
      float a = b_ts1[j];
      for (int k = 0; k < H; ++k) a = fmaf(w_ts1[j * H + k], hid[k], a);
      ts_out[b * n_speed + j] = a;
    }
  }
}

// BPTT.  dwp (B,T,2), dts (B,n_speed) or NULL -> djoined (B, jstride) (+=: several heads share the joined feature) and
// the parameter gradients (+=).
__global__ void __launch_bounds__(384) gru_cell_head_bwd_kernel(
    const float* __restrict__ joined, int jstride, const float* __restrict__ tp, const GruW w,
    const float* __restrict__ w_ts0, const float* __restrict__ b_ts0, const float* __restrict__ w_ts1,
    const float* __restrict__ wp, const float* __restrict__ h_all, const float* __restrict__ dwp,
    const float* __restrict__ dts, float* __restrict__ djoined, float* __restrict__ dw_ih, float* __restrict__ dw_hh,
    float* __restrict__ db_ih, float* __restrict__ db_hh, float* __restrict__ dw_out, float* __restrict__ db_out,
    float* __restrict__ dw_ts0, float* __restrict__ db_ts0, float* __restrict__ dw_ts1, float* __restrict__ db_ts1, int T,
    int H, int I, int learn_origin, int n_speed) {
  extern __shared__ float sm[];
  float* s_dwhh = sm;                    // [3H][H] per-CTA weight-gradient sums
  float* s_dwih = s_dwhh + 3 * H * H;    // [3H][I]
  float* s_dbih = s_dwih + 3 * H * I;    // [3H]
  float* s_dbhh = s_dbih + 3 * H;        // [3H]
  float* s_dwout = s_dbhh + 3 * H;       // [2][H]
  __shared__ float hp[kMaxH], hc[kMaxH], xin[4], gi[3 * kMaxH], gh[3 * kMaxH], dgi[3 * kMaxH], dgh[3 * kMaxH];
  __shared__ float dh[kMaxH], dhn[kMaxH], G[2], dxin[4], dbo[2], hid[kMaxH], dhid[kMaxH];
  const int b = blockIdx.x, j = threadIdx.x, nt = blockDim.x;
  for (int i = j; i < 3 * H * H + 3 * H * I + 6 * H + 2 * H; i += nt) sm[i] = 0.f;
  if (j < H) dh[j] = 0.f;
  if (j < 2) {
    G[j] = 0.f;
    dbo[j] = 0.f;
  }
  if (j < 4) dxin[j] = 0.f;
  const float* jp = joined + static_cast<long long>(b) * jstride;
  __syncthreads();
  for (int t = T; t >= 1; --t) {
    // state before / after step t, input of step t
    if (j < H) {
      hp[j] = h_all[(static_cast<long long>(b) * (T + 1) + t - 1) * H + j];
      hc[j] = h_all[(static_cast<long long>(b) * (T + 1) + t) * H + j];
    }
    if (j < 2) {
      xin[j] = t >= 2 ? wp[(static_cast<long long>(b) * T + t - 2) * 2 + j] : (learn_origin ? jp[H + j] : 0.f);
      // gradient of waypoint t: its own output + the carry x_{t+1} = x_t + ... + its use as input of step t+1
      G[j] = dwp[(static_cast<long long>(b) * T + t - 1) * 2 + j] + G[j] + dxin[j];
    }
    if (j >= 2 && j < I) xin[j] = tp[b * 2 + (j - 2)];
    __syncthreads();
    if (j < 3 * H) gate_rows(w, xin, hp, gi, gh, H, I, j);
    if (j < H) {   // dh_t = W_out^T G_t + what step t+1 sent back
      dhn[j] = dh[j] + w.w_out[j] * G[0] + w.w_out[H + j] * G[1];
      s_dwout[j] += G[0] * hc[j];
      s_dwout[H + j] += G[1] * hc[j];
    }
    if (j < 2) dbo[j] += G[j];
    __syncthreads();
    if (j < H) {
      const float r = sigm(gi[j] + gh[j]);
      const float z = sigm(gi[H + j] + gh[H + j]);
      const float n = tanhf(gi[2 * H + j] + r * gh[2 * H + j]);
      const float d = dhn[j];
      const float dn_pre = d * (1.f - z) * (1.f - n * n);
      const float dz_pre = d * (hp[j] - n) * z * (1.f - z);
      const float dr_pre = dn_pre * gh[2 * H + j] * r * (1.f - r);
      dgi[j] = dr_pre;
      dgi[H + j] = dz_pre;
      dgi[2 * H + j] = dn_pre;
      dgh[j] = dr_pre;
      dgh[H + j] = dz_pre;
      dgh[2 * H + j] = dn_pre * r;
      dh[j] = d * z;   // direct path h_{t-1} -> h_t; the W_hh path is added below
    }
    __syncthreads();
    if (j < 3 * H) {
      s_dbih[j] += dgi[j];
      s_dbhh[j] += dgh[j];
      for (int i = 0; i < I; ++i) s_dwih[j * I + i] += dgi[j] * xin[i];
      float* row = s_dwhh + static_cast<long long>(j) * H;
      const float g = dgh[j];
      for (int k = 0; k < H; ++k) row[k] += g * hp[k];
    }
    __syncthreads();
    if (j < H) {   // dh_{t-1} += W_hh^T dgh
      float a = 0.f;
      for (int q = 0; q < 3 * H; ++q) a = fmaf(w.w_hh[static_cast<long long>(q) * H + j], dgh[q], a);
      dh[j] += a;
    }
    if (j < I) {   // gradient of the step's input [x_{t-1}, target point]
      float a = 0.f;
      for (int q = 0; q < 3 * H; ++q) a = fmaf(w.w_ih[q * I + j], dgi[q], a);
      dxin[j] = a;
    }
    __syncthreads();
  }
  // x_0 and h_0 are slices of the joined feature
  float* dj = djoined + static_cast<long long>(b) * jstride;
  if (j < H) dhid[j] = 0.f;
  __syncthreads();
  if (dts != nullptr) {   // target-speed MLP backward
    if (j < H) {
      float a = b_ts0[j];
      for (int k = 0; k < H; ++k) a = fmaf(w_ts0[j * H + k], jp[k], a);
      hid[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (j < H) {
      float a = 0.f;
      for (int c = 0; c < n_speed; ++c) a = fmaf(w_ts1[c * H + j], dts[b * n_speed + c], a);
      dhid[j] = hid[j] > 0.f ? a : 0.f;
      for (int c = 0; c < n_speed; ++c) atomicAdd(dw_ts1 + c * H + j, dts[b * n_speed + c] * hid[j]);
    }
    if (j < n_speed) atomicAdd(db_ts1 + j, dts[b * n_speed + j]);
    __syncthreads();
    if (j < H) {
      atomicAdd(db_ts0 + j, dhid[j]);
      for (int k = 0; k < H; ++k) atomicAdd(dw_ts0 + j * H + k, dhid[j] * jp[k]);
    }
    __syncthreads();
  }
  if (j < H) {
    float a = dh[j];
    if (dts != nullptr)
      for (int q = 0; q < H; ++q) a = fmaf(w_ts0[q * H + j], dhid[q], a);
    dj[j] += a;
  }
  if (j < 2 && learn_origin) dj[H + j] += G[j] + dxin[j];
  // parameter gradients of this sample
  for (int i = j; i < 3 * H * H; i += nt) atomicAdd(dw_hh + i, s_dwhh[i]);
  for (int i = j; i < 3 * H * I; i += nt) atomicAdd(dw_ih + i, s_dwih[i]);
  for (int i = j; i < 3 * H; i += nt) {
    atomicAdd(db_ih + i, s_dbih[i]);
    atomicAdd(db_hh + i, s_dbhh[i]);
  }
  for (int i = j; i < 2 * H; i += nt) atomicAdd(dw_out + i, s_dwout[i]);
  if (j < 2) atomicAdd(db_out + j, dbo[j]);
}

}  // namespace

extern "C" int tfpp_gru_cell_head(const float* joined, int joined_stride, const float* target_point, const float* w_ih,
                                  const float* w_hh, const float* b_ih, const float* b_hh, const float* w_out,
                                  const float* b_out, const float* w_ts0, const float* b_ts0, const float* w_ts1,
                                  const float* b_ts1, float* waypoints, float* speed_logits, float* h_all, int batch,
                                  int steps, int hidden, int input_size, int learn_origin, int n_speed,
                                  tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(hidden >= 1 && hidden <= kMaxH && 3 * hidden <= 384, "hidden size <= 128");
  TFPP_CHECK_ARG(input_size == 2 || input_size == 4, "input = previous waypoint (2) [+ target point (2)]");
  TFPP_CHECK_ARG(input_size == 2 || target_point != nullptr, "target point required");
  TFPP_CHECK_ARG(n_speed <= hidden, "n_speed <= hidden");
  TFPP_CHECK_ARG(joined_stride >= hidden + (learn_origin ? 2 : 0), "joined row too short");
  if (batch <= 0) return TFPP_OK;
  GruW w{w_ih, w_hh, b_ih, b_hh, w_out, b_out};
  gru_cell_head_kernel<<<batch, 384, 0, stream>>>(joined, joined_stride, target_point, w, w_ts0, b_ts0, w_ts1, b_ts1,
                                                  waypoints, speed_logits, h_all, steps, hidden, input_size, learn_origin,
                                                  n_speed);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_gru_cell_head_bwd(const float* joined, int joined_stride, const float* target_point, const float* w_ih,
                                      const float* w_hh, const float* b_ih, const float* b_hh, const float* w_out,
                                      const float* b_out, const float* w_ts0, const float* b_ts0, const float* w_ts1,
                                      const float* waypoints, const float* h_all, const float* d_waypoints,
                                      const float* d_speed_logits, float* d_joined, float* dw_ih, float* dw_hh,
                                      float* db_ih, float* db_hh, float* dw_out, float* db_out, float* dw_ts0,
                                      float* db_ts0, float* dw_ts1, float* db_ts1, int batch, int steps, int hidden,
                                      int input_size, int learn_origin, int n_speed, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(hidden >= 1 && hidden <= kMaxH && 3 * hidden <= 384, "hidden size <= 128");
  TFPP_CHECK_ARG(input_size == 2 || input_size == 4, "input = previous waypoint (2) [+ target point (2)]");
  TFPP_CHECK_ARG(h_all != nullptr && waypoints != nullptr && d_waypoints != nullptr && d_joined != nullptr, "null buffer");
  if (batch <= 0) return TFPP_OK;
  GruW w{w_ih, w_hh, b_ih, b_hh, w_out, b_out};
  const size_t smem = sizeof(float) * (3 * hidden * hidden + 3 * hidden * input_size + 6 * hidden + 2 * hidden);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gru_cell_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr = true;
  }
  gru_cell_head_bwd_kernel<<<batch, 384, smem, stream>>>(joined, joined_stride, target_point, w, w_ts0, b_ts0, w_ts1, waypoints,
                                                         h_all, d_waypoints, d_speed_logits, d_joined, dw_ih, dw_hh, db_ih,
                                                         db_hh, dw_out, db_out, dw_ts0, db_ts0, dw_ts1, db_ts1, steps, hidden,
                                                         input_size, learn_origin, n_speed);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
