// Token-side kernels of TransFuser++: LayerNorm, the fusion self-attention core (320 tokens, 4 heads, head dims
// 18/54/144/378), the small decoder attention (11 queries x 65 memory tokens), the GRU path decoder and the small
// MLP heads.  Reference call sites are cited per entry point.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ LayerNorm
// nn.LayerNorm(C), eps 1e-5, biased variance (transfuser.py:288,388-389; model.py:123,137-143 norm1..3).
// One warp per row, two-pass statistics in fp32.  Writes bf16 (GEMM operand) and/or f32 (residual stream).
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x, int x_f32,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        bf16* __restrict__ y_bf16, float* __restrict__ y_f32,
                                                        float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                        int rows, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xf = static_cast<const float*>(x) + static_cast<long long>(row) * C;
  const bf16* xb = static_cast<const bf16*>(x) + static_cast<long long>(row) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += x_f32 ? xf[c] : bf2f(xb[c]);
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = (x_f32 ? xf[c] : bf2f(xb[c])) - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    if (save_mean) save_mean[row] = mean;
    if (save_rstd) save_rstd[row] = rstd;
  }
  for (int c = lane; c < C; c += 32) {
    const float v = ((x_f32 ? xf[c] : bf2f(xb[c])) - mean) * rstd * gamma[c] + beta[c];
    if (y_bf16) y_bf16[static_cast<long long>(row) * C + c] = f2bf(v);
    if (y_f32) y_f32[static_cast<long long>(row) * C + c] = v;
  }
}

// ------------------------------------------------------------------------------------------------ decoder attention
// nn.MultiheadAttention core inside nn.TransformerDecoderLayer (model.py:137-143): Tq <= 16 queries, Tk <= 512 keys,
// head_dim 32.  One CTA per (batch, head), one warp per query row.  q/k/v are row-strided f32 or bf16 views.
__global__ void __launch_bounds__(256) small_mha_kernel(const bf16* __restrict__ q, long long q_sb, long long q_sr,
                                                        const bf16* __restrict__ k, long long k_sb, long long k_sr,
                                                        const bf16* __restrict__ v, long long v_sb, long long v_sr,
                                                        bf16* __restrict__ out, long long o_sb, long long o_sr, int Tq,
                                                        int Tk, int hd, float scale, const unsigned long long* drop_rng,
                                                        float drop_p, unsigned drop_site) {
  const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);   // nn.MultiheadAttention(dropout=0.1) on the probabilities
  extern __shared__ float mha_sm[];
  float* ks = mha_sm;                    // [Tk][hd+1]
  float* vs = ks + Tk * (hd + 1);        // [Tk][hd+1]
  float* ps = vs + Tk * (hd + 1);        // [warps][Tk]
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < Tk * hd; i += blockDim.x) {
    const int r = i / hd, d = i % hd;
    ks[r * (hd + 1) + d] = bf2f(k[b * k_sb + r * k_sr + h * hd + d]);
    vs[r * (hd + 1) + d] = bf2f(v[b * v_sb + r * v_sr + h * hd + d]);
  }
  __syncthreads();
  float* pw = ps + warp * Tk;
  for (int r = warp; r < Tq; r += nwarps) {
    const bf16* qp = q + b * q_sb + r * q_sr + h * hd;
    float m = -INFINITY;
    for (int c = lane; c < Tk; c += 32) {
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(bf2f(qp[d]), ks[c * (hd + 1) + d], a);
      a *= scale;
      pw[c] = a;
      m = fmaxf(m, a);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < Tk; c += 32) {
      const float e = __expf(pw[c] - m);
      pw[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    if (drop.on) {
      const unsigned long long base = ((static_cast<unsigned long long>(b) * gridDim.y + h) * Tq + r) * Tk;
      for (int c = lane; c < Tk; c += 32) pw[c] *= drop_mult(drop, base + c);
      __syncwarp();
    }
    for (int d = lane; d < hd; d += 32) {
      float a = 0.f;
      for (int c = 0; c < Tk; ++c) a = fmaf(pw[c], vs[c * (hd + 1) + d], a);
      out[b * o_sb + r * o_sr + h * hd + d] = f2bf(a * inv);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ small dense heads
__device__ __forceinline__ float warp_dot(const float* __restrict__ w, const float* __restrict__ x, int n, int lane) {
  float a = 0.f;
  for (int i = lane; i < n; i += 32) a = fmaf(__ldg(w + i), x[i], a);
  return warp_sum(a);
}

// extra-sensor token, model.py:308-319: BatchNorm1d(1, affine=False)(ego_vel) ++ command -> Linear(7,128) ReLU
// -> Linear(128,256) ReLU -> + extra_sensor_pos_embed, written as memory row `row` of (B, rows_per_batch, 256).
__global__ void __launch_bounds__(256) extra_sensor_kernel(const float* __restrict__ ego_vel,
                                                           const float* __restrict__ command, float vel_mean,
                                                           float vel_invstd, int use_batch_stats, float eps,
                                                           float* __restrict__ running_mean,
                                                           float* __restrict__ running_var, float momentum,
                                                           const float* __restrict__ w0, const float* __restrict__ b0,
                                                           const float* __restrict__ w1, const float* __restrict__ b1,
                                                           const float* __restrict__ pos, bf16* __restrict__ mem_bf16,
                                                           float* __restrict__ mem_f32, int B, int n_cmd, int hidden,
                                                           int d_model, int rows_per_batch, int row) {
  extern __shared__ float es_sm[];
  float* in = es_sm;            // 1 + n_cmd
  float* hid = es_sm + 8;       // hidden
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (warp == 0) {
    float mean = vel_mean, invstd = vel_invstd;
    if (use_batch_stats) {  // training-mode BatchNorm1d: biased batch variance
      float s = 0.f;
      for (int i = lane; i < B; i += 32) s += ego_vel[i];
      mean = warp_sum(s) / B;
      float q = 0.f;
      for (int i = lane; i < B; i += 32) {
        const float d = ego_vel[i] - mean;
        q = fmaf(d, d, q);
      }
      const float var = warp_sum(q) / B;
      invstd = rsqrtf(var + eps);
      if (b == 0 && lane == 0 && running_mean != nullptr) {
        const float unbiased = B > 1 ? var * B / (B - 1.f) : var;
        running_mean[0] = (1.f - momentum) * running_mean[0] + momentum * mean;
        running_var[0] = (1.f - momentum) * running_var[0] + momentum * unbiased;
      }
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
  for (int j = warp; j < d_model; j += nwarps) {
    float a = warp_dot(w1 + static_cast<long long>(j) * hidden, hid, hidden, lane);
    if (lane == 0) {
      a = fmaxf(a + b1[j], 0.f) + pos[j];
      const long long o = (static_cast<long long>(b) * rows_per_batch + row) * d_model + j;
      if (mem_bf16) mem_bf16[o] = f2bf(a);
      if (mem_f32) mem_f32[o] = a;
    }
  }
}

// GRUWaypointsPredictorInterFuser.forward (model.py:857-867) + target_speed_network (model.py:118-119,358).
// joined: (B, n_wp + 1, D) f32 decoder output.  One CTA per sample; the 64-wide hidden state lives in shared memory,
// every gate pre-activation is one warp-level dot product (coalesced weight rows + shuffle reduction); gate order
// (r, z, n) as nn.GRU.  Outputs: checkpoints (B, n_wp, 2) = cumsum(decoder(h_t)), speed logits (B, n_speed).
__global__ void __launch_bounds__(256) planner_head_kernel(
    const float* __restrict__ joined, const float* __restrict__ target_point, const float* __restrict__ w_enc,
    const float* __restrict__ b_enc, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
    const float* __restrict__ b_ih, const float* __restrict__ b_hh, const float* __restrict__ w_dec,
    const float* __restrict__ b_dec, const float* __restrict__ w_ts0, const float* __restrict__ b_ts0,
    const float* __restrict__ w_ts1, const float* __restrict__ b_ts1, float* __restrict__ checkpoints,
    float* __restrict__ speed_logits, float* __restrict__ h_all, int n_wp, int D, int HS, int n_speed) {
  extern __shared__ float gr_sm[];
  const int n_rows = n_wp + (n_speed > 0 ? 1 : 0);   // n_speed == 0: no target-speed token (wp_decoder, model.py:165-171)
  float* x = gr_sm;                 // (n_wp + 1) * D
  float* h = x + (n_wp + 1) * D;    // HS
  float* gi = h + HS;               // 3 HS
  float* gh = gi + 3 * HS;          // 3 HS
  float* hid = gh + 3 * HS;         // D (target-speed hidden)
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < n_rows * D; i += blockDim.x) x[i] = joined[static_cast<long long>(b) * n_rows * D + i];
  if (threadIdx.x < HS) {
    const float tx = target_point[b * 2], ty = target_point[b * 2 + 1];
    h[threadIdx.x] = w_enc[threadIdx.x * 2] * tx + w_enc[threadIdx.x * 2 + 1] * ty + b_enc[threadIdx.x];
  }
  __syncthreads();
  float cx = 0.f, cy = 0.f;
  for (int t = 0; t < n_wp; ++t) {
    for (int j = warp; j < 3 * HS; j += nwarps) {
      const float a = warp_dot(w_ih + static_cast<long long>(j) * D, x + t * D, D, lane);
      const float c = warp_dot(w_hh + static_cast<long long>(j) * HS, h, HS, lane);
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
      const float hn = (1.f - z) * n + z * h[j];
      h[j] = hn;
      if (h_all) h_all[(static_cast<long long>(b) * n_wp + t) * HS + j] = hn;
    }
    __syncthreads();
    if (warp < 2) {
      const float d = warp_dot(w_dec + warp * HS, h, HS, lane) + b_dec[warp];
      if (lane == 0) {
        if (warp == 0) {
          cx += d;
          checkpoints[(static_cast<long long>(b) * n_wp + t) * 2] = cx;
        } else {
          cy += d;
          checkpoints[(static_cast<long long>(b) * n_wp + t) * 2 + 1] = cy;
        }
      }
    }
    __syncthreads();
  }
  // target speed MLP on the last query token
  if (n_speed == 0) return;
  const float* ts = x + n_wp * D;
  for (int j = warp; j < D; j += nwarps) {
    const float a = warp_dot(w_ts0 + static_cast<long long>(j) * D, ts, D, lane);
    if (lane == 0) hid[j] = fmaxf(a + b_ts0[j], 0.f);
  }
  __syncthreads();
  for (int j = warp; j < n_speed; j += nwarps) {
    const float a = warp_dot(w_ts1 + static_cast<long long>(j) * D, hid, D, lane);
    if (lane == 0) speed_logits[b * n_speed + j] = a + b_ts1[j];
  }
}

}  // namespace

#define STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int tfpp_layernorm(const void* x, int x_f32, const float* gamma, const float* beta, void* y_bf16,
                              float* y_f32, float* save_mean, float* save_rstd, int rows, int channels, float eps,
                              tfpp_stream_t stream_) {
  STREAM;
  layernorm_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(x, x_f32, gamma, beta, static_cast<bf16*>(y_bf16), y_f32,
                                                          save_mean, save_rstd, rows, channels, eps);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_small_mha_dropout(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb,
                                      long long k_sr, const void* v, long long v_sb, long long v_sr, void* out,
                                      long long o_sb, long long o_sr, int batch, int heads, int tq, int tk, int head_dim,
                                      const unsigned long long* drop_rng, float drop_p, unsigned drop_site,
                                      tfpp_stream_t stream_);

extern "C" int tfpp_small_mha(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb,
                              long long k_sr, const void* v, long long v_sb, long long v_sr, void* out, long long o_sb,
                              long long o_sr, int batch, int heads, int tq, int tk, int head_dim,
                              tfpp_stream_t stream_) {
  return tfpp_small_mha_dropout(q, q_sb, q_sr, k, k_sb, k_sr, v, v_sb, v_sr, out, o_sb, o_sr, batch, heads, tq, tk,
                                head_dim, nullptr, 0.f, 0u, stream_);
}

extern "C" int tfpp_small_mha_dropout(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb,
                                      long long k_sr, const void* v, long long v_sb, long long v_sr, void* out,
                                      long long o_sb, long long o_sr, int batch, int heads, int tq, int tk, int head_dim,
                                      const unsigned long long* drop_rng, float drop_p, unsigned drop_site,
                                      tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(tk <= 512 && head_dim <= 64, "small_mha: tk <= 512, head_dim <= 64");
  const size_t smem = sizeof(float) * (2 * tk * (head_dim + 1) + 8 * tk);
  static bool attr_set = false;
  if (!attr_set) {  // 257 memory tokens (bev_encoder backbone) need 76 KB
    cudaFuncSetAttribute(small_mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  dim3 grid(batch, heads);
  small_mha_kernel<<<grid, 256, smem, stream>>>(static_cast<const bf16*>(q), q_sb, q_sr, static_cast<const bf16*>(k),
                                                k_sb, k_sr, static_cast<const bf16*>(v), v_sb, v_sr,
                                                static_cast<bf16*>(out), o_sb, o_sr, tq, tk, head_dim,
                                                1.0f / sqrtf(static_cast<float>(head_dim)),
                                                drop_p > 0.f ? drop_rng : nullptr, drop_p, drop_site);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_extra_sensor_token(const float* ego_vel, const float* command, float vel_mean, float vel_var,
                                       int use_batch_stats, float* running_mean, float* running_var, const float* w0,
                                       const float* b0, const float* w1, const float* b1, const float* pos,
                                       void* mem_bf16, float* mem_f32, int batch, int n_cmd, int hidden, int d_model,
                                       int rows_per_batch, int row, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(n_cmd <= 7 && hidden <= 1024, "extra_sensor: n_cmd <= 7, hidden <= 1024");
  const float eps = 1e-5f;
  extra_sensor_kernel<<<batch, 256, sizeof(float) * (8 + hidden), stream>>>(
      ego_vel, command, vel_mean, rsqrtf(vel_var + eps), use_batch_stats, eps, running_mean, running_var, 0.1f, w0, b0,
      w1, b1, pos, static_cast<bf16*>(mem_bf16), mem_f32, batch, n_cmd, hidden, d_model, rows_per_batch, row);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_planner_head(const float* joined, const float* target_point, const float* w_enc, const float* b_enc,
                                 const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                 const float* w_dec, const float* b_dec, const float* w_ts0, const float* b_ts0,
                                 const float* w_ts1, const float* b_ts1, float* checkpoints, float* speed_logits,
                                 float* h_all, int batch, int n_wp, int d_model, int hidden, int n_speed,
                                 tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(hidden <= 256, "hidden <= 256");
  const size_t smem = sizeof(float) * ((n_wp + 1) * d_model + 7 * hidden + d_model);
  planner_head_kernel<<<batch, 256, smem, stream>>>(joined, target_point, w_enc, b_enc, w_ih, w_hh, b_ih, b_hh, w_dec,
                                                    b_dec, w_ts0, b_ts0, w_ts1, b_ts1, checkpoints, speed_logits, h_all,
                                                    n_wp, d_model, hidden, n_speed);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
