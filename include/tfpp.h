/* C ABI of libtfpp.so — the B200 (sm_100a) kernels under the TransFuser++ nn.Module surface.
 *
 * The reference (autonomousvision/carla_garage) has no native code and no FFI: its hot path is a
 * chain of torch library dispatches inside team_code/{transfuser,model,center_net}.py.  The drop-in boundary a user
 * sees is therefore the Python class surface (carla_garage_b200.nn mirrors it); this header is the boundary UNDER
 * it: plain pointers + sizes + a CUDA stream, int return code (0 = ok, see tfpp_last_error()), no torch types, no
 * allocation, no global mutable state except the last-error string (thread-local).  Every entry point names the
 * reference call site(s) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Layout conventions: feature maps are NHWC bf16 (B,H,W,C contiguous); token matrices are (rows, C) bf16 or f32;
 * user-facing model inputs/outputs stay NCHW f32 like the reference's.
 */
#ifndef TFPP_H_
#define TFPP_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tfpp_stream_t; /* cudaStream_t */

const char* tfpp_last_error(void);
int tfpp_abi_version(void);

/* ---- K1: LiDAR points -> BEV histogram.  team_code/data.py:873-906 (np.histogramdd x2, clip 5, /5, .T) ----
 * points: (B, n_points, 3) f32 device; counts: (B, 2, 256, 256) u32 workspace (zeroed inside);
 * out: (B, C, H, W) f32, C = 2 if use_ground_plane else 1, index [b][ch][y_bin][x_bin]. */
int tfpp_pillar_scatter(const float* points, int batch, int n_points, uint32_t* counts, float* out,
                        int use_ground_plane, float min_x, float max_x, float min_y, float max_y,
                        float pixels_per_meter, int hist_max, float split_z, float max_z, tfpp_stream_t stream);

/* ---- tcgen05 implicit-GEMM convolution / linear layer --------------------------------------------------------
 * out[pixel, n] = act( scale[n] * sum_{tap,c} A[pixel + shift(tap), c0(n) + c] * Wt[n, tap, c] + shift[n]
 *                      + res1[pixel, n] + res2[pixel, n] )
 * Replaces: timm RegNet 1x1 / grouped 3x3 convs (transfuser.py:216-219), nn.Linear in SelfAttention/Block
 * (transfuser.py:352-359,391-396), 1x1 channel maps (transfuser.py:233,237), FPN + decoder + head convs
 * (transfuser.py:131-137, transfuser_utils.py:675-704, model.py:75-90,148, center_net.py:43-47), decoder
 * projections (model.py:137-143).  A is NHWC bf16 (a_batch, H, W, a_channels); weights are bf16
 * (N, taps, w_kdim) K-major.  Pixel tiles are th x tw x nb = 128 rows. */
typedef struct {
  const void* a;
  int a_batch, height, width, a_channels; /* A tensor extents (a_batch may be 4*batch for stride-2 parity planes) */
  long long a_batch_stride;               /* elements between A images; 0 = contiguous (height*width*a_channels) */
  const void* w;
  int w_taps, w_kdim, n; /* weights (n, w_taps, w_kdim) bf16 */
  int batch;             /* output batch */
  int k_per_tile;        /* input channels reduced per n-tile (dense: a_channels; grouped: <= 64) */
  int a_c_per_ntile;     /* A channel offset added per n-tile (0 dense, 48 grouped) */
  int bn;                /* n tile: multiple of 16, <= 256 */
  int tw, th, nb;        /* pixel tile */
  int ntaps;
  int tap_dx[9], tap_dy[9], tap_db[9], tap_w[9];
  void* out;
  int out_f32;
  long long o_sb, o_sy, o_sx, o_sn; /* element strides of out */
  const void* res1;
  int res1_f32;
  long long r1_sb, r1_sy, r1_sx, r1_sn;
  const void* res2;
  int res2_f32;
  long long r2_sb, r2_sy, r2_sx, r2_sn;
  const float* scale; /* per output channel, may be NULL (=1) */
  const float* shift; /* per output channel, may be NULL (=0) */
  int act;            /* 0 none, 1 relu, 2 sigmoid, 3 gelu(erf) */
  int act_n_limit;    /* activation applies to channels < act_n_limit (0 = all) */
  float* stat_sum;    /* optional per-channel sum / sum of squares of the raw accumulator (BatchNorm batch stats) */
  float* stat_sq;
  /* training-mode dropout on the output (nn.Dropout of transfuser.py:379,395 and of nn.TransformerDecoderLayer,
   * model.py:137-140), applied AFTER the activation and BEFORE the residuals are added:
   * out = drop(act(scale*acc+shift)) + res1 + res2.  drop_rng = device pointer to {seed, step} (2 x uint64) or NULL;
   * the word of output element e (offset in elements from `out`) comes from Philox4x32-10, see tfpp_dropout. */
  const unsigned long long* drop_rng;
  float drop_p;
  unsigned drop_site;
} tfpp_conv_gemm_args;

int tfpp_conv_gemm(const tfpp_conv_gemm_args* args, tfpp_stream_t stream);

/* ---- tcgen05 weight-gradient GEMM ------------------------------------------------------------------------------
 * dw[co, tap_w, ci] += sum_pixels dy[pixel, co] * x[pixel + shift(tap), ci]   (fp32 atomics: zero dw first)
 * dy: NHWC bf16 (batch, height, width, cout); x: NHWC bf16 (x_batch, height, width, x_channels) (x_batch = 4*batch for
 * stride-2 parity planes, taps as in tfpp_conv_gemm).  dw is addressed through element strides, so it can be the
 * parameter's own (cout, cin, kh, kw) gradient.  Grouped (group_width > 0): input channel ci of group g pairs with
 * output channels of group g only and the ci index of dw is local to the group.  Replaces autograd's convolution_backward (weight) / addmm of train.py:898. */
typedef struct {
  const void* dy;
  const void* x;
  float* dw;
  int batch, height, width, cout; /* cout = channels of dy (may be zero-padded to a multiple of 8) */
  int cout_valid;                 /* rows of dw actually written (0 = cout) */
  int x_batch, x_channels;
  long long x_batch_stride; /* elements, 0 = contiguous */
  int cin;                  /* dense: number of input channels written (<= x_channels) */
  int group_width;          /* 0 = dense */
  long long dw_s_co, dw_s_tap, dw_s_ci; /* element strides of dw: (co, tap_w, ci) -> co*s_co + tap*s_tap + ci*s_ci */
  int ntaps;
  int tap_dx[9], tap_dy[9], tap_db[9], tap_w[9];
  int tw, th, nb;           /* 64-pixel tile */
  int bn;                   /* ci tile (dense), 0 = auto */
  int splits;               /* pixel-range splits, 0 = auto */
} tfpp_wgrad_args;

int tfpp_conv_wgrad(const tfpp_wgrad_args* args, tfpp_stream_t stream);

/* ---- small-channel 3x3 convolutions at high resolution (PerspectiveDecoder tail, transfuser_utils.py:690-704) ------
 * x NHWC bf16 (B,H,W,cin), w bf16 (cout, 9, cin) (tap = ky*3+kx reads x[y+ky-1][x+kx-1]; rows >= n_valid zero),
 * out NHWC bf16 (B,H,W,cout) or NCHW f32 (B,n_valid,H,W).  (cin, cout) in {(32,32),(32,16),(32,8),(16,32)}.
 * The input-gradient is the same call with the transposed, spatially flipped weight pack. */
int tfpp_smallc_conv3x3(const void* x, const void* w, const float* bias, void* out, int out_nchw_f32, int n_valid,
                        int act, int act_n_limit, int batch, int height, int width, int cin, int cout,
                        tfpp_stream_t stream);
/* dw[co, tap, ci] += sum_pixels dy[pix, co] * x[pix + tap, ci]; dy (B,H,W,cout_padded), x (B,H,W,cin=32). */
int tfpp_smallc_wgrad3x3(const void* dy, const void* x, float* dw, long long s_co, long long s_tap, long long s_ci,
                         int co_valid, int batch, int height, int width, int cout_padded, int cin,
                         tfpp_stream_t stream);

/* ---- stem conv: timm RegNet stem ConvNormAct(in,32,k3,s2,p1) fused with normalize_imagenet ---------------------
 * (team_code/transfuser.py:146-149,164-167; transfuser_utils.py:542-551).  x: NCHW f32 (B,cin<=3,H,W), w: f32
 * (32,cin,3,3); in_scale/in_shift: per-input-channel affine applied before zero padding (NULL = identity);
 * scale/shift: per-output-channel affine (folded eval BatchNorm) + act; out: NHWC bf16 (B,H/2,W/2,32);
 * stat_sum/stat_sq (32 f32 each, optional): batch statistics of the raw conv output (training BatchNorm). */
int tfpp_stem_conv(const float* x, const float* w, const float* in_scale, const float* in_shift, const float* scale,
                   const float* shift, int act, void* out, float* stat_sum, float* stat_sq, int batch, int cin,
                   int height, int width, tfpp_stream_t stream);

/* ---- BatchNorm2d (training semantics; timm BatchNormAct2d) ----------------------------------------------------
 * sum / sq: per-channel sums from the conv epilogue; writes the affine (scale, shift) the apply pass uses, the saved
 * mean / invstd for backward, and updates running stats (momentum 0.1, unbiased variance). */
int tfpp_bn_finalize(const float* sum, const float* sq, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                     int channels, float count, float eps, float momentum, tfpp_stream_t stream);

/* y = act(x * scale[c] + shift[c] (+ res)), NHWC bf16; pool_sum (B,C) f32 optional: per-sample channel sums of y
 * (the squeeze of timm SEModule).  Any of scale/shift (together) and res may be NULL; res_scale/res_shift (together,
 * optional) apply a per-channel affine to res first (the BatchNorm of the RegNet downsample branch). */
int tfpp_scale_shift_act(const void* x, const void* res, const float* scale, const float* shift, const float* res_scale,
                         const float* res_shift, int act, void* y, float* pool_sum, int batch, int hw, int channels,
                         tfpp_stream_t stream);

/* timm SEModule excite: gate = sigmoid(fc2(relu(fc1(pool_sum / hw)))); w1 (rd,C), w2 (C,rd) f32.
 * hidden (B,rd) f32 is required: it carries relu(fc1) between the two passes and is what the backward reads. */
int tfpp_se_gate(const float* pool_sum, int hw, const float* w1, const float* b1, const float* w2, const float* b2,
                 float* gate, float* hidden, int batch, int channels, int rd, tfpp_stream_t stream);

/* y[b,p,c] = x[b,p,c] * gate[b,c] (SE scale). */
int tfpp_channel_scale(const void* x, const float* gate, void* y, int batch, int hw, int channels, tfpp_stream_t stream);

/* (B,H,W,C) -> (4B,H/2,W/2,C) parity planes (plane q=(y&1)*2+(x&1) at batch q*B+b): stride-2 convs become taps. */
int tfpp_parity_split(const void* x, void* y, int batch, int height, int width, int channels, tfpp_stream_t stream);

/* AdaptiveAvgPool2d -> token rows (+pos_emb): transfuser.py:230-231,317-325. out rows [row0, row0+ph*pw) of
 * (B, rows_per_batch, C), f32 or bf16. */
int tfpp_avgpool_tokens(const void* x, const float* pos_emb, void* out, int out_f32, int batch, int height, int width,
                        int channels, int ph, int pw, int rows_per_batch, int row0, tfpp_stream_t stream);

/* F.interpolate(bilinear, align_corners=False) of a (B,sh,sw,C) source (f32 or bf16, arbitrary batch/row element
 * strides) to NHWC bf16 (B,dh,dw,C), optionally + add (transfuser.py:239-255,119-123; transfuser_utils.py:699-701). */
int tfpp_bilinear(const void* src, int src_f32, long long src_batch_stride, long long src_row_stride, const void* add,
                  void* out, int batch, int sh, int sw, int dh, int dw, int channels, tfpp_stream_t stream);

/* bev_semantic_decoder tail: resize NHWC bf16 (B,sh,sw,src_channels) -> NCHW f32 (B,channels,dh,dw) * mask(dh,dw)
 * (model.py:88-90,385). */
int tfpp_bilinear_nchw_mask(const void* src, const float* mask, float* out, int batch, int sh, int sw, int src_channels,
                            int channels, int dh, int dw, tfpp_stream_t stream);

int tfpp_nchw_f32_to_nhwc_bf16(const float* x, void* y, int batch, int channels, int hw, tfpp_stream_t stream);
int tfpp_nhwc_bf16_to_nchw_f32(const void* x, float* y, int batch, int channels, int hw, tfpp_stream_t stream);

/* ---- token-side kernels ---------------------------------------------------------------------------------------
 * nn.LayerNorm (transfuser.py:288,388-389; model.py:123,137-143): x f32 or bf16 (rows,C) -> bf16 and/or f32. */
int tfpp_layernorm(const void* x, int x_f32, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                   float* save_mean, float* save_rstd, int rows, int channels, float eps, tfpp_stream_t stream);

/* SelfAttention core (transfuser.py:367-376): qkv (B,T,3C) bf16 [q|k|v] -> out (B,T,C) bf16. T<=320, T%32==0. */
int tfpp_fusion_attn(const void* qkv, void* out, int batch, int tokens, int channels, int heads, tfpp_stream_t stream);

/* nn.MultiheadAttention core of the planner decoder (model.py:137-143): bf16 row-strided q/k/v views. */
int tfpp_small_mha(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb, long long k_sr,
                   const void* v, long long v_sb, long long v_sr, void* out, long long o_sb, long long o_sr, int batch,
                   int heads, int tq, int tk, int head_dim, tfpp_stream_t stream);

/* extra-sensor memory token (model.py:308-319): BatchNorm1d(1,affine=False)(ego_vel) ++ command -> MLP -> +pos. */
int tfpp_extra_sensor_token(const float* ego_vel, const float* command, float vel_mean, float vel_var,
                            int use_batch_stats, float* running_mean, float* running_var, const float* w0,
                            const float* b0, const float* w1, const float* b1, const float* pos, void* mem_bf16,
                            float* mem_f32, int batch, int n_cmd, int hidden, int d_model, int rows_per_batch, int row,
                            tfpp_stream_t stream);

/* GRUWaypointsPredictorInterFuser + target_speed_network (model.py:118-119,357-358,857-867). joined (B,n_wp+1,D) f32. */
int tfpp_planner_head(const float* joined, const float* target_point, const float* w_enc, const float* b_enc,
                      const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* w_dec,
                      const float* b_dec, const float* w_ts0, const float* b_ts0, const float* w_ts1,
                      const float* b_ts1, float* checkpoints, float* speed_logits, float* h_all, int batch, int n_wp,
                      int d_model, int hidden, int n_speed, tfpp_stream_t stream);

/* LidarCenterNetHead.decode_heatmap (center_net.py:172-237): NCHW f32 maps (batch strides in elements) ->
 * (B,k,9) [x,y,w,h,yaw,vel,brake,cls,score]. */
int tfpp_decode_heatmap(const float* heat, long long heat_sb, const float* wh, long long wh_sb, const float* offset,
                        long long off_sb, const float* yaw_cls, long long ycls_sb, const float* yaw_res,
                        long long yres_sb, float* out, int batch, int n_cls, int height, int width, int n_bins, int k,
                        float width_ratio, float height_ratio, tfpp_stream_t stream);

/* ==== backward / training kernels (adjoints of the ops above; they replace the autograd graph that
 * team_code/train.py:898 `loss.backward()` walks, and optim.AdamW(amsgrad=True).step() of train.py:527-531,908) ==== */

/* BatchNorm(+ReLU)(+SE gate / squeeze) backward. dy,y,raw NHWC bf16; gate,pool_grad (B,C) f32 optional;
 * s1,s2 (C) f32 zeroed by the caller: on return s1 = dbeta, s2 = dgamma. draw = grad wrt the conv output; dz_out
 * (optional) = masked incoming gradient (grad of the residual input of the block).  With ReLU and y == NULL the mask
 * is recomputed as raw * fwd_scale + fwd_shift > 0 (the forward's folded BatchNorm affine): one tensor less to read. */
int tfpp_bn_bwd(const void* dy, const void* y, const void* raw, const float* mean, const float* invstd,
                const float* gamma, const float* gate, const float* pool_grad, const float* fwd_scale,
                const float* fwd_shift, int act, float* s1, float* s2, void* draw, void* dz_out, int batch, int hw,
                int channels, tfpp_stream_t stream);

/* SE backward: dout = grad wrt (a2 * gate); outputs pool_grad (B,C) = dL/d(pool_sum) and fc1/fc2 gradients (+=).
 * dgate_sum (B,C) f32 zeroed by the caller and ws (B,C+rd) f32 are workspaces. */
int tfpp_se_bwd(const void* dout, const void* a2, const float* gate, const float* hidden, const float* pool_sum, int hw,
                const float* w1, const float* w2, float* dgate_sum, float* ws, float* dw1, float* db1, float* dw2,
                float* db2, float* pool_grad, int batch, int channels, int rd, tfpp_stream_t stream);

/* dz = dy_scale * dy * act'(y) -> NHWC bf16 (channels_padded), dbias += column sums.
 * layout: 0 = dy,y NHWC bf16; 1 = NCHW f32; 2 = NHWC f32 (token matrices). dz may be NULL (bias gradient only). */
int tfpp_act_bwd(const void* dy, const void* y, int layout, int act, int act_n_limit, float dy_scale, void* dz,
                 float* dbias, int batch, int hw, int channels, int channels_padded, tfpp_stream_t stream);

/* adjoints of tfpp_bilinear / tfpp_bilinear_nchw_mask / tfpp_avgpool_tokens(+residual) / tfpp_parity_split */
int tfpp_bilinear_bwd(const void* dout, void* dsrc, int dsrc_f32, long long src_batch_stride, long long src_row_stride,
                      int accumulate, int batch, int sh, int sw, int dh, int dw, int channels, tfpp_stream_t stream);
int tfpp_bilinear_nchw_mask_bwd(const float* dout, const float* mask, void* dsrc, int batch, int sh, int sw,
                                int src_channels, int channels, int dh, int dw, tfpp_stream_t stream);
int tfpp_pool_bwd_add(const void* dout, const void* dtok, int dtok_f32, void* out, int batch, int height, int width,
                      int channels, int ph, int pw, int rows_per_batch, int row0, tfpp_stream_t stream);
int tfpp_parity_merge(const void* x, void* y, int batch, int height, int width, int channels, tfpp_stream_t stream);
int tfpp_add_bf16(const void* a, const void* b, void* y, long long n, tfpp_stream_t stream);
int tfpp_cast_f32_bf16(const float* x, void* y, long long n, tfpp_stream_t stream);
/* out (groups*rows, C) bf16 = rows [row0, row0+rows) of every group of x (groups, group_rows, C) f32; dbias += sums */
int tfpp_cast_rows(const float* x, void* out, float* dbias, int groups, int group_rows, int row0, int rows,
                   int channels, tfpp_stream_t stream);
/* out (n) += sum_b x (batch, n) f32 */
int tfpp_batch_reduce(const float* x, float* out, int batch, long long n, tfpp_stream_t stream);

/* stem conv weight gradient: dw (32,cin,3,3) f32 += ; draw NHWC bf16 (B,H/2,W/2,32). */
int tfpp_stem_wgrad(const float* x, const void* draw, const float* in_scale, const float* in_shift, float* dw, int batch,
                    int cin, int height, int width, tfpp_stream_t stream);

int tfpp_layernorm_bwd(const void* dy, int dy_f32, const float* x, const float* mean, const float* rstd,
                       const float* gamma, const float* dres, float* dx, float* dgamma, float* dbeta, int rows,
                       int channels, tfpp_stream_t stream);

/* fusion attention backward: dqkv (B,T,3C) bf16; dkv_ws (B,T,2C) f32 workspace (zeroed inside). */
int tfpp_fusion_attn_bwd(const void* qkv, const void* dout, void* dqkv, float* dkv_ws, int batch, int tokens,
                         int channels, int heads, tfpp_stream_t stream);

int tfpp_small_mha_bwd(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb, long long k_sr,
                       const void* v, long long v_sb, long long v_sr, const void* dout, long long o_sb, long long o_sr,
                       void* dq, long long dq_sb, long long dq_sr, void* dk, long long dk_sb, long long dk_sr, void* dv,
                       long long dv_sb, long long dv_sr, int accumulate_kv, int batch, int heads, int tq, int tk,
                       int head_dim, tfpp_stream_t stream);

int tfpp_extra_sensor_token_bwd(const float* ego_vel, const float* command, float vel_mean, float vel_var,
                                int use_batch_stats, const float* w0, const float* b0, const float* w1, const float* b1,
                                const float* dmem, long long dmem_stride, float* dw0, float* db0, float* dw1, float* db1,
                                float* dpos, int batch, int n_cmd, int hidden, int d_model, tfpp_stream_t stream);

int tfpp_planner_head_bwd(const float* joined, const float* target_point, const float* h_all, const float* w_enc,
                          const float* b_enc, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                          const float* w_dec, const float* w_ts0, const float* b_ts0, const float* w_ts1,
                          const float* dcp, const float* dlogits, float* djoined, float* dw_enc, float* db_enc,
                          float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, float* dw_dec, float* db_dec,
                          float* dw_ts0, float* db_ts0, float* dw_ts1, float* db_ts1, int batch, int n_wp, int d_model,
                          int hidden, int n_speed, tfpp_stream_t stream);

/* ---- dropout (training mode; embd_pdrop / attn_pdrop / resid_pdrop = 0.1, config.py:364-366; decoder 0.1) -----------
 * Counter-based: element i of dropout site `site` is dropped iff word (i % 4) of
 * Philox4x32-10(counter = (i / 4 lo, i / 4 hi, site, step), key = seed) < p * 2^32, kept elements are scaled by
 * 1 / (1 - p).  rng = device pointer to {seed, step} (uint64 each; the caller bumps step once per training forward,
 * also inside a CUDA graph).  No mask is stored: backward kernels regenerate it.  rng == NULL or p == 0: identity.
 * In place over x (n elements, f32 or bf16): GPT.drop on pos_emb + tokens (transfuser.py:325) and its adjoint. */
int tfpp_dropout(void* x, int x_f32, long long n, const unsigned long long* rng, float p, unsigned site,
                 tfpp_stream_t stream);
/* tfpp_act_bwd with the mask of an output dropout multiplied in: dz = drop_mask(e) * dy_scale * dy * act'(y) where e is
 * the element offset pixel * channels + c (layouts 0 and 2). */
int tfpp_act_bwd_dropout(const void* dy, const void* y, int layout, int act, int act_n_limit, float dy_scale, void* dz,
                         float* dbias, int batch, int hw, int channels, int channels_padded,
                         const unsigned long long* rng, float p, unsigned site, tfpp_stream_t stream);
/* attention cores with dropout on the softmax probabilities (transfuser.py:374; nn.MultiheadAttention(dropout=0.1));
 * element index of P[b, h, q, k] = ((b * heads + h) * Tq + q) * Tk + k. */
int tfpp_fusion_attn_dropout(const void* qkv, void* out, int batch, int tokens, int channels, int heads,
                             const unsigned long long* rng, float p, unsigned site, tfpp_stream_t stream);
int tfpp_fusion_attn_bwd_dropout(const void* qkv, const void* dout, void* dqkv, float* dkv_ws, int batch, int tokens,
                                 int channels, int heads, const unsigned long long* rng, float p, unsigned site,
                                 tfpp_stream_t stream);
int tfpp_small_mha_dropout(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb, long long k_sr,
                           const void* v, long long v_sb, long long v_sr, void* out, long long o_sb, long long o_sr,
                           int batch, int heads, int tq, int tk, int head_dim, const unsigned long long* rng, float p,
                           unsigned site, tfpp_stream_t stream);
int tfpp_small_mha_bwd_dropout(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb,
                               long long k_sr, const void* v, long long v_sb, long long v_sr, const void* dout,
                               long long o_sb, long long o_sr, void* dq, long long dq_sb, long long dq_sr, void* dk,
                               long long dk_sb, long long dk_sr, void* dv, long long dv_sb, long long dv_sr,
                               int accumulate_kv, int batch, int heads, int tq, int tk, int head_dim,
                               const unsigned long long* rng, float p, unsigned site, tfpp_stream_t stream);

/* fused losses (model.py:394-445, center_net.py:77-123): scalar loss sums + d(weighted loss)/d(pre-activation).
 * The gradient outputs (dz*, dbias, dlogits/dcp excepted) may be NULL: loss values only.  w_dev / w2_dev (nullable):
 * loss weight(s) read on the device and multiplied into the gradients — the autograd boundary passes d(total)/d(loss)
 * this way without a host round trip (train.py:889-896 forms the total from the dict compute_loss returns). */
int tfpp_ce_map_loss(const float* logits, const long long* labels, const float* valid, float grad_scale,
                     const float* w_dev, float* loss_sum, void* dz_nhwc, float* dz_nchw, float* dbias, int batch,
                     int classes, int channels_padded, int hw, tfpp_stream_t stream);
int tfpp_l1_sigmoid_loss(const float* p, const float* target, float grad_scale, const float* w_dev, float* loss_sum,
                         void* dz_nhwc, float* dbias, int channels_padded, long long n, tfpp_stream_t stream);
int tfpp_center_head_loss(const float* maps, const float* t_heat, const float* t_wh, const float* t_off,
                          const long long* t_ycls, const float* t_yres, const float* pix_w, const float* avg_factor,
                          const float* w5, float* losses, void* dz, float* dbias, int batch, int hw, int n_cls,
                          int n_bins, int channels_padded, tfpp_stream_t stream);
int tfpp_planner_loss(const float* logits, const long long* labels, const float* class_w, const float* cp,
                      const float* cp_t, float w_ts, float w_cp, const float* w2_dev, float* losses, float* dlogits,
                      float* dcp, int batch, int n_cls, int n_cp, tfpp_stream_t stream);

/* Grouped 3x3 convolution of the RegNetY bottleneck (timm regnet.Bottleneck.conv2, group width 24; stride 1 or 2,
 * padding 1) on a haloed shared-memory tile: x (B,H,W,C) NHWC bf16, C % 72 == 0; w (C/24, 9, 24, 24) bf16 =
 * [group][ky*3+kx][out][in]; out (B,H/stride,W/stride,C) bf16.  Optional: per-channel affine + ReLU (eval-mode
 * BatchNorm fold) and BatchNorm batch statistics of the raw output (stat_sum/stat_sq, C floats each, +=).  The input
 * gradient of the stride-1 conv is the same call on dY with the transposed, spatially flipped pack.
 * Replaces nn.Conv2d(groups=C/24) + BatchNorm2d statistics in timm's regnet.Bottleneck (oracle/regnety.py). */
int tfpp_gconv3x3(const void* x, const void* w, void* out, const float* scale, const float* shift, int act,
                  float* stat_sum, float* stat_sq, int batch, int height, int width, int channels, int stride,
                  tfpp_stream_t stream);

/* Input gradient of the stride-2 tfpp_gconv3x3: dy (B,Ho,Wo,C) bf16, w_t = the transposed / flipped pack (the same one
 * the stride-1 input gradient uses), dx (B,2Ho,2Wo,C) bf16 (every element written). */
int tfpp_gconv3x3_dgrad_s2(const void* dy, const void* w_t, void* dx, int batch, int out_height, int out_width,
                           int channels, tfpp_stream_t stream);

/* Weight gradient of tfpp_gconv3x3: dw (C,24,3,3) f32 torch layout += sum over pixels of dY x X inside each group.
 * workspace: tfpp_gconv3x3_wgrad_workspace(...) floats of scratch (per-CTA partial sums, reduced without atomics). */
long long tfpp_gconv3x3_wgrad_workspace(int batch, int height, int width, int channels, int stride);
int tfpp_gconv3x3_wgrad(const void* dy, const void* x, float* dw, float* workspace, int batch, int height, int width,
                        int channels, int stride, tfpp_stream_t stream);

/* EXPERIMENTAL (op-level parity green on B200, not on the default path in round 1, see csrc/halo_umma.cu): dense 3x3 conv, stride 1, pad 1, on tcgen05
 * from a haloed shared-memory tile held in 8-channel planes (no im2col, the input is read once).  x (B,H,W,cin) NHWC
 * bf16, cin in {16,32,64}; w (9, cin/8, cout_padded, 8) bf16 = [tap][k chunk][n][8 k] (ops.pack_halo_umma_weight);
 * out NHWC bf16 (cout_padded channels) or NCHW f32 (n_valid channels).  Same role as tfpp_smallc_conv3x3. */
int tfpp_halo_conv3x3(const void* x, const void* w, const float* bias, void* out, int out_nchw_f32, int n_valid, int act,
                      int act_n_limit, int batch, int height, int width, int cin, int cout_padded, tfpp_stream_t stream);

/* EXPERIMENTAL, not yet run on a GPU (csrc/halo_umma.cu): the stride-1 RegNet group conv on tcgen05 from the haloed
 * plane layout.  Same contract as tfpp_gconv3x3(stride 1) except the weight pack: (C/24, 9, 4, 32, 8) bf16 =
 * [group][tap][k chunk][n][8 k], zero padded from 24 to 32 (ops.pack_halo_gconv_weight). */
int tfpp_halo_gconv3x3(const void* x, const void* w, void* out, const float* scale, const float* shift, int act,
                       float* stat_sum, float* stat_sq, int batch, int height, int width, int channels,
                       tfpp_stream_t stream);

/* Weight-pack refresh: out[i] = idx[i] >= 0 ? flat[idx[i]] : 0 for i < n (n % 8 == 0), cast to bf16 (out_f32 = 0) or
 * kept fp32.  One launch rebuilds every kernel-layout weight copy after the optimizer step; replaces the implicit
 * per-module weight reads of torch's conv / linear kernels (team_code/train.py:898-908 loop). */
int tfpp_gather_pack(const float* flat, const int* idx, void* out, long long n, int out_f32, tfpp_stream_t stream);

/* AdamW(amsgrad=True), torch semantics; grad_scale multiplies the gradient first (1/world_size after a sum
 * all-reduce).  dev_state (optional, 4 floats on the device: [step count, learning rate, 1 - beta1^step,
 * sqrt(1 - beta2^step)]; the caller sets the first two) replaces the host-side `step` / `lr` so the launch can be
 * replayed from a CUDA graph; the count is incremented and the bias corrections refreshed by the call.
 * param / grad / state buffers 16-byte aligned.  flags (optional, one byte per element): bit 0 = no weight decay
 * (the no_decay group of create_optimizer_groups, model.py:556-645), bit 1 = frozen (requires_grad=False,
 * train.py:495-508: element untouched). */
int tfpp_adamw_amsgrad(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                       long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                       float grad_scale, float* dev_state, const unsigned char* flags, tfpp_stream_t stream);

/* ---- fp32 parity mode (csrc/fp32_path.cu) ------------------------------------------------------------------------
 * BASELINE.json north_star: "within 1e-3 rel fp32 / 1e-2 bf16".  The same op contracts as their bf16 twins above with
 * fp32 feature maps / weight packs and fp32 CUDA-core contractions (train.py:374-377 is the reference's effective
 * precision: fp32, TF32 off).  carla_garage_b200.ops dispatches to them when the activation dtype is float32
 * (ops.set_precision('fp32')); they serve the end-to-end parity tests, never the benchmarked path. */
int tfpp_conv_gemm_f32(const tfpp_conv_gemm_args* args, tfpp_stream_t stream); /* a, w: f32; bn/tw/th/nb ignored */
int tfpp_gconv3x3_f32(const float* x, const float* w, float* out, const float* scale, const float* shift, int act,
                      float* stat_sum, float* stat_sq, int batch, int height, int width, int channels, int stride,
                      tfpp_stream_t stream);
int tfpp_stem_conv_f32(const float* x, const float* w, const float* in_scale, const float* in_shift, const float* scale,
                       const float* shift, int act, float* out, float* stat_sum, float* stat_sq, int batch, int cin,
                       int height, int width, tfpp_stream_t stream);
int tfpp_scale_shift_act_f32(const float* x, const float* res, const float* scale, const float* shift,
                             const float* res_scale, const float* res_shift, int act, float* y, float* pool_sum,
                             int batch, int hw, int channels, tfpp_stream_t stream);
int tfpp_channel_scale_f32(const float* x, const float* gate, float* y, int batch, int hw, int channels,
                           tfpp_stream_t stream);
int tfpp_parity_split_f32(const float* x, float* y, int batch, int height, int width, int channels, tfpp_stream_t stream);
int tfpp_avgpool_tokens_f32(const float* x, const float* pos_emb, float* out, int batch, int height, int width,
                            int channels, int ph, int pw, int rows_per_batch, int row0, tfpp_stream_t stream);
int tfpp_bilinear_f32(const float* src, long long src_batch_stride, long long src_row_stride, const float* add,
                      float* out, int batch, int sh, int sw, int dh, int dw, int channels, tfpp_stream_t stream);
int tfpp_bilinear_nchw_mask_f32(const float* src, const float* mask, float* out, int batch, int sh, int sw,
                                int src_channels, int channels, int dh, int dw, tfpp_stream_t stream);
/* attention core of both transformers (transfuser.py:367-376, model.py:137-143) on row-strided f32 q/k/v views;
 * probability dropout indexed like tfpp_fusion_attn_dropout / tfpp_small_mha_dropout */
int tfpp_mha_f32(const float* q, long long q_sb, long long q_sr, const float* k, long long k_sb, long long k_sr,
                 const float* v, long long v_sb, long long v_sr, float* out, long long o_sb, long long o_sr, int batch,
                 int heads, int tq, int tk, int head_dim, const unsigned long long* drop_rng, float drop_p,
                 unsigned drop_site, tfpp_stream_t stream);

/* ---- data-parallel gradient exchange over NVLink peer memory, fused with AdamW (csrc/peer_exchange.cu) -----------
 * Replaces DistributedDataParallel's NCCL all-reduce (train.py:516) + ZeroRedundancyOptimizer(AdamW) (train.py:527-531):
 * one kernel per step reduce-scatters the flat gradient out of the peers' buffers, steps AdamW(amsgrad) on the shard
 * this rank owns (gradient averaged over the ranks) and pushes the new parameters into every rank's flat parameter
 * buffer, between two flag barriers.  Buffers come from tfpp_peer_alloc (cudaMalloc + CUDA IPC handle, zero filled) and
 * are mapped into the other processes with tfpp_peer_open.  Plain kernels only: CUDA-graph capturable. */
#define TFPP_MAX_PEERS 8
int tfpp_peer_alloc(long long bytes, void** ptr, void* handle64);
int tfpp_peer_open(const void* handle64, void** ptr);
int tfpp_peer_close(void* ptr);
int tfpp_peer_free(void* ptr);
typedef struct {
  int world, rank;
  float* grad[TFPP_MAX_PEERS];      /* every rank's flat gradient (entry `rank` = the local buffer) */
  float* param[TFPP_MAX_PEERS];     /* every rank's flat parameter buffer */
  unsigned* flags[TFPP_MAX_PEERS];  /* every rank's barrier words: (2 * TFPP_MAX_PEERS + 2) uint32, zero initialised */
  float* exp_avg;                   /* local AdamW state, full length n (only the owned shard is touched) */
  float* exp_avg_sq;
  float* max_exp_avg_sq;
  long long n;                      /* elements of the flat buffers, multiple of 4 */
  float beta1, beta2, eps, weight_decay;
  float* dev_state;                 /* [step, lr, 1-beta1^step, sqrt(1-beta2^step)] as in tfpp_adamw_amsgrad */
  const unsigned char* opt_flags;   /* per-element flags as in tfpp_adamw_amsgrad, or NULL */
} tfpp_peer_step_args;
int tfpp_peer_adamw_step(const tfpp_peer_step_args* args, tfpp_stream_t stream);
int tfpp_peer_barrier(const tfpp_peer_step_args* args, int slot, tfpp_stream_t stream); /* flag barrier only (tests) */

/* ---- ensemble bounding-box merge (csrc/nms.cu) --------------------------------------------------------------------
 * sensor_agent.py:445-491: per frame, the union of the ensemble members' decoded boxes (center_net.py:172-237 rows
 * (x, y, half w, half h, yaw, ..., score), `stride` floats each, score last) is thresholded (score > conf_threshold,
 * model.py:449), optionally converted image -> vehicle frame (transfuser_utils.py:388-406) and merged by rotated-IoU
 * non-maximum suppression (transfuser_utils.py:409-452; shapely polygon IoU restated as convex clipping).
 * boxes / out_boxes: (batch, num_boxes <= 512, stride) f32; out_boxes holds the kept boxes, highest score first, rows
 * >= out_count[b] zero; out_index (optional, (batch, num_boxes) int32): source row of every kept box, -1 after. */
int tfpp_nms_rotated(const float* boxes, int batch, int num_boxes, int stride, float conf_threshold, float iou_threshold,
                     int to_vehicle, float pixels_per_meter, float min_x, float min_y, float* out_boxes, int* out_count,
                     int* out_index, tfpp_stream_t stream);

/* ---- CenterNet training targets on the device (csrc/targets.cu) ---------------------------------------------------
 * CARLA_Data.get_targets (data.py:698-791) + gaussian_target.py:11-61,160-183 + angle2class (center_net.py:240-254):
 * boxes (batch, max_boxes <= 128, 8) f32 = (x, y, extent_x, extent_y, yaw, speed, brake, class) in BEV pixels of the
 * img_h x img_w LiDAR image, counts (batch) int32 valid boxes per sample (NULL = all).  Outputs in the layout the train
 * loop moves to the device (train.py:693-766): center_heatmap (B,num_classes,H,W), wh / offset / pixel_weight (B,2,H,W),
 * yaw_class (B,H,W) int64, yaw_res (B,1,H,W), velocity (B,1,H,W) or NULL, brake (B,H,W) int64 or NULL, avg_factor (B). */
int tfpp_centernet_targets(const float* boxes, const int* counts, int batch, int max_boxes, int feat_h, int feat_w,
                           int img_h, int img_w, int num_classes, int num_dir_bins, float* center_heatmap, float* wh,
                           float* offset, long long* yaw_class, float* yaw_res, float* velocity, long long* brake,
                           float* pixel_weight, float* avg_factor, tfpp_stream_t stream);

/* fp32 parity mode, backward half (csrc/fp32_path_bwd.cu): adjoints with fp32 gradient storage, same contracts as
 * tfpp_conv_wgrad (dense), tfpp_gconv3x3_wgrad / _dgrad_s2, tfpp_stem_wgrad, tfpp_bn_bwd, the reduce pass of
 * tfpp_se_bwd (then call tfpp_se_bwd with dout = NULL), tfpp_act_bwd_dropout (nchw: dy / y NCHW f32, else NHWC f32),
 * tfpp_bilinear_bwd, tfpp_bilinear_nchw_mask_bwd, tfpp_pool_bwd_add, tfpp_add_bf16, tfpp_cast_rows, and the attention
 * adjoint of both transformers (workspace: 2 * batch * heads * tq * tk floats). */
int tfpp_conv_wgrad_f32(const tfpp_wgrad_args* args, tfpp_stream_t stream);
int tfpp_gconv3x3_wgrad_f32(const float* dy, const float* x, float* dw, int batch, int height, int width, int channels,
                            int stride, tfpp_stream_t stream);
int tfpp_gconv3x3_dgrad_s2_f32(const float* dy, const float* w_t, float* dx, int batch, int out_height, int out_width,
                               int channels, tfpp_stream_t stream);
int tfpp_stem_wgrad_f32(const float* x, const float* draw, const float* in_scale, const float* in_shift, float* dw,
                        int batch, int cin, int height, int width, tfpp_stream_t stream);
int tfpp_bn_bwd_f32(const float* dy, const float* y, const float* raw, const float* mean, const float* invstd,
                    const float* gamma, const float* gate, const float* pool_grad, const float* fwd_scale,
                    const float* fwd_shift, int act, float* s1, float* s2, float* draw, float* dz_out, int batch, int hw,
                    int channels, tfpp_stream_t stream);
int tfpp_se_bwd_reduce_f32(const float* dout, const float* a2, float* dgate_sum, int batch, int hw, int channels,
                           tfpp_stream_t stream);
int tfpp_act_bwd_f32(const float* dy, const float* y, int nchw, int act, int act_n_limit, float dy_scale, float* dz,
                     float* dbias, int batch, int hw, int channels, int channels_padded,
                     const unsigned long long* drop_rng, float drop_p, unsigned drop_site, tfpp_stream_t stream);
int tfpp_bilinear_bwd_f32(const float* dout, float* dsrc, long long src_batch_stride, long long src_row_stride,
                          int accumulate, int batch, int sh, int sw, int dh, int dw, int channels, tfpp_stream_t stream);
int tfpp_bilinear_nchw_mask_bwd_f32(const float* dout, const float* mask, float* dsrc, int batch, int sh, int sw,
                                    int src_channels, int channels, int dh, int dw, tfpp_stream_t stream);
int tfpp_pool_bwd_add_f32(const float* dout, const float* dtok, float* out, int batch, int height, int width,
                          int channels, int ph, int pw, int rows_per_batch, int row0, tfpp_stream_t stream);
int tfpp_add_f32(const float* a, const float* b, float* y, long long n, tfpp_stream_t stream);
int tfpp_copy_rows_f32(const float* x, float* out, float* dbias, int groups, int group_rows, int row0, int rows,
                       int channels, tfpp_stream_t stream);
int tfpp_mha_bwd_f32(const float* q, long long q_sb, long long q_sr, const float* k, long long k_sb, long long k_sr,
                     const float* v, long long v_sb, long long v_sr, const float* dout, long long o_sb, long long o_sr,
                     float* dq, long long dq_sb, long long dq_sr, float* dk, long long dk_sb, long long dk_sr, float* dv,
                     long long dv_sb, long long dv_sr, float* workspace, int accumulate_kv, int batch, int heads, int tq,
                     int tk, int head_dim, const unsigned long long* drop_rng, float drop_p, unsigned drop_site,
                     tfpp_stream_t stream);

/* ---- the original TransFuser planner head (config.transformer_decoder_join = False; csrc/gru_cell.cu) ---------------
 * GRUWaypointsPredictorTransFuser (model.py:870-913): hidden state h_0 = joined[:, :hidden], first input
 * x_0 = joined[:, hidden:hidden+2] (learn_origin) or 0; per step h = GRUCell([x, target_point], h), x += Linear(h);
 * waypoints (batch, steps, 2) = the x after every step.  Optionally the target-speed MLP (Linear-ReLU-Linear,
 * model.py:113-118,371-376) on joined[:, :hidden] -> speed_logits (batch, n_speed).  joined rows are joined_stride
 * floats apart.  h_all (batch, steps + 1, hidden) keeps the states for the backward pass (NULL in inference).
 * Backward: BPTT; d_joined (batch, joined_stride) and every parameter gradient are ACCUMULATED (+=). */
int tfpp_gru_cell_head(const float* joined, int joined_stride, const float* target_point, const float* w_ih,
                       const float* w_hh, const float* b_ih, const float* b_hh, const float* w_out, const float* b_out,
                       const float* w_ts0, const float* b_ts0, const float* w_ts1, const float* b_ts1, float* waypoints,
                       float* speed_logits, float* h_all, int batch, int steps, int hidden, int input_size,
                       int learn_origin, int n_speed, tfpp_stream_t stream);
int tfpp_gru_cell_head_bwd(const float* joined, int joined_stride, const float* target_point, const float* w_ih,
                           const float* w_hh, const float* b_ih, const float* b_hh, const float* w_out,
                           const float* b_out, const float* w_ts0, const float* b_ts0, const float* w_ts1,
                           const float* waypoints, const float* h_all, const float* d_waypoints,
                           const float* d_speed_logits, float* d_joined, float* dw_ih, float* dw_hh, float* db_ih,
                           float* db_hh, float* dw_out, float* db_out, float* dw_ts0, float* db_ts0, float* dw_ts1,
                           float* db_ts1, int batch, int steps, int hidden, int input_size, int learn_origin,
                           int n_speed, tfpp_stream_t stream);

/* K1 with CARLA_Data.align / the agent's half-sweep merge fused in (data.py:840-871, sensor_agent.py:381-425,
 * transfuser_utils.py:116-130): every point of sample b goes through n_xforms rigid transforms
 * p' = R(yaw)^T (p - t), xform = (batch, n_xforms, 4) doubles {tx, ty, tz, yaw}, in float64 like numpy, before the
 * float64 histogram of data.py:873-906 (split_z is the float64 lidar_split_height). */
int tfpp_pillar_scatter_aligned(const float* points, const double* xform, int n_xforms, int batch, int n_points,
                                unsigned int* counts, float* out, int use_ground_plane, float min_x, float max_x,
                                float min_y, float max_y, float pixels_per_meter, int hist_max, double split_z,
                                float max_z, tfpp_stream_t stream);

/* ---- bev_encoder backbone (reference team_code/bev_encoder.py) --------------------------------
 * nn.InstanceNorm2d(affine=False, eps) + activation (bev_encoder.py:126-137 bev_compressor, :253-262 UpsamplingConcat):
 * tfpp_instnorm_stats accumulates per-(sample, channel) sum / sum of squares into zero-initialised (batch, channels)
 * buffers; tfpp_instnorm_apply writes y = act((x - mean) * invstd) (biased variance) and, if given, mean / invstd for
 * the backward pass; tfpp_instnorm_bwd is the adjoint (s1, s2: zero-initialised (batch, channels) workspaces).
 * x / dx are NHWC with unit channel stride; *_pix_stride = elements between consecutive pixels (>= channels), so that
 * y / dy may live inside a wider tensor.  f32 = 1: float32 tensors (parity mode), 0: bf16. */
int tfpp_instnorm_stats(const void* x, int f32, long long x_pix_stride, int batch, int hw, int channels, float* sum,
                        float* sq, tfpp_stream_t stream);
int tfpp_instnorm_apply(const void* x, int f32, long long x_pix_stride, const float* sum, const float* sq, float eps,
                        int act, void* y, long long y_pix_stride, float* mean, float* invstd, int batch, int hw,
                        int channels, tfpp_stream_t stream);
int tfpp_instnorm_bwd(const void* dy, long long dy_pix_stride, const void* x, int f32, const float* mean,
                      const float* invstd, int act, float* s1, float* s2, void* dx, int batch, int hw, int channels,
                      tfpp_stream_t stream);

/* Camera -> BEV lift, bev_encoder.py:179-199 (F.grid_sample over the (depth, width, height) voxel grid of
 * transfuser_utils.py:596-665, sum over height, / bev_projection_normalizer, transpose, * valid_bev_pixels) in its
 * separable form: out[b, w, d, :] = wl[d, w] * V[b, d, x0[d, w], :] + wr[d, w] * V[b, d, x0[d, w] + 1, :],
 * V[b, d, x, :] = sum_y a_rows[d, y] * img[b, y, x, :].  img (batch, img_h, img_w, channels) NHWC, out (batch, width,
 * depth, channels) NHWC; a_rows (depth, img_h) f32, x0 (depth, width) int32 in [0, img_w - 2], wl / wr (depth, width)
 * f32 with the normaliser, the visibility mask and the zero padding folded in.  tfpp_bev_lift_bwd: adjoint wrt img;
 * ws = (batch, depth, img_w, channels) f32 workspace; accumulate = 1 adds to dimg. */
int tfpp_bev_lift(const void* img, int f32, const float* a_rows, const int* x0, const float* wl, const float* wr,
                  void* out, int batch, int img_h, int img_w, int channels, int depth, int width, tfpp_stream_t stream);
int tfpp_bev_lift_bwd(const void* dout, int f32, const float* a_rows, const int* x0, const float* wl, const float* wr,
                      float* ws, void* dimg, int accumulate, int batch, int img_h, int img_w, int channels, int depth,
                      int width, tfpp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TFPP_H_ */
