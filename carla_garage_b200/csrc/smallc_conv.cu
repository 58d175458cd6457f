// 3x3 convolutions with few channels (Cin, Cout <= 32) at high resolution: the full-resolution tail of the
// PerspectiveDecoder (transfuser_utils.py:690-704: 32->32 and 32->7 / 32->1 convs at 256x1024) and their gradients.
//
// These layers are HBM-bound (AI = 9*Cout flop per input byte ~ 288 flop/B is under the tcgen05 ridge only because the
// implicit-GEMM kernel re-reads every input tile 9 times through L2 in half-empty 128-byte TMA rows).  Here one CTA
// stages a haloed 10x34 pixel tile in shared memory ONCE (cp.async, zero fill = conv padding), keeps the 9 tap weights
// resident, and the 9 shifted products run on mma.sync m16n8k16 straight out of that tile: HBM traffic = input once
// (x1.33 halo) + output once.  Forward and input-gradient use the same kernel (dgrad = conv with the transposed,
// spatially flipped pack); the weight gradient has its own kernel below (ldmatrix.trans operands, pixel contraction).
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

constexpr int TH = 8, TW = 32;                 // output tile: 8 rows x 32 columns, one row per warp
constexpr int HH = TH + 2, HW_ = TW + 2;       // halo tile

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct SmallConvParams {
  const bf16* x;        // (B,H,W,CIN) NHWC
  const bf16* w;        // (COUT, 9, CIN) bf16, tap = ky*3+kx, reads x[y+ky-1][x+kx-1]
  const float* bias;    // optional (n_valid)
  void* out;
  int out_nchw_f32;     // 0: NHWC bf16 with COUT channels; 1: NCHW f32 with n_valid channels
  int n_valid;
  int act, act_n_limit;
  int B, H, W;
  int tiles_x, tiles_y;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256, 2) smallc_conv3x3_kernel(const SmallConvParams p) {
  constexpr int P = CIN + 8;                 // smem pixel pitch (bf16 elements): bank-conflict-free fragment loads
  constexpr int NT = COUT / 8;               // n tiles of 8 output channels
  constexpr int KS = CIN / 16;               // k steps per tap
  constexpr int CH16 = CIN / 8;              // 16-byte chunks per pixel
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* wsm = reinterpret_cast<bf16*>(smem_raw);                 // [COUT][9][P]
  bf16* halo0 = wsm + COUT * 9 * P;                              // [2][HH][HW_][P]
  constexpr int HALO_ELEMS = HH * HW_ * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;

  // weights -> smem (once per CTA)
  for (int i = threadIdx.x; i < COUT * 9 * CH16; i += blockDim.x) {
    const int row = i / CH16, ch = i % CH16;
    *reinterpret_cast<uint4*>(wsm + row * P + ch * 8) = *reinterpret_cast<const uint4*>(p.w + row * CIN + ch * 8);
  }
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = tiles_per_img * p.B;

  auto load_tile = [&](int tile, bf16* dst) {
    const int b = tile / tiles_per_img;
    const int r = tile % tiles_per_img;
    const int y0 = (r / p.tiles_x) * TH - 1, x0 = (r % p.tiles_x) * TW - 1;
    for (int i = threadIdx.x; i < HH * HW_ * CH16; i += blockDim.x) {
      const int ch = i % CH16, pix = i / CH16;
      const int hy = pix / HW_, hx = pix % HW_;
      const int yy = y0 + hy, xx = x0 + hx;
      const bool ok = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      const bf16* src = p.x + ((static_cast<long long>(b) * p.H + (ok ? yy : 0)) * p.W + (ok ? xx : 0)) * CIN + ch * 8;
      cp_async16(dst + pix * P + ch * 8, src, ok);
    }
    cp_async_commit();
  };

  int buf = 0;
  int tile = blockIdx.x;
  if (tile < num_tiles) load_tile(tile, halo0);
  for (; tile < num_tiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < num_tiles) load_tile(next, halo0 + (buf ^ 1) * HALO_ELEMS);
    if (next < num_tiles) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const bf16* halo = halo0 + buf * HALO_ELEMS;
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
    const uint32_t halo_u = static_cast<uint32_t>(__cvta_generic_to_shared(halo));
    const uint32_t wsm_u = static_cast<uint32_t>(__cvta_generic_to_shared(wsm));
    const int lmat = lane >> 3, lrow = lane & 7;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        // A: 16 pixels x 16 channels per m tile, one ldmatrix.x4 each (lanes 0-15: pixel rows, lanes 16-31: +8 channels)
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
          ldsm_x4(a[mt], halo_u + (((warp + ky) * HW_ + mt * 16 + (lane & 15) + kx) * P + ks * 16 + (lane >> 4) * 8) * 2);
        // B: weight rows (output channels) x 16 input channels: one ldmatrix.x4 feeds two n tiles
        if (NT >= 2) {
#pragma unroll
          for (int np = 0; np < NT / 2; ++np) {
            uint32_t bq[4];
            ldsm_x4(bq, wsm_u + ((((np * 2 + (lmat >> 1)) * 8 + lrow) * 9 + tap) * P + ks * 16 + (lmat & 1) * 8) * 2);
            mma16816(acc[0][2 * np], a[0], bq);
            mma16816(acc[1][2 * np], a[1], bq);
            mma16816(acc[0][2 * np + 1], a[0], bq + 2);
            mma16816(acc[1][2 * np + 1], a[1], bq + 2);
          }
        } else {
          uint32_t bq[2];
          ldsm_x2(bq, wsm_u + ((lrow * 9 + tap) * P + ks * 16 + (lmat & 1) * 8) * 2);
          mma16816(acc[0][0], a[0], bq);
          mma16816(acc[1][0], a[1], bq);
        }
      }
    }
    // epilogue
    const int b = tile / tiles_per_img;
    const int r = tile % tiles_per_img;
    const int oy = (r / p.tiles_x) * TH + warp, ox0 = (r % p.tiles_x) * TW;
    if (oy < p.H) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int ox = ox0 + mt * 16 + g + hf * 8;
          if (ox >= p.W) continue;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int n = nt * 8 + 2 * t4;
            float v0 = acc[mt][nt][hf * 2], v1 = acc[mt][nt][hf * 2 + 1];
            if (p.bias) {
              if (n < p.n_valid) v0 += __ldg(p.bias + n);
              if (n + 1 < p.n_valid) v1 += __ldg(p.bias + n + 1);
            }
            if (p.act != ACT_NONE) {
              if (p.act_n_limit == 0 || n < p.act_n_limit) v0 = apply_act(v0, p.act);
              if (p.act_n_limit == 0 || n + 1 < p.act_n_limit) v1 = apply_act(v1, p.act);
            }
            if (p.out_nchw_f32) {
              float* o = static_cast<float*>(p.out);
              const long long hw = static_cast<long long>(p.H) * p.W;
              const long long base = (static_cast<long long>(b) * p.n_valid) * hw + static_cast<long long>(oy) * p.W + ox;
              if (n < p.n_valid) o[base + n * hw] = v0;
              if (n + 1 < p.n_valid) o[base + (n + 1) * hw] = v1;
            } else {
              bf16* o = static_cast<bf16*>(p.out) + ((static_cast<long long>(b) * p.H + oy) * p.W + ox) * COUT + n;
              *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v0, v1);
            }
          }
        }
      }
    }
    __syncthreads();  // everyone is done with `buf` before the next iteration's prefetch overwrites it
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: dW[co][tap][ci] += sum_pixels dY[pix][co] * X[pix + tap][ci]  (fp32 atomics at the end)
// CTA tile = 8 x 32 pixels: dY tile [256][CO] and haloed X tile [340][CI] in smem; contraction over the 256 pixels with
// mma.sync, operands fetched with ldmatrix.trans (pixel-major storage -> channel-major fragments).  The 9 x (CI/8)
// (tap, n-tile) output blocks are distributed round-robin over the 8 warps and accumulate across all tiles of the CTA.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, const void* smem_row_ptr) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t* r, const void* smem_row_ptr) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row_ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(s));
}

struct SmallWgradParams {
  const bf16* dy;   // (B,H,W,CO)
  const bf16* x;    // (B,H,W,CI)
  float* dw;        // element (co, tap, ci) at co*s_co + tap*s_tap + ci*s_ci
  long long s_co, s_tap, s_ci;
  int co_valid;
  int B, H, W, tiles_x, tiles_y;
};

template <int CO, int CI>
__global__ void __launch_bounds__(256, 2) smallc_wgrad3x3_kernel(const SmallWgradParams p) {
  constexpr int PY = CO + 8, PX = CI + 8;
  constexpr int MT = CO / 16;                 // m tiles (output channels)
  constexpr int NTI = CI / 8;                 // n tiles (input channels)
  constexpr int UNITS = 9 * NTI;              // (tap, n tile) blocks
  constexpr int UPW = (UNITS + 7) / 8;        // blocks per warp
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* dys = reinterpret_cast<bf16*>(smem_raw);          // [TH*TW][PY]
  bf16* xs = dys + TH * TW * PY;                          // [HH*HW_][PX]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = tiles_per_img * p.B;
  float acc[UPW][MT][4];
#pragma unroll
  for (int u = 0; u < UPW; ++u)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[u][m][0] = acc[u][m][1] = acc[u][m][2] = acc[u][m][3] = 0.f;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int r = tile % tiles_per_img;
    const int ty0 = (r / p.tiles_x) * TH, tx0 = (r % p.tiles_x) * TW;
    __syncthreads();
    for (int i = threadIdx.x; i < TH * TW * (CO / 8); i += blockDim.x) {
      const int ch = i % (CO / 8), pix = i / (CO / 8);
      const int yy = ty0 + pix / TW, xx = tx0 + pix % TW;
      const bool ok = yy < p.H && xx < p.W;
      const bf16* src = p.dy + ((static_cast<long long>(b) * p.H + (ok ? yy : 0)) * p.W + (ok ? xx : 0)) * CO + ch * 8;
      cp_async16(dys + pix * PY + ch * 8, src, ok);
    }
    for (int i = threadIdx.x; i < HH * HW_ * (CI / 8); i += blockDim.x) {
      const int ch = i % (CI / 8), pix = i / (CI / 8);
      const int yy = ty0 - 1 + pix / HW_, xx = tx0 - 1 + pix % HW_;
      const bool ok = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      const bf16* src = p.x + ((static_cast<long long>(b) * p.H + (ok ? yy : 0)) * p.W + (ok ? xx : 0)) * CI + ch * 8;
      cp_async16(xs + pix * PX + ch * 8, src, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // contraction over the tile's pixels: k step = 16 consecutive pixels of one tile row
#pragma unroll 1
    for (int ks = 0; ks < TH * TW / 16; ++ks) {
      const int ry = ks / (TW / 16), rx = (ks % (TW / 16)) * 16;   // first pixel of the k step inside the tile
      // A fragments (m = co, k = pixel) for every m tile: matrices [k0-7][m0-7], [k0-7][m8-15], [k8-15][m0-7], [k8-15][m8-15]
      uint32_t a[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int mat = lane >> 3, rr = lane & 7;
        const int pix = ry * TW + rx + (mat >> 1) * 8 + rr;
        const int col = m * 16 + (mat & 1) * 8;
        uint32_t t[4];
        ldmatrix_x4_trans(t, dys + pix * PY + col);
        a[m][0] = t[0]; a[m][1] = t[1]; a[m][2] = t[2]; a[m][3] = t[3];
      }
#pragma unroll
      for (int u = 0; u < UPW; ++u) {
        const int unit = warp + u * 8;
        if (unit < UNITS) {
          const int tap = unit / NTI, nt = unit % NTI;
          const int ky = tap / 3, kx = tap % 3;
          // B fragment (k = pixel shifted by the tap, n = ci): matrices [k0-7][n0-7], [k8-15][n0-7]
          const int mat = (lane >> 3) & 1, rr = lane & 7;
          const int hpix = (ry + ky) * HW_ + rx + kx + mat * 8 + rr;
          uint32_t bfr[2];
          ldmatrix_x2_trans(bfr, xs + hpix * PX + nt * 8);
#pragma unroll
          for (int m = 0; m < MT; ++m) mma16816(acc[u][m], a[m], bfr);
        }
      }
    }
  }
  // flush: acc[u][m] holds rows co = m*16 + g (+8), cols ci = nt*8 + 2*t4 (+1)
#pragma unroll
  for (int u = 0; u < UPW; ++u) {
    const int unit = warp + u * 8;
    if (unit < UNITS) {
      const int tap = unit / NTI, nt = unit % NTI;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = m * 16 + g + (e >> 1) * 8;
          const int ci = nt * 8 + 2 * t4 + (e & 1);
          if (co < p.co_valid) atomicAdd(p.dw + co * p.s_co + tap * p.s_tap + ci * p.s_ci, acc[u][m][e]);
        }
      }
    }
  }
}

template <int CIN, int COUT>
int launch_conv(const SmallConvParams& p, cudaStream_t stream) {
  constexpr int P = CIN + 8;
  const size_t smem = sizeof(bf16) * (COUT * 9 * P + 2 * HH * HW_ * P);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(smallc_conv3x3_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr = true;
  }
  const int tiles = p.tiles_x * p.tiles_y * p.B;
  const int grid = tiles < 2 * TFPP_NUM_SMS ? tiles : 2 * TFPP_NUM_SMS;
  smallc_conv3x3_kernel<CIN, COUT><<<grid, 256, smem, stream>>>(p);
  return 0;
}

template <int CO, int CI>
int launch_wgrad(const SmallWgradParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(bf16) * (TH * TW * (CO + 8) + HH * HW_ * (CI + 8));
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(smallc_wgrad3x3_kernel<CO, CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr = true;
  }
  const int tiles = p.tiles_x * p.tiles_y * p.B;
  const int grid = tiles < 2 * TFPP_NUM_SMS ? tiles : 2 * TFPP_NUM_SMS;
  smallc_wgrad3x3_kernel<CO, CI><<<grid, 256, smem, stream>>>(p);
  return 0;
}

}  // namespace

extern "C" int tfpp_smallc_conv3x3(const void* x, const void* w, const float* bias, void* out, int out_nchw_f32,
                                   int n_valid, int act, int act_n_limit, int batch, int height, int width, int cin,
                                   int cout, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SmallConvParams p;
  p.x = static_cast<const bf16*>(x); p.w = static_cast<const bf16*>(w); p.bias = bias; p.out = out;
  p.out_nchw_f32 = out_nchw_f32; p.n_valid = n_valid; p.act = act; p.act_n_limit = act_n_limit;
  p.B = batch; p.H = height; p.W = width;
  p.tiles_x = ceil_div(width, TW); p.tiles_y = ceil_div(height, TH);
  TFPP_CHECK_ARG(n_valid <= cout, "n_valid <= cout");
  if (cin == 32 && cout == 32) launch_conv<32, 32>(p, stream);
  else if (cin == 32 && cout == 8) launch_conv<32, 8>(p, stream);
  else if (cin == 32 && cout == 16) launch_conv<32, 16>(p, stream);
  else if (cin == 16 && cout == 32) launch_conv<16, 32>(p, stream);
  else {
    tfpp_set_error("tfpp_smallc_conv3x3: unsupported (cin, cout) = (%d, %d)", cin, cout);
    return TFPP_ERR_ARG;
  }
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_smallc_wgrad3x3(const void* dy, const void* x, float* dw, long long s_co, long long s_tap,
                                    long long s_ci, int co_valid, int batch, int height, int width, int cout_padded,
                                    int cin, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SmallWgradParams p;
  p.dy = static_cast<const bf16*>(dy); p.x = static_cast<const bf16*>(x); p.dw = dw;
  p.s_co = s_co; p.s_tap = s_tap; p.s_ci = s_ci; p.co_valid = co_valid;
  p.B = batch; p.H = height; p.W = width;
  p.tiles_x = ceil_div(width, TW); p.tiles_y = ceil_div(height, TH);
  if (cout_padded == 32 && cin == 32) launch_wgrad<32, 32>(p, stream);
  else if (cout_padded == 16 && cin == 32) launch_wgrad<16, 32>(p, stream);
  else {
    tfpp_set_error("tfpp_smallc_wgrad3x3: unsupported (cout_padded, cin) = (%d, %d)", cout_padded, cin);
    return TFPP_ERR_ARG;
  }
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
