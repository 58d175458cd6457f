// Grouped 3x3 convolution of the RegNetY bottleneck (timm regnet.Bottleneck.conv2: group width 24, stride 1 or 2;
// oracle/regnety.py) and its input gradient.
//
// Per output pixel a group is a 216-long dot product for each of 24 channels: 18 flop per byte moved — an HBM-bound
// layer.  The implicit-GEMM tcgen05 path re-reads every input tile nine times through L2 and pads K from 24 to 64, so
// it ran ~10x above the traffic floor.  Here a CTA stages one haloed pixel tile of a 72-channel slab (3 groups) in
// shared memory ONCE (cp.async, zero fill = padding), each warp owns one group with that group's 9 x 24 x 24 weights
// resident in registers as mma fragments, and the nine shifted products run on mma.sync (m16n8k16 + m16n8k8 = K 24)
// straight out of the tile via ldmatrix.  HBM traffic = input once (x ~1.3 halo, mostly L2 hits) + output once.
// BatchNorm batch statistics (training) or the folded BatchNorm affine + ReLU (eval) ride in the epilogue.
// The input gradient of a stride-1 conv is the same kernel with the transposed, spatially flipped weight pack.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

constexpr int GW = 24;                  // group width (RegNetY-3.2GF: 24 for every stage)
constexpr int SLAB_C = 72;              // channels per CTA work item = 3 groups
constexpr int PIX_BYTES = SLAB_C * 2;   // 144 B per pixel in the tile: ldmatrix rows fall on distinct bank quads
constexpr int HALO_MAX = 340;           // pixels: 10 x 34, 18 x 18, 34 x 10 (stride 1) / 9 x 33, 17 x 17, 5 x 65 (stride 2)
constexpr int kWarps = 6;               // 2 warps per group
constexpr int kThreadsG = kWarps * 32;
constexpr int STAGE_BYTES = 16 * GW * 2;  // per-warp output staging: 16 pixels x 24 channels bf16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float* c, const uint32_t* a, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(b0));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gmem), "r"(sz) : "memory");
}

struct GConvParams {
  const bf16* x;       // (B,H,W,C) NHWC
  const bf16* w;       // (C/24, 9, 24, 24): [group][tap = ky*3+kx][out channel][in channel]
  bf16* out;           // (B,Ho,Wo,C)
  const float* scale;  // optional per-channel affine (eval-mode BatchNorm fold)
  const float* shift;
  int act;
  float* stat_sum;     // optional BatchNorm batch statistics of the raw output (C each)
  float* stat_sq;
  int B, H, W, C, Ho, Wo;
  int tw_log2, th;     // output tile: th rows x (1 << tw_log2) columns
  int tiles_x, tiles_y, slabs;
};

template <int STRIDE>
__global__ void __launch_bounds__(kThreadsG, 2) gconv3x3_kernel(const GConvParams p) {
  constexpr int TPIX = STRIDE == 1 ? 256 : 64;   // output pixels per tile
  constexpr int MT = TPIX / 16;                  // m tiles of 16 pixels per group
  constexpr int MT_PER_WARP = MT / 2;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* halo_base = smem_raw;                                   // [2][HALO_MAX][144 B]
  uint8_t* stage_base = smem_raw + 2 * HALO_MAX * PIX_BYTES;       // [kWarps][768 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t4 = lane & 3;
  const int group = warp >> 1, half = warp & 1;
  const int tw = 1 << p.tw_log2, th = p.th;
  const int hw = (tw - 1) * STRIDE + 3, hh = (th - 1) * STRIDE + 3;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int tiles_total = tiles_per_img * p.B;
  const int n_items = tiles_total * p.slabs;

  auto load_tile = [&](int item, int buf) {
    const int slab = item / tiles_total;
    const int tile = item - slab * tiles_total;
    const int b = tile / tiles_per_img;
    const int r = tile - b * tiles_per_img;
    const int iy0 = (r / p.tiles_x) * th * STRIDE - 1, ix0 = (r % p.tiles_x) * tw * STRIDE - 1;
    const uint32_t dst0 = smem_u32(halo_base + buf * (HALO_MAX * PIX_BYTES));
    const bf16* src0 = p.x + static_cast<long long>(b) * p.H * p.W * p.C + slab * SLAB_C;
    const int n_chunks = hh * hw * 9;
    for (int i = threadIdx.x; i < n_chunks; i += kThreadsG) {
      const int pix = i / 9, ch = i - pix * 9;
      const int hy = pix / hw, hx = pix - hy * hw;
      const int yy = iy0 + hy, xx = ix0 + hx;
      const bool ok = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      const bf16* src = src0 + (static_cast<long long>(ok ? yy : 0) * p.W + (ok ? xx : 0)) * p.C + ch * 8;
      cp_async16_zfill(dst0 + pix * PIX_BYTES + ch * 16, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // tap offsets inside the halo tile (bytes)
  int tapoff[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) tapoff[t] = ((t / 3) * hw + (t % 3)) * PIX_BYTES;

  uint32_t wk16[9][3][2], wk8[9][3];   // this warp's group weights as mma B fragments
  float sc[3][2], sh[3][2];
  float ssum[3][2], ssq[3][2];
  int cur_slab = -1;
  const bool has_stats = p.stat_sum != nullptr;
  const bool has_affine = p.scale != nullptr;

  auto flush_stats = [&](int slab) {
    if (!has_stats) return;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = ssum[nt][e], b2 = ssq[nt][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b2 += __shfl_xor_sync(0xffffffffu, b2, o);
        }
        if (gq == 0) {
          const int c = slab * SLAB_C + group * GW + nt * 8 + 2 * t4 + e;
          atomicAdd(p.stat_sum + c, a);
          atomicAdd(p.stat_sq + c, b2);
        }
        ssum[nt][e] = ssq[nt][e] = 0.f;
      }
  };

  int buf = 0;
  int item = blockIdx.x;
  if (item < n_items) load_tile(item, 0);
  for (; item < n_items; item += gridDim.x, buf ^= 1) {
    const int next = item + gridDim.x;
    if (next < n_items) {
      load_tile(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int slab = item / tiles_total;
    const int tile = item - slab * tiles_total;
    if (slab != cur_slab) {
      if (cur_slab >= 0) flush_stats(cur_slab);
      cur_slab = slab;
      const uint32_t* wg = reinterpret_cast<const uint32_t*>(p.w) +
                           static_cast<long long>(slab * 3 + group) * (9 * GW * GW / 2);
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const uint32_t* row = wg + (t * GW + nt * 8 + gq) * (GW / 2);   // [co][ci] row, 12 words
          wk16[t][nt][0] = __ldg(row + t4);
          wk16[t][nt][1] = __ldg(row + 4 + t4);
          wk8[t][nt] = __ldg(row + 8 + t4);
        }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = slab * SLAB_C + group * GW + nt * 8 + 2 * t4 + e;
          sc[nt][e] = has_affine ? __ldg(p.scale + c) : 1.f;
          sh[nt][e] = has_affine ? __ldg(p.shift + c) : 0.f;
          ssum[nt][e] = ssq[nt][e] = 0.f;
        }
    }
    const int b = tile / tiles_per_img;
    const int r = tile - b * tiles_per_img;
    const int oy0 = (r / p.tiles_x) * th, ox0 = (r % p.tiles_x) * tw;
    const uint32_t halo = smem_u32(halo_base + buf * (HALO_MAX * PIX_BYTES));
    uint8_t* stage = stage_base + warp * STAGE_BYTES;
    bf16* out_img = p.out + static_cast<long long>(b) * p.Ho * p.Wo * p.C + slab * SLAB_C + group * GW;

#pragma unroll 1
    for (int u = 0; u < MT_PER_WARP; ++u) {
      const int mt = half * MT_PER_WARP + u;
      // ldmatrix row address of this lane: pixel q of the m tile, channel sub-block by matrix index
      const int mat = lane >> 3, rr = lane & 7;
      const int q = mt * 16 + (mat & 1) * 8 + rr;
      const int ty = q >> p.tw_log2, tx = q & (tw - 1);
      const uint32_t a_base = halo + ((ty * STRIDE) * hw + tx * STRIDE) * PIX_BYTES + group * (GW * 2);
      const uint32_t a16 = a_base + (mat >> 1) * 16;   // x4: k 0-7 / 8-15
      const uint32_t a8 = a_base + 32;                 // x2: k 16-23 (lanes 16-31 repeat valid addresses)
      float acc[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        uint32_t a[4], a2[2];
        ldsm_x4(a, a16 + tapoff[t]);
        ldsm_x2(a2, a8 + tapoff[t]);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          mma16816(acc[nt], a, wk16[t][nt][0], wk16[t][nt][1]);
          mma1688(acc[nt], a2, wk8[t][nt]);
        }
      }
      // epilogue: rows gq and gq + 8 of the m tile
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int qo = mt * 16 + gq + hf * 8;
        const int oy = oy0 + (qo >> p.tw_log2), ox = ox0 + (qo & (tw - 1));
        const bool ok = oy < p.Ho && ox < p.Wo;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          float v0 = acc[nt][hf * 2], v1 = acc[nt][hf * 2 + 1];
          if (has_stats && ok) {
            ssum[nt][0] += v0;
            ssum[nt][1] += v1;
            ssq[nt][0] = fmaf(v0, v0, ssq[nt][0]);
            ssq[nt][1] = fmaf(v1, v1, ssq[nt][1]);
          }
          if (has_affine) {
            v0 = fmaf(v0, sc[nt][0], sh[nt][0]);
            v1 = fmaf(v1, sc[nt][1], sh[nt][1]);
          }
          if (p.act == ACT_RELU) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          }
          *reinterpret_cast<uint32_t*>(stage + (gq + hf * 8) * (GW * 2) + nt * 16 + t4 * 4) = pack_bf16x2(v0, v1);
        }
      }
      __syncwarp();
      // 16 pixels x 48 B = 48 chunks of 16 B: coalesced-as-possible vector stores
#pragma unroll
      for (int j = lane; j < 48; j += 32) {
        const int px = j / 3, part = j - px * 3;
        const int qo = mt * 16 + px;
        const int oy = oy0 + (qo >> p.tw_log2), ox = ox0 + (qo & (tw - 1));
        if (oy < p.Ho && ox < p.Wo) {
          const uint4 v = *reinterpret_cast<const uint4*>(stage + j * 16);
          *reinterpret_cast<uint4*>(out_img + (static_cast<long long>(oy) * p.Wo + ox) * p.C + part * 8) = v;
        }
      }
      __syncwarp();
    }
    __syncthreads();  // everyone is done with `buf` before the next iteration's prefetch overwrites it
  }
  if (cur_slab >= 0) flush_stats(cur_slab);
}

// ---------------------------------------------------------------------------------------------------------------
// Input gradient of the stride-2 conv:  dX[iy, ix] = sum over the taps (ky, kx) with (iy + 1 - ky, ix + 1 - kx) even of
// W[.,.,ky,kx]^T dY[(iy + 1 - ky) / 2, (ix + 1 - kx) / 2].  The four input parity classes (iy & 1, ix & 1) use 1, 2, 2
// and 4 taps.  A CTA stages a (TU+1) x (TV+1) tile of dY once and produces the 2TU x 2TV tile of dX class by class:
// an m tile is 16 consecutive pixels of one class (stride-1 rows in the dY tile), written back with pixel stride 2.
// Weights: the same transposed / flipped pack as the stride-1 input gradient (tap' = (2-ky)*3 + (2-kx)).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsG, 2) gconv3x3_dgrad_s2_kernel(const GConvParams p) {
  constexpr int TPIX = 128;                       // class pixels per tile (TU * TV)
  constexpr int MT = TPIX / 16;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* halo_base = smem_raw;                                   // [2][HALO_MAX][144 B]
  uint8_t* stage_base = smem_raw + 2 * HALO_MAX * PIX_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t4 = lane & 3;
  const int group = warp >> 1, half = warp & 1;
  const int tv = 1 << p.tw_log2, tu = p.th;       // dY tile: tu x tv (+1 halo row / column)
  const int hw = tv + 1, hh = tu + 1;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int tiles_total = tiles_per_img * p.B;
  const int n_items = tiles_total * p.slabs;
  // here p.x = dY (B,H,W,C) with H,W the OUTPUT size of the forward conv; p.out = dX (B,2H,2W,C)
  auto load_tile = [&](int item, int buf) {
    const int slab = item / tiles_total;
    const int tile = item - slab * tiles_total;
    const int b = tile / tiles_per_img;
    const int r = tile - b * tiles_per_img;
    const int u0 = (r / p.tiles_x) * tu, v0 = (r % p.tiles_x) * tv;
    const uint32_t dst0 = smem_u32(halo_base + buf * (HALO_MAX * PIX_BYTES));
    const bf16* src0 = p.x + static_cast<long long>(b) * p.H * p.W * p.C + slab * SLAB_C;
    const int n_chunks = hh * hw * 9;
    for (int i = threadIdx.x; i < n_chunks; i += kThreadsG) {
      const int pix = i / 9, ch = i - pix * 9;
      const int hy = pix / hw, hx = pix - hy * hw;
      const int yy = u0 + hy, xx = v0 + hx;
      const bool ok = yy < p.H && xx < p.W;
      const bf16* src = src0 + (static_cast<long long>(ok ? yy : 0) * p.W + (ok ? xx : 0)) * p.C + ch * 8;
      cp_async16_zfill(dst0 + pix * PIX_BYTES + ch * 16, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint32_t wk16[9][3][2], wk8[9][3];
  int cur_slab = -1;
  const int Wx = 2 * p.W;   // dX width
  int buf = 0;
  int item = blockIdx.x;
  if (item < n_items) load_tile(item, 0);
  for (; item < n_items; item += gridDim.x, buf ^= 1) {
    const int next = item + gridDim.x;
    if (next < n_items) {
      load_tile(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int slab = item / tiles_total;
    const int tile = item - slab * tiles_total;
    if (slab != cur_slab) {
      cur_slab = slab;
      const uint32_t* wg = reinterpret_cast<const uint32_t*>(p.w) +
                           static_cast<long long>(slab * 3 + group) * (9 * GW * GW / 2);
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const uint32_t* row = wg + (t * GW + nt * 8 + gq) * (GW / 2);
          wk16[t][nt][0] = __ldg(row + t4);
          wk16[t][nt][1] = __ldg(row + 4 + t4);
          wk8[t][nt] = __ldg(row + 8 + t4);
        }
    }
    const int b = tile / tiles_per_img;
    const int r = tile - b * tiles_per_img;
    const int u0 = (r / p.tiles_x) * tu, v0 = (r % p.tiles_x) * tv;
    const uint32_t halo = smem_u32(halo_base + buf * (HALO_MAX * PIX_BYTES));
    uint8_t* stage = stage_base + warp * STAGE_BYTES;
    bf16* out_img = p.out + static_cast<long long>(b) * (4ll * p.H * p.W) * p.C + slab * SLAB_C + group * GW;
    const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
      const int py = cls >> 1, px = cls & 1;
      if (((cls == 0 || cls == 3) ? 0 : 1) != half) continue;   // warp half 0: classes (0,0) + (1,1); half 1: the others
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int q = mt * 16 + (mat & 1) * 8 + rr;
        const int u = q >> p.tw_log2, v = q & (tv - 1);
        const uint32_t a_base = halo + (u * hw + v) * PIX_BYTES + group * (GW * 2);
        float acc[3][4];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          if (((ky + py) & 1) == 0) continue;          // iy + 1 - ky must be even
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (((kx + px) & 1) == 0) continue;
            const int doy = (py == 1 && ky == 0) ? 1 : 0, dox = (px == 1 && kx == 0) ? 1 : 0;
            const int tp = (2 - ky) * 3 + (2 - kx);    // index into the flipped pack
            uint32_t a[4], a2[2];
            const uint32_t off = (doy * hw + dox) * PIX_BYTES;
            ldsm_x4(a, a_base + off + (mat >> 1) * 16);
            ldsm_x2(a2, a_base + off + 32);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
              mma16816(acc[nt], a, wk16[tp][nt][0], wk16[tp][nt][1]);
              mma1688(acc[nt], a2, wk8[tp][nt]);
            }
          }
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
            *reinterpret_cast<uint32_t*>(stage + (gq + hf * 8) * (GW * 2) + nt * 16 + t4 * 4) =
                pack_bf16x2(acc[nt][hf * 2], acc[nt][hf * 2 + 1]);
        __syncwarp();
#pragma unroll
        for (int j = lane; j < 48; j += 32) {
          const int pxl = j / 3, part = j - pxl * 3;
          const int qo = mt * 16 + pxl;
          const int uu = u0 + (qo >> p.tw_log2), vv = v0 + (qo & (tv - 1));
          if (uu < p.H && vv < p.W) {
            const uint4 val = *reinterpret_cast<const uint4*>(stage + j * 16);
            *reinterpret_cast<uint4*>(out_img + (static_cast<long long>(2 * uu + py) * Wx + 2 * vv + px) * p.C + part * 8) = val;
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient: dW[co][ci][tap] += sum_pixels dY[pix][co] * X[pix*stride + tap - 1][ci] inside each group of 24.
// grid = (CTAs per slab, slabs).  A CTA stages the dY tile and the haloed X tile of one 72-channel slab, contracts
// over the tile's pixels with mma.sync (operands fetched with ldmatrix.trans: pixel-major storage -> channel-major
// fragments) and keeps its share of the 3 x 9 x 24 x 24 products in registers across all its tiles: warp = (group,
// kernel row).  Partials go to a workspace with plain stores; a second tiny kernel adds them into dW (no atomics).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpsW = 9;
constexpr int kThreadsW = kWarpsW * 32;
constexpr int SLAB_W = SLAB_C * GW * 9;   // 15552 weight-gradient elements per slab, torch layout [co][ci][ky][kx]

__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}

struct GWgradParams {
  const bf16* dy;   // (B,Ho,Wo,C)
  const bf16* x;    // (B,H,W,C)
  float* part;      // workspace [slabs][gridDim.x][SLAB_W]
  int B, H, W, C, Ho, Wo;
  int tw_log2, th;
  int tiles_x, tiles_y;
};

template <int STRIDE>
__global__ void __launch_bounds__(kThreadsW, 2) gconv3x3_wgrad_kernel(const GWgradParams p) {
  constexpr int TPIX = STRIDE == 1 ? 256 : 64;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* xs = smem_raw;                           // [HALO_MAX][144 B]
  uint8_t* dys = smem_raw + HALO_MAX * PIX_BYTES;   // [TPIX][144 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t4 = lane & 3;
  const int group = warp / 3, ky = warp - group * 3;
  const int slab = blockIdx.y;
  const int tw = 1 << p.tw_log2, th = p.th;
  const int hw = (tw - 1) * STRIDE + 3, hh = (th - 1) * STRIDE + 3;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int tiles_total = tiles_per_img * p.B;
  float acc[3][2][3][4];   // [kx][m tile][n tile][frag]
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 3; ++n) acc[a][m][n][0] = acc[a][m][n][1] = acc[a][m][n][2] = acc[a][m][n][3] = 0.f;
  const uint32_t xs_u = smem_u32(xs), dys_u = smem_u32(dys);
  const int mat = lane >> 3, rr = lane & 7;

  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int r = tile - b * tiles_per_img;
    const int oy0 = (r / p.tiles_x) * th, ox0 = (r % p.tiles_x) * tw;
    __syncthreads();  // previous tile fully consumed
    {
      const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
      const bf16* src0 = p.x + static_cast<long long>(b) * p.H * p.W * p.C + slab * SLAB_C;
      for (int i = threadIdx.x; i < hh * hw * 9; i += kThreadsW) {
        const int pix = i / 9, ch = i - pix * 9;
        const int hy = pix / hw, hx = pix - hy * hw;
        const int yy = iy0 + hy, xx = ix0 + hx;
        const bool ok = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
        cp_async16_zfill(xs_u + pix * PIX_BYTES + ch * 16,
                         src0 + (static_cast<long long>(ok ? yy : 0) * p.W + (ok ? xx : 0)) * p.C + ch * 8, ok);
      }
      const bf16* dsrc0 = p.dy + static_cast<long long>(b) * p.Ho * p.Wo * p.C + slab * SLAB_C;
      for (int i = threadIdx.x; i < TPIX * 9; i += kThreadsW) {
        const int q = i / 9, ch = i - q * 9;
        const int oy = oy0 + (q >> p.tw_log2), ox = ox0 + (q & (tw - 1));
        const bool ok = oy < p.Ho && ox < p.Wo;
        cp_async16_zfill(dys_u + q * PIX_BYTES + ch * 16,
                         dsrc0 + (static_cast<long long>(ok ? oy : 0) * p.Wo + (ok ? ox : 0)) * p.C + ch * 8, ok);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll 1
    for (int ks = 0; ks < TPIX / 16; ++ks) {
      // A fragments (m = out channel of the group, k = 16 tile pixels): m tile 0 = channels 0-15, m tile 1 = 16-23
      uint32_t a0[4], a1[4];
      {
        const int q = ks * 16 + (mat >> 1) * 8 + rr;
        ldsm_x4_t(a0, dys_u + q * PIX_BYTES + group * (GW * 2) + (mat & 1) * 16);
        uint32_t t2[2];
        const int q2 = ks * 16 + (mat & 1) * 8 + rr;
        ldsm_x2_t(t2, dys_u + q2 * PIX_BYTES + group * (GW * 2) + 32);
        a1[0] = t2[0]; a1[1] = 0u; a1[2] = t2[1]; a1[3] = 0u;
      }
      // B fragments (k = the same pixels shifted by the tap, n = in channel of the group)
      const int qb = ks * 16 + (mat & 1) * 8 + rr;
      const int ty = qb >> p.tw_log2, tx = qb & (tw - 1);
      const uint32_t xrow = xs_u + ((ty * STRIDE + ky) * hw + tx * STRIDE) * PIX_BYTES + group * (GW * 2);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        uint32_t b01[4], b2[2];
        ldsm_x4_t(b01, xrow + kx * PIX_BYTES + (mat >> 1) * 16);   // n tiles 0, 1: {b0,b1} each
        ldsm_x2_t(b2, xrow + kx * PIX_BYTES + 32);                 // n tile 2
        mma16816(acc[kx][0][0], a0, b01[0], b01[1]);
        mma16816(acc[kx][0][1], a0, b01[2], b01[3]);
        mma16816(acc[kx][0][2], a0, b2[0], b2[1]);
        mma16816(acc[kx][1][0], a1, b01[0], b01[1]);
        mma16816(acc[kx][1][1], a1, b01[2], b01[3]);
        mma16816(acc[kx][1][2], a1, b2[0], b2[1]);
      }
    }
  }
  // partials: torch layout inside the slab, element (co, ci, tap) at (co * 24 + ci) * 9 + tap
  float* dst = p.part + (static_cast<long long>(slab) * gridDim.x + blockIdx.x) * SLAB_W;
#pragma unroll
  for (int kx = 0; kx < 3; ++kx)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = m * 16 + gq + (e >> 1) * 8;
          const int ci = n * 8 + 2 * t4 + (e & 1);
          if (co < GW) dst[((group * GW + co) * GW + ci) * 9 + ky * 3 + kx] = acc[kx][m][n][e];
        }
}

__global__ void __launch_bounds__(256) gconv3x3_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw,
                                                                    int ctas, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long slab = i / SLAB_W, e = i - slab * SLAB_W;
  const float* src = part + slab * ctas * SLAB_W + e;
  float a = 0.f;
  for (int k = 0; k < ctas; ++k) a += src[static_cast<long long>(k) * SLAB_W];
  dw[i] += a;
}

}  // namespace

static void gconv_tile(int stride, int wo, int* tw_log2, int* th);

extern "C" int tfpp_gconv3x3(const void* x, const void* w, void* out, const float* scale, const float* shift, int act,
                             float* stat_sum, float* stat_sq, int batch, int height, int width, int channels, int stride,
                             tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(channels % SLAB_C == 0, "channels must be a multiple of 72 (3 groups of width 24)");
  TFPP_CHECK_ARG(stride == 1 || stride == 2, "stride 1 or 2");
  TFPP_CHECK_ARG(stride == 1 || (height % 2 == 0 && width % 2 == 0), "stride 2 needs even H, W");
  TFPP_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  TFPP_CHECK_ARG((stat_sum == nullptr) == (stat_sq == nullptr), "stat_sum and stat_sq go together");
  TFPP_CHECK_ARG(act == ACT_NONE || act == ACT_RELU, "activation: none or relu");
  GConvParams p;
  p.x = static_cast<const bf16*>(x); p.w = static_cast<const bf16*>(w); p.out = static_cast<bf16*>(out);
  p.scale = scale; p.shift = shift; p.act = act; p.stat_sum = stat_sum; p.stat_sq = stat_sq;
  p.B = batch; p.H = height; p.W = width; p.C = channels;
  p.Ho = height / stride; p.Wo = width / stride;
  int tw_log2;
  gconv_tile(stride, p.Wo, &tw_log2, &p.th);
  p.tw_log2 = tw_log2;
  p.tiles_x = ceil_div(p.Wo, 1 << tw_log2);
  p.tiles_y = ceil_div(p.Ho, p.th);
  p.slabs = channels / SLAB_C;
  const long long items = static_cast<long long>(p.tiles_x) * p.tiles_y * batch * p.slabs;
  if (items == 0) return TFPP_OK;
  const size_t smem = 2 * HALO_MAX * PIX_BYTES + kWarps * STAGE_BYTES;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gconv3x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(gconv3x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    attr = true;
  }
  const int grid = static_cast<int>(items < 2 * TFPP_NUM_SMS ? items : 2 * TFPP_NUM_SMS);
  if (stride == 1) gconv3x3_kernel<1><<<grid, kThreadsG, smem, stream>>>(p);
  else gconv3x3_kernel<2><<<grid, kThreadsG, smem, stream>>>(p);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

static void gconv_tile(int stride, int wo, int* tw_log2, int* th) {
  const int tpix = stride == 1 ? 256 : 64;
  const int tw_max = stride == 1 ? 32 : 16;
  int l = 3;  // at least 8 wide (an m tile is 16 consecutive tile pixels: 1 or 2 rows)
  while ((1 << l) < wo && (1 << l) < tw_max) ++l;
  *tw_log2 = l;
  *th = tpix >> l;
}

extern "C" long long tfpp_gconv3x3_wgrad_workspace(int batch, int height, int width, int channels, int stride) {
  int l, th;
  gconv_tile(stride, width / stride, &l, &th);
  const long long tiles = static_cast<long long>(ceil_div(width / stride, 1 << l)) * ceil_div(height / stride, th) * batch;
  const int slabs = channels / SLAB_C;
  long long ctas = 2 * TFPP_NUM_SMS / slabs;
  if (ctas < 1) ctas = 1;
  if (ctas > tiles) ctas = tiles;
  return ctas * slabs * SLAB_W;   // floats
}

extern "C" int tfpp_gconv3x3_wgrad(const void* dy, const void* x, float* dw, float* workspace, int batch, int height,
                                   int width, int channels, int stride, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(channels % SLAB_C == 0, "channels must be a multiple of 72 (3 groups of width 24)");
  TFPP_CHECK_ARG(stride == 1 || stride == 2, "stride 1 or 2");
  TFPP_CHECK_ARG(stride == 1 || (height % 2 == 0 && width % 2 == 0), "stride 2 needs even H, W");
  GWgradParams p;
  p.dy = static_cast<const bf16*>(dy); p.x = static_cast<const bf16*>(x); p.part = workspace;
  p.B = batch; p.H = height; p.W = width; p.C = channels; p.Ho = height / stride; p.Wo = width / stride;
  gconv_tile(stride, p.Wo, &p.tw_log2, &p.th);
  p.tiles_x = ceil_div(p.Wo, 1 << p.tw_log2);
  p.tiles_y = ceil_div(p.Ho, p.th);
  const long long tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * batch;
  if (tiles == 0) return TFPP_OK;
  const int slabs = channels / SLAB_C;
  long long ctas = 2 * TFPP_NUM_SMS / slabs;
  if (ctas < 1) ctas = 1;
  if (ctas > tiles) ctas = tiles;
  const int tpix = stride == 1 ? 256 : 64;
  const size_t smem = static_cast<size_t>(HALO_MAX + tpix) * PIX_BYTES;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gconv3x3_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (HALO_MAX + 256) * PIX_BYTES);
    cudaFuncSetAttribute(gconv3x3_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (HALO_MAX + 64) * PIX_BYTES);
    attr = true;
  }
  dim3 grid(static_cast<unsigned>(ctas), slabs);
  if (stride == 1) gconv3x3_wgrad_kernel<1><<<grid, kThreadsW, smem, stream>>>(p);
  else gconv3x3_wgrad_kernel<2><<<grid, kThreadsW, smem, stream>>>(p);
  TFPP_CHECK_LAUNCH();
  const long long total = static_cast<long long>(slabs) * SLAB_W;
  gconv3x3_wgrad_reduce_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      workspace, dw, static_cast<int>(ctas), total);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_gconv3x3_dgrad_s2(const void* dy, const void* w_t, void* dx, int batch, int out_height, int out_width,
                                      int channels, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(channels % SLAB_C == 0, "channels must be a multiple of 72 (3 groups of width 24)");
  GConvParams p;
  p.x = static_cast<const bf16*>(dy); p.w = static_cast<const bf16*>(w_t); p.out = static_cast<bf16*>(dx);
  p.scale = nullptr; p.shift = nullptr; p.act = ACT_NONE; p.stat_sum = nullptr; p.stat_sq = nullptr;
  p.B = batch; p.H = out_height; p.W = out_width; p.C = channels; p.Ho = 2 * out_height; p.Wo = 2 * out_width;
  int l = 3;
  while ((1 << l) < out_width && (1 << l) < 16) ++l;
  p.tw_log2 = l;
  p.th = 128 >> l;
  p.tiles_x = ceil_div(out_width, 1 << l);
  p.tiles_y = ceil_div(out_height, p.th);
  p.slabs = channels / SLAB_C;
  const long long items = static_cast<long long>(p.tiles_x) * p.tiles_y * batch * p.slabs;
  if (items == 0) return TFPP_OK;
  const size_t smem = 2 * HALO_MAX * PIX_BYTES + kWarps * STAGE_BYTES;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gconv3x3_dgrad_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    attr = true;
  }
  const int grid = static_cast<int>(items < 2 * TFPP_NUM_SMS ? items : 2 * TFPP_NUM_SMS);
  gconv3x3_dgrad_s2_kernel<<<grid, kThreadsG, smem, stream>>>(p);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
