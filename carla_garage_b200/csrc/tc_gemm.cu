// tcgen05 implicit-GEMM convolution / linear kernel for sm_100a.
//
//   out[pixel, n] = act(scale[n] * sum_{tap, c} A[pixel + shift(tap), c0 + c] * W[n, tap, c] + shift[n] + res1 + res2)
//
// One persistent CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane issues
// tcgen05.mma, accumulators live in TMEM, double buffered), warps 2-5 = epilogue (tcgen05.ld -> registers ->
// fused BatchNorm/bias/activation/residual/statistics -> global).  A tiles are 128 pixels x 64 channels fetched with
// 4-D TMA boxes (C, W, H, B) so that 3x3 taps are plain coordinate shifts with hardware zero fill (= conv padding);
// B tiles are BN x 64 boxes of the (N, taps, K) weight tensor.  Both land in the canonical K-major SWIZZLE_128B
// layout the UMMA shared-memory descriptors expect.
//
// Replaces (reference, torch library dispatches): timm RegNet 1x1 / grouped 3x3 convs iterated at
// team_code/transfuser.py:216-219, nn.Linear at transfuser.py:352-359,391-396, the 1x1 channel maps at
// transfuser.py:233,237, FPN/decoder/head convs (transfuser.py:131-137, transfuser_utils.py:675-704,
// model.py:75-90,148, center_net.py:43-47) and the decoder projections (model.py:137-143).
#include "../../include/tfpp.h"
#include "tc_common.cuh"

#include <cstdlib>

using namespace tc;

namespace {

constexpr int kThreads = 320;                       // TMA warp, MMA warp, 8 epilogue warps
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // bf16 elements = 128 bytes = one SWIZZLE_128B row
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;                      // 2 accumulator stages x 256 columns
constexpr int kAccStride = 256;

struct KParams {
  int batch, height, width;
  int kc_per_tap, a_c_per_ntile;
  int n, bn, n_tiles, m_tiles_x, m_tiles_y, m_tiles_b;
  int tw, th, nb;
  int ntaps;
  int tap_dx[9], tap_dy[9], tap_db[9], tap_w[9];
  int stages, b_stage_bytes;
  void* out;
  int out_f32;
  long long o_sb, o_sy, o_sx, o_sn;
  const void* res1;
  int res1_f32;
  long long r1_sb, r1_sy, r1_sx, r1_sn;
  const void* res2;
  int res2_f32;
  long long r2_sb, r2_sy, r2_sx, r2_sn;
  const float* scale;
  const float* shift;
  int act, act_n_limit;
  float* stat_sum;
  float* stat_sq;
  const unsigned long long* drop_rng;
  float drop_p;
  unsigned drop_site;
  int fast_layout;   // bf16/fp32 NHWC output with unit channel stride, <= 1 residual of the same kind, none/ReLU
  int stat_floats;   // 2 * n_tiles * bn when statistics are requested, else 0
};

struct TileCoord {
  int n0, x0, y0, b0, n_tile;
};

// m_mul / m_add: a CTA pair's work item covers the M tiles 2 * item + {0, 1} (an odd tail tile decodes to a batch index
// past the end: its TMA boxes are zero-filled and its rows are masked in the epilogue)
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int tile, int m_mul = 1, int m_add = 0) {
  TileCoord t;
  t.n_tile = tile % p.n_tiles;
  int m = (tile / p.n_tiles) * m_mul + m_add;
  t.n0 = t.n_tile * p.bn;
  t.x0 = (m % p.m_tiles_x) * p.tw;
  m /= p.m_tiles_x;
  t.y0 = (m % p.m_tiles_y) * p.th;
  t.b0 = (m / p.m_tiles_y) * p.nb;
  return t;
}

// PAIR: two CTAs of a cluster (cta_group::2) share one M = 256 x BN tile: CTA r stages its own 128 pixels of A and rows
// [r * BN / 2, (r + 1) * BN / 2) of the B tile, the leader issues M = 256 MMAs that read both CTAs' shared memory, so each
// SM pulls 16 KB + BN * 64 B per k-block through L2 instead of 16 KB + BN * 128 B (the GEMMs are L2 -> SM bound).
// LEAN: epilogue specialisations compiled as their own kernels (the K <= 576 layers are bound by the latency of the ten
// warps' epilogue instruction stream, and the 168-register generic epilogue must not pay for them): 1 = bf16 NHWC output +
// BatchNorm statistics, nothing else (RegNet convs of a training step); 2 = bf16 NHWC output + at most one bf16
// residual (input-gradient GEMMs); 3 = bf16 NHWC output + bias (+ ReLU) (convs / linears with a bias); 0 = everything.
template <bool PAIR, int LEAN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const KParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16 KB | B b_stage_bytes)] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kABytes + p.b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  float2* saff = reinterpret_cast<float2*>(bars + 2 * kMaxStages + 6);   // [256] per-tile {scale, shift}
  float* sstat = reinterpret_cast<float*>(saff + 256);                   // [2][stat_stride] BatchNorm partial sums
  float* trbuf = sstat + p.stat_floats;   // 8 x [32][33] epilogue transpose tiles (only when statistics / generic path)
  const int stat_stride = p.stat_floats / 2;
  if (p.stat_sum != nullptr)
    for (int i = threadIdx.x; i < 2 * stat_stride; i += blockDim.x) sstat[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader of the CTA pair
  const int m_tiles = p.m_tiles_x * p.m_tiles_y * p.m_tiles_b;
  // work items: (n tile, M tile) — or (n tile, pair of consecutive M tiles) for a CTA pair; both CTAs of a pair walk the
  // same item sequence
  const int num_tiles = p.n_tiles * (PAIR ? (m_tiles + 1) / 2 : m_tiles);
  const int first_tile = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int k_iters = p.ntaps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), PAIR ? 16 : 8);   // the leader's barrier collects both CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int b_rows = PAIR ? p.bn / 2 : p.bn;   // B rows this CTA stages
      const uint32_t tx_bytes = (kABytes + b_rows * kBlockK * 2) * (PAIR ? 2 : 1);   // the leader's barrier counts both CTAs
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const TileCoord t = decode_tile(p, tile, PAIR ? 2 : 1, static_cast<int>(rank));
        const int c_base = t.n_tile * p.a_c_per_ntile;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int kc = 0; kc < p.kc_per_tap; ++kc) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
            if (PAIR) {
              if (rank == 0) mbar_expect_tx(fb, tx_bytes);
              const uint32_t lb = mapa_shared(fb, 0);   // completion goes to the leader's barrier
              tma_load_4d_pair(smem_u32(sa), &tmap_a, lb, c_base + kc * kBlockK, t.x0 + p.tap_dx[tap],
                               t.y0 + p.tap_dy[tap], t.b0 + p.tap_db[tap]);
              tma_load_3d_pair(smem_u32(sa + kABytes), &tmap_b, lb, kc * kBlockK, p.tap_w[tap],
                               t.n0 + static_cast<int>(rank) * b_rows);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
              continue;
            }
            mbar_expect_tx(fb, tx_bytes);
            tma_load_4d(smem_u32(sa), &tmap_a, fb, c_base + kc * kBlockK, t.x0 + p.tap_dx[tap], t.y0 + p.tap_dy[tap],
                        t.b0 + p.tap_db[tap]);
            tma_load_3d(smem_u32(sa + kABytes), &tmap_b, fb, kc * kBlockK, p.tap_w[tap], t.n0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (rank == 0 && elect_one()) {  // one elected lane (of the leader CTA): descriptors stay in uniform registers
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, p.bn);
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint64_t adesc = make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = make_sw128_kmajor_desc(sa + kABytes);
#pragma unroll
          for (int kk = 0; kk < kBlockK / 16; ++kk) {
            // advance 16 bf16 = 32 bytes inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if (PAIR)
              umma_bf16_pair(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) ? 1u : 0u);
            else
              umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) ? 1u : 0u);
          }
          if (PAIR)
            umma_commit_pair(smem_u32(&empty_bar[stage]));   // frees the stage in both CTAs
          else
            umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (PAIR)
          umma_commit_pair(smem_u32(&tfull_bar[acc]));       // both CTAs' epilogues may read their 128 accumulator rows
        else
          umma_commit(smem_u32(&tfull_bar[acc]));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..9)
    // Two warps per SM sub-partition share a 32-lane quarter of the accumulator and split its 32-column slabs between
    // them (even / odd slab).  Every instruction's latency is exposed with so few warps, so the hot path keeps ~2
    // instructions per output element: loop invariants hoisted, the per-tile affine staged in shared memory, output in
    // 8-column groups (one 16-byte store for bf16, two for fp32) that also cover ragged last slabs.  A compact generic
    // path (dynamic loops over a shared-memory copy of the slab) serves every other layout.  BatchNorm statistics:
    // transpose the slab through shared memory so that lane == column and column sums are register accumulations.
    const int ew = warp - 2;                     // 0..7
    const int lane_group = warp & 3;             // TMEM lanes [32*lane_group, +32) are accessible to this warp
    const int half = ew >> 2;                    // slab parity owned by this warp
    const int row = lane_group * 32 + lane;      // tile row == pixel
    const int pix_per_img = p.th * p.tw;
    float* tr = trbuf + ew * (32 * 33);          // [32][33] fp32 transpose tile of this warp (if allocated)
    const float* __restrict__ g_scale = p.scale;
    const float* __restrict__ g_shift = p.shift;
    const bool has_stats = p.stat_sum != nullptr;
    const bool has_affine = (g_scale != nullptr) || (g_shift != nullptr);
    const int act = p.act;
    const DropCtx drop = drop_ctx(p.drop_rng, p.drop_p, p.drop_site);
    const bool fast_layout = p.fast_layout != 0;
    const bool out_f32 = p.out_f32 != 0;
    const bf16* __restrict__ res_b = (p.res1 != nullptr && !p.res1_f32) ? static_cast<const bf16*>(p.res1) : nullptr;
    const float* __restrict__ res_f = (p.res1 != nullptr && p.res1_f32) ? static_cast<const float*>(p.res1) : nullptr;
    bf16* __restrict__ out_b = static_cast<bf16*>(p.out);
    float* __restrict__ out_f = static_cast<float*>(p.out);
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader0 = PAIR ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const TileCoord t = decode_tile(p, tile, PAIR ? 2 : 1, static_cast<int>(rank));
      const int bl = row / pix_per_img;
      const int rem = row - bl * pix_per_img;
      const int yy = t.y0 + rem / p.tw;
      const int xx = t.x0 + rem % p.tw;
      const int bb = t.b0 + bl;
      const bool valid = (bb < p.batch) && (yy < p.height) && (xx < p.width);
      const long long o_base = bb * p.o_sb + yy * p.o_sy + xx * p.o_sx;
      const long long r1_base = bb * p.r1_sb + yy * p.r1_sy + xx * p.r1_sx;
      const long long r2_base = bb * p.r2_sb + yy * p.r2_sy + xx * p.r2_sx;
      const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
      const int nend = min(p.n, t.n0 + p.bn);  // columns [n0, nend) of this tile are real outputs
      const int ncols = nend - t.n0;
      // per-tile affine (BatchNorm fold / bias) staged once in shared memory: {scale, shift} per column
      if (has_affine) {
        asm volatile("bar.sync 2, 256;" ::: "memory");  // previous tile's readers are done
        for (int i = ew * 32 + lane; i < ncols; i += 256)
          saff[i] = make_float2(g_scale ? __ldg(g_scale + t.n0 + i) : 1.f, g_shift ? __ldg(g_shift + t.n0 + i) : 0.f);
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }

      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + acc * kAccStride;
      for (int c = half * 32; c < ncols; c += 64) {
        uint32_t r[32];
        const int n = t.n0 + c;
        const int ng = min(4, (ncols - c) >> 3);   // 8-column groups of this slab that are real outputs (fast path)
        if (LEAN != 0) {
          uint4 rq[4];
          const bool with_res = LEAN == 2 && res_b != nullptr;
          if (with_res && valid) {
            const uint4* rp = reinterpret_cast<const uint4*>(res_b + r1_base + n);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < ng) rq[q] = __ldg(rp + q);
          }
          __syncwarp();
          tmem_ld32_issue(taddr + c, r);
          tmem_ld_wait32(r);
          if (LEAN == 1) {
            const uint32_t trs = smem_u32(tr);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(trs + (lane * 33 + j) * 4), "r"(r[j]) : "memory");
            __syncwarp();
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;   // two accumulation chains
#pragma unroll
            for (int rr = 0; rr < 32; rr += 2) {
              float a0, a1;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a0) : "r"(trs + (rr * 33 + lane) * 4) : "memory");
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a1) : "r"(trs + ((rr + 1) * 33 + lane) * 4) : "memory");
              if (valid_mask != 0xffffffffu) {
                a0 = ((valid_mask >> rr) & 1u) ? a0 : 0.f;
                a1 = ((valid_mask >> (rr + 1)) & 1u) ? a1 : 0.f;
              }
              s0 += a0;
              s1 += a1;
              q0 = fmaf(a0, a0, q0);
              q1 = fmaf(a1, a1, q1);
            }
            if (n + lane < nend) {
              atomicAdd(&sstat[n + lane], s0 + s1);
              atomicAdd(&sstat[stat_stride + n + lane], q0 + q1);
            }
          }
          if (valid) {
            bf16* op = out_b + o_base + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < ng) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[q * 8 + j]);
                if (LEAN == 3) {  // + bias[n] (staged per tile in saff as {1, bias}), optional ReLU
#pragma unroll
                  for (int j = 0; j < 8; j += 2) {
                    const float4 a2 = *reinterpret_cast<const float4*>(&saff[c + q * 8 + j]);
                    v[j] += a2.y;
                    v[j + 1] += a2.w;
                  }
                  if (act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                  }
                }
                if (with_res) {
                  const uint32_t w[4] = {rq[q].x, rq[q].y, rq[q].z, rq[q].w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16x2(w[j]);
                    v[2 * j] += f.x;
                    v[2 * j + 1] += f.y;
                  }
                }
                *reinterpret_cast<uint4*>(op + q * 8) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                                   pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              }
            }
          }
          if (LEAN == 1) __syncwarp();   // the transpose tile is reused by the next slab
          continue;
        }
        // residual of this row's slab: issue the loads before the TMEM wait
        uint4 rq[4];
        float4 rf[8];
        if (fast_layout && valid) {
          if (res_b != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(res_b + r1_base + n);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < ng) rq[q] = __ldg(rp + q);
          } else if (res_f != nullptr) {
            const float4* rp = reinterpret_cast<const float4*>(res_f + r1_base + n);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < 2 * ng) rf[q] = __ldg(rp + q);
          }
        }
        __syncwarp();
        tmem_ld32_issue(taddr + c, r);
        tmem_ld_wait32(r);
        if (has_stats || !fast_layout) {
#pragma unroll
          for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
        }
        if (has_stats) {
          const int col = n + lane;
          float ssum = 0.f, ssq = 0.f;
          if (valid_mask == 0xffffffffu) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              const float a = tr[rr * 33 + lane];
              ssum += a;
              ssq = fmaf(a, a, ssq);
            }
          } else {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              const float a = ((valid_mask >> rr) & 1u) ? tr[rr * 33 + lane] : 0.f;
              ssum += a;
              ssq = fmaf(a, a, ssq);
            }
          }
          if (col < nend) {
            atomicAdd(&sstat[col], ssum);
            atomicAdd(&sstat[stat_stride + col], ssq);
          }
        }
        if (fast_layout) {
          if (valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < ng) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[q * 8 + j]);
                if (has_affine) {
#pragma unroll
                  for (int j = 0; j < 8; j += 2) {
                    const float4 a2 = *reinterpret_cast<const float4*>(&saff[c + q * 8 + j]);  // {sc0, sh0, sc1, sh1}
                    v[j] = fmaf(v[j], a2.x, a2.y);
                    v[j + 1] = fmaf(v[j + 1], a2.z, a2.w);
                  }
                }
                if (drop.on) {  // out = drop(act(affine(acc))) + res: activation first, residual after the mask
                  if (act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                  }
                  float dm[8];
                  drop_mult8(drop, static_cast<unsigned long long>(o_base + n + q * 8), dm);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] *= dm[j];
                }
                if (res_b != nullptr) {
                  const uint32_t w[4] = {rq[q].x, rq[q].y, rq[q].z, rq[q].w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16x2(w[j]);
                    v[2 * j] += f.x;
                    v[2 * j + 1] += f.y;
                  }
                } else if (res_f != nullptr) {
                  v[0] += rf[2 * q].x; v[1] += rf[2 * q].y; v[2] += rf[2 * q].z; v[3] += rf[2 * q].w;
                  v[4] += rf[2 * q + 1].x; v[5] += rf[2 * q + 1].y; v[6] += rf[2 * q + 1].z; v[7] += rf[2 * q + 1].w;
                }
                if (act == ACT_RELU && !drop.on) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (out_f32) {
                  float4* op = reinterpret_cast<float4*>(out_f + o_base + n + q * 8);
                  op[0] = make_float4(v[0], v[1], v[2], v[3]);
                  op[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
                  *reinterpret_cast<uint4*>(out_b + o_base + n + q * 8) =
                      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                 pack_bf16x2(v[6], v[7]));
                }
              }
            }
          }
        } else if (p.out != nullptr && valid) {
          // generic: walk the columns of this slab from the shared-memory copy (any strides / dtypes / tails)
          const int jn = min(32, nend - n);
          for (int j = 0; j < jn; ++j) {
            const int col = n + j;
            float v = tr[lane * 33 + j];
            if (has_affine) {
              const float2 a2 = saff[c + j];
              v = fmaf(v, a2.x, a2.y);
            }
            const long long off = o_base + col * p.o_sn;
            if (drop.on) {  // drop(act(.)) + residuals
              if (act != ACT_NONE && (p.act_n_limit == 0 || col < p.act_n_limit)) v = apply_act(v, act);
              v *= drop_mult(drop, static_cast<unsigned long long>(off));
            }
            if (p.res1) v += load_res(p.res1, p.res1_f32, r1_base + col * p.r1_sn);
            if (p.res2) v += load_res(p.res2, p.res2_f32, r2_base + col * p.r2_sn);
            if (!drop.on && act != ACT_NONE && (p.act_n_limit == 0 || col < p.act_n_limit)) v = apply_act(v, act);
            if (p.out_f32)
              static_cast<float*>(p.out)[off] = v;
            else
              static_cast<bf16*>(p.out)[off] = f2bf(v);
          }
        }
        if (has_stats || !fast_layout) __syncwarp();  // the transpose tile is reused by the next slab
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR)
          mbar_arrive_cluster(tempty_leader0 + acc * 8);   // the leader's MMA thread owns the accumulator hand-back
        else
          mbar_arrive(smem_u32(&tempty_bar[acc]));
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.stat_sum != nullptr) {
      // flush the per-CTA statistics: one global atomic per column per CTA instead of one per warp per tile
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = ew * 32 + lane;
      for (int i = et; i < p.n; i += 256) {
        const float a = sstat[i], b2 = sstat[stat_stride + i];
        if (a != 0.f || b2 != 0.f) {
          atomicAdd(p.stat_sum + i, a);
          atomicAdd(p.stat_sq + i, b2);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM until the very end
  if (warp == 1) {
    tc_fence_after();
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace

extern "C" int tfpp_conv_gemm(const tfpp_conv_gemm_args* a, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(a != nullptr && a->a != nullptr && a->w != nullptr, "null operand");
  TFPP_CHECK_ARG(a->bn >= 16 && a->bn <= 256 && a->bn % 16 == 0, "bn must be a multiple of 16 in [16,256]");
  TFPP_CHECK_ARG(a->tw * a->th * a->nb == kBlockM, "tw*th*nb must be 128");
  TFPP_CHECK_ARG(a->tw <= 256 && a->th <= 256 && a->nb <= 256, "tile extents must be <= 256");
  TFPP_CHECK_ARG(a->ntaps >= 1 && a->ntaps <= 9, "1..9 taps");
  TFPP_CHECK_ARG(a->a_channels % 8 == 0 && a->w_kdim % 8 == 0, "channel counts must be multiples of 8 (16 B TMA strides)");
  TFPP_CHECK_ARG((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
                 "operands must be 16 B aligned");
  TFPP_CHECK_ARG(a->a_batch_stride % 8 == 0, "a_batch_stride must be a multiple of 8 elements");
  TFPP_CHECK_ARG(a->k_per_tile >= 1 && a->k_per_tile <= a->w_kdim, "k_per_tile must be <= w_kdim");
  TFPP_CHECK_ARG(a->out != nullptr || a->stat_sum != nullptr, "nothing to produce");
  TFPP_CHECK_ARG((a->stat_sum == nullptr) == (a->stat_sq == nullptr), "stat_sum and stat_sq go together");

  KParams p;
  p.batch = a->batch;
  p.height = a->height;
  p.width = a->width;
  p.kc_per_tap = ceil_div(a->k_per_tile, kBlockK);
  p.a_c_per_ntile = a->a_c_per_ntile;
  p.n = a->n;
  p.bn = a->bn;
  p.n_tiles = ceil_div(a->n, a->bn);
  p.tw = a->tw;
  p.th = a->th;
  p.nb = a->nb;
  p.m_tiles_x = ceil_div(a->width, a->tw);
  p.m_tiles_y = ceil_div(a->height, a->th);
  p.m_tiles_b = ceil_div(a->batch, a->nb);
  p.ntaps = a->ntaps;
  for (int i = 0; i < 9; ++i) {
    p.tap_dx[i] = a->tap_dx[i];
    p.tap_dy[i] = a->tap_dy[i];
    p.tap_db[i] = a->tap_db[i];
    p.tap_w[i] = a->tap_w[i];
  }
  // CTA pairs (cta_group::2) for the wide, deep GEMMs: they are bound by the L2 -> SM operand feed and a pair halves the
  // B bytes each SM pulls (measured on B200, B=32 step shapes: K=1512 N=6048 805 -> 975 TFLOP/s, K=6048 N=1512 1031 ->
  // 1135; the K=576 RegNet 1x1 convs get slower — 9 k-blocks per tile do not amortise the cross-CTA hand-offs — hence the
  // K >= 1024 rule).  TFPP_GEMM_PAIR=0 switches the path off, =2 forces it wherever the shape allows (tests).
  static const int pair_mode = [] { const char* e = getenv("TFPP_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  const int m_tiles_total = p.m_tiles_x * p.m_tiles_y * p.m_tiles_b;
  const long long k_total = static_cast<long long>(p.ntaps) * a->k_per_tile;
  const bool pair = pair_mode != 0 && a->bn % 16 == 0 && a->bn >= 64 && m_tiles_total >= 2 &&
                    (pair_mode == 2 || (a->bn >= 128 && k_total >= 1024 && m_tiles_total * p.n_tiles >= 2 * TFPP_NUM_SMS));
  const int b_rows = pair ? a->bn / 2 : a->bn;
  p.b_stage_bytes = ((b_rows * kBlockK * 2 + 1023) / 1024) * 1024;
  const int stage_bytes = kABytes + p.b_stage_bytes;
  p.out = a->out;
  p.out_f32 = a->out_f32;
  p.o_sb = a->o_sb; p.o_sy = a->o_sy; p.o_sx = a->o_sx; p.o_sn = a->o_sn;
  p.res1 = a->res1; p.res1_f32 = a->res1_f32;
  p.r1_sb = a->r1_sb; p.r1_sy = a->r1_sy; p.r1_sx = a->r1_sx; p.r1_sn = a->r1_sn;
  p.res2 = a->res2; p.res2_f32 = a->res2_f32;
  p.r2_sb = a->r2_sb; p.r2_sy = a->r2_sy; p.r2_sx = a->r2_sx; p.r2_sn = a->r2_sn;
  p.scale = a->scale;
  p.shift = a->shift;
  p.act = a->act;
  p.act_n_limit = a->act_n_limit;
  p.stat_sum = a->stat_sum;
  p.stat_sq = a->stat_sq;
  p.drop_rng = a->drop_p > 0.f ? a->drop_rng : nullptr;
  p.drop_p = a->drop_p;
  p.drop_site = a->drop_site;
  TFPP_CHECK_ARG(a->drop_p >= 0.f && a->drop_p < 1.f, "dropout probability must be in [0, 1)");

  CUtensorMap tmap_a, tmap_b;
  {
    const cuuint64_t c = a->a_channels, w = a->width, h = a->height, b = a->a_batch;
    const cuuint64_t dims[4] = {c, w, h, b};
    const cuuint64_t img = a->a_batch_stride > 0 ? (cuuint64_t)a->a_batch_stride : h * w * c;
    const cuuint64_t strides[3] = {c * 2, w * c * 2, img * 2};
    const cuuint32_t box[4] = {kBlockK, (cuuint32_t)a->tw, (cuuint32_t)a->th, (cuuint32_t)a->nb};
    int rc = encode_map(&tmap_a, a->a, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const cuuint64_t k = a->w_kdim, t = a->w_taps, n = a->n;
    const cuuint64_t dims[3] = {k, t, n};
    const cuuint64_t strides[2] = {k * 2, t * k * 2};
    const cuuint32_t box[3] = {kBlockK, 1, (cuuint32_t)b_rows};
    int rc = encode_map(&tmap_b, a->w, 3, dims, strides, box);
    if (rc) return rc;
  }

  p.fast_layout = a->out != nullptr && a->o_sn == 1 && a->res2 == nullptr && (a->res1 == nullptr || a->r1_sn == 1) &&
                  a->act_n_limit == 0 && (a->act == ACT_NONE || a->act == ACT_RELU) && a->n % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && a->o_sx % 8 == 0 && a->o_sy % 8 == 0 && a->o_sb % 8 == 0 &&
                  (a->res1 == nullptr || ((reinterpret_cast<uintptr_t>(a->res1) & 15) == 0 && a->r1_sx % 8 == 0 &&
                                          a->r1_sy % 8 == 0 && a->r1_sb % 8 == 0));
  p.stat_floats = a->stat_sum ? 2 * p.n_tiles * p.bn : 0;
  const size_t stat_bytes = sizeof(float) * p.stat_floats;
  TFPP_CHECK_ARG(stat_bytes <= 13 * 1024, "too many channels for the shared-memory statistics buffer");
  const bool need_tr = a->stat_sum != nullptr || !p.fast_layout;
  const size_t tr_bytes = (need_tr ? 8 * 32 * 33 * sizeof(float) : 0) + 256 * sizeof(float2);
  const size_t fixed_bytes = 1024 /*align*/ + 256 /*barriers*/ + tr_bytes + stat_bytes;
  int stages = static_cast<int>((227 * 1024 - fixed_bytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  TFPP_CHECK_ARG(stages >= 2, "shared memory budget exceeded");
  p.stages = stages;
  const size_t smem_bytes = static_cast<size_t>(stages) * stage_bytes + fixed_bytes;
  // TFPP_GEMM_LEAN: 0 = generic epilogue everywhere, 2 = statistics / residual instantiations only, default = all
  static const int lean_level = [] { const char* e = getenv("TFPP_GEMM_LEAN"); return e ? atoi(e) : 3; }();
  const bool lean_on = lean_level > 0;
  int lean = 0;
  if (lean_on && p.fast_layout && !a->out_f32 && a->scale == nullptr && a->shift == nullptr && a->act == ACT_NONE &&
      p.drop_rng == nullptr) {
    if (a->stat_sum != nullptr && a->res1 == nullptr) lean = 1;
    if (a->stat_sum == nullptr && (a->res1 == nullptr || !a->res1_f32)) lean = 2;
  }
  if (lean_level >= 3 && p.fast_layout && !a->out_f32 && a->scale == nullptr && a->shift != nullptr && p.drop_rng == nullptr &&
      a->stat_sum == nullptr && a->res1 == nullptr)
    lean = 3;   // bias (+ ReLU): fast_layout already restricts the activation to none / ReLU
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const KParams);
  static const KernelFn kernels[2][4] = {
      {conv_gemm_kernel<false, 0>, conv_gemm_kernel<false, 1>, conv_gemm_kernel<false, 2>, conv_gemm_kernel<false, 3>},
      {conv_gemm_kernel<true, 0>, conv_gemm_kernel<true, 1>, conv_gemm_kernel<true, 2>, conv_gemm_kernel<true, 3>}};
  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < 8; ++i) {
      cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernels[i / 4][i % 4]),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        tfpp_set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
        return TFPP_ERR_CUDA;
      }
    }
    attr_set = true;
  }
  const int num_tiles = p.n_tiles * p.m_tiles_x * p.m_tiles_y * p.m_tiles_b;
  int sms = TFPP_NUM_SMS;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    static int cached_sms = 0;
    if (cached_sms == 0) cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cached_sms > 0) sms = cached_sms;
  }
  if (pair) {
    const int items = p.n_tiles * ((m_tiles_total + 1) / 2);
    const int pairs = items < sms / 2 ? items : sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernels[1][lean], tmap_a, tmap_b, p);
    if (e != cudaSuccess) {
      tfpp_set_error("%s:%d: CUDA: %s", __FILE__, __LINE__, cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    return TFPP_OK;
  }
  const int grid = num_tiles < sms ? num_tiles : sms;
  kernels[0][lean]<<<grid, kThreads, smem_bytes, stream>>>(tmap_a, tmap_b, p);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
