// EXPERIMENTAL — not on the product path in round 1 (the engine uses it only with TFPP_HALO_UMMA=1).  Written at the
// very end of round 1: op-level parity against F.conv2d is green on B200 (tests/test_ops_gpu.py::test_halo_umma_conv3x3,
// profiles/r01_halo_umma_optest_v19.log) but it has not been timed or run inside the training step yet.
//
// Dense 3x3 convolution (stride 1, padding 1) with few channels at high resolution on tcgen05, WITHOUT re-reading the
// input nine times: the haloed-tile kernels of smallc_conv.cu / gconv3x3.cu are bound by the legacy mma.sync pipe
// (~540 TFLOP/s, a quarter of tcgen05), and the implicit-GEMM kernel of tc_gemm.cu by nine L2 reads of every tile.
//
// Idea: keep the haloed tile in shared memory in channel-chunk-major planes — plane c holds 8 channels (16 bytes) of
// every tile pixel, pixels linear with a row pitch of PW = 64:  plane[c][y * 64 + x] (16 B each).  In the K-major,
// NON-swizzled UMMA operand layout a core matrix is 8 rows x 16 B stored contiguously, i.e. exactly 8 consecutive
// pixels of one plane; SBO = 128 B walks on along the pixels, LBO = plane size walks to the next 8 channels.  So the
// A operand of tap (ky, kx) for 128 consecutive linear pixels is just a descriptor whose start address is shifted by
// (ky * 64 + kx) pixels: nine descriptors over ONE tile, no im2col, no copies.  Linear pixels include the two halo
// columns of each row (x = 62, 63): those outputs are junk and dropped (3 %).
// One 4-D TMA box {8 ch, 64, 10, 1} per plane writes precisely this layout (hardware zero fill = padding).
#include "../../include/tfpp.h"
#include "tc_common.cuh"

#include <cstdlib>

using namespace tc;

namespace {

constexpr int kThreadsH = 192;            // group conv kernel: warp 0: TMA, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int kLoadWarps = 4;             // dense kernel: cp.async producer warps behind the epilogue warps
constexpr int PW = 64;                    // plane row pitch in pixels
constexpr int TWO = PW - 2;               // output columns per tile (62)
constexpr int THO = 8;                    // output rows per tile
constexpr int PH = THO + 2;
constexpr int MBLK = THO * PW / 128;      // 4 UMMA M blocks of 128 linear pixels
constexpr int PLANE_BYTES = (PH * PW + 10) * 16;  // + slack: the last taps of the last block read 2 pixels past; 10400 B
                                                  // = 32 (mod 128): the four planes of a pixel fall into different banks
constexpr int kStagesH = 2;
constexpr int kTmemColsH = 512;

struct HParams {
  const bf16* x;        // (B,H,W,K) NHWC bf16 input
  const bf16* w;        // (9, K/8, N, 8) bf16: [tap][k chunk][n][8 k] — already in the shared-memory operand layout
  const float* bias;    // optional (n_valid)
  void* out;
  int out_nchw_f32;     // 0: NHWC bf16 with N channels; 1: NCHW f32 with n_valid channels
  int n_valid, act, act_n_limit;
  int B, H, W, K, N;    // K = input channels (16/32/64), N = padded output channels (16/32/48/64)
  int tiles_x, tiles_y;
  int stages;           // input tiles in flight (2..4): DRAM latency x bandwidth needs ~100 KB per SM in flight
};

// EPI = 4 (validated) or 8 epilogue warps (two per TMEM lane quarter, alternating M blocks; not yet run)
template <int EPI>
__global__ void __launch_bounds__(32 * MBLK + 32 * EPI + 32 * kLoadWarps, 1) halo_umma_conv3x3_kernel(const HParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kchunks = p.K / 8;
  const int stage_bytes = kchunks * PLANE_BYTES;
  uint8_t* wsm = smem + p.stages * stage_bytes;                          // [9][K/8][N][8] bf16
  const int w_bytes = 9 * p.K * p.N * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + ((w_bytes + 127) & ~127));
  uint64_t* full_bar = bars;                  // [4] tile landed
  uint64_t* empty_bar = bars + 4;             // [4] tile consumed by the MMAs
  uint64_t* tfull_bar = bars + 8;             // [2] accumulators ready
  uint64_t* tempty_bar = bars + 10;           // [2] accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* sbias = reinterpret_cast<float*>(bars + 16);   // [64] per-column bias (0 beyond n_valid / without bias)

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it (the
  // M block of an MMA-issuing warp, hence its descriptors) in uniform registers instead of electing lane by lane
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = tiles_per_img * p.B;
  if (threadIdx.x < 64) sbias[threadIdx.x] = (p.bias != nullptr && static_cast<int>(threadIdx.x) < p.n_valid) ? __ldg(p.bias + threadIdx.x) : 0.f;

  // weights -> shared memory (generic proxy), made visible to the async proxy (UMMA) before the first MMA
  for (int i = threadIdx.x; i < w_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(p.w) + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(smem_u32(&full_bar[s]), kLoadWarps);
      mbar_init(smem_u32(&empty_bar[s]), MBLK);     // one commit per MMA-issuing warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), MBLK);
      mbar_init(smem_u32(&tempty_bar[s]), EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemColsH)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= MBLK + EPI) {
    // ================================================================ producers: cp.async straight into the planes
    // Round 2: the first version wrote each 8-channel plane with its own TMA box {8 ch, 64, 10, 1}, i.e. 2560 16-byte
    // elements per tile, and the TMA engine's per-element cost (not math, not DRAM) bounded the kernel at 1.13 ms
    // (profiles/r02_ncu_halo_conv.txt: tensor pipe 6.8 %, DRAM 11 %).  A tile row is one contiguous 64 x K x 2 B run of
    // the NHWC input, so four warps copy it with coalesced 16-byte cp.async (zero fill outside the image = padding);
    // consecutive lanes take consecutive 16-byte chunks of a pixel and land in consecutive planes.
    const int pt = (warp - (MBLK + EPI)) * 32 + lane;
    const int n_chunks = kchunks * PH * PW;
    const int S = p.stages;
    int issued = 0, signalled = 0;
    auto publish = [&](int pending) {  // the oldest unpublished tile has landed once <= `pending` newer groups are in flight
      switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> UMMA (async proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&full_bar[signalled % S]));
      ++signalled;
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      const int y0 = (r / p.tiles_x) * THO, x0 = (r % p.tiles_x) * TWO;
      const int stage = issued % S;
      mbar_wait(smem_u32(&empty_bar[stage]), ((issued / S) & 1) ^ 1);
      const uint32_t st = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
      for (int i = pt; i < n_chunks; i += 32 * kLoadWarps) {
        const int c = i % kchunks, pix = i / kchunks;
        const int ty = pix / PW, tx = pix - ty * PW;
        const int gy = y0 - 1 + ty, gx = x0 - 1 + tx;
        const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        const bf16* src = ok ? p.x + ((static_cast<long long>(b) * p.H + gy) * p.W + gx) * p.K + c * 8 : p.x;
        const uint32_t dst = st + static_cast<uint32_t>(c) * PLANE_BYTES + static_cast<uint32_t>(pix) * 16;
        const int bytes = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      ++issued;
      if (issued - signalled == S) publish(S - 1);
    }
    while (signalled < issued) publish(issued - signalled - 1);
  } else if (warp < MBLK) {
    // ================================================================ MMA issuers: warp w owns the 128-pixel M block w.
    // With N = 32..64 a tcgen05.mma is 16-32 cycles of tensor work but ~150 cycles of dependent descriptor /
    // uniform-register instructions to issue (profiles/r02_ncu_halo_conv_v2.txt: one issuing thread was busy 74 % of the
    // kernel): four issuers, each with its own commits, bring the issue rate to the tensor pipe's.
    // The whole warp walks the loop (waits, descriptor arithmetic) so that the compiler keeps the descriptors in uniform
    // registers; only the tcgen05 instructions themselves are issued by lane 0.
    {
      const int blk = warp;
      const bool leader = elect_one();
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = make_idesc_bf16(128, p.N);
      const uint32_t w_u = smem_u32(wsm);
      const uint32_t w_lbo = static_cast<uint32_t>(p.N * 16);
      const uint64_t b0 = make_nosw_kmajor_desc(w_u, w_lbo, 128);
      const uint64_t w_tap_16 = static_cast<uint64_t>(p.K * p.N * 2) >> 4;   // weight block of one tap, in 16-byte units
      const uint64_t w_k_16 = static_cast<uint64_t>(2 * w_lbo) >> 4;         // 16 input channels further
      const int ksteps = p.K / 16;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
        // one thread issues all 72 MMAs of a tile: keep its dependent instruction chain short — the two descriptors are
        // built once per tile and only their 16-byte-granular start-address fields are advanced (no carry: all of shared
        // memory fits the 14-bit field)
        const uint64_t a0 = make_nosw_kmajor_desc(st, PLANE_BYTES, 128);
        {
          const uint32_t d_tmem = tmem_base + acc * 256 + blk * p.N;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint64_t a_tap = a0 + static_cast<uint64_t>(blk * 128 + ky * PW + kx);
            const uint64_t b_tap = b0 + static_cast<uint64_t>(tap) * w_tap_16;
            for (int kk = 0; kk < ksteps; ++kk)
              if (leader)
                umma_bf16(d_tmem, a_tap + static_cast<uint64_t>(kk) * (2 * (PLANE_BYTES >> 4)),
                          b_tap + static_cast<uint64_t>(kk) * w_k_16, idesc, (tap | kk) ? 1u : 0u);
          }
        }
        if (leader) {
          umma_commit(smem_u32(&empty_bar[stage]));    // the tile's planes may be overwritten
          umma_commit(smem_u32(&tfull_bar[acc]));      // the accumulators are complete
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ================================================================ epilogue: lane == linear tile pixel
    const int lane_group = warp & 3;
    const int blk0 = EPI == 8 ? ((warp - MBLK) >> 2) : 0;   // with 8 warps: even / odd M blocks
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      const int y0 = (r / p.tiles_x) * THO, x0 = (r % p.tiles_x) * TWO;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      for (int blk = blk0; blk < MBLK; blk += (EPI == 8 ? 2 : 1)) {
        const int m = blk * 128 + lane_group * 32 + lane;   // linear pixel of the tile
        const int ty = m / PW, tx = m - ty * PW;
        const int oy = y0 + ty, ox = x0 + tx;
        const bool valid = tx < TWO && oy < p.H && ox < p.W;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + acc * 256 + blk * p.N;
        // Round 2: this loop was the kernel's limiter (profiles/r02_ncu_halo_conv.txt: ~3100 instructions per warp and
        // tile with one warp per scheduler, i.e. every latency exposed; tensor pipe 6.8 %, DRAM 11 %).  Bias from a
        // shared-memory table, the activation decided once per 16-column chunk, no per-element branches.
        const int act_cols = p.act == ACT_NONE ? 0 : (p.act_n_limit == 0 ? p.N : p.act_n_limit);
        for (int c = 0; c < p.N; c += 16) {
          float v[16];
          __syncwarp();
          tmem_ld16(taddr + c, v);
          if (!valid) continue;
          const float4* bp = reinterpret_cast<const float4*>(sbias + c);
          const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          v[8] += b2.x; v[9] += b2.y; v[10] += b2.z; v[11] += b2.w; v[12] += b3.x; v[13] += b3.y; v[14] += b3.z; v[15] += b3.w;
          if (c + 16 <= act_cols && p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (c < act_cols) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c + j < act_cols) v[j] = apply_act(v[j], p.act);
          }
          if (p.out_nchw_f32) {
            float* o = static_cast<float*>(p.out);
            const long long hw = static_cast<long long>(p.H) * p.W;
            const long long base = static_cast<long long>(b) * p.n_valid * hw + static_cast<long long>(oy) * p.W + ox;
            const int jn = min(16, p.n_valid - c);
            for (int j = 0; j < jn; ++j) o[base + (c + j) * hw] = v[j];
          } else {
            bf16* o = static_cast<bf16*>(p.out) + ((static_cast<long long>(b) * p.H + oy) * p.W + ox) * p.N + c;
            uint4* o4 = reinterpret_cast<uint4*>(o);
            o4[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                               pack_bf16x2(v[6], v[7]));
            o4[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                               pack_bf16x2(v[14], v[15]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemColsH) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENTAL, NOT YET RUN ON A GPU — the RegNet group conv (width 24, stride 1) on the same plane layout.
// A 72-channel slab = 3 groups = 9 planes (+ 1 more plane so that the zero-padded K = 32 of the last group reads
// defined memory: channels 72..79 of the slab window, real data or TMA zero fill, multiplied by zero weight rows).
// Group g uses planes 3g..3g+3 (K = 32: 24 real channels + 8 that meet zero weights) and its own 9 x [32 k][32 n]
// weight block (24 real output channels); one CTA works on one slab (grid = CTAs per slab x slabs) so the 55 KB of
// weights are loaded once.  Tile = 4 rows x 62 columns (2 UMMA blocks of 128 linear pixels), double buffered.
// Epilogue: 24 of the 32 accumulator columns per group -> bf16 NHWC; BatchNorm batch statistics (sum, sum of squares
// over the valid pixels) through a shared-memory transpose, accumulated per CTA and flushed once.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GTHO = 4;
constexpr int GPH = GTHO + 2;
constexpr int GMBLK = GTHO * PW / 128;                    // 2
constexpr int GPLANE_BYTES = (GPH * PW + 8) * 16;         // 6272
constexpr int GPLANES = 10;
constexpr int GW_BYTES = 3 * 9 * 32 * 32 * 2;             // 55296: [group][tap][k chunk 4][n 32][8]

struct HGParams {
  const bf16* w;         // (C/24, 9, 4, 32, 8) bf16: [group][tap][k chunk][n][8 k], zero padded from 24 to 32 both ways
  bf16* out;             // (B,H,W,C)
  const float* scale;    // optional folded BatchNorm (eval)
  const float* shift;
  int act;
  float* stat_sum;       // optional BatchNorm batch statistics (C each)
  float* stat_sq;
  int B, H, W, C;
  int tiles_x, tiles_y;
};

__global__ void __launch_bounds__(kThreadsH, 1) halo_umma_gconv3x3_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                          const HGParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int stage_bytes = GPLANES * GPLANE_BYTES;                   // 62720
  uint8_t* wsm = smem + kStagesH * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + GW_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 2;
  uint64_t* tfull_bar = bars + 4;
  uint64_t* tempty_bar = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* sstat = reinterpret_cast<float*>(bars + 10);                   // [2][72] per-CTA statistics
  float* trbuf = sstat + 2 * 72;                                        // 4 x [32][33] transpose tiles

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.y;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = tiles_per_img * p.B;

  for (int i = threadIdx.x; i < GW_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(p.w) + static_cast<long long>(slab) * (GW_BYTES / 16) + i);
  for (int i = threadIdx.x; i < 2 * 72; i += blockDim.x) sstat[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < kStagesH; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemColsH)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int r = tile - b * tiles_per_img;
        const int y0 = (r / p.tiles_x) * GTHO, x0 = (r % p.tiles_x) * TWO;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, static_cast<uint32_t>(GPLANES * GPH * PW * 16));
        uint8_t* st = smem + static_cast<size_t>(stage) * stage_bytes;
        for (int c = 0; c < GPLANES; ++c)
          tma_load_4d(smem_u32(st + c * GPLANE_BYTES), &tmap_x, fb, slab * 72 + c * 8, x0 - 1, y0 - 1, b);
        if (++stage == kStagesH) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elected lane: descriptors stay in uniform registers (see the dense kernel)
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = make_idesc_bf16(128, 32);
      const uint32_t w_u = smem_u32(wsm);
      const uint64_t b0 = make_nosw_kmajor_desc(w_u, 32 * 16, 128);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
        const uint64_t a0 = make_nosw_kmajor_desc(st, GPLANE_BYTES, 128);   // start-address fields advance in 16-byte units
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
          for (int blk = 0; blk < GMBLK; ++blk) {
            const uint32_t d_tmem = tmem_base + acc * 256 + (g * GMBLK + blk) * 32;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int ky = tap / 3, kx = tap - ky * 3;
              const uint64_t a_tap = a0 + static_cast<uint64_t>((g * 3) * (GPLANE_BYTES >> 4) + blk * 128 + ky * PW + kx);
              const uint64_t b_tap = b0 + static_cast<uint64_t>((g * 9 + tap) * ((32 * 32 * 2) >> 4));
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                umma_bf16(d_tmem, a_tap + static_cast<uint64_t>(kk * 2 * (GPLANE_BYTES >> 4)),
                          b_tap + static_cast<uint64_t>(kk * 2 * ((32 * 16) >> 4)), idesc, (tap | kk) ? 1u : 0u);
            }
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        umma_commit(smem_u32(&tfull_bar[acc]));
        if (++stage == kStagesH) {
          stage = 0;
          phase ^= 1;
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int lane_group = warp & 3;
    float* tr = trbuf + (warp - 2) * (32 * 33);
    const bool has_stats = p.stat_sum != nullptr;
    const bool has_affine = p.scale != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      const int y0 = (r / p.tiles_x) * GTHO, x0 = (r % p.tiles_x) * TWO;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      for (int blk = 0; blk < GMBLK; ++blk) {
        const int m = blk * 128 + lane_group * 32 + lane;
        const int ty = m / PW, tx = m - ty * PW;
        const int oy = y0 + ty, ox = x0 + tx;
        const bool valid = tx < TWO && oy < p.H && ox < p.W;
        const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
        for (int g = 0; g < 3; ++g) {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + acc * 256 +
                                 (g * GMBLK + blk) * 32;
          uint32_t rr[32];
          __syncwarp();
          tmem_ld32_issue(taddr, rr);
          tmem_ld_wait32(rr);
          const int c0 = slab * 72 + g * 24;
          if (has_stats) {
#pragma unroll
            for (int j = 0; j < 24; ++j) tr[lane * 33 + j] = __uint_as_float(rr[j]);
            __syncwarp();
            if (lane < 24) {
              float ssum = 0.f, ssq = 0.f;
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                const float a = ((valid_mask >> q) & 1u) ? tr[q * 33 + lane] : 0.f;
                ssum += a;
                ssq = fmaf(a, a, ssq);
              }
              atomicAdd(&sstat[g * 24 + lane], ssum);
              atomicAdd(&sstat[72 + g * 24 + lane], ssq);
            }
            __syncwarp();
          }
          if (valid) {
            float v[24];
#pragma unroll
            for (int j = 0; j < 24; ++j) {
              float x = __uint_as_float(rr[j]);
              if (has_affine) x = fmaf(x, __ldg(p.scale + c0 + j), __ldg(p.shift + c0 + j));
              if (p.act == ACT_RELU) x = fmaxf(x, 0.f);
              v[j] = x;
            }
            uint4* o4 = reinterpret_cast<uint4*>(p.out + ((static_cast<long long>(b) * p.H + oy) * p.W + ox) * p.C + c0);
#pragma unroll
            for (int q = 0; q < 3; ++q)
              o4[q] = make_uint4(pack_bf16x2(v[q * 8], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                                 pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (has_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = (warp - 2) * 32 + lane;
      for (int i = et; i < 72; i += 128) {
        atomicAdd(p.stat_sum + slab * 72 + i, sstat[i]);
        atomicAdd(p.stat_sq + slab * 72 + i, sstat[72 + i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemColsH) : "memory");
  }
}

}  // namespace

extern "C" int tfpp_halo_gconv3x3(const void* x, const void* w, void* out, const float* scale, const float* shift, int act,
                                  float* stat_sum, float* stat_sq, int batch, int height, int width, int channels,
                                  tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(channels % 72 == 0, "channels must be a multiple of 72 (3 groups of width 24)");
  TFPP_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  TFPP_CHECK_ARG((stat_sum == nullptr) == (stat_sq == nullptr), "stat_sum and stat_sq go together");
  TFPP_CHECK_ARG(act == ACT_NONE || act == ACT_RELU, "activation: none or relu");
  HGParams p;
  p.w = static_cast<const bf16*>(w); p.out = static_cast<bf16*>(out); p.scale = scale; p.shift = shift; p.act = act;
  p.stat_sum = stat_sum; p.stat_sq = stat_sq;
  p.B = batch; p.H = height; p.W = width; p.C = channels;
  p.tiles_x = ceil_div(width, TWO);
  p.tiles_y = ceil_div(height, GTHO);
  CUtensorMap tmap;
  {
    const cuuint64_t c = channels, w_ = width, h = height, b = batch;
    const cuuint64_t dims[4] = {c, w_, h, b};
    const cuuint64_t strides[3] = {c * 2, w_ * c * 2, h * w_ * c * 2};
    const cuuint32_t box[4] = {8, PW, GPH, 1};
    int rc = encode_map(&tmap, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  const size_t smem = 1024 + static_cast<size_t>(kStagesH) * GPLANES * GPLANE_BYTES + GW_BYTES + 80 + sizeof(float) * (2 * 72 + 4 * 32 * 33) + 64;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(halo_umma_gconv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr = true;
  }
  const long long tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * batch;
  if (tiles == 0) return TFPP_OK;
  const int slabs = channels / 72;
  long long per_slab = TFPP_NUM_SMS / slabs;
  if (per_slab < 1) per_slab = 1;
  if (per_slab > tiles) per_slab = tiles;
  dim3 grid(static_cast<unsigned>(per_slab), slabs);
  halo_umma_gconv3x3_kernel<<<grid, kThreadsH, smem, stream>>>(tmap, p);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_halo_conv3x3(const void* x, const void* w, const float* bias, void* out, int out_nchw_f32,
                                 int n_valid, int act, int act_n_limit, int batch, int height, int width, int cin,
                                 int cout_padded, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(cin == 16 || cin == 32 || cin == 64, "cin must be 16, 32 or 64");
  TFPP_CHECK_ARG(cout_padded % 16 == 0 && cout_padded >= 16 && cout_padded <= 64, "cout_padded must be 16..64, % 16");
  TFPP_CHECK_ARG(n_valid >= 1 && n_valid <= cout_padded, "n_valid <= cout_padded");
  TFPP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0, "operands must be 16-byte aligned");
  HParams p;
  p.x = static_cast<const bf16*>(x);
  p.w = static_cast<const bf16*>(w); p.bias = bias; p.out = out; p.out_nchw_f32 = out_nchw_f32;
  p.n_valid = n_valid; p.act = act; p.act_n_limit = act_n_limit;
  p.B = batch; p.H = height; p.W = width; p.K = cin; p.N = cout_padded;
  p.tiles_x = ceil_div(width, TWO);
  p.tiles_y = ceil_div(height, THO);
  const size_t w_smem = ((9 * cin * cout_padded * 2 + 127) & ~127) + 512;
  p.stages = 4;
  while (p.stages > 2 && 1024 + static_cast<size_t>(p.stages) * (cin / 8) * PLANE_BYTES + w_smem > 227 * 1024) --p.stages;
  const size_t smem = 1024 + static_cast<size_t>(p.stages) * (cin / 8) * PLANE_BYTES + w_smem;
  TFPP_CHECK_ARG(smem <= 227 * 1024, "shared memory budget exceeded");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(halo_umma_conv3x3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(halo_umma_conv3x3_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr = true;
  }
  const long long tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * batch;
  if (tiles == 0) return TFPP_OK;
  const int grid = static_cast<int>(tiles < TFPP_NUM_SMS ? tiles : TFPP_NUM_SMS);
  static const bool epi8 = [] { const char* e = getenv("TFPP_HALO_UMMA_EPI8"); return e != nullptr && e[0] == '1'; }();
  if (epi8) halo_umma_conv3x3_kernel<8><<<grid, 32 * MBLK + 32 * 8 + 32 * kLoadWarps, smem, stream>>>(p);
  else halo_umma_conv3x3_kernel<4><<<grid, 32 * MBLK + 32 * 4 + 32 * kLoadWarps, smem, stream>>>(p);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
