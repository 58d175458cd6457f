// tcgen05 weight-gradient GEMM:  dW[co, tap, ci] += sum_pixels dY[pixel, co] * X[pixel + shift(tap), ci].
//
// The contraction runs over pixels, which are the strided dimension of both NHWC operands, so both UMMA operands are
// MN-major: a TMA box {64 channels, 64 pixels} lands in shared memory as [pixel][64 ch] 128-byte swizzled rows, which
// is exactly the canonical MN-major SWIZZLE_128B layout (8-pixel K groups 1024 B apart, 64-channel MN blocks LBO
// apart).  One CTA owns one (co tile of 128, ci tile of BN, tap, pixel-range split): TMA producer warp, MMA warp,
// 4 epilogue warps that atomically add the fp32 TMEM accumulator into dW (split-K reduction in L2).
//
// Replaces autograd's cudnn_convolution_backward_weight / addmm for every conv / linear of the TransFuser++ step
// (team_code/train.py:898 loss.backward()).
#include "../../include/tfpp.h"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int kThreads = 192;
constexpr int kPix = 64;                        // pixels (K) per pipeline stage
constexpr int kRegion = kPix * 128;             // bytes of one [64 pixels][64 ch] block = 8 KB
constexpr int kMaxStages = 6;
constexpr int kTmemCols = 256;

struct WParams {
  int batch, height, width;     // extents of dY (pixel space the contraction runs over)
  int tw, th, nb;               // pixel tile, tw*th*nb == 64
  int p_tiles_x, p_tiles_y, p_tiles_b, p_tiles;
  int splits;                   // pixel-range splits (grid.z)
  int m_tiles, n_tiles, ntaps;
  int m_stride;                 // output channels advanced per m tile (128 dense, 96 grouped)
  int m_valid;                  // rows of the tile that are real outputs
  int bn;                       // ci tile width (multiple of 16, <= 256)
  int n_blocks;                 // 64-channel TMA boxes per B stage = ceil(bn / 64)
  int grouped;                  // 1: ci window follows the m tile (block-diagonal 24-wide groups)
  int group_width;
  int cout, cin;                // true channel counts (masking)
  int tap_dx[9], tap_dy[9], tap_db[9], tap_w[9];
  int stages;
  float* dw;                    // fp32; element (co, tap, ci) at co*s_co + tap*s_tap + ci*s_ci
  long long s_co, s_tap, s_ci;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = (2 + p.n_blocks) * kRegion;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile decode: blockIdx.x = (m_tile, n_tile), blockIdx.y = tap, blockIdx.z = split
  const int m_tile = blockIdx.x / p.n_tiles, n_tile = blockIdx.x % p.n_tiles;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int co0 = m_tile * p.m_stride;
  const int ci0 = p.grouped ? co0 : n_tile * p.bn;
  const int pt0 = static_cast<int>(static_cast<long long>(p.p_tiles) * split / p.splits);
  const int pt1 = static_cast<int>(static_cast<long long>(p.p_tiles) * (split + 1) / p.splits);
  const int k_iters = pt1 - pt0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tfull_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (k_iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = static_cast<uint32_t>(stage_bytes);
        for (int pt = pt0; pt < pt1; ++pt) {
          int m = pt;
          const int x0 = (m % p.p_tiles_x) * p.tw;
          m /= p.p_tiles_x;
          const int y0 = (m % p.p_tiles_y) * p.th;
          const int b0 = (m / p.p_tiles_y) * p.nb;
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, tx_bytes);
          uint8_t* st = smem + static_cast<size_t>(stage) * stage_bytes;
          tma_load_4d(smem_u32(st), &tmap_dy, fb, co0, x0, y0, b0);
          tma_load_4d(smem_u32(st + kRegion), &tmap_dy, fb, co0 + 64, x0, y0, b0);
          for (int nbk = 0; nbk < p.n_blocks; ++nbk)
            tma_load_4d(smem_u32(st + (2 + nbk) * kRegion), &tmap_x, fb, ci0 + nbk * 64, x0 + p.tap_dx[tap],
                        y0 + p.tap_dy[tap], b0 + p.tap_db[tap]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t idesc = make_idesc_bf16_mn(128, p.bn);
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + 2 * kRegion;
#pragma unroll
          for (int kk = 0; kk < kPix / 16; ++kk) {
            // 16 pixels = two 8-row K groups = 2048 bytes further into every region
            const uint64_t adesc = make_sw128_mnmajor_desc(sa + kk * 2048, kRegion);
            const uint64_t bdesc = make_sw128_mnmajor_desc(sb + kk * 2048, kRegion);
            umma_bf16(tmem_base, adesc, bdesc, idesc, (k | kk) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(smem_u32(tfull_bar));
      }
    } else {
      const int lane_group = warp & 3;
      const int row = lane_group * 32 + lane;  // output channel within the tile
      const int co = co0 + row;
      mbar_wait(smem_u32(tfull_bar), 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
      const bool row_ok = row < p.m_valid && co < p.cout;
      const int wt = p.tap_w[tap];
      const bool vec_ok = !p.grouped && p.s_ci == 1 && (p.s_co & 3) == 0 && ((wt * p.s_tap) & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;
      for (int c = 0; c < p.bn; c += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(taddr + c, v);
        if (!row_ok) continue;
        if (p.grouped) {
          const int g0 = (row / p.group_width) * p.group_width;  // first ci (window-relative) of this row's group
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int cw = c + j;
            if (cw >= g0 && cw < g0 + p.group_width)
              atomicAdd(p.dw + co * p.s_co + wt * p.s_tap + (cw - g0) * p.s_ci, v[j]);
          }
        } else if (vec_ok && ci0 + c + 16 <= p.cin) {
          // contiguous input-channel axis (1x1 convs, linears): 16 floats of this row as four 16-byte reductions —
          // the L2 atomic units are the bottleneck of the split-K epilogue, this issues a quarter of the operations
          float* dst = p.dw + co * p.s_co + wt * p.s_tap + (ci0 + c);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                         "f"(v[j + 2]), "f"(v[j + 3])
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int ci = ci0 + c + j;
            if (ci < p.cin) atomicAdd(p.dw + co * p.s_co + wt * p.s_tap + ci * p.s_ci, v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace

extern "C" int tfpp_conv_wgrad(const tfpp_wgrad_args* a, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(a != nullptr && a->dy != nullptr && a->x != nullptr && a->dw != nullptr, "null operand");
  TFPP_CHECK_ARG(a->cout % 8 == 0 && a->x_channels % 8 == 0, "channel counts must be multiples of 8");
  TFPP_CHECK_ARG(a->tw * a->th * a->nb == kPix, "tw*th*nb must be 64");
  TFPP_CHECK_ARG(a->ntaps >= 1 && a->ntaps <= 9, "1..9 taps");
  WParams p;
  p.batch = a->batch; p.height = a->height; p.width = a->width;
  p.tw = a->tw; p.th = a->th; p.nb = a->nb;
  p.p_tiles_x = ceil_div(a->width, a->tw);
  p.p_tiles_y = ceil_div(a->height, a->th);
  p.p_tiles_b = ceil_div(a->batch, a->nb);
  p.p_tiles = p.p_tiles_x * p.p_tiles_y * p.p_tiles_b;
  p.grouped = a->group_width > 0;
  p.group_width = a->group_width;
  p.cout = a->cout_valid > 0 ? a->cout_valid : a->cout;
  p.cin = a->cin;
  p.ntaps = a->ntaps;
  p.s_co = a->dw_s_co; p.s_tap = a->dw_s_tap; p.s_ci = a->dw_s_ci;
  for (int i = 0; i < 9; ++i) {
    p.tap_dx[i] = a->tap_dx[i]; p.tap_dy[i] = a->tap_dy[i]; p.tap_db[i] = a->tap_db[i]; p.tap_w[i] = a->tap_w[i];
  }
  if (p.grouped) {
    TFPP_CHECK_ARG(96 % a->group_width == 0, "group width must divide 96");
    p.m_stride = 96; p.m_valid = 96; p.bn = 96; p.n_tiles = 1;
  } else {
    p.m_stride = 128; p.m_valid = 128;
    int bn = a->bn > 0 ? a->bn : 256;
    if (bn > 256) bn = 256;
    const int cin16 = ((a->cin + 15) / 16) * 16;
    if (bn > cin16) bn = cin16;
    TFPP_CHECK_ARG(bn % 16 == 0 && bn >= 16, "bn must be a multiple of 16");
    p.bn = bn;
    p.n_tiles = ceil_div(a->cin, bn);
  }
  p.m_tiles = ceil_div(a->cout, p.m_stride);
  p.n_blocks = ceil_div(p.bn, 64);
  const int stage_bytes = (2 + p.n_blocks) * kRegion;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  p.dw = a->dw;
  const int tiles = p.m_tiles * p.n_tiles * p.ntaps;
  int splits = a->splits;
  if (splits <= 0) {
    // one CTA per SM (200 KB of stages): pick the split count that minimises waves x (pixel blocks per CTA + fixed
    // prologue/epilogue cost) instead of a fixed 2 CTAs/SM target — 300 CTAs on 148 SMs is three waves, not two
    const int fixed = 6 + p.bn / 32;   // prologue + TMEM drain + reductions, in units of one 64-pixel k block
    long long best_cost = -1;
    splits = 1;
    const int smax = p.p_tiles < 2 * TFPP_NUM_SMS ? p.p_tiles : 2 * TFPP_NUM_SMS;
    for (int sp = 1; sp <= smax; ++sp) {
      const long long waves = ceil_div_ll(static_cast<long long>(tiles) * sp, TFPP_NUM_SMS);
      const long long cost = waves * (ceil_div(p.p_tiles, sp) + fixed);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        splits = sp;
      }
    }
  }
  if (splits > p.p_tiles) splits = p.p_tiles;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.splits = splits;

  CUtensorMap tmap_dy, tmap_x;
  {
    const cuuint64_t c = a->cout, w = a->width, h = a->height, b = a->batch;
    const cuuint64_t dims[4] = {c, w, h, b};
    const cuuint64_t strides[3] = {c * 2, w * c * 2, h * w * c * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)a->tw, (cuuint32_t)a->th, (cuuint32_t)a->nb};
    int rc = encode_map(&tmap_dy, a->dy, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const cuuint64_t c = a->x_channels, w = a->width, h = a->height, b = a->x_batch;
    const cuuint64_t img = a->x_batch_stride > 0 ? (cuuint64_t)a->x_batch_stride : h * w * c;
    const cuuint64_t dims[4] = {c, w, h, b};
    const cuuint64_t strides[3] = {c * 2, w * c * 2, img * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)a->tw, (cuuint32_t)a->th, (cuuint32_t)a->nb};
    int rc = encode_map(&tmap_x, a->x, 4, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem_bytes = static_cast<size_t>(stages) * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(p.m_tiles * p.n_tiles, p.ntaps, splits);
  wgrad_kernel<<<grid, kThreads, smem_bytes, stream>>>(tmap_dy, tmap_x, p);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
