// Self-attention core of the TransFuser fusion GPT (transfuser.py:367-376) — forward:
//   per (batch, head): out = softmax(Q K^T / sqrt(hd)) V over T <= 320 tokens (8x32 image + 8x8 LiDAR anchors).
// qkv is the fused projection output (B, T, 3C) bf16 laid out [q | k | v]; head h owns channels [h*hd, (h+1)*hd) with
// hd = C / 4 = 18, 54, 144, 378: only 4-byte aligned, so global->shared staging uses 4-byte cp.async (zero fill for
// the ragged last chunk) into 16-byte aligned tiles, and everything after that is ldmatrix + mma.sync.
//
// One CTA = 64 queries of one head (grid 5 x heads x B), 8 warps, 2 CTAs per SM.  The head dimension streams through
// double-buffered 32-wide chunks: first the K chunks (S = Q K^T accumulates in registers, 16 rows x 160 keys per
// warp), then softmax in registers (row max / sum exchanged between the two key halves through shared memory), the
// probabilities are parked once as bf16 in shared memory, then the V chunks (O chunk = P V, ldmatrix.trans on the
// row-major V tile, written straight to global).  Scores never leave the SM.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

constexpr int kT = 320;            // max tokens
constexpr int kQ = 64;             // queries per CTA
constexpr int kDC = 32;            // head-dim chunk
constexpr int kPitchC = kDC + 8;   // bf16 pitch of chunk tiles (80 B rows: conflict-free ldmatrix)
constexpr int kPitchP = kT + 8;    // bf16 pitch of the probability tile (656 B rows)
constexpr int kNt = (kT / 2) / 8;  // n8 tiles per warp in S: 20
constexpr int kStageElems = (kT + kQ) * kPitchC;   // K/V chunk [320][40] + Q chunk [64][40]

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void cp_async4_zfill(uint32_t saddr, const void* gmem, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(gmem), "r"(sz) : "memory");
}

// rows [r_begin, r_begin + nrows) x columns [d0, d0 + 32) of a row-strided bf16 matrix -> tile[r][40], zero outside
// [0, T) x [0, hd).  16 words per row: half-warps copy 64 contiguous bytes.
__device__ __forceinline__ void stage_chunk(uint32_t tile, const bf16* src, long long row_stride, int r_begin, int nrows,
                                            int T, int d0, int hd) {
  for (int i = threadIdx.x; i < nrows * (kDC / 2); i += blockDim.x) {
    const int r = i >> 4, w = i & 15;
    const int d = d0 + 2 * w;
    const bool ok = (r_begin + r < T) && (d < hd);
    const bf16* g = src + (ok ? (r_begin + r) * row_stride + d : 0);
    cp_async4_zfill(tile + (r * kPitchC + 2 * w) * 2, g, ok);
  }
}

// keep flags (bit 0 / bit 1) of the probability pair (row q, keys col, col + 1) of head (b, h); col is even
__device__ __forceinline__ uint32_t drop_pair(const DropCtx& d, int b, int heads, int h, int T, int q, int col) {
  const unsigned long long idx = ((static_cast<unsigned long long>(b) * heads + h) * T + q) * T + col;
  const uint4 w = drop_words(d, idx >> 2);
  const uint32_t w0 = (idx & 2) ? w.z : w.x, w1 = (idx & 2) ? w.w : w.y;
  return (w0 >= d.thresh ? 1u : 0u) | (w1 >= d.thresh ? 2u : 0u);
}

template <bool DROP>
__global__ void __launch_bounds__(256, 2) fusion_attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                 int T, int C, int heads, float scale,
                                                                 const unsigned long long* drop_rng, float drop_p,
                                                                 unsigned drop_site) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  bf16* P = reinterpret_cast<bf16*>(att_smem);                 // [64][kPitchP]
  bf16* stage0 = P + kQ * kPitchP;                             // [2][kStageElems]
  float* scratch = reinterpret_cast<float*>(stage0 + 2 * kStageElems);   // [2][64]
  const int hd = C / heads;
  const int q0 = blockIdx.x * kQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const long long rs3 = 3ll * C;
  const bf16* base = qkv + static_cast<long long>(b) * T * rs3 + h * hd;
  const int nc = (hd + kDC - 1) / kDC;
  const int rb = warp & 3, kh = warp >> 2;
  const int keys_half = T / 2;          // T % 32 == 0: a multiple of 16
  const int npairs = keys_half / 16;    // pairs of n8 tiles per warp
  const int r0 = rb * 16 + g;

  auto issue = [&](int s, int buf) {
    const uint32_t tile = smem_u32(stage0 + buf * kStageElems);
    if (s < nc) {
      stage_chunk(tile, base + C, rs3, 0, T, T, s * kDC, hd);                                   // K chunk
      stage_chunk(tile + kT * kPitchC * 2, base, rs3, q0, kQ, T, s * kDC, hd);                  // Q chunk
    } else {
      stage_chunk(tile, base + 2 * C, rs3, 0, T, T, (s - nc) * kDC, hd);                        // V chunk
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[kNt][4];
#pragma unroll
  for (int i = 0; i < kNt; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  issue(0, 0);
  for (int s = 0; s < 2 * nc; ++s) {
    const int buf = s & 1;
    if (s + 1 < 2 * nc) {
      issue(s + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t kv = smem_u32(stage0 + buf * kStageElems);
    if (s < nc) {
      // ---- S += Q_chunk K_chunk^T
      const uint32_t qs = kv + kT * kPitchC * 2;
#pragma unroll
      for (int k0 = 0; k0 < kDC; k0 += 16) {
        uint32_t a[4];
        ldsm_x4(a, qs + ((rb * 16 + (lane & 15)) * kPitchC + k0 + (lane >> 4) * 8) * 2);
        const int mat = lane >> 3;
        const uint32_t brow = kv + ((kh * keys_half + (mat >> 1) * 8 + (lane & 7)) * kPitchC + k0 + (mat & 1) * 8) * 2;
#pragma unroll
        for (int np = 0; np < kNt / 2; ++np) {
          if (np < npairs) {
            uint32_t bq[4];
            ldsm_x4(bq, brow + np * 16 * kPitchC * 2);
            mma16816(acc[2 * np], a, bq[0], bq[1]);
            mma16816(acc[2 * np + 1], a, bq[2], bq[3]);
          }
        }
      }
      if (s == nc - 1) {
        // ---- softmax over the keys, in registers (F.softmax(dim=-1), transfuser.py:373)
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < kNt; ++nt) {
          if (nt < 2 * npairs) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] *= scale;
            m0 = fmaxf(m0, fmaxf(acc[nt][0], acc[nt][1]));
            m1 = fmaxf(m1, fmaxf(acc[nt][2], acc[nt][3]));
          }
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        if (t4 == 0) {
          scratch[kh * kQ + r0] = m0;
          scratch[kh * kQ + r0 + 8] = m1;
        }
        __syncthreads();
        m0 = fmaxf(scratch[r0], scratch[kQ + r0]);
        m1 = fmaxf(scratch[r0 + 8], scratch[kQ + r0 + 8]);
        __syncthreads();
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < kNt; ++nt) {
          if (nt < 2 * npairs) {
            acc[nt][0] = __expf(acc[nt][0] - m0);
            acc[nt][1] = __expf(acc[nt][1] - m0);
            acc[nt][2] = __expf(acc[nt][2] - m1);
            acc[nt][3] = __expf(acc[nt][3] - m1);
            s0 += acc[nt][0] + acc[nt][1];
            s1 += acc[nt][2] + acc[nt][3];
          }
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        if (t4 == 0) {
          scratch[kh * kQ + r0] = s0;
          scratch[kh * kQ + r0 + 8] = s1;
        }
        __syncthreads();
        const float i0 = 1.f / (scratch[r0] + scratch[kQ + r0]);
        const float i1 = 1.f / (scratch[r0 + 8] + scratch[kQ + r0 + 8]);
#pragma unroll
        for (int nt = 0; nt < kNt; ++nt) {
          if (nt < 2 * npairs) {
            const int col = kh * keys_half + nt * 8 + 2 * t4;
            float p00 = acc[nt][0] * i0, p01 = acc[nt][1] * i0, p10 = acc[nt][2] * i1, p11 = acc[nt][3] * i1;
            if (DROP) {  // attn_drop (transfuser.py:374): dropped probabilities leave the P V product; 1/(1-p) is applied to O
              const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);   // built here: no registers held over the main loop
              const uint32_t k0 = drop_pair(drop, b, heads, h, T, q0 + r0, col);
              const uint32_t k1 = drop_pair(drop, b, heads, h, T, q0 + r0 + 8, col);
              p00 = (k0 & 1u) ? p00 : 0.f;
              p01 = (k0 & 2u) ? p01 : 0.f;
              p10 = (k1 & 1u) ? p10 : 0.f;
              p11 = (k1 & 2u) ? p11 : 0.f;
            }
            *reinterpret_cast<uint32_t*>(P + r0 * kPitchP + col) = pack_bf16x2(p00, p01);
            *reinterpret_cast<uint32_t*>(P + (r0 + 8) * kPitchP + col) = pack_bf16x2(p10, p11);
          }
        }
      }
    } else {
      // ---- O[:, chunk] = P V_chunk: warp -> 16 rows x 16 of the chunk's 32 columns
      const int d0 = (s - nc) * kDC;
      float o[2][4];
      o[0][0] = o[0][1] = o[0][2] = o[0][3] = o[1][0] = o[1][1] = o[1][2] = o[1][3] = 0.f;
      const uint32_t prow = smem_u32(P) + ((rb * 16 + (lane & 15)) * kPitchP + (lane >> 4) * 8) * 2;
      const int mat = lane >> 3;
      const uint32_t vrow = kv + (((mat & 1) * 8 + (lane & 7)) * kPitchC + kh * 16 + (mat >> 1) * 8) * 2;
      for (int k0 = 0; k0 < T; k0 += 16) {
        uint32_t a[4], bq[4];
        ldsm_x4(a, prow + k0 * 2);
        ldsm_x4_t(bq, vrow + k0 * kPitchC * 2);
        mma16816(o[0], a, bq[0], bq[1]);
        mma16816(o[1], a, bq[2], bq[3]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = d0 + kh * 16 + nt * 8 + 2 * t4;
        if (col < hd) {  // hd is even: col + 1 < hd as well
          const int qr = q0 + r0;
          bf16* op = out + (static_cast<long long>(b) * T + qr) * C + h * hd + col;
          const float ik = DROP ? 1.f / (1.f - drop_p) : 1.f;
          if (qr < T) *reinterpret_cast<uint32_t*>(op) = pack_bf16x2(o[nt][0] * ik, o[nt][1] * ik);
          if (qr + 8 < T) *reinterpret_cast<uint32_t*>(op + 8ll * C) = pack_bf16x2(o[nt][2] * ik, o[nt][3] * ik);
        }
      }
    }
    __syncthreads();   // everyone is done with `buf` (and P is complete) before the next prefetch overwrites it
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  Same CTA decomposition (64 queries of one head).  With P = softmax(Q K^T * scale):
//   dV += P^T dO;  dP = dO V^T;  dS = P * (dP - rowsum(dP * P)) * scale;  dQ = dS K;  dK += dS^T Q.
// Four streaming phases over the 32-wide head-dim chunks, all double buffered:
//   A  K, Q chunks   -> S in registers -> softmax -> P (bf16, shared)
//   V  dO chunks     -> dV[:, chunk] += P^T dO_chunk            (operands via ldmatrix.trans: no transposed copies)
//   B  V, dO chunks  -> dP in registers -> dS overwrites P
//   C  K, Q chunks   -> dQ[:, chunk] = dS K_chunk (stored bf16);  dK[:, chunk] += dS^T Q_chunk
// dK / dV of a head receive contributions from all 5 query tiles: fp32 vector reductions (red.global.add.v2.f32) into
// a (B,T,2C) workspace, converted to bf16 by a second small kernel.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <bool DROP>
__global__ void __launch_bounds__(256, 2) fusion_attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                 bf16* __restrict__ dqkv, float* __restrict__ dkv, int T,
                                                                 int C, int heads, float scale,
                                                                 const unsigned long long* drop_rng, float drop_p,
                                                                 unsigned drop_site) {
  // Dropout on the probabilities: P is parked with the SIGN BIT as the "dropped" flag (P >= 0), so that both the
  // undropped softmax (needed for dS) and the mask survive in the one shared-memory tile:
  //   dV += (P o M)^T dO / (1-p);  dP' = (dO V^T) o M / (1-p);  dS = P o (dP' - rowsum(dP' o P)) * scale.
  const float inv_keep = DROP ? 1.f / (1.f - drop_p) : 1.f;
  extern __shared__ __align__(128) uint8_t att_smem[];
  bf16* P = reinterpret_cast<bf16*>(att_smem);                 // [64][kPitchP]: P, later dS
  bf16* stage0 = P + kQ * kPitchP;                             // [2][kStageElems]: big tile [320][40] + small tile [64][40]
  float* scratch = reinterpret_cast<float*>(stage0 + 2 * kStageElems);   // [2][64]
  const int hd = C / heads;
  const int q0 = blockIdx.x * kQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const long long rs3 = 3ll * C;
  const bf16* base = qkv + static_cast<long long>(b) * T * rs3 + h * hd;
  const bf16* dob = dout + static_cast<long long>(b) * T * C + h * hd;
  const int nc = (hd + kDC - 1) / kDC;
  const int rb = warp & 3, kh = warp >> 2;
  const int keys_half = T / 2;
  const int npairs = keys_half / 16;
  const int r0 = rb * 16 + g;
  const int mat = lane >> 3;
  const uint32_t p_u = smem_u32(P);

  // stage s of the 4*nc-long sequence: phase = s / nc (A, V, B, C), chunk = s % nc
  auto issue = [&](int s, int buf) {
    const uint32_t big = smem_u32(stage0 + buf * kStageElems);
    const uint32_t small = big + kT * kPitchC * 2;
    const int phase = s / nc, d0 = (s - phase * nc) * kDC;
    if (phase == 0 || phase == 3) {
      stage_chunk(big, base + C, rs3, 0, T, T, d0, hd);          // K chunk
      stage_chunk(small, base, rs3, q0, kQ, T, d0, hd);          // Q chunk
    } else {
      if (phase == 2) stage_chunk(big, base + 2 * C, rs3, 0, T, T, d0, hd);   // V chunk
      stage_chunk(small, dob, C, q0, kQ, T, d0, hd);             // dO chunk
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // row statistics shared by the two key halves of a row block
  auto pair_reduce = [&](float v0, float v1, bool is_max, float& o0, float& o1) {
    if (is_max) {
      v0 = fmaxf(v0, __shfl_xor_sync(0xffffffffu, v0, 1));
      v0 = fmaxf(v0, __shfl_xor_sync(0xffffffffu, v0, 2));
      v1 = fmaxf(v1, __shfl_xor_sync(0xffffffffu, v1, 1));
      v1 = fmaxf(v1, __shfl_xor_sync(0xffffffffu, v1, 2));
    } else {
      v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
      v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
    }
    __syncthreads();
    if (t4 == 0) {
      scratch[kh * kQ + r0] = v0;
      scratch[kh * kQ + r0 + 8] = v1;
    }
    __syncthreads();
    if (is_max) {
      o0 = fmaxf(scratch[r0], scratch[kQ + r0]);
      o1 = fmaxf(scratch[r0 + 8], scratch[kQ + r0 + 8]);
    } else {
      o0 = scratch[r0] + scratch[kQ + r0];
      o1 = scratch[r0 + 8] + scratch[kQ + r0 + 8];
    }
  };

  // pipeline step: prefetch stage s + 1, wait for stage s; returns the shared-memory address of stage s's big tile
  auto begin_stage = [&](int s) -> uint32_t {
    const int buf = s & 1;
    if (s + 1 < 4 * nc) {
      issue(s + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    return smem_u32(stage0 + buf * kStageElems);
  };

  float acc[kNt][4];
  // [64 x T] += small_chunk big_chunk^T   (S = Q K^T or dP = dO V^T): warp -> 16 rows x T/2 keys
  auto gemm_rows_keys = [&](uint32_t big) {
    const uint32_t small = big + kT * kPitchC * 2;
#pragma unroll
    for (int k0 = 0; k0 < kDC; k0 += 16) {
      uint32_t a[4];
      ldsm_x4(a, small + ((rb * 16 + (lane & 15)) * kPitchC + k0 + (lane >> 4) * 8) * 2);
      const uint32_t brow = big + ((kh * keys_half + (mat >> 1) * 8 + (lane & 7)) * kPitchC + k0 + (mat & 1) * 8) * 2;
#pragma unroll
      for (int np = 0; np < kNt / 2; ++np) {
        if (np < npairs) {
          uint32_t bq[4];
          ldsm_x4(bq, brow + np * 16 * kPitchC * 2);
          mma16816(acc[2 * np], a, bq[0], bq[1]);
          mma16816(acc[2 * np + 1], a, bq[2], bq[3]);
        }
      }
    }
  };
  // dst[T x chunk] += (P or dS)^T small_chunk: M = keys (16-key blocks round-robin over the warps), N = the chunk's 32
  // columns, contraction over the 64 queries.  Both operands come out of row-major tiles through ldmatrix.trans.
  auto accum_keys = [&](uint32_t big, float* dst, int d0, bool masked, float out_scale) {
    const uint32_t small = big + kT * kPitchC * 2;
    uint32_t bfr[4][2][4];   // [k step][column half]: B fragments of the small tile, reused by every key block
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
        ldsm_x4_t(bfr[ks][ch], small + ((ks * 16 + (mat & 1) * 8 + (lane & 7)) * kPitchC + ch * 16 + (mat >> 1) * 8) * 2);
    for (int kb = warp; kb < T / 16; kb += 8) {
      float o[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];   // A = (P or dS)^T: m = key, k = query; stored [query][key]
        ldsm_x4_t(a, p_u + ((ks * 16 + (mat >> 1) * 8 + (lane & 7)) * kPitchP + kb * 16 + (mat & 1) * 8) * 2);
        if (masked) {  // negative halves are dropped probabilities: they do not take part in P^T dO
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] &= ~(((a[e] >> 15) & 0x00010001u) * 0xffffu);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          mma16816(o[2 * ch], a, bfr[ks][ch][0], bfr[ks][ch][1]);
          mma16816(o[2 * ch + 1], a, bfr[ks][ch][2], bfr[ks][ch][3]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nt * 8 + 2 * t4;
        if (d0 + col < hd) {
          float* op = dst + static_cast<long long>(kb * 16 + g) * (2ll * C) + col;
          red_add_v2(op, o[nt][0] * out_scale, o[nt][1] * out_scale);
          red_add_v2(op + 8 * 2ll * C, o[nt][2] * out_scale, o[nt][3] * out_scale);
        }
      }
    }
  };
  float* dk_dst = dkv + static_cast<long long>(b) * T * (2ll * C) + h * hd;
  float* dv_dst = dk_dst + C;

  issue(0, 0);
  // ---- phase A: S = Q K^T -> softmax -> P
#pragma unroll
  for (int i = 0; i < kNt; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int c = 0; c < nc; ++c) {
    const uint32_t big = begin_stage(c);
    gemm_rows_keys(big);
    __syncthreads();
  }
  {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < kNt; ++nt) {
      if (nt < 2 * npairs) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] *= scale;
        m0 = fmaxf(m0, fmaxf(acc[nt][0], acc[nt][1]));
        m1 = fmaxf(m1, fmaxf(acc[nt][2], acc[nt][3]));
      }
    }
    float M0, M1;
    pair_reduce(m0, m1, true, M0, M1);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < kNt; ++nt) {
      if (nt < 2 * npairs) {
        acc[nt][0] = __expf(acc[nt][0] - M0);
        acc[nt][1] = __expf(acc[nt][1] - M0);
        acc[nt][2] = __expf(acc[nt][2] - M1);
        acc[nt][3] = __expf(acc[nt][3] - M1);
        s0 += acc[nt][0] + acc[nt][1];
        s1 += acc[nt][2] + acc[nt][3];
      }
    }
    float S0, S1;
    pair_reduce(s0, s1, false, S0, S1);
    const float i0 = 1.f / S0, i1 = 1.f / S1;
#pragma unroll
    for (int nt = 0; nt < kNt; ++nt) {
      if (nt < 2 * npairs) {
        const int col = kh * keys_half + nt * 8 + 2 * t4;
        float p00 = acc[nt][0] * i0, p01 = acc[nt][1] * i0, p10 = acc[nt][2] * i1, p11 = acc[nt][3] * i1;
        if (DROP) {
          const DropCtx drop = drop_ctx(drop_rng, drop_p, drop_site);
          const uint32_t k0 = drop_pair(drop, b, heads, h, T, q0 + r0, col);
          const uint32_t k1 = drop_pair(drop, b, heads, h, T, q0 + r0 + 8, col);
          p00 = (k0 & 1u) ? p00 : -p00;
          p01 = (k0 & 2u) ? p01 : -p01;
          p10 = (k1 & 1u) ? p10 : -p10;
          p11 = (k1 & 2u) ? p11 : -p11;
        }
        *reinterpret_cast<uint32_t*>(P + r0 * kPitchP + col) = pack_bf16x2(p00, p01);
        *reinterpret_cast<uint32_t*>(P + (r0 + 8) * kPitchP + col) = pack_bf16x2(p10, p11);
      }
    }
  }
  // ---- phase V: dV[:, chunk] += P^T dO_chunk   (begin_stage's barrier publishes P)
  for (int c = 0; c < nc; ++c) {
    const uint32_t big = begin_stage(nc + c);
    accum_keys(big, dv_dst + c * kDC, c * kDC, DROP, inv_keep);
    __syncthreads();
  }
  // ---- phase B: dP = dO V^T -> dS overwrites P
#pragma unroll
  for (int i = 0; i < kNt; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int c = 0; c < nc; ++c) {
    const uint32_t big = begin_stage(2 * nc + c);
    gemm_rows_keys(big);
    __syncthreads();
  }
  {
    // dS = P * (dP' - rowsum(dP' * P)) * scale, in place over P (a thread touches only its own fragment slots);
    // dP' = dP masked and scaled by the probability dropout (sign bit of the parked P = dropped)
    float part0 = 0.f, part1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < kNt; ++nt) {
      if (nt < 2 * npairs) {
        const int col = kh * keys_half + nt * 8 + 2 * t4;
        const uint32_t ra = *reinterpret_cast<const uint32_t*>(P + r0 * kPitchP + col);
        const uint32_t rb2 = *reinterpret_cast<const uint32_t*>(P + (r0 + 8) * kPitchP + col);
        const float2 pa = unpack_bf16x2(ra & 0x7fff7fffu), pb = unpack_bf16x2(rb2 & 0x7fff7fffu);
        if (DROP) {
          acc[nt][0] = (ra & 0x00008000u) ? 0.f : acc[nt][0] * inv_keep;
          acc[nt][1] = (ra & 0x80000000u) ? 0.f : acc[nt][1] * inv_keep;
          acc[nt][2] = (rb2 & 0x00008000u) ? 0.f : acc[nt][2] * inv_keep;
          acc[nt][3] = (rb2 & 0x80000000u) ? 0.f : acc[nt][3] * inv_keep;
        }
        part0 += acc[nt][0] * pa.x + acc[nt][1] * pa.y;
        part1 += acc[nt][2] * pb.x + acc[nt][3] * pb.y;
      }
    }
    float rs0, rs1;
    pair_reduce(part0, part1, false, rs0, rs1);
#pragma unroll
    for (int nt = 0; nt < kNt; ++nt) {
      if (nt < 2 * npairs) {
        const int col = kh * keys_half + nt * 8 + 2 * t4;
        const float2 pa = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(P + r0 * kPitchP + col) & 0x7fff7fffu);
        const float2 pb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(P + (r0 + 8) * kPitchP + col) & 0x7fff7fffu);
        *reinterpret_cast<uint32_t*>(P + r0 * kPitchP + col) =
            pack_bf16x2(pa.x * (acc[nt][0] - rs0) * scale, pa.y * (acc[nt][1] - rs0) * scale);
        *reinterpret_cast<uint32_t*>(P + (r0 + 8) * kPitchP + col) =
            pack_bf16x2(pb.x * (acc[nt][2] - rs1) * scale, pb.y * (acc[nt][3] - rs1) * scale);
      }
    }
  }
  // ---- phase C: dQ[:, chunk] = dS K_chunk;  dK[:, chunk] += dS^T Q_chunk
  for (int c = 0; c < nc; ++c) {
    const uint32_t big = begin_stage(3 * nc + c);
    const int d0 = c * kDC;
    {
      float o[2][4];   // warp -> 16 rows x 16 of the chunk's 32 columns
      o[0][0] = o[0][1] = o[0][2] = o[0][3] = o[1][0] = o[1][1] = o[1][2] = o[1][3] = 0.f;
      const uint32_t prow = p_u + ((rb * 16 + (lane & 15)) * kPitchP + (lane >> 4) * 8) * 2;
      const uint32_t krow = big + (((mat & 1) * 8 + (lane & 7)) * kPitchC + kh * 16 + (mat >> 1) * 8) * 2;
      for (int k0 = 0; k0 < T; k0 += 16) {
        uint32_t a[4], bq[4];
        ldsm_x4(a, prow + k0 * 2);
        ldsm_x4_t(bq, krow + k0 * kPitchC * 2);
        mma16816(o[0], a, bq[0], bq[1]);
        mma16816(o[1], a, bq[2], bq[3]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = d0 + kh * 16 + nt * 8 + 2 * t4;
        if (col < hd) {
          const int qr = q0 + r0;
          bf16* op = dqkv + (static_cast<long long>(b) * T + qr) * rs3 + h * hd + col;
          if (qr < T) *reinterpret_cast<uint32_t*>(op) = pack_bf16x2(o[nt][0], o[nt][1]);
          if (qr + 8 < T) *reinterpret_cast<uint32_t*>(op + 8 * rs3) = pack_bf16x2(o[nt][2], o[nt][3]);
        }
      }
    }
    accum_keys(big, dk_dst + d0, d0, false, 1.f);
    __syncthreads();
  }
}

// dqkv[:, :, C:3C] = bf16(dkv), 8 elements per thread
__global__ void __launch_bounds__(256) dkv_cast_kernel(const float* __restrict__ dkv, bf16* __restrict__ dqkv,
                                                       long long rows, int C) {
  const int per_row = 2 * C / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * per_row) return;
  const long long r = i / per_row;
  const int c = static_cast<int>(i - r * per_row) * 8;
  const float4 a = *reinterpret_cast<const float4*>(dkv + r * 2 * C + c);
  const float4 b2 = *reinterpret_cast<const float4*>(dkv + r * 2 * C + c + 4);
  *reinterpret_cast<uint4*>(dqkv + r * 3 * C + C + c) =
      make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b2.x, b2.y), pack_bf16x2(b2.z, b2.w));
}

}  // namespace

extern "C" int tfpp_fusion_attn(const void* qkv, void* out, int batch, int tokens, int channels, int heads,
                                tfpp_stream_t stream_) {
  return tfpp_fusion_attn_dropout(qkv, out, batch, tokens, channels, heads, nullptr, 0.f, 0u, stream_);
}

extern "C" int tfpp_fusion_attn_dropout(const void* qkv, void* out, int batch, int tokens, int channels, int heads,
                                        const unsigned long long* drop_rng, float drop_p, unsigned drop_site,
                                        tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(tokens <= kT && tokens % 32 == 0, "tokens must be a multiple of 32 and <= 320");
  TFPP_CHECK_ARG(channels % heads == 0 && (channels / heads) % 2 == 0, "even head dim required");
  const size_t smem = sizeof(bf16) * (kQ * kPitchP + 2 * kStageElems) + sizeof(float) * 2 * kQ;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fusion_attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fusion_attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr_set = true;
  }
  const int hd = channels / heads;
  dim3 grid(ceil_div(tokens, kQ), heads, batch);
  if (drop_rng != nullptr && drop_p > 0.f)
    fusion_attn_fwd_kernel<true><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out),
                                                              tokens, channels, heads,
                                                              1.0f / sqrtf(static_cast<float>(hd)), drop_rng, drop_p,
                                                              drop_site);
  else
    fusion_attn_fwd_kernel<false><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(qkv), static_cast<bf16*>(out),
                                                               tokens, channels, heads,
                                                               1.0f / sqrtf(static_cast<float>(hd)), nullptr, 0.f, 0u);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_fusion_attn_bwd(const void* qkv, const void* dout, void* dqkv, float* dkv_ws, int batch, int tokens,
                                    int channels, int heads, tfpp_stream_t stream_) {
  return tfpp_fusion_attn_bwd_dropout(qkv, dout, dqkv, dkv_ws, batch, tokens, channels, heads, nullptr, 0.f, 0u, stream_);
}

extern "C" int tfpp_fusion_attn_bwd_dropout(const void* qkv, const void* dout, void* dqkv, float* dkv_ws, int batch,
                                            int tokens, int channels, int heads, const unsigned long long* drop_rng,
                                            float drop_p, unsigned drop_site, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(tokens <= kT && tokens % 32 == 0, "tokens must be a multiple of 32 and <= 320");
  TFPP_CHECK_ARG(channels % heads == 0 && (channels / heads) % 2 == 0 && channels % 8 == 0,
                 "even head dim and channels % 8 == 0 required");
  const size_t smem = sizeof(bf16) * (kQ * kPitchP + 2 * kStageElems) + sizeof(float) * 2 * kQ;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fusion_attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fusion_attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long rows = static_cast<long long>(batch) * tokens;
  cudaMemsetAsync(dkv_ws, 0, sizeof(float) * rows * 2 * channels, stream);
  const int hd = channels / heads;
  dim3 grid(ceil_div(tokens, kQ), heads, batch);
  if (drop_rng != nullptr && drop_p > 0.f)
    fusion_attn_bwd_kernel<true><<<grid, 256, smem, stream>>>(
        static_cast<const bf16*>(qkv), static_cast<const bf16*>(dout), static_cast<bf16*>(dqkv), dkv_ws, tokens, channels,
        heads, 1.0f / sqrtf(static_cast<float>(hd)), drop_rng, drop_p, drop_site);
  else
    fusion_attn_bwd_kernel<false><<<grid, 256, smem, stream>>>(
        static_cast<const bf16*>(qkv), static_cast<const bf16*>(dout), static_cast<bf16*>(dqkv), dkv_ws, tokens, channels,
        heads, 1.0f / sqrtf(static_cast<float>(hd)), nullptr, 0.f, 0u);
  TFPP_CHECK_LAUNCH();
  dkv_cast_kernel<<<static_cast<unsigned>(ceil_div_ll(rows * 2 * channels / 8, 256)), 256, 0, stream>>>(
      dkv_ws, static_cast<bf16*>(dqkv), rows, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
