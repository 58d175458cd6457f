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

// ------------------------------------------------------------------------------------------------ fusion attention bwd
// Per (64-query tile, head, batch): recompute P = softmax(Q K^T * scale); dP = dO V^T; dS = P * (dP - rowsum(dP * P));
// dQ = dS K * scale (written, bf16); dK += dS^T Q * scale, dV += P^T dO (fp32 atomics into dkv (B,T,2C): the key/value
// gradients of one head receive contributions from all 5 query tiles).  mma.sync m16n8k16 bf16, fp32 accumulate.
constexpr int kAttT = 320;
constexpr int kAttQ = 64;
constexpr int kChunk = 64;
constexpr int kPitchK = kChunk + 8;   // bf16 row pitch of [rows][64] staging tiles
constexpr int kPitchT = kAttT + 8;    // bf16 row pitch of [64][320] tiles
constexpr int kPitchS = kAttT + 4;    // f32 row pitch of [64][320]
constexpr int kPitchQ = kAttQ + 8;    // bf16 row pitch of [rows][64 queries] transposed tiles

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void load_a(uint32_t* a, const bf16* tile, int pitch, int r0, int k0, int g, int t4) {
  const bf16* p = tile + (r0 + g) * pitch + k0 + 2 * t4;
  a[0] = *reinterpret_cast<const uint32_t*>(p);
  a[1] = *reinterpret_cast<const uint32_t*>(p + 8 * pitch);
  a[2] = *reinterpret_cast<const uint32_t*>(p + 8);
  a[3] = *reinterpret_cast<const uint32_t*>(p + 8 * pitch + 8);
}
__device__ __forceinline__ void load_b(uint32_t* b, const bf16* tile, int pitch, int n0, int k0, int g, int t4) {
  const bf16* p = tile + (n0 + g) * pitch + k0 + 2 * t4;
  b[0] = *reinterpret_cast<const uint32_t*>(p);
  b[1] = *reinterpret_cast<const uint32_t*>(p + 8);
}

// stage rows [r_begin, r_begin + nrows) x 64 head-dim columns [d0, d0+64) of a (.., 3C) / (.., C) matrix into
// tile[r][d] (bf16 pairs), zero outside [0,T) x [0,hd)
__device__ __forceinline__ void stage_rows(bf16* tile, int pitch, const bf16* src, long long row_stride, int col0,
                                           int r_begin, int nrows, int T, int d0, int hd) {
  for (int i = threadIdx.x; i < nrows * (kChunk / 2); i += blockDim.x) {
    const int r = i / (kChunk / 2), d = (i % (kChunk / 2)) * 2;
    uint32_t v = 0;
    if (r_begin + r < T && d0 + d < hd)
      v = *reinterpret_cast<const uint32_t*>(src + (r_begin + r) * row_stride + col0 + d0 + d);
    *reinterpret_cast<uint32_t*>(tile + r * pitch + d) = v;
  }
}
// same but transposed: tile[d][r]
__device__ __forceinline__ void stage_rows_t(bf16* tile, int pitch, const bf16* src, long long row_stride, int col0,
                                             int r_begin, int nrows, int T, int d0, int hd) {
  for (int i = threadIdx.x; i < nrows * (kChunk / 2); i += blockDim.x) {
    const int r = i / (kChunk / 2), d = (i % (kChunk / 2)) * 2;
    uint32_t v = 0;
    if (r_begin + r < T && d0 + d < hd)
      v = *reinterpret_cast<const uint32_t*>(src + (r_begin + r) * row_stride + col0 + d0 + d);
    const __nv_bfloat162 pr = *reinterpret_cast<__nv_bfloat162*>(&v);
    tile[d * pitch + r] = pr.x;
    tile[(d + 1) * pitch + r] = pr.y;
  }
}

__global__ void __launch_bounds__(256) fusion_attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                              bf16* __restrict__ dqkv, float* __restrict__ dkv, int T,
                                                              int C, int heads, float scale) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* dS = reinterpret_cast<bf16*>(att_smem);             // [64][kPitchT]   P (bf16), then dS in place
  bf16* PT = dS + kAttQ * kPitchT;                          // [320][kPitchQ]  P^T
  bf16* dST = PT + kAttT * kPitchQ;                         // [320][kPitchQ]  dS^T
  bf16* KV = dST + kAttT * kPitchQ;                         // [320][kPitchK] K / V chunk, or [64][kPitchT] transposed
  bf16* Qs = KV + kAttT * kPitchK;                          // [64][kPitchK] Q / dO chunk, or [64][kPitchQ] transposed
  float* scratch = reinterpret_cast<float*>(Qs + kAttQ * kPitchK);  // [2][64] cross-warp row reductions
  const int hd = C / heads;
  const int q0 = blockIdx.x * kAttQ, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const long long rs3 = 3ll * C;
  const bf16* base = qkv + static_cast<long long>(b) * T * rs3;
  const bf16* dob = dout + static_cast<long long>(b) * T * C;
  const int n_chunks = (hd + kChunk - 1) / kChunk;
  const int Tp = (T + 15) & ~15;
  const int rb = warp & 3, kh = warp >> 2;
  const int keys_half = Tp / 2, ntiles = keys_half / 8;
  constexpr int kNt = (kAttT / 2) / 8;  // 20
  const int r0 = rb * 16 + g;           // this thread's rows: r0 and r0 + 8

  // cross-warp (two key halves) row reduction helper: quad shuffle, then shared scratch
  auto row_reduce = [&](float v0, float v1, bool is_max, float& o0, float& o1) {
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
      scratch[kh * kAttQ + r0] = v0;
      scratch[kh * kAttQ + r0 + 8] = v1;
    }
    __syncthreads();
    if (is_max) {
      o0 = fmaxf(scratch[r0], scratch[kAttQ + r0]);
      o1 = fmaxf(scratch[r0 + 8], scratch[kAttQ + r0 + 8]);
    } else {
      o0 = scratch[r0] + scratch[kAttQ + r0];
      o1 = scratch[r0 + 8] + scratch[kAttQ + r0 + 8];
    }
  };

  // ---- two [64 x T] products with the same structure: S = Q K^T (pass 0) and dP = dO V^T (pass 1)
  for (int pass = 0; pass < 2; ++pass) {
    float acc[kNt][4];
#pragma unroll
    for (int i = 0; i < kNt; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int d0 = ch * kChunk;
      __syncthreads();
      if (pass == 0) {
        stage_rows(Qs, kPitchK, base, rs3, h * hd, q0, kAttQ, T, d0, hd);
        stage_rows(KV, kPitchK, base, rs3, C + h * hd, 0, Tp, T, d0, hd);
      } else {
        stage_rows(Qs, kPitchK, dob, C, h * hd, q0, kAttQ, T, d0, hd);
        stage_rows(KV, kPitchK, base, rs3, 2 * C + h * hd, 0, Tp, T, d0, hd);
      }
      __syncthreads();
      const int kmax = min(kChunk, ((hd - d0) + 15) & ~15);
      for (int k0 = 0; k0 < kmax; k0 += 16) {
        uint32_t a[4];
        load_a(a, Qs, kPitchK, rb * 16, k0, g, t4);
#pragma unroll
        for (int nt = 0; nt < kNt; ++nt) {
          if (nt < ntiles) {
            uint32_t bf[2];
            load_b(bf, KV, kPitchK, kh * keys_half + nt * 8, k0, g, t4);
            mma16816(acc[nt], a, bf);
          }
        }
      }
    }
    if (pass == 0) {
      // softmax over keys, all in registers (+ two cross-warp reductions); P -> dS buffer (row major) and P^T
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < kNt; ++nt) {
        if (nt < ntiles) {
          const int col = kh * keys_half + nt * 8 + 2 * t4;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[nt][e] = (col + (e & 1) < T) ? acc[nt][e] * scale : -INFINITY;
          }
          m0 = fmaxf(m0, fmaxf(acc[nt][0], acc[nt][1]));
          m1 = fmaxf(m1, fmaxf(acc[nt][2], acc[nt][3]));
        }
      }
      float M0, M1;
      row_reduce(m0, m1, true, M0, M1);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < kNt; ++nt) {
        if (nt < ntiles) {
          acc[nt][0] = __expf(acc[nt][0] - M0);
          acc[nt][1] = __expf(acc[nt][1] - M0);
          acc[nt][2] = __expf(acc[nt][2] - M1);
          acc[nt][3] = __expf(acc[nt][3] - M1);
          s0 += acc[nt][0] + acc[nt][1];
          s1 += acc[nt][2] + acc[nt][3];
        }
      }
      float S0, S1;
      row_reduce(s0, s1, false, S0, S1);
      const float i0 = 1.f / S0, i1 = 1.f / S1;
#pragma unroll
      for (int nt = 0; nt < kNt; ++nt) {
        if (nt < ntiles) {
          const int col = kh * keys_half + nt * 8 + 2 * t4;
          const float p00 = acc[nt][0] * i0, p01 = acc[nt][1] * i0, p10 = acc[nt][2] * i1, p11 = acc[nt][3] * i1;
          *reinterpret_cast<uint32_t*>(dS + r0 * kPitchT + col) = pack_bf16x2(p00, p01);
          *reinterpret_cast<uint32_t*>(dS + (r0 + 8) * kPitchT + col) = pack_bf16x2(p10, p11);
          PT[col * kPitchQ + r0] = f2bf(p00);
          PT[(col + 1) * kPitchQ + r0] = f2bf(p01);
          PT[col * kPitchQ + r0 + 8] = f2bf(p10);
          PT[(col + 1) * kPitchQ + r0 + 8] = f2bf(p11);
        }
      }
    } else {
      // dS = P * (dP - rowsum(dP * P)) * scale, in place over P (each thread touches only its own fragment slots)
      float part0 = 0.f, part1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < kNt; ++nt) {
        if (nt < ntiles) {
          const int col = kh * keys_half + nt * 8 + 2 * t4;
          const float2 pa = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dS + r0 * kPitchT + col));
          const float2 pb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dS + (r0 + 8) * kPitchT + col));
          part0 += acc[nt][0] * pa.x + acc[nt][1] * pa.y;
          part1 += acc[nt][2] * pb.x + acc[nt][3] * pb.y;
        }
      }
      float rs0, rs1;
      row_reduce(part0, part1, false, rs0, rs1);
#pragma unroll
      for (int nt = 0; nt < kNt; ++nt) {
        if (nt < ntiles) {
          const int col = kh * keys_half + nt * 8 + 2 * t4;
          const float2 pa = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dS + r0 * kPitchT + col));
          const float2 pb = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dS + (r0 + 8) * kPitchT + col));
          const float v00 = pa.x * (acc[nt][0] - rs0) * scale, v01 = pa.y * (acc[nt][1] - rs0) * scale;
          const float v10 = pb.x * (acc[nt][2] - rs1) * scale, v11 = pb.y * (acc[nt][3] - rs1) * scale;
          *reinterpret_cast<uint32_t*>(dS + r0 * kPitchT + col) = pack_bf16x2(v00, v01);
          *reinterpret_cast<uint32_t*>(dS + (r0 + 8) * kPitchT + col) = pack_bf16x2(v10, v11);
          dST[col * kPitchQ + r0] = f2bf(v00);
          dST[(col + 1) * kPitchQ + r0] = f2bf(v01);
          dST[col * kPitchQ + r0 + 8] = f2bf(v10);
          dST[(col + 1) * kPitchQ + r0 + 8] = f2bf(v11);
        }
      }
    }
  }
  __syncthreads();

  // ---- dQ = dS K  (64 x hd): per 64-column chunk; B operand = K^T chunk staged transposed [d][key]
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int d0 = ch * kChunk;
    __syncthreads();
    stage_rows_t(KV, kPitchT, base, rs3, C + h * hd, 0, Tp, T, d0, hd);
    __syncthreads();
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    for (int k0 = 0; k0 < Tp; k0 += 16) {
      uint32_t a[4];
      load_a(a, dS, kPitchT, rb * 16, k0, g, t4);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bf[2];
        load_b(bf, KV, kPitchT, kh * 32 + nt * 8, k0, g, t4);
        mma16816(o[nt], a, bf);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = d0 + kh * 32 + nt * 8 + 2 * t4;
      if (col < hd) {
        const int qr = q0 + r0;
        bf16* op = dqkv + (static_cast<long long>(b) * T + qr) * rs3 + h * hd + col;
        if (qr < T) *reinterpret_cast<uint32_t*>(op) = pack_bf16x2(o[nt][0], o[nt][1]);
        if (qr + 8 < T) *reinterpret_cast<uint32_t*>(op + 8 * rs3) = pack_bf16x2(o[nt][2], o[nt][3]);
      }
    }
  }

  // ---- dK += dS^T Q  and  dV += P^T dO  (T x hd each, contraction over this CTA's 64 queries)
  // A = dS^T / P^T [key][q]; B = Q^T / dO^T chunk staged transposed [d][q]; warp w owns keys [w*40, w*40+40)? T/8
  // warps is not a multiple of 16 in general, so each warp walks 16-key blocks round-robin.
  for (int which = 0; which < 2; ++which) {
    const bf16* A = which == 0 ? dST : PT;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int d0 = ch * kChunk;
      __syncthreads();
      if (which == 0) stage_rows_t(Qs, kPitchQ, base, rs3, h * hd, q0, kAttQ, T, d0, hd);
      else stage_rows_t(Qs, kPitchQ, dob, C, h * hd, q0, kAttQ, T, d0, hd);
      __syncthreads();
      for (int kb = warp; kb < Tp / 16; kb += 8) {
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
        for (int k0 = 0; k0 < kAttQ; k0 += 16) {
          uint32_t a[4];
          load_a(a, A, kPitchQ, kb * 16, k0, g, t4);
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            uint32_t bf[2];
            load_b(bf, Qs, kPitchQ, nt * 8, k0, g, t4);
            mma16816(o[nt], a, bf);
          }
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int col = d0 + nt * 8 + 2 * t4;
          if (col < hd) {
            const int key = kb * 16 + g;
            float* op = dkv + (static_cast<long long>(b) * T + key) * (2ll * C) + which * C + h * hd + col;
            if (key < T) {
              atomicAdd(op, o[nt][0]);
              atomicAdd(op + 1, o[nt][1]);
            }
            if (key + 8 < T) {
              atomicAdd(op + 8 * 2ll * C, o[nt][2]);
              atomicAdd(op + 8 * 2ll * C + 1, o[nt][3]);
            }
          }
        }
      }
    }
  }
}

// dqkv[:, :, C:3C] = bf16(dkv)
__global__ void __launch_bounds__(256) dkv_cast_kernel(const float* __restrict__ dkv, bf16* __restrict__ dqkv,
                                                       long long rows, int C) {
  const long long total = rows * 2 * C;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / (2 * C);
  const int c = static_cast<int>(i % (2 * C));
  dqkv[r * 3 * C + C + c] = f2bf(dkv[i]);
}

// ------------------------------------------------------------------------------------------------ decoder attention bwd
// One CTA per (batch, head); everything in fp32 shared memory (Tq <= 16, Tk <= 128, hd <= 64).
__global__ void __launch_bounds__(256) small_mha_bwd_kernel(const bf16* __restrict__ q, long long q_sb, long long q_sr,
                                                            const bf16* __restrict__ k, long long k_sb, long long k_sr,
                                                            const bf16* __restrict__ v, long long v_sb, long long v_sr,
                                                            const bf16* __restrict__ dout, long long o_sb, long long o_sr,
                                                            bf16* __restrict__ dq, long long dq_sb, long long dq_sr,
                                                            bf16* __restrict__ dk, long long dk_sb, long long dk_sr,
                                                            bf16* __restrict__ dv, long long dv_sb, long long dv_sr,
                                                            int accumulate_kv, int Tq, int Tk, int hd, float scale) {
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
    float dpl[4];  // Tk <= 128 -> up to 4 per lane
    int n = 0;
    for (int c = lane; c < Tk; c += 32, ++n) {
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(dos[r * P1 + d], vs[c * P1 + d], a);
      const float p = pr[c] * inv;
      pr[c] = p;
      dpl[n] = a;
      rsum = fmaf(a, p, rsum);
    }
    rsum = warp_sum(rsum);
    n = 0;
    __syncwarp();
    // keep P in a register copy for dV: store dS in ps after use -> need P too: pack dS into a second pass
    for (int c = lane; c < Tk; c += 32, ++n) {
      const float p = pr[c];
      dpl[n] = p * (dpl[n] - rsum) * scale;  // dS
    }
    // dV += P^T dO is accumulated after the loop (needs all rows) -> store P in place and dS in registers -> write dQ now
    __syncwarp();
    // dQ[r] = dS[r,:] K
    // stash dS into shared by swapping with P: we still need P for dV, so use two-step: write dS to a side buffer
    float* dsr = ps + Tq * Tk + r * Tk;  // second Tq x Tk block
    n = 0;
    for (int c = lane; c < Tk; c += 32, ++n) dsr[c] = dpl[n];
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
  for (int i = threadIdx.x; i < (n_wp + 1) * D; i += blockDim.x) x[i] = joined[static_cast<long long>(b) * (n_wp + 1) * D + i];
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
      djoined[(static_cast<long long>(b) * (n_wp + 1) + t) * D + k] = a;
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

extern "C" int tfpp_fusion_attn_bwd(const void* qkv, const void* dout, void* dqkv, float* dkv_ws, int batch, int tokens,
                                    int channels, int heads, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(tokens <= kAttT && tokens % 16 == 0, "tokens must be a multiple of 16 and <= 320");
  TFPP_CHECK_ARG(channels % heads == 0 && (channels / heads) % 2 == 0, "even head dim required");
  const size_t smem = sizeof(bf16) * (kAttQ * kPitchT + 2 * kAttT * kPitchQ + kAttT * kPitchK + kAttQ * kPitchK) +
                      sizeof(float) * 2 * kAttQ;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fusion_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr_set = true;
  }
  TFPP_CHECK_ARG(smem <= 227 * 1024, "fusion_attn_bwd shared memory budget exceeded");
  cudaError_t e = cudaMemsetAsync(dkv_ws, 0, sizeof(float) * 2ull * channels * tokens * batch, stream);
  if (e != cudaSuccess) {
    tfpp_set_error("memset: %s", cudaGetErrorString(e));
    return TFPP_ERR_CUDA;
  }
  const int hd = channels / heads;
  dim3 grid(ceil_div(tokens, kAttQ), heads, batch);
  fusion_attn_bwd_kernel<<<grid, 256, smem, stream>>>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(dout),
                                                      static_cast<bf16*>(dqkv), dkv_ws, tokens, channels, heads,
                                                      1.0f / sqrtf(static_cast<float>(hd)));
  TFPP_CHECK_LAUNCH();
  const long long rows = static_cast<long long>(batch) * tokens;
  dkv_cast_kernel<<<static_cast<int>(ceil_div_ll(rows * 2 * channels, 256)), 256, 0, stream>>>(
      dkv_ws, static_cast<bf16*>(dqkv), rows, channels);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_small_mha_bwd(const void* q, long long q_sb, long long q_sr, const void* k, long long k_sb,
                                  long long k_sr, const void* v, long long v_sb, long long v_sr, const void* dout,
                                  long long o_sb, long long o_sr, void* dq, long long dq_sb, long long dq_sr, void* dk,
                                  long long dk_sb, long long dk_sr, void* dv, long long dv_sb, long long dv_sr,
                                  int accumulate_kv, int batch, int heads, int tq, int tk, int head_dim,
                                  tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(tk <= 128 && tq <= 16 && head_dim <= 64, "small_mha_bwd: tq <= 16, tk <= 128, head_dim <= 64");
  const size_t smem = sizeof(float) * ((2 * tq + 2 * tk) * (head_dim + 1) + 2 * tq * tk);
  dim3 grid(batch, heads);
  small_mha_bwd_kernel<<<grid, 256, smem, stream>>>(
      static_cast<const bf16*>(q), q_sb, q_sr, static_cast<const bf16*>(k), k_sb, k_sr, static_cast<const bf16*>(v), v_sb,
      v_sr, static_cast<const bf16*>(dout), o_sb, o_sr, static_cast<bf16*>(dq), dq_sb, dq_sr, static_cast<bf16*>(dk),
      dk_sb, dk_sr, static_cast<bf16*>(dv), dv_sb, dv_sr, accumulate_kv, tq, tk, head_dim,
      1.0f / sqrtf(static_cast<float>(head_dim)));
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
