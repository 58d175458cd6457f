// Kernels of the ``bev_encoder`` backbone (reference team_code/bev_encoder.py):
//
//  * the camera -> BEV lift (bev_encoder.py:179-199: F.grid_sample of the perspective features at every voxel centre of a
//    256 x 256 x 96 volume, sum over height, normalise, transpose, mask) restated as what it is for a pinhole camera
//    without roll / pitch: the horizontal pixel coordinate of a voxel depends on (depth, width) only, the vertical one on
//    (depth, height) only, so the bilinear sample separates and
//        bev[b, w, d, :] = s(d, w) * ( wl(d, w) * V[b, d, x0(d, w), :] + wr(d, w) * V[b, d, x0(d, w) + 1, :] ),
//        V[b, d, x, :]   = sum_y A[d, y] * img[b, y, x, :],   A[d, y] = sum_h (vertical bilinear weight of row y)
//    — one 256 x 32 matrix A and three (256, 256) tables instead of 805 M gathered taps per sample (the 75 MB grid is
//    never read on the device).  The tables are derived on the host from the module's own ``grid`` /
//    ``bev_projection_normalizer`` / ``valid_bev_pixels`` parameters (carla_garage_b200/nn/bev_encoder.py).
//  * nn.InstanceNorm2d(affine=False) + ReLU / GELU forward and backward (bev_encoder.py:126-137,253-262), channel-strided
//    so that the compressor can write straight into the 40-channel [bev | lidar | 0] tensor the BEV stem convolves.
//
// Every kernel is templated on the activation type (bf16 production path / fp32 parity mode, ``f32`` flag of the C ABI).
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Vec4<bf16> {
  static __device__ __forceinline__ float4 load(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(bf16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
};
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) { return bf2f(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = f2bf(v); }

__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == ACT_GELU)  // d/dz [z * Phi(z)] = Phi(z) + z * phi(z)
    return 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
  return 1.f;
}

// ------------------------------------------------------------------------------------------------ instance norm
// grid (chunks, B), 256 threads; a thread owns one 4-channel group and walks the pixels of its chunk.
template <typename T>
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const T* __restrict__ x, long long pix_stride, int HW, int C,
                                                             int ppb, float* __restrict__ sum, float* __restrict__ sq) {
  extern __shared__ float sm[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, c4n = C / 4;
  const int rows = blockDim.x / c4n, cg = threadIdx.x % c4n, prow = threadIdx.x / c4n;
  const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
  const T* xb = x + static_cast<long long>(b) * HW * pix_stride + cg * 4;
  if (prow < rows) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int p = p0 + prow; p < p1; p += rows) {
      const float4 v = Vec4<T>::load(xb + p * pix_stride);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    float* a = sm + cg * 4;
    atomicAdd(a, s.x); atomicAdd(a + 1, s.y); atomicAdd(a + 2, s.z); atomicAdd(a + 3, s.w);
    a += C;
    atomicAdd(a, q.x); atomicAdd(a + 1, q.y); atomicAdd(a + 2, q.z); atomicAdd(a + 3, q.w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(sum + static_cast<long long>(b) * C + i, sm[i]);
    atomicAdd(sq + static_cast<long long>(b) * C + i, sm[C + i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const T* __restrict__ x, long long x_ps,
                                                             const float* __restrict__ sum, const float* __restrict__ sq,
                                                             float eps, int act, T* __restrict__ y, long long y_ps,
                                                             float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                                             int HW, int C, int ppb) {
  extern __shared__ float sm[];  // mean[C] | invstd[C]
  const int b = blockIdx.y, c4n = C / 4;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float m = sum[static_cast<long long>(b) * C + i] / HW;
    const float var = fmaxf(sq[static_cast<long long>(b) * C + i] / HW - m * m, 0.f);  // biased, as nn.InstanceNorm2d
    const float is = rsqrtf(var + eps);
    sm[i] = m;
    sm[C + i] = is;
    if (blockIdx.x == 0 && mean_out != nullptr) {
      mean_out[static_cast<long long>(b) * C + i] = m;
      invstd_out[static_cast<long long>(b) * C + i] = is;
    }
  }
  __syncthreads();
  const int rows = blockDim.x / c4n, cg = threadIdx.x % c4n, prow = threadIdx.x / c4n;
  if (prow >= rows) return;
  const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
  const T* xb = x + static_cast<long long>(b) * HW * x_ps + cg * 4;
  T* yb = y + static_cast<long long>(b) * HW * y_ps + cg * 4;
  const float4 m = *reinterpret_cast<const float4*>(sm + cg * 4);
  const float4 is = *reinterpret_cast<const float4*>(sm + C + cg * 4);
  for (int p = p0 + prow; p < p1; p += rows) {
    float4 v = Vec4<T>::load(xb + p * x_ps);
    v.x = apply_act((v.x - m.x) * is.x, act);
    v.y = apply_act((v.y - m.y) * is.y, act);
    v.z = apply_act((v.z - m.z) * is.z, act);
    v.w = apply_act((v.w - m.w) * is.w, act);
    Vec4<T>::store(yb + p * y_ps, v);
  }
}

// backward, pass 1: s1[b,c] = sum dz, s2[b,c] = sum dz * z with z = (x - mean) * invstd, dz = dy * act'(z)
template <typename T>
__global__ void __launch_bounds__(256) instnorm_bwd_reduce_kernel(const T* __restrict__ dy, long long dy_ps,
                                                                  const T* __restrict__ x, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, int act,
                                                                  float* __restrict__ s1, float* __restrict__ s2, int HW,
                                                                  int C, int ppb) {
  extern __shared__ float sm[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, c4n = C / 4;
  const int rows = blockDim.x / c4n, cg = threadIdx.x % c4n, prow = threadIdx.x / c4n;
  const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
  if (prow < rows) {
    const T* xb = x + static_cast<long long>(b) * HW * C + cg * 4;
    const T* db = dy + static_cast<long long>(b) * HW * dy_ps + cg * 4;
    const float4 m = *reinterpret_cast<const float4*>(mean + static_cast<long long>(b) * C + cg * 4);
    const float4 is = *reinterpret_cast<const float4*>(invstd + static_cast<long long>(b) * C + cg * 4);
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = p0 + prow; p < p1; p += rows) {
      const float4 v = Vec4<T>::load(xb + static_cast<long long>(p) * C);
      const float4 d = Vec4<T>::load(db + p * dy_ps);
      const float z[4] = {(v.x - m.x) * is.x, (v.y - m.y) * is.y, (v.z - m.z) * is.z, (v.w - m.w) * is.w};
      const float g[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dz = g[j] * act_grad(z[j], act);
        a1[j] += dz;
        a2[j] = fmaf(dz, z[j], a2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(sm + cg * 4 + j, a1[j]);
      atomicAdd(sm + C + cg * 4 + j, a2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(s1 + static_cast<long long>(b) * C + i, sm[i]);
    atomicAdd(s2 + static_cast<long long>(b) * C + i, sm[C + i]);
  }
}

// pass 2: dx = invstd * (dz - s1 / N - z * s2 / N)
template <typename T>
__global__ void __launch_bounds__(256) instnorm_bwd_apply_kernel(const T* __restrict__ dy, long long dy_ps,
                                                                 const T* __restrict__ x, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, int act,
                                                                 const float* __restrict__ s1, const float* __restrict__ s2,
                                                                 T* __restrict__ dx, int HW, int C, int ppb) {
  const int b = blockIdx.y, c4n = C / 4;
  const int rows = blockDim.x / c4n, cg = threadIdx.x % c4n, prow = threadIdx.x / c4n;
  if (prow >= rows) return;
  const int p0 = blockIdx.x * ppb, p1 = min(HW, p0 + ppb);
  const long long bc = static_cast<long long>(b) * C + cg * 4;
  const T* xb = x + static_cast<long long>(b) * HW * C + cg * 4;
  const T* db = dy + static_cast<long long>(b) * HW * dy_ps + cg * 4;
  T* ob = dx + static_cast<long long>(b) * HW * C + cg * 4;
  const float inv_n = 1.f / HW;
  float m[4], is[4], k1[4], k2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m[j] = mean[bc + j];
    is[j] = invstd[bc + j];
    k1[j] = s1[bc + j] * inv_n;
    k2[j] = s2[bc + j] * inv_n;
  }
  for (int p = p0 + prow; p < p1; p += rows) {
    const float4 v = Vec4<T>::load(xb + static_cast<long long>(p) * C);
    const float4 d = Vec4<T>::load(db + p * dy_ps);
    const float xv[4] = {v.x, v.y, v.z, v.w}, g[4] = {d.x, d.y, d.z, d.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = (xv[j] - m[j]) * is[j];
      const float dz = g[j] * act_grad(z, act);
      o[j] = is[j] * (dz - k1[j] - z * k2[j]);
    }
    Vec4<T>::store(ob + static_cast<long long>(p) * C, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ------------------------------------------------------------------------------------------------ camera -> BEV lift
// grid (D, B), 256 threads.  img (B, IH, IW, C) NHWC, out (B, W, D, C) NHWC (rows = width index, columns = depth index:
// the transpose of bev_encoder.py:193 is folded into the store).
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(bf16* p, const float* v) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                              pack_bf16x2(v[6], v[7]));
  }
};

// a thread owns 8 channels: 16-byte loads of the feature rows, 16-byte stores of the BEV cells (C % 8 == 0)
template <typename T>
__global__ void __launch_bounds__(256) bev_lift_kernel(const T* __restrict__ img, const float* __restrict__ A,
                                                       const int* __restrict__ x0, const float* __restrict__ wl,
                                                       const float* __restrict__ wr, T* __restrict__ out, int IH, int IW,
                                                       int C, int D, int W) {
  extern __shared__ __align__(16) float sm[];  // V[IW * C] | a[IH]
  float* V = sm;
  float* a = sm + IW * C;
  const int d = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < IH; i += blockDim.x) a[i] = A[d * IH + i];
  __syncthreads();
  const int row = IW * C;
  const T* ib = img + static_cast<long long>(b) * IH * row;
  for (int i = threadIdx.x * 8; i < row; i += blockDim.x * 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int y = 0; y < IH; ++y) {
      const float w = a[y];
      if (w != 0.f) {
        float v[8];
        Vec8<T>::load(ib + static_cast<long long>(y) * row + i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, v[j], acc[j]);
      }
    }
    Vec8<float>::store(V + i, acc);
  }
  __syncthreads();
  const int c8n = C / 8;
  for (int g = threadIdx.x; g < W * c8n; g += blockDim.x) {
    const int w = g / c8n, c = (g % c8n) * 8;
    const int t = d * W + w;
    const int xi = x0[t];
    const float l = wl[t], r = wr[t];
    float vl[8], vr[8], o[8];
    Vec8<float>::load(V + xi * C + c, vl);
    Vec8<float>::load(V + (xi + 1) * C + c, vr);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = l * vl[j] + r * vr[j];
    Vec8<T>::store(out + ((static_cast<long long>(b) * W + w) * D + d) * C + c, o);
  }
}

// backward 1: dV[b, d, x, :] = sum_w (wl, wr)(d, w) * dout[b, w, d, :] scattered to x0(d, w), x0(d, w) + 1   (fp32 workspace)
template <typename T>
__global__ void __launch_bounds__(256) bev_lift_bwd_scatter_kernel(const T* __restrict__ dout, const int* __restrict__ x0,
                                                                   const float* __restrict__ wl,
                                                                   const float* __restrict__ wr, float* __restrict__ ws,
                                                                   int IW, int C, int D, int W) {
  extern __shared__ float sm[];  // dV[IW * C]
  const int d = blockIdx.x, b = blockIdx.y;
  const int row = IW * C;
  for (int i = threadIdx.x; i < row; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int c = threadIdx.x % C, wrow = threadIdx.x / C, wstep = blockDim.x / C;
  for (int w = wrow; w < W; w += wstep) {
    const int t = d * W + w;
    const float l = wl[t], r = wr[t];
    if (l == 0.f && r == 0.f) continue;
    const float g = ldf<T>(dout + ((static_cast<long long>(b) * W + w) * D + d) * C + c);
    const int xi = x0[t];
    if (l != 0.f) atomicAdd(sm + xi * C + c, l * g);
    if (r != 0.f) atomicAdd(sm + (xi + 1) * C + c, r * g);
  }
  __syncthreads();
  float* o = ws + (static_cast<long long>(b) * D + d) * row;
  for (int i = threadIdx.x; i < row; i += blockDim.x) o[i] = sm[i];
}

// backward 2: dimg[b, y, x, :] (+)= sum_d A[d, y] * dV[b, d, x, :].  grid (ceil(IW*C / 256), B); A in shared memory.
template <typename T, int IHMAX>
__global__ void __launch_bounds__(256) bev_lift_bwd_rows_kernel(const float* __restrict__ ws, const float* __restrict__ A,
                                                                T* __restrict__ dimg, int accumulate, int IH, int IW,
                                                                int C, int D) {
  extern __shared__ float sa[];  // A[D * IH]
  for (int i = threadIdx.x; i < D * IH; i += blockDim.x) sa[i] = A[i];
  __syncthreads();
  const int row = IW * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row) return;
  const int b = blockIdx.y;
  float acc[IHMAX];
#pragma unroll
  for (int y = 0; y < IHMAX; ++y) acc[y] = 0.f;
  const float* wb = ws + static_cast<long long>(b) * D * row + i;
  for (int d = 0; d < D; ++d) {
    const float v = wb[static_cast<long long>(d) * row];
    const float* ar = sa + d * IH;
#pragma unroll
    for (int y = 0; y < IHMAX; ++y)
      if (y < IH) acc[y] = fmaf(ar[y], v, acc[y]);
  }
  T* ob = dimg + static_cast<long long>(b) * IH * row + i;
#pragma unroll
  for (int y = 0; y < IHMAX; ++y) {
    if (y < IH) {
      float v = acc[y];
      if (accumulate) v += ldf<T>(ob + static_cast<long long>(y) * row);
      stf<T>(ob + static_cast<long long>(y) * row, v);
    }
  }
}

void chunking(int batch, int hw, int* chunks, int* ppb) {
  int ch = TFPP_NUM_SMS * 4 / batch;
  if (ch < 1) ch = 1;
  int p = (hw + ch - 1) / ch;
  if (p < 16) p = 16;
  *chunks = (hw + p - 1) / p;
  *ppb = p;
}

}  // namespace

#define STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int tfpp_instnorm_stats(const void* x, int f32, long long x_pix_stride, int batch, int hw, int channels,
                                   float* sum, float* sq, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 4 == 0 && channels <= 1024 && x_pix_stride >= channels && x_pix_stride % 4 == 0,
                 "instnorm: channels % 4 == 0, <= 1024; pixel stride a multiple of 4 elements");
  int chunks, ppb;
  chunking(batch, hw, &chunks, &ppb);
  const dim3 grid(chunks, batch);
  const size_t smem = sizeof(float) * 2 * channels;
  if (f32)
    instnorm_stats_kernel<float><<<grid, 256, smem, stream>>>(static_cast<const float*>(x), x_pix_stride, hw, channels, ppb,
                                                              sum, sq);
  else
    instnorm_stats_kernel<bf16><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(x), x_pix_stride, hw, channels, ppb,
                                                             sum, sq);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_instnorm_apply(const void* x, int f32, long long x_pix_stride, const float* sum, const float* sq,
                                   float eps, int act, void* y, long long y_pix_stride, float* mean, float* invstd,
                                   int batch, int hw, int channels, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 4 == 0 && channels <= 1024 && x_pix_stride % 4 == 0 && y_pix_stride % 4 == 0,
                 "instnorm: channels % 4 == 0, <= 1024; pixel strides multiples of 4 elements");
  TFPP_CHECK_ARG((mean == nullptr) == (invstd == nullptr), "mean and invstd go together");
  int chunks, ppb;
  chunking(batch, hw, &chunks, &ppb);
  const dim3 grid(chunks, batch);
  const size_t smem = sizeof(float) * 2 * channels;
  if (f32)
    instnorm_apply_kernel<float><<<grid, 256, smem, stream>>>(static_cast<const float*>(x), x_pix_stride, sum, sq, eps, act,
                                                              static_cast<float*>(y), y_pix_stride, mean, invstd, hw,
                                                              channels, ppb);
  else
    instnorm_apply_kernel<bf16><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(x), x_pix_stride, sum, sq, eps, act,
                                                             static_cast<bf16*>(y), y_pix_stride, mean, invstd, hw,
                                                             channels, ppb);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_instnorm_bwd(const void* dy, long long dy_pix_stride, const void* x, int f32, const float* mean,
                                 const float* invstd, int act, float* s1, float* s2, void* dx, int batch, int hw,
                                 int channels, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels % 4 == 0 && channels <= 1024 && dy_pix_stride % 4 == 0,
                 "instnorm: channels % 4 == 0, <= 1024; pixel stride a multiple of 4 elements");
  int chunks, ppb;
  chunking(batch, hw, &chunks, &ppb);
  const dim3 grid(chunks, batch);
  const size_t smem = sizeof(float) * 2 * channels;
  if (f32) {
    instnorm_bwd_reduce_kernel<float><<<grid, 256, smem, stream>>>(static_cast<const float*>(dy), dy_pix_stride,
                                                                   static_cast<const float*>(x), mean, invstd, act, s1, s2,
                                                                   hw, channels, ppb);
    TFPP_CHECK_LAUNCH();
    instnorm_bwd_apply_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(dy), dy_pix_stride,
                                                               static_cast<const float*>(x), mean, invstd, act, s1, s2,
                                                               static_cast<float*>(dx), hw, channels, ppb);
  } else {
    instnorm_bwd_reduce_kernel<bf16><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(dy), dy_pix_stride,
                                                                  static_cast<const bf16*>(x), mean, invstd, act, s1, s2,
                                                                  hw, channels, ppb);
    TFPP_CHECK_LAUNCH();
    instnorm_bwd_apply_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(dy), dy_pix_stride,
                                                              static_cast<const bf16*>(x), mean, invstd, act, s1, s2,
                                                              static_cast<bf16*>(dx), hw, channels, ppb);
  }
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bev_lift(const void* img, int f32, const float* a_rows, const int* x0, const float* wl,
                             const float* wr, void* out, int batch, int img_h, int img_w, int channels, int depth,
                             int width, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels >= 8 && channels <= 256 && 256 % channels == 0, "bev_lift: channels in {8, 16, 32, 64, 128, 256}");
  const size_t smem = sizeof(float) * (static_cast<size_t>(img_w) * channels + img_h);
  TFPP_CHECK_ARG(smem <= 48 * 1024, "bev_lift: one image row block must fit 48 KB of shared memory");
  const dim3 grid(depth, batch);
  if (f32)
    bev_lift_kernel<float><<<grid, 256, smem, stream>>>(static_cast<const float*>(img), a_rows, x0, wl, wr,
                                                        static_cast<float*>(out), img_h, img_w, channels, depth, width);
  else
    bev_lift_kernel<bf16><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(img), a_rows, x0, wl, wr,
                                                       static_cast<bf16*>(out), img_h, img_w, channels, depth, width);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}

extern "C" int tfpp_bev_lift_bwd(const void* dout, int f32, const float* a_rows, const int* x0, const float* wl,
                                 const float* wr, float* ws, void* dimg, int accumulate, int batch, int img_h, int img_w,
                                 int channels, int depth, int width, tfpp_stream_t stream_) {
  STREAM;
  TFPP_CHECK_ARG(channels >= 1 && channels <= 256 && 256 % channels == 0, "bev_lift: channels must divide 256");
  TFPP_CHECK_ARG(img_h <= 32, "bev_lift_bwd: at most 32 feature rows");
  const int row = img_w * channels;
  const size_t smem1 = sizeof(float) * row, smem2 = sizeof(float) * depth * img_h;
  TFPP_CHECK_ARG(smem1 <= 48 * 1024 && smem2 <= 48 * 1024, "bev_lift_bwd: shared-memory budget exceeded");
  const dim3 g1(depth, batch), g2((row + 255) / 256, batch);
  if (f32) {
    bev_lift_bwd_scatter_kernel<float><<<g1, 256, smem1, stream>>>(static_cast<const float*>(dout), x0, wl, wr, ws, img_w,
                                                                   channels, depth, width);
    TFPP_CHECK_LAUNCH();
    bev_lift_bwd_rows_kernel<float, 32><<<g2, 256, smem2, stream>>>(ws, a_rows, static_cast<float*>(dimg), accumulate,
                                                                    img_h, img_w, channels, depth);
  } else {
    bev_lift_bwd_scatter_kernel<bf16><<<g1, 256, smem1, stream>>>(static_cast<const bf16*>(dout), x0, wl, wr, ws, img_w,
                                                                  channels, depth, width);
    TFPP_CHECK_LAUNCH();
    bev_lift_bwd_rows_kernel<bf16, 32><<<g2, 256, smem2, stream>>>(ws, a_rows, static_cast<bf16*>(dimg), accumulate, img_h,
                                                                   img_w, channels, depth);
  }
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
