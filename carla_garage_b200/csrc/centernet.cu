// CenterNet heat-map decode (K20).  Reference: LidarCenterNetHead.decode_heatmap, team_code/center_net.py:172-237,
// with get_local_maximum / get_topk_from_heatmap / transpose_and_gather_feat (team_code/gaussian_target.py:186-264)
// and class2angle (center_net.py:125-140).  One CTA per sample: 3x3 peak keep, block-wide bitonic sort of the
// classes*H*W (score, index) pairs in shared memory, gather + angle decode of the first k.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(1024) decode_heatmap_kernel(const float* __restrict__ heat, long long heat_sb,
                                                              const float* __restrict__ wh, long long wh_sb,
                                                              const float* __restrict__ offset, long long off_sb,
                                                              const float* __restrict__ yaw_cls, long long ycls_sb,
                                                              const float* __restrict__ yaw_res, long long yres_sb,
                                                              float* __restrict__ out, int n_cls, int H, int W,
                                                              int n_bins, int k, int npad, float width_ratio,
                                                              float height_ratio) {
  extern __shared__ __align__(8) uint8_t dec_sm[];
  float* score = reinterpret_cast<float*>(dec_sm);
  int* index = reinterpret_cast<int*>(score + npad);
  const int b = blockIdx.x;
  const int n = n_cls * H * W;
  const float* hp = heat + b * heat_sb;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    float s = -INFINITY;  // padding sorts last
    if (i < n) {
      const int x = i % W, y = (i / W) % H, c = i / (W * H);
      const float v = hp[i];
      float m = -INFINITY;  // F.max_pool2d pads with -inf (gaussian_target.py:197-198)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) m = fmaxf(m, hp[(c * H + yy) * W + xx]);
        }
      s = (m == v) ? v : 0.f;  // heat * keep (gaussian_target.py:199-200)
    }
    score[i] = s;
    index[i] = i;
  }
  __syncthreads();
  // bitonic sort, descending by score, ascending by index on ties (deterministic; torch.topk leaves ties unspecified)
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < npad / 2; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const float sl = score[lo], sh = score[hi];
        const int il = index[lo], ih = index[hi];
        const bool lo_first = (sl > sh) || (sl == sh && il < ih);  // lo should precede hi in descending order
        if (lo_first != desc) {
          score[lo] = sh; score[hi] = sl;
          index[lo] = ih; index[hi] = il;
        }
      }
      __syncthreads();
    }
  }
  const int hw = H * W;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int idx = index[j];
    const int cls = idx / hw;                 // center_net get_topk: topk_inds // (H*W)
    const int pix = idx % hw;
    const int y = pix / W, x = pix % W;
    const float w_ = wh[b * wh_sb + pix], h_ = wh[b * wh_sb + hw + pix];
    const float ox = offset[b * off_sb + pix], oy = offset[b * off_sb + hw + pix];
    int best = 0;
    float bv = yaw_cls[b * ycls_sb + pix];
    for (int c = 1; c < n_bins; ++c) {
      const float v = yaw_cls[b * ycls_sb + static_cast<long long>(c) * hw + pix];
      if (v > bv) { bv = v; best = c; }      // torch.argmax: first maximal index
    }
    const float per = 6.283185307179586f / static_cast<float>(n_bins);
    float yaw = static_cast<float>(best) * per + yaw_res[b * yres_sb + pix];
    if (yaw > 3.141592653589793f) yaw -= 6.283185307179586f;
    float* o = out + (static_cast<long long>(b) * k + j) * 9;
    o[0] = (static_cast<float>(x) + ox) * width_ratio;
    o[1] = (static_cast<float>(y) + oy) * height_ratio;
    o[2] = w_ * width_ratio;
    o[3] = h_ * height_ratio;
    o[4] = yaw;
    o[5] = 0.f;  // velocity / brake heads do not exist in single-frame mode (center_net.py:223-225)
    o[6] = 0.f;
    o[7] = static_cast<float>(cls);
    o[8] = score[j];
  }
}

}  // namespace

extern "C" int tfpp_decode_heatmap(const float* heat, long long heat_sb, const float* wh, long long wh_sb,
                                   const float* offset, long long off_sb, const float* yaw_cls, long long ycls_sb,
                                   const float* yaw_res, long long yres_sb, float* out, int batch, int n_cls, int height,
                                   int width, int n_bins, int k, float width_ratio, float height_ratio,
                                   tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n = n_cls * height * width;
  int npad = 1;
  while (npad < n) npad <<= 1;
  TFPP_CHECK_ARG(npad * 8 <= 200 * 1024, "heat map too large for the shared-memory sort");
  TFPP_CHECK_ARG(k <= n, "k larger than the heat map");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(decode_heatmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr_set = true;
  }
  decode_heatmap_kernel<<<batch, 1024, static_cast<size_t>(npad) * 8, stream>>>(
      heat, heat_sb, wh, wh_sb, offset, off_sb, yaw_cls, ycls_sb, yaw_res, yres_sb, out, n_cls, height, width, n_bins, k,
      npad, width_ratio, height_ratio);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
