// CenterNet training-target rasteriser on the device (the data loader's label side).
//
// Reference: CARLA_Data.get_targets (team_code/data.py:698-791) runs per sample inside the DataLoader workers: for
// every ground-truth box a Gaussian blob (gaussian_target.py:11-61, radius from gaussian_radius,
// gaussian_target.py:160-183, min_overlap 0.1, at least 2) is max-merged into the heat map of the box's class, and the
// extent / yaw bin + residual (center_net.py:240-254) / velocity / brake / sub-pixel offset / weight targets are written
// at the box's centre pixel (a later box overwrites an earlier one on the same pixel).  avg_factor = max(1, number of
// heat-map pixels equal to 1).  Here one CTA rasterises one sample straight into the (B, ...) label tensors the loss
// kernels read, so the labels never cross PCIe as dense maps: 30 boxes x 32 B instead of 13 x 64 x 64 floats.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

constexpr int kMaxBoxes = 128;

__device__ __forceinline__ double gaussian_radius_d(double height, double width, double min_overlap) {
  const double b1 = height + width;
  const double c1 = width * height * (1.0 - min_overlap) / (1.0 + min_overlap);
  const double r1 = (b1 - sqrt(b1 * b1 - 4.0 * c1)) / 2.0;
  const double b2 = 2.0 * (height + width);
  const double c2 = (1.0 - min_overlap) * width * height;
  const double r2 = (b2 - sqrt(b2 * b2 - 16.0 * c2)) / 8.0;
  const double a3 = 4.0 * min_overlap;
  const double b3 = -2.0 * min_overlap * (height + width);
  const double c3 = (min_overlap - 1.0) * width * height;
  const double r3 = (b3 + sqrt(b3 * b3 - 4.0 * a3 * c3)) / (2.0 * a3);
  return fmin(r1, fmin(r2, r3));
}

// numpy's float32 `a % b` for b > 0: fmodf, shifted into [0, b)
__device__ __forceinline__ float mod_pos(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && m < 0.f) m += b;
  return m;
}

__global__ void __launch_bounds__(256) centernet_targets_kernel(
    const float* __restrict__ boxes, const int* __restrict__ counts, int max_boxes, int H, int W, float wr, float hr,
    int C, int bins, float* __restrict__ heat, float* __restrict__ wh, float* __restrict__ offset,
    long long* __restrict__ yaw_class, float* __restrict__ yaw_res, float* __restrict__ velocity,
    long long* __restrict__ brake, float* __restrict__ pixel_weight, float* __restrict__ avg_factor) {
  __shared__ int s_cx[kMaxBoxes], s_cy[kMaxBoxes], s_r[kMaxBoxes], s_cls[kMaxBoxes];
  __shared__ int s_count;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int hw = H * W;
  int n = counts ? counts[b] : max_boxes;
  n = max(0, min(n, min(max_boxes, kMaxBoxes)));
  const float* bx = boxes + static_cast<long long>(b) * max_boxes * 8;
  float* heat_b = heat + static_cast<long long>(b) * C * hw;
  float* wh_b = wh + static_cast<long long>(b) * 2 * hw;
  float* off_b = offset + static_cast<long long>(b) * 2 * hw;
  float* pw_b = pixel_weight + static_cast<long long>(b) * 2 * hw;
  // 0. clear this sample's maps
  for (int i = tid; i < C * hw; i += blockDim.x) heat_b[i] = 0.f;
  for (int i = tid; i < 2 * hw; i += blockDim.x) {
    wh_b[i] = 0.f;
    off_b[i] = 0.f;
    pw_b[i] = 0.f;
  }
  for (int i = tid; i < hw; i += blockDim.x) {
    yaw_class[static_cast<long long>(b) * hw + i] = 0;
    yaw_res[static_cast<long long>(b) * hw + i] = 0.f;
    if (velocity) velocity[static_cast<long long>(b) * hw + i] = 0.f;
    if (brake) brake[static_cast<long long>(b) * hw + i] = 0;
  }
  if (tid == 0) s_count = 0;
  // 1. per-box geometry (data.py:745-760)
  for (int j = tid; j < n; j += blockDim.x) {
    const float* p = bx + j * 8;
    const float ctx = p[0] * wr, cty = p[1] * hr;
    const int cx = static_cast<int>(ctx), cy = static_cast<int>(cty);   // astype(int): truncation
    const float ex = p[2] * wr, ey = p[3] * hr;
    const int cls = static_cast<int>(p[7]);
    const bool ok = cx >= 0 && cx < W && cy >= 0 && cy < H && cls >= 0 && cls < C;   // the reference would raise
    int r = static_cast<int>(gaussian_radius_d(static_cast<double>(ey), static_cast<double>(ex), 0.1));
    r = max(2, r);
    s_cx[j] = cx;
    s_cy[j] = cy;
    s_r[j] = ok ? r : -1;
    s_cls[j] = cls;
  }
  __syncthreads();
  // 2. Gaussian blobs, max-merged (gaussian_target.py:34-61): non-negative floats order like their bit patterns
  for (int j = 0; j < n; ++j) {
    const int r = s_r[j];
    if (r < 0) continue;
    const int d = 2 * r + 1;
    const double sigma = static_cast<double>(d) / 6.0;
    const float two_s2 = static_cast<float>(2.0 * sigma * sigma);
    int* hp = reinterpret_cast<int*>(heat_b + static_cast<long long>(s_cls[j]) * hw);
    for (int t = tid; t < d * d; t += blockDim.x) {
      const int dy = t / d - r, dx = t % d - r;
      const int y = s_cy[j] + dy, x = s_cx[j] + dx;
      if (y < 0 || y >= H || x < 0 || x >= W) continue;
      const float r2 = static_cast<float>(dx * dx + dy * dy);
      float v = expf(-r2 / two_s2);
      if (v < 1.1920929e-07f) v = 0.f;   // h[h < eps * h.max()] = 0, h.max() == 1
      atomicMax(hp + y * W + x, __float_as_int(v));
    }
  }
  // 3. centre-pixel targets: the LAST box on a pixel wins (data.py:762-784 is a sequential loop)
  for (int j = tid; j < n; j += blockDim.x) {
    if (s_r[j] < 0) continue;
    bool last = true;
    for (int k = j + 1; k < n; ++k)
      if (s_r[k] >= 0 && s_cx[k] == s_cx[j] && s_cy[k] == s_cy[j]) last = false;
    if (!last) continue;
    const float* p = bx + j * 8;
    const int pix = s_cy[j] * W + s_cx[j];
    const float ctx = p[0] * wr, cty = p[1] * hr;
    wh_b[pix] = p[2] * wr;
    wh_b[hw + pix] = p[3] * hr;
    // angle2class (center_net.py:240-254) in float32 like numpy on a float32 angle
    const float two_pi = 6.283185307179586f;
    const float per = static_cast<float>(6.283185307179586 / static_cast<double>(bins));
    const float half = static_cast<float>(6.283185307179586 / static_cast<double>(bins) / 2.0);
    const float ang = mod_pos(p[4], two_pi);
    const float shifted = mod_pos(ang + half, two_pi);
    const float m = fmodf(shifted, per);
    const float div = (shifted - m) / per;
    float fl = floorf(div);
    if (div - fl > 0.5f) fl += 1.f;
    yaw_class[static_cast<long long>(b) * hw + pix] = static_cast<long long>(fl);
    yaw_res[static_cast<long long>(b) * hw + pix] = shifted - (fl * per + half);
    if (velocity) velocity[static_cast<long long>(b) * hw + pix] = p[5];
    if (brake) brake[static_cast<long long>(b) * hw + pix] = static_cast<long long>(__float2int_rn(p[6]));
    off_b[pix] = ctx - static_cast<float>(s_cx[j]);
    off_b[hw + pix] = cty - static_cast<float>(s_cy[j]);
    pw_b[pix] = 1.f;
    pw_b[hw + pix] = 1.f;
  }
  __syncthreads();
  // 4. avg_factor = max(1, #(heat == 1)) (data.py:786)
  int cnt = 0;
  for (int i = tid; i < C * hw; i += blockDim.x) cnt += heat_b[i] == 1.f ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0 && cnt) atomicAdd(&s_count, cnt);
  __syncthreads();
  if (tid == 0) avg_factor[b] = static_cast<float>(max(1, s_count));
}

}  // namespace

extern "C" int tfpp_centernet_targets(const float* boxes, const int* counts, int batch, int max_boxes, int feat_h,
                                      int feat_w, int img_h, int img_w, int num_classes, int num_dir_bins,
                                      float* center_heatmap, float* wh, float* offset, long long* yaw_class,
                                      float* yaw_res, float* velocity, long long* brake, float* pixel_weight,
                                      float* avg_factor, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(boxes != nullptr && center_heatmap != nullptr && wh != nullptr && offset != nullptr && yaw_class != nullptr &&
                     yaw_res != nullptr && pixel_weight != nullptr && avg_factor != nullptr,
                 "null buffer");
  TFPP_CHECK_ARG(max_boxes >= 1 && max_boxes <= kMaxBoxes, "1..128 boxes per sample");
  TFPP_CHECK_ARG(feat_h > 0 && feat_w > 0 && img_h > 0 && img_w > 0 && num_classes > 0 && num_dir_bins > 0, "bad sizes");
  if (batch <= 0) return TFPP_OK;
  const float wr = static_cast<float>(static_cast<double>(feat_w) / img_w);
  const float hr = static_cast<float>(static_cast<double>(feat_h) / img_h);
  centernet_targets_kernel<<<batch, 256, 0, stream>>>(boxes, counts, max_boxes, feat_h, feat_w, wr, hr, num_classes,
                                                      num_dir_bins, center_heatmap, wh, offset, yaw_class, yaw_res,
                                                      velocity, brake, pixel_weight, avg_factor);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
