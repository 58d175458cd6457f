// K1 — LiDAR point cloud -> BEV occupancy histogram ("pillar scatter").
//
// Reference: CARLA_Data.lidar_to_histogram_features, team_code/data.py:873-906 (two np.histogramdd calls on the
// CPU, clip at hist_max_per_pixel, divide, transpose).  Here: one thread per point, 12-byte point loads, u32
// atomicAdd into a (B,2,H,W) count grid, then a clip/scale pass that writes the [ch][y_bin][x_bin] f32 image.
// Integer-exact: counts are integers, outputs are k/5.  HBM-bound: 12 B per point in, 4 B x H x W per channel out.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

struct ScatterCfg {
  float min_x, max_x, min_y, max_y, ppm, split_z, max_z;
  double split_z_d;   // the split height as the float64 the reference's config holds (aligned path)
  int nx, ny;
};

__global__ void __launch_bounds__(256) pillar_count_kernel(const float* __restrict__ pts, long long total,
                                                           int n_points, uint32_t* __restrict__ counts, ScatterCfg c) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const float x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
    // data.py:896-898: drop z >= max_height; "below" is z <= split (compared in the cloud's dtype, f32 here)
    if (!(z < c.max_z)) continue;
    // np.histogramdd semantics (data.py:887): half-open bins [e_i, e_{i+1}), last bin closed; NaN never counted.
    if (!(x >= c.min_x && x <= c.max_x && y >= c.min_y && y <= c.max_y)) continue;
    // edges are exact multiples of 1/ppm (0.25): floor(x * ppm) is exact in f32, no rounding at bin borders
    int bx = static_cast<int>(floorf(x * c.ppm)) - static_cast<int>(c.min_x * c.ppm);
    int by = static_cast<int>(floorf(y * c.ppm)) - static_cast<int>(c.min_y * c.ppm);
    bx = min(bx, c.nx - 1);
    by = min(by, c.ny - 1);
    const int ch = (z <= c.split_z) ? 0 : 1;
    const int b = static_cast<int>(i / n_points);
    // stored transposed already: [b][ch][y_bin][x_bin]  (data.py:893 `.T`)
    atomicAdd(counts + ((static_cast<long long>(b) * 2 + ch) * c.ny + by) * c.nx + bx, 1u);
  }
}

// K1 with the data loader's / agent's rigid alignment fused in: CARLA_Data.align (data.py:840-871) and
// the agent's half-sweep merge (sensor_agent.py:381-425) move a past sweep into the current vehicle frame with
// transfuser_utils.algin_lidar (transfuser_utils.py:116-130): p' = R(yaw)^T (p - t), once for the ego motion and once
// more for the augmentation.  numpy does that in float64 (the translation is a float64 array) and histogramdd then bins
// the float64 coordinates, so this variant transforms and compares in double: same bins, and z is compared with the
// float64 thresholds (an f32 0.2 is ABOVE the float64 0.2, unlike in the unaligned f32 path).
// xf: per sample n_xf x {tx, ty, tz, yaw} doubles, applied in order.
__global__ void __launch_bounds__(256) pillar_count_aligned_kernel(const float* __restrict__ pts, const double* __restrict__ xf,
                                                                   int n_xf, long long total, int n_points,
                                                                   uint32_t* __restrict__ counts, ScatterCfg c) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int b = static_cast<int>(i / n_points);
    double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    for (int k = 0; k < n_xf; ++k) {
      const double* t = xf + (static_cast<long long>(b) * n_xf + k) * 4;
      const double cs = cos(t[3]), sn = sin(t[3]);
      const double dx = x - t[0], dy = y - t[1];
      x = cs * dx + sn * dy;     // rows of R^T
      y = -sn * dx + cs * dy;
      z = z - t[2];
    }
    if (!(z < static_cast<double>(c.max_z))) continue;
    const double min_x = c.min_x, max_x = c.max_x, min_y = c.min_y, max_y = c.max_y, ppm = c.ppm;
    if (!(x >= min_x && x <= max_x && y >= min_y && y <= max_y)) continue;
    int bx = static_cast<int>(floor(x * ppm)) - static_cast<int>(c.min_x * c.ppm);
    int by = static_cast<int>(floor(y * ppm)) - static_cast<int>(c.min_y * c.ppm);
    bx = min(bx, c.nx - 1);
    by = min(by, c.ny - 1);
    const int ch = (z <= static_cast<double>(c.split_z_d)) ? 0 : 1;
    atomicAdd(counts + ((static_cast<long long>(b) * 2 + ch) * c.ny + by) * c.nx + bx, 1u);
  }
}

__global__ void __launch_bounds__(256) pillar_finalize_kernel(const uint32_t* __restrict__ counts,
                                                              float* __restrict__ out, long long total, int plane,
                                                              int use_ground_plane, int hist_max) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n_ch = use_ground_plane ? 2 : 1;
  const long long b = i / (static_cast<long long>(n_ch) * plane);
  const long long rem = i - b * n_ch * plane;
  const int ch = static_cast<int>(rem / plane);
  const long long pix = rem - static_cast<long long>(ch) * plane;
  const int src_ch = use_ground_plane ? ch : 1;  // data.py:901-904: [below, above] or [above]
  uint32_t cnt = counts[(b * 2 + src_ch) * plane + pix];
  cnt = cnt > static_cast<uint32_t>(hist_max) ? static_cast<uint32_t>(hist_max) : cnt;
  // float64 division then cast to f32 in the reference (data.py:889,905); k/5 rounds identically via f32 division
  out[i] = static_cast<float>(static_cast<double>(cnt) / static_cast<double>(hist_max));
}

}  // namespace

static int pillar_scatter_impl(const float* points, const double* xform, int n_xforms, double split_z_d, int batch,
                               int n_points, uint32_t* counts, float* out, int use_ground_plane, float min_x, float max_x,
                               float min_y, float max_y, float pixels_per_meter, int hist_max, float split_z, float max_z,
                               tfpp_stream_t stream_);

extern "C" int tfpp_pillar_scatter(const float* points, int batch, int n_points, uint32_t* counts, float* out,
                                   int use_ground_plane, float min_x, float max_x, float min_y, float max_y,
                                   float pixels_per_meter, int hist_max, float split_z, float max_z,
                                   tfpp_stream_t stream_) {
  return pillar_scatter_impl(points, nullptr, 0, split_z, batch, n_points, counts, out, use_ground_plane, min_x, max_x, min_y,
                             max_y, pixels_per_meter, hist_max, split_z, max_z, stream_);
}

extern "C" int tfpp_pillar_scatter_aligned(const float* points, const double* xform, int n_xforms, int batch, int n_points,
                                           uint32_t* counts, float* out, int use_ground_plane, float min_x, float max_x,
                                           float min_y, float max_y, float pixels_per_meter, int hist_max, double split_z,
                                           float max_z, tfpp_stream_t stream_) {
  TFPP_CHECK_ARG(xform != nullptr && n_xforms >= 1 && n_xforms <= 4, "1..4 rigid transforms per sample");
  return pillar_scatter_impl(points, xform, n_xforms, split_z, batch, n_points, counts, out, use_ground_plane, min_x, max_x,
                             min_y, max_y, pixels_per_meter, hist_max, static_cast<float>(split_z), max_z, stream_);
}

static int pillar_scatter_impl(const float* points, const double* xform, int n_xforms, double split_z_d, int batch,
                               int n_points, uint32_t* counts, float* out, int use_ground_plane, float min_x, float max_x,
                               float min_y, float max_y, float pixels_per_meter, int hist_max, float split_z, float max_z,
                               tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(batch >= 0 && n_points >= 0, "negative sizes");
  TFPP_CHECK_ARG(counts != nullptr && out != nullptr, "null output");
  TFPP_CHECK_ARG(hist_max > 0 && pixels_per_meter > 0, "bad histogram config");
  ScatterCfg c;
  c.min_x = min_x; c.max_x = max_x; c.min_y = min_y; c.max_y = max_y;
  c.ppm = pixels_per_meter; c.split_z = split_z; c.max_z = max_z; c.split_z_d = split_z_d;
  c.nx = static_cast<int>((max_x - min_x) * pixels_per_meter);
  c.ny = static_cast<int>((max_y - min_y) * pixels_per_meter);
  const int plane = c.nx * c.ny;
  if (batch == 0) return TFPP_OK;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(uint32_t) * 2ull * plane * batch, stream);
  if (e != cudaSuccess) {
    tfpp_set_error("memset: %s", cudaGetErrorString(e));
    return TFPP_ERR_CUDA;
  }
  const long long total = static_cast<long long>(batch) * n_points;
  if (total > 0) {
    TFPP_CHECK_ARG(points != nullptr, "null points");
    long long blocks = ceil_div_ll(total, 256);
    const long long cap = static_cast<long long>(TFPP_NUM_SMS) * 16;
    if (blocks > cap) blocks = cap;
    if (xform != nullptr)
      pillar_count_aligned_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(points, xform, n_xforms, total, n_points, counts, c);
    else
      pillar_count_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(points, total, n_points, counts, c);
    TFPP_CHECK_LAUNCH();
  }
  const long long out_total = static_cast<long long>(batch) * (use_ground_plane ? 2 : 1) * plane;
  pillar_finalize_kernel<<<static_cast<int>(ceil_div_ll(out_total, 256)), 256, 0, stream>>>(
      counts, out, out_total, plane, use_ground_plane, hist_max);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
