// Ensemble bounding-box merge of the agent: confidence threshold + image -> vehicle conversion +
// rotated-rectangle non-maximum suppression over the union of the ensemble members' detections, one CTA per frame.
//
// Reference: sensor_agent.py:445-491 collects convert_features_to_bb_metric() of every ensemble member
// (model.py:447-459: score > bb_confidence_threshold, transfuser_utils.bb_image_to_vehicle_system,
// transfuser_utils.py:388-406) and merges them with transfuser_utils.non_maximum_suppression
// (transfuser_utils.py:409-433): argsort by confidence, greedy keep, drop every remaining box whose rotated IoU with
// the kept one exceeds iou_treshold_nms; the IoU is shapely's polygon intersection / union of two rotated rectangles
// (transfuser_utils.py:436-452).  On the host that is an O(M^2) Python loop over shapely objects per frame.
//
// Here: block bitonic sort of (score, index), corners of the sorted boxes in shared memory, the M x M "suppresses" bit
// matrix computed in parallel (Sutherland-Hodgman clip of one convex quad by the other, registers only), then one warp
// walks the sorted list and ORs rows of kept boxes into the removed mask.
#include "../../include/tfpp.h"
#include "common.cuh"

namespace {

constexpr int kMaxBoxes = 512;   // 3-5 ensemble members x top-100 detections
constexpr int kWords = kMaxBoxes / 32;

struct Quad {
  float x[4], y[4];
};

// rectangle with HALF extents (w, h), rotated by yaw (radians, counter-clockwise), centred at (cx, cy)
// (transfuser_utils.rect_polygon); corners counter-clockwise
__device__ __forceinline__ Quad make_quad(float cx, float cy, float w, float h, float yaw) {
  float s, c;
  sincosf(yaw, &s, &c);
  const float px[4] = {-w, w, w, -w}, py[4] = {-h, -h, h, h};
  Quad q;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    q.x[i] = cx + c * px[i] - s * py[i];
    q.y[i] = cy + s * px[i] + c * py[i];
  }
  return q;
}

// area of (convex quad a) intersected with (convex, counter-clockwise quad b): clip a by the four edges of b
__device__ float intersection_area(const Quad& a, const Quad& b) {
  float px[8], py[8], qx[8], qy[8];
  int n = 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    px[i] = a.x[i];
    py[i] = a.y[i];
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float ax = b.x[e], ay = b.y[e], bx = b.x[(e + 1) & 3], by = b.y[(e + 1) & 3];
    const float ex = bx - ax, ey = by - ay;
    int m = 0;
    for (int j = 0; j < n; ++j) {
      const int jp = j == 0 ? n - 1 : j - 1;
      const float sc = ex * (py[j] - ay) - ey * (px[j] - ax);
      const float sp = ex * (py[jp] - ay) - ey * (px[jp] - ax);
      if (sc >= 0.f) {
        if (sp < 0.f) {
          const float t = sp / (sp - sc);
          qx[m] = px[jp] + t * (px[j] - px[jp]);
          qy[m] = py[jp] + t * (py[j] - py[jp]);
          ++m;
        }
        qx[m] = px[j];
        qy[m] = py[j];
        ++m;
      } else if (sp >= 0.f) {
        const float t = sp / (sp - sc);
        qx[m] = px[jp] + t * (px[j] - px[jp]);
        qy[m] = py[jp] + t * (py[j] - py[jp]);
        ++m;
      }
    }
    n = m;
    if (n == 0) return 0.f;
    for (int j = 0; j < n; ++j) {
      px[j] = qx[j];
      py[j] = qy[j];
    }
  }
  float a2 = 0.f;
  for (int j = 0; j < n; ++j) {
    const int jn = j + 1 == n ? 0 : j + 1;
    a2 += px[j] * py[jn] - py[j] * px[jn];
  }
  return 0.5f * fabsf(a2);
}

__global__ void __launch_bounds__(256) nms_rotated_kernel(const float* __restrict__ boxes, int M, int stride,
                                                          float conf_thr, float iou_thr, int to_vehicle, float ppm,
                                                          float min_x, float min_y, float* __restrict__ out,
                                                          int* __restrict__ out_count, int* __restrict__ out_index) {
  extern __shared__ __align__(16) unsigned char nms_smem[];   // 62 KB carved below (above the 48 KB static limit)
  float* skey = reinterpret_cast<float*>(nms_smem);                       // [512]
  int* sidx = reinterpret_cast<int*>(skey + kMaxBoxes);                   // [512]
  float(*sqx)[4] = reinterpret_cast<float(*)[4]>(sidx + kMaxBoxes);       // [512][4]
  float(*sqy)[4] = reinterpret_cast<float(*)[4]>(&sqx[kMaxBoxes][0]);     // [512][4]
  float* sarea = &sqy[kMaxBoxes][0];
  float* srad = sarea + kMaxBoxes;
  float* scx = srad + kMaxBoxes;
  float* scy = scx + kMaxBoxes;
  unsigned(*smask)[kWords] = reinterpret_cast<unsigned(*)[kWords]>(scy + kMaxBoxes);   // [512][16]
  int* skeep = reinterpret_cast<int*>(&smask[kMaxBoxes][0]);              // [512]
  __shared__ int s_nvalid, s_nkeep;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* bx = boxes + static_cast<long long>(b) * M * stride;
  if (tid == 0) s_nvalid = 0;
  __syncthreads();
  // 1. keys: score of the boxes above the confidence threshold (model.py:449), -inf otherwise
  for (int i = tid; i < kMaxBoxes; i += blockDim.x) {
    float key = -INFINITY;
    if (i < M) {
      const float sc = bx[i * stride + stride - 1];
      if (sc > conf_thr) {
        key = sc;
        atomicAdd(&s_nvalid, 1);
      }
    }
    skey[i] = key;
    sidx[i] = i;
  }
  __syncthreads();
  // 2. bitonic sort, descending by score; ties: lower original index first (deterministic)
  for (int k = 2; k <= kMaxBoxes; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < kMaxBoxes; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool desc = (i & k) == 0;
          const float ka = skey[i], kb = skey[l];
          const int ia = sidx[i], ib = sidx[l];
          const bool a_first = ka > kb || (ka == kb && ia < ib);   // "a belongs before b" in descending order
          if (desc ? !a_first : a_first) {
            skey[i] = kb; skey[l] = ka;
            sidx[i] = ib; sidx[l] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  const int nv = s_nvalid;
  // 3. geometry of the sorted boxes (vehicle frame if requested: transfuser_utils.py:388-406)
  for (int i = tid; i < nv; i += blockDim.x) {
    const float* p = bx + sidx[i] * stride;
    float x = p[0], y = p[1], w = p[2], h = p[3], yaw = p[4];
    if (to_vehicle) {
      yaw = -yaw;
      const float tx = x - (-(min_x * ppm)), ty = y - (-(min_y * ppm));
      x = ty / ppm;
      y = tx / ppm;
      const float w2 = h / ppm, h2 = w / ppm;
      w = w2;
      h = h2;
    }
    const Quad q = make_quad(x, y, w, h, yaw);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      sqx[i][c] = q.x[c];
      sqy[i][c] = q.y[c];
    }
    sarea[i] = 4.f * fabsf(w * h);
    srad[i] = sqrtf(w * w + h * h);
    scx[i] = x;
    scy[i] = y;
  }
  for (int i = tid; i < kMaxBoxes * kWords; i += blockDim.x) (&smask[0][0])[i] = 0u;
  __syncthreads();
  // 4. suppression bits: box i (higher score) suppresses box j > i when IoU > threshold
  const int words = (nv + 31) / 32;
  for (int t = tid; t < nv * words; t += blockDim.x) {
    const int i = t / words, wj = t % words;
    unsigned bits = 0u;
    Quad qi;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      qi.x[c] = sqx[i][c];
      qi.y[c] = sqy[i][c];
    }
    for (int bit = 0; bit < 32; ++bit) {
      const int j = wj * 32 + bit;
      if (j <= i || j >= nv) continue;
      const float dx = scx[i] - scx[j], dy = scy[i] - scy[j], rr = srad[i] + srad[j];
      if (dx * dx + dy * dy > rr * rr) continue;   // circumscribed circles apart: no overlap
      Quad qj;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        qj.x[c] = sqx[j][c];
        qj.y[c] = sqy[j][c];
      }
      const float inter = intersection_area(qi, qj);
      const float uni = sarea[i] + sarea[j] - inter;
      if (uni > 0.f && inter / uni > iou_thr) bits |= 1u << bit;
    }
    smask[i][wj] = bits;
  }
  __syncthreads();
  // 5. greedy walk (transfuser_utils.py:418-431) by one warp: lane w owns word w of the removed mask
  if (tid < 32) {
    unsigned removed = 0u;
    int nkeep = 0;
    for (int i = 0; i < nv; ++i) {
      const unsigned wi = __shfl_sync(0xffffffffu, removed, i >> 5);
      const bool gone = (wi >> (i & 31)) & 1u;
      if (!gone) {
        if (tid < kWords) removed |= smask[i][tid];
        if (tid == 0) skeep[nkeep] = i;
        ++nkeep;
      }
    }
    if (tid == 0) s_nkeep = nkeep;
  }
  __syncthreads();
  // 6. kept boxes, highest confidence first (converted when requested); rows >= count are zero
  const int nk = s_nkeep;
  if (tid == 0) out_count[b] = nk;
  float* ob = out + static_cast<long long>(b) * M * stride;
  for (int t = tid; t < M * stride; t += blockDim.x) {
    const int r = t / stride, c = t % stride;
    float v = 0.f;
    if (r < nk) {
      const int src = sidx[skeep[r]];
      const float* p = bx + src * stride;
      v = p[c];
      if (to_vehicle) {
        if (c == 0) v = (p[1] + min_y * ppm) / ppm;
        else if (c == 1) v = (p[0] + min_x * ppm) / ppm;
        else if (c == 2) v = p[3] / ppm;
        else if (c == 3) v = p[2] / ppm;
        else if (c == 4) v = -p[4];
      }
    }
    ob[t] = v;
  }
  if (out_index != nullptr)
    for (int r = tid; r < M; r += blockDim.x) out_index[static_cast<long long>(b) * M + r] = r < nk ? sidx[skeep[r]] : -1;
}

}  // namespace

extern "C" int tfpp_nms_rotated(const float* boxes, int batch, int num_boxes, int stride, float conf_threshold,
                                float iou_threshold, int to_vehicle, float pixels_per_meter, float min_x, float min_y,
                                float* out_boxes, int* out_count, int* out_index, tfpp_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TFPP_CHECK_ARG(boxes != nullptr && out_boxes != nullptr && out_count != nullptr, "null buffer");
  TFPP_CHECK_ARG(num_boxes >= 1 && num_boxes <= kMaxBoxes, "1..512 boxes per frame");
  TFPP_CHECK_ARG(stride >= 6, "a box is (x, y, w, h, yaw, ..., score)");
  TFPP_CHECK_ARG(pixels_per_meter > 0.f, "pixels_per_meter must be positive");
  if (batch <= 0) return TFPP_OK;
  constexpr size_t kSmem = sizeof(float) * kMaxBoxes * (2 + 8 + 4) + sizeof(unsigned) * kMaxBoxes * kWords +
                           sizeof(int) * kMaxBoxes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(nms_rotated_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) {
      tfpp_set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
      return TFPP_ERR_CUDA;
    }
    attr_set = true;
  }
  nms_rotated_kernel<<<batch, 256, kSmem, stream>>>(boxes, num_boxes, stride, conf_threshold, iou_threshold, to_vehicle,
                                                pixels_per_meter, min_x, min_y, out_boxes, out_count, out_index);
  TFPP_CHECK_LAUNCH();
  return TFPP_OK;
}
