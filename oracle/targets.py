"""TEST INFRASTRUCTURE (oracle): CPU restatement of the CenterNet training-target rasteriser of the data loader.

Follows team_code/data.py:698-791 (``CARLA_Data.get_targets``), team_code/gaussian_target.py:11-61
(``gaussian2d`` / ``gen_gaussian_target``), gaussian_target.py:160-183 (``gaussian_radius``) and
team_code/center_net.py:240-254 (``angle2class``).  Pinned by goldens produced by the unmodified reference functions
run in the build container (tests/golden/make_targets_golden.py -> tests/golden/targets.npz).

Only tests/, smoke() and bench.py's cpu_baseline may import this module."""
import math

import numpy as np


def gaussian_radius(det_size, min_overlap):
  """gaussian_target.py:160-183, evaluated in float64 on the float32 extents."""
  height, width = float(det_size[0]), float(det_size[1])
  b1 = height + width
  c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
  r1 = (b1 - math.sqrt(b1**2 - 4 * c1)) / 2
  b2 = 2 * (height + width)
  c2 = (1 - min_overlap) * width * height
  r2 = (b2 - math.sqrt(b2**2 - 16 * c2)) / 8
  a3 = 4 * min_overlap
  b3 = -2 * min_overlap * (height + width)
  c3 = (min_overlap - 1) * width * height
  r3 = (b3 + math.sqrt(b3**2 - 4 * a3 * c3)) / (2 * a3)
  return min(r1, r2, r3)


def gaussian2d(radius):
  """gaussian_target.py:11-31 with sigma = (2 radius + 1) / 6, float32 like the heat map."""
  sigma = (2 * radius + 1) / 6
  x = np.arange(-radius, radius + 1, dtype=np.float32).reshape(1, -1)
  y = np.arange(-radius, radius + 1, dtype=np.float32).reshape(-1, 1)
  h = np.exp(-(x * x + y * y) / np.float32(2 * sigma * sigma))
  h[h < np.finfo(np.float32).eps * h.max()] = 0
  return h


def angle2class(angle, num_dir_bins):
  """center_net.py:240-254."""
  angle = angle % (2 * np.pi)
  per = 2 * np.pi / float(num_dir_bins)
  shifted = (angle + per / 2) % (2 * np.pi)
  cls = shifted // per
  return int(cls), shifted - (cls * per + per / 2)


def get_targets(gt_bboxes, feat_h=64, feat_w=64, img_h=256, img_w=256, num_classes=4, num_dir_bins=12):
  """data.py:698-791 for one sample.  gt_bboxes (N, 8) float32: x, y, extent_x, extent_y, yaw, speed, brake, class in
  BEV pixel coordinates.  Returns (dict of arrays, avg_factor)."""
  gt_bboxes = np.asarray(gt_bboxes, dtype=np.float32).reshape(-1, 8)
  wr, hr = float(feat_w / img_w), float(feat_h / img_h)
  t = {'center_heatmap_target': np.zeros([num_classes, feat_h, feat_w], np.float32),
       'wh_target': np.zeros([2, feat_h, feat_w], np.float32), 'offset_target': np.zeros([2, feat_h, feat_w], np.float32),
       'yaw_class_target': np.zeros([feat_h, feat_w], np.int32), 'yaw_res_target': np.zeros([1, feat_h, feat_w], np.float32),
       'velocity_target': np.zeros([1, feat_h, feat_w], np.float32), 'brake_target': np.zeros([feat_h, feat_w], np.int32),
       'pixel_weight': np.zeros([2, feat_h, feat_w], np.float32)}
  if gt_bboxes.shape[0] == 0:
    return t, 1
  for j in range(gt_bboxes.shape[0]):
    ctx, cty = np.float32(gt_bboxes[j, 0] * np.float32(wr)), np.float32(gt_bboxes[j, 1] * np.float32(hr))
    cx, cy = int(ctx), int(cty)
    ex, ey = np.float32(gt_bboxes[j, 2] * np.float32(wr)), np.float32(gt_bboxes[j, 3] * np.float32(hr))
    radius = max(2, int(gaussian_radius([ey, ex], 0.1)))
    heat = t['center_heatmap_target'][int(gt_bboxes[j, 7])]
    g = gaussian2d(radius)
    left, right = min(cx, radius), min(feat_w - cx, radius + 1)
    top, bottom = min(cy, radius), min(feat_h - cy, radius + 1)
    np.maximum(heat[cy - top:cy + bottom, cx - left:cx + right], g[radius - top:radius + bottom, radius - left:radius + right],
               out=heat[cy - top:cy + bottom, cx - left:cx + right])
    t['wh_target'][0, cy, cx], t['wh_target'][1, cy, cx] = ex, ey
    ycls, yres = angle2class(gt_bboxes[j, 4], num_dir_bins)
    t['yaw_class_target'][cy, cx] = ycls
    t['yaw_res_target'][0, cy, cx] = yres
    t['velocity_target'][0, cy, cx] = gt_bboxes[j, 5]
    t['brake_target'][cy, cx] = int(round(float(gt_bboxes[j, 6])))
    t['offset_target'][0, cy, cx] = ctx - np.float32(cx)
    t['offset_target'][1, cy, cx] = cty - np.float32(cy)
    t['pixel_weight'][:, cy, cx] = 1.0
  return t, max(1, int(np.equal(t['center_heatmap_target'], 1).sum()))
