"""TEST INFRASTRUCTURE (oracle): CPU restatement of the ensemble's bounding-box merge.

Follows team_code/transfuser_utils.py:409-450 (``non_maximum_suppression``, ``rect_polygon``, ``iou_bbs``) and the
image -> vehicle conversion of team_code/transfuser_utils.py:388-406 (``bb_image_to_vehicle_system``) as used by
model.py:447-459 (``convert_features_to_bb_metric``).  The reference computes the rotated-rectangle IoU with shapely
(``Polygon.intersection(...).area`` / ``union(...).area``); shapely is a third-party dependency that is absent from
/root/reference and from this image (the reference environment pins none in a lock file), so **parity is unpinned by
the live reference**: its published contract — exact area of the intersection of two convex polygons — is restated here
with Sutherland-Hodgman clipping in float64 and pinned by closed-form cases (tests/test_oracle.py::test_nms_oracle_*).

Only tests/, smoke() and bench.py's cpu_baseline may import this module."""
import numpy as np


def rect_corners(x, y, width, height, angle):
  """transfuser_utils.py:436-443: rectangle (+-width, +-height) (HALF extents) rotated by ``angle`` radians counter-
  clockwise about its centre, then translated to (x, y).  Returns (4, 2) float64, counter-clockwise."""
  c, s = np.cos(angle), np.sin(angle)
  pts = np.array([(-width, -height), (width, -height), (width, height), (-width, height)], dtype=np.float64)
  rot = np.array([[c, -s], [s, c]])
  return pts @ rot.T + np.array([x, y], dtype=np.float64)


def polygon_area(p):
  if len(p) < 3:
    return 0.0
  x, y = p[:, 0], p[:, 1]
  return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def clip_convex(subject, clip):
  """Sutherland-Hodgman: the part of convex polygon ``subject`` inside convex, counter-clockwise ``clip``."""
  out = [tuple(p) for p in subject]
  n = len(clip)
  for i in range(n):
    a, b = clip[i], clip[(i + 1) % n]
    inp, out = out, []
    if not inp:
      break

    def side(p):
      return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])

    for j, cur in enumerate(inp):
      prev = inp[j - 1]
      sc, sp = side(cur), side(prev)
      if sc >= 0:
        if sp < 0:
          t = sp / (sp - sc)
          out.append((prev[0] + t * (cur[0] - prev[0]), prev[1] + t * (cur[1] - prev[1])))
        out.append(cur)
      elif sp >= 0:
        t = sp / (sp - sc)
        out.append((prev[0] + t * (cur[0] - prev[0]), prev[1] + t * (cur[1] - prev[1])))
  return np.array(out, dtype=np.float64).reshape(-1, 2)


def iou_bbs(bb1, bb2):
  """transfuser_utils.py:446-452 (boxes are (x, y, half width, half height, yaw, ...))."""
  a = rect_corners(*[float(v) for v in bb1[:5]])
  b = rect_corners(*[float(v) for v in bb2[:5]])
  inter = polygon_area(clip_convex(a, b))
  union = polygon_area(a) + polygon_area(b) - inter
  return inter / union if union > 0 else 0.0


def non_maximum_suppression(bounding_boxes, iou_threshold):
  """transfuser_utils.py:409-433.  bounding_boxes: list (one entry per ensemble member) of lists of boxes; returns the
  kept boxes, highest confidence (last column) first."""
  boxes = [np.asarray(b, dtype=np.float64) for member in bounding_boxes if member is not None for b in member]
  if not boxes:
    return []
  boxes = np.stack(boxes)
  order = list(np.argsort(boxes[:, -1], kind='stable'))
  kept = []
  while order:
    idx = order.pop()
    kept.append(boxes[idx])
    order = [j for j in order if not iou_bbs(boxes[idx], boxes[j]) > iou_threshold]
  return kept


def bb_image_to_vehicle_system(box, pixels_per_meter, min_x, min_y):
  """transfuser_utils.py:388-406."""
  box = np.array(box, dtype=np.float64)
  box[4] = -box[4]
  box[:2] = box[:2] - np.array([-(min_x * pixels_per_meter), -(min_y * pixels_per_meter)])
  box[0], box[1] = box[1], box[0]
  box[2], box[3] = box[3], box[2]
  box[:4] = box[:4] / pixels_per_meter
  return box


def ensemble_boxes(decoded, conf_threshold, iou_threshold, pixels_per_meter=4.0, min_x=-32.0, min_y=-32.0):
  """sensor_agent.py:445-491 for one frame: ``decoded`` = list over ensemble members of (K, 9) decoded boxes in image
  coordinates (center_net.py:172-237); threshold (model.py:449), convert (model.py:451-457), NMS over the union."""
  members = []
  for d in decoded:
    d = np.asarray(d, dtype=np.float64)
    members.append([bb_image_to_vehicle_system(b, pixels_per_meter, min_x, min_y) for b in d[d[:, -1] > conf_threshold]])
  return non_maximum_suppression(members, iou_threshold)
