"""Inference-side helpers: CUDA-graph replay of the eval forward (the 20 Hz agent loop, sensor_agent.py:343-615, is
launch-latency bound at B=1), the ensemble reduction of SensorAgent.run_step (sensor_agent.py:445-485,527-552) and the
ensemble's bounding-box merge (rotated-IoU NMS, transfuser_utils.py:409-452) on the device."""
import numpy as np
import torch

from . import ops


class GraphedForward:
  """Captures ``model.eval()`` forward for one batch size into a CUDA graph with static input buffers.
  Call it like the module: ``outs = gf(rgb=..., lidar_bev=..., target_point=..., ego_vel=..., command=...)``;
  the returned tensors are the graph's static outputs (overwritten by the next call)."""

  def __init__(self, model, example_inputs):
    self.model = model.eval()
    self.static = {k: v.clone().cuda() for k, v in example_inputs.items()}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
      for _ in range(2):
        self.model(**self.static)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph), torch.no_grad():
      self.out = self.model(**self.static)

  def __call__(self, **inputs):
    for k, v in inputs.items():
      self.static[k].copy_(v, non_blocking=True)
    self.graph.replay()
    return self.out


def ensemble_outputs(outs):
  """sensor_agent.py:481-483,527-531: mean of the softmaxed target-speed logits and of the predicted checkpoints over
  the ensemble members.  outs: list of forward 10-tuples."""
  probs = torch.stack([torch.softmax(o[1], dim=1) for o in outs]).mean(0)
  checkpoints = torch.stack([o[2] for o in outs]).mean(0)
  return probs, checkpoints


def non_maximum_suppression(bounding_boxes, iou_treshhold):
  """Drop-in for transfuser_utils.non_maximum_suppression (transfuser_utils.py:409-433; the misspelt argument name is
  the reference's): ``bounding_boxes`` is a list over ensemble members of lists of metric boxes
  (convert_features_to_bb_metric output, score last); returns the kept boxes as a list of numpy arrays, highest score
  first.  The pairwise rotated IoUs and the greedy walk run in one kernel launch (csrc/nms.cu)."""
  flat = [np.asarray(b, dtype=np.float32) for member in bounding_boxes if member is not None for b in member]
  if not flat:
    return []
  if len(flat) > 512:
    raise ValueError('tfpp_nms_rotated merges at most 512 boxes per frame')
  boxes = torch.from_numpy(np.stack(flat)).cuda().unsqueeze(0).contiguous()
  out, count = ops.nms_rotated(boxes, float('-inf'), iou_treshhold)
  n = int(count[0])
  return list(out[0, :n].cpu().numpy())


class EnsembleForward:
  """BASELINE.json config 5 / sensor_agent.py:445-552 on the device: ``members`` LidarCenterNets (the agent loads one
  per checkpoint, sensor_agent.py:120-139) run the same batch of frames in eval mode from ONE CUDA graph; per frame the
  target-speed probabilities and checkpoints are averaged over the members and the members' decoded boxes
  (center_net.py:172-237) are thresholded, converted to the vehicle frame and merged by rotated-IoU NMS.
  Returns (probs (B,4), checkpoints (B,10,2), boxes (B, members*K, 9) highest score first + zero padding, counts (B,))."""

  def __init__(self, members, example_inputs):
    self.members = [m.eval() for m in members]
    self.cfg = members[0].config
    self.static = {k: v.clone().cuda() for k, v in example_inputs.items()}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
      for _ in range(2):
        self._run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph), torch.no_grad():
      self.out = self._run()

  def _run(self):
    cfg = self.cfg
    outs = [m(**self.static) for m in self.members]
    probs, checkpoints = ensemble_outputs(outs)
    decoded = torch.cat([m.head.get_bboxes(*o[6]) for m, o in zip(self.members, outs)], dim=1).contiguous()
    boxes, counts = ops.nms_rotated(decoded, cfg.bb_confidence_threshold, cfg.iou_treshold_nms, to_vehicle=True,
                                    pixels_per_meter=cfg.pixels_per_meter, min_x=cfg.min_x, min_y=cfg.min_y)
    return probs, checkpoints, boxes, counts

  def __call__(self, **inputs):
    for k, v in inputs.items():
      self.static[k].copy_(v, non_blocking=True)
    self.graph.replay()
    return self.out
