"""Inference-side helpers: CUDA-graph replay of the eval forward (the 20 Hz agent loop, sensor_agent.py:343-615, is
launch-latency bound at B=1) and the ensemble reduction of SensorAgent.run_step (sensor_agent.py:445-485,527-552)."""
import torch


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
