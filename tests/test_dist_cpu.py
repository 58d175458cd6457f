"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: bucketed flat-gradient all-reduce and the
flat parameter state (adjacency groups, gradient views)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from carla_garage_b200.training import allreduce_flat
  g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
  allreduce_flat(g, dist.group.WORLD, bucket_elems=96)  # 11 buckets, last one ragged
  ok = torch.equal(g, torch.arange(1000, dtype=torch.float32) * 3)
  ret[rank] = bool(ok)
  dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
  mgr = mp.Manager()
  ret = mgr.dict()
  port = 29500 + os.getpid() % 1000
  mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
  assert ret[0] and ret[1]


def test_flat_state_layout_cpu():
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn import LidarCenterNet
  from carla_garage_b200.training import FlatState
  net = LidarCenterNet(GlobalConfig())
  before = {k: v.clone() for k, v in net.state_dict().items()}
  st = FlatState(net)
  # values survive flattening, state_dict keys are untouched
  after = net.state_dict()
  assert set(after) == set(before)
  for k in ('backbone.image_encoder.s3.b7.conv2.conv.weight', 'join.layers.3.linear1.weight', 'checkpoint_query'):
    assert torch.equal(after[k], before[k])
  at = net.backbone.transformers[2].blocks[1].attn
  span = st.g_span(at.query.weight, at.value.weight)
  assert span.numel() == 3 * at.query.weight.numel()
  assert span.data_ptr() == st.g(at.query.weight).data_ptr()
  assert st.g(at.key.weight).data_ptr() == span.data_ptr() + 4 * at.query.weight.numel()
  heads = net.head.head_names()
  assert st.g_span(getattr(net.head, heads[0])[2].bias, getattr(net.head, heads[-1])[2].bias).numel() == 21
  n_train = sum(p.numel() for p in net.parameters() if p.requires_grad)
  assert n_train == 120351026 - 2 * 256 * 256
  assert n_train <= st.flat.numel() < n_train + 4 * len(st.params)  # 16-byte alignment padding before matrices only
  for q in st.params:
    if q.ndim >= 2 and q.numel() % 4 == 0:
      assert (q.data_ptr() - st.flat.data_ptr()) % 16 == 0 and (q.grad.data_ptr() - st.grad.data_ptr()) % 16 == 0
  # parameters are views of the flat buffer, gradients of the flat gradient
  p = net.change_channel.weight
  assert p.data_ptr() >= st.flat.data_ptr() and p.grad.data_ptr() == st.g(p).data_ptr()


def test_peer_shard_bounds_cover_the_buffer_once():
  """carla_garage_b200.peer.shard_bounds == the partition tfpp_peer_adamw_step uses (float4-granular equal chunks)."""
  from carla_garage_b200.peer import shard_bounds
  for n in (4, 8, 4 * 1000003 + 28, 120342512):
    for world in (1, 2, 3, 4, 8):
      spans = [shard_bounds(n, world, r) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == n
      for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
        assert hi == lo2 and lo % 4 == 0 and hi % 4 == 0 and lo <= hi
