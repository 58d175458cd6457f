"""N-GPU check of the NVLink peer exchange (csrc/peer_exchange.cu): launch with
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py
1. tfpp_peer_adamw_step == NCCL all-reduce + tfpp_adamw_amsgrad on synthetic buffers (eager and from a CUDA graph),
2. timing of both at the model's size (120 M parameters),
3. a captured Trainer step on the real model: replicas stay bit-identical, loss finite.
Prints one JSON line on rank 0; exit code 1 on any mismatch."""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', init_method='env://', device_id=torch.device('cuda', local))
  pg = dist.group.WORLD
  from carla_garage_b200 import _lib, ops, peer
  lib = _lib.load()
  res = {'world': world}
  ok = True

  # ---- 1. synthetic buffers -------------------------------------------------------------------------------------
  n = 4 * 1000003 + 4 * 7   # not a multiple of world * 4: ragged last shard
  x = peer.PeerExchange(pg, n)
  g = torch.Generator(device='cuda').manual_seed(1234)
  p0 = torch.randn(n, device='cuda', generator=g)
  g.manual_seed(100 + rank)
  flags = (torch.rand(n, device='cuda', generator=torch.Generator(device='cuda').manual_seed(7)) * 4).to(torch.uint8)

  def make_state():
    return types.SimpleNamespace(exp_avg=torch.zeros(n, device='cuda'), exp_avg_sq=torch.zeros(n, device='cuda'),
                                 max_exp_avg_sq=torch.zeros(n, device='cuda'),
                                 dev_state=torch.tensor([0.0, 3e-4, 0.0, 0.0], device='cuda'), flags=flags)

  st_peer, st_ref = make_state(), make_state()
  x.param.copy_(p0)
  p_ref = p0.clone()
  graph = None
  for step in range(4):
    grad = torch.randn(n, device='cuda', generator=g) * (1.0 + rank)
    x.grad.copy_(grad)
    # reference: NCCL sum + the single-GPU fused AdamW with grad_scale 1/world
    gsum = grad.clone()
    dist.all_reduce(gsum)
    _lib.check(lib.tfpp_adamw_amsgrad(p_ref.data_ptr(), gsum.data_ptr(), st_ref.exp_avg.data_ptr(), st_ref.exp_avg_sq.data_ptr(),
                                      st_ref.max_exp_avg_sq.data_ptr(), n, 0.0, 0.9, 0.999, 1e-8, 0.01, step + 1, 1.0 / world,
                                      st_ref.dev_state.data_ptr(), flags.data_ptr(), ops._stream()), 'tfpp_adamw_amsgrad')
    if step < 2:
      x.step(st_peer)
    else:  # the same launch sequence from a CUDA graph (what Trainer.capture records)
      if graph is None:
        torch.cuda.synchronize()
        dist.barrier()
        graph = torch.cuda.CUDAGraph()
        snap = [t.clone() for t in (x.param, st_peer.exp_avg, st_peer.exp_avg_sq, st_peer.max_exp_avg_sq, st_peer.dev_state)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
          x.step(st_peer)   # warm-up on a side stream (real step: undone below)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        dist.barrier()
        for t, s_ in zip((x.param, st_peer.exp_avg, st_peer.exp_avg_sq, st_peer.max_exp_avg_sq, st_peer.dev_state), snap):
          t.copy_(s_)
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.graph(graph):
          x.step(st_peer)
        # capture does not execute: nothing to undo
      graph.replay()
    torch.cuda.synchronize()
    lo, hi = x.shard
    err = float((x.param - p_ref).abs().max())
    serr = float((st_peer.exp_avg[lo:hi] - st_ref.exp_avg[lo:hi]).abs().max())
    same = [torch.zeros(1, device='cuda', dtype=torch.float64) for _ in range(world)]
    dist.all_gather(same, x.param.double().sum().reshape(1))
    replicas_equal = all(float(s) == float(same[0]) for s in same)
    res[f'step{step}'] = {'max_abs_err_param': err, 'max_abs_err_exp_avg_shard': serr, 'replicas_equal': replicas_equal}
    ok = ok and err < 2e-6 and serr < 1e-5 and replicas_equal
  x.bufs.close()

  # ---- 2. timing at the model's size ------------------------------------------------------------------------------
  n = 120_342_512 // 4 * 4
  x = peer.PeerExchange(pg, n)
  st = types.SimpleNamespace(exp_avg=torch.zeros(n, device='cuda'), exp_avg_sq=torch.zeros(n, device='cuda'),
                             max_exp_avg_sq=torch.zeros(n, device='cuda'),
                             dev_state=torch.tensor([0.0, 3e-4, 0.0, 0.0], device='cuda'), flags=None)
  x.grad.normal_()
  buf = torch.randn(n, device='cuda')

  def timed(fn, iters=10):
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
      fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

  def nccl_path():
    works = [dist.all_reduce(buf[s:s + 16 * 1024 * 1024], async_op=True) for s in range(0, n, 16 * 1024 * 1024)]
    for w in works:
      w.wait()
    _lib.check(lib.tfpp_adamw_amsgrad(x.param.data_ptr(), buf.data_ptr(), st.exp_avg.data_ptr(), st.exp_avg_sq.data_ptr(),
                                      st.max_exp_avg_sq.data_ptr(), n, 0.0, 0.9, 0.999, 1e-8, 0.01, 1, 1.0 / world,
                                      st.dev_state.data_ptr(), None, ops._stream()), 'tfpp_adamw_amsgrad')

  res['ms_peer_step'] = timed(lambda: x.step(st))
  res['ms_nccl_allreduce_plus_adamw'] = timed(nccl_path)
  res['peer_bytes_in_per_rank'] = (world - 1) * (n // world) * 4
  res['peer_read_gbs'] = res['peer_bytes_in_per_rank'] / (res['ms_peer_step'] * 1e-3) / 1e9
  x.bufs.close()

  # ---- 3. the real model: captured step, replicas identical ---------------------------------------------------------
  if os.environ.get('PEER_CHECK_MODEL', '1') == '1':
    from carla_garage_b200 import synth
    from carla_garage_b200.config import GlobalConfig
    from carla_garage_b200.nn import LidarCenterNet
    from carla_garage_b200.training import Trainer
    torch.manual_seed(rank)   # different initial weights per rank: the constructor's broadcast must fix that
    net = LidarCenterNet(GlobalConfig()).cuda().train()
    tr = Trainer(net, process_group=pg)
    assert tr.xchg is not None
    b = 4
    inp = {k: v.cuda() for k, v in synth.make_inputs(b, seed=50 + rank).items()}
    lab = {k: v.cuda().contiguous() for k, v in synth.make_labels(b, seed=60 + rank).items()}
    tr.capture(inp, lab)
    losses = []
    for _ in range(3):
      _, gl = tr.replay()
      losses.append(float(gl.sum()))
    torch.cuda.synchronize()
    chk = [torch.zeros(2, device='cuda', dtype=torch.float64) for _ in range(world)]
    dist.all_gather(chk, torch.stack([tr.st.flat.double().sum(), tr.st.flat.double().abs().sum()]))
    equal = all(torch.equal(c, chk[0]) for c in chk)
    res['model'] = {'graphs': 1 if tr.graph_opt is None else 2, 'replicas_equal': equal, 'losses': losses,
                    'launches_per_step': tr.launches_per_step}
    ok = ok and equal and all(l == l and abs(l) < 1e6 for l in losses) and tr.graph_opt is None
  res['ok'] = ok
  if rank == 0:
    print(json.dumps(res))
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == '__main__':
  main()
