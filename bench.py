#!/usr/bin/env python
"""bench.py — TransFuser++ training-step throughput on B200 (BASELINE.json metric: train samples/s).

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host CPU cores

A "step" = forward + fused losses + backward + gradient all-reduce + AdamW(amsgrad) on one synthetic batch of
32 samples per GPU (config "TransFuser++ train step bf16, RegNetY-3.2GF backbones, batch=32 on 1xB200"), weak scaling.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 32
FLOP_PER_SAMPLE_TRAIN = 306.4e9  # 3 x 51.06 GMAC x 2


def _clock_sampler(stop, samples):
  q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')
  while not stop.is_set():
    try:
      out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', '0'],
                           capture_output=True, text=True, timeout=5).stdout.strip()
      if out:
        samples.append([x.strip() for x in out.split(',')])
    except Exception:  # pylint: disable=broad-except
      pass
    stop.wait(0.2)


def _clock_summary(samples):
  sm, reasons, mx = [], set(), None
  for s in samples:
    try:
      sm.append(float(s[1]))
      mx = float(s[2])
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[4:8]):
        if v.lower().startswith('active'):
          reasons.add(name)
    except Exception:  # pylint: disable=broad-except
      continue
  sm.sort()
  return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
          'samples': len(sm)}


def pick_cpu_threads(batch=2):
  """torch's CPU kernels slow down badly when oversubscribed on many-core hosts (117 s per step with 128 threads vs
  ~4 s with 8 on the same network), so the reference arm uses the fastest of a few thread counts, chosen by timing one
  eval-mode forward each."""
  import torch
  from carla_garage_b200 import synth
  from oracle import tfpp_oracle as orc
  ncpu = os.cpu_count() or 1
  cands = sorted({c for c in (8, 16, 32, 64) if c <= ncpu} | ({ncpu} if ncpu <= 64 else set()))
  if len(cands) <= 1:
    return cands[0] if cands else ncpu
  sd = synth.golden_state(os.path.join(ROOT, 'tests', 'golden'))
  inp = synth.make_inputs(batch, seed=1234)
  best, best_t = cands[0], float('inf')
  for c in cands:
    torch.set_num_threads(c)
    with torch.no_grad():
      orc.forward(sd, **inp)  # warm the primitive caches for this thread count
      t0 = time.perf_counter()
      orc.forward(sd, **inp)
      dt = time.perf_counter() - t0
    if dt < best_t:
      best, best_t = c, dt
  return best


def cpu_reference_step_rate(batch, steps, warmup, threads, budget_s=150.0):
  """The reference algorithm (oracle/tfpp_oracle.py: CPU fp32 restatement pinned to the reference) as a full train
  step: forward (training BN), 10 losses, autograd backward, torch AdamW(amsgrad).  A step costs about a minute per
  sample on the host, so the run is bounded by time: at most ``steps`` timed steps, stopping once ``budget_s`` seconds
  of timed work are spent (at least one).  Returns (samples/s, seconds per step, timed steps run)."""
  import torch
  from carla_garage_b200 import synth
  from oracle import tfpp_oracle as orc
  torch.set_num_threads(threads)
  sd = synth.golden_state(os.path.join(ROOT, 'tests', 'golden'))
  params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
            if v.is_floating_point() and 'running' not in k and 'valid_bev' not in k and not k.startswith('loss_')}
  state = dict(sd)
  state.update(params)
  opt = torch.optim.AdamW(list(params.values()), lr=3e-4, amsgrad=True)
  inp = synth.make_inputs(batch, seed=1234)
  lab = synth.make_labels(batch, seed=1234)
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    out = orc.forward(state, **inp, training=True)
    loss = orc.total_loss(orc.compute_loss(sd, out, lab))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
      if sum(times) >= budget_s:
        break
  return batch * len(times) / sum(times), sum(times) / len(times), len(times)


WORKLOAD = 'TransFuser++ train step bf16, RegNetY-3.2GF backbones, batch=32 per GPU'


def run_reference(args):
  """Reference arm: the reference's own algorithm (the oracle port: the reference is pure Python/torch and cannot be
  pip-installed offline) on the host cores, same metric / unit / config as the B200 arm."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  threads = pick_cpu_threads()
  batch = 2  # smallest batch the training-mode BatchNorm1d of the velocity input accepts
  wu = min(max(args.warmup, 0), 1)
  rate, sec, ran = cpu_reference_step_rate(batch, max(args.steps, 1), wu, threads)
  sample = (f'{ran} train step(s) of batch {batch} ({sec:.1f} s each; time-bounded, {args.steps} requested) of '
            f'oracle/tfpp_oracle.py (torch CPU fp32), {threads} threads (fastest of 8/16/32/64)')
  line = {
      'impl': 'reference', 'metric': 'train_samples_per_s', 'value': rate, 'unit': 'samples/s', 'n_gpus': args.gpus,
      'steps': ran, 'warmup': wu, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': WORKLOAD, 'global_batch': PER_GPU_BATCH * args.gpus, 'per_gpu_batch': PER_GPU_BATCH,
                 'parallelism': f'dp{args.gpus}', 'bounded_sample': sample},
      'cpu_baseline': {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port', 'sample': sample},
      'e2e': {'value': rate, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200')
  ap.add_argument('--batch', type=int, default=PER_GPU_BATCH)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
    return

  import torch
  import torch.distributed as dist
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if not torch.cuda.is_available():
    raise RuntimeError('bench.py needs a CUDA device (the product has no CPU fallback); use --impl reference for CPU')
  torch.cuda.set_device(local_rank)
  pg = None
  if world > 1:
    dist.init_process_group('nccl', init_method='env://', device_id=torch.device('cuda', local_rank))
    pg = dist.group.WORLD
  from carla_garage_b200 import _lib, ops, synth
  from carla_garage_b200.config import GlobalConfig
  from carla_garage_b200.nn import LidarCenterNet
  from carla_garage_b200.training import Trainer
  warmup = max(args.warmup, 3)
  b = args.batch
  torch.manual_seed(0)
  cfg = GlobalConfig()
  backbone = os.environ.get('TFPP_BENCH_BACKBONE', cfg.backbone)  # 'bev_encoder': SURVEY.md §8 f3 (not the headline config)
  cfg.backbone = backbone
  net = LidarCenterNet(cfg)
  golden = os.path.join(ROOT, 'tests', 'golden')
  net.load_state_dict(synth.bev_state(golden) if backbone == 'bev_encoder' else synth.golden_state(golden), strict=True)
  net = net.cuda().train()
  tr = Trainer(net, process_group=pg)
  host_in = {k: v.pin_memory() for k, v in synth.make_inputs(b, seed=1234 + rank).items()}
  host_pts = synth.make_point_clouds(b, seed=1234 + rank).pin_memory()
  host_lab = {k: v.contiguous().pin_memory() for k, v in synth.make_labels(b, seed=1234 + rank).items()}
  dev_in = {k: v.cuda() for k, v in host_in.items()}
  dev_in['lidar_bev'] = ops.pillar_scatter(host_pts.cuda())  # the real voxelised LiDAR (K1)
  dev_lab = {k: v.cuda() for k, v in host_lab.items()}

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device='cuda')
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t)
    return ms

  # ---- device-resident throughput.  N=1: the whole step (K1 + fwd + losses + bwd + AdamW) replays from one CUDA
  # graph; N>1: two graphs (forward/backward, optimizer) with the NCCL all-reduce issued eagerly between them.
  dev_pts = host_pts.cuda()
  use_graph = os.environ.get('TFPP_NO_GRAPH', '0') != '1'
  if use_graph:
    tr.capture(dev_in, dev_lab, points=dev_pts, split=True if os.environ.get('TFPP_SPLIT_GRAPH', '0') == '1' else None)

  def step_resident():
    if use_graph:
      tr.replay()
    else:
      tr.step(dev_in, dev_lab)

  for _ in range(warmup):
    step_resident()
  stop, samples = threading.Event(), []
  th = threading.Thread(target=_clock_sampler, args=(stop, samples), daemon=True)
  if rank == 0:
    th.start()
  launches_per_step = tr.launches_per_step if use_graph else None
  _lib.reset_launch_count()
  prof_range = os.environ.get('TFPP_PROFILE_STEP', '0') == '1'  # ncu --profile-from-start off: timed steps only
  if prof_range:
    torch.cuda.profiler.start()
  ms = timed(step_resident, args.steps)
  if prof_range:
    torch.cuda.profiler.stop()
  launches = _lib.launch_count() if not use_graph else launches_per_step * args.steps
  stop.set()
  value = world * b * args.steps / (ms / 1e3)

  # ---- end to end: host buffers in pinned memory -> device every step (+ K1 on the device), loss read back
  h2d = sum(v.numel() * v.element_size() for k, v in host_in.items() if k != 'lidar_bev') + \
      host_pts.numel() * host_pts.element_size() + sum(v.numel() * v.element_size() for v in host_lab.values())
  loss_host = torch.zeros(10, pin_memory=True)

  prefetch = use_graph and os.environ.get('TFPP_PREFETCH', '1') == '1'  # H2D of step i+1 on a copy stream under step i (A/B on B200: e2e 521.6 -> 565.1 samples/s)
  if prefetch:
    tr.stage({k: v for k, v in host_in.items() if k != 'lidar_bev'}, host_lab, host_pts)

  def step_e2e():
    if prefetch:
      _, gl = tr.replay_staged()
      tr.stage({k: v for k, v in host_in.items() if k != 'lidar_bev'}, host_lab, host_pts)  # next step's inputs
      loss_host.copy_(gl, non_blocking=True)
      return
    if use_graph:
      _, gl = tr.replay({k: v for k, v in host_in.items() if k != 'lidar_bev'}, host_lab, host_pts)
      loss_host.copy_(gl, non_blocking=True)
      return
    inp = {k: v.cuda(non_blocking=True) for k, v in host_in.items() if k != 'lidar_bev'}
    inp['lidar_bev'] = ops.pillar_scatter(host_pts.cuda(non_blocking=True))
    lab = {k: v.cuda(non_blocking=True) for k, v in host_lab.items()}
    _, losses = tr.step(inp, lab)
    loss_host.copy_(torch.stack([losses[k] for k in sorted(losses)]), non_blocking=True)

  for _ in range(2):
    step_e2e()
  ms_e2e = timed(step_e2e, args.steps)
  e2e = world * b * args.steps / (ms_e2e / 1e3)

  # ---- roofline of the dominant kernel family (tcgen05 implicit GEMM: conv_gemm + conv_wgrad): one instrumented
  # step with CUDA events around every launch on the launching stream
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
  except Exception:  # pylint: disable=broad-except
    pass
  roof = None
  if world > 1:  # every collective of the run is behind us: the rest is rank-local work on rank 0
    barrier()
    dist.destroy_process_group()
  if rank == 0:
    # rank-local instrumented step: no collective inside (the other ranks are not running it)
    tr.local_only = True
    # one stream for this step: with the LiDAR branch / planner on side streams the per-launch times would include the
    # kernels they share the SMs with
    prev_no_overlap = os.environ.get('TFPP_NO_OVERLAP')
    os.environ['TFPP_NO_OVERLAP'] = '1'
    try:
      prof = ops.profile_gemm_launches(lambda: tr.step(dev_in, dev_lab))
      calls = _lib.profile_calls(lambda: tr.step(dev_in, dev_lab))   # every C-ABI entry point of one eager step
    finally:
      if prev_no_overlap is None:
        os.environ.pop('TFPP_NO_OVERLAP', None)
      else:
        os.environ['TFPP_NO_OVERLAP'] = prev_no_overlap
    tr.local_only = False
    peak = peaks.get('bf16_tflops_sustained', 1400.0)
    roof = {'bound': 'tensor', 'kernel': 'conv_gemm_kernel + wgrad_kernel (tcgen05)',
            'achieved': prof['tflops'], 'peak': peak, 'unit': 'TFLOP/s', 'frac': prof['tflops'] / peak,
            'peak_source': 'measured (MEASURED_PEAKS.json, sustained)' if peaks else 'fallback',
            # DRAM bytes of the largest-FLOP launch of the family (fusion QKV GEMM, M=10240 K=1512 N=4536) from
            # `ncu --set full` (profiles/r01_ncu_qkv_v13_details.txt): 44.8 MB read + 39.8 MB written, against 138 MB
            # of algorithmic operand + output bytes (operand re-reads hit L2; part of the output is still in L2)
            'traffic': 121.3e6, 'traffic_launch': 'conv_gemm_kernel<pair>, fusion MLP up-projection (M=10240, K=1512, N=6048): dram read 49.4 MB + write 71.8 MB, profiles/r02_ncu_gemm_mlp_pair.txt (algorithmic 173 MB; the tail of the 124 MB output is still in L2 when the kernel ends)',
            'launches_per_step': prof['launches'],
            # serial GEMM time over the (stream-overlapped) step time: an upper bound of the family's share
            'share_of_step': prof['ms'] / (ms / args.steps),
            'algorithmic_gflop_per_step': prof['gflop']}
    # device time of one eager single-stream step by C-ABI entry point (CUDA events around every call), largest first;
    # HBM-bound entries with a simple byte model also get their achieved bandwidth against the measured copy peak
    hbm_peak = peaks.get('hbm_gbs', 6500.0)
    n_param = tr.st.flat.numel()
    byte_model = {'tfpp_adamw_amsgrad': 36.0 * n_param,                                   # 5 reads + 4 writes of fp32
                  'tfpp_pillar_scatter': float(host_pts.numel() * 4 + b * 2 * 256 * 256 * 4 * 2 + b * 256 * 256 * 4)}
    tot_ms = sum(v[0] for v in calls.values())
    roof['step_kernels'] = [dict(entry=k, ms=round(v[0], 3), calls=v[1], share=round(v[0] / tot_ms, 4),
                                 **({'gb_s': round(byte_model[k] / (v[0] * 1e-3) / 1e9, 1),
                                     'frac_of_hbm_peak': round(byte_model[k] / (v[0] * 1e-3) / 1e9 / hbm_peak, 3)}
                                    if k in byte_model else {}))
                            for k, v in sorted(calls.items(), key=lambda kv: -kv[1][0])[:16]]
    roof['step_kernels_total_ms'] = round(tot_ms, 2)

  if rank != 0:
    return
  # ---- second half of BASELINE.json's metric: forward ms/frame at batch 1 (the agent's 20 Hz loop), graph replay
  from carla_garage_b200.inference import GraphedForward
  net.eval()
  one = {k: v[:1].contiguous() for k, v in dev_in.items()}
  gf = GraphedForward(net, one)
  for _ in range(5):
    gf(**one)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize()
  e0.record()
  for _ in range(50):
    gf(**one)
  e1.record()
  torch.cuda.synchronize()
  fwd_ms = e0.elapsed_time(e1) / 50
  inference = {'fwd_ms_per_frame': fwd_ms, 'batch': 1, 'mode': 'eval, CUDA-graph replay, inputs resident'}
  # ---- BASELINE.json config 5: sensor_agent.py inference, 3-member ensemble, batch 64, forward only (one CUDA graph:
  # 3 forwards + CenterNet decode + threshold / vehicle-frame conversion / rotated-IoU NMS over the union + averaging)
  if os.environ.get('TFPP_BENCH_ENSEMBLE', '1') == '1' and backbone == 'transFuser':
    from carla_garage_b200.inference import EnsembleForward
    del gf
    members = [net]
    base = synth.golden_state(os.path.join(ROOT, 'tests', 'golden'))
    for s_ in (1, 2):
      m_ = LidarCenterNet(GlobalConfig())
      g_ = torch.Generator().manual_seed(s_)
      m_.load_state_dict({k: (v + 0.01 * torch.randn(v.shape, generator=g_) if v.is_floating_point() and v.dim() >= 2 else v)
                          for k, v in base.items()}, strict=True)
      members.append(m_.cuda().eval())
    eb = 64
    big = {k: torch.cat([v] * (-(-eb // v.shape[0])))[:eb].contiguous() for k, v in dev_in.items()}
    ef = EnsembleForward(members, big)
    for _ in range(2):
      ef()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
      ef()
    e1.record()
    torch.cuda.synchronize()
    ens_ms = e0.elapsed_time(e1) / 5
    inference['ensemble'] = {'frames_per_s': eb / (ens_ms / 1e3), 'ms_per_batch': ens_ms, 'batch': eb, 'members': len(members),
                             'boxes_kept_mean': float(ef.out[3].float().mean()),
                             'includes': '3 eval forwards + decode_heatmap + confidence threshold + vehicle-frame '
                                         'conversion + rotated-IoU NMS over the union + ensemble means, one CUDA graph'}
  cpu = None
  if not args.no_cpu_baseline:
    threads = pick_cpu_threads()
    rate, sec, ran = cpu_reference_step_rate(2, 3, 1, threads, budget_s=30.0)
    cpu = {'value': rate, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
           'sample': f'{ran} train step(s) of batch 2 ({sec:.1f} s each) of oracle/tfpp_oracle.py (torch CPU fp32), '
                     f'{threads} threads (fastest of 8/16/32/64)'}
    from oracle import tfpp_oracle as orc
    sd = synth.golden_state(os.path.join(ROOT, 'tests', 'golden'))
    one_cpu = {k: v[:1] for k, v in synth.make_inputs(1, seed=1234).items()}
    with torch.no_grad():
      orc.forward(sd, **one_cpu)
      t0 = time.perf_counter()
      for _ in range(3):
        orc.forward(sd, **one_cpu)
    inference['cpu_ref_ms_per_frame'] = (time.perf_counter() - t0) / 3 * 1e3
    inference['cpu_cores'] = threads
  line = {
      'metric': 'train_samples_per_s', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
      'config': {'workload': (WORKLOAD if backbone == 'transFuser' else WORKLOAD.replace('TransFuser++', f'LidarCenterNet(backbone={backbone})')) if b == PER_GPU_BATCH else WORKLOAD.replace('batch=32', f'batch={b}') + (' (BASELINE.json config 3: DDP imitation training, batch=12/GPU)' if b == 12 else ''),
                 'global_batch': world * b, 'per_gpu_batch': b, 'parallelism': f'dp{world}',
                 'l2': 'per-step working set (~20 GB of activations) >> 126 MB L2; no flush needed',
                 'inputs': 'rgb (B,3,256,1024) f32 + 60k-point LiDAR cloud -> (B,1,256,256) BEV (reference default use_ground_plane=0)',
                 'dropout': ('on: embd/attn/resid_pdrop 0.1 + decoder 0.1, Philox4x32-10 masks regenerated in the backward kernels' if net.engine.dropout_enabled else 'off (TFPP_DROPOUT=0)'), 'cuda_graph': bool(use_graph), 'graphs': (1 if tr.graph_opt is None else 2) if use_graph else 0,
                 'exchange': ('none (1 GPU)' if world == 1 else ('NVLink peer-memory reduce-scatter + AdamW shard + all-gather in one kernel (csrc/peer_exchange.cu)' if tr.xchg is not None else 'NCCL all-reduce between two graphs')), 'model_tflop_per_step': world * b * FLOP_PER_SAMPLE_TRAIN / 1e12},
      'e2e': {'value': e2e, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 40,
              'ms_per_step': ms_e2e / args.steps},
      'gpu_launches': launches,
      'clocks': _clock_summary(samples),
      'roofline': roof,
      'cpu_baseline': cpu,
      'inference': inference,
      'model_flops_utilisation': (b * FLOP_PER_SAMPLE_TRAIN / (ms / args.steps / 1e3)) / 1e12 /
                                 peaks.get('bf16_tflops_sustained', 1400.0),
  }
  print(json.dumps(line))


if __name__ == '__main__':
  main()
