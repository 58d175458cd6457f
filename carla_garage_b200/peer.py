"""NVLink peer-memory buffers of the data-parallel step (csrc/peer_exchange.cu, include/tfpp.h "tfpp_peer_*").

One process per GPU (torchrun).  Every rank cudaMalloc's its flat parameter buffer, its flat gradient buffer and a few
barrier words through the C ABI, publishes the CUDA IPC handles through torch.distributed (any backend: the handles are
64 opaque bytes) and maps the other ranks' buffers.  torch sees the local buffers as ordinary tensors
(``__cuda_array_interface__``), so FlatState / the kernels use them like any other allocation; the fused
reduce-scatter + AdamW + all-gather kernel gets the table of peer pointers.

Replaces DistributedDataParallel + ZeroRedundancyOptimizer of the reference (train.py:516,527-531)."""
import ctypes

import torch

from . import _lib
from ._lib import PeerStepArgs, check


def shard_bounds(n, world, rank):
  """[lo, hi) of the flat buffer (n elements, n % 4 == 0) owned by ``rank``: float4-granular equal chunks — the same
  arithmetic as tfpp_peer_adamw_step."""
  n4 = n // 4
  chunk = -(-n4 // world)
  lo = min(chunk * rank, n4)
  return 4 * lo, 4 * min(lo + chunk, n4)


class _Raw:
  """A cudaMalloc'ed region exposed to torch through the CUDA array interface."""

  def __init__(self, ptr, nbytes, typestr, shape):
    self.ptr, self.nbytes = ptr, nbytes
    self.__cuda_array_interface__ = {'shape': shape, 'typestr': typestr, 'data': (ptr, False), 'version': 2,
                                     'strides': None}


class PeerBuffers:
  """Symmetric device allocations (same sizes on every rank) mapped into every process of ``group``."""

  def __init__(self, group, sizes):
    """sizes: dict name -> (numel, torch dtype) of this rank's buffers.  All ranks must pass the same dict."""
    self.group = group
    self.world = torch.distributed.get_world_size(group)
    self.rank = torch.distributed.get_rank(group)
    if self.world > 8:
      raise RuntimeError('the peer exchange addresses at most 8 ranks (one NVSwitch domain)')
    self.dev = torch.device('cuda', torch.cuda.current_device())
    lib = _lib.load()
    self.local, self.tensors, handles = {}, {}, {}
    for name, (numel, dtype) in sizes.items():
      item = torch.empty(0, dtype=dtype).element_size()
      ptr = ctypes.c_void_p()
      h = (ctypes.c_ubyte * 64)()
      check(lib.tfpp_peer_alloc(numel * item, ctypes.byref(ptr), h), 'tfpp_peer_alloc')
      self.local[name] = ptr.value
      handles[name] = bytes(h)
      typestr = {torch.float32: '<f4', torch.int32: '<i4', torch.uint8: '|u1'}[dtype]
      raw = _Raw(ptr.value, numel * item, typestr, (numel,))
      self.tensors[name] = torch.as_tensor(raw, device=self.dev)
      assert self.tensors[name].data_ptr() == ptr.value
    everyone = [None] * self.world
    torch.distributed.all_gather_object(everyone, (self.rank, torch.cuda.current_device(), handles), group=group)
    self.ptrs = {name: [0] * self.world for name in sizes}
    self._opened = []
    for r, _, hs in everyone:
      for name in sizes:
        if r == self.rank:
          self.ptrs[name][r] = self.local[name]
        else:
          p = ctypes.c_void_p()
          check(lib.tfpp_peer_open(hs[name], ctypes.byref(p)), 'tfpp_peer_open')
          self.ptrs[name][r] = p.value
          self._opened.append(p.value)
    torch.distributed.barrier(group=group)   # nobody runs ahead of a peer that has not mapped the buffers yet

  def close(self):
    lib = _lib.load()
    torch.cuda.synchronize()
    torch.distributed.barrier(group=self.group)
    for p in self._opened:
      lib.tfpp_peer_close(p)
    self._opened = []
    torch.distributed.barrier(group=self.group)
    for p in self.local.values():
      lib.tfpp_peer_free(p)
    self.local = {}


class PeerExchange:
  """The exchange step of one Trainer: tables of peer pointers + the fused kernel launch."""

  FLAG_WORDS = 64

  def __init__(self, group, numel):
    assert numel % 4 == 0
    self.numel = numel
    self.bufs = PeerBuffers(group, {'param': (numel, torch.float32), 'grad': (numel, torch.float32),
                                    'flags': (self.FLAG_WORDS, torch.int32)})
    self.world, self.rank = self.bufs.world, self.bufs.rank
    self.param, self.grad = self.bufs.tensors['param'], self.bufs.tensors['grad']
    self.shard = shard_bounds(numel, self.world, self.rank)

  def args(self, st, betas, eps, weight_decay):
    a = PeerStepArgs()
    a.world, a.rank = self.world, self.rank
    for r in range(self.world):
      a.grad[r], a.param[r], a.flags[r] = self.bufs.ptrs['grad'][r], self.bufs.ptrs['param'][r], self.bufs.ptrs['flags'][r]
    a.exp_avg, a.exp_avg_sq, a.max_exp_avg_sq = st.exp_avg.data_ptr(), st.exp_avg_sq.data_ptr(), st.max_exp_avg_sq.data_ptr()
    a.n = self.numel
    a.beta1, a.beta2, a.eps, a.weight_decay = betas[0], betas[1], eps, weight_decay
    a.dev_state = st.dev_state.data_ptr()
    a.opt_flags = st.flags.data_ptr() if st.flags is not None else None
    return a

  def step(self, st, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """barrier -> reduce-scatter + AdamW(shard, mean gradient) + parameter all-gather -> barrier, on the current stream."""
    a = self.args(st, betas, eps, weight_decay)
    check(_lib.load().tfpp_peer_adamw_step(ctypes.byref(a), torch.cuda.current_stream().cuda_stream),
          'tfpp_peer_adamw_step')

  def barrier(self, slot=0):
    a = PeerStepArgs()
    a.world, a.rank = self.world, self.rank
    for r in range(self.world):
      a.flags[r] = self.bufs.ptrs['flags'][r]
    check(_lib.load().tfpp_peer_barrier(ctypes.byref(a), slot, torch.cuda.current_stream().cuda_stream),
          'tfpp_peer_barrier')
