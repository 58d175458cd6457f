"""TEST INFRASTRUCTURE (oracle): numpy restatement of the dropout random stream of carla_garage_b200/csrc/common.cuh —
Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter =
(i / 4 lo, i / 4 hi, site, step), key = (seed lo, seed hi); element i takes word i % 4 and is dropped iff
word < p * 2^32.  The reference's dropout (nn.Dropout, transfuser.py:325,374,379,395; nn.TransformerDecoderLayer,
model.py:137-140) draws from torch's generator, which no independent implementation can match bit for bit; parity of
the dropout path is therefore checked by feeding THESE masks to the fp32 oracle at the reference's dropout sites.
Pinned by the known-answer vectors of the Random123 distribution (tests/test_oracle.py)."""
import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
  """Vectorised over numpy uint32 arrays (counters) with scalar keys; returns the four output words."""
  c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
  k0, k1 = np.uint32(k0), np.uint32(k1)
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = M0 * c0.astype(np.uint64)
      p1 = M1 * c2.astype(np.uint64)
      hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
      hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
      c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
      k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
      k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
  return c0, c1, c2, c3


def words(n, seed, step, site):
  """The first n 32-bit words of dropout site ``site`` at step ``step``."""
  n4 = (n + 3) // 4
  idx = np.arange(n4, dtype=np.uint64)
  lo, hi = (idx & MASK32).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32)
  ones = np.ones(n4, dtype=np.uint32)
  w = philox4x32_10(lo, hi, ones * np.uint32(site), ones * np.uint32(step & 0xFFFFFFFF), seed & 0xFFFFFFFF,
                    (seed >> 32) & 0xFFFFFFFF)
  return np.stack(w, axis=1).reshape(-1)[:n]


def threshold(p):
  t = float(np.float32(p)) * 4294967296.0
  return 0xFFFFFFFF if t >= 4294967295.0 else int(t)


def multiplier(shape, p, seed, step, site):
  """float32 tensor of ``shape``: 0 where dropped, 1/(1-p) where kept (C-order element index)."""
  n = int(np.prod(shape))
  keep = words(n, seed, step, site) >= np.uint32(threshold(p))
  inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
  return torch.from_numpy((keep.astype(np.float32) * inv).reshape(shape))


class DropoutStream:
  """Hands out multipliers in the order the engine numbers its dropout sites (one per call)."""

  def __init__(self, seed, step, first_site=0):
    self.seed, self.step, self.site = int(seed), int(step), int(first_site)

  def __call__(self, t, p):
    if p <= 0.0:
      return t
    m = multiplier(tuple(t.shape), p, self.seed, self.step, self.site)
    self.site += 1
    return t * m.to(t.dtype)
