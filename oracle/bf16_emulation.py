"""ORACLE-side test infrastructure: what an IDEAL bf16-storage implementation of the forward can achieve.

Runs oracle/tfpp_oracle.py with every conv2d / linear rounding its weight, its input and its output to bf16
(fp32 accumulation) - the storage precision of the B200 path (NHWC bf16 activations, bf16 tensor-core operands).
The relative error of this emulation against the fp32 oracle is the noise floor of bf16 storage for a given
state_dict; tests/golden/make_golden.py records it per tap so that the GPU parity test can assert
"error <= max(1e-2, 2 x bf16 floor)" instead of a tolerance the arithmetic cannot meet on a randomly initialised
(ill-conditioned) network.  Not used by the product.
"""
import contextlib

import torch
import torch.nn.functional as real_F

from . import tfpp_oracle as orc


def _r(t):
  return t.to(torch.bfloat16).float()


class _RoundingF:

  def __getattr__(self, name):
    return getattr(real_F, name)

  @staticmethod
  def conv2d(x, w, b=None, **kw):
    y = real_F.conv2d(_r(x) if x.shape[1] > 3 else x, _r(w) if x.shape[1] > 3 else w, b, **kw)
    return _r(y)

  @staticmethod
  def linear(x, w, b=None):
    if w.shape[0] * w.shape[1] < 64 * 64:  # tiny heads run in fp32 on the device too
      return real_F.linear(x, w, b)
    return real_F.linear(_r(x), _r(w), b)


@contextlib.contextmanager
def bf16_storage():
  saved = orc.F
  orc.F = _RoundingF()
  try:
    yield
  finally:
    orc.F = saved


def forward(sd, inputs, training=False, taps=None):
  with bf16_storage(), torch.no_grad():
    return orc.forward(sd, **inputs, training=training, taps=taps)
