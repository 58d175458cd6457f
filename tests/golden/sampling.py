"""Shared sub-sampling rule for golden fixtures (generator and tests must agree)."""


def sample(t, max_hw=16, max_c=48):
  """(B,C,H,W) or (B,H,W) tensor -> strided subsample with <= max_c channels and <= max_hw rows/cols."""
  if t.dim() == 4:
    cs = max(1, t.shape[1] // max_c)
    hs = max(1, t.shape[2] // max_hw)
    ws = max(1, t.shape[3] // max_hw)
    return t[:, ::cs, ::hs, ::ws].contiguous()
  if t.dim() == 3 and t.shape[-1] >= 64 and t.shape[-2] >= 64:
    hs = max(1, t.shape[1] // max_hw)
    ws = max(1, t.shape[2] // max_hw)
    return t[:, ::hs, ::ws].contiguous()
  return t.contiguous()
