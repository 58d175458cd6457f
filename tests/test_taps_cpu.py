"""Host-side tap tables of the implicit-GEMM convolutions, checked on the CPU against torch: a tiny emulator of
tfpp_conv_gemm's addressing contract (include/tfpp.h: out[b, y, x, n] = sum_taps sum_c a[b + db, y + dy, x + dx, c] *
w[n, tap_w, c], zero outside the map) evaluates the tables ops.py hands to the kernels — 3x3, the 3x3 input gradient, the
stride-2 parity-plane forms used by the BEV stem of the bev_encoder backbone (forward, input gradient, weight gradient)."""
import torch
import torch.nn.functional as F

from carla_garage_b200 import ops


def emulate_conv_gemm(a, w, taps, batch=None):
  """a (Ba, H, W, C), w (N, T, K) -> (B, H, W, N) following the kernel's addressing contract."""
  ab, h, wd, _ = a.shape
  b = ab if batch is None else batch
  out = torch.zeros((b, h, wd, w.shape[0]), dtype=torch.float64)
  for dx, dy, db, tw in taps:
    for bb in range(b):
      src = torch.zeros((h, wd, a.shape[3]), dtype=torch.float64)
      if 0 <= bb + db < ab:
        y0, y1 = max(0, -dy), min(h, h - dy)
        x0, x1 = max(0, -dx), min(wd, wd - dx)
        if y1 > y0 and x1 > x0:
          src[y0:y1, x0:x1] = a[bb + db, y0 + dy:y1 + dy, x0 + dx:x1 + dx]
      out[bb] += src @ w[:, tw, :].t()
  return out


def emulate_conv_wgrad(dy, x, taps, w_taps):
  """dw[co, tap_w, ci] = sum_pixels dy[pixel, co] * x[pixel + tap, ci] (contract of tfpp_conv_wgrad)."""
  b, h, wd, co = dy.shape
  dw = torch.zeros((co, w_taps, x.shape[3]), dtype=torch.float64)
  for dx, dy_, db, tw in taps:
    for bb in range(b):
      if not 0 <= bb + db < x.shape[0]:
        continue
      y0, y1 = max(0, -dy_), min(h, h - dy_)
      x0, x1 = max(0, -dx), min(wd, wd - dx)
      if y1 > y0 and x1 > x0:
        dw[:, tw, :] += torch.einsum('yxo,yxi->oi', dy[bb, y0:y1, x0:x1], x[bb + db, y0 + dy_:y1 + dy_, x0 + dx:x1 + dx])
  return dw


def parity_split(x):
  """tfpp_parity_split: (B, H, W, C) -> (4B, H/2, W/2, C), plane (py, px) of sample b at index (py * 2 + px) * B + b."""
  return torch.cat([x[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=0)


def test_3x3_forward_and_dgrad_taps():
  g = torch.Generator().manual_seed(0)
  x = torch.randn(5, 6, 7, 4, generator=g, dtype=torch.float64)   # (B, H, W, C)
  w = torch.randn(3, 4, 3, 3, generator=g, dtype=torch.float64)
  want = F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
  got = emulate_conv_gemm(x, ops.pack_conv_weight(w, dt=torch.float64), ops.TAPS_3X3)
  assert torch.allclose(got, want, atol=1e-12)
  dy = torch.randn(want.shape, generator=g, dtype=torch.float64)
  xr = x.clone().requires_grad_(True)
  (F.conv2d(xr.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1) * dy).sum().backward()
  got_dx = emulate_conv_gemm(dy, ops.pack_conv_weight_t(w, dt=torch.float64), ops.TAPS_3X3_DGRAD)
  assert torch.allclose(got_dx, xr.grad, atol=1e-12)


def test_stride2_parity_plane_forward_dgrad_wgrad():
  """The BEV stem of the bev_encoder backbone (Engine.bev_stem / Backward.bev_stem): one implicit GEMM over the four
  parity planes forward, one per plane for the input gradient, one over the planes for the weight gradient."""
  g = torch.Generator().manual_seed(1)
  b, h, wd, cin, cout = 2, 8, 10, 5, 3
  x = torch.randn(b, h, wd, cin, generator=g, dtype=torch.float64)
  w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
  xr = x.clone().requires_grad_(True)
  wr = w.clone().requires_grad_(True)
  want = F.conv2d(xr.permute(0, 3, 1, 2), wr, stride=2, padding=1).permute(0, 2, 3, 1)     # (B, H/2, W/2, Cout)
  planes = parity_split(x)
  got = emulate_conv_gemm(planes, ops.pack_conv_weight(w, dt=torch.float64), ops.taps_3x3_stride2(b), batch=b)
  assert torch.allclose(got, want.detach(), atol=1e-12)
  dy = torch.randn(want.shape, generator=g, dtype=torch.float64)
  (want * dy).sum().backward()
  wt = ops.pack_conv_weight_t(w, dt=torch.float64)
  dx = torch.zeros_like(x)
  for py in (0, 1):
    for px in (0, 1):
      dx[:, py::2, px::2] = emulate_conv_gemm(dy, wt, ops.taps_3x3_stride2_dgrad(py, px))
  assert torch.allclose(dx, xr.grad, atol=1e-12)
  dw = emulate_conv_wgrad(dy, planes, ops.taps_3x3_stride2(b), 9)                            # (Cout, 9, Cin)
  assert torch.allclose(dw.transpose(1, 2).reshape(cout, cin, 3, 3), wr.grad, atol=1e-12)


def test_n_tile_choice_is_valid_and_fills_the_machine():
  """ops.pick_bn_for: legal tile widths, the wide tile for many-round problems, no worse than the wide tile otherwise."""
  for n in (8, 21, 32, 64, 66, 72, 128, 216, 256, 320, 512, 576, 1512, 2048, 4536, 6048):
    for m_tiles in (1, 3, 16, 17, 64, 80, 256, 1024, 8192):
      bn = ops.pick_bn_for(n, m_tiles)
      assert 16 <= bn <= 256 and bn % 16 == 0
      wide = ops.pick_bn(n)
      cost = lambda t: -(-(m_tiles * -(-n // t)) // 148) * (128 + t)
      if m_tiles * -(-n // wide) >= 4 * 148:
        assert bn == wide
      else:
        assert cost(bn) <= cost(wide) * 1.26   # (the padding penalty may trade a little of the round model)
