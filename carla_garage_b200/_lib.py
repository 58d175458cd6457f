"""ctypes binding of libtfpp.so (include/tfpp.h).  Fails loudly when the library is missing: there is no CPU or
torch fallback anywhere in the product path."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libtfpp.so')
_lib = None

c_int, c_float, c_ll, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_void_p


class ConvGemmArgs(ctypes.Structure):
  """tfpp_conv_gemm_args (include/tfpp.h)."""
  _fields_ = [
      ('a', c_void_p), ('a_batch', c_int), ('height', c_int), ('width', c_int), ('a_channels', c_int),
      ('a_batch_stride', c_ll),
      ('w', c_void_p), ('w_taps', c_int), ('w_kdim', c_int), ('n', c_int),
      ('batch', c_int), ('k_per_tile', c_int), ('a_c_per_ntile', c_int), ('bn', c_int),
      ('tw', c_int), ('th', c_int), ('nb', c_int), ('ntaps', c_int),
      ('tap_dx', c_int * 9), ('tap_dy', c_int * 9), ('tap_db', c_int * 9), ('tap_w', c_int * 9),
      ('out', c_void_p), ('out_f32', c_int),
      ('o_sb', c_ll), ('o_sy', c_ll), ('o_sx', c_ll), ('o_sn', c_ll),
      ('res1', c_void_p), ('res1_f32', c_int),
      ('r1_sb', c_ll), ('r1_sy', c_ll), ('r1_sx', c_ll), ('r1_sn', c_ll),
      ('res2', c_void_p), ('res2_f32', c_int),
      ('r2_sb', c_ll), ('r2_sy', c_ll), ('r2_sx', c_ll), ('r2_sn', c_ll),
      ('scale', c_void_p), ('shift', c_void_p), ('act', c_int), ('act_n_limit', c_int),
      ('stat_sum', c_void_p), ('stat_sq', c_void_p),
      ('drop_rng', c_void_p), ('drop_p', c_float), ('drop_site', ctypes.c_uint),
  ]


class WgradArgs(ctypes.Structure):
  """tfpp_wgrad_args (include/tfpp.h)."""
  _fields_ = [
      ('dy', c_void_p), ('x', c_void_p), ('dw', c_void_p),
      ('batch', c_int), ('height', c_int), ('width', c_int), ('cout', c_int), ('cout_valid', c_int),
      ('x_batch', c_int), ('x_channels', c_int), ('x_batch_stride', c_ll),
      ('cin', c_int), ('group_width', c_int), ('dw_s_co', c_ll), ('dw_s_tap', c_ll), ('dw_s_ci', c_ll), ('ntaps', c_int),
      ('tap_dx', c_int * 9), ('tap_dy', c_int * 9), ('tap_db', c_int * 9), ('tap_w', c_int * 9),
      ('tw', c_int), ('th', c_int), ('nb', c_int), ('bn', c_int), ('splits', c_int),
  ]


class PeerStepArgs(ctypes.Structure):
  """tfpp_peer_step_args (include/tfpp.h)."""
  _fields_ = [
      ('world', c_int), ('rank', c_int),
      ('grad', c_void_p * 8), ('param', c_void_p * 8), ('flags', c_void_p * 8),
      ('exp_avg', c_void_p), ('exp_avg_sq', c_void_p), ('max_exp_avg_sq', c_void_p),
      ('n', c_ll), ('beta1', c_float), ('beta2', c_float), ('eps', c_float), ('weight_decay', c_float),
      ('dev_state', c_void_p), ('opt_flags', c_void_p),
  ]


P, I, F, L = c_void_p, c_int, c_float, c_ll
_PROTOS = {
    'tfpp_abi_version': [],
    'tfpp_pillar_scatter': [P, I, I, P, P, I, F, F, F, F, F, I, F, F, P],
    'tfpp_pillar_scatter_aligned': [P, P, I, I, I, P, P, I, F, F, F, F, F, I, ctypes.c_double, F, P],
    'tfpp_instnorm_stats': [P, I, L, I, I, I, P, P, P],
    'tfpp_instnorm_apply': [P, I, L, P, P, F, I, P, L, P, P, I, I, I, P],
    'tfpp_instnorm_bwd': [P, L, P, I, P, P, I, P, P, P, I, I, I, P],
    'tfpp_bev_lift': [P, I, P, P, P, P, P, I, I, I, I, I, I, P],
    'tfpp_bev_lift_bwd': [P, I, P, P, P, P, P, P, I, I, I, I, I, I, I, P],
    'tfpp_conv_gemm': [ctypes.POINTER(ConvGemmArgs), P],
    'tfpp_conv_wgrad': [ctypes.POINTER(WgradArgs), P],
    'tfpp_smallc_conv3x3': [P, P, P, P, I, I, I, I, I, I, I, I, I, P],
    'tfpp_smallc_wgrad3x3': [P, P, P, L, L, L, I, I, I, I, I, I, P],
    'tfpp_stem_conv': [P, P, P, P, P, P, I, P, P, P, I, I, I, I, P],
    'tfpp_bn_finalize': [P, P, P, P, P, P, P, P, P, P, I, F, F, F, P],
    'tfpp_scale_shift_act': [P, P, P, P, P, P, I, P, P, I, I, I, P],
    'tfpp_se_gate': [P, I, P, P, P, P, P, P, I, I, I, P],
    'tfpp_channel_scale': [P, P, P, I, I, I, P],
    'tfpp_parity_split': [P, P, I, I, I, I, P],
    'tfpp_avgpool_tokens': [P, P, P, I, I, I, I, I, I, I, I, I, P],
    'tfpp_bilinear': [P, I, L, L, P, P, I, I, I, I, I, I, P],
    'tfpp_bilinear_nchw_mask': [P, P, P, I, I, I, I, I, I, I, P],
    'tfpp_nchw_f32_to_nhwc_bf16': [P, P, I, I, I, P],
    'tfpp_nhwc_bf16_to_nchw_f32': [P, P, I, I, I, P],
    'tfpp_layernorm': [P, I, P, P, P, P, P, P, I, I, F, P],
    'tfpp_fusion_attn': [P, P, I, I, I, I, P],
    'tfpp_small_mha': [P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, P],
    'tfpp_extra_sensor_token': [P, P, F, F, I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, P],
    'tfpp_planner_head': [P] * 17 + [I, I, I, I, I, P],
    'tfpp_decode_heatmap': [P, L, P, L, P, L, P, L, P, L, P, I, I, I, I, I, I, F, F, P],
    'tfpp_bn_bwd': [P, P, P, P, P, P, P, P, P, P, I, P, P, P, P, I, I, I, P],
    'tfpp_gconv3x3': [P, P, P, P, P, I, P, P, I, I, I, I, I, P],
    'tfpp_gconv3x3_dgrad_s2': [P, P, P, I, I, I, I, P],
    'tfpp_gconv3x3_wgrad': [P, P, P, P, I, I, I, I, I, P],
    'tfpp_gconv3x3_wgrad_workspace': [I, I, I, I, I],
    'tfpp_halo_gconv3x3': [P, P, P, P, P, I, P, P, I, I, I, I, P],
    'tfpp_halo_conv3x3': [P, P, P, P, I, I, I, I, I, I, I, I, I, P],
    'tfpp_gather_pack': [P, P, P, L, I, P],
    'tfpp_se_bwd': [P, P, P, P, P, I, P, P, P, P, P, P, P, P, P, I, I, I, P],
    'tfpp_act_bwd': [P, P, I, I, I, F, P, P, I, I, I, I, P],
    'tfpp_bilinear_bwd': [P, P, I, L, L, I, I, I, I, I, I, I, P],
    'tfpp_bilinear_nchw_mask_bwd': [P, P, P, I, I, I, I, I, I, I, P],
    'tfpp_pool_bwd_add': [P, P, I, P, I, I, I, I, I, I, I, I, P],
    'tfpp_parity_merge': [P, P, I, I, I, I, P],
    'tfpp_add_bf16': [P, P, P, L, P],
    'tfpp_cast_f32_bf16': [P, P, L, P],
    'tfpp_cast_rows': [P, P, P, I, I, I, I, I, P],
    'tfpp_batch_reduce': [P, P, I, L, P],
    'tfpp_stem_wgrad': [P, P, P, P, P, I, I, I, I, P],
    'tfpp_layernorm_bwd': [P, I, P, P, P, P, P, P, P, P, I, I, P],
    'tfpp_fusion_attn_bwd': [P, P, P, P, I, I, I, I, P],
    'tfpp_small_mha_bwd': [P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, I, P],
    'tfpp_extra_sensor_token_bwd': [P, P, F, F, I, P, P, P, P, P, L, P, P, P, P, P, I, I, I, I, P],
    'tfpp_planner_head_bwd': [P] * 28 + [I, I, I, I, I, P],
    'tfpp_ce_map_loss': [P, P, P, F, P, P, P, P, P, I, I, I, I, P],
    'tfpp_l1_sigmoid_loss': [P, P, F, P, P, P, P, I, L, P],
    'tfpp_center_head_loss': [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, P],
    'tfpp_planner_loss': [P, P, P, P, P, F, F, P, P, P, P, I, I, I, P],
    'tfpp_dropout': [P, I, L, P, F, I, P],
    'tfpp_act_bwd_dropout': [P, P, I, I, I, F, P, P, I, I, I, I, P, F, I, P],
    'tfpp_fusion_attn_dropout': [P, P, I, I, I, I, P, F, I, P],
    'tfpp_fusion_attn_bwd_dropout': [P, P, P, P, I, I, I, I, P, F, I, P],
    'tfpp_small_mha_dropout': [P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, P, F, I, P],
    'tfpp_small_mha_bwd_dropout': [P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, I, P, F, I,
                                   P],
    'tfpp_adamw_amsgrad': [P, P, P, P, P, L, F, F, F, F, F, I, F, P, P, P],
    'tfpp_gru_cell_head': [P, I, P] + [P] * 10 + [P, P, P, I, I, I, I, I, I, P],
    'tfpp_gru_cell_head_bwd': [P, I, P] + [P] * 9 + [P] * 4 + [P] * 11 + [I, I, I, I, I, I, P],
    'tfpp_centernet_targets': [P, P, I, I, I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P],
    'tfpp_nms_rotated': [P, I, I, I, F, F, I, F, F, F, P, P, P, P],
    # NVLink peer-memory gradient exchange fused with AdamW (csrc/peer_exchange.cu)
    'tfpp_peer_alloc': [L, ctypes.POINTER(c_void_p), P],
    'tfpp_peer_open': [P, ctypes.POINTER(c_void_p)],
    'tfpp_peer_close': [P],
    'tfpp_peer_free': [P],
    'tfpp_peer_adamw_step': [ctypes.POINTER(PeerStepArgs), P],
    'tfpp_peer_barrier': [ctypes.POINTER(PeerStepArgs), I, P],
    # fp32 parity mode (csrc/fp32_path.cu)
    'tfpp_conv_gemm_f32': [ctypes.POINTER(ConvGemmArgs), P],
    'tfpp_gconv3x3_f32': [P, P, P, P, P, I, P, P, I, I, I, I, I, P],
    'tfpp_stem_conv_f32': [P, P, P, P, P, P, I, P, P, P, I, I, I, I, P],
    'tfpp_scale_shift_act_f32': [P, P, P, P, P, P, I, P, P, I, I, I, P],
    'tfpp_channel_scale_f32': [P, P, P, I, I, I, P],
    'tfpp_parity_split_f32': [P, P, I, I, I, I, P],
    'tfpp_avgpool_tokens_f32': [P, P, P, I, I, I, I, I, I, I, I, P],
    'tfpp_bilinear_f32': [P, L, L, P, P, I, I, I, I, I, I, P],
    'tfpp_bilinear_nchw_mask_f32': [P, P, P, I, I, I, I, I, I, I, P],
    'tfpp_mha_f32': [P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, P, F, I, P],
    'tfpp_conv_wgrad_f32': [ctypes.POINTER(WgradArgs), P],
    'tfpp_gconv3x3_wgrad_f32': [P, P, P, I, I, I, I, I, P],
    'tfpp_gconv3x3_dgrad_s2_f32': [P, P, P, I, I, I, I, P],
    'tfpp_stem_wgrad_f32': [P, P, P, P, P, I, I, I, I, P],
    'tfpp_bn_bwd_f32': [P, P, P, P, P, P, P, P, P, P, I, P, P, P, P, I, I, I, P],
    'tfpp_se_bwd_reduce_f32': [P, P, P, I, I, I, P],
    'tfpp_act_bwd_f32': [P, P, I, I, I, F, P, P, I, I, I, I, P, F, I, P],
    'tfpp_bilinear_bwd_f32': [P, P, L, L, I, I, I, I, I, I, I, P],
    'tfpp_bilinear_nchw_mask_bwd_f32': [P, P, P, I, I, I, I, I, I, I, P],
    'tfpp_pool_bwd_add_f32': [P, P, P, I, I, I, I, I, I, I, I, P],
    'tfpp_add_f32': [P, P, P, L, P],
    'tfpp_copy_rows_f32': [P, P, P, I, I, I, I, I, P],
    'tfpp_mha_bwd_f32': [P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, L, L, P, I, I, I, I, I, I, P, F, I, P],
}


def exported_symbols():
  return sorted(_PROTOS) + ['tfpp_last_error']


def load():
  """Load libtfpp.so once; raise if it was not built (run `python __graft_entry__.py`)."""
  global _lib  # pylint: disable=global-statement
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(f'{LIB_PATH} is missing: build the CUDA extension first (python __graft_entry__.py). '
                       'carla_garage_b200 has no CPU/torch fallback.')
  lib = ctypes.CDLL(LIB_PATH)
  lib.tfpp_last_error.restype = ctypes.c_char_p
  lib.tfpp_last_error.argtypes = []
  for name, args in _PROTOS.items():
    fn = getattr(lib, name)
    fn.restype = c_ll if name.endswith('_workspace') else c_int
    fn.argtypes = args
  _lib = _Timed(lib)
  return _lib


class _Timed:
  """The loaded library; with ``CALL_PROFILE`` set (profile_calls) every entry point that takes a stream is bracketed by
  CUDA events on the launching stream, so that bench.py can report where an eager step spends its device time."""

  def __init__(self, lib):
    object.__setattr__(self, '_lib', lib)

  def __getattr__(self, name):
    fn = getattr(self._lib, name)
    if CALL_PROFILE[0] is None or not name.startswith('tfpp_') or name.endswith(('_workspace', 'last_error')) or \
        name.startswith('tfpp_peer_') and not name.endswith(('_step', '_barrier')):
      return fn

    def timed(*args):
      import torch  # pylint: disable=import-outside-toplevel
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      rc = fn(*args)
      e1.record()
      CALL_PROFILE[0].append((name, e0, e1))
      return rc
    return timed


CALL_PROFILE = [None]


def profile_calls(fn):
  """Run fn() once with CUDA events around every C-ABI call; returns {entry point: (milliseconds, calls)}.  Meaningful on
  ONE stream only (overlapping streams would charge a kernel for its neighbours)."""
  import torch  # pylint: disable=import-outside-toplevel
  CALL_PROFILE[0] = []
  try:
    fn()
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1 in CALL_PROFILE[0]:
      a = agg.setdefault(name, [0.0, 0])
      a[0] += e0.elapsed_time(e1)
      a[1] += 1
  finally:
    CALL_PROFILE[0] = None
  return {k: (v[0], v[1]) for k, v in agg.items()}


# kernels launched per C-ABI call (for bench.py's gpu_launches claim); default 1
_KERNELS_PER_CALL = {'tfpp_peer_adamw_step': 4, 'tfpp_adamw_amsgrad': 2, 'tfpp_pillar_scatter': 2, 'tfpp_pillar_scatter_aligned': 2, 'tfpp_bn_bwd': 2, 'tfpp_instnorm_bwd': 2, 'tfpp_bev_lift_bwd': 2, 'tfpp_se_bwd': 4, 'tfpp_se_gate': 2, 'tfpp_gconv3x3_wgrad': 2, 'tfpp_fusion_attn_bwd': 2}
_LAUNCHES = [0]


def reset_launch_count():
  _LAUNCHES[0] = 0


def launch_count():
  return _LAUNCHES[0]


def check(rc, what):
  _LAUNCHES[0] += _KERNELS_PER_CALL.get(what, 1)
  if rc != 0:
    raise RuntimeError(f'{what} failed (rc={rc}): {load().tfpp_last_error().decode()}')
