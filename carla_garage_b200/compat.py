"""Import shims so the *unmodified* reference modules (``/root/reference/team_code``)
import under this image (torch 2.11, NumPy 2, no carla / timm / laspy / ujson / imgaug / shapely).

Nothing here computes anything on the hot path; it only makes ``import config``, ``import model`` … succeed
(SURVEY.md §8c lists what is missing).  ``timm.create_model`` is routed to a caller-supplied factory:

* tests / golden generation pass the plain-torch RegNetY restatement from ``oracle/regnety.py`` so that the
  reference's ``LidarCenterNet`` runs verbatim on CPU;
* the drop-in package never needs it (the product backbone builds its own CUDA RegNet).

Reference call sites that force each stub: ``team_code/config.py:6,16-24`` (carla.WeatherParameters),
``team_code/data.py:6,15,22`` (ujson, laspy, imgaug), ``team_code/transfuser_utils.py:15-16`` (shapely),
``team_code/transfuser.py:9,25,52`` (timm), ``team_code/video_swin_transformer.py:11`` (timm.models.layers),
``team_code/data.py:212-226`` (np.string_).
"""
import os
import sys
import types

REFERENCE_TEAM_CODE = os.environ.get('TFPP_REFERENCE_DIR', '/root/reference/team_code')


class _Anything:
  """Attribute/call sink: every attribute access or call returns another sink."""

  def __init__(self, *a, **k):
    pass

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return _Anything()

  def __call__(self, *a, **k):
    return _Anything()


def _module(name, **attrs):
  m = types.ModuleType(name)
  m.__dict__.update(attrs)

  def _fallback(attr):  # PEP 562 fallback for anything not listed
    if attr.startswith('__'):
      raise AttributeError(attr)
    return _Anything()

  m.__getattr__ = _fallback
  sys.modules[name] = m
  return m


def install(timm_factory=None, add_reference_to_path=True):
  """Install the stubs (idempotent). ``timm_factory(name, pretrained, features_only, in_chans)`` -> nn.Module."""
  import numpy as np
  import torch

  if not hasattr(np, 'string_'):
    np.string_ = np.bytes_
  for alias, target in (('float', float), ('int', int)):
    if alias not in np.__dict__:
      setattr(np, alias, target)

  if 'carla' not in sys.modules:
    weather = type('WeatherParameters', (), {})
    for preset in ('ClearNoon', 'CloudySunset', 'WetSunset', 'MidRainSunset', 'WetCloudySunset', 'HardRainNoon',
                   'SoftRainSunset', 'ClearSunset'):
      setattr(weather, preset, preset)
    _module('carla', WeatherParameters=weather)
  for name in ('laspy', 'ujson', 'shapely', 'filterpy', 'filterpy.kalman', 'torchmetrics', 'diskcache_stub'):
    if name not in sys.modules:
      try:
        __import__(name)
      except Exception:  # pylint: disable=broad-except
        _module(name)
  if 'shapely.geometry' not in sys.modules:
    try:
      __import__('shapely.geometry')
    except Exception:  # pylint: disable=broad-except
      geom = _module('shapely.geometry', Polygon=_Anything)
      sys.modules['shapely'].geometry = geom
  if 'imgaug' not in sys.modules:
    try:
      __import__('imgaug')
    except Exception:  # pylint: disable=broad-except
      aug = _module('imgaug.augmenters')
      ia = _module('imgaug', augmenters=aug)
      del ia

  if 'timm' not in sys.modules or getattr(sys.modules['timm'], '_tfpp_shim', False):
    try:
      import timm  # noqa: F401  pylint: disable=unused-import,import-outside-toplevel
      real_timm = not getattr(sys.modules['timm'], '_tfpp_shim', False)
    except Exception:  # pylint: disable=broad-except
      real_timm = False
    if not real_timm:

      def create_model(name, pretrained=False, features_only=False, in_chans=3, **kwargs):
        del kwargs
        if timm_factory is None:
          raise RuntimeError('timm is not installed and no timm_factory was given to compat.install()')
        return timm_factory(name, pretrained=pretrained, features_only=features_only, in_chans=in_chans)

      class DropPath(torch.nn.Module):
        """Stochastic depth; identity when p == 0 or in eval (timm.models.layers.DropPath)."""

        def __init__(self, drop_prob=0.0):
          super().__init__()
          self.drop_prob = drop_prob

        def forward(self, x):
          if self.drop_prob == 0.0 or not self.training:
            return x
          keep = 1.0 - self.drop_prob
          mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
          return x * mask / keep

      layers = _module('timm.models.layers', DropPath=DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
      models = _module('timm.models', layers=layers)
      t = _module('timm', create_model=create_model, models=models)
      t._tfpp_shim = True  # pylint: disable=protected-access

  if add_reference_to_path and os.path.isdir(REFERENCE_TEAM_CODE) and REFERENCE_TEAM_CODE not in sys.path:
    sys.path.insert(0, REFERENCE_TEAM_CODE)


def reference_available():
  return os.path.isdir(REFERENCE_TEAM_CODE)
