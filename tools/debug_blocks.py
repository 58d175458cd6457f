"""Per-block isolation on the GPU: feed each engine block the ORACLE's input and compare with the oracle's output."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carla_garage_b200 import synth, ops
from carla_garage_b200.config import GlobalConfig
from carla_garage_b200.nn import LidarCenterNet
from oracle import tfpp_oracle as orc

G = os.path.join(ROOT, 'tests', 'golden')
sd = synth.golden_state(G)
net = LidarCenterNet(GlobalConfig()); net.load_state_dict(sd); net = net.cuda().eval()
eng = net.engine
training = '--train' in sys.argv
if training:
  net.train()

def rel(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return float((a - b).norm() / (b.norm() + 1e-30))

def to_dev(x):  # NCHW f32 cpu -> NHWC bf16 cuda
  return ops.nchw_to_nhwc(x.cuda().contiguous())

torch.set_num_threads(os.cpu_count())
inp = synth.make_inputs(2, seed=11)
with torch.no_grad():
  x_img = orc._conv_bn(sd, 'backbone.image_encoder.stem', orc.normalize_imagenet(inp['rgb']), training, stride=2)
  x_lid = orc._conv_bn(sd, 'backbone.lidar_encoder.stem', inp['lidar_bev'], training, stride=2)
  g_img = eng.stem(inp['rgb'].cuda(), net.backbone.image_encoder['stem'], training, True)
  g_lid = eng.stem(inp['lidar_bev'].cuda(), net.backbone.lidar_encoder['stem'], training, False)
  print('stem img', rel(ops.nhwc_to_nchw(g_img), x_img), 'lid', rel(ops.nhwc_to_nchw(g_lid), x_lid))
  for i in range(4):
    for name, enc in (('image', net.backbone.image_encoder), ('lidar', net.backbone.lidar_encoder)):
      x = x_img if name == 'image' else x_lid
      for j, blk in enumerate(enc[f's{i+1}']):
        p = f'backbone.{name}_encoder.s{i+1}.b{j+1}'
        want = orc.regnet_block(sd, p, x, training, stride=2 if j == 0 else 1)
        got = ops.nhwc_to_nchw(eng.regnet_block(to_dev(x), blk, training))
        print(f'{name} s{i+1}.b{j+1} C={want.shape[1]} {tuple(want.shape[2:])} rel={rel(got, want):.3e} '
              f'|x|={float(x.norm()):.3e} |y|={float(want.norm()):.3e}', flush=True)
        x = want
      if name == 'image': x_img = x
      else: x_lid = x
    w_img, w_lid = orc.fuse_features(sd, 'backbone', x_img, x_lid, i, orc.DEFAULT_CFG | dict(vars(GlobalConfig())) if False else orc.DEFAULT_CFG)
    g_img, g_lid = eng.fuse(to_dev(x_img), to_dev(x_lid), i, training)
    print(f'fuse {i}: img {rel(ops.nhwc_to_nchw(g_img), w_img):.3e} lid {rel(ops.nhwc_to_nchw(g_lid), w_lid):.3e} '
          f'delta img {rel(ops.nhwc_to_nchw(g_img) - x_img.cuda(), w_img - x_img):.3e}', flush=True)
    x_img, x_lid = w_img, w_lid
