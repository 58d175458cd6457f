"""Per-shape roofline of the tcgen05 GEMM family from the per-shape timing table bench.py dumps (TFPP_GEMM_DUMP):
ideal time = max(algorithmic bytes / HBM rate, FLOPs / tensor rate) with the measured peaks of MEASURED_PEAKS.json
(6.57 TB/s copy, 1437.7 TFLOP/s sustained cuBLAS bf16).  usage: python tools/gemm_roofline.py profiles/r02_final_gemm_shapes.txt"""
import re
import sys

HBM, TENSOR = 6565.8e9, 1437.7e12
rows = []
for line in open(sys.argv[1]):
  m = re.match(r'\s*([\d.]+) ms\s+n=\s*(\d+)\s+avg\s+([\d.]+) us\s+([\d.]+) TFLOP/s\s+(gemm|wgrad)\s+b(\d+) (\d+)x(\d+) (.*)', line)
  if not m:
    continue
  tot_ms, n, avg_us, tf, kind, b, h, w, rest = m.groups()
  n, avg_us, b, h, w = int(n), float(avg_us), int(b), int(h), int(w)
  pix = b * h * w
  kv = dict(re.findall(r'([a-z]+)(\d+)', rest))
  taps = int(kv.get('taps', 1))
  if kind == 'gemm':
    nn, k = int(kv['n']), int(kv['k'])
    k_alg = 24 if kv.get('grp') == '1' else k
    flops = 2.0 * pix * nn * k_alg * taps
    byts = 2.0 * (pix * k + nn * k * taps + pix * nn)            # A once, weights once, output once (bf16)
  else:
    co, ci = int(kv['cout']), int(kv['cin'])
    gw = int(kv.get('grp', 0))
    flops = 2.0 * pix * co * (gw or ci) * taps
    byts = 2.0 * pix * (co + ci) + 4.0 * co * (gw or ci) * taps  # dy + x once (bf16), dW fp32
  ideal_us = max(byts / HBM, flops / TENSOR) * 1e6
  rows.append((float(tot_ms), n, avg_us, float(tf), ideal_us, 'hbm' if byts / HBM > flops / TENSOR else 'tensor', kind, f'b{b} {h}x{w} {rest.strip()}', flops))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
ideal = sum(r[1] * r[4] for r in rows) / 1e3
fl = sum(r[1] * r[8] for r in rows)
print(f'# {len(rows)} shapes, {sum(r[1] for r in rows)} launches, {tot:.2f} ms measured (CUDA events around every launch: includes ~4 us of launch gap each),')
print(f'# roofline-ideal {ideal:.2f} ms  ->  family at {ideal / tot:.2f} of its per-shape roofline; {fl / tot / 1e9:.0f} TFLOP/s average = {fl / tot / 1e9 / 1437.7:.3f} of the sustained tensor peak')
print('#   total ms  launches  avg us  TFLOP/s  ideal us  bound   frac   shape')
for t, n, a, tf, i, bound, kind, shape, _ in rows[:60]:
  print(f'  {t:9.3f}  {n:8d}  {a:6.1f}  {tf:7.1f}  {i:8.1f}  {bound:6s}  {i / a:5.2f}   {kind} {shape}')
