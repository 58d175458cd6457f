"""SASS census of carla_garage_b200/lib/libtfpp.so: per kernel, how many tcgen05 (UTC*MMA), TMEM (LDTM/STTM), TMA
(UTMALDG/UTMASTG/UBLKCP), legacy tensor (HMMA) and cp.async (LDGSTS) instructions it contains.
  python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'carla_garage_b200', 'lib', 'libtfpp.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([('UTC*MMA (tcgen05.mma)', r'\bUTC\w*MMA'), ('LDTM/STTM (tcgen05.ld/st)', r'\b(LDTM|STTM)'),
                                ('UTMALDG/UTMASTG/UBLKCP (TMA)', r'\b(UTMALDG|UTMASTG|UBLKCP)'), ('HMMA (mma.sync)', r'\bHMMA'),
                                ('LDGSTS (cp.async)', r'\bLDGSTS'), ('RED/ATOM', r'\b(RED|ATOM|ATOMG)\b'),
                                ('instructions', r'^\s+/\*[0-9a-f]{4,6}\*/')])
cur, rows = None, collections.OrderedDict()
for line in sass.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    name = re.sub(r'\(.*', '', name)
    cur = rows.setdefault(name, collections.Counter())
    continue
  if cur is None:
    continue
  for k, p in pats.items():
    if re.search(p, line):
      cur[k] += 1
print(f'# SASS census of {os.path.relpath(lib, ROOT)} (sm_100a), {len(rows)} kernels; columns: ' + ' | '.join(pats))
tot = collections.Counter()
for name, c in sorted(rows.items(), key=lambda kv: -(kv[1]['UTC*MMA (tcgen05.mma)'] * 1000 + kv[1]['HMMA (mma.sync)'])):
  print(f'{name[:70]:70s} ' + ' '.join(f'{c[k]:6d}' for k in pats))
  tot.update(c)
print(f'{"TOTAL":70s} ' + ' '.join(f'{tot[k]:6d}' for k in pats))
