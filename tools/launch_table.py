"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches / total us / share, for the
launches [start, end) (default: the second half = the second, warm step of tools/profile_step.py)."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path) as f:
  lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
  if r.get('Metric Name') != 'gpu__time_duration.sum':
    continue
  v = float(r['Metric Value'].replace(',', ''))
  unit = r['Metric Unit']
  us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
  name = re.sub(r'\(.*$', '', r['Kernel Name'])
  rows.append((name, us))
n = len(rows)
start = int(sys.argv[2]) if len(sys.argv) > 2 else n // 2
end = int(sys.argv[3]) if len(sys.argv) > 3 else n
sel = rows[start:end]
agg = OrderedDict()
for name, us in sel:
  a = agg.setdefault(name, [0, 0.0])
  a[0] += 1
  a[1] += us
tot = sum(a[1] for a in agg.values())
print('| kernel | launches | total us | share |\n|---|---:|---:|---:|')
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print('| `%s` | %d | %.1f | %.1f%% |' % (name, cnt, us, 100 * us / tot))
print('\nsum %.2f ms over %d launches (launches %d..%d of %d)' % (tot / 1e3, len(sel), start, end, n))
