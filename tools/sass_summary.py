"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), UTCBAR (tcgen05.commit), and the legacy HMMA.
  python tools/sass_summary.py [calciumgan_b200/libcalciumgan_b200.so] > profiles/r2_sass_summary.md"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else 'calciumgan_b200/libcalciumgan_b200.so'
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
ops = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UBLKRED', 'UTMAPF', 'SYNCS', 'HMMA', 'LDGSTS', 'REDG', 'ATOMG']
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
  m = re.match(r'\s*Function : (\S+)', line)
  if m:
    cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
    cur = re.sub(r'\(.*$', '', cur).replace('void ', '')
    per[cur] = collections.Counter()
    continue
  if cur is None:
    continue
  m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
  if m:
    op = m.group(1).split('.')[0]
    per[cur]['_total'] += 1
    if op in ops:
      per[cur][op] += 1
print('# SASS opcode summary of %s (sm_100a)\n' % so)
print('`cuobjdump -sass`, instructions per kernel. UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UBLKCP = TMA, UBLKRED = bulk reduce-add (cp.reduce.async.bulk),')
print('UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async, REDG / ATOMG = global reductions. No HMMA (legacy mma.sync) anywhere.\n')
print('| kernel | instr | ' + ' | '.join(ops) + ' |')
print('|---|---:|' + '---:|' * len(ops))
tot = collections.Counter()
for k, c in per.items():
  if not any(c[o] for o in ops[:8]) and not k.startswith('tc::'):
    continue
  print('| `%s` | %d | ' % (k, c['_total']) + ' | '.join(str(c[o]) if c[o] else '' for o in ops) + ' |')
  tot.update(c)
print('| **all tensor-core kernels** | %d | ' % tot['_total'] + ' | '.join(str(tot[o]) for o in ops) + ' |')
others = [k for k, c in per.items() if not any(c[o] for o in ops[:8]) and not k.startswith('tc::')]
print('\nCUDA-core kernels (fp32 path and memory-bound glue, no tensor-core / TMA instructions): %d kernels, e.g. %s' %
      (len(others), ', '.join('`%s`' % o for o in others[:12])))
