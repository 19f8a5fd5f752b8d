"""SASS opcode histogram of every kernel in libsdfnmpc_b200.so (evidence of the Blackwell-native instructions:
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, DMMA = mma.sync f64, LDGSTS = cp.async).
    python tools/sass_hist.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'sdf_nmpc_b200', 'lib', 'libsdfnmpc_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kern, hist, size = None, collections.OrderedDict(), {}
for line in out.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(anonymous namespace\)::|sdfnmpc::|void ', '', kern).split('(')[0]
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ['UTCHMMA', 'UTCBAR', 'UTCATOMSWS', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'DMMA', 'HMMA', 'LDGSTS', 'SYNCS', 'DFMA', 'FFMA', 'MUFU', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL', 'BAR', 'WARPSYNC']
print('# SASS opcode histogram per kernel (`cuobjdump -sass sdf_nmpc_b200/lib/libsdfnmpc_b200.so`, sm_100a)\n')
print('| kernel | instructions | ' + ' | '.join(KEY) + ' |')
print('|---|---|' + '---|' * len(KEY))
for k, c in hist.items():
    print(f'| `{k}` | {sum(c.values())} | ' + ' | '.join(str(c.get(o, 0)) for o in KEY) + ' |')
print('\nUTCHMMA / UTCBAR / UTCATOMSWS / LDTM = tcgen05.mma / commit / alloc / ld (5th-generation tensor cores, TMEM); UBLKCP = cp.async.bulk '
      '(bulk-copy engine); DMMA = fp64 tensor-core mma.sync (the QP kernel); LDGSTS = cp.async; LDL / STL = local-memory (spill) traffic.')
