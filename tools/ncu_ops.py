"""Aggregate `ncu --page source` instruction counts of one kernel by SASS opcode.
usage: python tools/ncu_ops.py <report.ncu-rep> <kernel substring> [top]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + kern], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(src.split('\n')) if len(r) > 5]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
ops, samp = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[0] == hdr[0]:
        break
    sass = r[idx['Source']].strip()
    if sass.startswith('@'):
        sass = sass.split(None, 1)[1]
    op = sass.split()[0].split('.')[0] if sass else '?'
    ops[op] += int(r[idx['Instructions Executed']]); samp[op] += int(r[idx['# Samples']])
ti, ts = sum(ops.values()), sum(samp.values())
print('total warp-instr', ti)
for op, c in ops.most_common(top):
    print(f'{op:12s} inst {100 * c / ti:5.1f}%  samp {100 * samp[op] / ts:5.1f}%')
