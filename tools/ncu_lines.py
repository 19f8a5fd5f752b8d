"""Attribute ncu SASS-level counters to CUDA source lines: nvdisasm line info of the cubin x `ncu --page source --csv`.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel substring> <cubin file stem (sdf_api|nmpc_api|encoder_api)> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern, stem = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
binary = sys.argv[5] if len(sys.argv) > 5 else None      # object / library the report was taken with (default: in-tree build)
ranges = sys.argv[6] if len(sys.argv) > 6 else None      # 'name:lo-hi,name:lo-hi' line ranges of the phase file to aggregate
phase_file = sys.argv[7] if len(sys.argv) > 7 else 'qp_core.cuh'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', binary or os.path.join(root, 'sdf_nmpc_b200/lib/libsdfnmpc_b200.so')], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith(stem)][0]
sass = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
amap, cur, infunc = {}, None, False
for l in sass:
    if l.startswith('.text.'):
        infunc = kern in l
        continue
    if l.startswith('.section'):
        infunc = False
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        amap[int(m.group(1), 16)] = (cur, m.group(2))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + re.split(r'I[A-Z]', kern)[0]], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split('\n')))
rows = [r for r in rows if len(r) > 5]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[1:]:          # first launch of the kernel only (the report may hold several)
    if r[0] == hdr[0]:
        break
    data.append(r)
base = int(data[0][0], 16)
inst, samp, stall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in data:
    off = int(r[0], 16) - base
    key = amap.get(off, (('?', 0), ''))[0]
    inst[key] += int(r[idx['Instructions Executed']]); samp[key] += int(r[idx['# Samples']])
    for c in stall_cols:
        stall[key][c] += int(r[idx[c]])
ti, ts = sum(inst.values()), sum(samp.values())
print('total warp-instr', ti, 'samples', ts)
files = {}
for key, c in samp.most_common(top):
    f, n = key
    if f not in files:
        try:
            files[f] = open(os.path.join(root, 'sdf_nmpc_b200/csrc', f)).read().split('\n')
        except OSError:
            files[f] = []
    txt = files[f][n - 1].strip()[:80] if files[f] and 0 < n <= len(files[f]) else ''
    st = ' '.join(f'{k[6:]}={v}' for k, v in stall[key].most_common(2))
    print(f'{f}:{n:4d} samp {100 * c / ts:5.1f}% inst {100 * inst[key] / ti:5.1f}% [{st}] {txt}')

if ranges:
    print('--- phases (qp_core.cuh line ranges) ---')
    acc = collections.OrderedDict()
    spec = [(r.split(':')[0], *map(int, r.split(':')[1].split('-'))) for r in ranges.split(',')]
    for key in samp:
        f, n = key
        name = 'other:' + f
        if f == phase_file:
            name = phase_file + ':unassigned'
            for nm, lo, hi in spec:
                if lo <= n <= hi:
                    name = nm
                    break
        a = acc.setdefault(name, [0, 0])
        a[0] += samp[key]; a[1] += inst[key]
    for name, (sa, ins) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f'{name:28s} samp {100 * sa / ts:5.1f}%  inst {100 * ins / ti:5.1f}%')
