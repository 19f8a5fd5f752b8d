"""Development aid: per-phase cycles of the QP kernel (a -DSDFNMPC_QP_TRACE build through SDFNMPC_LIB), instance 0,
IPM iteration 2, for a given number of instances (= resident warps up to 1184)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {0: 'elim0', 1: 'factor', 2: 'rollout0', 3: 'steps0', 8: 'elim1', 9: 'corr sweep', 10: 'rollout1', 11: 'steps1',
         16: 'prologue operands', 17: 'hessian build', 18: 'initial point', 19: 'verify pass 2', 20: 'outputs', 21: 'verify multipliers', 22: 'verify pass 1'}
for B in sys.argv[1:] or ['148', '1184', '4096']:
    f = os.path.join(ROOT, 'gpurun_out', f'qpt_{B}.txt')
    env = dict(os.environ, SDFNMPC_QP_TRACE_FILE=f, SDFNMPC_LIB=os.path.join(ROOT, 'sdf_nmpc_b200/lib/variants/lib_trace.so'))
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools/qp_scaling.py'), B], env=env, capture_output=True)
    v = [int(x) for x in open(f).read().split()]
    it = sum(v[i] for i in (0, 1, 2, 3, 8, 9, 10, 11))
    print(f'B={B}: iteration {it} cycles;', ', '.join(f'{NAMES[i]} {v[i]}' for i in sorted(NAMES)))
