"""Development experiment (CPU, host emulation of the device QP core): IPM iteration counts for different initial points.
usage: python tools/qp_init_sweep.py "tag=-DSDFNMPC_QP_T0=0.1 -DSDFNMPC_QP_MU0=0.1" ..."""
import ctypes, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import make_problem
from test_host_emulation import _emul_step
from oracle import rti
from sdf_nmpc_b200 import synth

net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
cfg = rti.default_cfg()
B, steps = 128, 8
for spec in sys.argv[1:] or ['base=']:
    tag, flags = spec.split('=', 1)
    out = os.path.join(tempfile.mkdtemp(prefix='qpinit_'), 'lib.so')
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-DSDFNMPC_HOST_EMUL'] + flags.split() +
                   ['-x', 'c++', os.path.join(ROOT, 'tests', 'host_emul', 'emul.cpp'), '-o', out], check=True)
    lib = ctypes.CDLL(out)
    prob, w = make_problem(B, cfg, seed=1000 + 3200)
    xe, ue = rti.init_traj(prob['x0'], cfg)
    xe, ue = np.ascontiguousarray(xe), np.ascontiguousarray(ue)
    its_all, bad, u_hist = [], 0, []
    x = prob['x0'].copy()
    for step in range(steps):
        u0, st, its, *_ = _emul_step(lib, net, cfg, prob, xe, ue)
        its_all.append(its.copy()); bad += int((st != 0).sum()); u_hist.append(u0.copy())
    its_all = np.array(its_all)
    print(f'{tag:10s} mean {its_all.mean():.2f} first-step {its_all[0].mean():.2f} later {its_all[1:].mean():.2f} max {its_all.max()} bad {bad}', flush=True)
    np.save(f'/tmp/qpinit_{tag}.npy', np.array(u_hist))
