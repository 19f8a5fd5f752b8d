"""Offline estimate of the QP kernel's tail: iteration counts of the device algorithm (host emulation) for the bench
workload, then list scheduling of the instances on 148 x 8 warp slots in launch order vs longest-first."""
import os, sys, heapq, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_host_emulation as T
from conftest import make_problem
from oracle import rti
from sdf_nmpc_b200 import synth
import subprocess, tempfile, ctypes

def main(B=4096, steps=4):
    out = os.path.join(tempfile.mkdtemp(prefix='sdfnmpc_emul_'), 'libhost_emul.so')
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-DSDFNMPC_HOST_EMUL', '-x', 'c++',
                    os.path.join(ROOT, 'tests', 'host_emul', 'emul.cpp'), '-o', out], check=True)
    lib = ctypes.CDLL(out)
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg()
    prob, w = make_problem(B, cfg)
    xe, ue = rti.init_traj(prob['x0'], cfg)
    xe, ue = np.ascontiguousarray(xe), np.ascontiguousarray(ue)
    hist = []
    for s in range(steps):
        t = time.time()
        u0, st, its, *_ = T._emul_step(lib, net, cfg, prob, xe, ue)
        hist.append(its.copy())
        print('step', s, 'iters mean %.2f max %d' % (its.mean(), its.max()), np.bincount(its), '%.0fs' % (time.time() - t), flush=True)
    np.save(os.path.join(ROOT, 'gpurun_out', 'qp_iters_emul.npy'), np.stack(hist))
    for s in range(1, steps):
        prev, cur = hist[s - 1], hist[s]
        for name, order in (('launch order', np.arange(B)), ('LPT by previous step', np.argsort(-prev, kind='stable')),
                            ('LPT oracle', np.argsort(-cur, kind='stable'))):
            slots = [0.0] * (148 * 8)
            heapq.heapify(slots)
            for i in order:
                t0 = heapq.heappop(slots)
                heapq.heappush(slots, t0 + cur[i] + 0.5)
            print('step', s, name, 'makespan', max(slots), 'ideal', (cur.sum() + 0.5 * B) / (148 * 8))

if __name__ == '__main__':
    main(*(int(a) for a in sys.argv[1:]))
