"""Development: CUDA-event timing of the distance-field labeller (DfComputer.get_df) and the collision checker."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.collision import ColChecker, DfComputer

DMAX, HF, VF = 5.0, 0.7592, 0.4903
imgs = torch.from_numpy(np.clip(np.stack([synth.depth_frame(seed=3 + i, spherical=True) for i in range(8)]) / DMAX, 0, 1).astype(np.float32)).cuda()
rng = np.random.default_rng(0)


def best(fn, n=5):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        t.append(a.elapsed_time(b))
    return min(t)


for n in (5000, 40000):
    pts = torch.from_numpy(np.stack([rng.uniform(0.2, 5.5, n), rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n)], 1).astype(np.float32)).cuda()
    p2i = torch.from_numpy(rng.integers(0, 8, n).astype(np.int32)).cuda()
    sd = DfComputer(True, DMAX, HF, VF, 1.0, is_spherical=True)
    ud = DfComputer(False, DMAX, HF, VF, 1.0, is_spherical=True)
    cc = ColChecker(DMAX, HF, VF, 0.3, is_spherical=True)
    ts, tu, tc = best(lambda: sd.get_df(imgs, pts, p2i)), best(lambda: ud.get_df(imgs, pts, p2i)), best(lambda: cc.check_image_points(imgs, pts, p2i))
    G = sd.grid.shape[0]
    print(f'n={n}: signed {ts:.3f} ms ({n * G / ts / 1e6:.1f} G label evaluations/s, {G} voxels), unsigned {tu:.3f} ms, colcheck {tc:.4f} ms', flush=True)
