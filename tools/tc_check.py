"""Development check of the tcgen05 SDF kernel against the CUDA-core kernel and the oracle."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.sdf import SdfEvaluator
from oracle import sdf as osdf

def run(impl, net, pts, lat, idx):
    os.environ['SDFNMPC_SDF_IMPL'] = impl
    ev = SdfEvaluator(net)
    ev.set_latents(torch.from_numpy(lat).cuda())
    p = torch.from_numpy(pts).cuda(); i = torch.from_numpy(idx).cuda()
    d, g = ev.eval(p, pt2inst=i)
    torch.cuda.synchronize()
    t = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ev.eval(p, pt2inst=i); e1.record(); torch.cuda.synchronize(); t.append(e0.elapsed_time(e1))
    return d.cpu().numpy(), g.cpu().numpy(), min(t)

rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
for arch, kw in (('B', dict(seed=0, **synth.ARCH_B, calibrated=True)), ('A', dict(seed=1, **synth.ARCH_A, bias_scale=0.1))):
    net = synth.make_sdf_net(**kw)
    for n in (128, 1000, 100_000, 1_000_000):
        pts = synth.sdf_query_points(n, seed=3)
        lat = np.random.default_rng(11).standard_normal((4, 128)).astype(np.float32)
        idx = (np.arange(n) * 4 // n).astype(np.int32)
        d1, g1, t1 = run('ffma', net, pts, lat, idx)
        d2, g2, t2 = run('tc', net, pts, lat, idx)
        m = min(n, 4096)
        dr, gr = osdf.value_and_grad(net, pts[:m], lat[idx[:m]])
        print(f'arch {arch} n {n}: ffma {t1:.3f} ms  tc {t2:.3f} ms ({n / t2 / 1e3:.1f} Mpts/s, {n * 615168 / t2 / 1e9:.1f} TFLOP/s) | '
              f'tc vs ffma d {rel(d2, d1):.2e} g {rel(g2, g1):.2e} | tc vs oracle d {rel(d2[:m], dr):.2e} g {rel(g2[:m], gr):.2e} | '
              f'ffma vs oracle d {rel(d1[:m], dr):.2e} g {rel(g1[:m], gr):.2e}', flush=True)
