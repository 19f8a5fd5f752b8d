"""Development: SDF tensor-core kernel throughput at 1e6 points + error against the float64 oracle, and the in-step call shape."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.sdf import SdfEvaluator
from oracle import sdf as osdf

net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
ev = SdfEvaluator(net)
z = synth.latent(seed=2)
ev.set_latents(torch.from_numpy(z).cuda())
for n in (86016, 1_000_000):
    pts = torch.from_numpy(synth.sdf_query_points(n, seed=3)).cuda()
    sdf = torch.empty(n, device='cuda'); grad = torch.empty(n, 3, device='cuda')
    for _ in range(3):
        ev.eval(pts, out=(sdf, grad))
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ev.eval(pts, out=(sdf, grad))
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
    print(f'n={n}: {best:.4f} ms, {n / best / 1e3:.1f} Mpts/s, {n * 615168 / best / 1e9:.1f} TFLOP/s', flush=True)
p = synth.sdf_query_points(4096, seed=3)
d, g = ev.eval(torch.from_numpy(p).cuda())
dr, gr = osdf.value_and_grad(net, p, z)
print(f'value err {np.abs(d.cpu().numpy() - dr).max() / np.abs(dr).max():.2e}, grad err {np.abs(g.cpu().numpy() - gr).max() / np.abs(gr).max():.2e}')
