import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.sdf import SdfEvaluator
os.environ.setdefault('SDFNMPC_SDF_IMPL', 'tc')
net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
ev = SdfEvaluator(net)
ev.set_latents(torch.from_numpy(synth.latent(seed=2)).cuda())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
pts = torch.from_numpy(synth.sdf_query_points(n, seed=3)).cuda()
for _ in range(3):
    d, g = ev.eval(pts)
torch.cuda.synchronize()
print('ok', float(d.sum()), float(g.sum()))
