"""Timing of the closed-loop step with flags.recursive_feasibility + stability on the 4096-instance workload (bench.py's
extra.terminal_ingredients) beside the plain OCP.  python tools/term_time.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from sdf_nmpc_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda', 0)
net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
w = synth.nmpc_instances(B)
x0 = torch.from_numpy(w['x0']).to(dev)
print(json.dumps(bench.terminal_variant_bench(net, w, dev, B, x0, steps=20)))
