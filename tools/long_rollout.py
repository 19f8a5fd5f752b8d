"""Robustness check: long closed-loop rollout of many instances; counts non-zero QP statuses, NaNs and iteration statistics."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.nmpc import BatchedNmpc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
model = sys.argv[3] if len(sys.argv) > 3 else 'att'
cfg = Config(); cfg.mpc.model = model
term = len(sys.argv) > 4 and sys.argv[4] == 'term'          # flags.recursive_feasibility + stability (lifted terminal rows)
if term:
    cfg.flags.recursive_feasibility = True; cfg.flags.stability = True
    cfg.mpc.braking_dist = dict(degree=4, coeffs=[float(c) for c in synth.braking_coeffs(4)])
    cfg.mpc.weights.slack_brake = [100.0, 10.0]
net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
w = synth.nmpc_instances(B, model=model)
mpc = BatchedNmpc(cfg, net, B)
mpc.set_sdf_flag(1.0); mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
yrefN, WN = w['yrefN'], w['WN']
if term:
    yrefN = np.concatenate([yrefN, np.zeros((B, 1))], 1); WN = np.concatenate([WN, np.full((B, 1), 3.0)], 1)
mpc.set_ref_arrays(w['yref'], w['W'], yrefN, WN, w['qd'])
x = torch.from_numpy(w['x0']).cuda()
bad = 0; its_max = 0; its_sum = 0.0; res_max = 0.0
for s in range(steps):
    mpc.set_x0(x); mpc.solve()
    st = mpc.status
    bad += int((st != 0).sum().item())
    its_max = max(its_max, int(mpc.qp_iters.max().item())); its_sum += float(mpc.qp_iters.double().mean().item())
    res_max = max(res_max, float(mpc.get_qp_residual()[:, 0].max().item()))
    x = mpc.plant_step(x.clone(), mpc.get_u(), 0.075)
    if not torch.isfinite(x).all():
        print('non-finite state at step', s); break
print(f'{"terminal flags on, " if term else ""}model {model} B {B} steps {steps}: non-zero statuses {bad}, mean its {its_sum / steps:.2f}, max its {its_max}, max verified KKT residual {res_max:.2e}, |x| max {float(x.abs().max()):.2f}')
