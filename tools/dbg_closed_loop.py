"""Debug: per-step control/state differences GPU vs oracle in closed loop, per model."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from conftest import make_problem
from oracle import rti, models
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.nmpc import BatchedNmpc

model = sys.argv[1] if len(sys.argv) > 1 else 'att_tau'
B, steps = 8, 10
net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
ocfg = rti.default_cfg(model)
prob, w = make_problem(B, ocfg, model=model)
cfg = Config(); cfg.mpc.model = model
mpc = BatchedNmpc(cfg, net, B, 'cuda:0')
mpc.set_sdf_flag(1.0); mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo']); mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
M = models.get(model)
x = prob['x0'].copy(); xg = torch.from_numpy(x).cuda()
xb, ub = rti.init_traj(x, ocfg)
for step in range(steps):
    p = dict(prob); p['x0'] = x
    xb, ub, out = rti.rti_step(net, xb, ub, p, ocfg)
    mpc.set_x0(xg); mpc.solve()
    u0 = mpc.get_u()
    du = np.abs(u0.cpu().numpy() - out['u0']).max(1)
    x = M.rk4(x, out['u0'], ocfg['dt'][0], ocfg)
    xg = mpc.plant_step(xg.clone(), u0, ocfg['dt'][0])
    dx = np.abs(xg.cpu().numpy() - x).max(1)
    print(step, 'du', np.array2string(du, precision=1), 'dx', np.array2string(dx, precision=1), 'st', mpc.status.cpu().numpy(), out['status'], 'qpit', mpc.qp_iters.cpu().numpy(), flush=True)
