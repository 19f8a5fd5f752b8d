"""Per-kernel device time of one RTI step at small batch (development aid)."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.nmpc import BatchedNmpc
net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
for B in (1, 8, 64):
    w = synth.nmpc_instances(B)
    mpc = BatchedNmpc(Config(), net, B)
    mpc.set_sdf_flag(1.0); mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
    mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
    x = torch.from_numpy(w['x0']).cuda()
    for _ in range(5):
        mpc.set_x0(x); mpc.solve()
    torch.cuda.synchronize()
    mpc.timing_begin(20)
    for _ in range(20):
        mpc.set_x0(x); mpc.solve()
    ms, n = mpc.timing_end()
    print(B, {k: round(1e3 * v / n, 1) for k, v in ms.items()}, 'us; iters', float(mpc.qp_iters.double().mean()))
