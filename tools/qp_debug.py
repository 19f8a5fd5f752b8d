"""Development aid: run the 24-instance RTI parity problem of tests/test_gpu_nmpc.py and print status / iterations
per step (use with a -DSDFNMPC_QP_DEBUG_INST=<b> variant library through SDFNMPC_LIB)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import make_problem  # noqa: E402


def main():
    import torch
    from oracle import rti
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    ocfg = rti.default_cfg('att')
    prob, w = make_problem(B, ocfg, seed=1000, flag=1.0)
    mpc = BatchedNmpc(Config(), net, B, 'cuda:0')
    mpc.set_sdf_flag(1.0)
    mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
    mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
    for step in range(3):
        mpc.set_x0(prob['x0'])
        mpc.solve()
        torch.cuda.synchronize()
        print('step', step, 'status', mpc.status.cpu().numpy().tolist(), 'iters', mpc.qp_iters.cpu().numpy().tolist())
        print('u0[3]', mpc.get_u()[min(3, B - 1)].cpu().numpy())


if __name__ == '__main__':
    main()
