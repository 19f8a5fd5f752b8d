"""Development aid: QP-kernel time against the number of instances (1 / 4 / 8 warps per SM resident, then several waves)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(Bn):
    import torch
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    w = synth.nmpc_instances(Bn)
    mpc = BatchedNmpc(Config(), net, Bn)
    mpc.set_sdf_flag(1.0); mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
    mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
    x = torch.from_numpy(w['x0']).cuda()

    def step():
        nonlocal x
        mpc.set_x0(x); mpc.solve()
        x = mpc.plant_step(x.clone(), mpc.get_u(), 0.075)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    mpc.timing_begin(10)
    its, mx = 0.0, 0
    for _ in range(10):
        step()
        its += float(mpc.qp_iters.double().mean().item())
        mx = max(mx, int(mpc.qp_iters.max().item()))
    ms, n = mpc.timing_end()
    return dict(B=Bn, qp_ms=ms['qp'] / n, lin_ms=ms['linearize'] / n, sdf_ms=ms['sdf'] / n, iters=its / 10, max_iters=mx,
                ok=int((mpc.status == 0).sum().item()))


if __name__ == '__main__':
    for Bn in [int(a) for a in sys.argv[1:]] or [148, 592, 1184, 2368, 4096, 8192]:
        print(json.dumps(one(Bn)), flush=True)
