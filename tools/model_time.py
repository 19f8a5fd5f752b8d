"""Closed-loop step time of every model (att / att_tau / acc / props) on the 4096-instance workload, per kernel.
python tools/model_time.py [B] [models...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sdf_nmpc_b200 import synth  # noqa: E402
from sdf_nmpc_b200.config import Config  # noqa: E402
from sdf_nmpc_b200.nmpc import BatchedNmpc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
models = sys.argv[2:] or ['att', 'att_tau', 'acc', 'props']
net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
for model in models:
    cfg = Config()
    cfg.mpc.model = model
    w = synth.nmpc_instances(B, model=model)
    mpc = BatchedNmpc(cfg, net, B)
    mpc.set_sdf_flag(1.0)
    mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
    mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
    x = torch.from_numpy(w['x0']).cuda()
    dt = float(mpc.dt[0])

    def step():
        mpc.set_x0(x)
        mpc.solve()
        mpc.plant_step(x, mpc.u0, dt)

    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    mpc.timing_begin(3)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    kms, n = mpc.timing_end()
    print(json.dumps(dict(model=model, ms_per_step=ms, solves_per_s=B / ms * 1e3, qp_iters_mean=float(mpc.qp_iters.double().mean().item()),
                          status_ok=int((mpc.status == 0).sum().item()), kernels_ms={k: round(v / max(n, 1), 4) for k, v in kms.items()})), flush=True)
    del mpc
    torch.cuda.empty_cache()
