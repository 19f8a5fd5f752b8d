"""Development timing probe (not the bench contract): SDF sweep, RTI step, encoder.  Writes gpurun_out/quick_perf.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import synth  # noqa: E402
from sdf_nmpc_b200.config import Config  # noqa: E402
from sdf_nmpc_b200.nmpc import BatchedNmpc  # noqa: E402
from sdf_nmpc_b200.sdf import SdfEvaluator  # noqa: E402
from sdf_nmpc_b200.vae import VaeWrapper  # noqa: E402


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    out = {}
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    ev = SdfEvaluator(net)
    ev.set_latents(torch.from_numpy(synth.latent(seed=2)).cuda())
    F_PT = 615168
    for n in (10_000, 100_000, 1_000_000, 10_000_000):
        pts = torch.from_numpy(synth.sdf_query_points(n, seed=3)).cuda()
        sdf = torch.empty(n, device='cuda'); grad = torch.empty(n, 3, device='cuda')
        ms = timeit(lambda: ev.eval(pts, out=(sdf, grad)), it=5 if n >= 1_000_000 else 20)
        out[f'sdf_n{n}'] = dict(ms=ms, mpts_s=n / ms / 1e3, tflops=n * F_PT / ms / 1e9)
        print('sdf', n, out[f'sdf_n{n}'], flush=True)
    for B in (1, 256, 4096):
        cfg = Config()
        w = synth.nmpc_instances(B)
        mpc = BatchedNmpc(cfg, net, B)
        mpc.set_sdf_flag(1.0)
        mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
        mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
        x0 = torch.from_numpy(w['x0']).cuda()
        mpc.set_x0(x0)
        ms = timeit(lambda: mpc.solve(), warm=3, it=10)
        it = mpc.qp_iters.cpu().numpy()
        st = mpc.status.cpu().numpy()
        out[f'rti_B{B}'] = dict(ms=ms, solves_s=B / ms * 1e3, qp_iters_mean=float(it.mean()), qp_iters_max=int(it.max()),
                                status_ok=int((st == 0).sum()))
        print('rti', B, out[f'rti_B{B}'], flush=True)
        del mpc
    enc = synth.make_encoder_net(seed=5)
    for B in (1, 64):
        cfg = Config()
        vae = VaeWrapper(cfg, enc, max_batch=B)
        frames = np.stack([synth.depth_frame(seed=i) for i in range(min(B, 4))])
        frames = np.concatenate([frames] * (B // len(frames) + 1))[:B]
        vae.set_img(frames)
        ms = timeit(lambda: vae.encode(to_numpy=False), warm=2, it=5)
        out[f'enc_B{B}'] = dict(ms=ms, frames_s=B / ms * 1e3, tflops=B * 4.34e9 / ms / 1e9)
        print('enc', B, out[f'enc_B{B}'], flush=True)
        del vae
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/quick_perf.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
