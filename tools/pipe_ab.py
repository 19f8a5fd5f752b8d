"""Experiment: the QP kernel's last wave is ~half empty at 4096 instances (4096 / (148 SMs x 8 warps) = 3.46 waves;
tools/qp_sched_sim.py).  Independent instances can be stepped as G groups on G streams, so that one group's SDF / linearise /
first QP wave fills the tail of another group's QP.  Times K closed-loop steps of B instances for G = 1, 2, 4 and a few batch
sizes.   python tools/pipe_ab.py [B] [K]  -> gpurun_out/pipe_ab.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.nmpc import BatchedNmpc


def run(B, G, K, net, w, priority=False):
    dev = torch.device('cuda', 0)
    n = B // G
    groups = []
    for g in range(G):
        s = torch.cuda.Stream(priority=(-1 if (priority and g % 2 == 0) else 0)) if G > 1 else torch.cuda.current_stream()
        sl = slice(g * n, (g + 1) * n)
        with torch.cuda.stream(s):
            mpc = BatchedNmpc(Config(), net, n, dev)
            mpc.set_sdf_flag(1.0)
            mpc.set_latent(w['latent'][sl], w['W_p_Bo'][sl], w['W_R_Bo'][sl])
            mpc.set_ref_arrays(w['yref'][sl], w['W'][sl], w['yrefN'][sl], w['WN'][sl], w['qd'][sl])
            x = torch.from_numpy(w['x0'][sl]).to(dev)
        groups.append((s, mpc, x))
    torch.cuda.synchronize()
    dt = 0.075

    def step():
        for s, mpc, x in groups:
            with torch.cuda.stream(s):
                mpc.set_x0(x); mpc.solve(); mpc.plant_step(x, mpc.u0, dt)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s, _, _ in groups:
        s.wait_event(e0)
    for _ in range(K):
        step()
    for s, _, _ in groups:
        ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ok = sum(int((m.status == 0).sum().item()) for _, m, _ in groups)
    u = torch.cat([m.u0 for _, m, _ in groups]).cpu()
    return dict(B=B, G=G, priority=priority, ms_per_step=ms, solves_per_s=B / ms * 1e3, status_ok=ok), u


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    out = []
    w = synth.nmpc_instances(B)
    ref = None
    for G, pr in ((1, False), (2, False), (4, False), (2, True), (8, False)):
        r, u = run(B, G, K, net, w, pr)
        if ref is None:
            ref = u
        r['max_du_vs_single'] = float((u - ref).abs().max())
        print(r, flush=True)
        out.append(r)
    for Bs in (3552, 4736, 8192):
        ws = synth.nmpc_instances(Bs)
        for G in (1, 2):
            r, _ = run(Bs, G, K, net, ws)
            print(r, flush=True)
            out.append(r)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'pipe_ab.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
