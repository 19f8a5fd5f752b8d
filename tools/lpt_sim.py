"""Would longest-first ordering of the instances shorten the QP kernel?  Measures the IPM iteration counts of consecutive closed-loop
steps and list-schedules them on the kernel's warp slots (development aid)."""
import sys, heapq, json; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from sdf_nmpc_b200 import synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.nmpc import BatchedNmpc
B = 4096
net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
w = synth.nmpc_instances(B)
mpc = BatchedNmpc(Config(), net, B)
mpc.set_sdf_flag(1.0); mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
x = torch.from_numpy(w['x0']).cuda()
its = []
for _ in range(8):
    mpc.set_x0(x); mpc.solve()
    its.append(mpc.qp_iters.cpu().numpy().copy())
    x = mpc.plant_step(x.clone(), mpc.get_u(), 0.075)
def makespan(d, order, slots=1184):
    h = [0.0] * slots
    heapq.heapify(h)
    end = 0.0
    for i in order:
        t = heapq.heappop(h) + d[i]
        end = max(end, t)
        heapq.heappush(h, t)
    return end
for k in range(1, 8):
    cur, prev = its[k], its[k - 1]
    d = 200.0 + 113.0 * cur
    ideal = d.sum() / 1184
    m0 = makespan(d, np.arange(B))
    m1 = makespan(d, np.argsort(-cur, kind='stable'))
    m2 = makespan(d, np.argsort(-prev, kind='stable'))
    print(json.dumps(dict(step=k, mean=float(cur.mean()), max=int(cur.max()), corr=float(np.corrcoef(cur, prev)[0, 1]),
                          ideal=ideal, natural=m0 / ideal, lpt_true=m1 / ideal, lpt_prev=m2 / ideal)))

# ---- can the iteration count be predicted from what is known BEFORE the solve (constraint values at the linearisation point)? ----
print('--- proxy features')
cfg = Config()
lh, uh = cfg.h_bounds(float(net['max_df']))
feats, targ = [], []
x = torch.from_numpy(w['x0']).cuda()
mpc2 = BatchedNmpc(Config(), net, B)
mpc2.set_sdf_flag(1.0); mpc2.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
mpc2.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
for k in range(8):
    mpc2.set_x0(x); mpc2.solve()
    hval = mpc2.get_diag()[0].cpu().numpy()[:, 1:]           # [B, N, 3] at the linearisation point of THIS solve
    xs, us = (t.cpu().numpy() for t in mpc2.get_matrices())
    lo = np.asarray(lh)[None, None] - hval; hi = hval - np.asarray(uh)[None, None]
    viol = np.maximum(np.maximum(lo, hi), 0.0)
    near = (np.maximum(lo, hi) > -0.1)
    f = np.stack([viol[..., 0].sum(1), viol[..., 1].sum(1), viol[..., 2].sum(1), near[..., 0].sum(1), near[..., 1].sum(1), near[..., 2].sum(1),
                  (np.abs(us) > 0.95).sum((1, 2)), np.ones(B)], 1)
    feats.append(f); targ.append(mpc2.qp_iters.cpu().numpy().astype(float))
    x = mpc2.plant_step(x.clone(), mpc2.get_u(), 0.075)
for k in range(1, 8):
    coef, *_ = np.linalg.lstsq(feats[k - 1], targ[k - 1], rcond=None)
    pred = feats[k] @ coef
    d = 200.0 + 113.0 * targ[k]
    ideal = d.sum() / 1184
    print(json.dumps(dict(step=k, corr_pred=float(np.corrcoef(pred, targ[k])[0, 1]), natural=makespan(d, np.arange(B)) / ideal,
                          lpt_pred=makespan(d, np.argsort(-pred, kind='stable')) / ideal,
                          lpt_true=makespan(d, np.argsort(-targ[k], kind='stable')) / ideal)))
