"""Oracle: one SQP(-RTI) iteration of the sdf-nmpc OCP, batched numpy float64.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  **Parity unpinned** for the
acados semantics (third-party, source not in the reference tree): the choices
below restate SURVEY.md section 8a-5 [from memory of acados 0.3.1] and define what
"the reference QP" means for this repository:

  1. stage cost 1/2||y-yref||^2_W, its Gauss-Newton Hessian/gradient and the slack
     penalties are scaled by the stage length Ts_k; terminal scaling 1
     (``ocp.py:54-61,85-92``; acados cost "scaling").
  2. Levenberg-Marquardt: + Ts_k*lm*I on the (nu+nx) stage block, + lm*I at N
     (``ocp.py:119-120``).
  3. QP variables per stage [du; dx]; ``Ocp.solve`` writes the measured state into node 0 of the iterate
     (``solver.set(0, 'x', x0)``, ``ocp.py:161``) BEFORE the RTI step, so node 0 is linearised at x0 and
     dx_0 = x0 - xbar_0 = 0 (``ocp.py:160-161,169``).
  4. nonlinear rows h (hfov, vfov, sdf - order ``gen_model.py:42,68``) at nodes
     1..N-1 and N; none at node 0 (switch ``h_at_node0``).
  5. soft rows: lh <= h + s_l, h - s_u <= uh, s >= 0, penalty z's + 1/2 s'Zs
     (``base_model.py:142-168``; ``default.yaml:49-50``).
  6. ERK4, one step per interval, exact sensitivities of the discrete map.
  7. full step, no globalisation; status = QP status.
The QP is condensed to the N*nu inputs (``FULL_CONDENSING_HPIPM``, ``ocp.py:113``)
and solved by ``oracle.qp.solve_dense``.
"""
import numpy as np

from . import model_att as M
from . import models as MODELS
from . import qp as QP
from . import sdf as SDF


def default_cfg(model='att'):
    """`model` in {'att', 'att_tau', 'acc', 'props'} (gen_model.py:15-20); bounds and hover input come from the model."""
    c = dict(M.DEFAULT)
    mdl = MODELS.get(model)
    c.update(model=model, ax=4.0, ay=4.0, az=3.0, tau_roll=0.12, tau_pitch=0.12)
    if model == 'props':        # robot.alloc / inertia / limits.wp of config/default.yaml:79-87,108
        motors = [[0.09, -0.09, -0.005, 0, 0, -1], [-0.09, 0.09, -0.005, 0, 0, -1], [0.09, 0.09, -0.005, 0, 0, 1], [-0.09, -0.09, -0.005, 0, 0, 1]]
        Gf, Gt = MODELS.props_allocation(motors, 0.02246, 0.00020673)
        c.update(cf=0.02246, ct=0.00020673, motors=motors, Gf=Gf, Gt=Gt, inertia=np.array([0.017, 0.018, 0.028]), wp_max=25.0)
    c.update(N=20, T=1.5, lm=10.0, lbu=mdl.lbu.copy(), ubu=mdl.ubu.copy(),
             slack_df=(200.0, 50.0), slack_fov=(20.0, 0.0), size_xy=0.22, bound_margin=0.15,
             h_at_node0=False, qp_tol=1e-8, qp_iter_max=100)
    c['dt'] = np.full(c['N'], c['T'] / c['N'])
    return c


def h_bounds(cfg, max_df):
    """(lh, uh) for [hfov, vfov, sdf]  (cost_const_helpers.py:67-75, gen_model.py:35)."""
    hl, vl = cfg['hfov'] * cfg['fov_ratio'], cfg['vfov'] * cfg['fov_ratio']
    return (np.array([-hl, -vl, cfg['size_xy'] + cfg['bound_margin']]), np.array([hl, vl, max_df + 0.2]))


def slack_weights(cfg):
    """(z [3], Z [3]) in h order (ocp.py:85-92)."""
    return (np.array([cfg['slack_fov'][0], cfg['slack_fov'][0], cfg['slack_df'][0]]),
            np.array([cfg['slack_fov'][1], cfg['slack_fov'][1], cfg['slack_df'][1]]))


def poly3(v, deg, coeffs):
    """3-variate polynomial of ``utils/math.py:294-321`` (terms ordered by total degree, then a, then b; c = td - a - b):
    value [B] and gradient [B,3]."""
    val = np.zeros(v.shape[0]); grad = np.zeros((v.shape[0], 3))
    i = 0
    for td in range(deg + 1):
        for a in range(td + 1):
            for b in range(td + 1 - a):
                c = td - a - b
                val += coeffs[i] * v[:, 0] ** a * v[:, 1] ** b * v[:, 2] ** c
                if a: grad[:, 0] += coeffs[i] * a * v[:, 0] ** (a - 1) * v[:, 1] ** b * v[:, 2] ** c
                if b: grad[:, 1] += coeffs[i] * b * v[:, 0] ** a * v[:, 1] ** (b - 1) * v[:, 2] ** c
                if c: grad[:, 2] += coeffs[i] * c * v[:, 0] ** a * v[:, 1] ** b * v[:, 2] ** (c - 1)
                i += 1
    assert i == len(coeffs)
    return val, grad


def term_rows(net, xN, prob, cfg):
    """Terminal rows of ``flags.recursive_feasibility`` / ``flags.stability`` (``gen_model.py:73-118``; 'att' state layout
    [p, q, v]) at the terminal node: values h [B,m], gradients wrt x [B,m,nx], bounds, soft mask and slack weights.  Row order:
    rec_feas = sdf_flag(p) - flag * poly(v) in [size.xy, max_df] (soft with ``weights.slack_brake``, hard when that is None);
    hfov / vfov (hard) at the braking end point p + flag * poly(v) v / sqrt(v.v + 1e-4) (``gen_model.py:105-111``);
    with stability the velocity box |v_i| <= limits.v_i (hard, ``cost_const_helpers.py:93-102``)."""
    t = cfg['term']
    B, nx = xN.shape
    flag = prob['flag']
    v = xN[:, 7:10]
    pv, gpv = poly3(v, t['degree'], t['coeffs'])
    hh, Jp = eval_h(net, xN, prob, cfg)                  # sdf_flag value and d/dp (row 2)
    rows_h, rows_C, lo, hi, soft, z, Z = [], [], [], [], [], [], []
    C = np.zeros((B, nx)); C[:, 0:3] = Jp[:, 2]; C[:, 7:10] = -flag[:, None] * gpv
    rows_h.append(hh[:, 2] - flag * pv); rows_C.append(C)
    lo.append(cfg['size_xy']); hi.append(net['max_df'])
    sb = t.get('slack_brake')
    soft.append(sb is not None); z.append(sb[0] if sb else 0.0); Z.append(sb[1] if sb else 0.0)
    # braking end point and its derivative
    n = np.sqrt((v * v).sum(1) + 1e-4)
    dirv = v / n[:, None]
    d = flag * pv
    xE = xN.copy(); xE[:, 0:3] = xN[:, 0:3] + d[:, None] * dirv
    dE_dv = d[:, None, None] * (np.eye(3)[None] / n[:, None, None] - np.einsum('bi,bj->bij', v, v) / (n ** 3)[:, None, None]) \
        + np.einsum('bi,bj->bij', dirv, flag[:, None] * gpv)
    hf, JE = M.h_fov(xE, prob['W_p_Co'], prob['W_R_Co'], flag, cfg)        # d/dp_E [B,2,3]
    hl, vl = cfg['hfov'] * cfg['fov_ratio'], cfg['vfov'] * cfg['fov_ratio']
    for r, lim in ((0, hl), (1, vl)):
        if r == 1 and not t.get('vfov', True):
            continue
        C = np.zeros((B, nx)); C[:, 0:3] = JE[:, r]; C[:, 7:10] = np.einsum('bp,bpv->bv', JE[:, r], dE_dv)
        rows_h.append(hf[:, r]); rows_C.append(C); lo.append(-lim); hi.append(lim); soft.append(False); z.append(0.0); Z.append(0.0)
    if t.get('stability', False):
        for i in range(3):
            C = np.zeros((B, nx)); C[:, 7 + i] = 1.0
            rows_h.append(v[:, i].copy()); rows_C.append(C); lo.append(-t['vlim'][i]); hi.append(t['vlim'][i])
            soft.append(False); z.append(0.0); Z.append(0.0)
    return (np.stack(rows_h, 1), np.stack(rows_C, 1), np.array(lo), np.array(hi), np.array(soft, bool), np.array(z), np.array(Z))


def init_traj(x0, cfg):
    """ocp.py:144-149: x_k = x0, u_k = u_hover."""
    B, N = x0.shape[0], cfg['N']
    xb = np.repeat(x0[:, None, :], N + 1, 1).copy()
    ub = np.broadcast_to(MODELS.get(cfg.get('model', 'att')).u_hover(cfg), (B, N, 4)).copy()
    return xb, ub


def eval_h(net, xk, prob, cfg):
    """h [B,3] and dh/dp [B,3,3] at one node."""
    flag = prob['flag']
    hf, Jf = M.h_fov(xk, prob['W_p_Co'], prob['W_R_Co'], flag, cfg)
    c0 = M.sdf_query(xk, prob['W_p_Co'], prob['W_R_Co'])
    # the network runs in float32 behind a double interface (L4CasADi)
    d, gs = SDF.value_and_grad(net, c0.astype(np.float32), prob['latent'], dtype=prob.get('sdf_dtype', np.float32))
    d, gs = d.astype(np.float64), gs.astype(np.float64)
    h3 = flag * d + (1.0 - flag) * net['max_df']
    J3 = flag[:, None] * np.einsum('bc,bpc->bp', gs, prob['W_R_Co'])
    return np.concatenate([hf, h3[:, None]], 1), np.concatenate([Jf, J3[:, None, :]], 1)


def linearize(net, xb, ub, prob, cfg):
    """All stage quantities of the OCP-QP at (xb, ub)."""
    B, N = xb.shape[0], cfg['N']
    dt = cfg['dt']
    mdl = MODELS.get(cfg.get('model', 'att'))
    nyN, nx = mdl.nyN, mdl.nx
    nw = nx + 4
    A = np.zeros((B, N, nx, nx)); Bm = np.zeros((B, N, nx, 4)); b = np.zeros((B, N, nx))
    H = np.zeros((B, N, nw, nw)); g = np.zeros((B, N, nw))
    C = np.zeros((B, N + 1, 3, nx)); hbar = np.zeros((B, N + 1, 3))
    for k in range(N):
        xn, A[:, k], Bm[:, k] = mdl.rk4_sens(xb[:, k], ub[:, k], dt[k], cfg)
        b[:, k] = xn - xb[:, k + 1]
        y, J = mdl.y_fun(xb[:, k], ub[:, k], prob['qd'][:, k], cfg)
        Wk = prob['W'][:, k]
        H[:, k] = dt[k] * (np.einsum('byi,by,byj->bij', J, Wk, J) + cfg['lm'] * np.eye(nw)[None])
        g[:, k] = dt[k] * np.einsum('byi,by->bi', J, Wk * (y - prob['yref'][:, k]))
        if cfg.get('sdf_cost_w', 0.0) > 0:      # flags.sdf_cost: stage cost row (1 - 0.5 sdf)^4, weight 20 (gen_model.py:65-66)
            hk, Jp = eval_h(net, xb[:, k], prob, cfg)
            qq = 1.0 - 0.5 * hk[:, 2]
            rr, Jr = qq ** 4, -2.0 * (qq ** 3)[:, None] * Jp[:, 2, :]
            H[:, k, 4:7, 4:7] += dt[k] * cfg['sdf_cost_w'] * np.einsum('bi,bj->bij', Jr, Jr)
            g[:, k, 4:7] += dt[k] * cfg['sdf_cost_w'] * rr[:, None] * Jr
    yN, JN = mdl.y_fun(xb[:, N], np.zeros((B, 4)), prob['qd'][:, N], cfg)
    yN, JN = yN[:, :nyN], JN[:, :nyN, 4:]
    term = cfg.get('term')
    if term and term.get('stability', False):
        # quad_rollpitchyawrate.py:52-55: yN = flag * [p, q_e3]; gen_model.py:139-147: extra terminal output flag * v.v
        f = prob['flag']
        v = xb[:, N, 7:10]
        Jv = np.zeros((B, 1, nx)); Jv[:, 0, 7:10] = 2.0 * f[:, None] * v
        yN = np.concatenate([f[:, None] * yN, (f * (v * v).sum(1))[:, None]], 1)
        JN = np.concatenate([f[:, None, None] * JN, Jv], 1)
    HN = np.einsum('byi,by,byj->bij', JN, prob['WN'], JN) + cfg['lm'] * np.eye(nx)[None]
    gN = np.einsum('byi,by->bi', JN, prob['WN'] * (yN - prob['yrefN']))
    k0 = 0 if cfg['h_at_node0'] else 1
    for k in range(k0, N + 1):
        hbar[:, k], Jp = eval_h(net, xb[:, k], prob, cfg)
        C[:, k, :, 0:3] = Jp
    out = dict(A=A, B=Bm, b=b, H=H, g=g, HN=HN, gN=gN, C=C, hbar=hbar, k0=k0)
    if term:
        out['term'] = term_rows(net, xb[:, N], prob, cfg)
    return out


def condense(lin, xb, ub, prob, cfg, max_df):
    """Dense QP in U = (du_0..du_{N-1}) with the row list used for active-set comparison:
    rows = [u lower (N*4), u upper (N*4), h lower (nodes k0..N, 3 each), h upper (same)]."""
    B, N = xb.shape[0], cfg['N']
    nv = 4 * N
    nx = xb.shape[2]
    Gam = np.zeros((B, N + 1, nx, nv)); gam = np.zeros((B, N + 1, nx))
    gam[:, 0] = prob['x0'] - xb[:, 0]
    for k in range(N):
        Gam[:, k + 1] = lin['A'][:, k] @ Gam[:, k]
        Gam[:, k + 1, :, 4 * k:4 * k + 4] += lin['B'][:, k]
        gam[:, k + 1] = np.einsum('bij,bj->bi', lin['A'][:, k], gam[:, k]) + lin['b'][:, k]
    P = np.zeros((B, nv, nv)); q = np.zeros((B, nv))
    for k in range(N):
        T = np.zeros((B, nx + 4, nv)); T[:, np.arange(4), 4 * k + np.arange(4)] = 1.0; T[:, 4:] = Gam[:, k]
        t = np.concatenate([np.zeros((B, 4)), gam[:, k]], 1)
        HT = lin['H'][:, k] @ T
        P += np.swapaxes(T, 1, 2) @ HT
        q += np.einsum('bin,bi->bn', T, np.einsum('bij,bj->bi', lin['H'][:, k], t) + lin['g'][:, k])
    TN = Gam[:, N]
    P += np.swapaxes(TN, 1, 2) @ lin['HN'] @ TN
    q += np.einsum('bin,bi->bn', TN, np.einsum('bij,bj->bi', lin['HN'], gam[:, N]) + lin['gN'])
    P = 0.5 * (P + np.swapaxes(P, 1, 2))
    lh, uh = h_bounds(cfg, max_df)
    zl, Zl = slack_weights(cfg)
    k0 = lin['k0']
    nk = N + 1 - k0
    I = np.broadcast_to(np.eye(nv), (B, nv, nv))
    ubar = ub.reshape(B, nv)
    rows_G = [I, -I]
    rows_d = [np.tile(cfg['lbu'], N)[None] - ubar, -(np.tile(cfg['ubu'], N)[None] - ubar)]
    CG = np.einsum('bkrx,bkxn->bkrn', lin['C'][:, k0:], Gam[:, k0:]).reshape(B, nk * 3, nv)
    Cg = np.einsum('bkrx,bkx->bkr', lin['C'][:, k0:], gam[:, k0:]).reshape(B, nk * 3)
    hb = lin['hbar'][:, k0:].reshape(B, nk * 3)
    # rows the mission flags switch off (gen_model.py:41,67-70) are REMOVED here: cfg['h_off'] = bool[3] in h order
    keep = np.tile(~np.asarray(cfg.get('h_off', [False] * 3), bool), nk)
    if cfg.get('term'):
        keep[-1] = False        # gen_model.py:69-70: no terminal sdf row next to the recursive-feasibility row
    CG, Cg, hb = CG[:, keep], Cg[:, keep], hb[:, keep]
    nh = int(keep.sum())
    rows_G += [CG, -CG]
    rows_d += [np.tile(lh, nk)[keep][None] - hb - Cg, -(np.tile(uh, nk)[keep][None] - hb - Cg)]
    soft = np.concatenate([np.zeros(2 * nv, bool), np.ones(2 * nh, bool)])
    # slack penalties of stage rows (nodes k0..N-1) are scaled by Ts_k, terminal rows by 1
    sc = np.concatenate([np.repeat(cfg['dt'][k0:N], 3), np.ones(3)])
    zz = np.concatenate([np.zeros(2 * nv), (np.tile(zl, nk) * sc)[keep], (np.tile(zl, nk) * sc)[keep]])
    ZZ = np.concatenate([np.zeros(2 * nv), (np.tile(Zl, nk) * sc)[keep], (np.tile(Zl, nk) * sc)[keep]])
    nt = 0
    if cfg.get('term'):         # rows of term_rows: full-width gradients at node N, appended as [lower; upper]
        hT, CT, loT, hiT, sfT, zT, ZT = lin['term']
        nt = hT.shape[1]
        TG = np.einsum('brx,bxn->brn', CT, Gam[:, N]); Tg = np.einsum('brx,bx->br', CT, gam[:, N])
        rows_G += [TG, -TG]
        rows_d += [loT[None] - hT - Tg, -(hiT[None] - hT - Tg)]
        soft = np.concatenate([soft, sfT, sfT]); zz = np.concatenate([zz, zT, zT]); ZZ = np.concatenate([ZZ, ZT, ZT])
    G = np.concatenate(rows_G, 1); d = np.concatenate(rows_d, 1)
    return dict(P=P, q=q, G=G, d=d, soft=soft, z=zz, Z=ZZ, Gam=Gam, gam=gam, nk=nk, keep=keep, nt=nt)


def rti_step(net, xb, ub, prob, cfg):
    """One SQP-RTI iteration (preparation + feedback).  Returns new (xb, ub) and diagnostics."""
    B, N = xb.shape[0], cfg['N']
    xb = xb.copy()
    xb[:, 0] = prob['x0']                      # ocp.py:161: solver.set(0, 'x', x0) precedes the solve
    lin = linearize(net, xb, ub, prob, cfg)
    cq = condense(lin, xb, ub, prob, cfg, net['max_df'])
    sol = QP.solve_dense(cq['P'], cq['q'], cq['G'], cq['d'], cq['soft'], cq['z'], cq['Z'],
                         tol=cfg['qp_tol'], max_iter=cfg['qp_iter_max'])
    U = sol['v']
    dx = np.einsum('bkxn,bn->bkx', cq['Gam'], U) + cq['gam']
    xn = xb + dx
    un = ub + U.reshape(B, N, 4)
    nv, nk = 4 * N, cq['nk']
    lam = sol['lam']
    keep = cq['keep']
    nh = int(keep.sum())

    def full(v):            # removed rows report zero multiplier / slack
        o = np.zeros((B, nk * 3))
        o[:, keep] = v
        return o.reshape(B, nk, 3)

    out = dict(u0=un[:, 0].copy(), status=sol['status'], qp_iters=sol['iters'],
               lam_lbu=lam[:, :nv].reshape(B, N, 4), lam_ubu=lam[:, nv:2 * nv].reshape(B, N, 4),
               lam_lh=full(lam[:, 2 * nv:2 * nv + nh]), lam_uh=full(lam[:, 2 * nv + nh:2 * nv + 2 * nh]),
               slack_l=full(sol['s'][:, 2 * nv:2 * nv + nh]), slack_u=full(sol['s'][:, 2 * nv + nh:2 * nv + 2 * nh]),
               du=U.reshape(B, N, 4), dx=dx, lin=lin, qp=cq)
    nt = cq['nt']
    if nt:
        o = 2 * nv + 2 * nh
        out.update(lam_term_l=lam[:, o:o + nt], lam_term_u=lam[:, o + nt:o + 2 * nt],
                   slack_term_l=sol['s'][:, o:o + nt], slack_term_u=sol['s'][:, o + nt:o + 2 * nt])
    return xn, un, out


def active_set(out, thr=1e-6):
    """Boolean masks of active rows: multiplier > thr*max(1, max multiplier)."""
    allm = np.concatenate([out[k].reshape(out[k].shape[0], -1) for k in ('lam_lbu', 'lam_ubu', 'lam_lh', 'lam_uh')], 1)
    ref = thr * np.maximum(1.0, allm.max(1))
    return {k: out[k] > ref.reshape((-1,) + (1,) * (out[k].ndim - 1)) for k in ('lam_lbu', 'lam_ubu', 'lam_lh', 'lam_uh')}


def closed_loop(net, prob, cfg, steps, plant_dt=None, move_ref=True):
    """RK4 plant = same model; SQP-RTI once per control step, warm start = previous solution (shift 0)."""
    x = prob['x0'].copy()
    xb, ub = init_traj(x, cfg)
    plant_dt = cfg['dt'][0] if plant_dt is None else plant_dt
    hist = []
    p = dict(prob)
    for _ in range(steps):
        p['x0'] = x
        xb, ub, out = rti_step(net, xb, ub, p, cfg)
        hist.append(dict(x=x.copy(), u0=out['u0'].copy(), status=out['status'].copy()))
        x = MODELS.get(cfg.get('model', 'att')).rk4(x, out['u0'], plant_dt, cfg)
    return x, xb, ub, hist


def sqp_solve(net, xb, ub, prob, cfg, tol=1e-6, max_iter=100):
    """Full SQP (config 5): repeat the RTI iteration until the full step is below `tol` (inf-norm over [du; dx]) or
    `max_iter` iterations; converged instances are frozen.  Returns (xb, ub, iters [B], converged [B], last out).
    Termination on the Newton-type step is this repository's definition (acados tests four KKT residuals; with exact
    QP solves and full steps the step vanishes exactly when they do)."""
    B = xb.shape[0]
    xb, ub = xb.copy(), ub.copy()
    iters = np.zeros(B, int)
    active = np.ones(B, bool)
    out = None
    step_hist = []
    for _ in range(max_iter):
        if not active.any():
            break
        xn, un, out = rti_step(net, xb, ub, prob, cfg)
        step = np.maximum(np.abs(out['dx']).reshape(B, -1).max(1), np.abs(out['du']).reshape(B, -1).max(1))
        xb[active], ub[active] = xn[active], un[active]
        iters[active] += 1
        step_hist.append(np.where(active, step, np.nan))
        active &= step > tol
    return xb, ub, iters, ~active, dict(last=out, steps=np.array(step_hist))
