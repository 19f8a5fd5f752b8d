"""Independent cross-checks of the (unpinned) QP / RTI oracle: scipy on a tiny explicit-slack QP, KKT conditions
of the dense solution, and the documented RTI semantics."""
import numpy as np
import pytest

from conftest import make_problem


def _small_qp(rng, nv=6, mh=5, ms=4):
    A = rng.standard_normal((nv, nv))
    P = A @ A.T + 0.5 * np.eye(nv)
    q = rng.standard_normal(nv) * 3
    G = rng.standard_normal((mh + ms, nv))
    d = rng.standard_normal(mh + ms) - 0.5
    soft = np.array([False] * mh + [True] * ms)
    z = np.where(soft, rng.uniform(0.5, 3, mh + ms), 0.0)
    Z = np.where(soft, rng.choice([0.0, 2.0], mh + ms), 0.0)
    return P, q, G, d, soft, z, Z


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_dense_ipm_vs_scipy(seed):
    from scipy.optimize import minimize
    from oracle import qp as QP
    rng = np.random.default_rng(seed)
    P, q, G, d, soft, z, Z = _small_qp(rng)
    sol = QP.solve_dense(P[None], q[None], G[None], d[None], soft, z, Z)
    assert sol['status'][0] == 0
    v = sol['v'][0]
    nv, ns = len(q), int(soft.sum())
    si = np.where(soft)[0]

    def obj(x):
        vv, s = x[:nv], x[nv:]
        return 0.5 * vv @ P @ vv + q @ vv + z[si] @ s + 0.5 * (Z[si] * s) @ s

    def cons(x):
        vv, s = x[:nv], x[nv:]
        r = G @ vv - d
        r[si] += s
        return np.concatenate([r, s])

    # feasible start for SLSQP: hard rows may be infeasible at 0, so start from the IPM point perturbed
    x0 = np.concatenate([v + 1e-3, np.maximum(sol['s'][0][si], 0) + 1e-2])
    ref = minimize(obj, x0, constraints=[dict(type='ineq', fun=cons)], method='SLSQP', options=dict(ftol=1e-14, maxiter=500))
    # SLSQP sometimes reports "positive directional derivative" at this ftol although it sits on the optimum, so
    # the comparison is on the objective (the IPM point is feasible and at least as good) and, if it converged, on v
    f_ipm = obj(np.concatenate([v, sol['s'][0][si]]))
    assert f_ipm <= ref.fun + 1e-6 * max(1, abs(ref.fun))
    assert cons(np.concatenate([v, sol['s'][0][si]])).min() > -1e-7
    if ref.success:
        assert abs(f_ipm - ref.fun) <= 1e-6 * max(1, abs(ref.fun)) and np.abs(v - ref.x[:nv]).max() <= 1e-4
    # KKT of the returned point
    lam, mu, s = sol['lam'][0], sol['mu'][0], sol['s'][0]
    assert np.abs(P @ v + q - G.T @ lam).max() < 1e-6
    assert (G @ v - d + s).min() > -1e-7 and lam.min() >= 0 and np.abs(lam * (G @ v - d + s)).max() < 1e-6
    assert np.abs((z + Z * s - lam - mu)[soft]).max() < 1e-6


def test_rti_step_semantics_and_kkt():
    from oracle import rti
    from sdf_nmpc_b200 import synth
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg()
    B = 6
    prob, w = make_problem(B, cfg)
    xb, ub = rti.init_traj(prob['x0'], cfg)
    assert np.allclose(ub[:, :, 0], 9.81 / 20) and np.array_equal(xb[:, 7], prob['x0'])      # ocp.py:144-149
    xn, un, out = rti.rti_step(net, xb, ub, prob, cfg)
    assert np.all(out['status'] == 0)
    assert np.abs(xn[:, 0] - prob['x0']).max() < 1e-9                                          # dx_0 = x0 - xbar_0
    assert (un >= cfg['lbu'] - 1e-7).all() and (un <= cfg['ubu'] + 1e-7).all()                  # hard input bounds
    lin = out['lin']
    # linearised dynamics hold for the step (full step, no globalisation)
    dx, du = out['dx'], out['du']
    for k in range(cfg['N']):
        r = dx[:, k + 1] - (np.einsum('bij,bj->bi', lin['A'][:, k], dx[:, k]) + np.einsum('bij,bj->bi', lin['B'][:, k], du[:, k]) + lin['b'][:, k])
        assert np.abs(r).max() < 1e-9
    # no nonlinear rows at node 0, scaling of the stage cost by Ts and LM on the diagonal
    assert lin['k0'] == 1 and np.all(lin['hbar'][:, 0] == 0)
    y, J = __import__('oracle.model_att', fromlist=['x']).y_fun(xb[:, 3], ub[:, 3], prob['qd'][:, 3], cfg)
    H3 = cfg['dt'][3] * (np.einsum('byi,by,byj->bij', J, prob['W'][:, 3], J) + cfg['lm'] * np.eye(14))
    assert np.abs(H3 - lin['H'][:, 3]).max() < 1e-12
    # slack solution is consistent: s = max(0, violation) on soft rows
    hl = lin['hbar'][:, 1:] + np.einsum('bkrx,bkx->bkr', lin['C'][:, 1:], dx[:, 1:])
    lh, uh = rti.h_bounds(cfg, net['max_df'])
    assert np.abs(out['slack_l'] - np.maximum(lh - hl, 0)).max() < 1e-5
    assert np.abs(out['slack_u'] - np.maximum(hl - uh, 0)).max() < 1e-5
    # flag = 0: h = [0, 0, max_df] and no soft row is active (gen_model.py:61)
    prob0 = dict(prob); prob0['flag'] = np.zeros(B)
    _, _, out0 = rti.rti_step(net, xb, ub, prob0, cfg)
    assert np.allclose(out0['lin']['hbar'][:, 1:, :2], 0) and np.allclose(out0['lin']['hbar'][:, 1:, 2], net['max_df'])
    assert out0['lam_lh'].max() < 1e-6 and out0['lam_uh'].max() < 1e-6


def test_hover_start_first_input():
    """SURVEY Appendix D sanity: hover start, v_ref = (3,0,0): thrust ~ g/gamma and positive pitch."""
    from oracle import rti
    from sdf_nmpc_b200 import synth
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg()
    N = cfg['N']
    x0 = np.array([[0, 0, 2, 1, 0, 0, 0, 0, 0, 0.0]])
    yref = np.zeros((1, N, 11)); yref[0, :, 0] = 3 * 0.075 * np.arange(N); yref[0, :, 2] = 2; yref[0, :, 4] = 3
    Wrow = np.array([0, 0, 5, 10, 3, 3, 3, 50, 50, 3, 0.05])
    prob = dict(x0=x0, yref=yref, W=np.broadcast_to(Wrow, (1, N, 11)).copy(), yrefN=np.array([[4.5, 0, 2, 0]]), WN=Wrow[None, :4].copy(),
                qd=np.broadcast_to(np.array([1.0, 0, 0, 0]), (1, N + 1, 4)).copy(), flag=np.zeros(1), latent=synth.latent(seed=2)[None],
                W_p_Co=np.array([[0.18, 0, 1.975]]), W_R_Co=np.eye(3)[None])
    xb, ub = rti.init_traj(x0, cfg)
    _, _, out = rti.rti_step(net, xb, ub, prob, cfg)
    u0 = out['u0'][0]
    assert abs(u0[0] - 9.81 / 20) < 0.05 and u0[2] > 0.05 and abs(u0[1]) < 1e-6 and abs(u0[3]) < 1e-6
