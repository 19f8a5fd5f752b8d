"""Oracle: dense convex QP solved the way the reference does it - full condensing
followed by a dense primal-dual interior-point method (HPIPM's role under
``qp_solver='FULL_CONDENSING_HPIPM'``, ``sdf_nmpc/ocp.py:113-116``).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  **Parity unpinned**: HPIPM is
a third-party dependency (bundled with acados 0.3.1) that is neither in the
reference tree nor installable here; this is a textbook Mehrotra
predictor-corrector restated from the published algorithm (Mehrotra 1992;
Frison & Diehl, "HPIPM", 2020, section 3) and only the *optimum* is
comparable, not HPIPM's iterates.  Cross-checked against scipy in the tests.

Problem (batched over the leading axis):
    min_v  1/2 v'Pv + q'v + sum_{i soft} ( z_i s_i + 1/2 Z_i s_i^2 )
    s.t.   G_i v - d_i        >= 0      (hard rows)
           G_i v - d_i + s_i  >= 0,  s_i >= 0   (soft rows)
Slacks are eliminated from the Newton system row by row (they couple to exactly
one row each), which is also what HPIPM does.
"""
import numpy as np


def solve_dense(P, q, G, d, soft, z, Z, tol=1e-8, max_iter=100, mu0=1.0, verbose=False, tol_res=None):
    """Returns dict(v, lam, s, mu, iters [B], status [B] (0 ok, 2 max-iter, 4 NaN/fact. failure))."""
    B, nv = q.shape
    tol_res = 10.0 * tol if tol_res is None else tol_res   # residual noise floor is ~Theta*eps, see DESIGN.md
    m = d.shape[1]
    soft = np.asarray(soft, bool)
    sf = soft.astype(float)[None, :]
    z = np.broadcast_to(np.asarray(z, float), (B, m)) * sf
    Z = np.broadcast_to(np.asarray(Z, float), (B, m)) * sf
    v = np.zeros((B, nv))
    s = np.where(soft[None], 0.5, 0.0) * np.ones((B, m))
    r = np.einsum('bmn,bn->bm', G, v) - d + s
    t = np.maximum(r, 0.5)
    lam = np.where(soft[None], 0.5 * (z + Z * s), mu0 / t)
    mu = np.where(soft[None], 0.5 * (z + Z * s), 0.0)
    mu = np.maximum(mu, 1e-3 * sf)
    lam = np.maximum(lam, 1e-3)
    ncomp = m + int(soft.sum())
    iters = np.zeros(B, int)
    status = np.full(B, 2, int)
    done = np.zeros(B, bool)
    safe = np.zeros(B, bool)
    stuck = np.zeros(B, int)            # iterations spent at complementarity <= tol without meeting the residual test
    hist = []
    s_safe = lambda: np.where(soft[None], s, 1.0)
    GT = np.swapaxes(G, 1, 2)
    for it in range(max_iter + 1):
        Gv = np.einsum('bmn,bn->bm', G, v)
        r_p = t - (Gv - d + s)
        r_s = (z + Z * s - lam - mu) * sf
        r_v = np.einsum('bij,bj->bi', P, v) + q - np.einsum('bnm,bm->bn', GT, lam)
        comp = lam * t + mu * s
        mubar = comp.sum(1) / ncomp
        res = np.maximum.reduce([np.abs(r_p).max(1), np.abs(r_s).max(1), np.abs(r_v).max(1)])
        cmax = np.maximum((lam * t).max(1), (mu * s).max(1))
        bad = ~np.isfinite(res) | ~np.isfinite(mubar)
        # Noise floor: with a bound active at t ~ 1e-11 and a multiplier of O(10) the condensed Newton matrix carries Theta ~ 1e12
        # and the recomputed stationarity residual jitters between 1e-8 and 1e-5 from one iteration to the next (seen with the
        # stiff 'props' model).  Ten iterations at converged complementarity with the residual within 1e3 x tol_res count as
        # converged - the structured solvers track the residual by its exact (1 - alpha) contraction and do not see this jitter.
        stuck = np.where((cmax <= tol) & (res <= 1e3 * tol_res), stuck + 1, 0)
        newly = (~done) & (((res <= tol_res) & (cmax <= tol)) | (stuck >= 10) | bad)
        status[newly & ~bad] = 0
        status[newly & bad] = 4
        iters[newly] = it
        done |= newly
        if verbose:
            print(it, res.max(), cmax.max(), np.abs(r_p).max(), np.abs(r_s).max(), np.abs(r_v).max(), t.min(), lam.max())
        if done.all() or it == max_iter:
            break
        ss = s_safe()
        Th = lam / t
        zs = Z + mu / ss
        Dm = np.where(soft[None], zs + Th, 1.0)
        Tt = np.where(soft[None], Th * zs / Dm, Th)      # = Th (1 - Th/D), without the cancellation
        M = P + np.einsum('bnm,bm,bmk->bnk', GT, Tt, G)
        try:
            Lc = np.linalg.cholesky(M)
        except np.linalg.LinAlgError:
            Lc = None

        def newton(kap, nu):
            rho = (-r_s + kap / t + nu / ss + Th * r_p) * sf
            coef = np.where(soft[None], -kap / t + Th * (rho / Dm - r_p), -kap / t - Th * r_p)
            rhs = -r_v - np.einsum('bnm,bm->bn', GT, coef)
            if Lc is not None:
                LT = np.swapaxes(Lc, 1, 2)
                dv = np.linalg.solve(LT, np.linalg.solve(Lc, rhs[..., None]))[..., 0]
                # one step of iterative refinement (M is ill-conditioned close to the solution)
                rr = rhs - np.einsum('bij,bj->bi', M, dv)
                dv = dv + np.linalg.solve(LT, np.linalg.solve(Lc, rr[..., None]))[..., 0]
            else:
                dv = np.linalg.solve(M, rhs[..., None])[..., 0]
            gd = np.einsum('bmn,bn->bm', G, dv)
            ds = (rho - Th * gd) / Dm * sf
            dt = gd + ds - r_p
            dl = (kap - lam * dt) / t
            dm = (nu - mu * ds) / ss * sf
            return dv, ds, dt, dl, dm

        def max_step(x, dx, mask=None):
            with np.errstate(divide='ignore', invalid='ignore'):
                a = np.where(dx < 0, -x / dx, np.inf)
            if mask is not None:
                a = np.where(mask, a, np.inf)
            return np.minimum(a.min(1), 1e30)

        # Mehrotra's corrector can cycle on poorly centred iterates: once the largest product has not dropped
        # by 30 % over two iterations an instance switches to plain damped path following (sigma >= 0.2, no
        # second-order term).  [robustness heuristic of this repository, not HPIPM's]
        hist.append(cmax.copy())
        if it >= 2:
            safe |= cmax > 0.7 * hist[-3]
        so = np.where(safe, 0.0, 1.0)[:, None]
        dv, ds, dt, dl, dm = newton(-lam * t, -mu * s)
        a_aff = np.minimum(1.0, np.minimum.reduce([max_step(t, dt), max_step(lam, dl),
                                                   max_step(s, ds, soft[None]), max_step(mu, dm, soft[None])]))
        aa = a_aff[:, None]
        mu_aff = ((lam + aa * dl) * (t + aa * dt) + (mu + aa * dm) * (s + aa * ds)).sum(1) / ncomp
        sig = np.clip((mu_aff / np.maximum(mubar, 1e-300)) ** 3, np.where(safe, 0.2, 0.0), 1.0)
        tau = np.maximum(sig * mubar, 0.1 * tol)[:, None]   # do not drive products below the tolerance
        dv, ds, dt, dl, dm = newton(tau - lam * t - so * dl * dt, (tau - mu * s - so * dm * ds) * sf)
        a = np.minimum(1.0, 0.995 * np.minimum.reduce([max_step(t, dt), max_step(lam, dl),
                                                       max_step(s, ds, soft[None]), max_step(mu, dm, soft[None])]))
        a = np.where(done, 0.0, a)[:, None]
        v = v + a * dv
        s = s + a * ds
        t = t + a * dt
        lam = lam + a * dl
        mu = mu + a * dm
    iters[~done] = max_iter
    return dict(v=v, lam=lam, s=s, mu=mu, t=t, iters=iters, status=status)
