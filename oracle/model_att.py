"""Oracle: quadrotor model 'att' (roll/pitch/yaw-rate inputs), batched numpy float64.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Restates, with hand-derived analytic Jacobians:
  * dynamics        ``sdf_nmpc/model/quad_rollpitchyawrate.py:20-42``
    (helpers ``sdf_nmpc/utils/math.py:7-54,169-192``)
  * NLS outputs     ``quad_rollpitchyawrate.py:49-55``
  * FoV constraints ``sdf_nmpc/model/cost_const_helpers.py:48-75``
  * SDF query point ``sdf_nmpc/gen_model.py:46-51``
  * ERK4 + forward sensitivities: acados ``sim_erk`` (4 stages, 1 step) on the
    variational ODE - SURVEY.md 8a-5 item 6 [from memory of acados 0.3.1].
All functions take a leading batch axis.  State x = [p(3), q(4, scalar first), v(3)],
input u = [T/m, roll, pitch, wz] normalised.
"""
import numpy as np

DEFAULT = dict(g=9.81, gamma=20.0, roll=0.7, pitch=0.7, wz=3.0, mass=1.46,
               B_p_C=np.array([0.18, 0.0, -0.025]), B_R_C=np.eye(3), fov_off=0.05,
               hfov=0.7592, vfov=0.4903, fov_ratio=0.9)

NX, NU, NY, NYN = 10, 4, 11, 4


def _parts(x, u, c):
    q = x[:, 3:7]
    n = np.linalg.norm(q, axis=1, keepdims=True)
    qh = q / n
    psi = 2.0 * np.arctan2(q[:, 3], q[:, 0])
    gam, r, pt, wz = c['gamma'] * u[:, 0], c['roll'] * u[:, 1], c['pitch'] * u[:, 2], c['wz'] * u[:, 3]
    return q, n, qh, psi, gam, r, pt, wz


def f_expl(x, u, c=DEFAULT):
    q, n, qh, psi, gam, r, pt, wz = _parts(x, u, c)
    cps, sps, cr, sr, cp, sp = np.cos(psi), np.sin(psi), np.cos(r), np.sin(r), np.cos(pt), np.sin(pt)
    a = np.stack([gam * (cps * cr * sp + sps * sr), gam * (sps * cr * sp - cps * sr), gam * cr * cp - c['g']], 1)
    dq = 0.5 * wz[:, None] * np.stack([-qh[:, 3], qh[:, 2], -qh[:, 1], qh[:, 0]], 1)
    return np.concatenate([x[:, 7:10], dq, a], 1)


_M = np.array([[0, 0, 0, -1.0], [0, 0, 1.0, 0], [0, -1.0, 0, 0], [1.0, 0, 0, 0]])


def f_jac(x, u, c=DEFAULT):
    """(df/dx [n,10,10], df/du [n,10,4])."""
    nb = x.shape[0]
    q, n, qh, psi, gam, r, pt, wz = _parts(x, u, c)
    cps, sps, cr, sr, cp, sp = np.cos(psi), np.sin(psi), np.cos(r), np.sin(r), np.cos(pt), np.sin(pt)
    fx = np.zeros((nb, 10, 10))
    fu = np.zeros((nb, 10, 4))
    fx[:, 0, 7] = fx[:, 1, 8] = fx[:, 2, 9] = 1.0
    dqh = (np.eye(4)[None] - qh[:, :, None] * qh[:, None, :]) / n[:, :, None]
    fx[:, 3:7, 3:7] = 0.5 * wz[:, None, None] * (_M[None] @ dqh)
    fu[:, 3:7, 3] = 0.5 * c['wz'] * (qh @ _M.T)
    den = q[:, 0] ** 2 + q[:, 3] ** 2
    dpsi0, dpsi3 = -2.0 * q[:, 3] / den, 2.0 * q[:, 0] / den
    da_dpsi = np.stack([gam * (-sps * cr * sp + cps * sr), gam * (cps * cr * sp + sps * sr), np.zeros(nb)], 1)
    fx[:, 7:10, 3] = da_dpsi * dpsi0[:, None]
    fx[:, 7:10, 6] = da_dpsi * dpsi3[:, None]
    dirv = np.stack([cps * cr * sp + sps * sr, sps * cr * sp - cps * sr, cr * cp], 1)
    fu[:, 7:10, 0] = c['gamma'] * dirv
    fu[:, 7:10, 1] = c['roll'] * gam[:, None] * np.stack([-cps * sr * sp + sps * cr, -sps * sr * sp - cps * cr, -sr * cp], 1)
    fu[:, 7:10, 2] = c['pitch'] * gam[:, None] * np.stack([cps * cr * cp, sps * cr * cp, -cr * sp], 1)
    return fx, fu


def rk4(x, u, dt, c=DEFAULT):
    k1 = f_expl(x, u, c)
    k2 = f_expl(x + 0.5 * dt * k1, u, c)
    k3 = f_expl(x + 0.5 * dt * k2, u, c)
    k4 = f_expl(x + dt * k3, u, c)
    return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def rk4_sens(x, u, dt, c=DEFAULT):
    """One ERK4 step with forward sensitivities: (x+ [n,10], A [n,10,10], B [n,10,4])."""
    nb = x.shape[0]
    S0 = np.concatenate([np.broadcast_to(np.eye(10), (nb, 10, 10)), np.zeros((nb, 10, 4))], 2)

    def vde(xs, S):
        fx, fu = f_jac(xs, u, c)
        dS = fx @ S
        dS[:, :, 10:] += fu
        return f_expl(xs, u, c), dS

    k1, K1 = vde(x, S0)
    k2, K2 = vde(x + 0.5 * dt * k1, S0 + 0.5 * dt * K1)
    k3, K3 = vde(x + 0.5 * dt * k2, S0 + 0.5 * dt * K2)
    k4, K4 = vde(x + dt * k3, S0 + dt * K3)
    xn = x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    S = S0 + dt / 6.0 * (K1 + 2 * K2 + 2 * K3 + K4)
    return xn, S[:, :, :10], S[:, :, 10:]


def y_fun(x, u, qd, c=DEFAULT):
    """Stage NLS output y [n,11] and Jacobian wrt [u;x] [n,11,14]; yN = y[:, :4] (u-columns are zero there)."""
    nb = x.shape[0]
    q, n, qh, psi, gam, r, pt, wz = _parts(x, u, c)
    cr, sr, cp, sp = np.cos(r), np.sin(r), np.cos(pt), np.sin(pt)
    qe3 = -qd[:, 0] * qh[:, 3] - qd[:, 1] * qh[:, 2] + qd[:, 2] * qh[:, 1] + qd[:, 3] * qh[:, 0]
    y = np.concatenate([x[:, 0:3], qe3[:, None], x[:, 7:10], r[:, None], pt[:, None], wz[:, None],
                        (gam * cr * cp - c['g'])[:, None]], 1)
    J = np.zeros((nb, 11, 14))
    J[:, 0, 4] = J[:, 1, 5] = J[:, 2, 6] = 1.0
    dqh = (np.eye(4)[None] - qh[:, :, None] * qh[:, None, :]) / n[:, :, None]
    coef = np.stack([qd[:, 3], qd[:, 2], -qd[:, 1], -qd[:, 0]], 1)
    J[:, 3, 7:11] = np.einsum('bi,bij->bj', coef, dqh)
    J[:, 4, 11] = J[:, 5, 12] = J[:, 6, 13] = 1.0
    J[:, 7, 1] = c['roll']
    J[:, 8, 2] = c['pitch']
    J[:, 9, 3] = c['wz']
    J[:, 10, 0] = c['gamma'] * cr * cp
    J[:, 10, 1] = -gam * sr * cp * c['roll']
    J[:, 10, 2] = -gam * cr * sp * c['pitch']
    return y, J


def sdf_query(x, W_p_Co, W_R_Co):
    """Co_p_B = W_R_Co^T (p - W_p_Co)  (gen_model.py:46-51)."""
    return np.einsum('bji,bj->bi', W_R_Co, x[:, 0:3] - W_p_Co)


def h_fov(x, W_p_Co, W_R_Co, flag, c=DEFAULT):
    """FoV rows (h1,h2) [n,2] and d/dp [n,2,3]  (cost_const_helpers.py:56-65)."""
    c0 = sdf_query(x, W_p_Co, W_R_Co)
    o = c['B_R_C'].T @ c['B_p_C'] + np.array([c['fov_off'], 0.0, 0.0])
    cc = c0 + o
    cx, cy, cz = cc[:, 0], cc[:, 1], cc[:, 2]
    rho2 = cx * cx + cy * cy
    rho = np.sqrt(rho2)
    h = np.stack([np.arctan2(cy, cx), np.arctan2(cz, rho)], 1) * flag[:, None]
    g1 = np.stack([-cy / rho2, cx / rho2, np.zeros_like(cx)], 1)
    g2 = np.stack([-cz * cx / rho, -cz * cy / rho, rho], 1) / (rho2 + cz * cz)[:, None]
    G = np.stack([g1, g2], 1)                                   # d h / d c  [n,2,3]
    Jp = np.einsum('bkc,bpc->bkp', G, W_R_Co) * flag[:, None, None]   # d c / d p = R^T
    return h, Jp


def u_to_TRPYr(u, c=DEFAULT):
    """[gamma*mass, roll, pitch, wz] (quad_rollpitchyawrate.py:46)."""
    return np.stack([c['gamma'] * u[:, 0] * c['mass'], c['roll'] * u[:, 1], c['pitch'] * u[:, 2], c['wz'] * u[:, 3]], 1)


def u_to_acc(x, u, c=DEFAULT):
    """[W_R_B^T W_a, wz] (quad_rollpitchyawrate.py:45)."""
    q, n, qh, psi, gam, r, pt, wz = _parts(x, u, c)
    a = f_expl(x, u, c)[:, 7:10]
    cps, sps, cr, sr, cp, sp = np.cos(psi), np.sin(psi), np.cos(r), np.sin(r), np.cos(pt), np.sin(pt)
    z = np.zeros_like(cps)
    Rz = np.stack([np.stack([cps, -sps, z], 1), np.stack([sps, cps, z], 1), np.stack([z, z, z + 1], 1)], 1)
    Vb = np.stack([np.stack([cp, sr * sp, cr * sp], 1), np.stack([z, cr, -sr], 1), np.stack([-sp, sr * cp, cr * cp], 1)], 1)
    R = Rz @ Vb
    return np.concatenate([np.einsum('bji,bj->bi', R, a), wz[:, None]], 1)
