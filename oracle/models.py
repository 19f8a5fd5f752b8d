"""Oracle: the reference's reachable quadrotor models behind one interface, batched numpy float64.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

``get(name)`` returns the model selected by ``cfg.mpc.model`` (``sdf_nmpc/gen_model.py:12-21``):

  * ``att``      ``sdf_nmpc/model/quad_rollpitchyawrate.py``      hand-derived Jacobians, ``oracle/model_att.py``
  * ``att_tau``  ``sdf_nmpc/model/quad_rollpitchyawrate_tau.py``  first-order roll/pitch lag, full-quaternion attitude
  * ``acc``      ``sdf_nmpc/model/quad_acc.py``                   world-frame acceleration inputs
  * ``props``    ``sdf_nmpc/model/quad_props.py``                 13 states [p, q, v, w], rotor-speed inputs, allocation matrices

For ``att_tau`` and ``acc`` the model functions are written once on a small array-valued forward-mode type ``Jet``
(value [n] + all 14 tangents [n,14] at once, ordered [u; x] like the QP), so that the ERK4 step, the NLS outputs
and their exact derivatives come from the same expression - this is what CasADi's AD does for the reference
(``base_model.py:83-109``).  ``tests/test_oracle_golden.py`` pins the values against the reference's own numpy math
helpers and the derivatives against central finite differences.
"""
import numpy as np

from . import model_att as A

NX, NU, NW, NY = 10, 4, 14, 11


class Jet:
    """v: [n], d: [n, m] (m tangent directions)."""
    __slots__ = ('v', 'd')
    __array_priority__ = 1000
    __array_ufunc__ = None          # numpy scalars/arrays on the left defer to __radd__/__rmul__

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def lift(a, like):
        return a if isinstance(a, Jet) else Jet(np.broadcast_to(np.asarray(a, float), like.v.shape), np.zeros_like(like.d))

    def __add__(self, o):
        return Jet(self.v + o.v, self.d + o.d) if isinstance(o, Jet) else Jet(self.v + o, self.d)
    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, -self.d)

    def __sub__(self, o):
        return Jet(self.v - o.v, self.d - o.d) if isinstance(o, Jet) else Jet(self.v - o, self.d)

    def __rsub__(self, o):
        return Jet(o - self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v * o.v, self.d * o.v[:, None] + o.d * self.v[:, None])
        o = np.asarray(o, float)
        return Jet(self.v * o, self.d * (o[:, None] if o.ndim else o))
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Jet):
            q = self.v / o.v
            return Jet(q, (self.d - q[:, None] * o.d) / o.v[:, None])
        return self * (1.0 / np.asarray(o, float))

    def __rtruediv__(self, o):
        return Jet.lift(o, self) / self


def _un(fv, fd):
    def f(a):
        if isinstance(a, Jet):
            return Jet(fv(a.v), fd(a.v)[:, None] * a.d)
        return fv(a)
    return f


sin = _un(np.sin, np.cos)
cos = _un(np.cos, lambda v: -np.sin(v))
sqrt = _un(np.sqrt, lambda v: 0.5 / np.sqrt(v))
asin = _un(np.arcsin, lambda v: 1.0 / np.sqrt(1.0 - v * v))


def atan2(y, x):
    if isinstance(y, Jet) or isinstance(x, Jet):
        ref = y if isinstance(y, Jet) else x
        y, x = Jet.lift(y, ref), Jet.lift(x, ref)
        r2 = x.v * x.v + y.v * y.v
        return Jet(np.arctan2(y.v, x.v), (x.v[:, None] * y.d - y.v[:, None] * x.d) / r2[:, None])
    return np.arctan2(y, x)


def _cols(a):
    return [a[:, i] for i in range(a.shape[1])]


def _seed(x, u):
    """Jets of the components of x [n,nx] and u [n,4], tangent columns ordered [u; x]."""
    n, nx = x.shape
    nw = nx + NU
    E = np.eye(nw)
    xs = [Jet(x[:, i], np.broadcast_to(E[NU + i], (n, nw)).copy()) for i in range(nx)]
    us = [Jet(u[:, j], np.broadcast_to(E[j], (n, nw)).copy()) for j in range(NU)]
    return xs, us


def _unit_quat(x):
    q0, q1, q2, q3 = x[3], x[4], x[5], x[6]
    n = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3)
    return q0 / n, q1 / n, q2 / n, q3 / n


def _qe3(qd, h):
    """Fourth component of q_d (x) conj(q)  (utils/math.py hamilton_prod, invert)."""
    return qd[:, 2] * h[1] + qd[:, 3] * h[0] - qd[:, 0] * h[3] - qd[:, 1] * h[2]


# ----------------------------------------------------------------- att_tau (quad_rollpitchyawrate_tau.py:22-52) -----
def _f_att_tau(x, u, c):
    q0, q1, q2, q3 = _unit_quat(x)
    roll = atan2(2.0 * (q0 * q1 + q2 * q3), 1.0 - 2.0 * (q1 * q1 + q2 * q2))        # quat2euler, utils/math.py:57-70
    pitch = asin(2.0 * (q0 * q2 - q3 * q1))
    gam, wz = c['gamma'] * u[0], c['wz'] * u[3]
    dot_roll = (c['roll'] * u[1] - roll) / c['tau_roll']
    dot_pitch = (c['pitch'] * u[2] - pitch) / c['tau_pitch']
    # deuler_avel_map(eta) @ [dot_roll, dot_pitch, 0], rows 0 and 1 as written in utils/math.py:209-226
    w0 = dot_roll + (sin(pitch) * sin(roll) / cos(pitch)) * dot_pitch
    w1 = cos(roll) * dot_pitch
    dq = [0.5 * (-(q1 * w0) - q2 * w1 - q3 * wz), 0.5 * (q0 * w0 + q2 * wz - q3 * w1),
          0.5 * (q0 * w1 - q1 * wz + q3 * w0), 0.5 * (q0 * wz + q1 * w1 - q2 * w0)]
    a = [gam * (2.0 * (q1 * q3 + q0 * q2)), gam * (2.0 * (q2 * q3 - q0 * q1)),
         gam * (q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3) - c['g']]
    return [x[7], x[8], x[9]] + dq + a


def _y_att_tau(x, u, qd, c):
    h = _unit_quat(x)
    az = (c['gamma'] * u[0]) * (h[0] * h[0] - h[1] * h[1] - h[2] * h[2] + h[3] * h[3]) - c['g']
    return [x[0], x[1], x[2], _qe3(qd, h), x[7], x[8], x[9], c['roll'] * u[1], c['pitch'] * u[2], c['wz'] * u[3], az]


# ----------------------------------------------------------------- acc (quad_acc.py:20-45) ------------------------
def _f_acc(x, u, c):
    q0, q1, q2, q3 = _unit_quat(x)
    hw = 0.5 * (c['wz'] * u[3])
    return [x[7], x[8], x[9], -(hw * q3), hw * q2, -(hw * q1), hw * q0, c['ax'] * u[0], c['ay'] * u[1], c['az'] * u[2]]


def _y_acc(x, u, qd, c):
    h = _unit_quat(x)
    return [x[0], x[1], x[2], _qe3(qd, h), x[7], x[8], x[9], c['ax'] * u[0], c['ay'] * u[1], c['az'] * u[2], c['wz'] * u[3]]


# ----------------------------------------------------------------- props (quad_props.py:18-59) -------------------
def axis_rot(axis, angle):
    """utils/math.py:357-364."""
    c, s_ = np.cos(angle), np.sin(angle)
    if axis == 'x':
        return np.array([[1, 0, 0], [0, c, -s_], [0, s_, c]])
    if axis == 'y':
        return np.array([[c, 0, s_], [0, 1, 0], [-s_, 0, c]])
    return np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])


def props_allocation(motors, cf, ct):
    """(Gf, Gt) [3,4] with the force coefficient folded in, as quad_props.py:18-27 builds them: propeller frames
    Rz(i pi / (n/2)) Ry(beta_i) Rx((-1)^i alpha_i), GTMRP_matrix (utils/math.py:385-396), then both scaled by cf."""
    px, py, pz, alpha, beta, sign = zip(*motors)
    n = len(sign)
    R = [axis_rot('z', i * (np.pi / (n / 2))) @ axis_rot('y', beta[i]) @ axis_rot('x', (-1) ** i * alpha[i]) for i in range(n)]
    pos = np.array([px, py, pz], float).T
    Rz = [r @ np.array([0.0, 0.0, 1.0]) for r in R]
    Gf = np.column_stack(Rz)
    Gt = np.column_stack([np.cross(pos[i], Rz[i]) + ct / cf * sign[i] * Rz[i] for i in range(n)])
    return cf * Gf, cf * Gt


def _quat2rot(q0, q1, q2, q3):
    """utils/math.py:7-23, rows."""
    return [[q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2.0 * (q1 * q2 - q0 * q3), 2.0 * (q1 * q3 + q0 * q2)],
            [2.0 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2.0 * (q2 * q3 - q0 * q1)],
            [2.0 * (q1 * q3 - q0 * q2), 2.0 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3]]


def _f_props(x, u, c):
    q0, q1, q2, q3 = _unit_quat(x)                     # quad_props.py:30 normalises before anything else
    w = x[10:13]
    s = [(c['wp_max'] * uj) * (c['wp_max'] * uj) for uj in u]
    Gf, Gt, J, m = c['Gf'], c['Gt'], c['inertia'], c['mass']
    F = [(Gf[i, 0] * s[0] + Gf[i, 1] * s[1] + Gf[i, 2] * s[2] + Gf[i, 3] * s[3]) / m for i in range(3)]
    tq = [Gt[i, 0] * s[0] + Gt[i, 1] * s[1] + Gt[i, 2] * s[2] + Gt[i, 3] * s[3] for i in range(3)]
    R = _quat2rot(q0, q1, q2, q3)
    a = [R[i][0] * F[0] + R[i][1] * F[1] + R[i][2] * F[2] for i in range(3)]
    a[2] = a[2] - c['g']
    dq = [0.5 * (-(q1 * w[0]) - q2 * w[1] - q3 * w[2]), 0.5 * (q0 * w[0] + q2 * w[2] - q3 * w[1]),
          0.5 * (q0 * w[1] - q1 * w[2] + q3 * w[0]), 0.5 * (q0 * w[2] + q1 * w[1] - q2 * w[0])]
    Jw = [J[0] * w[0], J[1] * w[1], J[2] * w[2]]
    cr = [w[1] * Jw[2] - w[2] * Jw[1], w[2] * Jw[0] - w[0] * Jw[2], w[0] * Jw[1] - w[1] * Jw[0]]
    dw = [(tq[i] - cr[i]) / J[i] for i in range(3)]
    return [x[7], x[8], x[9]] + dq + a + dw


def _y_props(x, u, qd, c):
    h = _unit_quat(x)
    roll = atan2(2.0 * (h[0] * h[1] + h[2] * h[3]), 1.0 - 2.0 * (h[1] * h[1] + h[2] * h[2]))
    pitch = asin(2.0 * (h[0] * h[2] - h[3] * h[1]))
    return [x[0], x[1], x[2], roll, pitch, _qe3(qd, h)] + list(x[7:13]) + [c['wp_max'] * uj for uj in u]


def props_u_to_acc(x, u, c):
    """[W_R_B^T W_a, w_z] (quad_props.py:53), numpy [n,13], [n,4] -> [n,4]."""
    f = _f_props(_cols(x), _cols(u), c)
    q = _unit_quat(_cols(x))
    R = _quat2rot(*q)
    Wa = f[7:10]
    return np.stack([R[0][i] * Wa[0] + R[1][i] * Wa[1] + R[2][i] * Wa[2] for i in range(3)] + [x[:, 12]], 1)


def _stack_v(lst, n):
    return np.stack([np.broadcast_to(e.v if isinstance(e, Jet) else e, (n,)) for e in lst], 1)


def _stack_d(lst, n):
    nw = next(e.d.shape[1] for e in lst if isinstance(e, Jet))
    return np.stack([e.d if isinstance(e, Jet) else np.zeros((n, nw)) for e in lst], 1)


def _rk4_generic(f, x, u, dt, c):
    def ax(x, s, k):
        return [xi + s * ki for xi, ki in zip(x, k)]
    k1 = f(x, u, c)
    k2 = f(ax(x, 0.5 * dt, k1), u, c)
    k3 = f(ax(x, 0.5 * dt, k2), u, c)
    k4 = f(ax(x, dt, k3), u, c)
    return [xi + (dt / 6.0) * (a + 2.0 * b + 2.0 * cc + d) for xi, a, b, cc, d in zip(x, k1, k2, k3, k4)]


class JetModel:
    nx, ny = NX, NY

    def __init__(self, name, f, y, nyN, lbu, u_hover, nx=NX, ny=NY):
        self.name, self._f, self._y, self.nyN = name, f, y, nyN
        self.nx, self.ny = nx, ny
        self.lbu, self.ubu = np.asarray(lbu, float), np.ones(4)
        self._u_hover = u_hover

    def u_hover(self, c):
        return self._u_hover(c)

    def f_expl(self, x, u, c):
        return _stack_v(self._f(_cols(x), _cols(u), c), x.shape[0])

    def rk4(self, x, u, dt, c):
        return _stack_v(_rk4_generic(self._f, _cols(x), _cols(u), dt, c), x.shape[0])

    def rk4_sens(self, x, u, dt, c):
        xs, us = _seed(x, u)
        xn = _rk4_generic(self._f, xs, us, dt, c)
        n = x.shape[0]
        S = _stack_d(xn, n)
        return _stack_v(xn, n), S[:, :, NU:], S[:, :, :NU]

    def y_fun(self, x, u, qd, c):
        xs, us = _seed(x, u)
        y = self._y(xs, us, qd, c)
        n = x.shape[0]
        return _stack_v(y, n), _stack_d(y, n)


class _Att:
    name, nyN, nx, ny = 'att', 4, NX, NY
    lbu, ubu = np.array([0.0, -1, -1, -1]), np.ones(4)

    @staticmethod
    def u_hover(c):
        return np.array([c['g'] / c['gamma'], 0, 0, 0])

    f_expl = staticmethod(A.f_expl)
    rk4 = staticmethod(A.rk4)
    rk4_sens = staticmethod(A.rk4_sens)
    y_fun = staticmethod(A.y_fun)


_MODELS = {
    'att': _Att(),
    # quad_rollpitchyawrate_tau.py:44,57-58
    'att_tau': JetModel('att_tau', _f_att_tau, _y_att_tau, 4, [0, -1, -1, -1], lambda c: np.array([c['g'] / c['gamma'], 0, 0, 0])),
    # quad_acc.py:35,50-51
    'acc': JetModel('acc', _f_acc, _y_acc, 7, [-1, -1, -1, -1], lambda c: np.zeros(4)),
    # quad_props.py:12-15,62-63.  DEVIATION from quad_props.py:51-52: the reference sets u_hover to the hover rotor SPEED wh = 12.6
    # although the inputs are normalised speeds bounded to [0, 1] (wp = u * limits.wp, line 35).  Started there, the first
    # linearisation sits at 625 x the hover thrust and every QP of the first control steps ends at the iteration cap (measured with
    # this oracle).  The hover input used here is wh / limits.wp, the value for which f(x_hover, u_hover) = 0.
    'props': JetModel('props', _f_props, _y_props, 12, [0, 0, 0, 0],
                      lambda c: np.full(4, np.sqrt(c['mass'] * c['g'] / 4 / c['cf']) / c['wp_max']), nx=13, ny=16),
}
MODEL_ID = {'att': 0, 'att_tau': 1, 'acc': 2, 'props': 3}


def get(name):
    return _MODELS[name]
