"""Batched reference generation - the step that feeds ``BatchedNmpc.set_ref*``.

``BatchedRefGen`` keeps the method names and the exact arithmetic of ``sdf_nmpc.ref_gen.RefGen``
(``sdf_nmpc/ref_gen.py:7-128``) with a leading batch axis: every method returns a ``RefTraj`` holding
``p [B,N+1,3]``, ``q [B,N+1,4]``, ``v [B,N+1,3]``, ``wz [B,N+1]`` and the position weight ``Wp`` (the joystick mode zeroes
it, ``ref_gen.py:121``).  ``RefTraj.arrays(model, weights)`` applies the model's ``formate_ref`` row layout to all nodes
at once and yields the ``yref / W / yrefN / WN / qd`` arrays of ``BatchedNmpc.set_ref_arrays``.

Faithful-to-the-letter details of the reference that are kept on purpose:
  * ``force_yaw_current`` compares ``cfg.ref.yaw_mode`` with the string ``'curent'`` (``ref_gen.py:12``), so the yaw mode
    ``'current'`` takes the final ``else`` branch (identity quaternion) like it does there;
  * the heuristic ``vref = min(cfg.ref.vref, total_distance)`` and ``np.arange(0, total, T/N*vref)`` sampling
    (``ref_gen.py:57-61``); padding repeats the last position / quaternion with zero velocity (``ref_gen.py:92-96``);
  * the reference's ``from_x0`` and stop-and-turn branches return ``N`` identical entries; here all ``N+1`` nodes carry
    that entry (node N would raise ``IndexError`` in a caller that walks ``range(N+1)``).
All tensors are float64 on the device of ``x0``; nothing here launches custom kernels - it is bookkeeping that precedes
the hot path.
"""
import math
from dataclasses import dataclass

import torch


def quat2yaw(q):
    """utils/math.py:73-82, batched on the last axis."""
    return torch.atan2(2 * (q[..., 0] * q[..., 3] + q[..., 1] * q[..., 2]), 1 - 2 * (q[..., 2] ** 2 + q[..., 3] ** 2))


def yaw2quat(yaw):
    """utils/math.py:142-160 with roll = pitch = 0."""
    h = 0.5 * yaw
    z = torch.zeros_like(h)
    return torch.stack([torch.cos(h), z, z, torch.sin(h)], -1)


_WROW = {}


def _unit_quat(B, dev, dt):
    q = torch.zeros(B, 4, dtype=dt, device=dev)
    q[:, 0] = 1.0
    return q


@dataclass
class RefTraj:
    p: torch.Tensor     # [B,N+1,3]
    q: torch.Tensor     # [B,N+1,4]
    v: torch.Tensor     # [B,N+1,3]
    wz: torch.Tensor    # [B,N+1]
    zero_Wp: bool = False

    def arrays(self, model, weights, wh=None):
        """(yref [B,N,ny], W [B,N,ny], yrefN [B,nyN], WN [B,nyN], qd [B,N+1,4]) with the row layout of the model's
        ``formate_ref`` (quad_rollpitchyawrate.py:62-65, quad_rollpitchyawrate_tau.py:61-64, quad_acc.py:54-58,
        quad_props.py:66-69 - ``wh``: the hover rotor speed that model uses as input reference).
        ``weights``: mapping with pos, vel, att, rates (3 each) and acc (scalar) - one of ``cfg.mpc.weights.set_const_*``."""
        B, N1, _ = self.p.shape
        dev, dt = self.p.device, self.p.dtype
        zero = torch.zeros(B, N1, 1, device=dev, dtype=dt)
        Wp = [0.0, 0.0, 0.0] if self.zero_Wp else list(weights['pos'])
        Wq, Wv, Ww, Wa = list(weights['att']), list(weights['vel']), list(weights['rates']), float(weights['acc'])
        if model == 'props':
            hov = torch.full((B, N1, 4), float(wh), device=dev, dtype=dt)
            y = torch.cat([self.p, zero, zero, zero, self.v, zero, zero, self.wz[..., None], hov], -1)
            Wrow = Wp + Wq + Wv + Ww + [Wa] * 4
            nyN = 12
        elif model == 'acc':
            y = torch.cat([self.p, zero, self.v, zero, zero, zero, self.wz[..., None]], -1)
            Wrow = Wp + [Wq[2]] + Wv + [Wa, Wa, Wa] + [Ww[2]]
            nyN = 7
        else:
            y = torch.cat([self.p, zero, self.v, zero, zero, self.wz[..., None], zero], -1)
            Wrow = Wp + [Wq[2]] + Wv + [Wq[0], Wq[1]] + [Ww[2]] + [Wa]
            nyN = 4
        key = (model, tuple(Wrow), str(dev), dt)
        if key not in _WROW:             # device constant, built once (no host-to-device copy per control step)
            _WROW[key] = torch.tensor(Wrow, device=dev, dtype=dt)
        W = _WROW[key].expand(B, N1, len(Wrow))
        return (y[:, :-1].contiguous(), W[:, :-1].contiguous(), y[:, -1, :nyN].contiguous(), W[:, -1, :nyN].contiguous(),
                self.q.contiguous())


class BatchedRefGen:
    def __init__(self, cfg):
        self.cfg = cfg
        self.N = int(cfg.mpc.N)
        self.x0 = None
        self.force_yaw_current = (cfg.ref.yaw_mode == 'curent')     # sic (ref_gen.py:12)

    def set_x0(self, x0):
        self.x0 = torch.as_tensor(x0, dtype=torch.float64)
        if self.x0.ndim == 1:
            self.x0 = self.x0[None]

    def _const(self, p, q, v, wz=None):
        B, N1 = p.shape[0], self.N + 1
        wz = torch.zeros(B, dtype=p.dtype, device=p.device) if wz is None else wz
        return RefTraj(p[:, None].expand(B, N1, 3).clone(), q[:, None].expand(B, N1, 4).clone(),
                       v[:, None].expand(B, N1, 3).clone(), wz[:, None].expand(B, N1).clone())

    def from_x0(self):
        """ref_gen.py:18-24: hover at the current position with the current yaw."""
        x0 = self.x0
        return self._const(x0[:, :3], yaw2quat(quat2yaw(x0[:, 3:7])), torch.zeros_like(x0[:, :3]))

    def gen_ref_joystick(self, vwref):
        """ref_gen.py:100-128: (vx, vy, vz, wz) in [-1, 1] scaled by cfg.ref.vref / wzref."""
        x0 = self.x0
        vw = torch.as_tensor(vwref, dtype=x0.dtype, device=x0.device).reshape(-1, 4).expand(x0.shape[0], 4)
        v = vw[:, :3] * float(self.cfg.ref.vref)
        wz = vw[:, 3] * float(self.cfg.ref.wzref)
        q_cur = yaw2quat(quat2yaw(x0[:, 3:7]))
        if self.force_yaw_current:
            q = q_cur
        elif self.cfg.ref.yaw_mode == 'align':
            far = torch.linalg.norm(v[:, :2], dim=1) > float(self.cfg.ref.yaw_align_dmin)
            q = torch.where(far[:, None], yaw2quat(torch.atan2(v[:, 1], v[:, 0])), q_cur)
        else:
            q = _unit_quat(x0.shape[0], x0.device, x0.dtype)
        tr = self._const(x0[:, :3], q, v, wz)
        k = torch.arange(self.N + 1, dtype=x0.dtype, device=x0.device)
        tr.p = x0[:, None, :3] + v[:, None, :] * (k[None, :, None] * float(self.cfg.mpc.T) / self.N)
        tr.zero_Wp = True
        return tr

    def gen_ref_list_wps(self, wps_p, wps_q=None, n_wps=None):
        """ref_gen.py:26-98.  wps_p [B,W,3], wps_q [B,W,4] (default identity), n_wps [B] = valid way-points per instance
        (default W; padding entries are ignored).  Way-points are followed at cfg.ref.vref from the current position."""
        cfg, N = self.cfg, self.N
        x0 = self.x0
        dev, dt = x0.device, x0.dtype
        wps_p = torch.as_tensor(wps_p, dtype=dt, device=dev)
        B, Wn, _ = wps_p.shape
        if wps_q is None:
            wps_q = _unit_quat(B, dev, dt)[:, None].expand(B, Wn, 4)
        wps_q = torch.as_tensor(wps_q, dtype=dt, device=dev)
        n_wps = torch.full((B,), Wn, device=dev, dtype=torch.long) if n_wps is None else torch.as_tensor(n_wps, device=dev).long()
        path_p = torch.cat([x0[:, None, :3], wps_p], 1)                       # [B,W+1,3]
        path_q = torch.cat([x0[:, None, 3:7], wps_q], 1)
        path_yaw = quat2yaw(path_q)
        seg_valid = torch.arange(Wn, device=dev)[None, :] < n_wps[:, None]    # segment j = path[j] -> path[j+1]
        dist = torch.linalg.norm(path_p[:, 1:] - path_p[:, :-1], dim=2) * seg_valid
        cum = torch.cat([torch.zeros(B, 1, dtype=dt, device=dev), torch.cumsum(dist, 1)], 1)   # [B,W+1]
        total = cum[:, -1]
        vref = torch.clamp(total, max=float(cfg.ref.vref))                    # min(vref, total_distance)
        step = float(cfg.mpc.T) / N * vref
        moving = total != 0                                                   # `if total_distance / 1e-3:`
        n_even = torch.where(moving, torch.ceil(total / torch.where(moving, step, torch.ones_like(step))), torch.zeros_like(total))
        n_even = torch.clamp(n_even, max=N + 1).long()                        # loop breaks once N+1 entries exist
        k = torch.arange(N + 1, device=dev)
        d = k[None, :].to(dt) * step[:, None]                                 # np.arange(0, total, step)[k]
        # searchsorted(cum, d) - 1 on the valid prefix of cum (padding made strictly larger than any d)
        cum_s = torch.where(torch.arange(Wn + 1, device=dev)[None, :] <= n_wps[:, None], cum, torch.full_like(cum, math.inf))
        seg = torch.searchsorted(cum_s, d.contiguous(), right=False) - 1
        seg = torch.minimum(torch.clamp(seg, min=0), (n_wps - 1)[:, None].clamp(min=0))
        g3 = seg[..., None].expand(B, N + 1, 3)
        p_a, p_b = torch.gather(path_p, 1, g3), torch.gather(path_p, 1, g3 + 1)
        dseg = torch.gather(dist, 1, seg)
        direction = (p_b - p_a) / dseg[..., None]
        p = p_a + direction * (d - torch.gather(cum, 1, seg))[..., None]
        v = direction * vref[:, None, None]
        # yaw
        q0 = path_q[:, 0]
        if self.force_yaw_current:
            q = q0[:, None].expand(B, N + 1, 4)
        elif cfg.ref.yaw_mode == 'ref':
            q = yaw2quat(torch.gather(path_yaw, 1, seg + 1))
        elif cfg.ref.yaw_mode == 'align':
            far = torch.linalg.norm(path_p[:, 1, :2] - x0[:, :2], dim=1) > float(cfg.ref.yaw_align_dmin)
            qa = yaw2quat(torch.atan2(v[..., 1], v[..., 0]) + float(cfg.ref.align_yaw_offset))
            q = torch.where(far[:, None, None], qa, q0[:, None].expand(B, N + 1, 4))
        else:
            q = _unit_quat(B, dev, dt)[:, None].expand(B, N + 1, 4)
        # padding: repeat the last generated (p, q) with zero velocity; empty trajectory -> last way-point
        valid = k[None, :] < n_even[:, None]
        last = (n_even - 1).clamp(min=0)
        lp = torch.gather(p, 1, last[:, None, None].expand(B, 1, 3))
        lq = torch.gather(q.expand(B, N + 1, 4), 1, last[:, None, None].expand(B, 1, 4))
        endp = torch.gather(path_p, 1, n_wps[:, None, None].expand(B, 1, 3))
        endq = torch.gather(path_q, 1, n_wps[:, None, None].expand(B, 1, 4))
        none = (n_even == 0)[:, None, None]
        lp, lq = torch.where(none, endp, lp), torch.where(none, endq, lq)
        p = torch.where(valid[..., None], p, lp.expand(B, N + 1, 3))
        q = torch.where(valid[..., None], q.expand(B, N + 1, 4), lq.expand(B, N + 1, 4))
        v = torch.where(valid[..., None], v, torch.zeros_like(v))
        tr = RefTraj(p, q.clone(), v, torch.zeros(B, N + 1, dtype=dt, device=dev))
        # stop and turn (ref_gen.py:35-50)
        if cfg.ref.stop_and_turn.enable:
            yaw_curr = path_yaw[:, 0]
            yaw_r = yaw_curr.clone()
            if cfg.ref.yaw_mode == 'topic':
                yaw_r = path_yaw[:, 1].clone()
            elif cfg.ref.yaw_mode == 'align':
                dxy = path_p[:, 1, :2] - x0[:, :2]
                far = torch.linalg.norm(dxy, dim=1) > float(cfg.ref.yaw_align_dmin)
                yaw_r = torch.where(far, torch.atan2(dxy[:, 1], dxy[:, 0]), yaw_r) + float(cfg.ref.align_yaw_offset)
            turn = (yaw_curr - yaw_r).abs() > float(cfg.ref.stop_and_turn.dang_min)
            st = self._const(x0[:, :3], yaw2quat(yaw_r), torch.zeros_like(x0[:, :3]))
            m = turn[:, None, None]
            tr = RefTraj(torch.where(m, st.p, tr.p), torch.where(m, st.q, tr.q), torch.where(m, st.v, tr.v), tr.wz)
            tr.turned = turn            # [B] instances that took the stop-and-turn branch
        return tr


# ------------------------------------------------------------------------------------------------------------------
# Single-instance facade with the reference's own classes and call sequence (``sdf_nmpc/utils/reference.py:6-53``,
# ``sdf_nmpc/ref_gen.py:7-128``): ``refgen.x0 = x; traj = refgen.gen_ref_joystick(vw); mpc.set_ref(traj[k], k)``.
# ------------------------------------------------------------------------------------------------------------------
class Ref:
    """One node's reference: ``p, q, v, wz`` and - set by the caller, as in the reference, or through ``use_weights`` - the
    weights ``Wp, Wq, Wv, Ww, Wa`` that ``Nmpc.set_ref`` reads.  ``W_on`` / ``W_off`` keep the reference's naming
    (``utils/reference.py:15-28``: ``W_on`` holds ``set_const_off`` and vice versa)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.p, self.q, self.v, self.wz = [0., 0., 0.], [1., 0., 0., 0.], [0., 0., 0.], 0.
        w = cfg.mpc.weights

        def pack(s):
            from .config import AttrDict
            return AttrDict({'Wp': s.pos, 'Wq': s.att, 'Wv': s.vel, 'Ww': s.rates, 'Wa': s.acc})
        self.W_on, self.W_off = pack(w.set_const_off), pack(w.set_const_on)

    def use_weights(self, constraints_on=True):
        """Convenience (not in the reference, whose callers assign the five attributes themselves): take the weight set
        ``mpc.weights.set_const_on`` / ``set_const_off`` of the configuration."""
        s = self.cfg.mpc.weights.set_const_on if constraints_on else self.cfg.mpc.weights.set_const_off
        self.Wp, self.Wq, self.Wv, self.Ww, self.Wa = list(s.pos), list(s.att), list(s.vel), list(s.rates), s.acc
        return self

    def hover_at_state(self, x):
        """utils/reference.py:40-44."""
        import numpy as np
        x = torch.as_tensor(np.asarray(x, float))
        self.p = x[:3].numpy().copy()
        self.q = yaw2quat(quat2yaw(x[3:7])).numpy()
        self.v, self.wz = [0., 0., 0.], 0.


class Waypoint:
    """utils/reference.py:47-53."""

    def __init__(self, p, q=(1, 0, 0, 0)):
        import numpy as np
        self.p, self.q = np.array(p, dtype=float), np.array(q, dtype=float)

    def __str__(self):
        return f'{self.p}, yaw {float(quat2yaw(torch.as_tensor(self.q))):.4f}'


class RefGen:
    """``sdf_nmpc.ref_gen.RefGen`` (``ref_gen.py:7-128``): set ``x0``, then ``from_x0()``, ``gen_ref_list_wps(wps)`` or
    ``gen_ref_joystick(vwref)`` return a list of ``Ref`` (copies of ``self.ref``, so weights assigned to it carry over).
    The arithmetic is ``BatchedRefGen``'s with a batch of one on the CPU (pinned to the reference's outputs,
    ``tests/golden/refgen.npz``); like there, ``from_x0`` and the stop-and-turn branch return N entries, the others N + 1."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.x0 = None
        self.ref = Ref(cfg)
        self._b = BatchedRefGen(cfg)

    def _reset(self):
        self.ref = Ref(self.cfg)

    def _list(self, tr, n):
        import copy
        out = []
        for k in range(n):
            r = copy.copy(self.ref)
            r.p, r.q, r.v, r.wz = tr.p[0, k].numpy().copy(), tr.q[0, k].numpy().copy(), tr.v[0, k].numpy().copy(), float(tr.wz[0, k])
            if getattr(tr, 'zero_Wp', False):
                r.Wp = [0, 0, 0]
            out.append(r)
        return out

    def _x0(self):
        import numpy as np
        self._b.set_x0(torch.as_tensor(np.asarray(self.x0, float)[None, :]))

    def from_x0(self):
        self._x0()
        return self._list(self._b.from_x0(), self.cfg.mpc.N)

    def gen_ref_joystick(self, vwref):
        import numpy as np
        self._x0()
        return self._list(self._b.gen_ref_joystick(np.asarray(vwref, float)[None, :4]), self.cfg.mpc.N + 1)

    def gen_ref_list_wps(self, wps):
        import numpy as np
        self._reset()
        self._x0()
        tr = self._b.gen_ref_list_wps(np.stack([w.p for w in wps])[None], np.stack([w.q for w in wps])[None])
        n = self.cfg.mpc.N + 1
        turned = getattr(tr, 'turned', None)
        if turned is not None and bool(turned[0]):
            n = self.cfg.mpc.N              # the stop-and-turn branch (ref_gen.py:35-50) returns N entries
        return self._list(tr, n)
