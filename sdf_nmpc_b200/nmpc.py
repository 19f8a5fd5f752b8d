"""Command interface of the controller, batched over independent NMPC instances.

``BatchedNmpc`` keeps the method names and argument meaning of ``sdf_nmpc.controller.Nmpc``
(``sdf_nmpc/controller.py:7-142``) with a leading batch axis; ``Nmpc`` is the single-instance view with
the reference's exact signatures (numpy float64 in/out).  Where the reference pushes ``yref/W/p`` node
by node into acados through ctypes and calls ``solve_for_x0`` (``sdf_nmpc/ocp.py:159-170``), this class
hands device pointers to ``sdfnmpc_solver_rti_step`` (one SQP-RTI iteration for all instances).

The per-node parameter matrix ``p[N+1, 145]`` of the reference (``controller.py:35,45-61``) is stored
de-duplicated: flag, camera pose and latent are per instance (the reference writes the same value into
every row), only ``q_d`` is per node.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .config import Config
from .sdf import SdfEvaluator, _ptr, _stream

NX, NU, NY, NH = 10, 4, 11, 3                       # models att / att_tau / acc; 'props': nx 13 (+ 1 pad on the device), ny 16
MODEL_ID = {'att': 0, 'att_tau': 1, 'acc': 2, 'props': 3}      # mpc.model values reachable in gen_model.py:15-20


def braking_coeffs(cfg):
    """Coefficients of the braking-distance polynomial (``gen_model.py:76-78``): ``mpc.braking_dist.coeffs`` if the YAML carries
    them inline, else the ``.npy`` file ``mpc.braking_dist.coeff_file`` in the cache directory (``$SDFNMPC_CACHE_DIR`` or
    ``~/.cache/sdf_nmpc``, where the reference's ``scripts/find_amax.py`` writes it)."""
    import os
    bd = cfg.mpc.braking_dist
    deg = int(bd.degree)
    if bd.get('coeffs') is not None:
        co = np.asarray(bd.coeffs, float).ravel()
    else:
        root = os.environ.get('SDFNMPC_CACHE_DIR', os.path.join(os.path.expanduser('~'), '.cache', 'sdf_nmpc'))
        path = os.path.join(root, bd.coeff_file)
        if not os.path.exists(path):
            raise FileNotFoundError(f'flags.recursive_feasibility needs the braking-distance coefficients {path} '
                                    '(scripts/find_amax.py of the reference) or mpc.braking_dist.coeffs in the YAML')
        co = np.load(path).astype(float).ravel()
    n = (deg + 1) * (deg + 2) * (deg + 3) // 6
    if deg > 5 or co.size != n:
        raise ValueError(f'braking_dist: degree {deg} needs {n} coefficients (degree <= 5), got {co.size}')
    return deg, co


def r_tilde_ratio(T, phi, theta, r1, r2, r3, dt, g=9.81):
    """The expression ``utils/stability.py:6-42`` derives with sympy (smallest ``r_tilde`` with ``U~' (r_tilde I) U~ >= U' R U``, ``U = [T - g,
    phi, theta]``, ``U~ = dt (T R e3 - g e3)`` at zero yaw): input cost over the squared velocity change of one step.  Pinned to the
    reference's own symbolic result in ``tests/golden/stability.npz``."""
    num = r1 * (T - g) ** 2 + r2 * phi ** 2 + r3 * theta ** 2
    den = dt ** 2 * (T ** 2 - 2.0 * T * g * np.cos(phi) * np.cos(theta) + g ** 2)
    return num, den


def stability_r_tilde(w, lim, dt, g=9.81):
    """``get_r_tilde_max`` (``utils/stability.py:45-75``): the largest ``r_tilde_ratio`` over the input box ``[0, gamma] x [-roll, roll] x
    [-pitch, pitch]`` - a dense grid here, SLSQP from a random start in the reference."""
    T, ph, th = np.meshgrid(np.linspace(0, lim.gamma, 161), np.linspace(-lim.roll, lim.roll, 41), np.linspace(-lim.pitch, lim.pitch, 41), indexing='ij')
    num, den = r_tilde_ratio(T, ph, th, w.acc, w.att[0], w.att[1], dt, g)
    return float(np.max(np.where(den > 1e-9, num / np.maximum(den, 1e-9), 0.0)))


def stability_p_term(cfg, g=9.81):
    """Terminal weight ``p_term`` of ``gen_model.py:121-147`` (upper bound of the stage cost over the braking manoeuvre).
    ``r_tilde`` (``utils/stability.py:6-75``) is the largest ratio input cost / (dt^2 |T R e3 - g e3|^2 ) over the input box - the
    reference solves its sympy expression with SLSQP from a random start; here a dense grid plus the same ratio in closed form.
    NOTE the reference computes ``p_term`` into ``model.extra_WN`` but ``controller.set_ref`` fills ``WN`` with ``W[:nyN]``
    (``controller.py:140-142``), so the weight that reaches acados for the fifth terminal output is ``Wv[0]``; ``set_ref`` here
    does the same, and callers of ``set_ref_arrays`` choose the weight themselves."""
    # gen_model.py:129-134 reads cfg.mpc.weights.vel / .att / .rates / .acc, which predate the weight SETS of default.yaml:36-48
    # (the flag cannot be switched on with the shipped YAML); the set used while constraints are on stands in for them
    w = cfg.mpc.weights if 'vel' in cfg.mpc.weights else cfg.mpc.weights.set_const_on
    lim, m = cfg.robot.limits, cfg.mpc
    dt = m.T / m.N
    max_vel_error = (2 * cfg.ref.vref) ** 2 * max(w.vel)
    max_att = np.array([lim.roll, lim.pitch, lim.wz])
    max_att_error = float(max_att @ np.diag(list(w.att[:2]) + list(w.rates[2:])) @ max_att)
    max_thrust_error = max(w.acc * (lim.gamma - g) ** 2, w.acc * g ** 2)
    sc_max = max_vel_error + max_att_error + max_thrust_error
    r_tilde = stability_r_tilde(w, lim, dt, g)
    return max(r_tilde + max_vel_error, sc_max / m.stability.a_b_min ** 2 / dt ** 2)


def props_allocation(motors, cf, ct):
    """(Gf, Gt) [3, 4] of ``quad_props.py:18-27``: propeller frames ``Rz(i pi / (n/2)) Ry(beta_i) Rx((-1)^i alpha_i)``, the
    allocation of a generically tilted multi-rotor (``utils/math.py:385-396``), both scaled by the force coefficient."""
    def rot(axis, a):
        c, s = np.cos(a), np.sin(a)
        return np.array({'x': [[1, 0, 0], [0, c, -s], [0, s, c]], 'y': [[c, 0, s], [0, 1, 0], [-s, 0, c]],
                         'z': [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[axis], float)
    px, py, pz, alpha, beta, sign = zip(*motors)
    n = len(sign)
    ez = [rot('z', i * (np.pi / (n / 2))) @ rot('y', beta[i]) @ rot('x', (-1) ** i * alpha[i]) @ np.array([0.0, 0.0, 1.0]) for i in range(n)]
    pos = np.array([px, py, pz], float).T
    Gf = np.column_stack(ez)
    Gt = np.column_stack([np.cross(pos[i], ez[i]) + ct / cf * sign[i] * ez[i] for i in range(n)])
    return cf * Gf, cf * Gt


class BatchedNmpc:
    def __init__(self, cfg, sdf_net=None, batch=1, device='cuda:0'):
        """``sdf_net``: packed network (``weights.sdf_from_module`` / ``synth.make_sdf_net``); ``None`` loads the TorchScript file
        ``cfg.nn.sdf_weights`` from the data directory like ``gen_model.py:32`` does."""
        if sdf_net is None:
            from . import weights
            sdf_net = weights.load_sdf(cfg)
        if cfg.mpc.model not in MODEL_ID:
            raise NotImplementedError(f"mpc.model == '{cfg.mpc.model}' is not built (available: {sorted(MODEL_ID)})")
        self.model = cfg.mpc.model
        self.nyN = {'acc': 7, 'props': 12}.get(self.model, 4)   # quad_acc.py:15 / quad_props.py:15 / quad_rollpitchyawrate*.py:15
        fl = cfg.get('flags', {})
        # flags.recursive_feasibility / stability (gen_model.py:73-149): terminal rows lifted onto auxiliary nodes behind node N
        # (csrc/term_lift.cuh).  `stability` only acts inside the recursive-feasibility branch, as in the reference.
        self.lift = 0
        if fl.get('enable_sdf', True) and fl.get('recursive_feasibility', False):
            if self.model != 'att':
                raise NotImplementedError("recursive feasibility and stability are implemented only for mpc.model == 'att' (gen_model.py:74)")
            self.lift = 2 if fl.get('stability', False) else 1
            if self.lift == 2:
                self.nyN = 5                                    # + flag * v.v (gen_model.py:139-147)
        # stage dimensions: nx as the reference counts it, nxd as the device stores it ('props': 13 states + one pad entry that
        # stays zero - the QP's staging copies move 16-byte units, see model_att.cuh)
        self.nx, self.nxd, self.ny = (13, 14, 16) if self.model == 'props' else (NX, NX, NY)
        self.cfg = cfg
        self.B = int(batch)
        self.N = int(cfg.mpc.N)
        self.T = cfg.mpc.T
        self.device = torch.device(device)
        self.sdf = SdfEvaluator(sdf_net, device)
        self.lib = self.sdf.lib
        self.L = self.sdf.size_latent
        lim = cfg.robot.limits
        self.g = 9.81
        self.dt = cfg.shooting_dt()
        lh, uh = cfg.h_bounds(self.sdf.max_df)
        # mission flags (gen_model.py:26-70): a row the reference would not add is kept with bounds that can never be
        # active, which leaves the optimum unchanged (slack 0, multiplier 0)
        off = [not fl.get('enable_sdf', True) or not (cfg.sensor.hfov < 3.14),
               not fl.get('enable_sdf', True) or not fl.get('vfov_constraint', True),
               not fl.get('enable_sdf', True) or not fl.get('sdf_constraint', True)]
        lh, uh = np.where(off, -1e6, lh), np.where(off, 1e6, uh)
        self.sdf_cost_w = 20.0 if (fl.get('enable_sdf', True) and fl.get('sdf_cost', False)) else 0.0
        zl, Zl = cfg.slack_weights()
        d = _lib.SolverDesc()
        d.N, d.max_batch, d.qp_iter_max, d.h_at_node0 = self.N, self.B, int(cfg.b200.qp_iter_max), int(bool(cfg.b200.h_at_node0))
        d.lm, d.qp_tol, d.qp_tol_res = float(cfg.mpc.lm_reg), float(cfg.b200.qp_tol), float(cfg.b200.qp_tol_res)
        d.g, d.gamma, d.roll, d.pitch, d.wz, d.mass = self.g, lim.gamma, lim.roll, lim.pitch, lim.wz, cfg.robot.mass
        d.cam_off = (C.c_double * 3)(*cfg.cam_offset())
        d.lbu = (C.c_double * 4)(-1 if self.model == 'acc' else 0, -1, -1, -1)   # quad_rollpitchyawrate.py:58-59, quad_acc.py:50-51
        d.model = MODEL_ID[self.model]
        if self.model == 'props':
            # quad_props.py:18-38,51-52,62-63 on robot.alloc / robot.inertia / limits.wp (default.yaml:78-87,108)
            al = cfg.robot.alloc
            Gf, Gt = props_allocation(al.motors, float(al.cf), float(al.ct))
            inertia = np.asarray(cfg.robot.inertia, float)
            d.lbu = (C.c_double * 4)(0, 0, 0, 0)
            d.wp_max = float(lim.wp)
            d.Gf_m = (C.c_double * 12)(*(Gf / cfg.robot.mass).ravel())
            d.JGt = (C.c_double * 12)(*(Gt / inertia[:, None]).ravel())
            d.inertia = (C.c_double * 3)(*inertia)
            self.wh = float(np.sqrt(cfg.robot.mass * self.g / 4 / al.cf))        # hover rotor speed (quad_props.py:51)
            # DEVIATION: quad_props.py:52 starts Ocp.init at u = wh (a rotor speed) although the inputs are normalised speeds
            # in [0, 1]; from there the first QPs end at the iteration cap (oracle/models.py).  The hover input is wh / limits.wp.
            d.u_hover_props = self.wh / float(lim.wp)
        d.ax, d.ay, d.az = float(lim.ax), float(lim.ay), float(lim.az)
        d.tau_roll = d.tau_pitch = 0.12                   # quad_rollpitchyawrate_tau.py:19-20 (constants in the reference)
        d.sdf_cost_w = self.sdf_cost_w
        d.ubu = (C.c_double * 4)(1, 1, 1, 1)
        d.lh, d.uh = (C.c_double * 3)(*lh), (C.c_double * 3)(*uh)
        d.zl, d.Zl = (C.c_double * 3)(*zl), (C.c_double * 3)(*Zl)
        for k in range(self.N):
            d.dt[k] = float(self.dt[k])
        if self.lift:
            deg, co = braking_coeffs(cfg)
            d.term_lift, d.brake_degree = self.lift, deg
            d.brake_coeffs = (C.c_double * 56)(*co)
            d.brake_bounds = (C.c_double * 2)(float(cfg.robot.size.xy), float(self.sdf.max_df))      # gen_model.py:116
            sb = cfg.mpc.weights.get('slack_brake')
            sb = None if isinstance(sb, str) else sb                # default.yaml:51 spells `None`, which YAML reads as a string
            d.slack_brake = (C.c_double * 2)(*(sb if sb else (-1.0, 0.0)))                           # None: hard row
            hl, vl = cfg.sensor.hfov * cfg.mpc.fov_ratio, cfg.sensor.vfov * cfg.mpc.fov_ratio
            d.fov_term = (C.c_double * 2)(hl, vl if fl.get('vfov_constraint', True) else -1.0)
            d.vlim = (C.c_double * 3)(float(lim.vx), float(lim.vy), float(lim.vz))
            d.z_hard = float(cfg.b200.get('z_hard', 1e4))
            self.p_term = stability_p_term(cfg, self.g) if self.lift == 2 else None
        self.NI = self.N + self.lift                      # node stride of the library's own arrays (sdfnmpc_solver_horizon)
        self.handle = C.c_void_p()
        _lib.check(self.lib.sdfnmpc_solver_create(C.byref(self.handle), C.byref(d), self.sdf.handle), 'sdfnmpc_solver_create')
        self._desc = d                                    # the descriptor this handle was built from (diagnostics, tests)
        # command clipping (controller.py:18-26)
        self.cmd_acc_hover = np.array([0, 0, 0, 0])
        self.cmd_acc_min = [-lim.ax, -lim.ay, -lim.az, -lim.wz]
        self.cmd_acc_max = [lim.ax, lim.ay, lim.az, lim.wz]
        self.cmd_TRPYr_hover = np.array([cfg.robot.mass * self.g, 0, 0, 0])
        self.cmd_TRPYr_min = [0, -lim.roll, -lim.pitch, -lim.wz]
        self.cmd_TRPYr_max = [lim.gamma, lim.roll, lim.pitch, lim.wz]
        self.cmd_props_hover = np.array([cfg.robot.mass * self.g / 4] * 4)       # controller.py:27 (as written there)
        self.cmd_props_min = [0] * 4
        self.cmd_props_max = [lim.get('wp', 0.0)] * 4
        f64 = dict(dtype=torch.float64, device=self.device)
        B, N = self.B, self.N
        self.u0 = torch.zeros(B, NU, **f64)
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._timed = False
        self.status = torch.zeros(B, dtype=torch.int32, device=self.device)
        self.qp_iters = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._pin = {}
        self.reset()

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.lib.sdfnmpc_solver_destroy(h)
            self.handle = None

    # ------------------------------------------------------------------ state ----------
    def _t(self, a, shape=None, dtype=torch.float64):
        t = torch.as_tensor(a, dtype=dtype, device=self.device)
        return t.reshape(shape).contiguous() if shape is not None else t.contiguous()

    def _layout(self):
        """Offsets (in doubles) of the per-step inputs inside ONE device block: a host caller fills a pinned mirror with the
        same layout and uploads it with a single copy (``Nmpc``)."""
        B, N = self.B, self.N
        sizes = (('x0', (B, self.nxd)), ('y', (B, N, self.ny)), ('W', (B, N, self.ny)), ('yN', (B, self.nyN)), ('WN', (B, self.nyN)),
                 ('qd', (B, N + 1, 4)), ('flag', (B,)), ('pose', (B, 12)))
        off, lay = 0, {}
        for name, shp in sizes:
            n = int(np.prod(shp))
            lay[name] = (off, n, shp)
            off += n
        return lay, off

    def _bind(self, block):
        """Point the input views at `block` (layout of ``_layout``)."""
        lay, _ = self._layout()
        view = lambda k: block[lay[k][0]:lay[k][0] + lay[k][1]].view(lay[k][2])
        self._in = block
        self._x0 = view('x0')                             # static address: the RTI step may be replayed as a CUDA graph
        self.flag = view('flag')
        self.pose = view('pose')                          # W_p_Co(3), W_R_Co row-major(9)
        self.qd = view('qd')
        self.y, self.yN, self.W, self.WN = view('y'), view('yN'), view('W'), view('WN')

    def reset(self):
        """controller.py:32-42."""
        B = self.B
        _, total = self._layout()
        self._blocks = [torch.zeros(total, dtype=torch.float64, device=self.device)]
        self._bind(self._blocks[0])
        self.x0 = None
        self.latent = torch.zeros(B, max(self.L, 1), dtype=torch.float32, device=self.device)
        self.fail_count = torch.zeros(B, dtype=torch.int64, device=self.device)
        self._latent_dirty = True
        self._graph = None
        self.set_sdf_flag(False)

    def select_inputs(self, i):
        """Double-buffered inputs for a host caller (eager mode): while a control step runs on block ``1 - i`` the references of the
        next step are uploaded into block ``i`` on a copy stream; the state feedback, which is all that has to wait for the
        previous step, goes last.  The second block is created on first use as a copy of the first."""
        if self._graph is not None:
            raise _lib.SdfNmpcError('select_inputs: the captured control step reads the block it was captured with')
        if len(self._blocks) == 1:
            self._blocks.append(self._blocks[0].clone())
        had_x0 = self.x0 is not None
        self._bind(self._blocks[int(i)])
        if had_x0:
            self.x0 = self._x0
        return self._in

    def set_sdf_flag(self, flag):
        """controller.py:45-47 (scalar or per-instance)."""
        self.flag[:] = self._t(flag).to(torch.float64)

    def set_latent(self, latent, W_p_Bo, W_R_Bo):
        """controller.py:50-54: camera-observation pose from the body pose at image time + latent."""
        R = self._t(W_R_Bo, (-1, 3, 3))
        p = self._t(W_p_Bo, (-1, 3))
        if not hasattr(self, '_BpC'):
            self._BpC = self._t(np.asarray(self.cfg.sensor.B_p_C, float))
            self._BRC = self._t(np.asarray(self.cfg.sensor.B_R_C, float))
        BpC, BRC = self._BpC, self._BRC
        # 3 x 3 products written out as broadcast sums: no library GEMM launches on this path
        self.pose[:, 0:3] = (R * BpC).sum(-1) + p
        self.pose[:, 3:12] = (R[:, :, :, None] * BRC[None]).sum(2).reshape(-1, 9)
        self.latent[:] = self._t(latent, (-1, max(self.L, 1)), torch.float32)
        self._latent_dirty = True

    def reset_latent(self):
        self.pose.zero_()
        self.latent.zero_()
        self._latent_dirty = True

    def set_x0(self, x0):
        """controller.py:65-69: the first call initialises the trajectory (ocp.py:144-149)."""
        if x0 is not self._x0:
            self._x0[:, :self.nx].copy_(torch.as_tensor(x0, dtype=torch.float64).reshape(self.B, -1)[:, :self.nx], non_blocking=True)
        if self.x0 is None:
            _lib.check(self.lib.sdfnmpc_solver_init_traj(self.handle, _ptr(self._x0), self.B, _stream()), 'init_traj')
        self.x0 = self._x0

    def set_ref_arrays(self, yref, W, yrefN, WN, qd):
        self.y[:] = self._t(yref, (self.B, self.N, self.ny))
        self.W[:] = self._t(W, (self.B, self.N, self.ny))
        self.yN[:] = self._t(yrefN, (self.B, self.nyN))
        self.WN[:] = self._t(WN, (self.B, self.nyN))
        self.qd[:] = self._t(qd, (self.B, self.N + 1, 4))

    def formate_ref(self, ref):
        """quad_rollpitchyawrate.py:62-65 / quad_rollpitchyawrate_tau.py:61-64 (same rows).  For 'acc' the rows are those
        of quad_acc.py:54-58; that function reads an undefined name in the reference (it cannot run there), so the row
        order it spells out is used with the weights of `ref` itself."""
        if self.model == 'props':          # quad_props.py:66-69
            yr = np.concatenate([ref.p, [0, 0, 0], ref.v, [0, 0, ref.wz], [self.wh] * 4])
            W = np.concatenate([ref.Wp, ref.Wq, ref.Wv, ref.Ww, [ref.Wa] * 4])
            return yr, W
        if self.model == 'acc':
            yr = np.concatenate([ref.p, [0], ref.v, [0, 0, 0], [ref.wz]])
            W = np.concatenate([ref.Wp, ref.Wq[2:], ref.Wv, [ref.Wa, ref.Wa, ref.Wa], [ref.Ww[2]]])
            return yr, W
        yr = np.concatenate([ref.p, [0], ref.v, [0, 0], [ref.wz], [0]])
        W = np.concatenate([ref.Wp, [ref.Wq[2]], ref.Wv, ref.Wq[:2], ref.Ww[2:], [ref.Wa]])
        return yr, W

    def set_ref(self, ref, k, instance=None):
        """controller.py:133-142 for one node (all instances, or one)."""
        sel = slice(None) if instance is None else instance
        y, W = self.formate_ref(ref)
        self.qd[sel, k] = self._t(np.asarray(ref.q, float))
        if k < self.N:
            self.y[sel, k] = self._t(y)
            self.W[sel, k] = self._t(W)
        else:
            self.yN[sel] = self._t(y[:self.nyN])
            self.WN[sel] = self._t(W[:self.nyN])

    # ------------------------------------------------------------------ control iteration ----------
    def shift(self, k=None):
        k = self.cfg.mpc.shift if k is None else k
        if k > 0:
            _lib.check(self.lib.sdfnmpc_solver_shift(self.handle, self.B, int(k), _stream()), 'shift')

    def solve(self):
        """controller.py:72-81 - one SQP-RTI iteration for every instance; returns fail_count [B]."""
        if self.x0 is None:
            raise _lib.SdfNmpcError('set_x0 must be called before solve')
        if self._graph is not None:
            if self._latent_dirty:
                self.sdf.set_latents(self.latent)
                self._latent_dirty = False
            self._ev[0].record()
            self._graph.replay()
            self._ev[1].record()
            self._timed = True
            return self.fail_count
        self.shift()
        self._rti_iteration(timed=not torch.cuda.is_current_stream_capturing())    # an outer graph may be capturing this step
        _lib.check(self.lib.sdfnmpc_fail_count_update(_ptr(self.status), _ptr(self.fail_count), self.B, _stream()), 'fail_count_update')
        return self.fail_count

    def enable_graph(self):
        """Capture one control step (warm-start shift, linearise, SDF, QP, fail count) as a CUDA graph; ``solve()`` then
        replays it: one launch instead of four to six, which is what a single-instance caller pays for (``--workload
        single``).  Inputs live at static addresses (``_layout``); latents are folded outside the graph when they change."""
        if self.x0 is None:
            raise _lib.SdfNmpcError('set_x0 must be called before enable_graph')
        if self._latent_dirty:
            self.sdf.set_latents(self.latent)
            self._latent_dirty = False
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self.shift()
                self._rti_iteration(timed=False)
                _lib.check(self.lib.sdfnmpc_fail_count_update(_ptr(self.status), _ptr(self.fail_count), self.B, _stream()), 'fail_count_update')
        torch.cuda.current_stream().wait_stream(side)
        self._graph = g

    def _rti_iteration(self, timed=True):
        """ocp.py:159-170 without the warm-start shift and the fail-count bookkeeping (both belong to the control step)."""
        lat = _ptr(self.latent) if self._latent_dirty else C.c_void_p(0)
        if timed:
            self._ev[0].record()
        _lib.check(self.lib.sdfnmpc_solver_rti_step(
            self.handle, self.B, _ptr(self.x0), _ptr(self.y), _ptr(self.W), _ptr(self.yN), _ptr(self.WN), _ptr(self.qd),
            _ptr(self.pose), _ptr(self.flag), lat, _ptr(self.u0), _ptr(self.status), _ptr(self.qp_iters), _stream()),
            'sdfnmpc_solver_rti_step')
        if timed:
            self._ev[1].record()
            self._timed = True
        self._latent_dirty = False

    def solve_host(self, x0, out=None, sync=True):
        """End-to-end step with HOST buffers: pinned H2D of the state feedback, solve, D2H of the controls.  With
        ``sync=False`` the copies are only enqueued: the returned pinned arrays are valid after the caller synchronises the
        stream (lets a caller append its own device work and read everything back with ONE synchronisation)."""
        pin = self._pin.get('x0')
        if pin is None:
            pin = self._pin['x0'] = torch.zeros(self.B, self.nxd, dtype=torch.float64).pin_memory()
            self._pin['u0'] = torch.empty(self.B, NU, dtype=torch.float64).pin_memory()
            self._pin['st'] = torch.empty(self.B, dtype=torch.int32).pin_memory()
        pin[:, :self.nx].copy_(torch.as_tensor(x0, dtype=torch.float64).reshape(self.B, -1)[:, :self.nx])
        self._x0.copy_(pin, non_blocking=True)
        self.set_x0(self._x0)
        self.solve()
        self._pin['u0'].copy_(self.u0, non_blocking=True)
        self._pin['st'].copy_(self.status, non_blocking=True)
        if sync:
            torch.cuda.current_stream().synchronize()
        return self._pin['u0'].numpy(), self._pin['st'].numpy()

    # ------------------------------------------------------------------ getters ----------
    def _state_ptrs(self):
        ptrs = [C.c_void_p() for _ in range(4)]
        _lib.check(self.lib.sdfnmpc_solver_state(self.handle, *[C.byref(p) for p in ptrs]), 'solver_state')
        return ptrs

    def _view(self, ptr, shape, dtype, copy=True):
        # zero-copy torch view of library-owned device memory, then a stream-ordered copy
        class _Dev:
            pass
        d = _Dev()
        d.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr='<f8' if dtype == torch.float64 else '<f4',
                                          data=(int(ptr.value), False), version=2)
        t = torch.as_tensor(d, device=self.device)
        return t.clone() if copy else t

    def get_matrices(self):
        """controller.py:84-92 -> (x [B,N+1,10], u [B,N,4]) device tensors (copies)."""
        xb, ub, _, _ = self._state_ptrs()
        x = self._view(xb, (self.B, self.NI + 1, self.nxd), torch.float64)[:, :self.N + 1]
        return (x[:, :, :self.nx] if self.nxd != self.nx else x), self._view(ub, (self.B, self.NI, NU), torch.float64)[:, :self.N]

    def set_matrices(self, x, u):
        """Overwrite the SQP iterate (x [B,N+1,10], u [B,N,4]) - the batched form of the per-node ``solver.set(k, 'x'|'u', .)``
        calls of ``ocp.py:144-149``; used to warm-start from an external trajectory."""
        xb, ub, _, _ = self._state_ptrs()
        self._view(xb, (self.B, self.NI + 1, self.nxd), torch.float64, copy=False)[:, :self.N + 1, :self.nx].copy_(self._t(x, (self.B, self.N + 1, self.nx)))
        self._view(ub, (self.B, self.NI, NU), torch.float64, copy=False)[:, :self.N].copy_(self._t(u, (self.B, self.N, NU)))
        if self.x0 is None:
            self._x0[:, :self.nx] = self._t(x, (self.B, self.N + 1, self.nx))[:, 0]
            self.x0 = self._x0

    def get_multipliers(self):
        """QP multipliers of the last step: lam_hard [B,N,8] (lbu, ubu), lam_soft [B,N+1,6] (lh, uh).  With lifted terminal rows
        (flags.recursive_feasibility) lam_soft has N + 1 + lift nodes: node N + 1 = [hfov_E, vfov_E, rec_feas], node N + 2 = velocity box."""
        _, _, lh, ls = self._state_ptrs()
        return (self._view(lh, (self.B, self.NI, 2 * NU), torch.float64)[:, :self.N], self._view(ls, (self.B, self.NI + 1, 2 * NH), torch.float64))

    def get_diag(self):
        ptrs = [C.c_void_p() for _ in range(3)]
        _lib.check(self.lib.sdfnmpc_solver_diag(self.handle, *[C.byref(p) for p in ptrs]), 'solver_diag')
        return (self._view(ptrs[0], (self.B, self.NI + 1, NH), torch.float64), self._view(ptrs[1], (self.B, self.NI + 1), torch.float32)[:, :self.N + 1],
                self._view(ptrs[2], (self.B, self.NI + 1, 3), torch.float32)[:, :self.N + 1])

    def get_u(self):
        return self.u0

    def get_cmd_TRPYr(self):
        """controller.py:106-108: [gamma*mass, roll, pitch, wz] clipped."""
        cmd = torch.empty(self.B, 4, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sdfnmpc_cmd_trpyr(self.handle, self.B, _ptr(self.u0), _ptr(cmd), _stream()), 'cmd_trpyr')
        return cmd

    def get_cmd_acc(self):
        """controller.py:101-103: [W_R_B^T W_a, wz] clipped to the acceleration limits."""
        cmd = torch.empty(self.B, 4, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sdfnmpc_cmd_acc(self.handle, self.B, _ptr(self.x0), _ptr(self.u0), _ptr(cmd), _stream()), 'cmd_acc')
        return cmd

    def get_cmd_props(self):
        """controller.py:111-113: rotor speeds u * limits.wp clipped to [0, limits.wp] (model 'props' only)."""
        if self.model != 'props':
            raise _lib.SdfNmpcError(f"get_cmd_props: model '{self.model}' has no u_to_props (quad_props.py:52)")
        cmd = torch.empty(self.B, 4, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sdfnmpc_cmd_props(self.handle, self.B, _ptr(self.u0), _ptr(cmd), _stream()), 'cmd_props')
        return cmd

    def get_openloop_traj(self):
        """controller.py:116-122: predicted (position, quaternion) per node, node 0 = x0."""
        x, _ = self.get_matrices()
        x[:, 0] = self.x0[:, :self.nx]
        return x[:, :, 0:3], x[:, :, 3:7]

    def eval(self, k):
        """controller.py:125-130: SDF value at node k of the current solution, ignoring the flag."""
        xb = self._view(self._state_ptrs()[0], (self.B, self.NI + 1, self.nxd), torch.float64, copy=False)   # node k read in place
        R = self.pose[:, 3:12].reshape(-1, 3, 3)
        c0 = torch.einsum('bji,bj->bi', R, xb[:, k, 0:3] - self.pose[:, 0:3]).to(torch.float32)
        if self._latent_dirty:
            self.sdf.set_latents(self.latent)
        d, _ = self.sdf.eval(c0, pts_per_inst=1, want_grad=False)
        return d

    def get_t(self):
        """``Ocp.get_t`` (``ocp.py:170,177-178``): time of the last solve in seconds - here the device time of the last
        SQP-RTI iteration (linearise + SDF + QP), from CUDA events on the launching stream."""
        if not self._timed:
            return 0.0
        self._ev[1].synchronize()
        return 1e-3 * self._ev[0].elapsed_time(self._ev[1])

    def get_qp_residual(self):
        """[B,2]: verified KKT residual of each instance's last QP, inf-norm of the SQP step it produced."""
        p = C.c_void_p()
        _lib.check(self.lib.sdfnmpc_solver_qp_residual(self.handle, C.byref(p)), 'qp_residual')
        return self._view(p, (self.B, 2), torch.float64)

    def solve_sqp(self, tol=1e-6, max_iter=100):
        """Full SQP (BASELINE configs[4]): repeat the RTI iteration until the full step [du; dx] of an instance is below
        `tol` (inf-norm) or `max_iter` iterations.  Instances are independent, so all of them keep iterating on the
        device; iteration count and controls of each are recorded at ITS convergence.  Returns (u0 [B,4], sqp_iters [B],
        converged [B])."""
        done = torch.zeros(self.B, dtype=torch.bool, device=self.device)
        iters = torch.zeros(self.B, dtype=torch.int64, device=self.device)
        u_conv = torch.zeros(self.B, NU, dtype=torch.float64, device=self.device)
        if self.x0 is None:
            raise _lib.SdfNmpcError('set_x0 must be called before solve_sqp')
        self.shift()                       # once per control step, not once per SQP iteration
        for _ in range(max_iter):
            self._rti_iteration()
            step = self.get_qp_residual()[:, 1]
            iters += (~done).to(torch.int64)
            newly = (~done) & (step <= tol)
            u_conv = torch.where(newly[:, None], self.u0, u_conv)
            done |= newly
            if bool(done.all()):
                break
        u_conv = torch.where(done[:, None], u_conv, self.u0)
        _lib.check(self.lib.sdfnmpc_fail_count_update(_ptr(self.status), _ptr(self.fail_count), self.B, _stream()), 'fail_count_update')
        return u_conv, iters, done

    def timing_begin(self, max_steps):
        """Per-kernel CUDA-event timing of the next ``max_steps`` solves (bench.py roofline)."""
        _lib.check(self.lib.sdfnmpc_solver_timing_begin(self.handle, int(max_steps)), 'timing_begin')

    def timing_end(self):
        ms = (C.c_float * 3)()
        steps = C.c_int32()
        _lib.check(self.lib.sdfnmpc_solver_timing_end(self.handle, ms, C.byref(steps)), 'timing_end')
        return dict(linearize=ms[0], sdf=ms[1], qp=ms[2]), int(steps.value)

    def plant_step(self, x, u, dt):
        """x <- RK4(x, u, dt) with the controller's own model (closed-loop rollouts)."""
        if x.shape[1] != self.nxd:             # 'props': the device state carries one pad entry
            xd = torch.zeros(x.shape[0], self.nxd, dtype=torch.float64, device=self.device)
            xd[:, :self.nx] = x
            _lib.check(self.lib.sdfnmpc_plant_step(self.handle, x.shape[0], _ptr(xd), _ptr(u), float(dt), _stream()), 'plant_step')
            x.copy_(xd[:, :self.nx])
            return x
        _lib.check(self.lib.sdfnmpc_plant_step(self.handle, x.shape[0], _ptr(x), _ptr(u), float(dt), _stream()), 'plant_step')
        return x


class Nmpc:
    """Single-instance facade with the reference's signatures (``sdf_nmpc/controller.py``).

    Like the reference, the setters only write host (numpy) arrays; ``solve()`` pushes everything at once
    (``ocp.py:159-168`` pushes yref / W / p node by node into acados): here ONE pinned host-to-device copy of the input block,
    one CUDA-graph replay of the control step and one read-back of (u0, status, fail count)."""

    def __init__(self, cfg, rebuild=False, sdf_net=None, device='cuda:0', use_graph=True):
        """``Nmpc(cfg, rebuild=False)`` as in ``controller.py:9``: the SDF comes from ``cfg.nn.sdf_weights`` in the data
        directory (``gen_model.py:32``); ``rebuild`` is accepted and ignored (nothing is code-generated here).  A packed
        network may be given instead, as ``sdf_net=`` or in second position."""
        if isinstance(rebuild, dict):
            sdf_net, rebuild = rebuild, False
        self.cfg = cfg if isinstance(cfg, Config) else Config(cfg)
        self._b = BatchedNmpc(self.cfg, sdf_net, 1, device)
        self.N, self.T = self._b.N, self._b.T
        self._use_graph = bool(use_graph)
        lay, total = self._b._layout()
        self._lay = lay
        self._h = torch.zeros(total, dtype=torch.float64).pin_memory()
        hv = self._h.numpy()
        self._hv = {k: hv[o:o + n].reshape(shp) for k, (o, n, shp) in lay.items()}
        self._u0 = torch.zeros(NU, dtype=torch.float64).pin_memory()
        self._st = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._fc = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.cmd_acc_hover, self.cmd_acc_min, self.cmd_acc_max = self._b.cmd_acc_hover, self._b.cmd_acc_min, self._b.cmd_acc_max
        self.cmd_TRPYr_hover, self.cmd_TRPYr_min, self.cmd_TRPYr_max = self._b.cmd_TRPYr_hover, self._b.cmd_TRPYr_min, self._b.cmd_TRPYr_max
        self.cmd_props_hover, self.cmd_props_min, self.cmd_props_max = self._b.cmd_props_hover, self._b.cmd_props_min, self._b.cmd_props_max
        self.reset()

    def reset(self):
        self._b.reset()
        self._h.zero_()
        self.x0 = None
        self.fail_count = 0

    def set_sdf_flag(self, flag):
        self._hv['flag'][:] = float(flag)

    def set_latent(self, latent, W_p_Bo, W_R_Bo):
        R = np.asarray(W_R_Bo, float).reshape(3, 3)
        self._hv['pose'][0, 0:3] = R @ np.asarray(self.cfg.sensor.B_p_C, float) + np.asarray(W_p_Bo, float).reshape(3)
        self._hv['pose'][0, 3:12] = (R @ np.asarray(self.cfg.sensor.B_R_C, float)).reshape(9)
        self._b.latent[:] = torch.as_tensor(np.asarray(latent, np.float32).reshape(1, -1))
        self._b._latent_dirty = True

    def reset_latent(self):
        self._hv['pose'][:] = 0.0
        self._b.latent.zero_()
        self._b._latent_dirty = True

    def set_x0(self, x0):
        self.x0 = np.asarray(x0, float)[:self._b.nx]
        self._hv['x0'][0, :self._b.nx] = self.x0

    def set_ref(self, ref, k):
        y, W = self._b.formate_ref(ref)
        self._hv['qd'][0, k] = np.asarray(ref.q, float)
        if k < self.N:
            self._hv['y'][0, k], self._hv['W'][0, k] = y, W
        else:
            self._hv['yN'][0], self._hv['WN'][0] = y[:self._b.nyN], W[:self._b.nyN]

    def solve(self):
        if self.x0 is None:
            raise _lib.SdfNmpcError('set_x0 must be called before solve')
        b = self._b
        b._in.copy_(self._h, non_blocking=True)
        b.set_x0(b._x0)                      # first call: Ocp.init (ocp.py:144-149) from the uploaded state
        if self._use_graph and b._graph is None:
            b.solve()                        # one eager step first (module loading), then capture
            b.enable_graph()
        else:
            b.solve()
        self._u0.copy_(b.u0[0], non_blocking=True)
        self._st.copy_(b.status, non_blocking=True)
        self._fc.copy_(b.fail_count, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.fail_count = int(self._fc[0])
        if self.fail_count:
            print('solver failed: status', int(self._st[0]))
        return self.fail_count

    def get_u(self):
        return self._u0.numpy().copy()

    def get_t(self):
        return self._b.get_t()

    def get_matrices(self):
        x, u = self._b.get_matrices()
        return x[0].cpu().numpy(), u[0].cpu().numpy()

    def get_cmd_TRPYr(self):
        return self._b.get_cmd_TRPYr()[0].cpu().numpy()

    def get_cmd_acc(self):
        return self._b.get_cmd_acc()[0].cpu().numpy()

    def get_cmd_props(self):
        return self._b.get_cmd_props()[0].cpu().numpy()

    def get_openloop_traj(self):
        p, q = self._b.get_openloop_traj()
        p, q = p[0].cpu().numpy(), q[0].cpu().numpy()
        return [(p[k], q[k]) for k in range(self.N + 1)]

    def eval(self, k):
        return self._b.eval(k).cpu().numpy()
