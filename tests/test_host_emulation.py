"""CPU-side check of the DEVICE algorithms: tests/host_emul/emul.cpp compiles the very headers the CUDA kernels
include (linearize_core.cuh, qp_core.cuh, model_att.cuh, dual.cuh) with one lane per warp, and the result is
compared with the oracle.  This is a development aid for a container without a GPU; the product never loads it."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, make_problem


@pytest.fixture(scope='module')
def emul():
    out = os.path.join(tempfile.mkdtemp(prefix='sdfnmpc_emul_'), 'libhost_emul.so')
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-DSDFNMPC_HOST_EMUL', '-x', 'c++',
                    os.path.join(ROOT, 'tests', 'host_emul', 'emul.cpp'), '-o', out], check=True)
    return ctypes.CDLL(out)


def _P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _emul_step(lib, net, cfg, prob, xe, ue):
    from oracle import rti, sdf as SDF
    B, N = xe.shape[0], cfg['N']
    pose = np.ascontiguousarray(np.concatenate([prob['W_p_Co'], prob['W_R_Co'].reshape(B, 9)], 1))
    from oracle import models
    mid, nyN = models.MODEL_ID[cfg['model']], models.get(cfg['model']).nyN
    mcv = np.array([cfg['g'], cfg['gamma'], cfg['roll'], cfg['pitch'], cfg['wz'], cfg['mass'],
                    *(cfg['B_R_C'].T @ cfg['B_p_C'] + np.array([cfg['fov_off'], 0, 0])),
                    cfg['ax'], cfg['ay'], cfg['az'], cfg['tau_roll'], cfg['tau_pitch']])
    props = cfg['model'] == 'props'
    nx, ny = (14, 16) if props else (10, 11)          # device state of 'props': 13 + one pad entry (model_att.cuh)
    nw = nx + 4
    if props:
        mcv = np.concatenate([mcv, [cfg['wp_max']], (cfg['Gf'] / cfg['mass']).ravel(),
                              (cfg['Gt'] / cfg['inertia'][:, None]).ravel(), cfg['inertia']])
    lh, uh = rti.h_bounds(cfg, net['max_df'])
    zs, Zs = rti.slack_weights(cfg)
    AB = np.zeros((B, N, nx, nw)); bd = np.zeros((B, N, nx)); J = np.zeros((B, N + 1, ny, nw)); yr = np.zeros((B, N + 1, ny))
    C = np.zeros((B, N + 1, 3, nw)); hv = np.zeros((B, N + 1, 3)); qp = np.zeros((B, N + 1, 3), np.float32)
    c = np.ascontiguousarray
    yref, yrefN, qd, flag, W, WN, x0 = (c(prob[k]) for k in ('yref', 'yrefN', 'qd', 'flag', 'W', 'WN', 'x0'))
    if props:
        x0 = c(np.pad(x0, ((0, 0), (0, 1))))
    lib.emul_linearize(mid, nyN, B, N, _P(mcv), _P(xe), _P(ue), _P(yref), _P(yrefN), _P(qd), _P(pose), _P(flag), _P(cfg['dt']),
                       _P(AB), _P(bd), _P(J), _P(yr), _P(C), _P(hv), _P(qp), _P(x0))
    lin = dict(AB=AB.copy(), b=bd.copy(), J=J.copy(), hv=hv.copy(), C=C.copy(), qp=qp.copy(), yr=yr.copy())
    d, g = SDF.value_and_grad(net, qp.reshape(-1, 3), np.repeat(prob['latent'], N + 1, 0), dtype=np.float32)
    d, g = c(d.astype(np.float32)), c(g.astype(np.float32))
    u0 = np.zeros((B, 4)); lamh = np.zeros((B, N, 8)); lams = np.zeros((B, N + 1, 6))
    st = np.zeros(B, np.int32); its = np.zeros(B, np.int32)
    D = ctypes.c_double
    (lib.emul_qp_props if props else lib.emul_qp)(B, N, 1, 100, nyN, D(cfg.get('sdf_cost_w', 0.0)), D(cfg['lm']), D(1e-8), D(1e-7), D(net['max_df']), _P(flag), _P(AB), _P(bd), _P(J), _P(yr), _P(C),
                _P(hv), _P(W), _P(WN), _P(cfg['dt']), _P(cfg['lbu']), _P(cfg['ubu']), _P(lh), _P(uh), _P(zs), _P(Zs), _P(d), _P(g),
                _P(pose), _P(x0), _P(xe), _P(ue), _P(u0), _P(lamh), _P(lams), _P(st), _P(its))
    return u0, st, its, lamh, lams, lin


@pytest.mark.parametrize('model', ['att', 'att_tau', 'acc'])
def test_linearize_and_qp_cores_match_oracle(emul, model):
    from oracle import rti
    from sdf_nmpc_b200 import synth
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg(model)
    B = 12
    prob, w = make_problem(B, cfg, model=model)
    xb, ub = rti.init_traj(prob['x0'], cfg)
    xe, ue = np.ascontiguousarray(xb.copy()), np.ascontiguousarray(ub.copy())
    for step in range(3):
        xo, uo, out = rti.rti_step(net, xb, ub, prob, cfg)
        u0, st, its, lamh, lams, lin = _emul_step(emul, net, cfg, prob, xe, ue)
        if step == 0:   # dual-number linearisation == the oracle's Jacobians (hand-derived for 'att')
            L = out['lin']
            assert np.abs(lin['AB'][..., 4:] - L['A']).max() < 1e-12 and np.abs(lin['AB'][..., :4] - L['B']).max() < 1e-12
            assert np.abs(lin['b'] - L['b']).max() < 1e-12
            assert np.abs(lin['hv'][:, 1:, :2] - L['hbar'][:, 1:, :2]).max() < 1e-12
            assert np.abs(lin['C'][:, 1:, :2, 4:] - L['C'][:, 1:, :2]).max() < 1e-10
        assert np.all(st == 0) and its.max() <= 30
        assert np.abs(u0 - out['u0']).max() <= 1e-5
        assert np.abs(xe - xo).max() <= 1e-4 and np.abs(ue - uo).max() <= 1e-4
        assert np.abs(lamh[:, :, :4] - out['lam_lbu']).max() < 1e-2 and np.abs(lams[:, 1:, :3] - out['lam_lh']).max() < 1e-2
        xb, ub = xo, uo


def test_props_cores_match_oracle(emul):
    """Model 'props' (quad_props.py): 13 states + one pad entry on the device, nu 4, ny 16 / 12 - linearisation by dual numbers and
    the scalar Riccati IPM instantiated for that shape against the numpy oracle over three RTI steps."""
    from oracle import rti
    from sdf_nmpc_b200 import synth
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg('props')
    B = 8
    prob, w = make_problem(B, cfg, model='props')
    xb, ub = rti.init_traj(prob['x0'], cfg)
    xe = np.ascontiguousarray(np.pad(xb, ((0, 0), (0, 0), (0, 1))))
    ue = np.ascontiguousarray(ub.copy())
    for step in range(3):
        xo, uo, out = rti.rti_step(net, xb, ub, prob, cfg)
        u0, st, its, lamh, lams, lin = _emul_step(emul, net, cfg, prob, xe, ue)
        if step == 0:
            L = out['lin']
            assert np.abs(lin['AB'][:, :, :13, 4:17] - L['A']).max() < 1e-11 and np.abs(lin['AB'][:, :, :13, :4] - L['B']).max() < 1e-11
            assert np.abs(lin['b'][:, :, :13] - L['b']).max() < 1e-11
            # the pad state: identity dynamics, no coupling, zero defect
            assert np.abs(lin['AB'][:, :, 13, :17]).max() == 0 and np.all(lin['AB'][:, :, 13, 17] == 1) and np.abs(lin['AB'][:, :, :13, 17]).max() == 0
            assert np.abs(lin['hv'][:, 1:, :2] - L['hbar'][:, 1:, :2]).max() < 1e-12
        assert np.all(st == 0) and its.max() <= 40
        assert np.abs(u0 - out['u0']).max() <= 1e-5
        assert np.abs(xe[:, :, :13] - xo).max() <= 1e-4 and np.abs(ue - uo).max() <= 1e-4 and np.abs(xe[:, :, 13]).max() == 0
        xb, ub = xo, uo


def lift_arrays(cfg, max_df, z_hard=1e4):
    """Per-node bounds / slack weights [N + L + 1, 3] of the lifted terminal rows (term_lift.cuh) as the host side builds them."""
    from oracle import rti
    t = cfg['term']
    N, L = cfg['N'], 2 if t['stability'] else 1
    lh, uh = rti.h_bounds(cfg, max_df)
    zs, Zs = rti.slack_weights(cfg)
    nlh, nuh, nzs, nZs = (np.tile(a, (N + L + 1, 1)) for a in (lh, uh, zs, Zs))
    nlh[N, 2], nuh[N, 2] = -50.0, 50.0                # no terminal sdf row next to the rec_feas row (gen_model.py:69-70)
    hl, vl = cfg['hfov'] * cfg['fov_ratio'], cfg['vfov'] * cfg['fov_ratio']
    sb = t['slack_brake']
    nlh[N + 1], nuh[N + 1] = [-hl, -vl, cfg['size_xy']], [hl, vl, max_df]
    nzs[N + 1], nZs[N + 1] = [z_hard, z_hard, sb[0] if sb else z_hard], [0, 0, sb[1] if sb else 0]
    if L == 2:
        nlh[N + 2], nuh[N + 2], nzs[N + 2], nZs[N + 2] = -np.array(t['vlim']), t['vlim'], z_hard, 0
    return nlh, nuh, nzs, nZs


def _emul_lift_step(lib, net, cfg, prob, xe, ue):
    from oracle import rti, sdf as SDF
    t = cfg['term']
    L = 2 if t['stability'] else 1
    B, N = xe.shape[0], cfg['N']
    NI, nyN, nx, ny, nw = N + L, 5 if t['stability'] else 4, 10, 11, 14
    c = np.ascontiguousarray
    pose = c(np.concatenate([prob['W_p_Co'], prob['W_R_Co'].reshape(B, 9)], 1))
    mcv = np.array([cfg['g'], cfg['gamma'], cfg['roll'], cfg['pitch'], cfg['wz'], cfg['mass'],
                    *(cfg['B_R_C'].T @ cfg['B_p_C'] + np.array([cfg['fov_off'], 0, 0])),
                    cfg['ax'], cfg['ay'], cfg['az'], cfg['tau_roll'], cfg['tau_pitch']])
    lh, uh = rti.h_bounds(cfg, net['max_df'])
    zs, Zs = rti.slack_weights(cfg)
    node = lift_arrays(cfg, net['max_df'])
    dti = c(np.concatenate([cfg['dt'], np.ones(L)]))
    AB = np.zeros((B, NI, nx, nw)); bd = np.zeros((B, NI, nx)); J = np.zeros((B, NI + 1, ny, nw)); yr = np.zeros((B, NI + 1, ny))
    C = np.zeros((B, NI + 1, 3, nw)); hv = np.zeros((B, NI + 1, 3)); qp = np.zeros((B, NI + 1, 3), np.float32)
    yref, yrefN, qd, flag, W, WN, x0 = (c(prob[k]) for k in ('yref', 'yrefN', 'qd', 'flag', 'W', 'WN', 'x0'))
    lib.emul_set_lift(N, int(t['stability']), *[_P(a) for a in node])
    lib.emul_linearize(0, nyN, B, NI, _P(mcv), _P(xe), _P(ue), _P(yref), _P(yrefN), _P(qd), _P(pose), _P(flag), _P(dti),
                       _P(AB), _P(bd), _P(J), _P(yr), _P(C), _P(hv), _P(qp), _P(x0))
    d, g = SDF.value_and_grad(net, qp.reshape(-1, 3), np.repeat(prob['latent'], NI + 1, 0), dtype=np.float32)
    d, g = c(d.astype(np.float32)), c(g.astype(np.float32))
    D = ctypes.c_double
    co = c(np.asarray(t['coeffs'], float))
    lib.emul_term_lift(B, NI, N, L, t['degree'], _P(co), len(co), _P(mcv), _P(pose), _P(flag), D(net['max_df']), _P(d), _P(g),
                       _P(xe), _P(AB), _P(bd), _P(C), _P(hv))
    u0 = np.zeros((B, 4)); lamh = np.zeros((B, NI, 8)); lams = np.zeros((B, NI + 1, 6))
    st = np.zeros(B, np.int32); its = np.zeros(B, np.int32)
    lib.emul_qp(B, NI, 1, 100, nyN, D(0.0), D(cfg['lm']), D(1e-8), D(1e-7), D(net['max_df']), _P(flag), _P(AB), _P(bd), _P(J), _P(yr), _P(C),
                _P(hv), _P(W), _P(WN), _P(dti), _P(cfg['lbu']), _P(cfg['ubu']), _P(lh), _P(uh), _P(zs), _P(Zs), _P(d), _P(g),
                _P(pose), _P(x0), _P(xe), _P(ue), _P(u0), _P(lamh), _P(lams), _P(st), _P(its))
    lib.emul_set_lift(-1, 0, None, None, None, None)
    return u0, st, its, lams, dict(AB=AB, hv=hv)


def term_problem(cfg, prob, stability, slack_brake):
    """cfg / prob of the terminal-ingredient variants (flags.recursive_feasibility, + flags.stability)."""
    from sdf_nmpc_b200 import synth
    c, p = dict(cfg), dict(prob)
    c['term'] = dict(degree=4, coeffs=synth.braking_coeffs(4), slack_brake=slack_brake, stability=stability, vlim=[1.5, 1.5, 1.0], vfov=True)
    if stability:        # fifth terminal output flag * v.v (gen_model.py:139-147) with a weight and a reference of its own
        B = prob['x0'].shape[0]
        p['WN'] = np.concatenate([prob['WN'], np.full((B, 1), 3.0)], 1)
        p['yrefN'] = np.concatenate([prob['yrefN'], np.full((B, 1), 0.3)], 1)
    return c, p


@pytest.mark.parametrize('stability,slack_brake', [(False, (100.0, 10.0)), (True, (100.0, 10.0)), (True, None)])
def test_lifted_terminal_rows_match_oracle(emul, stability, slack_brake):
    """flags.recursive_feasibility / stability (gen_model.py:73-149): the device statement LIFTS the terminal rows onto auxiliary
    nodes and replaces hard rows by an exact L1 penalty (term_lift.cuh); the oracle adds them as full-width terminal rows, hard
    where the reference's are hard.  Same controls, trajectories and terminal multipliers over closed RTI steps."""
    from oracle import rti
    from sdf_nmpc_b200 import synth
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg('att')
    B = 8
    prob, w = make_problem(B, cfg)
    cfg, prob = term_problem(cfg, prob, stability, slack_brake)
    N, L = cfg['N'], 2 if stability else 1
    xb, ub = rti.init_traj(prob['x0'], cfg)
    xe = np.ascontiguousarray(np.concatenate([xb, np.repeat(xb[:, -1:], L, 1)], 1))
    ue = np.ascontiguousarray(np.concatenate([ub, np.repeat(ub[:, -1:], L, 1)], 1))
    active = 0.0
    for step in range(3):
        xo, uo, out = rti.rti_step(net, xb, ub, prob, cfg)
        u0, st, its, lams, lin = _emul_lift_step(emul, net, cfg, prob, xe, ue)
        hT, CT = out['lin']['term'][0], out['lin']['term'][1]
        if step == 0:       # lifted rows: node N + 1 carries [hfov_E, vfov_E, rec_feas]; the oracle orders them [rec_feas, hfov_E, vfov_E]
            assert np.abs(lin['hv'][:, N + 1, [2, 0, 1]] - hT[:, :3]).max() < 1e-6          # float32 network value inside
            assert np.abs(lin['AB'][:, N, [2, 0, 1], 4:] - CT[:, :3]).max() < 1e-5
        assert np.all(st == 0) and np.all(out['status'] == 0) and its.max() <= 40
        assert np.abs(u0 - out['u0']).max() <= 5e-5          # each side continues from its own iterate
        assert np.abs(xe[:, :N + 1] - xo).max() <= 1e-4 and np.abs(ue[:, :N] - uo).max() <= 1e-4
        assert np.abs(lams[:, N + 1, [2, 0, 1]] - out['lam_term_l'][:, :3]).max() < 1e-2
        assert np.abs(lams[:, N + 1, [5, 3, 4]] - out['lam_term_u'][:, :3]).max() < 1e-2
        active = max(active, out['lam_term_l'].max(), out['lam_term_u'].max())
        xb, ub = xo, uo
    assert active > 1.0          # the terminal rows do bind in this sample


def test_qp_core_robust_over_many_instances(emul):
    """Riccati IPM converges (status 0) on a few hundred random instances over repeated RTI steps, including the
    ones that made plain Mehrotra cycle before the stall safeguard."""
    from oracle import rti
    from sdf_nmpc_b200 import synth
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg()
    B = 192
    prob, w = make_problem(B, cfg, seed=1000 + 3200)      # block containing instance 3292
    xe, ue = rti.init_traj(prob['x0'], cfg)
    xe, ue = np.ascontiguousarray(xe), np.ascontiguousarray(ue)
    for step in range(12):
        u0, st, its, *_ = _emul_step(emul, net, cfg, prob, xe, ue)
        assert np.all(st == 0), (step, np.where(st != 0)[0])
        assert its.max() <= 40


def test_cxx_port_matches_numpy_oracle():
    """oracle/cxx_port.CxxRti (the CPU arm bench.py times: C++ linearisation + IPM, libtorch batch-1 SDF) gives the numpy
    oracle's controls on the same closed-loop steps."""
    import subprocess as sp
    from oracle import cxx_port, rti, models
    from sdf_nmpc_b200 import synth
    if not os.path.exists(cxx_port.LIB):
        os.makedirs(os.path.dirname(cxx_port.LIB), exist_ok=True)
        sp.run(['g++', '-O3', '-std=c++17', '-shared', '-fPIC', '-DSDFNMPC_HOST_EMUL', '-x', 'c++',
                os.path.join(ROOT, 'tests', 'host_emul', 'emul.cpp'), '-o', cxx_port.LIB], check=True)
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = rti.default_cfg()
    B = 3
    prob, w = make_problem(B, cfg)
    port = cxx_port.CxxRti(net, cfg, prob, sdf='torch')
    xb, ub = rti.init_traj(prob['x0'], cfg)
    x = prob['x0'].copy()
    for step in range(2):
        p = dict(prob); p['x0'] = x
        xb, ub, out = rti.rti_step(net, xb, ub, p, cfg)
        u0, st, its = port.step(x)
        assert np.all(st == 0) and np.abs(u0 - out['u0']).max() <= 2e-5
        x = models.get('att').rk4(x, out['u0'], cfg['dt'][0], cfg)
