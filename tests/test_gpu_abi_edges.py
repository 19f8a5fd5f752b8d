"""Edge cases of the C ABI on the device (empty / ragged inputs, maximum sizes, argument errors): invalid arguments
come back as SDFNMPC_E_INVALID / E_STATE instead of a launch, an empty SDF batch is a no-op, a batch smaller than the handle's
capacity solves exactly the instances it names, and the longest horizon the QP kernel maps onto a warp (N = 31) matches the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_problem

pytestmark = pytest.mark.gpu
E_INVALID, E_STATE = -1, -3


def _mpc(B, cuda_device, cfg=None):
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    return net, BatchedNmpc(cfg or Config(), net, B, cuda_device)


def test_invalid_arguments_are_refused(cuda_device):
    import torch
    from sdf_nmpc_b200.nmpc import _ptr
    net, mpc = _mpc(4, cuda_device)
    lib = mpc.lib
    # solver_create: horizon outside 1..31, empty batch, unknown model, terminal rows on another model / past the warp's 32 nodes
    base = mpc._desc
    clone = lambda: type(base).from_buffer_copy(base)
    for field, value in (('N', 0), ('N', 32), ('max_batch', 0), ('model', 7)):
        d = clone()
        setattr(d, field, value)
        h = C.c_void_p()
        assert lib.sdfnmpc_solver_create(C.byref(h), C.byref(d), mpc.sdf.handle) == E_INVALID, field
    d = clone(); d.term_lift, d.model = 1, 2
    assert lib.sdfnmpc_solver_create(C.byref(C.c_void_p()), C.byref(d), mpc.sdf.handle) == E_INVALID
    d = clone(); d.term_lift, d.N, d.brake_degree = 2, 30, 4
    d.fov_term[0] = 0.6
    assert lib.sdfnmpc_solver_create(C.byref(C.c_void_p()), C.byref(d), mpc.sdf.handle) == E_INVALID
    # rti_step: more instances than the handle holds, none at all, null handle
    args = [_ptr(mpc._x0), _ptr(mpc.y), _ptr(mpc.W), _ptr(mpc.yN), _ptr(mpc.WN), _ptr(mpc.qd), _ptr(mpc.pose), _ptr(mpc.flag), C.c_void_p(0),
            _ptr(mpc.u0), _ptr(mpc.status), _ptr(mpc.qp_iters), C.c_void_p(0)]
    assert lib.sdfnmpc_solver_rti_step(mpc.handle, 5, *args) == E_INVALID
    assert lib.sdfnmpc_solver_rti_step(mpc.handle, 0, *args) == E_INVALID
    assert lib.sdfnmpc_solver_rti_step(C.c_void_p(0), 4, *args) == E_INVALID
    torch.cuda.synchronize()


def test_empty_and_unprepared_sdf_batches(cuda_device):
    import torch
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.sdf import SdfEvaluator
    from sdf_nmpc_b200.nmpc import _ptr
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    ev = SdfEvaluator(net, cuda_device)
    pts = torch.zeros(8, 3, dtype=torch.float32, device=cuda_device)
    d = torch.full((8,), 7.0, dtype=torch.float32, device=cuda_device)
    g = torch.full((8, 3), 7.0, dtype=torch.float32, device=cuda_device)
    lib = ev.lib
    # before any latent: call order violated
    assert lib.sdfnmpc_sdf_eval(ev.handle, _ptr(pts), C.c_void_p(0), 8, 8, _ptr(d), _ptr(g), C.c_void_p(0)) == E_STATE
    ev.set_latents(torch.from_numpy(synth.latent(seed=2)).to(cuda_device))
    # an empty batch is a no-op: nothing launched, outputs untouched
    n0 = lib.sdfnmpc_launch_count()
    assert lib.sdfnmpc_sdf_eval(ev.handle, _ptr(pts), C.c_void_p(0), 8, 0, _ptr(d), _ptr(g), C.c_void_p(0)) == 0
    assert lib.sdfnmpc_launch_count() == n0
    torch.cuda.synchronize()
    assert float(d.min()) == 7.0 and float(g.min()) == 7.0
    # negative count, missing points, more instances than latents
    assert lib.sdfnmpc_sdf_eval(ev.handle, _ptr(pts), C.c_void_p(0), 8, -1, _ptr(d), _ptr(g), C.c_void_p(0)) == E_INVALID
    assert lib.sdfnmpc_sdf_eval(ev.handle, C.c_void_p(0), C.c_void_p(0), 8, 8, _ptr(d), _ptr(g), C.c_void_p(0)) == E_INVALID
    assert lib.sdfnmpc_sdf_eval(ev.handle, _ptr(pts), C.c_void_p(0), 4, 8, _ptr(d), _ptr(g), C.c_void_p(0)) == E_INVALID     # 2 instances, 1 latent


def test_partial_batch_solves_only_the_named_instances(cuda_device):
    """A handle built for 8 instances steps the first 5: those get the controls of a 5-instance solver, the rest are left alone."""
    import torch
    from oracle import rti
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc, _ptr
    B, Bp = 8, 5
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    ocfg = rti.default_cfg()
    prob, w = make_problem(B, ocfg)
    big, small = BatchedNmpc(Config(), net, B, cuda_device), BatchedNmpc(Config(), net, Bp, cuda_device)
    for m, n in ((big, B), (small, Bp)):
        m.set_sdf_flag(1.0)
        m.set_latent(w['latent'][:n], w['W_p_Bo'][:n], w['W_R_Bo'][:n])
        m.set_ref_arrays(w['yref'][:n], w['W'][:n], w['yrefN'][:n], w['WN'][:n], w['qd'][:n])
        m.set_x0(prob['x0'][:n])
    small.solve()
    big.sdf.set_latents(big.latent)
    big.u0.fill_(-9.0)
    big.status.fill_(77)
    args = [_ptr(big.x0), _ptr(big.y), _ptr(big.W), _ptr(big.yN), _ptr(big.WN), _ptr(big.qd), _ptr(big.pose), _ptr(big.flag), C.c_void_p(0),
            _ptr(big.u0), _ptr(big.status), _ptr(big.qp_iters), C.c_void_p(0)]
    assert big.lib.sdfnmpc_solver_rti_step(big.handle, Bp, *args) == 0
    torch.cuda.synchronize()
    assert torch.equal(big.u0[:Bp], small.u0) and int(big.status[:Bp].abs().sum()) == 0
    assert float(big.u0[Bp:].max()) == -9.0 and int(big.status[Bp:].min()) == 77


def test_maximum_horizon_matches_oracle(cuda_device):
    """N = 31: node k of an instance is lane k of its warp, so 32 nodes is the longest OCP the QP kernel takes."""
    from oracle import rti
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc
    B = 6
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = Config()
    cfg.mpc.N = 31
    cfg.mpc.T = 31 * 0.075
    ocfg = rti.default_cfg()
    ocfg['N'] = 31
    ocfg['dt'] = np.asarray(cfg.shooting_dt(), float)
    prob, w = make_problem(B, ocfg)
    mpc = BatchedNmpc(cfg, net, B, cuda_device)
    mpc.set_sdf_flag(1.0)
    mpc.set_latent(w['latent'], w['W_p_Bo'], w['W_R_Bo'])
    mpc.set_ref_arrays(w['yref'], w['W'], w['yrefN'], w['WN'], w['qd'])
    xb, ub = rti.init_traj(prob['x0'], ocfg)
    for step in range(2):
        mpc.set_x0(prob['x0'])
        assert int(mpc.solve().max().item()) == 0
        xb, ub, out = rti.rti_step(net, xb, ub, prob, ocfg)
        assert np.array_equal(mpc.status.cpu().numpy(), out['status'])
        assert np.abs(mpc.get_u().cpu().numpy() - out['u0']).max() <= 1e-4, step
    cfg.mpc.N = 32
    with pytest.raises(Exception):
        BatchedNmpc(cfg, net, B, cuda_device)
