"""The drop-in boundary with the reference's signatures and call sequence:
``Nmpc(cfg, rebuild=False)`` and ``VaeWrapper(cfg)`` load TorchScript archives named by ``cfg.nn.*_weights`` from the data
directory (``controller.py:9-16``, ``gen_model.py:32``, ``vae.py:8-13``).  The LFS files are absent offline, so stand-in
archives with the reference's parameter names are written to a temporary data directory (tests/refstub.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class Ref:
    def __init__(self, p, q, v, wz):
        self.p, self.q, self.v, self.wz = p, q, v, wz
        self.Wp, self.Wv, self.Wq, self.Ww, self.Wa = [0, 0, 5], [3, 3, 3], [50, 50, 10], [0, 0, 3], 0.05


@pytest.fixture(scope='module')
def data_dir(tmp_path_factory):
    import torch
    import sdf_nmpc_b200
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    import refstub
    d = tmp_path_factory.mktemp('data')
    cfg = Config()
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    enc = synth.make_encoder_net(seed=5)
    torch.jit.script(refstub.NeuralDFStub(net).eval()).save(str(d / cfg.nn.sdf_weights))
    torch.jit.script(refstub.VaeStub(enc, synth.make_decoder_net(seed=7)).eval()).save(str(d / cfg.nn.vae_weights))
    sdf_nmpc_b200.set_data_dir(d)
    yield d, net, enc
    sdf_nmpc_b200.set_data_dir(None)


def test_reference_call_sequence_runs_verbatim(data_dir, cuda_device):
    """the reference's control loop with only the import changed."""
    from oracle import rti
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import Nmpc
    from sdf_nmpc_b200.vae import VaeWrapper
    _, net, enc = data_dir
    cfg = Config()
    vae = VaeWrapper(cfg)                      # vae.py:8
    mpc = Nmpc(cfg, rebuild=False)             # controller.py:9
    img = synth.depth_frame(seed=1)
    vae.set_img(img)
    latent = vae.encode()                      # [1, 128]
    assert latent.shape == (1, 128)
    decoded = vae.decode()                     # vae.py:42-45: [H, W] in [0, 1]
    assert decoded.shape == (270, 480) and 0 <= decoded.min() and decoded.max() <= 1
    x0 = np.array([0, 0, 2, 1, 0, 0, 0, 0, 0, 0], float)
    mpc.set_sdf_flag(True)
    mpc.set_latent(latent, x0[:3], np.eye(3))
    vref, dt = np.array([3.0, 0.5, 0.0]), 0.075
    us = []
    for step in range(3):                      # step 0 eager, step 1 captures the graph, step 2 replays it
        mpc.set_x0(x0)
        for k in range(mpc.N + 1):
            mpc.set_ref(Ref(x0[:3] + vref * k * dt, np.array([1.0, 0, 0, 0]), vref, 0.0), k)
        assert mpc.solve() == 0
        us.append(mpc.get_u())
        assert mpc.get_t() > 0
    x, u = mpc.get_matrices()
    assert x.shape == (21, 10) and u.shape == (20, 4) and np.allclose(u[0], us[-1])
    cmd = mpc.get_cmd_TRPYr()
    assert cmd.shape == (4,) and abs(cmd[0] - min(20.0, us[-1][0] * 20.0 * cfg.robot.mass)) < 1e-9
    assert mpc.get_cmd_acc().shape == (4,) and len(mpc.get_openloop_traj()) == 21
    assert np.isfinite(mpc.eval(3)).all()
    # the loaded weights are the synthetic ones: same controls as the oracle on the same problem
    ocfg = rti.default_cfg()
    prob = dict(x0=x0[None], flag=np.ones(1), latent=latent.astype(np.float32),
                W_p_Co=(np.eye(3) @ ocfg['B_p_C'] + x0[:3])[None], W_R_Co=(np.eye(3) @ ocfg['B_R_C'])[None])
    yref = np.zeros((1, 20, 11)); W = np.zeros((1, 20, 11)); qd = np.zeros((1, 21, 4)); qd[..., 0] = 1
    for k in range(20):
        yref[0, k, 0:3] = x0[:3] + vref * k * dt; yref[0, k, 4:7] = vref
        W[0, k] = [0, 0, 5, 10, 3, 3, 3, 50, 50, 3, 0.05]
    prob.update(yref=yref, W=W, yrefN=np.concatenate([x0[:3] + vref * 20 * dt, [0]])[None], WN=np.array([[0, 0, 5, 10.0]]), qd=qd)
    xb, ub = rti.init_traj(prob['x0'], ocfg)
    for step in range(3):
        xb, ub, out = rti.rti_step(net, xb, ub, prob, ocfg)
        assert np.abs(out['u0'][0] - us[step]).max() <= 1e-4, (step, out['u0'][0], us[step])


def test_encoder_from_archive_matches_packed_weights(data_dir, cuda_device):
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.vae import VaeWrapper
    _, net, enc = data_dir
    img = synth.depth_frame(seed=2)
    a = VaeWrapper(Config()); a.set_img(img)
    b = VaeWrapper(Config(), enc); b.set_img(img)
    assert np.array_equal(a.encode(), b.encode())


def test_terminal_flags_follow_the_reference(cuda_device):
    """gen_model.py:73-149: `stability` only acts inside the recursive-feasibility branch (alone it leaves the plain OCP), the branch
    is stated for the 'att' layout only (gen_model.py:74 asserts), and it needs the braking-distance coefficients."""
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B)
    cfg = Config()
    cfg.flags['stability'] = True
    mpc = BatchedNmpc(cfg, net, 2, cuda_device)
    assert mpc.lift == 0 and mpc.nyN == 4 and mpc.NI == mpc.N
    cfg.flags['recursive_feasibility'] = True
    cfg.mpc.braking_dist = dict(degree=4, coeff_file='does/not/exist.npy')
    with pytest.raises(FileNotFoundError):
        BatchedNmpc(cfg, net, 2, cuda_device)
    cfg.mpc.braking_dist = dict(degree=4, coeffs=[0.0] * 34)       # degree 4 has 35 terms
    with pytest.raises(ValueError):
        BatchedNmpc(cfg, net, 2, cuda_device)
    cfg.mpc.model = 'att_tau'
    with pytest.raises(NotImplementedError):
        BatchedNmpc(cfg, net, 2, cuda_device)


def test_single_instance_facade_with_terminal_flags(cuda_device):
    """The reference's call sequence with flags.recursive_feasibility + stability on: ``set_ref`` fills the five terminal outputs /
    weights with ``y[:nyN]`` / ``W[:nyN]`` (controller.py:140-142), the CUDA-graph replay includes the term-lift kernel, and the
    facade gives the controls of the batched solver fed with the same arrays."""
    import torch
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import BatchedNmpc, Nmpc
    net = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True)
    cfg = Config()
    cfg.flags.recursive_feasibility = True
    cfg.flags.stability = True
    cfg.mpc.braking_dist = dict(degree=4, coeffs=[float(c) for c in synth.braking_coeffs(4)])
    cfg.mpc.weights.slack_brake = [100.0, 10.0]
    mpc = Nmpc(cfg, sdf_net=net, device=cuda_device)
    ref = BatchedNmpc(cfg, net, 1, cuda_device)
    assert mpc._b.nyN == 5 and mpc._b.NI == mpc.N + 2
    latent = synth.latent(seed=2)[None]
    x0 = np.array([0, 0, 2, 1, 0, 0, 0, 1.5, 0.2, 0], float)
    vref, dt = np.array([3.0, 0.5, 0.0]), 0.075
    for m in (mpc, ref):
        m.set_sdf_flag(True)
    mpc.set_latent(latent, x0[:3], np.eye(3))
    ref.set_latent(latent, x0[None, :3], np.eye(3)[None])
    N = mpc.N
    yref = np.zeros((1, N, 11)); W = np.zeros((1, N, 11)); qd = np.zeros((1, N + 1, 4)); qd[..., 0] = 1
    for k in range(N):
        yref[0, k, 0:3] = x0[:3] + vref * k * dt; yref[0, k, 4:7] = vref
        W[0, k] = [0, 0, 5, 10, 3, 3, 3, 50, 50, 3, 0.05]
    # terminal: y[:5] = [p, q_e3, v_x] and W[:5] = [Wp, Wq_e3, Wv_x] - what controller.py:140-142 hands to acados
    yN = np.concatenate([x0[:3] + vref * N * dt, [0.0, vref[0]]])[None]
    WN = np.array([[0, 0, 5, 10.0, 3.0]])
    ref.set_ref_arrays(yref, W, yN, WN, qd)
    for step in range(3):                      # eager, graph capture, graph replay
        mpc.set_x0(x0)
        for k in range(N + 1):
            mpc.set_ref(Ref(x0[:3] + vref * k * dt, np.array([1.0, 0, 0, 0]), vref, 0.0), k)
        assert mpc.solve() == 0
        ref.set_x0(x0[None])
        assert int(ref.solve().max().item()) == 0
        assert np.abs(mpc.get_u() - ref.get_u().cpu().numpy()[0]).max() <= 1e-9, step
    x, u = mpc.get_matrices()
    assert x.shape == (N + 1, 10) and u.shape == (N, 4)
    assert len(mpc.get_openloop_traj()) == N + 1
