"""Generate the committed golden fixtures from the REFERENCE's own code.

Run only where ``/root/reference`` exists (this container):
    python tests/golden/make_golden.py

It imports the reference's torch modules (``sdf_nmpc.network.neural_df``,
``sdf_nmpc.network.vae``, ``sdf_nmpc.utils.preprocessing``) and the numpy
branches of ``sdf_nmpc.utils.math`` (with a stub ``casadi`` module, which those
files import but do not need on the numeric path), loads the repo's seeded
synthetic weights into them and stores inputs' seeds + the reference outputs.
Weights are NOT stored: they are regenerated from the seed by
``sdf_nmpc_b200.synth`` (numpy PCG64, platform independent).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.modules.setdefault('casadi', types.ModuleType('casadi'))

from sdf_nmpc_b200 import synth  # noqa: E402


def ref_neural_df(net, dtype=torch.float32):
    from sdf_nmpc.network.neural_df import NeuralDF
    sizes = [net['W1'].shape[0], net['W2'].shape[0], net['W3'].shape[0], net['W4'].shape[0]]
    m = NeuralDF(signed=True, size_latent=net['size_latent'], nb_freqs=net['nb_freqs'], res=net['res'],
                 embed=net['embed'], act=net['act'], dropout_rate=0.1, w0=net['w0'], layer_sizes=sizes,
                 max_df=net['max_df'])
    sd = m.state_dict()
    sd['layers.main1.0.weight'] = torch.from_numpy(net['W1']); sd['layers.main1.0.bias'] = torch.from_numpy(net['b1'])
    sd['layers.main1.3.weight'] = torch.from_numpy(net['W2']); sd['layers.main1.3.bias'] = torch.from_numpy(net['b2'])
    sd['layers.main2.0.weight'] = torch.from_numpy(net['W3']); sd['layers.main2.0.bias'] = torch.from_numpy(net['b3'])
    sd['layers.main2.3.weight'] = torch.from_numpy(net['W4']); sd['layers.main2.3.bias'] = torch.from_numpy(net['b4'])
    sd['layers.df.0.weight'] = torch.from_numpy(net['W5']); sd['layers.df.0.bias'] = torch.from_numpy(net['b5'])
    m.load_state_dict(sd)
    # the direction table of the synthetic net must be the reference's own buffer
    assert np.array_equal(m.embed.dirs.numpy(), net['dirs']), 'direction table differs from the reference'
    return m.to(dtype).eval()


def sdf_case(name, n, **kw):
    net = synth.make_sdf_net(**kw)
    pts = synth.sdf_query_points(n, seed=3)
    rng = np.random.default_rng(11)
    n_lat = 4
    lat = rng.standard_normal((n_lat, net['size_latent'])).astype(np.float32)
    idx = (np.arange(n) * n_lat // n).astype(np.int32)
    x = np.concatenate([pts, lat[idx]], 1)
    out = {}
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        m = ref_neural_df(net, dt)
        xt = torch.from_numpy(x).to(dt).requires_grad_(True)
        d = m(xt)
        (g,) = torch.autograd.grad(d.sum(), xt)
        out['d_' + tag] = d.detach().numpy()[:, 0]
        out['g_' + tag] = g.detach().numpy()[:, :3]
    # what L4CasADi executes: batch-1 forward + jacrev (gen_model.py:38-39), first 8 points
    m32 = ref_neural_df(net, torch.float32)
    jac = torch.func.jacrev(lambda v: m32(v[None])[0])
    out['jacrev_f32'] = np.stack([jac(torch.from_numpy(x[i]))[0].detach().numpy() for i in range(8)])
    # scripted module keeps the attributes the packer reads (gen_model.py:32-35)
    ts = torch.jit.script(m32)
    assert float(ts.max_df) == net['max_df'] and ts.res == net['res']
    np.savez_compressed(os.path.join(HERE, f'sdf_{name}.npz'), n=n, n_lat=n_lat, lat_seed=11, pts_seed=3,
                        kw=np.array(repr(kw)), **out)
    print(name, 'd range', out['d_f32'].min(), out['d_f32'].max(),
          'f32-f64 val', np.abs(out['d_f32'] - out['d_f64']).max() / np.abs(out['d_f64']).max(),
          'grad', np.abs(out['g_f32'] - out['g_f64']).max() / np.abs(out['g_f64']).max())


def ref_encoder(net):
    from sdf_nmpc.network.vae import Vae
    vae = Vae(size_latent=net['size_latent'], shape_imgs=[270, 480], dropout_rate=0.1, batchnorm=True)
    enc = vae.encoder
    sd = enc.state_dict()
    t = torch.from_numpy
    sd['layers.resnet.0.weight'] = t(net['conv0_w']); sd['layers.resnet.0.bias'] = t(net['conv0_b'])
    for i, blk in enumerate(net['blocks']):
        p = f'layers.resnet.{3 + i}.'
        sd[p + 'layers.0.weight'] = t(blk['conv1_w'])
        sd[p + 'layers.3.weight'] = t(blk['conv2_w'])
        for j, key in ((1, 'bn1'), (4, 'bn2')):
            sd[p + f'layers.{j}.weight'] = t(blk[key]['weight']); sd[p + f'layers.{j}.bias'] = t(blk[key]['bias'])
            sd[p + f'layers.{j}.running_mean'] = t(blk[key]['mean']); sd[p + f'layers.{j}.running_var'] = t(blk[key]['var'])
        if 'sc_w' in blk:
            sd[p + 'shortcut.0.weight'] = t(blk['sc_w'])
            sd[p + 'shortcut.1.weight'] = t(blk['sc_bn']['weight']); sd[p + 'shortcut.1.bias'] = t(blk['sc_bn']['bias'])
            sd[p + 'shortcut.1.running_mean'] = t(blk['sc_bn']['mean']); sd[p + 'shortcut.1.running_var'] = t(blk['sc_bn']['var'])
    sd['layers.mean.weight'] = t(net['mean_w']); sd['layers.mean.bias'] = t(net['mean_b'])
    enc.load_state_dict(sd)
    return enc.eval()


def encoder_case():
    from sdf_nmpc.utils import preprocessing
    net = synth.make_encoder_net(seed=5)
    enc = ref_encoder(net)
    out = {}
    for tag, spherical, is_depth in (('depth', False, True), ('range', True, False)):
        frames = np.stack([synth.depth_frame(seed=1 + i, spherical=spherical) for i in range(2)])
        pre = [preprocessing.ToDevice('cpu'), preprocessing.Reshape([1, 270, 480]), preprocessing.ClipDistance(5, 1000)]
        if is_depth:
            pre.append(preprocessing.Depth2Range([1, 270, 480], 0.7592, 0.4903, 'cpu'))
        pipe = torch.nn.Sequential(*pre)
        lat, pres = [], []
        with torch.no_grad():
            for f in frames:
                img = pipe(f)                       # vae.py:31-32
                pres.append(img.numpy()[0, 0])
                lat.append(enc(img).numpy()[0])     # vae.py:37-40
            enc64 = enc.double()
            lat64 = [enc64(torch.from_numpy(p)[None, None].double()).numpy()[0] for p in pres]
            enc.float()
        out['latent_' + tag] = np.stack(lat)
        out['latent64_' + tag] = np.stack(lat64)
        pre_img = np.stack(pres)
        out['pre_rows_' + tag] = pre_img[:, ::45, ::40].copy()   # sparse sample of the preprocessed image
        out['pre_sum_' + tag] = pre_img.astype(np.float64).sum((1, 2))
        print(tag, 'latent range', out['latent_' + tag].min(), out['latent_' + tag].max(),
              'f32-f64', np.abs(out['latent_' + tag] - out['latent64_' + tag]).max() / np.abs(out['latent64_' + tag]).max())
    # a small stand-alone ResBlock case for fast CPU tests (stride 2 with shortcut, and stride 1)
    np.savez_compressed(os.path.join(HERE, 'encoder.npz'), seed=5, frame_seed0=1, **out)


def ref_decoder(net):
    from sdf_nmpc.network.vae import Vae
    dec = Vae(size_latent=net['size_latent'], shape_imgs=[270, 480], dropout_rate=0.1, batchnorm=True).decoder
    sd = dec.state_dict()
    t = torch.from_numpy
    sd['layers.resnet.0.weight'] = t(net['lin_w']); sd['layers.resnet.0.bias'] = t(net['lin_b'])
    for i, blk in enumerate(net['blocks']):
        p = f'layers.resnet.{4 + i}.'
        sd[p + 'layers.0.weight'] = t(blk['ct1_w']); sd[p + 'layers.3.weight'] = t(blk['ct2_w']); sd[p + 'shortcut.0.weight'] = t(blk['sc_w'])
        for q, key in ((p + 'layers.1', 'bn1'), (p + 'layers.4', 'bn2'), (p + 'shortcut.1', 'sc_bn')):
            sd[q + '.weight'] = t(blk[key]['weight']); sd[q + '.bias'] = t(blk[key]['bias'])
            sd[q + '.running_mean'] = t(blk[key]['mean']); sd[q + '.running_var'] = t(blk[key]['var'])
    sd['layers.resnet.8.weight'] = t(net['out_w']); sd['layers.resnet.8.bias'] = t(net['out_b'])
    dec.load_state_dict(sd)
    return dec.eval()


def decoder_case():
    """vae.py:42-45 on two seeded latents: a sparse sample of the decoded image (every 9th row, 8th column) and its sums,
    from the reference's Decoder in float32 and float64."""
    net = synth.make_decoder_net(seed=7)
    dec = ref_decoder(net)
    z = synth.latent(seed=11, n=2).astype(np.float32)
    with torch.no_grad():
        img32 = dec(torch.from_numpy(z)).numpy()[:, 0]
        img64 = dec.double()(torch.from_numpy(z).double()).numpy()[:, 0]
    np.savez_compressed(os.path.join(HERE, 'decoder.npz'), seed=7, latent_seed=11, rows32=img32[:, ::9, ::8].copy(),
                        rows64=img64[:, ::9, ::8].copy(), sum32=img32.astype(np.float64).sum((1, 2)), sum64=img64.sum((1, 2)))
    print('decoder: image range', img32.min(), img32.max(), 'f32-f64', np.abs(img32 - img64).max())


def model_att_case():
    """f_expl, y, yN, h values of model 'att' composed from the reference's numpy math helpers exactly as
    model/quad_rollpitchyawrate.py:20-55, cost_const_helpers.py:56-65 and gen_model.py:46-51 compose them."""
    from sdf_nmpc.utils import math as rm
    rng = np.random.default_rng(7)
    n = 64
    g, gam, rmax, pmax, wmax = 9.81, 20.0, 0.7, 0.7, 3.0
    X = np.concatenate([rng.uniform(-2, 2, (n, 3)), rng.standard_normal((n, 4)), rng.uniform(-3, 3, (n, 3))], 1)
    X[: n // 2, 3:7] /= np.linalg.norm(X[: n // 2, 3:7], axis=1, keepdims=True)   # half unit, half un-normalised
    U = np.concatenate([rng.uniform(0, 1, (n, 1)), rng.uniform(-1, 1, (n, 3))], 1)
    QD = rng.standard_normal((n, 4)); QD /= np.linalg.norm(QD, axis=1, keepdims=True)
    Pco = rng.uniform(-1, 1, (n, 3))
    eul = rng.uniform(-0.5, 0.5, (n, 3)); eul[:, 2] = rng.uniform(-np.pi, np.pi, n)
    B_p_C = np.array([0.18, 0, -0.025]); B_R_C = rm.euler2rot(np.zeros(3)); off = 0.05
    F, Y, Hfov, C0, RCO = [], [], [], [], []
    for i in range(n):
        x, u = X[i], U[i]
        p = x[:3]; q = x[3:7] / np.linalg.norm(x[3:7]); v = x[7:]
        th = np.arctan2(q[3], q[0]); qyaw = np.array([np.cos(th), 0, 0, np.sin(th)])
        gamma, roll, pitch, wz = u[0] * gam, u[1] * rmax, u[2] * pmax, u[3] * wmax
        V_R_B = rm.euler2rot(np.array([roll, pitch, 0.0])); W_R_V = rm.quat2rot(qyaw)
        W_a = W_R_V @ V_R_B @ np.array([0, 0, gamma]) + np.array([0, 0, -g])
        dq = rm.hamilton_prod(q, np.array([0, 0, 0, wz])) / 2
        F.append(np.concatenate([v, dq, W_a]))
        q_e = rm.hamilton_prod(QD[i], rm.invert(q))
        Y.append(np.concatenate([p, [q_e[3]], v, [roll, pitch, wz, W_a[2]]]))
        W_R_Co = rm.euler2rot(eul[i]); RCO.append(W_R_Co)
        c0 = W_R_Co.T @ (p - Pco[i]); C0.append(c0)
        c = c0 + B_R_C.T @ B_p_C + np.array([off, 0, 0])
        Hfov.append([np.arctan2(c[1], c[0]), np.arctan2(c[2], np.linalg.norm(c[:2]))])
    np.savez_compressed(os.path.join(HERE, 'model_att.npz'), X=X, U=U, QD=QD, Pco=Pco, Rco=np.array(RCO),
                        F=np.array(F), Y=np.array(Y), Hfov=np.array(Hfov), C0=np.array(C0),
                        vfov=rm.get_vfov(0.7592, 1.778, False))
    print('model_att: f range', np.abs(F).max())


def model_variants_case():
    """f_expl, y of models 'att_tau' and 'acc' composed from the reference's numpy math helpers exactly as
    model/quad_rollpitchyawrate_tau.py:22-52 and model/quad_acc.py:20-45 compose them (same X, U, QD as model_att.npz,
    attitudes kept away from the pitch singularity of the Euler map)."""
    from sdf_nmpc.utils import math as rm
    G = np.load(os.path.join(HERE, 'model_att.npz'))
    X, U, QD = G['X'].copy(), G['U'], G['QD']
    rng = np.random.default_rng(11)
    n = len(X)
    eul = np.stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.6, 0.6, n), rng.uniform(-np.pi, np.pi, n)], 1)
    scale = np.where(np.arange(n) < n // 2, 1.0, rng.uniform(0.5, 2.0, n))     # half unit, half un-normalised
    for i in range(n):
        cy, sy, cp, sp, cr, sr = (f(a / 2) for a in eul[i, ::-1] for f in (np.cos, np.sin))
        X[i, 3:7] = scale[i] * np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                                         cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])
    g, gam, rmax, pmax, wmax, ax, ay, az, tau = 9.81, 20.0, 0.7, 0.7, 3.0, 4.0, 4.0, 3.0, 0.12
    Ft, Yt, Fa, Ya = [], [], [], []
    cmd = dict(att=[], att_tau=[], acc=[])     # u_to_acc of each model: [W_R_B^T W_a, wz]
    for i in range(n):
        x, u = X[i], U[i]
        p = x[:3]; q = x[3:7] / np.linalg.norm(x[3:7]); v = x[7:]
        q_e = rm.hamilton_prod(QD[i], rm.invert(q))
        # att_tau
        eta = rm.quat2euler(q)
        gamma, roll_des, pitch_des, wz = u[0] * gam, u[1] * rmax, u[2] * pmax, u[3] * wmax
        W_R_B = rm.quat2rot(q)
        W_a = W_R_B @ np.array([0, 0, gamma]) + np.array([0, 0, -g])
        w = rm.deuler_avel_map(eta) @ np.array([(roll_des - eta[0]) / tau, (pitch_des - eta[1]) / tau, 0])
        dq = rm.hamilton_prod(q, np.array([0, w[0], w[1], wz])) / 2
        cmd['att_tau'].append(np.concatenate([W_R_B.T @ W_a, [wz]]))
        Ft.append(np.concatenate([v, dq, W_a]))
        Yt.append(np.concatenate([p, [q_e[3]], v, [roll_des, pitch_des, wz, W_a[2]]]))
        # acc
        W_a = np.array([u[0] * ax, u[1] * ay, u[2] * az])
        dq = rm.hamilton_prod(q, np.array([0, 0, 0, wz])) / 2
        cmd['acc'].append(np.concatenate([W_R_B.T @ W_a, [wz]]))
        Fa.append(np.concatenate([v, dq, W_a]))
        Ya.append(np.concatenate([p, [q_e[3]], v, W_a, [wz]]))
        # att (quad_rollpitchyawrate.py:27-45)
        th = np.arctan2(q[3], q[0]); qyaw = np.array([np.cos(th), 0, 0, np.sin(th)])
        R_att = rm.quat2rot(qyaw) @ rm.euler2rot(np.array([roll_des, pitch_des, 0.0]))
        W_a = R_att @ np.array([0, 0, gamma]) + np.array([0, 0, -g])
        cmd['att'].append(np.concatenate([R_att.T @ W_a, [wz]]))
    np.savez_compressed(os.path.join(HERE, 'model_variants.npz'), X=X, U=U, QD=QD, F_att_tau=np.array(Ft), Y_att_tau=np.array(Yt),
                        F_acc=np.array(Fa), Y_acc=np.array(Ya), **{'CMDACC_' + k: np.array(v) for k, v in cmd.items()})
    print('model_variants: f range', np.abs(Ft).max(), np.abs(Fa).max())


def model_props_case():
    """Model 'props': allocation matrices from the reference's GTMRP_matrix / axis_rot on the robot.alloc block of the reference's
    default.yaml, and f_expl, y, u_to_acc, u_to_props composed from its numpy math helpers exactly as model/quad_props.py:18-59
    composes them (same X[:, :10], U, QD as model_variants.npz; body rates drawn here; inputs mapped to [0, 1])."""
    from math import pi
    from sdf_nmpc.utils import math as rm
    from sdf_nmpc.utils.config import Config
    cfg = Config('/root/reference/sdf_nmpc/config/default.yaml')
    G = np.load(os.path.join(HERE, 'model_variants.npz'))
    rng = np.random.default_rng(31)
    n = len(G['X'])
    X = np.concatenate([G['X'], rng.uniform(-2, 2, (n, 3))], 1)
    U = 0.5 * (G['U'] + 1.0) * np.array([1.0, 0.5, 0.5, 0.5]) + np.array([0.0, 0.25, 0.25, 0.25])
    QD = G['QD']
    # quad_props.py:18-27
    px, py, pz, alpha, beta, sign = zip(*cfg.robot.alloc.motors)
    cf = [cfg.robot.alloc.cf] * len(sign)
    ct = [cfg.robot.alloc.ct] * len(sign)
    R = [rm.axis_rot('z', i * (pi / (len(sign) / 2))) @ rm.axis_rot('y', beta[i]) @ rm.axis_rot('x', (-1) ** i * alpha[i]) for i in range(len(sign))]
    p = np.array([px, py, pz]).T
    Gf, Gt = rm.GTMRP_matrix(R, p, sign, cf, ct)
    Gf = cf * Gf
    Gt = cf * Gt
    m, g = cfg.robot.mass, 9.81
    J = np.diag(cfg.robot.inertia); Jinv = np.linalg.inv(J)
    F, Y, CA, CP = [], [], [], []
    for i in range(n):
        x, u = X[i], U[i]
        p_ = x[0:3]; q = x[3:7] / np.linalg.norm(x[3:7]); eta = rm.quat2euler(q); v = x[7:10]; w = x[10:]
        wp = u * cfg.robot.limits.wp
        W_R_B = rm.quat2rot(q)
        W_a = W_R_B @ Gf @ wp ** 2 / m + np.array([0, 0, -g])
        F.append(np.concatenate([v, rm.hamilton_prod(q, np.concatenate([[0], w])) / 2, W_a, Jinv @ (Gt @ wp ** 2 - np.cross(w, J @ w))]))
        q_e = rm.hamilton_prod(QD[i], rm.invert(q))
        Y.append(np.concatenate([p_, eta[:2], [q_e[3]], v, w, wp]))
        CA.append(np.concatenate([W_R_B.T @ W_a, [w[2]]]))
        CP.append(wp)
    wh = np.sqrt(m * g / 4 / cfg.robot.alloc.cf)
    np.savez_compressed(os.path.join(HERE, 'model_props.npz'), X=X, U=U, QD=QD, Gf=Gf, Gt=Gt, F=np.array(F), Y=np.array(Y),
                        CMDACC=np.array(CA), CMDPROPS=np.array(CP), wh=wh, wp_max=float(cfg.robot.limits.wp),
                        inertia=np.array(cfg.robot.inertia, float), mass=m, cf=cfg.robot.alloc.cf, ct=cfg.robot.alloc.ct,
                        motors=np.array(cfg.robot.alloc.motors, float))
    print('model_props: f range', np.abs(F).max(), 'Gt', np.abs(Gt).max())


def refgen_case():
    """Trajectories of the reference's own RefGen (sdf_nmpc/ref_gen.py) on random states / way-points for every yaw mode,
    with and without stop-and-turn, plus joystick references."""
    import copy
    from sdf_nmpc.ref_gen import RefGen
    from sdf_nmpc.utils.config import Config
    from sdf_nmpc.utils.reference import Waypoint
    from sdf_nmpc.utils.math import yaw2quat
    cfg0 = Config('/root/reference/sdf_nmpc/config/default.yaml')
    rng = np.random.default_rng(21)
    out = {}
    B, Wmax, N = 16, 4, cfg0.mpc.N
    X0 = np.concatenate([rng.uniform(-2, 2, (B, 3)), np.zeros((B, 4)), rng.uniform(-1, 1, (B, 3))], 1)
    X0[:, 3:7] = np.array([yaw2quat(y) for y in rng.uniform(-np.pi, np.pi, B)])
    nw = rng.integers(1, Wmax + 1, B)
    WP = X0[:, None, :3] + rng.uniform(-3, 3, (B, Wmax, 3))
    WP[0, 0] = X0[0, :3] + np.array([0.03, 0.02, 0.0])          # closer than yaw_align_dmin
    WP[1, 0] = X0[1, :3] + np.array([0.5, 0.2, 0.1]); nw[1] = 1  # shorter than vref: vref saturates
    WP[2] = X0[2, None, :3] + np.cumsum(rng.uniform(0.2, 0.6, (Wmax, 3)), 0); nw[2] = Wmax   # ends inside the horizon
    WQ = np.array([[yaw2quat(y) for y in row] for row in rng.uniform(-np.pi, np.pi, (B, Wmax))])
    out.update(X0=X0, WP=WP, WQ=WQ, nw=nw)
    for mode in ('align', 'ref', 'current', 'zero'):
        for sat in (False, True):
            cfg = copy.deepcopy(cfg0)
            cfg.ref.yaw_mode = mode
            cfg.ref.stop_and_turn.enable = sat
            cfg.ref.align_yaw_offset = 0.1 if mode == 'align' else 0
            P, Q, V, Wz = [], [], [], []
            for b in range(B):
                rg = RefGen(cfg)
                rg.x0 = X0[b]
                traj = rg.gen_ref_list_wps([Waypoint(WP[b, j], WQ[b, j]) for j in range(nw[b])])
                traj = list(traj) + [traj[-1]] * (N + 1 - len(traj))        # stop-and-turn returns N entries
                P.append([np.asarray(r.p, float) for r in traj[:N + 1]]); Q.append([np.asarray(r.q, float) for r in traj[:N + 1]])
                V.append([np.asarray(r.v, float) for r in traj[:N + 1]]); Wz.append([float(r.wz) for r in traj[:N + 1]])
            tag = f'{mode}_{int(sat)}'
            out.update({f'P_{tag}': np.array(P), f'Q_{tag}': np.array(Q), f'V_{tag}': np.array(V), f'Wz_{tag}': np.array(Wz)})
    VW = rng.uniform(-1, 1, (B, 4)); VW[0, :2] = 0.01
    out['VW'] = VW
    for mode in ('align', 'zero'):
        cfg = copy.deepcopy(cfg0); cfg.ref.yaw_mode = mode
        P, Q, V, Wz = [], [], [], []
        for b in range(B):
            rg = RefGen(cfg); rg.x0 = X0[b]
            traj = rg.gen_ref_joystick(VW[b])
            P.append([np.asarray(r.p, float) for r in traj]); Q.append([np.asarray(r.q, float) for r in traj])
            V.append([np.asarray(r.v, float) for r in traj]); Wz.append([float(r.wz) for r in traj])
        out.update({f'JP_{mode}': np.array(P), f'JQ_{mode}': np.array(Q), f'JV_{mode}': np.array(V), f'JWz_{mode}': np.array(Wz)})
    rg = RefGen(cfg0)
    hov = []
    for b in range(B):
        rg.x0 = X0[b]; r = rg.from_x0()[0]
        hov.append(np.concatenate([r.p, r.q, r.v, [r.wz]]))
    out['HOVER'] = np.array(hov)
    np.savez_compressed(os.path.join(HERE, 'refgen.npz'), **out)
    print('refgen: cases', len(out))


def stability_case():
    """``utils/stability.py:6-42``: the reference's symbolic r_tilde expression evaluated at seeded points of the input box."""
    import importlib.util
    import sympy
    spec = importlib.util.spec_from_file_location('ref_stability', '/root/reference/sdf_nmpc/utils/stability.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    expr = mod.symbolic_r_tilde()
    f = sympy.lambdify(sympy.symbols('T phi theta g m dt r_1 r_2 r_3'), expr, 'numpy')
    rng = np.random.default_rng(21)
    n = 64
    T, phi, theta = rng.uniform(0.0, 20.0, n), rng.uniform(-0.7, 0.7, n), rng.uniform(-0.7, 0.7, n)
    r = rng.uniform(0.01, 60.0, (3, n))
    dt = rng.uniform(0.01, 0.1, n)
    val = np.array([float(f(T[i], phi[i], theta[i], 9.81, 1.0, dt[i], r[0, i], r[1, i], r[2, i])) for i in range(n)])
    np.savez_compressed(os.path.join(HERE, 'stability.npz'), T=T, phi=phi, theta=theta, r=r, dt=dt, r_tilde=val, expr=str(expr))
    print('stability: r_tilde range', val.min(), val.max())


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(4)
    sdf_case('archB', 512, seed=0, **synth.ARCH_B, bias_scale=0.0)
    sdf_case('archA_bias', 512, seed=1, **synth.ARCH_A, bias_scale=0.1)
    sdf_case('archB_cal', 512, seed=0, **synth.ARCH_B, calibrated=True)
    sdf_case('small_state_relu', 256, seed=4, embed='cube', nb_freqs=3, layer_sizes=(64, 64, 32, 16), res='state',
             act='relu', w0=1.0)
    sdf_case('small_latent_softplus', 256, seed=5, embed='ico', nb_freqs=2, layer_sizes=(64, 32, 32, 16), res='latent',
             act='softplus', w0=1.0, size_latent=32)
    decoder_case()
    model_att_case()
    model_variants_case()
    model_props_case()
    refgen_case()
    stability_case()
    encoder_case()
