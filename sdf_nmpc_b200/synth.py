"""Seeded synthetic weights and workloads.

The reference's pretrained networks (``sdf_nmpc/data/*.pt``) are Git-LFS
pointers that are absent offline, so every experiment uses seeded synthetic
weights with the *distributions* of the reference initialisers
(``sdf_nmpc/utils/layer_init.py:5-25``: SIREN ``U(+-sqrt(6/fan_in)/w0)`` for the
SDF, Xavier-uniform for the convolutions) drawn from numpy's PCG64 so that the
same bits are produced here and on the GPU box.  Workloads follow SURVEY.md
section 8d (configs 1-5).  Everything here is numpy; nothing touches the oracle.
"""
import numpy as np

from .embedding import directions

ARCH_A = dict(embed='oct', nb_freqs=5, layer_sizes=(256, 256, 128, 64))   # df_train.py:98-102 literals
ARCH_B = dict(embed='dod', nb_freqs=5, layer_sizes=(256, 256, 128, 64))   # best match to the shipped file size


def make_sdf_net(seed=0, embed='dod', nb_freqs=5, layer_sizes=(256, 256, 128, 64), size_latent=128,
                 res='full', w0=20.0, act='sin', max_df=1.0, bias_scale=0.0, calibrated=False):
    """Synthetic NeuralDF parameters as a dict of numpy float32 arrays.

    Shapes follow ``sdf_nmpc/network/neural_df.py:61-89``.  ``calibrated``
    applies the SURVEY 8c recipe (output bias 0.75, output weight x3) so that
    the SDF constraint 0.37 <= d is active on part of a horizon only.
    """
    rng = np.random.default_rng(seed)
    dirs = directions(embed)
    E = 3 + 2 * nb_freqs * dirs.shape[1]
    L = size_latent
    h1, h2, h3, h4 = layer_sizes
    in3 = h2 + (E + L) * (res == 'full') + E * (res == 'state') + L * (res == 'latent')

    def lin(n_out, n_in):
        bound = np.sqrt(6.0 / n_in) / w0 if act == 'sin' else np.sqrt(1.0 / n_in)
        W = rng.uniform(-bound, bound, size=(n_out, n_in)).astype(np.float32)
        b = (rng.uniform(-bias_scale, bias_scale, size=(n_out,)) if bias_scale > 0 else np.zeros(n_out)).astype(np.float32)
        return W, b

    net = dict(dirs=dirs, nb_freqs=int(nb_freqs), w0=float(w0), act=act, res=res, max_df=float(max_df),
               size_latent=int(L), embed=embed, seed=int(seed))
    net['W1'], net['b1'] = lin(h1, E + L)
    net['W2'], net['b2'] = lin(h2, h1)
    net['W3'], net['b3'] = lin(h3, in3)
    net['W4'], net['b4'] = lin(h4, h3)
    net['W5'], net['b5'] = lin(1, h4)
    if calibrated:
        net['W5'] = (net['W5'] * 3.0).astype(np.float32)
        net['b5'] = np.full((1,), 0.75, np.float32)
    return net


ENCODER_BLOCKS = ((64, 2), (128, 2), (256, 2), (512, 1))   # (size_in, stride): network/vae.py:23-26


def make_encoder_net(seed=0, size_latent=128, nb_chan=1):
    """Synthetic VAE-encoder parameters (``sdf_nmpc/network/vae.py:19-32``,
    ``resnet.py:32-46``) with *randomised* BatchNorm statistics so that BN
    folding is exercised.  Conv weights [Cout,Cin,kh,kw] float32."""
    rng = np.random.default_rng(seed)

    def conv(cout, cin, k):
        fan_in, fan_out = cin * k * k, cout * k * k
        bound = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-bound, bound, size=(cout, cin, k, k)).astype(np.float32)

    def bn(c):
        return dict(weight=rng.uniform(0.5, 1.5, c).astype(np.float32), bias=rng.uniform(-0.2, 0.2, c).astype(np.float32),
                    mean=(0.1 * rng.standard_normal(c)).astype(np.float32), var=rng.uniform(0.5, 1.5, c).astype(np.float32),
                    eps=1e-5)

    net = dict(size_latent=int(size_latent), nb_chan=int(nb_chan), seed=int(seed))
    net['conv0_w'] = conv(64, nb_chan, 7)
    net['conv0_b'] = rng.uniform(-0.05, 0.05, 64).astype(np.float32)
    blocks = []
    for cin, stride in ENCODER_BLOCKS:
        cout = cin * stride
        blk = dict(stride=stride, conv1_w=conv(cout, cin, 3), bn1=bn(cout), conv2_w=conv(cout, cout, 3), bn2=bn(cout))
        if stride != 1:
            blk['sc_w'] = conv(cout, cin, 1)
            blk['sc_bn'] = bn(cout)
        blocks.append(blk)
    net['blocks'] = blocks
    bound = 1.0 / np.sqrt(2048.0)
    net['mean_w'] = rng.uniform(-bound, bound, size=(size_latent, 2048)).astype(np.float32)
    net['mean_b'] = rng.uniform(-bound, bound, size=(size_latent,)).astype(np.float32)
    return net


def make_decoder_net(seed=0, size_latent=128):
    """Synthetic VAE-decoder parameters (``sdf_nmpc/network/vae.py:63-90``, ``resnet.py:59-111``), torch layouts:
    ConvTranspose2d weights [Cin, Cout, kh, kw]; BatchNorm statistics randomised as for the encoder."""
    rng = np.random.default_rng(seed)

    def convt(cin, cout, k, gain=1.0):
        # a stride-2 transposed convolution spreads each input over k*k/4 outputs: bound chosen to keep activations O(1)
        bound = gain * np.sqrt(6.0 / (cin * k * k / 2.0))
        return rng.uniform(-bound, bound, size=(cin, cout, k, k)).astype(np.float32)

    def bn(c):
        return dict(weight=rng.uniform(0.5, 1.5, c).astype(np.float32), bias=rng.uniform(-0.2, 0.2, c).astype(np.float32),
                    mean=(0.1 * rng.standard_normal(c)).astype(np.float32), var=rng.uniform(0.5, 1.5, c).astype(np.float32),
                    eps=1e-5)

    J = 512 * 8 * 15
    bound = 1.0 / np.sqrt(size_latent)
    net = dict(size_latent=int(size_latent), seed=int(seed),
               lin_w=rng.uniform(-bound, bound, size=(J, size_latent)).astype(np.float32),
               lin_b=rng.uniform(-bound, bound, size=(J,)).astype(np.float32))
    blocks = []
    c = 512
    for _ in range(4):
        co = c // 2
        blocks.append(dict(ct1_w=convt(c, co, 3), bn1=bn(co), ct2_w=convt(co, co, 3, 0.7), bn2=bn(co), sc_w=convt(c, co, 1, 0.5),
                           sc_bn=bn(co)))
        c = co
    net['blocks'] = blocks
    net['out_w'] = convt(c, 1, 5, 0.5)
    net['out_b'] = rng.uniform(-0.2, 0.2, size=(1,)).astype(np.float32)
    return net


# --------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------

def braking_coeffs(degree=4):
    """Synthetic coefficients of the 3-variate braking-distance polynomial (``utils/math.py:294-321`` term order: total degree,
    then a, then b; c = td - a - b).  The reference's file ``mpc.braking_dist.coeff_file`` is fitted by ``scripts/find_amax.py`` into
    the user's cache directory and is not shipped; this stand-in has the same shape: a small offset plus anisotropic quadratic
    growth with the speed and two higher-order terms."""
    vals = {(0, 0, 0): 0.02, (2, 0, 0): 0.10, (0, 2, 0): 0.11, (0, 0, 2): 0.14, (0, 0, 3): 0.01, (2, 2, 0): 0.004}
    co = []
    for td in range(degree + 1):
        for a in range(td + 1):
            for b in range(td + 1 - a):
                co.append(vals.get((a, b, td - a - b), 0.0))
    return np.array(co)


def sdf_query_points(n, seed=3):
    """Config 2: points uniform in the frustum box [0,5]x[-2.5,2.5]x[-1.5,1.5], float32 [n,3]."""
    rng = np.random.default_rng(seed)
    lo = np.array([0.0, -2.5, -1.5], np.float32)
    hi = np.array([5.0, 2.5, 1.5], np.float32)
    return (lo + (hi - lo) * rng.random((n, 3), dtype=np.float32)).astype(np.float32)


def latent(seed=2, size_latent=128, n=None):
    rng = np.random.default_rng(seed)
    shape = (size_latent,) if n is None else (n, size_latent)
    return rng.standard_normal(shape).astype(np.float32)


def yaw2quat(yaw):
    yaw = np.asarray(yaw, float)
    z = np.zeros_like(yaw)
    return np.stack([np.cos(0.5 * yaw), z, z, np.sin(0.5 * yaw)], -1)


def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


PROPS_WH = float(np.sqrt(1.46 * 9.81 / 4 / 0.02246))     # hover rotor speed sqrt(m g / 4 / cf), quad_props.py:51 with default.yaml:78-81


def nmpc_instances(B, N=20, dt=0.075, seed=1000, vref=3.0, weights=None, first=0, model='att'):
    """Config 3/4: B randomised NMPC instances (per-instance seed ``seed + first + i``).

    Returns a dict of float64 arrays: x0 [B,10], yref [B,N,11], W [B,N,11], yrefN [B,4], WN [B,4],
    qd [B,N+1,4], W_p_Bo [B,3], W_R_Bo [B,3,3], latent [B,128] (float32).
    ``weights`` = dict(pos,vel,att,rates,acc); default = constraints-on set (default.yaml:43-48).
    ``model``: 'att' (default), 'att_tau' (x0 additionally tilted by up to 0.2 rad in roll/pitch) or 'acc' (yrefN/WN
    are [B,7] and the output rows follow quad_acc.py:40-58) or 'props' (x0 [B,13] with a tilted attitude and a small body
    rate, yref / W [B,N,16], yrefN / WN [B,12], rows of quad_props.py:56-69 with the hover rotor speed as input reference).
    """
    if weights is None:
        weights = dict(pos=[0, 0, 5], vel=[3, 3, 3], att=[50, 50, 10], rates=[0, 0, 3], acc=0.05)
    nx, ny = (13, 16) if model == 'props' else (10, 11)
    x0 = np.zeros((B, nx))
    yref = np.zeros((B, N, ny))
    nyN = 7 if model == 'acc' else (12 if model == 'props' else 4)
    yrefN = np.zeros((B, nyN))
    qd = np.zeros((B, N + 1, 4))
    lat = np.zeros((B, 128), np.float32)
    Rbo = np.zeros((B, 3, 3))
    for i in range(B):
        rng = np.random.default_rng(seed + first + i)
        p = rng.uniform(-1, 1, 3) + np.array([0, 0, 2.0])
        yaw = rng.uniform(-np.pi, np.pi)
        v = rng.uniform(-3, 3, 3)
        nv = np.linalg.norm(v)
        if nv > 3.0:
            v *= 3.0 / nv
        lat[i] = rng.standard_normal(128).astype(np.float32)
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        v_ref = vref * d
        x0[i, :3], x0[i, 3:7], x0[i, 7:10] = p, yaw2quat(yaw), v
        if model in ('att_tau', 'props'):
            r, pt = rng.uniform(-0.2, 0.2, 2)
            qr = np.array([np.cos(r / 2), np.sin(r / 2), 0, 0]); qp = np.array([np.cos(pt / 2), 0, np.sin(pt / 2), 0])
            x0[i, 3:7] = _qmul(_qmul(yaw2quat(yaw), qp), qr)
        if model == 'props':
            x0[i, 10:13] = rng.uniform(-0.3, 0.3, 3)
        k = np.arange(N + 1)[:, None]
        p_ref = p[None] + v_ref[None] * k * dt
        yaw_ref = np.arctan2(v_ref[1], v_ref[0])
        qd[i] = yaw2quat(yaw_ref)
        yref[i, :, 0:3] = p_ref[:N]
        yrefN[i, 0:3] = p_ref[N]
        if model == 'props':
            yref[i, :, 6:9] = v_ref
            yref[i, :, 12:16] = PROPS_WH
            yrefN[i, 6:9] = v_ref
        else:
            yref[i, :, 4:7] = v_ref
        if model == 'acc':
            yrefN[i, 4:7] = v_ref
        c, s = np.cos(yaw), np.sin(yaw)
        Rbo[i] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    w = weights
    Wrow = np.array(list(w['pos']) + [w['att'][2]] + list(w['vel']) + [w['att'][0], w['att'][1], w['rates'][2], w['acc']], float)
    if model == 'acc':
        Wrow = np.array(list(w['pos']) + [w['att'][2]] + list(w['vel']) + [w['acc']] * 3 + [w['rates'][2]], float)
    if model == 'props':
        Wrow = np.array(list(w['pos']) + list(w['att']) + list(w['vel']) + list(w['rates']) + [w['acc']] * 4, float)
    W = np.broadcast_to(Wrow, (B, N, ny)).copy()
    WN = np.broadcast_to(Wrow[:nyN], (B, nyN)).copy()
    return dict(x0=x0, yref=yref, W=W, yrefN=yrefN, WN=WN, qd=qd, W_p_Bo=x0[:, :3].copy(), W_R_Bo=Rbo, latent=lat)


def depth_frame(seed=1, H=270, W=480, hfov=0.7592, vfov=0.4903, spherical=False, dmax=5.0, invalid_frac=0.02):
    """Config 1/5: one synthetic depth (pinhole, metres along x) or spherical range image:
    3 random spheres r in [0.3,1] m at 1.5-4.5 m and a back wall at 6 m, ray-cast with the pixel models of
    ``utils/visualization.py:80-109`` (consistent with ``preprocessing.py:20-27``); 2 % of pixels set to 0."""
    rng = np.random.default_rng(seed)
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    if spherical:
        az = hfov * (1 - u / (W / 2))
        el = vfov * (1 - v / (H / 2))
        ray = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)], -1)
    else:
        ray = np.stack([np.ones((H, W)), np.tan(hfov) * (1 - u / (W / 2)) * np.ones((H, 1)),
                        np.tan(vfov) * (1 - v / (H / 2)) * np.ones((1, W))], -1)
    rn = ray / np.linalg.norm(ray, axis=-1, keepdims=True)
    t_hit = 6.0 / np.maximum(rn[..., 0], 1e-6)              # wall x = 6 m
    for _ in range(3):
        r = rng.uniform(0.3, 1.0)
        c = np.array([rng.uniform(1.5, 4.5), rng.uniform(-1.5, 1.5), rng.uniform(-0.8, 0.8)])
        b = rn @ c
        disc = b * b - (c @ c - r * r)
        t = b - np.sqrt(np.maximum(disc, 0))
        ok = (disc > 0) & (t > 0)
        t_hit = np.where(ok & (t < t_hit), t, t_hit)
    img = t_hit if spherical else t_hit * rn[..., 0]        # range, or depth along the optical axis
    img = img.astype(np.float32)
    img[rng.random((H, W)) < invalid_frac] = 0.0
    return img
