"""Oracle: depth-image preprocessing + VAE encoder (inference path), numpy.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Restates
  * ``sdf_nmpc/vae.py:15-40``                      (pipeline order, encoder returns the mean only)
  * ``sdf_nmpc/utils/preprocessing.py:5-30,67-79`` (Depth2Range, ClipDistance)
  * ``sdf_nmpc/network/vae.py:19-38``              (conv7x7/2+bias, ELU, maxpool3/2, 4 ResBlocks,
                                                    AdaptiveAvgPool2d((2,2)), Flatten, Linear)
  * ``sdf_nmpc/network/resnet.py:32-56``           (ResBlock: conv-BN-ReLU-conv-BN + shortcut, ReLU)
with BatchNorm in eval mode (running statistics) and dropout = identity.
``net`` is the dict layout of ``sdf_nmpc_b200.synth.make_encoder_net``.
"""
import numpy as np


def preprocess(img, dmax=5.0, mm_resolution=1000, hfov=0.7592, vfov=0.4903, is_depth=True, is_normalized=False,
               dtype=np.float32):
    """[...,H,W] raw -> dmax-normalised range image in [0,1] (float32 arithmetic like the reference)."""
    x = np.asarray(img, np.float32)
    H, W = x.shape[-2:]
    if not is_normalized:
        x = np.clip(x / np.float32(dmax / mm_resolution * 1000), 0, 1)
    if is_depth:
        u = np.arange(W, dtype=np.float32)[None, :]
        v = np.arange(H, dtype=np.float32)[:, None]
        th, tv = np.tan(np.float32(hfov)), np.tan(np.float32(vfov))
        yz = np.sqrt(1 + (th * (1 - 2 * u / np.float32(W))) ** 2 + (tv * (1 - 2 * v / np.float32(H))) ** 2).astype(np.float32)
        x = np.clip(x * yz, 0, 1)
    return x.astype(dtype)


def conv2d(x, w, stride, pad):
    """x [B,C,H,W], w [Co,C,k,k] -> [B,Co,Ho,Wo] (cross-correlation, like torch)."""
    B, C, H, W = x.shape
    Co, _, k, _ = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    win = np.lib.stride_tricks.sliding_window_view(xp, (k, k), axis=(2, 3))[:, :, ::stride, ::stride]   # [B,C,Ho,Wo,k,k]
    Ho, Wo = win.shape[2], win.shape[3]
    cols = win.transpose(0, 2, 3, 1, 4, 5).reshape(B * Ho * Wo, C * k * k)
    out = cols @ w.reshape(Co, -1).T
    return out.reshape(B, Ho, Wo, Co).transpose(0, 3, 1, 2)


def batchnorm(x, bn):
    s = bn['weight'] / np.sqrt(bn['var'] + bn['eps'])
    return x * s[None, :, None, None] + (bn['bias'] - bn['mean'] * s)[None, :, None, None]


def maxpool3s2(x):
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), constant_values=-np.inf)
    win = np.lib.stride_tricks.sliding_window_view(xp, (3, 3), axis=(2, 3))[:, :, ::2, ::2]
    return win.max((-1, -2))


def adaptive_avgpool2(x):
    B, C, H, W = x.shape
    out = np.zeros((B, C, 2, 2), x.dtype)
    for i in range(2):
        y0, y1 = (i * H) // 2, -((-(i + 1) * H) // 2)
        for j in range(2):
            x0, x1 = (j * W) // 2, -((-(j + 1) * W) // 2)
            out[:, :, i, j] = x[:, :, y0:y1, x0:x1].mean((2, 3))
    return out


def encode(net, x, dtype=np.float64):
    """x [B,H,W] preprocessed image -> latent mean [B,L]."""
    cast = lambda a: np.asarray(a, dtype)
    bn = lambda d: {k: (cast(v) if k != 'eps' else v) for k, v in d.items()}
    x = cast(x)[:, None]
    x = conv2d(x, cast(net['conv0_w']), 2, 3) + cast(net['conv0_b'])[None, :, None, None]
    x = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))          # ELU
    x = maxpool3s2(x)
    for blk in net['blocks']:
        s = blk['stride']
        y = np.maximum(batchnorm(conv2d(x, cast(blk['conv1_w']), s, 1), bn(blk['bn1'])), 0)
        y = batchnorm(conv2d(y, cast(blk['conv2_w']), 1, 1), bn(blk['bn2']))
        sc = batchnorm(conv2d(x, cast(blk['sc_w']), s, 0), bn(blk['sc_bn'])) if s != 1 else x
        x = np.maximum(y + sc, 0)
    f = adaptive_avgpool2(x).reshape(x.shape[0], -1)
    return f @ cast(net['mean_w']).T + cast(net['mean_b'])
