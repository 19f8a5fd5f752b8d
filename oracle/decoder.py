"""Oracle: VAE decoder (inference path), numpy float64.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Restates
  * ``sdf_nmpc/vae.py:42-45``              (decode = decoder(latent) reshaped to the image)
  * ``sdf_nmpc/network/vae.py:63-90``      (Linear, ELU, Unflatten(512, 8, 15), 4 ResBlockDeconv, ConvTranspose2d(32, 1, 5, 1, 2),
                                            Upsample(size, 'bilinear'), Sigmoid)
  * ``sdf_nmpc/network/resnet.py:59-111``  (ResBlockDeconv: convT-BN-ReLU-convT-BN + convT 1x1 / BN shortcut, ReLU)
with BatchNorm in eval mode and dropout = identity.  The transposed convolution is written from its definition (every input
pixel scatters its k x k stamp into the output), independent of the sub-pixel convolution form the CUDA path uses.
``net`` is the dict layout of ``sdf_nmpc_b200.synth.make_decoder_net``.  Pinned to the reference's own ``Decoder`` by
``tests/golden/decoder.npz`` (``tests/golden/make_golden.py:decoder_case``).
"""
import numpy as np

from .encoder import batchnorm


def conv_transpose2d(x, w, stride, pad, out_pad):
    """x [B,Ci,H,W], w [Ci,Co,k,k] (torch ConvTranspose2d layout) -> [B,Co,(H-1)s-2p+k+op,...]: out[s*i - p + ky] += x[i] w[ky]."""
    B, Ci, H, W = x.shape
    _, Co, k, _ = w.shape
    Ho, Wo = (H - 1) * stride - 2 * pad + k + out_pad, (W - 1) * stride - 2 * pad + k + out_pad
    full = np.zeros((B, Co, (H - 1) * stride + k + out_pad, (W - 1) * stride + k + out_pad), x.dtype)
    for ky in range(k):
        for kx in range(k):
            full[:, :, ky:ky + (H - 1) * stride + 1:stride, kx:kx + (W - 1) * stride + 1:stride] += np.einsum('bchw,cd->bdhw', x, w[:, :, ky, kx])
    return full[:, :, pad:pad + Ho, pad:pad + Wo]


def upsample_bilinear(x, size, dtype=np.float64):
    """torch.nn.Upsample(size, mode='bilinear'), align_corners False: src = max(0, scale (dst + 0.5) - 0.5)."""
    H, W = x.shape[-2:]
    Ho, Wo = size

    def axis(n_in, n_out):
        src = np.maximum((np.arange(n_out, dtype=dtype) + 0.5) * dtype(n_in / n_out) - 0.5, 0)
        i0 = np.floor(src).astype(int)
        i1 = np.minimum(i0 + 1, n_in - 1)
        return i0, i1, (src - i0).astype(dtype)
    y0, y1, ly = axis(H, Ho)
    x0, x1, lx = axis(W, Wo)
    top = x[..., y0, :][..., x0] * (1 - lx) + x[..., y0, :][..., x1] * lx
    bot = x[..., y1, :][..., x0] * (1 - lx) + x[..., y1, :][..., x1] * lx
    return top * (1 - ly)[:, None] + bot * ly[:, None]


def decode(net, z, shape, dtype=np.float64, return_logits=False):
    """z [B,L] latent -> [B,H,W] reconstruction in [0,1]."""
    cast = lambda a: np.asarray(a, dtype)
    bn = lambda d: {k: (cast(v) if k != 'eps' else v) for k, v in d.items()}
    x = cast(z) @ cast(net['lin_w']).T + cast(net['lin_b'])
    x = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    x = x.reshape(-1, 512, 8, 15)
    for blk in net['blocks']:
        y = np.maximum(batchnorm(conv_transpose2d(x, cast(blk['ct1_w']), 2, 1, 1), bn(blk['bn1'])), 0)
        y = batchnorm(conv_transpose2d(y, cast(blk['ct2_w']), 1, 1, 0), bn(blk['bn2']))
        sc = batchnorm(conv_transpose2d(x, cast(blk['sc_w']), 2, 0, 1), bn(blk['sc_bn']))
        x = np.maximum(y + sc, 0)
    x = conv_transpose2d(x, cast(net['out_w']), 1, 2, 0) + cast(net['out_b'])[None, :, None, None]
    x = upsample_bilinear(x[:, 0], shape, dtype)
    return x if return_logits else 1 / (1 + np.exp(-x))
