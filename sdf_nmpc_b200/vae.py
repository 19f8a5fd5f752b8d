"""Depth-image encoder wrapper, batched - host mirror of ``sdf_nmpc.vae.VaeWrapper``
(``sdf_nmpc/vae.py:7-45``): ``set_img`` + ``encode`` keep their names and meaning, with an optional leading
batch axis.  Preprocessing (``utils/preprocessing.py``: ToDevice, Reshape, ClipDistance, Depth2Range) is fused
into the first convolution kernel.  ``decode`` (``vae.py:42-45``) runs the decoder of ``network/vae.py:63-90`` through
``sdfnmpc_decoder_*`` (transposed convolutions on the encoder's tensor-core kernel, see ``csrc/encoder_api.cu``).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .sdf import _ptr, _stream


def _fp(a):
    return a.ctypes.data_as(_lib.FP)


class VaeWrapper:
    """``enc_net``: dict as produced by ``weights.encoder_from_module`` / ``synth.make_encoder_net``."""

    def __init__(self, cfg, enc_net=None, max_batch=1, device='cuda:0', dec_net=None):
        """``VaeWrapper(cfg)`` as in ``vae.py:8``: encoder and decoder come from the TorchScript file ``cfg.nn.vae_weights`` in
        the data directory (``vae.py:11-13``) unless packed dictionaries are given (``enc_net`` alone: no ``decode``)."""
        if not torch.cuda.is_available():
            raise _lib.SdfNmpcError('CUDA is not available: sdf_nmpc_b200 has no CPU path')
        if enc_net is None:
            from . import weights
            enc_net, dec_net = weights.load_vae(cfg)
        self.cfg = cfg
        self.lib = _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        s = cfg.sensor
        self.H, self.W = int(s.shape_imgs[-2]), int(s.shape_imgs[-1])
        self.L = int(enc_net['size_latent'])
        self.max_batch = int(max_batch)
        d = _lib.EncoderDesc(self.H, self.W, self.L, self.max_batch, int(bool(s.is_depth)), int(bool(s.is_normalized)),
                             float(s.dmax / s.mm_resolution * 1000), float(s.hfov), float(s.vfov))
        self._keep = []

        def arr(a):
            a = np.ascontiguousarray(a, np.float32)
            self._keep.append(a)
            return _fp(a)

        w = _lib.EncoderWeights()
        w.conv0_w, w.conv0_b = arr(enc_net['conv0_w']), arr(enc_net['conv0_b'])
        null = C.cast(None, _lib.FP)
        for i, blk in enumerate(enc_net['blocks']):
            w.conv1_w[i], w.conv2_w[i] = arr(blk['conv1_w']), arr(blk['conv2_w'])
            for j, key in enumerate(('weight', 'bias', 'mean', 'var')):
                w.bn1[i][j], w.bn2[i][j] = arr(blk['bn1'][key]), arr(blk['bn2'][key])
                w.sc_bn[i][j] = arr(blk['sc_bn'][key]) if 'sc_bn' in blk else null
            w.sc_w[i] = arr(blk['sc_w']) if 'sc_w' in blk else null
        w.mean_w, w.mean_b = arr(enc_net['mean_w']), arr(enc_net['mean_b'])
        w.bn_eps = float(enc_net['blocks'][0]['bn1'].get('eps', 1e-5))
        self.handle = C.c_void_p()
        _lib.check(self.lib.sdfnmpc_encoder_create(C.byref(self.handle), C.byref(d), C.byref(w)), 'sdfnmpc_encoder_create')
        self.img = None
        self.latent = None
        self.decoded = None
        self.dec_handle = None
        if dec_net is not None:
            dd = _lib.DecoderDesc(self.H, self.W, int(dec_net['size_latent']), self.max_batch)
            v = _lib.DecoderWeights()
            v.lin_w, v.lin_b = arr(dec_net['lin_w']), arr(dec_net['lin_b'])
            for i, blk in enumerate(dec_net['blocks']):
                v.ct1_w[i], v.ct2_w[i], v.sc_w[i] = arr(blk['ct1_w']), arr(blk['ct2_w']), arr(blk['sc_w'])
                for j, key in enumerate(('weight', 'bias', 'mean', 'var')):
                    v.bn1[i][j], v.bn2[i][j], v.sc_bn[i][j] = arr(blk['bn1'][key]), arr(blk['bn2'][key]), arr(blk['sc_bn'][key])
            v.out_w, v.out_b = arr(dec_net['out_w']), arr(dec_net['out_b'])
            v.bn_eps = float(dec_net['blocks'][0]['bn1'].get('eps', 1e-5))
            self.dec_handle = C.c_void_p()
            _lib.check(self.lib.sdfnmpc_decoder_create(C.byref(self.dec_handle), C.byref(dd), C.byref(v)), 'sdfnmpc_decoder_create')

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.lib.sdfnmpc_encoder_destroy(h)
            self.handle = None
        h = getattr(self, 'dec_handle', None)
        if h:
            self.lib.sdfnmpc_decoder_destroy(h)
            self.dec_handle = None

    def set_img(self, img):
        """Raw sensor image(s) [H,W] or [B,H,W] (numpy or tensor), as ``vae.py:31-32`` takes them."""
        t = torch.as_tensor(img)
        t = t.to(self.device, torch.float32, non_blocking=True).reshape(-1, t.shape[-2], t.shape[-1])
        if tuple(t.shape[-2:]) != (self.H, self.W):   # Reshape (preprocessing.py:92-96)
            t = torch.nn.functional.interpolate(t[:, None], size=(self.H, self.W), mode='bilinear')[:, 0]
        self.img = t.contiguous()

    def encode(self, to_numpy=True):
        """``vae.py:37-40``: latent mean(s) [B, L]."""
        B = self.img.shape[0]
        out = torch.empty(B, self.L, dtype=torch.float32, device=self.device)
        for s in range(0, B, self.max_batch):
            n = min(self.max_batch, B - s)
            _lib.check(self.lib.sdfnmpc_encoder_encode(self.handle, _ptr(self.img[s:s + n]), n, _ptr(out[s:s + n]), _stream()),
                       'sdfnmpc_encoder_encode')
        self.latent = out
        return out.cpu().numpy() if to_numpy else out

    def encode_host(self, imgs):
        """numpy in / numpy out through the C ABI with host buffers (H2D + D2H inside the call)."""
        imgs = np.ascontiguousarray(imgs, np.float32).reshape(-1, self.H, self.W)
        out = np.empty((imgs.shape[0], self.L), np.float32)
        for s in range(0, imgs.shape[0], self.max_batch):
            n = min(self.max_batch, imgs.shape[0] - s)
            _lib.check(self.lib.sdfnmpc_encoder_encode_host(self.handle, imgs[s:s + n].ctypes.data_as(C.c_void_p), n,
                                                            out[s:s + n].ctypes.data_as(C.c_void_p)), 'encode_host')
        return out

    def set_latent(self, latent):
        self.latent = torch.as_tensor(latent, dtype=torch.float32, device=self.device)

    def decode(self, to_numpy=True):
        """``vae.py:42-45``: reconstruction of the current latent(s) (from ``encode`` or ``set_latent``), values in [0, 1];
        [H, W] for a single latent as the reference reshapes it, [B, H, W] for a batch."""
        if not self.dec_handle:
            raise _lib.SdfNmpcError('VaeWrapper was built without decoder weights (dec_net)')
        lat = self.latent.reshape(-1, self.L).contiguous()
        B = lat.shape[0]
        out = torch.empty(B, self.H, self.W, dtype=torch.float32, device=self.device)
        for s in range(0, B, self.max_batch):
            n = min(self.max_batch, B - s)
            _lib.check(self.lib.sdfnmpc_decoder_decode(self.dec_handle, _ptr(lat[s:s + n]), n, _ptr(out[s:s + n]), _stream()),
                       'sdfnmpc_decoder_decode')
        self.decoded = out[0] if B == 1 else out
        return self.decoded.cpu().numpy() if to_numpy else self.decoded
