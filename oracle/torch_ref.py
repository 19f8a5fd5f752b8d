"""Oracle: torch restatements of the reference's two networks, for the CPU / library BASELINES of bench.py.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  The reference runs both networks through libtorch:
  * the SDF as a TorchScript ``NeuralDF`` wrapped by L4CasADi (``sdf_nmpc/gen_model.py:27,32-39``): batch 1,
    ``torch.set_num_threads(1)``, forward + a traced ``torch.func.jacrev`` per shooting node, on the CPU;
  * the depth encoder as ``Vae.encoder`` + ``utils/preprocessing`` on ``cfg.nn.device`` (``sdf_nmpc/vae.py:8-40``,
    default ``cuda`` -> cuDNN), batch 1.
The reference's Python cannot travel to the GPU box (and its pretrained ``.pt`` files are LFS pointers), so the
baselines are timed on these modules: the same layer sequence written out with ``torch.nn.functional`` calls on the
repo's seeded synthetic weights, i.e. the same libtorch / cuDNN kernels the reference would run.  Pinned by
``tests/test_oracle_golden.py`` against the fixtures the reference's own modules produced.

Follows ``sdf_nmpc/network/neural_df.py:61-103``, ``utils/embeddings.py:106-111``, ``utils/activation.py:12-13``,
``network/vae.py:19-38``, ``network/resnet.py:32-56``, ``utils/preprocessing.py:5-30,67-96``.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


class SdfTorch(torch.nn.Module):
    """NeuralDF.forward on a ``net`` dict of ``sdf_nmpc_b200.synth.make_sdf_net`` (input [n, 3 + L], output [n, 1])."""

    def __init__(self, net, dtype=torch.float32):
        super().__init__()
        t = lambda a: torch.nn.Parameter(torch.as_tensor(np.asarray(a), dtype=dtype), requires_grad=False)
        self.W1, self.b1, self.W2, self.b2 = t(net['W1']), t(net['b1']), t(net['W2']), t(net['b2'])
        self.W3, self.b3, self.W4, self.b4 = t(net['W3']), t(net['b3']), t(net['W4']), t(net['b4'])
        self.W5, self.b5 = t(net['W5']), t(net['b5'])
        self.dirs = t(net['dirs'])
        nf = int(net['nb_freqs'])
        self.nf = nf
        self.freq = t(2.0 ** np.linspace(0, nf - 1, nf)) if nf > 0 else None
        self.w0, self.act, self.res = float(net['w0']), net['act'], net['res']
        self.max_df = float(net['max_df'])

    def _a(self, x):
        if self.act == 'sin':
            return torch.sin(self.w0 * x)
        if self.act == 'relu':
            return torch.relu(x)
        return F.softplus(x)

    def forward(self, x):
        s, z = x[:, :3], x[:, 3:]
        if self.nf > 0:
            proj = s @ self.dirs
            xb = (proj[..., None] * self.freq).flatten(1)
            e = torch.cat([s, torch.sin(xb), torch.sin(xb + 0.5 * math.pi)], 1)
        else:
            e = s
        a = self._a(F.linear(torch.cat([e, z], 1), self.W1, self.b1))
        a = self._a(F.linear(a, self.W2, self.b2))
        if self.res in ('full', 'state'):
            a = torch.cat([a, e], 1)
        if self.res in ('full', 'latent'):
            a = torch.cat([a, z], 1)
        a = self._a(F.linear(a, self.W3, self.b3))
        a = self._a(F.linear(a, self.W4, self.b4))
        return F.linear(a, self.W5, self.b5)


def sdf_l4c_pair(net):
    """(forward, jacobian) callables the way L4CasADi 1.4.1 builds them for ``model_expects_batch_dim=True``
    (``gen_model.py:38-39``): TorchScript forward on [1, 131] and a ``torch.jit.trace`` of ``jacrev``."""
    from torch.fx.experimental.proxy_tensor import make_fx
    m = SdfTorch(net).eval()
    fwd = torch.jit.script(m)
    x0 = torch.zeros(1, 3 + int(net['size_latent']))
    # L4CasADi 1.4.1 traces functorch's jacrev with torch.jit.trace directly; under torch 2.11 that trace fails inside vmap, so
    # the jacrev is first lowered to an aten graph (make_fx) and THAT graph is traced: the same TorchScript executor, the same
    # kernels (one forward, one reverse sweep, batch 1)
    jac = torch.jit.trace(make_fx(torch.func.jacrev(lambda v: m(v)[0]))(x0), x0, check_trace=False)
    return fwd, jac


class EncoderTorch(torch.nn.Module):
    """Preprocessing + Vae.encoder (mean head) on an ``enc`` dict of ``synth.make_encoder_net``; input [B, H, W] metres."""

    def __init__(self, enc, hfov=0.7592, vfov=0.4903, max_range=5.0, is_depth=True, shape=(270, 480)):
        super().__init__()
        t = lambda a: torch.nn.Parameter(torch.as_tensor(np.asarray(a), dtype=torch.float32), requires_grad=False)
        self.conv0_w, self.conv0_b = t(enc['conv0_w']), t(enc['conv0_b'])
        self.blocks = []
        for i, blk in enumerate(enc['blocks']):
            d = dict(stride=int(blk['stride']))
            for key in ('conv1_w', 'conv2_w', 'sc_w'):
                if key in blk:
                    d[key] = t(blk[key]); self.register_parameter(f'b{i}_{key}', d[key])
            for key in ('bn1', 'bn2', 'sc_bn'):
                if key in blk:
                    for f in ('weight', 'bias', 'mean', 'var'):
                        d[key + f] = t(blk[key][f]); self.register_parameter(f'b{i}_{key}_{f}', d[key + f])
                    d[key + 'eps'] = float(blk[key]['eps'])
            self.blocks.append(d)
        self.mean_w, self.mean_b = t(enc['mean_w']), t(enc['mean_b'])
        self.max_range, self.is_depth = float(max_range), bool(is_depth)
        H, W = shape
        # Depth2Range (utils/preprocessing.py:20-27): range = depth * sqrt(1 + y^2 + z^2) of the pinhole ray
        u = torch.arange(W, dtype=torch.float32)[None, :]
        v = torch.arange(H, dtype=torch.float32)[:, None]
        y = math.tan(hfov) * (1 - u / (W / 2))
        z = math.tan(vfov) * (1 - v / (H / 2))
        self.yz_sqrt = torch.nn.Parameter(torch.sqrt(1 + y ** 2 + z ** 2), requires_grad=False)

    def _bn(self, x, d, key):
        return F.batch_norm(x, d[key + 'mean'], d[key + 'var'], d[key + 'weight'], d[key + 'bias'], False, 0.0, d[key + 'eps'])

    def forward(self, img):
        x = torch.clamp(img[:, None] / self.max_range, 0, 1)
        if self.is_depth:
            x = torch.clamp(x * self.yz_sqrt, 0, 1)
        x = F.elu(F.conv2d(x, self.conv0_w, self.conv0_b, stride=2, padding=3))
        x = F.max_pool2d(x, 3, 2, 1)
        for d in self.blocks:
            y = F.relu(self._bn(F.conv2d(x, d['conv1_w'], None, stride=d['stride'], padding=1), d, 'bn1'))
            y = self._bn(F.conv2d(y, d['conv2_w'], None, stride=1, padding=1), d, 'bn2')
            sc = self._bn(F.conv2d(x, d['sc_w'], None, stride=d['stride']), d, 'sc_bn') if 'sc_w' in d else x
            x = F.relu(y + sc)
        x = F.adaptive_avg_pool2d(x, (2, 2)).flatten(1)
        return F.linear(x, self.mean_w, self.mean_b)
