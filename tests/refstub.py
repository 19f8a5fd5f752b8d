"""Stand-ins for the reference's TorchScript archives (``sdf_nmpc/data/sdf_90_25664.pt``, ``vae.pt`` - Git-LFS objects that
are absent offline): torch modules with the parameter names and attributes the reference keeps on its scripted modules
(``network/neural_df.py:28-37,61-89``, ``network/vae.py:19-32``, ``network/resnet.py:32-46``), filled with the repo's seeded
synthetic weights and saved with ``torch.jit.script`` so that the drop-in constructors load them exactly like
``gen_model.py:32`` / ``vae.py:11`` load the real files.  Only names, shapes and attributes matter here (the loaders read the
state dict); ``forward`` is a placeholder."""
import torch
from torch import nn


class _Embed(nn.Module):
    def __init__(self, dirs):
        super().__init__()
        self.register_buffer('dirs', torch.as_tensor(dirs))


class NeuralDFStub(nn.Module):
    def __init__(self, net):
        super().__init__()
        self.nb_freqs, self.res, self.w0 = int(net['nb_freqs']), str(net['res']), float(net['w0'])
        self.activation, self.embeddings = str(net['act']), str(net['embed'])
        self.size_latent, self.max_df, self.signed = int(net['size_latent']), float(net['max_df']), True
        lin = lambda W, b: self._lin(W, b)
        self.embed = _Embed(net['dirs'])
        self.layers = nn.ModuleDict(dict(
            main1=nn.Sequential(lin(net['W1'], net['b1']), nn.Identity(), nn.Identity(), lin(net['W2'], net['b2']), nn.Identity(), nn.Identity()),
            main2=nn.Sequential(lin(net['W3'], net['b3']), nn.Identity(), nn.Identity(), lin(net['W4'], net['b4']), nn.Identity(), nn.Identity()),
            df=nn.Sequential(lin(net['W5'], net['b5']))))

    @staticmethod
    def _lin(W, b):
        m = nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            m.weight.copy_(torch.as_tensor(W)); m.bias.copy_(torch.as_tensor(b))
        return m

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x[:, :1]


def _bn(d):
    m = nn.BatchNorm2d(d['weight'].shape[0])
    with torch.no_grad():
        m.weight.copy_(torch.as_tensor(d['weight'])); m.bias.copy_(torch.as_tensor(d['bias']))
        m.running_mean.copy_(torch.as_tensor(d['mean'])); m.running_var.copy_(torch.as_tensor(d['var']))
    return m


def _conv(W, bias=None, stride=1, pad=0):
    m = nn.Conv2d(W.shape[1], W.shape[0], W.shape[2], stride, pad, bias=bias is not None)
    with torch.no_grad():
        m.weight.copy_(torch.as_tensor(W))
        if bias is not None:
            m.bias.copy_(torch.as_tensor(bias))
    return m


class _ResBlockStub(nn.Module):
    def __init__(self, blk):
        super().__init__()
        s = int(blk['stride'])
        self.layers = nn.Sequential(_conv(blk['conv1_w'], None, s, 1), _bn(blk['bn1']), nn.ReLU(), _conv(blk['conv2_w'], None, 1, 1), _bn(blk['bn2']))
        self.shortcut = nn.Sequential(_conv(blk['sc_w'], None, s, 0), _bn(blk['sc_bn'])) if 'sc_w' in blk else nn.Sequential()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x


class _EncoderStub(nn.Module):
    def __init__(self, enc):
        super().__init__()
        mean = nn.Linear(enc['mean_w'].shape[1], enc['mean_w'].shape[0])
        with torch.no_grad():
            mean.weight.copy_(torch.as_tensor(enc['mean_w'])); mean.bias.copy_(torch.as_tensor(enc['mean_b']))
        self.layers = nn.ModuleDict(dict(
            resnet=nn.Sequential(_conv(enc['conv0_w'], enc['conv0_b'], 2, 3), nn.ELU(), nn.MaxPool2d(3, 2, 1),
                                 *[_ResBlockStub(b) for b in enc['blocks']]),
            mean=mean))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x


def _convt(W, bias=None, stride=1, pad=0, out_pad=0):
    m = nn.ConvTranspose2d(W.shape[0], W.shape[1], W.shape[2], stride, pad, out_pad, bias=bias is not None)
    with torch.no_grad():
        m.weight.copy_(torch.as_tensor(W))
        if bias is not None:
            m.bias.copy_(torch.as_tensor(bias))
    return m


class _ResBlockDeconvStub(nn.Module):
    def __init__(self, blk):
        super().__init__()
        self.layers = nn.Sequential(_convt(blk['ct1_w'], None, 2, 1, 1), _bn(blk['bn1']), nn.ReLU(), _convt(blk['ct2_w'], None, 1, 1, 0),
                                    _bn(blk['bn2']))
        self.shortcut = nn.Sequential(_convt(blk['sc_w'], None, 2, 0, 1), _bn(blk['sc_bn']))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x


class _DecoderStub(nn.Module):
    """Parameter names of ``network/vae.py:63-90`` (Sequential indices 0 Linear, 4-7 ResBlockDeconv, 8 ConvTranspose2d)."""

    def __init__(self, dec):
        super().__init__()
        lin = nn.Linear(dec['lin_w'].shape[1], dec['lin_w'].shape[0])
        with torch.no_grad():
            lin.weight.copy_(torch.as_tensor(dec['lin_w'])); lin.bias.copy_(torch.as_tensor(dec['lin_b']))
        self.layers = nn.ModuleDict(dict(
            resnet=nn.Sequential(lin, nn.ELU(), nn.Identity(), nn.Identity(), *[_ResBlockDeconvStub(b) for b in dec['blocks']],
                                 _convt(dec['out_w'], dec['out_b'], 1, 2, 0))))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x


class VaeStub(nn.Module):
    def __init__(self, enc, dec=None):
        super().__init__()
        self.encoder = _EncoderStub(enc)
        if dec is not None:
            self.decoder = _DecoderStub(dec)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x
