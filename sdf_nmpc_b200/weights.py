"""Weight packing: reference torch modules / TorchScript archives -> plain numpy dictionaries.

``sdf_from_module`` reads a ``NeuralDF`` (eager or ``torch.jit.script``-ed, as ``gen_model.py:32`` loads it)
through its ``state_dict`` and the attributes the reference keeps on the scripted module
(``neural_df.py:28-37``: nb_freqs, res, w0, activation, embeddings, size_latent, max_df).
``encoder_from_module`` does the same for ``Vae.encoder`` (``network/vae.py:19-32``).
The dictionaries have the layout of ``synth.make_sdf_net`` / ``synth.make_encoder_net``.
"""
import struct

import numpy as np

ACT_CODES = {'sin': 0, 'relu': 1, 'softplus': 2}
SDF_KEYS = ('W1', 'b1', 'W2', 'b2', 'W3', 'b3', 'W4', 'b4', 'W5', 'b5')


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


def sdf_from_module(module):
    sd = {k: v for k, v in module.state_dict().items()}
    net = dict(nb_freqs=int(module.nb_freqs), w0=float(module.w0), act=str(module.activation), res=str(module.res),
               max_df=float(module.max_df), size_latent=int(module.size_latent), embed=str(module.embeddings))
    if 'layers.embeddings.0.dirs' in sd:
        net['dirs'] = _np(sd['layers.embeddings.0.dirs'])
    elif 'embed.dirs' in sd:
        net['dirs'] = _np(sd['embed.dirs'])
    else:                                   # embed == 'none': identity embedding
        net['dirs'] = np.eye(3, dtype=np.float32)
        net['nb_freqs'] = 0
    for name, key in (('W1', 'layers.main1.0'), ('W2', 'layers.main1.3'), ('W3', 'layers.main2.0'),
                      ('W4', 'layers.main2.3'), ('W5', 'layers.df.0')):
        net[name] = _np(sd[key + '.weight'])
        net['b' + name[1]] = _np(sd[key + '.bias'])
    return net


def encoder_from_module(encoder):
    sd = encoder.state_dict()
    net = dict(size_latent=int(sd['layers.mean.weight'].shape[0]), nb_chan=int(sd['layers.resnet.0.weight'].shape[1]))
    net['conv0_w'], net['conv0_b'] = _np(sd['layers.resnet.0.weight']), _np(sd['layers.resnet.0.bias'])
    blocks = []
    for i in range(4):
        p = f'layers.resnet.{3 + i}.'

        def bn(q):
            return dict(weight=_np(sd[q + '.weight']), bias=_np(sd[q + '.bias']), mean=_np(sd[q + '.running_mean']),
                        var=_np(sd[q + '.running_var']), eps=1e-5)
        blk = dict(conv1_w=_np(sd[p + 'layers.0.weight']), bn1=bn(p + 'layers.1'), conv2_w=_np(sd[p + 'layers.3.weight']),
                   bn2=bn(p + 'layers.4'))
        blk['stride'] = 2 if (p + 'shortcut.0.weight') in sd else 1
        if blk['stride'] == 2:
            blk['sc_w'], blk['sc_bn'] = _np(sd[p + 'shortcut.0.weight']), bn(p + 'shortcut.1')
        blocks.append(blk)
    net['blocks'] = blocks
    net['mean_w'], net['mean_b'] = _np(sd['layers.mean.weight']), _np(sd['layers.mean.bias'])
    return net


def decoder_from_module(decoder):
    """``Vae.decoder`` (``network/vae.py:63-90``) -> the dictionary layout of ``synth.make_decoder_net``."""
    sd = decoder.state_dict()
    net = dict(size_latent=int(sd['layers.resnet.0.weight'].shape[1]), lin_w=_np(sd['layers.resnet.0.weight']),
               lin_b=_np(sd['layers.resnet.0.bias']))
    if net['lin_w'].shape[0] != 512 * 8 * 15:
        raise ValueError('decoder: the Linear layer must feed the [512, 8, 15] feature map of network/vae.py:73-75')

    def bn(q):
        return dict(weight=_np(sd[q + '.weight']), bias=_np(sd[q + '.bias']), mean=_np(sd[q + '.running_mean']),
                    var=_np(sd[q + '.running_var']), eps=1e-5)
    blocks = []
    for i in range(4):
        p = f'layers.resnet.{4 + i}.'
        blocks.append(dict(ct1_w=_np(sd[p + 'layers.0.weight']), bn1=bn(p + 'layers.1'), ct2_w=_np(sd[p + 'layers.3.weight']),
                           bn2=bn(p + 'layers.4'), sc_w=_np(sd[p + 'shortcut.0.weight']), sc_bn=bn(p + 'shortcut.1')))
    net['blocks'] = blocks
    net['out_w'], net['out_b'] = _np(sd['layers.resnet.8.weight']), _np(sd['layers.resnet.8.bias'])
    return net


def load_sdf(cfg):
    """``gen_model.py:32-34``: ``torch.jit.load(default_data_dir() / cfg.nn.sdf_weights)`` -> packed dictionary."""
    import os

    import torch

    from . import default_data_dir
    path = os.path.join(default_data_dir(), cfg.nn.sdf_weights)
    if not os.path.isfile(path) or os.path.getsize(path) < 1024:
        raise FileNotFoundError(f'{path}: TorchScript SDF weights not found (the reference ships them through Git LFS); '
                                'put the file there, call sdf_nmpc_b200.set_data_dir(...) or pass sdf_net=')
    return sdf_from_module(torch.jit.load(path, map_location='cpu').eval())


def load_encoder(cfg):
    """``vae.py:11-13``: ``torch.jit.load(default_data_dir() / cfg.nn.vae_weights)`` -> packed encoder dictionary."""
    import os

    import torch

    from . import default_data_dir
    path = os.path.join(default_data_dir(), cfg.nn.vae_weights)
    if not os.path.isfile(path) or os.path.getsize(path) < 1024:
        raise FileNotFoundError(f'{path}: TorchScript VAE weights not found (the reference ships them through Git LFS); '
                                'put the file there, call sdf_nmpc_b200.set_data_dir(...) or pass enc_net=')
    return encoder_from_module(torch.jit.load(path, map_location='cpu').eval().encoder)


def load_vae(cfg):
    """Encoder AND decoder dictionaries of the TorchScript VAE ``cfg.nn.vae_weights`` (``vae.py:11-13``)."""
    import os

    import torch

    from . import default_data_dir
    path = os.path.join(default_data_dir(), cfg.nn.vae_weights)
    if not os.path.isfile(path) or os.path.getsize(path) < 1024:
        raise FileNotFoundError(f'{path}: TorchScript VAE weights not found (the reference ships them through Git LFS); '
                                'put the file there, call sdf_nmpc_b200.set_data_dir(...) or pass enc_net= / dec_net=')
    vae = torch.jit.load(path, map_location='cpu').eval()
    return encoder_from_module(vae.encoder), (decoder_from_module(vae.decoder) if hasattr(vae, 'decoder') else None)


def sdf_desc_fields(net):
    E = 3 + 2 * int(net['nb_freqs']) * net['dirs'].shape[1]
    return dict(n_dirs=int(net['dirs'].shape[1]), nb_freqs=int(net['nb_freqs']),
                hidden=[int(net[k].shape[0]) for k in ('W1', 'W2', 'W3', 'W4')], size_latent=int(net['size_latent']),
                res_state=int(net['res'] in ('full', 'state')), res_latent=int(net['res'] in ('full', 'latent')),
                act=ACT_CODES[net['act']], w0=float(net['w0']), max_df=float(net['max_df'])), E


def save_sdf_blob(net, path):
    """Flat binary the C library loads for the CasADi-ABI symbols (``sdfnmpc_l4c_load_blob``)."""
    f, _ = sdf_desc_fields(net)
    with open(path, 'wb') as fh:
        fh.write(struct.pack('<ii', 0x42464453, 1))
        fh.write(struct.pack('<ii4iiiiiff', f['n_dirs'], f['nb_freqs'], *f['hidden'], f['size_latent'], f['res_state'],
                             f['res_latent'], f['act'], f['w0'], f['max_df']))
        fh.write(np.ascontiguousarray(net['dirs'], np.float32).tobytes())
        for k in SDF_KEYS:
            fh.write(np.ascontiguousarray(net[k], np.float32).tobytes())
    return path
