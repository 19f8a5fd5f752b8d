import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


# the synthetic SDF networks the golden fixtures were generated with (tests/golden/make_golden.py)
def sdf_cases():
    from sdf_nmpc_b200 import synth
    return {
        'archB': dict(seed=0, **synth.ARCH_B, bias_scale=0.0),
        'archA_bias': dict(seed=1, **synth.ARCH_A, bias_scale=0.1),
        'archB_cal': dict(seed=0, **synth.ARCH_B, calibrated=True),
        'small_state_relu': dict(seed=4, embed='cube', nb_freqs=3, layer_sizes=(64, 64, 32, 16), res='state', act='relu', w0=1.0),
        'small_latent_softplus': dict(seed=5, embed='ico', nb_freqs=2, layer_sizes=(64, 32, 32, 16), res='latent',
                                      act='softplus', w0=1.0, size_latent=32),
    }


def golden_sdf_inputs(name):
    """(net, pts [n,3], latents [4,L], idx [n], fixture)"""
    from sdf_nmpc_b200 import synth
    G = np.load(os.path.join(GOLDEN, f'sdf_{name}.npz'))
    net = synth.make_sdf_net(**sdf_cases()[name])
    n = int(G['n'])
    pts = synth.sdf_query_points(n, seed=int(G['pts_seed']))
    lat = np.random.default_rng(int(G['lat_seed'])).standard_normal((int(G['n_lat']), net['size_latent'])).astype(np.float32)
    idx = (np.arange(n) * int(G['n_lat']) // n).astype(np.int32)
    return net, pts, lat, idx, G


def relmax(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(np.asarray(b, np.float64)).max())


def make_problem(B, cfg, seed=1000, flag=1.0, model='att'):
    """Oracle-side problem dict for B randomised instances."""
    from sdf_nmpc_b200 import synth
    w = synth.nmpc_instances(B, N=cfg['N'], dt=cfg['dt'][0], seed=seed, model=model)
    prob = dict(x0=w['x0'], yref=w['yref'], W=w['W'], yrefN=w['yrefN'], WN=w['WN'], qd=w['qd'],
                flag=np.full(B, float(flag)), latent=w['latent'])
    prob['W_p_Co'] = np.einsum('bij,j->bi', w['W_R_Bo'], cfg['B_p_C']) + w['W_p_Bo']
    prob['W_R_Co'] = w['W_R_Bo'] @ cfg['B_R_C']
    return prob, w


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return 'cuda:0'
