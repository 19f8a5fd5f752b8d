"""GPU parity of the VAE decoder (``VaeWrapper.decode``, sdf_nmpc/vae.py:42-45) through the C ABI against the committed outputs
of the reference's own ``Decoder`` (tests/golden/decoder.npz) and the float64 oracle.  Tolerance: 1e-4 max-norm relative on the
decoded image (values in [0, 1]), the encoder's latent tolerance."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax

pytestmark = pytest.mark.gpu
IMG_TOL = 1e-4


def _wrapper(dec, max_batch, device, enc_seed=5):
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.vae import VaeWrapper
    return VaeWrapper(Config(), synth.make_encoder_net(seed=enc_seed), max_batch=max_batch, device=device, dec_net=dec)


def test_decoder_matches_reference_fixture(cuda_device):
    from sdf_nmpc_b200 import synth
    G = np.load(os.path.join(GOLDEN, 'decoder.npz'))
    dec = synth.make_decoder_net(seed=int(G['seed']))
    z = synth.latent(seed=int(G['latent_seed']), n=2).astype(np.float32)
    vae = _wrapper(dec, 2, cuda_device)
    vae.set_latent(z)
    img = vae.decode()
    assert img.shape == (2, 270, 480) and img.min() >= 0 and img.max() <= 1
    assert relmax(img[:, ::9, ::8], G['rows32']) <= IMG_TOL, relmax(img[:, ::9, ::8], G['rows32'])
    assert relmax(img[:, ::9, ::8], G['rows64']) <= IMG_TOL
    assert np.abs(img.astype(np.float64).sum((1, 2)) / G['sum64'] - 1).max() <= 1e-5
    # one latent, as the reference calls it: [H, W]; same small-batch (split-K) path, so the image is the batch's row bit for bit
    vae.set_latent(z[1])
    one = vae.decode()
    assert one.shape == (270, 480) and np.array_equal(one, img[1])


def test_decoder_whole_image_vs_oracle_and_batch_regimes(cuda_device):
    """Every pixel against the float64 oracle, for the small-batch path (split-K, up to four latents) and the batch path (six)."""
    from oracle import decoder as OD
    from sdf_nmpc_b200 import synth
    dec = synth.make_decoder_net(seed=3)
    z = (1.5 * synth.latent(seed=21, n=6)).astype(np.float32)
    ref = OD.decode(dec, z, (270, 480))
    vae = _wrapper(dec, 6, cuda_device)
    vae.set_latent(z)
    big = vae.decode()
    vae.set_latent(z[:3])
    small = vae.decode()
    assert relmax(big, ref) <= IMG_TOL, relmax(big, ref)
    assert relmax(small, ref[:3]) <= IMG_TOL, relmax(small, ref[:3])
    assert np.array_equal(vae.decode(), small)                       # reproducible
    assert ref.std() > 0.01                                          # the synthetic decoder does not saturate to a constant


def test_encode_decode_roundtrip_api(cuda_device):
    """set_img -> encode -> decode with the reference's call sequence (vae.py:31-45); numbers checked against the oracles."""
    from oracle import decoder as OD, encoder as OE
    from sdf_nmpc_b200 import synth
    dec = synth.make_decoder_net(seed=7)
    vae = _wrapper(dec, 1, cuda_device)
    frame = synth.depth_frame(seed=4)
    vae.set_img(frame)
    lat = vae.encode()
    img = vae.decode()
    ref = OD.decode(dec, OE.encode(synth.make_encoder_net(seed=5), OE.preprocess(frame[None])), (270, 480))[0]
    assert lat.shape == (1, 128) and img.shape == (270, 480)
    assert relmax(img, ref) <= 2 * IMG_TOL, relmax(img, ref)        # latent error (1e-4) + decoder error


def test_decoder_tensor_core_and_cuda_core_paths_agree(cuda_device):
    import subprocess
    import sys
    import tempfile
    from conftest import ROOT
    from sdf_nmpc_b200 import synth
    dec = synth.make_decoder_net(seed=9)
    z = synth.latent(seed=31, n=5).astype(np.float32)
    vae = _wrapper(dec, 5, cuda_device)
    vae.set_latent(z)
    img_tc = vae.decode()
    out = os.path.join(tempfile.mkdtemp(), 'img.npy')
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests');"
            "from test_gpu_decoder import _wrapper; from sdf_nmpc_b200 import synth;"
            "vae = _wrapper(synth.make_decoder_net(seed=9), 5, 'cuda:0'); vae.set_latent(synth.latent(seed=31, n=5).astype(np.float32));"
            "np.save(%r, vae.decode())") % (ROOT, ROOT, out)
    subprocess.run([sys.executable, '-c', code], check=True, env=dict(os.environ, SDFNMPC_ENC_IMPL='ffma'))
    img_ffma = np.load(out)
    assert relmax(img_tc, img_ffma) <= IMG_TOL, relmax(img_tc, img_ffma)
    assert not np.array_equal(img_tc, img_ffma)


@pytest.mark.parametrize('shape', [(96, 160), (125, 333)])
def test_other_image_shapes_encode_and_decode(shape, cuda_device):
    """sensor.shape_imgs other than the default 270 x 480 (odd sizes exercise the partial tiles of the fused first layer, the
    adaptive pooling bins and the bilinear up-sampling): latents and reconstructions against the float64 oracles."""
    from oracle import decoder as OD, encoder as OE
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.vae import VaeWrapper
    H, W = shape
    cfg = Config()
    cfg.sensor.shape_imgs = [1, H, W]
    enc, dec = synth.make_encoder_net(seed=13), synth.make_decoder_net(seed=14)
    frames = np.stack([synth.depth_frame(seed=70 + i, H=H, W=W) for i in range(3)])
    vae = VaeWrapper(cfg, enc, max_batch=3, device=cuda_device, dec_net=dec)
    vae.set_img(frames)
    lat = vae.encode()
    pre = OE.preprocess(frames, hfov=cfg.sensor.hfov, vfov=cfg.sensor.vfov)
    lat_ref = OE.encode(enc, pre)
    assert relmax(lat, lat_ref) <= 1e-4, relmax(lat, lat_ref)
    img = vae.decode()
    ref = OD.decode(dec, lat_ref, (H, W))
    assert img.shape == (3, H, W) and relmax(img, ref) <= 2 * IMG_TOL, relmax(img, ref)
