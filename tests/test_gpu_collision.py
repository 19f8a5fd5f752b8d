"""GPU collision labels (sdfnmpc_colcheck, mirror of ColChecker.check_image_points) against the numpy restatement of the
reference's Warp kernel and against geometric known answers of the synthetic scene."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HF, VF, DMAX = 0.7592, 0.4903, 5.0


@pytest.mark.parametrize('spherical,is_depth,outside', [(True, False, 'free'), (False, True, 'col'), (True, False, 'extrapolate')])
def test_colcheck_matches_oracle(spherical, is_depth, outside, cuda_device):
    from oracle import colcheck as OC
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.collision import ColChecker
    frames = np.stack([synth.depth_frame(seed=3 + i, spherical=spherical) for i in range(3)]).astype(np.float32)
    imgs = np.clip(frames / DMAX, 0, 1).astype(np.float32)          # normalised by dmax, like the preprocessing output
    rng = np.random.default_rng(5)
    n = 3 * 2000
    pts = np.stack([rng.uniform(-0.5, 6.0, n), rng.uniform(-3, 3, n), rng.uniform(-2, 2, n)], 1).astype(np.float32)
    pts[:50] *= 0.02                                                 # inside the safe ball
    p2i = rng.integers(0, 3, n).astype(np.int32)
    cc = ColChecker(DMAX, HF, VF, 0.3, is_depth=is_depth, is_spherical=spherical, outside=outside, device=cuda_device)
    got = cc.check_image_points(imgs, pts, p2i).cpu().numpy()
    want = OC.check(imgs, pts, p2i, DMAX, HF, VF, 0.3, cc.outside, is_depth, spherical)
    # float32 atan2 / tan differ by an ulp between libm and CUDA: a point exactly on a pixel or frustum edge may flip
    assert (got != want).mean() <= 2e-3, (got != want).mean()
    assert 0.05 < got.mean() < 0.95
    # default point-to-image map: points split evenly over the images
    got2 = cc.check_image_points(imgs, pts).cpu().numpy()
    want2 = OC.check(imgs, pts, np.repeat(np.arange(3), n // 3), DMAX, HF, VF, 0.3, cc.outside, is_depth, spherical)
    assert (got2 != want2).mean() <= 2e-3


def test_colcheck_known_answers(cuda_device):
    """A wall at 2 m seen by a spherical range sensor: points in front of it are free, points behind it collide, points
    past dmax collide, points inside the safe ball or outside the frustum ('free' mode) do not."""
    from sdf_nmpc_b200.collision import ColChecker
    H, W = 90, 160
    img = np.full((H, W), 2.0 / DMAX, np.float32)
    cc = ColChecker(DMAX, HF, VF, 0.3, is_depth=False, is_spherical=True, outside='free', device=cuda_device)
    pts = np.array([[1.0, 0, 0], [1.9, 0.1, 0.05], [2.1, 0, 0], [3.0, 0.5, -0.2], [5.5, 0, 0], [0.1, 0.05, 0], [1.0, 3.0, 0], [-1.0, 0, 0]], np.float32)
    got = cc.check_image_points(img, pts).cpu().numpy()
    assert got.tolist() == [False, False, True, True, True, False, False, False]
    cc_col = ColChecker(DMAX, HF, VF, 0.3, is_depth=False, is_spherical=True, outside='col', device=cuda_device)
    assert cc_col.check_image_points(img, pts).cpu().numpy().tolist() == [False, False, True, True, True, False, True, True]
    assert cc.check_image_points(img, np.zeros((0, 3), np.float32)).numel() == 0
