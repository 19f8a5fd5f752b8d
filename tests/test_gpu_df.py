"""GPU distance-field labels (``DfComputer.get_df`` through sdfnmpc_df_signed / sdfnmpc_df_unsigned, mirror of
sdf_nmpc/utils/df_computer.py) against the numpy restatement (oracle/df_computer.py; parity unpinned: the reference's Warp kernels
cannot run offline) and geometric known answers.  Labels come from float32 atan2 / tan / sin / cos, which differ by an ulp
between libm and CUDA, so a point whose nearest differing voxel sits exactly on a pixel edge may pick the next voxel: the test
bounds the FRACTION of points that differ (like tests/test_gpu_collision.py) and requires the rest to agree to float32 rounding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HF, VF, DMAX = 0.7592, 0.4903, 5.0


def _scene(spherical, n_img=2, n=300, seed=5):
    from sdf_nmpc_b200 import synth
    frames = np.stack([synth.depth_frame(seed=3 + i, spherical=spherical) for i in range(n_img)]).astype(np.float32)
    imgs = np.clip(frames / DMAX, 0, 1).astype(np.float32)
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(0.2, 5.5, n), rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n)], 1).astype(np.float32)
    return imgs, pts, rng.integers(0, n_img, n).astype(np.int32)


@pytest.mark.parametrize('spherical,is_depth', [(True, False), (False, True)])
def test_signed_field_matches_oracle(spherical, is_depth, cuda_device):
    from oracle import df_computer as OD
    from sdf_nmpc_b200.collision import DfComputer
    imgs, pts, p2i = _scene(spherical)
    df = DfComputer(True, DMAX, HF, VF, 1.0, is_depth=is_depth, is_spherical=spherical, device=cuda_device)
    sdf, grad = df.get_df(imgs, pts, p2i)
    sdf, grad = sdf.cpu().numpy(), grad.cpu().numpy()
    dist, grid = df.distances.cpu().numpy(), df.grid.cpu().numpy()
    # the voxel table: same construction as the reference's repeat / repeat_interleave form (CPU linspace may differ by an ulp)
    odist, ogrid = OD.dist_grid()
    assert abs(len(grid) - len(ogrid)) <= 8 and 15000 < len(grid) < 20000
    if len(grid) == len(ogrid):
        assert np.abs(grid - ogrid).max() <= 1e-6 and np.abs(dist - odist).max() <= 1e-6
    want, wgrad = OD.get_sdf(imgs, pts, p2i, DMAX, HF, VF, is_depth, spherical, (dist, grid))
    same = np.abs(sdf - want) <= 1e-6
    assert same.mean() >= 0.97, same.mean()
    assert np.abs(sdf - want).max() <= 0.11                      # a flipped label moves the answer by a voxel at most
    assert np.abs(grad[same] - wgrad[same]).max() <= 1e-5 or (np.abs(grad[same] - wgrad[same]).max(1) <= 1e-5).mean() >= 0.97
    assert sdf.min() == np.float32(-0.3) and sdf.max() == 1.0 and len(np.unique(sdf)) > 20
    nz = np.abs(grad).sum(1) > 0
    assert np.allclose(np.linalg.norm(grad[nz], axis=1), 1, atol=1e-5) and not grad[(sdf == 1.0) | (sdf == np.float32(-0.3))].any()
    # default point-to-image map + single image [H, W]
    s1, _ = df.get_df(imgs[0], pts[p2i == 0])
    assert np.array_equal(s1.cpu().numpy(), sdf[p2i == 0])


@pytest.mark.parametrize('spherical,is_depth', [(True, False), (False, True)])
def test_unsigned_field_matches_oracle(spherical, is_depth, cuda_device):
    from oracle import df_computer as OD
    from sdf_nmpc_b200.collision import DfComputer
    imgs, pts, p2i = _scene(spherical, n=400, seed=9)
    imgs[1, 100:140, 200:260] = 0.0                              # an invalid patch (whole pooling windows of zeros)
    df = DfComputer(False, DMAX, HF, VF, 1.0, is_depth=is_depth, is_spherical=spherical, device=cuda_device)
    udf, grad = df.get_df(imgs, pts, p2i)
    udf, grad = udf.cpu().numpy(), grad.cpu().numpy()
    want, wgrad = OD.get_udf(imgs, pts, p2i, DMAX, HF, VF, is_depth, spherical)
    assert np.abs(udf - want).max() <= 2e-5, np.abs(udf - want).max()
    close = np.abs(grad - wgrad).max(1) <= 1e-3
    assert close.mean() >= 0.98, close.mean()                    # equidistant pixels may swap the arg-min
    assert 0.1 < (udf < 1.0).mean() < 0.95 and udf.min() >= 0 and udf.max() == 1.0


def test_constant_range_known_answers(cuda_device):
    """Range image of a sphere of radius 2 m around a spherical sensor: signed distance of a point at range r is 2 - r up to the
    voxel size of its shell; saturation at 1 / -0.3 with zero gradient; the gradient points away from the surface."""
    from sdf_nmpc_b200.collision import DfComputer
    img = np.full((90, 160), 2.0 / DMAX, np.float32)
    pts = np.array([[1.9, 0, 0], [1.5, 0.2, 0.1], [2.15, 0, 0], [0.5, 0, 0], [2.6, 0.1, 0]], np.float32)
    df = DfComputer(True, DMAX, HF, VF, 1.0, is_depth=False, is_spherical=True, device=cuda_device)
    sdf, grad = (a.cpu().numpy() for a in df.get_df(img, pts))
    r = np.linalg.norm(pts, axis=1)
    assert abs(sdf[0] - 0.1) <= 0.011 and abs(sdf[1] - (2 - r[1])) <= 0.06 and abs(sdf[2] + 0.15) <= 0.03
    assert sdf[3] == 1.0 and sdf[4] == np.float32(-0.3) and not grad[3].any() and not grad[4].any()
    assert grad[0][0] < -0.9 and grad[2][0] < -0.9
    assert df.get_df(img, np.zeros((0, 3), np.float32))[0].numel() == 0
