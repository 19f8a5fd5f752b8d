"""CPU checks of the distance-field oracle (oracle/df_computer.py): its vectorised collision labels equal the statement-by-statement
loop of oracle/colcheck.py, and the labeller gives the geometric answers of a flat wall."""
import numpy as np

HF, VF, DMAX = 0.7592, 0.4903, 5.0


def test_vectorised_colcheck_equals_loop():
    from oracle import colcheck as OC, df_computer as OD
    from sdf_nmpc_b200 import synth
    frames = np.stack([synth.depth_frame(seed=3 + i, spherical=True) for i in range(2)]).astype(np.float32)
    imgs = np.clip(frames / DMAX, 0, 1).astype(np.float32)
    rng = np.random.default_rng(5)
    n = 600
    pts = np.stack([rng.uniform(-0.5, 6, n), rng.uniform(-3, 3, n), rng.uniform(-2, 2, n)], 1).astype(np.float32)
    pts[:20] *= 0.02
    p2i = rng.integers(0, 2, n).astype(np.int32)
    for outside in (0, 1, 2):
        for sph, dep in ((True, False), (False, True)):
            a = OC.check(imgs, pts, p2i, DMAX, HF, VF, 0.3, outside, dep, sph)
            b = OD.colcheck_vec(imgs, pts, p2i, DMAX, HF, VF, 0.3, outside, dep, sph)
            assert np.array_equal(a, b)


def test_wall_known_answers():
    """Spherical range image of a sphere of radius 2 m around the sensor (constant range): a point at range r has signed distance
    2 - r up to the voxel resolution of its shell; deep inside / far outside saturate at -0.3 / 1 with zero gradient."""
    from oracle import df_computer as OD
    img = np.full((1, 90, 160), 2.0 / DMAX, np.float32)
    pts = np.array([[1.9, 0, 0], [1.5, 0.2, 0.1], [2.15, 0, 0], [0.5, 0, 0], [2.6, 0.1, 0]], np.float32)
    sdf, grad = OD.get_sdf(img, pts, np.zeros(5, int), DMAX, HF, VF, False, True)
    r = np.linalg.norm(pts, axis=1)
    assert abs(sdf[0] - 0.1) <= 0.011 and abs(sdf[1] - (2 - r[1])) <= 0.06 and abs(sdf[2] + 0.15) <= 0.03
    assert sdf[3] == 1.0 and sdf[4] == np.float32(-0.3) and not grad[3].any() and not grad[4].any()
    assert grad[0][0] < -0.9 and grad[2][0] < -0.9      # free: away from the wall (-x); inside: towards free space (-x)
    udf, ug = OD.get_udf(np.full((1, 90, 160), 2.0 / DMAX, np.float32), pts[:2], np.zeros(2, int), DMAX, HF, VF, False, False)
    assert abs(udf[0] - 0.1) <= 0.02 and ug[0][0] < -0.9


def test_minpool_ignores_zeros():
    from oracle import df_computer as OD
    img = np.ones((1, 10, 10), np.float32) * 0.8
    img[0, :5, :5] = 0.0                   # an all-invalid window stays 0
    img[0, 0, 7] = 0.0; img[0, 1, 8] = 0.3   # zeros are skipped, the smallest valid value wins
    mp = OD.minpool_nonzero(img, 5, DMAX)
    assert mp.shape == (1, 2, 2) and mp[0, 0, 0] == 0 and mp[0, 0, 1] == np.float32(0.3) and mp[0, 1, 0] == np.float32(0.8)
