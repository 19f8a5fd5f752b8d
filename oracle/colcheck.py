"""Oracle: ground-truth collision labels of points against range / depth images, numpy float32.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Restates the Warp kernel ``ColChecker._kernel_colcheck``
(``sdf_nmpc/utils/collision_checker.py:23-91``) statement by statement.  Parity unpinned: ``warp-lang`` is not installable
offline, so no reference outputs exist; the function is short enough to check against the source by eye, and
``tests/test_gpu_collision.py`` adds geometric known answers."""
import numpy as np


def check(imgs, pts, p2i, dmax, hfov, vfov, safe_ball, outside, is_depth, is_spherical):
    imgs = np.asarray(imgs, np.float32)
    pts = np.asarray(pts, np.float32)
    B, H, W = imgs.shape
    f = np.float32
    out = np.zeros(len(pts), bool)
    for t, p in enumerate(pts):
        ln = np.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2], dtype=f)
        if not ln > f(safe_ball):
            continue
        val = p[0] if is_depth else ln
        if val >= f(dmax):
            out[t] = True
            continue
        az = np.arctan2(p[1], p[0], dtype=f)
        el = np.arctan2(p[2], np.sqrt(p[0] * p[0] + p[1] * p[1], dtype=f), dtype=f) if is_spherical else np.arctan2(p[2], p[0], dtype=f)
        if outside == 2:
            az, el = np.clip(az, f(-hfov), f(hfov)), np.clip(el, f(-vfov), f(vfov))
        elif abs(az) >= f(hfov) or abs(el) >= f(vfov):
            out[t] = outside == 1
            continue
        if is_spherical:
            u = int(f(W) / f(2) * (f(1) - az / f(hfov)))
            v = int(f(H) / f(2) * (f(1) - el / f(vfov)))
        else:
            u = int(f(W) / f(2) * (f(1) - np.tan(az, dtype=f) / np.tan(f(hfov), dtype=f)))
            v = int(f(H) / f(2) * (f(1) - np.tan(el, dtype=f) / np.tan(f(vfov), dtype=f)))
        u, v = min(max(u, 0), W - 1), min(max(v, 0), H - 1)
        if val >= imgs[int(p2i[t]), v, u] * f(dmax):
            out[t] = True
    return out
