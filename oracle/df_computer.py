"""Oracle: ground-truth distance-field labels of points against range / depth images, numpy / torch-CPU float32.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Restates ``sdf_nmpc/utils/df_computer.py``:
  * ``generate_dist_grid`` (:31-53)                   with the reference's own torch expressions (repeat / repeat_interleave), on the CPU
  * ``get_sdf`` + ``_get_mindist_occgrid`` (:179-221) on top of ``oracle.colcheck`` ('extrapolate', safe ball 0)
  * ``get_udf`` + ``_kernel_pixel_wise_udf`` (:84-176) statement by statement, quirks included (see ``csrc/df_api.cu``)
Parity unpinned, like ``oracle/colcheck.py``: ``warp-lang`` is not installable offline, so no reference outputs exist;
``tests/test_gpu_df.py`` adds geometric known answers (a wall at a known distance).
"""
import numpy as np

from . import colcheck as OC

GRID_PARAMS = [(0, 0.1, 0.01), (0.1, 0.2, 0.02), (0.2, 0.3, 0.03), (0.3, 0.5, 0.05), (0.5, 1, 0.1)]
MIN_DF, MAX_DF = np.float32(-0.3), np.float32(1.0)


def dist_grid(grid_params=GRID_PARAMS):
    import torch
    dist, grid = [], []
    for dmin, dmax, step in grid_params:
        nb = int(2. * dmax / step) + 1
        c = torch.linspace(-dmax, dmax, nb)
        g = torch.stack([c.repeat(nb).repeat(nb), c.repeat_interleave(nb).repeat(nb), c.repeat_interleave(nb).repeat_interleave(nb)], dim=1)
        d = torch.linalg.norm(g, dim=1)
        idx = (d > dmin) * (d <= dmax)
        grid.append(g[idx]); dist.append(d[idx])
    return torch.cat(dist).numpy(), torch.cat(grid).numpy()


def colcheck_vec(imgs, pts, p2i, dmax, hfov, vfov, safe_ball, outside, is_depth, is_spherical):
    """Vectorised form of ``oracle.colcheck.check`` (same float32 statements; tests/test_oracle_df.py holds the two equal)."""
    f = np.float32
    imgs = np.asarray(imgs, f); p = np.asarray(pts, f)
    B, H, W = imgs.shape
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    ln = np.sqrt(x * x + y * y + z * z, dtype=f)
    val = x if is_depth else ln
    az = np.arctan2(y, x, dtype=f)
    el = np.arctan2(z, np.sqrt(x * x + y * y, dtype=f), dtype=f) if is_spherical else np.arctan2(z, x, dtype=f)
    out_fov = np.zeros(len(p), bool)
    if outside == 2:
        az, el = np.clip(az, f(-hfov), f(hfov)), np.clip(el, f(-vfov), f(vfov))
    else:
        out_fov = (np.abs(az) >= f(hfov)) | (np.abs(el) >= f(vfov))
    with np.errstate(invalid='ignore', over='ignore'):
        if is_spherical:
            u = (f(W) / f(2) * (f(1) - az / f(hfov))).astype(np.int64)
            v = (f(H) / f(2) * (f(1) - el / f(vfov))).astype(np.int64)
        else:
            u = (f(W) / f(2) * (f(1) - np.tan(az, dtype=f) / np.tan(f(hfov), dtype=f))).astype(np.int64)
            v = (f(H) / f(2) * (f(1) - np.tan(el, dtype=f) / np.tan(f(vfov), dtype=f))).astype(np.int64)
    u, v = np.clip(u, 0, W - 1), np.clip(v, 0, H - 1)
    hit = val >= imgs[np.asarray(p2i, np.int64), v, u] * f(dmax)
    col = np.where(val >= f(dmax), True, np.where(out_fov, outside == 1, hit))
    return np.where(ln > f(safe_ball), col, False)


def get_sdf(imgs, pts, p2i, dmax, hfov, vfov, is_depth, is_spherical, grid=None):
    f = np.float32
    dist, grid = dist_grid() if grid is None else grid
    pts = np.asarray(pts, f); p2i = np.asarray(p2i, np.int64)
    args = (dmax, hfov, vfov, 0, 2, is_depth, is_spherical)
    inside = colcheck_vec(imgs, pts, p2i, *args)
    sign = np.where(inside, f(-1), f(1))
    sdf = np.empty(len(pts), f); grad = np.zeros((len(pts), 3), f)
    for t in range(len(pts)):
        occ = colcheck_vec(imgs, pts[t] + grid, np.full(len(grid), p2i[t]), *args)
        d = np.where(occ != inside[t], dist, MAX_DF)
        g = int(np.argmin(d))                       # first of equal minima, like torch.min on the CPU
        v = np.clip(sign[t] * d[g], MIN_DF, MAX_DF)
        sdf[t] = v
        if v != MIN_DF and v != MAX_DF:
            grad[t] = -sign[t] * (grid[g] / np.sqrt((grid[g] * grid[g]).sum(dtype=f), dtype=f))
    return sdf, grad


def minpool_nonzero(imgs, k, dmax):
    imgs = np.asarray(imgs, np.float32)
    B, H, W = imgs.shape
    win = imgs.reshape(B, H // k, k, W // k, k).transpose(0, 1, 3, 2, 4).reshape(B, H // k, W // k, k * k)
    any_nz = (win != 0).any(-1)
    m = np.where(win == 0, np.float32(dmax), win).min(-1)
    return np.where(any_nz, m, np.float32(0))


def get_udf(imgs, pts, p2i, dmax, hfov, vfov, is_depth, is_spherical, pool=5):
    f = np.float32
    mp = minpool_nonzero(imgs, pool, dmax)
    B, H, W = mp.shape
    pts = np.asarray(pts, f)
    ph = (np.arange(H * W) % W).astype(f); pv = (np.arange(H * W) // W).astype(f)
    if is_spherical:
        x = np.ones(H * W, f)
        y = np.tan(f(hfov), dtype=f) * (f(1) - f(2) * ph / f(W))
        z = np.tan(f(vfov), dtype=f) * (f(1) - f(2) * pv / f(H))
    else:
        az = f(hfov) * (f(1) - f(2) * ph / f(W)); el = f(vfov) * (f(1) - f(2) * pv / f(H))
        x = np.cos(el, dtype=f) * np.cos(az, dtype=f); y = np.cos(el, dtype=f) * np.sin(az, dtype=f); z = np.sin(el, dtype=f)
    dirs = np.stack([x, y, z], 1)
    udf = np.empty(len(pts), f); grad = np.zeros((len(pts), 3), f)
    for t, p in enumerate(pts):
        pix = dirs * mp[int(p2i[t])].reshape(-1, 1) * f(dmax)
        rel = (pix - p).astype(f)
        dp = np.sqrt((rel * rel).sum(1, dtype=f), dtype=f)
        d_bg = f(dmax) - p[0] if is_depth else f(dmax) - np.sqrt((p * p).sum(dtype=f), dtype=f)
        wall = dp > d_bg
        d = np.where(x == 0, f(dmax), np.where(wall, d_bg, dp)).astype(f)
        rel = np.where((wall & (x != 0))[:, None], np.array([dmax, p[1], p[2]], f)[None], rel)
        q = int(np.argmin(d))
        v = np.clip(d[q], f(0), MAX_DF)
        udf[t] = v
        if v != MAX_DF:
            grad[t] = -(rel[q] / np.sqrt((rel[q] * rel[q]).sum(dtype=f), dtype=f))
    return udf, grad
