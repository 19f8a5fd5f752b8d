"""Ground-truth collision labels in range / depth images - mirror of ``sdf_nmpc.utils.collision_checker.ColChecker``
(``sdf_nmpc/utils/collision_checker.py:6-125``; the reference runs it as a Warp kernel).  Scores closed-loop rollouts
against the synthetic depth scene.  One CUDA thread per point, through ``sdfnmpc_colcheck``."""
import ctypes as C

import torch

from . import _lib
from .sdf import _ptr, _stream


class ColChecker:
    """outside -- 'col', 'free', or 'extrapolate'."""

    def __init__(self, dmax, hfov, vfov, safe_ball_size, is_depth=False, is_spherical=False, outside='free', device=None):
        assert outside in ['free', 'col', 'extrapolate']
        self.dmax, self.hfov, self.vfov, self.safe_ball_size = float(dmax), float(hfov), float(vfov), float(safe_ball_size)
        self.is_depth, self.is_spherical = bool(is_depth), bool(is_spherical)
        self.outside = 0 if outside == 'free' else 1 if outside == 'col' else 2
        self.device = torch.device(device if device is not None else 'cuda:0')
        self.lib = _lib.load()

    def check_image_points(self, imgs, points, p_to_i=None):
        """imgs [B,H,W] or [H,W], normalised by dmax; points [n,3] un-normalised Cartesian coordinates in the sensor frame;
        p_to_i [n] image index per point (default: points split evenly, in order, over the images).  Returns bool [n]."""
        imgs = torch.as_tensor(imgs, dtype=torch.float32, device=self.device)
        if imgs.ndim == 2:
            imgs = imgs.unsqueeze(0)
        if imgs.ndim != 3:
            raise AssertionError('input image must have size [B, H, W] or [H, W]')
        imgs = imgs.contiguous()
        pts = torch.as_tensor(points, dtype=torch.float32, device=self.device).reshape(-1, 3).contiguous()
        n = pts.shape[0]
        p2i = None
        if p_to_i is not None:
            p2i = torch.as_tensor(p_to_i, device=self.device).to(torch.int32).contiguous()
        elif n % imgs.shape[0] != 0:
            raise AssertionError('without p_to_i the number of points must be a multiple of the number of images')
        out = torch.zeros(n, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.sdfnmpc_colcheck(_ptr(imgs), imgs.shape[0], imgs.shape[1], imgs.shape[2], _ptr(pts),
                                             _ptr(p2i) if p2i is not None else C.c_void_p(0), n, self.dmax, self.hfov, self.vfov,
                                             self.safe_ball_size, self.outside, int(self.is_depth), int(self.is_spherical), _ptr(out),
                                             _stream()), 'sdfnmpc_colcheck')
        return out.bool()


class DfComputer:
    """Ground-truth signed / unsigned distance-field labels of points against depth or range images - mirror of
    ``sdf_nmpc.utils.df_computer.DfComputer`` (``utils/df_computer.py:7-221``; constructor and ``get_df`` keep their signatures).
    The reference evaluates ``ColChecker`` on a [points x voxels] cloud in batches of ``batch_size`` points and reduces with
    torch; here ``sdfnmpc_df_signed`` / ``sdfnmpc_df_unsigned`` give a CTA to each point and reduce over its voxels / pixels
    on chip, so ``batch_size`` is accepted and unused."""

    # (dmin, dmax, step) of the concentric voxel shells (df_computer.py:29)
    GRID_PARAMS = [(0, 0.1, 0.01), (0.1, 0.2, 0.02), (0.2, 0.3, 0.03), (0.3, 0.5, 0.05), (0.5, 1, 0.1)]
    POOL = 5       # min-pool window of the unsigned field (df_computer.py:140)

    def __init__(self, signed, dmax, hfov, vfov, max_df, is_depth=False, is_spherical=False, batch_size=5000, device=None):
        self.signed, self.is_depth, self.is_spherical = bool(signed), bool(is_depth), bool(is_spherical)
        self.hfov, self.vfov, self.dmax = float(hfov), float(vfov), float(dmax)
        self.min_df, self.max_df = -0.3, 1.0           # fixed in the reference (df_computer.py:16-17), whatever max_df is passed
        self.batch_size = batch_size
        self.device = torch.device(device if device is not None else 'cuda:0')
        self.lib = _lib.load()
        if self.signed:
            self.distances, self.grid = self.generate_dist_grid(self.GRID_PARAMS)

    def generate_dist_grid(self, grid_params):
        """Voxel offsets of the shells dmin < |offset| <= dmax at resolution step (df_computer.py:31-53): x runs fastest, then y,
        then z, shells concatenated in order - the order decides which of several equidistant voxels gives the gradient."""
        grids, dists = [], []
        for dmin, dmax, step in grid_params:
            n = int(2. * dmax / step) + 1
            c = torch.linspace(-dmax, dmax, n, device=self.device)
            z, y, x = torch.meshgrid(c, c, c, indexing='ij')
            g = torch.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], dim=1)
            d = torch.linalg.norm(g, dim=1)
            keep = (d > dmin) & (d <= dmax)
            grids.append(g[keep])
            dists.append(d[keep])
        return torch.cat(dists).contiguous(), torch.cat(grids).contiguous()

    def _prep(self, imgs, points, p_to_i):
        imgs = torch.as_tensor(imgs, dtype=torch.float32, device=self.device)
        if imgs.ndim == 2:
            imgs = imgs.unsqueeze(0)
        if imgs.ndim != 3:
            raise AssertionError('input image must have size [B, H, W] or [H, W]')
        pts = torch.as_tensor(points, dtype=torch.float32, device=self.device).reshape(-1, 3).contiguous()
        p2i = None
        if p_to_i is not None:
            p2i = torch.as_tensor(p_to_i, device=self.device).to(torch.int32).contiguous()
        elif pts.shape[0] % imgs.shape[0] != 0:
            raise AssertionError('without p_to_i the number of points must be a multiple of the number of images')
        return imgs.contiguous(), pts, p2i

    def get_df(self, imgs, points, p_to_i=None):
        """imgs [B,H,W] or [H,W] normalised by dmax; points [n,3] un-normalised, sensor frame; p_to_i [n] image per point
        (default: even split in order).  Returns (df [n], grad [n,3]) device tensors."""
        imgs, pts, p2i = self._prep(imgs, points, p_to_i)
        return self.get_sdf(imgs, pts, p2i) if self.signed else self.get_udf(imgs, pts, p2i)

    def get_sdf(self, imgs, points, p_to_i):
        n = points.shape[0]
        sdf = torch.empty(n, dtype=torch.float32, device=self.device)
        grad = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.sdfnmpc_df_signed(_ptr(imgs), imgs.shape[0], imgs.shape[1], imgs.shape[2], _ptr(points),
                                              _ptr(p_to_i) if p_to_i is not None else C.c_void_p(0), n, _ptr(self.grid),
                                              _ptr(self.distances), self.grid.shape[0], self.dmax, self.hfov, self.vfov,
                                              int(self.is_depth), int(self.is_spherical), self.min_df, self.max_df, _ptr(sdf), _ptr(grad),
                                              _stream()), 'sdfnmpc_df_signed')
        return sdf, grad

    def get_udf(self, imgs, points, p_to_i):
        n = points.shape[0]
        B, H, W = imgs.shape
        assert H % self.POOL == 0 and W % self.POOL == 0
        pooled = torch.empty(B, H // self.POOL, W // self.POOL, dtype=torch.float32, device=self.device)
        udf = torch.empty(n, dtype=torch.float32, device=self.device)
        grad = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.sdfnmpc_df_unsigned(_ptr(imgs), B, H, W, self.POOL, _ptr(pooled), _ptr(points),
                                                _ptr(p_to_i) if p_to_i is not None else C.c_void_p(0), n, self.dmax, self.hfov, self.vfov,
                                                int(self.is_depth), int(self.is_spherical), self.max_df, _ptr(udf), _ptr(grad), _stream()),
                   'sdfnmpc_df_unsigned')
        return udf, grad
