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
