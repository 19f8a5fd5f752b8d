"""Batched neural-SDF evaluator: the drop-in for the ``l4casadi.L4CasADi(sdf, ..., with_jacobian=True)``
object of ``sdf_nmpc/gen_model.py:38-39`` - value and input-gradient, for many points and latents at once.
torch is used for device memory and streams only; the arithmetic runs in ``libsdfnmpc_b200.so``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .weights import SDF_KEYS, sdf_desc_fields


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SdfEvaluator:
    """``net`` is a dict as produced by ``weights.sdf_from_module`` / ``synth.make_sdf_net``."""

    def __init__(self, net, device='cuda:0'):
        if not torch.cuda.is_available():
            raise _lib.SdfNmpcError('CUDA is not available: sdf_nmpc_b200 has no CPU path')
        self.lib = _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.net = net
        f, self.E = sdf_desc_fields(net)
        self.size_latent = f['size_latent']
        self.max_df = f['max_df']
        desc = _lib.SdfDesc(f['n_dirs'], f['nb_freqs'], (C.c_int32 * 4)(*f['hidden']), f['size_latent'], f['res_state'],
                            f['res_latent'], f['act'], f['w0'], f['max_df'])
        arrs = [np.ascontiguousarray(net['dirs'], np.float32)] + [np.ascontiguousarray(net[k], np.float32) for k in SDF_KEYS]
        self.handle = C.c_void_p()
        _lib.check(self.lib.sdfnmpc_sdf_create(C.byref(self.handle), C.byref(desc), *[a.ctypes.data_as(C.c_void_p) for a in arrs]),
                   'sdfnmpc_sdf_create')
        self.n_inst = 0

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h:
            self.lib.sdfnmpc_sdf_destroy(h)
            self.handle = None

    def set_latents(self, latents):
        """latents: [n_inst, L] float32 CUDA tensor (or [L])."""
        z = torch.as_tensor(latents, dtype=torch.float32, device=self.device).reshape(-1, max(self.size_latent, 1)).contiguous()
        _lib.check(self.lib.sdfnmpc_sdf_set_latents(self.handle, _ptr(z), z.shape[0], _stream()), 'sdfnmpc_sdf_set_latents')
        self.n_inst = z.shape[0]

    def eval(self, pts, pt2inst=None, pts_per_inst=None, want_grad=True, out=None):
        """pts [n,3] float32 CUDA tensor -> (sdf [n], grad [n,3]).  Point i uses latent ``pt2inst[i]`` if given,
        else ``i // pts_per_inst`` (default: one latent for all points)."""
        pts = pts.to(self.device, torch.float32).contiguous()
        n = pts.shape[0]
        if out is None:
            sdf = torch.empty(n, dtype=torch.float32, device=self.device)
            grad = torch.empty(n, 3, dtype=torch.float32, device=self.device) if want_grad else None
        else:
            sdf, grad = out
        if pt2inst is not None:
            pt2inst = pt2inst.to(self.device, torch.int32).contiguous()
        ppi = int(pts_per_inst) if pts_per_inst else max(n, 1)
        _lib.check(self.lib.sdfnmpc_sdf_eval(self.handle, _ptr(pts), _ptr(pt2inst), ppi, n, _ptr(sdf), _ptr(grad), _stream()),
                   'sdfnmpc_sdf_eval')
        return sdf, grad

    def eval_host(self, pts, latent):
        """numpy in / numpy out through the C ABI with host buffers (H2D + D2H inside the call)."""
        pts = np.ascontiguousarray(pts, np.float32)
        lat = np.ascontiguousarray(latent, np.float32).reshape(-1)
        n = pts.shape[0]
        sdf = np.empty(n, np.float32)
        grad = np.empty((n, 3), np.float32)
        _lib.check(self.lib.sdfnmpc_sdf_eval_host(self.handle, pts.ctypes.data_as(C.c_void_p), lat.ctypes.data_as(C.c_void_p),
                                                  n, sdf.ctypes.data_as(C.c_void_p), grad.ctypes.data_as(C.c_void_p)),
                   'sdfnmpc_sdf_eval_host')
        return sdf, grad

    # ---- CasADi external-function ABI (what acados links as libsdf_l4c.so) ----
    def bind_l4c(self):
        _lib.check(self.lib.sdfnmpc_l4c_bind(self.handle), 'sdfnmpc_l4c_bind')

    def l4c_call(self, i0):
        """Calls the exported ``sdf_l4c`` / ``jac_sdf_l4c`` symbols exactly as CasADi-generated C does."""
        i0 = np.ascontiguousarray(i0, np.float64).reshape(-1)
        val = np.zeros(1)
        jac = np.zeros(i0.size)
        DP = C.POINTER(C.c_double)
        arg = (DP * 2)(i0.ctypes.data_as(DP), val.ctypes.data_as(DP))
        res = (DP * 1)(val.ctypes.data_as(DP))
        if self.lib.sdf_l4c(arg, res, None, None, 0) != 0:
            raise _lib.SdfNmpcError('sdf_l4c failed')
        res = (DP * 1)(jac.ctypes.data_as(DP))
        if self.lib.jac_sdf_l4c(arg, res, None, None, 0) != 0:
            raise _lib.SdfNmpcError('jac_sdf_l4c failed')
        return float(val[0]), jac
