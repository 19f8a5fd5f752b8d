"""Oracle, compiled flavour: the CPU baseline of bench.py (``cpu_baseline`` leg and ``--impl reference``).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  The reference's CPU path is acados 0.3.1 (C) + L4CasADi 1.4.1
(libtorch), neither of which can be built offline; the numpy oracle (``oracle/rti.py``) restates its semantics but runs a
dense condensed IPM in Python at ~13 solves/s per core, which flatters any GPU number.  This module is the faster, more
honest CPU arm:

  * dynamics + sensitivities, cost / constraint linearisation and the structure-exploiting interior-point QP run in C++:
    ``tests/host_emul/emul.cpp`` compiled by ``__graft_entry__.build()`` into ``oracle/_ref/libhost_emul.so`` (the scalar
    statement of the solver, ``csrc/qp_core.cuh`` + ``linearize_core.cuh``, one lane per instance) - pinned against the numpy
    oracle by ``tests/test_host_emulation.py``;
  * the neural SDF runs the way L4CasADi runs it (``sdf_nmpc/gen_model.py:27,38-39``): libtorch on the CPU, one thread,
    batch 1, a TorchScript forward and a traced ``jacrev`` per shooting node (``oracle/torch_ref.py``).

``kind`` stays "port": a CPU restatement of the reference's path, not its binaries.
"""
import ctypes
import os

import numpy as np

from . import models as MODELS
from . import rti

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'oracle', '_ref', 'libhost_emul.so')


def load():
    if not os.path.exists(LIB):
        raise FileNotFoundError(f'{LIB}: run __graft_entry__.build() first')
    return ctypes.CDLL(LIB)


def _P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CxxRti:
    """One SQP-RTI iteration per call for a block of instances; state (warm start) lives in ``xe`` / ``ue``."""

    def __init__(self, net, cfg, prob, sdf='torch'):
        self.lib, self.net, self.cfg, self.prob = load(), net, cfg, prob
        B, N = prob['x0'].shape[0], cfg['N']
        self.B, self.N = B, N
        c = np.ascontiguousarray
        self.pose = c(np.concatenate([prob['W_p_Co'], prob['W_R_Co'].reshape(B, 9)], 1))
        self.mid, self.nyN = MODELS.MODEL_ID[cfg['model']], MODELS.get(cfg['model']).nyN
        self.mcv = np.array([cfg['g'], cfg['gamma'], cfg['roll'], cfg['pitch'], cfg['wz'], cfg['mass'],
                             *(cfg['B_R_C'].T @ cfg['B_p_C'] + np.array([cfg['fov_off'], 0, 0])),
                             cfg['ax'], cfg['ay'], cfg['az'], cfg['tau_roll'], cfg['tau_pitch']])
        self.lh, self.uh = rti.h_bounds(cfg, net['max_df'])
        self.zs, self.Zs = rti.slack_weights(cfg)
        self.AB = np.zeros((B, N, 10, 14)); self.bd = np.zeros((B, N, 10)); self.J = np.zeros((B, N + 1, 11, 14))
        self.yr = np.zeros((B, N + 1, 11)); self.C = np.zeros((B, N + 1, 3, 14)); self.hv = np.zeros((B, N + 1, 3))
        self.qp = np.zeros((B, N + 1, 3), np.float32)
        self.arr = {k: c(prob[k]) for k in ('yref', 'yrefN', 'qd', 'flag', 'W', 'WN')}
        self.xe, self.ue = (c(a) for a in rti.init_traj(prob['x0'], cfg))
        self.u0 = np.zeros((B, 4)); self.lamh = np.zeros((B, N, 8)); self.lams = np.zeros((B, N + 1, 6))
        self.st = np.zeros(B, np.int32); self.its = np.zeros(B, np.int32)
        self.sdf_d = np.zeros((B, N + 1), np.float32); self.sdf_g = np.zeros((B, N + 1, 3), np.float32)
        self.sdf_mode = sdf
        if sdf == 'torch':
            import torch
            from . import torch_ref
            torch.set_num_threads(1)                       # gen_model.py:27
            self.torch = torch
            self.fwd, self.jac = torch_ref.sdf_l4c_pair(net)
            self.xin = torch.zeros(1, 3 + int(net['size_latent']))

    def _sdf(self):
        if self.sdf_mode != 'torch':
            from . import sdf as SDF
            d, g = SDF.value_and_grad(self.net, self.qp.reshape(-1, 3), np.repeat(self.prob['latent'], self.N + 1, 0), dtype=np.float32)
            self.sdf_d[:] = d.reshape(self.B, -1); self.sdf_g[:] = g.reshape(self.B, -1, 3)
            return
        torch = self.torch
        with torch.no_grad():
            for b in range(self.B):
                self.xin[0, 3:] = torch.from_numpy(self.prob['latent'][b])
                for k in range(self.N + 1):
                    self.xin[0, :3] = torch.from_numpy(self.qp[b, k])
                    self.sdf_d[b, k] = float(self.fwd(self.xin)[0, 0])          # sdf_l4c
                    self.sdf_g[b, k] = self.jac(self.xin)[0, 0, :3].numpy()      # jac_sdf_l4c (1 x 131, state columns kept)

    def step(self, x0):
        cfg, a, D = self.cfg, self.arr, ctypes.c_double
        x0 = np.ascontiguousarray(x0)
        self.lib.emul_linearize(self.mid, self.nyN, self.B, self.N, _P(self.mcv), _P(self.xe), _P(self.ue), _P(a['yref']), _P(a['yrefN']),
                                _P(a['qd']), _P(self.pose), _P(a['flag']), _P(cfg['dt']), _P(self.AB), _P(self.bd), _P(self.J), _P(self.yr),
                                _P(self.C), _P(self.hv), _P(self.qp), _P(x0))
        self._sdf()
        self.lib.emul_qp(self.B, self.N, 1, 100, self.nyN, D(cfg.get('sdf_cost_w', 0.0)), D(cfg['lm']), D(1e-8), D(1e-7), D(self.net['max_df']),
                         _P(a['flag']), _P(self.AB), _P(self.bd), _P(self.J), _P(self.yr), _P(self.C), _P(self.hv), _P(a['W']), _P(a['WN']),
                         _P(cfg['dt']), _P(cfg['lbu']), _P(cfg['ubu']), _P(self.lh), _P(self.uh), _P(self.zs), _P(self.Zs), _P(self.sdf_d),
                         _P(self.sdf_g), _P(self.pose), _P(x0), _P(self.xe), _P(self.ue), _P(self.u0), _P(self.lamh), _P(self.lams),
                         _P(self.st), _P(self.its))
        return self.u0, self.st, self.its
