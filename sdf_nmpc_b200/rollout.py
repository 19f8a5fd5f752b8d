"""Monte-Carlo closed-loop rollouts of many independent environments on one GPU (BASELINE configs[3] / configs[4]):

    depth frame -> encoder -> latent -> reference generation -> SQP-RTI (or full SQP) -> RK4 plant -> collision label

Every stage is one of the batched device paths of this package; the loop itself is host bookkeeping, like the ROS node
that drives the reference controller.  Environments are sharded over ranks by ``dist.shard_range``; per-environment result
records are gathered with one collective (``dist.gather_records``).
"""
import numpy as np
import torch

from . import synth
from .collision import ColChecker
from .nmpc import BatchedNmpc
from .ref_gen import BatchedRefGen
from .vae import VaeWrapper

REC_WIDTH = 16      # [x(10), fail steps, collision steps, min SDF seen, mean QP iterations, mean SQP iterations, distance to goal]


class Rollout:
    """Closed-loop Monte-Carlo rollout of environments [first, first + n_env) on one GPU.

    Construction does everything that happens once (synthetic scenes, handles, first encoding); ``step()`` is one control
    step for all environments: reference generation -> SQP-RTI (or full SQP) -> statistics -> RK4 plant -> collision label.
    In RTI mode with a static scene the whole step - a hundred small tensor operations around four kernels of this library -
    is captured ONCE as a CUDA graph (``capture()``) and replayed, so the bookkeeping costs one launch per step instead of
    ~1 ms of launch overhead (r1: 273 k environment-steps/s per GPU against 593 k solves/s of the solver alone)."""

    def __init__(self, cfg, sdf_net, enc_net, n_env, first=0, mode='rti', encode_every=1, sqp_tol=5e-2, sqp_max_iter=10,
                 goal_offset=(4.0, 0.0, 0.0), n_scenes=8, device='cuda:0'):
        self.cfg, self.mode, self.encode_every, self.sqp_tol, self.sqp_max_iter = cfg, mode, encode_every, sqp_tol, sqp_max_iter
        self.n_env, self.dt = n_env, float(cfg.shooting_dt()[0])
        dev = self.dev = torch.device(device)
        spherical = bool(cfg.sensor.is_spherical)
        w = synth.nmpc_instances(n_env, N=cfg.mpc.N, dt=self.dt, first=first, model=cfg.mpc.model)
        # a small pool of synthetic scenes, one per environment (seeded by the global environment index)
        pool = np.stack([synth.depth_frame(seed=200 + s, spherical=spherical) for s in range(n_scenes)]).astype(np.float32)
        scene = torch.from_numpy((np.arange(first, first + n_env) % n_scenes).astype(np.int64)).to(dev)
        self.frames = torch.from_numpy(pool).to(dev)[scene]                    # [n_env, H, W] raw sensor images
        self.vae = VaeWrapper(cfg, enc_net, max_batch=min(n_env, 256), device=device)   # encode() walks larger batches in chunks of 256
        self.mpc = BatchedNmpc(cfg, sdf_net, n_env, device)
        self.rg = BatchedRefGen(cfg)
        self.cc = ColChecker(cfg.sensor.dmax, cfg.sensor.hfov, cfg.sensor.vfov, cfg.robot.size.xy, is_depth=bool(cfg.sensor.is_depth),
                             is_spherical=spherical, outside='free', device=device)
        self.imgs_norm = torch.clamp(self.frames / float(cfg.sensor.dmax), 0, 1)    # what ColChecker expects (normalised by dmax)
        x = torch.from_numpy(w['x0']).to(dev)
        x[:, 7:] = 0.0
        self.x = x
        self.p0 = x[:, :3].clone()
        yaw0 = 2.0 * torch.atan2(x[:, 6], x[:, 3])
        c0, s0 = torch.cos(yaw0), torch.sin(yaw0)
        z, o = torch.zeros_like(c0), torch.ones_like(c0)
        self.R0 = torch.stack([torch.stack([c0, -s0, z], 1), torch.stack([s0, c0, z], 1), torch.stack([z, z, o], 1)], 1)   # W_R_Bo (yaw only)
        self.goal = (self.p0 + torch.einsum('bij,j->bi', self.R0, torch.tensor(goal_offset, dtype=torch.float64, device=dev)))[:, None, :].contiguous()
        BpC = torch.tensor(np.asarray(cfg.sensor.B_p_C, float), device=dev)
        BRC = torch.tensor(np.asarray(cfg.sensor.B_R_C, float), device=dev)
        self.W_p_Co = torch.einsum('bij,j->bi', self.R0, BpC) + self.p0        # camera pose when the frame was taken (scene is static)
        self.W_R_Co = self.R0 @ BRC
        if cfg.mpc.model == 'props':
            raise NotImplementedError("rollout records hold the 10-entry state of the att / att_tau / acc models; 'props' is driven through BatchedNmpc directly")
        self.mpc.set_sdf_flag(1.0)
        self.weights = cfg.mpc.weights.set_const_on
        z64 = lambda: torch.zeros(n_env, dtype=torch.float64, device=dev)
        self.fails, self.cols, self.it_qp, self.it_sqp = z64(), z64(), z64(), z64()
        self.min_sdf = torch.full((n_env,), float('inf'), dtype=torch.float64, device=dev)
        self.steps_done = 0
        self.graph = None
        self._encode()

    def _encode(self):
        self.vae.set_img(self.frames)
        self.mpc.set_latent(self.vae.encode(to_numpy=False), self.p0, self.R0)

    def step(self):
        if self.graph is not None:
            self.graph.replay()
            self.steps_done += 1
            return
        if self.steps_done > 0 and self.encode_every and self.steps_done % self.encode_every == 0:
            self._encode()
        self._step_body()
        self.steps_done += 1

    def _step_body(self):
        mpc, x = self.mpc, self.x
        self.rg.set_x0(x)
        tr = self.rg.gen_ref_list_wps(self.goal)
        mpc.set_ref_arrays(*tr.arrays(self.cfg.mpc.model, self.weights))
        mpc.set_x0(x)
        if self.mode == 'sqp':
            u, its, _ = mpc.solve_sqp(tol=self.sqp_tol, max_iter=self.sqp_max_iter)
            self.it_sqp += its.double()
        else:
            mpc.solve()
            u = mpc.get_u()
            self.it_sqp += 1.0
        self.fails += (mpc.status != 0).double()
        self.it_qp += mpc.qp_iters.double()
        torch.minimum(self.min_sdf, mpc.eval(0).double(), out=self.min_sdf)
        mpc.plant_step(x, u, self.dt)                                          # in place: x is the rollout's state
        Co_p = torch.einsum('bji,bj->bi', self.W_R_Co, x[:, :3] - self.W_p_Co)  # body position in the frame of the depth image
        self.cols += self.cc.check_image_points(self.imgs_norm, Co_p.float()).double()

    def capture(self):
        """RTI mode, static scene: one control step as a CUDA graph (after at least one eager step)."""
        if self.mode != 'rti' or self.encode_every:
            raise ValueError('only the RTI rollout with a static scene (encode_every = 0) is captured')
        if self.steps_done == 0:
            self.step()
        if self.mpc._latent_dirty:
            self.mpc.sdf.set_latents(self.mpc.latent)
            self.mpc._latent_dirty = False
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self._step_body()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = g

    def run(self, steps):
        for _ in range(steps):
            self.step()
        return self.records()

    def records(self):
        n = max(self.steps_done, 1)
        rec = torch.zeros(self.n_env, REC_WIDTH, dtype=torch.float64, device=self.dev)
        rec[:, :10] = self.x
        rec[:, 10], rec[:, 11], rec[:, 12] = self.fails, self.cols, self.min_sdf
        rec[:, 13], rec[:, 14] = self.it_qp / n, self.it_sqp / n
        rec[:, 15] = torch.linalg.norm(self.goal[:, 0] - self.x[:, :3], dim=1)
        return rec


def rollout(cfg, sdf_net, enc_net, n_env, steps, first=0, mode='rti', encode_every=1, sqp_tol=5e-2, sqp_max_iter=10,
            goal_offset=(4.0, 0.0, 0.0), n_scenes=8, device='cuda:0', graph=False):
    """Run `steps` control steps for environments [first, first + n_env).  mode: 'rti' (one iteration per step) or 'sqp'
    (iterate to `sqp_tol`).  encode_every: re-encode the environments' depth frames every k steps (0 = only once).
    Returns a float64 tensor [n_env, REC_WIDTH] on the device."""
    r = Rollout(cfg, sdf_net, enc_net, n_env, first, mode, encode_every, sqp_tol, sqp_max_iter, goal_offset, n_scenes, device)
    if graph:
        r.capture()
        steps -= r.steps_done
    return r.run(steps)
