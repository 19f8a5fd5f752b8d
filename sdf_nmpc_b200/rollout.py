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


def rollout(cfg, sdf_net, enc_net, n_env, steps, first=0, mode='rti', encode_every=1, sqp_tol=5e-2, sqp_max_iter=10,
            goal_offset=(4.0, 0.0, 0.0), n_scenes=8, device='cuda:0', timers=None):
    """Run `steps` control steps for environments [first, first + n_env).  mode: 'rti' (one iteration per step) or 'sqp'
    (iterate to `sqp_tol`).  encode_every: re-encode the environments' depth frames every k steps (0 = only once).
    Returns a float64 tensor [n_env, REC_WIDTH] on the device."""
    dev = torch.device(device)
    spherical = bool(cfg.sensor.is_spherical)
    w = synth.nmpc_instances(n_env, N=cfg.mpc.N, dt=float(cfg.shooting_dt()[0]), first=first, model=cfg.mpc.model)
    # a small pool of synthetic scenes, one per environment (seeded by the global environment index)
    pool = np.stack([synth.depth_frame(seed=200 + s, spherical=spherical) for s in range(n_scenes)]).astype(np.float32)
    scene = torch.from_numpy((np.arange(first, first + n_env) % n_scenes).astype(np.int64)).to(dev)
    pool_d = torch.from_numpy(pool).to(dev)
    frames = pool_d[scene]                                             # [n_env, H, W] raw sensor images
    vae = VaeWrapper(cfg, enc_net, max_batch=min(n_env, 256), device=device)      # encode() walks larger batches in chunks of 256
    mpc = BatchedNmpc(cfg, sdf_net, n_env, device)
    rg = BatchedRefGen(cfg)
    cc = ColChecker(cfg.sensor.dmax, cfg.sensor.hfov, cfg.sensor.vfov, cfg.robot.size.xy, is_depth=bool(cfg.sensor.is_depth),
                    is_spherical=spherical, outside='free', device=device)
    imgs_norm = torch.clamp(frames / float(cfg.sensor.dmax), 0, 1)    # what ColChecker expects (normalised by dmax)

    x = torch.from_numpy(w['x0']).to(dev)
    x[:, 7:] = 0.0
    p0 = x[:, :3].clone()
    yaw0 = 2.0 * torch.atan2(x[:, 6], x[:, 3])
    c0, s0 = torch.cos(yaw0), torch.sin(yaw0)
    z, o = torch.zeros_like(c0), torch.ones_like(c0)
    R0 = torch.stack([torch.stack([c0, -s0, z], 1), torch.stack([s0, c0, z], 1), torch.stack([z, z, o], 1)], 1)   # W_R_Bo (yaw only)
    goal = p0 + torch.einsum('bij,j->bi', R0, torch.tensor(goal_offset, dtype=torch.float64, device=dev))
    BpC = torch.tensor(np.asarray(cfg.sensor.B_p_C, float), device=dev)
    BRC = torch.tensor(np.asarray(cfg.sensor.B_R_C, float), device=dev)
    W_p_Co = torch.einsum('bij,j->bi', R0, BpC) + p0                   # camera pose when the frame was taken (scene is static)
    W_R_Co = R0 @ BRC
    mpc.set_sdf_flag(1.0)
    weights = cfg.mpc.weights.set_const_on

    fails = torch.zeros(n_env, dtype=torch.float64, device=dev)
    cols = torch.zeros_like(fails)
    min_sdf = torch.full_like(fails, float('inf'))
    it_qp = torch.zeros_like(fails)
    it_sqp = torch.zeros_like(fails)
    for step in range(steps):
        if step == 0 or (encode_every and step % encode_every == 0):
            vae.set_img(frames)
            lat = vae.encode(to_numpy=False)
            mpc.set_latent(lat, p0, R0)
        rg.set_x0(x)
        tr = rg.gen_ref_list_wps(goal[:, None, :])
        mpc.set_ref_arrays(*tr.arrays(cfg.mpc.model, weights))
        mpc.set_x0(x)
        if mode == 'sqp':
            u, its, _ = mpc.solve_sqp(tol=sqp_tol, max_iter=sqp_max_iter)
            it_sqp += its.double()
        else:
            mpc.solve()
            u = mpc.get_u()
            it_sqp += 1.0
        fails += (mpc.status != 0).double()
        it_qp += mpc.qp_iters.double()
        min_sdf = torch.minimum(min_sdf, mpc.eval(0).double())
        x = mpc.plant_step(x.clone(), u, float(cfg.shooting_dt()[0]))
        Co_p = torch.einsum('bji,bj->bi', W_R_Co, x[:, :3] - W_p_Co)   # body position in the frame of the depth image
        cols += cc.check_image_points(imgs_norm, Co_p.float()).double()
    rec = torch.zeros(n_env, REC_WIDTH, dtype=torch.float64, device=dev)
    rec[:, :10] = x
    rec[:, 10], rec[:, 11], rec[:, 12] = fails, cols, min_sdf
    rec[:, 13], rec[:, 14] = it_qp / steps, it_sqp / steps
    rec[:, 15] = torch.linalg.norm(goal - x[:, :3], dim=1)
    return rec
