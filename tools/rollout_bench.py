"""configs[3] / configs[4] drivers: Monte-Carlo closed-loop rollouts with result gather.
    python tools/rollout_bench.py rti 4096 50          # config 4 flavour on one GPU (per-GPU shard of the 65 536)
    python tools/rollout_bench.py sqp 1024 20          # config 5: per-step batched encoding + full SQP
  multi-GPU: python -m torch.distributed.run --nproc-per-node N ... tools/rollout_bench.py rti 65536 50  (total environments)"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdf_nmpc_b200 import dist, synth
from sdf_nmpc_b200.config import Config
from sdf_nmpc_b200.rollout import rollout

mode = sys.argv[1] if len(sys.argv) > 1 else 'rti'
total = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
if world > 1:
    torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dev = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
cfg = Config()
if mode == 'sqp':
    cfg.sensor.is_depth, cfg.sensor.is_spherical = False, True
    cfg.sensor.vfov = cfg.sensor.hfov / cfg.sensor.aspect_ratio
lo, hi = dist.shard_range(total, rank, world)
net, enc = synth.make_sdf_net(seed=0, **synth.ARCH_B, calibrated=True), synth.make_encoder_net(seed=5)
kw = dict(mode=mode, encode_every=1 if mode == 'sqp' else 0, device=dev)
rollout(cfg, net, enc, min(hi - lo, 256), 2, first=lo, **kw)        # warm-up (allocations, first-touch)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
t0 = time.perf_counter()
rec = rollout(cfg, net, enc, hi - lo, steps, first=lo, **kw)
allrec = dist.gather_records(rec) if world > 1 else rec
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    r = allrec.cpu().numpy()
    print(json.dumps(dict(mode=mode, environments=int(r.shape[0]), steps=steps, n_gpus=world, seconds=dt, env_steps_per_s=r.shape[0] * steps / dt,
                          fail_steps=float(r[:, 10].sum()), envs_with_collision=int((r[:, 11] > 0).sum()), min_sdf_seen=float(r[:, 12].min()),
                          mean_qp_iters=float(r[:, 13].mean()), mean_sqp_iters=float(r[:, 14].mean()), mean_goal_dist=float(r[:, 15].mean()))))
