"""Multi-GPU plumbing: instances are independent NMPC problems, so the path shards trivially - contiguous
blocks of instances per rank, no data-path collective - and only the fixed-width per-instance result records
of a rollout are gathered (one all-gather).  ``torch.distributed`` with NCCL on the GPU box,
``gloo`` in the CPU tests.
"""
import os

import torch
import torch.distributed as dist

RECORD_WIDTH = 16     # u0(4) status(1) qp_iters(1) x(10)


def env_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kw['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of ``total`` instances owned by ``rank`` (ragged tails go to the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(u0, status, qp_iters, x):
    """[B,16] float64 result record per instance."""
    return torch.cat([u0.double(), status.double()[:, None], qp_iters.double()[:, None], x.double()], 1).contiguous()


def gather_records(local, total=None):
    """All-gather of per-instance records.  Equal shards use ``all_gather_into_tensor`` (one NCCL launch);
    ragged shards fall back to ``all_gather`` on padded blocks.  Returns [total, width] on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0], local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    m = max(sizes)
    pad = torch.zeros((m, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    blocks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(blocks, pad)
    return torch.cat([b[:s] for b, s in zip(blocks, sizes)], 0)


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
