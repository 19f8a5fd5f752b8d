"""Sharding + result gather of the multi-GPU path, exercised with world_size 2 over gloo on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def test_shard_range_partitions():
    from sdf_nmpc_b200.dist import shard_range
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 3, 8) == (3 * 8192, 4 * 8192)


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from sdf_nmpc_b200 import dist as D, synth
    D.init('gloo')
    lo, hi = D.shard_range(total, rank, world)
    w = synth.nmpc_instances(hi - lo, first=lo)              # the same seeded instances whoever owns them
    x = torch.from_numpy(w['x0'])
    u0 = x[:, :4] * 0.5
    rec = D.pack_records(u0, torch.zeros(hi - lo, dtype=torch.int32), torch.full((hi - lo,), 9, dtype=torch.int32), x)
    allrec = D.gather_records(rec)
    t = D.max_over_ranks(10.0 + rank, x.device)
    D.barrier()
    if rank == 0:
        q.put((allrec.numpy(), t))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('total', [64, 37])
def test_gather_records_world2_gloo(total):
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.dist import RECORD_WIDTH
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    allrec, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = synth.nmpc_instances(total)['x0']
    assert allrec.shape == (total, RECORD_WIDTH)
    assert np.array_equal(allrec[:, 6:], ref) and np.array_equal(allrec[:, :4], ref[:, :4] * 0.5)
    assert np.all(allrec[:, 4] == 0) and np.all(allrec[:, 5] == 9)
    assert t == 11.0
