"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + proof gathering used by bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hekaton_system_b200.shard import derive_stage1_requests, gather_records, shard_range, stage_barrier


def test_shard_range_matches_reference_chunking():
    assert list(shard_range(64, 8, 0)) == list(range(0, 8))
    assert list(shard_range(64, 8, 7)) == list(range(56, 64))
    assert [i for r in range(4) for i in shard_range(16, 4, r)] == list(range(16))
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)          # node.rs:472: n_sub % workers == 0


def _worker(rank, world, port, n_sub, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard_range(n_sub, world, rank)
    # a fake 'proof record' per local subcircuit: [subcircuit index, index^2 + 7, rank]
    rec = torch.tensor([[i, i * i + 7, rank] for i in idx], dtype=torch.int64)
    allrec = gather_records(rec, world)
    # the stage barrier of a run: every rank's commitment records -> one challenge, the same on all ranks
    chal, host = stage_barrier(rec * 3 + 1, world)
    q.put((rank, allrec.tolist(), chal, host.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_sub, world = 8, 2
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_sub, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(world)]
    got = {r[0]: r[1] for r in res}
    [p.join(timeout=60) for p in ps]
    expect = [[i, i * i + 7, i // (n_sub // world)] for i in range(n_sub)]
    assert got[0] == expect and got[1] == expect
    # barrier: both ranks derived the same challenge from the SAME gathered records (rank-major order), and it is
    # the stub applied to all of them -- no rank can have a stage-1 request before every stage-0 record arrived
    import numpy as np
    chals = {r[0]: r[2] for r in res}
    hosts = {r[0]: r[3] for r in res}
    assert chals[0] == chals[1] and hosts[0] == hosts[1]
    per = n_sub // world
    flat = [[[3 * v + 1 for v in expect[i]] for i in range(r * per, (r + 1) * per)] for r in range(world)]
    assert hosts[0] == [sum(rows, []) for rows in flat]
    assert chals[0] == derive_stage1_requests(np.array(hosts[0], dtype=np.int64))
