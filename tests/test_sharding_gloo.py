"""N>1 path on the CPU: two gloo ranks shard a batch of subframes, each decodes its share (with the oracle
standing in for the device), and the reduced result equals the single-process one."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as o
    from srsue_b200.shard import shard_range, reduce_metrics
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cell = o.make_cell(6, 1, 1)
    cfg = o.make_cfg(cell, sf_idx=1, cfi=1, qm=2, tbs=152)
    n = 7
    lo, hi = shard_range(n, rank, world)
    ok_bits = 0
    for i in range(lo, hi):
        tb, iq, _ = o.gen_subframe(cell, cfg, 100 + i, 10.0)
        rc, pl, _, _ = o.ue_dl_decode(cell, cfg, iq)
        ok_bits += 152 * int(rc == 0 and np.array_equal(pl, tb))
    t, u = reduce_metrics(1.0 + rank, ok_bits, dist)
    if rank == 0:
        q.put((t, u, lo, hi))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t, u, lo, hi = q.get(timeout=10)
    assert t == 2.0                    # max over ranks
    assert u == 7 * 152                # sum over ranks: all seven subframes decode
    assert (lo, hi) == (0, 4)


def test_shard_helpers():
    from srsue_b200.shard import shard_range, balance_by_work
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    parts = balance_by_work([13 * 5824, 6 * 1024, 13 * 5824, 1 * 176, 5 * 6144, 5 * 6144], 2)
    assert sorted(sum(parts, [])) == list(range(6))
    loads = [sum([13 * 5824, 6 * 1024, 13 * 5824, 1 * 176, 5 * 6144, 5 * 6144][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 6144 * 5
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
