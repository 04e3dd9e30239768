"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding and the summary gathers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smjp_b200.distributed import SummaryGather, gather_lognorm, gather_summaries, shard_by_work, shard_range


def test_shard_ranges_partition_the_chains():
    for total, world in [(4096, 8), (10, 3), (7, 8), (65536, 4), (1, 2)]:
        r = [shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1


def test_shard_by_work_balances_n_squared():
    rng = np.random.default_rng(0)
    n = rng.integers(100, 900, size=5000)
    b = shard_by_work(n, 8)
    assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) > 0)
    work = np.array([np.sum(n[b[i]: b[i + 1]].astype(float) ** 2) for i in range(8)])
    assert work.max() / work.mean() < 1.02


def _worker(rank, world, port, total, K, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    g = torch.Generator().manual_seed(1)
    full_t = torch.rand((total, K), dtype=torch.float64, generator=g)
    full_j = torch.randint(0, 50, (total, K), dtype=torch.int32, generator=g)
    full_l = torch.rand(total, dtype=torch.float64, generator=g)
    t, j = gather_summaries(full_t[lo:hi].clone(), full_j[lo:hi].clone())
    l = gather_lognorm(full_l[lo:hi].clone())
    ok = bool(torch.equal(t, full_t) and torch.equal(j, full_j) and torch.equal(l, full_l))
    # the overlapped gather bench.py keeps inside its timed loop (synchronous on CPU tensors)
    sg = SummaryGather(total, K, rank, world, "cpu")
    for rep in range(2):
        sg.launch(full_t[lo:hi] + rep, full_j[lo:hi] + rep)
    t2, j2 = sg.result()
    ok = ok and bool(torch.equal(t2, full_t + 1) and torch.equal(j2, full_j + 1)) and sg.launched == 2
    q.put((rank, ok, tuple(t.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [11, 64])
def test_gather_summaries_world_size_2_gloo(total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and all(shape == (total, 3) for _, _, shape in res)
