"""N > 1 host path on CPU: world_size-2 gloo processes exercise the parameter sharding,
the summary gather (equal and ragged shards) and the verdict AND-reduce."""

from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcmc_ref_b200.sharding import all_passed, gather_draw_blocks, gather_summaries, shard_bounds, verdict_from_summary


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000, 100_000):
        for world in (1, 2, 3, 8):
            bounds = [shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_params: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_params, world, rank)
        # row p of the summary is [p, p + 0.5, ...]: the gathered matrix must be in parameter order
        local = torch.arange(lo, hi, dtype=torch.float64)[:, None] + torch.arange(16, dtype=torch.float64)[None, :] / 32
        full = gather_summaries(local, n_params)
        want = torch.arange(n_params, dtype=torch.float64)[:, None] + torch.arange(16, dtype=torch.float64)[None, :] / 32
        ok_order = bool(torch.equal(full, want))
        full2 = gather_summaries(local)  # sizes discovered by a collective
        ok_auto = bool(torch.equal(full2, want))
        v_all = all_passed(True)
        v_one = all_passed(rank != 1)
        # ONE collective per step: the verdict is the AND of the gathered flags column (column 15, value 3 =
        # mean and std passed); rank 1's last parameter fails in the second variant
        flags = torch.full((hi - lo, 20), 3.0, dtype=torch.float64)
        v_col_all = bool(verdict_from_summary(gather_summaries(flags, n_params).clone()))
        if rank == 1:
            flags[-1, 15] = 1.0
        v_col_one = bool(verdict_from_summary(gather_summaries(flags, n_params)))
        # draw-block gather of the sharded log-density (value = draw index)
        dlo, dhi = shard_bounds(2 * n_params + 1, world, rank)
        lp = gather_draw_blocks(torch.arange(dlo, dhi, dtype=torch.float64), 2 * n_params + 1)
        ok_lp = bool(torch.equal(lp, torch.arange(2 * n_params + 1, dtype=torch.float64)))
        q.put((rank, ok_order, ok_auto, v_all, v_one, v_col_all, v_col_one, ok_lp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_params", [10, 11])
def test_gather_and_verdict_world2(n_params):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_params, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_order, ok_auto, v_all, v_one, v_col_all, v_col_one, ok_lp in results:
        assert ok_order and ok_auto and ok_lp, rank
        assert v_all is True and v_one is False
        assert v_col_all is True and v_col_one is False


def test_cpulist_and_numa_binding_without_a_gpu():
    """sysfs cpulist syntax; without a CUDA device (this container) the binding reports why it did nothing and
    leaves the affinity alone."""
    import os

    from mcmc_ref_b200 import sharding

    assert sharding._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding._cpulist("5") == {5} and sharding._cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = sharding.bind_host_to_gpu(0)
    assert set(info) >= {"numa_nodes", "gpu_numa_node", "cpus_bound"}
    import torch

    if not torch.cuda.is_available():
        assert info["cpus_bound"] is None and os.sched_getaffinity(0) == before
