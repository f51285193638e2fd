"""CPU tests (gloo, world_size 2 and 3) of the pair sharding + result gather used at N > 1."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_partition(pkg):
    import importlib

    sh = importlib.import_module("nct_b200.sharding")
    for n, world in [(32, 8), (9, 2), (5, 4), (1, 1)]:
        seen = []
        for r in range(world):
            seen += sh.shard_indices(n, r, world)
        assert sorted(seen) == list(range(n))
    assert sh.shard_indices(32, 3, 8) == [3, 11, 19, 27]  # BASELINE config 3: 32 pairs over 8 ranks, 4 each
    with pytest.raises(ValueError):
        sh.shard_indices(4, 2, 2)


def _worker(rank, world, port, n_pairs, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.load_package()
    import importlib

    sh = importlib.import_module("nct_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sh.shard_indices(n_pairs, rank, world)
    # stand-in for the per-pair transfer: a result image that encodes the pair index; pairs.txt images differ in size
    local = [torch.full((4 + i % 3, 5 + i % 2, 3), i, dtype=torch.uint8) for i in mine]
    out = sh.gather_results(local, n_pairs, rank, world, dist)
    if rank == 0:
        assert all(tuple(t.shape) == (4 + i % 3, 5 + i % 2, 3) for i, t in enumerate(out))
        q.put([int(t[0, 0, 0]) for t in out])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pairs", [(2, 9), (2, 4), (3, 7), (3, 2), (2, 1)])  # the last two: ranks without any pair
def test_gather_over_gloo(world, n_pairs):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + n_pairs) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == list(range(n_pairs))
