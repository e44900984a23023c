"""world_size-2 `gloo` test (CPU) of the N>1 host logic: contiguous image sharding with no data-path collective, summaries
gathered over torch.distributed.  The per-image work here is the CPU oracle (test infrastructure) standing in for the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from oracle_lib import load_oracle
    from visioncortex_b200 import synth
    from visioncortex_b200.runtime import make_config
    from visioncortex_b200.sharding import run_sharded, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = load_oracle()
    cfg = make_config(hierarchical=0, is_same_color_a=2)
    images = [synth.g2_scene(96, 64, seed=100 + k) for k in range(7)]

    def process(img):
        r = oracle.runner_run(img, cfg, pools=False)
        return (int(r["num_slots"]), int(len(r["clusters_output"])), int(r["cluster_indices"].astype(np.uint64).sum()))

    got = run_sharded(images, process, world, rank, dist)
    mine = list(shard_range(len(images), world, rank))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, got))


def test_sharding_two_ranks_gloo():
    sys.path.insert(0, ROOT)
    from visioncortex_b200.sharding import owner_of, shard_range

    # pure arithmetic: blocks are contiguous, disjoint, cover everything, and match the owner function
    for n in (1, 7, 8, 512):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                blk = list(shard_range(n, world, r))
                assert all(owner_of(k, n, world) == r for k in blk)
                seen += blk
            assert seen == list(range(n))

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    (r0, mine0, got0), (r1, mine1, got1) = res
    assert sorted(mine0 + mine1) == list(range(7)) and not set(mine0) & set(mine1)
    assert got0 == got1 and sorted(got0) == list(range(7))
