"""world_size-2 gloo test of the multi-process plumbing of the partitioned path (no GPU): every rank plans its own
share with the product's partition planner and the ranks exchange handle blobs the way bench.py does."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hostsim
    from sparse_gslam_b200 import dist as sdist
    from sparse_gslam_b200 import graphgen as gg
    g = gg.make_small(seed=3, P=120, L=20, E_l=300, n_closures=10)
    hostsim.use_ghost_landmarks(True)       # the product's default layout on > 1 GPUs: ghost landmark rows, and every
    hostsim.use_filtered_structure(True)    # process runs the symbolic phase on its own share of the edges only
    hs = hostsim.HostSim(g, jac_numeric=False, world=world)
    mine = hs.partition_stats()[rank]
    blob = bytes([rank]) * 64
    blobs = sdist.exchange_blobs(blob)
    stats = [None] * world
    dist.all_gather_object(stats, mine)
    sdist.barrier()
    dist.destroy_process_group()
    q.put((rank, blobs, stats))


def test_two_rank_partition_and_handle_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, blobs, stats in res:
        assert blobs == [bytes([r]) * 64 for r in range(world)]       # rank order preserved
        assert sum(s["nP"] for s in stats) == 119 and sum(s["nL_owned"] for s in stats) == 20
        assert sum(s["nL"] for s in stats) > 20 and all(s["halo_t"] == 0 for s in stats)   # ghost rows, no remote landmark reads
        assert sum(s["n_pl_owned"] for s in stats) == 300
