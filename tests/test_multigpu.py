"""GPU tier, needs >= 2 GPUs (skipped otherwise): the row-block partitioned path, one process per GPU, against the
single-GPU result and the oracle. Halo gathers and PCG reductions run over NVLink peer memory."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _collect(procs, q, world, timeout=600):
    """One result per rank; a worker that dies (non-zero exit code) fails the test at once instead of stalling q.get."""
    import queue
    import time
    res, t0 = {}, time.time()
    while len(res) < world:
        try:
            r, out = q.get(timeout=2)
            res[r] = out
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.kill()
                raise AssertionError(f"multi-GPU workers failed: exit codes {[p.exitcode for p in procs]}")
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


LAYOUTS = {  # planner mode of the partitioned path -> environment of the workers
    "owner-only": {"SGB_GHOST_LANDMARKS": "0"},                              # remote t / W / b_l gathered over NVLink
    "ghost-rows": {"SGB_GHOST_LANDMARKS": "1", "SGB_PARTITION_FULL": "1"},   # ghost landmark rows, whole structure per rank
    "ghost-rows-filtered": {"SGB_GHOST_LANDMARKS": "1", "SGB_PARTITION_FULL": "0"},  # the default: per-rank symbolic phase
}


def _worker(rank, world, port, q, layout):
    sys.path.insert(0, ROOT)
    os.environ.update(LAYOUTS[layout])
    export_h = layout != "ghost-rows-filtered"   # a rank-filtered handle cannot export blocks in the global order
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparse_gslam_b200 import SparseOptimizerB200, capi
    from sparse_gslam_b200 import dist as sdist
    from sparse_gslam_b200 import graphgen as gg
    out = {}
    g = gg.make_c5(rows=40, cols=40)
    opt = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC, device=rank)
    assert opt.initialize_partitioned(g, world, rank, sdist.exchange_blobs)
    out["info"] = opt.partition_info()
    lin = opt.linearize(hessian=export_h)
    out["H"], out["b"], out["chi2"] = lin["H"], lin["b"], lin["chi2"]
    sdist.barrier()
    ok, x, it, rel = opt.solve_once(50.0)
    out["x"], out["solve"] = x, (ok, it, rel)
    sdist.barrier()
    n, st = opt.optimize(8)
    out["lm"] = (n, [s["trials"] for s in st], [s["chi2"] for s in st])
    out["est"] = opt.estimates()
    out["chi2_after"] = opt.active_chi2()
    sdist.barrier()
    gp = g.pose_only(phi=1.0)
    gn = SparseOptimizerB200(capi.ALGO_GN, device=rank)
    assert gn.initialize_partitioned(gp, world, rank, sdist.exchange_blobs)
    out["gn"] = gn.optimize(3)[0]
    out["gn_est"] = gn.estimates()[0]
    sdist.barrier()
    opt.close()
    gn.close()
    dist.destroy_process_group()
    q.put((rank, out))


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_matches_single_gpu(world, layout):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from sparse_gslam_b200 import SparseOptimizerB200, capi
    from sparse_gslam_b200 import graphgen as gg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, layout)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world)
    g = gg.make_c5(rows=40, cols=40)
    one = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC)
    one.initialize_optimization(g)
    lin = one.linearize()
    b = sum(res[r]["b"] for r in range(world))
    if res[0]["H"] is not None:
        H = sum(res[r]["H"] for r in range(world))
        np.testing.assert_array_equal(H, lin["H"])     # each block is written by exactly one rank, same arithmetic
    np.testing.assert_array_equal(b, lin["b"])
    for r in range(world):
        np.testing.assert_allclose(res[r]["chi2"], lin["chi2"], rtol=1e-13)
    ok, x1, it1, _ = one.solve_once(50.0)
    x = sum(res[r]["x"] for r in range(world))
    assert ok and all(res[r]["solve"][0] for r in range(world))
    np.testing.assert_allclose(x, x1, rtol=1e-8, atol=1e-12)
    n1, s1 = one.optimize(8)
    p1, l1 = one.estimates()
    for r in range(world):
        n, trials, chis = res[r]["lm"]
        assert n == n1 and trials == [s["trials"] for s in s1]
        np.testing.assert_allclose(chis, [s["chi2"] for s in s1], rtol=1e-9)
        np.testing.assert_allclose(res[r]["est"][0], p1, atol=1e-8)
        np.testing.assert_allclose(res[r]["est"][1], l1, atol=1e-8)
        np.testing.assert_allclose(res[r]["chi2_after"], one.active_chi2(), rtol=1e-9)
    gn = SparseOptimizerB200(capi.ALGO_GN)
    gn.initialize_optimization(g.pose_only(phi=1.0))
    assert gn.optimize(3)[0] == 3 == res[0]["gn"]
    np.testing.assert_allclose(res[1]["gn_est"], gn.estimates()[0], atol=1e-8)
