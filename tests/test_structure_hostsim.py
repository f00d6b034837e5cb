"""CPU tier: the product's host-side symbolic phase (sgb_structure.cpp) and the kernel row bodies (sgb_rows.h, run
serially by tests/hostsim) against the oracle. No GPU, no compute call into libsgb.so."""
import ctypes as C
import os

import numpy as np
import pytest

import hostsim
from oracle.cpu_oracle import ALGO_GN, ALGO_LM, JAC_ANALYTIC, JAC_G2O_NUMERIC, Oracle
from sparse_gslam_b200 import capi
from sparse_gslam_b200 import graphgen as gg

STRUCT_KEYS = ("kind", "index", "offset", "row", "col", "nrows", "ncols", "pose_hidx", "lm_hidx")


def assert_same_structure(so, sh):
    for k in ("n_free", "n_blocks", "dim"):
        assert so[k] == sh[k], k
    for k in STRUCT_KEYS:
        assert np.array_equal(so[k], sh[k]), k


@pytest.fixture(autouse=True)
def _owner_only_layout_unless_requested():
    """The planner's default is ghost landmark rows; the tests of the owner-only layout (halo gathers of t / W / b_l)
    switch them off, the `ghost_landmarks` fixture switches them back on."""
    hostsim.use_ghost_landmarks(False)
    hostsim.use_filtered_structure(False)
    yield
    hostsim.use_ghost_landmarks(False)
    hostsim.use_filtered_structure(False)



@pytest.mark.parametrize("maker", [lambda: gg.make_small(seed=0), lambda: gg.make_small(seed=7, P=120, L=20, E_l=300, n_closures=10),
                                   lambda: gg.make_c4_window(1003), lambda: gg.make_c5(rows=20, cols=20),
                                   lambda: gg.make_c1().pose_only(phi=10.0)])
def test_structure_bit_exact(maker):
    """The ordered block list (row, col, nrows, ncols), scalar offsets and Hessian indices equal the oracle's."""
    g = maker()
    o = Oracle(g)
    assert o.initialize_optimization()
    hs = hostsim.HostSim(g)
    assert hs.status == capi.OK, hs.error
    assert_same_structure(o.structure(), hs.structure())


def test_structure_with_fixed_landmark_and_inactive_vertices():
    g = gg.make_small(seed=2)
    g.lm_fixed = g.lm_fixed.copy()
    g.lm_fixed[3] = 1
    g.pose_fixed = g.pose_fixed.copy()
    g.pose_fixed[5] = 1
    # a pose and a landmark that no edge touches
    g.pose_est = np.vstack([g.pose_est, [[50.0, 50.0, 0.0]]])
    g.pose_gt = np.vstack([g.pose_gt, [[50.0, 50.0, 0.0]]])
    g.pose_id = np.append(g.pose_id, g.pose_id.max() + 1).astype(np.int32)
    g.pose_fixed = np.append(g.pose_fixed, 0).astype(np.uint8)
    g.lm_est = np.vstack([g.lm_est, [[3.0, 0.1]]])
    g.lm_gt = np.vstack([g.lm_gt, [[3.0, 0.1]]])
    g.lm_id = np.append(g.lm_id, g.lm_id.max() + 1).astype(np.int32)
    g.lm_fixed = np.append(g.lm_fixed, 0).astype(np.uint8)
    o = Oracle(g)
    assert o.initialize_optimization()
    hs = hostsim.HostSim(g)
    assert hs.status == capi.OK
    so, sh = o.structure(), hs.structure()
    assert_same_structure(so, sh)
    assert sh["pose_hidx"][-1] == -1 and sh["lm_hidx"][-1] == -1 and sh["lm_hidx"][3] == -1 and sh["pose_hidx"][5] == -1
    for numeric in (True, False):
        hs = hostsim.HostSim(g, jac_numeric=numeric)
        lo = o.linearize(JAC_G2O_NUMERIC if numeric else JAC_ANALYTIC)
        lh = hs.linearize()
        np.testing.assert_allclose(lh["H"], lo["H"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(lh["b"], lo["b"], rtol=1e-12, atol=1e-8)
        np.testing.assert_allclose(lh["chi2"], lo["chi2"], rtol=1e-12)


def test_duplicate_and_reversed_edges_share_a_block():
    g = gg.make_small(seed=1)
    k = 10  # duplicate one odometry edge, once in the same and once in the opposite direction
    z = g.pp_z[k]
    c, s = np.cos(z[2]), np.sin(z[2])
    zinv = np.array([-(c * z[0] + s * z[1]), (s * z[0] - c * z[1]), -z[2]])
    g.pp_i = np.concatenate([g.pp_i, [g.pp_i[k], g.pp_j[k]]]).astype(np.int32)
    g.pp_j = np.concatenate([g.pp_j, [g.pp_j[k], g.pp_i[k]]]).astype(np.int32)
    g.pp_z = np.vstack([g.pp_z, z + [0.01, -0.01, 0.002], zinv])
    g.pp_info = np.vstack([g.pp_info, g.pp_info[k] * 0.5, g.pp_info[k] * 0.25])
    g.pp_phi = np.concatenate([g.pp_phi, [0.0, 0.0]])
    top = max(g.pp_seq.max(), g.pl_seq.max())
    g.pp_seq = np.concatenate([g.pp_seq, [top + 1, top + 2]])
    # duplicate a pose-line edge too
    g.pl_pose = np.append(g.pl_pose, g.pl_pose[4]).astype(np.int32)
    g.pl_lm = np.append(g.pl_lm, g.pl_lm[4]).astype(np.int32)
    g.pl_z = np.vstack([g.pl_z, g.pl_z[4] + [0.01, 0.003]])
    g.pl_info = np.vstack([g.pl_info, g.pl_info[4] * 0.7])
    g.pl_seq = np.append(g.pl_seq, top + 3)
    o = Oracle(g)
    o.initialize_optimization()
    hs = hostsim.HostSim(g, jac_numeric=False)
    assert_same_structure(o.structure(), hs.structure())
    lo, lh = o.linearize(JAC_ANALYTIC), hs.linearize()
    np.testing.assert_allclose(lh["H"], lo["H"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(lh["b"], lo["b"], rtol=1e-12, atol=1e-8)


def test_rows_linearize_and_solve_match_oracle(small_graph):
    g = small_graph
    o = Oracle(g)
    o.initialize_optimization()
    for numeric, jac in ((True, JAC_G2O_NUMERIC), (False, JAC_ANALYTIC)):
        hs = hostsim.HostSim(g, jac_numeric=numeric, tol=1e-12)
        lo, lh = o.linearize(jac), hs.linearize()
        np.testing.assert_allclose(lh["H"], lo["H"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(lh["b"], lo["b"], rtol=1e-12, atol=1e-8)
        for lam in (0.0, 0.5, 300.0):
            ok, xo = o.solve_once(lam, jac)
            flag, xh, iters, rel = hs.solve_once(lam)
            assert ok and flag == 0 and iters > 0
            np.testing.assert_allclose(xh, xo, rtol=1e-8, atol=1e-10)


def test_rows_lm_and_gn_match_oracle():
    g = gg.make_small(seed=5, n_closures=8)
    o = Oracle(g)
    o.initialize_optimization()
    hs = hostsim.HostSim(g, jac_numeric=False, tol=1e-12)
    n1, s1 = o.optimize(6, ALGO_LM, JAC_ANALYTIC)
    n2, s2 = hs.optimize(6, capi.ALGO_LM)
    assert n1 == n2
    for a, b in zip(s1, s2):
        assert a["trials"] == b["trials"]
        np.testing.assert_allclose(b["chi2"], a["chi2"], rtol=1e-9)
        np.testing.assert_allclose(b["lambda_"], a["lambda_"], rtol=1e-6)
    po, lo = o.estimates()
    ph, lh = hs.estimates()
    np.testing.assert_allclose(ph, po, atol=1e-8)
    np.testing.assert_allclose(lh, lo, atol=1e-8)
    gp = g.pose_only(phi=1.0)
    o = Oracle(gp)
    o.initialize_optimization()
    hs = hostsim.HostSim(gp, tol=1e-12)
    assert o.optimize(20, ALGO_GN)[0] == hs.optimize(20, capi.ALGO_GN)[0] == 20
    np.testing.assert_allclose(hs.estimates()[0], o.estimates()[0], atol=1e-9)


def test_unsupported_and_invalid_graphs():
    g = gg.make_small(seed=0)
    bad = g.copy()
    bad.pp_j = bad.pp_j.copy()
    bad.pp_j[0] = 10_000
    assert hostsim.HostSim(bad).status == capi.ERR_INVALID
    bad = g.copy()
    bad.lm_id = (bad.lm_id - gg.LANDMARK_ID0).astype(np.int32)  # landmark ids below pose ids
    assert hostsim.HostSim(bad).status == capi.ERR_UNSUPPORTED
    empty = g.copy()
    for k in ("pp_i", "pp_j", "pp_z", "pp_info", "pp_phi", "pp_seq", "pl_pose", "pl_lm", "pl_z", "pl_info", "pl_seq"):
        setattr(empty, k, getattr(empty, k)[:0])
    assert hostsim.HostSim(empty).status == capi.ERR_NOT_INITIALIZED


def test_library_exports_every_declared_symbol():
    """libsgb.so loads without a GPU and exports everything include/sgb_capi.h declares (no compute calls here)."""
    import re
    from sparse_gslam_b200 import build
    build.build()
    L = capi.load()
    hdr = open(os.path.join(os.path.dirname(capi.HERE), "include", "sgb_capi.h")).read()
    declared = sorted(set(re.findall(r"^(?:sgb_status|const char\*|int32_t|void)\s+(sgb_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in sgb_capi.h but not exported"
    assert sorted(capi.EXPORTS) == declared
    assert b"sm_100a" in L.sgb_version()
    opt = capi.Options()
    L.sgb_default_options(C.byref(opt))
    assert opt.pcg_tolerance == 1e-10 and opt.lm_max_trials == 10 and opt.jacobian_mode == capi.JAC_G2O_NUMERIC
    if L.sgb_device_count() == 0:  # no GPU here: creation must fail loudly, not fall back
        h = C.c_void_p()
        assert L.sgb_create(None, C.byref(h)) == capi.ERR_NO_DEVICE
        from sparse_gslam_b200 import SgbError, SparseOptimizerB200
        with pytest.raises(SgbError):
            SparseOptimizerB200()


# ------------------------------------------------------------------ row-block partition (multi-GPU path, virtual ranks)
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_partition_covers_graph_and_matches_single_rank(world):
    """Every row / edge is owned exactly once; linearise, solve and LM over `world` virtual ranks (peer tables wired
    like the NVLink mappings) equal the single-rank result."""
    g = gg.make_small(seed=3, P=120, L=20, E_l=300, n_closures=10)
    one = hostsim.HostSim(g, jac_numeric=False, tol=1e-12)
    many = hostsim.HostSim(g, jac_numeric=False, tol=1e-12, world=world)
    assert many.status == capi.OK, many.error
    st = one.structure()
    ps = many.partition_stats()
    nP = int((st["kind"] == 0).sum())
    nL = int((st["kind"] == 1).sum())
    assert sum(p["nP"] for p in ps) == nP and sum(p["nL"] for p in ps) == nL
    n_pp = int((g.pose_fixed[g.pp_i] == 0).sum() if False else g.n_pp)
    assert sum(p["n_pp_owned"] for p in ps) == n_pp and sum(p["n_pl_owned"] for p in ps) == g.n_pl
    assert all(p["n_pp"] >= p["n_pp_owned"] and p["n_pl"] >= p["n_pl_owned"] for p in ps)
    l1, lm = one.linearize(), many.linearize()
    np.testing.assert_array_equal(lm["H"], l1["H"])      # same arithmetic, same order => bit-identical
    np.testing.assert_array_equal(lm["b"], l1["b"])
    np.testing.assert_allclose(lm["chi2"], l1["chi2"], rtol=1e-13)
    assert many.check_hlp() == 0.0 and one.check_hlp() == 0.0
    f1, x1, it1, _ = one.solve_once(0.3)
    fm, xm, itm, _ = many.solve_once(0.3)
    # the preconditioner's 4-pose blocks are rank-local (they never span a cut and their alignment follows the rank's
    # first row), so the iteration count may differ a little between partitions; the solution may not
    assert f1 == fm == 0 and abs(it1 - itm) <= max(2, 0.25 * it1)
    np.testing.assert_allclose(xm, x1, rtol=1e-9, atol=1e-12)
    n1, s1 = one.optimize(6, capi.ALGO_LM)
    nm, sm = many.optimize(6, capi.ALGO_LM)
    assert n1 == nm and [s["trials"] for s in s1] == [s["trials"] for s in sm]
    p1, q1 = one.estimates()
    for r in range(world):  # every replica of the estimates agrees
        pm, qm = many.estimates(r)
        np.testing.assert_allclose(pm, p1, atol=1e-9)
        np.testing.assert_allclose(qm, q1, atol=1e-9)


def test_partition_c5_halo_is_small():
    g = gg.make_c5(rows=40, cols=40)
    many = hostsim.HostSim(g, jac_numeric=False, world=4)
    ps = many.partition_stats()
    assert all(p["nP"] == 400 or p["nP"] == 399 for p in ps)
    # boustrophedon rows are contiguous in pose id: only the rows next to a cut are halo
    assert max(p["halo_p"] for p in ps) < 0.25 * 400 * 8
    one = hostsim.HostSim(g, jac_numeric=False)
    np.testing.assert_array_equal(many.linearize()["H"], one.linearize()["H"])
    gp = g.pose_only(phi=1.0)
    a, b = hostsim.HostSim(gp, world=3), hostsim.HostSim(gp)
    assert a.optimize(3, capi.ALGO_GN)[0] == b.optimize(3, capi.ALGO_GN)[0] == 3
    np.testing.assert_allclose(a.estimates(2)[0], b.estimates()[0], atol=1e-9)


def test_more_ranks_than_rows():
    from test_oracle import tiny_graph
    g = tiny_graph()
    many = hostsim.HostSim(g, jac_numeric=False, world=4)
    one = hostsim.HostSim(g, jac_numeric=False)
    assert many.status == capi.OK
    np.testing.assert_array_equal(many.linearize()["H"], one.linearize()["H"])
    many.optimize(3, capi.ALGO_LM)
    one.optimize(3, capi.ALGO_LM)
    np.testing.assert_allclose(many.estimates(3)[0], one.estimates()[0], atol=1e-12)


@pytest.mark.parametrize("lam", [1e-3, 10.0])
def test_preconditioner_blocks_are_the_inverse_schur_diagonal_blocks(lam):
    """setup_chunk against numpy: S = Hpp + lam I - Hpl (Hll + lam I)^-1 Hpl^T from the oracle's dense Hessian, cut into
    4-pose (12x12) diagonal blocks, inverted; the kernels store the rows in single precision."""
    from oracle.cpu_oracle import JAC_ANALYTIC, Oracle
    g = gg.make_small(seed=5, P=63, L=12, E_l=170, n_closures=6)  # 62 free poses: the last chunk is partial
    o = Oracle(g)
    assert o.initialize_optimization()
    st = o.structure()
    lin = o.linearize(JAC_ANALYTIC)
    H = o.dense_hessian(lin, st)
    nP = int((st["kind"] == 0).sum())
    n3 = 3 * nP
    Hd = H + lam * np.eye(H.shape[0])
    S = Hd[:n3, :n3] - Hd[:n3, n3:] @ np.linalg.inv(Hd[n3:, n3:]) @ Hd[n3:, :n3]
    hs = hostsim.HostSim(g, jac_numeric=False)
    ok, C = hs.preconditioner(lam, nP)
    assert ok and nP % 4 != 0
    for c0 in range(0, nP, 4):
        c1 = min(nP, c0 + 4)
        Dinv = np.linalg.inv(S[3 * c0:3 * c1, 3 * c0:3 * c1])
        for p in range(c0, c1):
            got = C[p].astype(np.float64)[:, :3 * (c1 - c0)]
            ref = Dinv[3 * (p - c0):3 * (p - c0) + 3, :]
            assert np.abs(got - ref).max() <= 2e-6 * np.abs(Dinv).max(), (c0, p)
            assert np.all(C[p][:, 3 * (c1 - c0):] == 0)  # no coupling to the poses missing from a partial chunk


def _random_graph(rng, p_range=(2, 40)):
    """Small random graph with the collisions the structure builder has to get right: duplicate and reversed edges,
    several fixed poses, fixed landmarks, unobserved (inactive) vertices, shuffled ids, shuffled insertion ranks."""
    P, L = int(rng.integers(*p_range)), int(rng.integers(0, 9))
    gt = np.cumsum(rng.normal(size=(P, 3)) * [0.5, 0.2, 0.2], axis=0)
    lines = np.stack([rng.uniform(0.5, 6.0, L), rng.uniform(-np.pi, np.pi, L)], 1) if L else np.zeros((0, 2))
    n_pp = int(rng.integers(1, 3 * P))
    pp_i = rng.integers(0, P, n_pp)
    pp_j = (pp_i + rng.integers(1, P, n_pp)) % P
    if n_pp > 3:  # duplicates and a reversed duplicate
        pp_i[-1], pp_j[-1] = pp_i[0], pp_j[0]
        pp_i[-2], pp_j[-2] = pp_j[1], pp_i[1]
    n_pl = int(rng.integers(0, 4 * P)) if L else 0
    pl_p, pl_l = rng.integers(0, P, n_pl), rng.integers(0, max(L, 1), n_pl)
    if n_pl > 2:
        pl_p[-1], pl_l[-1] = pl_p[0], pl_l[0]
    z_pp = gg.se2_between(gt[pp_i], gt[pp_j]) + rng.normal(size=(n_pp, 3)) * 0.02
    z_pl = (gg.line_in_pose_frame(gt[pl_p], lines[pl_l]) + rng.normal(size=(n_pl, 2)) * 0.02) if n_pl else np.zeros((0, 2))
    def spd(n, d):
        out = []
        for _ in range(n):
            a = rng.normal(size=(d, d))
            m = a @ a.T + d * np.eye(d)
            out.append([m[0, 0], m[0, 1], m[0, 2], m[1, 1], m[1, 2], m[2, 2]] if d == 3 else [m[0, 0], m[0, 1], m[1, 1]])
        return np.array(out, np.float64).reshape(n, 6 if d == 3 else 3)
    pose_fixed = (rng.random(P) < 0.15).astype(np.uint8)
    pose_fixed[0] = 1
    lm_fixed = (rng.random(L) < 0.15).astype(np.uint8)
    pid = rng.permutation(P).astype(np.int32) * 3  # ids define the Hessian order, not the array positions
    lid = (gg.LANDMARK_ID0 + rng.permutation(L) * 2).astype(np.int32)
    seq = rng.permutation(n_pp + n_pl).astype(np.int64)
    return gg.Graph(name="fuzz", pose_id=pid, pose_est=gt + rng.normal(size=(P, 3)) * 0.05, pose_fixed=pose_fixed, pose_gt=gt,
                    lm_id=lid, lm_est=lines + rng.normal(size=(L, 2)) * 0.02, lm_fixed=lm_fixed, lm_gt=lines,
                    pp_i=pp_i.astype(np.int32), pp_j=pp_j.astype(np.int32), pp_z=z_pp, pp_info=spd(n_pp, 3),
                    pp_phi=np.where(rng.random(n_pp) < 0.3, 1.0, 0.0), pp_seq=seq[:n_pp],
                    pl_pose=pl_p.astype(np.int32), pl_lm=pl_l.astype(np.int32), pl_z=z_pl, pl_info=spd(n_pl, 2), pl_seq=seq[n_pp:])


@pytest.mark.parametrize("seed", range(40))
def test_random_graphs_structure_and_system_match_oracle(seed):
    """Fuzz of the host symbolic phase + the row bodies against the oracle: index mapping, block list (bit-exact), H, b,
    chi2 and one damped solve on random graphs full of collisions (duplicates, reversed edges, fixed and inactive vertices,
    permuted ids and insertion ranks, DCS on a random subset)."""
    from oracle.cpu_oracle import JAC_ANALYTIC, Oracle
    g = _random_graph(np.random.default_rng(1000 + seed))
    o = Oracle(g)
    hs = hostsim.HostSim(g, jac_numeric=False, tol=1e-12)
    ok_o = o.initialize_optimization()
    if not ok_o:
        assert hs.status != capi.OK
        return
    assert hs.status == capi.OK, hs.error
    so, sh = o.structure(), hs.structure()
    for k in ("n_free", "n_blocks", "dim"):
        assert so[k] == sh[k], k
    for k in ("kind", "index", "offset", "row", "col", "nrows", "ncols", "pose_hidx", "lm_hidx"):
        assert np.array_equal(so[k], sh[k]), k
    lo, lh = o.linearize(JAC_ANALYTIC), hs.linearize()
    scale = max(1.0, np.abs(lo["H"]).max())
    assert np.abs(lh["H"] - lo["H"]).max() <= 1e-12 * scale
    assert np.abs(lh["b"] - lo["b"]).max() <= 1e-11 * max(1.0, np.abs(lo["b"]).max())
    np.testing.assert_allclose(lh["chi2"], lo["chi2"], rtol=1e-12)
    lam = 10.0  # damped enough for every random graph to be positive definite
    ok, xo = o.solve_once(lam, JAC_ANALYTIC)
    f, xh, it, rel = hs.solve_once(lam)
    assert ok and f == 0, (ok, f)
    assert np.abs(xh - xo).max() <= 1e-8 * max(1e-3, np.abs(xo).max())


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("world", [2, 3])
def test_random_graphs_partitioned_match_single_rank(seed, world):
    """The same fuzz through the row-block partition planner: `world` virtual ranks against one."""
    g = _random_graph(np.random.default_rng(2000 + seed))
    one = hostsim.HostSim(g, jac_numeric=False, tol=1e-12)
    many = hostsim.HostSim(g, jac_numeric=False, tol=1e-12, world=world)
    assert one.status == many.status
    if one.status != capi.OK:
        return
    l1, lm = one.linearize(), many.linearize()
    np.testing.assert_array_equal(lm["H"], l1["H"])
    np.testing.assert_array_equal(lm["b"], l1["b"])
    assert many.check_hlp() == 0.0
    f1, x1, _, _ = one.solve_once(10.0)
    fm, xm, _, _ = many.solve_once(10.0)
    assert f1 == fm == 0
    assert np.abs(xm - x1).max() <= 1e-8 * max(1e-3, np.abs(x1).max())
    n1, _ = one.optimize(3, capi.ALGO_LM)
    nm, _ = many.optimize(3, capi.ALGO_LM)
    assert n1 == nm
    p1, q1 = one.estimates()
    for r in range(world):
        pm, qm = many.estimates(r)
        np.testing.assert_allclose(pm, p1, atol=1e-8)
        np.testing.assert_allclose(qm, q1, atol=1e-8)


@pytest.fixture(params=[False, True], ids=["whole-structure", "rank-filtered"])
def ghost_landmarks(request):
    """Ghost landmark rows on; `rank-filtered`: every virtual rank also builds its own structure from the edges it needs
    only (build_structure(..., world, rank)), the way libsgb runs the symbolic phase on > 1 GPUs."""
    hostsim.use_ghost_landmarks(True)
    hostsim.use_filtered_structure(request.param)
    yield
    hostsim.use_filtered_structure(False)
    hostsim.use_ghost_landmarks(False)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_ghost_landmarks_match_single_rank(world, ghost_landmarks):
    """Planner mode with ghost rows (sgb_partition.h): every rank keeps a full copy of the landmark rows its poses
    observe. No landmark quantity is read from another rank any more (halo_t == 0), and the system, the solve and the LM
    trajectory are the single-rank ones -- also when every rank only ever saw its own share of the edges."""
    g = gg.make_small(seed=3, P=120, L=20, E_l=300, n_closures=10)
    hostsim.use_ghost_landmarks(False)
    one = hostsim.HostSim(g, jac_numeric=False, tol=1e-12)
    plain = hostsim.HostSim(g, jac_numeric=False, tol=1e-12, world=world)
    hostsim.use_ghost_landmarks(True)
    many = hostsim.HostSim(g, jac_numeric=False, tol=1e-12, world=world)
    assert many.status == capi.OK, many.error
    ps, pp = many.partition_stats(), plain.partition_stats()
    assert all(p["halo_t"] == 0 for p in ps) and any(p["halo_t"] > 0 for p in pp)
    # pushed pose halos: no column of a rank's matrices names a row of another rank any more, the halo copies do
    assert all(p["remote_cols"] == 0 and p["nH"] > 0 for p in ps) and any(p["remote_cols"] > 0 for p in pp)
    assert sum(p["nL"] for p in ps) > sum(p["nL"] for p in pp)          # the ghost rows
    assert sum(p["n_pl_owned"] for p in ps) == g.n_pl and sum(p["n_pp_owned"] for p in ps) == g.n_pp
    l1, lm = one.linearize(), many.linearize()
    np.testing.assert_array_equal(lm["H"], l1["H"])
    np.testing.assert_array_equal(lm["b"], l1["b"])
    np.testing.assert_allclose(lm["chi2"], l1["chi2"], rtol=1e-13)
    assert many.check_hlp() == 0.0
    f1, x1, it1, _ = one.solve_once(0.3)
    fm, xm, itm, _ = many.solve_once(0.3)
    assert f1 == fm == 0 and abs(it1 - itm) <= max(2, 0.25 * it1)
    np.testing.assert_allclose(xm, x1, rtol=1e-9, atol=1e-12)
    n1, s1 = one.optimize(6, capi.ALGO_LM)
    nm, sm = many.optimize(6, capi.ALGO_LM)
    assert n1 == nm and [s["trials"] for s in s1] == [s["trials"] for s in sm]
    np.testing.assert_allclose([s["chi2"] for s in sm], [s["chi2"] for s in s1], rtol=1e-9)
    p1, q1 = one.estimates()
    for r in range(world):
        pm, qm = many.estimates(r)
        np.testing.assert_allclose(pm, p1, atol=1e-9)
        np.testing.assert_allclose(qm, q1, atol=1e-9)


@pytest.mark.parametrize("seed", range(10))
def test_ghost_landmarks_fuzz(seed, ghost_landmarks):
    g = _random_graph(np.random.default_rng(3000 + seed))
    hostsim.use_ghost_landmarks(False)
    one = hostsim.HostSim(g, jac_numeric=False, tol=1e-12)
    hostsim.use_ghost_landmarks(True)
    many = hostsim.HostSim(g, jac_numeric=False, tol=1e-12, world=3)
    assert one.status == many.status
    if one.status != capi.OK:
        return
    l1, lm = one.linearize(), many.linearize()
    np.testing.assert_array_equal(lm["H"], l1["H"])
    np.testing.assert_array_equal(lm["b"], l1["b"])
    np.testing.assert_allclose(lm["chi2"], l1["chi2"], rtol=1e-12)
    assert many.check_hlp() == 0.0
    f1, x1, _, _ = one.solve_once(10.0)
    fm, xm, _, _ = many.solve_once(10.0)
    assert f1 == fm == 0
    assert np.abs(xm - x1).max() <= 1e-8 * max(1e-3, np.abs(x1).max())
    assert one.optimize(3, capi.ALGO_LM)[0] == many.optimize(3, capi.ALGO_LM)[0]
    np.testing.assert_allclose(many.estimates(2)[0], one.estimates()[0], atol=1e-8)
    np.testing.assert_allclose(many.estimates(1)[1], one.estimates()[1], atol=1e-8)


def test_ctypes_mirrors_have_the_layout_of_the_c_structs(tmp_path):
    """The Python host side talks to libsgb.so through ctypes mirrors of the structs of include/sgb_capi.h: a C program
    compiled against the header prints sizeof and the offset of the last member of each, which must be what ctypes lays
    out (a member added on one side only would shift everything behind it silently)."""
    import subprocess
    pairs = [("sgb_options", capi.Options), ("sgb_graph_soa", capi.GraphSoA), ("sgb_graph_delta", capi.GraphDelta),
             ("sgb_iter_stat", capi.IterStat), ("sgb_structure_info", capi.StructureInfo), ("sgb_timings", capi.Timings),
             ("sgb_partition_info", capi.PartitionInfo), ("sgb_block_matrix", capi.BlockMatrix),
             ("sgb_device_values", capi.DeviceValues), ("sgb_pg_info", capi.PgInfo)]
    lines = []
    for cname, cls in pairs:   # the mirrors use the header's member names: offsetof of a misnamed member does not compile
        last = cls._fields_[-1][0]
        lines.append('printf("%s %%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));' % (cname, cname, cname, last))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sgb_capi.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(os.path.dirname(capi.HERE), "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    for (cname, cls), ln in zip(pairs, out):
        name, size, off = ln.split()
        last = cls._fields_[-1][0]
        assert name == cname and int(size) == C.sizeof(cls) and int(off) == getattr(cls, last).offset, (ln, C.sizeof(cls))
