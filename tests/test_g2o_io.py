"""Graph fixture / wire format: g2o text files through the C ABI (sgb_g2o_*). CPU tier: the reader
and writer are host code; the optimisation of a loaded file is in the gpu tier."""
import os

import numpy as np
import pytest

from oracle.cpu_oracle import ALGO_LM, JAC_ANALYTIC, Oracle
from sparse_gslam_b200 import capi
from sparse_gslam_b200 import graphgen as gg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_pose_line.g2o")
FIELDS = ("pose_id", "pose_est", "pose_fixed", "lm_id", "lm_est", "lm_fixed", "pp_i", "pp_j", "pp_z", "pp_info", "pp_phi",
          "pl_pose", "pl_lm", "pl_z", "pl_info")


def test_golden_file_parses_to_known_arrays():
    g = gg.load_g2o(GOLDEN)
    assert (g.P, g.L, g.n_pp, g.n_pl) == (3, 1, 3, 3)
    np.testing.assert_array_equal(g.pose_id, [0, 1, 2])
    np.testing.assert_array_equal(g.lm_id, [10000000])
    np.testing.assert_array_equal(g.pose_fixed, [1, 0, 0])
    np.testing.assert_array_equal(g.pp_i, [0, 1, 0])
    np.testing.assert_array_equal(g.pp_j, [1, 2, 2])
    np.testing.assert_array_equal(g.pp_z[2], [2, 0, 0])
    np.testing.assert_array_equal(g.pp_info[0], [100, 0, 0, 100, 0, 400])
    np.testing.assert_array_equal(g.pp_phi, [0, 0, 1.0])
    np.testing.assert_array_equal(g.pl_pose, [0, 1, 2])
    np.testing.assert_array_equal(g.pl_info[1], [900, 0, 2500])
    # line order = insertion rank across both edge kinds (g2o internalId)
    np.testing.assert_array_equal(g.pp_seq, [0, 2, 5])
    np.testing.assert_array_equal(g.pl_seq, [1, 3, 4])
    # and the oracle can optimise it: chi2 decreases, the wall ends near rho = 2
    o = Oracle(g)
    o.initialize_optimization()
    c0 = o.chi2()[0]
    n, _ = o.optimize(10, ALGO_LM, JAC_ANALYTIC)
    assert n >= 1 and o.chi2()[0] < 0.5 * c0
    assert abs(o.estimates()[1][0, 0] - 2.0) < 0.05


@pytest.mark.parametrize("maker", [lambda: gg.make_small(seed=3), lambda: gg.make("c4"),
                                   lambda: gg.make_small(seed=5, phi=1.0).pose_only(phi=0.75)])
def test_save_load_round_trip_is_exact(tmp_path, maker):
    g = maker()
    path = str(tmp_path / "g.g2o")
    gg.save_g2o(g, path)
    h = gg.load_g2o(path)
    for k in FIELDS:
        np.testing.assert_array_equal(getattr(h, k), getattr(g, k), err_msg=k)   # %.17g round-trips doubles exactly
    # the relative insertion order of all edges survives (sequence numbers themselves are renumbered 0..)
    def order(x):
        seq = np.concatenate([x.pp_seq, x.pl_seq])
        kind = np.concatenate([np.zeros(x.n_pp, int), np.ones(x.n_pl, int)])
        idx = np.concatenate([np.arange(x.n_pp), np.arange(x.n_pl)])
        o = np.lexsort((idx, kind, seq))
        return kind[o], idx[o]
    ka, ia = order(g)
    kb, ib = order(h)
    np.testing.assert_array_equal(ka, kb)
    np.testing.assert_array_equal(ia, ib)
    # save(load(x)) is a fixed point
    path2 = str(tmp_path / "g2.g2o")
    gg.save_g2o(h, path2)
    assert open(path).read() == open(path2).read()


def test_malformed_files_are_rejected_with_a_message(tmp_path):
    bad = tmp_path / "bad.g2o"
    bad.write_text("VERTEX_SE2 0 0 0 0\nEDGE_SE2 0 7 1 0 0 1 0 0 1 0 1\n")
    with pytest.raises(ValueError, match="bad.g2o:2: EDGE_SE2 references an unknown VERTEX_SE2"):
        gg.load_g2o(str(bad))   # errors found when the references are resolved carry the line of the offending entry
    bad.write_text("# c\nVERTEX_SE2 0 0 0 0\nVERTEX_SE2 1 1 0 0\nEDGE_SE2 0 1 1 0 0 1 0 0 1 0 1\n\nFIX 9\n")
    with pytest.raises(ValueError, match="bad.g2o:6: FIX of an unknown vertex 9"):
        gg.load_g2o(str(bad))
    bad.write_text("VERTEX_SE2 0 0 0 0\nVERTEX_SE2 1 1 0 0\nEDGE_SE2 0 1 1 0 0 1 0 0 1 0 1\nROBUST_KERNEL_DCS 3 1.0\n")
    with pytest.raises(ValueError, match="bad.g2o:4: ROBUST_KERNEL_DCS on an unknown edge"):
        gg.load_g2o(str(bad))
    bad.write_text("VERTEX_SE2 0 0 0\n")
    with pytest.raises(ValueError, match="bad.g2o:1: malformed VERTEX_SE2"):
        gg.load_g2o(str(bad))
    with pytest.raises(ValueError, match="cannot open"):
        gg.load_g2o(str(tmp_path / "missing.g2o"))


@pytest.mark.gpu
def test_loaded_file_optimises_like_the_original_graph(tmp_path):
    from sparse_gslam_b200 import SparseOptimizerB200
    g = gg.make("c4")
    path = str(tmp_path / "c4.g2o")
    gg.save_g2o(g, path)
    h = gg.load_g2o(path)
    res = []
    for x in (g, h):
        opt = SparseOptimizerB200(capi.ALGO_LM, jacobian_mode=capi.JAC_ANALYTIC)
        assert opt.initialize_optimization(x)
        n, _ = opt.optimize(15)
        res.append((n, opt.estimates(), opt.active_chi2()))
    assert res[0][0] == res[1][0]
    np.testing.assert_array_equal(res[0][1][0], res[1][1][0])   # same inputs bit for bit => same result bit for bit
    np.testing.assert_array_equal(res[0][1][1], res[1][1][1])
    assert res[0][2] == res[1][2]
