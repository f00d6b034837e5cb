// sgb_g2o_adapter.h -- header-only g2o plugin that puts the B200 backend behind the interface the reference
// configures in src/sparse_gslam/src/graphs.cpp:9-23:
//
//     opt.setAlgorithm(new g2o::OptimizationAlgorithmLevenberg(make_unique<BlockSolver<..>>(make_unique<LinearSolverEigen<..>>())));
//  becomes
//     opt.setAlgorithm(new g2o::OptimizationAlgorithmB200(SGB_ALGO_LM));        // setup_lm_opt
//     opt.setAlgorithm(new g2o::OptimizationAlgorithmB200(SGB_ALGO_GN));        // setup_pose_opt
//
// It implements g2o::OptimizationAlgorithm (init / solve / updateStructure / computeMarginals), i.e. it replaces the
// algorithm + BlockSolver + LinearSolver triple at once, so that LM damping and the accept/reject logic stay on the
// GPU (SURVEY.md 8b "recommended level"). init() flattens the optimizer's active graph into the SoA of
// include/sgb_capi.h and uploads it; solve(i) runs one iteration on the device and writes the estimates back to the
// g2o vertices (the caller reads them right after optimize(), drone.cpp:164-165,186, and uses host push/pop).
//
// Compiles against real g2o headers (define SGB_USE_REAL_G2O and include them first) or against
// g2o_compat/g2o_compat.h (this repository's tests). Links libsgb.so only.
#pragma once
#include <algorithm>
#include <cstdio>
#include <unordered_map>
#include <vector>

#include "../../include/sgb_capi.h"
#ifndef SGB_USE_REAL_G2O
#include "g2o_compat/g2o_compat.h"
#endif

namespace g2o {

class OptimizationAlgorithmB200 : public OptimizationAlgorithm
#ifndef SGB_USE_REAL_G2O
    , public SparseOptimizer::ErrorEvaluator
#endif
{
 public:
  explicit OptimizationAlgorithmB200(int algo, const sgb_options* options = nullptr) : _algo(algo) {
    sgb_options o;
    sgb_default_options(&o);
    o.incremental = 1;  // the landmark graph grows by updateInitialization once per key-frame (drone.cpp:152-153)
    if (options) o = *options;
    _incremental = o.incremental != 0;
    sgb_status st = sgb_create(&o, &_h);
    if (st != SGB_OK) {
      std::fprintf(stderr, "OptimizationAlgorithmB200: sgb_create failed with status %d (no CUDA device? there is no CPU fallback)\n", (int)st);
      _h = nullptr;
    }
  }
  ~OptimizationAlgorithmB200() override { if (_h) sgb_destroy(_h); }
  OptimizationAlgorithmB200(const OptimizationAlgorithmB200&) = delete;
  OptimizationAlgorithmB200& operator=(const OptimizationAlgorithmB200&) = delete;

  // OptimizationAlgorithmWithHessian::init: here the whole symbolic phase (index mapping is the optimizer's)
  bool init(bool online = false) override {
    if (!_h || !_optimizer) return false;
#ifndef SGB_USE_REAL_G2O
    _optimizer->setErrorEvaluator(this);
#endif
    // online: updateStructure() has already extended the device-resident graph by the new vertices and edges
    if (online && _extended) {
      _extended = false;
      return true;
    }
    _extended = false;
    return upload(false);
  }

  SolverResult solve(int iteration, bool online = false) override {
    if (!_h) return Fail;
    if (iteration > 0 || !_fresh) {
      // estimates may have been changed on the host between solve() calls (host-side push/pop): re-send them
      if (!sendEstimates()) return Fail;
    }
    _fresh = false;
    int32_t result = SGB_RESULT_FAIL;
    sgb_iter_stat stat;
    sgb_status st = sgb_step(_h, _algo, iteration, online ? 1 : 0, &result, &stat);
    if (st != SGB_OK) {
      std::fprintf(stderr, "OptimizationAlgorithmB200::solve: %s\n", sgb_last_error(_h));
      return Fail;
    }
    _last = stat;
    if (!fetchEstimates()) return Fail;
    return result == SGB_RESULT_OK ? OK : (result == SGB_RESULT_TERMINATE ? Terminate : Fail);
  }

  // OptimizationAlgorithmWithHessian::updateStructure (reached from SparseOptimizer::updateInitialization, drone.cpp:153):
  // the optimizer's active sets already contain the new vertices and edges. The active graph is flattened again on the
  // host; when the graph the device holds is a prefix of it (the reference only ever appends: new pose, new landmarks,
  // new edges) only the tail travels (sgb_update_graph) and the estimates of the old vertices stay on the device.
  bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) override {
    if (!_h || !_optimizer) return false;
    _extended = false;
    if (!_incremental || !_uploaded) return true;   // nothing resident to extend: init() uploads the whole graph
    if (!upload(true)) return false;
    _extended = true;
    return true;
  }

  // Pure virtual in g2o::OptimizationAlgorithm (libg2o 2020.5.29, g2o/core/optimization_algorithm.h); the algorithms the
  // reference installs (graphs.cpp:9-23) implement it as BlockSolver::computeMarginals -> LinearSolver::solvePattern.
  // blockIndices holds (row, col) Hessian indices of the optimizer's index mapping; spinv is re-created with the
  // optimizer's block layout and receives one dense block per request: the block of the inverse of the Hessian at the
  // vertices' current estimates, each column solved for on the device (sgb_compute_marginals).
  bool computeMarginals(SparseBlockMatrix<MatrixX>& spinv, const std::vector<std::pair<int, int>>& blockIndices) override {
    if (!_h || !_optimizer) return false;
    const auto& iv = _optimizer->indexMapping();
    const int n = (int)iv.size();
    if (n == 0) return false;
    if (!_uploaded && !upload()) return false;   // computeMarginals before any optimize(): flatten the graph now
    if (_hidx_of.empty() && !buildHessianIndexMap()) return false;
    if (!sendEstimates()) return false;
    std::vector<int> rbi(n);
    int acc = 0;
    for (int i = 0; i < n; ++i) { acc += iv[i]->dimension(); rbi[i] = acc; }
    spinv = SparseBlockMatrix<MatrixX>(rbi.data(), rbi.data(), n, n, true);
    std::vector<int32_t> br, bc;
    for (const auto& rc : blockIndices) {
      if (rc.first < 0 || rc.first >= n || rc.second < 0 || rc.second >= n) return false;
      auto ir = _hidx_of.find(iv[rc.first]), ic = _hidx_of.find(iv[rc.second]);
      if (ir == _hidx_of.end() || ic == _hidx_of.end()) return false;
      br.push_back(ir->second);
      bc.push_back(ic->second);
    }
    size_t total = 0;
    for (const auto& rc : blockIndices) total += (size_t)iv[rc.first]->dimension() * iv[rc.second]->dimension();
    std::vector<double> vals(total ? total : 1);
    sgb_status st = sgb_compute_marginals(_h, (int32_t)br.size(), br.data(), bc.data(), vals.data());
    if (st != SGB_OK) {
      std::fprintf(stderr, "OptimizationAlgorithmB200::computeMarginals: %s\n", sgb_last_error(_h));
      return false;
    }
    size_t o = 0;
    for (const auto& rc : blockIndices) {
      auto* blk = spinv.block(rc.first, rc.second, true);
      const int nr = iv[rc.first]->dimension(), nc = iv[rc.second]->dimension();
      for (int j = 0; j < nc; ++j)
        for (int i = 0; i < nr; ++i) (*blk)(i, j) = vals[o + (size_t)j * nr + i];
      o += (size_t)nr * nc;
    }
    return true;
  }

  const sgb_iter_stat& lastIteration() const { return _last; }
  sgb_handle* handle() const { return _h; }

  // computeActiveErrors()/activeChi2() of the caller's gate (drone.cpp:164-165), evaluated on the device
  bool chi2(number_t* plain, number_t* robust)
#ifndef SGB_USE_REAL_G2O
      override
#endif
  {
    if (!_h || !sendEstimates()) return false;
    double c[2] = {0, 0};
    if (sgb_chi2(_h, c) != SGB_OK) return false;
    *plain = c[0];
    *robust = c[1];
    return true;
  }

 private:
  // flattens the optimizer's active graph; extend == true: send only what lies beyond the graph the device holds, if that
  // graph is a prefix of the new one (same vertices and edges in the same order), otherwise the whole graph
  bool upload(bool extend = false) {
    const auto& verts = _optimizer->activeVertices();
    const auto& edges = _optimizer->activeEdges();
    std::vector<OptimizableGraph::Vertex*> old_poses, old_lms;
    old_poses.swap(_poses); old_lms.swap(_lms);
    std::vector<const HyperGraph::Edge*> old_pp, old_pl;
    old_pp.swap(_pp_edges); old_pl.swap(_pl_edges);
    std::unordered_map<const HyperGraph::Vertex*, int> pidx, lidx;
    std::vector<int32_t> pose_id, lm_id;
    std::vector<uint8_t> pose_fixed, lm_fixed;
    for (auto* v : verts) {
      if (v->dimension() == 3) { pidx[v] = (int)_poses.size(); _poses.push_back(v); pose_id.push_back(v->id()); pose_fixed.push_back(v->fixed()); }
      else if (v->dimension() == 2) { lidx[v] = (int)_lms.size(); _lms.push_back(v); lm_id.push_back(v->id()); lm_fixed.push_back(v->fixed()); }
      else { std::fprintf(stderr, "OptimizationAlgorithmB200: unsupported vertex dimension %d\n", v->dimension()); return false; }
    }
    _pose_est.assign(3 * _poses.size(), 0.0);
    _lm_est.assign(2 * _lms.size(), 0.0);
    gatherEstimates();
    std::vector<int32_t> pp_i, pp_j, pl_p, pl_l;
    std::vector<double> pp_z, pp_info, pp_phi, pl_z, pl_info;
    std::vector<int64_t> pp_seq, pl_seq;
    for (auto* e : edges) {
      if (e->dimension() == 3) {
        auto* ee = static_cast<EdgeSE2*>(e);
        pp_i.push_back(pidx.at(e->vertex(0))); pp_j.push_back(pidx.at(e->vertex(1)));
        const SE2& z = ee->measurement();
        pp_z.insert(pp_z.end(), {z[0], z[1], z[2]});
        const auto& I = ee->information();
        pp_info.insert(pp_info.end(), {I(0, 0), I(0, 1), I(0, 2), I(1, 1), I(1, 2), I(2, 2)});
        pp_phi.push_back(e->robustKernel() ? e->robustKernel()->delta() : 0.0);  // the reference only uses RobustKernelDCS
        pp_seq.push_back(e->internalId());
        _pp_edges.push_back(e);
      } else if (e->dimension() == 2) {
        auto* ee = static_cast<EdgeSE2RhoTheta*>(e);
        pl_p.push_back(pidx.at(e->vertex(0))); pl_l.push_back(lidx.at(e->vertex(1)));
        const auto& z = ee->measurement();
        pl_z.insert(pl_z.end(), {z[0], z[1]});
        const auto& I = ee->information();
        pl_info.insert(pl_info.end(), {I(0, 0), I(0, 1), I(1, 1)});
        pl_seq.push_back(e->internalId());
        _pl_edges.push_back(e);
      } else { std::fprintf(stderr, "OptimizationAlgorithmB200: unsupported edge dimension %d\n", e->dimension()); return false; }
    }
    auto is_prefix = [](const auto& a, const auto& b) { return a.size() <= b.size() && std::equal(a.begin(), a.end(), b.begin()); };
    if (extend && _uploaded && is_prefix(old_poses, _poses) && is_prefix(old_lms, _lms) && is_prefix(old_pp, _pp_edges) &&
        is_prefix(old_pl, _pl_edges)) {
      const size_t P0 = old_poses.size(), L0 = old_lms.size(), E0 = old_pp.size(), F0 = old_pl.size();
      sgb_graph_delta d{};
      d.n_new_poses = (int32_t)(_poses.size() - P0); d.pose_id = pose_id.data() + P0; d.pose_est = _pose_est.data() + 3 * P0; d.pose_fixed = pose_fixed.data() + P0;
      d.n_new_landmarks = (int32_t)(_lms.size() - L0); d.lm_id = lm_id.data() + L0; d.lm_est = _lm_est.data() + 2 * L0; d.lm_fixed = lm_fixed.data() + L0;
      d.n_new_pp = (int32_t)(pp_i.size() - E0); d.pp_i = pp_i.data() + E0; d.pp_j = pp_j.data() + E0; d.pp_z = pp_z.data() + 3 * E0;
      d.pp_info = pp_info.data() + 6 * E0; d.pp_phi = pp_phi.data() + E0; d.pp_seq = pp_seq.data() + E0;
      d.n_new_pl = (int32_t)(pl_p.size() - F0); d.pl_pose = pl_p.data() + F0; d.pl_lm = pl_l.data() + F0; d.pl_z = pl_z.data() + 2 * F0;
      d.pl_info = pl_info.data() + 3 * F0; d.pl_seq = pl_seq.data() + F0;
      sgb_status su = sgb_update_graph(_h, &d);
      if (su != SGB_OK) { std::fprintf(stderr, "OptimizationAlgorithmB200::updateStructure: %s\n", sgb_last_error(_h)); return false; }
      // the host may have moved old estimates since the last solve (pop() after a rejected key-frame): solve() re-sends them
      _hidx_of.clear();
      _fresh = false;
      return true;
    }
    sgb_graph_soa g{};
    g.n_poses = (int32_t)_poses.size(); g.pose_id = pose_id.data(); g.pose_est = _pose_est.data(); g.pose_fixed = pose_fixed.data();
    g.n_landmarks = (int32_t)_lms.size(); g.lm_id = lm_id.data(); g.lm_est = _lm_est.data(); g.lm_fixed = lm_fixed.data();
    g.n_pp = (int32_t)pp_i.size(); g.pp_i = pp_i.data(); g.pp_j = pp_j.data(); g.pp_z = pp_z.data(); g.pp_info = pp_info.data();
    g.pp_phi = pp_phi.data(); g.pp_seq = pp_seq.data();
    g.n_pl = (int32_t)pl_p.size(); g.pl_pose = pl_p.data(); g.pl_lm = pl_l.data(); g.pl_z = pl_z.data(); g.pl_info = pl_info.data();
    g.pl_seq = pl_seq.data();
    sgb_status st = sgb_set_graph(_h, &g);
    if (st != SGB_OK) { std::fprintf(stderr, "OptimizationAlgorithmB200::init: %s\n", sgb_last_error(_h)); return false; }
    _hidx_of.clear();  // rebuilt on demand by computeMarginals
    _uploaded = true;
    _fresh = true;
    return true;
  }
  // vertex -> Hessian index of the backend's structure (what sgb_compute_marginals addresses blocks by)
  bool buildHessianIndexMap() {
    std::vector<int32_t> ph(_poses.size()), lh(_lms.size());
    if (sgb_get_structure(_h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ph.data(), lh.data()) != SGB_OK) return false;
    _hidx_of.clear();
    for (size_t i = 0; i < _poses.size(); ++i) if (ph[i] >= 0) _hidx_of[_poses[i]] = ph[i];
    for (size_t i = 0; i < _lms.size(); ++i) if (lh[i] >= 0) _hidx_of[_lms[i]] = lh[i];
    return true;
  }
  void gatherEstimates() {
    for (size_t i = 0; i < _poses.size(); ++i) _poses[i]->getEstimateData(&_pose_est[3 * i]);
    for (size_t i = 0; i < _lms.size(); ++i) _lms[i]->getEstimateData(&_lm_est[2 * i]);
  }
  bool sendEstimates() {
    gatherEstimates();
    return sgb_set_estimates(_h, _pose_est.data(), _lm_est.data()) == SGB_OK;
  }
  bool fetchEstimates() {
    if (sgb_get_estimates(_h, _pose_est.data(), _lm_est.data()) != SGB_OK) return false;
    for (size_t i = 0; i < _poses.size(); ++i) if (!_poses[i]->fixed()) _poses[i]->setEstimateData(&_pose_est[3 * i]);
    for (size_t i = 0; i < _lms.size(); ++i) if (!_lms[i]->fixed()) _lms[i]->setEstimateData(&_lm_est[2 * i]);
    return true;
  }

  int _algo;
  sgb_handle* _h = nullptr;
  bool _fresh = false, _uploaded = false, _incremental = false, _extended = false;
  std::vector<const HyperGraph::Edge*> _pp_edges, _pl_edges;   // edges of the device-resident graph, in its order
  std::vector<OptimizableGraph::Vertex*> _poses, _lms;
  std::vector<double> _pose_est, _lm_est;
  std::unordered_map<const HyperGraph::Vertex*, int32_t> _hidx_of;
  sgb_iter_stat _last{};
};

// The narrower drop-in: only the innermost object of graphs.cpp:11,19 changes,
//     g2o::make_unique<LinearSolverEigen<PoseMatrixType>>()   becomes   g2o::make_unique<LinearSolverB200<PoseMatrixType>>()
// g2o's BlockSolver keeps linearisation, damping and the LM logic on the host and hands the damped block matrix over
// once per trial; the solve itself (Schur elimination + PCG) runs on the device. init() is called once per optimize()
// (like LinearSolverEigen's, it forgets the pattern analysis).
template <typename MatrixType>
class LinearSolverB200 : public LinearSolver<MatrixType> {
 public:
  explicit LinearSolverB200(const sgb_options* options = nullptr) {
    sgb_options o;
    sgb_default_options(&o);
    if (options) o = *options;
    if (sgb_create(&o, &_h) != SGB_OK) {
      std::fprintf(stderr, "LinearSolverB200: sgb_create failed (no CUDA device? there is no CPU fallback)\n");
      _h = nullptr;
    }
  }
  ~LinearSolverB200() override { if (_h) sgb_destroy(_h); }
  LinearSolverB200(const LinearSolverB200&) = delete;
  LinearSolverB200& operator=(const LinearSolverB200&) = delete;

  bool init() override {
    _pattern = false;
    return _h != nullptr;
  }
  bool solve(const SparseBlockMatrix<MatrixType>& A, number_t* x, number_t* b) override {
    if (!_h) return false;
    const auto& cols = A.blockCols();
    const int n = (int)cols.size();
    if (!_pattern) {
      std::vector<int32_t> dim(n), col_ptr(n + 1, 0), row_idx;
      for (int c = 0; c < n; ++c) {
        dim[c] = A.colsOfBlock(c);
        for (const auto& kv : cols[c])
          if (kv.first <= c) row_idx.push_back(kv.first);  // upper triangle, rows ascending (std::map order)
        col_ptr[c + 1] = (int32_t)row_idx.size();
      }
      sgb_block_matrix M{n, dim.data(), col_ptr.data(), row_idx.data()};
      if (sgb_linear_set_pattern(_h, &M) != SGB_OK) {
        std::fprintf(stderr, "LinearSolverB200::solve: %s\n", sgb_last_error(_h));
        return false;
      }
      _pattern = true;
    }
    _values.clear();
    for (int c = 0; c < n; ++c)
      for (const auto& kv : cols[c]) {
        if (kv.first > c) continue;
        const auto& blk = *kv.second;  // column-major
        _values.insert(_values.end(), blk.data(), blk.data() + (size_t)blk.rows() * blk.cols());
      }
    sgb_status st = sgb_linear_solve(_h, _values.data(), b, x, &_pcg_iters, nullptr);
    if (st != SGB_OK && st != SGB_ERR_SOLVE_FAILED) std::fprintf(stderr, "LinearSolverB200::solve: %s\n", sgb_last_error(_h));
    return st == SGB_OK;
  }
  int lastPcgIterations() const { return _pcg_iters; }

 private:
  sgb_handle* _h = nullptr;
  bool _pattern = false;
  int32_t _pcg_iters = 0;
  std::vector<double> _values;
};

}  // namespace g2o
