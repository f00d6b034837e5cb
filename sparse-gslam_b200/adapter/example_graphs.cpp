// example_graphs.cpp -- what the reference's src/sparse_gslam/src/graphs.cpp + the optimiser calls of
// drone.cpp:146-165 look like on the B200 backend. Builds a small landmark graph through the g2o-style API, runs
// initializeOptimization(); push(); optimize(15); computeActiveErrors(); activeChi2(); computeMarginals, then the pose
// graph of submap_loop_closer.cpp:204-288 through setup_pose_opt + optimize(20) with DCS closures, and prints the results.
// With a file prefix as argument it also dumps both graphs (g2o text) and the final estimates: tests/test_adapter.py
// feeds the dumped graphs to the CPU oracle and compares.
// Compile:  g++ -std=c++17 example_graphs.cpp -I../../include -L.. -lsgb -Wl,-rpath,'$ORIGIN/..' -o example_graphs
#include <cstdio>
#include <deque>
#include <random>
#include <string>
#include <vector>

#include "sgb_g2o_adapter.h"

// ---- graphs.cpp:9-23 with the two changed lines
void setup_lm_opt(g2o::SparseOptimizer& opt) {
  opt.setAlgorithm(new g2o::OptimizationAlgorithmB200(SGB_ALGO_LM));
  opt.setVerbose(false);
  opt.setComputeBatchStatistics(false);
}
void setup_pose_opt(g2o::SparseOptimizer& opt) {
  opt.setAlgorithm(new g2o::OptimizationAlgorithmB200(SGB_ALGO_GN));
  opt.setVerbose(false);
  opt.setComputeBatchStatistics(false);
}

// ---- the narrower drop-in: LinearSolverB200 behind g2o's own BlockSolver (graphs.cpp:11 / :19 with one changed word)
static int linear_solver_example() {
  // H of 3 poses + 2 landmarks: block-tridiagonal pose part, every pose sees landmark 0, pose 2 sees landmark 1
  const int rbi[5] = {3, 6, 9, 11, 13};
  g2o::SparseBlockMatrix<g2o::MatrixX> A(rbi, rbi, 5, 5);
  std::mt19937 rng(7);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  auto fill = [&](int r, int c) {
    auto* b = A.block(r, c, true);
    for (int j = 0; j < b->cols(); ++j)
      for (int i = 0; i < b->rows(); ++i) (*b)(i, j) = (r == c) ? ((i == j) ? 12.0 : ((i < j) ? 0.3 * (i + j + 1) : 0.0)) : u(rng);
    if (r == c)
      for (int j = 0; j < b->cols(); ++j)
        for (int i = j + 1; i < b->rows(); ++i) (*b)(i, j) = (*b)(j, i);  // symmetric diagonal blocks
  };
  for (int k = 0; k < 5; ++k) fill(k, k);
  fill(0, 1); fill(1, 2); fill(0, 3); fill(1, 3); fill(2, 3); fill(2, 4);
  double b[13], x[13], dense[13][13] = {};
  for (double& v : b) v = u(rng);
  for (int c = 0; c < 5; ++c)
    for (const auto& kv : A.blockCols()[c])
      for (int j = 0; j < kv.second->cols(); ++j)
        for (int i = 0; i < kv.second->rows(); ++i) {
          int gi = A.rowBaseOfBlock(kv.first) + i, gj = A.colBaseOfBlock(c) + j;
          dense[gi][gj] = dense[gj][gi] = (*kv.second)(i, j);
        }
  g2o::LinearSolverB200<g2o::MatrixX> solver;
  if (!solver.init() || !solver.solve(A, x, b)) return 3;
  double worst = 0.0;
  for (int i = 0; i < 13; ++i) {
    double r = -b[i];
    for (int j = 0; j < 13; ++j) r += dense[i][j] * x[j];
    worst = std::max(worst, std::fabs(r));
  }
  std::printf("LinearSolverB200: %d PCG iterations, max |Ax - b| = %.3e\n", solver.lastPcgIterations(), worst);
  return worst < 1e-8 ? 0 : 4;
}

// ---- dumps: the graph as g2o text (the wire format of sgb_g2o_load, line order = insertion order) and the estimates, so that
// the parity test can hand the very same graph to the CPU oracle and compare final states (tests/test_adapter.py)
static void dump_vertices(std::FILE* f, std::deque<g2o::VertexSE2>& poses, std::deque<g2o::VertexRhoTheta>& lms) {
  for (auto& v : poses) std::fprintf(f, "VERTEX_SE2 %d %.17g %.17g %.17g\n", v.id(), v.estimate()[0], v.estimate()[1], v.estimate()[2]);
  for (auto& v : lms) std::fprintf(f, "VERTEX_RHOTHETA %d %.17g %.17g\n", v.id(), v.estimate()[0], v.estimate()[1]);
  for (auto& v : poses) if (v.fixed()) std::fprintf(f, "FIX %d\n", v.id());
}
static void dump_pp(std::FILE* f, g2o::EdgeSE2& e) {
  const auto& I = e.information();
  std::fprintf(f, "EDGE_SE2 %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", e.vertex(0)->id(), e.vertex(1)->id(),
               e.measurement()[0], e.measurement()[1], e.measurement()[2], I(0, 0), I(0, 1), I(0, 2), I(1, 1), I(1, 2), I(2, 2));
}
static void dump_pl(std::FILE* f, g2o::EdgeSE2RhoTheta& e) {
  const auto& I = e.information();
  std::fprintf(f, "EDGE_SE2_RHOTHETA %d %d %.17g %.17g %.17g %.17g %.17g\n", e.vertex(0)->id(), e.vertex(1)->id(),
               e.measurement()[0], e.measurement()[1], I(0, 0), I(0, 1), I(1, 1));
}
static void dump_estimates(std::FILE* f, const char* tag, std::deque<g2o::VertexSE2>& poses, std::deque<g2o::VertexRhoTheta>& lms) {
  for (auto& v : poses) std::fprintf(f, "%s POSE %d %.17g %.17g %.17g\n", tag, v.id(), v.estimate()[0], v.estimate()[1], v.estimate()[2]);
  for (auto& v : lms) std::fprintf(f, "%s LINE %d %.17g %.17g\n", tag, v.id(), v.estimate()[0], v.estimate()[1]);
}

// usage: example_graphs [dump-prefix]   (with a prefix: <prefix>_lm.g2o, <prefix>_pose.g2o, <prefix>_result.txt are written)
int main(int argc, char** argv) {
  const std::string prefix = argc > 1 ? argv[1] : "";
  std::FILE* res = prefix.empty() ? nullptr : std::fopen((prefix + "_result.txt").c_str(), "w");
  g2o::SparseOptimizer opt;
  setup_lm_opt(opt);
  std::deque<g2o::VertexSE2> poses;
  std::deque<g2o::EdgeSE2> odom;
  std::deque<g2o::VertexRhoTheta> lms;
  std::deque<g2o::EdgeSE2RhoTheta> obs;
  std::mt19937 rng(1);
  std::normal_distribution<double> n01(0.0, 1.0);
  const int P = 40;
  // two walls: y = 2 (rho 2, alpha pi/2) and x = 25 (rho 25, alpha 0)
  const double walls[2][2] = {{2.0, 1.5707963267948966}, {25.0, 0.0}};
  for (int l = 0; l < 2; ++l) {
    lms.emplace_back();
    lms.back().setId(10000000 + l);
    g2o::Vector2 e; e[0] = walls[l][0] + 0.05; e[1] = walls[l][1] - 0.02;
    lms.back().setEstimate(e);
    opt.addVertex(&lms.back());
  }
  std::vector<std::pair<int, int>> edge_log;  // insertion order: (type, index)
  for (int k = 0; k < P; ++k) {
    poses.emplace_back();
    poses.back().setId(k);
    poses.back().setEstimate(g2o::SE2(0.5 * k + (k ? 0.03 * n01(rng) : 0.0), k ? 0.03 * n01(rng) : 0.0, k ? 0.01 * n01(rng) : 0.0));
    if (k == 0) poses.back().setFixed(true);
    opt.addVertex(&poses.back());
    if (k > 0) {
      odom.emplace_back();
      odom.back().vertices()[0] = &poses[k - 1];
      odom.back().vertices()[1] = &poses[k];
      odom.back().setMeasurement(g2o::SE2(0.5 + 0.02 * n01(rng), 0.02 * n01(rng), 0.01 * n01(rng)));
      odom.back().information()(0, 0) = 2500; odom.back().information()(1, 1) = 2500; odom.back().information()(2, 2) = 10000;
      opt.addEdge(&odom.back());
      edge_log.emplace_back(0, (int)odom.size() - 1);
    }
    for (int l = 0; l < 2; ++l) {
      obs.emplace_back();
      obs.back().vertices()[0] = &poses[k];
      obs.back().vertices()[1] = &lms[l];
      g2o::Vector2 z;
      z[0] = (l == 0 ? 2.0 : 25.0 - 0.5 * k) + 0.03 * n01(rng);
      z[1] = walls[l][1] + 0.02 * n01(rng);
      obs.back().setMeasurement(z);
      obs.back().information()(0, 0) = 1111; obs.back().information()(1, 1) = 2500;
      opt.addEdge(&obs.back());
      edge_log.emplace_back(1, (int)obs.size() - 1);
    }
  }
  if (!prefix.empty()) {
    std::FILE* f = std::fopen((prefix + "_lm.g2o").c_str(), "w");
    if (!f) return 10;
    dump_vertices(f, poses, lms);
    for (auto& tk : edge_log) tk.first == 0 ? dump_pp(f, odom[tk.second]) : dump_pl(f, obs[tk.second]);
    std::fclose(f);
  }
  // ---- drone.cpp:146-165
  if (!opt.initializeOptimization()) return 1;
  opt.push();
  int n = opt.optimize(15, false);
  opt.computeActiveErrors();
  double chi2_after = opt.activeChi2();
  std::printf("optimize returned %d, chi2 = %.6f, last pose = (%.4f, %.4f, %.4f)\n", n, chi2_after, poses.back().estimate()[0],
              poses.back().estimate()[1], poses.back().estimate()[2]);
  if (res) {
    std::fprintf(res, "LM ITERATIONS %d\nLM CHI2 %.17g\n", n, chi2_after);
    dump_estimates(res, "LM", poses, lms);
  }
  // ---- SparseOptimizer::computeMarginals -> OptimizationAlgorithm::computeMarginals (pure virtual in g2o): marginal
  // covariance blocks of the last pose, of the first landmark and their cross block
  {
    g2o::SparseBlockMatrix<g2o::MatrixX> spinv;
    const int hp = poses.back().hessianIndex(), hl = lms.front().hessianIndex();
    std::vector<std::pair<int, int>> want = {{hp, hp}, {hl, hl}, {hp, hl}};
    if (!opt.computeMarginals(spinv, want)) return 5;
    const g2o::MatrixX* bp = spinv.block(hp, hp);
    const g2o::MatrixX* bl = spinv.block(hl, hl);
    const g2o::MatrixX* bx = spinv.block(hp, hl);
    if (!bp || !bl || !bx || bp->rows() != 3 || bl->rows() != 2 || bx->rows() != 3 || bx->cols() != 2) return 6;
    std::printf("marginals: var(x, y, theta) of the last pose = (%.3e, %.3e, %.3e), var(rho, alpha) of wall 0 = (%.3e, %.3e)\n",
                (*bp)(0, 0), (*bp)(1, 1), (*bp)(2, 2), (*bl)(0, 0), (*bl)(1, 1));
    if (!((*bp)(0, 0) > 0 && (*bp)(1, 1) > 0 && (*bp)(2, 2) > 0 && (*bl)(0, 0) > 0 && (*bl)(1, 1) > 0)) return 7;
    if (std::fabs((*bp)(0, 1) - (*bp)(1, 0)) > 1e-9 * ((*bp)(0, 0) + (*bp)(1, 1))) return 8;  // symmetric
    if (res) {
      std::fprintf(res, "MARGINAL %d %d", hp, hp);
      for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) std::fprintf(res, " %.17g", (*bp)(i, j));
      std::fprintf(res, "\nMARGINAL %d %d", hl, hl);
      for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) std::fprintf(res, " %.17g", (*bl)(i, j));
      std::fprintf(res, "\nMARGINAL %d %d", hp, hl);
      for (int j = 0; j < 2; ++j) for (int i = 0; i < 3; ++i) std::fprintf(res, " %.17g", (*bx)(i, j));
      std::fprintf(res, "\n");
    }
  }
  opt.discardTop();
  // ---- the online path of drone.cpp:152-156: two more key-frames arrive; updateInitialization(new vertices, new edges);
  // push(); optimize(15, true). Only the new vertices and edges travel to the device (sgb_update_graph).
  for (int step = 0; step < 2; ++step) {
    g2o::HyperGraph::VertexSet vset;
    g2o::HyperGraph::EdgeSet eset;
    const int k = (int)poses.size();
    poses.emplace_back();
    poses.back().setId(k);
    poses.back().setEstimate(poses[k - 1].estimate() * g2o::SE2(0.5, 0.0, 0.0));
    opt.addVertex(&poses.back());
    vset.insert(&poses.back());
    odom.emplace_back();
    odom.back().vertices()[0] = &poses[k - 1];
    odom.back().vertices()[1] = &poses[k];
    odom.back().setMeasurement(g2o::SE2(0.5 + 0.02 * n01(rng), 0.02 * n01(rng), 0.01 * n01(rng)));
    odom.back().information()(0, 0) = 2500; odom.back().information()(1, 1) = 2500; odom.back().information()(2, 2) = 10000;
    opt.addEdge(&odom.back());
    eset.insert(&odom.back());
    edge_log.emplace_back(0, (int)odom.size() - 1);
    for (int l = 0; l < 2; ++l) {
      obs.emplace_back();
      obs.back().vertices()[0] = &poses[k];
      obs.back().vertices()[1] = &lms[l];
      g2o::Vector2 z;
      z[0] = (l == 0 ? 2.0 : 25.0 - 0.5 * k) + 0.03 * n01(rng);
      z[1] = walls[l][1] + 0.02 * n01(rng);
      obs.back().setMeasurement(z);
      obs.back().information()(0, 0) = 1111; obs.back().information()(1, 1) = 2500;
      opt.addEdge(&obs.back());
      eset.insert(&obs.back());
      edge_log.emplace_back(1, (int)obs.size() - 1);
    }
    if (step == 1 && !prefix.empty()) {   // the graph the last online optimize() starts from
      std::FILE* f = std::fopen((prefix + "_online.g2o").c_str(), "w");
      if (!f) return 10;
      dump_vertices(f, poses, lms);
      for (auto& tk : edge_log) tk.first == 0 ? dump_pp(f, odom[tk.second]) : dump_pl(f, obs[tk.second]);
      std::fclose(f);
    }
    if (!opt.updateInitialization(vset, eset)) return 13;
    opt.push();
    int non = opt.optimize(15, true);
    opt.computeActiveErrors();
    std::printf("online key-frame %d: optimize returned %d, chi2 = %.6f\n", k, non, opt.activeChi2());
    if (non <= 0) return 14;
    opt.discardTop();
    if (step == 1 && res) {
      std::fprintf(res, "ON ITERATIONS %d\nON CHI2 %.17g\n", non, opt.activeChi2());
      dump_estimates(res, "ON", poses, lms);
    }
  }
  delete opt.algorithm();
  if (!(n > 0 && chi2_after < 400.0)) return 2;

  // ---- the pose graph: setup_pose_opt + submap_loop_closer.cpp:204-288 (poses and odometry copied from the optimised
  // landmark graph, DCS closure edges) + optimize(20)
  g2o::SparseOptimizer popt;
  setup_pose_opt(popt);
  std::deque<g2o::VertexSE2> pposes;
  std::deque<g2o::EdgeSE2> pedges;
  std::deque<g2o::RobustKernelDCS> kernels;
  for (int k = 0; k < P; ++k) {
    pposes.emplace_back();
    pposes.back().setId(k);
    pposes.back().setEstimate(poses[k].estimate());
    if (k == 0) pposes.back().setFixed(true);
    popt.addVertex(&pposes.back());
    if (k > 0) {
      pedges.emplace_back();
      pedges.back().vertices()[0] = &pposes[k - 1];
      pedges.back().vertices()[1] = &pposes[k];
      pedges.back().setMeasurement(poses[k - 1].estimate().inverse() * poses[k].estimate());   // submap_loop_closer.cpp:216
      pedges.back().information() = odom[k - 1].information();                                // :217
      popt.addEdge(&pedges.back());
    }
  }
  const int closures[3][2] = {{2, 30}, {5, 38}, {10, 25}};
  for (int c = 0; c < 3; ++c) {
    const int a = closures[c][0], b = closures[c][1];
    pedges.emplace_back();
    pedges.back().vertices()[0] = &pposes[a];
    pedges.back().vertices()[1] = &pposes[b];
    g2o::SE2 rel = poses[a].estimate().inverse() * poses[b].estimate();
    // the last closure is a false one (far off): DCS must down-weight it
    pedges.back().setMeasurement(g2o::SE2(rel[0] + (c == 2 ? 1.5 : 0.02 * n01(rng)), rel[1] + (c == 2 ? -1.0 : 0.02 * n01(rng)), rel[2] + 0.01 * n01(rng)));
    pedges.back().information()(0, 0) = 400; pedges.back().information()(1, 1) = 400; pedges.back().information()(2, 2) = 2500;
    kernels.emplace_back();
    kernels.back().setDelta(0.75);                                                            // datasets/mit-killian/slam.yaml:38
    pedges.back().setRobustKernel(&kernels.back());                                           // submap_loop_closer.cpp:279-281
    popt.addEdge(&pedges.back());
  }
  if (!prefix.empty()) {
    std::FILE* f = std::fopen((prefix + "_pose.g2o").c_str(), "w");
    if (!f) return 10;
    std::deque<g2o::VertexRhoTheta> none;
    dump_vertices(f, pposes, none);
    for (auto& e : pedges) dump_pp(f, e);
    int k = 0;
    for (auto& e : pedges) { if (e.robustKernel()) std::fprintf(f, "ROBUST_KERNEL_DCS %d %.17g\n", k, e.robustKernel()->delta()); ++k; }
    std::fclose(f);
  }
  if (!popt.initializeOptimization()) return 11;
  int np = popt.optimize(20);
  popt.computeActiveErrors();
  std::printf("pose graph: optimize returned %d, chi2 = %.6f (robust %.6f)\n", np, popt.activeChi2(), popt.activeRobustChi2());
  if (res) {
    std::fprintf(res, "GN ITERATIONS %d\nGN CHI2 %.17g\n", np, popt.activeChi2());
    std::deque<g2o::VertexRhoTheta> none;
    dump_estimates(res, "GN", pposes, none);
    std::fclose(res);
  }
  delete popt.algorithm();
  if (np != 20) return 12;
  return linear_solver_example();
}
