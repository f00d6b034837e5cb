// example_graphs.cpp -- what the reference's src/sparse_gslam/src/graphs.cpp + the optimiser calls of
// drone.cpp:146-165 look like on the B200 backend. Builds a small landmark graph through the g2o-style API, runs
// initializeOptimization(); push(); optimize(15); computeActiveErrors(); activeChi2(), and prints the result.
// Compile:  g++ -std=c++17 example_graphs.cpp -I../../include -L.. -lsgb -Wl,-rpath,'$ORIGIN/..' -o example_graphs
#include <cstdio>
#include <deque>
#include <random>

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

int main() {
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
    }
  }
  // ---- drone.cpp:146-165
  if (!opt.initializeOptimization()) return 1;
  opt.push();
  int n = opt.optimize(15, false);
  opt.computeActiveErrors();
  double chi2_after = opt.activeChi2();
  std::printf("optimize returned %d, chi2 = %.6f, last pose = (%.4f, %.4f, %.4f)\n", n, chi2_after, poses.back().estimate()[0],
              poses.back().estimate()[1], poses.back().estimate()[2]);
  opt.discardTop();
  delete opt.algorithm();
  if (!(n > 0 && chi2_after < 400.0)) return 2;
  return linear_solver_example();
}
