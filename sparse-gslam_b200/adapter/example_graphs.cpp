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
  return (n > 0 && chi2_after < 400.0) ? 0 : 2;
}
