// sgb_coarse.h -- coarse space of the two-level preconditioner of the reduced pose system (single GPU, small graphs).
//
// The block-Jacobi preconditioner (setup_chunk, sgb_rows.h) removes the stiff local modes of the Schur complement
// S = (Hpp + lambda I) - Hpl (Hll + lambda I)^-1 Hpl^T; what the PCG then spends its iterations on are the smooth
// modes ALONG the trajectory (a chain of key-frames bends as a whole), and the iteration count grows with the chain
// length. The reference has no counterpart: LinearSolverEigen factorises exactly. The product adds, to the block-Jacobi
// term, the exact solve on a coarse space of piecewise-LINEAR functions over the pose chain (pose ids are temporal,
// drone.cpp:121): a node every h consecutive pose rows, every pose row i interpolating between its two neighbours,
//
//     M^-1 = blockJacobi^-1 + R^T (R S R^T)^-1 R,      R[node n][row i] = hat_n(i) I3,
//
// a symmetric positive definite operator like the one it extends (additive two-level Schwarz), so the contract of the
// solve is unchanged: the PCG converges to the same x at the same tolerance, in 2-5x fewer iterations on chain-shaped
// graphs of 50-1200 poses (tests/experiments/precond_study.py). nn <= kCzMaxNodes nodes, coarse dimension
// nc = 3 nn <= 120: the coarse matrix is formed (from host-built gather lists: deterministic sums, no atomics),
// factorised and explicitly inverted by ONE CTA once per LM trial (k_setup_coarse), and applied as a dense nc x nc
// matrix-vector product from shared memory inside the resident PCG kernel (sgb_resident.cuh).
//
// The bodies below are shared by the kernels and by the host-side test harness (tests/hostsim): every routine takes
// (tid, nth) and a barrier functor -- the CTA's threads with __syncthreads on the device, (0, 1) and a no-op on the
// host -- so the harness executes the very code the device runs, element by element.
#pragma once
#include "sgb_rows.h"

namespace sgb {

constexpr int kCzMaxNodes = 40;            // coarse dimension <= 120
constexpr int kCzMaxDim = 3 * kCzMaxNodes;

// leading dimension of the shared-memory copies of an nc x nc matrix of doubles: odd, so that rows AND columns are
// read conflict-free (a double spans two banks: stride s doubles is conflict-free iff s is odd)
SGB_HD int cz_ld(int nc) { return nc | 1; }

// hat-function weights of pose row i: left node i / h with weight 1 - (i % h) / h, right node i / h + 1
SGB_HD double cz_wr(int i, int h) { return (double)(i % h) / (double)h; }

// G[gi] = sum of weight x (3x2 block of Hpl) over the items of G-nonzero gi (one (node, landmark) pair), row-major
SGB_HD void coarse_g_row(const DevGraph& g, int gi) {
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int q = g.cz_g_ptr[gi]; q < g.cz_g_ptr[gi + 1]; ++q) {
    const int e = g.cz_g_e[q];
    const double w = g.cz_g_w[q];
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[c] += w * g.Hpl.vals[sell_vaddr(e, 6, c)];
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) g.cz_G[6 * (size_t)gi + c] = acc[c];
}

// 3x3 block (node m, node n), m <= n, of R S R^T, written into the LOWER triangle of A (nc x nc, leading dimension ld):
// the Hpp items, lambda R R^T, minus the Schur terms G[g1] W_l G[g2]^T of the landmarks both nodes reach
SGB_HD void coarse_block(const DevGraph& g, int m, int n, double lambda, double* A, int ld) {
  const int b = m * g.cz_nn + n;
  double acc[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int q = g.cz_p_ptr[b]; q < g.cz_p_ptr[b + 1]; ++q) {
    const int e = g.cz_p_e[q];
    const double w = g.cz_p_w[q];
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] += w * g.Hpp.vals[sell_vaddr(e, 9, c)];
  }
  const double rr = lambda * g.cz_rr[b];
  acc[0] += rr; acc[4] += rr; acc[8] += rr;
  const double* W = g.Hll_inv[g.rank];
  for (int q = g.cz_t_ptr[b]; q < g.cz_t_ptr[b + 1]; ++q) {
    const int g1 = g.cz_t_g[2 * (size_t)q], g2 = g.cz_t_g[2 * (size_t)q + 1];
    const int l = g.cz_g_lm[g1];
    const double w11 = W[l], w12 = W[(size_t)g.capL + l], w22 = W[2 * (size_t)g.capL + l];
    const double* G1 = g.cz_G + 6 * (size_t)g1;
    const double* G2 = g.cz_G + 6 * (size_t)g2;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double bw0 = G1[2 * r] * w11 + G1[2 * r + 1] * w12, bw1 = G1[2 * r] * w12 + G1[2 * r + 1] * w22;
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[3 * r + c] -= bw0 * G2[2 * c] + bw1 * G2[2 * c + 1];
    }
  }
  // block (m, n) of the symmetric matrix lives at rows 3m.., columns 3n..; its transpose is the lower-triangle copy
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int row = 3 * n + c, col = 3 * m + r;
      if (row >= col) A[(size_t)row * ld + col] = acc[3 * r + c];
    }
}

// One CTA (or the host with tid = 0, nth = 1): coarse matrix -> its explicit inverse.
//   1. G values, then the lower triangle of A = R S R^T in `A` (nc x ld, shared memory on the device);
//      a node no pose row reaches (the last one when nP = 1 mod h) gets an identity block
//   2. Cholesky A = L L^T in place (right-looking, two barriers per column), 1 / L_kk in dinv
//   3. X = L^-1, one column per thread, stored TRANSPOSED in the strict upper triangle of the same array
//   4. out = X^T X (nc x nc, row-major, both triangles): symmetric positive (semi-)definite by construction whatever the
//      rounding of X -- what a preconditioner must be; a Gauss-Jordan inverse is only as definite as cond(A) eps allows
// A pivot that is not positive / not finite switches the coarse term off for this solve (out = 0, *fail = 1).
template <class Barrier>
SGB_HD void coarse_factor(const DevGraph& g, double lambda, double* A, double* dinv, double* out, int* fail, int tid,
                                 int nth, Barrier barrier) {
  const int nn = g.cz_nn, nc = 3 * nn, ld = cz_ld(nc);
  for (int gi = tid; gi < g.cz_ng; gi += nth) coarse_g_row(g, gi);
  for (int i = tid; i < nc * ld; i += nth) A[i] = 0.0;
  barrier();
  for (int b = tid; b < nn * nn; b += nth) {
    const int m = b / nn, n = b % nn;
    if (m <= n) coarse_block(g, m, n, lambda, A, ld);
  }
  barrier();
  for (int d = tid; d < nc; d += nth)
    if (A[(size_t)d * ld + d] == 0.0) A[(size_t)d * ld + d] = 1.0;  // unreached node: nothing else in its row / column
  bool ok = true;
  const int sx = nth >= 32 ? 32 : 1, sy = nth / sx, tx = tid % sx, ty = tid / sx;  // nth = 1 or a multiple of 32
  for (int k = 0; k < nc; ++k) {
    barrier();
    const double p = A[(size_t)k * ld + k];
    if (!(p > 0.0) || !(p < 1e300)) { ok = false; break; }  // the same value in every thread: a uniform exit
    const double inv = 1.0 / sqrt(p);
    for (int i = k + 1 + tid; i < nc; i += nth) A[(size_t)i * ld + k] *= inv;
    barrier();
    for (int i = k + 1 + ty; i < nc; i += sy) {  // trailing update of the lower triangle, a 2-D tile of threads
      const double lik = A[(size_t)i * ld + k];
      for (int j = k + 1 + tx; j <= i; j += sx) A[(size_t)i * ld + j] -= lik * A[(size_t)j * ld + k];
    }
    if (tid == 0) dinv[k] = inv;  // (k, k) itself is not read again: the factor's diagonal is kept as its reciprocal
  }
  barrier();
  if (!ok) {
    for (int i = tid; i < nc * nc; i += nth) out[i] = 0.0;
    if (tid == 0) *fail = 1;
    return;
  }
  for (int j = tid; j < nc; j += nth) {  // column j of X = L^-1: X_jj = 1 / L_jj, X_ij = -(sum_{k=j}^{i-1} L_ik X_kj) / L_ii
    double* xr = A + (size_t)j * ld;     // row j of the array holds X_kj at column k > j
    for (int i = j + 1; i < nc; ++i) {
      const double* Li = A + (size_t)i * ld;
      double s = Li[j] * dinv[j];
      for (int k = j + 1; k < i; ++k) s += Li[k] * xr[k];
      xr[i] = -s * dinv[i];
    }
  }
  barrier();
  for (int i = ty; i < nc; i += sy)
    for (int j = i + tx; j < nc; j += sx) {
      const double* xi = A + (size_t)i * ld;
      const double* xj = A + (size_t)j * ld;
      double s = (i == j ? dinv[j] : xi[j]) * dinv[j];  // k = j
      for (int k = j + 1; k < nc; ++k) s += xi[k] * xj[k];
      out[(size_t)i * nc + j] = s;
      out[(size_t)j * nc + i] = s;
    }
  if (tid == 0) *fail = 0;
}

}  // namespace sgb
