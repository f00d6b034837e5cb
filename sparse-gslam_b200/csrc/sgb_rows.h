// sgb_rows.h -- the per-row / per-edge bodies of every kernel on the hot path, written once and called from the
// CUDA kernels (sgb_kernels.cu). The host-side test harness under tests/hostsim executes the same bodies
// serially to check indexing before GPU time is spent; the product never runs them on the CPU.
//
// Reference behaviour restated here: a9 EdgeSE2, a4-a6 EdgeSE2RhoTheta, a10
// constructQuadraticForm, a11 DCS, a15 buildSystem, a18 setLambda, a20 update. Summation order inside a
// Hessian block follows the reference: contributions are added in edge insertion order.
#pragma once
#include "sgb_math.h"
#include "sgb_types.h"

#if defined(__CUDA_ARCH__)
#define SGB_LDG(p) __ldg(p)    // read-only for the lifetime of the kernel
#define SGB_LDS(p) __ldcs(p)   // read-only AND touched once per pass (the Hessian block values streamed by the PCG):
                               // evict-first, so the stream does not push the gathered vectors out of L1/L2
#define SGB_LDCG(p) __ldcg(p)  // written by other CTAs (or, through NVLink, other GPUs) while the kernel runs: no L1
#else
#define SGB_LDG(p) (*(p))
#define SGB_LDS(p) (*(p))
#define SGB_LDCG(p) (*(p))
#endif

namespace sgb {
// two consecutive doubles at a 16-byte aligned address with one L1-bypassing 128-bit load (one L2 request, not two)
struct Pair64 { double a, b; };
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ Pair64 ldcg_pair(const double* p) {
  double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
}
#else
inline Pair64 ldcg_pair(const double* p) { return {p[0], p[1]}; }
#endif
}  // namespace sgb

namespace sgb {

struct LinAcc {
  double chi = 0.0;    // sum of e^T Omega e over the edges this row owns
  double chi_r = 0.0;  // robustified
  double maxd = 0.0;   // max |diag H|
};

SGB_HD size_t sell_vaddr(int e, int NC, int c) { return (size_t)(e & ~31) * NC + (size_t)c * 32 + (e & 31); }
SGB_HD int sell_width(const Sell& m, int slice) { return (m.sbase[slice + 1] - m.sbase[slice]) >> 5; }

// ---------------------------------------------------------------------------------------------------------
// one pose-pose edge evaluated at (xi, xj): error, Jacobians, weighted information, -Omega e
struct PPTerm {
  double e[3], A[9], B[9], om[6], omr[3], chi, chi_r;
};
SGB_HD void pp_term(const DevGraph& g, int k, const double* pose, PPTerm& t) {
  int i = SGB_LDG(&g.pp_i[k]), j = SGB_LDG(&g.pp_j[k]);
  double xi[3] = {pose[3 * i], pose[3 * i + 1], pose[3 * i + 2]};
  double xj[3] = {pose[3 * j], pose[3 * j + 1], pose[3 * j + 2]};
  double zinv[3];
  for (int c = 0; c < 3; ++c) zinv[c] = SGB_LDG(&g.pp_zinv[(size_t)c * g.n_pp + k]);
  for (int c = 0; c < 6; ++c) t.om[c] = SGB_LDG(&g.pp_info[(size_t)c * g.n_pp + k]);
  double si, ci, sz, cz;
  sgb_sincos(xi[2], &si, &ci);
  sgb_sincos(zinv[2], &sz, &cz);
  pp_error(xi, xj, zinv, ci, si, cz, sz, t.e);
  pp_jacobians(xi, xj, ci, si, cz, sz, t.A, t.B);
  t.chi = sym3_quad(t.om, t.e);
  t.chi_r = t.chi;
  double w = 1.0;
  if (g.has_robust) {
    double phi = SGB_LDG(&g.pp_phi[k]);
    if (phi > 0.0) dcs_robustify(phi, t.chi, &t.chi_r, &w);
  }
  if (w != 1.0)
    for (int c = 0; c < 6; ++c) t.om[c] *= w;  // robustInformation = rho[1] * Omega
  sym3_mul(t.om, t.e, t.omr);
  for (int c = 0; c < 3; ++c) t.omr[c] = -t.omr[c];
}
// error only (chi2 evaluation after a trial step)
SGB_HD void pp_chi(const DevGraph& g, int k, const double* pose, double* chi, double* chi_r) {
  int i = SGB_LDG(&g.pp_i[k]), j = SGB_LDG(&g.pp_j[k]);
  double xi[3] = {pose[3 * i], pose[3 * i + 1], pose[3 * i + 2]};
  double xj[3] = {pose[3 * j], pose[3 * j + 1], pose[3 * j + 2]};
  double zinv[3], om[6], e[3];
  for (int c = 0; c < 3; ++c) zinv[c] = SGB_LDG(&g.pp_zinv[(size_t)c * g.n_pp + k]);
  for (int c = 0; c < 6; ++c) om[c] = SGB_LDG(&g.pp_info[(size_t)c * g.n_pp + k]);
  double si, ci, sz, cz;
  sgb_sincos(xi[2], &si, &ci);
  sgb_sincos(zinv[2], &sz, &cz);
  pp_error(xi, xj, zinv, ci, si, cz, sz, e);
  double c2 = sym3_quad(om, e), cr = c2, w;
  if (g.has_robust) {
    double phi = SGB_LDG(&g.pp_phi[k]);
    if (phi > 0.0) dcs_robustify(phi, c2, &cr, &w);
  }
  *chi = c2;
  *chi_r = cr;
}

// M (3x3 row-major) = X^T * sym(om) * Y for 3x3 row-major X, Y
SGB_HD void xt_om_y(const double X[9], const double om[6], const double Y[9], double M[9]) {
  double XtO[9];
  const double O[9] = {om[0], om[1], om[2], om[1], om[3], om[4], om[2], om[4], om[5]};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) XtO[3 * r + c] = X[r] * O[c] + X[3 + r] * O[3 + c] + X[6 + r] * O[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M[3 * r + c] = XtO[3 * r] * Y[c] + XtO[3 * r + 1] * Y[3 + c] + XtO[3 * r + 2] * Y[6 + c];
}

struct PLTerm {
  double e[2], A[6], B[4], om[3], omr[2], chi;
};
SGB_HD void pl_term(const DevGraph& g, int k, const double* pose, const double* lm, bool want_A, bool want_B,
                    PLTerm& t) {
  int p = SGB_LDG(&g.pl_p[k]), l = SGB_LDG(&g.pl_l[k]);
  double x[3] = {pose[3 * p], pose[3 * p + 1], pose[3 * p + 2]};
  double ln[2] = {lm[2 * l], lm[2 * l + 1]};
  double z[2] = {SGB_LDG(&g.pl_z[k]), SGB_LDG(&g.pl_z[(size_t)g.n_pl + k])};
  for (int c = 0; c < 3; ++c) t.om[c] = SGB_LDG(&g.pl_info[(size_t)c * g.n_pl + k]);
  if (g.jac_numeric) {
    pl_error_literal(x, ln, z, t.e);
    pl_jac_numeric(x, ln, z, want_A, want_B, t.A, t.B);
  } else {
    double sgn, ca, sa;
    pl_error_closed(x, ln, z, t.e, &sgn, &ca, &sa);
    pl_jac_analytic(x, sgn, ca, sa, t.A, t.B);
  }
  t.chi = sym2_quad(t.om, t.e);
  sym2_mul(t.om, t.e, t.omr);
  t.omr[0] = -t.omr[0];
  t.omr[1] = -t.omr[1];
}
SGB_HD double pl_chi(const DevGraph& g, int k, const double* pose, const double* lm) {
  int p = SGB_LDG(&g.pl_p[k]), l = SGB_LDG(&g.pl_l[k]);
  double x[3] = {pose[3 * p], pose[3 * p + 1], pose[3 * p + 2]};
  double ln[2] = {lm[2 * l], lm[2 * l + 1]};
  double z[2] = {SGB_LDG(&g.pl_z[k]), SGB_LDG(&g.pl_z[(size_t)g.n_pl + k])};
  double om[3], e[2];
  for (int c = 0; c < 3; ++c) om[c] = SGB_LDG(&g.pl_info[(size_t)c * g.n_pl + k]);
  if (g.jac_numeric) {
    pl_error_literal(x, ln, z, e);
  } else {
    double sgn, ca, sa;
    pl_error_closed(x, ln, z, e, &sgn, &ca, &sa);
  }
  return sym2_quad(om, e);
}
// 3x2 block A^T Omega B of a pose-line edge (row-major 3x2)
SGB_HD void pl_offdiag(const PLTerm& t, double blk[6]) {
  double AtO[6];  // 3x2 = A^T (2x3)^T * Omega(2x2)
  for (int r = 0; r < 3; ++r) {
    AtO[2 * r] = t.A[r] * t.om[0] + t.A[3 + r] * t.om[1];
    AtO[2 * r + 1] = t.A[r] * t.om[1] + t.A[3 + r] * t.om[2];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 2; ++c) blk[2 * r + c] = AtO[2 * r] * t.B[c] + AtO[2 * r + 1] * t.B[2 + c];
}

// ---------------------------------------------------------------------------------------------------------
// current / trial estimate buffers of this rank
SGB_HD const double* cur_pose(const DevGraph& g) { return g.pose_buf[g.cur][g.rank]; }
SGB_HD const double* cur_lm(const DevGraph& g) { return g.lm_buf[g.cur][g.rank]; }

// sum of A^T Omega B over the duplicate chain headed by edge k, oriented as (row of vertex `from_h`, col other)
SGB_HD void pp_chain_block(const DevGraph& g, int k, const PPTerm& t, int row_h, const double* pose, double blk[9]) {
  // block as seen from vertex 0 of the head edge: A^T Omega B  (row hi, col hj)
  double m[9];
  xt_om_y(t.A, t.om, t.B, m);
  int hi = SGB_LDG(&g.pp_hi[k]);
  bool head_is_row = (hi == row_h);
  if (head_is_row) {
    for (int c = 0; c < 9; ++c) blk[c] = m[c];
  } else {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) blk[3 * r + c] = m[3 * c + r];
  }
  for (int d = SGB_LDG(&g.pp_dup[k]); d >= 0; d = SGB_LDG(&g.pp_dup[d])) {
    PPTerm u;
    pp_term(g, d, pose, u);
    double m2[9];
    xt_om_y(u.A, u.om, u.B, m2);
    if (SGB_LDG(&g.pp_hi[d]) == row_h) {
      for (int c = 0; c < 9; ++c) blk[c] += m2[c];
    } else {  // this duplicate runs the other way: its (row hj, col hi) view is the transpose
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) blk[3 * r + c] += m2[3 * c + r];
    }
  }
}

// linearise + assemble, owned pose row lp: diagonal block, gradient, and the off-diagonal blocks of THIS row
// one entry of a pose row's incidence list: the edge's contribution to the row's diagonal block H and gradient bv, the
// off-diagonal block(s) this row owns (written to Hpp / Hpl), the edge's chi2 if it is accounted on this row
SGB_HD void lin_pose_entry(const DevGraph& g, const double* pose, const double* lm, int it, double H[9], double bv[3], LinAcc& acc) {
  {
    int packed = SGB_LDG(&g.pinc[it]);
    int k = packed >> 2, role = (packed >> 1) & 1, type = packed & 1;
    if (type == 0) {
      PPTerm t;
      pp_term(g, k, pose, t);
      int hi = SGB_LDG(&g.pp_hi[k]), hj = SGB_LDG(&g.pp_hj[k]);
      if (role == 0) {
        double M[9];
        xt_om_y(t.A, t.om, t.A, M);
        for (int c = 0; c < 9; ++c) H[c] += M[c];
        for (int r = 0; r < 3; ++r) bv[r] += t.A[r] * t.omr[0] + t.A[3 + r] * t.omr[1] + t.A[6 + r] * t.omr[2];
        acc.chi += t.chi;
        acc.chi_r += t.chi_r;
        int e_ij = SGB_LDG(&g.pp_e_ij[k]);
        if (e_ij >= 0) {  // both vertices free, this edge leads its vertex pair: block (row hi, col hj)
          double blk[9];
          pp_chain_block(g, k, t, hi, pose, blk);
          for (int c = 0; c < 9; ++c) g.Hpp.vals[sell_vaddr(e_ij, 9, c)] = blk[c];
        }
      } else {
        double M[9];
        xt_om_y(t.B, t.om, t.B, M);
        for (int c = 0; c < 9; ++c) H[c] += M[c];
        for (int r = 0; r < 3; ++r) bv[r] += t.B[r] * t.omr[0] + t.B[3 + r] * t.omr[1] + t.B[6 + r] * t.omr[2];
        if (hi < 0) {  // vertex 0 is fixed: the edge's chi2 is owned here
          acc.chi += t.chi;
          acc.chi_r += t.chi_r;
        }
        int e_ji = SGB_LDG(&g.pp_e_ji[k]);
        if (e_ji >= 0) {  // block (row hj, col hi) = (A^T Omega B)^T summed over the chain
          double blk[9];
          pp_chain_block(g, k, t, hj, pose, blk);
          for (int c = 0; c < 9; ++c) g.Hpp.vals[sell_vaddr(e_ji, 9, c)] = blk[c];
        }
      }
    } else {
      PLTerm t;
      int hl = SGB_LDG(&g.pl_hl[k]);
      pl_term(g, k, pose, lm, true, hl >= 0, t);
      double AtO[6];
      for (int r = 0; r < 3; ++r) {
        AtO[2 * r] = t.A[r] * t.om[0] + t.A[3 + r] * t.om[1];
        AtO[2 * r + 1] = t.A[r] * t.om[1] + t.A[3 + r] * t.om[2];
      }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) H[3 * r + c] += AtO[2 * r] * t.A[c] + AtO[2 * r + 1] * t.A[3 + c];
      for (int r = 0; r < 3; ++r) bv[r] += t.A[r] * t.omr[0] + t.A[3 + r] * t.omr[1];
      acc.chi += t.chi;
      acc.chi_r += t.chi;
      int e_pl = SGB_LDG(&g.pl_e_pl[k]);
      if (e_pl >= 0) {
        double blk[6];
        pl_offdiag(t, blk);
        for (int d = SGB_LDG(&g.pl_dup[k]); d >= 0; d = SGB_LDG(&g.pl_dup[d])) {
          PLTerm u;
          pl_term(g, d, pose, lm, true, true, u);
          double m2[6];
          pl_offdiag(u, m2);
          for (int c = 0; c < 6; ++c) blk[c] += m2[c];
        }
        for (int c = 0; c < 6; ++c) g.Hpl.vals[sell_vaddr(e_pl, 6, c)] = blk[c];
      }
    }
  }
}
// linearise + assemble, owned pose row lp: diagonal block, gradient, and the off-diagonal blocks of THIS row
SGB_HD void lin_pose_row(const DevGraph& g, int lp, LinAcc& acc) {
  const double* pose = cur_pose(g);
  const double* lm = cur_lm(g);
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double bv[3] = {0, 0, 0};
  int beg = g.pinc_ptr[lp], end = g.pinc_ptr[lp + 1];
  for (int it = beg; it < end; ++it) lin_pose_entry(g, pose, lm, it, H, bv, acc);
  int ed = g.hpp_diag[lp];
  for (int c = 0; c < 9; ++c) g.Hpp.vals[sell_vaddr(ed, 9, c)] = H[c];
  for (int r = 0; r < 3; ++r) g.b_p[3 * (size_t)lp + r] = bv[r];
  acc.maxd = fmax(acc.maxd, fmax(fabs(H[0]), fmax(fabs(H[4]), fabs(H[8]))));
}
#if defined(__CUDACC__)
// The same row on FOUR lanes: lane `sub` takes the incidence entries sub, sub + 4, ... (each entry is an independent edge
// evaluation with its own chain of gathers), the four partial sums of the diagonal block and the gradient are then added
// in lane order ((p0 + p1) + p2) + p3 -- a fixed order that depends on the row's incidence list only, so the result is
// the same for every partition of the graph; it differs from the serial row above by rounding. All four lanes of a row's
// group must call (active == false for the lanes of a group beyond the last row).
__device__ __forceinline__ void lin_pose_row_lanes(const DevGraph& g, int lp, int sub, bool active, LinAcc& acc) {
  const double* pose = cur_pose(g);
  const double* lm = cur_lm(g);
  double v[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // H (9), bv (3)
  if (active) {
    const int beg = g.pinc_ptr[lp], end = g.pinc_ptr[lp + 1];
    for (int it = beg + sub; it < end; it += 4) lin_pose_entry(g, pose, lm, it, v, v + 9, acc);
  }
  const unsigned lane = threadIdx.x & 31u, base = lane & ~3u, mask = 0xfu << base;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const double p0 = __shfl_sync(mask, v[q], (int)base), p1 = __shfl_sync(mask, v[q], (int)base + 1);
    const double p2 = __shfl_sync(mask, v[q], (int)base + 2), p3 = __shfl_sync(mask, v[q], (int)base + 3);
    v[q] = ((p0 + p1) + p2) + p3;
  }
  if (active && sub == 0) {
    const int ed = g.hpp_diag[lp];
    for (int c = 0; c < 9; ++c) g.Hpp.vals[sell_vaddr(ed, 9, c)] = v[c];
    for (int r = 0; r < 3; ++r) g.b_p[3 * (size_t)lp + r] = v[9 + r];
    acc.maxd = fmax(acc.maxd, fmax(fabs(v[0]), fmax(fabs(v[4]), fabs(v[8]))));
  }
}
#endif

// linearise + assemble, owned landmark row ll: 2x2 diagonal block, gradient, and the landmark-major copy of the
// pose-line blocks (recomputed here so that every rank writes only rows it owns)
SGB_HD void lin_lm_row(const DevGraph& g, int ll, LinAcc& acc) {
  const double* pose = cur_pose(g);
  const double* lm = cur_lm(g);
  double h11 = 0, h12 = 0, h22 = 0, b0 = 0, b1 = 0;
  int beg = g.linc_ptr[ll], end = g.linc_ptr[ll + 1];
  for (int it = beg; it < end; ++it) {
    int k = SGB_LDG(&g.linc[it]);
    int e_lp = SGB_LDG(&g.pl_e_lp[k]);
    PLTerm t;
    pl_term(g, k, pose, lm, e_lp >= 0, true, t);
    // B^T Omega B
    double BtO[4];
    for (int r = 0; r < 2; ++r) {
      BtO[2 * r] = t.B[r] * t.om[0] + t.B[2 + r] * t.om[1];
      BtO[2 * r + 1] = t.B[r] * t.om[1] + t.B[2 + r] * t.om[2];
    }
    h11 += BtO[0] * t.B[0] + BtO[1] * t.B[2];
    h12 += BtO[0] * t.B[1] + BtO[1] * t.B[3];
    h22 += BtO[2] * t.B[1] + BtO[3] * t.B[3];
    b0 += t.B[0] * t.omr[0] + t.B[2] * t.omr[1];
    b1 += t.B[1] * t.omr[0] + t.B[3] * t.omr[1];
    if (SGB_LDG(&g.pl_hp[k]) < 0 && ll < g.nL_owned) {  // pose fixed: chi2 owned by the landmark row (of the owner)
      acc.chi += t.chi;
      acc.chi_r += t.chi;
    }
    if (e_lp >= 0) {
      double blk[6];
      pl_offdiag(t, blk);
      for (int d = SGB_LDG(&g.pl_dup[k]); d >= 0; d = SGB_LDG(&g.pl_dup[d])) {
        PLTerm u;
        pl_term(g, d, pose, lm, true, true, u);
        double m2[6];
        pl_offdiag(u, m2);
        for (int c = 0; c < 6; ++c) blk[c] += m2[c];
      }
      for (int c = 0; c < 6; ++c) g.Hlp.vals[sell_vaddr(e_lp, 6, c)] = blk[c];
    }
  }
  g.Hll[ll] = h11;
  g.Hll[(size_t)g.nL + ll] = h12;
  g.Hll[2 * (size_t)g.nL + ll] = h22;
  g.b_l[g.rank][2 * (size_t)ll] = b0;
  g.b_l[g.rank][2 * (size_t)ll + 1] = b1;
  acc.maxd = fmax(acc.maxd, fmax(fabs(h11), fabs(h22)));
}

// ---------------------------------------------------------------------------------------------------------
// trial set-up (BlockSolver::setLambda is applied on the fly; H itself is never modified => restoreDiagonal is free)
SGB_HD bool setup_lm_row(const DevGraph& g, int ll, double lambda) {
  double inv[3];
  bool ok = inv2_spd(g.Hll[ll] + lambda, g.Hll[(size_t)g.nL + ll], g.Hll[2 * (size_t)g.nL + ll] + lambda, inv);
  double* W = g.Hll_inv[g.rank];
  W[ll] = inv[0];
  W[(size_t)g.capL + ll] = inv[1];
  W[2 * (size_t)g.capL + ll] = inv[2];
  return ok;
}
// Schur diagonal block S_ii = Hpp_ii + lambda I - sum_l Hpl_il (Hll_l + lambda I)^-1 Hpl_il^T, its inverse (the
// block-Jacobi preconditioner) and the reduced right-hand side bt_i = b_i - sum_l Hpl_il (Hll_l+lambda I)^-1 b_l.
// Landmark quantities may live on another rank: gathered through the peer tables.
SGB_HD bool setup_pose_row(const DevGraph& g, int lp, double lambda) {
  double M[9];
  int ed = g.hpp_diag[lp];
  for (int c = 0; c < 9; ++c) M[c] = g.Hpp.vals[sell_vaddr(ed, 9, c)];
  M[0] += lambda;
  M[4] += lambda;
  M[8] += lambda;
  double bt[3] = {g.b_p[3 * (size_t)lp], g.b_p[3 * (size_t)lp + 1], g.b_p[3 * (size_t)lp + 2]};
  if (g.Hpl.rows > 0) {
    int slice = lp >> 5, lane = lp & 31;
    int w = sell_width(g.Hpl, slice);
    int base = g.Hpl.sbase[slice];
    for (int k = 0; k < w; ++k) {
      int e = base + k * 32 + lane;
      int enc = SGB_LDG(&g.Hpl.col[e]);
      if (enc < 0) continue;
      int o = enc >> kOwnerShift, l = enc & kLocalMask;
      double B[6];
      for (int c = 0; c < 6; ++c) B[c] = g.Hpl.vals[sell_vaddr(e, 6, c)];
      const double* W = g.Hll_inv[o];
      double w11 = SGB_LDCG(&W[l]), w12 = SGB_LDCG(&W[(size_t)g.capL + l]), w22 = SGB_LDCG(&W[2 * (size_t)g.capL + l]);
      double BW[6];
      for (int r = 0; r < 3; ++r) {
        BW[2 * r] = B[2 * r] * w11 + B[2 * r + 1] * w12;
        BW[2 * r + 1] = B[2 * r] * w12 + B[2 * r + 1] * w22;
      }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[3 * r + c] -= BW[2 * r] * B[2 * c] + BW[2 * r + 1] * B[2 * c + 1];
      const double* bl = g.b_l[o];
      double bl0 = SGB_LDCG(&bl[2 * (size_t)l]), bl1 = SGB_LDCG(&bl[2 * (size_t)l + 1]);
      for (int r = 0; r < 3; ++r) bt[r] -= BW[2 * r] * bl0 + BW[2 * r + 1] * bl1;
    }
  }
  double Mi[9];
  bool ok = inv3_spd(M, Mi);  // positive-definiteness check of the 3x3 diagonal block (the preconditioner is setup_chunk's)
  for (int r = 0; r < 3; ++r) g.bt[3 * (size_t)lp + r] = bt[r];
  return ok;
}

// Packed upper triangle of a symmetric N x N matrix: entry (i, j), i <= j. Every index below is a compile-time constant
// once the loops are unrolled, so the 78 doubles of a 12 x 12 block live in registers (the first build kept a full
// 144-double array that was indexed with run-time column positions: 1.2 KB of local memory per thread, 0.91 ms per launch
// on the 1M-pose graph, instruction- and local-memory-bound).
template <int N>
SGB_HD constexpr int tri_idx(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }

// In-place inverse of a symmetric positive definite N x N matrix held as its packed upper triangle, through the Cholesky
// factor A = U^T U; false when a pivot is not positive / not finite (the matrix is then left in an unspecified state).
template <int N>
SGB_HD bool inv_spd_packed(double* A) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {  // row j of U
    double dj = A[tri_idx<N>(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) dj -= A[tri_idx<N>(k, j)] * A[tri_idx<N>(k, j)];
    ok = ok && (dj > 0.0) && (dj < 1e300);
    const double inv = 1.0 / sqrt(dj);
    A[tri_idx<N>(j, j)] = inv;  // the diagonal holds 1 / U_jj from here on
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double v = A[tri_idx<N>(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= A[tri_idx<N>(k, j)] * A[tri_idx<N>(k, i)];
      A[tri_idx<N>(j, i)] = v * inv;
    }
  }
#pragma unroll
  for (int j = 1; j < N; ++j) {  // V = U^-1 (upper), column by column; entry (i, j) only needs U_kj with k >= i
#pragma unroll
    for (int i = 0; i < j; ++i) {
      double v = 0.0;
#pragma unroll
      for (int k = i; k < j; ++k) v -= (k == i ? A[tri_idx<N>(i, i)] : A[tri_idx<N>(i, k)]) * A[tri_idx<N>(k, j)];
      A[tri_idx<N>(i, j)] = v * A[tri_idx<N>(j, j)];
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)  // A^-1 = V V^T: (i, j), i <= j, = sum_{k >= j} V_ik V_jk; rows > i and columns > j are still V
#pragma unroll
    for (int j = i; j < N; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = j; k < N; ++k) v += A[tri_idx<N>(i, k)] * A[tri_idx<N>(j, k)];
      A[tri_idx<N>(i, j)] = v;
    }
  return ok;
}

// Block-Jacobi preconditioner with blocks of kChunk consecutive pose rows: the 12 x 12 diagonal block of the Schur
// complement S = (Hpp + lambda I) - Hpl (Hll + lambda I)^-1 Hpl^T -- pose-pose blocks inside the chunk (the odometry
// chain, mostly) plus the coupling of chunk-mates that observe the same landmark -- inverted exactly. Pose row lp
// stores its three rows of the inverse, see precond_mul. One thread per chunk (once per LM trial); the block is kept as
// its packed upper triangle in registers (tri_idx). A chunk never spans two 32-row slices (32 is a multiple of kChunk).
// Poses missing from the last chunk get identity.
SGB_HD bool setup_chunk(const DevGraph& g, int ch, double lambda) {
  constexpr int N = 3 * kChunk;
  const int p0 = ch * kChunk;
  const int np = (g.nP - p0 < kChunk) ? g.nP - p0 : kChunk;
  double D[(N * (N + 1)) / 2];
#pragma unroll
  for (int i = 0; i < (N * (N + 1)) / 2; ++i) D[i] = 0.0;
  const int slice = p0 >> 5;
  const int wpp = sell_width(g.Hpp, slice), bpp = g.Hpp.sbase[slice];
  const bool has_pl = g.Hpl.rows > 0;
  const int wpl = has_pl ? sell_width(g.Hpl, slice) : 0, bpl = has_pl ? g.Hpl.sbase[slice] : 0;
#pragma unroll
  for (int a = 0; a < kChunk; ++a) {
    if (a < np) {
      const int lane = (p0 + a) & 31;
      for (int k = 0; k < wpp; ++k) {
        const int e = bpp + k * 32 + lane;
        const int enc = SGB_LDG(&g.Hpp.col[e]);
        if (enc < 0) continue;
        if ((enc >> kOwnerShift) != g.rank) continue;
        const int b = (enc & kLocalMask) - p0;
        if (b < a || b >= np) continue;  // upper triangle of the chunk (the diagonal block included)
        double v[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) v[c] = g.Hpp.vals[sell_vaddr(e, 9, c)];
#pragma unroll
        for (int bb = a; bb < kChunk; ++bb)
          if (bb == b) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c)
                if (bb > a || c >= r) D[tri_idx<N>(3 * a + r, 3 * bb + c)] += v[3 * r + c];
          }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) D[tri_idx<N>(3 * a + r, 3 * a + r)] += lambda;
      for (int k = 0; k < wpl; ++k) {
        const int e = bpl + k * 32 + lane;
        const int enc = SGB_LDG(&g.Hpl.col[e]);
        if (enc < 0) continue;
        const int o = enc >> kOwnerShift, l = enc & kLocalMask;
        double B[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) B[c] = g.Hpl.vals[sell_vaddr(e, 6, c)];
        const double* W = g.Hll_inv[o];
        const double w11 = SGB_LDCG(&W[l]), w12 = SGB_LDCG(&W[(size_t)g.capL + l]), w22 = SGB_LDCG(&W[2 * (size_t)g.capL + l]);
        double BW[6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          BW[2 * r] = B[2 * r] * w11 + B[2 * r + 1] * w12;
          BW[2 * r + 1] = B[2 * r] * w12 + B[2 * r + 1] * w22;
        }
#pragma unroll
        for (int bb = a; bb < kChunk; ++bb) {  // chunk-mates (and the pose itself) that observe the same landmark
          if (bb >= np) continue;
          const int lane_b = (p0 + bb) & 31;
          for (int k2 = 0; k2 < wpl; ++k2) {
            const int e2 = bpl + k2 * 32 + lane_b;
            if (SGB_LDG(&g.Hpl.col[e2]) != enc) continue;
            double C[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) C[c] = bb == a ? B[c] : g.Hpl.vals[sell_vaddr(e2, 6, c)];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c)
                if (bb > a || c >= r) D[tri_idx<N>(3 * a + r, 3 * bb + c)] -= BW[2 * r] * C[2 * c] + BW[2 * r + 1] * C[2 * c + 1];
            break;
          }
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r) D[tri_idx<N>(3 * a + r, 3 * a + r)] = 1.0;
    }
  }
  const bool ok = inv_spd_packed<N>(D);
#pragma unroll
  for (int a = 0; a < kChunk; ++a)
    if (a < np) {
      // 36 entries per pose row = 9 float4, float4 q of pose lp at index q * nP + lp
#pragma unroll
      for (int q = 0; q < (3 * N) / 4; ++q) {
        float f[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = 4 * q + u, r = idx / N, m = idx % N, i = 3 * a + r;
          f[u] = (float)(m >= i ? D[tri_idx<N>(i, m)] : D[tri_idx<N>(m, i)]);
        }
        float* dst = g.Cinv + ((size_t)q * g.nP + p0 + a) * 4;
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);  // one 16-byte store
#else
        dst[0] = f[0]; dst[1] = f[1]; dst[2] = f[2]; dst[3] = f[3];
#endif
      }
    }
  return ok;
}

// ---------------------------------------------------------------------------------------------------------
// implicit Schur-complement operator  q = (Hpp + lambda I) v - Hpl (Hll + lambda I)^-1 Hpl^T v
//
// The SELL row loops below are software-pipelined by hand: the column indices of the NEXT pair of blocks are
// loaded before the current pair is consumed, and the (independent) value / gather loads of two blocks are issued
// together, so that one thread keeps ~2 x 12 loads in flight instead of serialising "index -> gather -> FMA" per
// block. Padding entries (col < 0) sit at the end of a row and contribute exact zeros; the order of the additions
// is the plain k = 0, 1, 2, ... order, so results do not depend on the unrolling.
SGB_HD int sell_col_or_pad(const int32_t* col, int e, bool valid) { return valid ? SGB_LDG(&col[e]) : -1; }

// Landmark-major pass over one slice of the grouped Hlp (sgb_types.h): this lane's share of
// u_l = sum_i Hpl_il^T v_i for row srow[slice] + lane / G. The G lanes of a row are then summed
// with lm_group_sum (device: shuffles; host harness: the same butterfly over an array).
// vtab[o] = pose-vector segment of rank o (p during PCG, x_p during back-substitution)
#ifndef SGB_LM_UNROLL
#define SGB_LM_UNROLL 2  // blocks whose loads one lane keeps in flight in the landmark pass
#endif
// what a warp needs to know about a slice of the grouped Hlp before it can issue the first useful load; the
// persistent kernel fetches it one task ahead so that it is off the chain of dependent latencies
struct LmSliceMeta {
  int e_begin, steps, shift, row0, row_end;
};
SGB_HD LmSliceMeta lm_slice_meta(const DevGraph& g, int slice) {
  LmSliceMeta m;
  m.e_begin = g.Hlp.sbase[slice];
  m.steps = (g.Hlp.sbase[slice + 1] - m.e_begin) >> 5;
  m.shift = g.Hlp.sshift[slice];
  m.row0 = g.Hlp.srow[slice];
  m.row_end = g.Hlp.srow[slice + 1];
  return m;
}
SGB_HD void lm_gather_lane(const DevGraph& g, const LmSliceMeta& m, int lane, double* const* vtab, double* u0_out,
                           double* u1_out) {
  constexpr int U = SGB_LM_UNROLL;
  const int steps = m.steps;
  const int base = m.e_begin + lane;
  const int32_t* col = g.Hlp.col;
  double u0 = 0, u1 = 0;
  int enc[U];
  for (int u = 0; u < U; ++u) enc[u] = sell_col_or_pad(col, base + 32 * u, u < steps);
  for (int j = 0; j < steps; j += U) {
    int nxt[U];
    for (int u = 0; u < U; ++u) nxt[u] = sell_col_or_pad(col, base + (j + U + u) * 32, j + U + u < steps);
    double v[U][3], a[U][6];
    for (int u = 0; u < U; ++u) {
      for (int c = 0; c < 3; ++c) v[u][c] = 0.0;
      for (int c = 0; c < 6; ++c) a[u][c] = 0.0;
      if (enc[u] >= 0) {
        const double* pv = vtab[enc[u] >> kOwnerShift] + 3 * (size_t)(enc[u] & kLocalMask);
        const double* pa = g.Hlp.vals + sell_vaddr(base + (j + u) * 32, 6, 0);
        for (int c = 0; c < 3; ++c) v[u][c] = SGB_LDCG(pv + c);
        for (int c = 0; c < 6; ++c) a[u][c] = SGB_LDS(pa + 32 * c);
      }
    }
    for (int u = 0; u < U; ++u) {
      u0 += a[u][0] * v[u][0] + a[u][2] * v[u][1] + a[u][4] * v[u][2];
      u1 += a[u][1] * v[u][0] + a[u][3] * v[u][1] + a[u][5] * v[u][2];
    }
    for (int u = 0; u < U; ++u) enc[u] = nxt[u];
  }
  *u0_out = u0;
  *u1_out = u1;
}
#if defined(__CUDACC__)
// butterfly over the G = 32 >> shift neighbouring lanes that share a row: every lane of the group ends with the sum
__device__ __forceinline__ void lm_group_sum(int shift, double& u0, double& u1) {
  for (int off = 1; off < (32 >> shift); off <<= 1) {
    u0 += __shfl_xor_sync(0xffffffffu, u0, off);
    u1 += __shfl_xor_sync(0xffffffffu, u1, off);
  }
}
#endif
#if !defined(__CUDA_ARCH__)
inline void lm_group_sum_host(int shift, double u0[32], double u1[32]) {
  for (int off = 1; off < (32 >> shift); off <<= 1) {
    double a[32], b[32];
    for (int l = 0; l < 32; ++l) { a[l] = u0[l] + u0[l ^ off]; b[l] = u1[l] + u1[l ^ off]; }
    for (int l = 0; l < 32; ++l) { u0[l] = a[l]; u1[l] = b[l]; }
  }
}
#endif
// Finish of a landmark row (= local landmark `ll`): mode 0, phase A of the Schur product, t_l = W_l u_l with
// W_l = (Hll_l + lambda I)^-1; mode 1, back-substitution, x_l = W_l (b_l - u_l). The operands that do not depend on
// the gather (W_l, b_l) are fetched by lm_row_prefetch BEFORE the gather loop so that their latency overlaps it.
struct LmRowOperands {
  double w11, w12, w22, b0, b1;
};
SGB_HD void lm_row_prefetch(const DevGraph& g, int ll, int mode, LmRowOperands& o) {
  const double* W = g.Hll_inv[g.rank];
  o.w11 = W[ll];
  o.w12 = W[(size_t)g.capL + ll];
  o.w22 = W[2 * (size_t)g.capL + ll];
  o.b0 = o.b1 = 0.0;
  if (mode == 1) {
    const double* bl = g.b_l[g.rank];
    o.b0 = bl[2 * (size_t)ll];
    o.b1 = bl[2 * (size_t)ll + 1];
  }
}
SGB_HD void lm_row_finish(const DevGraph& g, int ll, int mode, const LmRowOperands& o, double u0, double u1) {
  if (mode == 0) {
    double* t = g.t[g.rank];
    t[2 * (size_t)ll] = o.w11 * u0 + o.w12 * u1;
    t[2 * (size_t)ll + 1] = o.w12 * u0 + o.w22 * u1;
  } else {
    u0 = o.b0 - u0;
    u1 = o.b1 - u1;
    g.x_l[2 * (size_t)ll] = o.w11 * u0 + o.w12 * u1;
    g.x_l[2 * (size_t)ll + 1] = o.w12 * u0 + o.w22 * u1;
  }
}
// One slice of the landmark-major pass, executed by one warp (all 32 lanes must call): mode 0 = phase A on p,
// mode 1 = back-substitution on x_p.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void lm_slice_pass(const DevGraph& g, const LmSliceMeta& m, int mode) {
  const int lane = threadIdx.x & 31;
  const int sh = m.shift;
  const int row = m.row0 + (lane >> (5 - sh));
  const bool writer = (lane & ((32 >> sh) - 1)) == 0 && row < m.row_end;
  LmRowOperands o;
  if (writer) lm_row_prefetch(g, row, mode, o);
  double u0, u1;
  lm_gather_lane(g, m, lane, mode == 0 ? g.p : g.x_p, &u0, &u1);
  lm_group_sum(sh, u0, u1);
  if (writer) lm_row_finish(g, row, mode, o, u0, u1);
}
__device__ __forceinline__ void lm_slice_pass(const DevGraph& g, int slice, int mode) {
  lm_slice_pass(g, lm_slice_meta(g, slice), mode);
}
// the slices sl0, sl0 + stride, ... of one warp, with the next slice's descriptor fetched while the current one runs
__device__ __forceinline__ void lm_slices_pass(const DevGraph& g, int sl0, int stride, int mode) {
  const int n = g.Hlp.nslices;
  if (sl0 >= n) return;
  LmSliceMeta m = lm_slice_meta(g, sl0);
  for (int sl = sl0; sl < n; sl += stride) {
    LmSliceMeta nxt = m;
    if (sl + stride < n) nxt = lm_slice_meta(g, sl + stride);
    lm_slice_pass(g, m, mode);
    m = nxt;
  }
}
#else
inline void lm_slice_pass(const DevGraph& g, int slice, int mode) {
  double u0[32], u1[32];
  const LmSliceMeta m = lm_slice_meta(g, slice);
  for (int lane = 0; lane < 32; ++lane) lm_gather_lane(g, m, lane, mode == 0 ? g.p : g.x_p, &u0[lane], &u1[lane]);
  const int sh = g.Hlp.sshift[slice];
  lm_group_sum_host(sh, u0, u1);
  for (int rr = 0; rr < (1 << sh); ++rr) {
    int row = g.Hlp.srow[slice] + rr, lane = rr << (5 - sh);
    if (row >= g.Hlp.srow[slice + 1]) break;
    LmRowOperands o;
    lm_row_prefetch(g, row, mode, o);
    lm_row_finish(g, row, mode, o, u0[lane], u1[lane]);
  }
}
inline void lm_slices_pass(const DevGraph& g, int sl0, int stride, int mode) {
  for (int sl = sl0; sl < g.Hlp.nslices; sl += stride) lm_slice_pass(g, sl, mode);
}
#endif
// phase B (pose-major): w_i = lambda z_i + sum_j Hpp_ij z_j - sum_l Hpl_il t_l, followed by the two direction
// recurrences of the single-reduction PCG for this row, d_i = z_i + beta d_i and s_i = w_i + beta s_i (so that w
// itself never travels to memory); returns z_i . w_i. z = the peer-visible operator input g.p.
struct PoseRowMeta {  // first entry and width of a pose row in Hpp and Hpl (fetched one row ahead by the PCG kernel)
  int bpp, wpp, bpl, wpl;
};
SGB_HD PoseRowMeta pose_row_meta(const DevGraph& g, int lp) {
  const int slice = lp >> 5, lane = lp & 31;
  PoseRowMeta m;
  m.wpp = sell_width(g.Hpp, slice);
  m.bpp = g.Hpp.sbase[slice] + lane;
  const bool has_pl = g.Hpl.rows > 0;
  m.wpl = has_pl ? sell_width(g.Hpl, slice) : 0;
  m.bpl = has_pl ? g.Hpl.sbase[slice] + lane : 0;
  return m;
}
SGB_HD double schur_phaseB_row(const DevGraph& g, int lp, const PoseRowMeta& m, double lambda, double beta) {
  const double* vown = g.p[g.rank] + 3 * (size_t)lp;
  const int wpp = m.wpp, bpp = m.bpp, wpl = m.wpl, bpl = m.bpl;
  double* d = g.d + 3 * (size_t)lp;
  double* sv = g.s + 3 * (size_t)lp;
  const int32_t* cpp = g.Hpp.col;
  const int32_t* cpl = g.Hpl.col;
  int enc0 = sell_col_or_pad(cpp, bpp, wpp > 0), enc1 = sell_col_or_pad(cpp, bpp + 32, wpp > 1);
  int l0 = sell_col_or_pad(cpl, bpl, wpl > 0), l1 = sell_col_or_pad(cpl, bpl + 32, wpl > 1);  // issued early
  double vi0 = SGB_LDCG(vown), vi1 = SGB_LDCG(vown + 1), vi2 = SGB_LDCG(vown + 2);
  double q0 = lambda * vi0, q1 = lambda * vi1, q2 = lambda * vi2;
  for (int k = 0; k < wpp; k += 2) {
    const int e0 = bpp + k * 32, e1 = e0 + 32;
    const int n0 = sell_col_or_pad(cpp, e0 + 64, k + 2 < wpp), n1 = sell_col_or_pad(cpp, e1 + 64, k + 3 < wpp);
    double v[2][3] = {{0, 0, 0}, {0, 0, 0}};
    double a[2][9] = {{0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0}};
    if (enc0 >= 0) {
      const double* pv = g.p[enc0 >> kOwnerShift] + 3 * (size_t)(enc0 & kLocalMask);
      const double* pa = g.Hpp.vals + sell_vaddr(e0, 9, 0);
      for (int c = 0; c < 3; ++c) v[0][c] = SGB_LDCG(pv + c);
      for (int c = 0; c < 9; ++c) a[0][c] = SGB_LDS(pa + 32 * c);
    }
    if (enc1 >= 0) {
      const double* pv = g.p[enc1 >> kOwnerShift] + 3 * (size_t)(enc1 & kLocalMask);
      const double* pa = g.Hpp.vals + sell_vaddr(e1, 9, 0);
      for (int c = 0; c < 3; ++c) v[1][c] = SGB_LDCG(pv + c);
      for (int c = 0; c < 9; ++c) a[1][c] = SGB_LDS(pa + 32 * c);
    }
    for (int u = 0; u < 2; ++u) {
      q0 += a[u][0] * v[u][0] + a[u][1] * v[u][1] + a[u][2] * v[u][2];
      q1 += a[u][3] * v[u][0] + a[u][4] * v[u][1] + a[u][5] * v[u][2];
      q2 += a[u][6] * v[u][0] + a[u][7] * v[u][1] + a[u][8] * v[u][2];
    }
    enc0 = n0;
    enc1 = n1;
  }
  for (int k = 0; k < wpl; k += 2) {
    const int e0 = bpl + k * 32, e1 = e0 + 32;
    const int n0 = sell_col_or_pad(cpl, e0 + 64, k + 2 < wpl), n1 = sell_col_or_pad(cpl, e1 + 64, k + 3 < wpl);
    double tv[2][2] = {{0, 0}, {0, 0}}, a[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
    if (l0 >= 0) {
      const double* pt = g.t[l0 >> kOwnerShift] + 2 * (size_t)(l0 & kLocalMask);
      const double* pa = g.Hpl.vals + sell_vaddr(e0, 6, 0);
      Pair64 tt = ldcg_pair(pt);
      tv[0][0] = tt.a;
      tv[0][1] = tt.b;
      for (int c = 0; c < 6; ++c) a[0][c] = SGB_LDS(pa + 32 * c);
    }
    if (l1 >= 0) {
      const double* pt = g.t[l1 >> kOwnerShift] + 2 * (size_t)(l1 & kLocalMask);
      const double* pa = g.Hpl.vals + sell_vaddr(e1, 6, 0);
      Pair64 tt = ldcg_pair(pt);
      tv[1][0] = tt.a;
      tv[1][1] = tt.b;
      for (int c = 0; c < 6; ++c) a[1][c] = SGB_LDS(pa + 32 * c);
    }
    for (int u = 0; u < 2; ++u) {
      q0 -= a[u][0] * tv[u][0] + a[u][1] * tv[u][1];
      q1 -= a[u][2] * tv[u][0] + a[u][3] * tv[u][1];
      q2 -= a[u][4] * tv[u][0] + a[u][5] * tv[u][1];
    }
    l0 = n0;
    l1 = n1;
  }
  // (fetching these six operands, or the next row's descriptor, ahead of the row loops was measured: the extra live
  // registers cost more in spills at the 80-register budget than the shorter dependency chain gains)
  const double dold0 = d[0], dold1 = d[1], dold2 = d[2], sold0 = sv[0], sold1 = sv[1], sold2 = sv[2];
  const double d0 = vi0 + beta * dold0, d1 = vi1 + beta * dold1, d2 = vi2 + beta * dold2;
  const double s0 = q0 + beta * sold0, s1 = q1 + beta * sold1, s2 = q2 + beta * sold2;
  d[0] = d0;
  d[1] = d1;
  d[2] = d2;
  sv[0] = s0;
  sv[1] = s1;
  sv[2] = s2;
  return vi0 * q0 + vi1 * q1 + vi2 * q2;
}
SGB_HD double schur_phaseB_row(const DevGraph& g, int lp, double lambda, double beta) {
  return schur_phaseB_row(g, lp, pose_row_meta(g, lp), lambda, beta);
}
// The same row with U blocks in flight instead of two. On a graph that fits L2 a row is a chain of dependent L2
// latencies -- one per pair of blocks above -- and a thread owns a single row, so registers are not scarce: with U = 8 a
// whole row of Hpp (and of Hpl) is requested at once. Blocks are accumulated in the same order with the same expressions,
// so the result is bit-identical to the two-block version (padding adds +0.0).
template <int U>
SGB_HD double schur_phaseB_row_u(const DevGraph& g, int lp, const PoseRowMeta& m, double lambda, double beta) {
  const double* vown = g.p[g.rank] + 3 * (size_t)lp;
  const int wpp = m.wpp, bpp = m.bpp, wpl = m.wpl, bpl = m.bpl;
  double* d = g.d + 3 * (size_t)lp;
  double* sv = g.s + 3 * (size_t)lp;
  const int32_t* cpp = g.Hpp.col;
  const int32_t* cpl = g.Hpl.col;
  int enc[U], ln[U];
#pragma unroll
  for (int u = 0; u < U; ++u) enc[u] = sell_col_or_pad(cpp, bpp + 32 * u, wpp > u);
#pragma unroll
  for (int u = 0; u < U; ++u) ln[u] = sell_col_or_pad(cpl, bpl + 32 * u, wpl > u);
  // the six operands of the direction recurrences are requested up front as well
  const double dold0 = SGB_LDCG(d), dold1 = SGB_LDCG(d + 1), dold2 = SGB_LDCG(d + 2);
  const double sold0 = SGB_LDCG(sv), sold1 = SGB_LDCG(sv + 1), sold2 = SGB_LDCG(sv + 2);
  double vi0 = SGB_LDCG(vown), vi1 = SGB_LDCG(vown + 1), vi2 = SGB_LDCG(vown + 2);
  double q0 = lambda * vi0, q1 = lambda * vi1, q2 = lambda * vi2;
  for (int k = 0; k < wpp; k += U) {
    int nx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) nx[u] = sell_col_or_pad(cpp, bpp + 32 * (k + U + u), k + U + u < wpp);
    double v[U][3], a[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      for (int c = 0; c < 3; ++c) v[u][c] = 0.0;
      for (int c = 0; c < 9; ++c) a[u][c] = 0.0;
      if (enc[u] >= 0) {
        const double* pv = g.p[enc[u] >> kOwnerShift] + 3 * (size_t)(enc[u] & kLocalMask);
        const double* pa = g.Hpp.vals + sell_vaddr(bpp + 32 * (k + u), 9, 0);
        for (int c = 0; c < 3; ++c) v[u][c] = SGB_LDCG(pv + c);
        for (int c = 0; c < 9; ++c) a[u][c] = SGB_LDG(pa + 32 * c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      q0 += a[u][0] * v[u][0] + a[u][1] * v[u][1] + a[u][2] * v[u][2];
      q1 += a[u][3] * v[u][0] + a[u][4] * v[u][1] + a[u][5] * v[u][2];
      q2 += a[u][6] * v[u][0] + a[u][7] * v[u][1] + a[u][8] * v[u][2];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) enc[u] = nx[u];
  }
  for (int k = 0; k < wpl; k += U) {
    int nx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) nx[u] = sell_col_or_pad(cpl, bpl + 32 * (k + U + u), k + U + u < wpl);
    double tv[U][2], a[U][6];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      tv[u][0] = tv[u][1] = 0.0;
      for (int c = 0; c < 6; ++c) a[u][c] = 0.0;
      if (ln[u] >= 0) {
        const double* pt = g.t[ln[u] >> kOwnerShift] + 2 * (size_t)(ln[u] & kLocalMask);
        const double* pa = g.Hpl.vals + sell_vaddr(bpl + 32 * (k + u), 6, 0);
        Pair64 tt = ldcg_pair(pt);
        tv[u][0] = tt.a;
        tv[u][1] = tt.b;
        for (int c = 0; c < 6; ++c) a[u][c] = SGB_LDG(pa + 32 * c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      q0 -= a[u][0] * tv[u][0] + a[u][1] * tv[u][1];
      q1 -= a[u][2] * tv[u][0] + a[u][3] * tv[u][1];
      q2 -= a[u][4] * tv[u][0] + a[u][5] * tv[u][1];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) ln[u] = nx[u];
  }
  const double d0 = vi0 + beta * dold0, d1 = vi1 + beta * dold1, d2 = vi2 + beta * dold2;
  const double s0 = q0 + beta * sold0, s1 = q1 + beta * sold1, s2 = q2 + beta * sold2;
  d[0] = d0;
  d[1] = d1;
  d[2] = d2;
  sv[0] = s0;
  sv[1] = s1;
  sv[2] = s2;
  return vi0 * q0 + vi1 * q1 + vi2 * q2;
}
// the rows lp0, lp0 + stride, ... of one thread
SGB_HD double schur_phaseB_rows(const DevGraph& g, int lp0, int stride, double lambda, double beta) {
  double acc = 0.0;
  for (int lp = lp0; lp < g.nP; lp += stride) acc += schur_phaseB_row(g, lp, pose_row_meta(g, lp), lambda, beta);
  return acc;
}
template <int U>
SGB_HD double schur_phaseB_rows_u(const DevGraph& g, int lp0, int stride, double lambda, double beta) {
  if (U <= 2) return schur_phaseB_rows(g, lp0, stride, lambda, beta);
  double acc = 0.0;
  for (int lp = lp0; lp < g.nP; lp += stride) acc += schur_phaseB_row_u<(U > 2 ? U : 4)>(g, lp, pose_row_meta(g, lp), lambda, beta);
  return acc;
}
// z_i = sum_j Cinv_ij r_j over the kChunk poses j of row lp's chunk (rr = their residuals, 3 per pose, zeros for poses
// beyond nP); returns r_i . z_i
SGB_HD double precond_mul(const DevGraph& g, int lp, const double rr[3 * kChunk], double z[3]) {
  constexpr int N = 3 * kChunk;
  float ci[3 * N];  // nine 16-byte streaming loads, consecutive lanes = consecutive 16-byte words
#pragma unroll
  for (int q = 0; q < (3 * N) / 4; ++q) {
#if defined(__CUDA_ARCH__)
    const float4 v = __ldcs(reinterpret_cast<const float4*>(g.Cinv) + (size_t)q * g.nP + lp);
    ci[4 * q] = v.x; ci[4 * q + 1] = v.y; ci[4 * q + 2] = v.z; ci[4 * q + 3] = v.w;
#else
    const float* v = g.Cinv + ((size_t)q * g.nP + lp) * 4;
    ci[4 * q] = v[0]; ci[4 * q + 1] = v[1]; ci[4 * q + 2] = v[2]; ci[4 * q + 3] = v[3];
#endif
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c) acc += (double)ci[r * N + c] * rr[c];
    z[r] = acc;
  }
  const int a = lp & (kChunk - 1);
  return rr[3 * a] * z[0] + rr[3 * a + 1] * z[1] + rr[3 * a + 2] * z[2];
}
// host form (test harness): the chunk-mates' residuals are read from the residual vector rvec [3 * nP]
SGB_HD double precond_row_from(const DevGraph& g, int lp, const double* rvec, double z[3]) {
  double rr[3 * kChunk];
  const int p0 = lp & ~(kChunk - 1);
  for (int j = 0; j < kChunk; ++j)
    for (int c = 0; c < 3; ++c) rr[3 * j + c] = (p0 + j < g.nP) ? rvec[3 * (size_t)(p0 + j) + c] : 0.0;
  return precond_mul(g, lp, rr, z);
}
#if defined(__CUDACC__)
// device form: the kChunk lanes that hold the rows of one chunk (consecutive lanes: row index and lane index agree in
// their low bits) exchange their residuals with shuffles. EVERY lane of such a group must call it (active = false and
// r = 0 for a lane whose row is beyond nP).
__device__ __forceinline__ double precond_row_shfl(const DevGraph& g, int lp, bool active, const double r[3], double z[3]) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned base = lane & ~(unsigned)(kChunk - 1);
  const unsigned mask = ((1u << kChunk) - 1u) << base;
  double rr[3 * kChunk];
#pragma unroll
  for (int j = 0; j < kChunk; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) rr[3 * j + c] = __shfl_sync(mask, r[c], (int)(base + j));
  z[0] = z[1] = z[2] = 0.0;
  return active ? precond_mul(g, lp, rr, z) : 0.0;
}
#endif

// Pushed pose halos (sgb_partition.h): the owner of row lp writes the row's new operator input z and step x into the halo
// copies of every rank whose matrices reference the row. Plain strong stores through the peer mapping: they are ordered
// before the cross-rank barrier that ends the phase by the barrier's release chain (CTA barrier -> gpu-scope fence +
// ticket -> system-scope release of the publishing CTA), so the consumer's gathers after the barrier read them locally.
SGB_HD void push_halo_row(const DevGraph& g, int lp, const double z[3], const double x[3]) {
  const int b = SGB_LDG(&g.send_ptr[lp]), e = SGB_LDG(&g.send_ptr[lp + 1]);
  for (int k = b; k < e; ++k) {
    const int enc = SGB_LDG(&g.send_dst[k]);
    const int q = enc >> kOwnerShift;
    const size_t o = 3 * ((size_t)g.capP + (size_t)g.halo_base_at[q] + (size_t)(enc & kLocalMask));
    double* pz = g.p[q] + o;
    double* px = g.x_p[q] + o;
#if defined(__CUDA_ARCH__)
    for (int c = 0; c < 3; ++c) {
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(pz + c), "d"(z[c]) : "memory");
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(px + c), "d"(x[c]) : "memory");
    }
#else
    for (int c = 0; c < 3; ++c) { pz[c] = z[c]; px[c] = x[c]; }
#endif
  }
}

// SparseOptimizer::update for one owned free vertex: reads the current estimate, writes the new one into buffer
// `dst` of EVERY rank (estimates are replicated; the owner pushes its rows over NVLink); returns the vertex's share
// of computeScale = sum_j x_j (lambda x_j + b_j)
SGB_HD double update_pose_row(const DevGraph& g, int lp, double lambda, int dst) {
  int p = g.pose_of_l[lp];
  const double* x = g.x_p[g.rank] + 3 * (size_t)lp;
  const double* b = g.b_p + 3 * (size_t)lp;
  const double* src = cur_pose(g);
  double in[3] = {src[3 * (size_t)p], src[3 * (size_t)p + 1], src[3 * (size_t)p + 2]};
  double u[3] = {x[0], x[1], x[2]};
  double out[3];
  pose_oplus(in, u, out);
  for (int o = 0; o < g.world; ++o) {
    double* d = g.pose_buf[dst][o] + 3 * (size_t)p;
    d[0] = out[0];
    d[1] = out[1];
    d[2] = out[2];
  }
  return u[0] * (lambda * u[0] + b[0]) + u[1] * (lambda * u[1] + b[1]) + u[2] * (lambda * u[2] + b[2]);
}
SGB_HD double update_lm_row(const DevGraph& g, int ll, double lambda, int dst) {
  int l = g.lm_of_l[ll];
  const double* x = g.x_l + 2 * (size_t)ll;
  const double* b = g.b_l[g.rank] + 2 * (size_t)ll;
  const double* src = cur_lm(g);
  double in[2] = {src[2 * (size_t)l], src[2 * (size_t)l + 1]};
  double u[2] = {x[0], x[1]};
  double out[2];
  lm_oplus(in, u, out);
  for (int o = 0; o < g.world; ++o) {
    double* d = g.lm_buf[dst][o] + 2 * (size_t)l;
    d[0] = out[0];
    d[1] = out[1];
  }
  return u[0] * (lambda * u[0] + b[0]) + u[1] * (lambda * u[1] + b[1]);
}

}  // namespace sgb
