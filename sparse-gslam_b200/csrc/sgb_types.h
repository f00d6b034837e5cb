// sgb_types.h -- device-resident layout of one graph (pointers into HBM) shared by the kernels, the host
// orchestration and the host-side test harness. See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

namespace sgb {

// Sliced-ELL (SELL-32) block matrix: rows grouped in slices of 32, every slice padded to its widest row.
// Entry e = sbase[slice] + k*32 + lane addresses the k-th block of row (slice*32 + lane); its NC values live at
// vals[(e & ~31) * NC + c * 32 + (e & 31)], c = 0..NC-1, so that one warp reading component c of its k-th blocks
// touches 32 consecutive doubles (one 256-byte, fully coalesced request).
struct Sell {
  int32_t rows;          // number of rows
  int32_t nslices;
  const int32_t* sbase;  // [nslices + 1] entry offset of each slice (multiple of 32)
  const int32_t* col;    // [entries] column (block index) or -1 for padding
  double* vals;          // [entries * NC]
};

constexpr int SELL_C = 32;

struct DevGraph {
  // ---- sizes
  int32_t P_all, L_all;  // all vertices (array order of sgb_set_graph)
  int32_t Pf, Lf;        // free (Hessian-indexed) poses / landmarks
  int32_t n_pp, n_pl;    // active edges
  int32_t has_robust;    // any DCS edge
  int32_t jac_numeric;   // 1 = g2o central differences for pose-line edges
  // ---- estimates: [3*P_all], [2*L_all]
  double* pose;
  double* lm;
  // ---- vertex maps
  const int32_t* pose_of_h;  // [Pf] free pose -> pose array index
  const int32_t* lm_of_h;    // [Lf]
  // ---- active pose-pose edges (insertion order), SoA
  const int32_t* pp_i;       // [n_pp] pose array index of vertex 0
  const int32_t* pp_j;
  const int32_t* pp_hi;      // [n_pp] free index of vertex 0 or -1 (fixed)
  const int32_t* pp_hj;
  const double* pp_zinv;     // [3][n_pp] inverse measurement (x, y, theta), component-major
  const double* pp_info;     // [6][n_pp] upper triangle, component-major
  const double* pp_phi;      // [n_pp] DCS delta (<= 0: none); only read when has_robust
  const int32_t* pp_e_ij;    // [n_pp] SELL entry of block (row hi, col hj) in Hpp; -1 if a vertex is fixed or the
                             //        edge is not the first (leader) of its vertex pair
  const int32_t* pp_e_ji;    // [n_pp] SELL entry of block (row hj, col hi)
  const int32_t* pp_dup;     // [n_pp] next edge on the same vertex pair (chain), -1 = none
  // ---- active pose-line edges
  const int32_t* pl_p;       // [n_pl] pose array index
  const int32_t* pl_l;       // [n_pl] landmark array index
  const int32_t* pl_hp;      // free index or -1
  const int32_t* pl_hl;
  const double* pl_z;        // [2][n_pl]
  const double* pl_info;     // [3][n_pl]
  const int32_t* pl_e_pl;    // [n_pl] SELL entry in Hpl (pose-major), -1 if not leader / a vertex fixed
  const int32_t* pl_e_lp;    // [n_pl] SELL entry in Hlp (landmark-major)
  const int32_t* pl_dup;     // [n_pl] duplicate chain
  // ---- incidence lists (insertion order within a vertex)
  const int32_t* pinc_ptr;   // [Pf + 1]
  const int32_t* pinc;       // packed: (edge << 2) | (role << 1) | type ; type 0 = pose-pose, 1 = pose-line
  const int32_t* linc_ptr;   // [Lf + 1]
  const int32_t* linc;       // pose-line edge index
  // ---- Hessian (un-reduced, lambda NOT included) and gradient
  Sell Hpp;                  // Pf x Pf, 3x3 blocks row-major (NC = 9), both triangles
  Sell Hpl;                  // Pf x Lf, 3x2 blocks row-major (NC = 6), pose-major
  Sell Hlp;                  // Lf x Pf, the same 3x2 blocks (NC = 6), landmark-major; row r = landmark lp_row2h[r]
  const int32_t* hpp_diag;   // [Pf] SELL entry of the diagonal block
  const int32_t* lp_row2h;   // [Lf] Hlp row -> free landmark
  const int32_t* lp_h2row;   // [Lf]
  double* Hll;               // [3][Lf] (11,12,22)
  double* b;                 // [3*Pf + 2*Lf] Hessian order
  // ---- per-trial quantities
  double* Hll_inv;           // [3][Lf] (Hll + lambda I)^-1
  double* Minv;              // [9][Pf] block-Jacobi preconditioner = inverse of the Schur diagonal block
  double* bt;                // [3*Pf] reduced right-hand side
  double* x;                 // [3*Pf + 2*Lf] step
  // ---- PCG vectors [3*Pf], t [2*Lf]
  double* r;
  double* z;
  double* p;
  double* q;
  double* t;
  // ---- trial estimates
  double* pose_trial;
  double* lm_trial;
};

// scalars of the optimiser kept on the device (LM / GN control, reductions)
struct DevScalars {
  double chi2;          // activeChi2 of the last evaluation
  double chi2_robust;   // activeRobustChi2
  double chi_lin;       // activeRobustChi2 at the last linearisation point
  double max_diag;      // max |H_vv(d,d)|
  double lambda;
  double ni;
  double current_chi;   // LM currentChi
  double temp_chi;
  double rho;
  double scale;         // computeScale
  double rz0, rz, pq;   // PCG
  double pcg_rel;       // sqrt(rz / rz0) at exit
  int32_t pcg_iters;
  int32_t pcg_flag;     // 0 converged, 1 max iterations, 2 breakdown (not SPD / non-finite)
  int32_t setup_fail;   // non-invertible diagonal block seen in the trial set-up
  int32_t accepted;     // LM: last trial accepted
  int32_t trials;
  int32_t result;       // SGB_RESULT_*
  int32_t again;        // LM: run another trial
  int32_t pad;
};

}  // namespace sgb
