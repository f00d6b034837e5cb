// sgb_types.h -- device-resident layout of one rank's share of a graph (pointers into HBM), shared by the kernels,
// the host orchestration and the host-side test harness. See DESIGN.md "Data layout in HBM".
//
// Multi-GPU model: the reduced pose system is partitioned by contiguous blocks of free-pose rows; a landmark is
// owned by the rank that owns its first observer. Every rank stores the rows it owns (Hessian rows, gradient,
// PCG vectors). Column indices are ENCODED as (owner << kOwnerShift) | local index, and the few vectors that
// other ranks gather from (p, t, x_pose, b_landmark, Hll^-1) live in a peer-mapped arena, so a kernel reads a
// remote entry with an ordinary load through NVLink. world == 1 is the same code with one owner.
#pragma once
#include <stdint.h>

namespace sgb {

constexpr int kMaxRanks = 8;
constexpr int kChunk = 4;  // pose rows per block of the block-Jacobi preconditioner (12x12 scalar blocks)
constexpr int kOwnerShift = 26;  // up to 64M rows per rank
constexpr int kLocalMask = (1 << kOwnerShift) - 1;

// Sliced-ELL (SELL-32) block matrix: rows grouped in slices of 32, every slice padded to its widest row.
// Entry e = sbase[slice] + k*32 + lane addresses the k-th block of row (slice*32 + lane); its NC values live at
// vals[(e & ~31) * NC + c * 32 + (e & 31)], c = 0..NC-1, so that one warp reading component c of its k-th blocks
// touches 32 consecutive doubles (one 256-byte, fully coalesced request).
//
// The landmark-major matrix uses the GROUPED form: slice s holds 2^sshift[s] rows starting at srow[s] and gives each
// of them G = 32 >> sshift[s] lanes; lane (rr * G + kk) of the warp that owns the slice walks the blocks
// k = kk, kk + G, ... of row srow[s] + rr, and entry sbase[s] + j*32 + lane is its j-th block -- the same "32
// consecutive entries per warp step" addressing, with the row sum finished by a shuffle reduction over the G lanes.
// Neighbouring lanes hold neighbouring observers of one landmark (consecutive key-frames), so the gathers coalesce.
struct Sell {
  int32_t rows;          // number of (local) rows
  int32_t nslices;
  const int32_t* sbase;  // [nslices + 1] entry offset of each slice (multiple of 32)
  const int32_t* col;    // [entries] encoded column or -1 for padding
  double* vals;          // [entries * NC]
  const int32_t* srow;   // grouped form: [nslices + 1] first row of each slice; NULL = uniform 32-row slices
  const int32_t* sshift; // grouped form: [nslices] log2(rows in the slice)
};

// cross-rank signalling slots (one per rank, in the peer-mapped arena)
struct Mailbox {
  unsigned long long flag[2][kMaxRanks];  // [parity][source rank] = sequence number of the last message
  double val[2][kMaxRanks][4];
  // flag-in-data slots of the persistent PCG kernel's reductions: value k of source rank r travels as two 8-byte words
  // {low 32 bits of the double | seq << 32} and {high 32 bits | seq << 32}; an 8-byte store is single-copy atomic, so the
  // reader needs no separate flag and the writer no fence between value and flag (one NVLink round trip less)
  unsigned long long ll[2][kMaxRanks][2][2];
};

struct DevGraph {
  // ---- partition
  int32_t world, rank;
  int32_t nP, nL;        // rows kept by this rank: free poses / free landmarks (landmarks: owned rows first, then ghosts)
  int32_t nL_owned;      // landmark rows [0, nL_owned) are owned (updated, exported, chi2-accounted here); [nL_owned, nL)
                         // are ghost copies of rows other ranks own (sgb_partition.h); == nL unless ghosts are enabled
  int32_t ghosts;        // 1: ghost landmarks are in use, the pose-major pass reads no landmark quantity of another rank
  int32_t capP, capL;    // max rows owned by any rank (array strides in the arena)
  int32_t P_all, L_all;  // all vertices (array order of sgb_set_graph); estimates are replicated on every rank
  int32_t n_pp, n_pl;    // local edges: incident to an owned row
  int32_t n_pp_owned, n_pl_owned;  // the first n_*_owned local edges have their chi2 accounted on this rank
  int32_t has_robust;    // any DCS edge
  int32_t jac_numeric;   // 1 = g2o central differences for pose-line edges
  int32_t cur;           // which estimate buffer is current (0/1); the other one holds the LM trial
  int32_t pushed;        // 1: pose halos are pushed (sgb_partition.h): p / x_p of this rank hold, behind the capP own rows,
                         // copies of the remote rows its matrices reference, written by their owners; no remote gathers
  // ---- pushed halos: destinations of the owned rows other ranks need (CSR over the owned rows)
  const int32_t* send_ptr;            // [nP + 1]
  const int32_t* send_dst;            // (destination rank << kOwnerShift) | k: this row is the k-th row of this rank in
                                      // the destination's halo
  int32_t halo_base_at[kMaxRanks];    // first halo slot of this rank's rows on every destination (exchanged at connect)
  // ---- estimates, replicated: est[buffer][rank] -> [3*P_all] / [2*L_all]; est[b][rank] is this rank's copy
  double* pose_buf[2][kMaxRanks];
  double* lm_buf[2][kMaxRanks];
  // ---- vertex maps
  const int32_t* pose_of_l;  // [nP] local free pose -> pose array index
  const int32_t* lm_of_l;    // [nL]
  // ---- local pose-pose edges (owned first, insertion order inside each group), component-major SoA
  const int32_t* pp_i;       // [n_pp] pose array index of vertex 0
  const int32_t* pp_j;
  const int32_t* pp_hi;      // [n_pp] GLOBAL free index of vertex 0 or -1 (fixed)
  const int32_t* pp_hj;
  const double* pp_zinv;     // [3][n_pp] inverse measurement (x, y, theta)
  const double* pp_info;     // [6][n_pp] upper triangle
  const double* pp_phi;      // [n_pp] DCS delta (<= 0: none); only read when has_robust
  const int32_t* pp_e_ij;    // [n_pp] local Hpp entry of block (row hi, col hj); -1 unless row hi is owned, both
                             //        vertices are free and the edge is the first (leader) of its vertex pair
  const int32_t* pp_e_ji;    // [n_pp] local Hpp entry of block (row hj, col hi); -1 unless row hj is owned, ...
  const int32_t* pp_dup;     // [n_pp] next local edge on the same vertex pair (chain), -1 = none
  // ---- local pose-line edges
  const int32_t* pl_p;       // [n_pl] pose array index
  const int32_t* pl_l;       // [n_pl] landmark array index
  const int32_t* pl_hp;      // GLOBAL free index or -1
  const int32_t* pl_hl;
  const double* pl_z;        // [2][n_pl]
  const double* pl_info;     // [3][n_pl]
  const int32_t* pl_e_pl;    // [n_pl] local Hpl entry (pose row owned, leader), else -1
  const int32_t* pl_e_lp;    // [n_pl] local Hlp entry (landmark row owned, leader), else -1
  const int32_t* pl_dup;     // [n_pl] duplicate chain
  // ---- incidence lists of the owned rows (insertion order within a vertex)
  const int32_t* pinc_ptr;   // [nP + 1]
  const int32_t* pinc;       // packed: (local edge << 2) | (role << 1) | type ; type 0 = pose-pose, 1 = pose-line
  const int32_t* linc_ptr;   // [nL + 1]
  const int32_t* linc;       // local pose-line edge index
  // ---- Hessian rows owned by this rank (un-reduced, lambda NOT included) and gradient
  Sell Hpp;                  // nP x Pf, 3x3 blocks row-major (NC = 9), both triangles; columns = encoded poses
  Sell Hpl;                  // nP x Lf, 3x2 blocks row-major (NC = 6); columns = encoded landmarks
  Sell Hlp;                  // nL x Pf, the same 3x2 blocks (NC = 6), landmark-major; row r = local landmark r
  const int32_t* hpp_diag;   // [nP] Hpp entry of the diagonal block
  const int32_t* hlp_edge;   // [Hlp entries] local pose-line edge that leads the vertex pair stored at this entry, -1 = padding
                             //               (the landmark rows are linearised by one lane per entry, k_lin_lm)
  const int32_t* lfix_ptr;   // [nL + 1] CSR over the landmark rows: their observations from FIXED poses (no Hessian block;
  const int32_t* lfix;       //          usually only the first key-frame's), local edge indices in insertion order
  double* Hll;               // [3][nL] (11,12,22), stride nL
  double* b_p;               // [3*nP]
  // ---- vectors other ranks gather from: tbl[rank] is this rank's own array
  double* b_l[kMaxRanks];     // [2*capL] gradient of the owned landmarks
  double* Hll_inv[kMaxRanks]; // [3][capL] (Hll + lambda I)^-1, stride capL
  double* x_p[kMaxRanks];     // [3*capP] pose step
  double* p[kMaxRanks];       // [3*capP] the vector the Schur operator is applied to: the preconditioned residual z
  double* t[kMaxRanks];       // [2*capL] (Hll + lambda I)^-1 Hpl^T p
  Mailbox* mbox[kMaxRanks];
  // ---- local only
  double* x_l;               // [2*nL] landmark step
  float* Cinv;               // [9][nP] float4: block-Jacobi preconditioner with blocks of kChunk = 4 consecutive pose rows: pose
                             //          row lp holds its 3 rows (x 12 columns) of the inverse 12x12 Schur diagonal block.
                             //          Stored in SINGLE precision (the products are formed in double): a preconditioner
                             //          only has to be symmetric positive definite -- (i, j) and (j, i) are rounded from
                             //          the same double -- and it does not enter the solution the PCG converges to
  double* bt;                // [3*nP] reduced right-hand side
  double* r;                 // residual
  double* d;                 // search direction
  double* s;                 // S d
  // ---- coarse space of the two-level preconditioner (sgb_coarse.h; single GPU, the resident four-lane solve only).
  // cz_h = 0: off. Node n sits at pose row n * cz_h; gather lists built by the host (plan_coarse, sgb_partition.h):
  int32_t cz_h, cz_nn;       // node spacing in pose rows (a multiple of 8), number of nodes (coarse dimension 3 * cz_nn)
  int32_t cz_ng;             // (node, landmark) pairs with a non-zero G = R Hpl block
  const int32_t* cz_g_ptr;   // [cz_ng + 1] CSR over the pairs: Hpl entries and hat weights that sum to the 3x2 block
  const int32_t* cz_g_e;
  const double* cz_g_w;
  const int32_t* cz_g_lm;    // [cz_ng] local landmark row of the pair
  const int32_t* cz_p_ptr;   // [cz_nn^2 + 1] CSR over node pairs (m * cz_nn + n, m <= n): Hpp entries and weight products
  const int32_t* cz_p_e;
  const double* cz_p_w;
  const double* cz_rr;       // [cz_nn^2] (R R^T)(m, n): what lambda multiplies
  const int32_t* cz_t_ptr;   // [cz_nn^2 + 1] CSR over node pairs: Schur terms (g1, g2) = two G pairs of one landmark
  const int32_t* cz_t_g;     // [2 * terms]
  double* cz_G;              // [6 * cz_ng] G values of the current linearisation
  double* cz_A;              // [(3 cz_nn)^2] (R S R^T)^-1 of the current trial, row-major
  int32_t* cz_fail;          // 1: the last factorisation met a non-positive pivot (coarse term switched off for that solve)
};

// scalars of the optimiser kept on the device (LM / GN control, reductions); identical on every rank
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
  unsigned long long xseq;  // sequence number of the last cross-rank message this rank took part in
  int32_t pcg_iters;
  int32_t pcg_flag;     // 0 converged, 1 max iterations, 2 breakdown (not SPD / non-finite)
  int32_t setup_fail;   // non-invertible diagonal block seen in the trial set-up (this rank)
  int32_t accepted;     // LM: last trial accepted
  int32_t trials;
  int32_t result;       // SGB_RESULT_*
  int32_t again;        // LM: run another trial
  int32_t pad;
  unsigned long long pcg_phase_ns[4];  // k_pcg: accumulated wall time of phases A-D seen by CTA 0 (reset per optimize)
};

}  // namespace sgb
