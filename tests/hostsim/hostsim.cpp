// hostsim.cpp -- TEST HARNESS ONLY (never shipped, never loaded by the sparse_gslam_b200 package).
//
// Executes the row bodies of the CUDA kernels (sparse-gslam_b200/csrc/sgb_rows.h) serially on the host, over the
// structure produced by the real host-side symbolic phase and partition planner (sgb_structure.cpp,
// sgb_partition.cpp), so that the scatter maps, SELL addressing, owner/local column encoding and per-row arithmetic
// can be checked against the oracle in the CPU-only test tier before GPU time is spent. `world` virtual ranks live
// in one process: their peer tables point at each other's arrays exactly as the NVLink peer mappings do on the
// GPUs, and every phase is run rank after rank (bulk-synchronous), mirroring sgb_backend.cu / k_pcg step by step.
// The product library libsgb.so does not contain this file and has no CPU path.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/sgb_capi.h"
#include "../../sparse-gslam_b200/csrc/sgb_partition.h"
#include "../../sparse-gslam_b200/csrc/sgb_coarse.h"
#include "../../sparse-gslam_b200/csrc/sgb_rows.h"
#include "../../sparse-gslam_b200/csrc/sgb_structure.h"

using namespace sgb;

struct Rank {
  CoarsePlan CZ;
  std::vector<double> cz_work, cz_dinv;  // the factorisation's scratch (shared memory on the device)
  int cz_fail = 0;
  Structure Sr;  // this rank's own (filtered) structure when the rank-filtered symbolic phase is simulated
  LocalPlan P;
  DevGraph G;
  std::vector<std::vector<double>> dbl;
  std::vector<std::vector<float>> flt;
  std::vector<std::vector<int32_t>> ints;
  double* D(size_t n) { dbl.emplace_back(std::max<size_t>(n, 1), 0.0); return dbl.back().data(); }
  const int32_t* I(const std::vector<int32_t>& v) { ints.push_back(v); if (ints.back().empty()) ints.back().push_back(0); return ints.back().data(); }
};

struct hs_handle {
  Structure S;
  int world = 1;
  std::vector<std::unique_ptr<Rank>> R;
  std::string err;
  double lambda = 0, ni = 2;
  double tol = 1e-10;
  int maxit = 0;
  int cur = 0;
};

static void mk_sell(Rank* r, Sell* out, const HostSell& s, int NC) {
  out->rows = s.rows;
  out->nslices = s.nslices;
  out->sbase = r->I(s.sbase);
  out->col = r->I(s.col);
  out->vals = r->D((size_t)s.entries() * NC);
  out->srow = s.grouped() ? r->I(s.srow) : nullptr;
  out->sshift = s.grouped() ? r->I(s.sshift) : nullptr;
}

extern "C" {

void hs_use_ghost_landmarks(int on) { partition_use_ghost_landmarks(on != 0); }
static int g_coarse_nodes = 0;
// two-level preconditioner (sgb_coarse.h): at most `max_nodes` coarse nodes, 0 = block-Jacobi only (world == 1)
void hs_use_coarse(int max_nodes) { g_coarse_nodes = max_nodes; }
static bool g_filtered = false;
// every virtual rank builds its own rank-filtered structure (build_structure(..., world, rank)), as libsgb does on >1 GPUs
void hs_use_filtered_structure(int on) { g_filtered = on != 0; }

hs_handle* hs_create(const sgb_graph_soa* g, int jac_numeric, double tol, int maxit, int world, int* status) {
  hs_handle* h = new hs_handle();
  h->world = std::max(1, world);
  sgb_status st = build_structure(*g, h->S, h->err);
  if (status) *status = st;
  if (st != SGB_OK) return h;
  build_block_list(h->S);
  h->tol = tol > 0 ? tol : 1e-10;
  h->maxit = maxit > 0 ? maxit : std::max(100, 12 * h->S.Pf);
  size_t np = 3 * (size_t)h->S.P_all, nl = 2 * (size_t)h->S.L_all;
  const bool filtered = g_filtered && h->world > 1 && partition_ghost_landmarks();
  for (int rk = 0; rk < h->world; ++rk) {
    h->R.emplace_back(new Rank());
    Rank* r = h->R.back().get();
    if (filtered) {
      st = build_structure(*g, r->Sr, h->err, h->world, rk);
      if (st != SGB_OK) { if (status) *status = st; return h; }
    }
    const Structure& S = filtered ? r->Sr : h->S;
    st = partition(S, h->world, rk, r->P, h->err);
    if (st != SGB_OK) { if (status) *status = st; return h; }
    build_export(h->S, r->P);
    const LocalPlan& P = r->P;
    DevGraph& G = r->G;
    std::memset(&G, 0, sizeof G);
    G.world = h->world; G.rank = rk; G.nP = P.nP; G.nL = P.nL; G.capP = P.capP; G.capL = P.capL;
    G.nL_owned = P.nL_owned; G.ghosts = (h->world > 1 && partition_ghost_landmarks()) ? 1 : 0;
    G.P_all = S.P_all; G.L_all = S.L_all; G.n_pp = P.n_pp; G.n_pl = P.n_pl;
    G.n_pp_owned = P.n_pp_owned; G.n_pl_owned = P.n_pl_owned;
    G.has_robust = S.has_robust; G.jac_numeric = jac_numeric; G.cur = 0;
    G.pose_of_l = r->I(P.pose_of_l); G.lm_of_l = r->I(P.lm_of_l);
    G.pp_i = r->I(P.pp_i); G.pp_j = r->I(P.pp_j); G.pp_hi = r->I(P.pp_hi); G.pp_hj = r->I(P.pp_hj);
    G.pp_e_ij = r->I(P.pp_e_ij); G.pp_e_ji = r->I(P.pp_e_ji); G.pp_dup = r->I(P.pp_dup);
    G.pl_p = r->I(P.pl_p); G.pl_l = r->I(P.pl_l); G.pl_hp = r->I(P.pl_hp); G.pl_hl = r->I(P.pl_hl);
    G.pl_e_pl = r->I(P.pl_e_pl); G.pl_e_lp = r->I(P.pl_e_lp); G.pl_dup = r->I(P.pl_dup);
    G.pinc_ptr = r->I(P.pinc_ptr); G.pinc = r->I(P.pinc); G.linc_ptr = r->I(P.linc_ptr); G.linc = r->I(P.linc);
    G.hpp_diag = r->I(P.hpp_diag);
    double* zinv = r->D(3 * (size_t)P.n_pp); double* info = r->D(6 * (size_t)P.n_pp); double* phi = r->D(P.n_pp);
    for (int k = 0; k < P.n_pp; ++k) {
      int s = S.pp_src[P.pp_g[k]];
      double x = g->pp_z[3 * (size_t)s], y = g->pp_z[3 * (size_t)s + 1], th = g->pp_z[3 * (size_t)s + 2];
      double thi = normalize_theta(-th), c = std::cos(thi), sn = std::sin(thi);
      zinv[k] = c * (-x) - sn * (-y);
      zinv[(size_t)P.n_pp + k] = sn * (-x) + c * (-y);
      zinv[2 * (size_t)P.n_pp + k] = thi;
      for (int c6 = 0; c6 < 6; ++c6) info[(size_t)c6 * P.n_pp + k] = g->pp_info[6 * (size_t)s + c6];
      phi[k] = g->pp_phi ? g->pp_phi[s] : 0.0;
    }
    G.pp_zinv = zinv; G.pp_info = info; G.pp_phi = phi;
    double* z = r->D(2 * (size_t)P.n_pl); double* linfo = r->D(3 * (size_t)P.n_pl);
    for (int k = 0; k < P.n_pl; ++k) {
      int s = S.pl_src[P.pl_g[k]];
      z[k] = g->pl_z[2 * (size_t)s]; z[(size_t)P.n_pl + k] = g->pl_z[2 * (size_t)s + 1];
      for (int c3 = 0; c3 < 3; ++c3) linfo[(size_t)c3 * P.n_pl + k] = g->pl_info[3 * (size_t)s + c3];
    }
    G.pl_z = z; G.pl_info = linfo;
    mk_sell(r, &G.Hpp, P.Hpp, 9); mk_sell(r, &G.Hpl, P.Hpl, 6); mk_sell(r, &G.Hlp, P.Hlp, 6);
    size_t n3 = 3 * (size_t)P.nP, n2 = 2 * (size_t)P.nL;
    G.Hll = r->D(3 * (size_t)P.nL); G.b_p = r->D(n3); G.x_l = r->D(n2);
    r->flt.emplace_back(3 * 3 * kChunk * (size_t)std::max(P.nP, 1), 0.0f); G.Cinv = r->flt.back().data(); G.bt = r->D(n3); G.r = r->D(n3); G.d = r->D(n3); G.s = r->D(n3);
    // the "arena": arrays other ranks reach into
    for (int b = 0; b < 2; ++b) {
      G.pose_buf[b][rk] = r->D(np); G.lm_buf[b][rk] = r->D(nl);
      std::copy(g->pose_est, g->pose_est + np, G.pose_buf[b][rk]);
      std::copy(g->lm_est, g->lm_est + nl, G.lm_buf[b][rk]);
    }
    G.p[rk] = r->D(3 * ((size_t)P.capP + P.nH)); G.x_p[rk] = r->D(3 * ((size_t)P.capP + P.nH));  // own rows + halo copies
    G.pushed = P.pushed ? 1 : 0;
    if (P.pushed) { G.send_ptr = r->I(P.send_ptr); G.send_dst = r->I(P.send_dst); }
    G.t[rk] = r->D(2 * (size_t)P.capL); G.b_l[rk] = r->D(2 * (size_t)P.capL); G.Hll_inv[rk] = r->D(3 * (size_t)P.capL);
    if (h->world == 1 && g_coarse_nodes > 0) {
      plan_coarse(P, coarse_spacing(P.nP, std::min(g_coarse_nodes, kCzMaxNodes)), r->CZ);
      const CoarsePlan& C = r->CZ;
      if (C.nn > 0) {
        G.cz_h = C.h; G.cz_nn = C.nn; G.cz_ng = C.ng;
        G.cz_g_ptr = C.g_ptr.data(); G.cz_g_e = C.g_e.data(); G.cz_g_w = C.g_w.data(); G.cz_g_lm = C.g_lm.data();
        G.cz_p_ptr = C.p_ptr.data(); G.cz_p_e = C.p_e.data(); G.cz_p_w = C.p_w.data(); G.cz_rr = C.rr.data();
        G.cz_t_ptr = C.t_ptr.data(); G.cz_t_g = C.t_g.data();
        const int nc = 3 * C.nn;
        G.cz_G = r->D(6 * (size_t)C.ng); G.cz_A = r->D((size_t)nc * nc); G.cz_fail = &r->cz_fail;
        r->cz_work.assign((size_t)nc * cz_ld(nc), 0.0); r->cz_dinv.assign(nc, 0.0);
      }
    }
  }
  // "sgb_comm_connect": cross-link the peer tables
  for (int a = 0; a < h->world; ++a)
    for (int b = 0; b < h->world; ++b) {
      DevGraph& A = h->R[a]->G;
      const DevGraph& B = h->R[b]->G;
      for (int s = 0; s < 2; ++s) { A.pose_buf[s][b] = B.pose_buf[s][b]; A.lm_buf[s][b] = B.lm_buf[s][b]; }
      A.p[b] = B.p[b]; A.x_p[b] = B.x_p[b]; A.t[b] = B.t[b]; A.b_l[b] = B.b_l[b]; A.Hll_inv[b] = B.Hll_inv[b];
      // pushed halos: where rank a's rows start in rank b's halo; the two sides of the plan must agree (sgb_comm_connect)
      A.halo_base_at[b] = h->R[b]->P.halo_base[a];
      if (a != b && h->R[a]->P.pushed && h->R[b]->P.halo_cnt[a] != h->R[a]->P.send_cnt[b]) {
        h->err = "halo plan mismatch between ranks";
        if (status) *status = SGB_ERR_COMM;
      }
    }
  return h;
}
// wall time (ms) of the host symbolic phase of ONE rank: build_structure (+ rank filter) and the partition plan
double hs_time_symbolic(const sgb_graph_soa* g, int world, int rank, int filtered, int* n_pp, int* n_pl) {
  Structure S;
  LocalPlan P;
  std::string err;
  auto t0 = std::chrono::steady_clock::now();
  if (build_structure(*g, S, err, filtered ? world : 1, filtered ? rank : 0) != SGB_OK) return -1.0;
  if (partition_consume(S, world, rank, P, err) != SGB_OK) return -2.0;
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (n_pp) *n_pp = P.n_pp;
  if (n_pl) *n_pl = P.n_pl;
  return ms;
}
void hs_destroy(hs_handle* h) { delete h; }
const char* hs_error(hs_handle* h) { return h->err.c_str(); }

void hs_info(hs_handle* h, sgb_structure_info* o) {
  const Structure& S = h->S;
  o->n_free = S.Pf + S.Lf; o->n_free_poses = S.Pf; o->n_free_landmarks = S.Lf; o->n_blocks = (int)S.blk_row.size();
  o->scalar_dim = S.dim; o->n_active_pp = S.n_pp; o->n_active_pl = S.n_pl; o->coarse_nodes = h->R[0]->G.cz_h > 0 ? h->R[0]->G.cz_nn : 0; o->block_values = S.block_values;
}
void hs_structure(hs_handle* h, int32_t* kind, int32_t* index, int32_t* offset, int32_t* br, int32_t* bc, int32_t* bnr,
                  int32_t* bnc, int32_t* ph, int32_t* lh) {
  const Structure& S = h->S;
  auto cp = [](int32_t* d, const std::vector<int32_t>& v) { if (d && !v.empty()) std::memcpy(d, v.data(), v.size() * 4); };
  cp(kind, S.ord_kind); cp(index, S.ord_index); cp(offset, S.ord_offset);
  cp(br, S.blk_row); cp(bc, S.blk_col); cp(bnr, S.blk_nr); cp(bnc, S.blk_nc);
  if (ph) for (int i = 0; i < S.P_all; ++i) ph[i] = S.pose_h[i];
  if (lh) for (int i = 0; i < S.L_all; ++i) lh[i] = S.lm_h[i] >= 0 ? S.Pf + S.lm_h[i] : -1;
}
// padding statistics of the three SELL matrices summed over ranks: entries (incl. padding) and real blocks
void hs_sell_stats(hs_handle* h, int64_t* out /*[6]*/) {
  for (int i = 0; i < 6; ++i) out[i] = 0;
  for (auto& r : h->R) {
    const HostSell* m[3] = {&r->P.Hpp, &r->P.Hpl, &r->P.Hlp};
    for (int i = 0; i < 3; ++i) {
      out[2 * i] += m[i]->entries();
      for (int32_t c : m[i]->col) out[2 * i + 1] += c >= 0;
    }
  }
}
// per-rank partition summary: [world][8] = nP, nL, n_pp, n_pl, n_pp_owned, n_pl_owned, halo_p, halo_t
void hs_partition_stats(hs_handle* h, int64_t* out) {
  for (int k = 0; k < h->world; ++k) {
    const LocalPlan& P = h->R[k]->P;
    int64_t remote = 0;  // matrix columns that still name a row of another rank (0 with pushed halos)
    for (int32_t c : P.Hpp.col) remote += c >= 0 && (c >> kOwnerShift) != P.rank;
    for (int32_t c : P.Hlp.col) remote += c >= 0 && (c >> kOwnerShift) != P.rank;
    int64_t v[11] = {P.nP, P.nL, P.n_pp, P.n_pl, P.n_pp_owned, P.n_pl_owned, P.halo_p, P.halo_t, P.nL_owned, remote, P.nH};
    std::copy(v, v + 11, out + 11 * k);
  }
}

static void set_cur(hs_handle* h, int cur) {
  h->cur = cur;
  for (auto& r : h->R) r->G.cur = cur;
}

static void linearize(hs_handle* h, double chi[3]) {
  LinAcc acc;
  for (auto& r : h->R) {
    DevGraph& G = r->G;
    for (int lp = 0; lp < G.nP; ++lp) lin_pose_row(G, lp, acc);
    for (int ll = 0; ll < G.nL; ++ll) lin_lm_row(G, ll, acc);
  }
  chi[0] = acc.chi; chi[1] = acc.chi_r; chi[2] = acc.maxd;
}
static void chi2_edges(hs_handle* h, int buf, double chi[2]) {
  double c = 0, cr = 0;
  for (auto& r : h->R) {
    DevGraph& G = r->G;
    const double* pose = G.pose_buf[buf][G.rank];
    const double* lm = G.lm_buf[buf][G.rank];
    for (int k = 0; k < G.n_pp_owned; ++k) { double a, b; pp_chi(G, k, pose, &a, &b); c += a; cr += b; }
    for (int k = 0; k < G.n_pl_owned; ++k) { double a = pl_chi(G, k, pose, lm); c += a; cr += a; }
  }
  chi[0] = c; chi[1] = cr;
}

static void gather_vec(hs_handle* h, bool step, double* out) {
  const Structure& S = h->S;
  for (auto& r : h->R) {
    const DevGraph& G = r->G;
    const LocalPlan& P = r->P;
    const double* dp = step ? G.x_p[G.rank] : G.b_p;
    const double* dl = step ? G.x_l : G.b_l[G.rank];
    for (int l = 0; l < 3 * P.nP; ++l) out[3 * (size_t)P.p_begin + l] = dp[l];
    for (int l = 0; l < P.nL_owned; ++l) {
      size_t o = 3 * (size_t)S.Pf + 2 * (size_t)P.lm_global[l];
      out[o] = dl[2 * l];
      out[o + 1] = dl[2 * l + 1];
    }
  }
}

void hs_linearize(hs_handle* h, double* b, double* Hblocks, double* chi2) {
  const Structure& S = h->S;
  double chi[3];
  linearize(h, chi);
  if (chi2) { chi2[0] = chi[0]; chi2[1] = chi[1]; }
  if (b) gather_vec(h, false, b);
  if (Hblocks) {
    size_t o = 0;
    for (size_t k = 0; k < S.blk_row.size(); ++k) {
      int kind = S.blk_kind[k];
      // the owner is the rank that names itself (with rank-filtered structures a rank only knows the owners of the
      // blocks it has something to do with); take the values from it
      int owner = -1;
      for (int rk = 0; rk < h->world && owner < 0; ++rk)
        if (h->R[rk]->P.blk_owner[k] == rk) owner = rk;
      if (owner < 0) { o += (kind == 0 ? 9 : kind == 1 ? 6 : 4); continue; }
      const Rank& R = *h->R[owner];
      int e = R.P.blk_entry[k];
      const DevGraph& G = R.G;
      if (kind == 0) {
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Hblocks[o + c * 3 + r] = G.Hpp.vals[sell_vaddr(e, 9, 3 * r + c)];
        o += 9;
      } else if (kind == 1) {
        for (int c = 0; c < 2; ++c) for (int r = 0; r < 3; ++r) Hblocks[o + c * 3 + r] = G.Hpl.vals[sell_vaddr(e, 6, 2 * r + c)];
        o += 6;
      } else {
        double h11 = G.Hll[e], h12 = G.Hll[(size_t)G.nL + e], h22 = G.Hll[2 * (size_t)G.nL + e];
        Hblocks[o] = h11; Hblocks[o + 1] = h12; Hblocks[o + 2] = h12; Hblocks[o + 3] = h22;
        o += 4;
      }
    }
  }
}
// the landmark-major copy Hlp must hold exactly the blocks of Hpl: returns the max abs difference
double hs_check_hlp(hs_handle* h) {
  const Structure& S = h->S;
  double worst = 0;
  // global (pose hp, landmark hl) -> value from Hpl; compare with Hlp rows
  for (auto& r : h->R) {
    const DevGraph& G = r->G;
    const LocalPlan& P = r->P;
    for (int row = 0; row < G.nL; ++row) {
      int ll = row;
      int hl = P.lm_global[ll];
      int w = P.Hlp.width(P.Hlp.slice_of(row));
      for (int k = 0; k < w; ++k) {
        int e = P.Hlp.entry(row, k);
        int enc = G.Hlp.col[e];
        if (enc < 0) continue;
        int o = enc >> kOwnerShift, lp = enc & kLocalMask;
        if (P.pushed && lp >= P.capP) {  // a halo copy: the slot names the global row
          int hp = P.halo_src[lp - P.capP];
          o = hp / P.chunkP;
          lp = hp % P.chunkP;
        }
        const Rank& Q = *h->R[o];
        // find hl in pose row lp of rank o
        int s2 = lp >> 5, l2 = lp & 31, w2 = sell_width(Q.G.Hpl, s2);
        bool found = false;
        for (int k2 = 0; k2 < w2; ++k2) {
          int e2 = Q.G.Hpl.sbase[s2] + k2 * 32 + l2;
          if (Q.G.Hpl.col[e2] == (Q.P.enc_lm_here.empty() ? Q.P.enc_lm[hl] : Q.P.enc_lm_here[hl])) {
            for (int c = 0; c < 6; ++c) worst = std::max(worst, std::fabs(Q.G.Hpl.vals[sell_vaddr(e2, 6, c)] - G.Hlp.vals[sell_vaddr(e, 6, c)]));
            found = true;
          }
        }
        if (!found) worst = 1e300;
      }
    }
  }
  (void)S;
  return worst;
}

// mirrors k_setup_* + k_pcg (single-reduction PCG, sgb_kernels.cuh) + k_backsub; returns the pcg flag (0 ok, 1 maxit,
// 2 breakdown), iterations in *iters
static int solve(hs_handle* h, double lambda, int* iters, double* rel) {
  bool ok = true;
  for (auto& r : h->R) for (int ll = 0; ll < r->G.nL; ++ll) ok &= setup_lm_row(r->G, ll, lambda);
  for (auto& r : h->R) for (int lp = 0; lp < r->G.nP; ++lp) ok &= setup_pose_row(r->G, lp, lambda);
  for (auto& r : h->R) for (int ch = 0; ch < (r->G.nP + kChunk - 1) / kChunk; ++ch) ok &= setup_chunk(r->G, ch, lambda);
  // two-level preconditioner (world == 1): k_setup_coarse, then the coarse residual rc = R r follows r through the
  // recurrence rc -= alpha R s, and every z gets R^T (R S R^T)^-1 rc on top of the block-Jacobi term
  DevGraph& G0 = h->R[0]->G;
  const bool cz = h->world == 1 && G0.cz_h > 0;
  const int cz_nc = cz ? 3 * G0.cz_nn : 0;
  std::vector<double> rc(cz_nc, 0.0), yc(cz_nc, 0.0), tmp(cz_nc, 0.0);
  if (cz) coarse_factor(G0, lambda, h->R[0]->cz_work.data(), h->R[0]->cz_dinv.data(), G0.cz_A, G0.cz_fail, 0, 1, [] {});
  auto cz_restrict = [&](const double* v, double* out) {
    std::fill(out, out + cz_nc, 0.0);
    for (int i = 0; i < G0.nP; ++i) {
      const double wr = cz_wr(i, G0.cz_h), wl = 1.0 - wr;
      const int n0 = i / G0.cz_h;
      for (int c = 0; c < 3; ++c) {
        out[3 * n0 + c] += wl * v[3 * (size_t)i + c];
        out[3 * (n0 + 1) + c] += wr * v[3 * (size_t)i + c];
      }
    }
  };
  auto cz_solve = [&]() {
    for (int t = 0; t < cz_nc; ++t) {
      double a = 0.0;
      for (int j = 0; j < cz_nc; ++j) a += G0.cz_A[(size_t)j * cz_nc + t] * rc[j];
      yc[t] = a;
    }
  };
  auto cz_add = [&](int i, const double* r3, double z[3]) {  // z += (R^T yc)_i, returns r_i . (R^T yc)_i
    const double wr = cz_wr(i, G0.cz_h), wl = 1.0 - wr;
    const int n0 = i / G0.cz_h;
    double dot = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double v = wl * yc[3 * n0 + c] + wr * yc[3 * (n0 + 1) + c];
      z[c] += v;
      dot += r3[c] * v;
    }
    return dot;
  };
  double gam = 0;
  for (auto& rk : h->R) {
    DevGraph& G = rk->G;
    double* x = G.x_p[G.rank]; double* zin = G.p[G.rank];
    for (int i = 0; i < 3 * G.nP; ++i) { x[i] = 0; G.r[i] = G.bt[i]; G.d[i] = 0; G.s[i] = 0; }
    if (cz) { cz_restrict(G.r, rc.data()); cz_solve(); }
    for (int lp = 0; lp < G.nP; ++lp) {  // the device exchanges the chunk's residuals with shuffles; here they are in G.r
      double z[3];
      gam += precond_row_from(G, lp, G.r, z);
      if (cz) gam += cz_add(lp, G.r + 3 * (size_t)lp, z);
      for (int c = 0; c < 3; ++c) zin[3 * lp + c] = z[c];
      const double x0[3] = {0.0, 0.0, 0.0};
      if (G.pushed) push_halo_row(G, lp, z, x0);
    }
  }
  double gam0 = gam, gam_old = 0, alpha_old = 0;
  int it = 0, flag = 0;
  bool any_lm = false;  // with ghost rows capL is a per-rank quantity
  for (auto& rk : h->R) any_lm = any_lm || rk->G.capL > 0;
  if (!(gam0 > 0.0)) {
    flag = (gam0 == 0.0) ? 0 : 2;
  } else {
    double target = h->tol * h->tol * gam0;
    flag = 1;
    while (true) {
      if (!(gam == gam)) { flag = 2; break; }
      if (gam <= target) { flag = 0; break; }
      if (it >= h->maxit) { flag = 1; break; }
      double beta = it == 0 ? 0.0 : gam / gam_old;
      if (any_lm) for (auto& rk : h->R) for (int sl = 0; sl < rk->G.Hlp.nslices; ++sl) lm_slice_pass(rk->G, sl, 0);
      double del = 0;
      for (auto& rk : h->R) for (int lp = 0; lp < rk->G.nP; ++lp) del += schur_phaseB_row(rk->G, lp, lambda, beta);
      double denom = it == 0 ? del : del - beta * gam / alpha_old;
      if (!(denom > 0.0)) { flag = 2; break; }
      double alpha = gam / denom;
      double gnew = 0;
      for (auto& rk : h->R) {
        DevGraph& G = rk->G;
        double* x = G.x_p[G.rank]; double* zin = G.p[G.rank];
        for (size_t o = 0; o < 3 * (size_t)G.nP; ++o) {
          x[o] += alpha * G.d[o];
          G.r[o] = G.r[o] - alpha * G.s[o];
        }
        if (cz) {
          cz_restrict(G.s, tmp.data());
          for (int t = 0; t < cz_nc; ++t) rc[t] -= alpha * tmp[t];
          cz_solve();
        }
        for (int lp = 0; lp < G.nP; ++lp) {
          double z[3];
          gnew += precond_row_from(G, lp, G.r, z);
          if (cz) gnew += cz_add(lp, G.r + 3 * (size_t)lp, z);
          for (int c = 0; c < 3; ++c) zin[3 * (size_t)lp + c] = z[c];
          if (G.pushed) push_halo_row(G, lp, z, x + 3 * (size_t)lp);
        }
      }
      ++it;
      gam_old = gam; alpha_old = alpha; gam = gnew;
    }
  }
  for (auto& rk : h->R) for (int sl = 0; sl < rk->G.Hlp.nslices; ++sl) lm_slice_pass(rk->G, sl, 1);
  if (iters) *iters = it;
  if (rel) *rel = gam0 > 0 ? std::sqrt(std::fabs(gam) / gam0) : 0.0;
  if (!ok) flag = 2;
  return flag;
}

int hs_solve_once(hs_handle* h, double lambda, double* x, int* iters, double* rel) {
  double chi[3];
  linearize(h, chi);
  int flag = solve(h, lambda, iters, rel);
  if (x) gather_vec(h, true, x);
  return flag;
}

static double update_all(hs_handle* h, double lambda, int dst) {
  double scale = 0;
  for (auto& rk : h->R) {
    DevGraph& G = rk->G;
    for (int lp = 0; lp < G.nP; ++lp) scale += update_pose_row(G, lp, lambda, dst);
    for (int ll = 0; ll < G.nL_owned; ++ll) scale += update_lm_row(G, ll, lambda, dst);
  }
  return scale;
}

int hs_optimize(hs_handle* h, int algo, int max_iters, sgb_iter_stat* stats) {
  int done = 0, result = SGB_RESULT_OK;
  bool ok = true;
  for (int it = 0; it < max_iters && ok; ++it) {
    double chi[3];
    linearize(h, chi);
    double currentChi = chi[1], chi_lin = chi[1];
    int trials = 0, pcg_total = 0;
    double rho = 0, rel = 0;
    if (algo == SGB_ALGO_GN) {
      int iters = 0;
      int flag = solve(h, 0.0, &iters, &rel);
      pcg_total = iters;
      update_all(h, 0.0, h->cur);
      result = flag != 2 ? SGB_RESULT_OK : SGB_RESULT_FAIL;
      trials = 1;
    } else {
      if (it == 0) { h->lambda = 1e-5 * chi[2]; h->ni = 2; }
      while (true) {
        int iters = 0;
        int flag = solve(h, h->lambda, &iters, &rel);
        pcg_total += iters;
        double scale = update_all(h, h->lambda, h->cur ^ 1);
        double c2[2];
        chi2_edges(h, h->cur ^ 1, c2);
        double tempChi = flag != 2 ? c2[1] : DBL_MAX;
        rho = (currentChi - tempChi) / (scale + 1e-3);
        bool lambda_finite = true;
        if (rho > 0 && std::isfinite(tempChi)) {
          double a = 2 * rho - 1, alpha = std::min(1.0 - a * a * a, 2.0 / 3.0);
          h->lambda *= std::max(1.0 / 3.0, alpha);
          h->ni = 2;
          currentChi = tempChi;
          set_cur(h, h->cur ^ 1);
        } else {
          h->lambda *= h->ni;
          h->ni *= 2;
          lambda_finite = std::isfinite(h->lambda);
        }
        if (lambda_finite) trials++;
        bool again = rho < 0 && trials < 10 && lambda_finite;
        if (!again) { result = (trials == 10 || rho == 0 || !lambda_finite) ? SGB_RESULT_TERMINATE : SGB_RESULT_OK; break; }
      }
    }
    if (stats) {
      stats[it].iteration = it; stats[it].trials = trials; stats[it].result = result; stats[it].pcg_iters = pcg_total;
      stats[it].chi2 = currentChi; stats[it].lambda = algo == SGB_ALGO_LM ? h->lambda : 0.0; stats[it].rho = rho;
      stats[it].chi2_before = chi_lin; stats[it].pcg_residual = rel;
    }
    ok = result == SGB_RESULT_OK;
    ++done;
  }
  return result == SGB_RESULT_FAIL ? 0 : done;
}

// estimates of virtual rank `rank` (all replicas must agree)
void hs_get_estimates(hs_handle* h, int rank, double* pose, double* lm) {
  const DevGraph& G = h->R[rank]->G;
  if (pose) std::copy(G.pose_buf[h->cur][rank], G.pose_buf[h->cur][rank] + 3 * (size_t)h->S.P_all, pose);
  if (lm) std::copy(G.lm_buf[h->cur][rank], G.lm_buf[h->cur][rank] + 2 * (size_t)h->S.L_all, lm);
}
void hs_chi2(hs_handle* h, double* chi2) { chi2_edges(h, h->cur, chi2); }

// block-Jacobi preconditioner of rank 0 after a set-up at `lambda` (linearises first): out [nP][3][12], row-major
int hs_preconditioner(hs_handle* h, double lambda, float* out) {
  double chi[3];
  linearize(h, chi);
  bool ok = true;
  for (auto& r : h->R) for (int ll = 0; ll < r->G.nL; ++ll) ok &= setup_lm_row(r->G, ll, lambda);
  const DevGraph& G = h->R[0]->G;
  for (int ch = 0; ch < (G.nP + kChunk - 1) / kChunk; ++ch) ok &= setup_chunk(G, ch, lambda);
  for (int lp = 0; lp < G.nP; ++lp)
    for (int idx = 0; idx < 9 * kChunk; ++idx) out[(size_t)lp * 9 * kChunk + idx] = G.Cinv[((size_t)(idx >> 2) * G.nP + lp) * 4 + (idx & 3)];
  return ok ? 0 : 1;
}

// coarse space of the two-level preconditioner (sgb_coarse.h) of rank 0: info[0] = node spacing h (0 = not planned),
// info[1] = nodes, info[2] = 1 when the last factorisation failed; out (may be NULL) = (R S R^T)^-1 of the last solve, [3nn][3nn]
void hs_coarse(hs_handle* h, int32_t* info, double* out) {
  const Rank& r = *h->R[0];
  info[0] = r.G.cz_h; info[1] = r.G.cz_nn; info[2] = r.cz_fail;
  if (out && r.G.cz_h > 0) std::memcpy(out, r.G.cz_A, sizeof(double) * 9 * (size_t)r.G.cz_nn * r.G.cz_nn);
}

}  // extern "C"

// timing of the host symbolic phase alone (seconds): out[0] = build_structure, out[1] = partition(world, rank 0)
#include <chrono>
extern "C" int hs_time_structure(const sgb_graph_soa* g, int world, double* out) {
  Structure S;
  LocalPlan P;
  std::string err;
  auto t0 = std::chrono::steady_clock::now();
  sgb_status st = build_structure(*g, S, err);
  auto t1 = std::chrono::steady_clock::now();
  if (st == SGB_OK) st = partition(S, world, 0, P, err);
  auto t2 = std::chrono::steady_clock::now();
  out[0] = std::chrono::duration<double>(t1 - t0).count();
  out[1] = std::chrono::duration<double>(t2 - t1).count();
  return st;
}
