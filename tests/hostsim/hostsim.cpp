// hostsim.cpp -- TEST HARNESS ONLY (never shipped, never loaded by the sparse_gslam_b200 package).
//
// Executes the row bodies of the CUDA kernels (sparse-gslam_b200/csrc/sgb_rows.h) serially on the host, over the
// structure produced by the real host-side symbolic phase (sgb_structure.cpp), so that the scatter maps, SELL
// addressing and per-row arithmetic can be checked against the oracle in the CPU-only test tier before GPU time is
// spent. The orchestration below mirrors sgb_backend.cu / k_pcg step by step. The product library libsgb.so does not
// contain this file and has no CPU path.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sgb_capi.h"
#include "../../sparse-gslam_b200/csrc/sgb_rows.h"
#include "../../sparse-gslam_b200/csrc/sgb_structure.h"

using namespace sgb;

struct hs_handle {
  Structure S;
  DevGraph G;
  std::vector<std::vector<double>> dbl;
  std::vector<std::vector<int32_t>> ints;
  std::string err;
  double lambda = 0, ni = 2;
  double tol = 1e-10;
  int maxit = 0;
  double* D(size_t n) { dbl.emplace_back(std::max<size_t>(n, 1), 0.0); return dbl.back().data(); }
  const int32_t* I(const std::vector<int32_t>& v) { ints.push_back(v); if (ints.back().empty()) ints.back().push_back(0); return ints.back().data(); }
};

static void mk_sell(hs_handle* h, Sell* out, const HostSell& s, int NC) {
  out->rows = s.rows;
  out->nslices = s.nslices;
  out->sbase = h->I(s.sbase);
  out->col = h->I(s.col);
  out->vals = h->D((size_t)s.entries() * NC);
}

extern "C" {

hs_handle* hs_create(const sgb_graph_soa* g, int jac_numeric, double tol, int maxit, int* status) {
  hs_handle* h = new hs_handle();
  sgb_status st = build_structure(*g, h->S, h->err);
  if (status) *status = st;
  if (st != SGB_OK) return h;
  const Structure& S = h->S;
  DevGraph& G = h->G;
  std::memset(&G, 0, sizeof G);
  G.P_all = S.P_all; G.L_all = S.L_all; G.Pf = S.Pf; G.Lf = S.Lf; G.n_pp = S.n_pp; G.n_pl = S.n_pl;
  G.has_robust = S.has_robust; G.jac_numeric = jac_numeric;
  h->tol = tol > 0 ? tol : 1e-10;
  h->maxit = maxit > 0 ? maxit : std::max(100, 12 * S.Pf);
  size_t np = 3 * (size_t)S.P_all, nl = 2 * (size_t)S.L_all;
  G.pose = h->D(np); G.lm = h->D(nl); G.pose_trial = h->D(np); G.lm_trial = h->D(nl);
  std::copy(g->pose_est, g->pose_est + np, G.pose); std::copy(g->pose_est, g->pose_est + np, G.pose_trial);
  std::copy(g->lm_est, g->lm_est + nl, G.lm); std::copy(g->lm_est, g->lm_est + nl, G.lm_trial);
  G.pose_of_h = h->I(S.pose_of_h); G.lm_of_h = h->I(S.lm_of_h);
  G.pp_i = h->I(S.pp_i); G.pp_j = h->I(S.pp_j); G.pp_hi = h->I(S.pp_hi); G.pp_hj = h->I(S.pp_hj);
  G.pp_e_ij = h->I(S.pp_e_ij); G.pp_e_ji = h->I(S.pp_e_ji); G.pp_dup = h->I(S.pp_dup);
  G.pl_p = h->I(S.pl_p); G.pl_l = h->I(S.pl_l); G.pl_hp = h->I(S.pl_hp); G.pl_hl = h->I(S.pl_hl);
  G.pl_e_pl = h->I(S.pl_e_pl); G.pl_e_lp = h->I(S.pl_e_lp); G.pl_dup = h->I(S.pl_dup);
  G.pinc_ptr = h->I(S.pinc_ptr); G.pinc = h->I(S.pinc); G.linc_ptr = h->I(S.linc_ptr); G.linc = h->I(S.linc);
  G.hpp_diag = h->I(S.hpp_diag); G.lp_row2h = h->I(S.lp_row2h); G.lp_h2row = h->I(S.lp_h2row);
  double* zinv = h->D(3 * (size_t)S.n_pp); double* info = h->D(6 * (size_t)S.n_pp); double* phi = h->D(S.n_pp);
  for (int k = 0; k < S.n_pp; ++k) {
    int s = S.pp_src[k];
    double x = g->pp_z[3 * (size_t)s], y = g->pp_z[3 * (size_t)s + 1], th = g->pp_z[3 * (size_t)s + 2];
    double thi = normalize_theta(-th), c = std::cos(thi), sn = std::sin(thi);
    zinv[k] = c * (-x) - sn * (-y);
    zinv[(size_t)S.n_pp + k] = sn * (-x) + c * (-y);
    zinv[2 * (size_t)S.n_pp + k] = thi;
    for (int c6 = 0; c6 < 6; ++c6) info[(size_t)c6 * S.n_pp + k] = g->pp_info[6 * (size_t)s + c6];
    phi[k] = g->pp_phi ? g->pp_phi[s] : 0.0;
  }
  G.pp_zinv = zinv; G.pp_info = info; G.pp_phi = phi;
  double* z = h->D(2 * (size_t)S.n_pl); double* linfo = h->D(3 * (size_t)S.n_pl);
  for (int k = 0; k < S.n_pl; ++k) {
    int s = S.pl_src[k];
    z[k] = g->pl_z[2 * (size_t)s]; z[(size_t)S.n_pl + k] = g->pl_z[2 * (size_t)s + 1];
    for (int c3 = 0; c3 < 3; ++c3) linfo[(size_t)c3 * S.n_pl + k] = g->pl_info[3 * (size_t)s + c3];
  }
  G.pl_z = z; G.pl_info = linfo;
  mk_sell(h, &G.Hpp, S.Hpp, 9); mk_sell(h, &G.Hpl, S.Hpl, 6); mk_sell(h, &G.Hlp, S.Hlp, 6);
  size_t n3 = 3 * (size_t)S.Pf, n2 = 2 * (size_t)S.Lf;
  G.Hll = h->D(3 * (size_t)S.Lf); G.Hll_inv = h->D(3 * (size_t)S.Lf); G.b = h->D(n3 + n2); G.x = h->D(n3 + n2);
  G.Minv = h->D(9 * (size_t)S.Pf); G.bt = h->D(n3); G.r = h->D(n3); G.z = h->D(n3); G.p = h->D(n3); G.q = h->D(n3); G.t = h->D(n2);
  return h;
}
void hs_destroy(hs_handle* h) { delete h; }
const char* hs_error(hs_handle* h) { return h->err.c_str(); }

void hs_info(hs_handle* h, sgb_structure_info* o) {
  const Structure& S = h->S;
  o->n_free = S.Pf + S.Lf; o->n_free_poses = S.Pf; o->n_free_landmarks = S.Lf; o->n_blocks = (int)S.blk_row.size();
  o->scalar_dim = S.dim; o->n_active_pp = S.n_pp; o->n_active_pl = S.n_pl; o->reserved = 0; o->block_values = S.block_values;
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
// padding statistics of the three SELL matrices: entries (incl. padding) and real blocks
void hs_sell_stats(hs_handle* h, int64_t* out /*[6]*/) {
  const HostSell* m[3] = {&h->S.Hpp, &h->S.Hpl, &h->S.Hlp};
  for (int i = 0; i < 3; ++i) {
    out[2 * i] = m[i]->entries();
    int64_t real = 0;
    for (int32_t c : m[i]->col) real += c >= 0;
    out[2 * i + 1] = real;
  }
}

static void linearize(hs_handle* h, double chi[3]) {
  DevGraph& G = h->G;
  LinAcc acc;
  for (int hp = 0; hp < G.Pf; ++hp) lin_pose_row(G, hp, acc);
  for (int hl = 0; hl < G.Lf; ++hl) lin_lm_row(G, hl, acc);
  chi[0] = acc.chi; chi[1] = acc.chi_r; chi[2] = acc.maxd;
}
static void chi2_edges(hs_handle* h, const double* pose, const double* lm, double chi[2]) {
  DevGraph& G = h->G;
  double c = 0, cr = 0;
  for (int k = 0; k < G.n_pp; ++k) { double a, b; pp_chi(G, k, pose, &a, &b); c += a; cr += b; }
  for (int k = 0; k < G.n_pl; ++k) { double a = pl_chi(G, k, pose, lm); c += a; cr += a; }
  chi[0] = c; chi[1] = cr;
}

void hs_linearize(hs_handle* h, double* b, double* Hblocks, double* chi2) {
  const Structure& S = h->S;
  DevGraph& G = h->G;
  double chi[3];
  linearize(h, chi);
  if (chi2) { chi2[0] = chi[0]; chi2[1] = chi[1]; }
  if (b) std::copy(G.b, G.b + S.dim, b);
  if (Hblocks) {
    size_t o = 0;
    for (size_t k = 0; k < S.blk_row.size(); ++k) {
      int kind = S.blk_kind[k], e = S.blk_entry[k];
      if (kind == 0) {
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Hblocks[o + c * 3 + r] = G.Hpp.vals[sell_vaddr(e, 9, 3 * r + c)];
        o += 9;
      } else if (kind == 1) {
        for (int c = 0; c < 2; ++c) for (int r = 0; r < 3; ++r) Hblocks[o + c * 3 + r] = G.Hpl.vals[sell_vaddr(e, 6, 2 * r + c)];
        o += 6;
      } else {
        double h11 = G.Hll[e], h12 = G.Hll[(size_t)S.Lf + e], h22 = G.Hll[2 * (size_t)S.Lf + e];
        Hblocks[o] = h11; Hblocks[o + 1] = h12; Hblocks[o + 2] = h12; Hblocks[o + 3] = h22;
        o += 4;
      }
    }
    // the landmark-major copy must hold the same blocks
  }
}

// mirrors k_setup_* + k_pcg + k_backsub; returns pcg flag (0 ok, 1 maxit, 2 breakdown), iterations in *iters
static int solve(hs_handle* h, double lambda, int* iters, double* rel) {
  DevGraph& G = h->G;
  bool ok = true;
  for (int hl = 0; hl < G.Lf; ++hl) ok &= setup_lm_row(G, hl, lambda);
  for (int hp = 0; hp < G.Pf; ++hp) ok &= setup_pose_row(G, hp, lambda);
  double rz = 0;
  for (int hp = 0; hp < G.Pf; ++hp) {
    double r[3] = {G.bt[3 * hp], G.bt[3 * hp + 1], G.bt[3 * hp + 2]}, z[3];
    rz += precond_row(G, hp, r, z);
    for (int c = 0; c < 3; ++c) { G.x[3 * hp + c] = 0; G.r[3 * hp + c] = r[c]; G.p[3 * hp + c] = z[c]; }
  }
  double rz0 = rz;
  int it = 0, flag = 0;
  if (!(rz0 > 0.0)) {
    flag = (rz0 == 0.0) ? 0 : 2;
  } else {
    double target = h->tol * h->tol * rz0;
    flag = 1;
    while (it < h->maxit) {
      for (int row = 0; row < G.Lf; ++row) schur_phaseA_row(G, row, G.p);
      double pq = 0;
      for (int hp = 0; hp < G.Pf; ++hp) pq += schur_phaseB_row(G, hp, G.p, lambda, G.q);
      if (!(pq > 0.0)) { flag = 2; break; }
      double alpha = rz / pq, rzn = 0;
      for (int hp = 0; hp < G.Pf; ++hp) {
        double r[3], z[3];
        for (int c = 0; c < 3; ++c) { size_t o = 3 * (size_t)hp + c; G.x[o] += alpha * G.p[o]; r[c] = G.r[o] - alpha * G.q[o]; G.r[o] = r[c]; }
        rzn += precond_row(G, hp, r, z);
        for (int c = 0; c < 3; ++c) G.z[3 * (size_t)hp + c] = z[c];
      }
      ++it;
      if (!(rzn == rzn)) { flag = 2; break; }
      if (rzn <= target) { rz = rzn; flag = 0; break; }
      double beta = rzn / rz;
      rz = rzn;
      for (size_t o = 0; o < 3 * (size_t)G.Pf; ++o) G.p[o] = G.z[o] + beta * G.p[o];
    }
  }
  for (int row = 0; row < G.Lf; ++row) backsub_lm_row(G, row);
  if (iters) *iters = it;
  if (rel) *rel = rz0 > 0 ? std::sqrt(std::fabs(rz) / rz0) : 0.0;
  if (!ok) flag = 2;
  return flag;
}

int hs_solve_once(hs_handle* h, double lambda, double* x, int* iters, double* rel) {
  double chi[3];
  linearize(h, chi);
  int flag = solve(h, lambda, iters, rel);
  if (x) std::copy(h->G.x, h->G.x + h->S.dim, x);
  return flag;
}

int hs_optimize(hs_handle* h, int algo, int max_iters, sgb_iter_stat* stats) {
  DevGraph& G = h->G;
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
      for (int hp = 0; hp < G.Pf; ++hp) update_pose_row(G, hp, 0.0, G.pose, G.pose);
      for (int hl = 0; hl < G.Lf; ++hl) update_lm_row(G, hl, 0.0, G.lm, G.lm);
      result = flag != 2 ? SGB_RESULT_OK : SGB_RESULT_FAIL;
      trials = 1;
    } else {
      if (it == 0) { h->lambda = 1e-5 * chi[2]; h->ni = 2; }
      while (true) {
        int iters = 0;
        int flag = solve(h, h->lambda, &iters, &rel);
        pcg_total += iters;
        double scale = 0;
        for (int hp = 0; hp < G.Pf; ++hp) scale += update_pose_row(G, hp, h->lambda, G.pose, G.pose_trial);
        for (int hl = 0; hl < G.Lf; ++hl) scale += update_lm_row(G, hl, h->lambda, G.lm, G.lm_trial);
        double c2[2];
        chi2_edges(h, G.pose_trial, G.lm_trial, c2);
        double tempChi = flag != 2 ? c2[1] : DBL_MAX;
        rho = (currentChi - tempChi) / (scale + 1e-3);
        bool lambda_finite = true;
        if (rho > 0 && std::isfinite(tempChi)) {
          double a = 2 * rho - 1, alpha = std::min(1.0 - a * a * a, 2.0 / 3.0);
          h->lambda *= std::max(1.0 / 3.0, alpha);
          h->ni = 2;
          currentChi = tempChi;
          std::swap(G.pose, G.pose_trial);
          std::swap(G.lm, G.lm_trial);
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

void hs_get_estimates(hs_handle* h, double* pose, double* lm) {
  if (pose) std::copy(h->G.pose, h->G.pose + 3 * (size_t)h->S.P_all, pose);
  if (lm) std::copy(h->G.lm, h->G.lm + 2 * (size_t)h->S.L_all, lm);
}
void hs_chi2(hs_handle* h, double* chi2) { chi2_edges(h, h->G.pose, h->G.lm, chi2); }

}  // extern "C"
