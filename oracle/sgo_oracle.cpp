/*
 * sgo_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE). PARITY UNPINNED (see sgo_oracle.h).
 *
 * Restates, step by step, what the reference executes inside
 * g2o::SparseOptimizer::optimize() for its LM landmark graph and GN pose graph
 * (reference call sites: src/sparse_gslam/src/graphs.cpp:9-23, drone.cpp:146-156,
 * submap_loop_closer.cpp:286-288, log_runner.cpp:203-204). The g2o/Eigen parts are
 * not vendored in /root/reference; they are restated from the published
 * libg2o 2020.5.29 algorithm. Single-threaded like
 * the reference build (no OpenMP; Eigen Simplicial Cholesky is serial).
 */
#include "sgo_oracle.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <vector>

namespace {

// ---------------------------------------------------------------- A.1 scalars and SE2
// g2o/stuff/misc.h normalize_theta (floor-based wrap into [-pi, pi))
inline double normalize_theta(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  double multiplier = std::floor(theta / (2 * M_PI));
  theta = theta - multiplier * 2 * M_PI;
  if (theta >= M_PI) theta -= 2 * M_PI;
  if (theta < -M_PI) theta += 2 * M_PI;
  return theta;
}

// g2o/types/slam2d/se2.h
struct SE2 {
  double x, y, th;
};
inline SE2 se2_mul(const SE2& a, const SE2& b) {
  double c = std::cos(a.th), s = std::sin(a.th);
  SE2 r;
  r.x = a.x + (c * b.x - s * b.y);
  r.y = a.y + (s * b.x + c * b.y);
  r.th = normalize_theta(a.th + b.th);
  return r;
}
inline SE2 se2_inv(const SE2& a) {
  SE2 r;
  r.th = normalize_theta(-a.th);
  double c = std::cos(r.th), s = std::sin(r.th);
  double mx = -a.x, my = -a.y;
  r.x = c * mx - s * my;
  r.y = s * mx + c * my;
  return r;
}

// ls_extractor/utils.h:23-30 checkRhoTheta and :32-45 transform_line
inline void transform_line(const double line[2], double tx, double ty, double angle, double out[2]) {
  double rho = line[0], al = line[1];
  al += angle;
  if (al > M_PI) al -= 2 * M_PI;
  if (al < -M_PI) al += 2 * M_PI;
  double nx = std::cos(al), ny = std::sin(al);
  rho += tx * nx + ty * ny;
  if (rho < 0.0) {
    rho = -rho;
    al += M_PI;
    if (al > M_PI) al -= 2 * M_PI;
  }
  out[0] = rho;
  out[1] = al;
}

// g2o_bindings/edge_se2_rhotheta.cpp:9-16
inline void pl_error(const double pose[3], const double line[2], const double z[2], double e[2]) {
  SE2 p{pose[0], pose[1], pose[2]};
  SE2 pinv = se2_inv(p);
  double r[2];
  transform_line(line, pinv.x, pinv.y, pinv.th, r);
  e[0] = z[0] - r[0];
  e[1] = z[1] - r[1];
  e[1] = normalize_theta(e[1]);
}

// VertexSE2::oplusImpl (A.1): t += u (world frame), theta = normalize(theta + u2)
inline void pose_oplus(double p[3], const double u[3]) {
  p[0] += u[0];
  p[1] += u[1];
  p[2] = normalize_theta(p[2] + u[2]);
}
// g2o_bindings/vertex_rhotheta.cpp:30-34: no wrap (return value of normalize_theta discarded)
inline void lm_oplus(double l[2], const double u[2]) {
  l[0] += u[0];
  l[1] += u[1];
}

struct EdgePP {
  int i, j;
  double z[3];
  SE2 zinv;  // EdgeSE2::setMeasurement caches the inverse
  double info[9];
  double phi;
  int64_t seq;
};
struct EdgePL {
  int p, l;
  double z[2];
  double info[4];
  int64_t seq;
};

// -------------------------------------------------------------- minimum-degree ordering on the block graph
// (the reference uses Eigen's scalar AMD; any fill-reducing ordering gives the same result up to rounding)
std::vector<int> min_degree_order(int n, const std::vector<std::vector<int>>& adj0) {
  std::vector<std::vector<int>> adj(adj0), elems(n), evars(n);
  std::vector<char> is_elem(n, 0), dead(n, 0), edead(n, 0);
  std::vector<int> deg(n), mark(n, -1), wstamp(n, -1), w(n, 0);
  std::vector<int> head(n + 1, -1), nxt(n, -1), prv(n, -1);
  auto bucket_insert = [&](int v) {
    int d = deg[v];
    nxt[v] = head[d];
    prv[v] = -1;
    if (head[d] >= 0) prv[head[d]] = v;
    head[d] = v;
  };
  auto bucket_remove = [&](int v) {
    int d = deg[v];
    if (prv[v] >= 0) nxt[prv[v]] = nxt[v]; else head[d] = nxt[v];
    if (nxt[v] >= 0) prv[nxt[v]] = prv[v];
  };
  for (int i = 0; i < n; ++i) {
    deg[i] = (int)adj[i].size();
    bucket_insert(i);
  }
  std::vector<int> order;
  order.reserve(n);
  int mind = 0;
  std::vector<int> Lp;
  for (int k = 0; k < n; ++k) {
    while (mind <= n && head[mind] < 0) ++mind;
    int p = head[mind];
    bucket_remove(p);
    order.push_back(p);
    dead[p] = 1;
    Lp.clear();
    mark[p] = k;
    for (int v : adj[p])
      if (!dead[v] && mark[v] != k) { mark[v] = k; Lp.push_back(v); }
    for (int e : elems[p]) {
      if (edead[e]) continue;
      for (int v : evars[e])
        if (!dead[v] && mark[v] != k) { mark[v] = k; Lp.push_back(v); }
      edead[e] = 1;
      std::vector<int>().swap(evars[e]);
    }
    std::vector<int>().swap(adj[p]);
    std::vector<int>().swap(elems[p]);
    is_elem[p] = 1;
    evars[p] = Lp;
    int lp = (int)Lp.size();
    // w(e) = |L_e \ L_p|
    for (int i : Lp)
      for (int e : elems[i]) {
        if (edead[e]) continue;
        if (wstamp[e] != k) { wstamp[e] = k; w[e] = (int)evars[e].size(); }
        w[e] -= 1;
      }
    for (int i : Lp) {
      bucket_remove(i);
      // prune variable adjacency: drop dead vars and vars now covered by element p
      auto& a = adj[i];
      size_t o = 0;
      for (size_t t = 0; t < a.size(); ++t) {
        int v = a[t];
        if (dead[v] || mark[v] == k) continue;
        a[o++] = v;
      }
      a.resize(o);
      auto& el = elems[i];
      o = 0;
      long d = (long)a.size() + (lp - 1);
      for (size_t t = 0; t < el.size(); ++t) {
        int e = el[t];
        if (edead[e]) continue;
        if (w[e] <= 0) continue;  // e subset of L_p: absorbed (killed below)
        d += w[e];
        el[o++] = e;
      }
      el.resize(o);
      el.push_back(p);
      long cap = (long)(n - k - 1);
      long dn = std::min<long>(std::min<long>(d, (long)deg[i] + (lp - 1)), cap);
      if (dn < 0) dn = 0;
      deg[i] = (int)dn;
    }
    for (int i : Lp) {
      bucket_insert(i);
      if (deg[i] < mind) mind = deg[i];
    }
  }
  return order;
}

// -------------------------------------------------------------- sparse LDLt (up-looking, elimination tree)
struct LDLt {
  int n = 0;
  std::vector<int> Cp, Ci;      // upper(P A P^T) in CSC
  std::vector<double> Cx;
  std::vector<int> amap;        // A entry -> C position
  std::vector<int> perm, pinv;  // perm[k] = original index of permuted k
  std::vector<int64_t> Lp;       // 64-bit: nnz(L) can exceed 2^31 on the 1M-pose graphs
  std::vector<int> Li, Parent, Lnz, Flag, Pattern;
  std::vector<double> Lx, D, Y;
  bool analyzed = false;

  // Ap/Ai: upper-triangular CSC of A (row <= col)
  void analyze(int n_, const std::vector<int>& Ap, const std::vector<int>& Ai, const std::vector<int>& perm_) {
    n = n_;
    perm = perm_;
    pinv.assign(n, 0);
    for (int k = 0; k < n; ++k) pinv[perm[k]] = k;
    int nnz = Ap[n];
    std::vector<int> cnt(n + 1, 0);
    for (int c = 0; c < n; ++c)
      for (int p = Ap[c]; p < Ap[c + 1]; ++p) {
        int i = pinv[Ai[p]], j = pinv[c];
        cnt[std::max(i, j) + 1]++;
      }
    Cp.assign(n + 1, 0);
    for (int c = 0; c < n; ++c) Cp[c + 1] = Cp[c] + cnt[c + 1];
    Ci.assign(nnz, 0);
    Cx.assign(nnz, 0.0);
    amap.assign(nnz, 0);
    std::vector<int> pos(Cp.begin(), Cp.end() - 1);
    for (int c = 0; c < n; ++c)
      for (int p = Ap[c]; p < Ap[c + 1]; ++p) {
        int i = pinv[Ai[p]], j = pinv[c];
        int cc = std::max(i, j), rr = std::min(i, j);
        int q = pos[cc]++;
        Ci[q] = rr;
        amap[p] = q;
      }
    // symbolic: etree + column counts
    Lp.assign(n + 1, 0);
    Parent.assign(n, -1);
    Lnz.assign(n, 0);
    Flag.assign(n, 0);
    for (int k = 0; k < n; ++k) {
      Parent[k] = -1;
      Flag[k] = k;
      Lnz[k] = 0;
      for (int p = Cp[k]; p < Cp[k + 1]; ++p) {
        int i = Ci[p];
        if (i < k) {
          for (; Flag[i] != k; i = Parent[i]) {
            if (Parent[i] == -1) Parent[i] = k;
            Lnz[i]++;
            Flag[i] = k;
          }
        }
      }
    }
    int64_t tot = 0;
    for (int k = 0; k < n; ++k) {
      Lp[k] = tot;
      tot += Lnz[k];
    }
    Lp[n] = tot;
    Li.assign(tot, 0);
    Lx.assign(tot, 0.0);
    D.assign(n, 0.0);
    Y.assign(n, 0.0);
    Pattern.assign(n, 0);
    analyzed = true;
  }
  // returns false on zero / non-finite pivot (Eigen SimplicialLDLT reports NumericalIssue)
  bool factorize(const std::vector<double>& Ax) {
    for (size_t p = 0; p < Ax.size(); ++p) Cx[amap[p]] = Ax[p];
    for (int k = 0; k < n; ++k) {
      Y[k] = 0.0;
      int top = n;
      Flag[k] = k;
      Lnz[k] = 0;
      for (int p = Cp[k]; p < Cp[k + 1]; ++p) {
        int i = Ci[p];
        if (i <= k) {
          Y[i] += Cx[p];
          int len = 0;
          for (; Flag[i] != k; i = Parent[i]) {
            Pattern[len++] = i;
            Flag[i] = k;
          }
          while (len > 0) Pattern[--top] = Pattern[--len];
        }
      }
      D[k] = Y[k];
      Y[k] = 0.0;
      for (; top < n; ++top) {
        int i = Pattern[top];
        double yi = Y[i];
        Y[i] = 0.0;
        int64_t p2 = Lp[i] + Lnz[i];
        for (int64_t p = Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
        double l_ki = yi / D[i];
        D[k] -= l_ki * yi;
        Li[p2] = k;
        Lx[p2] = l_ki;
        Lnz[i]++;
      }
      if (D[k] == 0.0 || !std::isfinite(D[k])) return false;
    }
    return true;
  }
  void solve(const double* b, double* x) const {
    std::vector<double> y(n);
    for (int k = 0; k < n; ++k) y[k] = b[perm[k]];
    for (int j = 0; j < n; ++j) {
      int64_t p2 = Lp[j] + Lnz[j];
      for (int64_t p = Lp[j]; p < p2; ++p) y[Li[p]] -= Lx[p] * y[j];
    }
    for (int j = 0; j < n; ++j) y[j] /= D[j];
    for (int j = n - 1; j >= 0; --j) {
      int64_t p2 = Lp[j] + Lnz[j];
      for (int64_t p = Lp[j]; p < p2; ++p) y[j] -= Lx[p] * y[Li[p]];
    }
    for (int k = 0; k < n; ++k) x[perm[k]] = y[k];
  }
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct sgo_handle {
  // graph
  std::vector<int> pose_id, lm_id;
  std::vector<double> pose, lm;  // estimates
  std::vector<uint8_t> pose_fixed, lm_fixed;
  std::vector<EdgePP> pp;
  std::vector<EdgePL> pl;
  // active sets (A.5)
  struct ERef { int type, idx; int64_t seq; };
  std::vector<ERef> active_edges;          // sorted by internal id
  std::vector<int> pose_hidx, lm_hidx;     // hessian index, -1 fixed/inactive
  std::vector<char> pose_active, lm_active;
  struct VRef { int kind, idx; };
  std::vector<VRef> ivmap;
  bool initialized = false;
  // structure (A.5 buildStructure)
  bool structure_built = false;
  std::vector<int> voff;                     // scalar offset per hessian index
  std::vector<int> vdim;
  int ndim = 0;
  std::vector<std::map<int, int>> blockCols; // col -> (row -> value offset)
  std::vector<int> diag_off;                 // per hessian idx
  struct EMap { int off_ii, off_jj, off_ij; bool transposed; };
  std::vector<EMap> emap;                    // per active edge
  std::vector<double> Hval, b, x;
  std::vector<double> pose_b, lm_b;          // unused helper
  // per-edge linearisation results
  std::vector<double> pp_err, pp_A, pp_B, pl_err, pl_A, pl_B;
  // linear solver (A.7)
  std::vector<int> Ap, Ai;
  std::vector<int> a_src;                    // CCS entry -> index into Hval
  std::vector<double> Ax;
  LDLt chol;
  bool solver_init = false;
  // LM state
  double lambda = 0, ni = 2;
  double prof[4] = {0, 0, 0, 0};
};

namespace {

void edge_errors(sgo_handle* h) {
  // SparseOptimizer::computeActiveErrors
  for (auto& er : h->active_edges) {
    if (er.type == 0) {
      const EdgePP& e = h->pp[er.idx];
      SE2 xi{h->pose[3 * e.i], h->pose[3 * e.i + 1], h->pose[3 * e.i + 2]};
      SE2 xj{h->pose[3 * e.j], h->pose[3 * e.j + 1], h->pose[3 * e.j + 2]};
      SE2 d = se2_mul(e.zinv, se2_mul(se2_inv(xi), xj));  // EdgeSE2::computeError (A.2)
      double* o = &h->pp_err[3 * er.idx];
      o[0] = d.x; o[1] = d.y; o[2] = d.th;
    } else {
      const EdgePL& e = h->pl[er.idx];
      pl_error(&h->pose[3 * e.p], &h->lm[2 * e.l], e.z, &h->pl_err[2 * er.idx]);
    }
  }
}

inline double pp_chi2(const EdgePP& e, const double* err) {
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = e.info[3 * r] * err[0] + e.info[3 * r + 1] * err[1] + e.info[3 * r + 2] * err[2];
  return err[0] * t[0] + err[1] * t[1] + err[2] * t[2];
}
inline double pl_chi2(const EdgePL& e, const double* err) {
  double t0 = e.info[0] * err[0] + e.info[1] * err[1];
  double t1 = e.info[2] * err[0] + e.info[3] * err[1];
  return err[0] * t0 + err[1] * t1;
}
// RobustKernelDCS::robustify (A.4)
inline void dcs(double phi, double e2, double rho[3]) {
  double scale = (2.0 * phi) / (phi + e2);
  if (scale >= 1.0) {
    rho[0] = e2; rho[1] = 1.0; rho[2] = 0.0;
  } else {
    rho[0] = scale * e2 * scale; rho[1] = scale * scale; rho[2] = 0.0;
  }
}

void chi2_sums(sgo_handle* h, double out[2]) {
  double c = 0, cr = 0;
  for (auto& er : h->active_edges) {
    if (er.type == 0) {
      const EdgePP& e = h->pp[er.idx];
      double v = pp_chi2(e, &h->pp_err[3 * er.idx]);
      c += v;
      if (e.phi > 0) { double rho[3]; dcs(e.phi, v, rho); cr += rho[0]; } else cr += v;
    } else {
      double v = pl_chi2(h->pl[er.idx], &h->pl_err[2 * er.idx]);
      c += v;
      cr += v;
    }
  }
  out[0] = c;
  out[1] = cr;
}

bool build_structure(sgo_handle* h) {
  int nf = (int)h->ivmap.size();
  h->voff.assign(nf, 0);
  h->vdim.assign(nf, 0);
  int off = 0;
  for (int i = 0; i < nf; ++i) {
    h->vdim[i] = h->ivmap[i].kind == 0 ? 3 : 2;
    h->voff[i] = off;
    off += h->vdim[i];
  }
  h->ndim = off;
  h->blockCols.assign(nf, {});
  long voffset = 0;
  auto alloc = [&](int r, int c) -> int {
    auto it = h->blockCols[c].find(r);
    if (it != h->blockCols[c].end()) return it->second;
    int o = (int)voffset;
    voffset += h->vdim[r] * h->vdim[c];
    h->blockCols[c][r] = o;
    return o;
  };
  h->diag_off.assign(nf, 0);
  for (int i = 0; i < nf; ++i) h->diag_off[i] = alloc(i, i);
  h->emap.assign(h->active_edges.size(), {});
  for (size_t k = 0; k < h->active_edges.size(); ++k) {
    auto& er = h->active_edges[k];
    int i0, i1;
    if (er.type == 0) { i0 = h->pose_hidx[h->pp[er.idx].i]; i1 = h->pose_hidx[h->pp[er.idx].j]; }
    else { i0 = h->pose_hidx[h->pl[er.idx].p]; i1 = h->lm_hidx[h->pl[er.idx].l]; }
    sgo_handle::EMap m{-1, -1, -1, false};
    if (i0 >= 0) m.off_ii = h->diag_off[i0];
    if (i1 >= 0) m.off_jj = h->diag_off[i1];
    if (i0 >= 0 && i1 >= 0) {
      m.transposed = i0 > i1;
      int r = std::min(i0, i1), c = std::max(i0, i1);
      m.off_ij = alloc(r, c);
    }
    h->emap[k] = m;
  }
  h->Hval.assign(voffset, 0.0);
  h->b.assign(h->ndim, 0.0);
  h->x.assign(h->ndim, 0.0);
  h->structure_built = true;
  h->solver_init = false;
  return true;
}

// Numeric Jacobian of EdgeSE2RhoTheta: BaseBinaryEdge::linearizeOplus default (A.3)
void pl_jac_numeric(sgo_handle* h, const EdgePL& e, bool pfree, bool lfree, double A[6], double B[4]) {
  const double delta = 1e-9;
  const double scalar = 1 / (2 * delta);
  double* pose = &h->pose[3 * e.p];
  double* line = &h->lm[2 * e.l];
  if (pfree) {
    double add[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d) {
      double bak[3] = {pose[0], pose[1], pose[2]};  // push
      add[d] = delta;
      pose_oplus(pose, add);
      double e1[2];
      pl_error(pose, line, e.z, e1);
      pose[0] = bak[0]; pose[1] = bak[1]; pose[2] = bak[2];  // pop
      add[d] = -delta;
      pose_oplus(pose, add);
      double e2[2];
      pl_error(pose, line, e.z, e2);
      pose[0] = bak[0]; pose[1] = bak[1]; pose[2] = bak[2];
      add[d] = 0.0;
      A[0 * 3 + d] = scalar * (e1[0] - e2[0]);
      A[1 * 3 + d] = scalar * (e1[1] - e2[1]);
    }
  }
  if (lfree) {
    double add[2] = {0, 0};
    for (int d = 0; d < 2; ++d) {
      double bak[2] = {line[0], line[1]};
      add[d] = delta;
      lm_oplus(line, add);
      double e1[2];
      pl_error(pose, line, e.z, e1);
      line[0] = bak[0]; line[1] = bak[1];
      add[d] = -delta;
      lm_oplus(line, add);
      double e2[2];
      pl_error(pose, line, e.z, e2);
      line[0] = bak[0]; line[1] = bak[1];
      add[d] = 0.0;
      B[0 * 2 + d] = scalar * (e1[0] - e2[0]);
      B[1 * 2 + d] = scalar * (e1[1] - e2[1]);
    }
  }
}
// Appendix B closed form (valid away from the switching sets)
void pl_jac_analytic(const double pose[3], const double line[2], double A[6], double B[4]) {
  double ca = std::cos(line[1]), sa = std::sin(line[1]);
  double q = line[0] - pose[0] * ca - pose[1] * sa;
  double s = q >= 0 ? 1.0 : -1.0;
  A[0] = s * ca; A[1] = s * sa; A[2] = 0;
  A[3] = 0; A[4] = 0; A[5] = 1;
  B[0] = -s; B[1] = -s * (pose[0] * sa - pose[1] * ca);
  B[2] = 0; B[3] = -1;
}

// column-major accessors for an (nr x nc) block at Hval+off
inline double& blk(std::vector<double>& H, int off, int nr, int r, int c) { return H[off + c * nr + r]; }

// BlockSolver::buildSystem (A.5) + BaseBinaryEdge::constructQuadraticForm (A.4)
void build_system(sgo_handle* h, int jac_mode) {
  std::fill(h->Hval.begin(), h->Hval.end(), 0.0);
  std::fill(h->b.begin(), h->b.end(), 0.0);
  for (size_t k = 0; k < h->active_edges.size(); ++k) {
    auto& er = h->active_edges[k];
    const auto& m = h->emap[k];
    if (er.type == 0) {
      const EdgePP& e = h->pp[er.idx];
      bool ifree = m.off_ii >= 0, jfree = m.off_jj >= 0;
      if (!ifree && !jfree) continue;
      // EdgeSE2::linearizeOplus (analytic, A.2)
      const double* pi = &h->pose[3 * e.i];
      const double* pj = &h->pose[3 * e.j];
      double thetai = pi[2];
      double dx = pj[0] - pi[0], dy = pj[1] - pi[1];
      double si = std::sin(thetai), ci = std::cos(thetai);
      double Ai[9] = {-ci, -si, -si * dx + ci * dy, si, -ci, -ci * dx - si * dy, 0, 0, -1};
      double Bj[9] = {ci, si, 0, -si, ci, 0, 0, 0, 1};
      double cz = std::cos(e.zinv.th), sz = std::sin(e.zinv.th);
      double Z[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
      double A[9], B[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double sa = 0, sb = 0;
          for (int t = 0; t < 3; ++t) { sa += Z[3 * r + t] * Ai[3 * t + c]; sb += Z[3 * r + t] * Bj[3 * t + c]; }
          A[3 * r + c] = sa;
          B[3 * r + c] = sb;
        }
      std::memcpy(&h->pp_A[9 * er.idx], A, sizeof A);
      std::memcpy(&h->pp_B[9 * er.idx], B, sizeof B);
      const double* err = &h->pp_err[3 * er.idx];
      double om[9];
      std::memcpy(om, e.info, sizeof om);
      double omega_r[3];
      for (int r = 0; r < 3; ++r) omega_r[r] = -(om[3 * r] * err[0] + om[3 * r + 1] * err[1] + om[3 * r + 2] * err[2]);
      if (e.phi > 0) {
        double rho[3];
        dcs(e.phi, pp_chi2(e, err), rho);
        for (int t = 0; t < 9; ++t) om[t] *= rho[1];   // robustInformation: rho[1]*Omega (2nd-order term commented out upstream)
        for (int r = 0; r < 3; ++r) omega_r[r] *= rho[1];
      }
      if (ifree) {
        double AtO[9];  // A^T * Omega
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int t = 0; t < 3; ++t) s += A[3 * t + r] * om[3 * t + c];
            AtO[3 * r + c] = s;
          }
        int hi = h->pose_hidx[e.i];
        for (int r = 0; r < 3; ++r) {
          double s = 0;
          for (int t = 0; t < 3; ++t) s += A[3 * t + r] * omega_r[t];
          h->b[h->voff[hi] + r] += s;
        }
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int t = 0; t < 3; ++t) s += AtO[3 * r + t] * A[3 * t + c];
            blk(h->Hval, m.off_ii, 3, r, c) += s;
          }
        if (jfree) {
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
              double s = 0;
              for (int t = 0; t < 3; ++t) s += AtO[3 * r + t] * B[3 * t + c];
              if (m.transposed) blk(h->Hval, m.off_ij, 3, c, r) += s;  // _hessianTransposed += B^T AtO^T
              else blk(h->Hval, m.off_ij, 3, r, c) += s;
            }
        }
      }
      if (jfree) {
        int hj = h->pose_hidx[e.j];
        for (int r = 0; r < 3; ++r) {
          double s = 0;
          for (int t = 0; t < 3; ++t) s += B[3 * t + r] * omega_r[t];
          h->b[h->voff[hj] + r] += s;
        }
        double BtO[9];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int t = 0; t < 3; ++t) s += B[3 * t + r] * om[3 * t + c];
            BtO[3 * r + c] = s;
          }
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int t = 0; t < 3; ++t) s += BtO[3 * r + t] * B[3 * t + c];
            blk(h->Hval, m.off_jj, 3, r, c) += s;
          }
      }
    } else {
      const EdgePL& e = h->pl[er.idx];
      bool pfree = m.off_ii >= 0, lfree = m.off_jj >= 0;
      if (!pfree && !lfree) continue;
      double* A = &h->pl_A[6 * er.idx];
      double* B = &h->pl_B[4 * er.idx];
      if (jac_mode == SGO_JAC_G2O_NUMERIC) pl_jac_numeric(h, e, pfree, lfree, A, B);
      else pl_jac_analytic(&h->pose[3 * e.p], &h->lm[2 * e.l], A, B);
      const double* err = &h->pl_err[2 * er.idx];
      const double* om = e.info;
      double omega_r[2] = {-(om[0] * err[0] + om[1] * err[1]), -(om[2] * err[0] + om[3] * err[1])};
      if (pfree) {
        double AtO[6];  // 3x2
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 2; ++c) AtO[2 * r + c] = A[0 * 3 + r] * om[0 * 2 + c] + A[1 * 3 + r] * om[1 * 2 + c];
        int hi = h->pose_hidx[e.p];
        for (int r = 0; r < 3; ++r) h->b[h->voff[hi] + r] += A[r] * omega_r[0] + A[3 + r] * omega_r[1];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) blk(h->Hval, m.off_ii, 3, r, c) += AtO[2 * r] * A[c] + AtO[2 * r + 1] * A[3 + c];
        if (lfree) {
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 2; ++c) {
              double s = AtO[2 * r] * B[c] + AtO[2 * r + 1] * B[2 + c];
              if (m.transposed) blk(h->Hval, m.off_ij, 2, c, r) += s;  // block is (lm row, pose col): 2x3
              else blk(h->Hval, m.off_ij, 3, r, c) += s;               // block is (pose row, lm col): 3x2
            }
        }
      }
      if (lfree) {
        int hl = h->lm_hidx[e.l];
        for (int r = 0; r < 2; ++r) h->b[h->voff[hl] + r] += B[r] * omega_r[0] + B[2 + r] * omega_r[1];
        double BtO[4];
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 2; ++c) BtO[2 * r + c] = B[r] * om[c] + B[2 + r] * om[2 + c];
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 2; ++c) blk(h->Hval, m.off_jj, 2, r, c) += BtO[2 * r] * B[c] + BtO[2 * r + 1] * B[2 + c];
      }
    }
  }
}

// LinearSolverEigen::solve (A.7): scalar upper-triangular CCS, ordering + symbolic once per optimize(), numeric every call
void solver_init_pattern(sgo_handle* h) {
  int n = h->ndim;
  int nf = (int)h->ivmap.size();
  h->Ap.assign(n + 1, 0);
  h->Ai.clear();
  h->a_src.clear();
  for (int c = 0; c < nf; ++c) {
    int nc = h->vdim[c];
    for (int cc = 0; cc < nc; ++cc) {
      int gcol = h->voff[c] + cc;
      for (auto& kv : h->blockCols[c]) {
        int r = kv.first, nr = h->vdim[r];
        for (int rr = 0; rr < nr; ++rr) {
          int grow = h->voff[r] + rr;
          if (grow > gcol) continue;
          h->Ai.push_back(grow);
          h->a_src.push_back(kv.second + cc * nr + rr);
        }
      }
      h->Ap[gcol + 1] = (int)h->Ai.size();
    }
  }
  h->Ax.assign(h->Ai.size(), 0.0);
  // ordering on the vertex (block) graph, expanded to scalars
  std::vector<std::vector<int>> adj(nf);
  for (int c = 0; c < nf; ++c)
    for (auto& kv : h->blockCols[c])
      if (kv.first != c) { adj[c].push_back(kv.first); adj[kv.first].push_back(c); }
  std::vector<int> border = min_degree_order(nf, adj);
  std::vector<int> perm;
  perm.reserve(n);
  for (int v : border)
    for (int d = 0; d < h->vdim[v]; ++d) perm.push_back(h->voff[v] + d);
  h->chol.analyze(n, h->Ap, h->Ai, perm);
  h->solver_init = true;
}

bool linear_solve(sgo_handle* h, double lambda) {
  if (!h->solver_init) solver_init_pattern(h);
  for (size_t p = 0; p < h->Ax.size(); ++p) h->Ax[p] = h->Hval[h->a_src[p]];
  if (lambda != 0.0) {
    // BlockSolver::setLambda: every scalar diagonal entry += lambda (restored afterwards; here applied to the CCS copy only)
    for (int c = 0; c < h->ndim; ++c) h->Ax[h->Ap[c + 1] - 1] += lambda;
  }
  if (!h->chol.factorize(h->Ax)) return false;
  h->chol.solve(h->b.data(), h->x.data());
  return true;
}

// SparseOptimizer::update (A.6)
void apply_update(sgo_handle* h, const double* x) {
  for (size_t i = 0; i < h->ivmap.size(); ++i) {
    auto& v = h->ivmap[i];
    if (v.kind == 0) pose_oplus(&h->pose[3 * v.idx], x + h->voff[i]);
    else lm_oplus(&h->lm[2 * v.idx], x + h->voff[i]);
  }
}

}  // namespace

extern "C" {

sgo_handle* sgo_create(void) { return new sgo_handle(); }
void sgo_destroy(sgo_handle* h) { delete h; }

int sgo_set_graph(sgo_handle* h, const sgo_graph* g) {
  int P = g->n_poses, L = g->n_landmarks;
  h->pose_id.resize(P);
  h->lm_id.resize(L);
  for (int i = 0; i < P; ++i) h->pose_id[i] = g->pose_id ? g->pose_id[i] : i;
  for (int i = 0; i < L; ++i) h->lm_id[i] = g->lm_id ? g->lm_id[i] : 10000000 + i;
  h->pose.assign(g->pose_est, g->pose_est + 3 * (size_t)P);
  h->lm.assign(g->lm_est, g->lm_est + 2 * (size_t)L);
  h->pose_fixed.assign(P, 0);
  h->lm_fixed.assign(L, 0);
  if (g->pose_fixed) h->pose_fixed.assign(g->pose_fixed, g->pose_fixed + P);
  if (g->lm_fixed) h->lm_fixed.assign(g->lm_fixed, g->lm_fixed + L);
  h->pp.resize(g->n_pp);
  for (int k = 0; k < g->n_pp; ++k) {
    EdgePP& e = h->pp[k];
    e.i = g->pp_i[k];
    e.j = g->pp_j[k];
    if (e.i < 0 || e.i >= P || e.j < 0 || e.j >= P) return -1;
    for (int t = 0; t < 3; ++t) e.z[t] = g->pp_z[3 * k + t];
    e.zinv = se2_inv(SE2{e.z[0], e.z[1], e.z[2]});
    const double* u = &g->pp_info[6 * (size_t)k];
    double full[9] = {u[0], u[1], u[2], u[1], u[3], u[4], u[2], u[4], u[5]};
    std::memcpy(e.info, full, sizeof full);
    e.phi = g->pp_phi ? g->pp_phi[k] : 0.0;
    e.seq = g->pp_seq ? g->pp_seq[k] : k;
  }
  h->pl.resize(g->n_pl);
  for (int k = 0; k < g->n_pl; ++k) {
    EdgePL& e = h->pl[k];
    e.p = g->pl_pose[k];
    e.l = g->pl_lm[k];
    if (e.p < 0 || e.p >= P || e.l < 0 || e.l >= L) return -1;
    e.z[0] = g->pl_z[2 * k];
    e.z[1] = g->pl_z[2 * k + 1];
    const double* u = &g->pl_info[3 * (size_t)k];
    e.info[0] = u[0]; e.info[1] = u[1]; e.info[2] = u[1]; e.info[3] = u[2];
    e.seq = g->pl_seq ? g->pl_seq[k] : (int64_t)g->n_pp + k;
  }
  h->pp_err.assign(3 * (size_t)g->n_pp, 0.0);
  h->pp_A.assign(9 * (size_t)g->n_pp, 0.0);
  h->pp_B.assign(9 * (size_t)g->n_pp, 0.0);
  h->pl_err.assign(2 * (size_t)g->n_pl, 0.0);
  h->pl_A.assign(6 * (size_t)g->n_pl, 0.0);
  h->pl_B.assign(4 * (size_t)g->n_pl, 0.0);
  h->initialized = false;
  h->structure_built = false;
  return 0;
}

// SparseOptimizer::initializeOptimization + buildIndexMapping (A.5)
int sgo_initialize(sgo_handle* h) {
  int P = (int)h->pose_id.size(), L = (int)h->lm_id.size();
  h->initialized = false;
  h->structure_built = false;
  if (h->pp.empty() && h->pl.empty()) {
    std::fprintf(stderr, "sgo: Attempt to initialize an empty graph\n");
    return 0;
  }
  h->pose_active.assign(P, 0);
  h->lm_active.assign(L, 0);
  h->active_edges.clear();
  for (size_t k = 0; k < h->pp.size(); ++k) {
    auto& e = h->pp[k];
    if (h->pose_fixed[e.i] && h->pose_fixed[e.j]) continue;  // allVerticesFixed
    h->active_edges.push_back({0, (int)k, e.seq});
    h->pose_active[e.i] = 1;
    h->pose_active[e.j] = 1;
  }
  for (size_t k = 0; k < h->pl.size(); ++k) {
    auto& e = h->pl[k];
    if (h->pose_fixed[e.p] && h->lm_fixed[e.l]) continue;
    h->active_edges.push_back({1, (int)k, e.seq});
    h->pose_active[e.p] = 1;
    h->lm_active[e.l] = 1;
  }
  std::stable_sort(h->active_edges.begin(), h->active_edges.end(),
                   [](const sgo_handle::ERef& a, const sgo_handle::ERef& b) { return a.seq < b.seq; });
  // active vertices sorted by id
  struct AV { int id, kind, idx; };
  std::vector<AV> av;
  for (int i = 0; i < P; ++i) if (h->pose_active[i]) av.push_back({h->pose_id[i], 0, i});
  for (int i = 0; i < L; ++i) if (h->lm_active[i]) av.push_back({h->lm_id[i], 1, i});
  std::stable_sort(av.begin(), av.end(), [](const AV& a, const AV& b) { return a.id < b.id; });
  h->pose_hidx.assign(P, -1);
  h->lm_hidx.assign(L, -1);
  h->ivmap.clear();
  if (av.empty()) return 0;
  for (auto& v : av) {  // nothing is ever marginalised in the reference => single pass (k = 0)
    bool fixed = v.kind == 0 ? h->pose_fixed[v.idx] : h->lm_fixed[v.idx];
    if (fixed) continue;
    int hi = (int)h->ivmap.size();
    if (v.kind == 0) h->pose_hidx[v.idx] = hi; else h->lm_hidx[v.idx] = hi;
    h->ivmap.push_back({v.kind, v.idx});
  }
  h->initialized = true;
  build_structure(h);
  return 1;
}

int sgo_num_free(sgo_handle* h) { return (int)h->ivmap.size(); }
int sgo_scalar_dim(sgo_handle* h) { return h->ndim; }
int sgo_num_blocks(sgo_handle* h) {
  int n = 0;
  for (auto& c : h->blockCols) n += (int)c.size();
  return n;
}
int64_t sgo_block_values_size(sgo_handle* h) { return (int64_t)h->Hval.size(); }
void sgo_get_order(sgo_handle* h, int32_t* kind, int32_t* index, int32_t* offset) {
  for (size_t i = 0; i < h->ivmap.size(); ++i) {
    kind[i] = h->ivmap[i].kind;
    index[i] = h->ivmap[i].idx;
    offset[i] = h->voff[i];
  }
}
void sgo_get_blocks(sgo_handle* h, int32_t* row, int32_t* col, int32_t* nr, int32_t* nc) {
  int k = 0;
  for (size_t c = 0; c < h->blockCols.size(); ++c)
    for (auto& kv : h->blockCols[c]) {
      row[k] = kv.first;
      col[k] = (int)c;
      nr[k] = h->vdim[kv.first];
      nc[k] = h->vdim[c];
      ++k;
    }
}
void sgo_get_hessian_index(sgo_handle* h, int32_t* ph, int32_t* lh) {
  std::copy(h->pose_hidx.begin(), h->pose_hidx.end(), ph);
  std::copy(h->lm_hidx.begin(), h->lm_hidx.end(), lh);
}

int sgo_linearize(sgo_handle* h, int jac_mode, double* pp_err, double* pp_A, double* pp_B, double* pl_err,
                  double* pl_A, double* pl_B, double* b, double* Hblocks, double* chi2) {
  if (!h->initialized) return -1;
  edge_errors(h);
  build_system(h, jac_mode);
  if (pp_err) std::copy(h->pp_err.begin(), h->pp_err.end(), pp_err);
  if (pp_A) std::copy(h->pp_A.begin(), h->pp_A.end(), pp_A);
  if (pp_B) std::copy(h->pp_B.begin(), h->pp_B.end(), pp_B);
  if (pl_err) std::copy(h->pl_err.begin(), h->pl_err.end(), pl_err);
  if (pl_A) std::copy(h->pl_A.begin(), h->pl_A.end(), pl_A);
  if (pl_B) std::copy(h->pl_B.begin(), h->pl_B.end(), pl_B);
  if (b) std::copy(h->b.begin(), h->b.end(), b);
  if (Hblocks) {
    // concatenated in column-major / row-ascending block order
    size_t o = 0;
    for (size_t c = 0; c < h->blockCols.size(); ++c)
      for (auto& kv : h->blockCols[c]) {
        int sz = h->vdim[kv.first] * h->vdim[c];
        std::copy(h->Hval.begin() + kv.second, h->Hval.begin() + kv.second + sz, Hblocks + o);
        o += sz;
      }
  }
  if (chi2) chi2_sums(h, chi2);
  return 0;
}

int sgo_solve_once(sgo_handle* h, int jac_mode, double lambda, double* x) {
  if (!h->initialized) return -1;
  edge_errors(h);
  build_system(h, jac_mode);
  h->solver_init = false;
  bool ok = linear_solve(h, lambda);
  if (x) std::copy(h->x.begin(), h->x.end(), x);
  return ok ? 0 : 1;
}

void sgo_chi2(sgo_handle* h, double* chi2) {
  edge_errors(h);
  chi2_sums(h, chi2);
}

// SparseOptimizer::optimize (A.6) with OptimizationAlgorithmLevenberg / GaussNewton ::solve
int sgo_optimize(sgo_handle* h, int algo, int iters, int jac_mode, sgo_iter_stat* stats) {
  if (!h->initialized || h->ivmap.empty()) {
    std::fprintf(stderr, "sgo: 0 vertices to optimize, maybe forgot to call initializeOptimization()\n");
    return -1;
  }
  double t_start = now_s();
  double t_lin = 0, t_solve = 0;
  // algorithm->init(online=false): solver.init -> linearSolver.init() forces a new symbolic analysis
  h->solver_init = false;
  int done = 0;
  int result = 1;
  bool ok = true;
  for (int it = 0; it < iters && ok; ++it) {
    if (it == 0) build_structure(h);
    double chi[2];
    double t0 = now_s();
    edge_errors(h);
    chi2_sums(h, chi);
    double currentChi = chi[1];
    double chi_before = currentChi;
    build_system(h, jac_mode);
    t_lin += now_s() - t0;
    int trials = 0;
    double rho = 0;
    if (algo == SGO_ALGO_GN) {
      t0 = now_s();
      bool ok2 = linear_solve(h, 0.0);
      t_solve += now_s() - t0;
      apply_update(h, h->x.data());
      result = ok2 ? 1 : -1;
      trials = 1;
      // reported chi2 for GN: value before the update (g2o prints the same)
    } else {
      if (it == 0) {
        // computeLambdaInit: tau * max |diag H|
        double maxDiagonal = 0;
        for (size_t i = 0; i < h->ivmap.size(); ++i) {
          int d = h->vdim[i];
          for (int j = 0; j < d; ++j) maxDiagonal = std::max(std::fabs(h->Hval[h->diag_off[i] + j * d + j]), maxDiagonal);
        }
        h->lambda = 1e-5 * maxDiagonal;
        h->ni = 2;
      }
      double tempChi = currentChi;
      do {
        std::vector<double> pose_bak(h->pose), lm_bak(h->lm);  // optimizer.push()
        t0 = now_s();
        bool ok2 = linear_solve(h, h->lambda);
        t_solve += now_s() - t0;
        apply_update(h, h->x.data());
        t0 = now_s();
        edge_errors(h);
        chi2_sums(h, chi);
        t_lin += now_s() - t0;
        tempChi = chi[1];
        if (!ok2) tempChi = DBL_MAX;
        rho = currentChi - tempChi;
        double scale = 0;  // computeScale
        for (int j = 0; j < h->ndim; ++j) scale += h->x[j] * (h->lambda * h->x[j] + h->b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          double scaleFactor = std::max(1. / 3., alpha);
          h->lambda *= scaleFactor;
          h->ni = 2;
          currentChi = tempChi;
          // discardTop
        } else {
          h->lambda *= h->ni;
          h->ni *= 2;
          h->pose.swap(pose_bak);  // pop
          h->lm.swap(lm_bak);
          if (!std::isfinite(h->lambda)) break;
        }
        trials++;
      } while (rho < 0 && trials < 10);
      if (trials == 10 || rho == 0 || !std::isfinite(h->lambda)) result = 2; else result = 1;
    }
    if (stats) {
      stats[it].iteration = it;
      stats[it].trials = trials;
      stats[it].result = result;
      stats[it].pad = 0;
      stats[it].chi2 = currentChi;
      stats[it].lambda = h->lambda;
      stats[it].rho = rho;
      stats[it].chi2_before = chi_before;
    }
    ok = (result == 1);
    ++done;
  }
  h->prof[0] = (double)h->chol.Li.size();
  h->prof[1] = t_lin;
  h->prof[2] = t_solve;
  h->prof[3] = now_s() - t_start;
  if (result == -1) return 0;
  return done;
}

void sgo_get_estimates(sgo_handle* h, double* pose_est, double* lm_est) {
  if (pose_est) std::copy(h->pose.begin(), h->pose.end(), pose_est);
  if (lm_est) std::copy(h->lm.begin(), h->lm.end(), lm_est);
}
void sgo_set_estimates(sgo_handle* h, const double* pose_est, const double* lm_est) {
  if (pose_est) std::copy(pose_est, pose_est + h->pose.size(), h->pose.begin());
  if (lm_est) std::copy(lm_est, lm_est + h->lm.size(), h->lm.begin());
}
void sgo_last_profile(sgo_handle* h, double* out) { std::memcpy(out, h->prof, sizeof h->prof); }

}  // extern "C"
