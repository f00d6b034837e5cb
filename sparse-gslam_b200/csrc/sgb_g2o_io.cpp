// sgb_g2o_io.cpp -- graph fixture / wire format: g2o's text format for the types on the path.
//
//   VERTEX_SE2 id x y theta                              [g2o VertexSE2::read/write]
//   EDGE_SE2 i j dx dy dtheta I11 I12 I13 I22 I23 I33    [g2o EdgeSE2::read/write]
//   FIX id ...                                           [g2o SparseOptimizer::load]
//   VERTEX_RHOTHETA id rho theta
//   EDGE_SE2_RHOTHETA i l rho theta I11 I12 I22
//   ROBUST_KERNEL_DCS k delta                            (k = index of the EDGE_SE2 line, 0-based)
// The reference registers the tags VERTEX_RHOTHETA / EDGE_SE2_RHOTHETA with g2o's factory
// (vertex_rhotheta.cpp:43, edge_se2_rhotheta.cpp:24) but leaves their read()/write() empty
// (vertex_rhotheta.cpp:36-42, edge_se2_rhotheta.cpp:18-23), so it cannot store a graph; the field order used here is
// the estimate / measurement followed by the upper triangle of the information matrix, g2o's convention for every
// slam2d type. g2o does not serialise robust kernels; the extra ROBUST_KERNEL_DCS lines carry the per-edge DCS delta
// set at submap_loop_closer.cpp:41,57,283 (unknown tags are skipped by g2o's loader, and by this one).
// The line order of the two edge kinds is kept as the insertion rank (g2o internalId) of the edges.
// Host only, no CUDA.
#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sgb_capi.h"

struct sgb_graph_file {
  std::vector<int32_t> pose_id, lm_id, pp_i, pp_j, pl_pose, pl_lm;
  std::vector<uint8_t> pose_fixed, lm_fixed;
  std::vector<double> pose_est, lm_est, pp_z, pp_info, pp_phi, pl_z, pl_info;
  std::vector<int64_t> pp_seq, pl_seq;
};

namespace {
void set_err(char* buf, int32_t len, const std::string& msg) {
  if (buf && len > 0) std::snprintf(buf, (size_t)len, "%s", msg.c_str());
}
}  // namespace

extern "C" {

sgb_status sgb_g2o_load(const char* path, sgb_graph_file** out, char* errbuf, int32_t errlen) {
  if (!path || !out) return SGB_ERR_INVALID;
  *out = nullptr;
  std::ifstream is(path);
  if (!is) {
    set_err(errbuf, errlen, std::string("cannot open ") + path);
    return SGB_ERR_INVALID;
  }
  auto* f = new sgb_graph_file();
  std::unordered_map<int64_t, int32_t> pose_of, lm_of;
  // every entry that is resolved after the whole file has been read remembers its line: errors found there (an edge or
  // a FIX naming an unknown vertex, a robust kernel on an unknown edge) are reported with it
  struct RawPP { int64_t i, j, line; };
  struct RawPL { int64_t p, l, line; };
  struct RawFix { int64_t id, line; };
  struct RawDcs { int64_t k; double d; int64_t line; };
  std::vector<RawPP> rpp;
  std::vector<RawPL> rpl;
  std::vector<RawFix> fixed_ids;
  std::vector<RawDcs> dcs;
  std::string line, tag;
  int64_t seq = 0, lineno = 0;
  auto fail = [&](const std::string& what) {
    set_err(errbuf, errlen, std::string(path) + ":" + std::to_string(lineno) + ": " + what);
    delete f;
    return SGB_ERR_INVALID;
  };
  while (std::getline(is, line)) {
    ++lineno;
    std::istringstream ls(line);
    if (!(ls >> tag) || tag[0] == '#') continue;
    if (tag == "VERTEX_SE2") {
      int64_t id;
      double x, y, t;
      if (!(ls >> id >> x >> y >> t)) return fail("malformed VERTEX_SE2");
      if (pose_of.count(id) || lm_of.count(id)) return fail("duplicate vertex id");
      pose_of[id] = (int32_t)f->pose_id.size();
      f->pose_id.push_back((int32_t)id);
      f->pose_est.insert(f->pose_est.end(), {x, y, t});
    } else if (tag == "VERTEX_RHOTHETA") {
      int64_t id;
      double r, t;
      if (!(ls >> id >> r >> t)) return fail("malformed VERTEX_RHOTHETA");
      if (pose_of.count(id) || lm_of.count(id)) return fail("duplicate vertex id");
      lm_of[id] = (int32_t)f->lm_id.size();
      f->lm_id.push_back((int32_t)id);
      f->lm_est.insert(f->lm_est.end(), {r, t});
    } else if (tag == "EDGE_SE2") {
      RawPP e;
      double v[9];
      if (!(ls >> e.i >> e.j)) return fail("malformed EDGE_SE2");
      for (double& d : v)
        if (!(ls >> d)) return fail("malformed EDGE_SE2");
      e.line = lineno;
      rpp.push_back(e);
      f->pp_z.insert(f->pp_z.end(), v, v + 3);
      f->pp_info.insert(f->pp_info.end(), v + 3, v + 9);
      f->pp_seq.push_back(seq++);
    } else if (tag == "EDGE_SE2_RHOTHETA") {
      RawPL e;
      double v[5];
      if (!(ls >> e.p >> e.l)) return fail("malformed EDGE_SE2_RHOTHETA");
      for (double& d : v)
        if (!(ls >> d)) return fail("malformed EDGE_SE2_RHOTHETA");
      e.line = lineno;
      rpl.push_back(e);
      f->pl_z.insert(f->pl_z.end(), v, v + 2);
      f->pl_info.insert(f->pl_info.end(), v + 2, v + 5);
      f->pl_seq.push_back(seq++);
    } else if (tag == "FIX") {
      int64_t id;
      while (ls >> id) fixed_ids.push_back({id, lineno});
    } else if (tag == "ROBUST_KERNEL_DCS") {
      int64_t k;
      double d;
      if (!(ls >> k >> d)) return fail("malformed ROBUST_KERNEL_DCS");
      dcs.push_back({k, d, lineno});
    }  // anything else: not a type of this path, skipped like g2o skips unknown tags
  }
  f->pose_fixed.assign(f->pose_id.size(), 0);
  f->lm_fixed.assign(f->lm_id.size(), 0);
  for (auto& fx : fixed_ids) {
    const int64_t id = fx.id;
    lineno = fx.line;
    auto a = pose_of.find(id);
    if (a != pose_of.end()) { f->pose_fixed[a->second] = 1; continue; }
    auto b = lm_of.find(id);
    if (b != lm_of.end()) { f->lm_fixed[b->second] = 1; continue; }
    return fail("FIX of an unknown vertex " + std::to_string(id));
  }
  for (auto& e : rpp) {
    lineno = e.line;
    auto a = pose_of.find(e.i), b = pose_of.find(e.j);
    if (a == pose_of.end() || b == pose_of.end()) return fail("EDGE_SE2 references an unknown VERTEX_SE2");
    f->pp_i.push_back(a->second);
    f->pp_j.push_back(b->second);
  }
  for (auto& e : rpl) {
    lineno = e.line;
    auto a = pose_of.find(e.p);
    auto b = lm_of.find(e.l);
    if (a == pose_of.end() || b == lm_of.end()) return fail("EDGE_SE2_RHOTHETA references an unknown vertex");
    f->pl_pose.push_back(a->second);
    f->pl_lm.push_back(b->second);
  }
  f->pp_phi.assign(f->pp_i.size(), 0.0);
  for (auto& kd : dcs) {
    lineno = kd.line;
    if (kd.k < 0 || kd.k >= (int64_t)f->pp_phi.size()) return fail("ROBUST_KERNEL_DCS on an unknown edge");
    f->pp_phi[(size_t)kd.k] = kd.d;
  }
  *out = f;
  return SGB_OK;
}

void sgb_g2o_view(const sgb_graph_file* f, sgb_graph_soa* g) {
  if (!f || !g) return;
  std::memset(g, 0, sizeof *g);
  g->n_poses = (int32_t)f->pose_id.size();
  g->pose_id = f->pose_id.data();
  g->pose_est = f->pose_est.data();
  g->pose_fixed = f->pose_fixed.data();
  g->n_landmarks = (int32_t)f->lm_id.size();
  g->lm_id = f->lm_id.data();
  g->lm_est = f->lm_est.data();
  g->lm_fixed = f->lm_fixed.data();
  g->n_pp = (int32_t)f->pp_i.size();
  g->pp_i = f->pp_i.data();
  g->pp_j = f->pp_j.data();
  g->pp_z = f->pp_z.data();
  g->pp_info = f->pp_info.data();
  g->pp_phi = f->pp_phi.data();
  g->pp_seq = f->pp_seq.data();
  g->n_pl = (int32_t)f->pl_pose.size();
  g->pl_pose = f->pl_pose.data();
  g->pl_lm = f->pl_lm.data();
  g->pl_z = f->pl_z.data();
  g->pl_info = f->pl_info.data();
  g->pl_seq = f->pl_seq.data();
}

void sgb_g2o_free(sgb_graph_file* f) { delete f; }

sgb_status sgb_g2o_save(const char* path, const sgb_graph_soa* g) {
  if (!path || !g) return SGB_ERR_INVALID;
  FILE* fp = std::fopen(path, "w");
  if (!fp) return SGB_ERR_INVALID;
  auto pid = [&](int i) { return g->pose_id ? g->pose_id[i] : i; };
  auto lid = [&](int i) { return g->lm_id ? g->lm_id[i] : 10000000 + i; };
  for (int i = 0; i < g->n_poses; ++i)
    std::fprintf(fp, "VERTEX_SE2 %d %.17g %.17g %.17g\n", pid(i), g->pose_est[3 * (size_t)i], g->pose_est[3 * (size_t)i + 1],
                 g->pose_est[3 * (size_t)i + 2]);
  for (int i = 0; i < g->n_landmarks; ++i)
    std::fprintf(fp, "VERTEX_RHOTHETA %d %.17g %.17g\n", lid(i), g->lm_est[2 * (size_t)i], g->lm_est[2 * (size_t)i + 1]);
  for (int i = 0; i < g->n_poses; ++i)
    if (g->pose_fixed && g->pose_fixed[i]) std::fprintf(fp, "FIX %d\n", pid(i));
  for (int i = 0; i < g->n_landmarks; ++i)
    if (g->lm_fixed && g->lm_fixed[i]) std::fprintf(fp, "FIX %d\n", lid(i));
  // edges in insertion order: two-way merge of the two kinds by their sequence numbers (ties: pose-pose first)
  std::vector<int32_t> opp(g->n_pp), opl(g->n_pl);
  for (int k = 0; k < g->n_pp; ++k) opp[k] = k;
  for (int k = 0; k < g->n_pl; ++k) opl[k] = k;
  auto spp = [&](int k) { return g->pp_seq ? g->pp_seq[k] : (int64_t)k; };
  auto spl = [&](int k) { return g->pl_seq ? g->pl_seq[k] : (int64_t)g->n_pp + k; };
  std::stable_sort(opp.begin(), opp.end(), [&](int a, int b) { return spp(a) < spp(b); });
  std::stable_sort(opl.begin(), opl.end(), [&](int a, int b) { return spl(a) < spl(b); });
  std::vector<std::pair<int64_t, double>> dcs;
  size_t a = 0, b = 0;
  int64_t written_pp = 0;
  while (a < opp.size() || b < opl.size()) {
    bool take_pp = b >= opl.size() || (a < opp.size() && spp(opp[a]) <= spl(opl[b]));
    if (take_pp) {
      int k = opp[a++];
      const double* z = g->pp_z + 3 * (size_t)k;
      const double* w = g->pp_info + 6 * (size_t)k;
      std::fprintf(fp, "EDGE_SE2 %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", pid(g->pp_i[k]), pid(g->pp_j[k]),
                   z[0], z[1], z[2], w[0], w[1], w[2], w[3], w[4], w[5]);
      if (g->pp_phi && g->pp_phi[k] > 0.0) dcs.push_back({written_pp, g->pp_phi[k]});
      ++written_pp;
    } else {
      int k = opl[b++];
      const double* z = g->pl_z + 2 * (size_t)k;
      const double* w = g->pl_info + 3 * (size_t)k;
      std::fprintf(fp, "EDGE_SE2_RHOTHETA %d %d %.17g %.17g %.17g %.17g %.17g\n", pid(g->pl_pose[k]), lid(g->pl_lm[k]), z[0], z[1],
                   w[0], w[1], w[2]);
    }
  }
  for (auto& kd : dcs) std::fprintf(fp, "ROBUST_KERNEL_DCS %" PRId64 " %.17g\n", kd.first, kd.second);
  bool ok = std::ferror(fp) == 0;  // a failed write (full disk ...) is an error even when fclose succeeds
  ok = (std::fclose(fp) == 0) && ok;
  return ok ? SGB_OK : SGB_ERR_INVALID;
}

}  // extern "C"
