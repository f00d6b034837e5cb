// sgb_structure.cpp -- see sgb_structure.h. Host-only, no CUDA.
#include "sgb_structure.h"

#include <algorithm>
#include <numeric>

namespace sgb {

namespace {

struct ERef {
  int64_t seq;
  int32_t type;  // 0 pose-pose, 1 pose-line
  int32_t idx;   // index in the caller's arrays
};

// rows -> sorted unique column lists -> SELL-32 with per-row entry lookup
struct RowLists {
  std::vector<int32_t> ptr, col;  // CSR, columns ascending
  int find(int r, int c) const {
    auto b = col.begin() + ptr[r], e = col.begin() + ptr[r + 1];
    auto it = std::lower_bound(b, e, c);
    return (it != e && *it == c) ? (int)(it - b) : -1;
  }
};

RowLists make_rows(int rows, std::vector<std::pair<int32_t, int32_t>>& rc) {
  std::sort(rc.begin(), rc.end());
  rc.erase(std::unique(rc.begin(), rc.end()), rc.end());
  RowLists L;
  L.ptr.assign(rows + 1, 0);
  for (auto& p : rc) L.ptr[p.first + 1]++;
  for (int r = 0; r < rows; ++r) L.ptr[r + 1] += L.ptr[r];
  L.col.resize(rc.size());
  for (size_t k = 0; k < rc.size(); ++k) L.col[k] = rc[k].second;  // already grouped by row, ascending
  return L;
}

// SELL from row lists; row_of_sell[r] gives the logical row stored at SELL row r (identity when empty)
void make_sell(const RowLists& L, int rows, const std::vector<int32_t>* row_of_sell, HostSell& S) {
  S.rows = rows;
  S.nslices = (rows + 31) / 32;
  S.sbase.assign(S.nslices + 1, 0);
  for (int s = 0; s < S.nslices; ++s) {
    int w = 0;
    for (int lane = 0; lane < 32; ++lane) {
      int r = s * 32 + lane;
      if (r >= rows) break;
      int lr = row_of_sell ? (*row_of_sell)[r] : r;
      w = std::max(w, L.ptr[lr + 1] - L.ptr[lr]);
    }
    S.sbase[s + 1] = S.sbase[s] + w * 32;
  }
  S.col.assign((size_t)S.sbase[S.nslices], -1);
  for (int r = 0; r < rows; ++r) {
    int lr = row_of_sell ? (*row_of_sell)[r] : r;
    int s = r >> 5, lane = r & 31;
    for (int k = 0; k < L.ptr[lr + 1] - L.ptr[lr]; ++k) S.col[(size_t)S.sbase[s] + k * 32 + lane] = L.col[L.ptr[lr] + k];
  }
}

inline int sell_entry(const HostSell& S, int sell_row, int k) { return S.sbase[sell_row >> 5] + k * 32 + (sell_row & 31); }

}  // namespace

sgb_status build_structure(const sgb_graph_soa& g, Structure& S, std::string& err) {
  S = Structure();
  const int P = g.n_poses, L = g.n_landmarks;
  if (P < 0 || L < 0 || g.n_pp < 0 || g.n_pl < 0) { err = "negative size"; return SGB_ERR_INVALID; }
  if ((P > 0 && !g.pose_est) || (L > 0 && !g.lm_est)) { err = "missing estimates"; return SGB_ERR_INVALID; }
  if ((g.n_pp > 0 && (!g.pp_i || !g.pp_j || !g.pp_z || !g.pp_info)) ||
      (g.n_pl > 0 && (!g.pl_pose || !g.pl_lm || !g.pl_z || !g.pl_info))) {
    err = "missing edge arrays";
    return SGB_ERR_INVALID;
  }
  S.P_all = P;
  S.L_all = L;
  auto pfixed = [&](int i) { return g.pose_fixed ? g.pose_fixed[i] != 0 : false; };
  auto lfixed = [&](int i) { return g.lm_fixed ? g.lm_fixed[i] != 0 : false; };
  auto pid = [&](int i) { return g.pose_id ? g.pose_id[i] : i; };
  auto lid = [&](int i) { return g.lm_id ? g.lm_id[i] : 10000000 + i; };

  if (g.n_pp + g.n_pl == 0) { err = "Attempt to initialize an empty graph"; return SGB_ERR_NOT_INITIALIZED; }

  // ---- active edges (level 0, all vertices in the set, not all fixed), sorted by internal id
  std::vector<ERef> act;
  act.reserve((size_t)g.n_pp + g.n_pl);
  std::vector<char> pact(P, 0), lact(L, 0);
  for (int k = 0; k < g.n_pp; ++k) {
    int i = g.pp_i[k], j = g.pp_j[k];
    if (i < 0 || i >= P || j < 0 || j >= P || i == j) { err = "pose-pose edge with a bad vertex index"; return SGB_ERR_INVALID; }
    if (pfixed(i) && pfixed(j)) continue;
    act.push_back({g.pp_seq ? g.pp_seq[k] : (int64_t)k, 0, k});
    pact[i] = pact[j] = 1;
    if (g.pp_phi && g.pp_phi[k] > 0.0) S.has_robust = true;
  }
  for (int k = 0; k < g.n_pl; ++k) {
    int p = g.pl_pose[k], l = g.pl_lm[k];
    if (p < 0 || p >= P || l < 0 || l >= L) { err = "pose-line edge with a bad vertex index"; return SGB_ERR_INVALID; }
    if (pfixed(p) && lfixed(l)) continue;
    act.push_back({g.pl_seq ? g.pl_seq[k] : (int64_t)g.n_pp + k, 1, k});
    pact[p] = 1;
    lact[l] = 1;
  }
  std::stable_sort(act.begin(), act.end(), [](const ERef& a, const ERef& b) {
    if (a.seq != b.seq) return a.seq < b.seq;
    if (a.type != b.type) return a.type < b.type;
    return a.idx < b.idx;
  });

  // ---- index mapping: active vertices sorted by id; fixed -> -1; nothing is marginalised
  std::vector<int32_t> porder, lorder;
  for (int i = 0; i < P; ++i) if (pact[i] && !pfixed(i)) porder.push_back(i);
  for (int i = 0; i < L; ++i) if (lact[i] && !lfixed(i)) lorder.push_back(i);
  std::stable_sort(porder.begin(), porder.end(), [&](int a, int b) { return pid(a) < pid(b); });
  std::stable_sort(lorder.begin(), lorder.end(), [&](int a, int b) { return lid(a) < lid(b); });
  if (!porder.empty() && !lorder.empty() && pid(porder.back()) >= lid(lorder.front())) {
    err = "unsupported vertex ordering: every pose id must be smaller than every landmark id (reference: drone.h:22)";
    return SGB_ERR_UNSUPPORTED;
  }
  S.Pf = (int)porder.size();
  S.Lf = (int)lorder.size();
  if (S.Pf + S.Lf == 0) { err = "0 vertices to optimize"; return SGB_ERR_NOT_INITIALIZED; }
  S.pose_h.assign(P, -1);
  S.lm_h.assign(L, -1);
  S.pose_of_h = porder;
  S.lm_of_h = lorder;
  for (int h = 0; h < S.Pf; ++h) S.pose_h[porder[h]] = h;
  for (int h = 0; h < S.Lf; ++h) S.lm_h[lorder[h]] = h;
  S.dim = 3 * S.Pf + 2 * S.Lf;
  const int nfree = S.Pf + S.Lf;
  S.ord_kind.resize(nfree);
  S.ord_index.resize(nfree);
  S.ord_offset.resize(nfree);
  for (int h = 0; h < S.Pf; ++h) { S.ord_kind[h] = 0; S.ord_index[h] = porder[h]; S.ord_offset[h] = 3 * h; }
  for (int h = 0; h < S.Lf; ++h) { S.ord_kind[S.Pf + h] = 1; S.ord_index[S.Pf + h] = lorder[h]; S.ord_offset[S.Pf + h] = 3 * S.Pf + 2 * h; }

  // ---- per-type active edge arrays (insertion order)
  for (auto& e : act) (e.type == 0 ? S.pp_src : S.pl_src).push_back(e.idx);
  S.n_pp = (int)S.pp_src.size();
  S.n_pl = (int)S.pl_src.size();
  S.pp_i.resize(S.n_pp); S.pp_j.resize(S.n_pp); S.pp_hi.resize(S.n_pp); S.pp_hj.resize(S.n_pp);
  S.pp_e_ij.assign(S.n_pp, -1); S.pp_e_ji.assign(S.n_pp, -1); S.pp_dup.assign(S.n_pp, -1);
  for (int k = 0; k < S.n_pp; ++k) {
    int s = S.pp_src[k];
    S.pp_i[k] = g.pp_i[s]; S.pp_j[k] = g.pp_j[s];
    S.pp_hi[k] = S.pose_h[g.pp_i[s]]; S.pp_hj[k] = S.pose_h[g.pp_j[s]];
  }
  S.pl_p.resize(S.n_pl); S.pl_l.resize(S.n_pl); S.pl_hp.resize(S.n_pl); S.pl_hl.resize(S.n_pl);
  S.pl_e_pl.assign(S.n_pl, -1); S.pl_e_lp.assign(S.n_pl, -1); S.pl_dup.assign(S.n_pl, -1);
  for (int k = 0; k < S.n_pl; ++k) {
    int s = S.pl_src[k];
    S.pl_p[k] = g.pl_pose[s]; S.pl_l[k] = g.pl_lm[s];
    S.pl_hp[k] = S.pose_h[g.pl_pose[s]]; S.pl_hl[k] = S.lm_h[g.pl_lm[s]];
  }

  // ---- incidence lists in global insertion order
  {
    S.pinc_ptr.assign(S.Pf + 1, 0);
    S.linc_ptr.assign(S.Lf + 1, 0);
    for (int k = 0; k < S.n_pp; ++k) {
      if (S.pp_hi[k] >= 0) S.pinc_ptr[S.pp_hi[k] + 1]++;
      if (S.pp_hj[k] >= 0) S.pinc_ptr[S.pp_hj[k] + 1]++;
    }
    for (int k = 0; k < S.n_pl; ++k) {
      if (S.pl_hp[k] >= 0) S.pinc_ptr[S.pl_hp[k] + 1]++;
      if (S.pl_hl[k] >= 0) S.linc_ptr[S.pl_hl[k] + 1]++;
    }
    for (int r = 0; r < S.Pf; ++r) S.pinc_ptr[r + 1] += S.pinc_ptr[r];
    for (int r = 0; r < S.Lf; ++r) S.linc_ptr[r + 1] += S.linc_ptr[r];
    S.pinc.resize(S.pinc_ptr[S.Pf]);
    S.linc.resize(S.linc_ptr[S.Lf]);
    std::vector<int32_t> pp_pos(S.pinc_ptr.begin(), S.pinc_ptr.end() - 1), lp_pos(S.linc_ptr.begin(), S.linc_ptr.end() - 1);
    int kpp = 0, kpl = 0;
    for (auto& e : act) {  // merged walk keeps the cross-type insertion order
      if (e.type == 0) {
        int k = kpp++;
        if (S.pp_hi[k] >= 0) S.pinc[pp_pos[S.pp_hi[k]]++] = (k << 2) | (0 << 1) | 0;
        if (S.pp_hj[k] >= 0) S.pinc[pp_pos[S.pp_hj[k]]++] = (k << 2) | (1 << 1) | 0;
      } else {
        int k = kpl++;
        if (S.pl_hp[k] >= 0) S.pinc[pp_pos[S.pl_hp[k]]++] = (k << 2) | (0 << 1) | 1;
        if (S.pl_hl[k] >= 0) S.linc[lp_pos[S.pl_hl[k]]++] = k;
      }
    }
  }
  if (S.n_pp >= (1 << 29) || S.n_pl >= (1 << 29)) { err = "too many edges for the packed incidence encoding"; return SGB_ERR_UNSUPPORTED; }

  // ---- pose-pose pairs: leader = first edge (insertion order) of an unordered free pair, others chained
  std::vector<std::pair<int32_t, int32_t>> rc;
  {
    std::vector<std::pair<uint64_t, int32_t>> keyed;
    for (int k = 0; k < S.n_pp; ++k) {
      int a = S.pp_hi[k], b = S.pp_hj[k];
      if (a < 0 || b < 0) continue;
      uint64_t key = ((uint64_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b);
      keyed.push_back({key, k});
    }
    std::stable_sort(keyed.begin(), keyed.end());  // ties keep insertion order (k ascending = seq ascending)
    rc.reserve(2 * keyed.size() + S.Pf);
    for (int h = 0; h < S.Pf; ++h) rc.push_back({h, h});
    for (size_t t = 0; t < keyed.size(); ++t) {
      bool leader = (t == 0) || keyed[t].first != keyed[t - 1].first;
      if (leader) {
        int a = (int)(keyed[t].first >> 32), b = (int)(keyed[t].first & 0xffffffffu);
        rc.push_back({a, b});
        rc.push_back({b, a});
        S.n_pairs_pp++;
      } else {
        S.pp_dup[keyed[t - 1].second] = keyed[t].second;
      }
    }
    RowLists rows = make_rows(S.Pf, rc);
    make_sell(rows, S.Pf, nullptr, S.Hpp);
    S.hpp_diag.resize(S.Pf);
    for (int h = 0; h < S.Pf; ++h) S.hpp_diag[h] = sell_entry(S.Hpp, h, rows.find(h, h));
    for (size_t t = 0; t < keyed.size(); ++t) {
      bool leader = (t == 0) || keyed[t].first != keyed[t - 1].first;
      if (!leader) continue;
      int k = keyed[t].second;
      int a = S.pp_hi[k], b = S.pp_hj[k];
      S.pp_e_ij[k] = sell_entry(S.Hpp, a, rows.find(a, b));
      S.pp_e_ji[k] = sell_entry(S.Hpp, b, rows.find(b, a));
    }
    // reference block list, pose columns
    for (int c = 0; c < S.Pf; ++c) {
      for (int q = rows.ptr[c]; q < rows.ptr[c + 1]; ++q) {
        int r = rows.col[q];  // symmetric pattern: the rows of column c are the columns of row c
        if (r > c) break;
        S.blk_row.push_back(r); S.blk_col.push_back(c); S.blk_nr.push_back(3); S.blk_nc.push_back(3);
        S.blk_kind.push_back(0);
        // values of block (row r, col c) live in row r's SELL row
        S.blk_entry.push_back(sell_entry(S.Hpp, r, rows.find(r, c)));
      }
    }
  }
  // ---- pose-line pairs
  {
    std::vector<std::pair<uint64_t, int32_t>> keyed;
    for (int k = 0; k < S.n_pl; ++k) {
      int a = S.pl_hp[k], b = S.pl_hl[k];
      if (a < 0 || b < 0) continue;
      keyed.push_back({((uint64_t)a << 32) | (uint32_t)b, k});
    }
    std::stable_sort(keyed.begin(), keyed.end());
    std::vector<std::pair<int32_t, int32_t>> rc_pl, rc_lp;
    for (size_t t = 0; t < keyed.size(); ++t) {
      bool leader = (t == 0) || keyed[t].first != keyed[t - 1].first;
      if (leader) {
        int a = (int)(keyed[t].first >> 32), b = (int)(keyed[t].first & 0xffffffffu);
        rc_pl.push_back({a, b});
        rc_lp.push_back({b, a});
        S.n_pairs_pl++;
      } else {
        S.pl_dup[keyed[t - 1].second] = keyed[t].second;
      }
    }
    RowLists rows_pl = make_rows(S.Pf, rc_pl);
    RowLists rows_lp = make_rows(S.Lf, rc_lp);
    make_sell(rows_pl, S.Pf, nullptr, S.Hpl);
    // landmark rows sorted by descending observer count so that a 32-row slice pads little
    S.lp_row2h.resize(S.Lf);
    std::iota(S.lp_row2h.begin(), S.lp_row2h.end(), 0);
    std::stable_sort(S.lp_row2h.begin(), S.lp_row2h.end(), [&](int a, int b) {
      return rows_lp.ptr[a + 1] - rows_lp.ptr[a] > rows_lp.ptr[b + 1] - rows_lp.ptr[b];
    });
    S.lp_h2row.assign(S.Lf, 0);
    for (int r = 0; r < S.Lf; ++r) S.lp_h2row[S.lp_row2h[r]] = r;
    make_sell(rows_lp, S.Lf, &S.lp_row2h, S.Hlp);
    for (size_t t = 0; t < keyed.size(); ++t) {
      bool leader = (t == 0) || keyed[t].first != keyed[t - 1].first;
      if (!leader) continue;
      int k = keyed[t].second;
      int a = S.pl_hp[k], b = S.pl_hl[k];
      S.pl_e_pl[k] = sell_entry(S.Hpl, a, rows_pl.find(a, b));
      S.pl_e_lp[k] = sell_entry(S.Hlp, S.lp_h2row[b], rows_lp.find(b, a));
    }
    // reference block list, landmark columns: pose rows ascending, then the diagonal
    for (int hl = 0; hl < S.Lf; ++hl) {
      int c = S.Pf + hl;
      for (int q = rows_lp.ptr[hl]; q < rows_lp.ptr[hl + 1]; ++q) {
        int r = rows_lp.col[q];
        S.blk_row.push_back(r); S.blk_col.push_back(c); S.blk_nr.push_back(3); S.blk_nc.push_back(2);
        S.blk_kind.push_back(1);
        S.blk_entry.push_back(sell_entry(S.Hpl, r, rows_pl.find(r, hl)));
      }
      S.blk_row.push_back(c); S.blk_col.push_back(c); S.blk_nr.push_back(2); S.blk_nc.push_back(2);
      S.blk_kind.push_back(2);
      S.blk_entry.push_back(hl);
    }
  }
  S.block_values = 0;
  for (size_t k = 0; k < S.blk_row.size(); ++k) S.block_values += (int64_t)S.blk_nr[k] * S.blk_nc[k];
  return SGB_OK;
}

}  // namespace sgb
