// sgb_structure.cpp -- see sgb_structure.h. Host-only, no CUDA.
//
// The whole symbolic phase is linear time: vertex pairs are grouped with counting sorts (bucket by the smaller
// Hessian index, tiny in-bucket sorts), row lists are filled in an order that leaves them sorted, and every
// scatter-map entry is known at fill time (no searches). The pose-pose part, the pose-line part and the
// incidence lists are independent and run on three host threads. A 1M-pose / 5.5M-edge graph takes ~0.2 s.
#include "sgb_structure.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>

namespace sgb {

namespace {

// CSR row lists, columns ascending
struct RowLists {
  std::vector<int32_t> ptr, col;
};

// SELL from row lists; row_of_sell[r] gives the logical row stored at SELL row r (identity when null)
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
  S.col.resize((size_t)S.sbase[S.nslices]);
  // slices are disjoint ranges of S.col: padding fill and column scatter on two host threads for the big matrices
  auto fill = [&](int s0, int s1) {
    std::fill(S.col.begin() + S.sbase[s0], S.col.begin() + S.sbase[s1], -1);
    const int r1 = std::min(rows, s1 * 32);
    for (int r = s0 * 32; r < r1; ++r) {
      int lr = row_of_sell ? (*row_of_sell)[r] : r;
      int s = r >> 5, lane = r & 31;
      int32_t* dst = S.col.data() + (size_t)S.sbase[s] + lane;
      const int32_t* src = L.col.data() + L.ptr[lr];
      for (int k = 0, n = L.ptr[lr + 1] - L.ptr[lr]; k < n; ++k) dst[(size_t)k * 32] = src[k];
    }
  };
  if (rows > 400000) {
    const int mid = S.nslices / 2;
    std::thread t(fill, 0, mid);
    fill(mid, S.nslices);
    t.join();
  } else {
    fill(0, S.nslices);
  }
}

inline int sell_entry(const HostSell& S, int sell_row, int k) { return S.sbase[sell_row >> 5] + k * 32 + (sell_row & 31); }

// Vertex pairs (a, b) of the edges 0..n-1 (a = bucket key in [0, nbuckets), either < 0: skipped), grouped so that equal
// pairs are adjacent and ordered by (a, b, edge index). O(n + nbuckets) plus tiny in-bucket sorts.
struct PairItem {
  int32_t b, k;
};
struct Grouped {
  std::vector<int32_t> start;  // [nbuckets + 1]
  std::vector<PairItem> items;
};
template <class KeyA, class KeyB>
void group_pairs(int n, int nbuckets, KeyA key_a, KeyB key_b, Grouped& G) {
  G.start.assign((size_t)nbuckets + 1, 0);
  for (int k = 0; k < n; ++k) {
    int a = key_a(k);
    if (a >= 0 && key_b(k) >= 0) G.start[a + 1]++;
  }
  for (int a = 0; a < nbuckets; ++a) G.start[a + 1] += G.start[a];
  G.items.resize((size_t)G.start[nbuckets]);
  std::vector<int32_t> pos(G.start.begin(), G.start.end() - 1);
  for (int k = 0; k < n; ++k) {
    int a = key_a(k), b = key_b(k);
    if (a >= 0 && b >= 0) G.items[(size_t)pos[a]++] = {b, k};
  }
  auto sort_buckets = [&](int a0, int a1) {
    for (int a = a0; a < a1; ++a) {
      PairItem* lo = G.items.data() + G.start[a];
      int m = G.start[a + 1] - G.start[a];
      if (m <= 1) continue;
      if (m <= 24) {  // stable insertion sort by b (k is already ascending inside a bucket)
        for (int i = 1; i < m; ++i) {
          PairItem v = lo[i];
          int j = i - 1;
          while (j >= 0 && lo[j].b > v.b) { lo[j + 1] = lo[j]; --j; }
          lo[j + 1] = v;
        }
      } else {
        std::stable_sort(lo, lo + m, [](const PairItem& x, const PairItem& y) { return x.b < y.b; });
      }
    }
  };
  if (n > 1000000) {  // buckets are independent: two more host threads on the big graphs
    const int c1 = nbuckets / 3, c2 = 2 * (nbuckets / 3);
    std::thread t1(sort_buckets, 0, c1), t2(sort_buckets, c1, c2);
    sort_buckets(c2, nbuckets);
    t1.join();
    t2.join();
  } else {
    sort_buckets(0, nbuckets);
  }
}

// order of the active edges of one type: caller indices sorted by (seq, index); identity when seq is absent or sorted
void order_by_seq(std::vector<int32_t>& idx, const int64_t* seq) {
  if (!seq) return;
  bool sorted = true;
  for (size_t t = 1; t < idx.size() && sorted; ++t) sorted = seq[idx[t - 1]] <= seq[idx[t]];
  if (!sorted) std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return seq[a] < seq[b]; });
}

}  // namespace

sgb_status build_structure(const sgb_graph_soa& g, Structure& S, std::string& err, int world, int rank) {
  static const bool prof = std::getenv("SGB_PROFILE") != nullptr;
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[build_structure] %-14s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
    tp0 = t;
  };
  S = Structure();
  lap("reset");
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
  if (g.n_pp >= (1 << 29) || g.n_pl >= (1 << 29)) { err = "too many edges for the packed incidence encoding"; return SGB_ERR_UNSUPPORTED; }

  // ---- active edges (level 0, all vertices in the set, not all fixed), sorted by internal id within each type
  const bool filtered = world > 1;
  S.filtered = filtered;
  S.fworld = filtered ? world : 1;
  S.frank = filtered ? rank : 0;
  std::vector<char> pact(P, 0), lact(L, 0);
  if (!filtered) {
    S.pp_src.reserve(g.n_pp);
    S.pl_src.reserve(g.n_pl);
  }
  // (filtered build, large graph: the scan only marks vertices -- every writer stores the same 1 -- so it is split over host
  // threads; every rank of a multi-GPU job runs this part on the WHOLE graph, it is what does not shrink with the rank count)
  const int nth_scan = (filtered && (size_t)g.n_pp + g.n_pl > 400000) ? 4 : 1;
  {
    std::vector<int> bad(nth_scan, 0), robust(nth_scan, 0);
    auto scan = [&](int t) {
      const int a0 = (int)((int64_t)g.n_pp * t / nth_scan), a1 = (int)((int64_t)g.n_pp * (t + 1) / nth_scan);
      const int b0 = (int)((int64_t)g.n_pl * t / nth_scan), b1 = (int)((int64_t)g.n_pl * (t + 1) / nth_scan);
      for (int k = a0; k < a1; ++k) {
        int i = g.pp_i[k], j = g.pp_j[k];
        if (i < 0 || i >= P || j < 0 || j >= P || i == j) { bad[t] = 1; return; }
        if (pfixed(i) && pfixed(j)) continue;
        if (!filtered) S.pp_src.push_back(k);
        pact[i] = pact[j] = 1;
        if (g.pp_phi && g.pp_phi[k] > 0.0) robust[t] = 1;
      }
      for (int k = b0; k < b1; ++k) {
        int p = g.pl_pose[k], l = g.pl_lm[k];
        if (p < 0 || p >= P || l < 0 || l >= L) { bad[t] = 2; return; }
        if (pfixed(p) && lfixed(l)) continue;
        if (!filtered) S.pl_src.push_back(k);
        pact[p] = 1;
        lact[l] = 1;
      }
    };
    if (nth_scan == 1) {
      scan(0);
    } else {
      std::vector<std::thread> th;
      for (int t = 1; t < nth_scan; ++t) th.emplace_back(scan, t);
      scan(0);
      for (auto& t : th) t.join();
    }
    for (int t = 0; t < nth_scan; ++t) {
      if (bad[t] == 1) { err = "pose-pose edge with a bad vertex index"; return SGB_ERR_INVALID; }
      if (bad[t] == 2) { err = "pose-line edge with a bad vertex index"; return SGB_ERR_INVALID; }
      if (robust[t]) S.has_robust = true;
    }
  }
  lap("active scan");
  // global insertion rank of an active edge (ties: pose-pose first, then caller index -- both orders are stable)
  // (the lambdas are used after S.pp_src / S.pl_src have been filled, further down for the filtered build)
  auto seq_pp = [&](int k) { return g.pp_seq ? g.pp_seq[S.pp_src[k]] : (int64_t)S.pp_src[k]; };
  auto seq_pl = [&](int k) { return g.pl_seq ? g.pl_seq[S.pl_src[k]] : (int64_t)g.n_pp + S.pl_src[k]; };

  // ---- index mapping: active vertices sorted by id; fixed -> -1; nothing is marginalised
  std::vector<int32_t>& porder = S.pose_of_h;
  std::vector<int32_t>& lorder = S.lm_of_h;
  for (int i = 0; i < P; ++i) if (pact[i] && !pfixed(i)) porder.push_back(i);
  for (int i = 0; i < L; ++i) if (lact[i] && !lfixed(i)) lorder.push_back(i);
  if (g.pose_id && !std::is_sorted(porder.begin(), porder.end(), [&](int a, int b) { return pid(a) < pid(b); }))
    std::stable_sort(porder.begin(), porder.end(), [&](int a, int b) { return pid(a) < pid(b); });
  if (g.lm_id && !std::is_sorted(lorder.begin(), lorder.end(), [&](int a, int b) { return lid(a) < lid(b); }))
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
  for (int h = 0; h < S.Pf; ++h) S.pose_h[porder[h]] = h;
  for (int h = 0; h < S.Lf; ++h) S.lm_h[lorder[h]] = h;
  S.dim = 3 * S.Pf + 2 * S.Lf;
  const int nfree = S.Pf + S.Lf;
  S.ord_kind.resize(nfree);
  S.ord_index.resize(nfree);
  S.ord_offset.resize(nfree);
  for (int h = 0; h < S.Pf; ++h) { S.ord_kind[h] = 0; S.ord_index[h] = porder[h]; S.ord_offset[h] = 3 * h; }
  for (int h = 0; h < S.Lf; ++h) { S.ord_kind[S.Pf + h] = 1; S.ord_index[S.Pf + h] = lorder[h]; S.ord_offset[S.Pf + h] = 3 * S.Pf + 2 * h; }

  lap("index mapping");
  if (filtered) {
    // ---- the edges this rank needs, selected with the GLOBAL Hessian indices just computed: pose-pose edges with an
    // endpoint in its row block; every edge of a landmark that one of its poses observes (the rank keeps a full copy of
    // that landmark's row: owned or ghost, sgb_partition.h). A landmark without any free observer belongs to rank 0.
    const int chunk = partition_chunk(S.Pf, world);
    auto mine = [&](int h) { return h >= 0 && h / chunk == rank; };
    std::vector<char> keep(L, 0), free_obs(L, 0);
    const int nth = nth_scan;
    auto run = [&](auto&& body) {  // body(t) on nth host threads
      if (nth == 1) { body(0); return; }
      std::vector<std::thread> th;
      for (int t = 1; t < nth; ++t) th.emplace_back(body, t);
      body(0);
      for (auto& t : th) t.join();
    };
    run([&](int t) {  // marks only: every writer stores the same 1
      const int b0 = (int)((int64_t)g.n_pl * t / nth), b1 = (int)((int64_t)g.n_pl * (t + 1) / nth);
      for (int k = b0; k < b1; ++k) {
        int p = g.pl_pose[k], l = g.pl_lm[k];
        if (pfixed(p) && lfixed(l)) continue;
        int hp = S.pose_h[p];
        if (hp >= 0) {
          free_obs[l] = 1;
          if (mine(hp)) keep[l] = 1;
        }
      }
    });
    if (rank == 0)
      for (int l = 0; l < L; ++l)
        if (lact[l] && !free_obs[l]) keep[l] = 1;
    std::vector<std::vector<int32_t>> sel_pp(nth), sel_pl(nth);  // per thread, caller order inside; concatenated in thread order
    run([&](int t) {
      const int a0 = (int)((int64_t)g.n_pp * t / nth), a1 = (int)((int64_t)g.n_pp * (t + 1) / nth);
      const int b0 = (int)((int64_t)g.n_pl * t / nth), b1 = (int)((int64_t)g.n_pl * (t + 1) / nth);
      for (int k = a0; k < a1; ++k) {
        int i = g.pp_i[k], j = g.pp_j[k];
        if (pfixed(i) && pfixed(j)) continue;
        if (mine(S.pose_h[i]) || mine(S.pose_h[j])) sel_pp[t].push_back(k);
      }
      for (int k = b0; k < b1; ++k) {
        int p = g.pl_pose[k], l = g.pl_lm[k];
        if (pfixed(p) && lfixed(l)) continue;
        if (mine(S.pose_h[p]) || (S.lm_h[l] >= 0 && keep[l])) sel_pl[t].push_back(k);
      }
    });
    for (int t = 0; t < nth; ++t) {
      S.pp_src.insert(S.pp_src.end(), sel_pp[t].begin(), sel_pp[t].end());
      S.pl_src.insert(S.pl_src.end(), sel_pl[t].begin(), sel_pl[t].end());
    }
    S.lm_present.assign(S.Lf, 0);
    for (int l = 0; l < L; ++l)
      if (S.lm_h[l] >= 0 && keep[l]) S.lm_present[S.lm_h[l]] = 1;
    lap("rank filter");
  }
  order_by_seq(S.pp_src, g.pp_seq);
  order_by_seq(S.pl_src, g.pl_seq);
  S.n_pp = (int)S.pp_src.size();
  S.n_pl = (int)S.pl_src.size();
  lap("edge order");
  // ---- per-type active edge arrays (insertion order); independent per edge: split over host threads when large
  S.pp_i.resize(S.n_pp); S.pp_j.resize(S.n_pp); S.pp_hi.resize(S.n_pp); S.pp_hj.resize(S.n_pp);
  S.pl_p.resize(S.n_pl); S.pl_l.resize(S.n_pl); S.pl_hp.resize(S.n_pl); S.pl_hl.resize(S.n_pl);
  auto fill_edges = [&](int a0, int a1, int b0, int b1) {
    for (int k = a0; k < a1; ++k) {
      int s = S.pp_src[k];
      S.pp_i[k] = g.pp_i[s]; S.pp_j[k] = g.pp_j[s];
      S.pp_hi[k] = S.pose_h[g.pp_i[s]]; S.pp_hj[k] = S.pose_h[g.pp_j[s]];
    }
    for (int k = b0; k < b1; ++k) {
      int s = S.pl_src[k];
      S.pl_p[k] = g.pl_pose[s]; S.pl_l[k] = g.pl_lm[s];
      S.pl_hp[k] = S.pose_h[g.pl_pose[s]]; S.pl_hl[k] = S.lm_h[g.pl_lm[s]];
    }
  };
  if ((size_t)S.n_pp + S.n_pl > 200000) {
    const int nth = 4;
    std::vector<std::thread> th;
    for (int t = 0; t < nth; ++t)
      th.emplace_back(fill_edges, (int)((int64_t)S.n_pp * t / nth), (int)((int64_t)S.n_pp * (t + 1) / nth),
                      (int)((int64_t)S.n_pl * t / nth), (int)((int64_t)S.n_pl * (t + 1) / nth));
    S.pp_e_ij.assign(S.n_pp, -1); S.pp_e_ji.assign(S.n_pp, -1); S.pp_dup.assign(S.n_pp, -1);
    S.pl_e_pl.assign(S.n_pl, -1); S.pl_k_lp.assign(S.n_pl, -1); S.pl_dup.assign(S.n_pl, -1);
    for (auto& t : th) t.join();
  } else {
    fill_edges(0, S.n_pp, 0, S.n_pl);
    S.pp_e_ij.assign(S.n_pp, -1); S.pp_e_ji.assign(S.n_pp, -1); S.pp_dup.assign(S.n_pp, -1);
    S.pl_e_pl.assign(S.n_pl, -1); S.pl_k_lp.assign(S.n_pl, -1); S.pl_dup.assign(S.n_pl, -1);
  }

  lap("edge arrays");
  // ---- incidence lists in global insertion order (two-way merge of the two edge types)
  auto build_incidence = [&]() {
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
    while (kpp < S.n_pp || kpl < S.n_pl) {
      bool take_pp = kpl >= S.n_pl || (kpp < S.n_pp && seq_pp(kpp) <= seq_pl(kpl));
      if (take_pp) {
        int k = kpp++;
        if (S.pp_hi[k] >= 0) S.pinc[pp_pos[S.pp_hi[k]]++] = (k << 2) | (0 << 1) | 0;
        if (S.pp_hj[k] >= 0) S.pinc[pp_pos[S.pp_hj[k]]++] = (k << 2) | (1 << 1) | 0;
      } else {
        int k = kpl++;
        if (S.pl_hp[k] >= 0) S.pinc[pp_pos[S.pl_hp[k]]++] = (k << 2) | (0 << 1) | 1;
        if (S.pl_hl[k] >= 0) S.linc[lp_pos[S.pl_hl[k]]++] = k;
      }
    }
  };

  // ---- pose-pose pairs: leader = first edge (insertion order) of an unordered free pair, others chained.
  // Row r of Hpp = [cols < r ascending | r | cols > r ascending]; positions are known when they are filled.
  auto build_pp = [&]() {
    Grouped G;
    group_pairs(S.n_pp, S.Pf, [&](int k) { int a = S.pp_hi[k], b = S.pp_hj[k]; return (a < 0 || b < 0) ? -1 : std::min(a, b); },
                [&](int k) { return std::max(S.pp_hi[k], S.pp_hj[k]); }, G);
    // leaders -> distinct pairs t = 0.. in (a, b) order
    std::vector<int32_t> pa, pb, pk;
    pa.reserve(G.items.size()); pb.reserve(G.items.size()); pk.reserve(G.items.size());
    std::vector<int32_t> lc(S.Pf, 0), uc(S.Pf, 0);
    for (int a = 0; a < S.Pf; ++a)
      for (int q = G.start[a]; q < G.start[a + 1]; ++q) {
        if (q > G.start[a] && G.items[q].b == G.items[q - 1].b) {
          S.pp_dup[G.items[q - 1].k] = G.items[q].k;
          continue;
        }
        pa.push_back(a); pb.push_back(G.items[q].b); pk.push_back(G.items[q].k);
        uc[a]++; lc[G.items[q].b]++;
      }
    const int np = (int)pa.size();
    S.n_pairs_pp = np;
    RowLists rows;
    rows.ptr.assign(S.Pf + 1, 0);
    for (int r = 0; r < S.Pf; ++r) rows.ptr[r + 1] = rows.ptr[r] + lc[r] + 1 + uc[r];
    rows.col.resize(rows.ptr[S.Pf]);
    std::vector<int32_t> lstart(S.Pf + 1, 0);
    for (int r = 0; r < S.Pf; ++r) lstart[r + 1] = lstart[r] + lc[r];
    std::vector<int32_t> low_pair(np), up_idx(np), low_idx(np), lpos(S.Pf, 0), upos(S.Pf, 0);
    for (int r = 0; r < S.Pf; ++r) rows.col[rows.ptr[r] + lc[r]] = r;
    for (int t = 0; t < np; ++t) {
      int a = pa[t], b = pb[t];
      low_idx[t] = lpos[b];
      low_pair[lstart[b] + lpos[b]] = t;
      rows.col[rows.ptr[b] + lpos[b]++] = a;
      up_idx[t] = lc[a] + 1 + upos[a];
      rows.col[rows.ptr[a] + lc[a] + 1 + upos[a]++] = b;
    }
    make_sell(rows, S.Pf, nullptr, S.Hpp);
    S.hpp_diag.resize(S.Pf);
    auto entries = [&](int t0, int t1, int h0, int h1) {  // independent per pair / per row
      for (int h = h0; h < h1; ++h) S.hpp_diag[h] = sell_entry(S.Hpp, h, lc[h]);
      for (int t = t0; t < t1; ++t) {
        int k = pk[t];
        int e_up = sell_entry(S.Hpp, pa[t], up_idx[t]);    // block (row a, col b), a < b
        int e_low = sell_entry(S.Hpp, pb[t], low_idx[t]);  // block (row b, col a)
        bool fwd = S.pp_hi[k] == pa[t];
        S.pp_e_ij[k] = fwd ? e_up : e_low;
        S.pp_e_ji[k] = fwd ? e_low : e_up;
      }
    };
    if (np > 400000) {
      std::thread t(entries, 0, np / 2, 0, S.Pf / 2);
      entries(np / 2, np, S.Pf / 2, S.Pf);
      t.join();
    } else {
      entries(0, np, 0, S.Pf);
    }
  };

  // ---- pose-line pairs
  auto build_pl = [&]() {
    Grouped G;
    group_pairs(S.n_pl, S.Pf, [&](int k) { return S.pl_hp[k]; }, [&](int k) { return S.pl_hl[k]; }, G);
    std::vector<int32_t> pa, pb, pk;
    pa.reserve(G.items.size()); pb.reserve(G.items.size()); pk.reserve(G.items.size());
    RowLists rows_pl, rows_lp;
    rows_pl.ptr.assign(S.Pf + 1, 0);
    rows_lp.ptr.assign(S.Lf + 1, 0);
    for (int a = 0; a < S.Pf; ++a)
      for (int q = G.start[a]; q < G.start[a + 1]; ++q) {
        if (q > G.start[a] && G.items[q].b == G.items[q - 1].b) {
          S.pl_dup[G.items[q - 1].k] = G.items[q].k;
          continue;
        }
        pa.push_back(a); pb.push_back(G.items[q].b); pk.push_back(G.items[q].k);
        rows_pl.ptr[a + 1]++; rows_lp.ptr[G.items[q].b + 1]++;
      }
    const int np = (int)pa.size();
    S.n_pairs_pl = np;
    for (int r = 0; r < S.Pf; ++r) rows_pl.ptr[r + 1] += rows_pl.ptr[r];
    for (int r = 0; r < S.Lf; ++r) rows_lp.ptr[r + 1] += rows_lp.ptr[r];
    rows_pl.col.resize(np);
    rows_lp.col.resize(np);
    std::vector<int32_t> idx_pl(np), pos_pl(S.Pf, 0);
    // pairs ascend by (pose, landmark): both row lists come out sorted. The landmark-major lists (scattered writes)
    // are filled on a second thread while this one builds the pose-major SELL.
    auto lm_major = [&]() {
      std::vector<int32_t> pos_lp(S.Lf, 0);
      for (int t = 0; t < np; ++t) {
        int a = pa[t], b = pb[t];
        S.pl_k_lp[pk[t]] = pos_lp[b];
        rows_lp.col[rows_lp.ptr[b] + pos_lp[b]++] = a;
      }
    };
    auto pose_major = [&]() {
      for (int t = 0; t < np; ++t) {
        int a = pa[t];
        idx_pl[t] = pos_pl[a];
        rows_pl.col[rows_pl.ptr[a] + pos_pl[a]++] = pb[t];
      }
      make_sell(rows_pl, S.Pf, nullptr, S.Hpl);
      for (int t = 0; t < np; ++t) S.pl_e_pl[pk[t]] = sell_entry(S.Hpl, pa[t], idx_pl[t]);
    };
    if (np > 1000000) {
      std::thread t(lm_major);
      pose_major();
      t.join();
    } else {
      lm_major();
      pose_major();
    }
    S.lp_ptr.swap(rows_lp.ptr);
    S.lp_col.swap(rows_lp.col);
  };

  if ((size_t)S.n_pp + S.n_pl > 200000) {
    auto timed = [&](const char* what, auto&& f) {
      auto t0 = std::chrono::steady_clock::now();
      f();
      if (prof) std::fprintf(stderr, "[build_structure]   %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    std::thread t1([&] { timed("incidence", build_incidence); }), t2([&] { timed("pose-pose", build_pp); });
    timed("pose-line", build_pl);
    t1.join();
    t2.join();
  } else {
    build_incidence();
    build_pp();
    build_pl();
  }

  lap("inc+pp+pl");
  // the reference-order block list itself (sgb_get_structure, the parity hooks) is built on demand: build_block_list
  S.n_blocks = (int64_t)S.Pf + S.n_pairs_pp + S.n_pairs_pl + S.Lf;
  S.block_values = 9 * ((int64_t)S.Pf + S.n_pairs_pp) + 6 * S.n_pairs_pl + 4 * (int64_t)S.Lf;
  lap("block list");
  return SGB_OK;
}

// g2o's SparseBlockMatrix iteration order: block columns ascending, inside a column the rows
// ascending, upper triangle only; pose columns first, then the landmark columns (pose rows, then the diagonal).
void build_block_list(Structure& S) {
  if (S.blocks_built) return;
  const size_t nb = (size_t)S.n_blocks;
  S.blk_row.resize(nb); S.blk_col.resize(nb); S.blk_nr.resize(nb); S.blk_nc.resize(nb);
  S.blk_kind.resize(nb); S.blk_entry.resize(nb);
  auto find_in_row = [](const HostSell& M, int row, int col) {
    for (int k = 0, w = M.width(M.slice_of(row)); k < w; ++k) {
      int e = M.entry(row, k);
      if (M.col[e] == col) return e;
    }
    return -1;
  };
  size_t o = 0;
  auto put = [&](int r, int c, int nr, int nc, int kind, int entry) {
    S.blk_row[o] = r; S.blk_col[o] = c; S.blk_nr[o] = nr; S.blk_nc[o] = nc; S.blk_kind[o] = kind; S.blk_entry[o] = entry;
    ++o;
  };
  for (int c = 0; c < S.Pf; ++c) {
    // the pattern is symmetric: the rows of column c are the columns of row c (stored ascending)
    for (int k = 0, w = S.Hpp.width(S.Hpp.slice_of(c)); k < w; ++k) {
      int r = S.Hpp.col[S.Hpp.entry(c, k)];
      if (r < 0 || r >= c) break;
      put(r, c, 3, 3, 0, find_in_row(S.Hpp, r, c));  // the values of block (r, c) live in SELL row r
    }
    put(c, c, 3, 3, 0, S.hpp_diag[c]);
  }
  for (int hl = 0; hl < S.Lf; ++hl) {
    int c = S.Pf + hl;
    for (int q = S.lp_ptr[hl]; q < S.lp_ptr[hl + 1]; ++q) put(S.lp_col[q], c, 3, 2, 1, find_in_row(S.Hpl, S.lp_col[q], hl));
    put(c, c, 2, 2, 2, hl);
  }
  S.blocks_built = true;
}

}  // namespace sgb
