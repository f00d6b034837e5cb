// sgb_partition.cpp -- see sgb_partition.h. Host-only.
#include "sgb_partition.h"

#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <thread>

#include "sgb_types.h"

namespace sgb {

static int g_pushed = -1;
void partition_use_pushed_halos(bool on) { g_pushed = on ? 1 : 0; }
bool partition_pushed_halos() {
  if (g_pushed < 0) {
    const char* e = std::getenv("SGB_PUSHED_HALOS");
    g_pushed = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pushed == 1;
}
static int g_ghosts = -1;  // -1: not decided yet (environment), 0 / 1
void partition_use_ghost_landmarks(bool on) { g_ghosts = on ? 1 : 0; }
bool partition_ghost_landmarks() {
  if (g_ghosts < 0) {
    const char* e = std::getenv("SGB_GHOST_LANDMARKS");
    g_ghosts = (e && e[0] == '0') ? 0 : 1;
  }
  return g_ghosts == 1;
}

namespace {

// the k-th stored block of a SELL row and its column; returns the row width
inline int sell_row_cols(const HostSell& M, int sell_row, std::vector<int32_t>& cols) {
  cols.clear();
  int s = sell_row >> 5, lane = sell_row & 31;
  int w = (M.sbase[s + 1] - M.sbase[s]) >> 5;
  for (int k = 0; k < w; ++k) {
    int c = M.col[(size_t)M.sbase[s] + k * 32 + lane];
    if (c < 0) break;  // columns are packed at the front of a row
    cols.push_back(c);
  }
  return (int)cols.size();
}
inline int entry_k(const HostSell& M, int sell_row, int entry) { return M.k_of(sell_row, entry); }
inline int entry_of(const HostSell& M, int sell_row, int k) { return M.entry(sell_row, k); }

// flat row lists (CSR): row r holds col[ptr[r] .. ptr[r+1])
struct FlatRows {
  std::vector<int32_t> ptr, col;
  explicit FlatRows(int n = 0) : ptr(1, 0) { ptr.reserve((size_t)n + 1); }
  void close_row() { ptr.push_back((int32_t)col.size()); }
  int rows() const { return (int)ptr.size() - 1; }
  int size(int r) const { return ptr[r + 1] - ptr[r]; }
};

// SELL row s of the result holds list row order[s] (identity when order is null)
void build_local_sell(const FlatRows& rows, const std::vector<int32_t>* order, HostSell& S) {
  int n = rows.rows();
  S.rows = n;
  S.nslices = (n + 31) / 32;
  S.sbase.assign(S.nslices + 1, 0);
  for (int s = 0; s < S.nslices; ++s) {
    int w = 0;
    for (int lane = 0; lane < 32 && s * 32 + lane < n; ++lane) {
      int r = s * 32 + lane;
      w = std::max(w, rows.size(order ? (*order)[r] : r));
    }
    S.sbase[s + 1] = S.sbase[s] + w * 32;
  }
  S.col.assign((size_t)S.sbase[S.nslices], -1);
  for (int r = 0; r < n; ++r) {
    int lr = order ? (*order)[r] : r;
    int32_t* dst = S.col.data() + (size_t)S.sbase[r >> 5] + (r & 31);
    const int32_t* src = rows.col.data() + rows.ptr[lr];
    for (int k = 0, m = rows.size(lr); k < m; ++k) dst[(size_t)k * 32] = src[k];
  }
}

inline int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// Grouped SELL of the landmark-major matrix; the rows come sorted by descending length, so the first row of a
// slice is its longest. A slice whose rows have up to `len` blocks gives each row
// G = pow2_ceil(len / target_steps) lanes (<= 32), i.e. holds 32 / G rows: every warp task is then about
// `target_steps` block-steps long whatever the observer count (20 observers per wall in a grid world, hundreds
// for the hub landmarks of a corridor graph), which is what balances the landmark pass of the Schur product.
void build_grouped_sell(const FlatRows& rows, int target_steps, HostSell& S) {
  const int n = rows.rows();
  S.rows = n;
  S.sbase.assign(1, 0);
  S.srow.clear();
  S.sshift.clear();
  S.row_slice.assign(n, 0);
  for (int r = 0; r < n;) {
    int len = rows.size(r);
    int G = std::min(32, pow2_ceil((len + target_steps - 1) / std::max(1, target_steps)));
    int rh = 32 / G, sh = 0;
    // the rows of one run come sorted by descending length; where two runs meet (owned rows, then ghost rows) a later
    // row of the slice may be longer than its first: widen until the slice's longest row fits
    for (bool grown = true; grown;) {
      grown = false;
      for (int q = r + 1; q < std::min(n, r + rh); ++q)
        if (rows.size(q) > len) { len = rows.size(q); grown = true; }
      if (grown) {
        G = std::min(32, pow2_ceil((len + target_steps - 1) / std::max(1, target_steps)));
        rh = 32 / G;
      }
    }
    while ((1 << sh) < rh) ++sh;
    int w = (len + G - 1) / G * G;
    int s = (int)S.sshift.size();
    S.srow.push_back(r);
    S.sshift.push_back(sh);
    S.sbase.push_back(S.sbase.back() + w * rh);
    for (int q = r; q < std::min(n, r + rh); ++q) S.row_slice[q] = s;
    r += rh;
  }
  S.nslices = (int)S.sshift.size();
  S.srow.push_back(n);
  S.col.assign((size_t)S.sbase[S.nslices], -1);
  for (int r = 0; r < n; ++r) {
    const int32_t* src = rows.col.data() + rows.ptr[r];
    for (int k = 0, m = rows.size(r); k < m; ++k) S.col[(size_t)S.entry(r, k)] = src[k];
  }
}

// block-steps per lane in the landmark pass: a large matrix is throughput-bound (fewer, longer lane loops issue
// fewer warp-steps), a small one latency-bound (more lanes per row shorten the dependent chain)
int lm_target_steps(const FlatRows& rows) {
  static const int forced = [] {
    const char* e = std::getenv("SGB_LM_STEPS");  // tuning knob
    return e ? std::atoi(e) : 0;
  }();
  if (forced > 0) return forced;
  const size_t blocks = rows.col.size();
  if (blocks >= ((size_t)1 << 20)) return 10;
  // Latency-bound sizes: a warp of the persistent kernel (3 CTAs x 148 SMs x 8 warps) takes the slices w, w + warps, ...
  // one after the other, each slice costing its number of block-steps. Few steps per slice means many slices: on one of
  // eight shards of a 1M-pose graph 2 steps give 13 000 slices = FOUR per warp (8 step-units), 5 steps give 3 250 = one
  // per warp (5 step-units). Pick the target that minimises ceil(slices / warps) x steps; small graphs (fewer slices
  // than warps whatever the target) keep the shortest chain, 2.
  const int warps = 148 * 3 * 8;
  int best = 2;
  long best_cost = -1;
  for (int ts : {2, 3, 4, 5, 6, 8, 10}) {
    long slices = 0, steps_max = 0;
    for (int r = 0, n = rows.rows(); r < n;) {
      const int len = rows.size(r);
      const int G = std::min(32, pow2_ceil((len + ts - 1) / ts));
      slices++;
      steps_max = std::max<long>(steps_max, (len + G - 1) / G);
      r += 32 / G;
    }
    const long cost = ((slices + warps - 1) / warps) * std::max<long>(steps_max, 1);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ts; }
  }
  return best;
}

// world == 1: the local plan IS the global structure -- same rows, same edge order, same pose-major SELL layout;
// only the landmark numbering (row order of the landmark-major matrix) differs from the Hessian order. Plain copies
// on four host threads instead of the general owner/halo classification.
// `consume` (the product path: the Structure belongs to the handle and nobody reads these arrays again): the per-edge
// arrays and the incidence lists are MOVED into the plan instead of copied (~250 MB on the 1M-pose graph).
sgb_status partition_single(const Structure& S, LocalPlan& P, Structure* consume) {
  P.n_pp = P.n_pp_owned = S.n_pp;
  P.n_pl = P.n_pl_owned = S.n_pl;
  auto pose_side = [&]() {
    P.pp_g.resize(S.n_pp);
    std::iota(P.pp_g.begin(), P.pp_g.end(), 0);
    if (!consume) {
      P.pp_i = S.pp_i; P.pp_j = S.pp_j; P.pp_hi = S.pp_hi; P.pp_hj = S.pp_hj;
      P.pp_e_ij = S.pp_e_ij; P.pp_e_ji = S.pp_e_ji; P.pp_dup = S.pp_dup;
    }
    P.Hpp = S.Hpp;  // enc_pose(c) == c for a single owner
    P.hpp_diag = S.hpp_diag;
  };
  auto hpl_side = [&]() {
    P.Hpl = S.Hpl;
    for (auto& c : P.Hpl.col)
      if (c >= 0) c = P.enc_lm[c];
    if (!consume) {
      P.pinc_ptr = S.pinc_ptr;
      P.pinc = S.pinc;
    }
  };
  auto linc_side = [&]() {  // incidence lists of the landmarks in the plan's row order (independent of the SELL build)
    P.linc_ptr.assign(P.nL + 1, 0);
    P.linc.reserve(S.linc.size());
    for (int l = 0; l < P.nL; ++l) {
      int hl = P.lm_global[l];
      P.linc.insert(P.linc.end(), S.linc.begin() + S.linc_ptr[hl], S.linc.begin() + S.linc_ptr[hl + 1]);
      P.linc_ptr[l + 1] = (int)P.linc.size();
    }
  };
  auto lm_side = [&]() {
    FlatRows obs(P.nL);
    obs.col.reserve(S.lp_col.size());
    for (int l = 0; l < P.nL; ++l) {
      int hl = P.lm_global[l];
      obs.col.insert(obs.col.end(), S.lp_col.begin() + S.lp_ptr[hl], S.lp_col.begin() + S.lp_ptr[hl + 1]);
      obs.close_row();
    }
    build_grouped_sell(obs, lm_target_steps(obs), P.Hlp);
    // needs Hlp: the landmark-major entry of every leading pose-line edge (independent per edge: two threads when large)
    P.pl_e_lp.assign(S.n_pl, -1);
    auto entries = [&](int k0, int k1) {
      for (int k = k0; k < k1; ++k)
        if (S.pl_e_pl[k] >= 0) P.pl_e_lp[k] = P.Hlp.entry(P.enc_lm[S.pl_hl[k]] & kLocalMask, S.pl_k_lp[k]);
    };
    if (S.n_pl > 1000000) {
      std::thread t(entries, 0, S.n_pl / 2);
      entries(S.n_pl / 2, S.n_pl);
      t.join();
    } else {
      entries(0, S.n_pl);
    }
  };
  auto pl_side = [&]() {
    P.pl_g.resize(S.n_pl);
    std::iota(P.pl_g.begin(), P.pl_g.end(), 0);
    if (!consume) {
      P.pl_p = S.pl_p; P.pl_l = S.pl_l; P.pl_hp = S.pl_hp; P.pl_hl = S.pl_hl;
      P.pl_e_pl = S.pl_e_pl; P.pl_dup = S.pl_dup;
    }
  };
  if ((size_t)S.n_pp + S.n_pl > 200000) {
    std::thread t1(pose_side), t2(hpl_side), t3(pl_side), t4(linc_side);
    lm_side();
    t1.join();
    t2.join();
    t3.join();
    t4.join();
  } else {
    pose_side();
    hpl_side();
    pl_side();
    linc_side();
    lm_side();
  }
  if (consume) {  // after the threads: lm_side reads S.pl_e_pl / S.pl_hl
    Structure& M = *consume;
    P.pp_i.swap(M.pp_i); P.pp_j.swap(M.pp_j); P.pp_hi.swap(M.pp_hi); P.pp_hj.swap(M.pp_hj);
    P.pp_e_ij.swap(M.pp_e_ij); P.pp_e_ji.swap(M.pp_e_ji); P.pp_dup.swap(M.pp_dup);
    P.pinc_ptr.swap(M.pinc_ptr); P.pinc.swap(M.pinc);
    P.pl_p.swap(M.pl_p); P.pl_l.swap(M.pl_l); P.pl_hp.swap(M.pl_hp); P.pl_hl.swap(M.pl_hl);
    P.pl_e_pl.swap(M.pl_e_pl); P.pl_dup.swap(M.pl_dup);
  }
  return SGB_OK;
}

}  // namespace

static sgb_status partition_impl(const Structure& S, int world, int rank, LocalPlan& P, std::string& err, Structure* consume);
sgb_status partition(const Structure& S, int world, int rank, LocalPlan& P, std::string& err) {
  return partition_impl(S, world, rank, P, err, nullptr);
}
sgb_status partition_consume(Structure& S, int world, int rank, LocalPlan& P, std::string& err) {
  return partition_impl(S, world, rank, P, err, world == 1 ? &S : nullptr);
}
static sgb_status partition_impl(const Structure& S, int world, int rank, LocalPlan& P, std::string& err, Structure* consume) {
  P = LocalPlan();
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) { err = "bad world/rank"; return SGB_ERR_INVALID; }
  P.world = world;
  P.rank = rank;
  P.chunkP = partition_chunk(S.Pf, world);
  if (S.filtered && (S.fworld != world || S.frank != rank)) { err = "structure was filtered for another rank"; return SGB_ERR_INVALID; }
  if (S.filtered && !(world > 1 && partition_ghost_landmarks())) { err = "a rank-filtered structure needs ghost landmark rows"; return SGB_ERR_INVALID; }
  // landmarks this rank knows about: all of them, or (filtered build) the ones it keeps a row of
  auto present = [&](int hl) { return !S.filtered || S.lm_present[hl] != 0; };
  if (P.chunkP > kLocalMask) { err = "too many rows per rank for the column encoding"; return SGB_ERR_UNSUPPORTED; }
  auto owner_p = [&](int hp) { return hp / P.chunkP; };
  P.p_begin = std::min(S.Pf, rank * P.chunkP);
  int p_end = std::min(S.Pf, (rank + 1) * P.chunkP);
  P.nP = std::max(0, p_end - P.p_begin);
  P.capP = P.chunkP;

  // ---- landmark ownership: rank of the first free observer (global Hlp rows list observers ascending)
  std::vector<int32_t> lm_owner(S.Lf, 0), lm_first(S.Lf, -1);
  for (int k = 0; k < S.n_pl; ++k) {
    int hp = S.pl_hp[k], hl = S.pl_hl[k];
    if (hp < 0 || hl < 0) continue;
    if (lm_first[hl] < 0 || hp < lm_first[hl]) lm_first[hl] = hp;
  }
  // Local landmark numbering = row order of the landmark-major matrix: per owner, descending observer count
  // (stable in the Hessian index), so that the slices of the grouped SELL are uniform and no row -> landmark
  // indirection is needed on the device. One stable counting sort over (owner, -count) covers every rank.
  std::vector<int32_t> count(world, 0);
  P.enc_lm.assign(S.Lf, -1);  // stays -1 for a landmark this rank knows nothing about (filtered build)
  {
    int maxc = 0;
    for (int hl = 0; hl < S.Lf; ++hl) {
      lm_owner[hl] = !present(hl) ? -1 : (lm_first[hl] >= 0 ? owner_p(lm_first[hl]) : 0);
      maxc = std::max(maxc, S.lp_ptr[hl + 1] - S.lp_ptr[hl]);
    }
    const size_t nb = (size_t)world * ((size_t)maxc + 1);
    std::vector<int32_t> start(nb + 1, 0);
    auto key = [&](int hl) { return (size_t)lm_owner[hl] * ((size_t)maxc + 1) + (size_t)(maxc - (S.lp_ptr[hl + 1] - S.lp_ptr[hl])); };
    int n_present = 0;
    for (int hl = 0; hl < S.Lf; ++hl)
      if (lm_owner[hl] >= 0) { start[key(hl) + 1]++; ++n_present; }
    for (size_t b = 0; b < nb; ++b) start[b + 1] += start[b];
    std::vector<int32_t> sorted(n_present);
    for (int hl = 0; hl < S.Lf; ++hl)
      if (lm_owner[hl] >= 0) sorted[start[key(hl)]++] = hl;
    for (int q = 0; q < n_present; ++q) {
      int hl = sorted[q], o = lm_owner[hl];
      P.enc_lm[hl] = (o << kOwnerShift) | count[o];
      if (o == rank) P.lm_global.push_back(hl);
      count[o]++;
    }
  }
  P.nL = P.nL_owned = count[rank];
  P.capL = *std::max_element(count.begin(), count.end());
  // global free landmark -> row on THIS rank (owned rows first, then the ghost copies), -1 = not kept here
  std::vector<int32_t> loc_lm(S.Lf, -1);
  for (int l = 0; l < P.nL_owned; ++l) loc_lm[P.lm_global[l]] = l;
  if (world > 1 && partition_ghost_landmarks()) {
    std::vector<int32_t> nghost(world, 0), mine;
    std::vector<char> seen(world);
    for (int hl = 0; hl < S.Lf; ++hl) {
      if (lm_owner[hl] < 0) continue;
      std::fill(seen.begin(), seen.end(), 0);
      for (int q = S.lp_ptr[hl]; q < S.lp_ptr[hl + 1]; ++q) {
        int o = owner_p(S.lp_col[q]);
        if (o == lm_owner[hl] || seen[o]) continue;
        seen[o] = 1;
        nghost[o]++;
        if (o == rank) mine.push_back(hl);
      }
    }
    std::stable_sort(mine.begin(), mine.end(), [&](int a, int b) {
      return S.lp_ptr[a + 1] - S.lp_ptr[a] > S.lp_ptr[b + 1] - S.lp_ptr[b];
    });
    for (int hl : mine) {
      loc_lm[hl] = (int)P.lm_global.size();
      P.lm_global.push_back(hl);
    }
    P.nL = (int)P.lm_global.size();
    // with ghost rows no landmark-sized array is read by another rank: its stride is this rank's own row count (a
    // filtered build does not even know the other ranks' counts)
    P.capL = P.nL;
  }
  if (P.capL > kLocalMask) { err = "too many landmarks per rank for the column encoding"; return SGB_ERR_UNSUPPORTED; }
  P.pose_of_l.resize(P.nP);
  for (int l = 0; l < P.nP; ++l) P.pose_of_l[l] = S.pose_of_h[P.p_begin + l];
  P.lm_of_l.resize(P.nL);
  for (int l = 0; l < P.nL; ++l) P.lm_of_l[l] = S.lm_of_h[P.lm_global[l]];

  if (world == 1) return partition_single(S, P, consume);

  auto pose_local = [&](int hp) { return hp >= 0 && owner_p(hp) == rank; };
  P.enc_lm_here.resize(S.Lf);
  for (int hl = 0; hl < S.Lf; ++hl) P.enc_lm_here[hl] = loc_lm[hl] >= 0 ? ((rank << kOwnerShift) | loc_lm[hl]) : P.enc_lm[hl];
  // (filtered build: an absent landmark keeps enc -1 and is never referenced -- none of this rank's poses observes it)
  auto lm_local = [&](int hl) { return hl >= 0 && loc_lm[hl] >= 0; };          // owned or ghost row here
  auto lm_owned = [&](int hl) { return hl >= 0 && lm_owner[hl] == rank; };

  // ---- local edges: owned (chi2 accounted here) first, then the halo edges; global order inside each group
  std::vector<int32_t> pp_g2l(S.n_pp, -1), pl_g2l(S.n_pl, -1);
  P.pp_g.reserve(world == 1 ? S.n_pp : S.n_pp / world + 1024);
  P.pl_g.reserve(world == 1 ? S.n_pl : S.n_pl / world + 1024);
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = 0; k < S.n_pp; ++k) {
      int hi = S.pp_hi[k], hj = S.pp_hj[k];
      bool local = pose_local(hi) || pose_local(hj);
      if (!local) continue;
      bool owned = hi >= 0 ? pose_local(hi) : pose_local(hj);
      if ((pass == 0) != owned) continue;
      pp_g2l[k] = (int)P.pp_g.size();
      P.pp_g.push_back(k);
    }
    if (pass == 0) P.n_pp_owned = (int)P.pp_g.size();
  }
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = 0; k < S.n_pl; ++k) {
      int hp = S.pl_hp[k], hl = S.pl_hl[k];
      bool local = pose_local(hp) || lm_local(hl);
      if (!local) continue;
      bool owned = hp >= 0 ? pose_local(hp) : lm_owned(hl);
      if ((pass == 0) != owned) continue;
      pl_g2l[k] = (int)P.pl_g.size();
      P.pl_g.push_back(k);
    }
    if (pass == 0) P.n_pl_owned = (int)P.pl_g.size();
  }
  P.n_pp = (int)P.pp_g.size();
  P.n_pl = (int)P.pl_g.size();

  // ---- local SELL matrices with encoded columns (pose rows and landmark rows are independent: two host threads)
  int64_t halo_p_pose = 0, halo_p_lm = 0;
  FlatRows rows(P.nP), rows_pl(P.nP), obs(P.nL);
  auto build_pose_rows = [&]() {
    std::vector<int32_t> cols;
    rows.col.reserve(S.Hpp.col.size() / world + 1024);
    rows_pl.col.reserve(S.Hpl.col.size() / world + 1024);
    for (int l = 0; l < P.nP; ++l) {
      int hp = P.p_begin + l;
      sell_row_cols(S.Hpp, hp, cols);
      for (int c : cols) {
        rows.col.push_back(enc_pose(c, P.chunkP));
        if (owner_p(c) != rank) halo_p_pose++;
      }
      rows.close_row();
      sell_row_cols(S.Hpl, hp, cols);
      for (int c : cols) {
        rows_pl.col.push_back(P.enc_lm_here[c]);
        if (loc_lm[c] < 0) P.halo_t++;
      }
      rows_pl.close_row();
    }
  };
  auto finish_pose_rows = [&]() {
    build_local_sell(rows, nullptr, P.Hpp);
    build_local_sell(rows_pl, nullptr, P.Hpl);
    P.hpp_diag.resize(P.nP);
    for (int l = 0; l < P.nP; ++l) {
      int hp = P.p_begin + l;
      P.hpp_diag[l] = entry_of(P.Hpp, l, entry_k(S.Hpp, hp, S.hpp_diag[hp]));
    }
  };
  auto build_lm_rows = [&]() {
    // owned landmarks sorted by descending observer count (stable) to keep the SELL padding small
    obs.col.reserve(S.lp_col.size() / world + 1024);
    for (int l = 0; l < P.nL; ++l) {
      int hl = P.lm_global[l];
      for (int q = S.lp_ptr[hl]; q < S.lp_ptr[hl + 1]; ++q) {
        int c = S.lp_col[q];
        obs.col.push_back(enc_pose(c, P.chunkP));
        if (owner_p(c) != rank) halo_p_lm++;
      }
      obs.close_row();
    }
  };
  auto finish_lm_rows = [&]() {
    build_grouped_sell(obs, lm_target_steps(obs), P.Hlp);  // rows already sorted by descending length
  };
  const bool threaded = (size_t)P.n_pp + P.n_pl > 200000;
  if (threaded) {
    std::thread t(build_lm_rows);
    build_pose_rows();
    t.join();
  } else {
    build_pose_rows();
    build_lm_rows();
  }
  P.halo_p = halo_p_pose + halo_p_lm;

  // ---- pushed pose halos (sgb_partition.h): with ghost rows the only remote quantities left in the operator passes are
  // pose-vector entries; give them local halo slots and work out who has to push what to whom
  P.pushed = world > 1 && partition_ghost_landmarks() && partition_pushed_halos();
  if (P.pushed) {
    auto glob = [&](int enc) { return (enc >> kOwnerShift) * P.chunkP + (enc & kLocalMask); };
    std::vector<int32_t>& need = P.halo_src;
    for (int enc : rows.col) if ((enc >> kOwnerShift) != rank) need.push_back(glob(enc));
    for (int enc : obs.col) if ((enc >> kOwnerShift) != rank) need.push_back(glob(enc));
    std::sort(need.begin(), need.end());
    need.erase(std::unique(need.begin(), need.end()), need.end());
    P.nH = (int)need.size();
    if ((int64_t)P.capP + P.nH > kLocalMask) { err = "too many halo rows for the column encoding"; return SGB_ERR_UNSUPPORTED; }
    for (int hp : need) P.halo_cnt[owner_p(hp)]++;
    for (int o = 1; o < world; ++o) P.halo_base[o] = P.halo_base[o - 1] + P.halo_cnt[o - 1];
    auto slot_enc = [&](int enc) {
      int s = (int)(std::lower_bound(need.begin(), need.end(), glob(enc)) - need.begin());
      return (rank << kOwnerShift) | (P.capP + s);
    };
    // sender side, before the columns are re-encoded: rows with a pose-pose block in a row of q ...
    std::vector<std::vector<int32_t>> send_rows(world);
    for (int l = 0; l < P.nP; ++l)
      for (int q = rows.ptr[l]; q < rows.ptr[l + 1]; ++q) {
        int o = rows.col[q] >> kOwnerShift;
        if (o != rank) send_rows[o].push_back(l);
      }
    // ... and rows that observe a landmark some pose of q observes (q keeps a full copy of that landmark's row)
    {
      std::vector<char> has(world);
      for (int lr = 0; lr < P.nL; ++lr) {
        std::fill(has.begin(), has.end(), 0);
        bool any_remote = false;
        for (int q = obs.ptr[lr]; q < obs.ptr[lr + 1]; ++q) {
          int o = obs.col[q] >> kOwnerShift;
          has[o] = 1;
          any_remote |= o != rank;
        }
        if (!any_remote) continue;
        for (int q = obs.ptr[lr]; q < obs.ptr[lr + 1]; ++q) {
          if ((obs.col[q] >> kOwnerShift) != rank) continue;
          int l = obs.col[q] & kLocalMask;
          for (int o = 0; o < world; ++o)
            if (has[o] && o != rank) send_rows[o].push_back(l);
        }
      }
    }
    std::vector<int32_t> cnt(P.nP + 1, 0);
    for (int o = 0; o < world; ++o) {
      auto& v = send_rows[o];
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      P.send_cnt[o] = (int32_t)v.size();
      for (int l : v) cnt[l + 1]++;
    }
    P.send_ptr.assign(P.nP + 1, 0);
    for (int l = 0; l < P.nP; ++l) P.send_ptr[l + 1] = P.send_ptr[l] + cnt[l + 1];
    P.send_dst.assign(P.send_ptr[P.nP], 0);
    std::vector<int32_t> pos(P.send_ptr.begin(), P.send_ptr.end() - 1);
    for (int o = 0; o < world; ++o)
      for (size_t k = 0; k < send_rows[o].size(); ++k) P.send_dst[pos[send_rows[o][k]]++] = (o << kOwnerShift) | (int32_t)k;
    // the gathers become local
    for (auto& enc : rows.col) if ((enc >> kOwnerShift) != rank) enc = slot_enc(enc);
    for (auto& enc : obs.col) if ((enc >> kOwnerShift) != rank) enc = slot_enc(enc);
  }
  if (threaded) {
    std::thread t(finish_lm_rows);
    finish_pose_rows();
    t.join();
  } else {
    finish_pose_rows();
    finish_lm_rows();
  }

  // ---- per-edge arrays
  auto build_pp_edges = [&]() {
  P.pp_i.resize(P.n_pp); P.pp_j.resize(P.n_pp); P.pp_hi.resize(P.n_pp); P.pp_hj.resize(P.n_pp);
  P.pp_e_ij.assign(P.n_pp, -1); P.pp_e_ji.assign(P.n_pp, -1); P.pp_dup.assign(P.n_pp, -1);
  for (int l = 0; l < P.n_pp; ++l) {
    int k = P.pp_g[l];
    P.pp_i[l] = S.pp_i[k]; P.pp_j[l] = S.pp_j[k]; P.pp_hi[l] = S.pp_hi[k]; P.pp_hj[l] = S.pp_hj[k];
    if (S.pp_dup[k] >= 0) P.pp_dup[l] = pp_g2l[S.pp_dup[k]];  // chain members share both endpoints => all local
    int hi = S.pp_hi[k], hj = S.pp_hj[k];
    if (S.pp_e_ij[k] >= 0) {  // leader of a free pair
      if (pose_local(hi)) P.pp_e_ij[l] = entry_of(P.Hpp, hi - P.p_begin, entry_k(S.Hpp, hi, S.pp_e_ij[k]));
      if (pose_local(hj)) P.pp_e_ji[l] = entry_of(P.Hpp, hj - P.p_begin, entry_k(S.Hpp, hj, S.pp_e_ji[k]));
    }
  }
  };
  auto build_pl_edges = [&]() {
  P.pl_p.resize(P.n_pl); P.pl_l.resize(P.n_pl); P.pl_hp.resize(P.n_pl); P.pl_hl.resize(P.n_pl);
  P.pl_e_pl.assign(P.n_pl, -1); P.pl_e_lp.assign(P.n_pl, -1); P.pl_dup.assign(P.n_pl, -1);
  for (int l = 0; l < P.n_pl; ++l) {
    int k = P.pl_g[l];
    P.pl_p[l] = S.pl_p[k]; P.pl_l[l] = S.pl_l[k]; P.pl_hp[l] = S.pl_hp[k]; P.pl_hl[l] = S.pl_hl[k];
    if (S.pl_dup[k] >= 0) P.pl_dup[l] = pl_g2l[S.pl_dup[k]];
    int hp = S.pl_hp[k], hl = S.pl_hl[k];
    if (S.pl_e_pl[k] >= 0) {
      if (pose_local(hp)) P.pl_e_pl[l] = entry_of(P.Hpl, hp - P.p_begin, entry_k(S.Hpl, hp, S.pl_e_pl[k]));
      if (lm_local(hl)) P.pl_e_lp[l] = entry_of(P.Hlp, loc_lm[hl], S.pl_k_lp[k]);
    }
  }

  };
  // ---- incidence lists of owned rows, local edge ids
  auto build_incidence = [&]() {
  P.pinc_ptr.assign(P.nP + 1, 0);
  P.pinc.reserve(P.nP > 0 ? (size_t)(S.pinc_ptr[P.p_begin + P.nP] - S.pinc_ptr[P.p_begin]) : 0);
  for (int l = 0; l < P.nP; ++l) {
    int hp = P.p_begin + l;
    for (int q = S.pinc_ptr[hp]; q < S.pinc_ptr[hp + 1]; ++q) {
      int packed = S.pinc[q];
      int k = packed >> 2, low = packed & 3;
      int lk = (low & 1) ? pl_g2l[k] : pp_g2l[k];
      P.pinc.push_back((lk << 2) | low);
    }
    P.pinc_ptr[l + 1] = (int)P.pinc.size();
  }
  P.linc_ptr.assign(P.nL + 1, 0);
  for (int l = 0; l < P.nL; ++l) {
    int hl = P.lm_global[l];
    for (int q = S.linc_ptr[hl]; q < S.linc_ptr[hl + 1]; ++q) P.linc.push_back(pl_g2l[S.linc[q]]);
    P.linc_ptr[l + 1] = (int)P.linc.size();
  }

  };
  if (threaded) {
    std::thread t1(build_pp_edges), t2(build_incidence);
    build_pl_edges();
    t1.join();
    t2.join();
  } else {
    build_pp_edges();
    build_pl_edges();
    build_incidence();
  }
  return SGB_OK;
}

// where every block of the reference-order list lives: owner rank and, on the owner, the local entry
int coarse_spacing(int nP, int max_nodes) {
  if (nP < 24 || max_nodes < 3) return 0;  // a handful of rows: block-Jacobi alone converges in a few iterations
  for (int h : {8, 16, 32, 64})
    if ((nP + h - 1) / h + 1 <= max_nodes) return h;
  return 0;
}

// Gather lists of the coarse matrix R S R^T (sgb_coarse.h) for hat functions with a node every h pose rows. Everything is
// emitted in a fixed order (rows ascending, blocks of a row in storage order), so the device sums are reproducible.
void plan_coarse(const LocalPlan& P, int h, CoarsePlan& C) {
  C = CoarsePlan();
  if (h <= 0 || P.world != 1 || P.nP <= 0) return;
  const int nP = P.nP, nn = (nP + h - 1) / h + 1;
  C.h = h;
  C.nn = nn;
  struct NodeW { int n[2]; double w[2]; int cnt; };
  auto nodes_of = [&](int i) {
    NodeW r;
    const double wr = (double)(i % h) / (double)h;
    r.n[0] = i / h; r.w[0] = 1.0 - wr; r.cnt = 1;
    if (wr != 0.0) { r.n[1] = i / h + 1; r.w[1] = wr; r.cnt = 2; }
    return r;
  };
  const int npair = nn * nn;
  // ---- Hpp items and R R^T per node pair (m <= n): count, then fill
  C.p_ptr.assign((size_t)npair + 1, 0);
  C.rr.assign(npair, 0.0);
  auto each_pp = [&](auto&& f) {
    for (int i = 0; i < nP; ++i) {
      const NodeW a = nodes_of(i);
      const int w = P.Hpp.width(P.Hpp.slice_of(i));
      for (int k = 0; k < w; ++k) {
        const int e = P.Hpp.entry(i, k), enc = P.Hpp.col[e];
        if (enc < 0) continue;
        const NodeW b = nodes_of(enc & kLocalMask);
        for (int x = 0; x < a.cnt; ++x)
          for (int y = 0; y < b.cnt; ++y)
            if (a.n[x] <= b.n[y]) f(a.n[x] * nn + b.n[y], e, a.w[x] * b.w[y]);
      }
    }
  };
  each_pp([&](int b, int, double) { C.p_ptr[b + 1]++; });
  for (int b = 0; b < npair; ++b) C.p_ptr[b + 1] += C.p_ptr[b];
  C.p_e.resize(C.p_ptr[npair]);
  C.p_w.resize(C.p_ptr[npair]);
  {
    std::vector<int32_t> pos(C.p_ptr.begin(), C.p_ptr.end() - 1);
    each_pp([&](int b, int e, double w) { C.p_e[pos[b]] = e; C.p_w[pos[b]++] = w; });
  }
  for (int i = 0; i < nP; ++i) {
    const NodeW a = nodes_of(i);
    for (int x = 0; x < a.cnt; ++x)
      for (int y = 0; y < a.cnt; ++y)
        if (a.n[x] <= a.n[y]) C.rr[a.n[x] * nn + a.n[y]] += a.w[x] * a.w[y];
  }
  // ---- G = R Hpl: its non-zero (landmark, node) pairs, ordered by landmark then node; the items of a pair in row order
  const int nL = P.nL;
  std::vector<int32_t> key_cnt((size_t)nL * nn + 1, 0);
  auto each_pl = [&](auto&& f) {
    for (int i = 0; i < nP; ++i) {
      const NodeW a = nodes_of(i);
      const int w = P.Hpl.rows > 0 ? P.Hpl.width(P.Hpl.slice_of(i)) : 0;
      for (int k = 0; k < w; ++k) {
        const int e = P.Hpl.entry(i, k), enc = P.Hpl.col[e];
        if (enc < 0) continue;
        const int l = enc & kLocalMask;
        for (int x = 0; x < a.cnt; ++x) f(l * nn + a.n[x], e, a.w[x]);
      }
    }
  };
  each_pl([&](int key, int, double) { key_cnt[key + 1]++; });
  std::vector<int32_t> g_of_key((size_t)nL * nn, -1);
  C.g_ptr.assign(1, 0);
  for (int key = 0; key < nL * nn; ++key)
    if (key_cnt[key + 1] > 0) {
      g_of_key[key] = (int)C.g_lm.size();
      C.g_lm.push_back(key / nn);
      C.g_ptr.push_back(C.g_ptr.back() + key_cnt[key + 1]);
    }
  C.ng = (int)C.g_lm.size();
  C.g_e.resize(C.g_ptr.back());
  C.g_w.resize(C.g_ptr.back());
  {
    std::vector<int32_t> pos(C.g_ptr.begin(), C.g_ptr.end() - 1);
    each_pl([&](int key, int e, double w) { const int gi = g_of_key[key]; C.g_e[pos[gi]] = e; C.g_w[pos[gi]++] = w; });
  }
  // ---- Schur terms: the pairs of one landmark, (g1, g2) with node(g1) <= node(g2), grouped by node pair
  C.t_ptr.assign((size_t)npair + 1, 0);
  auto each_term = [&](auto&& f) {
    int g0 = 0;
    while (g0 < C.ng) {
      int g1 = g0;
      while (g1 < C.ng && C.g_lm[g1] == C.g_lm[g0]) ++g1;
      const int l = C.g_lm[g0];
      // the pairs of landmark l are g0 .. g1-1 in node order; node of pair gi = its key % nn: recover it from the keys
      for (int a = g0; a < g1; ++a)
        for (int b = a; b < g1; ++b) f(a, b, l);
      g0 = g1;
    }
  };
  std::vector<int32_t> g_node(C.ng, 0);
  for (int key = 0; key < nL * nn; ++key)
    if (g_of_key[key] >= 0) g_node[g_of_key[key]] = key % nn;
  each_term([&](int a, int b, int) { C.t_ptr[g_node[a] * nn + g_node[b] + 1]++; });
  for (int b = 0; b < npair; ++b) C.t_ptr[b + 1] += C.t_ptr[b];
  C.t_g.resize(2 * (size_t)C.t_ptr[npair]);
  {
    std::vector<int32_t> pos(C.t_ptr.begin(), C.t_ptr.end() - 1);
    each_term([&](int a, int b, int) {
      const int q = pos[g_node[a] * nn + g_node[b]]++;
      C.t_g[2 * (size_t)q] = a;
      C.t_g[2 * (size_t)q + 1] = b;
    });
  }
}

void build_export(Structure& S, LocalPlan& P) {
  if (P.export_built) return;
  build_block_list(S);
  auto owner_p = [&](int hp) { return hp / P.chunkP; };
  P.blk_owner.resize(S.blk_row.size());
  P.blk_entry.assign(S.blk_row.size(), -1);
  for (size_t b = 0; b < S.blk_row.size(); ++b) {
    int kind = S.blk_kind[b];
    if (kind == 2) {
      int hl = S.blk_entry[b], o = P.enc_lm[hl] < 0 ? -1 : (P.enc_lm[hl] >> kOwnerShift);
      P.blk_owner[b] = o;
      if (o == P.rank) P.blk_entry[b] = P.enc_lm[hl] & kLocalMask;
    } else {
      int hp = S.blk_row[b];  // block (row r, col c) is stored in pose row r
      P.blk_owner[b] = owner_p(hp);
      if (owner_p(hp) == P.rank) {
        const HostSell& G = kind == 0 ? S.Hpp : S.Hpl;
        const HostSell& Lc = kind == 0 ? P.Hpp : P.Hpl;
        P.blk_entry[b] = entry_of(Lc, hp - P.p_begin, entry_k(G, hp, S.blk_entry[b]));
      }
    }
  }
  P.export_built = true;
}

}  // namespace sgb
