// sgb_structure.h -- host-side symbolic phase: g2o's index mapping and block structure, plus the scatter maps
// (SELL entries, incidence lists, duplicate chains) the kernels assemble through.
//
// Restates SparseOptimizer::initializeOptimization / buildIndexMapping and BlockSolver::buildStructure of the
// un-vendored g2o; the reference triggers them at drone.cpp:148,
// submap_loop_closer.cpp:286 and log_runner.cpp:203.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sgb_capi.h"

namespace sgb {

// Sliced-ELL block pattern. Uniform form (pose-major matrices): 32-row slices, entry of (row, k) =
// sbase[row / 32] + k * 32 + row % 32. Grouped form (the landmark-major matrix): slice s holds 2^sshift[s] rows
// starting at srow[s] and gives each of them G = 32 >> sshift[s] lanes; the k-th block of its row rr sits at
// sbase[s] + (k / G) * 32 + rr * G + k % G, so that a warp whose lane (rr * G + kk) walks k = kk, kk + G, ... reads 32
// consecutive entries per step and neighbouring lanes hold neighbouring blocks of the same row.
struct HostSell {
  int rows = 0, nslices = 0;
  std::vector<int32_t> sbase;  // [nslices + 1] entry offset of each slice (multiple of 32)
  std::vector<int32_t> col;    // [entries]
  std::vector<int32_t> srow, sshift, row_slice;  // grouped form only: [nslices + 1], [nslices], [rows]
  int64_t entries() const { return (int64_t)col.size(); }
  bool grouped() const { return !sshift.empty(); }
  int slice_of(int row) const { return grouped() ? row_slice[row] : row >> 5; }
  int shift(int s) const { return grouped() ? sshift[s] : 5; }
  int first_row(int s) const { return grouped() ? srow[s] : s << 5; }
  int width(int s) const { return (sbase[s + 1] - sbase[s]) >> shift(s); }  // blocks per row, padding included
  int entry(int row, int k) const {
    int s = slice_of(row);
    if (!grouped()) return sbase[s] + (k << 5) + (row & 31);
    int gsh = 5 - sshift[s], G = 1 << gsh;
    return sbase[s] + ((k >> gsh) << 5) + ((row - srow[s]) << gsh) + (k & (G - 1));
  }
  int k_of(int row, int e) const {
    int s = slice_of(row), d = e - sbase[s];
    if (!grouped()) return d >> 5;
    int gsh = 5 - sshift[s];
    return ((d >> 5) << gsh) + (d & ((1 << gsh) - 1));
  }
};

struct Structure {
  int P_all = 0, L_all = 0, Pf = 0, Lf = 0, n_pp = 0, n_pl = 0, dim = 0;
  bool has_robust = false;
  // Rank-filtered build (multi-GPU with ghost landmark rows): the index mapping is the global one, but only the edges
  // rank `frank` of `fworld` needs were kept -- pose-pose edges with an endpoint in its row block, and every edge of a
  // landmark one of its poses observes. Rows of other ranks are empty; lm_present marks the landmarks this rank keeps.
  bool filtered = false;
  int fworld = 1, frank = 0;
  std::vector<char> lm_present;             // [Lf], filtered builds only
  // vertex maps
  std::vector<int32_t> pose_h, lm_h;        // array index -> free index (-1 fixed / inactive)
  std::vector<int32_t> pose_of_h, lm_of_h;  // free index -> array index
  // active edges in insertion order: index into the caller's arrays
  std::vector<int32_t> pp_src, pl_src;
  // per active edge
  std::vector<int32_t> pp_i, pp_j, pp_hi, pp_hj, pp_e_ij, pp_e_ji, pp_dup;
  std::vector<int32_t> pl_p, pl_l, pl_hp, pl_hl, pl_e_pl, pl_k_lp, pl_dup;  // pl_k_lp: position in the landmark's row
  // incidence
  std::vector<int32_t> pinc_ptr, pinc, linc_ptr, linc;
  // matrices
  HostSell Hpp, Hpl;
  std::vector<int32_t> hpp_diag;
  std::vector<int32_t> lp_ptr, lp_col;  // landmark-major row lists (CSR over free landmarks, observers ascending); the
                                        // landmark-major SELL copy itself is laid out per rank by the partition planner
  // ---- g2o block structure (what "symbolic structure bit-exact" is checked on)
  std::vector<int32_t> ord_kind, ord_index, ord_offset;  // per Hessian index
  std::vector<int32_t> blk_row, blk_col, blk_nr, blk_nc; // column-major, rows ascending
  // where each reference block lives on the device: kind 0 = Hpp entry, 1 = Hpl entry, 2 = Hll (index = hl)
  std::vector<int32_t> blk_kind, blk_entry;
  int64_t n_blocks = 0, block_values = 0;
  bool blocks_built = false;  // blk_* are filled on demand by build_block_list()
  // algorithmic sizes for the roofline
  int64_t n_pairs_pp = 0, n_pairs_pl = 0;   // distinct off-diagonal blocks
};

// Returns SGB_OK or an error code with a message. seq arrays may be NULL.
// world > 1 requests the rank-filtered build described at Structure::filtered: the scan of the active vertices and the
// index mapping still cover the whole graph (so Hessian indices are the reference's), everything per edge -- the bulk of
// the symbolic phase -- only covers the edges rank `rank` needs, i.e. ~1/world of the work per process.
sgb_status build_structure(const sgb_graph_soa& g, Structure& out, std::string& err, int world = 1, int rank = 0);
// free-pose rows per rank of the row-block partition (the one formula both the filter and the planner use)
inline int partition_chunk(int Pf, int world) { int c = (Pf + world - 1) / world; return c < 1 ? 1 : c; }
// fills Structure::blk_* (idempotent); only the structure / parity hooks of the C ABI need it
void build_block_list(Structure& S);

}  // namespace sgb
