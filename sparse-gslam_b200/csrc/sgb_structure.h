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

struct HostSell {
  int rows = 0, nslices = 0;
  std::vector<int32_t> sbase;  // [nslices + 1]
  std::vector<int32_t> col;    // [entries]
  int64_t entries() const { return (int64_t)col.size(); }
};

struct Structure {
  int P_all = 0, L_all = 0, Pf = 0, Lf = 0, n_pp = 0, n_pl = 0, dim = 0;
  bool has_robust = false;
  // vertex maps
  std::vector<int32_t> pose_h, lm_h;        // array index -> free index (-1 fixed / inactive)
  std::vector<int32_t> pose_of_h, lm_of_h;  // free index -> array index
  // active edges in insertion order: index into the caller's arrays
  std::vector<int32_t> pp_src, pl_src;
  // per active edge
  std::vector<int32_t> pp_i, pp_j, pp_hi, pp_hj, pp_e_ij, pp_e_ji, pp_dup;
  std::vector<int32_t> pl_p, pl_l, pl_hp, pl_hl, pl_e_pl, pl_e_lp, pl_dup;
  // incidence
  std::vector<int32_t> pinc_ptr, pinc, linc_ptr, linc;
  // matrices
  HostSell Hpp, Hpl, Hlp;
  std::vector<int32_t> hpp_diag, lp_row2h, lp_h2row;
  // ---- g2o block structure (what "symbolic structure bit-exact" is checked on)
  std::vector<int32_t> ord_kind, ord_index, ord_offset;  // per Hessian index
  std::vector<int32_t> blk_row, blk_col, blk_nr, blk_nc; // column-major, rows ascending
  // where each reference block lives on the device: kind 0 = Hpp entry, 1 = Hpl entry, 2 = Hll (index = hl)
  std::vector<int32_t> blk_kind, blk_entry;
  int64_t block_values = 0;
  // algorithmic sizes for the roofline
  int64_t n_pairs_pp = 0, n_pairs_pl = 0;   // distinct off-diagonal blocks
};

// Returns SGB_OK or an error code with a message. seq arrays may be NULL.
sgb_status build_structure(const sgb_graph_soa& g, Structure& out, std::string& err);

}  // namespace sgb
