// sgb_partition.h -- row-block partition of the symbolic structure across the GPUs of one box.
//
// Rank r owns the free-pose rows [r*chunkP, (r+1)*chunkP) and every landmark whose first (lowest-index) free
// observer it owns. A rank keeps: the Hessian rows it owns (SELL matrices with ENCODED columns, see sgb_types.h),
// the edges incident to those rows, and their incidence lists. world == 1 reproduces the global structure.
// Host-only, no CUDA. The reference has no counterpart (single process, single thread).
#pragma once
#include <vector>

#include "sgb_structure.h"

namespace sgb {

struct LocalPlan {
  int world = 1, rank = 0;
  int chunkP = 0;            // pose rows per rank (last rank may own fewer)
  int nP = 0, nL = 0, capP = 0, capL = 0;
  int nL_owned = 0;          // the first nL_owned landmark rows are owned; the rest are ghost copies (see below)
  int p_begin = 0;           // first global free pose owned
  int n_pp = 0, n_pl = 0, n_pp_owned = 0, n_pl_owned = 0;
  std::vector<int32_t> pose_of_l, lm_of_l;       // local row -> vertex array index
  std::vector<int32_t> lm_global;                // local landmark -> global free landmark
  std::vector<int32_t> enc_lm;                   // [Lf] global free landmark -> encoded (owner, local)
  std::vector<int32_t> enc_lm_here;              // [Lf] the same as THIS rank's matrices encode it: its own row when it
                                                 //      keeps the landmark (owned or ghost), the owner's otherwise
  std::vector<int32_t> pp_g, pl_g;               // local edge -> global ACTIVE edge index (Structure order)
  std::vector<int32_t> pp_i, pp_j, pp_hi, pp_hj, pp_e_ij, pp_e_ji, pp_dup;
  std::vector<int32_t> pl_p, pl_l, pl_hp, pl_hl, pl_e_pl, pl_e_lp, pl_dup;
  std::vector<int32_t> pinc_ptr, pinc, linc_ptr, linc;
  HostSell Hpp, Hpl, Hlp;
  std::vector<int32_t> hpp_diag;
  // reference-order export: for every block of Structure::blk_* the owner rank and the local entry
  std::vector<int32_t> blk_owner, blk_entry;
  bool export_built = false;  // filled on demand by build_export()
  // halo statistics (distinct remote entries this rank gathers per PCG iteration)
  int64_t halo_p = 0, halo_t = 0;
  // ---- pushed pose halos (ghost-row layout, world > 1). The pose-vector entries other ranks need (operator input z and
  // step x of boundary rows) are WRITTEN into the consumers' memory by their owner at the end of the vector-update phase
  // instead of being gathered through NVLink in the two operator passes: every column of the local matrices that names
  // a pose of another rank is re-encoded as (this rank, capP + slot) and reads a halo copy kept behind the rank's own
  // rows. halo_src = the global free-pose indices of the slots, ascending, i.e. grouped by source rank:
  // [halo_base[o], halo_base[o] + halo_cnt[o]) come from rank o. The sender side follows from the symmetry of the
  // structure (Hpp is symmetric; a landmark row lists all observers): rank r pushes its row i to rank q iff row i has a
  // pose-pose block with a row of q, or observes a landmark that some pose of q observes. send_dst[send_ptr[l] ..
  // send_ptr[l+1]) = (q << kOwnerShift) | k: local row l is the k-th of this rank's rows in q's halo. Both sides sort by
  // global index, so the k-th pushed row lands in slot halo_base_q[r] + k; the counts are cross-checked at connect time.
  bool pushed = false;
  int nH = 0;
  std::vector<int32_t> halo_src;
  int32_t halo_base[8] = {0, 0, 0, 0, 0, 0, 0, 0}, halo_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<int32_t> send_ptr, send_dst;
  int32_t send_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // rows this rank pushes to every other rank
};

// Ghost landmarks (world > 1, opt-in): a rank also keeps a full copy of the row of every landmark its poses observe
// but another rank owns -- all its edges (linearised locally from the replicated estimates), its Hll / W / b_l / t.
// The pose-major pass then never reads a landmark quantity of another rank, and the barrier after the landmark pass
// of the PCG only has to span the GPU. Updates, chi2 and the exported system stay with the owner.
// Default: on (validated on 2 and 8 GPUs, round 2); SGB_GHOST_LANDMARKS=0 in the environment or
// partition_use_ghost_landmarks(false) select the owner-only layout, where the pose-major pass gathers t / W / b_l of
// remote landmarks through NVLink.
void partition_use_ghost_landmarks(bool on);
bool partition_ghost_landmarks();
// Pushed pose halos (LocalPlan::pushed; needs ghost rows). Default on; SGB_PUSHED_HALOS=0 keeps the NVLink gathers.
void partition_use_pushed_halos(bool on);
bool partition_pushed_halos();

// Coarse space of the two-level preconditioner (sgb_coarse.h): gather lists for R S R^T over the rows of ONE rank
// (world == 1). nn = 0: not planned.
struct CoarsePlan {
  int h = 0, nn = 0, ng = 0;
  std::vector<int32_t> g_ptr, g_e, g_lm, p_ptr, p_e, t_ptr, t_g;
  std::vector<double> g_w, p_w, rr;
};
// node spacing for a graph of nP pose rows: the smallest of 8 / 16 / 32 / 64 that needs at most max_nodes nodes
// (<= kCzMaxNodes, sgb_coarse.h); 0 = the graph is too long for a dense coarse solve (or too short to need one)
int coarse_spacing(int nP, int max_nodes);
void plan_coarse(const LocalPlan& P, int h, CoarsePlan& C);

inline int enc_pose(int hp, int chunkP) { return ((hp / chunkP) << 26) | (hp % chunkP); }

sgb_status partition(const Structure& S, int world, int rank, LocalPlan& out, std::string& err);
// the same for a Structure nobody else reads afterwards: with world == 1 its per-edge arrays (pp_*, pl_*, pinc*) are
// moved into the plan and left empty in S (pp_src / pl_src, the SELL patterns and the index maps stay)
sgb_status partition_consume(Structure& S, int world, int rank, LocalPlan& out, std::string& err);
// fills Structure::blk_* and LocalPlan::blk_owner / blk_entry (idempotent); only the parity hooks need it
void build_export(Structure& S, LocalPlan& P);

}  // namespace sgb
