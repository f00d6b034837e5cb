// sgb_posegraph.cu -- the reference's PoseGraph (include/graphs.h:27-40) kept on the device, with the edits the
// reference makes around setup_pose_opt's optimiser: state copy from the landmark graph with
// relative-pose re-measurement (submap_loop_closer.cpp:206-223), closure insertion (:272-285), optimisation (:286-288,
// log_runner.cpp:203-204) and false-closure removal (log_runner.cpp:182-190). The store keeps the VALUES in HBM
// (estimates, measurements, information, DCS deltas, in insertion order) and mirrors only the small INDEX arrays on
// the host, which is all the host symbolic phase needs (sgb_set_graph_device); a closure therefore costs the upload
// of one edge, not of the graph. There is no CPU path: the arithmetic lives in the kernels below.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sgb_capi.h"
#include "sgb_edits.h"
#include "sgb_internal.h"

using namespace sgb;

namespace {

constexpr int kPgThreads = 256;
constexpr int kChainThreads = 512;  // the single block that scans the tile products
constexpr int kTileItems = 8;       // poses per thread
constexpr int kTile = kPgThreads * kTileItems;

// Inclusive scan of one SE2 element per thread under composition (Hillis-Steele in shared memory, fixed order =>
// deterministic). Returns this thread's inclusive value; *excl (may be NULL) receives the exclusive one (identity for
// thread 0). NT = blockDim.x.
template <int NT>
__device__ __forceinline__ Se2 block_scan_se2(Se2 v, double (*sx)[NT], double (*sy)[NT], double (*sth)[NT], Se2* excl) {
  const int t = threadIdx.x;
  int cur = 0;
  sx[0][t] = v.x; sy[0][t] = v.y; sth[0][t] = v.th;
  __syncthreads();
  for (int off = 1; off < NT; off <<= 1) {
    Se2 w{sx[cur][t], sy[cur][t], sth[cur][t]};
    if (t >= off) w = se2_mul(Se2{sx[cur][t - off], sy[cur][t - off], sth[cur][t - off]}, w);
    sx[cur ^ 1][t] = w.x; sy[cur ^ 1][t] = w.y; sth[cur ^ 1][t] = w.th;
    cur ^= 1;
    __syncthreads();
  }
  if (excl) *excl = t > 0 ? Se2{sx[cur][t - 1], sy[cur][t - 1], sth[cur][t - 1]} : Se2{0.0, 0.0, 0.0};
  return Se2{sx[cur][t], sy[cur][t], sth[cur][t]};
}

// The chain copy est[k] = est[-1] * z_0 * ... * z_k, z_k = lm[k]^-1 * lm[k+1] (submap_loop_closer.cpp:206-223) is an
// inclusive scan under SE2 composition. Reduce-then-scan over tiles of kTile poses, three launches:
//   k_pg_remeasure  one thread per kTileItems consecutive poses: the re-measured z (written once) and their product;
//                   the block's ordered product goes to tile_prod[tile]
//   k_pg_scan_tiles ONE block: tile_prefix[tile] = est[-1] * prod(tile_prod[0..tile))
//   k_pg_apply      per tile: scan of the thread products in shared memory, then every thread walks its poses from its
//                   prefix and writes the estimates
// The reference composes strictly left to right; composition is associative, so the two agree up to rounding (tests:
// 1e-11 relative on the host bodies, 1e-7 m over 1e6 compositions on the device).
__global__ void __launch_bounds__(kPgThreads) k_pg_remeasure(const double* __restrict__ lm, int count, double* __restrict__ z_out,
                                                            double* __restrict__ tile_prod) {
  __shared__ double sx[2][kPgThreads], sy[2][kPgThreads], sth[2][kPgThreads];
  const int k0 = min(count, blockIdx.x * kTile + threadIdx.x * kTileItems), k1 = min(count, k0 + kTileItems);
  Se2 loc{0.0, 0.0, 0.0};
  for (int k = k0; k < k1; ++k) {
    Se2 z = relative_measurement(lm + 3 * (size_t)k, lm + 3 * (size_t)(k + 1));
    se2_store(z_out + 3 * (size_t)k, z);
    loc = se2_mul(loc, z);
  }
  Se2 incl = block_scan_se2<kPgThreads>(loc, sx, sy, sth, nullptr);
  if (threadIdx.x == kPgThreads - 1) se2_store(tile_prod + 3 * (size_t)blockIdx.x, incl);
}
__global__ void __launch_bounds__(kChainThreads) k_pg_scan_tiles(const double* __restrict__ prev, const double* __restrict__ tile_prod,
                                                                int ntiles, double* __restrict__ tile_prefix) {
  __shared__ double sx[2][kChainThreads], sy[2][kChainThreads], sth[2][kChainThreads];
  const int t = threadIdx.x;
  const int chunk = (ntiles + kChainThreads - 1) / kChainThreads;
  const int k0 = min(ntiles, t * chunk), k1 = min(ntiles, k0 + chunk);
  Se2 loc{0.0, 0.0, 0.0};
  for (int k = k0; k < k1; ++k) loc = se2_mul(loc, se2_load(tile_prod + 3 * (size_t)k));
  Se2 excl;
  block_scan_se2<kChainThreads>(loc, sx, sy, sth, &excl);
  Se2 run = se2_load(prev);  // the vertex the chain hangs from
  if (t > 0) run = se2_mul(run, excl);
  for (int k = k0; k < k1; ++k) {
    se2_store(tile_prefix + 3 * (size_t)k, run);
    run = se2_mul(run, se2_load(tile_prod + 3 * (size_t)k));
  }
}
__global__ void __launch_bounds__(kPgThreads) k_pg_apply(const double* __restrict__ z, const double* __restrict__ tile_prefix, int count,
                                                        double* __restrict__ est) {
  __shared__ double sx[2][kPgThreads], sy[2][kPgThreads], sth[2][kPgThreads];
  const int k0 = min(count, blockIdx.x * kTile + threadIdx.x * kTileItems), k1 = min(count, k0 + kTileItems);
  Se2 zz[kTileItems];
  Se2 loc{0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < kTileItems; ++i)
    if (k0 + i < k1) {
      zz[i] = se2_load(z + 3 * (size_t)(k0 + i));
      loc = se2_mul(loc, zz[i]);
    }
  Se2 excl;
  block_scan_se2<kPgThreads>(loc, sx, sy, sth, &excl);
  Se2 run = se2_load(tile_prefix + 3 * (size_t)blockIdx.x);
  if (threadIdx.x > 0) run = se2_mul(run, excl);
#pragma unroll
  for (int i = 0; i < kTileItems; ++i)
    if (k0 + i < k1) {
      run = se2_mul(run, zz[i]);
      se2_store(est + 3 * (size_t)(k0 + i), run);
    }
}

// one thread per active closure: chi2 = e^T Omega e at the current estimates (no robust kernel), flag = chi2 > thr
__global__ void __launch_bounds__(kPgThreads) k_pg_closure_chi2(const double* __restrict__ est, const int32_t* __restrict__ ei,
                                                               const int32_t* __restrict__ ej, const double* __restrict__ z,
                                                               const double* __restrict__ info, const int32_t* __restrict__ slots,
                                                               int n, double thr, double* __restrict__ chi_out,
                                                               uint8_t* __restrict__ remove_out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    size_t s = (size_t)slots[k];
    double c = pp_edge_chi2(est + 3 * (size_t)ei[s], est + 3 * (size_t)ej[s], z + 3 * s, info + 6 * s);
    chi_out[k] = c;
    remove_out[k] = c > thr ? 1 : 0;
  }
}

int pg_grid(int n) { return std::max(1, std::min((n + kPgThreads - 1) / kPgThreads, 148 * 8)); }

}  // namespace

struct sgb_pose_graph {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  // device values, capacity-grown
  double *d_est = nullptr, *d_z = nullptr, *d_info = nullptr, *d_phi = nullptr;
  int32_t *d_ei = nullptr, *d_ej = nullptr;
  int capP = 0, capE = 0;
  // scratch
  double* d_tmp = nullptr;  // staging for host estimates / closure chi2
  double* d_tiles = nullptr;  // [2][3*tiles] tile products and prefixes of the chain scan
  size_t cap_tiles = 0;
  int32_t* d_slots = nullptr;
  uint8_t* d_flags = nullptr;
  size_t cap_tmp = 0, cap_slots = 0;
  // host mirror of the indices
  std::vector<int32_t> id, ei, ej, closures;  // closures: edge slot of every closure ever added
  std::vector<uint8_t> active, is_closure;
  std::vector<double> closure_chi2;
  bool has_robust = false;
  double last_edit_ms = 0.0;
  int64_t launches = 0;
};

#define PG_CUDA(call)                                                 \
  do {                                                                \
    cudaError_t _e = (call);                                          \
    if (_e != cudaSuccess) {                                          \
      pg->err = std::string(#call) + ": " + cudaGetErrorString(_e);   \
      return SGB_ERR_CUDA;                                            \
    }                                                                 \
  } while (0)

namespace {

template <class T>
sgb_status grow(sgb_pose_graph* pg, T** p, size_t old_n, size_t new_n) {
  T* q = nullptr;
  PG_CUDA(cudaMalloc((void**)&q, std::max<size_t>(new_n, 1) * sizeof(T)));
  if (*p && old_n) PG_CUDA(cudaMemcpyAsync(q, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, pg->stream));
  PG_CUDA(cudaStreamSynchronize(pg->stream));
  if (*p) cudaFree(*p);
  *p = q;
  return SGB_OK;
}
sgb_status reserve(sgb_pose_graph* pg, int nP, int nE) {
  sgb_status st;
  if (nP > pg->capP) {
    int cap = std::max(nP, std::max(1024, 2 * pg->capP));
    if ((st = grow(pg, &pg->d_est, 3 * pg->id.size(), 3 * (size_t)cap)) != SGB_OK) return st;
    pg->capP = cap;
  }
  if (nE > pg->capE) {
    int cap = std::max(nE, std::max(1024, 2 * pg->capE));
    size_t used = pg->ei.size();
    if ((st = grow(pg, &pg->d_z, 3 * used, 3 * (size_t)cap)) != SGB_OK) return st;
    if ((st = grow(pg, &pg->d_info, 6 * used, 6 * (size_t)cap)) != SGB_OK) return st;
    if ((st = grow(pg, &pg->d_phi, used, (size_t)cap)) != SGB_OK) return st;
    if ((st = grow(pg, &pg->d_ei, used, (size_t)cap)) != SGB_OK) return st;
    if ((st = grow(pg, &pg->d_ej, used, (size_t)cap)) != SGB_OK) return st;
    pg->capE = cap;
  }
  return SGB_OK;
}
sgb_status reserve_tmp(sgb_pose_graph* pg, size_t doubles, size_t slots) {
  if (doubles > pg->cap_tmp) {
    if (pg->d_tmp) cudaFree(pg->d_tmp);
    pg->d_tmp = nullptr;
    PG_CUDA(cudaMalloc((void**)&pg->d_tmp, doubles * sizeof(double)));
    pg->cap_tmp = doubles;
  }
  if (slots > pg->cap_slots) {
    if (pg->d_slots) cudaFree(pg->d_slots);
    if (pg->d_flags) cudaFree(pg->d_flags);
    pg->d_slots = nullptr;
    pg->d_flags = nullptr;
    PG_CUDA(cudaMalloc((void**)&pg->d_slots, slots * sizeof(int32_t)));
    PG_CUDA(cudaMalloc((void**)&pg->d_flags, slots));
    pg->cap_slots = slots;
  }
  return SGB_OK;
}

// lm_dev: device pointer to [3*(count+1)] estimates, predecessor first
sgb_status append_impl(sgb_pose_graph* pg, const double* lm_dev, int count, const int32_t* ids, const double* info) {
  if (pg->id.empty()) { pg->err = "append: the store has no first vertex (call sgb_pg_reset)"; return SGB_ERR_NOT_INITIALIZED; }
  if (count <= 0) return SGB_OK;
  if (!info) { pg->err = "append: information matrices missing"; return SGB_ERR_INVALID; }
  const int p0 = (int)pg->id.size(), e0 = (int)pg->ei.size();
  sgb_status st = reserve(pg, p0 + count, e0 + count);
  if (st != SGB_OK) return st;
  std::vector<int32_t> ni(count), nj(count);
  for (int k = 0; k < count; ++k) {
    ni[k] = p0 - 1 + k;
    nj[k] = p0 + k;
  }
  const int ntiles = (count + kTile - 1) / kTile;
  if ((size_t)ntiles > pg->cap_tiles) {
    if (pg->d_tiles) cudaFree(pg->d_tiles);
    pg->d_tiles = nullptr;
    pg->cap_tiles = 0;
    PG_CUDA(cudaMalloc((void**)&pg->d_tiles, 6 * (size_t)ntiles * sizeof(double)));  // products, then prefixes
    pg->cap_tiles = (size_t)ntiles;
  }
  PG_CUDA(cudaMemcpyAsync(pg->d_info + 6 * (size_t)e0, info, 6 * (size_t)count * sizeof(double), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaMemsetAsync(pg->d_phi + e0, 0, (size_t)count * sizeof(double), pg->stream));
  PG_CUDA(cudaMemcpyAsync(pg->d_ei + e0, ni.data(), (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaMemcpyAsync(pg->d_ej + e0, nj.data(), (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaEventRecord(pg->ev0, pg->stream));
  k_pg_remeasure<<<ntiles, kPgThreads, 0, pg->stream>>>(lm_dev, count, pg->d_z + 3 * (size_t)e0, pg->d_tiles);
  k_pg_scan_tiles<<<1, kChainThreads, 0, pg->stream>>>(pg->d_est + 3 * (size_t)(p0 - 1), pg->d_tiles, ntiles, pg->d_tiles + 3 * (size_t)ntiles);
  k_pg_apply<<<ntiles, kPgThreads, 0, pg->stream>>>(pg->d_z + 3 * (size_t)e0, pg->d_tiles + 3 * (size_t)ntiles, count, pg->d_est + 3 * (size_t)p0);
  PG_CUDA(cudaGetLastError());
  PG_CUDA(cudaEventRecord(pg->ev1, pg->stream));
  PG_CUDA(cudaStreamSynchronize(pg->stream));  // ni / nj / info are the caller's and this frame's
  float ms = 0.f;
  cudaEventElapsedTime(&ms, pg->ev0, pg->ev1);
  pg->last_edit_ms = ms;
  pg->launches += 3;
  int next_id = pg->id.back() + 1;
  for (int k = 0; k < count; ++k) {
    pg->id.push_back(ids ? ids[k] : next_id + k);
    pg->ei.push_back(ni[k]);
    pg->ej.push_back(nj[k]);
    pg->active.push_back(1);
    pg->is_closure.push_back(0);
  }
  return SGB_OK;
}

}  // namespace

extern "C" {

sgb_status sgb_pg_create(int32_t device, sgb_pose_graph** out) {
  if (!out) return SGB_ERR_INVALID;
  *out = nullptr;
  if (sgb_device_count() <= 0) return SGB_ERR_NO_DEVICE;  // no CPU path
  sgb_pose_graph* pg = new sgb_pose_graph();
  auto fail = [&](cudaError_t) {
    if (pg->ev0) cudaEventDestroy(pg->ev0);
    if (pg->ev1) cudaEventDestroy(pg->ev1);
    if (pg->stream) cudaStreamDestroy(pg->stream);
    delete pg;
    return SGB_ERR_CUDA;
  };
  cudaError_t e;
  if (device >= 0 && (e = cudaSetDevice(device)) != cudaSuccess) return fail(e);
  if ((e = cudaGetDevice(&pg->device)) != cudaSuccess) return fail(e);
  if ((e = cudaStreamCreateWithFlags(&pg->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&pg->ev0)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&pg->ev1)) != cudaSuccess) return fail(e);
  *out = pg;
  return SGB_OK;
}

void sgb_pg_destroy(sgb_pose_graph* pg) {
  if (!pg) return;
  cudaSetDevice(pg->device);
  if (pg->stream) cudaStreamSynchronize(pg->stream);
  for (void* p : {(void*)pg->d_est, (void*)pg->d_z, (void*)pg->d_info, (void*)pg->d_phi, (void*)pg->d_ei, (void*)pg->d_ej,
                  (void*)pg->d_tmp, (void*)pg->d_slots, (void*)pg->d_flags, (void*)pg->d_tiles})
    if (p) cudaFree(p);
  if (pg->ev0) cudaEventDestroy(pg->ev0);
  if (pg->ev1) cudaEventDestroy(pg->ev1);
  if (pg->stream) cudaStreamDestroy(pg->stream);
  delete pg;
}

const char* sgb_pg_last_error(const sgb_pose_graph* pg) { return pg ? pg->err.c_str() : "null pose graph"; }

sgb_status sgb_pg_reset(sgb_pose_graph* pg, int32_t first_id, const double first_est[3]) {
  if (!pg || !first_est) return SGB_ERR_INVALID;
  PG_CUDA(cudaSetDevice(pg->device));
  pg->id.clear(); pg->ei.clear(); pg->ej.clear(); pg->closures.clear();
  pg->active.clear(); pg->is_closure.clear(); pg->closure_chi2.clear();
  pg->has_robust = false;
  sgb_status st = reserve(pg, 1, 0);
  if (st != SGB_OK) return st;
  PG_CUDA(cudaMemcpyAsync(pg->d_est, first_est, 3 * sizeof(double), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaStreamSynchronize(pg->stream));
  pg->id.push_back(first_id);
  return SGB_OK;
}

sgb_status sgb_pg_append_from_lm(sgb_pose_graph* pg, sgb_handle* lm, int32_t lm_first, int32_t count, const int32_t* ids,
                                 const double* info) {
  if (!pg || !lm) return SGB_ERR_INVALID;
  PG_CUDA(cudaSetDevice(pg->device));
  HandleView v;
  if (!handle_view(lm, &v)) { pg->err = "append: the landmark-graph handle has no graph"; return SGB_ERR_NOT_INITIALIZED; }
  if (v.device != pg->device) { pg->err = "append: the landmark graph lives on another device"; return SGB_ERR_INVALID; }
  if (lm_first < 1 || count < 0 || lm_first + count > v.n_poses) { pg->err = "append: pose range outside the landmark graph"; return SGB_ERR_INVALID; }
  PG_CUDA(cudaStreamSynchronize(v.stream));  // the landmark graph's optimise has finished writing its estimates
  return append_impl(pg, v.pose_est + 3 * (size_t)(lm_first - 1), count, ids, info);
}

sgb_status sgb_pg_append_from_host(sgb_pose_graph* pg, const double* lm_est, int32_t count, const int32_t* ids,
                                   const double* info) {
  if (!pg || (count > 0 && !lm_est) || count < 0) return SGB_ERR_INVALID;
  PG_CUDA(cudaSetDevice(pg->device));
  sgb_status st = reserve_tmp(pg, 3 * (size_t)(count + 1), 0);
  if (st != SGB_OK) return st;
  PG_CUDA(cudaMemcpyAsync(pg->d_tmp, lm_est, 3 * (size_t)(count + 1) * sizeof(double), cudaMemcpyHostToDevice, pg->stream));
  return append_impl(pg, pg->d_tmp, count, ids, info);
}

sgb_status sgb_pg_add_closure(sgb_pose_graph* pg, int32_t from, int32_t to, const double z[3], const double info[6],
                              double dcs_phi, int32_t* closure_index) {
  if (!pg || !z || !info) return SGB_ERR_INVALID;
  const int P = (int)pg->id.size();
  if (from < 0 || from >= P || to < 0 || to >= P || from == to) { pg->err = "closure: bad vertex"; return SGB_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(pg->device));
  const int e0 = (int)pg->ei.size();
  sgb_status st = reserve(pg, P, e0 + 1);
  if (st != SGB_OK) return st;
  PG_CUDA(cudaMemcpyAsync(pg->d_z + 3 * (size_t)e0, z, 3 * sizeof(double), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaMemcpyAsync(pg->d_info + 6 * (size_t)e0, info, 6 * sizeof(double), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaMemcpyAsync(pg->d_phi + e0, &dcs_phi, sizeof(double), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaMemcpyAsync(pg->d_ei + e0, &from, sizeof(int32_t), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaMemcpyAsync(pg->d_ej + e0, &to, sizeof(int32_t), cudaMemcpyHostToDevice, pg->stream));
  PG_CUDA(cudaStreamSynchronize(pg->stream));
  pg->ei.push_back(from);
  pg->ej.push_back(to);
  pg->active.push_back(1);
  pg->is_closure.push_back(1);
  if (closure_index) *closure_index = (int32_t)pg->closures.size();
  pg->closures.push_back(e0);
  pg->closure_chi2.push_back(0.0);
  if (dcs_phi > 0.0) pg->has_robust = true;
  pg->last_edit_ms = 0.0;
  return SGB_OK;
}

sgb_status sgb_pg_optimize(sgb_pose_graph* pg, sgb_handle* solver, int32_t algo, int32_t max_iters, int32_t* iters_done,
                           sgb_iter_stat* stats) {
  if (iters_done) *iters_done = -1;
  if (!pg || !solver) return SGB_ERR_INVALID;
  if (handle_device(solver) != pg->device) { pg->err = "optimize: the solver handle lives on another device than the store"; return SGB_ERR_INVALID; }
  PG_CUDA(cudaSetDevice(pg->device));
  const int P = (int)pg->id.size();
  std::vector<int32_t> ci, cj, slot;
  ci.reserve(pg->ei.size()); cj.reserve(pg->ei.size()); slot.reserve(pg->ei.size());
  for (size_t s = 0; s < pg->ei.size(); ++s)
    if (pg->active[s]) {
      ci.push_back(pg->ei[s]);
      cj.push_back(pg->ej[s]);
      slot.push_back((int32_t)s);
    }
  std::vector<uint8_t> fixed(P, 0);
  if (P) fixed[0] = 1;  // drone.cpp:75
  sgb_graph_soa g;
  std::memset(&g, 0, sizeof g);
  g.n_poses = P;
  g.pose_id = pg->id.data();
  g.pose_fixed = fixed.data();
  g.n_pp = (int32_t)ci.size();
  g.pp_i = ci.data();
  g.pp_j = cj.data();
  sgb_device_values dv;
  std::memset(&dv, 0, sizeof dv);
  dv.pose_est = pg->d_est;
  dv.pp_z = pg->d_z;
  dv.pp_info = pg->d_info;
  dv.pp_phi = pg->d_phi;
  dv.pp_slot = slot.data();
  dv.has_robust = pg->has_robust ? 1 : 0;
  PG_CUDA(cudaStreamSynchronize(pg->stream));  // every edit is complete before the solver's stream reads the store
  sgb_status st = sgb_set_graph_device(solver, &g, &dv);
  if (st != SGB_OK) { pg->err = std::string("optimize: ") + sgb_last_error(solver); return st; }
  st = sgb_optimize(solver, algo, max_iters, 0, iters_done, stats);
  if (st != SGB_OK) { pg->err = std::string("optimize: ") + sgb_last_error(solver); return st; }
  HandleView v;
  if (!handle_view(solver, &v) || v.device != pg->device) { pg->err = "optimize: solver handle on another device"; return SGB_ERR_INVALID; }
  PG_CUDA(cudaMemcpyAsync(pg->d_est, v.pose_est, 3 * (size_t)P * sizeof(double), cudaMemcpyDeviceToDevice, v.stream));
  PG_CUDA(cudaStreamSynchronize(v.stream));
  return SGB_OK;
}

sgb_status sgb_pg_prune_closures(sgb_pose_graph* pg, double threshold, int32_t* n_removed, double* chi2_out,
                                 uint8_t* active_out) {
  if (!pg) return SGB_ERR_INVALID;
  PG_CUDA(cudaSetDevice(pg->device));
  std::vector<int32_t> slots, which;
  for (size_t c = 0; c < pg->closures.size(); ++c)
    if (pg->active[pg->closures[c]]) {
      slots.push_back(pg->closures[c]);
      which.push_back((int32_t)c);
    }
  const int n = (int)slots.size();
  int removed = 0;
  pg->last_edit_ms = 0.0;
  if (n > 0) {
    sgb_status st = reserve_tmp(pg, (size_t)n, (size_t)n);
    if (st != SGB_OK) return st;
    PG_CUDA(cudaMemcpyAsync(pg->d_slots, slots.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, pg->stream));
    PG_CUDA(cudaEventRecord(pg->ev0, pg->stream));
    k_pg_closure_chi2<<<pg_grid(n), kPgThreads, 0, pg->stream>>>(pg->d_est, pg->d_ei, pg->d_ej, pg->d_z, pg->d_info, pg->d_slots, n,
                                                                threshold, pg->d_tmp, pg->d_flags);
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaEventRecord(pg->ev1, pg->stream));
    std::vector<double> chi(n);
    std::vector<uint8_t> flags(n);
    PG_CUDA(cudaMemcpyAsync(chi.data(), pg->d_tmp, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, pg->stream));
    PG_CUDA(cudaMemcpyAsync(flags.data(), pg->d_flags, (size_t)n, cudaMemcpyDeviceToHost, pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, pg->ev0, pg->ev1);
    pg->last_edit_ms = ms;
    pg->launches += 1;
    for (int k = 0; k < n; ++k) {
      pg->closure_chi2[which[k]] = chi[k];
      if (flags[k]) {  // opt.removeEdge + closures.erase + false_closures.insert
        pg->active[slots[k]] = 0;
        ++removed;
      }
    }
  }
  if (n_removed) *n_removed = removed;
  for (size_t c = 0; c < pg->closures.size(); ++c) {
    if (chi2_out) chi2_out[c] = pg->closure_chi2[c];
    if (active_out) active_out[c] = pg->active[pg->closures[c]];
  }
  return SGB_OK;
}

sgb_status sgb_pg_get_info(const sgb_pose_graph* pg, sgb_pg_info* out) {
  if (!pg || !out) return SGB_ERR_INVALID;
  out->n_poses = (int32_t)pg->id.size();
  out->n_edges = (int32_t)pg->ei.size();
  out->n_closures = (int32_t)pg->closures.size();
  int a = 0;
  for (int32_t s : pg->closures) a += pg->active[s];
  out->n_active_closures = a;
  out->last_edit_ms = pg->last_edit_ms;
  out->kernel_launches = pg->launches;
  return SGB_OK;
}

sgb_status sgb_pg_download(sgb_pose_graph* pg, int32_t* pose_id, double* pose_est, int32_t* pp_i, int32_t* pp_j, double* pp_z,
                           double* pp_info, double* pp_phi, uint8_t* pp_active, uint8_t* pp_is_closure) {
  if (!pg) return SGB_ERR_INVALID;
  PG_CUDA(cudaSetDevice(pg->device));
  const size_t P = pg->id.size(), E = pg->ei.size();
  if (pose_id && P) std::memcpy(pose_id, pg->id.data(), P * sizeof(int32_t));
  if (pp_i && E) std::memcpy(pp_i, pg->ei.data(), E * sizeof(int32_t));
  if (pp_j && E) std::memcpy(pp_j, pg->ej.data(), E * sizeof(int32_t));
  if (pp_active && E) std::memcpy(pp_active, pg->active.data(), E);
  if (pp_is_closure && E) std::memcpy(pp_is_closure, pg->is_closure.data(), E);
  if (pose_est && P) PG_CUDA(cudaMemcpyAsync(pose_est, pg->d_est, 3 * P * sizeof(double), cudaMemcpyDeviceToHost, pg->stream));
  if (pp_z && E) PG_CUDA(cudaMemcpyAsync(pp_z, pg->d_z, 3 * E * sizeof(double), cudaMemcpyDeviceToHost, pg->stream));
  if (pp_info && E) PG_CUDA(cudaMemcpyAsync(pp_info, pg->d_info, 6 * E * sizeof(double), cudaMemcpyDeviceToHost, pg->stream));
  if (pp_phi && E) PG_CUDA(cudaMemcpyAsync(pp_phi, pg->d_phi, E * sizeof(double), cudaMemcpyDeviceToHost, pg->stream));
  PG_CUDA(cudaStreamSynchronize(pg->stream));
  return SGB_OK;
}

}  // extern "C"
