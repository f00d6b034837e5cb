// sgb_backend.cu -- handle, device memory, host orchestration of the LM / GN iteration and the C ABI
// (include/sgb_capi.h). There is no CPU path in this file: every compute entry point needs a CUDA device.
//
// Control flow restated from the un-vendored g2o:
//   SparseOptimizer::optimize -> OptimizationAlgorithm{Levenberg,GaussNewton}::solve, called by the reference at
//   src/sparse_gslam/src/drone.cpp:150,155, submap_loop_closer.cpp:287, log_runner.cpp:204.
#include <cuda_runtime.h>
#include <malloc.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sgb_capi.h"
#include "sgb_internal.h"
#include "sgb_kernels.cuh"
#include "sgb_partition.h"
#include "sgb_resident.cuh"
#include "sgb_structure.h"

using namespace sgb;

namespace {

struct PhaseEvents {
  cudaEvent_t e[5];
};

}  // namespace

struct sgb_handle {
  sgb_options opt;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  bool has_graph = false;
  Structure S;
  LocalPlan LP;
  DevGraph G;
  std::vector<void*> allocs;  // individually cudaMalloc'ed: the IPC-exported arena (world > 1) and push/pop backups
  // Slab pool for everything else: sgb_set_graph is called once per key-frame by the reference's caller
  // (drone.cpp:146-156), so the device memory of the previous graph is recycled instead of freed and re-allocated
  // (cudaFree/cudaMalloc of ~60 arrays cost up to 0.5 s on a 1M-pose graph).
  struct Slab { char* base; size_t cap, off; };
  std::vector<Slab> slabs;
  // pinned staging ring for host -> device uploads of pageable caller / planner memory (a pageable cudaMemcpy
  // runs at a fraction of the PCIe rate); the host copies chunk k+1 while the DMA engine moves chunk k
  static constexpr int kStages = 4;
  static constexpr size_t kStageBytes = (size_t)8 << 20;
  // component-major edge data assembled on the host before its upload; kept between calls so that a re-initialised
  // graph of similar size does not page-fault ~300 MB of fresh allocations again
  std::vector<double> e_zinv, e_info, e_phi, e_z, e_linfo;
  char* stage[kStages] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[kStages] = {nullptr, nullptr, nullptr, nullptr};
  int stage_cur = 0;
  // peer-mapped arena: estimates (2 buffers), p, x_p, t, b_l, Hll_inv, mailbox -- same offsets on every rank
  char* arena = nullptr;
  size_t arena_bytes = 0;
  // Layout of a rank's arena, kept in its first bytes so that peers can read it after opening the IPC handle: array
  // sizes that depend on rank-local quantities (halo slots, ghost rows) then need no agreement between the ranks.
  struct ArenaHeader {
    unsigned long long magic;
    unsigned long long off_pose[2], off_lm[2], off_p, off_xp, off_mbox, off_t, off_bl, off_hllinv;
    int32_t halo_base[kMaxRanks], halo_cnt[kMaxRanks];  // this rank's halo: slots [base[o], base[o] + cnt[o]) come from rank o
  };
  static constexpr unsigned long long kArenaMagic = 0x5347424152454e41ull;  // "SGBARENA"
  ArenaHeader layout[kMaxRanks];  // [rank] = this rank's own layout; peers' filled by sgb_comm_connect
  void* peer_base[kMaxRanks] = {nullptr};
  // Multi-GPU arenas are recycled like the rest of the graph memory (the reference re-initialises once per key-frame):
  // the exported allocation only grows, its IPC handle is exported once, and a peer mapping is kept open for as long as
  // the peer keeps presenting the same handle -- cudaMalloc / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle / cudaFree of
  // 8 ranks x 7 peers were a third of the end-to-end step at 8 GPUs.
  char* arena_raw = nullptr;
  size_t arena_raw_cap = 0;
  bool arena_exported = false;
  cudaIpcMemHandle_t arena_handle;
  void* peer_map[kMaxRanks] = {nullptr};          // open IPC mappings (survive sgb_set_graph)
  cudaIpcMemHandle_t peer_handle[kMaxRanks];      // the handle each mapping was opened from
  bool connected = false;
  DevScalars* d_sc = nullptr;
  DevScalars* h_sc = nullptr;  // pinned
  double* d_part_p = nullptr;
  double* d_part_l = nullptr;
  double* d_part_e = nullptr;
  unsigned long long* d_bar = nullptr;
  double* d_pose0 = nullptr;
  double* d_lm0 = nullptr;
  std::vector<std::pair<double*, double*>> stack;  // SparseOptimizer::push/pop backups (device)
  int pcg_blocks = 1;
  int pcg_threads = kThreads;  // threads per CTA of the persistent PCG kernel (256 / 288 / 320, see k_pcg)
  bool no_small = std::getenv("SGB_NO_SMALL") != nullptr;  // tuning runs: always the throughput build of k_pcg
  int pcg_cluster = 0;  // > 0: the PCG grid is one thread-block cluster of this many CTAs (small graph, one GPU)
  ResPlan res;          // valid: the graph fits one cluster's shared memory -> the cluster-resident solve (sgb_resident.cuh)
  ResPlan res_block;    // valid: the graph fits ONE 256-thread CTA -> the resident solve inside the batched kernel
  ResPlan res4;         // valid: the resident solve with four lanes per pose row (one CTA of up to 1024 threads, or a cluster)
  CoarsePlan cz;        // host gather lists of the two-level preconditioner's coarse matrix (res4.cz_nc > 0; sgb_coarse.h)
  // LinearSolver-level entry (sgb_linear_set_pattern / sgb_linear_solve): per input block its value offset, kind and
  // the SELL entries it lands in; device copies live in the pooled memory of the current graph
  struct LinearMap {
    bool valid = false;
    int n_blocks = 0, n3 = 0, n2 = 0;
    int64_t n_values = 0;
    const int32_t *d_off = nullptr, *d_e1 = nullptr, *d_e2 = nullptr, *d_kind = nullptr, *d_lmg = nullptr;
    double *d_vals = nullptr, *d_b = nullptr;
  } lin;
  // ---- the online path (sgb_update_graph, opt.incremental): index mirror of the graph on the host, raw values on the
  // device in grow-only arrays of their own (the pooled memory is recycled by every re-planning)
  struct Online {
    bool valid = false;
    std::vector<int32_t> pose_id, lm_id, pp_i, pp_j, pl_pose, pl_lm;
    std::vector<uint8_t> pose_fixed, lm_fixed;
    std::vector<int64_t> pp_seq, pl_seq;
    int64_t next_seq = 0;
    bool any_phi = false;
    double *d_pose = nullptr, *d_lm = nullptr, *d_pp_z = nullptr, *d_pp_info = nullptr, *d_pp_phi = nullptr, *d_pl_z = nullptr,
           *d_pl_info = nullptr;
    size_t cap_pose = 0, cap_lm = 0, cap_pp = 0, cap_pl = 0;  // in vertices / edges
  } on;
  sgb_timings tm;
  PhaseEvents ev;
  bool lm_state_valid = false;
};

#define SGB_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      h->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
      return SGB_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

namespace {

int grid_for(int n) {
  int b = (n + kThreads - 1) / kThreads;
  return std::max(1, std::min(b, 148 * 8));
}
// k_lin_lm: one warp per slice of the landmark-major matrix; k_lin_pose: one thread per pose row
int lin_lm_grid(const DevGraph& G) { return grid_for(32 * G.Hlp.nslices); }
int lin_pose_grid(const DevGraph& G) { return grid_for(G.nP); }

// individually allocated (and individually freed) device memory
template <class T>
sgb_status dalloc_raw(sgb_handle* h, T** out, size_t n) {
  *out = nullptr;
  void* p = nullptr;
  size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  SGB_CUDA(cudaMalloc(&p, bytes));
  h->allocs.push_back(p);
  *out = (T*)p;
  return SGB_OK;
}
// pooled device memory, recycled by the next sgb_set_graph (256-byte aligned)
template <class T>
sgb_status dalloc(sgb_handle* h, T** out, size_t n) {
  *out = nullptr;
  size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
  for (auto& sl : h->slabs)
    if (sl.cap - sl.off >= bytes) {
      *out = (T*)(sl.base + sl.off);
      sl.off += bytes;
      return SGB_OK;
    }
  size_t total = 0;
  for (auto& sl : h->slabs) total += sl.cap;
  size_t cap = std::max(bytes, std::max<size_t>(total / 2, (size_t)1 << 20));
  void* p = nullptr;
  SGB_CUDA(cudaMalloc(&p, cap));
  h->slabs.push_back({(char*)p, cap, bytes});
  *out = (T*)p;
  return SGB_OK;
}
// host (pageable) -> device on the handle's stream, through the pinned staging ring
sgb_status h2d(sgb_handle* h, void* dst, const void* src, size_t bytes) {
  if (bytes < ((size_t)256 << 10)) {
    SGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return SGB_OK;
  }
  if (!h->stage[0]) {  // first large upload of this handle (handles of small graphs never pin memory)
    for (int b = 0; b < sgb_handle::kStages; ++b) {
      SGB_CUDA(cudaMallocHost((void**)&h->stage[b], sgb_handle::kStageBytes));
      SGB_CUDA(cudaEventCreateWithFlags(&h->stage_ev[b], cudaEventDisableTiming));
    }
  }
  for (size_t off = 0; off < bytes; off += sgb_handle::kStageBytes) {
    size_t n = std::min(sgb_handle::kStageBytes, bytes - off);
    int b = h->stage_cur;
    h->stage_cur = (b + 1) % sgb_handle::kStages;
    SGB_CUDA(cudaEventSynchronize(h->stage_ev[b]));  // the previous transfer out of this buffer has finished
    if (n >= ((size_t)4 << 20)) {  // a single thread copies at ~7 GB/s, the link takes several times that: split the chunk
      const size_t half = n / 2;
      std::thread helper([&]() { std::memcpy(h->stage[b] + half, (const char*)src + off + half, n - half); });
      std::memcpy(h->stage[b], (const char*)src + off, half);
      helper.join();
    } else {
      std::memcpy(h->stage[b], (const char*)src + off, n);
    }
    SGB_CUDA(cudaMemcpyAsync((char*)dst + off, h->stage[b], n, cudaMemcpyHostToDevice, h->stream));
    SGB_CUDA(cudaEventRecord(h->stage_ev[b], h->stream));
  }
  return SGB_OK;
}
template <class T>
sgb_status upload(sgb_handle* h, const T** out, const std::vector<T>& v) {
  T* d = nullptr;
  sgb_status s = dalloc(h, &d, v.size());
  if (s != SGB_OK) return s;
  if (!v.empty() && (s = h2d(h, d, v.data(), v.size() * sizeof(T))) != SGB_OK) return s;
  *out = d;
  return SGB_OK;
}
sgb_status upload_sell(sgb_handle* h, Sell* out, const HostSell& s, int NC) {
  out->rows = s.rows;
  out->nslices = s.nslices;
  sgb_status st = upload(h, &out->sbase, s.sbase);
  if (st != SGB_OK) return st;
  st = upload(h, &out->col, s.col);
  if (st != SGB_OK) return st;
  out->srow = nullptr;
  out->sshift = nullptr;
  if (s.grouped()) {
    if ((st = upload(h, &out->srow, s.srow)) != SGB_OK) return st;
    if ((st = upload(h, &out->sshift, s.sshift)) != SGB_OK) return st;
  }
  st = dalloc(h, &out->vals, (size_t)s.entries() * NC);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaMemsetAsync(out->vals, 0, std::max<size_t>((size_t)s.entries() * NC, 1) * sizeof(double), h->stream));
  return SGB_OK;
}

void close_peers(sgb_handle* h) {
  for (int r = 0; r < kMaxRanks; ++r) {
    if (h->peer_map[r]) cudaIpcCloseMemHandle(h->peer_map[r]);
    h->peer_map[r] = nullptr;
  }
}
void free_graph(sgb_handle* h) {
  for (int r = 0; r < kMaxRanks; ++r) h->peer_base[r] = nullptr;  // the mappings themselves stay open (peer_map)
  h->connected = false;
  h->arena = nullptr;
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
  for (auto& sl : h->slabs) sl.off = 0;
  h->stack.clear();
  h->has_graph = false;
  h->lm_state_valid = false;
}

sgb_status need_graph(sgb_handle* h) {
  if (!h) return SGB_ERR_INVALID;
  if (!h->has_graph) {
    h->err = "0 vertices to optimize, maybe forgot to call initializeOptimization()";
    return SGB_ERR_NOT_INITIALIZED;
  }
  if (h->LP.world > 1 && !h->connected) {
    h->err = "partitioned graph: call sgb_comm_connect() with every rank's handle first";
    return SGB_ERR_COMM;
  }
  return SGB_OK;
}

// ---- launches (all on the handle's stream) --------------------------------------------------------------
sgb_status launch_linearize(sgb_handle* h) {
  DevGraph& G = h->G;
  int gp = lin_pose_grid(G), gl = lin_lm_grid(G);
  if (G.nP > 0) k_lin_pose<<<gp, kThreads, 0, h->stream>>>(G, h->d_part_p);
  if (G.nL > 0) k_lin_lm<<<gl, kThreads, 0, h->stream>>>(G, h->d_part_l);
  h->tm.kernel_launches += (G.nP > 0) + (G.nL > 0);
  h->tm.linearizations++;
  SGB_CUDA(cudaGetLastError());
  return SGB_OK;
}
sgb_status launch_finalize_lin(sgb_handle* h, int init_lambda) {
  DevGraph& G = h->G;
  double tau = h->opt.lm_tau > 0 ? h->opt.lm_tau : 1e-5;
  k_finalize_lin<<<1, kThreads, 0, h->stream>>>(G, h->d_sc, h->d_part_p, G.nP > 0 ? lin_pose_grid(G) : 0, h->d_part_l,
                                                G.nL > 0 ? lin_lm_grid(G) : 0, init_lambda, tau, h->opt.lm_user_lambda);
  h->tm.kernel_launches++;
  SGB_CUDA(cudaGetLastError());
  return SGB_OK;
}
sgb_status launch_setup(sgb_handle* h, double lambda_override, int use_override) {
  DevGraph& G = h->G;
  if (G.nL > 0) k_setup_lm<<<grid_for(G.nL), kThreads, 0, h->stream>>>(G, h->d_sc, lambda_override, use_override);
  if (G.world > 1) {  // pose rows read (Hll + lambda I)^-1 of landmarks owned by other ranks
    k_xbarrier<<<1, 32, 0, h->stream>>>(G, h->d_sc);
    h->tm.kernel_launches++;
  }
  if (G.nP > 0) {
    const int nch = (G.nP + kChunk - 1) / kChunk;
    k_setup_chunk<<<std::max(1, std::min((nch + 127) / 128, 148 * 16)), 128, 0, h->stream>>>(G, h->d_sc, lambda_override, use_override);
    k_setup_pose<<<grid_for(G.nP), kThreads, 0, h->stream>>>(G, h->d_sc, lambda_override, use_override);
  }
  h->tm.kernel_launches += 2 * (G.nP > 0) + (G.nL > 0);
  if (G.cz_h > 0) {  // coarse matrix of the two-level preconditioner -> its inverse (needs k_setup_lm's (Hll + lambda I)^-1)
    const int nc = 3 * G.cz_nn;
    k_setup_coarse<<<1, 1024, ((size_t)nc * cz_ld(nc) + nc) * sizeof(double), h->stream>>>(G, h->d_sc, lambda_override, use_override);
    h->tm.kernel_launches++;
  }
  SGB_CUDA(cudaGetLastError());
  return SGB_OK;
}
sgb_status launch_pcg(sgb_handle* h, double lambda_override, int use_override) {
  DevGraph G = h->G;
  DevScalars* sc = h->d_sc;
  double* part = h->d_part_e;
  unsigned long long* bar = h->d_bar;
  PcgParams prm;
  prm.tol = h->opt.pcg_tolerance > 0 ? h->opt.pcg_tolerance : 1e-10;
  prm.maxit = h->opt.pcg_max_iters > 0 ? h->opt.pcg_max_iters : std::max(100, 4 * 3 * h->S.Pf);
  prm.lambda_override = lambda_override;
  prm.use_override = use_override;
  SGB_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned long long), h->stream));
  static const bool no_res1 = std::getenv("SGB_NO_RESIDENT1") != nullptr;
  if (h->res4.valid && G.world == 1) {  // four lanes per pose row, everything on chip
    ResPlan rp = h->res4;
    const bool cz = rp.cz_nc > 0 && G.cz_h > 0 && G.cz_A != nullptr;
    if (rp.ncta == 1) {
      if (rp.bt <= 512) {
        if (cz) k_pcg_res4<512, 4, true><<<1, rp.bt, (size_t)rp.bytes, h->stream>>>(G, sc, prm, rp);
        else k_pcg_res4<512, 4, false><<<1, rp.bt, (size_t)rp.bytes, h->stream>>>(G, sc, prm, rp);
      } else {
        if (cz) k_pcg_res4<1024, 2, true><<<1, rp.bt, (size_t)rp.bytes, h->stream>>>(G, sc, prm, rp);
        else k_pcg_res4<1024, 2, false><<<1, rp.bt, (size_t)rp.bytes, h->stream>>>(G, sc, prm, rp);
      }
      SGB_CUDA(cudaGetLastError());
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(rp.ncta);
      cfg.blockDim = dim3(rp.bt);
      cfg.dynamicSmemBytes = (size_t)rp.bytes;
      cfg.stream = h->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = rp.ncta;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (rp.bt <= 512) {
        if (cz) SGB_CUDA(cudaLaunchKernelEx(&cfg, k_pcg_res4<512, 4, true>, G, sc, prm, rp));
        else SGB_CUDA(cudaLaunchKernelEx(&cfg, k_pcg_res4<512, 4, false>, G, sc, prm, rp));
      } else {
        if (cz) SGB_CUDA(cudaLaunchKernelEx(&cfg, k_pcg_res4<1024, 2, true>, G, sc, prm, rp));
        else SGB_CUDA(cudaLaunchKernelEx(&cfg, k_pcg_res4<1024, 2, false>, G, sc, prm, rp));
      }
    }
  } else if (h->res_block.valid && G.world == 1 && !no_res1) {  // the whole graph in ONE CTA: block barriers only
    ResPlan rp = h->res_block;
    k_pcg_res1<<<1, rp.bt, (size_t)rp.bytes, h->stream>>>(G, sc, prm, rp);
    SGB_CUDA(cudaGetLastError());
  } else if (h->res.valid) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->res.ncta);
    cfg.blockDim = dim3(h->res.bt);
    cfg.dynamicSmemBytes = (size_t)h->res.bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = h->res.ncta;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ResPlan rp = h->res;
    SGB_CUDA(cudaLaunchKernelEx(&cfg, k_pcg_res, G, sc, prm, rp));
  } else if (h->pcg_cluster > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->pcg_cluster);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = h->pcg_cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SGB_CUDA(cudaLaunchKernelEx(&cfg, k_pcg_cluster, G, sc, part, bar, prm));
  } else {
    void* args[] = {&G, &sc, &part, &bar, &prm};
    // at most one CTA per SM and a single GPU: the latency-oriented build of the same kernel
    const bool small = h->pcg_blocks <= h->sm_count && G.world == 1 && !h->no_small && h->pcg_threads == kThreads;
    void* fn = small ? (void*)k_pcg_small
                     : (h->pcg_threads == 320 ? (void*)k_pcg<320> : h->pcg_threads == 288 ? (void*)k_pcg<288> : (void*)k_pcg<256>);
    SGB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(h->pcg_blocks), dim3(small ? kThreads : h->pcg_threads), args, 0, h->stream));
  }
  h->tm.kernel_launches++;
  return SGB_OK;
}
sgb_status launch_backsub_update(sgb_handle* h, int dst, double lambda_override, int use_override) {
  DevGraph& G = h->G;
  if (G.nL > 0) {
    k_backsub<<<grid_for(32 * G.Hlp.nslices), kThreads, 0, h->stream>>>(G);  // one warp per slice of the grouped Hlp
    h->tm.kernel_launches++;
  }
  k_update<<<grid_for(G.nP + G.nL_owned), kThreads, 0, h->stream>>>(G, h->d_sc, dst, h->d_part_p, lambda_override, use_override);
  h->tm.kernel_launches++;
  if (G.world > 1) {  // every replica of the estimates has received every owner's rows
    k_xbarrier<<<1, 32, 0, h->stream>>>(G, h->d_sc);
    h->tm.kernel_launches++;
  }
  SGB_CUDA(cudaGetLastError());
  return SGB_OK;
}
sgb_status launch_chi2(sgb_handle* h, int buf) {
  DevGraph& G = h->G;
  k_chi2_edges<<<grid_for(G.n_pp_owned + G.n_pl_owned), kThreads, 0, h->stream>>>(G, buf, h->d_part_e);
  h->tm.kernel_launches++;
  SGB_CUDA(cudaGetLastError());
  return SGB_OK;
}
sgb_status read_scalars(sgb_handle* h) {
  SGB_CUDA(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->stream));
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  return SGB_OK;
}
float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// one OptimizationAlgorithm::solve(iteration)
sgb_status do_step(sgb_handle* h, int algo, int iteration, int* result, sgb_iter_stat* stat) {
  DevGraph& G = h->G;
  sgb_status st;
  cudaStream_t s = h->stream;
  SGB_CUDA(cudaEventRecord(h->ev.e[0], s));
  if ((st = launch_linearize(h)) != SGB_OK) return st;
  if ((st = launch_finalize_lin(h, (algo == SGB_ALGO_LM && (iteration == 0 || !h->lm_state_valid)) ? 1 : 0)) != SGB_OK) return st;
  SGB_CUDA(cudaEventRecord(h->ev.e[1], s));
  int pcg_total = 0, trials = 0;
  bool first = true;
  float t_lin = 0, t_setup = 0, t_pcg = 0, t_upd = 0;
  if (algo == SGB_ALGO_GN) {
    if ((st = launch_setup(h, 0.0, 1)) != SGB_OK) return st;
    SGB_CUDA(cudaEventRecord(h->ev.e[2], s));
    if ((st = launch_pcg(h, 0.0, 1)) != SGB_OK) return st;
    SGB_CUDA(cudaEventRecord(h->ev.e[3], s));
    if ((st = launch_backsub_update(h, G.cur, 0.0, 1)) != SGB_OK) return st;
    k_gn_control<<<1, 32, 0, s>>>(G, h->d_sc);
    h->tm.kernel_launches++;
    SGB_CUDA(cudaGetLastError());
    SGB_CUDA(cudaEventRecord(h->ev.e[4], s));
    if ((st = read_scalars(h)) != SGB_OK) return st;
    *result = h->h_sc->result == 1 ? SGB_RESULT_OK : SGB_RESULT_FAIL;
    pcg_total = h->h_sc->pcg_iters;
    trials = 1;
    t_lin = ev_ms(h->ev.e[0], h->ev.e[1]);
    t_setup = ev_ms(h->ev.e[1], h->ev.e[2]);
    t_pcg = ev_ms(h->ev.e[2], h->ev.e[3]);
    t_upd = ev_ms(h->ev.e[3], h->ev.e[4]);
  } else {
    h->lm_state_valid = true;
    int max_trials = h->opt.lm_max_trials > 0 ? h->opt.lm_max_trials : 10;
    while (true) {
      if (!first) SGB_CUDA(cudaEventRecord(h->ev.e[1], s));
      if ((st = launch_setup(h, 0.0, 0)) != SGB_OK) return st;
      SGB_CUDA(cudaEventRecord(h->ev.e[2], s));
      if ((st = launch_pcg(h, 0.0, 0)) != SGB_OK) return st;
      SGB_CUDA(cudaEventRecord(h->ev.e[3], s));
      if ((st = launch_backsub_update(h, G.cur ^ 1, 0.0, 0)) != SGB_OK) return st;
      if ((st = launch_chi2(h, G.cur ^ 1)) != SGB_OK) return st;
      k_lm_control<<<1, kThreads, 0, s>>>(G, h->d_sc, h->d_part_e, grid_for(G.n_pp_owned + G.n_pl_owned), h->d_part_p,
                                          grid_for(G.nP + G.nL_owned), max_trials);
      h->tm.kernel_launches++;
      SGB_CUDA(cudaGetLastError());
      SGB_CUDA(cudaEventRecord(h->ev.e[4], s));
      if ((st = read_scalars(h)) != SGB_OK) return st;
      if (first) t_lin = ev_ms(h->ev.e[0], h->ev.e[1]);
      t_setup += ev_ms(h->ev.e[1], h->ev.e[2]);
      t_pcg += ev_ms(h->ev.e[2], h->ev.e[3]);
      t_upd += ev_ms(h->ev.e[3], h->ev.e[4]);
      pcg_total += h->h_sc->pcg_iters;
      first = false;
      if (h->h_sc->accepted) G.cur ^= 1;  // discardTop: the trial estimates become current (on every rank)
      if (!h->h_sc->again) break;
    }
    trials = h->h_sc->trials;
    *result = h->h_sc->result;
  }
  h->tm.linearize_ms += t_lin;
  h->tm.setup_ms += t_setup;
  h->tm.pcg_ms += t_pcg;
  h->tm.update_ms += t_upd;
  h->tm.pcg_iters += pcg_total;
  h->tm.trials += trials;
  for (int i = 0; i < 4; ++i) h->tm.pcg_phase_ms[i] = 1e-6 * (double)h->h_sc->pcg_phase_ns[i];
  if (stat) {
    stat->iteration = iteration;
    stat->trials = trials;
    stat->result = *result;
    stat->pcg_iters = pcg_total;
    stat->chi2 = algo == SGB_ALGO_LM ? h->h_sc->current_chi : h->h_sc->chi2_robust;
    stat->lambda = algo == SGB_ALGO_LM ? h->h_sc->lambda : 0.0;
    stat->rho = algo == SGB_ALGO_LM ? h->h_sc->rho : 0.0;
    stat->chi2_before = h->h_sc->chi_lin;
    stat->pcg_residual = h->h_sc->pcg_rel;
  }
  return SGB_OK;
}

sgb_status do_optimize(sgb_handle* h, int algo, int max_iters, int* iters_done, sgb_iter_stat* stats) {
  if (algo != SGB_ALGO_LM && algo != SGB_ALGO_GN) {
    h->err = "unknown algorithm";
    return SGB_ERR_INVALID;
  }
  std::memset(&h->tm, 0, sizeof h->tm);
  SGB_CUDA(cudaMemsetAsync(&h->d_sc->pcg_phase_ns, 0, sizeof h->d_sc->pcg_phase_ns, h->stream));
  cudaEvent_t t0, t1;
  SGB_CUDA(cudaEventCreate(&t0));
  SGB_CUDA(cudaEventCreate(&t1));
  SGB_CUDA(cudaEventRecord(t0, h->stream));
  int done = 0, result = SGB_RESULT_OK;
  bool ok = true;
  sgb_status st = SGB_OK;
  for (int i = 0; i < max_iters && ok; ++i) {
    sgb_iter_stat tmp;
    st = do_step(h, algo, i, &result, &tmp);
    if (st != SGB_OK) break;
    if (stats) stats[i] = tmp;
    ok = (result == SGB_RESULT_OK);
    ++done;
  }
  cudaEventRecord(t1, h->stream);
  cudaEventSynchronize(t1);
  h->tm.total_ms = ev_ms(t0, t1);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  if (st != SGB_OK) return st;
  if (iters_done) *iters_done = (result == SGB_RESULT_FAIL) ? 0 : done;
  return SGB_OK;
}

}  // namespace

extern "C" {

static void online_free(sgb_handle* h);

const char* sgb_version(void) { return "sparse-gslam_b200 0.1 (sm_100a)"; }

int32_t sgb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void sgb_default_options(sgb_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);
  o->device = -1;
  o->jacobian_mode = SGB_JAC_G2O_NUMERIC;
  o->pcg_tolerance = 1e-10;
  o->pcg_max_iters = 0;
  o->verbose = 0;
  o->lm_tau = 1e-5;
  o->lm_user_lambda = 0.0;
  o->lm_max_trials = 10;
}

sgb_status sgb_create(const sgb_options* opt, sgb_handle** out) {
  if (!out) return SGB_ERR_INVALID;
  *out = nullptr;
  {
    // sgb_set_graph builds ~1 GB of host-side symbolic arrays on a 1M-pose graph and frees them at the next call
    // (the reference re-initialises once per key-frame). With glibc's defaults every array above the mmap threshold
    // is unmapped and page-faulted in again each time (~20 % of the host symbolic phase): keep freed memory in the
    // heap instead. Process-wide and performance-only; SGB_KEEP_HOST_MEMORY=0 leaves malloc alone.
    static const bool once = [] {
      const char* e = std::getenv("SGB_KEEP_HOST_MEMORY");
      if (!(e && e[0] == '0')) {
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
        mallopt(M_TRIM_THRESHOLD, 1 << 30);  // int arguments: 1 GiB is the largest power of two that fits
      }
      return true;
    }();
    (void)once;
  }
  if (sgb_device_count() <= 0) return SGB_ERR_NO_DEVICE;  // no CPU fallback by design
  sgb_handle* h = new sgb_handle();
  if (opt) h->opt = *opt; else sgb_default_options(&h->opt);
  std::memset(&h->tm, 0, sizeof h->tm);
  std::memset(&h->G, 0, sizeof h->G);
  auto fail = [&](const char* what, cudaError_t e) {
    std::fprintf(stderr, "sgb_create: %s: %s\n", what, cudaGetErrorString(e));
    delete h;
    return SGB_ERR_CUDA;
  };
  cudaError_t e;
  if (h->opt.device >= 0) {
    if ((e = cudaSetDevice(h->opt.device)) != cudaSuccess) return fail("cudaSetDevice", e);
  }
  if ((e = cudaGetDevice(&h->device)) != cudaSuccess) return fail("cudaGetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  h->sm_count = prop.multiProcessorCount;
  if (!prop.cooperativeLaunch) {
    delete h;
    return SGB_ERR_UNSUPPORTED;
  }
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
  for (auto& ev : h->ev.e)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail("cudaEventCreate", e);
  if ((e = cudaMalloc((void**)&h->d_sc, sizeof(DevScalars))) != cudaSuccess) return fail("cudaMalloc", e);
  cudaMemset(h->d_sc, 0, sizeof(DevScalars));
  if ((e = cudaMallocHost((void**)&h->h_sc, sizeof(DevScalars))) != cudaSuccess) return fail("cudaMallocHost", e);
  std::memset(h->h_sc, 0, sizeof(DevScalars));
  size_t pb = 4 * (size_t)kMaxBlocks * sizeof(double);  // k_pcg double-buffers two partial sums
  if ((e = cudaMalloc((void**)&h->d_part_p, pb)) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMalloc((void**)&h->d_part_l, pb)) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMalloc((void**)&h->d_part_e, pb)) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMalloc((void**)&h->d_bar, sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc", e);

  *out = h;
  return SGB_OK;
}

void sgb_destroy(sgb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_graph(h);
  online_free(h);
  close_peers(h);
  if (h->arena_raw) cudaFree(h->arena_raw);
  for (auto& sl : h->slabs) cudaFree(sl.base);
  h->slabs.clear();
  cudaFree(h->d_sc);
  cudaFreeHost(h->h_sc);
  cudaFree(h->d_part_p);
  cudaFree(h->d_part_l);
  cudaFree(h->d_part_e);
  cudaFree(h->d_bar);
  for (int b = 0; b < sgb_handle::kStages; ++b) {
    if (h->stage[b]) cudaFreeHost(h->stage[b]);
    if (h->stage_ev[b]) cudaEventDestroy(h->stage_ev[b]);
  }
  for (auto& ev : h->ev.e) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* sgb_last_error(const sgb_handle* h) { return h ? h->err.c_str() : "null handle"; }

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void fill_peer_tables(sgb_handle* h) {
  DevGraph& G = h->G;
  const int me = h->LP.rank;
  for (int r = 0; r < kMaxRanks; ++r) {
    char* base = (char*)h->peer_base[r];
    // unconnected slots alias this rank (never dereferenced: owners < world)
    const sgb_handle::ArenaHeader& L = base ? h->layout[r] : h->layout[me];
    if (!base) base = h->arena;
    for (int bsel = 0; bsel < 2; ++bsel) {
      G.pose_buf[bsel][r] = (double*)(base + L.off_pose[bsel]);
      G.lm_buf[bsel][r] = (double*)(base + L.off_lm[bsel]);
    }
    G.p[r] = (double*)(base + L.off_p);
    G.x_p[r] = (double*)(base + L.off_xp);
    G.t[r] = (double*)(base + L.off_t);
    G.b_l[r] = (double*)(base + L.off_bl);
    G.Hll_inv[r] = (double*)(base + L.off_hllinv);
    G.mbox[r] = (Mailbox*)(base + L.off_mbox);
    G.halo_base_at[r] = base ? L.halo_base[me] : 0;
  }
}

// Edge values that already live on the device (sgb_set_graph_device): gather them into the component-major local-edge
// arrays, one thread per local edge; EdgeSE2::setMeasurement's cached inverse is formed here as on the host path.
__global__ void __launch_bounds__(kThreads) k_gather_pp(const int32_t* __restrict__ slot, int n, const double* __restrict__ z,
                                                       const double* __restrict__ info, const double* __restrict__ phi,
                                                       double* __restrict__ zinv_o, double* __restrict__ info_o,
                                                       double* __restrict__ phi_o) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    size_t s = (size_t)slot[k];
    double x = z[3 * s], y = z[3 * s + 1], th = z[3 * s + 2];
    double thi = normalize_theta(-th);
    double c = cos(thi), sn = sin(thi);
    zinv_o[k] = c * (-x) - sn * (-y);
    zinv_o[(size_t)n + k] = sn * (-x) + c * (-y);
    zinv_o[2 * (size_t)n + k] = thi;
    for (int c6 = 0; c6 < 6; ++c6) info_o[(size_t)c6 * n + k] = info[6 * s + c6];
    phi_o[k] = phi ? phi[s] : 0.0;
  }
}
__global__ void __launch_bounds__(kThreads) k_gather_pl(const int32_t* __restrict__ slot, int n, const double* __restrict__ z,
                                                       const double* __restrict__ info, double* __restrict__ z_o,
                                                       double* __restrict__ info_o) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    size_t s = (size_t)slot[k];
    z_o[k] = z[2 * s];
    z_o[(size_t)n + k] = z[2 * s + 1];
    for (int c3 = 0; c3 < 3; ++c3) info_o[(size_t)c3 * n + k] = info[3 * s + c3];
  }
}

// out[map[k]] = k for every k with map[k] >= 0 (out pre-filled with -1): the inverse of an injective scatter map
__global__ void __launch_bounds__(kThreads) k_invert_map(const int32_t* __restrict__ map, int n, int32_t* __restrict__ out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int e = map[k];
    if (e >= 0) out[e] = k;
  }
}

// ---- LinearSolver-level entry: block values handed over in g2o's SparseBlockMatrix order ----------------------------
// one thread per input block: copies its values (column-major, like Eigen) into the SELL / Hll slots of the device
// layout. kind 0 = pose diagonal (Hpp entry e1), 1 = landmark diagonal (Hll index e1), 2 = pose-pose off-diagonal
// (Hpp entry e1 holds the block, e2 its transpose), 3 = pose-landmark (Hpl entry e1 and Hlp entry e2, same 3x2 block)
__global__ void __launch_bounds__(kThreads) k_scatter_blocks(DevGraph g, const double* __restrict__ vals, int n,
                                                            const int32_t* __restrict__ off, const int32_t* __restrict__ e1,
                                                            const int32_t* __restrict__ e2, const int32_t* __restrict__ kind) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double* in = vals + off[k];
    const int a1 = e1[k], a2 = e2[k], kd = kind[k];
    if (kd == 0) {
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) g.Hpp.vals[sell_vaddr(a1, 9, 3 * a + b)] = in[b * 3 + a];
    } else if (kd == 1) {
      g.Hll[a1] = in[0];
      g.Hll[(size_t)g.nL + a1] = in[2];
      g.Hll[2 * (size_t)g.nL + a1] = in[3];
    } else if (kd == 2) {
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          g.Hpp.vals[sell_vaddr(a1, 9, 3 * a + b)] = in[b * 3 + a];
          g.Hpp.vals[sell_vaddr(a2, 9, 3 * a + b)] = in[a * 3 + b];
        }
    } else {
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 2; ++b) {
          double v = in[b * 3 + a];
          g.Hpl.vals[sell_vaddr(a1, 6, 2 * a + b)] = v;
          g.Hlp.vals[sell_vaddr(a2, 6, 2 * a + b)] = v;
        }
    }
  }
}
// right-hand side in Hessian order -> b_p (poses, same order) and b_l (landmarks in the device's row order)
__global__ void __launch_bounds__(kThreads) k_scatter_rhs(DevGraph g, const double* __restrict__ b, const int32_t* __restrict__ lmg) {
  const int n3 = 3 * g.nP, n = n3 + 2 * g.nL;
  double* bl = g.b_l[g.rank];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < n3) g.b_p[i] = b[i];
    else {
      int q = i - n3, l = q >> 1, c = q & 1;
      bl[q] = b[n3 + 2 * lmg[l] + c];
    }
  }
}

static sgb_status set_graph_impl(sgb_handle* h, const sgb_graph_soa* g_in, int world, int rank,
                                 const sgb_device_values* dv = nullptr) {
  if (!h || !g_in) return SGB_ERR_INVALID;
  // device-resident values: the host symbolic phase only reads the index arrays; give it non-NULL value pointers
  // (never dereferenced on the host) so that its argument checks are the same on both paths
  sgb_graph_soa g_local = *g_in;
  if (dv) {
    g_local.pose_est = dv->pose_est; g_local.lm_est = dv->lm_est;
    g_local.pp_z = dv->pp_z; g_local.pp_info = dv->pp_info; g_local.pp_phi = nullptr;
    g_local.pl_z = dv->pl_z; g_local.pl_info = dv->pl_info;
  }
  const sgb_graph_soa* g = &g_local;
  static const bool prof = std::getenv("SGB_PROFILE") != nullptr;  // per-phase host timing of this call on stderr
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[sgb_set_graph] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
    tp0 = t;
  };
  SGB_CUDA(cudaSetDevice(h->device));
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  free_graph(h);
  h->lin = sgb_handle::LinearMap();
  lap("free");
  // Multi-GPU with ghost landmark rows: every rank runs the per-edge part of the symbolic phase on the edges IT needs only
  // (~1/world of the graph) -- the scan of the active vertices and the index mapping stay global, so Hessian indices are
  // the reference's. SGB_PARTITION_FULL=1 keeps the whole structure on every rank (the structure / Hessian export hooks
  // of the parity tests need it).
  static const bool full_env = [] { const char* e = std::getenv("SGB_PARTITION_FULL"); return e && e[0] == '1'; }();
  const bool filtered = world > 1 && partition_ghost_landmarks() && !full_env;
  sgb_status st = build_structure(*g, h->S, h->err, filtered ? world : 1, filtered ? rank : 0);
  if (st != SGB_OK) return st;
  lap("build_structure");
  if (dv && dv->has_robust) h->S.has_robust = true;
  st = partition_consume(h->S, world, rank, h->LP, h->err);
  if (st != SGB_OK) return st;
  lap("partition");
  const Structure& S = h->S;
  const LocalPlan& P = h->LP;
  DevGraph& G = h->G;
  std::memset(&G, 0, sizeof G);
  G.world = world; G.rank = rank; G.nP = P.nP; G.nL = P.nL; G.capP = P.capP; G.capL = P.capL;
  G.nL_owned = P.nL_owned;
  G.ghosts = (world > 1 && partition_ghost_landmarks()) ? 1 : 0;
  G.P_all = S.P_all; G.L_all = S.L_all; G.n_pp = P.n_pp; G.n_pl = P.n_pl;
  G.n_pp_owned = P.n_pp_owned; G.n_pl_owned = P.n_pl_owned;
  G.has_robust = S.has_robust ? 1 : 0;
  G.jac_numeric = h->opt.jacobian_mode == SGB_JAC_G2O_NUMERIC ? 1 : 0;
  G.cur = 0;
  // ---- arena (identical layout on every rank: sizes depend on global quantities only)
  G.pushed = P.pushed ? 1 : 0;
  size_t np = 3 * (size_t)S.P_all, nl = 2 * (size_t)S.L_all;
  size_t off = align256(sizeof(sgb_handle::ArenaHeader));
  auto take = [&](size_t doubles) { size_t o = off; off = align256(off + std::max<size_t>(doubles, 1) * sizeof(double)); return o; };
  sgb_handle::ArenaHeader& AL = h->layout[rank];
  std::memset(&AL, 0, sizeof AL);
  AL.magic = sgb_handle::kArenaMagic;
  AL.off_pose[0] = take(np); AL.off_pose[1] = take(np);
  AL.off_lm[0] = take(nl); AL.off_lm[1] = take(nl);
  // p and x_p: the rank's own rows, then the halo copies of the remote rows its matrices reference (pushed halos)
  AL.off_p = take(3 * ((size_t)P.capP + P.nH)); AL.off_xp = take(3 * ((size_t)P.capP + P.nH));
  AL.off_mbox = off; off = align256(off + sizeof(Mailbox));
  AL.off_t = take(2 * (size_t)P.capL); AL.off_bl = take(2 * (size_t)P.capL); AL.off_hllinv = take(3 * (size_t)P.capL);
  for (int r = 0; r < kMaxRanks; ++r) { AL.halo_base[r] = P.halo_base[r]; AL.halo_cnt[r] = P.halo_cnt[r]; }
  h->arena_bytes = off;
  if (world > 1) {
    if (h->arena_raw_cap < h->arena_bytes) {  // grow-only, with head-room for the next (slightly larger) graph
      if (h->arena_raw) SGB_CUDA(cudaFree(h->arena_raw));
      h->arena_raw = nullptr;
      h->arena_raw_cap = 0;
      h->arena_exported = false;
      size_t cap = h->arena_bytes + h->arena_bytes / 4;
      SGB_CUDA(cudaMalloc((void**)&h->arena_raw, cap));
      h->arena_raw_cap = cap;
    }
    h->arena = h->arena_raw;
  } else if ((st = dalloc(h, &h->arena, h->arena_bytes)) != SGB_OK) {
    return st;
  }
  SGB_CUDA(cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
  SGB_CUDA(cudaMemcpyAsync(h->arena, &AL, sizeof AL, cudaMemcpyHostToDevice, h->stream));  // h->layout outlives the copy
  for (int r = 0; r < kMaxRanks; ++r) h->peer_base[r] = nullptr;
  h->peer_base[rank] = h->arena;
  fill_peer_tables(h);
  h->connected = (world == 1);
  if ((st = dalloc(h, &h->d_pose0, np)) != SGB_OK) return st;
  if ((st = dalloc(h, &h->d_lm0, nl)) != SGB_OK) return st;
  if (np) {
    if (dv) SGB_CUDA(cudaMemcpyAsync(h->d_pose0, dv->pose_est, np * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    else if ((st = h2d(h, h->d_pose0, g->pose_est, np * sizeof(double))) != SGB_OK) return st;
    for (int bsel = 0; bsel < 2; ++bsel)
      SGB_CUDA(cudaMemcpyAsync(G.pose_buf[bsel][rank], h->d_pose0, np * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (nl) {
    if (dv) SGB_CUDA(cudaMemcpyAsync(h->d_lm0, dv->lm_est, nl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    else if ((st = h2d(h, h->d_lm0, g->lm_est, nl * sizeof(double))) != SGB_OK) return st;
    for (int bsel = 0; bsel < 2; ++bsel)
      SGB_CUDA(cudaMemcpyAsync(G.lm_buf[bsel][rank], h->d_lm0, nl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  lap("arena+estimates");
  // edge data, component-major, local edges; EdgeSE2::setMeasurement caches the inverse. The SoA transposes run on
  // host threads while this thread uploads the symbolic maps.
  std::vector<double>&zinv = h->e_zinv, &info = h->e_info, &phi = h->e_phi, &z = h->e_z, &linfo = h->e_linfo;
  if (dv) {  // values stay on the device: upload only the slot of every local edge and gather there
    std::vector<int32_t> slot_pp(P.n_pp), slot_pl(P.n_pl);
    for (int k = 0; k < P.n_pp; ++k) {
      int s = S.pp_src[P.pp_g[k]];
      slot_pp[k] = dv->pp_slot ? dv->pp_slot[s] : s;
    }
    for (int k = 0; k < P.n_pl; ++k) {
      int s = S.pl_src[P.pl_g[k]];
      slot_pl[k] = dv->pl_slot ? dv->pl_slot[s] : s;
    }
    const int32_t *d_spp = nullptr, *d_spl = nullptr;
    double *d_zinv = nullptr, *d_info = nullptr, *d_phi = nullptr, *d_z = nullptr, *d_linfo = nullptr;
    if ((st = upload(h, &d_spp, slot_pp)) != SGB_OK) return st;
    if ((st = upload(h, &d_spl, slot_pl)) != SGB_OK) return st;
    if ((st = dalloc(h, &d_zinv, 3 * (size_t)P.n_pp)) != SGB_OK) return st;
    if ((st = dalloc(h, &d_info, 6 * (size_t)P.n_pp)) != SGB_OK) return st;
    if ((st = dalloc(h, &d_phi, (size_t)P.n_pp)) != SGB_OK) return st;
    if ((st = dalloc(h, &d_z, 2 * (size_t)P.n_pl)) != SGB_OK) return st;
    if ((st = dalloc(h, &d_linfo, 3 * (size_t)P.n_pl)) != SGB_OK) return st;
    if (P.n_pp > 0) k_gather_pp<<<grid_for(P.n_pp), kThreads, 0, h->stream>>>(d_spp, P.n_pp, dv->pp_z, dv->pp_info, dv->pp_phi, d_zinv, d_info, d_phi);
    if (P.n_pl > 0) k_gather_pl<<<grid_for(P.n_pl), kThreads, 0, h->stream>>>(d_spl, P.n_pl, dv->pl_z, dv->pl_info, d_z, d_linfo);
    SGB_CUDA(cudaGetLastError());
    G.pp_zinv = d_zinv; G.pp_info = d_info; G.pp_phi = d_phi; G.pl_z = d_z; G.pl_info = d_linfo;
    SGB_CUDA(cudaStreamSynchronize(h->stream));  // slot_pp / slot_pl are locals
  }
  const size_t h_pp = dv ? 0 : (size_t)P.n_pp, h_pl = dv ? 0 : (size_t)P.n_pl;  // edges whose values come from the host
  zinv.resize(3 * h_pp);
  info.resize(6 * h_pp);
  phi.resize(h_pp);
  z.resize(2 * h_pl);
  linfo.resize(3 * h_pl);
  auto fill_pp = [&](int k0, int k1) {
    for (int k = k0; k < k1; ++k) {
      int s = S.pp_src[P.pp_g[k]];
      double x = g->pp_z[3 * (size_t)s], y = g->pp_z[3 * (size_t)s + 1], th = g->pp_z[3 * (size_t)s + 2];
      double thi = normalize_theta(-th);
      double c = std::cos(thi), sn = std::sin(thi);
      zinv[k] = c * (-x) - sn * (-y);
      zinv[(size_t)P.n_pp + k] = sn * (-x) + c * (-y);
      zinv[2 * (size_t)P.n_pp + k] = thi;
      for (int c6 = 0; c6 < 6; ++c6) info[(size_t)c6 * P.n_pp + k] = g->pp_info[6 * (size_t)s + c6];
      phi[k] = g->pp_phi ? g->pp_phi[s] : 0.0;
    }
  };
  auto fill_pl = [&](int k0, int k1) {
    for (int k = k0; k < k1; ++k) {
      int s = S.pl_src[P.pl_g[k]];
      z[k] = g->pl_z[2 * (size_t)s];
      z[(size_t)P.n_pl + k] = g->pl_z[2 * (size_t)s + 1];
      for (int c3 = 0; c3 < 3; ++c3) linfo[(size_t)c3 * P.n_pl + k] = g->pl_info[3 * (size_t)s + c3];
    }
  };
  std::vector<std::thread> workers;
  {
    const int nth = (h_pp + h_pl > 200000) ? (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 0;
    if (dv) {
      // nothing to transpose on the host
    } else if (nth == 0) {
      fill_pp(0, P.n_pp);
      fill_pl(0, P.n_pl);
    } else {
      for (int t = 0; t < nth; ++t) {
        int a0 = (int)((int64_t)P.n_pp * t / nth), a1 = (int)((int64_t)P.n_pp * (t + 1) / nth);
        int b0 = (int)((int64_t)P.n_pl * t / nth), b1 = (int)((int64_t)P.n_pl * (t + 1) / nth);
        workers.emplace_back([=, &fill_pp, &fill_pl]() { fill_pp(a0, a1); fill_pl(b0, b1); });
      }
    }
  }
  struct Joiner {  // the workers reference locals of this frame: never leave it (early error returns) before they end
    std::vector<std::thread>& w;
    ~Joiner() { for (auto& t : w) if (t.joinable()) t.join(); }
  } joiner{workers};
#define UP(field, vec) if ((st = upload(h, &G.field, vec)) != SGB_OK) return st
  UP(pose_of_l, P.pose_of_l); UP(lm_of_l, P.lm_of_l);
  UP(pp_i, P.pp_i); UP(pp_j, P.pp_j); UP(pp_hi, P.pp_hi); UP(pp_hj, P.pp_hj);
  UP(pp_e_ij, P.pp_e_ij); UP(pp_e_ji, P.pp_e_ji); UP(pp_dup, P.pp_dup);
  UP(pl_p, P.pl_p); UP(pl_l, P.pl_l); UP(pl_hp, P.pl_hp); UP(pl_hl, P.pl_hl);
  UP(pl_e_pl, P.pl_e_pl); UP(pl_e_lp, P.pl_e_lp); UP(pl_dup, P.pl_dup);
  UP(pinc_ptr, P.pinc_ptr); UP(pinc, P.pinc); UP(linc_ptr, P.linc_ptr); UP(linc, P.linc);
  UP(hpp_diag, P.hpp_diag);
  if (P.pushed) { UP(send_ptr, P.send_ptr); UP(send_dst, P.send_dst); }
  // maps of the lane-per-block landmark linearisation: entry -> leading edge, and the observations from fixed poses
  // (function scope: alive until the stream has been synchronised at the end of this call)
  // The entry -> edge map is the inverse of pl_e_lp, which is on the device already: scattered there (k_invert_map). The
  // observations from fixed poses are few (the reference fixes the first pose only): one scan of the edges' pose indices
  // finds them, then they are bucketed by landmark row, in insertion order inside a row.
  std::vector<int32_t> lfix_ptr((size_t)P.nL + 1, 0), lfix;
  {
    struct Fx { int32_t row, gedge, k; };  // landmark row on this rank, global (insertion-ordered) edge index, local edge
    std::vector<Fx> fx;
    for (int k = 0; k < P.n_pl; ++k) {
      if (!(P.pl_hp[k] < 0 && P.pl_hl[k] >= 0)) continue;
      const int32_t enc = P.enc_lm_here.empty() ? P.enc_lm[P.pl_hl[k]] : P.enc_lm_here[P.pl_hl[k]];
      if (enc < 0 || (enc >> kOwnerShift) != rank) continue;  // a landmark row kept elsewhere
      fx.push_back({enc & kLocalMask, P.pl_g[k], k});
    }
    std::sort(fx.begin(), fx.end(), [](const Fx& a, const Fx& b) { return a.row != b.row ? a.row < b.row : a.gedge < b.gedge; });
    for (auto& f : fx) lfix_ptr[(size_t)f.row + 1]++;
    for (int l = 0; l < P.nL; ++l) lfix_ptr[(size_t)l + 1] += lfix_ptr[(size_t)l];
    for (auto& f : fx) lfix.push_back(f.k);
  }
  UP(lfix_ptr, lfix_ptr); UP(lfix, lfix);
  {
    int32_t* d_map = nullptr;
    if ((st = dalloc(h, &d_map, (size_t)P.Hlp.entries())) != SGB_OK) return st;
    SGB_CUDA(cudaMemsetAsync(d_map, 0xff, std::max<size_t>((size_t)P.Hlp.entries(), 1) * sizeof(int32_t), h->stream));
    if (P.n_pl > 0) k_invert_map<<<grid_for(P.n_pl), kThreads, 0, h->stream>>>(G.pl_e_lp, P.n_pl, d_map);
    SGB_CUDA(cudaGetLastError());
    G.hlp_edge = d_map;
  }
  lap("upload maps");
  for (auto& t : workers) t.join();
  workers.clear();
  if (!dv) {
    UP(pp_zinv, zinv); UP(pp_info, info); UP(pp_phi, phi);
    UP(pl_z, z); UP(pl_info, linfo);
  }
#undef UP
  lap("edge data");
  if ((st = upload_sell(h, &G.Hpp, P.Hpp, 9)) != SGB_OK) return st;
  if ((st = upload_sell(h, &G.Hpl, P.Hpl, 6)) != SGB_OK) return st;
  if ((st = upload_sell(h, &G.Hlp, P.Hlp, 6)) != SGB_OK) return st;
  size_t n3 = 3 * (size_t)P.nP, n2 = 2 * (size_t)P.nL;
  if ((st = dalloc(h, &G.Hll, 3 * (size_t)P.nL)) != SGB_OK) return st;
  if ((st = dalloc(h, &G.b_p, n3)) != SGB_OK) return st;
  if ((st = dalloc(h, &G.x_l, n2)) != SGB_OK) return st;
  if ((st = dalloc(h, &G.Cinv, 3 * 3 * kChunk * (size_t)P.nP)) != SGB_OK) return st;
  if ((st = dalloc(h, &G.bt, n3)) != SGB_OK) return st;
  if ((st = dalloc(h, &G.r, n3)) != SGB_OK) return st;
  if ((st = dalloc(h, &G.d, n3)) != SGB_OK) return st;
  if ((st = dalloc(h, &G.s, n3)) != SGB_OK) return st;
  SGB_CUDA(cudaMemsetAsync(h->d_sc, 0, sizeof(DevScalars), h->stream));
  // persistent PCG grid: every CTA must be co-resident
  // Threads per CTA: the block size whose grid needs the fewest rows per thread (ceil), larger blocks paying a small
  // penalty for their smaller register budget; ties go to the smaller block. SGB_PCG_THREADS forces one (tuning runs).
  {
    static const int forced = [] { const char* e = std::getenv("SGB_PCG_THREADS"); return e ? std::atoi(e) : 0; }();
    const int items = std::max(P.nP, 32 * P.Hlp.nslices);
    const int cand[3] = {256, 288, 320};
    const double pen[3] = {1.0, 1.04, 1.15};
    double best = 0.0;
    int best_bt = kThreads, best_blocks = 1;
    // Measured (C5, LM it/s with 256 / 288 / 320 threads; rows per thread at 256 threads in brackets): one GPU (8.8)
    // 41.9 / 40.6 / 39.6; two GPUs (4.4) 77.2 / 76.0 / 72.3; four GPUs (2.2) 122.7 / 129.6 / 124.8; eight GPUs (1.1)
    // 175.0 / 197.4 / -- . A phase lasts as many whole rows as its busiest thread owns only when that number is small:
    // three -> two rows (4 GPUs) and two -> one (8 GPUs) pay, five -> four and nine -> eight do not, and the smaller
    // register budget of the larger blocks always costs a little. The larger blocks are considered up to three rows per thread.
    int per_sm256 = 0;
    SGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm256, (const void*)k_pcg<256>, 256, 0));
    const bool quantised = (double)P.nP / ((double)std::max(1, per_sm256) * h->sm_count * 256.0) <= 3.0;
    for (int c = 0; c < 3; ++c) {
      if (forced > 0 && cand[c] != forced) continue;
      if (forced <= 0 && c > 0 && !quantised) continue;
      int per_sm = 0;
      const void* fn = cand[c] == 320 ? (const void*)k_pcg<320> : cand[c] == 288 ? (const void*)k_pcg<288> : (const void*)k_pcg<256>;
      SGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, cand[c], 0));
      const int limit = std::max(1, per_sm * h->sm_count);
      const int want = std::max(1, (items + cand[c] - 1) / cand[c]);
      const int blocks = std::min(std::min(limit, want), kMaxBlocks);
      const double threads = (double)blocks * cand[c];
      const double cost = std::ceil(std::max(1, P.nP) / threads) * pen[c];
      if (best == 0.0 || cost < best - 1e-9) { best = cost; best_bt = cand[c]; best_blocks = blocks; }
    }
    h->pcg_threads = best_bt;
    h->pcg_blocks = best_blocks;
    if (prof) std::fprintf(stderr, "[sgb_set_graph] pcg grid %d x %d threads (%d rows, %.2f rows per thread)\n", h->pcg_blocks,
                           h->pcg_threads, P.nP, P.nP / ((double)h->pcg_blocks * h->pcg_threads));
  }
  const int want = std::max(1, (std::max(P.nP, 32 * P.Hlp.nslices) + kThreads - 1) / kThreads);
  // a graph that fits <= 16 CTAs runs its PCG as one thread-block cluster (hardware barrier, partials through DSMEM)
  h->pcg_cluster = 0;
  static const bool no_cluster = std::getenv("SGB_NO_CLUSTER") != nullptr;
  if (world == 1 && want <= 16 && !no_cluster) {
    int cb = 1;
    while (cb < want) cb <<= 1;
    static bool attr_ok = cudaFuncSetAttribute(k_pcg_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (cb <= 8 || attr_ok) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cb);
      cfg.blockDim = dim3(kThreads);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cb;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, k_pcg_cluster, &cfg) == cudaSuccess && nclusters >= 1) h->pcg_cluster = cb;
    }
    cudaGetLastError();  // a refused cluster shape is not an error: the cooperative grid is used instead
    if (prof) std::fprintf(stderr, "[sgb_set_graph] pcg cluster %d (wanted %d CTAs)\n", h->pcg_cluster, want);
  }
  // a graph whose per-CTA share of the matrices fits the shared memory of a cluster: the cluster-resident solve
  // (all three plans are reset HERE: a graph without free poses skips the planning below, and a plan left over from the
  // previous graph of a recycled handle would launch the resident kernel with that graph's geometry)
  h->res = ResPlan();
  h->res4 = ResPlan();
  h->res_block = ResPlan();
  h->cz = CoarsePlan();
  static const bool no_res = std::getenv("SGB_NO_RESIDENT") != nullptr;
  if (world == 1 && !no_res && !no_cluster && P.nP > 0) {
    static const bool np_ok = cudaFuncSetAttribute(k_pcg_res, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    auto cap_of = [](const HostSell& M, int spc, int ncta) {
      int cap = 0;
      for (int c = 0; c < ncta; ++c) {
        int s0 = std::min(M.nslices, c * spc), s1 = std::min(M.nslices, (c + 1) * spc);
        cap = std::max(cap, M.sbase[s1] - M.sbase[s0]);
      }
      return cap;
    };
    for (int bt : {256, 128, 64}) {
      const int spc = bt / 32;
      int need = std::max((P.nP + bt - 1) / bt, (P.Hlp.nslices + spc - 1) / spc);
      int cb = 1;
      while (cb < need) cb <<= 1;
      if (cb > 16 || (cb > 8 && !np_ok)) continue;
      ResPlan rp = ResPlan();
      rp.valid = 0; rp.bt = bt; rp.ncta = cb;
      rp.cap_pp = cap_of(P.Hpp, spc, cb); rp.cap_pl = cap_of(P.Hpl, spc, cb); rp.cap_lp = cap_of(P.Hlp, spc, cb);
      rp.nz = (3 * P.nP + 1) & ~1; rp.nt = std::max(2, 2 * P.nL);
      rp.cap_sl = spc; rp.cap_lr = 32 * spc;  // a slice holds at most 32 landmark rows
      rp.rows_cta = bt;
      rp.bytes = (int)res_offsets(rp).total;
      if (rp.bytes > 224 * 1024) continue;
      if (cudaFuncSetAttribute(k_pcg_res, cudaFuncAttributeMaxDynamicSharedMemorySize, rp.bytes) != cudaSuccess) { cudaGetLastError(); continue; }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cb);
      cfg.blockDim = dim3(bt);
      cfg.dynamicSmemBytes = (size_t)rp.bytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cb;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, k_pcg_res, &cfg) == cudaSuccess && nclusters >= 1) {
        rp.valid = 1;
        h->res = rp;
        break;
      }
      cudaGetLastError();
    }
    {  // four lanes per pose row: one CTA if the graph fits it, else the smallest cluster
      h->res4 = ResPlan();
      h->cz = CoarsePlan();
      static const bool no_res4 = std::getenv("SGB_NO_RES4") != nullptr;
      static const bool np4_ok = cudaFuncSetAttribute(k_pcg_res4<512, 4, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                                 cudaFuncSetAttribute(k_pcg_res4<1024, 2, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                                 cudaFuncSetAttribute(k_pcg_res4<512, 4, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                                 cudaFuncSetAttribute(k_pcg_res4<1024, 2, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
      // Two-level preconditioner (sgb_coarse.h): planned first WITH the coarse inverse in every CTA's shared memory, then
      // without it. sgb_options.coarse_nodes, else the environment (SGB_COARSE=0: off, SGB_COARSE_NODES: cap), decide.
      static const int cz_default = [] {  // on by default (validated on hardware in round 2); SGB_COARSE=0 switches it off
        const char* on = std::getenv("SGB_COARSE");
        if (on && std::atoi(on) == 0) return 0;
        const char* e = std::getenv("SGB_COARSE_NODES");
        return e ? std::atoi(e) : kCzMaxNodes;
      }();
      const int cz_nodes = std::max(0, std::min(h->opt.coarse_nodes != 0 ? h->opt.coarse_nodes : cz_default, kCzMaxNodes));
      const int cz_h = cz_nodes > 0 ? coarse_spacing(P.nP, cz_nodes) : 0;
      auto plan4 = [&](int h_cz) {
        for (int pass = 0; pass < 2; ++pass)        // pass 0: single CTA, pass 1: clusters
          for (int ib = 0; ib < 3; ++ib) {
            static const int kBt[2][3] = {{256, 512, 1024}, {1024, 512, 256}};  // smallest CTA that holds the rows / fewest CTAs
            const int bt = kBt[pass][ib];
            const int rows_cta = bt / 4, spc_p = rows_cta / 32, spc_l = bt / 32;
            if (spc_p < 1) continue;
            if (h_cz > 0 && rows_cta % h_cz != 0) continue;  // a segment of the coarse space never spans two CTAs
            int need = std::max((P.nP + rows_cta - 1) / rows_cta, pass == 0 ? 1 : (P.Hlp.nslices + spc_l - 1) / spc_l);
            int cb = 1;
            while (cb < need) cb <<= 1;
            if ((pass == 0) != (cb == 1)) continue;
            if (cb > 16 || (cb > 8 && !np4_ok)) continue;
            ResPlan rp = ResPlan();
            rp.valid = 0; rp.bt = bt; rp.ncta = cb; rp.rows_cta = rows_cta;
            rp.cap_pp = cap_of(P.Hpp, spc_p, cb); rp.cap_pl = cap_of(P.Hpl, spc_p, cb);
            rp.cap_lp = cb == 1 ? (int)P.Hlp.entries() : cap_of(P.Hlp, spc_l, cb);
            rp.nz = (3 * P.nP + 1) & ~1; rp.nt = std::max(2, 2 * P.nL);
            rp.cap_sl = cb == 1 ? std::max(1, P.Hlp.nslices) : spc_l;
            rp.cap_lr = cb == 1 ? std::max(1, P.nL) : 32 * spc_l;
            rp.cz_h = h_cz;
            rp.cz_nc = h_cz > 0 ? 3 * ((P.nP + h_cz - 1) / h_cz + 1) : 0;
            rp.bytes = (int)res_offsets(rp).total;
            if (rp.bytes > (h_cz > 0 ? 218 : 224) * 1024) continue;  // the coarse build keeps 5.4 KB more of static shared memory
            const void* fn4 = h_cz > 0 ? (bt <= 512 ? (const void*)k_pcg_res4<512, 4, true> : (const void*)k_pcg_res4<1024, 2, true>)
                                       : (bt <= 512 ? (const void*)k_pcg_res4<512, 4, false> : (const void*)k_pcg_res4<1024, 2, false>);
            if (cudaFuncSetAttribute(fn4, cudaFuncAttributeMaxDynamicSharedMemorySize, (h_cz > 0 ? 218 : 224) * 1024) != cudaSuccess) { cudaGetLastError(); continue; }
            if (cb > 1) {
              cudaLaunchConfig_t cfg = {};
              cfg.gridDim = dim3(cb);
              cfg.blockDim = dim3(bt);
              cfg.dynamicSmemBytes = (size_t)rp.bytes;
              cudaLaunchAttribute at[1];
              at[0].id = cudaLaunchAttributeClusterDimension;
              at[0].val.clusterDim.x = cb;
              at[0].val.clusterDim.y = 1;
              at[0].val.clusterDim.z = 1;
              cfg.attrs = at;
              cfg.numAttrs = 1;
              int nclusters = 0;
              cudaError_t eo = cudaOccupancyMaxActiveClusters(&nclusters, fn4, &cfg);
              if (!(eo == cudaSuccess && nclusters >= 1)) { cudaGetLastError(); continue; }
            }
            rp.valid = 1;
            h->res4 = rp;
            return true;
          }
        return false;
      };
      if (!no_res4) {
        bool with_cz = false;
        if (cz_h > 0) {
          const bool cz_attr = cudaFuncSetAttribute(k_setup_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           (int)(((size_t)kCzMaxDim * cz_ld(kCzMaxDim) + kCzMaxDim) * sizeof(double))) == cudaSuccess;
          // a coarser space (fewer nodes, a smaller inverse) when the finest one does not fit next to the matrices
          for (int hc = cz_h; hc <= 64 && cz_attr && !with_cz; hc *= 2) with_cz = plan4(hc);
          cudaGetLastError();
        }
        if (!with_cz) plan4(0);
        if (h->res4.valid && h->res4.cz_nc > 0) {  // gather lists of the coarse matrix, scratch and the inverse
          plan_coarse(P, h->res4.cz_h, h->cz);
          // a landmark reached from k nodes contributes k (k + 1) / 2 Schur terms: a graph whose every landmark is seen from
          // every part of the trajectory would need lists out of proportion to its size -- it keeps the plain preconditioner
          if (h->cz.t_g.size() / 2 > ((size_t)4 << 20)) {
            h->cz = CoarsePlan();
            plan4(0);
          }
        }
        if (h->res4.valid && h->res4.cz_nc > 0 && h->cz.nn > 0) {
          const CoarsePlan& C = h->cz;
          sgb_status stc;
          G.cz_h = C.h; G.cz_nn = C.nn; G.cz_ng = C.ng;
          if ((stc = upload(h, &G.cz_g_ptr, C.g_ptr)) != SGB_OK || (stc = upload(h, &G.cz_g_e, C.g_e)) != SGB_OK ||
              (stc = upload(h, &G.cz_g_w, C.g_w)) != SGB_OK || (stc = upload(h, &G.cz_g_lm, C.g_lm)) != SGB_OK ||
              (stc = upload(h, &G.cz_p_ptr, C.p_ptr)) != SGB_OK || (stc = upload(h, &G.cz_p_e, C.p_e)) != SGB_OK ||
              (stc = upload(h, &G.cz_p_w, C.p_w)) != SGB_OK || (stc = upload(h, &G.cz_rr, C.rr)) != SGB_OK ||
              (stc = upload(h, &G.cz_t_ptr, C.t_ptr)) != SGB_OK || (stc = upload(h, &G.cz_t_g, C.t_g)) != SGB_OK ||
              (stc = dalloc(h, &G.cz_G, 6 * (size_t)C.ng)) != SGB_OK ||
              (stc = dalloc(h, &G.cz_A, (size_t)9 * C.nn * C.nn)) != SGB_OK || (stc = dalloc(h, &G.cz_fail, 1)) != SGB_OK)
            return stc;
        }
      }
      if (prof) std::fprintf(stderr, "[sgb_set_graph] resident solve, four lanes per row: %s (%d CTAs x %d threads, %d bytes), coarse space: %d nodes every %d rows\n",
                             h->res4.valid ? "yes" : "no", h->res4.ncta, h->res4.bt, h->res4.bytes, h->res4.cz_nc / 3, h->res4.cz_h);
    }
    {  // the same question for one 256-thread CTA (sgb_optimize_batch: one graph per CTA)
      h->res_block = ResPlan();
      ResPlan rp = ResPlan();
      rp.valid = 0; rp.bt = kThreads; rp.ncta = 1;
      rp.cap_pp = cap_of(P.Hpp, kThreads / 32, 1); rp.cap_pl = cap_of(P.Hpl, kThreads / 32, 1);
      rp.cap_lp = (int)P.Hlp.entries();  // a single CTA keeps every landmark-major slice
      rp.nz = (3 * P.nP + 1) & ~1; rp.nt = std::max(2, 2 * P.nL);
      rp.cap_sl = std::max(1, P.Hlp.nslices); rp.cap_lr = std::max(1, P.nL);
      rp.rows_cta = kThreads;
      rp.bytes = (int)res_offsets(rp).total;
      if (P.nP <= kThreads && rp.bytes <= 200 * 1024 &&
          cudaFuncSetAttribute(k_pcg_res1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess) {
        rp.valid = 1;
        // the batched kernel (k_lm_block) applies the two-level preconditioner too when the graph's coarse lists exist (they
        // were planned with the four-lane solve above), a segment fits a warp and the inverse fits next to the matrices
        static const bool no_cz_block = std::getenv("SGB_COARSE_BATCH") != nullptr && std::atoi(std::getenv("SGB_COARSE_BATCH")) == 0;
        if (G.cz_h > 0 && G.cz_h <= 32 && !no_cz_block) {
          ResPlan rc = rp;
          rc.cz_h = G.cz_h;
          rc.cz_nc = 3 * G.cz_nn;
          rc.bytes = (int)res_offsets(rc).total;
          if (rc.bytes <= 200 * 1024) rp = rc;
        }
        h->res_block = rp;
      }
      cudaGetLastError();
    }
    if (prof) std::fprintf(stderr, "[sgb_set_graph] resident solve: %s (%d CTAs x %d threads, %d bytes of shared memory)\n",
                           h->res.valid ? "yes" : "no", h->res.ncta, h->res.bt, h->res.bytes);
  }
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  lap("matrices+sync");
  h->has_graph = true;
  h->lm_state_valid = false;
  return SGB_OK;
}

// grow-only device array of the online store: keeps the first `keep` elements when it has to move
static sgb_status online_reserve(sgb_handle* h, double** p, size_t* cap, size_t need, size_t width, size_t keep) {
  if (need <= *cap) return SGB_OK;
  size_t ncap = std::max(need + need / 2, (size_t)256);
  double* q = nullptr;
  SGB_CUDA(cudaMalloc((void**)&q, ncap * width * sizeof(double)));
  if (*p && keep) SGB_CUDA(cudaMemcpyAsync(q, *p, keep * width * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (*p) {
    SGB_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(*p);
  }
  *p = q;
  *cap = ncap;
  return SGB_OK;
}
static void online_free(sgb_handle* h) {
  auto& O = h->on;
  for (double** p : {&O.d_pose, &O.d_lm, &O.d_pp_z, &O.d_pp_info, &O.d_pp_phi, &O.d_pl_z, &O.d_pl_info}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  O = sgb_handle::Online();
}
// (re-)plans the graph held by the online store: host symbolic phase on the index mirror, values gathered on the device
static sgb_status online_replan(sgb_handle* h) {
  auto& O = h->on;
  sgb_graph_soa g;
  std::memset(&g, 0, sizeof g);
  g.n_poses = (int32_t)O.pose_id.size(); g.pose_id = O.pose_id.data(); g.pose_fixed = O.pose_fixed.data();
  g.n_landmarks = (int32_t)O.lm_id.size(); g.lm_id = O.lm_id.data(); g.lm_fixed = O.lm_fixed.data();
  g.n_pp = (int32_t)O.pp_i.size(); g.pp_i = O.pp_i.data(); g.pp_j = O.pp_j.data(); g.pp_seq = O.pp_seq.data();
  g.n_pl = (int32_t)O.pl_pose.size(); g.pl_pose = O.pl_pose.data(); g.pl_lm = O.pl_lm.data(); g.pl_seq = O.pl_seq.data();
  sgb_device_values dv;
  std::memset(&dv, 0, sizeof dv);
  dv.pose_est = O.d_pose; dv.lm_est = O.d_lm;
  dv.pp_z = O.d_pp_z; dv.pp_info = O.d_pp_info; dv.pp_phi = O.d_pp_phi;
  dv.pl_z = O.d_pl_z; dv.pl_info = O.d_pl_info;
  dv.has_robust = O.any_phi ? 1 : 0;
  return set_graph_impl(h, &g, 1, 0, &dv);
}
// appends vertices / edges to the store: index mirror on the host, values to the device arrays (H2D of the delta only)
static sgb_status online_append(sgb_handle* h, const sgb_graph_delta* d) {
  auto& O = h->on;
  const size_t P0 = O.pose_id.size(), L0 = O.lm_id.size(), E0 = O.pp_i.size(), F0 = O.pl_pose.size();
  const size_t nP = (size_t)d->n_new_poses, nL = (size_t)d->n_new_landmarks, nE = (size_t)d->n_new_pp, nF = (size_t)d->n_new_pl;
  if ((nP && !d->pose_est) || (nL && !d->lm_est) || (nE && (!d->pp_i || !d->pp_j || !d->pp_z || !d->pp_info)) ||
      (nF && (!d->pl_pose || !d->pl_lm || !d->pl_z || !d->pl_info))) {
    h->err = "graph delta with missing arrays";
    return SGB_ERR_INVALID;
  }
  for (size_t k = 0; k < nE; ++k)
    if (d->pp_i[k] < 0 || d->pp_j[k] < 0 || (size_t)d->pp_i[k] >= P0 + nP || (size_t)d->pp_j[k] >= P0 + nP) { h->err = "graph delta: pose-pose edge with a bad vertex index"; return SGB_ERR_INVALID; }
  for (size_t k = 0; k < nF; ++k)
    if (d->pl_pose[k] < 0 || d->pl_lm[k] < 0 || (size_t)d->pl_pose[k] >= P0 + nP || (size_t)d->pl_lm[k] >= L0 + nL) { h->err = "graph delta: pose-line edge with a bad vertex index"; return SGB_ERR_INVALID; }
  sgb_status st;
  if ((st = online_reserve(h, &O.d_pose, &O.cap_pose, P0 + nP, 3, P0)) != SGB_OK) return st;
  if ((st = online_reserve(h, &O.d_lm, &O.cap_lm, L0 + nL, 2, L0)) != SGB_OK) return st;
  { size_t c1 = O.cap_pp, c2 = O.cap_pp, c3 = O.cap_pp;
    if ((st = online_reserve(h, &O.d_pp_z, &c1, E0 + nE, 3, E0)) != SGB_OK) return st;
    if ((st = online_reserve(h, &O.d_pp_info, &c2, E0 + nE, 6, E0)) != SGB_OK) return st;
    if ((st = online_reserve(h, &O.d_pp_phi, &c3, E0 + nE, 1, E0)) != SGB_OK) return st;
    O.cap_pp = c1; }
  { size_t c1 = O.cap_pl, c2 = O.cap_pl;
    if ((st = online_reserve(h, &O.d_pl_z, &c1, F0 + nF, 2, F0)) != SGB_OK) return st;
    if ((st = online_reserve(h, &O.d_pl_info, &c2, F0 + nF, 3, F0)) != SGB_OK) return st;
    O.cap_pl = c1; }
  if (nP && (st = h2d(h, O.d_pose + 3 * P0, d->pose_est, 3 * nP * sizeof(double))) != SGB_OK) return st;
  if (nL && (st = h2d(h, O.d_lm + 2 * L0, d->lm_est, 2 * nL * sizeof(double))) != SGB_OK) return st;
  if (nE) {
    if ((st = h2d(h, O.d_pp_z + 3 * E0, d->pp_z, 3 * nE * sizeof(double))) != SGB_OK) return st;
    if ((st = h2d(h, O.d_pp_info + 6 * E0, d->pp_info, 6 * nE * sizeof(double))) != SGB_OK) return st;
    if (d->pp_phi) { if ((st = h2d(h, O.d_pp_phi + E0, d->pp_phi, nE * sizeof(double))) != SGB_OK) return st; }
    else SGB_CUDA(cudaMemsetAsync(O.d_pp_phi + E0, 0, nE * sizeof(double), h->stream));
  }
  if (nF) {
    if ((st = h2d(h, O.d_pl_z + 2 * F0, d->pl_z, 2 * nF * sizeof(double))) != SGB_OK) return st;
    if ((st = h2d(h, O.d_pl_info + 3 * F0, d->pl_info, 3 * nF * sizeof(double))) != SGB_OK) return st;
  }
  SGB_CUDA(cudaStreamSynchronize(h->stream));  // the caller may free its arrays when the call returns
  for (size_t i = 0; i < nP; ++i) { O.pose_id.push_back(d->pose_id ? d->pose_id[i] : (int32_t)(P0 + i)); O.pose_fixed.push_back(d->pose_fixed ? d->pose_fixed[i] : 0); }
  for (size_t i = 0; i < nL; ++i) { O.lm_id.push_back(d->lm_id ? d->lm_id[i] : (int32_t)(10000000 + L0 + i)); O.lm_fixed.push_back(d->lm_fixed ? d->lm_fixed[i] : 0); }
  for (size_t k = 0; k < nE; ++k) {
    O.pp_i.push_back(d->pp_i[k]); O.pp_j.push_back(d->pp_j[k]);
    O.pp_seq.push_back(d->pp_seq ? d->pp_seq[k] : O.next_seq + (int64_t)k);
    if (d->pp_phi && d->pp_phi[k] > 0.0) O.any_phi = true;
  }
  for (size_t k = 0; k < nF; ++k) {
    O.pl_pose.push_back(d->pl_pose[k]); O.pl_lm.push_back(d->pl_lm[k]);
    O.pl_seq.push_back(d->pl_seq ? d->pl_seq[k] : O.next_seq + (int64_t)(nE + k));
  }
  for (size_t k = 0; k < nE; ++k) O.next_seq = std::max(O.next_seq, O.pp_seq[E0 + k] + 1);
  for (size_t k = 0; k < nF; ++k) O.next_seq = std::max(O.next_seq, O.pl_seq[F0 + k] + 1);
  return SGB_OK;
}

sgb_status sgb_set_graph(sgb_handle* h, const sgb_graph_soa* g) {
  if (!h || !g) return SGB_ERR_INVALID;
  if (!h->opt.incremental) return set_graph_impl(h, g, 1, 0);
  // incremental handle: the graph goes into the online store first, then it is planned from there
  SGB_CUDA(cudaSetDevice(h->device));
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  auto& O = h->on;
  O.pose_id.clear(); O.lm_id.clear(); O.pp_i.clear(); O.pp_j.clear(); O.pl_pose.clear(); O.pl_lm.clear();
  O.pose_fixed.clear(); O.lm_fixed.clear(); O.pp_seq.clear(); O.pl_seq.clear();
  O.next_seq = 0; O.any_phi = false; O.valid = false;
  sgb_graph_delta d;
  std::memset(&d, 0, sizeof d);
  d.n_new_poses = g->n_poses; d.pose_id = g->pose_id; d.pose_est = g->pose_est; d.pose_fixed = g->pose_fixed;
  d.n_new_landmarks = g->n_landmarks; d.lm_id = g->lm_id; d.lm_est = g->lm_est; d.lm_fixed = g->lm_fixed;
  d.n_new_pp = g->n_pp; d.pp_i = g->pp_i; d.pp_j = g->pp_j; d.pp_z = g->pp_z; d.pp_info = g->pp_info; d.pp_phi = g->pp_phi; d.pp_seq = g->pp_seq;
  d.n_new_pl = g->n_pl; d.pl_pose = g->pl_pose; d.pl_lm = g->pl_lm; d.pl_z = g->pl_z; d.pl_info = g->pl_info; d.pl_seq = g->pl_seq;
  // sgb_graph_soa's default insertion rank: pose-pose edges in array order, then the pose-line edges
  std::vector<int64_t> seq_pp, seq_pl;
  if (!g->pp_seq) { seq_pp.resize(std::max(0, g->n_pp)); for (int k = 0; k < g->n_pp; ++k) seq_pp[k] = k; d.pp_seq = seq_pp.data(); }
  if (!g->pl_seq) { seq_pl.resize(std::max(0, g->n_pl)); for (int k = 0; k < g->n_pl; ++k) seq_pl[k] = (int64_t)g->n_pp + k; d.pl_seq = seq_pl.data(); }
  if (g->n_poses < 0 || g->n_landmarks < 0 || g->n_pp < 0 || g->n_pl < 0) { h->err = "negative size"; return SGB_ERR_INVALID; }
  sgb_status st = online_append(h, &d);
  if (st != SGB_OK) return st;
  if ((st = online_replan(h)) != SGB_OK) return st;
  O.valid = true;
  return SGB_OK;
}

sgb_status sgb_update_graph(sgb_handle* h, const sgb_graph_delta* d) {
  if (!h || !d) return SGB_ERR_INVALID;
  if (!h->opt.incremental) { h->err = "sgb_update_graph needs a handle created with sgb_options.incremental != 0"; return SGB_ERR_UNSUPPORTED; }
  if (!h->has_graph || !h->on.valid) { h->err = "sgb_update_graph: no graph to extend (call sgb_set_graph first)"; return SGB_ERR_NOT_INITIALIZED; }
  if (d->n_new_poses < 0 || d->n_new_landmarks < 0 || d->n_new_pp < 0 || d->n_new_pl < 0) { h->err = "negative size"; return SGB_ERR_INVALID; }
  SGB_CUDA(cudaSetDevice(h->device));
  auto& O = h->on;
  // the estimates of the existing vertices are the ones on the device: move them into the store before the pooled
  // memory that holds them is recycled by the re-planning
  const size_t P0 = O.pose_id.size(), L0 = O.lm_id.size();
  const DevGraph& G = h->G;
  if (P0) SGB_CUDA(cudaMemcpyAsync(O.d_pose, G.pose_buf[G.cur][G.rank], 3 * P0 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (L0) SGB_CUDA(cudaMemcpyAsync(O.d_lm, G.lm_buf[G.cur][G.rank], 2 * L0 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  sgb_status st = online_append(h, d);
  if (st != SGB_OK) return st;
  return online_replan(h);
}

sgb_status sgb_set_graph_device(sgb_handle* h, const sgb_graph_soa* indices, const sgb_device_values* dv) {
  if (!h || !indices || !dv) return SGB_ERR_INVALID;
  if ((indices->n_poses > 0 && !dv->pose_est) || (indices->n_landmarks > 0 && !dv->lm_est) ||
      (indices->n_pp > 0 && (!dv->pp_z || !dv->pp_info)) || (indices->n_pl > 0 && (!dv->pl_z || !dv->pl_info))) {
    h->err = "sgb_set_graph_device: missing device value array";
    return SGB_ERR_INVALID;
  }
  return set_graph_impl(h, indices, 1, 0, dv);
}

sgb_status sgb_set_graph_partitioned(sgb_handle* h, const sgb_graph_soa* g, int32_t world, int32_t rank) {
  return set_graph_impl(h, g, world, rank);
}

sgb_status sgb_comm_get_handle(sgb_handle* h, void* out64) {
  if (!h || !out64) return SGB_ERR_INVALID;
  if (!h->has_graph) return SGB_ERR_NOT_INITIALIZED;
  SGB_CUDA(cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (h->LP.world == 1 || h->arena != h->arena_raw) {  // a single-rank arena lives in the pooled memory: export it as it is
    cudaIpcMemHandle_t mh;
    SGB_CUDA(cudaIpcGetMemHandle(&mh, h->arena));
    std::memcpy(out64, &mh, 64);
    return SGB_OK;
  }
  if (!h->arena_exported) {
    SGB_CUDA(cudaIpcGetMemHandle(&h->arena_handle, h->arena_raw));
    h->arena_exported = true;
  }
  std::memcpy(out64, &h->arena_handle, 64);
  return SGB_OK;
}

sgb_status sgb_comm_connect(sgb_handle* h, const void* handles, int32_t world) {
  if (!h || !handles) return SGB_ERR_INVALID;
  if (!h->has_graph) return SGB_ERR_NOT_INITIALIZED;
  if (world != h->LP.world) { h->err = "world size mismatch"; return SGB_ERR_INVALID; }
  SGB_CUDA(cudaSetDevice(h->device));
  for (int r = 0; r < world; ++r) {
    if (r == h->LP.rank) continue;
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, (const char*)handles + 64 * (size_t)r, 64);
    void* p = h->peer_map[r];
    if (!p || std::memcmp(&mh, &h->peer_handle[r], 64) != 0) {  // a new (or re-allocated) peer arena
      if (p) cudaIpcCloseMemHandle(p);
      h->peer_map[r] = p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        h->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
        return SGB_ERR_COMM;
      }
      h->peer_map[r] = p;
      h->peer_handle[r] = mh;
    }
    h->peer_base[r] = p;
    // the peer's arena layout sits in its first bytes (written by its sgb_set_graph, which synchronised its stream)
    SGB_CUDA(cudaMemcpy(&h->layout[r], p, sizeof(sgb_handle::ArenaHeader), cudaMemcpyDeviceToHost));
    if (h->layout[r].magic != sgb_handle::kArenaMagic) { h->err = "peer arena without a layout header (version mismatch?)"; return SGB_ERR_COMM; }
    // pushed halos: the rows this rank pushes to r must be exactly the slots r reserved for it
    if (h->LP.pushed && h->layout[r].halo_cnt[h->LP.rank] != h->LP.send_cnt[r]) {
      h->err = "halo plan mismatch between ranks (" + std::to_string(h->LP.send_cnt[r]) + " rows to push, " +
               std::to_string(h->layout[r].halo_cnt[h->LP.rank]) + " slots reserved by the peer)";
      return SGB_ERR_COMM;
    }
  }
  fill_peer_tables(h);
  h->connected = true;
  return SGB_OK;
}

sgb_status sgb_get_partition_info(const sgb_handle* h, sgb_partition_info* o) {
  if (!h || !o) return SGB_ERR_INVALID;
  if (!h->has_graph) return SGB_ERR_NOT_INITIALIZED;
  const LocalPlan& P = h->LP;
  o->world = P.world; o->rank = P.rank; o->n_poses = P.nP; o->n_landmarks = P.nL;
  o->n_pp = P.n_pp; o->n_pl = P.n_pl; o->n_pp_owned = P.n_pp_owned; o->n_pl_owned = P.n_pl_owned;
  o->halo_pose_gathers = P.halo_p; o->halo_landmark_gathers = P.halo_t;
  o->hpp_entries = P.Hpp.entries(); o->hpl_entries = P.Hpl.entries(); o->hlp_entries = P.Hlp.entries();
  int64_t real = 0;
  for (int32_t c : P.Hpp.col) real += c >= 0;
  o->hpp_blocks = real;
  real = 0;
  for (int32_t c : P.Hpl.col) real += c >= 0;
  o->hpl_blocks = real;
  return SGB_OK;
}

sgb_status sgb_get_structure_info(const sgb_handle* h, sgb_structure_info* o) {
  if (!h || !o) return SGB_ERR_INVALID;
  if (!h->has_graph) return SGB_ERR_NOT_INITIALIZED;
  const Structure& S = h->S;
  o->n_free = S.Pf + S.Lf;
  o->n_free_poses = S.Pf;
  o->n_free_landmarks = S.Lf;
  o->n_blocks = (int32_t)S.n_blocks;
  o->scalar_dim = S.dim;
  o->n_active_pp = S.n_pp;
  o->n_active_pl = S.n_pl;
  o->coarse_nodes = h->G.cz_h > 0 ? h->G.cz_nn : 0;
  o->block_values = S.block_values;
  return SGB_OK;
}

sgb_status sgb_get_structure(const sgb_handle* h, int32_t* kind, int32_t* index, int32_t* offset, int32_t* br,
                             int32_t* bc, int32_t* bnr, int32_t* bnc, int32_t* ph, int32_t* lh) {
  if (!h) return SGB_ERR_INVALID;
  if (!h->has_graph) return SGB_ERR_NOT_INITIALIZED;
  if ((br || bc || bnr || bnc) && h->S.filtered) {
    const_cast<sgb_handle*>(h)->err = "the block list needs the whole structure on this rank: set SGB_PARTITION_FULL=1";
    return SGB_ERR_UNSUPPORTED;
  }
  if (br || bc || bnr || bnc)  // the Hessian order and the per-vertex indices alone need no block list
    build_block_list(const_cast<sgb_handle*>(h)->S);  // built on first use; the handle is not shared between threads
  const Structure& S = h->S;
  auto cp = [](int32_t* dst, const std::vector<int32_t>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(int32_t)); };
  cp(kind, S.ord_kind); cp(index, S.ord_index); cp(offset, S.ord_offset);
  cp(br, S.blk_row); cp(bc, S.blk_col); cp(bnr, S.blk_nr); cp(bnc, S.blk_nc);
  if (ph) for (int i = 0; i < S.P_all; ++i) ph[i] = S.pose_h[i];
  if (lh) for (int i = 0; i < S.L_all; ++i) lh[i] = S.lm_h[i] >= 0 ? S.Pf + S.lm_h[i] : -1;
  return SGB_OK;
}

// copies this rank's owned part of a Hessian-ordered vector (pose part from `dp` [3*nP], landmark part from `dl` [2*nL])
static sgb_status gather_owned_vector(sgb_handle* h, const double* dp, const double* dl, double* out) {
  const Structure& S = h->S;
  const LocalPlan& P = h->LP;
  std::fill(out, out + S.dim, 0.0);
  if (P.nP) SGB_CUDA(cudaMemcpy(out + 3 * (size_t)P.p_begin, dp, 3 * (size_t)P.nP * sizeof(double), cudaMemcpyDeviceToHost));
  if (P.nL) {
    std::vector<double> tmp(2 * (size_t)P.nL);
    SGB_CUDA(cudaMemcpy(tmp.data(), dl, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int l = 0; l < P.nL_owned; ++l) {  // ghost rows belong to their owners' share
      size_t o = 3 * (size_t)S.Pf + 2 * (size_t)P.lm_global[l];
      out[o] = tmp[2 * (size_t)l];
      out[o + 1] = tmp[2 * (size_t)l + 1];
    }
  }
  return SGB_OK;
}

sgb_status sgb_linearize(sgb_handle* h, double* b, double* Hblocks, double* chi2) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  if ((st = launch_linearize(h)) != SGB_OK) return st;
  if ((st = launch_finalize_lin(h, 0)) != SGB_OK) return st;
  if ((st = read_scalars(h)) != SGB_OK) return st;
  const Structure& S = h->S;
  const LocalPlan& P = h->LP;
  DevGraph& G = h->G;
  if (chi2) { chi2[0] = h->h_sc->chi2; chi2[1] = h->h_sc->chi2_robust; }
  if (b && (st = gather_owned_vector(h, G.b_p, G.b_l[P.rank], b)) != SGB_OK) return st;
  if (Hblocks && h->S.filtered) {
    h->err = "the Hessian export needs the whole structure on this rank: set SGB_PARTITION_FULL=1";
    return SGB_ERR_UNSUPPORTED;
  }
  if (Hblocks) {
    build_export(h->S, h->LP);
    std::vector<double> hpp((size_t)P.Hpp.entries() * 9), hpl((size_t)P.Hpl.entries() * 6), hll(3 * (size_t)P.nL);
    if (!hpp.empty()) SGB_CUDA(cudaMemcpy(hpp.data(), G.Hpp.vals, hpp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (!hpl.empty()) SGB_CUDA(cudaMemcpy(hpl.data(), G.Hpl.vals, hpl.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (!hll.empty()) SGB_CUDA(cudaMemcpy(hll.data(), G.Hll, hll.size() * sizeof(double), cudaMemcpyDeviceToHost));
    size_t o = 0;
    for (size_t k = 0; k < S.blk_row.size(); ++k) {
      int kind = S.blk_kind[k], e = P.blk_entry[k];
      int sz = S.blk_nr[k] * S.blk_nc[k];
      if (P.blk_owner[k] != P.rank) {  // owned by another rank: left zero (sum over ranks = full matrix)
        for (int c = 0; c < sz; ++c) Hblocks[o + c] = 0.0;
      } else if (kind == 0) {  // 3x3 row-major on the device -> column-major out
        for (int c = 0; c < 3; ++c)
          for (int r = 0; r < 3; ++r) Hblocks[o + c * 3 + r] = hpp[sell_vaddr(e, 9, 3 * r + c)];
      } else if (kind == 1) {  // 3x2
        for (int c = 0; c < 2; ++c)
          for (int r = 0; r < 3; ++r) Hblocks[o + c * 3 + r] = hpl[sell_vaddr(e, 6, 2 * r + c)];
      } else {
        double h11 = hll[e], h12 = hll[(size_t)P.nL + e], h22 = hll[2 * (size_t)P.nL + e];
        Hblocks[o] = h11; Hblocks[o + 1] = h12; Hblocks[o + 2] = h12; Hblocks[o + 3] = h22;
      }
      o += sz;
    }
  }
  return SGB_OK;
}

sgb_status sgb_solve_once(sgb_handle* h, double lambda, double* x, int32_t* pcg_iters, double* rel) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  if ((st = launch_linearize(h)) != SGB_OK) return st;
  if ((st = launch_finalize_lin(h, 0)) != SGB_OK) return st;
  if ((st = launch_setup(h, lambda, 1)) != SGB_OK) return st;
  if ((st = launch_pcg(h, lambda, 1)) != SGB_OK) return st;
  if (h->G.nL > 0) {
    k_backsub<<<grid_for(32 * h->G.Hlp.nslices), kThreads, 0, h->stream>>>(h->G);
    h->tm.kernel_launches++;
  }
  k_gn_control<<<1, 32, 0, h->stream>>>(h->G, h->d_sc);
  h->tm.kernel_launches++;
  if ((st = read_scalars(h)) != SGB_OK) return st;
  bool ok = h->h_sc->result == 1;
  if (pcg_iters) *pcg_iters = h->h_sc->pcg_iters;
  if (rel) *rel = h->h_sc->pcg_rel;
  if (x && (st = gather_owned_vector(h, h->G.x_p[h->LP.rank], h->G.x_l, x)) != SGB_OK) return st;
  if (!ok) {
    h->err = "linear solve failed (system not positive definite)";
    return SGB_ERR_SOLVE_FAILED;
  }
  return SGB_OK;
}


sgb_status sgb_linear_set_pattern(sgb_handle* h, const sgb_block_matrix* A) {
  if (!h || !A || A->n_block_cols <= 0 || !A->block_dim || !A->col_ptr || !A->row_idx) return SGB_ERR_INVALID;
  const int n = A->n_block_cols;
  int n3 = 0;
  while (n3 < n && A->block_dim[n3] == 3) ++n3;
  for (int i = n3; i < n; ++i)
    if (A->block_dim[i] != 2) { h->err = "linear: block dimensions must be 3 (poses) followed by 2 (landmarks)"; return SGB_ERR_UNSUPPORTED; }
  const int n2 = n - n3;
  if (n3 == 0) { h->err = "linear: a matrix without 3x3 pose blocks is not supported"; return SGB_ERR_UNSUPPORTED; }
  // The matrix of a graph whose edges are its off-diagonal blocks has exactly this pattern: build that graph and let
  // the ordinary symbolic phase lay it out. A vertex without off-diagonal blocks hangs from an extra FIXED pose (no row).
  std::vector<int32_t> pp_i, pp_j, pl_p, pl_l, blk_edge((size_t)A->col_ptr[n], -1), deg(n, 0);
  std::vector<char> has_diag(n, 0);
  for (int c = 0; c < n; ++c) {
    int prev = -1;
    for (int q = A->col_ptr[c]; q < A->col_ptr[c + 1]; ++q) {
      int r = A->row_idx[q];
      if (r < 0 || r > c || r <= prev) { h->err = "linear: rows must be ascending and in the upper triangle"; return SGB_ERR_INVALID; }
      prev = r;
      if (r == c) { has_diag[c] = 1; continue; }
      ++deg[r]; ++deg[c];
      if (c < n3) { blk_edge[q] = (int32_t)pp_i.size(); pp_i.push_back(r); pp_j.push_back(c); }
      else if (r < n3) { blk_edge[q] = (int32_t)pl_p.size(); pl_p.push_back(r); pl_l.push_back(c - n3); }
      else { h->err = "linear: landmark-landmark off-diagonal blocks are not supported"; return SGB_ERR_UNSUPPORTED; }
    }
    if (!has_diag[c]) { h->err = "linear: missing diagonal block"; return SGB_ERR_INVALID; }
  }
  bool need_anchor = false;
  for (int v = 0; v < n; ++v) need_anchor |= deg[v] == 0;
  const int P = n3 + (need_anchor ? 1 : 0);
  if (need_anchor)
    for (int v = 0; v < n; ++v)
      if (deg[v] == 0) {
        if (v < n3) { pp_i.push_back(n3); pp_j.push_back(v); }
        else { pl_p.push_back(n3); pl_l.push_back(v - n3); }
      }
  std::vector<double> pose_est(3 * (size_t)P, 0.0), lm_est(2 * (size_t)std::max(n2, 1), 0.0);
  std::vector<uint8_t> fixed(P, 0);
  if (need_anchor) fixed[n3] = 1;
  std::vector<double> zpp(3 * pp_i.size() + 1, 0.0), ipp(6 * pp_i.size() + 1, 0.0), zpl(2 * pl_p.size() + 1, 0.0), ipl(3 * pl_p.size() + 1, 0.0);
  sgb_graph_soa g;
  std::memset(&g, 0, sizeof g);
  g.n_poses = P; g.pose_est = pose_est.data(); g.pose_fixed = fixed.data();
  g.n_landmarks = n2; g.lm_est = lm_est.data();
  g.n_pp = (int32_t)pp_i.size(); g.pp_i = pp_i.data(); g.pp_j = pp_j.data(); g.pp_z = zpp.data(); g.pp_info = ipp.data();
  g.n_pl = (int32_t)pl_p.size(); g.pl_pose = pl_p.data(); g.pl_lm = pl_l.data(); g.pl_z = zpl.data(); g.pl_info = ipl.data();
  sgb_status st = set_graph_impl(h, &g, 1, 0);
  if (st != SGB_OK) return st;
  const Structure& S = h->S;
  const LocalPlan& LP = h->LP;
  if (S.Pf != n3 || S.Lf != n2) { h->err = "linear: internal error (free vertex count)"; return SGB_ERR_INVALID; }
  std::vector<int32_t> loc_pp(pp_i.size(), -1), loc_pl(pl_p.size(), -1), lm_local(n2, -1);
  for (int k = 0; k < LP.n_pp; ++k) loc_pp[S.pp_src[LP.pp_g[k]]] = k;
  for (int k = 0; k < LP.n_pl; ++k) loc_pl[S.pl_src[LP.pl_g[k]]] = k;
  for (int l = 0; l < LP.nL; ++l) lm_local[LP.lm_global[l]] = l;
  const int nb = A->col_ptr[n];
  std::vector<int32_t> off(nb), e1(nb, -1), e2(nb, -1), kind(nb);
  int64_t o = 0;
  for (int c = 0; c < n; ++c)
    for (int q = A->col_ptr[c]; q < A->col_ptr[c + 1]; ++q) {
      int r = A->row_idx[q];
      if (o > INT32_MAX) { h->err = "linear: more than 2^31 values"; return SGB_ERR_UNSUPPORTED; }
      off[q] = (int32_t)o;
      o += (int64_t)A->block_dim[r] * A->block_dim[c];
      if (r == c) {
        if (c < n3) { kind[q] = 0; e1[q] = LP.hpp_diag[c]; }
        else { kind[q] = 1; e1[q] = lm_local[c - n3]; }
      } else if (c < n3) {
        int k = loc_pp[blk_edge[q]];
        kind[q] = 2; e1[q] = LP.pp_e_ij[k]; e2[q] = LP.pp_e_ji[k];
      } else {
        int k = loc_pl[blk_edge[q]];
        kind[q] = 3; e1[q] = LP.pl_e_pl[k]; e2[q] = LP.pl_e_lp[k];
      }
      if (e1[q] < 0 || (kind[q] >= 2 && e2[q] < 0)) { h->err = "linear: internal error (block without a slot)"; return SGB_ERR_INVALID; }
    }
  auto& L = h->lin;
  L.n_blocks = nb; L.n3 = n3; L.n2 = n2; L.n_values = o;
  if ((st = upload(h, &L.d_off, off)) != SGB_OK) return st;
  if ((st = upload(h, &L.d_e1, e1)) != SGB_OK) return st;
  if ((st = upload(h, &L.d_e2, e2)) != SGB_OK) return st;
  if ((st = upload(h, &L.d_kind, kind)) != SGB_OK) return st;
  if ((st = upload(h, &L.d_lmg, LP.lm_global)) != SGB_OK) return st;
  if ((st = dalloc(h, &L.d_vals, (size_t)o)) != SGB_OK) return st;
  if ((st = dalloc(h, &L.d_b, (size_t)S.dim)) != SGB_OK) return st;
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  L.valid = true;
  return SGB_OK;
}

sgb_status sgb_linear_solve(sgb_handle* h, const double* values, const double* b, double* x, int32_t* pcg_iters, double* rel) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  if (!h->lin.valid) { h->err = "linear: call sgb_linear_set_pattern first"; return SGB_ERR_NOT_INITIALIZED; }
  if (!values || !b || !x) return SGB_ERR_INVALID;
  SGB_CUDA(cudaSetDevice(h->device));
  auto& L = h->lin;
  std::memset(&h->tm, 0, sizeof h->tm);
  if ((st = h2d(h, L.d_vals, values, (size_t)L.n_values * sizeof(double))) != SGB_OK) return st;
  if ((st = h2d(h, L.d_b, b, (size_t)h->S.dim * sizeof(double))) != SGB_OK) return st;
  cudaEvent_t t0 = h->ev.e[0], t1 = h->ev.e[1];
  SGB_CUDA(cudaEventRecord(t0, h->stream));
  k_scatter_blocks<<<grid_for(L.n_blocks), kThreads, 0, h->stream>>>(h->G, L.d_vals, L.n_blocks, L.d_off, L.d_e1, L.d_e2, L.d_kind);
  k_scatter_rhs<<<grid_for(h->S.dim), kThreads, 0, h->stream>>>(h->G, L.d_b, L.d_lmg);
  h->tm.kernel_launches += 2;
  SGB_CUDA(cudaGetLastError());
  if ((st = launch_setup(h, 0.0, 1)) != SGB_OK) return st;   // lambda is already inside A (BlockSolver::setLambda)
  if ((st = launch_pcg(h, 0.0, 1)) != SGB_OK) return st;
  if (h->G.nL > 0) {
    k_backsub<<<grid_for(32 * h->G.Hlp.nslices), kThreads, 0, h->stream>>>(h->G);
    h->tm.kernel_launches++;
  }
  k_gn_control<<<1, 32, 0, h->stream>>>(h->G, h->d_sc);
  h->tm.kernel_launches++;
  SGB_CUDA(cudaEventRecord(t1, h->stream));
  if ((st = read_scalars(h)) != SGB_OK) return st;
  h->tm.total_ms = ev_ms(t0, t1);
  h->tm.pcg_iters = h->h_sc->pcg_iters;
  if (pcg_iters) *pcg_iters = h->h_sc->pcg_iters;
  if (rel) *rel = h->h_sc->pcg_rel;
  if ((st = gather_owned_vector(h, h->G.x_p[h->LP.rank], h->G.x_l, x)) != SGB_OK) return st;
  if (h->h_sc->result != 1) {
    // LinearSolver::solve() == false: breakdown, or the iteration cap was hit with a residual above 1e-6 (pcg_usable)
    h->err = h->h_sc->pcg_flag == 1 ? "linear solve failed (PCG hit its iteration cap without converging)"
                                    : "linear solve failed (matrix not positive definite)";
    return SGB_ERR_SOLVE_FAILED;
  }
  return SGB_OK;
}

sgb_status sgb_compute_marginals(sgb_handle* h, int32_t n_blocks, const int32_t* block_row, const int32_t* block_col,
                                 double* out_values) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  if (n_blocks < 0 || (n_blocks > 0 && (!block_row || !block_col || !out_values))) return SGB_ERR_INVALID;
  if (h->LP.world != 1) { h->err = "marginals: single GPU only"; return SGB_ERR_UNSUPPORTED; }
  SGB_CUDA(cudaSetDevice(h->device));
  const Structure& S = h->S;
  const int nf = S.Pf + S.Lf;
  auto bdim = [&](int i) { return i < S.Pf ? 3 : 2; };
  auto boff = [&](int i) { return i < S.Pf ? 3 * i : 3 * S.Pf + 2 * (i - S.Pf); };
  std::vector<size_t> out_off(n_blocks);
  size_t total = 0;
  for (int k = 0; k < n_blocks; ++k) {
    if (block_row[k] < 0 || block_row[k] >= nf || block_col[k] < 0 || block_col[k] >= nf) {
      h->err = "marginals: block index out of range";
      return SGB_ERR_INVALID;
    }
    out_off[k] = total;
    total += (size_t)bdim(block_row[k]) * bdim(block_col[k]);
  }
  if (n_blocks == 0) return SGB_OK;
  // requests grouped by block column: every scalar column of H^-1 is solved for once
  std::vector<int32_t> order(n_blocks);
  for (int k = 0; k < n_blocks; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return block_col[a] < block_col[b]; });
  if ((st = launch_linearize(h)) != SGB_OK) return st;   // H at the current estimates (b_p / b_l are overwritten below)
  if ((st = launch_finalize_lin(h, 0)) != SGB_OK) return st;
  double* d_rhs = nullptr;
  int32_t* d_lmg = nullptr;
  SGB_CUDA(cudaMalloc((void**)&d_rhs, std::max<size_t>(S.dim, 1) * sizeof(double)));
  cudaError_t ce = cudaMalloc((void**)&d_lmg, std::max<size_t>(h->LP.lm_global.size(), 1) * sizeof(int32_t));
  if (ce != cudaSuccess) { cudaFree(d_rhs); h->err = cudaGetErrorString(ce); return SGB_ERR_CUDA; }
  auto cleanup = [&]() { cudaFree(d_rhs); cudaFree(d_lmg); };
  if (!h->LP.lm_global.empty())
    cudaMemcpyAsync(d_lmg, h->LP.lm_global.data(), h->LP.lm_global.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
  std::vector<double> x(S.dim);
  st = SGB_OK;
  for (size_t q = 0; q < order.size() && st == SGB_OK;) {
    const int c = block_col[order[q]];
    size_t q_end = q;
    while (q_end < order.size() && block_col[order[q_end]] == c) ++q_end;
    for (int d = 0; d < bdim(c) && st == SGB_OK; ++d) {
      const double one = 1.0;
      cudaMemsetAsync(d_rhs, 0, (size_t)S.dim * sizeof(double), h->stream);
      cudaMemcpyAsync(d_rhs + boff(c) + d, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream);
      k_scatter_rhs<<<grid_for(S.dim), kThreads, 0, h->stream>>>(h->G, d_rhs, d_lmg);
      h->tm.kernel_launches++;
      if ((st = launch_setup(h, 0.0, 1)) != SGB_OK) break;
      if ((st = launch_pcg(h, 0.0, 1)) != SGB_OK) break;
      if (h->G.nL > 0) {
        k_backsub<<<grid_for(32 * h->G.Hlp.nslices), kThreads, 0, h->stream>>>(h->G);
        h->tm.kernel_launches++;
      }
      k_gn_control<<<1, 32, 0, h->stream>>>(h->G, h->d_sc);
      h->tm.kernel_launches++;
      if ((st = read_scalars(h)) != SGB_OK) break;
      if (h->h_sc->result != 1) {
        h->err = "marginals: Hessian not positive definite";
        st = SGB_ERR_SOLVE_FAILED;
        break;
      }
      if ((st = gather_owned_vector(h, h->G.x_p[0], h->G.x_l, x.data())) != SGB_OK) break;
      for (size_t t = q; t < q_end; ++t) {
        const int k = order[t], r = block_row[k], nr = bdim(r);
        for (int i = 0; i < nr; ++i) out_values[out_off[k] + (size_t)d * nr + i] = x[boff(r) + i];
      }
    }
    q = q_end;
  }
  cudaStreamSynchronize(h->stream);
  cleanup();
  return st;
}

sgb_status sgb_optimize(sgb_handle* h, int32_t algo, int32_t max_iters, int32_t online, int32_t* iters_done,
                        sgb_iter_stat* stats) {
  (void)online;  // the structure is rebuilt by sgb_set_graph; online only skips buildStructure in g2o
  if (iters_done) *iters_done = -1;
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  h->lm_state_valid = false;  // lambda is re-initialised at iteration 0 of every optimize() call
  return do_optimize(h, algo, max_iters, iters_done, stats);
}

static sgb_status copy_into_both(sgb_handle* h, const double* pose, const double* lm, cudaMemcpyKind kind) {
  size_t np = 3 * (size_t)h->S.P_all * sizeof(double), nl = 2 * (size_t)h->S.L_all * sizeof(double);
  int r = h->LP.rank;
  for (int bsel = 0; bsel < 2; ++bsel) {
    if (pose && np) SGB_CUDA(cudaMemcpyAsync(h->G.pose_buf[bsel][r], pose, np, kind, h->stream));
    if (lm && nl) SGB_CUDA(cudaMemcpyAsync(h->G.lm_buf[bsel][r], lm, nl, kind, h->stream));
  }
  return SGB_OK;
}

sgb_status sgb_optimize_resident(sgb_handle* h, int32_t algo, int32_t max_iters, int32_t* iters_done,
                                 sgb_iter_stat* stats) {
  if (iters_done) *iters_done = -1;
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  if ((st = copy_into_both(h, h->d_pose0, h->d_lm0, cudaMemcpyDeviceToDevice)) != SGB_OK) return st;
  if (h->G.world > 1) {  // no rank may start pushing new estimates before every replica has been reset
    k_xbarrier<<<1, 32, 0, h->stream>>>(h->G, h->d_sc);
    SGB_CUDA(cudaGetLastError());
  }
  h->lm_state_valid = false;
  return do_optimize(h, algo, max_iters, iters_done, stats);
}

sgb_status sgb_step(sgb_handle* h, int32_t algo, int32_t iteration, int32_t online, int32_t* result, sgb_iter_stat* stat) {
  (void)online;
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  if (iteration == 0) h->lm_state_valid = false;
  int r = SGB_RESULT_OK;
  st = do_step(h, algo, iteration, &r, stat);
  if (result) *result = r;
  return st;
}

sgb_status sgb_get_estimates(sgb_handle* h, double* pose_est, double* lm_est) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  const DevGraph& G = h->G;
  if (pose_est && h->S.P_all)
    SGB_CUDA(cudaMemcpyAsync(pose_est, G.pose_buf[G.cur][G.rank], 3 * (size_t)h->S.P_all * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (lm_est && h->S.L_all)
    SGB_CUDA(cudaMemcpyAsync(lm_est, G.lm_buf[G.cur][G.rank], 2 * (size_t)h->S.L_all * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  return SGB_OK;
}

sgb_status sgb_set_estimates(sgb_handle* h, const double* pose_est, const double* lm_est) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  if ((st = copy_into_both(h, pose_est, lm_est, cudaMemcpyHostToDevice)) != SGB_OK) return st;
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  return SGB_OK;
}

sgb_status sgb_push(sgb_handle* h) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  double *p = nullptr, *l = nullptr;
  size_t np = 3 * (size_t)h->S.P_all, nl = 2 * (size_t)h->S.L_all;
  if ((st = dalloc_raw(h, &p, np)) != SGB_OK) return st;
  if ((st = dalloc_raw(h, &l, nl)) != SGB_OK) return st;
  const DevGraph& G = h->G;
  if (np) SGB_CUDA(cudaMemcpyAsync(p, G.pose_buf[G.cur][G.rank], np * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (nl) SGB_CUDA(cudaMemcpyAsync(l, G.lm_buf[G.cur][G.rank], nl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->stack.push_back({p, l});
  return SGB_OK;
}

static void release_top(sgb_handle* h) {
  auto top = h->stack.back();
  h->stack.pop_back();
  for (void* q : {(void*)top.first, (void*)top.second}) {
    auto it = std::find(h->allocs.begin(), h->allocs.end(), q);
    if (it != h->allocs.end()) h->allocs.erase(it);
    cudaFree(q);
  }
}

sgb_status sgb_pop(sgb_handle* h) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  if (h->stack.empty()) { h->err = "pop on an empty stack"; return SGB_ERR_INVALID; }
  SGB_CUDA(cudaSetDevice(h->device));
  auto top = h->stack.back();
  if ((st = copy_into_both(h, top.first, top.second, cudaMemcpyDeviceToDevice)) != SGB_OK) return st;
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  release_top(h);
  return SGB_OK;
}

sgb_status sgb_discard_top(sgb_handle* h) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  if (h->stack.empty()) { h->err = "discardTop on an empty stack"; return SGB_ERR_INVALID; }
  SGB_CUDA(cudaSetDevice(h->device));
  SGB_CUDA(cudaStreamSynchronize(h->stream));
  release_top(h);
  return SGB_OK;
}

sgb_status sgb_chi2(sgb_handle* h, double* chi2) {
  sgb_status st = need_graph(h);
  if (st != SGB_OK) return st;
  SGB_CUDA(cudaSetDevice(h->device));
  if ((st = launch_chi2(h, h->G.cur)) != SGB_OK) return st;
  k_finalize_chi<<<1, kThreads, 0, h->stream>>>(h->G, h->d_sc, h->d_part_e, grid_for(h->G.n_pp_owned + h->G.n_pl_owned));
  h->tm.kernel_launches++;
  SGB_CUDA(cudaGetLastError());
  if ((st = read_scalars(h)) != SGB_OK) return st;
  if (chi2) { chi2[0] = h->h_sc->chi2; chi2[1] = h->h_sc->chi2_robust; }
  return SGB_OK;
}

static sgb_status optimize_batch_impl(sgb_handle* const* hs, int32_t n, int32_t algo, int32_t max_iters,
                                      int32_t* iters_done, sgb_iter_stat* last_stats, bool reset) {
  if (!hs || n <= 0) return SGB_ERR_INVALID;
  sgb_handle* h = hs[0];  // errors are reported on the first handle; its stream carries the launch
  if (!h) return SGB_ERR_INVALID;
  if (algo != SGB_ALGO_LM && algo != SGB_ALGO_GN) { h->err = "unknown algorithm"; return SGB_ERR_INVALID; }
  for (int i = 0; i < n; ++i) {
    if (iters_done) iters_done[i] = -1;
    if (!hs[i] || !hs[i]->has_graph) { h->err = "batch: a handle has no graph (initializeOptimization not called)"; return SGB_ERR_NOT_INITIALIZED; }
    if (hs[i]->device != h->device) { h->err = "batch: all handles must live on the same device"; return SGB_ERR_INVALID; }
    if (hs[i]->LP.world != 1) { h->err = "batch: partitioned graphs are not batched"; return SGB_ERR_UNSUPPORTED; }
  }
  SGB_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  for (int i = 0; i < n; ++i) {  // work queued earlier on the other handles' streams must be finished
    if (hs[i] != h) SGB_CUDA(cudaStreamSynchronize(hs[i]->stream));
    if (reset) {
      sgb_handle* q = hs[i];
      size_t np = 3 * (size_t)q->S.P_all * sizeof(double), nl = 2 * (size_t)q->S.L_all * sizeof(double);
      for (int b = 0; b < 2; ++b) {
        if (np) SGB_CUDA(cudaMemcpyAsync(q->G.pose_buf[b][0], q->d_pose0, np, cudaMemcpyDeviceToDevice, s));
        if (nl) SGB_CUDA(cudaMemcpyAsync(q->G.lm_buf[b][0], q->d_lm0, nl, cudaMemcpyDeviceToDevice, s));
      }
    }
  }
  std::vector<BatchItem> items(n);
  int smem_bytes = 0;
  for (int i = 0; i < n; ++i) {
    items[i].g = hs[i]->G;
    items[i].sc = hs[i]->d_sc;
    const ResPlan& r = hs[i]->res_block;
    items[i].res = ResPlanFwd{r.valid, r.bt, r.ncta, r.cap_pp, r.cap_pl, r.cap_lp, r.nz, r.nt, r.bytes, r.cap_sl, r.cap_lr, r.rows_cta, r.cz_nc, r.cz_h};
    if (r.valid) smem_bytes = std::max(smem_bytes, r.bytes);
  }
  static_assert(sizeof(ResPlanFwd) == sizeof(ResPlan), "ResPlanFwd mirrors ResPlan");
  if (smem_bytes > 0) {
    cudaError_t ea = n <= h->sm_count ? cudaFuncSetAttribute(k_lm_block<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
                                      : cudaFuncSetAttribute(k_lm_block<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (ea != cudaSuccess) {  // no room: every graph takes the global-memory solve
      cudaGetLastError();
      smem_bytes = 0;
      for (auto& it : items) it.res.valid = 0;
    }
  }
  BatchItem* d_items = nullptr;
  BatchResult* d_res = nullptr;
  SGB_CUDA(cudaMalloc((void**)&d_items, sizeof(BatchItem) * (size_t)n));
  SGB_CUDA(cudaMalloc((void**)&d_res, sizeof(BatchResult) * (size_t)n));
  std::vector<BatchResult> res(n);
  BatchParams prm;
  prm.algo = algo;
  prm.max_iters = max_iters;
  prm.max_trials = h->opt.lm_max_trials > 0 ? h->opt.lm_max_trials : 10;
  prm.tau = h->opt.lm_tau > 0 ? h->opt.lm_tau : 1e-5;
  prm.user_lambda = h->opt.lm_user_lambda;
  prm.pcg.tol = h->opt.pcg_tolerance > 0 ? h->opt.pcg_tolerance : 1e-10;
  prm.pcg.maxit = h->opt.pcg_max_iters;
  prm.pcg.lambda_override = 0.0;
  prm.pcg.use_override = 0;
  cudaEvent_t t0, t1;
  SGB_CUDA(cudaEventCreate(&t0));
  SGB_CUDA(cudaEventCreate(&t1));
  cudaError_t e = cudaMemcpyAsync(d_items, items.data(), sizeof(BatchItem) * (size_t)n, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaEventRecord(t0, s);
  if (e == cudaSuccess) {
    if (n <= h->sm_count) k_lm_block<8><<<n, kThreads, smem_bytes, s>>>(d_items, prm, d_res);
    else k_lm_block<2><<<n, kThreads, smem_bytes, s>>>(d_items, prm, d_res);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaEventRecord(t1, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(res.data(), d_res, sizeof(BatchResult) * (size_t)n, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, t0, t1);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(d_items);
  cudaFree(d_res);
  if (e != cudaSuccess) {
    h->err = std::string("sgb_optimize_batch: ") + cudaGetErrorString(e);
    return SGB_ERR_CUDA;
  }
  for (int i = 0; i < n; ++i) {
    sgb_handle* q = hs[i];
    q->G.cur = res[i].cur;
    q->lm_state_valid = false;
    std::memset(&q->tm, 0, sizeof q->tm);
    q->tm.total_ms = ms;  // the launch is shared: every handle reports the batch time
    q->tm.pcg_iters = res[i].pcg_iters_total;
    q->tm.trials = res[i].trials;
    q->tm.kernel_launches = (i == 0) ? 1 : 0;
    if (iters_done) iters_done[i] = res[i].iters_done;
    if (last_stats) {
      sgb_iter_stat& st = last_stats[i];
      st.iteration = std::max(0, res[i].iters_done - 1);
      st.trials = res[i].trials;
      st.result = res[i].result;
      st.pcg_iters = res[i].pcg_iters;
      st.chi2 = res[i].chi2;
      st.lambda = res[i].lambda;
      st.rho = res[i].rho;
      st.chi2_before = res[i].chi2_before;
      st.pcg_residual = res[i].pcg_rel;
    }
  }
  return SGB_OK;
}

sgb_status sgb_optimize_batch(sgb_handle* const* handles, int32_t n, int32_t algo, int32_t max_iters, int32_t* iters_done,
                              sgb_iter_stat* last_stats) {
  return optimize_batch_impl(handles, n, algo, max_iters, iters_done, last_stats, false);
}
sgb_status sgb_optimize_batch_resident(sgb_handle* const* handles, int32_t n, int32_t algo, int32_t max_iters,
                                       int32_t* iters_done, sgb_iter_stat* last_stats) {
  return optimize_batch_impl(handles, n, algo, max_iters, iters_done, last_stats, true);
}

sgb_status sgb_get_timings(const sgb_handle* h, sgb_timings* out) {
  if (!h || !out) return SGB_ERR_INVALID;
  *out = h->tm;
  return SGB_OK;
}

}  // extern "C"

namespace sgb {
bool handle_view(sgb_handle* h, HandleView* out) {
  if (!h || !h->has_graph) return false;
  out->device = h->device;
  out->stream = h->stream;
  out->n_poses = h->S.P_all;
  out->n_landmarks = h->S.L_all;
  out->pose_est = h->G.pose_buf[h->G.cur][h->G.rank];
  out->lm_est = h->G.lm_buf[h->G.cur][h->G.rank];
  return true;
}
int handle_device(const sgb_handle* h) { return h ? h->device : -1; }
void handle_set_error(sgb_handle* h, const std::string& msg) {
  if (h) h->err = msg;
}
}  // namespace sgb
