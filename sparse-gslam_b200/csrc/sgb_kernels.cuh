// sgb_kernels.cuh -- CUDA kernels of the hot path (sm_100a). FP64, bandwidth-bound: no tensor cores.
//
//   k_lin_pose / k_lin_lm   linearise + assemble (vertex-centric gather through the host-built symbolic map:
//                           every Hessian block and gradient segment is written exactly once, no atomics,
//                           contributions summed in the reference's edge insertion order)      [SURVEY 8a a9,a10,a15]
//   k_chi2_edges            per-edge error + chi2 with warp-shuffle reductions (coalesced SoA loads) [a20]
//   k_setup_lm / k_setup_pose  damping, batched 2x2 landmark inverses, Schur diagonal + reduced rhs  [a18]
//   k_pcg                   persistent cooperative PCG on the implicit Schur complement, block-Jacobi   [a19 replaced]
//   k_backsub / k_update    landmark back-substitution, oplus into the trial buffers, computeScale   [a20, a7, a8]
//   k_lm_control            Levenberg-Marquardt gain ratio / lambda logic on the device              [a16]
//
// Multi-GPU (one process per GPU, row-block partition): the same kernels; rows are rank-local, remote
// vector entries are gathered through NVLink peer pointers (sgb_types.h), and the scalar reductions / barriers are
// exchanged through peer-mapped mailboxes written with st.release.sys and polled with ld.acquire.sys -- no host
// round trip and no separate collective kernel inside the PCG iteration.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <float.h>

#include "sgb_rows.h"
#include "sgb_coarse.h"

namespace sgb {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 4096;  // partial-sum slots per reduction
// persistent PCG kernel: CTAs per SM the register budget is sized for (the software-pipelined SpMV rows keep two
// blocks' worth of loads in flight per thread and need ~80 registers)
#ifndef SGB_PCG_MIN_BLOCKS
#define SGB_PCG_MIN_BLOCKS 3
#endif

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in every thread; fixed reduction tree (deterministic)
__device__ __forceinline__ double block_sum(double v, double* smem /*[32]*/) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  double r = (lane < nw) ? smem[lane] : 0.0;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ double block_max(double v, double* smem) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  double r = (lane < nw) ? smem[lane] : 0.0;
  r = warp_max(r);
  return r;
}
// every thread of the block gets sum(part[0..n)) in a fixed order
__device__ __forceinline__ double reduce_partials(const double* part, int n, double* smem) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += __ldcg(&part[i]);
  return block_sum(v, smem);
}

// ------------------------------------------------------------------------------------------------ cross-rank
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// One-directional system-scope fences (PTX ISA 8.6+). Acquire side of a relaxed poll that observed a peer's release:
// SASS = CCTL.IVALL only (invalidate L1), no MEMBAR -- against MEMBAR.SC.SYS for __threadfence_system(), which cost
// ~3 us per cross-GPU barrier when it was tried here (2 GPUs, C5: 130 -> 159 us per PCG iteration). Release side:
// MEMBAR.ALL.SYS (everything this thread observed or wrote is performed system-wide before the stores that follow).
__device__ __forceinline__ void fence_acquire_sys() { asm volatile("fence.acquire.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_release_sys() { asm volatile("fence.release.sys;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// All-reduce of up to 4 block-uniform doubles across the ranks (op 0 = sum, 1 = max), also a barrier: everything
// this GPU wrote before the caller's preceding grid-wide sync is visible to the peers once they pass.
// Called by every thread of a block; `pusher` blocks (one per GPU) publish this rank's contribution into every
// peer's mailbox, every calling block then waits for all contributions in its OWN mailbox and combines them in rank
// order (identical result on every block of every rank). seq must be the same on all ranks and increase by one per call.
__device__ __forceinline__ void xreduce(const DevGraph& g, unsigned long long seq, double* vals, int nv, int op, bool pusher) {
  if (g.world == 1) return;
  const int slot = (int)(seq & 1ull);
  const int t = threadIdx.x;
  if (pusher && t < g.world) {
    Mailbox* mb = g.mbox[t];
    for (int k = 0; k < nv; ++k) st_relaxed_sys_f64(&mb->val[slot][g.rank][k], vals[k]);
    st_release_sys_u64(&mb->flag[slot][g.rank], seq);  // release: orders the value stores above (same thread) before the flag
  }
  const Mailbox* me = g.mbox[g.rank];
  if (t < g.world) {
    // relaxed polling, then ONE system-scope acquire fence: the peer published with st.release.sys, so relaxed
    // observation + fence.acquire.sys is the morally-strong acquire pattern of the PTX memory model -- everything the
    // peer wrote before its release is visible afterwards
    while (ld_relaxed_sys_u64(&me->flag[slot][t]) < seq) {
    }
    fence_acquire_sys();
  }
  __syncthreads();
  double out[4];
  for (int k = 0; k < nv; ++k) out[k] = op == 0 ? 0.0 : -DBL_MAX;
  for (int o = 0; o < g.world; ++o)
    for (int k = 0; k < nv; ++k) {
      double v = ld_relaxed_sys_f64(&me->val[slot][o][k]);
      out[k] = op == 0 ? out[k] + v : fmax(out[k], v);
    }
  for (int k = 0; k < nv; ++k) vals[k] = out[k];
}
// pure cross-rank barrier, launched between kernels (after k_update pushed the new estimates to every replica)
__global__ void k_xbarrier(DevGraph g, DevScalars* sc) {
  unsigned long long seq = sc->xseq + 1;
  xreduce(g, seq, nullptr, 0, 0, true);
  if (threadIdx.x == 0) sc->xseq = seq;
}

// ------------------------------------------------------------------------------------------------ linearise
// part layout: [0]=chi [1]=chi_r [2]=maxd, each kMaxBlocks wide
// One thread per pose row. (Four lanes per row -- lin_pose_row_lanes, kept in sgb_rows.h -- were measured on the 1M-pose
// graph: 662 us instead of 357: with consecutive lanes on the same row the writes of the k-th off-diagonal block of
// consecutive rows are no longer coalesced, and the 12 x 4 shuffles of the combine step spill at 128 registers.)
__global__ void __launch_bounds__(kThreads) k_lin_pose(DevGraph g, double* part) {
  __shared__ double sm[32];
  LinAcc acc;
  for (int lp = blockIdx.x * blockDim.x + threadIdx.x; lp < g.nP; lp += gridDim.x * blockDim.x) lin_pose_row(g, lp, acc);
  double c = block_sum(acc.chi, sm), cr = block_sum(acc.chi_r, sm), m = block_max(acc.maxd, sm);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = c;
    part[kMaxBlocks + blockIdx.x] = cr;
    part[2 * kMaxBlocks + blockIdx.x] = m;
  }
}
// Landmark rows, one LANE per stored block of the landmark-major matrix (one warp per slice of the grouped Hlp, like the
// landmark pass of the PCG): the lane evaluates the edge (and the duplicates chained to it) that produces its block,
// writes the block -- 32 consecutive entries per warp step, coalesced -- and contributes B^T Omega B and B^T omega to the
// row's 2x2 diagonal block and gradient. The first build gave a landmark row to ONE thread that walked its ~20 edges one
// after the other: 200 000 threads with a serial chain of dependent gathers each, 0.47 ms on the 1M-pose graph.
// The row sums are formed in a CANONICAL order -- stored blocks in ascending observer order, then the observations from
// fixed poses in insertion order -- whatever the number of lanes the layout gives the row: every step the 32 lanes park
// their contributions in shared memory and the row's first lane adds them in block order. The sums are therefore still
// bit-identical for every partition of the graph (asserted on 2 / 4 / 8 GPUs), though no longer in g2o's insertion order
// (they differ from the serial row body lin_lm_row by rounding; parity with the oracle is checked at 1e-12).
__global__ void __launch_bounds__(kThreads) k_lin_lm(DevGraph g, double* part) {
  __shared__ double sm[32];
  __shared__ double stage[kThreads / 32][5][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const double* pose = cur_pose(g);
  const double* lm = cur_lm(g);
  LinAcc acc;
  auto add_edge = [&](const PLTerm& t, double* c) {  // B^T Omega B (11, 12, 22) and B^T omega_r of one edge
    double BtO[4];
    for (int r = 0; r < 2; ++r) {
      BtO[2 * r] = t.B[r] * t.om[0] + t.B[2 + r] * t.om[1];
      BtO[2 * r + 1] = t.B[r] * t.om[1] + t.B[2 + r] * t.om[2];
    }
    c[0] += BtO[0] * t.B[0] + BtO[1] * t.B[2];
    c[1] += BtO[0] * t.B[1] + BtO[1] * t.B[3];
    c[2] += BtO[2] * t.B[1] + BtO[3] * t.B[3];
    c[3] += t.B[0] * t.omr[0] + t.B[2] * t.omr[1];
    c[4] += t.B[1] * t.omr[0] + t.B[3] * t.omr[1];
  };
  for (int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sl < g.Hlp.nslices; sl += nwarps) {
    const LmSliceMeta m = lm_slice_meta(g, sl);
    const int G = 32 >> m.shift;
    const int row = m.row0 + (lane >> (5 - m.shift));
    const bool writer = (lane & (G - 1)) == 0 && row < m.row_end;
    double h[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < m.steps; ++j) {
      const int e = m.e_begin + 32 * j + lane;
      const int k = SGB_LDG(&g.hlp_edge[e]);
      double c[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (k >= 0) {
        PLTerm t;
        pl_term(g, k, pose, lm, true, true, t);
        add_edge(t, c);
        double blk[6];
        pl_offdiag(t, blk);
        for (int d = SGB_LDG(&g.pl_dup[k]); d >= 0; d = SGB_LDG(&g.pl_dup[d])) {
          PLTerm u;
          pl_term(g, d, pose, lm, true, true, u);
          add_edge(u, c);
          double m2[6];
          pl_offdiag(u, m2);
          for (int q = 0; q < 6; ++q) blk[q] += m2[q];
        }
        for (int q = 0; q < 6; ++q) g.Hlp.vals[sell_vaddr(e, 6, q)] = blk[q];
      }
      for (int q = 0; q < 5; ++q) stage[wib][q][lane] = c[q];
      __syncwarp();
      if (writer)
        for (int kk = 0; kk < G; ++kk)
          for (int q = 0; q < 5; ++q) h[q] += stage[wib][q][lane + kk];
      __syncwarp();
    }
    if (writer) {
      for (int q = SGB_LDG(&g.lfix_ptr[row]); q < SGB_LDG(&g.lfix_ptr[row + 1]); ++q) {  // observations from fixed poses
        PLTerm t;
        pl_term(g, SGB_LDG(&g.lfix[q]), pose, lm, false, true, t);
        add_edge(t, h);
        if (row < g.nL_owned) {  // pose fixed: the edge's chi2 is owned by the landmark row (of the owner)
          acc.chi += t.chi;
          acc.chi_r += t.chi;
        }
      }
      g.Hll[row] = h[0];
      g.Hll[(size_t)g.nL + row] = h[1];
      g.Hll[2 * (size_t)g.nL + row] = h[2];
      g.b_l[g.rank][2 * (size_t)row] = h[3];
      g.b_l[g.rank][2 * (size_t)row + 1] = h[4];
      acc.maxd = fmax(acc.maxd, fmax(fabs(h[0]), fabs(h[2])));
    }
  }
  double c = block_sum(acc.chi, sm), cr = block_sum(acc.chi_r, sm), m = block_max(acc.maxd, sm);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = c;
    part[kMaxBlocks + blockIdx.x] = cr;
    part[2 * kMaxBlocks + blockIdx.x] = m;
  }
}
// sums the partials of the two linearise kernels (and of all ranks); initialises LM state at iteration 0
__global__ void __launch_bounds__(kThreads) k_finalize_lin(DevGraph g, DevScalars* sc, const double* part_p, int nb_p,
                                                          const double* part_l, int nb_l, int init_lambda, double tau,
                                                          double user_lambda) {
  __shared__ double sm[32];
  double v[2];
  v[0] = reduce_partials(part_p, nb_p, sm) + reduce_partials(part_l, nb_l, sm);
  v[1] = reduce_partials(part_p + kMaxBlocks, nb_p, sm) + reduce_partials(part_l + kMaxBlocks, nb_l, sm);
  double m = 0.0;
  for (int i = threadIdx.x; i < nb_p; i += blockDim.x) m = fmax(m, part_p[2 * kMaxBlocks + i]);
  for (int i = threadIdx.x; i < nb_l; i += blockDim.x) m = fmax(m, part_l[2 * kMaxBlocks + i]);
  m = block_max(m, sm);
  unsigned long long seq = sc->xseq;
  __syncthreads();
  xreduce(g, ++seq, v, 2, 0, true);
  xreduce(g, ++seq, &m, 1, 1, true);
  if (threadIdx.x == 0) {
    sc->xseq = seq;
    sc->chi2 = v[0];
    sc->chi2_robust = v[1];
    sc->chi_lin = v[1];
    sc->max_diag = m;
    sc->current_chi = v[1];
    sc->temp_chi = v[1];
    sc->trials = 0;
    sc->rho = 0.0;
    sc->again = 0;
    if (init_lambda) {  // OptimizationAlgorithmLevenberg::computeLambdaInit
      sc->lambda = user_lambda > 0.0 ? user_lambda : tau * m;
      sc->ni = 2.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ chi2 only
// one thread per owned edge, coalesced component-major SoA loads, warp-shuffle + block reduction
__global__ void __launch_bounds__(kThreads) k_chi2_edges(DevGraph g, int buf, double* part) {
  __shared__ double sm[32];
  const double* pose = g.pose_buf[buf][g.rank];
  const double* lm = g.lm_buf[buf][g.rank];
  double c = 0.0, cr = 0.0;
  int n = g.n_pp_owned + g.n_pl_owned;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    if (k < g.n_pp_owned) {
      double a, b;
      pp_chi(g, k, pose, &a, &b);
      c += a;
      cr += b;
    } else {
      double a = pl_chi(g, k - g.n_pp_owned, pose, lm);
      c += a;
      cr += a;
    }
  }
  c = block_sum(c, sm);
  cr = block_sum(cr, sm);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = c;
    part[kMaxBlocks + blockIdx.x] = cr;
  }
}
__global__ void __launch_bounds__(kThreads) k_finalize_chi(DevGraph g, DevScalars* sc, const double* part, int nb) {
  __shared__ double sm[32];
  double v[2];
  v[0] = reduce_partials(part, nb, sm);
  v[1] = reduce_partials(part + kMaxBlocks, nb, sm);
  unsigned long long seq = sc->xseq;
  __syncthreads();
  xreduce(g, ++seq, v, 2, 0, true);
  if (threadIdx.x == 0) {
    sc->xseq = seq;
    sc->chi2 = v[0];
    sc->chi2_robust = v[1];
  }
}

// ------------------------------------------------------------------------------------------------ trial set-up
__global__ void __launch_bounds__(kThreads) k_setup_lm(DevGraph g, DevScalars* sc, double lambda_override, int use_override) {
  double lambda = use_override ? lambda_override : sc->lambda;
  bool ok = true;
  for (int ll = blockIdx.x * blockDim.x + threadIdx.x; ll < g.nL; ll += gridDim.x * blockDim.x) ok &= setup_lm_row(g, ll, lambda);
  if (!ok) atomicOr(&sc->setup_fail, 1);
}
// one thread per block of the block-Jacobi preconditioner (kChunk pose rows)
__global__ void __launch_bounds__(128) k_setup_chunk(DevGraph g, DevScalars* sc, double lambda_override, int use_override) {
  double lambda = use_override ? lambda_override : sc->lambda;
  bool ok = true;
  const int nch = (g.nP + kChunk - 1) / kChunk;
  for (int ch = blockIdx.x * blockDim.x + threadIdx.x; ch < nch; ch += gridDim.x * blockDim.x) ok &= setup_chunk(g, ch, lambda);
  if (!ok) atomicOr(&sc->setup_fail, 1);
}
__global__ void __launch_bounds__(kThreads) k_setup_pose(DevGraph g, DevScalars* sc, double lambda_override, int use_override) {
  double lambda = use_override ? lambda_override : sc->lambda;
  bool ok = true;
  for (int lp = blockIdx.x * blockDim.x + threadIdx.x; lp < g.nP; lp += gridDim.x * blockDim.x) ok &= setup_pose_row(g, lp, lambda);
  if (!ok) atomicOr(&sc->setup_fail, 1);
}

// ------------------------------------------------------------------------------------------------ PCG
// grid-wide barrier for the persistent kernel (all CTAs co-resident: cooperative launch)
__device__ __forceinline__ void grid_sync(unsigned long long* bar, unsigned int nblocks, unsigned long long& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += nblocks;
    __threadfence();
    atomicAdd(bar, 1ull);
    while (*((volatile unsigned long long*)bar) < epoch) {
    }
    __threadfence();
  }
  __syncthreads();
}

struct PcgParams {
  double tol;
  int maxit;
  double lambda_override;
  int use_override;
};

// Grid-wide AND cross-rank all-reduce (sum of up to 2 values) + barrier in ONE step for the persistent kernel:
// every CTA deposits its partial and takes a ticket; the CTA that draws the last ticket reduces the partials in
// index order (deterministic) and publishes the rank's sum into the mailbox of every rank (its own included, so
// world == 1 is the same code); every CTA then polls its own mailbox until all ranks have published and combines the
// values in rank order. One ticket round + one publish replaces "grid barrier, redundant reduction in every CTA,
// second exchange". nv == 0 is a pure barrier. Writes made before the call are visible to every CTA of every rank
// after it (fence before the ticket, fence + release store by the publisher, L1-bypassing loads afterwards).
__device__ __forceinline__ void grid_xreduce(const DevGraph& g, unsigned long long* bar, unsigned int nb,
                                             unsigned long long& epoch, unsigned long long& seq, double* part,
                                             double* vals, int nv, double* sm, int* s_last, double* cl_part = nullptr,
                                             bool local_only = false) {
  if (cl_part != nullptr) {
    // Small graph: the whole grid is ONE thread-block cluster (<= 16 CTAs). The hardware cluster barrier (release /
    // acquire at cluster scope, ~0.2 us) replaces the barrier through global memory (~1.8 us), and the partial sums
    // are read from the other CTAs' shared memory (DSMEM) and added in CTA order. cl_part = this CTA's [2][2] slots,
    // double-buffered by the parity of seq like the global-memory variant.
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    ++seq;
    const int slot = (int)(seq & 1ull);
    if (threadIdx.x == 0)
      for (int k = 0; k < nv; ++k) cl_part[slot * 2 + k] = vals[k];
    cl.sync();
    if (nv > 0) {
      if (threadIdx.x < 32) {
        double v[2] = {0.0, 0.0};
        if (threadIdx.x < nb) {
          const double* rp = cl.map_shared_rank(cl_part, threadIdx.x);
          for (int k = 0; k < nv; ++k) v[k] = rp[slot * 2 + k];
        }
        for (int k = 0; k < nv; ++k) {
          double acc = 0.0;
          for (unsigned int o = 0; o < nb; ++o) acc += __shfl_sync(0xffffffffu, v[k], (int)o);
          if (threadIdx.x == 0) sm[k] = acc;
        }
      }
      __syncthreads();
      for (int k = 0; k < nv; ++k) vals[k] = sm[k];
      __syncthreads();  // sm is reused by the caller's next block_sum
    }
    return;
  }
  if (bar == nullptr) {  // the whole solve runs in ONE thread block (batched small graphs): a block barrier is enough,
    __syncthreads();     // the values are already block-wide sums
    return;
  }
  epoch += nb;
  // A barrier that only spans this GPU must not consume a cross-rank sequence number: the mailbox slots alternate with
  // the parity of seq, and "a rank is at most one cross-rank barrier ahead" only holds if consecutive CROSS-RANK
  // barriers take consecutive numbers (with the local barrier counted, phases C(k) and B(k+1) shared a slot and a fast
  // rank overwrote values a slow rank was still waiting for: the 2-GPU C5 run with ghost landmarks hung). Its own
  // partial-sum slot alternates with the number of barriers this CTA has passed.
  if (!(g.world == 1 || local_only)) ++seq;
  const int slot = (g.world == 1 || local_only) ? (int)((epoch / nb) & 1ull) : (int)(seq & 1ull);
  if (g.world == 1 || local_only) {
    // One GPU (or a barrier that only has to span this GPU: with ghost landmarks nobody reads another rank's t):
    // flat barrier, gpu scope only. Every CTA deposits its partial (double-buffered by the parity of seq:
    // a CTA can be at most one barrier ahead of the slowest reader), arrives with a release reduction, spins on an
    // acquire load, and then sums ALL partials itself in index order -- the same value in every CTA, and no
    // "last CTA reduces and publishes" hop on the critical path.
    double* mypart = part + (size_t)slot * 2 * kMaxBlocks;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 0; k < nv; ++k) __stcg(&mypart[k * kMaxBlocks + blockIdx.x], vals[k]);
      red_release_gpu_add_u64(bar, 1ull);
      while (ld_acquire_gpu_u64(bar) < epoch) {
      }
    }
    __syncthreads();
    for (int k = 0; k < nv; ++k) vals[k] = reduce_partials(mypart + k * kMaxBlocks, (int)nb, sm);
    return;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < nv; ++k) part[k * kMaxBlocks + blockIdx.x] = vals[k];
    __threadfence();
    unsigned long long ticket = atomicAdd(bar, 1ull);
    *s_last = (ticket == epoch - 1ull) ? 1 : 0;
  }
  __syncthreads();
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  if (*s_last) {
    double loc[2] = {0.0, 0.0};
    if (nv > 0) {  // a pure barrier (nv == 0) publishes right away
      __threadfence();
      for (int k = 0; k < nv; ++k) loc[k] = reduce_partials(part + k * kMaxBlocks, (int)nb, sm);
    }
    if (threadIdx.x < g.world) {
      Mailbox* mb = g.mbox[threadIdx.x];
      if (nv == 0) {
        st_release_sys_u64(&mb->flag[slot][g.rank], seq);  // release: through the CTA barriers and the ticket, everything
                                                            // this GPU wrote in the phase
      } else {
        // everything this GPU wrote in the phase went to its OWN memory (peers read it through NVLink): one system
        // release fence orders it, then the values travel with their own flags -- no second fence behind a remote store
        fence_release_sys();
        for (int k = 0; k < nv; ++k) {
          unsigned long long bits = (unsigned long long)__double_as_longlong(loc[k]);
          st_relaxed_sys_u64(&mb->ll[slot][g.rank][k][0], (bits & 0xffffffffull) | tag);
          st_relaxed_sys_u64(&mb->ll[slot][g.rank][k][1], (bits >> 32) | tag);
        }
      }
    }
  }
  const Mailbox* me = g.mbox[g.rank];
  if (nv == 0) {
    if (threadIdx.x < g.world) {
      while (ld_relaxed_sys_u64(&me->flag[slot][threadIdx.x]) < seq) {
      }
      fence_acquire_sys();  // acquire side of the peer's st.release.sys (see xreduce)
    }
    __syncthreads();
    return;
  }
  __syncthreads();  // the last CTA's reduce_partials is done with sm
  if (threadIdx.x < g.world) {
    for (int k = 0; k < nv; ++k) {
      unsigned long long lo, hi;
      do {
        lo = ld_relaxed_sys_u64(&me->ll[slot][threadIdx.x][k][0]);
        hi = ld_relaxed_sys_u64(&me->ll[slot][threadIdx.x][k][1]);
      } while ((lo & 0xffffffff00000000ull) != tag || (hi & 0xffffffff00000000ull) != tag);
      sm[threadIdx.x * 2 + k] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    }
    fence_acquire_sys();  // the publisher fenced (release, system scope) before its flag-in-data stores: acquire side
  }
  __syncthreads();
  for (int k = 0; k < nv; ++k) {
    double acc = 0.0;
    for (int o = 0; o < g.world; ++o) acc += sm[o * 2 + k];  // rank order: the same sum on every CTA of every rank
    vals[k] = acc;
  }
  __syncthreads();  // sm is reused by the caller's next block_sum
}

// Preconditioned conjugate gradient on S = (Hpp + lambda I) - Hpl (Hll + lambda I)^-1 Hpl^T, M = blockdiag(S).
// One launch per GPU runs the whole solve; every CTA of every rank evaluates the same scalars from the same sums, so
// control flow is uniform across the grid and across the GPUs without a host round trip or a collective kernel.
//
// Single-reduction (Chronopoulos-Gear) form: the operator is applied to the preconditioned residual z (kept in the
// peer-visible slot g.p, which other ranks gather from), and the search direction d and s = S d follow by recurrence:
//   gamma = r.z ; beta = gamma / gamma_old ; w = S z ; delta = z.w ; d = z + beta d ; s = w + beta s
//   alpha = gamma / (delta - beta gamma / alpha_old) ; x += alpha d ; r -= alpha s ; z = M^-1 r
// Mathematically the textbook PCG iterates (tests: same iteration counts and step accuracy on the C1-C3 graphs), with
// THREE grid-wide (and cross-GPU) synchronisations per iteration instead of four, and the direction update folded
// into the operator pass so that w never travels to memory:
//   phase A  t = W Hlp^T z (landmark rows)                     | barrier: every rank's t is complete
//   phase B  w = (Hpp + lambda) z - Hpl t, z.w, d and s rows   | barrier + all-reduce of delta
//   phase C  x += alpha d, r -= alpha s, z = M^-1 r, r.z       | barrier + all-reduce of gamma: every z is complete
struct PcgOut {
  double gam0, gam;
  int iters, flag;
};
// The solve itself, shared by the persistent multi-CTA kernel (k_pcg: tid / nthreads span the grid, `bar` is the grid
// barrier counter) and the one-CTA-per-graph batched kernel (k_lm_block: tid / nthreads span the block, bar == NULL).
// ph_ns / t_ph: shared-memory phase timers of the calling CTA.
template <int U = 2>  // blocks in flight per thread in the pose-major pass (see schur_phaseB_row_u)
__device__ __forceinline__ void pcg_solve(const DevGraph& g, const int tid, const int nthreads, const unsigned int nb,
                                          double* part, unsigned long long* bar, unsigned long long& seq,
                                          const PcgParams& prm, const double lambda, double* sm, int* s_last,
                                          unsigned long long* ph_ns, unsigned long long* t_ph, PcgOut& out,
                                          double* cl_part = nullptr) {
#define SGB_PHASE_LAP(i)                        \
  do {                                          \
    if (threadIdx.x == 0) {                     \
      unsigned long long _t = globaltimer_ns(); \
      ph_ns[i] += _t - *t_ph;                   \
      *t_ph = _t;                               \
    }                                           \
  } while (0)
  unsigned long long epoch = 0;
  double* x = g.x_p[g.rank];
  double* zin = g.p[g.rank];

  // x = 0, r = bt, z = M^-1 r, d = s = 0. The kChunk lanes of a preconditioner block run the loop together (the
  // condition is uniform inside such a group: row and lane index agree in their low bits).
  double acc = 0.0;
  for (int lp = tid; (lp & ~(kChunk - 1)) < g.nP; lp += nthreads) {
    const bool act = lp < g.nP;
    double r[3] = {0.0, 0.0, 0.0}, z[3];
    if (act)
      for (int c = 0; c < 3; ++c) r[c] = g.bt[3 * (size_t)lp + c];
    acc += precond_row_shfl(g, lp, act, r, z);
    if (act) {
      for (int c = 0; c < 3; ++c) {
        size_t o = 3 * (size_t)lp + c;
        x[o] = 0.0;
        g.r[o] = r[c];
        zin[o] = z[c];
        g.d[o] = 0.0;
        g.s[o] = 0.0;
      }
      if (g.pushed) {
        const double x0[3] = {0.0, 0.0, 0.0};
        push_halo_row(g, lp, z, x0);
      }
    }
  }
  double gam = block_sum(acc, sm);
  grid_xreduce(g, bar, nb, epoch, seq, part, &gam, 1, sm, s_last, cl_part);  // also: every rank's z segment is complete
  SGB_PHASE_LAP(3);
  const double gam0 = gam, target = prm.tol * prm.tol * gam0;
  double gam_old = 0.0, alpha_old = 0.0;
  int it = 0, flag = 1;
  if (!(gam0 > 0.0)) {
    flag = (gam0 == 0.0) ? 0 : 2;  // zero right-hand side: x = 0 is exact; negative / NaN: M not SPD
  } else {
    while (true) {
      if (!(gam == gam)) {
        flag = 2;
        break;
      }
      if (gam <= target) {
        flag = 0;
        break;
      }
      if (it >= prm.maxit) break;  // flag 1
      const double beta = it == 0 ? 0.0 : gam / gam_old;
      if (g.capL > 0) {
        lm_slices_pass(g, tid >> 5, nthreads >> 5, 0);
        // every rank's t segment is complete; with ghost landmarks t is only read on the GPU that wrote it
        grid_xreduce(g, bar, nb, epoch, seq, part, nullptr, 0, sm, s_last, cl_part, g.world > 1 && g.ghosts != 0);
      }
      SGB_PHASE_LAP(0);
      acc = schur_phaseB_rows_u<U>(g, tid, nthreads, lambda, beta);
      double del = block_sum(acc, sm);
      grid_xreduce(g, bar, nb, epoch, seq, part, &del, 1, sm, s_last, cl_part);
      SGB_PHASE_LAP(1);
      const double denom = it == 0 ? del : del - beta * gam / alpha_old;  // = d.S d
      if (!(denom > 0.0)) {  // S not positive definite (or NaN): g2o's "Cholesky failure" analogue
        flag = 2;
        break;
      }
      const double alpha = gam / denom;
      acc = 0.0;
      for (int lp = tid; (lp & ~(kChunk - 1)) < g.nP; lp += nthreads) {
        const bool act = lp < g.nP;
        double r[3] = {0.0, 0.0, 0.0}, z[3], xn[3] = {0.0, 0.0, 0.0};
        if (act)
          for (int c = 0; c < 3; ++c) {
            size_t o = 3 * (size_t)lp + c;
            xn[c] = x[o] + alpha * g.d[o];
            x[o] = xn[c];
            r[c] = g.r[o] - alpha * g.s[o];
            g.r[o] = r[c];
          }
        acc += precond_row_shfl(g, lp, act, r, z);
        if (act) {
          for (int c = 0; c < 3; ++c) zin[3 * (size_t)lp + c] = z[c];
          if (g.pushed) push_halo_row(g, lp, z, xn);
        }
      }
      ++it;
      gam_old = gam;
      alpha_old = alpha;
      gam = block_sum(acc, sm);
      grid_xreduce(g, bar, nb, epoch, seq, part, &gam, 1, sm, s_last, cl_part);  // also: every rank's z segment is complete
      SGB_PHASE_LAP(2);
    }
  }
#undef SGB_PHASE_LAP
  out.gam0 = gam0;
  out.gam = gam;
  out.iters = it;
  out.flag = flag;
}

// BT = threads per CTA. The persistent grid is three CTAs per SM; a thread owns the rows tid, tid + nthreads, ... and a
// phase lasts as long as its busiest thread, so the phase time is quantised in whole rows: 125 000 rows (one of eight
// ranks of the 1M-pose graph) over 3 x 148 x 256 = 113 664 threads is TWO rows for every thread that matters, over
// 3 x 148 x 288 = 127 872 threads it is one. The host picks the block size that minimises ceil(rows / threads) (the
// register budget shrinks with the block: 80 / 72 / 64 registers for 256 / 288 / 320 threads).
template <int BT>
__global__ void __launch_bounds__(BT, SGB_PCG_MIN_BLOCKS) k_pcg(DevGraph g, DevScalars* sc, double* part,
                                                                unsigned long long* bar, PcgParams prm) {
  __shared__ double sm[32];
  __shared__ int s_last;
  // wall time of the three phases of an iteration as seen by thread 0 of each CTA (barrier waits included); CTA 0's
  // copy is reported. Kept in shared memory: it is touched once per phase and must not cost registers.
  __shared__ unsigned long long ph_ns[4], t_ph;
  if (threadIdx.x == 0) {
    ph_ns[0] = ph_ns[1] = ph_ns[2] = ph_ns[3] = 0;
    t_ph = globaltimer_ns();
  }
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long seq = sc->xseq;
  const double lambda = prm.use_override ? prm.lambda_override : sc->lambda;
  PcgOut out;
  pcg_solve(g, tid, gridDim.x * blockDim.x, gridDim.x, part, bar, seq, prm, lambda, sm, &s_last, ph_ns, &t_ph, out);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) sc->pcg_phase_ns[i] += ph_ns[i];
    sc->xseq = seq;
    sc->rz0 = out.gam0;
    sc->rz = out.gam;
    sc->pcg_iters = out.iters;
    sc->pcg_flag = out.flag;
    sc->pcg_rel = out.gam0 > 0.0 ? sqrt(fabs(out.gam) / out.gam0) : 0.0;
  }
}

// The same kernel for a graph whose grid is at most one CTA per SM (every thread owns at most one pose row and the
// graph sits in L2): occupancy is irrelevant, the dependent-latency chain of a row is what counts, so the pose-major
// pass keeps eight blocks in flight (up to 255 registers).
__global__ void __launch_bounds__(kThreads, 1) k_pcg_small(DevGraph g, DevScalars* sc, double* part, unsigned long long* bar,
                                                          PcgParams prm) {
  __shared__ double sm[32];
  __shared__ int s_last;
  __shared__ unsigned long long ph_ns[4], t_ph;
  if (threadIdx.x == 0) {
    ph_ns[0] = ph_ns[1] = ph_ns[2] = ph_ns[3] = 0;
    t_ph = globaltimer_ns();
  }
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long seq = sc->xseq;
  const double lambda = prm.use_override ? prm.lambda_override : sc->lambda;
  PcgOut out;
  pcg_solve<8>(g, tid, gridDim.x * blockDim.x, gridDim.x, part, bar, seq, prm, lambda, sm, &s_last, ph_ns, &t_ph, out);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) sc->pcg_phase_ns[i] += ph_ns[i];
    sc->xseq = seq;
    sc->rz0 = out.gam0;
    sc->rz = out.gam;
    sc->pcg_iters = out.iters;
    sc->pcg_flag = out.flag;
    sc->pcg_rel = out.gam0 > 0.0 ? sqrt(fabs(out.gam) / out.gam0) : 0.0;
  }
}

// The same solve for a graph small enough for ONE thread-block cluster (launched with a cluster dimension equal to the
// grid, <= 16 CTAs, one GPU): see grid_xreduce. bar must be non-NULL only to tell pcg_solve that the grid has more than
// one CTA; it is never dereferenced on this path.
__global__ void __launch_bounds__(kThreads, 1) k_pcg_cluster(DevGraph g, DevScalars* sc, double* part,
                                                            unsigned long long* bar, PcgParams prm) {
  __shared__ double sm[32];
  __shared__ double cl_part[4];
  __shared__ int s_last;
  __shared__ unsigned long long ph_ns[4], t_ph;
  if (threadIdx.x == 0) {
    ph_ns[0] = ph_ns[1] = ph_ns[2] = ph_ns[3] = 0;
    t_ph = globaltimer_ns();
  }
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long seq = sc->xseq;
  const double lambda = prm.use_override ? prm.lambda_override : sc->lambda;
  PcgOut out;
  pcg_solve<8>(g, tid, gridDim.x * blockDim.x, gridDim.x, part, bar, seq, prm, lambda, sm, &s_last, ph_ns, &t_ph, out, cl_part);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) sc->pcg_phase_ns[i] += ph_ns[i];
    sc->xseq = seq;
    sc->rz0 = out.gam0;
    sc->rz = out.gam;
    sc->pcg_iters = out.iters;
    sc->pcg_flag = out.flag;
    sc->pcg_rel = out.gam0 > 0.0 ? sqrt(fabs(out.gam) / out.gam0) : 0.0;
  }
  cooperative_groups::this_cluster().sync();  // no CTA may exit while another one can still read its shared memory
}

// ------------------------------------------------------------------------------------------------ update
__global__ void __launch_bounds__(kThreads) k_backsub(DevGraph g) {
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sl < g.Hlp.nslices; sl += nwarps) lm_slice_pass(g, sl, 1);
}
// SparseOptimizer::update into estimate buffer `dst` of every rank (the LM trial buffer, or the current one for GN)
// + computeScale partials
__global__ void __launch_bounds__(kThreads) k_update(DevGraph g, DevScalars* sc, int dst, double* part, double lambda_override,
                                                    int use_override) {
  __shared__ double sm[32];
  double lambda = use_override ? lambda_override : sc->lambda;
  double s = 0.0;
  int n = g.nP + g.nL_owned;  // ghost landmark rows are updated by their owners
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (v < g.nP) s += update_pose_row(g, v, lambda, dst);
    else s += update_lm_row(g, v - g.nP, lambda, dst);
  }
  if (g.world > 1) __threadfence_system();
  s = block_sum(s, sm);
  if (threadIdx.x == 0) part[2 * kMaxBlocks + blockIdx.x] = s;
}

// Did the linear solve produce a usable step? flag 0 = converged; flag 2 = breakdown (not SPD / non-finite) = g2o's
// solve() == false. flag 1 = the iteration limit was hit: LinearSolverEigen either solves exactly or fails, so an
// arbitrarily inexact step must not pass as a success -- it counts only if the preconditioned residual had already
// dropped by six orders of magnitude (step error ~1e-6 relative), otherwise the trial is rejected (LM) / the iteration
// fails (GN, sgb_linear_solve).
__host__ __device__ __forceinline__ bool pcg_usable(int flag, double rel) { return flag == 0 || (flag == 1 && rel <= 1e-6); }

// OptimizationAlgorithmLevenberg::solve, the part after the trial's chi2 is known: gain ratio, lambda /
// nu update, accept / reject, Terminate conditions, with g2o's constants. c / cr = activeChi2 / activeRobustChi2 of the
// trial estimates, scale = computeScale sum, solve_ok = the linear solve did not break down. One thread.
__device__ __forceinline__ void lm_control_update(DevScalars* sc, double c, double cr, double scale, bool solve_ok,
                                                  int max_trials) {
  sc->chi2 = c;
  sc->chi2_robust = cr;
  double tempChi = solve_ok ? cr : DBL_MAX;
  double rho = sc->current_chi - tempChi;
  scale += 1e-3;
  rho /= scale;
  sc->scale = scale;
  sc->temp_chi = tempChi;
  double lambda = sc->lambda, ni = sc->ni;
  int accepted = 0;
  bool lambda_finite = true;
  if (rho > 0.0 && isfinite(tempChi)) {
    double a = 2.0 * rho - 1.0;
    double alpha = 1.0 - a * a * a;
    alpha = fmin(alpha, 2.0 / 3.0);
    double scaleFactor = fmax(1.0 / 3.0, alpha);
    lambda *= scaleFactor;
    ni = 2.0;
    sc->current_chi = tempChi;
    accepted = 1;
  } else {
    lambda *= ni;
    ni *= 2.0;
    lambda_finite = isfinite(lambda);
  }
  int trials = sc->trials + (lambda_finite ? 1 : 0);  // g2o breaks out before qmax++ when lambda overflows
  sc->lambda = lambda;
  sc->ni = ni;
  sc->rho = rho;
  sc->accepted = accepted;
  sc->trials = trials;
  int again = (rho < 0.0 && trials < max_trials && lambda_finite) ? 1 : 0;
  sc->again = again;
  if (!again) sc->result = (trials == max_trials || rho == 0.0 || !lambda_finite) ? 2 : 1;
  sc->setup_fail = 0;
}
__global__ void __launch_bounds__(kThreads) k_lm_control(DevGraph g, DevScalars* sc, const double* part_chi, int nb_chi,
                                                        const double* part_scale, int nb_scale, int max_trials) {
  __shared__ double sm[32];
  double v[3];
  v[0] = reduce_partials(part_chi, nb_chi, sm);
  v[1] = reduce_partials(part_chi + kMaxBlocks, nb_chi, sm);
  v[2] = reduce_partials(part_scale + 2 * kMaxBlocks, nb_scale, sm);
  double fail = (double)sc->setup_fail;
  unsigned long long seq = sc->xseq;
  __syncthreads();
  xreduce(g, ++seq, v, 3, 0, true);
  xreduce(g, ++seq, &fail, 1, 1, true);
  if (threadIdx.x == 0) {
    sc->xseq = seq;
    lm_control_update(sc, v[0], v[1], v[2], pcg_usable(sc->pcg_flag, sc->pcg_rel) && (fail == 0.0), max_trials);
  }
}
// Gauss-Newton: the solve succeeded iff no rank saw a breakdown
__global__ void k_gn_control(DevGraph g, DevScalars* sc) {
  double fail = (double)sc->setup_fail;
  unsigned long long seq = sc->xseq;
  __syncthreads();
  xreduce(g, ++seq, &fail, 1, 1, true);
  if (threadIdx.x == 0) {
    sc->xseq = seq;
    bool ok = pcg_usable(sc->pcg_flag, sc->pcg_rel) && (fail == 0.0);
    sc->result = ok ? 1 : -1;
    sc->setup_fail = 0;
  }
}

// ------------------------------------------------------------------------------------------------ batched graphs
// SparseOptimizer::optimize() for MANY small independent graphs (the sliding-window landmark graphs the reference
// re-optimises once per key-frame, drone.cpp:146-156): ONE launch, one thread block per graph, the complete LM / GN
// loop on the device -- linearise, damp, Schur set-up, PCG, back-substitution, update, chi2 of the trial, gain ratio
// and lambda control -- with block barriers where the single-graph path has kernel boundaries or grid barriers. The
// per-row bodies (sgb_rows.h), the PCG (pcg_solve) and the LM control (lm_control_update) are the very same code.
struct ResPlanFwd {  // = ResPlan of sgb_resident.cuh (declared here so that BatchItem can carry one)
  int valid, bt, ncta, cap_pp, cap_pl, cap_lp, nz, nt, bytes, cap_sl, cap_lr, rows_cta, cz_nc, cz_h;
};
struct BatchItem {
  DevGraph g;
  DevScalars* sc;  // the handle's scalar block (kept coherent for later sgb_chi2 / sgb_step calls)
  ResPlanFwd res;  // valid: the graph fits the shared memory of its CTA -> the resident solve (sgb_resident.cuh)
};
struct BatchParams {
  int algo, max_iters, max_trials;
  double tau, user_lambda;
  PcgParams pcg;   // maxit <= 0: per graph, max(100, 12 * free poses)
};
struct BatchResult {
  int iters_done, result, cur, trials, pcg_iters /* last iteration */, pcg_iters_total;
  double chi2, lambda, rho, chi2_before, pcg_rel;
};

// U = blocks in flight per thread in the pose-major pass: 8 when the batch has at most one graph per SM (registers are
// free, the latency chain of a row is what counts), 2 otherwise (more resident CTAs per SM)
// the CTA-resident solve of sgb_resident.cuh (defined there; this header is included first)
__device__ void pcg_resident_block(const DevGraph& g, const PcgParams& prm, double lambda, const ResPlanFwd& rp,
                                   unsigned char* res_smem, double* sm, int* s_last, double* czs, unsigned long long& seq, PcgOut& out);
// where the resident solve keeps the coarse inverse inside its shared memory (the factorisation's work array until then)
__device__ double* pcg_resident_ainv(const ResPlanFwd& rp, unsigned char* res_smem);

template <int U>
__global__ void __launch_bounds__(kThreads) k_lm_block(const BatchItem* items, BatchParams prm, BatchResult* results) {
  extern __shared__ __align__(16) unsigned char lm_block_smem[];
  __shared__ ResPlanFwd res;
  __shared__ DevGraph g;
  __shared__ DevScalars sc;
  __shared__ double sm[32];
  __shared__ int s_last, s_ok;
  __shared__ unsigned long long ph_ns[4], t_ph;
  __shared__ double czs[5 * kCzMaxDim];  // two-level preconditioner: [cqL | cqR | rc | yc | reciprocal pivots] (sgb_coarse.h)
  {  // this block's graph descriptor -> shared memory (word copy)
    const int* src = reinterpret_cast<const int*>(&items[blockIdx.x].g);
    int* dst = reinterpret_cast<int*>(&g);
    for (int i = threadIdx.x; i < (int)(sizeof(DevGraph) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
    const int* s2 = reinterpret_cast<const int*>(items[blockIdx.x].sc);
    int* d2 = reinterpret_cast<int*>(&sc);
    for (int i = threadIdx.x; i < (int)(sizeof(DevScalars) / sizeof(int)); i += blockDim.x) d2[i] = s2[i];
    if (threadIdx.x == 0) {
      ph_ns[0] = ph_ns[1] = ph_ns[2] = ph_ns[3] = 0;
      t_ph = 0;
      res = items[blockIdx.x].res;
    }
  }
  __syncthreads();
  const int tid = threadIdx.x, nth = blockDim.x;
  PcgParams pcg = prm.pcg;
  if (pcg.maxit <= 0) pcg.maxit = max(100, 12 * g.nP);
  unsigned long long seq = 0;
  int done = 0, result = 1, pcg_total = 0, pcg_all = 0, trials_last = 0;
  double chi_before = 0.0;
  for (int iter = 0; iter < prm.max_iters && result == 1; ++iter) {
    // ---- linearise + assemble at the current estimates
    LinAcc acc;
    for (int lp = tid; lp < g.nP; lp += nth) lin_pose_row(g, lp, acc);
    for (int ll = tid; ll < g.nL; ll += nth) lin_lm_row(g, ll, acc);
    const double c = block_sum(acc.chi, sm), cr = block_sum(acc.chi_r, sm), md = block_max(acc.maxd, sm);
    if (tid == 0) {
      sc.chi2 = c;
      sc.chi2_robust = sc.chi_lin = sc.current_chi = sc.temp_chi = cr;
      sc.max_diag = md;
      sc.trials = 0;
      sc.rho = 0.0;
      sc.again = 0;
      sc.setup_fail = 0;
      if (prm.algo == 0 && iter == 0) {  // OptimizationAlgorithmLevenberg::computeLambdaInit
        sc.lambda = prm.user_lambda > 0.0 ? prm.user_lambda : prm.tau * md;
        sc.ni = 2.0;
      }
    }
    __syncthreads();
    chi_before = cr;
    pcg_total = 0;
    while (true) {  // LM: trials of this iteration; GN: exactly one pass
      const double lambda = prm.algo == 0 ? sc.lambda : 0.0;
      const int dst = prm.algo == 0 ? (g.cur ^ 1) : g.cur;
      bool ok = true;
      for (int ll = tid; ll < g.nL; ll += nth) ok &= setup_lm_row(g, ll, lambda);
      __syncthreads();
      for (int lp = tid; lp < g.nP; lp += nth) ok &= setup_pose_row(g, lp, lambda);
      for (int ch = tid; ch < (g.nP + kChunk - 1) / kChunk; ch += nth) ok &= setup_chunk(g, ch, lambda);
      if (tid == 0) s_ok = 1;
      __syncthreads();
      if (!ok) atomicAnd(&s_ok, 0);
      PcgOut po;
      if (res.valid) {
        __syncthreads();  // Hll_inv, bt and the preconditioner rows written above are complete
        if (res.cz_nc > 0) {  // coarse matrix -> its inverse (global g.cz_A), worked on where the solve will stage it
          coarse_factor(g, lambda, pcg_resident_ainv(res, lm_block_smem), czs + 4 * kCzMaxDim, g.cz_A, g.cz_fail, tid, nth,
                        [] { __syncthreads(); });
          __syncthreads();
        }
        pcg_resident_block(g, pcg, lambda, res, lm_block_smem, sm, &s_last, czs, seq, po);
        __syncthreads();  // x_p is complete
      } else {
        pcg_solve<U>(g, tid, nth, 1u, nullptr, nullptr, seq, pcg, lambda, sm, &s_last, ph_ns, &t_ph, po);
      }
      pcg_total += po.iters;
      pcg_all += po.iters;
      // ---- back-substitution, update into the trial (LM) or current (GN) estimates, computeScale
      for (int sl = tid >> 5; sl < g.Hlp.nslices; sl += nth >> 5) lm_slice_pass(g, sl, 1);
      __syncthreads();
      double s = 0.0;
      for (int v = tid; v < g.nP + g.nL_owned; v += nth)
        s += v < g.nP ? update_pose_row(g, v, lambda, dst) : update_lm_row(g, v - g.nP, lambda, dst);
      const double scale = block_sum(s, sm);
      __syncthreads();
      const bool solve_ok = pcg_usable(po.flag, po.gam0 > 0.0 ? sqrt(fabs(po.gam) / po.gam0) : 0.0) && s_ok != 0;
      if (prm.algo != 0) {  // Gauss-Newton
        if (tid == 0) {
          sc.pcg_iters = po.iters;
          sc.pcg_flag = po.flag;
          sc.pcg_rel = po.gam0 > 0.0 ? sqrt(fabs(po.gam) / po.gam0) : 0.0;
          sc.result = solve_ok ? 1 : -1;
        }
        __syncthreads();
        break;
      }
      // ---- chi2 of the trial estimates + LM control
      const double* pose = g.pose_buf[dst][g.rank];
      const double* lm = g.lm_buf[dst][g.rank];
      double tc = 0.0, tcr = 0.0;
      for (int k = tid; k < g.n_pp_owned + g.n_pl_owned; k += nth) {
        if (k < g.n_pp_owned) {
          double a, b;
          pp_chi(g, k, pose, &a, &b);
          tc += a;
          tcr += b;
        } else {
          double a = pl_chi(g, k - g.n_pp_owned, pose, lm);
          tc += a;
          tcr += a;
        }
      }
      tc = block_sum(tc, sm);
      tcr = block_sum(tcr, sm);
      if (tid == 0) {
        sc.pcg_iters = po.iters;
        sc.pcg_flag = po.flag;
        sc.pcg_rel = po.gam0 > 0.0 ? sqrt(fabs(po.gam) / po.gam0) : 0.0;
        lm_control_update(&sc, tc, tcr, scale, solve_ok, prm.max_trials);
        if (sc.accepted) g.cur ^= 1;  // discardTop: the trial estimates become current
      }
      __syncthreads();
      if (!sc.again) break;
    }
    result = sc.result;
    trials_last = prm.algo == 0 ? sc.trials : 1;
    ++done;
  }
  if (tid == 0) {
    BatchResult r;
    r.iters_done = result == -1 ? 0 : done;
    r.result = result;
    r.cur = g.cur;
    r.trials = trials_last;
    r.pcg_iters = pcg_total;
    r.pcg_iters_total = pcg_all;
    r.chi2 = prm.algo == 0 ? sc.current_chi : sc.chi2_robust;
    r.lambda = prm.algo == 0 ? sc.lambda : 0.0;
    r.rho = prm.algo == 0 ? sc.rho : 0.0;
    r.chi2_before = chi_before;
    r.pcg_rel = sc.pcg_rel;
    results[blockIdx.x] = r;
    *items[blockIdx.x].sc = sc;
  }
}

}  // namespace sgb
