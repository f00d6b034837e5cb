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
#pragma once
#include <cuda_runtime.h>
#include <float.h>

#include "sgb_rows.h"

namespace sgb {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 4096;  // partial-sum slots per reduction

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

// ------------------------------------------------------------------------------------------------ linearise
// part layout: [0]=chi [1]=chi_r [2]=maxd, each kMaxBlocks wide
__global__ void __launch_bounds__(kThreads) k_lin_pose(DevGraph g, double* part) {
  __shared__ double sm[32];
  LinAcc acc;
  for (int hp = blockIdx.x * blockDim.x + threadIdx.x; hp < g.Pf; hp += gridDim.x * blockDim.x) lin_pose_row(g, hp, acc);
  double c = block_sum(acc.chi, sm), cr = block_sum(acc.chi_r, sm), m = block_max(acc.maxd, sm);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = c;
    part[kMaxBlocks + blockIdx.x] = cr;
    part[2 * kMaxBlocks + blockIdx.x] = m;
  }
}
__global__ void __launch_bounds__(kThreads) k_lin_lm(DevGraph g, double* part) {
  __shared__ double sm[32];
  LinAcc acc;
  for (int hl = blockIdx.x * blockDim.x + threadIdx.x; hl < g.Lf; hl += gridDim.x * blockDim.x) lin_lm_row(g, hl, acc);
  double c = block_sum(acc.chi, sm), cr = block_sum(acc.chi_r, sm), m = block_max(acc.maxd, sm);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = c;
    part[kMaxBlocks + blockIdx.x] = cr;
    part[2 * kMaxBlocks + blockIdx.x] = m;
  }
}
// sums the partials of the two linearise kernels; initialises LM state at iteration 0
__global__ void __launch_bounds__(kThreads) k_finalize_lin(DevScalars* sc, const double* part_p, int nb_p,
                                                          const double* part_l, int nb_l, int init_lambda, double tau,
                                                          double user_lambda) {
  __shared__ double sm[32];
  double c = reduce_partials(part_p, nb_p, sm) + reduce_partials(part_l, nb_l, sm);
  double cr = reduce_partials(part_p + kMaxBlocks, nb_p, sm) + reduce_partials(part_l + kMaxBlocks, nb_l, sm);
  double m = 0.0;
  for (int i = threadIdx.x; i < nb_p; i += blockDim.x) m = fmax(m, part_p[2 * kMaxBlocks + i]);
  for (int i = threadIdx.x; i < nb_l; i += blockDim.x) m = fmax(m, part_l[2 * kMaxBlocks + i]);
  m = block_max(m, sm);
  if (threadIdx.x == 0) {
    sc->chi2 = c;
    sc->chi2_robust = cr;
    sc->chi_lin = cr;
    sc->max_diag = m;
    sc->current_chi = cr;
    sc->temp_chi = cr;
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
// one thread per edge, coalesced component-major SoA loads, warp-shuffle + block reduction
__global__ void __launch_bounds__(kThreads) k_chi2_edges(DevGraph g, const double* pose, const double* lm, double* part) {
  __shared__ double sm[32];
  double c = 0.0, cr = 0.0;
  int n = g.n_pp + g.n_pl;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    if (k < g.n_pp) {
      double a, b;
      pp_chi(g, k, pose, &a, &b);
      c += a;
      cr += b;
    } else {
      double a = pl_chi(g, k - g.n_pp, pose, lm);
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
__global__ void __launch_bounds__(kThreads) k_finalize_chi(DevScalars* sc, const double* part, int nb) {
  __shared__ double sm[32];
  double c = reduce_partials(part, nb, sm);
  double cr = reduce_partials(part + kMaxBlocks, nb, sm);
  if (threadIdx.x == 0) {
    sc->chi2 = c;
    sc->chi2_robust = cr;
  }
}

// ------------------------------------------------------------------------------------------------ trial set-up
__global__ void __launch_bounds__(kThreads) k_setup_lm(DevGraph g, DevScalars* sc, double lambda_override, int use_override) {
  double lambda = use_override ? lambda_override : sc->lambda;
  bool ok = true;
  for (int hl = blockIdx.x * blockDim.x + threadIdx.x; hl < g.Lf; hl += gridDim.x * blockDim.x) ok &= setup_lm_row(g, hl, lambda);
  if (!ok) atomicOr(&sc->setup_fail, 1);
}
__global__ void __launch_bounds__(kThreads) k_setup_pose(DevGraph g, DevScalars* sc, double lambda_override, int use_override) {
  double lambda = use_override ? lambda_override : sc->lambda;
  bool ok = true;
  for (int hp = blockIdx.x * blockDim.x + threadIdx.x; hp < g.Pf; hp += gridDim.x * blockDim.x) ok &= setup_pose_row(g, hp, lambda);
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

// Preconditioned conjugate gradient on S = (Hpp + lambda I) - Hpl (Hll + lambda I)^-1 Hpl^T, M = blockdiag(S).
// One launch runs the whole solve; every CTA evaluates the same scalars from the same partial sums in the same
// order, so control flow is uniform across the grid without any host round trip.
__global__ void __launch_bounds__(kThreads) k_pcg(DevGraph g, DevScalars* sc, double* part, unsigned long long* bar,
                                                 PcgParams prm) {
  __shared__ double sm[32];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  const unsigned int nb = gridDim.x;
  unsigned long long epoch = 0;
  const double lambda = prm.use_override ? prm.lambda_override : sc->lambda;
  double* part_rz = part;
  double* part_pq = part + kMaxBlocks;

  // x = 0, r = bt, z = Minv r, p = z
  double acc = 0.0;
  for (int hp = tid; hp < g.Pf; hp += nthreads) {
    double r[3] = {g.bt[3 * (size_t)hp], g.bt[3 * (size_t)hp + 1], g.bt[3 * (size_t)hp + 2]}, z[3];
    acc += precond_row(g, hp, r, z);
    for (int c = 0; c < 3; ++c) {
      g.x[3 * (size_t)hp + c] = 0.0;
      g.r[3 * (size_t)hp + c] = r[c];
      g.p[3 * (size_t)hp + c] = z[c];
    }
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) part_rz[blockIdx.x] = acc;
  grid_sync(bar, nb, epoch);
  double rz = reduce_partials(part_rz, nb, sm);
  const double rz0 = rz;
  int it = 0, flag = 0;
  if (!(rz0 > 0.0)) {
    flag = (rz0 == 0.0) ? 0 : 2;  // zero right-hand side: x = 0 is exact; negative / NaN: M not SPD
  } else {
    const double target = prm.tol * prm.tol * rz0;
    flag = 1;
    while (it < prm.maxit) {
      if (g.Lf > 0) {
        for (int row = tid; row < g.Lf; row += nthreads) schur_phaseA_row(g, row, g.p);
        grid_sync(bar, nb, epoch);
      }
      acc = 0.0;
      for (int hp = tid; hp < g.Pf; hp += nthreads) acc += schur_phaseB_row(g, hp, g.p, lambda, g.q);
      acc = block_sum(acc, sm);
      if (threadIdx.x == 0) part_pq[blockIdx.x] = acc;
      grid_sync(bar, nb, epoch);
      double pq = reduce_partials(part_pq, nb, sm);
      if (!(pq > 0.0)) {  // S not positive definite (or NaN): g2o's "Cholesky failure" analogue
        flag = 2;
        break;
      }
      double alpha = rz / pq;
      acc = 0.0;
      for (int hp = tid; hp < g.Pf; hp += nthreads) {
        double r[3], z[3];
        for (int c = 0; c < 3; ++c) {
          size_t o = 3 * (size_t)hp + c;
          g.x[o] += alpha * g.p[o];
          r[c] = g.r[o] - alpha * g.q[o];
          g.r[o] = r[c];
        }
        acc += precond_row(g, hp, r, z);
        for (int c = 0; c < 3; ++c) g.z[3 * (size_t)hp + c] = z[c];
      }
      acc = block_sum(acc, sm);
      if (threadIdx.x == 0) part_rz[blockIdx.x] = acc;
      grid_sync(bar, nb, epoch);
      double rzn = reduce_partials(part_rz, nb, sm);
      ++it;
      if (!(rzn == rzn)) {
        flag = 2;
        break;
      }
      if (rzn <= target) {
        rz = rzn;
        flag = 0;
        break;
      }
      double beta = rzn / rz;
      rz = rzn;
      for (int hp = tid; hp < g.Pf; hp += nthreads)
        for (int c = 0; c < 3; ++c) {
          size_t o = 3 * (size_t)hp + c;
          g.p[o] = g.z[o] + beta * g.p[o];
        }
      grid_sync(bar, nb, epoch);
    }
  }
  if (tid == 0) {
    sc->rz0 = rz0;
    sc->rz = rz;
    sc->pcg_iters = it;
    sc->pcg_flag = flag;
    sc->pcg_rel = rz0 > 0.0 ? sqrt(fabs(rz) / rz0) : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------ update
__global__ void __launch_bounds__(kThreads) k_backsub(DevGraph g) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < g.Lf; row += gridDim.x * blockDim.x) backsub_lm_row(g, row);
}
// SparseOptimizer::update into dst (trial buffers for LM, in place for GN) + computeScale partials
__global__ void __launch_bounds__(kThreads) k_update(DevGraph g, DevScalars* sc, const double* pose_src, double* pose_dst,
                                                    const double* lm_src, double* lm_dst, double* part, double lambda_override,
                                                    int use_override) {
  __shared__ double sm[32];
  double lambda = use_override ? lambda_override : sc->lambda;
  double s = 0.0;
  int n = g.Pf + g.Lf;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    if (v < g.Pf) s += update_pose_row(g, v, lambda, pose_src, pose_dst);
    else s += update_lm_row(g, v - g.Pf, lambda, lm_src, lm_dst);
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) part[2 * kMaxBlocks + blockIdx.x] = s;
}

// OptimizationAlgorithmLevenberg::solve, the part after the trial's chi2 is known
__global__ void __launch_bounds__(kThreads) k_lm_control(DevScalars* sc, const double* part_chi, int nb_chi,
                                                        const double* part_scale, int nb_scale, int max_trials) {
  __shared__ double sm[32];
  double c = reduce_partials(part_chi, nb_chi, sm);
  double cr = reduce_partials(part_chi + kMaxBlocks, nb_chi, sm);
  double scale = reduce_partials(part_scale + 2 * kMaxBlocks, nb_scale, sm);
  if (threadIdx.x == 0) {
    sc->chi2 = c;
    sc->chi2_robust = cr;
    bool ok2 = (sc->pcg_flag != 2) && (sc->setup_fail == 0);
    double tempChi = ok2 ? cr : DBL_MAX;
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
}

}  // namespace sgb
