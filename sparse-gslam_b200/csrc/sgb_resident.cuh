// sgb_resident.cuh -- the PCG solve for a graph that fits ONE thread-block cluster, with everything on chip.
//
// A small graph (the growing landmark graph the reference re-optimises once per key-frame, drone.cpp:146-156; the
// intel-lab-sized graphs of BASELINE config 1) is latency-bound: with the matrices in L2 a PCG iteration is three
// barriers plus three chains of dependent L2 round trips (index -> gather -> multiply), ~10 us, and an optimize(15) of
// a 300-pose graph runs ~2000 such iterations. Here the cluster keeps, for the whole solve,
//   * every CTA's own rows of Hpp / Hpl (pose-major) and its own slices of Hlp (landmark-major) in shared memory
//     (column indices and block values: they are constant during a solve),
//   * a full copy of the two vectors the operator gathers from -- the preconditioned residual z and t = W Hlp^T z --
//     in the shared memory of EVERY CTA: the thread that produces an entry stores it into all copies through
//     distributed shared memory, so every gather is a local shared-memory read,
//   * the per-row state (x, r, d, s) in registers of the thread that owns the row, the row's preconditioner rows in the
//     shared memory of its CTA,
// and synchronises with the hardware cluster barrier. Nothing but the final x travels to global memory.
// One thread owns one pose row, one warp one slice of the landmark-major matrix (the host sizes the cluster for that).
// The arithmetic per row is that of sgb_rows.h (same expressions, same order of the blocks).
#pragma once
#include "sgb_coarse.h"
#include "sgb_kernels.cuh"

namespace sgb {

struct ResPlan {   // host-computed, the same for every CTA of the launch
  int valid;
  int bt;          // threads = pose rows per CTA (multiple of 32)
  int ncta;        // cluster size
  int cap_pp, cap_pl, cap_lp;  // SELL entries of the largest per-CTA share
  int nz, nt;      // doubles of the replicated vectors z (3 per pose) and t (2 per landmark), padded to even
  int bytes;       // dynamic shared memory per CTA
  int cap_sl, cap_lr;  // landmark-major slices / landmark rows of the largest per-CTA share
  int rows_cta;        // pose rows per CTA: bt (one thread per row) or bt / 4 (four lanes per row)
  int cz_nc, cz_h;     // two-level preconditioner (sgb_coarse.h, four-lane solve only): coarse dimension (0 = off), node spacing
};

// byte offsets inside the dynamic shared memory (doubles first, then floats, then ints: natural alignment)
struct ResOffsets {
  size_t vpp, vpl, vlp, z, t, w, rvec, cinv, ainv, cpp, cpl, clp, meta, total;
};
__host__ __device__ inline ResOffsets res_offsets(const ResPlan& p) {
  ResOffsets o;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 15) & ~(size_t)15; return r; };
  o.vpp = take((size_t)p.cap_pp * 9 * 8);
  o.vpl = take((size_t)p.cap_pl * 6 * 8);
  o.vlp = take((size_t)p.cap_lp * 6 * 8);
  o.z = take((size_t)p.nz * 8);
  o.t = take((size_t)p.nt * 8);
  o.w = take((size_t)p.cap_lr * 3 * 8);
  o.rvec = take((size_t)p.rows_cta * 3 * 8);
  o.cinv = take((size_t)p.rows_cta * 9 * 16);
  o.ainv = take((size_t)p.cz_nc * (size_t)cz_ld(p.cz_nc) * 8);   // the coarse inverse, odd leading dimension
  o.cpp = take((size_t)p.cap_pp * 4);
  o.cpl = take((size_t)p.cap_pl * 4);
  o.clp = take((size_t)p.cap_lp * 4);
  o.meta = take((size_t)p.cap_sl * 5 * 4);
  o.total = off;
  return o;
}

#if defined(__CUDACC__)
// ---- cluster barrier on shared-memory mbarriers. cooperative_groups' cluster.sync() compiles to MEMBAR.ALL.GPU + ERRBAR +
// UCGABAR_ARV / UCGABAR_WAIT + CCTL.IVALL in EVERY warp; in the resident solve of the key-frame stream (4 CTAs x 16 warps,
// three barriers per PCG iteration) 55 % of the stall samples sat on those instructions, ~1.2 us per barrier
// (profiles/r2_res4_stream_*). Here every CTA owns one mbarrier expecting one arrival per CTA of the cluster: after the CTA's
// own block barrier, thread t < ncta arrives (release, cluster scope) on the mbarrier of CTA t through its shared::cluster
// address, and every thread waits (acquire, cluster scope) on its own CTA's mbarrier -- one remote arrive (~215 cycles) and
// a hardware-slept try_wait instead of the grid-style barrier. The distributed-shared-memory stores made before the block
// barrier are ordered before the arrive by the barrier + release chain, and visible after the acquire.
__device__ __forceinline__ unsigned res_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void res_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(res_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void res_mbar_arrive_remote(unsigned long long* bar, unsigned cta) {
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(res_smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void res_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RES_WAIT_%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra RES_WAIT_%=;\n"
      "}\n" ::"r"(res_smem_u32(bar)), "r"(parity) : "memory");
}

// store into the shared memory of CTA `rk` of the cluster at the address that `local_ptr` has in this CTA (mapa + st.shared::cluster:
// no generic-address conversion and no read of the CTA-id special register per store, which cooperative_groups'
// map_shared_rank costs inside a loop -- 7 % of the stall samples of the first build sat on S2R SR_CgaCtaId)
__device__ __forceinline__ void res_st_dsmem(const double* local_ptr, unsigned rk, double v) {
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(res_smem_u32(local_ptr)), "r"(rk));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}

// z_i = sum_j Cinv_ij r_j with the row's nine float4 in shared memory (layout [q][row of the CTA])
__device__ __forceinline__ double precond_row_res(const float4* cinv_s, int bt, int lr, bool active, const double r[3], double z[3]) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned base = lane & ~(unsigned)(kChunk - 1);
  const unsigned mask = ((1u << kChunk) - 1u) << base;
  double rr[3 * kChunk];
#pragma unroll
  for (int j = 0; j < kChunk; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) rr[3 * j + c] = __shfl_sync(mask, r[c], (int)(base + j));
  z[0] = z[1] = z[2] = 0.0;
  if (!active) return 0.0;
  constexpr int N = 3 * kChunk;
  float ci[3 * N];
#pragma unroll
  for (int q = 0; q < (3 * N) / 4; ++q) {
    const float4 v = cinv_s[q * bt + lr];
    ci[4 * q] = v.x; ci[4 * q + 1] = v.y; ci[4 * q + 2] = v.z; ci[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int rrow = 0; rrow < 3; ++rrow) {
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c) acc += (double)ci[rrow * N + c] * rr[c];
    z[rrow] = acc;
  }
  const int a = (int)(lane & (kChunk - 1));
  return rr[3 * a] * z[0] + rr[3 * a + 1] * z[1] + rr[3 * a + 2] * z[2];
}

// cooperative copy global -> shared by the whole CTA
template <class T>
__device__ __forceinline__ void stage_copy(T* dst, const T* src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// The solve itself. CL = true: the calling grid is ONE cluster of rp.ncta CTAs (k_pcg_res); CL = false: a single CTA
// owns the whole graph (the batched kernel k_lm_block_res: one graph per CTA), the cluster barrier becomes a block
// barrier and "every CTA's copy" is the CTA's own. rp.bt = blockDim.x threads, rp.bytes of dynamic shared memory at
// `res_smem`. Same recurrences, scalars and exit flags as pcg_solve (sgb_kernels.cuh). x is written to g.x_p.
// CZ = true (single CTA only, node spacing <= 32 rows): the coarse term of the two-level preconditioner, as in
// pcg_resident_solve4 below; czs = [cqL | cqR | rc | yc], kCzMaxDim doubles each, shared memory of the caller.
template <bool CL, bool CZ>
__device__ __forceinline__ void pcg_resident_solve(const DevGraph& g, const PcgParams& prm, const double lambda, const ResPlan& rp,
                                                   unsigned char* res_smem, double* sm, double* cl_part, int* s_last, double* czs,
                                                   unsigned long long& seq, PcgOut& out) {
  static_assert(!(CL && CZ), "the coarse term of the one-lane solve is built for a single CTA");
  namespace cg = cooperative_groups;
  const ResOffsets of = res_offsets(rp);
  double* vpp = reinterpret_cast<double*>(res_smem + of.vpp);
  double* vpl = reinterpret_cast<double*>(res_smem + of.vpl);
  double* vlp = reinterpret_cast<double*>(res_smem + of.vlp);
  double* z_s = reinterpret_cast<double*>(res_smem + of.z);
  double* t_s = reinterpret_cast<double*>(res_smem + of.t);
  float4* cinv_s = reinterpret_cast<float4*>(res_smem + of.cinv);
  int32_t* cpp = reinterpret_cast<int32_t*>(res_smem + of.cpp);
  int32_t* cpl = reinterpret_cast<int32_t*>(res_smem + of.cpl);
  int32_t* clp = reinterpret_cast<int32_t*>(res_smem + of.clp);
  double* w_s = reinterpret_cast<double*>(res_smem + of.w);       // [3][landmark rows of this CTA's slices]: (Hll + lambda I)^-1
  int32_t* meta_s = reinterpret_cast<int32_t*>(res_smem + of.meta);  // per slice: first entry (relative), steps, shift, first row, row end

  const int bt = rp.bt, ncta = CL ? (int)gridDim.x : 1, cta = CL ? (int)blockIdx.x : 0;
  const int spc = bt >> 5;  // 32-row slices per CTA
  const bool has_pl = g.Hpl.rows > 0;
  // ---- this CTA's share of the matrices -> shared memory
  const int ns_p = g.Hpp.nslices;
  const int sp0 = min(ns_p, cta * spc), sp1 = min(ns_p, (cta + 1) * spc);
  const int e0pp = g.Hpp.sbase[sp0], npp = g.Hpp.sbase[sp1] - e0pp;
  const int e0pl = has_pl ? g.Hpl.sbase[sp0] : 0, npl = has_pl ? g.Hpl.sbase[sp1] - e0pl : 0;
  const int ns_l = g.Hlp.nslices;
  // a cluster CTA owns as many slices as it has warps; a single CTA owns them all (its warps take them in turns)
  const int sl0 = CL ? min(ns_l, cta * spc) : 0, sl1 = CL ? min(ns_l, (cta + 1) * spc) : ns_l;
  const int e0lp = ns_l > 0 ? g.Hlp.sbase[sl0] : 0, nlp = ns_l > 0 ? g.Hlp.sbase[sl1] - e0lp : 0;
  stage_copy(cpp, g.Hpp.col + e0pp, npp);
  stage_copy(vpp, g.Hpp.vals + (size_t)e0pp * 9, npp * 9);
  if (has_pl) {
    stage_copy(cpl, g.Hpl.col + e0pl, npl);
    stage_copy(vpl, g.Hpl.vals + (size_t)e0pl * 6, npl * 6);
  }
  if (nlp > 0) {
    stage_copy(clp, g.Hlp.col + e0lp, nlp);
    stage_copy(vlp, g.Hlp.vals + (size_t)e0lp * 6, nlp * 6);
  }
  const int row0 = cta * bt;
  const int lr = (int)threadIdx.x, lp = row0 + lr;
  const bool act = lp < g.nP;
  if (act) {
    const float4* cg4 = reinterpret_cast<const float4*>(g.Cinv);
#pragma unroll
    for (int q = 0; q < 9; ++q) cinv_s[q * bt + lr] = cg4[(size_t)q * g.nP + lp];
  }
  // ---- coarse space (single CTA: row0 = 0, a segment of h <= 32 rows is h consecutive lanes of one warp)
  double* ainv_s = reinterpret_cast<double*>(res_smem + of.ainv);
  double* cqL = czs;
  double* cqR = czs + kCzMaxDim;
  double* rc_s = czs + 2 * kCzMaxDim;
  double* yc_s = czs + 3 * kCzMaxDim;
  const int cz_nc = CZ ? rp.cz_nc : 0, cz_h = CZ ? rp.cz_h : 8, cz_ldc = cz_ld(cz_nc), cz_nn = cz_nc / 3;
  const double cz_r = (CZ && act) ? cz_wr(lp, cz_h) : 0.0, cz_l = (CZ && act) ? 1.0 - cz_r : 0.0;
  const int cz_n0 = lp / cz_h;
  if (CZ) {
    for (int i = (int)threadIdx.x; i < cz_nc * cz_nc; i += bt) ainv_s[(i / cz_nc) * cz_ldc + (i % cz_nc)] = g.cz_A[i];
    for (int i = (int)threadIdx.x; i < 4 * kCzMaxDim; i += bt) czs[i] = 0.0;
  }
  // hat-weighted sums of a row vector over the segments: shuffle trees over the h lanes of a segment, its first lane stores
  auto cz_restrict = [&](const double* v) {
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      double a = cz_l * v[cc], b = cz_r * v[cc];
      for (int off = 1; off < cz_h; off <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
      }
      if ((lr & (cz_h - 1)) == 0 && cz_n0 < cz_nn - 1) {
        cqL[3 * cz_n0 + cc] = a;
        cqR[3 * (cz_n0 + 1) + cc] = b;
      }
    }
    __syncthreads();
  };
  auto cz_update = [&](double alpha, bool init) {
    for (int t = (int)threadIdx.x; t < cz_nc; t += bt) {
      const double q = cqR[t] + cqL[t];
      rc_s[t] = init ? q : rc_s[t] - alpha * q;
    }
    __syncthreads();
    for (int t = (int)threadIdx.x; t < cz_nc; t += bt) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int j = 0; j < cz_nc; j += 3) {
        a0 += ainv_s[j * cz_ldc + t] * rc_s[j];
        a1 += ainv_s[(j + 1) * cz_ldc + t] * rc_s[j + 1];
        a2 += ainv_s[(j + 2) * cz_ldc + t] * rc_s[j + 2];
      }
      yc_s[t] = (a0 + a1) + a2;
    }
    __syncthreads();
  };
  auto cz_add = [&](const double* rr, double* zz) {  // z += (R^T yc) of this row, returns r . (R^T yc)
    double dot = 0.0;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const double zc = cz_l * yc_s[3 * cz_n0 + cc] + cz_r * yc_s[3 * (cz_n0 + 1) + cc];
      zz[cc] += zc;
      dot += rr[cc] * zc;
    }
    return dot;
  };
  // ---- per-thread constants of the solve
  const int slice = lp >> 5, lane = lp & 31;
  int wpp = 0, bpp = 0, wpl = 0, bpl = 0;
  if (act) {
    wpp = sell_width(g.Hpp, slice);
    bpp = g.Hpp.sbase[slice] - e0pp + lane;
    if (has_pl) {
      wpl = sell_width(g.Hpl, slice);
      bpl = g.Hpl.sbase[slice] - e0pl + lane;
    }
  }
  // descriptors of this CTA's landmark-major slices and the W of their rows -> shared memory
  const int lrow0 = sl1 > sl0 ? g.Hlp.srow[sl0] : 0, nlrows = sl1 > sl0 ? g.Hlp.srow[sl1] - lrow0 : 0;
  for (int q = (int)threadIdx.x; q < sl1 - sl0; q += (int)blockDim.x) {
    const int sl = sl0 + q;
    meta_s[5 * q] = g.Hlp.sbase[sl] - e0lp;
    meta_s[5 * q + 1] = (g.Hlp.sbase[sl + 1] - g.Hlp.sbase[sl]) >> 5;
    meta_s[5 * q + 2] = g.Hlp.sshift[sl];
    meta_s[5 * q + 3] = g.Hlp.srow[sl];
    meta_s[5 * q + 4] = g.Hlp.srow[sl + 1];
  }
  for (int q = (int)threadIdx.x; q < nlrows; q += (int)blockDim.x) {
    const double* W = g.Hll_inv[g.rank];
    w_s[q] = W[lrow0 + q];
    w_s[nlrows + q] = W[(size_t)g.capL + lrow0 + q];
    w_s[2 * nlrows + q] = W[2 * (size_t)g.capL + lrow0 + q];
  }
  unsigned long long epoch = 0;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm);  // non-NULL marker only: never dereferenced on the cluster path
  // barrier (+ all-reduce of one block-wide sum, already known to every thread of the CTA). Cluster: every CTA WRITES its
  // partial into the slot it owns in every CTA's shared memory (DSMEM stores are fire-and-forget), one hardware cluster
  // barrier, then every thread adds the ncta local slots in CTA order -- no remote read (215 cycles) and no block barrier
  // after the cluster barrier. cl_part = [2][16] doubles, double-buffered by the parity of the reduction count: a slot is
  // rewritten two reductions later, i.e. after another cluster barrier that every reader has passed.
  auto sync_sum = [&](double* v, int nv) {
    if (CL) {
      cg::cluster_group cl = cg::this_cluster();
      ++seq;
      double* mine = cl_part + (seq & 1ull) * 16;
      if (nv > 0 && (int)threadIdx.x < ncta) cl.map_shared_rank(mine, (int)threadIdx.x)[cta] = v[0];
      cl.sync();
      if (nv > 0) {
        double acc = 0.0;
        for (int o = 0; o < ncta; ++o) acc += mine[o];
        v[0] = acc;
      }
    } else {
      __syncthreads();
    }
  };
  (void)bar; (void)epoch; (void)s_last;
  // store into every CTA's copy of a replicated vector
  auto put_z = [&](const double* zz) {
    if (CL) {
      cg::cluster_group cl = cg::this_cluster();
      for (int rk = 0; rk < ncta; ++rk) {
        double* zr = cl.map_shared_rank(z_s, rk) + 3 * lp;
        zr[0] = zz[0]; zr[1] = zz[1]; zr[2] = zz[2];
      }
    } else {
      z_s[3 * lp] = zz[0]; z_s[3 * lp + 1] = zz[1]; z_s[3 * lp + 2] = zz[2];
    }
  };

  // ---- x = 0, r = bt, z = M^-1 r, d = s = 0
  double x[3] = {0.0, 0.0, 0.0}, r[3] = {0.0, 0.0, 0.0}, d[3] = {0.0, 0.0, 0.0}, s[3] = {0.0, 0.0, 0.0}, z[3];
  if (act)
    for (int c = 0; c < 3; ++c) r[c] = g.bt[3 * (size_t)lp + c];
  __syncthreads();  // the staged arrays and cinv_s of the chunk-mates are complete
  double acc = precond_row_res(cinv_s, bt, lr, act, r, z);
  if (CL) cg::this_cluster().sync();  // nobody writes into another CTA's shared memory before that CTA has started
  if (CZ) {
    cz_restrict(r);
    cz_update(0.0, true);
    if (act) acc += cz_add(r, z);
  }
  if (act) put_z(z);
  double gam = block_sum(acc, sm);
  sync_sum(&gam, 1);  // also: every copy of z is complete
  const double gam0 = gam, target = prm.tol * prm.tol * gam0;
  double gam_old = 0.0, alpha_old = 0.0;
  int it = 0, flag = 1;
  if (!(gam0 > 0.0)) {
    flag = (gam0 == 0.0) ? 0 : 2;
  } else {
    while (true) {
      if (!(gam == gam)) { flag = 2; break; }
      if (gam <= target) { flag = 0; break; }
      if (it >= prm.maxit) break;
      const double beta = it == 0 ? 0.0 : gam / gam_old;
      // ---- phase A: t = W Hlp^T z, this warp's slice
      if (ns_l > 0) {
        for (int q = lr >> 5; q < sl1 - sl0; q += spc) {  // one slice per warp in a cluster, all slices in turns in a single CTA
          const int lm_e = meta_s[5 * q] + (lr & 31), lm_steps = meta_s[5 * q + 1], lm_sh = meta_s[5 * q + 2];
          const int lm_row = meta_s[5 * q + 3] + ((lr & 31) >> (5 - lm_sh));
          const bool lm_writer = ((lr & 31) & ((32 >> lm_sh) - 1)) == 0 && lm_row < meta_s[5 * q + 4];
          double u0 = 0.0, u1 = 0.0;
          for (int j = 0; j < lm_steps; ++j) {
            const int e = lm_e + 32 * j;
            const int col = clp[e];
            if (col >= 0) {
              const double* pv = z_s + 3 * (col & kLocalMask);
              const double* pa = vlp + (size_t)(e & ~31) * 6 + (e & 31);
              const double v0 = pv[0], v1 = pv[1], v2 = pv[2];
              u0 += pa[0] * v0 + pa[64] * v1 + pa[128] * v2;
              u1 += pa[32] * v0 + pa[96] * v1 + pa[160] * v2;
            }
          }
          lm_group_sum(lm_sh, u0, u1);
          if (lm_writer) {
            const double w11 = w_s[lm_row - lrow0], w12 = w_s[nlrows + lm_row - lrow0], w22 = w_s[2 * nlrows + lm_row - lrow0];
            const double t0 = w11 * u0 + w12 * u1, t1 = w12 * u0 + w22 * u1;
            if (CL) {
              cg::cluster_group cl = cg::this_cluster();
              for (int rk = 0; rk < ncta; ++rk) {
                double* tr = cl.map_shared_rank(t_s, rk) + 2 * lm_row;
                tr[0] = t0;
                tr[1] = t1;
              }
            } else {
              t_s[2 * lm_row] = t0;
              t_s[2 * lm_row + 1] = t1;
            }
          }
        }
        sync_sum(nullptr, 0);  // every copy of t is complete
      }
      // ---- phase B: w = (Hpp + lambda) z - Hpl t, delta = z.w, d = z + beta d, s = w + beta s
      acc = 0.0;
      if (act) {
        const double vi0 = z[0], vi1 = z[1], vi2 = z[2];  // this row's z is still in registers
        double q0 = lambda * vi0, q1 = lambda * vi1, q2 = lambda * vi2;
        for (int k = 0; k < wpp; ++k) {
          const int e = bpp + 32 * k;
          const int col = cpp[e];
          if (col < 0) continue;
          const double* pv = z_s + 3 * (col & kLocalMask);
          const double* pa = vpp + (size_t)(e & ~31) * 9 + (e & 31);
          const double v0 = pv[0], v1 = pv[1], v2 = pv[2];
          q0 += pa[0] * v0 + pa[32] * v1 + pa[64] * v2;
          q1 += pa[96] * v0 + pa[128] * v1 + pa[160] * v2;
          q2 += pa[192] * v0 + pa[224] * v1 + pa[256] * v2;
        }
        for (int k = 0; k < wpl; ++k) {
          const int e = bpl + 32 * k;
          const int col = cpl[e];
          if (col < 0) continue;
          const double* pt = t_s + 2 * (col & kLocalMask);
          const double* pa = vpl + (size_t)(e & ~31) * 6 + (e & 31);
          const double t0 = pt[0], t1 = pt[1];
          q0 -= pa[0] * t0 + pa[32] * t1;
          q1 -= pa[64] * t0 + pa[96] * t1;
          q2 -= pa[128] * t0 + pa[160] * t1;
        }
        d[0] = vi0 + beta * d[0]; d[1] = vi1 + beta * d[1]; d[2] = vi2 + beta * d[2];
        s[0] = q0 + beta * s[0]; s[1] = q1 + beta * s[1]; s[2] = q2 + beta * s[2];
        acc = vi0 * q0 + vi1 * q1 + vi2 * q2;
      }
      if (CZ) cz_restrict(s);
      double del = block_sum(acc, sm);
      sync_sum(&del, 1);
      const double denom = it == 0 ? del : del - beta * gam / alpha_old;
      if (!(denom > 0.0)) { flag = 2; break; }
      const double alpha = gam / denom;
      if (CZ) cz_update(alpha, false);
      // ---- phase C: x += alpha d, r -= alpha s, z = M^-1 r, gamma = r.z; the new z goes into every CTA's copy
      if (act)
        for (int c = 0; c < 3; ++c) {
          x[c] += alpha * d[c];
          r[c] -= alpha * s[c];
        }
      acc = precond_row_res(cinv_s, bt, lr, act, r, z);
      if (CZ && act) acc += cz_add(r, z);
      if (act) put_z(z);
      ++it;
      gam_old = gam;
      alpha_old = alpha;
      gam = block_sum(acc, sm);
      sync_sum(&gam, 1);  // every copy of z is complete
    }
  }
  if (act)
    for (int c = 0; c < 3; ++c) g.x_p[g.rank][3 * (size_t)lp + c] = x[c];
  out.gam0 = gam0;
  out.gam = gam;
  out.iters = it;
  out.flag = flag;
}

// ---------------------------------------------------------------------------------------------------------------------
// Four lanes per pose row: lane c < 3 of a row's group owns COMPONENT c of the row (of x, r, d, s, z and of every product),
// the fourth lane idles in the pose-major phases. One thread per row left the SM empty -- 7 warps for a 200-pose graph,
// ~670 dependent instructions per warp and iteration at an IPC of 0.1 (ncu, profiles/r2_res1_stream_*) -- and every
// product serial: here a pose-major row costs 3 instead of 9 multiply-adds per block and lane, the preconditioner 12
// instead of 36, and a 256-row CTA runs 32 warps. rp.bt threads per CTA own rp.bt / 4 rows and rp.bt / 32 landmark slices.
//
// CZ = true adds the coarse term of the two-level preconditioner (sgb_coarse.h): z = blockJacobi^-1 r + R^T Ainv (R r).
// The coarse residual rc = R r is kept, replicated, in every CTA and follows r through its own recurrence
// rc -= alpha R s: the hat-weighted sums of s over a segment of h rows are formed by the warps that own the rows
// (shuffles, then a fixed-order sum of the warp partials), stored into every CTA's copy through distributed shared
// memory and published by the barrier that already carries delta = z.w -- no barrier is added across the cluster; the
// dense nc x nc product Ainv rc is evaluated redundantly by every CTA from its own shared-memory copy of Ainv.
// czs = [cqL | cqR | rc | yc] (kCzMaxDim doubles each) + warp partials [32 * 6], static shared memory of the kernel.
template <bool CL, int U, bool CZ>
__device__ __forceinline__ void pcg_resident_solve4(const DevGraph& g, const PcgParams& prm, const double lambda, const ResPlan& rp,
                                                    unsigned char* res_smem, double* sm, double* cl_part, unsigned long long* cbar,
                                                    double* czs, unsigned long long& seq, PcgOut& out) {
  namespace cg = cooperative_groups;
  const ResOffsets of = res_offsets(rp);
  double* vpp = reinterpret_cast<double*>(res_smem + of.vpp);
  double* vpl = reinterpret_cast<double*>(res_smem + of.vpl);
  double* vlp = reinterpret_cast<double*>(res_smem + of.vlp);
  double* z_s = reinterpret_cast<double*>(res_smem + of.z);
  double* t_s = reinterpret_cast<double*>(res_smem + of.t);
  float4* cinv_s = reinterpret_cast<float4*>(res_smem + of.cinv);   // [9][rows of the CTA]
  int32_t* cpp = reinterpret_cast<int32_t*>(res_smem + of.cpp);
  int32_t* cpl = reinterpret_cast<int32_t*>(res_smem + of.cpl);
  int32_t* clp = reinterpret_cast<int32_t*>(res_smem + of.clp);
  double* w_s = reinterpret_cast<double*>(res_smem + of.w);
  int32_t* meta_s = reinterpret_cast<int32_t*>(res_smem + of.meta);
  double* r_s = reinterpret_cast<double*>(res_smem + of.rvec);      // [3 * rows of the CTA]: residual, read by the chunk-mates
  double* ainv_s = reinterpret_cast<double*>(res_smem + of.ainv);   // [nc][ld]: (R S R^T)^-1 of this trial
  double* cqL = czs;                     // [3 * node]: sum of (left weight x s) over the segment to the node's right
  double* cqR = czs + kCzMaxDim;         // [3 * node]: sum of (right weight x s) over the segment to the node's left
  double* rc_s = czs + 2 * kCzMaxDim;    // coarse residual R r
  double* yc_s = czs + 3 * kCzMaxDim;    // Ainv rc
  double* cw_s = czs + 4 * kCzMaxDim;    // [warp][side][component] partial sums of one restriction

  const int bt = rp.bt, rows_cta = bt >> 2, ncta = CL ? (int)gridDim.x : 1, cta = CL ? (int)blockIdx.x : 0;
  const int spc_p = rows_cta >> 5;  // pose slices per CTA
  const int spc_l = bt >> 5;        // landmark-major slices per CTA = warps
  const bool has_pl = g.Hpl.rows > 0;
  const int ns_p = g.Hpp.nslices;
  const int sp0 = min(ns_p, cta * spc_p), sp1 = min(ns_p, (cta + 1) * spc_p);
  const int e0pp = g.Hpp.sbase[sp0], npp = g.Hpp.sbase[sp1] - e0pp;
  const int e0pl = has_pl ? g.Hpl.sbase[sp0] : 0, npl = has_pl ? g.Hpl.sbase[sp1] - e0pl : 0;
  const int ns_l = g.Hlp.nslices;
  const int sl0 = CL ? min(ns_l, cta * spc_l) : 0, sl1 = CL ? min(ns_l, (cta + 1) * spc_l) : ns_l;
  const int e0lp = ns_l > 0 ? g.Hlp.sbase[sl0] : 0, nlp = ns_l > 0 ? g.Hlp.sbase[sl1] - e0lp : 0;
  stage_copy(cpp, g.Hpp.col + e0pp, npp);
  stage_copy(vpp, g.Hpp.vals + (size_t)e0pp * 9, npp * 9);
  if (has_pl) {
    stage_copy(cpl, g.Hpl.col + e0pl, npl);
    stage_copy(vpl, g.Hpl.vals + (size_t)e0pl * 6, npl * 6);
  }
  if (nlp > 0) {
    stage_copy(clp, g.Hlp.col + e0lp, nlp);
    stage_copy(vlp, g.Hlp.vals + (size_t)e0lp * 6, nlp * 6);
  }
  const int row0 = cta * rows_cta;
  const int tid = (int)threadIdx.x;
  const int lr = tid >> 2, c = tid & 3, lp = row0 + lr;
  const bool rowact = lp < g.nP, act = rowact && c < 3;
  {  // preconditioner rows of this CTA's poses: nine float4 per row, copied by the row's four lanes
    const float4* cg4 = reinterpret_cast<const float4*>(g.Cinv);
    if (rowact)
      for (int q = c; q < 9; q += 4) cinv_s[q * rows_cta + lr] = cg4[(size_t)q * g.nP + lp];
  }
  const int lrow0 = sl1 > sl0 ? g.Hlp.srow[sl0] : 0, nlrows = sl1 > sl0 ? g.Hlp.srow[sl1] - lrow0 : 0;
  for (int q = tid; q < sl1 - sl0; q += bt) {
    const int sl = sl0 + q;
    meta_s[5 * q] = g.Hlp.sbase[sl] - e0lp;
    meta_s[5 * q + 1] = (g.Hlp.sbase[sl + 1] - g.Hlp.sbase[sl]) >> 5;
    meta_s[5 * q + 2] = g.Hlp.sshift[sl];
    meta_s[5 * q + 3] = g.Hlp.srow[sl];
    meta_s[5 * q + 4] = g.Hlp.srow[sl + 1];
  }
  for (int q = tid; q < nlrows; q += bt) {
    const double* W = g.Hll_inv[g.rank];
    w_s[q] = W[lrow0 + q];
    w_s[nlrows + q] = W[(size_t)g.capL + lrow0 + q];
    w_s[2 * nlrows + q] = W[2 * (size_t)g.capL + lrow0 + q];
  }
  // ---- per-thread constants
  const int slice = lp >> 5, lane_p = lp & 31;
  int wpp = 0, bpp = 0, wpl = 0, bpl = 0;
  if (act) {
    wpp = sell_width(g.Hpp, slice);
    bpp = g.Hpp.sbase[slice] - e0pp + lane_p;
    if (has_pl) {
      wpl = sell_width(g.Hpl, slice);
      bpl = g.Hpl.sbase[slice] - e0pl + lane_p;
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
  // ---- coarse space: hat weights of this lane's row, the staged inverse, cleared exchange slots
  const int cz_nc = CZ ? rp.cz_nc : 0, cz_h = CZ ? rp.cz_h : 8, cz_ldc = cz_ld(cz_nc), cz_nn = cz_nc / 3;
  const double cz_r = (CZ && rowact) ? cz_wr(lp, cz_h) : 0.0, cz_l = (CZ && rowact) ? 1.0 - cz_r : 0.0;
  const int cz_n0 = lp / cz_h;
  if (CZ) {
    for (int i = tid; i < cz_nc * cz_nc; i += bt) ainv_s[(i / cz_nc) * cz_ldc + (i % cz_nc)] = g.cz_A[i];
    for (int i = tid; i < 4 * kCzMaxDim; i += bt) czs[i] = 0.0;
  }
  unsigned cpar = 0;  // phase parity of this CTA's mbarrier
  auto sync_sum = [&](double* v, int nv) {
    if (CL) {
      ++seq;
      double* mine = cl_part + (seq & 1ull) * 16;
      if (nv > 0 && tid < ncta) res_st_dsmem(mine + cta, (unsigned)tid, v[0]);
      __syncthreads();                                   // this CTA's distributed-shared-memory stores are issued
      if (tid < ncta) res_mbar_arrive_remote(cbar, tid);  // one arrival on every CTA's mbarrier
      res_mbar_wait(cbar, cpar);                          // all CTAs have arrived here
      cpar ^= 1u;
      if (nv > 0) {
        double acc = 0.0;
        for (int o = 0; o < ncta; ++o) acc += mine[o];
        v[0] = acc;
      }
    } else {
      __syncthreads();
    }
  };
  auto put_z = [&](double zc) {  // component c of row lp into every CTA's copy of z
    if (CL) {
      for (int rk = 0; rk < ncta; ++rk) res_st_dsmem(z_s + 3 * lp + c, (unsigned)rk, zc);
    } else {
      z_s[3 * lp + c] = zc;
    }
  };
  // z_c = sum_j Cinv[c][j] r_j over the 12 residual components of the row's chunk (read from r_s), returns r_c z_c
  auto precond = [&](double rc, double* zc) {
    if (act) r_s[3 * lr + c] = rc;
    __syncwarp();
    double acc = 0.0;
    if (act) {
      const double* rr = r_s + 3 * (lr & ~(kChunk - 1));
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 v = cinv_s[(3 * c + q) * rows_cta + lr];
        acc += (double)v.x * rr[4 * q] + (double)v.y * rr[4 * q + 1] + (double)v.z * rr[4 * q + 2] + (double)v.w * rr[4 * q + 3];
      }
    }
    __syncwarp();
    *zc = acc;
    return rc * acc;
  };

  // component c of row lp of a vector (0 on idle lanes) -> the hat-weighted sums over this CTA's segments, into every
  // CTA's cqL / cqR. A warp holds 8 consecutive rows, all inside one segment (h and the CTA's first row are multiples of 8).
  auto cz_restrict = [&](double v) {
    double a = cz_l * v, b = cz_r * v;
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if (lane < 3) {
      cw_s[warp * 6 + lane] = a;
      cw_s[warp * 6 + 3 + lane] = b;
    }
    __syncthreads();
    const int wps = cz_h >> 3, nseg = rows_cta / cz_h;
    if (tid < 6 * nseg) {
      const int sgm = tid / 6, side = (tid % 6) / 3, cc = tid % 3;
      double acc = 0.0;
      for (int w = sgm * wps; w < (sgm + 1) * wps; ++w) acc += cw_s[w * 6 + side * 3 + cc];
      const int gs = cta * nseg + sgm;  // the segment between nodes gs and gs + 1
      if (gs < cz_nn - 1) {
        double* dst = (side == 0 ? cqL + 3 * gs : cqR + 3 * (gs + 1)) + cc;
        if (CL) {
          for (int rk = 0; rk < ncta; ++rk) res_st_dsmem(dst, (unsigned)rk, acc);
        } else {
          *dst = acc;
        }
      }
    }
  };
  // rc = R r (init) or rc -= alpha R s, then yc = Ainv rc: every CTA for itself, thread t owns coarse component t
  auto cz_update = [&](double alpha, bool init) {
    for (int t = tid; t < cz_nc; t += bt) {
      const double q = cqR[t] + cqL[t];
      rc_s[t] = init ? q : rc_s[t] - alpha * q;
    }
    __syncthreads();
    for (int t = tid; t < cz_nc; t += bt) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;  // nc is a multiple of 3
      for (int j = 0; j < cz_nc; j += 3) {
        a0 += ainv_s[j * cz_ldc + t] * rc_s[j];
        a1 += ainv_s[(j + 1) * cz_ldc + t] * rc_s[j + 1];
        a2 += ainv_s[(j + 2) * cz_ldc + t] * rc_s[j + 2];
      }
      yc_s[t] = (a0 + a1) + a2;
    }
    __syncthreads();
  };
  // the coarse part of z for this lane's component
  auto cz_prolong = [&]() { return cz_l * yc_s[3 * cz_n0 + c] + cz_r * yc_s[3 * (cz_n0 + 1) + c]; };

  // ---- x = 0, r = bt, z = M^-1 r, d = s = 0 (one component per lane)
  double x = 0.0, r = 0.0, d = 0.0, s = 0.0, z = 0.0;
  if (act) r = g.bt[3 * (size_t)lp + c];
  __syncthreads();  // staged arrays complete
  // the rows of a chunk whose pose does not exist contribute zeros
  if (!rowact && c < 3 && lr < rows_cta) r_s[3 * lr + c] = 0.0;
  __syncthreads();
  double acc = precond(r, &z);
  if (CL) {
    if (tid == 0) res_mbar_init(cbar, (unsigned)ncta);
    cg::this_cluster().sync();  // every CTA has started and initialised its mbarrier before anybody stores into it
  }
  if (CZ) {
    cz_restrict(r);
    sync_sum(nullptr, 0);  // every copy of cqL / cqR is complete
    cz_update(0.0, true);
    if (act) {
      const double zc = cz_prolong();
      z += zc;
      acc += r * zc;
    }
  }
  if (act) put_z(z);
  double gam = block_sum(acc, sm);
  sync_sum(&gam, 1);
  const double gam0 = gam, target = prm.tol * prm.tol * gam0;
  double gam_old = 0.0, alpha_old = 0.0;
  int it = 0, flag = 1;
  if (!(gam0 > 0.0)) {
    flag = (gam0 == 0.0) ? 0 : 2;
  } else {
    while (true) {
      if (!(gam == gam)) { flag = 2; break; }
      if (gam <= target) { flag = 0; break; }
      if (it >= prm.maxit) break;
      const double beta = it == 0 ? 0.0 : gam / gam_old;
      // ---- phase A: t = W Hlp^T z; warp w takes the slices w, w + warps, ... of this CTA
      if (ns_l > 0) {
        for (int q = warp; q < sl1 - sl0; q += spc_l) {
          const int lm_e = meta_s[5 * q] + lane, lm_steps = meta_s[5 * q + 1], lm_sh = meta_s[5 * q + 2];
          const int lm_row = meta_s[5 * q + 3] + (lane >> (5 - lm_sh));
          const bool lm_writer = (lane & ((32 >> lm_sh) - 1)) == 0 && lm_row < meta_s[5 * q + 4];
          double u0 = 0.0, u1 = 0.0;
          for (int j = 0; j < lm_steps; ++j) {
            const int e = lm_e + 32 * j;
            const int col = clp[e];
            if (col >= 0) {
              const double* pv = z_s + 3 * (col & kLocalMask);
              const double* pa = vlp + (size_t)(e & ~31) * 6 + (e & 31);
              const double v0 = pv[0], v1 = pv[1], v2 = pv[2];
              u0 += pa[0] * v0 + pa[64] * v1 + pa[128] * v2;
              u1 += pa[32] * v0 + pa[96] * v1 + pa[160] * v2;
            }
          }
          lm_group_sum(lm_sh, u0, u1);
          if (lm_writer) {
            const double w11 = w_s[lm_row - lrow0], w12 = w_s[nlrows + lm_row - lrow0], w22 = w_s[2 * nlrows + lm_row - lrow0];
            const double t0 = w11 * u0 + w12 * u1, t1 = w12 * u0 + w22 * u1;
            if (CL) {
              for (int rk = 0; rk < ncta; ++rk) {
                res_st_dsmem(t_s + 2 * lm_row, (unsigned)rk, t0);
                res_st_dsmem(t_s + 2 * lm_row + 1, (unsigned)rk, t1);
              }
            } else {
              t_s[2 * lm_row] = t0;
              t_s[2 * lm_row + 1] = t1;
            }
          }
        }
        sync_sum(nullptr, 0);
      }
      // ---- phase B, component c of row lp: w_c = lambda z_c + sum_k Hpp_k[c][:] z_k - sum_k Hpl_k[c][:] t_k
      acc = 0.0;
      if (act) {
        double q = lambda * z;
        // U blocks per step: their column indices, then their gathers and values, are independent shared-memory loads
        // (a block is three dependent round trips -- index, operands, multiply-add -- and a row has up to a dozen blocks)
        for (int k0 = 0; k0 < wpp; k0 += U) {
          int col[U];
#pragma unroll
          for (int u = 0; u < U; ++u) col[u] = k0 + u < wpp ? cpp[bpp + 32 * (k0 + u)] : -1;
          double a[U][3], v[U][3];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            a[u][0] = a[u][1] = a[u][2] = v[u][0] = v[u][1] = v[u][2] = 0.0;
            if (col[u] >= 0) {
              const int e = bpp + 32 * (k0 + u);
              const double* pv = z_s + 3 * (col[u] & kLocalMask);
              const double* pa = vpp + (size_t)(e & ~31) * 9 + (e & 31) + 96 * c;   // row c of the 3x3 block
              v[u][0] = pv[0]; v[u][1] = pv[1]; v[u][2] = pv[2];
              a[u][0] = pa[0]; a[u][1] = pa[32]; a[u][2] = pa[64];
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) q += a[u][0] * v[u][0] + a[u][1] * v[u][1] + a[u][2] * v[u][2];
        }
        for (int k0 = 0; k0 < wpl; k0 += U) {
          int col[U];
#pragma unroll
          for (int u = 0; u < U; ++u) col[u] = k0 + u < wpl ? cpl[bpl + 32 * (k0 + u)] : -1;
          double a[U][2], tv[U][2];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            a[u][0] = a[u][1] = tv[u][0] = tv[u][1] = 0.0;
            if (col[u] >= 0) {
              const int e = bpl + 32 * (k0 + u);
              const double* pt = t_s + 2 * (col[u] & kLocalMask);
              const double* pa = vpl + (size_t)(e & ~31) * 6 + (e & 31) + 64 * c;   // row c of the 3x2 block
              tv[u][0] = pt[0]; tv[u][1] = pt[1];
              a[u][0] = pa[0]; a[u][1] = pa[32];
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) q -= a[u][0] * tv[u][0] + a[u][1] * tv[u][1];
        }
        d = z + beta * d;
        s = q + beta * s;
        acc = z * q;
      }
      if (CZ) cz_restrict(s);  // R s rides the barrier of delta
      double del = block_sum(acc, sm);
      sync_sum(&del, 1);
      const double denom = it == 0 ? del : del - beta * gam / alpha_old;
      if (!(denom > 0.0)) { flag = 2; break; }
      const double alpha = gam / denom;
      if (CZ) cz_update(alpha, false);
      // ---- phase C
      if (act) {
        x += alpha * d;
        r -= alpha * s;
      }
      acc = precond(r, &z);
      if (CZ && act) {
        const double zc = cz_prolong();
        z += zc;
        acc += r * zc;
      }
      if (act) put_z(z);
      ++it;
      gam_old = gam;
      alpha_old = alpha;
      gam = block_sum(acc, sm);
      sync_sum(&gam, 1);
    }
  }
  if (act) g.x_p[g.rank][3 * (size_t)lp + c] = x;
  out.gam0 = gam0;
  out.gam = gam;
  out.iters = it;
  out.flag = flag;
}

__device__ double* pcg_resident_ainv(const ResPlanFwd& rpf, unsigned char* res_smem) {
  ResPlan rp;
  rp.valid = rpf.valid; rp.bt = rpf.bt; rp.ncta = rpf.ncta; rp.cap_pp = rpf.cap_pp; rp.cap_pl = rpf.cap_pl; rp.cap_lp = rpf.cap_lp;
  rp.nz = rpf.nz; rp.nt = rpf.nt; rp.bytes = rpf.bytes; rp.cap_sl = rpf.cap_sl; rp.cap_lr = rpf.cap_lr; rp.rows_cta = rpf.rows_cta;
  rp.cz_nc = rpf.cz_nc; rp.cz_h = rpf.cz_h;
  return reinterpret_cast<double*>(res_smem + res_offsets(rp).ainv);
}
// the batched kernel's entry (k_lm_block, sgb_kernels.cuh): one CTA owns the whole graph
__device__ void pcg_resident_block(const DevGraph& g, const PcgParams& prm, double lambda, const ResPlanFwd& rpf,
                                   unsigned char* res_smem, double* sm, int* s_last, double* czs, unsigned long long& seq, PcgOut& out) {
  ResPlan rp;
  rp.valid = rpf.valid; rp.bt = rpf.bt; rp.ncta = rpf.ncta; rp.cap_pp = rpf.cap_pp; rp.cap_pl = rpf.cap_pl; rp.cap_lp = rpf.cap_lp;
  rp.nz = rpf.nz; rp.nt = rpf.nt; rp.bytes = rpf.bytes; rp.cap_sl = rpf.cap_sl; rp.cap_lr = rpf.cap_lr; rp.rows_cta = rpf.rows_cta;
  rp.cz_nc = rpf.cz_nc; rp.cz_h = rpf.cz_h;
  if (rp.cz_nc > 0) pcg_resident_solve<false, true>(g, prm, lambda, rp, res_smem, sm, nullptr, s_last, czs, seq, out);
  else pcg_resident_solve<false, false>(g, prm, lambda, rp, res_smem, sm, nullptr, s_last, czs, seq, out);
}

// One damped system of one graph, launched as ONE cluster of rp.ncta CTAs of rp.bt threads with rp.bytes of dynamic
// shared memory (single GPU).
__global__ void __launch_bounds__(kThreads, 1) k_pcg_res(DevGraph g, DevScalars* sc, PcgParams prm, ResPlan rp) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  __shared__ double sm[32];
  __shared__ double cl_part[32];
  __shared__ int s_last;
  unsigned long long seq = 0;  // counts the reductions of this launch (parity of the partial-sum slots)
  const double lambda = prm.use_override ? prm.lambda_override : sc->lambda;
  PcgOut out;
  pcg_resident_solve<true, false>(g, prm, lambda, rp, res_smem, sm, cl_part, &s_last, nullptr, seq, out);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->rz0 = out.gam0;
    sc->rz = out.gam;
    sc->pcg_iters = out.iters;
    sc->pcg_flag = out.flag;
    sc->pcg_rel = out.gam0 > 0.0 ? sqrt(fabs(out.gam) / out.gam0) : 0.0;
  }
  cooperative_groups::this_cluster().sync();  // no CTA may exit while another one can still write into (or read from) its shared memory
}
#endif

}  // namespace sgb

namespace sgb {
#if defined(__CUDACC__)
// Four lanes per row (pcg_resident_solve4): up to 1024 threads per CTA, a cluster of such CTAs or a single one.
// MAXT = largest CTA (sets the register budget: 64 at 1024 threads, 128 at 512), U = blocks per step, CZ = with the coarse
// term of the two-level preconditioner (a separate instantiation: the plain solve keeps its code and registers)
template <int MAXT, int U, bool CZ>
__global__ void __launch_bounds__(MAXT, 1) k_pcg_res4(DevGraph g, DevScalars* sc, PcgParams prm, ResPlan rp) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  __shared__ double sm[32];
  __shared__ double cl_part[32];
  __shared__ __align__(8) unsigned long long cbar;
  __shared__ double czs[CZ ? 4 * kCzMaxDim + 32 * 6 : 1];
  unsigned long long seq = 0;
  const double lambda = prm.use_override ? prm.lambda_override : sc->lambda;
  PcgOut out;
  if (gridDim.x > 1) pcg_resident_solve4<true, U, CZ>(g, prm, lambda, rp, res_smem, sm, cl_part, &cbar, czs, seq, out);
  else pcg_resident_solve4<false, U, CZ>(g, prm, lambda, rp, res_smem, sm, cl_part, &cbar, czs, seq, out);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->rz0 = out.gam0;
    sc->rz = out.gam;
    sc->pcg_iters = out.iters;
    sc->pcg_flag = out.flag;
    sc->pcg_rel = out.gam0 > 0.0 ? sqrt(fabs(out.gam) / out.gam0) : 0.0;
  }
  if (gridDim.x > 1) cooperative_groups::this_cluster().sync();
}
// Coarse matrix of the two-level preconditioner -> its inverse (sgb_coarse.h), once per LM trial after k_setup_lm: ONE CTA,
// the matrix in (3 nn) x ld doubles of dynamic shared memory followed by 3 nn reciprocal pivots.
__global__ void __launch_bounds__(1024, 1) k_setup_coarse(DevGraph g, DevScalars* sc, double lambda_override, int use_override) {
  extern __shared__ __align__(16) unsigned char cz_smem[];
  double* A = reinterpret_cast<double*>(cz_smem);
  const int nc = 3 * g.cz_nn;
  double* dinv = A + (size_t)nc * cz_ld(nc);
  const double lambda = use_override ? lambda_override : sc->lambda;
  coarse_factor(g, lambda, A, dinv, g.cz_A, g.cz_fail, (int)threadIdx.x, (int)blockDim.x, [] { __syncthreads(); });
}
// The same solve for a graph that fits ONE CTA (rows <= blockDim.x): block barriers only, no cluster.
__global__ void __launch_bounds__(kThreads, 1) k_pcg_res1(DevGraph g, DevScalars* sc, PcgParams prm, ResPlan rp) {
  extern __shared__ __align__(16) unsigned char res_smem[];
  __shared__ double sm[32];
  __shared__ int s_last;
  unsigned long long seq = 0;
  const double lambda = prm.use_override ? prm.lambda_override : sc->lambda;
  PcgOut out;
  pcg_resident_solve<false, false>(g, prm, lambda, rp, res_smem, sm, nullptr, &s_last, nullptr, seq, out);
  if (threadIdx.x == 0) {
    sc->rz0 = out.gam0;
    sc->rz = out.gam;
    sc->pcg_iters = out.iters;
    sc->pcg_flag = out.flag;
    sc->pcg_rel = out.gam0 > 0.0 ? sqrt(fabs(out.gam) / out.gam0) : 0.0;
  }
}
#endif
}  // namespace sgb
