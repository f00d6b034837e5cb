// hostsim_edits.cpp -- TEST HARNESS ONLY (never shipped, never loaded by the sparse_gslam_b200 package).
// Executes the bodies of sparse-gslam_b200/csrc/sgb_edits.h on the host with the kernels' work decomposition
// (sgb_posegraph.cu, sgb_frontend.cu): one "thread" per element, and for the pose chain the same tiled
// reduce-then-scan structure (k_pg_remeasure / k_pg_scan_tiles / k_pg_apply) with virtual threads, so the CPU-only tier
// checks the arithmetic and the scan against the oracle before GPU time is spent.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../sparse-gslam_b200/csrc/sgb_edits.h"

using namespace sgb;

extern "C" {

// inclusive Hillis-Steele scan over n elements (one per virtual thread), as block_scan_se2 does in shared memory
static void hillis_steele(std::vector<Se2>& v) {
  const int n = (int)v.size();
  std::vector<Se2> w(n);
  for (int off = 1; off < n; off <<= 1) {
    for (int t = 0; t < n; ++t) w[t] = t >= off ? se2_mul(v[t - off], v[t]) : v[t];
    v.swap(w);
  }
}
// k_pg_remeasure / k_pg_scan_tiles / k_pg_apply with T threads per tile, `items` poses per thread, T2 scan threads
void hs_pg_append(const double* prev, const double* lm, int count, int T, int items, int T2, double* z_out, double* est_out) {
  const int tile = T * items, ntiles = (count + tile - 1) / tile;
  std::vector<Se2> tile_prod(ntiles), tile_prefix(ntiles);
  for (int b = 0; b < ntiles; ++b) {  // k_pg_remeasure
    std::vector<Se2> loc(T);
    for (int t = 0; t < T; ++t) {
      int k0 = std::min(count, b * tile + t * items), k1 = std::min(count, k0 + items);
      Se2 l{0.0, 0.0, 0.0};
      for (int k = k0; k < k1; ++k) {
        Se2 z = relative_measurement(lm + 3 * (size_t)k, lm + 3 * (size_t)(k + 1));
        se2_store(z_out + 3 * (size_t)k, z);
        l = se2_mul(l, z);
      }
      loc[t] = l;
    }
    hillis_steele(loc);
    tile_prod[b] = loc[T - 1];
  }
  {  // k_pg_scan_tiles
    const int chunk = (ntiles + T2 - 1) / T2;
    std::vector<Se2> loc(T2);
    for (int t = 0; t < T2; ++t) {
      int k0 = std::min(ntiles, t * chunk), k1 = std::min(ntiles, k0 + chunk);
      Se2 l{0.0, 0.0, 0.0};
      for (int k = k0; k < k1; ++k) l = se2_mul(l, tile_prod[k]);
      loc[t] = l;
    }
    hillis_steele(loc);
    for (int t = 0; t < T2; ++t) {
      int k0 = std::min(ntiles, t * chunk), k1 = std::min(ntiles, k0 + chunk);
      Se2 run = se2_load(prev);
      if (t > 0) run = se2_mul(run, loc[t - 1]);
      for (int k = k0; k < k1; ++k) {
        tile_prefix[k] = run;
        run = se2_mul(run, tile_prod[k]);
      }
    }
  }
  for (int b = 0; b < ntiles; ++b) {  // k_pg_apply
    std::vector<Se2> loc(T);
    for (int t = 0; t < T; ++t) {
      int k0 = std::min(count, b * tile + t * items), k1 = std::min(count, k0 + items);
      Se2 l{0.0, 0.0, 0.0};
      for (int k = k0; k < k1; ++k) l = se2_mul(l, se2_load(z_out + 3 * (size_t)k));
      loc[t] = l;
    }
    hillis_steele(loc);
    for (int t = 0; t < T; ++t) {
      int k0 = std::min(count, b * tile + t * items), k1 = std::min(count, k0 + items);
      Se2 run = tile_prefix[b];
      if (t > 0) run = se2_mul(run, loc[t - 1]);
      for (int k = k0; k < k1; ++k) {
        run = se2_mul(run, se2_load(z_out + 3 * (size_t)k));
        se2_store(est_out + 3 * (size_t)k, run);
      }
    }
  }
}

void hs_closure_chi2(const double* est, const int32_t* ei, const int32_t* ej, const double* z, const double* info, int n, double* chi_out) {
  for (int k = 0; k < n; ++k)
    chi_out[k] = pp_edge_chi2(est + 3 * (size_t)ei[k], est + 3 * (size_t)ej[k], z + 3 * (size_t)k, info + 6 * (size_t)k);
}

void hs_odom_information(const double* deltas, const int32_t* seg_ptr, int n_seg, double std_x, double std_y, double std_w,
                         double* z_out, double* cov_out, double* info_out) {
  for (int s = 0; s < n_seg; ++s) {
    Se2 pose;
    double cov[9];
    odom_propagate<double>(deltas + 3 * (size_t)seg_ptr[s], seg_ptr[s + 1] - seg_ptr[s], std_x * std_x, std_y * std_y, std_w * std_w, &pose, cov);
    inv3_general_upper(cov, info_out + 6 * (size_t)s);
    se2_store(z_out + 3 * (size_t)s, pose);
    if (cov_out)
      for (int i = 0; i < 9; ++i) cov_out[9 * (size_t)s + i] = cov[i];
  }
}

void hs_scan_point_covariances(const double* deltas, int n_windows, int n_scans, int scan_size, const float* beam, const float* pts,
                               float std_x, float std_y, float std_w, float var_r, float* cov_out, float* rt_out, uint8_t* valid_out) {
  for (int w = 0; w < n_windows; ++w)
    for (int i = 0; i < n_scans; ++i) {
      ScanFrame f;
      scan_frame(deltas + 3 * ((size_t)w * (n_scans - 1) + i), n_scans - 1 - i, std_x * std_x, std_y * std_y, std_w * std_w, &f);
      for (int j = 0; j < scan_size; ++j) {
        size_t p = ((size_t)w * n_scans + i) * scan_size + j;
        float x = pts[2 * p], y = pts[2 * p + 1];
        float c[4] = {0, 0, 0, 0}, rt[2] = {0, 0};
        bool ok = std::isfinite(x) && std::isfinite(y);
        if (ok) scan_point_cov(f, beam[2 * j], beam[2 * j + 1], var_r, x, y, c, rt);
        for (int q = 0; q < 4; ++q) cov_out[4 * p + q] = c[q];
        rt_out[2 * p] = rt[0];
        rt_out[2 * p + 1] = rt[1];
        valid_out[p] = ok ? 1 : 0;
      }
    }
}

void hs_line_fit_information(const float* pts, const float* pcov, const int32_t* seg_ptr, int n_seg, float* rt_out, float* cov_out,
                             double* info_out) {
  for (int s = 0; s < n_seg; ++s) {
    int a = seg_ptr[s], b = seg_ptr[s + 1];
    line_fit(pts + 2 * (size_t)a, pcov + 4 * (size_t)a, b - a, rt_out + 2 * (size_t)s, cov_out + 4 * (size_t)s);
    line_info(cov_out + 4 * (size_t)s, info_out + 3 * (size_t)s);
  }
}

}  // extern "C"
