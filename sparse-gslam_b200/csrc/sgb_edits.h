// sgb_edits.h -- per-element bodies of the kernels on either side of the optimiser:
//
//   N3  pose-graph edits on the device: relative-pose re-measurement and chained initial estimates when the
//       optimised landmark-graph poses are copied into the pose graph (reference
//       src/sparse_gslam/src/submap_loop_closer.cpp:206-223), and the chi2 test that removes false closures before
//       the final optimisation (src/sparse_gslam/src/log_runner.cpp:182-190);
//   N4  information-matrix producers: odometry covariance propagation (include/odom_error_propagator.h:6-46, inverted
//       at drone.cpp:128), scan-point covariances (src/multicloud2.cpp:56-83) and the least-squares line fit with its
//       covariance (src/ls_extractor/src/impl/smc.cpp:30-68, inverted at drone.cpp:203).
//
// Like sgb_rows.h these bodies are written once, called from the CUDA kernels (sgb_frontend.cu) and executed
// serially by the host-side test harness (tests/hostsim); the product never runs them on the CPU.
// SE2 follows the un-vendored g2o/types/slam2d/se2.h: operator*, inverse, normalize_theta.
#pragma once
#include "sgb_math.h"

namespace sgb {

struct Se2 {
  double x, y, th;
};
// g2o SE2::operator*=: t += R * t2; theta = normalize_theta(theta + theta2)
SGB_HD Se2 se2_mul(const Se2& a, const Se2& b) {
  double s, c;
  sgb_sincos(a.th, &s, &c);
  Se2 r;
  r.x = a.x + (c * b.x - s * b.y);
  r.y = a.y + (s * b.x + c * b.y);
  r.th = normalize_theta(a.th + b.th);
  return r;
}
// g2o SE2::inverse: R' = R^-1 (angle normalised), t' = R' * (-t)
SGB_HD Se2 se2_inv(const Se2& a) {
  Se2 r;
  r.th = normalize_theta(-a.th);
  double s, c;
  sgb_sincos(r.th, &s, &c);
  double mx = -a.x, my = -a.y;
  r.x = c * mx - s * my;
  r.y = s * mx + c * my;
  return r;
}
SGB_HD Se2 se2_load(const double* p) { return Se2{p[0], p[1], p[2]}; }
SGB_HD void se2_store(double* p, const Se2& a) {
  p[0] = a.x;
  p[1] = a.y;
  p[2] = a.th;
}

// submap_loop_closer.cpp:216: edge->setMeasurement((it - 1)->pose.estimate().inverse() * it->pose.estimate())
SGB_HD Se2 relative_measurement(const double* prev_est, const double* est) {
  return se2_mul(se2_inv(se2_load(prev_est)), se2_load(est));
}

// EdgeSE2::computeError + chi2() without the robust kernel (log_runner.cpp:183-184): e = z^-1 * (xi^-1 * xj)
SGB_HD double pp_edge_chi2(const double* xi, const double* xj, const double* z, const double* info6) {
  Se2 zi = se2_inv(se2_load(z));
  double zinv[3] = {zi.x, zi.y, zi.th};
  double si, ci, sz, cz, e[3];
  sgb_sincos(xi[2], &si, &ci);
  sgb_sincos(zinv[2], &sz, &cz);
  pp_error(xi, xj, zinv, ci, si, cz, sz, e);
  return sym3_quad(info6, e);
}

// ---------------------------------------------------------------------------------------------------------------
// N4 (a) OdomErrorPropagator<T>::step over one key-frame interval: starts from reset() (cov = 1e-6 I, pose = identity),
// applies `n` deltas (dx, dy, dtheta each), returns the accumulated pose (the odometry edge's measurement,
// drone.cpp:127) and the covariance (row-major 3x3).
//   J1 = [[1,0,dy ct - dx st],[0,1,-dx ct - dy st],[0,0,1]], J2 = [[ct,st,0],[-st,ct,0],[0,0,1]], ct/st of the pose angle
//   covu = diag(|dx dx| var_x, |dy dx| var_y, |dtheta dx| var_w);  cov = J1 cov J1^T + J2 covu J2^T;  pose *= delta
template <class T>
SGB_HD void mat3_abat(const T A[9], const T B[9], T out[9]) {  // (A * B) * A^T, evaluated left to right like Eigen
  T AB[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) AB[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out[3 * r + c] = AB[3 * r] * A[3 * c] + AB[3 * r + 1] * A[3 * c + 1] + AB[3 * r + 2] * A[3 * c + 2];
}
template <class T>
SGB_HD void odom_cov_step(T cov[9], const Se2& pose, double ddx, double ddy, double ddth, T var_x, T var_y, T var_w) {
  T dx = (T)ddx, dy = (T)ddy, th = (T)pose.th;
  T ct = (T)cos(th), st = (T)sin(th);  // the reference calls cos/sin on T (float overloads for T = float)
  T J1[9] = {(T)1, (T)0, dy * ct - dx * st, (T)0, (T)1, -dx * ct - dy * st, (T)0, (T)0, (T)1};
  T J2[9] = {ct, st, (T)0, -st, ct, (T)0, (T)0, (T)0, (T)1};
  // the diagonal is formed in double (Delta holds doubles) and stored into T
  T cu[9] = {(T)(fabs(ddx * ddx) * var_x), (T)0, (T)0, (T)0, (T)(fabs(ddy * ddx) * var_y), (T)0,
             (T)0, (T)0, (T)(fabs(ddth * ddx) * var_w)};
  T a[9], b[9];
  mat3_abat(J1, cov, a);
  mat3_abat(J2, cu, b);
  for (int i = 0; i < 9; ++i) cov[i] = a[i] + b[i];
}
template <class T>
SGB_HD void odom_propagate(const double* deltas /*[3*n]*/, int n, T var_x, T var_y, T var_w, Se2* pose_out, T cov[9]) {
  for (int i = 0; i < 9; ++i) cov[i] = (T)0;
  cov[0] = cov[4] = cov[8] = (T)1e-6;
  Se2 pose{0.0, 0.0, 0.0};
  for (int k = 0; k < n; ++k) {
    const double* d = deltas + 3 * (size_t)k;
    odom_cov_step<T>(cov, pose, d[0], d[1], d[2], var_x, var_y, var_w);
    pose = se2_mul(pose, Se2{d[0], d[1], d[2]});
  }
  *pose_out = pose;
}
// Eigen's fixed-size 3x3 inverse (cofactors / determinant); out = upper triangle 11,12,13,22,23,33 of cov^-1
SGB_HD void inv3_general_upper(const double m[9], double out6[6]) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c10 = m[5] * m[6] - m[3] * m[8];
  double c20 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c10 + m[2] * c20;
  double id = 1.0 / det;
  out6[0] = c00 * id;
  out6[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  out6[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  out6[3] = (m[0] * m[8] - m[2] * m[6]) * id;
  out6[4] = (m[2] * m[3] - m[0] * m[5]) * id;
  out6[5] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// ---------------------------------------------------------------------------------------------------------------
// N4 (b) scan-point covariance, single precision like the reference (multicloud2.cpp:62-83).
// Per scan of the window: cov_s = Juk cov Juk^T with Juk the Jacobian of the inverse of the accumulated odometry,
// pose_s = pose^-1, Jl = updateJacobian(pose_s.x, pose_s.y, pose_s.theta) (2x5, Jl(0,0) = Jl(1,1) = 1).
struct ScanFrame {
  float cov[9];  // 3x3 row-major, covariance of the scan's pose in the window frame
  float jl[10];  // 2x5 row-major
};
SGB_HD void scan_frame(const double* deltas /*[3*n]*/, int n, float var_x, float var_y, float var_w, ScanFrame* f) {
  Se2 pose;
  float cov[9];
  odom_propagate<float>(deltas, n, var_x, var_y, var_w, &pose, cov);
  float ct = cosf((float)pose.th), st = sinf((float)pose.th);
  // Eigen::Matrix3f << mixes float and double operands: each entry is evaluated in double and narrowed
  float juk[9] = {-ct, st, (float)(pose.y * ct + pose.x * st), -st, -ct, (float)(pose.y * st - pose.x * ct), 0.0f, 0.0f, -1.0f};
  mat3_abat<float>(juk, cov, f->cov);
  Se2 inv = se2_inv(pose);
  float dx = (float)inv.x, dy = (float)inv.y, th = (float)inv.th;
  float c2 = cosf(th), s2 = sinf(th);
  float* J = f->jl;
  for (int i = 0; i < 10; ++i) J[i] = 0.0f;
  J[0] = 1.0f;
  J[5 + 1] = 1.0f;
  J[2] = dy * c2 - dx * s2;
  J[3] = c2;
  J[4] = s2;
  J[5 + 2] = -dx * c2 - dy * s2;
  J[5 + 3] = -s2;
  J[5 + 4] = c2;
}
// one point of that scan: covp = var_r [[c^2, cs],[cs, s^2]] (c, s = beam direction), cov = Ja cov_s Ja^T + Jb covp Jb^T
// with Ja = Jl[:, 0:3], Jb = Jl[:, 3:5]; also rho/theta of the point. out_cov: 2x2 row-major.
SGB_HD void scan_point_cov(const ScanFrame& f, float beam_cos, float beam_sin, float var_r, float px, float py,
                           float out_cov[4], float out_rhotheta[2]) {
  const float c = beam_cos * beam_sin;
  float cp[4] = {beam_cos * beam_cos, c, c, beam_sin * beam_sin};
  for (int i = 0; i < 4; ++i) cp[i] *= var_r;
  const float* J = f.jl;
  float JaC[6];
  for (int r = 0; r < 2; ++r)
    for (int k = 0; k < 3; ++k) JaC[3 * r + k] = J[5 * r] * f.cov[k] + J[5 * r + 1] * f.cov[3 + k] + J[5 * r + 2] * f.cov[6 + k];
  float JbC[4];
  for (int r = 0; r < 2; ++r)
    for (int k = 0; k < 2; ++k) JbC[2 * r + k] = J[5 * r + 3] * cp[k] + J[5 * r + 4] * cp[2 + k];
  for (int r = 0; r < 2; ++r)
    for (int q = 0; q < 2; ++q) {
      float a = JaC[3 * r] * J[5 * q] + JaC[3 * r + 1] * J[5 * q + 1] + JaC[3 * r + 2] * J[5 * q + 2];
      float b = JbC[2 * r] * J[5 * q + 3] + JbC[2 * r + 1] * J[5 * q + 4];
      out_cov[2 * r + q] = a + b;
    }
  out_rhotheta[0] = sqrtf(px * px + py * py);
  out_rhotheta[1] = atan2f(py, px);
}

// ---------------------------------------------------------------------------------------------------------------
// N4 (c) LineSegment::leastSqFit (smc.cpp:30-68), single precision like the reference: rho/theta of the segment's
// points and their covariance from the per-point covariances; the pose-line edge information is the inverse of that
// covariance cast to double (drone.cpp:203). pts [2*n], pcov [4*n] (2x2 row-major per point).
SGB_HD void check_rho_theta_f(float rt[2]) {  // ls_extractor/utils.h:23-30
  const float pi = 3.14159265358979323846f;
  if (rt[0] < 0.0f) {
    rt[0] = -rt[0];
    rt[1] += pi;
    if (rt[1] > pi) rt[1] -= 2.0f * pi;
  }
}
struct F2 { float x, y; };
struct F4 { float a, b, c, d; };
// one 8-byte / 16-byte load per point / per covariance (pts is 8-byte, pcov 16-byte aligned: segment offsets are whole points)
SGB_HD F2 ld_pt(const float* p) {
#if defined(__CUDA_ARCH__)
  float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return F2{v.x, v.y};
#else
  return F2{p[0], p[1]};
#endif
}
SGB_HD F4 ld_cov(const float* p) {
#if defined(__CUDA_ARCH__)
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return F4{v.x, v.y, v.z, v.w};
#else
  return F4{p[0], p[1], p[2], p[3]};
#endif
}
SGB_HD void line_fit(const float* pts, const float* pcov, int n, float rhotheta[2], float cov[4]) {
  float sx = 0.0f, sy = 0.0f, sum_xy = 0.0f, sum_xx = 0.0f, sum_yy = 0.0f;
  for (int i = 0; i < n; ++i) {
    F2 pt = ld_pt(pts + 2 * (size_t)i);
    float x = pt.x, y = pt.y;
    sx += x;
    sy += y;
    sum_xy += x * y;
    sum_xx += x * x;
    sum_yy += y * y;
  }
  const float nf = (float)n;
  float mean_x = sx / nf, mean_y = sy / nf;
  sum_xx -= nf * (mean_x * mean_x);
  sum_yy -= nf * (mean_y * mean_y);
  sum_xy -= nf * (mean_x * mean_y);
  float d = sum_yy - sum_xx;
  rhotheta[1] = (float)(0.5 * (double)atan2f(-2 * sum_xy, d));  // float atan2, then the double literal 0.5 promotes
  float cos_t = cosf(rhotheta[1]), sin_t = sinf(rhotheta[1]);
  rhotheta[0] = mean_x * cos_t + mean_y * sin_t;
  check_rho_theta_f(rhotheta);
  cos_t = cosf(rhotheta[1]);
  sin_t = sinf(rhotheta[1]);
  float mean_x_sin = mean_x * sin_t, mean_y_cos = mean_y * cos_t;
  cov[0] = cov[1] = cov[2] = cov[3] = 0.0f;
  float inv_norm = (float)(1.0 / (double)(d * d + 4 * sum_xy * sum_xy));
  float cos_over_n = cos_t / nf, sin_over_n = sin_t / nf;
  for (int i = 0; i < n; ++i) {
    F2 pt = ld_pt(pts + 2 * (size_t)i);
    float dx = mean_x - pt.x, dy = mean_y - pt.y;
    float a10 = (dy * d + 2 * sum_xy * dx) * inv_norm;
    float a11 = (dx * d - 2 * sum_xy * dy) * inv_norm;
    float a00 = cos_over_n - mean_x_sin * a10 + mean_y_cos * a10;
    float a01 = sin_over_n - mean_x_sin * a11 + mean_y_cos * a11;
    const F4 C = ld_cov(pcov + 4 * (size_t)i);
    // cov += Ai * C * Ai^T, (Ai * C) first
    float m00 = a00 * C.a + a01 * C.c, m01 = a00 * C.b + a01 * C.d;
    float m10 = a10 * C.a + a11 * C.c, m11 = a10 * C.b + a11 * C.d;
    cov[0] += m00 * a00 + m01 * a01;
    cov[1] += m00 * a10 + m01 * a11;
    cov[2] += m10 * a00 + m11 * a01;
    cov[3] += m10 * a10 + m11 * a11;
  }
}
// bl_line.cov.cast<double>().inverse() -> (11, 12, 22); Eigen's 2x2 inverse: adjugate / determinant
SGB_HD void line_info(const float cov[4], double out3[3]) {
  double a = (double)cov[0], b = (double)cov[1], c = (double)cov[2], d = (double)cov[3];
  double id = 1.0 / (a * d - b * c);
  out3[0] = d * id;
  out3[1] = -b * id;
  out3[2] = a * id;
}

}  // namespace sgb
