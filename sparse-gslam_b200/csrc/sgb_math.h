// sgb_math.h -- per-edge arithmetic of the hot path, usable from device code and from the host-side
// test harness (tests/hostsim). All FP64 (g2o number_t = double).
//
// Restates (not copies) the behaviour of:
//   g2o/stuff/misc.h normalize_theta, g2o/types/slam2d/se2.h, edge_se2.cpp          [un-vendored g2o, SURVEY A.1-A.2]
//   reference src/sparse_gslam/src/g2o_bindings/edge_se2_rhotheta.cpp:9-16          (pose-line error)
//   reference src/ls_extractor/include/ls_extractor/utils.h:23-30,32-45             (checkRhoTheta, transform_line)
//   g2o BaseBinaryEdge::linearizeOplus numeric default, delta = 1e-9                [SURVEY A.3]
//   g2o RobustKernelDCS::robustify                                                  [SURVEY A.4]
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SGB_HD __host__ __device__ __forceinline__
#else
#define SGB_HD inline
#endif

namespace sgb {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * kPi;

SGB_HD void sgb_sincos(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(a, s, c);
#else
  *s = sin(a);
  *c = cos(a);
#endif
}

// floor-based wrap into [-pi, pi)
SGB_HD double normalize_theta(double t) {
  if (t >= -kPi && t < kPi) return t;
  double m = floor(t / kTwoPi);
  t = t - m * kTwoPi;
  if (t >= kPi) t -= kTwoPi;
  if (t < -kPi) t += kTwoPi;
  return t;
}

// RobustKernelDCS: rho0 = robustified chi2, rho1 = weight on Omega
SGB_HD void dcs_robustify(double phi, double e2, double* rho0, double* rho1) {
  double scale = (2.0 * phi) / (phi + e2);
  if (scale >= 1.0) {
    *rho0 = e2;
    *rho1 = 1.0;
  } else {
    *rho0 = scale * e2 * scale;
    *rho1 = scale * scale;
  }
}

// ------------------------------------------------------------------ EdgeSE2
// e = (zinv * (xi^-1 * xj)).toVector();  zinv = cached inverse measurement (x, y, theta)
SGB_HD void pp_error(const double xi[3], const double xj[3], const double zinv[3], double ci, double si, double cz,
                     double sz, double e[3]) {
  // xi^-1: theta' = normalize(-theta_i), t' = R(theta') * (-t_i); cos(theta') = ci, sin(theta') = -si
  double tix = -(ci * xi[0] + si * xi[1]);
  double tiy = (si * xi[0] - ci * xi[1]);
  double thi = normalize_theta(-xi[2]);
  // d = xi^-1 * xj
  double dx = tix + (ci * xj[0] + si * xj[1]);
  double dy = tiy + (-si * xj[0] + ci * xj[1]);
  double dth = normalize_theta(thi + xj[2]);
  // e = zinv * d
  e[0] = zinv[0] + (cz * dx - sz * dy);
  e[1] = zinv[1] + (sz * dx + cz * dy);
  e[2] = normalize_theta(zinv[2] + dth);
}

// analytic Jacobians of EdgeSE2 (row-major 3x3): A = d e / d xi, B = d e / d xj
SGB_HD void pp_jacobians(const double xi[3], const double xj[3], double ci, double si, double cz, double sz, double A[9],
                         double B[9]) {
  double dx = xj[0] - xi[0], dy = xj[1] - xi[1];
  double a00 = -ci, a01 = -si, a02 = -si * dx + ci * dy;
  double a10 = si, a11 = -ci, a12 = -ci * dx - si * dy;
  // Z = blockdiag(R(zinv.theta), 1)
  A[0] = cz * a00 - sz * a10; A[1] = cz * a01 - sz * a11; A[2] = cz * a02 - sz * a12;
  A[3] = sz * a00 + cz * a10; A[4] = sz * a01 + cz * a11; A[5] = sz * a02 + cz * a12;
  A[6] = 0.0; A[7] = 0.0; A[8] = -1.0;
  B[0] = cz * ci + sz * si; B[1] = cz * si - sz * ci; B[2] = 0.0;
  B[3] = sz * ci - cz * si; B[4] = sz * si + cz * ci; B[5] = 0.0;
  B[6] = 0.0; B[7] = 0.0; B[8] = 1.0;
}

// symmetric 3x3 information stored as upper triangle u = (11,12,13,22,23,33): y = Omega * x
SGB_HD void sym3_mul(const double u[6], const double x[3], double y[3]) {
  y[0] = u[0] * x[0] + u[1] * x[1] + u[2] * x[2];
  y[1] = u[1] * x[0] + u[3] * x[1] + u[4] * x[2];
  y[2] = u[2] * x[0] + u[4] * x[1] + u[5] * x[2];
}
SGB_HD double sym3_quad(const double u[6], const double x[3]) {
  double y[3];
  sym3_mul(u, x, y);
  return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}
// symmetric 2x2 as (11,12,22)
SGB_HD void sym2_mul(const double u[3], const double x[2], double y[2]) {
  y[0] = u[0] * x[0] + u[1] * x[1];
  y[1] = u[1] * x[0] + u[2] * x[1];
}
SGB_HD double sym2_quad(const double u[3], const double x[2]) {
  double y[2];
  sym2_mul(u, x, y);
  return x[0] * y[0] + x[1] * y[1];
}

// ------------------------------------------------------------------ EdgeSE2RhoTheta
// literal restatement of computeError (used by the g2o-numeric mode, where every rounding step matters)
SGB_HD void pl_error_literal(const double pose[3], const double line[2], const double z[2], double e[2]) {
  double th = normalize_theta(-pose[2]);
  double s, c;
  sgb_sincos(th, &s, &c);
  double tx = c * (-pose[0]) - s * (-pose[1]);
  double ty = s * (-pose[0]) + c * (-pose[1]);
  double rho = line[0], al = line[1];
  al += th;
  if (al > kPi) al -= kTwoPi;
  if (al < -kPi) al += kTwoPi;
  double ny, nx;
  sgb_sincos(al, &ny, &nx);
  rho += tx * nx + ty * ny;
  if (rho < 0.0) {
    rho = -rho;
    al += kPi;
    if (al > kPi) al -= kTwoPi;
  }
  e[0] = z[0] - rho;
  e[1] = normalize_theta(z[1] - al);
}

// closed form: same value up to rounding, one sincos; also returns the
// branch sign s and cos/sin of the world line angle for the analytic Jacobian
SGB_HD void pl_error_closed(const double pose[3], const double line[2], const double z[2], double e[2], double* sgn,
                            double* ca, double* sa) {
  double s, c;
  sgb_sincos(line[1], &s, &c);
  *ca = c;
  *sa = s;
  double q = line[0] - pose[0] * c - pose[1] * s;
  double al = line[1] + normalize_theta(-pose[2]);
  if (al > kPi) al -= kTwoPi;
  if (al < -kPi) al += kTwoPi;
  double rho = q;
  *sgn = 1.0;
  if (rho < 0.0) {
    rho = -rho;
    *sgn = -1.0;
    al += kPi;
    if (al > kPi) al -= kTwoPi;
  }
  e[0] = z[0] - rho;
  e[1] = normalize_theta(z[1] - al);
}

// A (2x3 row-major) = d e / d (tx, ty, theta), B (2x2 row-major) = d e / d (rho, alpha)
SGB_HD void pl_jac_analytic(const double pose[3], double sgn, double ca, double sa, double A[6], double B[4]) {
  A[0] = sgn * ca; A[1] = sgn * sa; A[2] = 0.0;
  A[3] = 0.0; A[4] = 0.0; A[5] = 1.0;
  B[0] = -sgn; B[1] = -sgn * (pose[0] * sa - pose[1] * ca);
  B[2] = 0.0; B[3] = -1.0;
}

// g2o central differences; VertexSE2::oplusImpl wraps theta, VertexRhoTheta::oplusImpl does not
// (reference vertex_rhotheta.cpp:30-34). want_A / want_B mirror "skip fixed vertices".
SGB_HD void pl_jac_numeric(const double pose[3], const double line[2], const double z[2], bool want_A, bool want_B,
                           double A[6], double B[4]) {
  const double delta = 1e-9;
  const double scalar = 1 / (2 * delta);
  if (want_A) {
    for (int d = 0; d < 3; ++d) {
      double pp[3] = {pose[0], pose[1], pose[2]};
      double pm[3] = {pose[0], pose[1], pose[2]};
      if (d < 2) {
        pp[d] += delta;
        pm[d] += -delta;
        // oplus also re-normalises theta: theta = normalize_theta(theta + 0)
        pp[2] = normalize_theta(pp[2] + 0.0);
        pm[2] = normalize_theta(pm[2] + 0.0);
      } else {
        pp[2] = normalize_theta(pp[2] + delta);
        pm[2] = normalize_theta(pm[2] + -delta);
      }
      double e1[2], e2[2];
      pl_error_literal(pp, line, z, e1);
      pl_error_literal(pm, line, z, e2);
      A[0 * 3 + d] = scalar * (e1[0] - e2[0]);
      A[1 * 3 + d] = scalar * (e1[1] - e2[1]);
    }
  }
  if (want_B) {
    for (int d = 0; d < 2; ++d) {
      double lp[2] = {line[0], line[1]};
      double lm[2] = {line[0], line[1]};
      lp[d] += delta;
      lm[d] += -delta;
      double e1[2], e2[2];
      pl_error_literal(pose, lp, z, e1);
      pl_error_literal(pose, lm, z, e2);
      B[0 * 2 + d] = scalar * (e1[0] - e2[0]);
      B[1 * 2 + d] = scalar * (e1[1] - e2[1]);
    }
  }
}

// ------------------------------------------------------------------ vertex updates (SparseOptimizer::update)
SGB_HD void pose_oplus(const double p[3], const double u[3], double out[3]) {
  out[0] = p[0] + u[0];
  out[1] = p[1] + u[1];
  out[2] = normalize_theta(p[2] + u[2]);
}
SGB_HD void lm_oplus(const double l[2], const double u[2], double out[2]) {
  out[0] = l[0] + u[0];
  out[1] = l[1] + u[1];  // no wrap: vertex_rhotheta.cpp:33 discards normalize_theta's return value
}

// ------------------------------------------------------------------ small dense helpers
// inverse of a symmetric 3x3 given row-major m[9]; returns false when not positive definite / non-finite
SGB_HD bool inv3_spd(const double m[9], double out[9]) {
  double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[8];
  double c00 = d * f - e * e;
  double c01 = c * e - b * f;
  double c02 = b * e - c * d;
  double det = a * c00 + b * c01 + c * c02;
  bool ok = (a > 0.0) && (a * d - b * b > 0.0) && (det > 0.0) && (det == det) && (det < 1e300);
  double id = 1.0 / det;
  out[0] = c00 * id;
  out[1] = c01 * id;
  out[2] = c02 * id;
  out[3] = out[1];
  out[4] = (a * f - c * c) * id;
  out[5] = (b * c - a * e) * id;
  out[6] = out[2];
  out[7] = out[5];
  out[8] = (a * d - b * b) * id;
  return ok;
}
// inverse of symmetric 2x2 (h11, h12, h22) -> (i11, i12, i22)
SGB_HD bool inv2_spd(double h11, double h12, double h22, double out[3]) {
  double det = h11 * h22 - h12 * h12;
  bool ok = (h11 > 0.0) && (det > 0.0) && (det == det) && (det < 1e300);
  double id = 1.0 / det;
  out[0] = h22 * id;
  out[1] = -h12 * id;
  out[2] = h11 * id;
  return ok;
}

}  // namespace sgb
