// sgo_frontend.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for the rows either side of the optimiser
// (SURVEY.md 8f N3 / N4). Sequential restatement, expression by expression, of in-repo reference code:
//   N3  src/sparse_gslam/src/submap_loop_closer.cpp:206-223 (state copy lm graph -> pose graph),
//       src/sparse_gslam/src/log_runner.cpp:182-190 (closure chi2 test)
//   N4  src/sparse_gslam/include/odom_error_propagator.h:6-46 (+ drone.cpp:127-128), src/sparse_gslam/src/multicloud2.cpp:56-83,
//       src/ls_extractor/src/impl/smc.cpp:30-68 (+ drone.cpp:203), src/ls_extractor/include/ls_extractor/utils.h:23-30
// Unlike the optimiser itself this code IS in the reference tree, but it cannot be compiled from there: it needs
// Eigen, g2o's SE2, ROS and PCL headers, none of which exist in this container. The tiny fixed-size
// matrix type below stands in for Eigen (products evaluated left to right, coefficient sums in index order).
// PARITY UNPINNED for the same reason as sgo_oracle.cpp: the reference ships no tests or vectors for these functions.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "sgo_oracle.h"

namespace {

inline double normalize_theta(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  double multiplier = std::floor(theta / (2 * M_PI));
  theta = theta - multiplier * 2 * M_PI;
  if (theta >= M_PI) theta -= 2 * M_PI;
  if (theta < -M_PI) theta += 2 * M_PI;
  return theta;
}
struct SE2 {
  double t[2] = {0, 0};
  double a = 0;
  SE2() {}
  SE2(double x, double y, double th) { t[0] = x; t[1] = y; a = th; }
  double operator[](int i) const { return i < 2 ? t[i] : a; }
  SE2& operator*=(const SE2& o) {
    double c = std::cos(a), s = std::sin(a);
    double rx = c * o.t[0] - s * o.t[1], ry = s * o.t[0] + c * o.t[1];
    t[0] += rx;
    t[1] += ry;
    a += o.a;
    a = normalize_theta(a);
    return *this;
  }
  SE2 operator*(const SE2& o) const { SE2 r(*this); r *= o; return r; }
  SE2 inverse() const {
    SE2 r;
    r.a = normalize_theta(-a);
    double c = std::cos(r.a), s = std::sin(r.a);
    double mx = t[0] * -1., my = t[1] * -1.;
    r.t[0] = c * mx - s * my;
    r.t[1] = s * mx + c * my;
    return r;
  }
};

template <class T, int R, int C>
struct Mat {
  T v[R][C];
  Mat() { for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) v[i][j] = T(0); }
  T& operator()(int i, int j) { return v[i][j]; }
  T operator()(int i, int j) const { return v[i][j]; }
};
template <class T, int R, int K, int C>
Mat<T, R, C> mul(const Mat<T, R, K>& a, const Mat<T, K, C>& b) {
  Mat<T, R, C> o;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) {
      T s = a(i, 0) * b(0, j);
      for (int k = 1; k < K; ++k) s += a(i, k) * b(k, j);
      o(i, j) = s;
    }
  return o;
}
template <class T, int R, int C>
Mat<T, C, R> tr(const Mat<T, R, C>& a) {
  Mat<T, C, R> o;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) o(j, i) = a(i, j);
  return o;
}
template <class T, int R, int C>
Mat<T, R, C> add(const Mat<T, R, C>& a, const Mat<T, R, C>& b) {
  Mat<T, R, C> o;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) o(i, j) = a(i, j) + b(i, j);
  return o;
}
template <class T, int R, int C, int BR, int BC>
Mat<T, BR, BC> block(const Mat<T, R, C>& a, int r0, int c0) {
  Mat<T, BR, BC> o;
  for (int i = 0; i < BR; ++i)
    for (int j = 0; j < BC; ++j) o(i, j) = a(r0 + i, c0 + j);
  return o;
}

// odom_error_propagator.h:6-15
template <class T, int R, int C>
void updateJacobian(Mat<T, R, C>& J, T dx, T dy, T theta) {
  T ct = std::cos(theta), st = std::sin(theta);
  J(0, 2) = dy * ct - dx * st;
  J(0, 3) = ct;
  J(0, 4) = st;
  J(1, 2) = -dx * ct - dy * st;
  J(1, 3) = -st;
  J(1, 4) = ct;
}
// odom_error_propagator.h:17-52
template <class T>
struct OdomErrorPropagator {
  SE2 pose;
  T var_x, var_y, var_w;
  Mat<T, 3, 3> cov;
  Mat<T, 3, 6> J;
  OdomErrorPropagator(T sx, T sy, T sw) : var_x(sx * sx), var_y(sy * sy), var_w(sw * sw) {
    reset();
    J(0, 0) = 1; J(1, 1) = 1; J(2, 2) = 1; J(2, 5) = 1;
  }
  void reset() {
    cov = Mat<T, 3, 3>();
    cov(0, 0) = cov(1, 1) = cov(2, 2) = T(1e-6);
    pose = SE2();
  }
  void step(const SE2& dpose) {
    updateJacobian(J, (T)dpose[0], (T)dpose[1], (T)pose[2]);
    Mat<T, 3, 3> covu;
    covu(0, 0) = (T)(std::abs(dpose[0] * dpose[0]) * var_x);
    covu(1, 1) = (T)(std::abs(dpose[1] * dpose[0]) * var_y);
    covu(2, 2) = (T)(std::abs(dpose[2] * dpose[0]) * var_w);
    Mat<T, 3, 3> J1 = block<T, 3, 6, 3, 3>(J, 0, 0), J2 = block<T, 3, 6, 3, 3>(J, 0, 3);
    cov = add(mul(mul(J1, cov), tr(J1)), mul(mul(J2, covu), tr(J2)));
    pose *= dpose;
  }
};

// ls_extractor/utils.h:23-30
inline void checkRhoTheta(float rt[2]) {
  if (rt[0] < 0) {
    rt[0] = -rt[0];
    rt[1] += (float)M_PI;
    if (rt[1] > (float)M_PI) rt[1] -= (float)(2 * M_PI);
  }
}

}  // namespace

extern "C" {

// submap_loop_closer.cpp:206-223: for every copied pose, in order:
//   edge->setMeasurement((it - 1)->pose.estimate().inverse() * it->pose.estimate());
//   pose->setEstimate(prev_vertex->estimate() * edge->measurement());  prev_vertex = pose;
void sgo_pg_append(const double* prev_pg_est, const double* lm_est, int32_t count, double* z_out, double* est_out) {
  SE2 prev(prev_pg_est[0], prev_pg_est[1], prev_pg_est[2]);
  for (int k = 0; k < count; ++k) {
    SE2 a(lm_est[3 * k], lm_est[3 * k + 1], lm_est[3 * k + 2]);
    SE2 b(lm_est[3 * k + 3], lm_est[3 * k + 4], lm_est[3 * k + 5]);
    SE2 z = a.inverse() * b;
    SE2 e = prev * z;
    for (int c = 0; c < 3; ++c) {
      z_out[3 * k + c] = z[c];
      est_out[3 * k + c] = e[c];
    }
    prev = e;
  }
}

// log_runner.cpp:182-184: edge.computeError(); edge.chi2()  -- EdgeSE2: e = (z^-1 * (xi^-1 * xj)).toVector(), chi2 = e^T Omega e
void sgo_closure_chi2(const double* est, const int32_t* ei, const int32_t* ej, const double* z, const double* info6, int32_t n,
                      double* chi_out) {
  for (int k = 0; k < n; ++k) {
    const double *pi = est + 3 * (size_t)ei[k], *pj = est + 3 * (size_t)ej[k];
    SE2 xi(pi[0], pi[1], pi[2]), xj(pj[0], pj[1], pj[2]), zz(z[3 * k], z[3 * k + 1], z[3 * k + 2]);
    SE2 d = zz.inverse() * (xi.inverse() * xj);
    double e[3] = {d[0], d[1], d[2]};
    const double* u = info6 + 6 * (size_t)k;
    double O[3][3] = {{u[0], u[1], u[2]}, {u[1], u[3], u[4]}, {u[2], u[4], u[5]}};
    double c = 0;
    for (int r = 0; r < 3; ++r) {
      double s = 0;
      for (int q = 0; q < 3; ++q) s += O[r][q] * e[q];
      c += e[r] * s;
    }
    chi_out[k] = c;
  }
}

// drone.cpp:84 (step per odometry message), :127-128 (measurement = pose, information = cov.inverse()), :143 (reset)
void sgo_odom_information(const double* deltas, const int32_t* seg_ptr, int32_t n_seg, double std_x, double std_y, double std_w,
                          double* z_out, double* cov_out, double* info_out) {
  OdomErrorPropagator<double> prop(std_x, std_y, std_w);
  for (int s = 0; s < n_seg; ++s) {
    prop.reset();
    for (int k = seg_ptr[s]; k < seg_ptr[s + 1]; ++k) prop.step(SE2(deltas[3 * k], deltas[3 * k + 1], deltas[3 * k + 2]));
    for (int c = 0; c < 3; ++c) z_out[3 * s + c] = prop.pose[c];
    const Mat<double, 3, 3>& m = prop.cov;
    if (cov_out)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov_out[9 * s + 3 * i + j] = m(i, j);
    // Eigen fixed-size 3x3 inverse: cofactor matrix / determinant
    double inv[3][3];
    double c00 = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
    double c10 = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
    double c20 = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
    double det = m(0, 0) * c00 + m(0, 1) * c10 + m(0, 2) * c20;
    double id = 1.0 / det;
    inv[0][0] = c00 * id;
    inv[0][1] = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * id;
    inv[0][2] = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * id;
    inv[1][1] = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * id;
    inv[1][2] = (m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2)) * id;
    inv[2][2] = (m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0)) * id;
    double* o = info_out + 6 * (size_t)s;
    o[0] = inv[0][0]; o[1] = inv[0][1]; o[2] = inv[0][2]; o[3] = inv[1][1]; o[4] = inv[1][2]; o[5] = inv[2][2];
  }
}

// multicloud2.cpp:56-83 for n_windows windows
void sgo_scan_point_covariances(const double* deltas, int32_t n_windows, int32_t n_scans, int32_t scan_size,
                                const float* beam_cos_sin, const float* pts, float std_x, float std_y, float std_w, float var_r,
                                float* cov_out, float* rhotheta_out, uint8_t* valid_out) {
  OdomErrorPropagator<float> odom_prop((float)(double)std_x, (float)(double)std_y, (float)(double)std_w);
  Mat<float, 2, 5> Jl;
  Jl(0, 0) = Jl(1, 1) = 1.0f;
  const int delta_offset = n_scans - 1;
  for (int w = 0; w < n_windows; ++w) {
    const double* wd = deltas + 3 * (size_t)w * delta_offset;
    for (int i = 0; i < n_scans; ++i) {
      odom_prop.reset();
      for (int j = i; j < delta_offset; ++j) odom_prop.step(SE2(wd[3 * j], wd[3 * j + 1], wd[3 * j + 2]));
      float ct = std::cos((float)odom_prop.pose[2]), st = std::sin((float)odom_prop.pose[2]);
      Mat<float, 3, 3> Juk;
      Juk(0, 0) = -ct; Juk(0, 1) = st; Juk(0, 2) = (float)(odom_prop.pose[1] * ct + odom_prop.pose[0] * st);
      Juk(1, 0) = -st; Juk(1, 1) = -ct; Juk(1, 2) = (float)(odom_prop.pose[1] * st - odom_prop.pose[0] * ct);
      Juk(2, 0) = 0.0f; Juk(2, 1) = 0.0f; Juk(2, 2) = -1.0f;
      odom_prop.cov = mul(mul(Juk, odom_prop.cov), tr(Juk));
      odom_prop.pose = odom_prop.pose.inverse();
      updateJacobian(Jl, (float)odom_prop.pose[0], (float)odom_prop.pose[1], (float)odom_prop.pose[2]);
      Mat<float, 2, 3> Ja = block<float, 2, 5, 2, 3>(Jl, 0, 0);
      Mat<float, 2, 2> Jb = block<float, 2, 5, 2, 2>(Jl, 0, 3);
      for (int j = 0; j < scan_size; ++j) {
        size_t p = ((size_t)w * n_scans + i) * scan_size + j;
        float x = pts[2 * p], y = pts[2 * p + 1];
        bool ok = std::isfinite(x) && std::isfinite(y);
        valid_out[p] = ok ? 1 : 0;
        for (int c = 0; c < 4; ++c) cov_out[4 * p + c] = 0.0f;
        rhotheta_out[2 * p] = rhotheta_out[2 * p + 1] = 0.0f;
        if (!ok) continue;
        rhotheta_out[2 * p] = std::sqrt(x * x + y * y);
        rhotheta_out[2 * p + 1] = std::atan2(y, x);
        const float cv = beam_cos_sin[2 * j], sv = beam_cos_sin[2 * j + 1];
        Mat<float, 2, 2> covp;
        const float c = cv * sv;
        covp(0, 0) = cv * cv; covp(0, 1) = c; covp(1, 0) = c; covp(1, 1) = sv * sv;
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) covp(a, b) *= var_r;
        Mat<float, 2, 2> r = add(mul(mul(Ja, odom_prop.cov), tr(Ja)), mul(mul(Jb, covp), tr(Jb)));
        cov_out[4 * p] = r(0, 0); cov_out[4 * p + 1] = r(0, 1); cov_out[4 * p + 2] = r(1, 0); cov_out[4 * p + 3] = r(1, 1);
      }
    }
  }
}

// smc.cpp:30-68 (leastSqFit) per segment; drone.cpp:203 information = cov.cast<double>().inverse()
void sgo_line_fit_information(const float* pts, const float* pcov, const int32_t* seg_ptr, int32_t n_seg, float* rhotheta_out,
                              float* cov_out, double* info_out) {
  for (int s = 0; s < n_seg; ++s) {
    const int start = seg_ptr[s], end = seg_ptr[s + 1];
    float xybar[2] = {0, 0}, Sx2y2[2] = {0, 0}, Sxy = 0.0;
    for (int it = start; it < end; ++it) {
      const float* pt = pts + 2 * (size_t)it;
      xybar[0] += pt[0]; xybar[1] += pt[1];
      Sxy += pt[0] * pt[1];
      Sx2y2[0] += pt[0] * pt[0]; Sx2y2[1] += pt[1] * pt[1];
    }
    int n = end - start;
    xybar[0] /= n; xybar[1] /= n;
    Sx2y2[0] -= n * (xybar[0] * xybar[0]); Sx2y2[1] -= n * (xybar[1] * xybar[1]);
    Sxy -= n * (xybar[0] * xybar[1]);
    float Sy2_Sx2 = Sx2y2[1] - Sx2y2[0];
    float rt[2];
    rt[1] = 0.5 * std::atan2(-2 * Sxy, Sy2_Sx2);
    float ct = std::cos(rt[1]), st = std::sin(rt[1]);
    rt[0] = xybar[0] * ct + xybar[1] * st;
    checkRhoTheta(rt);
    ct = std::cos(rt[1]), st = std::sin(rt[1]);
    float xbar_st = xybar[0] * st, ybar_ct = xybar[1] * ct;
    Mat<float, 2, 2> cov;
    float denum = 1.0 / (Sy2_Sx2 * Sy2_Sx2 + 4 * Sxy * Sxy);
    float ct_n = ct / n, st_n = st / n;
    for (int it = start; it < end; ++it) {
      Mat<float, 2, 2> Ai, C;
      float d[2] = {xybar[0] - pts[2 * (size_t)it], xybar[1] - pts[2 * (size_t)it + 1]};
      Ai(1, 0) = (d[1] * Sy2_Sx2 + 2 * Sxy * d[0]) * denum;
      Ai(1, 1) = (d[0] * Sy2_Sx2 - 2 * Sxy * d[1]) * denum;
      Ai(0, 0) = ct_n - xbar_st * Ai(1, 0) + ybar_ct * Ai(1, 0);
      Ai(0, 1) = st_n - xbar_st * Ai(1, 1) + ybar_ct * Ai(1, 1);
      const float* pc = pcov + 4 * (size_t)it;
      C(0, 0) = pc[0]; C(0, 1) = pc[1]; C(1, 0) = pc[2]; C(1, 1) = pc[3];
      cov = add(cov, mul(mul(Ai, C), tr(Ai)));
    }
    rhotheta_out[2 * s] = rt[0];
    rhotheta_out[2 * s + 1] = rt[1];
    cov_out[4 * s] = cov(0, 0); cov_out[4 * s + 1] = cov(0, 1); cov_out[4 * s + 2] = cov(1, 0); cov_out[4 * s + 3] = cov(1, 1);
    double a = cov(0, 0), b = cov(0, 1), c = cov(1, 0), d = cov(1, 1);
    double id = 1.0 / (a * d - b * c);
    info_out[3 * s] = d * id;
    info_out[3 * s + 1] = -b * id;
    info_out[3 * s + 2] = a * id;
  }
}

}  // extern "C"
