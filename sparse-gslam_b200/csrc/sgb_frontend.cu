// sgb_frontend.cu -- information-matrix producers as batch kernels: what feeds the edge SoA of the
// optimiser. Reference: include/odom_error_propagator.h:6-46 + drone.cpp:84,127-128,143 (odometry edges),
// src/multicloud2.cpp:56-83 (scan-point covariances), src/ls_extractor/src/impl/smc.cpp:30-68 + drone.cpp:203
// (line fit, pose-line edge information). The reference runs these one key-frame / one segment at a time on the CPU;
// here a whole log's worth (or many robots') is one launch. All three are streaming kernels (HBM-bound; the
// sequential recurrences are per key-frame interval / per segment and short), no tensor cores. No CPU path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <string>

#include "../../include/sgb_capi.h"
#include "sgb_edits.h"

using namespace sgb;

namespace {

constexpr int kFeThreads = 256;
int fe_grid(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + kFeThreads - 1) / kFeThreads, 148 * 16)); }

// one thread per key-frame interval
__global__ void __launch_bounds__(kFeThreads) k_odom_information(const double* __restrict__ deltas, const int32_t* __restrict__ seg_ptr,
                                                                int n_seg, double var_x, double var_y, double var_w,
                                                                double* __restrict__ z_out, double* __restrict__ cov_out,
                                                                double* __restrict__ info_out) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
    int a = seg_ptr[s], b = seg_ptr[s + 1];
    Se2 pose;
    double cov[9], inf[6];
    odom_propagate<double>(deltas + 3 * (size_t)a, b - a, var_x, var_y, var_w, &pose, cov);
    inv3_general_upper(cov, inf);
    se2_store(z_out + 3 * (size_t)s, pose);
    if (cov_out)
      for (int i = 0; i < 9; ++i) cov_out[9 * (size_t)s + i] = cov[i];
    for (int i = 0; i < 6; ++i) info_out[6 * (size_t)s + i] = inf[i];
  }
}

// one thread per (window, scan): the scan's pose covariance and point Jacobian in the frame of the newest scan
__global__ void __launch_bounds__(kFeThreads) k_scan_frames(const double* __restrict__ deltas, int n_windows, int n_scans, float var_x,
                                                           float var_y, float var_w, ScanFrame* __restrict__ frames) {
  const int total = n_windows * n_scans;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    int w = q / n_scans, i = q - w * n_scans;
    const double* d = deltas + 3 * ((size_t)w * (n_scans - 1) + i);  // deltas i .. n_scans-2 of window w
    ScanFrame f;
    scan_frame(d, n_scans - 1 - i, var_x, var_y, var_w, &f);
    frames[q] = f;
  }
}
// one thread per point; the frame of a scan is shared by scan_size consecutive threads (L1-resident)
__global__ void __launch_bounds__(kFeThreads) k_scan_points(const ScanFrame* __restrict__ frames, const float2* __restrict__ beam,
                                                           const float2* __restrict__ pts, long long n_points, int scan_size,
                                                           float var_r, float4* __restrict__ cov_out, float2* __restrict__ rt_out,
                                                           uint8_t* __restrict__ valid_out) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += (long long)gridDim.x * blockDim.x) {
    long long scan = p / scan_size;
    int j = (int)(p - scan * scan_size);
    float2 xy = pts[p];
    float c[4] = {0.f, 0.f, 0.f, 0.f}, rt[2] = {0.f, 0.f};
    bool ok = isfinite(xy.x) && isfinite(xy.y);
    if (ok) {
      float2 b = __ldg(&beam[j]);
      scan_point_cov(frames[scan], b.x, b.y, var_r, xy.x, xy.y, c, rt);
    }
    cov_out[p] = make_float4(c[0], c[1], c[2], c[3]);
    rt_out[p] = make_float2(rt[0], rt[1]);
    valid_out[p] = ok ? 1 : 0;
  }
}

// one thread per segment (the sums and the Jacobian accumulation run in the reference's point order)
__global__ void __launch_bounds__(kFeThreads) k_line_fit(const float* __restrict__ pts, const float* __restrict__ pcov,
                                                        const int32_t* __restrict__ seg_ptr, int n_seg, float* __restrict__ rt_out,
                                                        float* __restrict__ cov_out, double* __restrict__ info_out) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
    int a = seg_ptr[s], b = seg_ptr[s + 1];
    float rt[2], cov[4];
    double inf[3];
    line_fit(pts + 2 * (size_t)a, pcov + 4 * (size_t)a, b - a, rt, cov);
    line_info(cov, inf);
    rt_out[2 * (size_t)s] = rt[0];
    rt_out[2 * (size_t)s + 1] = rt[1];
    for (int i = 0; i < 4; ++i) cov_out[4 * (size_t)s + i] = cov[i];
    for (int i = 0; i < 3; ++i) info_out[3 * (size_t)s + i] = inf[i];
  }
}

thread_local std::string t_err;

// grow-only device workspace per host thread and device (the producers are called once per key-frame batch)
struct Workspace {
  int device = -1;
  char* base = nullptr;
  size_t cap = 0, off = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ~Workspace() {
    if (device < 0) return;
    cudaSetDevice(device);
    if (base) cudaFree(base);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (stream) cudaStreamDestroy(stream);
  }
};
thread_local Workspace t_ws;

#define FE_CUDA(call)                                              \
  do {                                                             \
    cudaError_t _e = (call);                                       \
    if (_e != cudaSuccess) {                                       \
      t_err = std::string(#call) + ": " + cudaGetErrorString(_e);  \
      return SGB_ERR_CUDA;                                         \
    }                                                              \
  } while (0)

sgb_status ws_begin(int device, size_t bytes, Workspace** out) {
  if (sgb_device_count() <= 0) { t_err = "no CUDA device (there is no CPU path)"; return SGB_ERR_NO_DEVICE; }
  if (device >= 0) FE_CUDA(cudaSetDevice(device));
  int dev = 0;
  FE_CUDA(cudaGetDevice(&dev));
  Workspace& w = t_ws;
  if (w.device != dev) {
    if (w.device >= 0) {
      cudaSetDevice(w.device);
      if (w.base) cudaFree(w.base);
      if (w.e0) cudaEventDestroy(w.e0);
      if (w.e1) cudaEventDestroy(w.e1);
      if (w.stream) cudaStreamDestroy(w.stream);
      cudaSetDevice(dev);
    }
    w = Workspace();
    w.device = dev;
    FE_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    FE_CUDA(cudaEventCreate(&w.e0));
    FE_CUDA(cudaEventCreate(&w.e1));
  }
  if (bytes > w.cap) {
    if (w.base) cudaFree(w.base);
    w.base = nullptr;
    w.cap = 0;
    size_t cap = bytes + bytes / 4;
    FE_CUDA(cudaMalloc((void**)&w.base, cap));
    w.cap = cap;
  }
  w.off = 0;
  *out = &w;
  return SGB_OK;
}
size_t al(size_t b) { return (b + 255) & ~(size_t)255; }
template <class T>
T* ws_take(Workspace* w, size_t n) {
  T* p = (T*)(w->base + w->off);
  w->off += al(std::max<size_t>(n, 1) * sizeof(T));
  return p;
}
sgb_status ws_finish(Workspace* w, double* kernel_ms) {
  FE_CUDA(cudaStreamSynchronize(w->stream));
  if (kernel_ms) {
    float ms = 0.f;
    FE_CUDA(cudaEventElapsedTime(&ms, w->e0, w->e1));
    *kernel_ms = ms;
  }
  return SGB_OK;
}

}  // namespace

extern "C" {

const char* sgb_frontend_last_error(void) { return t_err.c_str(); }

sgb_status sgb_odom_information(int32_t device, const double* deltas, const int32_t* seg_ptr, int32_t n_seg, double std_x,
                                double std_y, double std_w, double* z_out, double* cov_out, double* info_out, double* kernel_ms) {
  if (kernel_ms) *kernel_ms = 0.0;
  if (n_seg < 0 || (n_seg > 0 && (!seg_ptr || !z_out || !info_out))) { t_err = "odom_information: bad arguments"; return SGB_ERR_INVALID; }
  if (n_seg == 0) return SGB_OK;
  const size_t n_steps = (size_t)seg_ptr[n_seg];
  if (seg_ptr[0] != 0 || (n_steps > 0 && !deltas)) { t_err = "odom_information: bad segment table"; return SGB_ERR_INVALID; }
  for (int s = 0; s < n_seg; ++s)
    if (seg_ptr[s + 1] < seg_ptr[s]) { t_err = "odom_information: segment table not monotone"; return SGB_ERR_INVALID; }
  size_t bytes = al(3 * n_steps * 8 + 8) + al((size_t)(n_seg + 1) * 4) + al(3 * (size_t)n_seg * 8) + al(9 * (size_t)n_seg * 8) + al(6 * (size_t)n_seg * 8);
  Workspace* w;
  sgb_status st = ws_begin(device, bytes, &w);
  if (st != SGB_OK) return st;
  double* d_d = ws_take<double>(w, 3 * n_steps);
  int32_t* d_p = ws_take<int32_t>(w, (size_t)n_seg + 1);
  double* d_z = ws_take<double>(w, 3 * (size_t)n_seg);
  double* d_c = ws_take<double>(w, 9 * (size_t)n_seg);
  double* d_i = ws_take<double>(w, 6 * (size_t)n_seg);
  if (n_steps) FE_CUDA(cudaMemcpyAsync(d_d, deltas, 3 * n_steps * 8, cudaMemcpyHostToDevice, w->stream));
  FE_CUDA(cudaMemcpyAsync(d_p, seg_ptr, ((size_t)n_seg + 1) * 4, cudaMemcpyHostToDevice, w->stream));
  FE_CUDA(cudaEventRecord(w->e0, w->stream));
  k_odom_information<<<fe_grid(n_seg), kFeThreads, 0, w->stream>>>(d_d, d_p, n_seg, std_x * std_x, std_y * std_y, std_w * std_w, d_z,
                                                                  cov_out ? d_c : nullptr, d_i);
  FE_CUDA(cudaGetLastError());
  FE_CUDA(cudaEventRecord(w->e1, w->stream));
  FE_CUDA(cudaMemcpyAsync(z_out, d_z, 3 * (size_t)n_seg * 8, cudaMemcpyDeviceToHost, w->stream));
  if (cov_out) FE_CUDA(cudaMemcpyAsync(cov_out, d_c, 9 * (size_t)n_seg * 8, cudaMemcpyDeviceToHost, w->stream));
  FE_CUDA(cudaMemcpyAsync(info_out, d_i, 6 * (size_t)n_seg * 8, cudaMemcpyDeviceToHost, w->stream));
  return ws_finish(w, kernel_ms);
}

sgb_status sgb_scan_point_covariances(int32_t device, const double* deltas, int32_t n_windows, int32_t n_scans, int32_t scan_size,
                                      const float* beam_cos_sin, const float* pts, float std_x, float std_y, float std_w,
                                      float var_r, float* cov_out, float* rhotheta_out, uint8_t* valid_out, double* kernel_ms) {
  if (kernel_ms) *kernel_ms = 0.0;
  if (n_windows < 0 || n_scans < 1 || scan_size < 1) { t_err = "scan_point_covariances: bad sizes"; return SGB_ERR_INVALID; }
  if (n_windows == 0) return SGB_OK;
  if (!beam_cos_sin || !pts || !cov_out || !rhotheta_out || !valid_out || (n_scans > 1 && !deltas)) {
    t_err = "scan_point_covariances: missing array";
    return SGB_ERR_INVALID;
  }
  const size_t n_d = 3 * (size_t)n_windows * (n_scans - 1), n_f = (size_t)n_windows * n_scans;
  const long long n_pts = (long long)n_f * scan_size;
  size_t bytes = al(n_d * 8 + 8) + al(n_f * sizeof(ScanFrame)) + al(2 * (size_t)scan_size * 4) + al((size_t)n_pts * 8) +
                 al((size_t)n_pts * 16) + al((size_t)n_pts * 8) + al((size_t)n_pts);
  Workspace* w;
  sgb_status st = ws_begin(device, bytes, &w);
  if (st != SGB_OK) return st;
  double* d_d = ws_take<double>(w, n_d);
  ScanFrame* d_f = ws_take<ScanFrame>(w, n_f);
  float2* d_b = ws_take<float2>(w, scan_size);
  float2* d_p = ws_take<float2>(w, (size_t)n_pts);
  float4* d_c = ws_take<float4>(w, (size_t)n_pts);
  float2* d_r = ws_take<float2>(w, (size_t)n_pts);
  uint8_t* d_v = ws_take<uint8_t>(w, (size_t)n_pts);
  if (n_d) FE_CUDA(cudaMemcpyAsync(d_d, deltas, n_d * 8, cudaMemcpyHostToDevice, w->stream));
  FE_CUDA(cudaMemcpyAsync(d_b, beam_cos_sin, 2 * (size_t)scan_size * 4, cudaMemcpyHostToDevice, w->stream));
  FE_CUDA(cudaMemcpyAsync(d_p, pts, (size_t)n_pts * 8, cudaMemcpyHostToDevice, w->stream));
  FE_CUDA(cudaEventRecord(w->e0, w->stream));
  k_scan_frames<<<fe_grid((long long)n_f), kFeThreads, 0, w->stream>>>(d_d, n_windows, n_scans, std_x * std_x, std_y * std_y,
                                                                     std_w * std_w, d_f);
  k_scan_points<<<fe_grid(n_pts), kFeThreads, 0, w->stream>>>(d_f, d_b, d_p, n_pts, scan_size, var_r, d_c, d_r, d_v);
  FE_CUDA(cudaGetLastError());
  FE_CUDA(cudaEventRecord(w->e1, w->stream));
  FE_CUDA(cudaMemcpyAsync(cov_out, d_c, (size_t)n_pts * 16, cudaMemcpyDeviceToHost, w->stream));
  FE_CUDA(cudaMemcpyAsync(rhotheta_out, d_r, (size_t)n_pts * 8, cudaMemcpyDeviceToHost, w->stream));
  FE_CUDA(cudaMemcpyAsync(valid_out, d_v, (size_t)n_pts, cudaMemcpyDeviceToHost, w->stream));
  return ws_finish(w, kernel_ms);
}

sgb_status sgb_line_fit_information(int32_t device, const float* pts, const float* pcov, const int32_t* seg_ptr, int32_t n_seg,
                                    float* rhotheta_out, float* cov_out, double* info_out, double* kernel_ms) {
  if (kernel_ms) *kernel_ms = 0.0;
  if (n_seg < 0 || (n_seg > 0 && (!seg_ptr || !rhotheta_out || !cov_out || !info_out))) { t_err = "line_fit_information: bad arguments"; return SGB_ERR_INVALID; }
  if (n_seg == 0) return SGB_OK;
  const size_t n_pts = (size_t)seg_ptr[n_seg];
  if (seg_ptr[0] != 0 || (n_pts > 0 && (!pts || !pcov))) { t_err = "line_fit_information: bad segment table"; return SGB_ERR_INVALID; }
  for (int s = 0; s < n_seg; ++s)
    if (seg_ptr[s + 1] < seg_ptr[s]) { t_err = "line_fit_information: segment table not monotone"; return SGB_ERR_INVALID; }
  size_t bytes = al(n_pts * 8 + 8) + al(n_pts * 16 + 8) + al(((size_t)n_seg + 1) * 4) + al((size_t)n_seg * 8) + al((size_t)n_seg * 16) + al((size_t)n_seg * 24);
  Workspace* w;
  sgb_status st = ws_begin(device, bytes, &w);
  if (st != SGB_OK) return st;
  float* d_p = ws_take<float>(w, 2 * n_pts);
  float* d_c = ws_take<float>(w, 4 * n_pts);
  int32_t* d_s = ws_take<int32_t>(w, (size_t)n_seg + 1);
  float* d_rt = ws_take<float>(w, 2 * (size_t)n_seg);
  float* d_cv = ws_take<float>(w, 4 * (size_t)n_seg);
  double* d_in = ws_take<double>(w, 3 * (size_t)n_seg);
  if (n_pts) {
    FE_CUDA(cudaMemcpyAsync(d_p, pts, n_pts * 8, cudaMemcpyHostToDevice, w->stream));
    FE_CUDA(cudaMemcpyAsync(d_c, pcov, n_pts * 16, cudaMemcpyHostToDevice, w->stream));
  }
  FE_CUDA(cudaMemcpyAsync(d_s, seg_ptr, ((size_t)n_seg + 1) * 4, cudaMemcpyHostToDevice, w->stream));
  FE_CUDA(cudaEventRecord(w->e0, w->stream));
  k_line_fit<<<fe_grid(n_seg), kFeThreads, 0, w->stream>>>(d_p, d_c, d_s, n_seg, d_rt, d_cv, d_in);
  FE_CUDA(cudaGetLastError());
  FE_CUDA(cudaEventRecord(w->e1, w->stream));
  FE_CUDA(cudaMemcpyAsync(rhotheta_out, d_rt, (size_t)n_seg * 8, cudaMemcpyDeviceToHost, w->stream));
  FE_CUDA(cudaMemcpyAsync(cov_out, d_cv, (size_t)n_seg * 16, cudaMemcpyDeviceToHost, w->stream));
  FE_CUDA(cudaMemcpyAsync(info_out, d_in, (size_t)n_seg * 24, cudaMemcpyDeviceToHost, w->stream));
  return ws_finish(w, kernel_ms);
}

}  // extern "C"
