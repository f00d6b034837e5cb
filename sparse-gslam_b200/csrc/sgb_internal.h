// sgb_internal.h -- the few pieces of a handle that the other translation units of libsgb.so (sgb_posegraph.cu)
// need: its stream, its device-resident current estimates, its error string. Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/sgb_capi.h"

namespace sgb {

struct HandleView {
  int device;
  cudaStream_t stream;
  int n_poses, n_landmarks;  // array sizes of the graph given to sgb_set_graph*
  double* pose_est;          // [3*n_poses] current estimates on the device (this rank's replica)
  double* lm_est;            // [2*n_landmarks]
};
// false when the handle has no graph
bool handle_view(sgb_handle* h, HandleView* out);
// CUDA device of a handle (valid before any graph is set)
int handle_device(const sgb_handle* h);
void handle_set_error(sgb_handle* h, const std::string& msg);

}  // namespace sgb
