// Internal declarations shared by the translation units of libuwvkf.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/uwvkf.h"
#include "pose_engine.cuh"
#include "small_engine.cuh"

namespace uwvkf {
int fail(int code, const std::string &msg);
bool any_nan(const double *v, int64_t n);   // host scan (checkMeasurment() of the reference throws on NaN)
}

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return uwvkf::fail(UWVKF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

struct uwvkf_filter_s {
  int kind = 0;
  int64_t B = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0, launches_at_capture = 0;
  bool capturing = false;
  uint32_t *d_status = nullptr;
  double mean_tol = 1e-9;
  int mean_max_it = 100;
  // staging for host inputs
  double *d_z = nullptr, *d_Rpf = nullptr, *d_z2 = nullptr;
  uint8_t *d_valid = nullptr, *d_valid2 = nullptr;
  double *d_tmp = nullptr;
  size_t tmp_bytes = 0;
  // PoseUKF
  double *d_state = nullptr, *d_offsets = nullptr, *d_rot = nullptr, *d_Q = nullptr, *d_uwv = nullptr;
  uwvkf::pose::Params pp{};
  bool have_params = false, have_Q = false;
  double *d_tm = nullptr;        // [B][10]: latitudes of the predict's sigma points (pose_tmerc_prepass_kernel)
  double *nav_fused = nullptr;   // set around the fused launch of uwvkf_pose_step_host: the kernel writes the navigation solution itself
  // uwvkf_pose_step_host: two input slots / two navigation buffers, upload and download streams, events per slot
  struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    double *d_w[2] = {nullptr, nullptr}, *d_z[2] = {nullptr, nullptr}, *d_nav[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_main[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool out_pending[2] = {false, false};
    int64_t calls = 0;
  } *pipe = nullptr;
  // small filters: state[B][dim + dof*dof], controls, shared process noise
  int dim = 0, dof = 0;
  double Qh[36] = {0};
  double *d_ctrl0 = nullptr, *d_ctrl1 = nullptr, *d_sim = nullptr;  // velocity: gyro, efforts, simulator; bottom: velocity; indirect: pose_ref
  uwvkf::small::UwvParams uwv{};
  bool have_model = false;
  double tau = 0.0;
};

namespace uwvkf {
int ensure_tmp(uwvkf_handle h, size_t bytes);
int stage(uwvkf_handle h, const void *src, size_t bytes, size_t cap, int memspace, void **slot, const void **out);
// small filters (small_filters.cu)
int small_create(uwvkf_handle h);
void small_destroy(uwvkf_handle h);
int small_set_state(uwvkf_handle h, const double *mu, const double *cov, int64_t cov_stride, int64_t first, int64_t count, int memspace);
int small_get_state(uwvkf_handle h, double *mu, double *cov, int64_t first, int64_t count, int memspace);
int small_predict(uwvkf_handle h, double dt);
int small_ensemble_stats(uwvkf_handle h, double *dout);
}  // namespace uwvkf

#define REQUIRE(h, k)                                                                                             \
  if (!(h)) return uwvkf::fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");                                        \
  if ((h)->kind != (k)) return uwvkf::fail(UWVKF_ERR_INVALID_ARGUMENT, "wrong filter kind for this call");       \
  CK(cudaSetDevice((h)->device));
