// ptxas statistics of the headline instantiation only (build-time probe, not part of the library)
#include "../slam-uwv_kalman_filters_b200/csrc/internal.h"
#include "../slam-uwv_kalman_filters_b200/csrc/pose_fused.cuh"
#ifndef PROBE_WARPS
#define PROBE_WARPS 12
#endif
template __global__ void uwvkf::pose::pose_fused_kernel<1, uwvkf::pose::M_VELOCITY, PROBE_WARPS>(uwvkf::pose::Launch, const __grid_constant__ uwvkf::pose::Params);
