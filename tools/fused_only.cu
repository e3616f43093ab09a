// ptxas statistics of ONE instantiation of the fused kernel (build-time probe, not part of the library): tools/ptxas_probe.sh
#include "../slam-uwv_kalman_filters_b200/csrc/internal.h"
#include "../slam-uwv_kalman_filters_b200/csrc/pose_fused.cuh"
#ifndef PROBE_WARPS
#define PROBE_WARPS 12
#endif
#ifndef PROBE_KIND
#define PROBE_KIND M_VELOCITY
#endif
#ifndef PROBE_KIND2
#define PROBE_KIND2 M_NONE
#endif
#ifndef PROBE_PREDICT
#define PROBE_PREDICT 1
#endif
template __global__ void uwvkf::pose::pose_fused_kernel<PROBE_PREDICT, uwvkf::pose::PROBE_KIND, PROBE_WARPS, uwvkf::pose::PROBE_KIND2>(
    uwvkf::pose::Launch, const __grid_constant__ uwvkf::pose::Params);
