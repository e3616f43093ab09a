// uwvkf C ABI (include/uwvkf.h) + kernel launchers.  Thin host code over the CUDA kernels; no CPU path.
#include "internal.h"
#include "pose_fused.cuh"
#include "pose_marker.cuh"
#include <algorithm>
#include <cstdlib>

namespace uwvkf {
thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }
// checkMeasurment() of the reference throws on NaN (src/PoseUKF.cpp:478,494): HOST inputs are scanned for every batch size.
// Branch-free blocks of 4096 so that the host compiler vectorises the scan (memory bound: ~0.1 ms per million doubles).
bool any_nan(const double *v, int64_t n) {
  for (int64_t o = 0; o < n; o += 4096) {
    const int64_t e = std::min<int64_t>(n, o + 4096);
    int bad = 0;
    for (int64_t k = o; k < e; ++k) bad |= (v[k] != v[k]);
    if (bad) return true;
  }
  return false;
}
}  // namespace uwvkf

namespace {

using namespace uwvkf;

__device__ uint16_t g_lut[pose::NP];  // packed index -> (row | col << 8)

// ------------------------------------------------------------------------------------------------------
// PoseUKF step kernel: one warp per filter, persistent CTAs, filter block staged in shared memory.
template <int KMAX, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) pose_step_kernel(pose::Launch a, const __grid_constant__ pose::Params P) {
  using W = pose::WL<KMAX>;
  extern __shared__ __align__(16) double smem[];
  double *Qs = smem;
  uint16_t *lut = reinterpret_cast<uint16_t *>(smem + 1432);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *w = smem + 1432 + 360 + warp * W::TOTAL;
  for (int e = threadIdx.x; e < pose::NP; e += blockDim.x) {
    Qs[e] = a.Q[e];
    lut[e] = g_lut[e];
  }
  __syncthreads();
  for (long long f = (long long)blockIdx.x * WARPS + warp; f < a.B; f += (long long)gridDim.x * WARPS) {
    {
      const double2 *src = reinterpret_cast<const double2 *>(a.state + f * pose::BLK);
      double2 *dst = reinterpret_cast<double2 *>(w);
#pragma unroll 4
      for (int k = lane; k < pose::BLK / 2; k += 32) dst[k] = __ldcs(src + k);
    }
    __syncwarp();
    unsigned st = 0;
    bool dirty = false;
    for (int o = 0; o < a.nops; ++o) {
      const pose::Op &op = a.ops[o];
      if (op.type == 0) {
        dirty |= pose::predict<KMAX>(w, Qs, lut, a.offsets + f * pose::NOFF, a.rot_rate + f * 3, P, op.dt, lane, st);
        continue;
      }
      if (op.valid && !op.valid[f]) continue;
      const int kind = op.kind;
      const bool efforts = (kind == pose::M_BODY_EFFORTS || kind == pose::M_BODY_EFFORTS_VELOCITY_ONLY);
      const int m = efforts ? 6 : ((kind == pose::M_Z || kind == pose::M_PRESSURE) ? 1 : ((kind == pose::M_VELOCITY || kind == pose::M_ACCELERATION) ? 3 : 2));
      constexpr int MZ = (KMAX >= 48 || KMAX == 20) ? pose::MAXM : 3;
      double z[MZ], R[MZ * MZ];
      bool bad = false;
      if (m > MZ) continue;   // six-dimensional measurements are routed to the dense instantiation by the host
      for (int k = 0; k < m; ++k) { z[k] = op.z[f * m + k]; bad |= isnan(z[k]); }
      for (int k = 0; k < m * m; ++k) { R[k] = op.Rdev ? op.Rdev[f * m * m + k] : op.R[k]; bad |= isnan(R[k]); }
      if (bad) { st |= pose::ST_NAN_MEAS; continue; }
      switch (kind) {
        case pose::M_XY: dirty |= pose::update<pose::M_XY, KMAX>(w, lut, z, R, op.aux, false, P, lane, st); break;
        case pose::M_Z: dirty |= pose::update<pose::M_Z, KMAX>(w, lut, z, R, op.aux, false, P, lane, st); break;
        case pose::M_VELOCITY: dirty |= pose::update<pose::M_VELOCITY, KMAX>(w, lut, z, R, op.aux, false, P, lane, st); break;
        case pose::M_ACCELERATION: dirty |= pose::update<pose::M_ACCELERATION, KMAX>(w, lut, z, R, op.aux, false, P, lane, st); break;
        case pose::M_GEOGRAPHIC: {  // reference src/PoseUKF.cpp:572-578
          double px, py;
          tmerc_world_to_nav(P.tm, z[0], z[1], px, py);
          Vec3 lever = qrot(Quat{w[3], w[4], w[5], w[6]}, Vec3{op.aux[0], op.aux[1], op.aux[2]});
          z[0] = px - lever.x; z[1] = py - lever.y;
          dirty |= pose::update<pose::M_XY, KMAX>(w, lut, z, R, op.aux, true, P, lane, st);
          break;
        }
        case pose::M_DELAYED_XY: {  // reference src/PoseUKF.cpp:516-521
          z[0] = z[0] + (w[0] - op.aux[0]); z[1] = z[1] + (w[1] - op.aux[1]);
          dirty |= pose::update<pose::M_XY, KMAX>(w, lut, z, R, op.aux, false, P, lane, st);
          break;
        }
        default:
          if constexpr (KMAX == 20) {   // constrainVelocity (reference src/PoseUKF.cpp:198-219, 585-592): prefix 9, m = 6
            if (kind == pose::M_BODY_EFFORTS_VELOCITY_ONLY) {
              double *ctx = w + W::F;
              pose::efforts_context(w, a.rot_rate + f * 3, P, ctx);
              __syncwarp();
              dirty |= pose::update<pose::M_BODY_EFFORTS_VELOCITY_ONLY, KMAX>(w, lut, z, R, ctx, false, P, lane, st, a.uwv);
            }
          }
          if constexpr (KMAX >= 53) {
            if (kind == pose::M_PRESSURE) dirty |= pose::update<pose::M_PRESSURE, KMAX>(w, lut, z, R, op.aux, false, P, lane, st);
            else if (kind == pose::M_WATER_VELOCITY) dirty |= pose::update<pose::M_WATER_VELOCITY, KMAX>(w, lut, z, R, op.aux, true, P, lane, st);
            else if (efforts) {   // reference src/PoseUKF.cpp:581-602; the context is evaluated at the mean, once per filter
              double *ctx = w + W::F;
              pose::efforts_context(w, a.rot_rate + f * 3, P, ctx);
              __syncwarp();
              if (kind == pose::M_BODY_EFFORTS) dirty |= pose::update<pose::M_BODY_EFFORTS, KMAX>(w, lut, z, R, ctx, false, P, lane, st, a.uwv);
              else dirty |= pose::update<pose::M_BODY_EFFORTS_VELOCITY_ONLY, KMAX>(w, lut, z, R, ctx, false, P, lane, st, a.uwv);
            }
          }
          break;
      }
    }
    __syncwarp();
    if (dirty) {
      double2 *dst = reinterpret_cast<double2 *>(a.state + f * pose::BLK);
      const double2 *src = reinterpret_cast<const double2 *>(w);
#pragma unroll 4
      for (int k = lane; k < pose::BLK / 2; k += 32) __stcs(dst + k, src[k]);
    }
    if (st && lane == 0) a.status[f] |= st;
    __syncwarp();
  }
}

// Body-efforts update (measurementEfforts, reference src/PoseUKF.cpp:153-196, 581-602) in STREAMED form (pose::WLS): the
// covariance stays in HBM -- read by the factorisation, streamed through registers by the rank-6 downdate -- and only mu and the
// six leading columns live in shared memory: 8 warps per SM at 255 registers instead of 6.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) pose_efforts_kernel(pose::Launch a, const __grid_constant__ pose::Params P) {
  using W = pose::WLS;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *w = smem + warp * W::TOTAL;
  const pose::Op &op = a.ops[0];
  constexpr int M = 6;
  constexpr int NWB = pose::OFF_S + 303;          // mu (+ pad) and the six leading columns: what goes back from shared memory
  for (long long f = (long long)blockIdx.x * WARPS + warp; f < a.B; f += (long long)gridDim.x * WARPS) {
    double *blk = a.state + f * pose::BLK;
    __syncwarp();
    // the whole block into L2 now: the factorisation and the streamed downdate read it from there
    for (int o = lane * 128; o < pose::BLK * 8; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(blk) + o));
    for (int k = lane; k < pose::OFF_S; k += 32) w[k] = blk[k];
    __syncwarp();
    unsigned st = 0;
    bool go = !(op.valid && !op.valid[f]);
    double z[M], R[M * M];
    if (go) {
      bool bad = false;
#pragma unroll
      for (int k = 0; k < M; ++k) { z[k] = op.z[f * M + k]; bad |= isnan(z[k]); }
#pragma unroll
      for (int k = 0; k < M * M; ++k) { R[k] = op.R[k]; bad |= isnan(R[k]); }
      if (bad) { st |= pose::ST_NAN_MEAS; go = false; }
    }
    if (go) {
      double *ctx = w + W::F;   // the context is evaluated at the mean, once per filter (reference src/PoseUKF.cpp:585-598)
      pose::efforts_context(w, a.rot_rate + f * 3, P, ctx);
      __syncwarp();
      const bool dirty = pose::update<pose::M_BODY_EFFORTS, 53, W>(w, nullptr, z, R, ctx, false, P, lane, st, a.uwv, blk + pose::OFF_S);
      __syncwarp();
      if (dirty) {
        for (int k = lane; k < NWB; k += 32) __stcs(blk + k, w[k]);
      }
    }
    if (st && lane == 0) a.status[f] |= st;
  }
}

// host (mu[count][54], cov[count][53*53] or one shared cov) -> device blocks, and the Markov offsets
__global__ void pose_pack_kernel(const double *mu, const double *cov, long long cov_stride, double *state, double *offsets,
                                 long long first, long long count) {
  long long total = count * pose::BLK;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    long long f = g / pose::BLK;
    int k = (int)(g - f * pose::BLK);
    double v = 0.0;
    if (k < pose::DIM) {
      v = mu[f * pose::DIM + k];
      if (k >= pose::P_INERTIA && k < pose::P_WV) offsets[(first + f) * pose::NOFF + (k - pose::P_INERTIA)] = v;
      if (k == pose::P_RHO) offsets[(first + f) * pose::NOFF + 27] = v;
    } else if (k >= pose::OFF_S && k < pose::OFF_S + pose::NP) {
      unsigned ij = g_lut[k - pose::OFF_S];
      int i = ij & 0xff, j = ij >> 8;
      v = cov[f * cov_stride + i * pose::N + j];
    }
    state[(first + f) * pose::BLK + k] = v;
  }
}
__global__ void pose_unpack_kernel(const double *state, double *mu, double *cov, long long first, long long count) {
  if (mu) {
    long long total = count * pose::DIM;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
      long long f = g / pose::DIM;
      int k = (int)(g - f * pose::DIM);
      mu[g] = state[(first + f) * pose::BLK + k];
    }
  }
  if (cov) {
    long long total = count * pose::N * pose::N;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
      long long f = g / (pose::N * pose::N);
      int k = (int)(g - f * pose::N * pose::N);
      int i = k / pose::N, j = k - i * pose::N;
      cov[g] = state[(first + f) * pose::BLK + pose::OFF_S + pose::sidx(i, j)];
    }
  }
}
__global__ void pose_get_nav_kernel(const double *state, double *out, long long first, long long count) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < count * 10) { long long f = g / 10; int k = (int)(g - f * 10); out[g] = state[(first + f) * pose::BLK + k]; }
}
__global__ void pose_set_pose_kernel(double *state, const double *pose7, long long B) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < B * 7) { long long f = g / 7; int k = (int)(g - f * 7); state[f * pose::BLK + k] = pose7[g]; }
}
// getRotationRate(), reference src/PoseUKF.cpp:693-699
__global__ void pose_rotation_rate_kernel(const double *state, const double *rot_rate, double *out, long long B,
                                          const __grid_constant__ pose::Params P) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  const double *mu = state + f * pose::BLK;
  double lat = tmerc_nav_to_lat(P.tm, mu[0], mu[1]);
  Vec3 eb = qrot_inv(Quat{mu[3], mu[4], mu[5], mu[6]}, Vec3{kEarthW * cos(lat), 0.0, kEarthW * sin(lat)});
  out[f * 3 + 0] = rot_rate[f * 3 + 0] - mu[pose::P_BG + 0] - eb.x;
  out[f * 3 + 1] = rot_rate[f * 3 + 1] - mu[pose::P_BG + 1] - eb.y;
  out[f * 3 + 2] = rot_rate[f * 3 + 2] - mu[pose::P_BG + 2] - eb.z;
}

// device-math probe (uwvkf_debug_eval): one evaluation per thread, inputs in[n][8] -> out[n][4]
__global__ void debug_eval_kernel(int what, const double *in, double *out, int n, const __grid_constant__ pose::Params P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < n;
  const double *a = in + (size_t)(act ? i : 0) * 8;
  double o[4] = {0, 0, 0, 0};
  switch (what) {
    case UWVKF_DEBUG_TMERC_SINCOS_LAT: tmerc_nav_to_sincos_lat(P.tm, a[0], a[1], o[0], o[1]); o[2] = tmerc_nav_to_lat(P.tm, a[0], a[1]); break;
    case UWVKF_DEBUG_SO3_BOXPLUS: { Quat q = so3_boxplus(Quat{a[0], a[1], a[2], a[3]}, Vec3{a[4], a[5], a[6]}, a[7]); o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; break; }
    case UWVKF_DEBUG_SO3_BOXMINUS: { Vec3 v = so3_boxminus(Quat{a[0], a[1], a[2], a[3]}, Quat{a[4], a[5], a[6], a[7]}); o[0] = v.x; o[1] = v.y; o[2] = v.z; break; }
    case UWVKF_DEBUG_RSQRT: o[0] = fast_rsqrt3(a[0]); o[1] = fast_rsqrt(a[0]); o[2] = fast_rcp(a[0]); break;
    case UWVKF_DEBUG_WARP_SUM: o[0] = pose::wsum(act ? a[0] : 0.0); o[1] = warp_sum(act ? a[0] : 0.0); break;   // sums over each warp
  }
  if (act) for (int k = 0; k < 4; ++k) out[(size_t)i * 4 + k] = o[k];
}

// number of filters with each status bit set (metrics: non-SPD, NaN-refused, gate-rejected, mean-not-converged)
__global__ void status_count_kernel(const uint32_t *status, long long B, unsigned long long *out) {
  unsigned c[4] = {0, 0, 0, 0};
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (long long)gridDim.x * blockDim.x) {
    const uint32_t s = status[f];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] += (s >> k) & 1u;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned v = c[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + k, (unsigned long long)v);
  }
}

// ensemble statistics: out = [count, sum mu (DIM), sum mu mu^T (lower, row-major packed)]
template <int DIM>
__global__ void ensemble_stats_kernel(const double *state, long long stride, long long B, double *out) {
  constexpr int NOUT = 1 + DIM + DIM * (DIM + 1) / 2;
  constexpr int PER = (NOUT + 255) / 256;
  __shared__ double m[DIM];
  double acc[PER];
  int ia[PER], ib[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    acc[k] = 0.0;
    int o = threadIdx.x + 256 * k;
    ia[k] = -2; ib[k] = 0;
    if (o == 0) ia[k] = -1;
    else if (o <= DIM) { ia[k] = o - 1; ib[k] = -1; }
    else if (o < NOUT) {
      int e = o - 1 - DIM, i = 0;
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      ia[k] = i; ib[k] = e - i * (i + 1) / 2;
    }
  }
  for (long long f = blockIdx.x; f < B; f += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < DIM) m[threadIdx.x] = state[f * stride + threadIdx.x];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      if (ia[k] == -1) acc[k] += 1.0;
      else if (ia[k] >= 0) acc[k] += (ib[k] < 0) ? m[ia[k]] : m[ia[k]] * m[ib[k]];
    }
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int o = threadIdx.x + 256 * k;
    if (o < NOUT) atomicAdd(out + o, acc[k]);
  }
}

}  // namespace

namespace uwvkf {
int small_ensemble_stats(uwvkf_handle h, double *dout) {
  int grid = (int)std::min<int64_t>(h->B, (int64_t)h->sm_count * 8);
  const long long stride = h->dim + h->dof * h->dof;
  if (h->dim == 4) ensemble_stats_kernel<4><<<grid, 256, 0, h->stream>>>(h->d_state, stride, h->B, dout);
  else ensemble_stats_kernel<7><<<grid, 256, 0, h->stream>>>(h->d_state, stride, h->B, dout);
  return UWVKF_OK;
}
}  // namespace uwvkf

// ======================================================================================================
namespace {

#ifndef UWVKF_FUSED_MAX_WARPS
#define UWVKF_FUSED_MAX_WARPS 12
#endif
#ifndef UWVKF_UPDATE_WARPS
#define UWVKF_UPDATE_WARPS UWVKF_FUSED_MAX_WARPS
#endif
// 1: uwvkf_pose_predict_update2 runs (DVL, accelerometer) as ONE two-update kernel.  Measured on 262,144 filters: predict + DVL +
// accelerometer 5.00 ms against 2.49 + 1.43 ms for the fused predict + DVL kernel followed by the accelerometer kernel (the
// two-update kernel issues at 35 % instead of 47 %: 7,660 instructions per filter-step, spilled loop invariants, twice the
// instruction-cache misses); DVL + accelerometer without a predict 2.56 against 2.71 ms.  Off: the entry point launches twice.
#ifndef UWVKF_FUSED2
#define UWVKF_FUSED2 0
#endif
// 1: the full body-efforts update runs in pose_efforts_kernel (covariance streamed from HBM instead of resident in shared memory:
// 8 warps per SM instead of 6).  Parity green; 3.79 ms against 3.52 ms per 65,536 filters for the resident kernel -- the global
// loads of the factorisation and of the streamed downdate cost more than the two extra warps bring.  Off.
#ifndef UWVKF_EFFORTS_STREAMED
#define UWVKF_EFFORTS_STREAMED 0
#endif
#ifndef UWVKF_FUSED2_WARPS
#define UWVKF_FUSED2_WARPS 12
#endif
constexpr int kFused2Warps = UWVKF_FUSED2_WARPS;
constexpr int kFastWarps = 12, kSlowWarps = 6, kEffortsVelWarps = 10, kFusedMaxWarps = UWVKF_FUSED_MAX_WARPS, kUpdateMaxWarps = UWVKF_UPDATE_WARPS;
constexpr int kChunk = 4096;  // filters per staging chunk in set/get_state

}  // namespace
namespace uwvkf {
int ensure_tmp(uwvkf_handle h, size_t bytes) {
  if (h->tmp_bytes >= bytes) return UWVKF_OK;
  if (h->d_tmp) cudaFree(h->d_tmp);
  h->d_tmp = nullptr; h->tmp_bytes = 0;
  CK(cudaMalloc(&h->d_tmp, bytes));
  h->tmp_bytes = bytes;
  return UWVKF_OK;
}
}  // namespace uwvkf
namespace {

void fill_tmerc(TMercConst &tm, double lat0, double lon0) {
  const double a = 6378137.0, f = 1.0 / 298.257223563;
  const double n = f / (2 - f), n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
  tm.lat0 = lat0; tm.lon0 = lon0;
  tm.e = std::sqrt(f * (2 - f));
  tm.A = a / (1 + n) * (1 + n2 / 4 + n4 / 64 + n6 / 256);
  tm.Ainv = 1.0 / tm.A;
  double *al = tm.alpha, *be = tm.beta, *de = tm.delta;
  al[0] = n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800;
  al[1] = 13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360;
  al[2] = 61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440;
  al[3] = 49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600;
  al[4] = 34729 * n5 / 80640 - 3418889 * n6 / 1995840;
  al[5] = 212378941 * n6 / 319334400;
  be[0] = n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800;
  be[1] = n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720;
  be[2] = 17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720;
  be[3] = 4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600;
  be[4] = 4583 * n5 / 161280 - 108847 * n6 / 3991680;
  be[5] = 20648693 * n6 / 638668800;
  de[0] = 2 * n - 2 * n2 / 3 - 2 * n3 + 116 * n4 / 45 + 26 * n5 / 45 - 2854 * n6 / 675;
  de[1] = 7 * n2 / 3 - 8 * n3 / 5 - 227 * n4 / 45 + 2704 * n5 / 315 + 2323 * n6 / 945;
  de[2] = 56 * n3 / 15 - 136 * n4 / 35 - 1262 * n5 / 105 + 73814 * n6 / 2835;
  de[3] = 4279 * n4 / 630 - 332 * n5 / 35 - 399572 * n6 / 14175;
  de[4] = 4174 * n5 / 315 - 144838 * n6 / 6237;
  de[5] = 601676 * n6 / 22275;
  double tau = std::tan(lat0);
  double sig = std::sinh(tm.e * std::atanh(tm.e * tau / std::sqrt(1 + tau * tau)));
  double taup = tau * std::sqrt(1 + sig * sig) - sig * std::sqrt(1 + tau * tau);
  double chi0 = std::atan(taup), s = chi0;
  for (int j = 0; j < 6; ++j) s += al[j] * std::sin(2 * (j + 1) * chi0);
  tm.xi0 = s;
  tm.sin_xi0 = std::sin(s);
  tm.cos_xi0 = std::cos(s);
}

int meas_dim(int kind) {
  switch (kind) {
    case UWVKF_POSE_XY: case UWVKF_POSE_WATER_VELOCITY: case UWVKF_POSE_GEOGRAPHIC: case UWVKF_POSE_DELAYED_XY: return 2;
    case UWVKF_POSE_Z: case UWVKF_POSE_PRESSURE: return 1;
    case UWVKF_POSE_ACCELERATION: case UWVKF_POSE_VELOCITY: return 3;
    case UWVKF_POSE_BODY_EFFORTS: case UWVKF_POSE_BODY_EFFORTS_VELOCITY_ONLY: return 6;
  }
  return -1;
}

template <int KMAX, int WARPS>
int launch_pose(uwvkf_handle h, const pose::Launch &L) {
  using W = pose::WL<KMAX>;
  size_t smem = (size_t)(1432 + 360 + WARPS * W::TOTAL) * sizeof(double);
  static unsigned long long attr_set = 0;   // one bit per device: the attribute belongs to the device's context
  if (!(attr_set >> h->device & 1ull)) {
    CK(cudaFuncSetAttribute(pose_step_kernel<KMAX, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set |= 1ull << h->device;
  }
  long long need = (L.B + WARPS - 1) / WARPS;
  int grid = (int)std::min<long long>(need, h->sm_count);
  pose_step_kernel<KMAX, WARPS><<<grid, WARPS * 32, smem, h->stream>>>(L, h->pp);
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

// v2 fused kernel (pose_fused.cuh): WARPS per CTA is chosen per instantiation from its shared-memory footprint
template <int HP, int KIND, int KIND2 = pose::M_NONE>
constexpr int fused_warps() {
  using C = pose::FusedCfg<HP, KIND, KIND2>;
  int w = (232448 - C::CTA * 8) / (C::TOTAL * 8);
  // update-only kernels carry far less register state; two-update kernels more (UWVKF_FUSED2_WARPS)
  const int cap = KIND2 != pose::M_NONE ? kFused2Warps : (HP ? kFusedMaxWarps : kUpdateMaxWarps);
  return w > cap ? cap : w;
}
template <int HP, int KIND, int KIND2 = pose::M_NONE>
int launch_fused(uwvkf_handle h, const pose::Launch &L) {
  using C = pose::FusedCfg<HP, KIND, KIND2>;
  constexpr int WARPS = fused_warps<HP, KIND, KIND2>();
  size_t smem = (size_t)(C::CTA + WARPS * C::TOTAL) * sizeof(double);
  static unsigned long long attr_set = 0;   // one bit per device
  if (!(attr_set >> h->device & 1ull)) {
    CK(cudaFuncSetAttribute(pose::pose_fused_kernel<HP, KIND, WARPS, KIND2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set |= 1ull << h->device;
  }
  long long need = (L.B + WARPS - 1) / WARPS;
  int grid = (int)std::min<long long>(need, h->sm_count);
  if constexpr (C::TMP) {
    // latitudes of the predict's sigma points, thread per filter, on the same stream right ahead of the fused kernel
    pose::pose_tmerc_prepass_kernel<<<(unsigned)((L.B + 255) / 256), 256, 0, h->stream>>>(L.state, h->d_tm, L.B, h->pp);
    CK(cudaGetLastError());
    h->launches++;
  }
  pose::pose_fused_kernel<HP, KIND, WARPS, KIND2><<<grid, WARPS * 32, smem, h->stream>>>(L, h->pp);
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}
template <int HP>
int launch_fused_kind(uwvkf_handle h, const pose::Launch &L, int kind) {
  switch (kind) {
    case UWVKF_POSE_XY: case UWVKF_POSE_GEOGRAPHIC: case UWVKF_POSE_DELAYED_XY: return launch_fused<HP, pose::M_XY>(h, L);
    case UWVKF_POSE_Z: return launch_fused<HP, pose::M_Z>(h, L);
    case UWVKF_POSE_VELOCITY: return launch_fused<HP, pose::M_VELOCITY>(h, L);
    case UWVKF_POSE_ACCELERATION: return launch_fused<HP, pose::M_ACCELERATION>(h, L);
    case UWVKF_POSE_PRESSURE: return launch_fused<HP, pose::M_PRESSURE>(h, L);
    case UWVKF_POSE_WATER_VELOCITY: return launch_fused<HP, pose::M_WATER_VELOCITY>(h, L);
  }
  return fail(UWVKF_ERR_INVALID_ARGUMENT, "measurement kind not handled by the fused kernel");
}
bool engine_v1() {
  static const bool v1 = [] { const char *e = std::getenv("UWVKF_ENGINE"); return e && std::string(e) == "v1"; }();
  return v1;
}
int run_pose(uwvkf_handle h, int nops, const pose::Op *ops) {
  if (!h->have_params) return fail(UWVKF_ERR_NOT_INITIALIZED, "uwvkf_pose_set_params was not called");
  pose::Launch L{};
  L.state = h->d_state; L.offsets = h->d_offsets; L.rot_rate = h->d_rot; L.status = h->d_status; L.Q = h->d_Q;
  L.uwv = h->d_uwv;
  L.nav = h->nav_fused;
  L.tm = h->d_tm;
  L.B = h->B; L.nops = nops;
  bool slow = false, dense_only = false, full_efforts = false;
  for (int o = 0; o < nops; ++o) {
    L.ops[o] = ops[o];
    if (ops[o].type == 0 && !h->have_Q) return fail(UWVKF_ERR_NOT_INITIALIZED, "uwvkf_set_process_noise was not called");
    if (ops[o].type == 1 && (ops[o].kind == UWVKF_POSE_PRESSURE || ops[o].kind == UWVKF_POSE_WATER_VELOCITY)) slow = true;
    if (ops[o].type == 1 && (ops[o].kind == UWVKF_POSE_BODY_EFFORTS || ops[o].kind == UWVKF_POSE_BODY_EFFORTS_VELOCITY_ONLY)) {
      if (!h->d_uwv) return fail(UWVKF_ERR_NOT_INITIALIZED, "uwvkf_pose_set_model_parameters was not called");
      slow = dense_only = true;
      full_efforts |= ops[o].kind == UWVKF_POSE_BODY_EFFORTS;
    }
  }
  // engine selection: "panel + stream" kernel for predict / update / predict+update (pressure and ADCP through their
  // affine tail); the resident kernel (dense prefix 53 for pressure / ADCP) for op sequences the fused kernel has no
  // instantiation for and, with UWVKF_ENGINE=v1, for everything (cross-checks)
  const bool force_v1 = engine_v1();
  if (!force_v1 && !dense_only) {
    if (nops == 1 && ops[0].type == 0) return launch_fused<1, pose::M_NONE>(h, L);
    if (nops == 1 && ops[0].type == 1) return launch_fused_kind<0>(h, L, ops[0].kind);
    if (nops == 2 && ops[0].type == 0 && ops[1].type == 1) return launch_fused_kind<1>(h, L, ops[1].kind);
#if UWVKF_FUSED2
    // the 100 Hz step of a full PoseUKF (BASELINE config 3): predict + DVL + accelerometer in ONE pass over the filter
    if (nops == 3 && ops[0].type == 0 && ops[1].type == 1 && ops[2].type == 1 && ops[1].kind == UWVKF_POSE_VELOCITY &&
        ops[2].kind == UWVKF_POSE_ACCELERATION)
      return launch_fused<1, pose::M_VELOCITY, pose::M_ACCELERATION>(h, L);
    if (nops == 2 && ops[0].type == 1 && ops[1].type == 1 && ops[0].kind == UWVKF_POSE_VELOCITY && ops[1].kind == UWVKF_POSE_ACCELERATION)
      return launch_fused<0, pose::M_VELOCITY, pose::M_ACCELERATION>(h, L);
#endif
  }
#if UWVKF_EFFORTS_STREAMED
  // the full body-efforts update: streamed dense kernel (covariance in HBM, 8 warps per SM)
  if (!force_v1 && dense_only && full_efforts && nops == 1 && !ops[0].Rdev) {
    constexpr int WARPS = 8;
    const size_t smem = (size_t)WARPS * pose::WLS::TOTAL * sizeof(double);
    static unsigned long long attr_set = 0;   // one bit per device
    if (!(attr_set >> h->device & 1ull)) {
      CK(cudaFuncSetAttribute(pose_efforts_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set |= 1ull << h->device;
    }
    const int grid = (int)std::min<long long>((L.B + WARPS - 1) / WARPS, h->sm_count);
    pose_efforts_kernel<WARPS><<<grid, WARPS * 32, smem, h->stream>>>(L, h->pp);
    CK(cudaGetLastError());
    h->launches++;
    return UWVKF_OK;
  }
#endif
  // the velocity-only body-efforts update reads prefix 9 only: prefix-19 engine with six-row tables, 10 warps per SM
  if (!force_v1 && dense_only && !full_efforts && nops == 1) return launch_pose<20, kEffortsVelWarps>(h, L);
  return slow ? launch_pose<53, kSlowWarps>(h, L) : launch_pose<19, kFastWarps>(h, L);
}

}  // namespace
namespace uwvkf {
// stage a host/device input array on the device; returns the device pointer to use
int stage(uwvkf_handle h, const void *src, size_t bytes, size_t cap, int memspace, void **slot, const void **out) {
  if (!src) { *out = nullptr; return UWVKF_OK; }
  if (memspace == UWVKF_DEVICE) { *out = src; return UWVKF_OK; }
  if (!*slot) CK(cudaMalloc(slot, cap));
  CK(cudaMemcpyAsync(*slot, src, bytes, cudaMemcpyHostToDevice, h->stream));
  *out = *slot;
  return UWVKF_OK;
}
}  // namespace uwvkf
namespace {

int make_update_op(uwvkf_handle h, int kind, const double *z, const double *R, int R_per_filter, const double *aux,
                   const uint8_t *valid, int memspace, pose::Op &op, int slot = 0) {
  const int m = meas_dim(kind);
  if (m < 0) return fail(UWVKF_ERR_INVALID_ARGUMENT, "unknown PoseUKF measurement kind");
  if (!z || !R) return fail(UWVKF_ERR_INVALID_ARGUMENT, "z and R must not be NULL");
  std::memset(&op, 0, sizeof op);
  op.type = 1; op.kind = kind;
  // checkMeasurment(): NaN in mu or cov throws in the reference.  HOST inputs are scanned here whatever the batch size;
  // DEVICE inputs cannot be read by the host and are checked per filter in the kernel (UWVKF_ST_NAN_MEASUREMENT).
  if (memspace == UWVKF_HOST && any_nan(z, h->B * m)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "measurement contains NaN");
  if (!R_per_filter || memspace == UWVKF_HOST) {
    if (any_nan(R, R_per_filter ? h->B * m * m : m * m)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "measurement covariance contains NaN");
  }
  const void *p;
  // (slot 1: the second update of a two-update call has its own staging buffers)
  int rc = stage(h, z, (size_t)h->B * m * sizeof(double), (size_t)h->B * pose::MAXM * sizeof(double), memspace,
                 (void **)(slot ? &h->d_z2 : &h->d_z), &p);
  if (rc) return rc;
  op.z = (const double *)p;
  if (R_per_filter) {
    if (slot) return fail(UWVKF_ERR_UNSUPPORTED, "per-filter R is not available for the second update of a fused call");
    if (m > 3) return fail(UWVKF_ERR_UNSUPPORTED, "per-filter R is supported for measurement dimensions <= 3");
    rc = stage(h, R, (size_t)h->B * m * m * sizeof(double), (size_t)h->B * 9 * sizeof(double), memspace, (void **)&h->d_Rpf, &p);
    if (rc) return rc;
    op.Rdev = (const double *)p;
  } else {
    for (int k = 0; k < m * m; ++k) op.R[k] = R[k];
  }
  rc = stage(h, valid, (size_t)h->B, (size_t)h->B, memspace, (void **)(slot ? &h->d_valid2 : &h->d_valid), &p);
  if (rc) return rc;
  op.valid = (const uint8_t *)p;
  if (aux) {
    int na = (kind == UWVKF_POSE_WATER_VELOCITY) ? 1 : (kind == UWVKF_POSE_DELAYED_XY ? 2 : 3);
    for (int k = 0; k < na; ++k) { if (std::isnan(aux[k])) return fail(UWVKF_ERR_NAN_MEASUREMENT, "aux contains NaN"); op.aux[k] = aux[k]; }
  } else if (kind == UWVKF_POSE_WATER_VELOCITY || kind == UWVKF_POSE_DELAYED_XY) {
    return fail(UWVKF_ERR_INVALID_ARGUMENT, "this measurement kind needs aux");
  }
  return UWVKF_OK;
}

}  // namespace

extern "C" {

const char *uwvkf_last_error(void) { return uwvkf::g_err.c_str(); }
const char *uwvkf_version(void) { return "uwvkf 0.1 (sm_100a)"; }

int uwvkf_create(int filter_kind, int64_t batch, int device, uwvkf_handle *out) {
  if (!out || batch <= 0) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad arguments");
  if (filter_kind < UWVKF_FILTER_POSE || filter_kind > UWVKF_FILTER_INDIRECT_POSE) return fail(UWVKF_ERR_INVALID_ARGUMENT, "unknown filter kind");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(UWVKF_ERR_NO_DEVICE, "no CUDA device: uwvkf has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad device index");
  CK(cudaSetDevice(device));
  uwvkf_handle h = new uwvkf_filter_s();
  h->kind = filter_kind; h->B = batch; h->device = device;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (filter_kind != UWVKF_FILTER_POSE) {
    int rc = small_create(h);
    if (rc == UWVKF_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(UWVKF_ERR_CUDA, "cudaDeviceSynchronize failed after allocation");
    if (rc) { uwvkf_destroy(h); return rc; }
    *out = h;
    return UWVKF_OK;
  }
  std::vector<uint16_t> lut(pose::NP);
  for (int j = 0; j < pose::N; ++j) for (int i = j; i < pose::N; ++i) lut[pose::cidx(i, j)] = (uint16_t)(i | (j << 8));
  cudaError_t ce = cudaMemcpyToSymbol(g_lut, lut.data(), lut.size() * sizeof(uint16_t));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_state, (size_t)batch * pose::BLK * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_offsets, (size_t)batch * pose::NOFF * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_rot, (size_t)batch * 3 * sizeof(double));
#if UWVKF_TMERC_PREPASS
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_tm, (size_t)batch * 10 * sizeof(double));   // written before it is read, every predict
#endif
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_status, (size_t)batch * sizeof(uint32_t));
  if (ce == cudaSuccess) ce = cudaMalloc(&h->d_Q, (size_t)pose::NP * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_state, 0, (size_t)batch * pose::BLK * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_offsets, 0, (size_t)batch * pose::NOFF * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_rot, 0, (size_t)batch * 3 * sizeof(double));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_status, 0, (size_t)batch * sizeof(uint32_t));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_Q, 0, (size_t)pose::NP * sizeof(double));
  if (ce != cudaSuccess) {
    std::string msg = cudaGetErrorString(ce);
    uwvkf_destroy(h);
    return fail(UWVKF_ERR_CUDA, "allocation failed: " + msg);
  }
  // the memsets above ran on the legacy default stream and are asynchronous to the host; user streams (uwvkf_set_stream,
  // typically cudaStreamNonBlocking) do not order behind it, so they must have finished before the handle is handed out
  CK(cudaDeviceSynchronize());
  h->pp.mean_tol = h->mean_tol;
  h->pp.mean_max_it = h->mean_max_it;
  *out = h;
  return UWVKF_OK;
}

static void destroy_pipe(uwvkf_handle h) {
  auto *p = h->pipe;
  if (!p) return;
  if (p->s_in) cudaStreamSynchronize(p->s_in);
  if (p->s_out) cudaStreamSynchronize(p->s_out);
  for (int k = 0; k < 2; ++k) {
    cudaFree(p->d_w[k]); cudaFree(p->d_z[k]); cudaFree(p->d_nav[k]);
    if (p->ev_in[k]) cudaEventDestroy(p->ev_in[k]);
    if (p->ev_main[k]) cudaEventDestroy(p->ev_main[k]);
    if (p->ev_out[k]) cudaEventDestroy(p->ev_out[k]);
  }
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
  h->pipe = nullptr;
}

int uwvkf_destroy(uwvkf_handle h) {
  if (!h) return UWVKF_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_state); cudaFree(h->d_offsets); cudaFree(h->d_rot); cudaFree(h->d_tm); cudaFree(h->d_status); cudaFree(h->d_Q);
  cudaFree(h->d_z); cudaFree(h->d_Rpf); cudaFree(h->d_valid); cudaFree(h->d_tmp); cudaFree(h->d_uwv);
  cudaFree(h->d_z2); cudaFree(h->d_valid2);
  destroy_pipe(h);
  small_destroy(h);
  delete h;
  return UWVKF_OK;
}

int uwvkf_set_stream(uwvkf_handle h, void *s) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  if (h->capturing) return fail(UWVKF_ERR_INVALID_ARGUMENT, "cannot switch streams during a graph capture");
  CK(cudaSetDevice(h->device));
  // staged inputs (d_z, d_tmp, ...) and kernels may still be in flight on the old stream: drain it before the switch (and
  // the copy streams of the host-buffer pipeline, whose events refer to work on the old stream)
  CK(cudaStreamSynchronize(h->stream));
  if (h->pipe) {
    CK(cudaStreamSynchronize(h->pipe->s_in));
    CK(cudaStreamSynchronize(h->pipe->s_out));
    h->pipe->out_pending[0] = h->pipe->out_pending[1] = false;
  }
  h->stream = (cudaStream_t)s;
  return UWVKF_OK;
}
int uwvkf_synchronize(uwvkf_handle h) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->pipe) {   // downloads of uwvkf_pose_step_host run behind the kernels, on their own stream
    CK(cudaStreamSynchronize(h->pipe->s_out));
    h->pipe->out_pending[0] = h->pipe->out_pending[1] = false;
  }
  return UWVKF_OK;
}
int uwvkf_join(uwvkf_handle h) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  if (h->pipe) {
    for (int k = 0; k < 2; ++k) {
      if (h->pipe->out_pending[k]) CK(cudaStreamWaitEvent(h->stream, h->pipe->ev_out[k], 0));
      if (h->pipe->calls > k) CK(cudaStreamWaitEvent(h->stream, h->pipe->ev_in[k], 0));
    }
  }
  return UWVKF_OK;
}
int64_t uwvkf_batch(uwvkf_handle h) { return h ? h->B : 0; }
int64_t uwvkf_kernel_launches(uwvkf_handle h) { return h ? h->launches : 0; }

int uwvkf_set_mean_policy(uwvkf_handle h, double tol, int max_it) {
  if (!h || max_it < 1 || !(tol >= 0)) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad policy");
  h->pp.mean_tol = h->mean_tol = tol;
  h->pp.mean_max_it = h->mean_max_it = max_it;
  return UWVKF_OK;
}

int uwvkf_set_state(uwvkf_handle h, const double *mu, const double *cov, int64_t cov_stride, int64_t first, int64_t count,
                    int memspace) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  if (!mu || !cov || first < 0 || count <= 0 || first + count > h->B) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad range or null pointer");
  if (h->kind != UWVKF_FILTER_POSE) return small_set_state(h, mu, cov, cov_stride, first, count, memspace);
  if (cov_stride != 0 && cov_stride != pose::N * pose::N) return fail(UWVKF_ERR_INVALID_ARGUMENT, "cov_stride must be 0 (shared) or DOF*DOF");
  const size_t cb = (size_t)pose::N * pose::N * sizeof(double), mb = (size_t)pose::DIM * sizeof(double);
  if (memspace == UWVKF_DEVICE) {
    pose_pack_kernel<<<1024, 256, 0, h->stream>>>(mu, cov, cov_stride, h->d_state, h->d_offsets, first, count);
    CK(cudaGetLastError());
    h->launches++;
    return UWVKF_OK;
  }
  int64_t chunk = std::min<int64_t>(count, kChunk);
  size_t need = (size_t)chunk * mb + (cov_stride ? (size_t)chunk * cb : cb);
  int rc = ensure_tmp(h, need);
  if (rc) return rc;
  double *dmu = h->d_tmp, *dcov = h->d_tmp + (size_t)chunk * pose::DIM;
  if (!cov_stride) CK(cudaMemcpyAsync(dcov, cov, cb, cudaMemcpyHostToDevice, h->stream));
  for (int64_t o = 0; o < count; o += chunk) {
    int64_t c = std::min<int64_t>(chunk, count - o);
    CK(cudaMemcpyAsync(dmu, mu + o * pose::DIM, (size_t)c * mb, cudaMemcpyHostToDevice, h->stream));
    if (cov_stride) CK(cudaMemcpyAsync(dcov, cov + o * cov_stride, (size_t)c * cb, cudaMemcpyHostToDevice, h->stream));
    pose_pack_kernel<<<1024, 256, 0, h->stream>>>(dmu, dcov, cov_stride, h->d_state, h->d_offsets, first + o, c);
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));  // staging buffer is reused by the next chunk
  }
  return UWVKF_OK;
}

int uwvkf_get_state(uwvkf_handle h, double *mu, double *cov, int64_t first, int64_t count, int memspace) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  if ((!mu && !cov) || first < 0 || count <= 0 || first + count > h->B) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad range or null pointer");
  if (h->kind != UWVKF_FILTER_POSE) return small_get_state(h, mu, cov, first, count, memspace);
  const size_t cb = (size_t)pose::N * pose::N * sizeof(double), mb = (size_t)pose::DIM * sizeof(double);
  if (memspace == UWVKF_DEVICE) {
    pose_unpack_kernel<<<1024, 256, 0, h->stream>>>(h->d_state, mu, cov, first, count);
    CK(cudaGetLastError());
    h->launches++;
    return UWVKF_OK;
  }
  int64_t chunk = std::min<int64_t>(count, cov ? kChunk : (int64_t)1 << 20);
  int rc = ensure_tmp(h, (size_t)chunk * (mb + (cov ? cb : 0)));
  if (rc) return rc;
  double *dmu = h->d_tmp, *dcov = cov ? h->d_tmp + (size_t)chunk * pose::DIM : nullptr;
  for (int64_t o = 0; o < count; o += chunk) {
    int64_t c = std::min<int64_t>(chunk, count - o);
    pose_unpack_kernel<<<1024, 256, 0, h->stream>>>(h->d_state, mu ? dmu : nullptr, dcov, first + o, c);
    CK(cudaGetLastError());
    h->launches++;
    if (mu) CK(cudaMemcpyAsync(mu + o * pose::DIM, dmu, (size_t)c * mb, cudaMemcpyDeviceToHost, h->stream));
    if (cov) CK(cudaMemcpyAsync(cov + o * pose::N * pose::N, dcov, (size_t)c * cb, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return UWVKF_OK;
}

int uwvkf_set_process_noise(uwvkf_handle h, const double *Q) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  if (!Q) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null Q");
  if (h->kind != UWVKF_FILTER_POSE) {
    for (int k = 0; k < h->dof * h->dof; ++k) h->Qh[k] = Q[k];
    return UWVKF_OK;
  }
  std::vector<double> packed(pose::NP);
  for (int j = 0; j < pose::N; ++j) for (int i = j; i < pose::N; ++i) packed[pose::cidx(i, j)] = Q[i * pose::N + j];
  CK(cudaMemcpyAsync(h->d_Q, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_Q = true;
  return UWVKF_OK;
}

int uwvkf_get_status(uwvkf_handle h, uint32_t *status, int64_t first, int64_t count) {
  if (!h || !status || first < 0 || count <= 0 || first + count > h->B) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad arguments");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(status, h->d_status + first, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}
int uwvkf_status_counts(uwvkf_handle h, int64_t *counts) {
  if (!h || !counts) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad arguments");
  CK(cudaSetDevice(h->device));
  int rc = ensure_tmp(h, 4 * sizeof(unsigned long long));
  if (rc) return rc;
  unsigned long long *d = reinterpret_cast<unsigned long long *>(h->d_tmp), hc[4];
  CK(cudaMemsetAsync(d, 0, sizeof hc, h->stream));
  status_count_kernel<<<std::max(1, std::min(h->sm_count * 4, (int)((h->B + 255) / 256))), 256, 0, h->stream>>>(h->d_status, h->B, d);
  CK(cudaGetLastError());
  h->launches++;
  CK(cudaMemcpyAsync(hc, d, sizeof hc, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < 4; ++k) counts[k] = (int64_t)hc[k];
  return UWVKF_OK;
}
int uwvkf_clear_status(uwvkf_handle h) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_status, 0, (size_t)h->B * sizeof(uint32_t), h->stream));
  return UWVKF_OK;
}

int uwvkf_pose_set_params(uwvkf_handle h, const double *p) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!p) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null params");
  pose::Params &P = h->pp;
  fill_tmerc(P.tm, p[UWVKF_PP_LAT0], p[UWVKF_PP_LON0]);
  P.gyro_bias_tau = p[UWVKF_PP_GYRO_BIAS_TAU];
  P.acc_bias_tau = p[UWVKF_PP_ACC_BIAS_TAU];
  for (int k = 0; k < 3; ++k) {
    P.gyro_bias_offset[k] = p[UWVKF_PP_GYRO_BIAS_OFFSET + k];
    P.acc_bias_offset[k] = p[UWVKF_PP_ACC_BIAS_OFFSET + k];
    P.imu_in_body[k] = p[UWVKF_PP_IMU_IN_BODY + k];
  }
  P.inertia_tau = p[UWVKF_PP_INERTIA_TAU]; P.lin_damping_tau = p[UWVKF_PP_LIN_DAMPING_TAU];
  P.quad_damping_tau = p[UWVKF_PP_QUAD_DAMPING_TAU]; P.wv_tau = p[UWVKF_PP_WV_TAU]; P.wv_limits = p[UWVKF_PP_WV_LIMITS];
  P.wv_scale = p[UWVKF_PP_WV_SCALE]; P.adcp_bias_tau = p[UWVKF_PP_ADCP_BIAS_TAU];
  P.atm_pressure = p[UWVKF_PP_ATM_PRESSURE]; P.rho_tau = p[UWVKF_PP_WATER_DENSITY_TAU];
  h->have_params = true;
  return UWVKF_OK;
}

int uwvkf_pose_set_model_parameters(uwvkf_handle h, const double *uwv) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!uwv) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null model parameters");
  for (int k = 0; k < pose::NUWV; ++k) if (std::isnan(uwv[k])) return fail(UWVKF_ERR_INVALID_ARGUMENT, "model parameters contain NaN");
  if (!h->d_uwv) CK(cudaMalloc(&h->d_uwv, pose::NUWV * sizeof(double)));
  CK(cudaMemcpyAsync(h->d_uwv, uwv, pose::NUWV * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

int uwvkf_pose_set_offsets(uwvkf_handle h, const double *offsets, int64_t first, int64_t count) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!offsets || first < 0 || count <= 0 || first + count > h->B) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad arguments");
  CK(cudaMemcpyAsync(h->d_offsets + first * pose::NOFF, offsets, (size_t)count * pose::NOFF * sizeof(double),
                     cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

int uwvkf_pose_set_rotation_rate(uwvkf_handle h, const double *w, int memspace) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!w) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null rotation rate");
  // HOST: scanned (the reference throws, src/PoseUKF.cpp:494); DEVICE: the predict kernel refuses a NaN rotation rate per
  // filter (UWVKF_ST_NAN_MEASUREMENT, prediction skipped)
  if (memspace == UWVKF_HOST && any_nan(w, h->B * 3)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "rotation rate contains NaN");
  CK(cudaMemcpyAsync(h->d_rot, w, (size_t)h->B * 3 * sizeof(double),
                     memspace == UWVKF_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  return UWVKF_OK;
}

int uwvkf_pose_get_rotation_rate(uwvkf_handle h, double *out) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!out) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null output");
  if (!h->have_params) return fail(UWVKF_ERR_NOT_INITIALIZED, "uwvkf_pose_set_params was not called");
  int rc = ensure_tmp(h, (size_t)h->B * 3 * sizeof(double));
  if (rc) return rc;
  pose_rotation_rate_kernel<<<(unsigned)((h->B + 127) / 128), 128, 0, h->stream>>>(h->d_state, h->d_rot, h->d_tmp, h->B, h->pp);
  CK(cudaGetLastError());
  h->launches++;
  CK(cudaMemcpyAsync(out, h->d_tmp, (size_t)h->B * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

int uwvkf_predict(uwvkf_handle h, double dt) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  if (std::isnan(dt)) return fail(UWVKF_ERR_INVALID_ARGUMENT, "dt is NaN");
  if (h->kind != UWVKF_FILTER_POSE) return small_predict(h, dt);
  pose::Op op{};
  op.type = 0; op.dt = dt;
  return run_pose(h, 1, &op);
}

int uwvkf_pose_update(uwvkf_handle h, int meas_kind, const double *z, const double *R, int R_per_filter, const double *aux,
                      const uint8_t *valid, int memspace) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  pose::Op op;
  int rc = make_update_op(h, meas_kind, z, R, R_per_filter, aux, valid, memspace, op);
  if (rc) return rc;
  return run_pose(h, 1, &op);
}

int uwvkf_pose_predict_update(uwvkf_handle h, double dt, const double *w, int meas_kind, const double *z, const double *R,
                              int R_per_filter, const double *aux, const uint8_t *valid, int memspace) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (std::isnan(dt)) return fail(UWVKF_ERR_INVALID_ARGUMENT, "dt is NaN");
  if (w) {
    int rc = uwvkf_pose_set_rotation_rate(h, w, memspace);
    if (rc) return rc;
  }
  pose::Op ops[2];
  std::memset(&ops[0], 0, sizeof ops[0]);
  ops[0].type = 0; ops[0].dt = dt;
  int rc = make_update_op(h, meas_kind, z, R, R_per_filter, aux, valid, memspace, ops[1]);
  if (rc) return rc;
  return run_pose(h, 2, ops);
}

int uwvkf_pose_predict_update2(uwvkf_handle h, double dt, const double *w, int kind_a, const double *z_a, const double *R_a,
                               const uint8_t *valid_a, int kind_b, const double *z_b, const double *R_b, const uint8_t *valid_b,
                               int memspace) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (std::isnan(dt) || dt < 0.0) return fail(UWVKF_ERR_INVALID_ARGUMENT, "dt must be >= 0");
  if (w) {
    int rc = uwvkf_pose_set_rotation_rate(h, w, memspace);
    if (rc) return rc;
  }
  pose::Op ops[3];
  std::memset(&ops[0], 0, sizeof ops[0]);
  ops[0].type = 0; ops[0].dt = dt;
  int rc = make_update_op(h, kind_a, z_a, R_a, 0, nullptr, valid_a, memspace, ops[1], 0);
  if (rc) return rc;
  rc = make_update_op(h, kind_b, z_b, R_b, 0, nullptr, valid_b, memspace, ops[2], 1);
  if (rc) return rc;
  const bool pair_fused = UWVKF_FUSED2 && kind_a == UWVKF_POSE_VELOCITY && kind_b == UWVKF_POSE_ACCELERATION;
  if (pair_fused) return dt > 0.0 ? run_pose(h, 3, ops) : run_pose(h, 2, ops + 1);
  // any other pair: the same result from two launches (fused predict + first update, then the second update)
  rc = dt > 0.0 ? run_pose(h, 2, ops) : run_pose(h, 1, ops + 1);
  if (rc) return rc;
  return run_pose(h, 1, ops + 2);
}

int uwvkf_pose_update_features(uwvkf_handle h, int nfeat, const double *px, const double *pxcov, const double *feature_pos,
                               const double *marker_pose, const double *cov_marker, const double *camera, const double *cam_in_imu,
                               const uint8_t *valid, int memspace) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (nfeat < 1 || nfeat > pose::MAXFEAT) return fail(UWVKF_ERR_INVALID_ARGUMENT, "1 <= nfeat <= 8");
  if (!px || !pxcov || !feature_pos || !marker_pose || !cov_marker || !camera || !cam_in_imu) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null argument");
  if (!h->have_params) return fail(UWVKF_ERR_NOT_INITIALIZED, "uwvkf_pose_set_params was not called");
  pose::MarkerArgs M{};
  M.nfeat = nfeat;
  for (int k = 0; k < nfeat * 4; ++k) M.pxcov[k / 4][k % 4] = pxcov[k];
  for (int k = 0; k < nfeat * 3; ++k) M.feat[k / 3][k % 3] = feature_pos[k];
  for (int k = 0; k < 7; ++k) { M.marker_pose[k] = marker_pose[k]; M.cam_in_imu[k] = cam_in_imu[k]; }
  for (int k = 0; k < 36; ++k) M.cov_marker[k] = cov_marker[k];
  for (int k = 0; k < 4; ++k) M.camera[k] = camera[k];
  if (memspace == UWVKF_HOST && any_nan(px, h->B * nfeat * 2))   // checkMeasurment() throwing (src/PoseUKF.cpp:635)
    return fail(UWVKF_ERR_NAN_MEASUREMENT, "marker corner contains NaN");
  for (int k = 0; k < nfeat * 4; ++k) if (std::isnan(pxcov[k])) return fail(UWVKF_ERR_NAN_MEASUREMENT, "marker corner covariance contains NaN");
  const double *dpx = px;
  if (memspace == UWVKF_HOST) {
    const size_t bytes = (size_t)h->B * nfeat * 2 * sizeof(double);
    int rc = ensure_tmp(h, bytes);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_tmp, px, bytes, cudaMemcpyHostToDevice, h->stream));
    dpx = h->d_tmp;
  }
  const void *pv;
  int rc = stage(h, valid, (size_t)h->B, (size_t)h->B, memspace, (void **)&h->d_valid, &pv);
  if (rc) return rc;
  constexpr int WARPS = 6;
  const size_t smem = (size_t)WARPS * pose::MarkerLayout::TOTAL * sizeof(double);
  static unsigned long long attr_set = 0;   // one bit per device
  if (!(attr_set >> h->device & 1ull)) {
    CK(cudaFuncSetAttribute(pose::pose_marker_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set |= 1ull << h->device;
  }
  const int grid = (int)std::min<long long>((h->B + WARPS - 1) / WARPS, h->sm_count);
  pose::pose_marker_kernel<WARPS><<<grid, WARPS * 32, smem, h->stream>>>(h->d_state, h->d_status, dpx, (const uint8_t *)pv, h->B, M, h->pp);
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

/* ---- sample stream (the caller-side loop) ------------------------------------------------------------------ */
struct uwvkf_stream_s {
  uwvkf_handle h = nullptr;
  int slots = 0, head = 0, count = 0;        // ring: [head, head + count) are queued and not yet replayed
  cudaStream_t copy = nullptr;
  double t_filter = 0.0;
  bool have_time = false;
  struct Slot {
    double ts = 0.0;
    int kind = 0;
    double R[pose::MAXM * pose::MAXM] = {0}, aux[4] = {0};
    bool has_aux = false, has_valid = false, used = false;
    double *h_z = nullptr, *d_z = nullptr;   // pinned staging / HBM, [B][MAXM]
    uint8_t *h_v = nullptr, *d_v = nullptr;  // [B]
    cudaEvent_t copied = nullptr, consumed = nullptr;
  };
  std::vector<Slot> ring;
};

int uwvkf_stream_create(uwvkf_handle h, int slots, uwvkf_stream *out) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!out || slots < 2 || slots > 4096) return fail(UWVKF_ERR_INVALID_ARGUMENT, "2 <= slots <= 4096");
  uwvkf_stream s = new uwvkf_stream_s();
  s->h = h; s->slots = slots; s->ring.resize(slots);
  cudaError_t e = cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking);
  const size_t zb = (size_t)h->B * pose::MAXM * sizeof(double);
  for (auto &sl : s->ring) {
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_z, zb);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_z, zb);
    if (e == cudaSuccess) e = cudaMallocHost(&sl.h_v, (size_t)h->B);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_v, (size_t)h->B);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    std::string msg = cudaGetErrorString(e);
    uwvkf_stream_destroy(s);
    return fail(UWVKF_ERR_CUDA, "stream allocation failed: " + msg);
  }
  *out = s;
  return UWVKF_OK;
}

int uwvkf_stream_destroy(uwvkf_stream s) {
  if (!s) return UWVKF_OK;
  cudaSetDevice(s->h->device);
  if (s->copy) cudaStreamSynchronize(s->copy);
  cudaStreamSynchronize(s->h->stream);
  for (auto &sl : s->ring) {
    cudaFreeHost(sl.h_z); cudaFree(sl.d_z); cudaFreeHost(sl.h_v); cudaFree(sl.d_v);
    if (sl.copied) cudaEventDestroy(sl.copied);
    if (sl.consumed) cudaEventDestroy(sl.consumed);
  }
  if (s->copy) cudaStreamDestroy(s->copy);
  delete s;
  return UWVKF_OK;
}

int uwvkf_stream_pending(uwvkf_stream s) { return s ? s->count : 0; }
double uwvkf_stream_time(uwvkf_stream s) { return s ? s->t_filter : 0.0; }

int uwvkf_stream_push(uwvkf_stream s, double timestamp, int kind, const double *z, const double *R, const double *aux,
                      const uint8_t *valid) {
  if (!s) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null stream");
  uwvkf_handle h = s->h;
  CK(cudaSetDevice(h->device));
  const int m = kind == UWVKF_STREAM_ROTATION_RATE ? 3 : meas_dim(kind);
  if (m < 0 || !z || std::isnan(timestamp)) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad sample");
  if (kind != UWVKF_STREAM_ROTATION_RATE && !R) return fail(UWVKF_ERR_INVALID_ARGUMENT, "R must not be NULL");
  if (s->count == s->slots) return fail(UWVKF_ERR_BUSY, "sample ring is full: call uwvkf_stream_run");
  if (s->count > 0) {
    const auto &last = s->ring[(s->head + s->count - 1) % s->slots];
    if (timestamp < last.ts) return fail(UWVKF_ERR_INVALID_ARGUMENT, "samples must be pushed in time order");
  } else if (s->have_time && timestamp < s->t_filter) {
    return fail(UWVKF_ERR_INVALID_ARGUMENT, "sample is older than the filter time");
  }
  // validate everything before the slot is touched: a refused sample leaves the ring as it was (checkMeasurment())
  const size_t n = (size_t)h->B * m;
  if (any_nan(z, (int64_t)n)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "sample contains NaN");
  if (R && any_nan(R, m * m)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "R contains NaN");
  auto &sl = s->ring[(s->head + s->count) % s->slots];
  if (sl.used) CK(cudaEventSynchronize(sl.copied));   // the staging buffer's previous upload must have left the host
  std::memcpy(sl.h_z, z, n * sizeof(double));
  sl.ts = timestamp; sl.kind = kind;
  if (R) for (int k = 0; k < m * m; ++k) sl.R[k] = R[k];
  sl.has_aux = aux != nullptr;
  if (aux) {
    const int na = (kind == UWVKF_POSE_WATER_VELOCITY) ? 1 : (kind == UWVKF_POSE_DELAYED_XY ? 2 : 3);
    for (int k = 0; k < na; ++k) sl.aux[k] = aux[k];
  }
  sl.has_valid = valid != nullptr;
  if (valid) std::memcpy(sl.h_v, valid, (size_t)h->B);
  // the HBM copy of this slot may still be read by kernels of an earlier lap of the ring: order the upload behind them
  if (sl.used) CK(cudaStreamWaitEvent(s->copy, sl.consumed, 0));
  CK(cudaMemcpyAsync(sl.d_z, sl.h_z, n * sizeof(double), cudaMemcpyHostToDevice, s->copy));
  if (valid) CK(cudaMemcpyAsync(sl.d_v, sl.h_v, (size_t)h->B, cudaMemcpyHostToDevice, s->copy));
  CK(cudaEventRecord(sl.copied, s->copy));
  sl.used = true;
  s->count++;
  return UWVKF_OK;
}

int uwvkf_stream_run(uwvkf_stream s, double until) {
  if (!s) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null stream");
  uwvkf_handle h = s->h;
  CK(cudaSetDevice(h->device));
  while (s->count > 0) {
    auto &sl = s->ring[s->head];
    if (sl.ts > until) break;
    CK(cudaStreamWaitEvent(h->stream, sl.copied, 0));
    int rc = UWVKF_OK;
    if (sl.kind == UWVKF_STREAM_ROTATION_RATE) {
      rc = uwvkf_pose_set_rotation_rate(h, sl.d_z, UWVKF_DEVICE);          // control input of the following predictions
    } else {
      const double dt = s->have_time ? sl.ts - s->t_filter : 0.0;          // predictionStepFromSampleTime: the first sample latches
      const double *aux = sl.has_aux ? sl.aux : nullptr;
      const uint8_t *v = sl.has_valid ? sl.d_v : nullptr;
      if (dt > 0.0) rc = uwvkf_pose_predict_update(h, dt, nullptr, sl.kind, sl.d_z, sl.R, 0, aux, v, UWVKF_DEVICE);
      else rc = uwvkf_pose_update(h, sl.kind, sl.d_z, sl.R, 0, aux, v, UWVKF_DEVICE);
      if (rc == UWVKF_OK) { s->t_filter = sl.ts; s->have_time = true; }
    }
    if (rc) return rc;
    CK(cudaEventRecord(sl.consumed, h->stream));
    s->head = (s->head + 1) % s->slots;
    s->count--;
  }
  return UWVKF_OK;
}

/* ---- CUDA-graph execution of a repeating schedule --------------------------------------------------------- */
struct uwvkf_graph_s {
  uwvkf_handle h = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches_per_replay = 0;
};

int uwvkf_graph_begin(uwvkf_handle h) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  if (!h->stream) return fail(UWVKF_ERR_INVALID_ARGUMENT, "graph capture needs a user stream (uwvkf_set_stream): the legacy default stream cannot be captured");
  if (h->capturing) return fail(UWVKF_ERR_INVALID_ARGUMENT, "a capture is already in progress");
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
  h->capturing = true;
  h->launches_at_capture = h->launches;
  return UWVKF_OK;
}

int uwvkf_graph_end(uwvkf_handle h, uwvkf_graph *out) {
  if (!h || !out) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad arguments");
  CK(cudaSetDevice(h->device));
  if (!h->capturing) return fail(UWVKF_ERR_INVALID_ARGUMENT, "no capture in progress");
  h->capturing = false;
  uwvkf_graph g = new uwvkf_graph_s();
  g->h = h;
  g->launches_per_replay = h->launches - h->launches_at_capture;
  h->launches = h->launches_at_capture;                       // nothing ran during the capture
  cudaError_t e = cudaStreamEndCapture(h->stream, &g->graph);
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    std::string msg = cudaGetErrorString(e);
    cudaGetLastError();
    uwvkf_graph_destroy(g);
    return fail(UWVKF_ERR_CUDA, "graph capture failed (only calls with memspace = UWVKF_DEVICE can be captured): " + msg);
  }
  *out = g;
  return UWVKF_OK;
}

int uwvkf_graph_launch(uwvkf_graph g) {
  if (!g || !g->exec) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null graph");
  CK(cudaSetDevice(g->h->device));
  CK(cudaGraphLaunch(g->exec, g->h->stream));
  g->h->launches += g->launches_per_replay;
  return UWVKF_OK;
}

int uwvkf_graph_destroy(uwvkf_graph g) {
  if (!g) return UWVKF_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return UWVKF_OK;
}

int uwvkf_pose_step_host(uwvkf_handle h, double dt, const double *w_host, int meas_kind, const double *z_host, const double *R,
                         const double *aux, double *nav_out_host) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  const int m = meas_dim(meas_kind);
  if (m < 0 || m > 3) return fail(UWVKF_ERR_INVALID_ARGUMENT, "measurement kind not available in the pipelined step");
  if (!w_host || !z_host || !R || std::isnan(dt)) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null input or NaN dt");
  if (h->capturing) return fail(UWVKF_ERR_INVALID_ARGUMENT, "host-buffer calls cannot be captured into a graph");
  if (!h->pipe) {
    auto *p = new uwvkf_filter_s::HostPipe();
    h->pipe = p;
    cudaError_t e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
      e = cudaMalloc(&p->d_w[k], (size_t)h->B * 3 * sizeof(double));
      if (e == cudaSuccess) e = cudaMalloc(&p->d_z[k], (size_t)h->B * 3 * sizeof(double));
      if (e == cudaSuccess) e = cudaMalloc(&p->d_nav[k], (size_t)h->B * 10 * sizeof(double));
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_in[k], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_main[k], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_out[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {   // leave no half-built pipeline behind: the next call starts over
      const std::string msg = cudaGetErrorString(e);
      cudaGetLastError();
      destroy_pipe(h);
      return fail(UWVKF_ERR_CUDA, "pipeline allocation failed: " + msg);
    }
  }
  auto *p = h->pipe;
  const int slot = (int)(p->calls & 1);
  // host-side pacing (and the buffer rules of the header): the upload of the previous call has left its host buffers, the
  // download issued two calls ago (same slot) has reached its host buffer
  if (p->calls >= 1) CK(cudaEventSynchronize(p->ev_in[slot ^ 1]));
  if (p->out_pending[slot]) { CK(cudaEventSynchronize(p->ev_out[slot])); p->out_pending[slot] = false; }
  // upload: the slot's previous reader is the kernel of call i-2
  if (p->calls >= 2) CK(cudaStreamWaitEvent(p->s_in, p->ev_main[slot], 0));
  CK(cudaMemcpyAsync(p->d_w[slot], w_host, (size_t)h->B * 3 * sizeof(double), cudaMemcpyHostToDevice, p->s_in));
  CK(cudaMemcpyAsync(p->d_z[slot], z_host, (size_t)h->B * m * sizeof(double), cudaMemcpyHostToDevice, p->s_in));
  CK(cudaEventRecord(p->ev_in[slot], p->s_in));
  // fused step on the handle's stream (the navigation buffer of this slot was read back two calls ago: waited for above)
  CK(cudaStreamWaitEvent(h->stream, p->ev_in[slot], 0));
  // the fused kernel stores the navigation solution of every filter itself (Launch::nav): no second kernel, no second pass over mu
  const bool nav_in_kernel = nav_out_host && !engine_v1();
  h->nav_fused = nav_in_kernel ? p->d_nav[slot] : nullptr;
  int rc = uwvkf_pose_predict_update(h, dt, p->d_w[slot], meas_kind, p->d_z[slot], R, 0, aux, nullptr, UWVKF_DEVICE);
  h->nav_fused = nullptr;
  if (rc) return rc;
  if (nav_out_host && !nav_in_kernel) {
    rc = uwvkf_pose_get_navigation(h, p->d_nav[slot], 0, h->B, UWVKF_DEVICE);
    if (rc) return rc;
  }
  CK(cudaEventRecord(p->ev_main[slot], h->stream));
  if (nav_out_host) {
    CK(cudaStreamWaitEvent(p->s_out, p->ev_main[slot], 0));
    CK(cudaMemcpyAsync(nav_out_host, p->d_nav[slot], (size_t)h->B * 10 * sizeof(double), cudaMemcpyDeviceToHost, p->s_out));
    CK(cudaEventRecord(p->ev_out[slot], p->s_out));
    p->out_pending[slot] = true;
  }
  p->calls++;
  return UWVKF_OK;
}

int uwvkf_pose_get_navigation(uwvkf_handle h, double *out, int64_t first, int64_t count, int memspace) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!out || first < 0 || count <= 0 || first + count > h->B) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad arguments");
  double *dout = out;
  if (memspace == UWVKF_HOST) {
    int rc = ensure_tmp(h, (size_t)count * 10 * sizeof(double));
    if (rc) return rc;
    dout = h->d_tmp;
  }
  pose_get_nav_kernel<<<(unsigned)((count * 10 + 255) / 256), 256, 0, h->stream>>>(h->d_state, dout, first, count);
  CK(cudaGetLastError());
  h->launches++;
  if (memspace == UWVKF_HOST) {
    CK(cudaMemcpyAsync(out, dout, (size_t)count * 10 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return UWVKF_OK;
}

int uwvkf_pose_reset_with_external_pose(uwvkf_handle h, const double *pose7) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!pose7) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null pose");
  int rc = ensure_tmp(h, (size_t)h->B * 7 * sizeof(double));
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_tmp, pose7, (size_t)h->B * 7 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  pose_set_pose_kernel<<<(unsigned)((h->B * 7 + 255) / 256), 256, 0, h->stream>>>(h->d_state, h->d_tmp, h->B);
  CK(cudaGetLastError());
  h->launches++;
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

int uwvkf_debug_phase_cycles(uwvkf_handle h, uint64_t *out, int reset) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!out) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null output");
  CK(cudaStreamSynchronize(h->stream));
#ifdef UWVKF_PHASE_TIMING
  CK(cudaMemcpyFromSymbol(out, pose::g_phase_cycles, 40 * sizeof(uint64_t)));
  if (reset) {
    static const uint64_t zeros[40] = {0};
    CK(cudaMemcpyToSymbol(pose::g_phase_cycles, zeros, sizeof(zeros)));
  }
  return UWVKF_OK;
#else
  (void)reset;
  for (int i = 0; i < 40; ++i) out[i] = 0;
  return fail(UWVKF_ERR_INVALID_ARGUMENT, "library was built without -DUWVKF_PHASE_TIMING");
#endif
}

int uwvkf_debug_eval(uwvkf_handle h, int what, const double *in, double *out, int n) {
  REQUIRE(h, UWVKF_FILTER_POSE);
  if (!in || !out || n <= 0 || what < 0 || what > UWVKF_DEBUG_WARP_SUM) return fail(UWVKF_ERR_INVALID_ARGUMENT, "bad arguments");
  if (!h->have_params) return fail(UWVKF_ERR_NOT_INITIALIZED, "uwvkf_pose_set_params was not called");
  int rc = ensure_tmp(h, (size_t)n * 12 * sizeof(double));
  if (rc) return rc;
  double *din = h->d_tmp, *dout = h->d_tmp + (size_t)n * 8;
  CK(cudaMemcpyAsync(din, in, (size_t)n * 8 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  debug_eval_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(what, din, dout, n, h->pp);
  CK(cudaGetLastError());
  h->launches++;
  CK(cudaMemcpyAsync(out, dout, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

int64_t uwvkf_ensemble_stats_size(uwvkf_handle h) {
  if (!h) return 0;
  const int d = h->kind == UWVKF_FILTER_POSE ? pose::DIM : h->dim;
  return 1 + d + d * (d + 1) / 2;
}

int uwvkf_ensemble_stats(uwvkf_handle h, double *out, int memspace) {
  if (!h) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null handle");
  CK(cudaSetDevice(h->device));
  if (!out) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null output");
  const size_t bytes = (size_t)uwvkf_ensemble_stats_size(h) * sizeof(double);
  double *dout = out;
  if (memspace == UWVKF_HOST) {
    int rc = ensure_tmp(h, bytes);
    if (rc) return rc;
    dout = h->d_tmp;
  }
  CK(cudaMemsetAsync(dout, 0, bytes, h->stream));
  if (h->kind != UWVKF_FILTER_POSE) {
    int rc = small_ensemble_stats(h, dout);
    if (rc) return rc;
  } else {
    int grid = (int)std::min<int64_t>(h->B, (int64_t)h->sm_count * 8);
    ensemble_stats_kernel<pose::DIM><<<grid, 256, 0, h->stream>>>(h->d_state, pose::BLK, h->B, dout);
  }
  CK(cudaGetLastError());
  h->launches++;
  if (memspace == UWVKF_HOST) {
    CK(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return UWVKF_OK;
}

}  // extern "C"
