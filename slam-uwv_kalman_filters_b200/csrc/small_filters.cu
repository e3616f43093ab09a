// VelocityUKF / BottomUKF / IndirectPoseUKF: kernels (one thread per filter) and their C ABI entry points.
#include "internal.h"

namespace {

using namespace uwvkf;
using namespace uwvkf::small;

constexpr int kThreads = 64;

struct Mat36 { double v[36]; };
struct Vec3A { double v[3]; };

// ---- VelocityUKF ------------------------------------------------------------------------------------------
// predictionStepImpl, reference src/VelocityUKF.cpp:114-130 with processMotionModel :6-33
__global__ void __launch_bounds__(kThreads) velocity_predict_kernel(double *state, double *sim, const double *gyro,
                                                                    const double *efforts, Mat36 Q, const __grid_constant__ UwvParams P,
                                                                    double dt, long long B, uint32_t *status, Policy pol) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  Ukf<VectM<4>> u;
  u.pol = pol; u.status = 0;
  load_filter(u, state, f);
  double *sm = sim + f * 10;
  const Quat q{sm[0], sm[1], sm[2], sm[3]};                      // :121 model_state.orientation
  double w[3] = {gyro[f * 3], gyro[f * 3 + 1], gyro[f * 3 + 2]};
  double tau[6];
  for (int k = 0; k < 6; ++k) tau[k] = efforts[f * 6 + k];
  double Qs[16];
  for (int k = 0; k < 16; ++k) Qs[k] = dt * Q.v[k];              // :123
  bool ok = u.predict([&](const double *x, double *o) {
    double vel6[6] = {x[0], x[1], x[2], w[0], w[1], w[2]}, nv[6];
    sim_step(P, tau, q, vel6, dt, nv);                           // :18-21
    for (int k = 0; k < 3; ++k) o[k] = x[k] + (nv[k] - x[k]);    // :24-26
    Vec3 vn = qrot(q, Vec3{o[0], o[1], o[2]});                   // :30
    o[3] = x[3] + dt * vn.z;                                     // :31
  }, Qs);
  if (ok) store_filter(u, state, f);
  // :126-127 advance the private motion model
  double v6[6], nv[6];
  for (int k = 0; k < 6; ++k) v6[k] = sm[4 + k];
  sim_step(P, tau, q, v6, dt, nv);
  for (int k = 0; k < 6; ++k) sm[4 + k] = nv[k];
  if (u.status) status[f] |= u.status;
}

// integrateMeasurement(DVLMeasurement) :79-85 (KIND 0) / integrateMeasurement(PressureMeasurement) :106-112 (KIND 1)
template <int KIND>
__global__ void __launch_bounds__(kThreads) velocity_update_kernel(double *state, const double *z, Mat36 R, const uint8_t *valid,
                                                                   long long B, uint32_t *status, Policy pol) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  if (valid && !valid[f]) return;
  constexpr int M = KIND == 0 ? 3 : 1;
  double zz[M];
  bool bad = false;
  for (int k = 0; k < M; ++k) { zz[k] = z[f * M + k]; bad |= isnan(zz[k]); }
  if (bad) { status[f] |= ST_NAN_MEAS; return; }
  Ukf<VectM<4>> u;
  u.pol = pol; u.status = 0;
  load_filter(u, state, f);
  bool ok;
  if constexpr (KIND == 0)
    ok = u.update<VectM<3>>(zz, [](const double *x, double *o) { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; }, R.v);  // :35-40
  else
    ok = u.update<VectM<1>>(zz, [](const double *x, double *o) { o[0] = x[3]; }, R.v);                            // :42-47
  if (ok) store_filter(u, state, f);
  if (u.status) status[f] |= u.status;
}

// setupMotionModel :58-77
__global__ void velocity_setup_sim_kernel(const double *state, const double *gyro, double *sim, long long B) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  double *sm = sim + f * 10;
  sm[0] = sm[1] = sm[2] = 0; sm[3] = 1;
  for (int k = 0; k < 3; ++k) { sm[4 + k] = state[f * 20 + k]; sm[7 + k] = gyro[f * 3 + k]; }
}
// integrateMeasurement(GyroMeasurement) :87-98
__global__ void velocity_set_gyro_kernel(const double *w, double *gyro, double *sim, int have_model, long long B) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  for (int k = 0; k < 3; ++k) {
    double v = w[f * 3 + k];
    gyro[f * 3 + k] = v;
    if (have_model) sim[f * 10 + 7 + k] = v;
  }
}

// ---- BottomUKF --------------------------------------------------------------------------------------------
// predictionStepImpl src/BottomUKF.cpp:48-54 with processModel :5-16
__global__ void __launch_bounds__(kThreads) bottom_predict_kernel(double *state, const double *vel, Mat36 Q, double dt, long long B,
                                                                  uint32_t *status, Policy pol) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  Ukf<BottomM> u;
  u.pol = pol; u.status = 0;
  load_filter(u, state, f);
  const double vx = vel[f * 3], vy = vel[f * 3 + 1], vz = vel[f * 3 + 2];
  const double nxy = sqrt(vx * vx + vy * vy);
  const double scale = (nxy * nxy) * (dt * dt);                  // :50
  double Qs[9];
  for (int k = 0; k < 9; ++k) Qs[k] = scale * Q.v[k];
  bool ok = u.predict([&](const double *x, double *o) {
    o[0] = x[0] + dt * (-1. * vz);                               // :10-12
    o[1] = x[1]; o[2] = x[2]; o[3] = x[3];
  }, Qs);
  if (ok) store_filter(u, state, f);
  if (u.status) status[f] |= u.status;
}
// integrateMeasurement(RangeMeasurement, unit_direction, origin) :56-61 with measurementDistance :18-30
__global__ void __launch_bounds__(kThreads) bottom_range_kernel(double *state, const double *z, double R, Vec3A dir, Vec3A org,
                                                                const uint8_t *valid, long long B, uint32_t *status, Policy pol) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  if (valid && !valid[f]) return;
  double zz[1] = {z[f]};
  if (isnan(zz[0])) { status[f] |= ST_NAN_MEAS; return; }
  Ukf<BottomM> u;
  u.pol = pol; u.status = 0;
  load_filter(u, state, f);
  double Rm[1] = {R};
  bool ok = u.update<VectM<1>>(zz, [&](const double *x, double *o) {
    double b0 = 0. - org.v[0], b1 = 0. - org.v[1], b2 = -x[0] - org.v[2];
    const double *n = x + 1;
    double v = dir.v[0] * n[0] + dir.v[1] * n[1] + dir.v[2] * n[2];
    o[0] = (v != 0.) ? (b0 * n[0] + b1 * n[1] + b2 * n[2]) / v : 0.;
  }, Rm);
  if (ok) store_filter(u, state, f);
  if (u.status) status[f] |= u.status;
}
// integrateMeasurement(NormalType, cov) :63-67 with measurementNormal :32-37 (measurement on S2)
__global__ void __launch_bounds__(kThreads) bottom_normal_kernel(double *state, const double *z, Mat36 R, const uint8_t *valid,
                                                                 long long B, uint32_t *status, Policy pol) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  if (valid && !valid[f]) return;
  double zz[3] = {z[f * 3], z[f * 3 + 1], z[f * 3 + 2]};
  if (isnan(zz[0]) || isnan(zz[1]) || isnan(zz[2])) { status[f] |= ST_NAN_MEAS; return; }
  Ukf<BottomM> u;
  u.pol = pol; u.status = 0;
  load_filter(u, state, f);
  bool ok = u.update<S2M>(zz, [](const double *x, double *o) { o[0] = x[1]; o[1] = x[2]; o[2] = x[3]; }, R.v);
  if (ok) store_filter(u, state, f);
  if (u.status) status[f] |= u.status;
}

// ---- IndirectPoseUKF --------------------------------------------------------------------------------------
// predictionStepImpl src/IndirectPoseUKF.cpp:80-92 with processModel :7-20
__global__ void __launch_bounds__(kThreads) indirect_predict_kernel(double *state, Mat36 Q, double tau, double dt, long long B,
                                                                    uint32_t *status, Policy pol) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  Ukf<PoseErrM> u;
  u.pol = pol; u.status = 0;
  load_filter(u, state, f);
  double rot[9], Qe[36], t[9];
  qmat(Quat{u.mu[3], u.mu[4], u.mu[5], u.mu[6]}, rot);            // :82
  for (int k = 0; k < 36; ++k) Qe[k] = Q.v[k];
  const double fs = 2. / (tau * dt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {                                 // :85-86
      double s = 0;
      for (int k = 0; k < 3; ++k) s += rot[i * 3 + k] * (fs * Q.v[(3 + k) * 6 + 3 + j]);
      t[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += t[i * 3 + k] * rot[j * 3 + k];
      Qe[(3 + i) * 6 + 3 + j] = s;
    }
  for (int k = 0; k < 36; ++k) Qe[k] *= dt * dt;                  // :88
  bool ok = u.predict([&](const double *x, double *o) {
    for (int k = 0; k < 3; ++k) o[k] = x[k];
    Vec3 lg = so3_log(Quat{x[3], x[4], x[5], x[6]});              // :15
    double d[3] = {(-1.0 / tau) * lg.x, (-1.0 / tau) * lg.y, (-1.0 / tau) * lg.z};  // :16
    so3_bp(x + 3, d, dt, o + 3);                                  // :17
  }, Qe);
  if (ok) store_filter(u, state, f);
  if (u.status) status[f] |= u.status;
}

struct FeatArgs {
  int nfeat;
  double pxcov[8][4], fpos[8][3];
  double marker[7], covm[36], cam[4], cib[7];
};
// integrateMeasurement(marker_features, ...) src/IndirectPoseUKF.cpp:94-135 with measurementVisualLandmark :36-50
__global__ void __launch_bounds__(kThreads) indirect_features_kernel(double *state, const double *px, const double *pose_ref,
                                                                     const __grid_constant__ FeatArgs A, const uint8_t *valid,
                                                                     long long B, uint32_t *status, Policy pol) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  if (valid && !valid[f]) return;
  const double *p = state + f * 43;
  Ukf<PoseErrMarkerM> u;
  u.pol = pol; u.status = 0;
  for (int k = 0; k < 7; ++k) { u.mu[k] = p[k]; u.mu[7 + k] = A.marker[k]; }       // :100-103
  for (int k = 0; k < 144; ++k) u.S[k] = 0.0;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {                                                   // :104-106
      u.S[i * 12 + j] = p[7 + i * 6 + j];
      u.S[(6 + i) * 12 + 6 + j] = A.covm[i * 6 + j];
    }
  double ref[7];
  for (int k = 0; k < 7; ++k) ref[k] = pose_ref[f * 7 + k];
  const double fx = A.cam[0], fy = A.cam[1], cx = A.cam[2], cy = A.cam[3];
  const double fx2 = fx * fx, fy2 = fy * fy, fxy = fx * fy;                         // :109-111
  for (int i = 0; i < A.nfeat; ++i) {                                               // :114
    double u0 = px[(f * A.nfeat + i) * 2], v0 = px[(f * A.nfeat + i) * 2 + 1];
    if (isnan(u0) || isnan(v0)) { status[f] |= ST_NAN_MEAS; return; }               // :116
    double z[3] = {(u0 - cx) / fx, (v0 - cy) / fy, 1.0};                            // :119-121
    double nz = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    z[0] /= nz; z[1] /= nz; z[2] /= nz;
    double R[4] = {A.pxcov[i][0] / fx2, A.pxcov[i][1] / fxy, A.pxcov[i][2] / fxy, A.pxcov[i][3] / fy2};  // :122-124
    const double *fp = A.fpos[i];
    u.update<S2M>(z, [&](const double *x, double *o) {                              // :126
      double a[7], c[7];
      xf_mul(ref, x, a);          // body_in_nav * pose_error
      xf_mul(a, A.cib, c);        // * cam_in_body
      Vec3 m = qrot(Quat{x[10], x[11], x[12], x[13]}, Vec3{fp[0], fp[1], fp[2]});
      Vec3 d{m.x + x[7] - c[0], m.y + x[8] - c[1], m.z + x[9] - c[2]};
      Vec3 r = qrot_inv(Quat{c[3], c[4], c[5], c[6]}, d);
      double n = sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
      o[0] = r.x / n; o[1] = r.y / n; o[2] = r.z / n;
    }, R);
  }
  if (!(u.status & ST_NOT_SPD)) {                                                   // :134
    double *o = state + f * 43;
    for (int k = 0; k < 7; ++k) o[k] = u.mu[k];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) o[7 + i * 6 + j] = u.S[i * 12 + j];
  }
  if (u.status) status[f] |= u.status;
}
// getCorrectedPose :137-142
__global__ void indirect_corrected_pose_kernel(const double *state, const double *pose_ref, double *out, long long B) {
  long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  double r[7];
  xf_mul(pose_ref + f * 7, state + f * 43, r);
  for (int k = 0; k < 7; ++k) out[f * 7 + k] = r[k];
}

__global__ void small_pack_kernel(const double *mu, const double *cov, long long cov_stride, double *state, int dim, int nn,
                                  long long first, long long count) {
  const int w = dim + nn;
  long long total = count * w;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    long long f = g / w;
    int k = (int)(g - f * w);
    state[(first + f) * w + k] = k < dim ? mu[f * dim + k] : cov[f * cov_stride + (k - dim)];
  }
}
__global__ void small_unpack_kernel(const double *state, double *mu, double *cov, int dim, int nn, long long first, long long count) {
  const int w = dim + nn;
  long long total = count * w;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    long long f = g / w;
    int k = (int)(g - f * w);
    double v = state[(first + f) * w + k];
    if (k < dim) { if (mu) mu[f * dim + k] = v; }
    else if (cov) cov[f * nn + (k - dim)] = v;
  }
}

inline unsigned grid_for(int64_t B) { return (unsigned)((B + kThreads - 1) / kThreads); }
inline Policy policy_of(uwvkf_handle h) { return Policy{h->mean_tol, h->mean_max_it}; }
inline Mat36 mat_of(const double *src, int n) {
  Mat36 m{};
  for (int k = 0; k < n; ++k) m.v[k] = src[k];
  return m;
}
int copy_ctrl(uwvkf_handle h, double *dst, const double *src, int per, int memspace) {
  if (!src) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null input");
  if (memspace == UWVKF_HOST && any_nan(src, h->B * per)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "input contains NaN");
  CK(cudaMemcpyAsync(dst, src, (size_t)h->B * per * sizeof(double),
                     memspace == UWVKF_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  return UWVKF_OK;
}
// measurement inputs: stage z (+ valid) on the device, NaN-check small host batches like checkMeasurment()
int stage_meas(uwvkf_handle h, const double *z, int m, const uint8_t *valid, int memspace, const double **dz, const uint8_t **dv) {
  if (!z) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null measurement");
  if (memspace == UWVKF_HOST && any_nan(z, h->B * m)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "measurement contains NaN");
  const void *p;
  int rc = stage(h, z, (size_t)h->B * m * sizeof(double), (size_t)h->B * 16 * sizeof(double), memspace, (void **)&h->d_z, &p);
  if (rc) return rc;
  *dz = (const double *)p;
  rc = stage(h, valid, (size_t)h->B, (size_t)h->B, memspace, (void **)&h->d_valid, &p);
  if (rc) return rc;
  *dv = (const uint8_t *)p;
  return UWVKF_OK;
}

}  // namespace

namespace uwvkf {

int small_create(uwvkf_handle h) {
  switch (h->kind) {
    case UWVKF_FILTER_VELOCITY: h->dim = 4; h->dof = 4; break;
    case UWVKF_FILTER_BOTTOM: h->dim = 4; h->dof = 3; break;
    default: h->dim = 7; h->dof = 6; break;
  }
  const size_t w = (size_t)(h->dim + h->dof * h->dof);
  CK(cudaMalloc(&h->d_state, (size_t)h->B * w * sizeof(double)));
  CK(cudaMemset(h->d_state, 0, (size_t)h->B * w * sizeof(double)));
  CK(cudaMalloc(&h->d_status, (size_t)h->B * sizeof(uint32_t)));
  CK(cudaMemset(h->d_status, 0, (size_t)h->B * sizeof(uint32_t)));
  if (h->kind == UWVKF_FILTER_VELOCITY) {
    // constructor, reference src/VelocityUKF.cpp:49-56: efforts = 0, process noise = diag(1e-4) on the velocity
    CK(cudaMalloc(&h->d_ctrl0, (size_t)h->B * 3 * sizeof(double)));
    CK(cudaMalloc(&h->d_ctrl1, (size_t)h->B * 6 * sizeof(double)));
    CK(cudaMalloc(&h->d_sim, (size_t)h->B * 10 * sizeof(double)));
    CK(cudaMemset(h->d_ctrl0, 0, (size_t)h->B * 3 * sizeof(double)));
    CK(cudaMemset(h->d_ctrl1, 0, (size_t)h->B * 6 * sizeof(double)));
    CK(cudaMemset(h->d_sim, 0, (size_t)h->B * 10 * sizeof(double)));
    for (int k = 0; k < 3; ++k) h->Qh[k * 4 + k] = 0.0001;
  } else if (h->kind == UWVKF_FILTER_BOTTOM) {
    // constructor, src/BottomUKF.cpp:40-46: velocity = 0, process noise = identity
    CK(cudaMalloc(&h->d_ctrl0, (size_t)h->B * 3 * sizeof(double)));
    CK(cudaMemset(h->d_ctrl0, 0, (size_t)h->B * 3 * sizeof(double)));
    for (int k = 0; k < 3; ++k) h->Qh[k * 3 + k] = 1.0;
  } else {
    // constructor, src/IndirectPoseUKF.cpp:53-78: pose_ref = identity
    CK(cudaMalloc(&h->d_ctrl0, (size_t)h->B * 7 * sizeof(double)));
    std::vector<double> id((size_t)h->B * 7, 0.0);
    for (int64_t f = 0; f < h->B; ++f) id[f * 7 + 6] = 1.0;
    CK(cudaMemcpy(h->d_ctrl0, id.data(), id.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  return UWVKF_OK;
}

void small_destroy(uwvkf_handle h) {
  cudaFree(h->d_ctrl0); cudaFree(h->d_ctrl1); cudaFree(h->d_sim);
}

int small_set_state(uwvkf_handle h, const double *mu, const double *cov, int64_t cov_stride, int64_t first, int64_t count,
                    int memspace) {
  const int nn = h->dof * h->dof;
  if (cov_stride != 0 && cov_stride != nn) return fail(UWVKF_ERR_INVALID_ARGUMENT, "cov_stride must be 0 (shared) or DOF*DOF");
  const double *dmu = mu, *dcov = cov;
  if (memspace == UWVKF_HOST) {
    const size_t nm = (size_t)count * h->dim, nc = cov_stride ? (size_t)count * nn : (size_t)nn;
    int rc = ensure_tmp(h, (nm + nc) * sizeof(double));
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_tmp, mu, nm * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_tmp + nm, cov, nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    dmu = h->d_tmp; dcov = h->d_tmp + nm;
  }
  small_pack_kernel<<<1024, 256, 0, h->stream>>>(dmu, dcov, cov_stride, h->d_state, h->dim, nn, first, count);
  CK(cudaGetLastError());
  h->launches++;
  if (memspace == UWVKF_HOST) CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

int small_get_state(uwvkf_handle h, double *mu, double *cov, int64_t first, int64_t count, int memspace) {
  const int nn = h->dof * h->dof;
  double *dmu = mu, *dcov = cov;
  const size_t nm = (size_t)count * h->dim, nc = (size_t)count * nn;
  if (memspace == UWVKF_HOST) {
    int rc = ensure_tmp(h, (nm + nc) * sizeof(double));
    if (rc) return rc;
    dmu = mu ? h->d_tmp : nullptr; dcov = cov ? h->d_tmp + nm : nullptr;
  }
  small_unpack_kernel<<<1024, 256, 0, h->stream>>>(h->d_state, dmu, dcov, h->dim, nn, first, count);
  CK(cudaGetLastError());
  h->launches++;
  if (memspace == UWVKF_HOST) {
    if (mu) CK(cudaMemcpyAsync(mu, dmu, nm * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (cov) CK(cudaMemcpyAsync(cov, dcov, nc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return UWVKF_OK;
}

int small_predict(uwvkf_handle h, double dt) {
  const Policy pol = policy_of(h);
  if (h->kind == UWVKF_FILTER_VELOCITY) {
    // reference src/VelocityUKF.cpp:117-118
    if (!h->have_model) return fail(UWVKF_ERR_NOT_INITIALIZED, "Motion model is not initialized!");
    velocity_predict_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, h->d_sim, h->d_ctrl0, h->d_ctrl1, mat_of(h->Qh, 16),
                                                                        h->uwv, dt, h->B, h->d_status, pol);
  } else if (h->kind == UWVKF_FILTER_BOTTOM) {
    bottom_predict_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, h->d_ctrl0, mat_of(h->Qh, 9), dt, h->B, h->d_status, pol);
  } else {
    if (!(h->tau > 0)) return fail(UWVKF_ERR_NOT_INITIALIZED, "uwvkf_indirect_set_tau was not called");
    indirect_predict_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, mat_of(h->Qh, 36), h->tau, dt, h->B, h->d_status, pol);
  }
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

}  // namespace uwvkf

// ======================================================================================================
extern "C" {

int uwvkf_velocity_setup_motion_model(uwvkf_handle h, const double *uwv) {
  REQUIRE(h, UWVKF_FILTER_VELOCITY);
  if (!uwv) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null parameters");
  UwvParams &P = h->uwv;
  std::memcpy(P.inertia, uwv, 36 * sizeof(double));
  std::memcpy(P.damp_lin, uwv + 36, 36 * sizeof(double));
  std::memcpy(P.damp_quad, uwv + 72, 36 * sizeof(double));
  std::memcpy(P.cog, uwv + 108, 3 * sizeof(double));
  std::memcpy(P.cob, uwv + 111, 3 * sizeof(double));
  P.weight = uwv[114]; P.buoyancy = uwv[115];
  // inverse of the inertia matrix (Gauss-Jordan with partial pivoting)
  double a[6][12];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { a[i][j] = P.inertia[i * 6 + j]; a[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
    if (a[p][c] == 0.0) return fail(UWVKF_ERR_INVALID_ARGUMENT, "inertia matrix is singular");
    if (p != c) for (int k = 0; k < 12; ++k) std::swap(a[p][k], a[c][k]);
    double ip = 1.0 / a[c][c];
    for (int k = 0; k < 12; ++k) a[c][k] *= ip;
    for (int r = 0; r < 6; ++r) if (r != c) { double f = a[r][c]; for (int k = 0; k < 12; ++k) a[r][k] -= f * a[c][k]; }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) P.inertia_inv[i * 6 + j] = a[i][6 + j];
  velocity_setup_sim_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, h->d_ctrl0, h->d_sim, h->B);
  CK(cudaGetLastError());
  h->launches++;
  h->have_model = true;
  return UWVKF_OK;
}

int uwvkf_velocity_set_gyro(uwvkf_handle h, const double *w, int memspace) {
  REQUIRE(h, UWVKF_FILTER_VELOCITY);
  if (!w) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null input");
  if (memspace == UWVKF_HOST && any_nan(w, h->B * 3)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "measurement contains NaN");
  const void *p;
  int rc = stage(h, w, (size_t)h->B * 3 * sizeof(double), (size_t)h->B * 16 * sizeof(double), memspace, (void **)&h->d_z, &p);
  if (rc) return rc;
  velocity_set_gyro_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>((const double *)p, h->d_ctrl0, h->d_sim, h->have_model ? 1 : 0, h->B);
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

int uwvkf_velocity_set_efforts(uwvkf_handle h, const double *tau, int memspace) {
  REQUIRE(h, UWVKF_FILTER_VELOCITY);
  return copy_ctrl(h, h->d_ctrl1, tau, 6, memspace);
}

int uwvkf_velocity_update(uwvkf_handle h, int kind, const double *z, const double *R, const uint8_t *valid, int memspace) {
  REQUIRE(h, UWVKF_FILTER_VELOCITY);
  if (kind != UWVKF_VEL_DVL && kind != UWVKF_VEL_PRESSURE) return fail(UWVKF_ERR_INVALID_ARGUMENT, "unknown VelocityUKF measurement kind");
  const int m = kind == UWVKF_VEL_DVL ? 3 : 1;
  if (!R) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null R");
  if (any_nan(R, m * m)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "measurement covariance contains NaN");
  const double *dz; const uint8_t *dv;
  int rc = stage_meas(h, z, m, valid, memspace, &dz, &dv);
  if (rc) return rc;
  if (kind == UWVKF_VEL_DVL)
    velocity_update_kernel<0><<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, dz, mat_of(R, 9), dv, h->B, h->d_status, policy_of(h));
  else
    velocity_update_kernel<1><<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, dz, mat_of(R, 1), dv, h->B, h->d_status, policy_of(h));
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

int uwvkf_velocity_get_motion_model_state(uwvkf_handle h, double *sim) {
  REQUIRE(h, UWVKF_FILTER_VELOCITY);
  if (!sim) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null output");
  CK(cudaMemcpyAsync(sim, h->d_sim, (size_t)h->B * 10 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

int uwvkf_bottom_set_velocity(uwvkf_handle h, const double *v, int memspace) {
  REQUIRE(h, UWVKF_FILTER_BOTTOM);
  return copy_ctrl(h, h->d_ctrl0, v, 3, memspace);
}

int uwvkf_bottom_update_range(uwvkf_handle h, const double *z, const double *R, const double *unit_direction, const double *origin,
                              const uint8_t *valid, int memspace) {
  REQUIRE(h, UWVKF_FILTER_BOTTOM);
  if (!R || !unit_direction || !origin) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null argument");
  if (std::isnan(R[0])) return fail(UWVKF_ERR_NAN_MEASUREMENT, "measurement covariance contains NaN");
  const double *dz; const uint8_t *dv;
  int rc = stage_meas(h, z, 1, valid, memspace, &dz, &dv);
  if (rc) return rc;
  Vec3A d{{unit_direction[0], unit_direction[1], unit_direction[2]}}, o{{origin[0], origin[1], origin[2]}};
  bottom_range_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, dz, R[0], d, o, dv, h->B, h->d_status, policy_of(h));
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

int uwvkf_bottom_update_normal(uwvkf_handle h, const double *z, const double *R, const uint8_t *valid, int memspace) {
  REQUIRE(h, UWVKF_FILTER_BOTTOM);
  if (!R) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null R");
  const double *dz; const uint8_t *dv;
  int rc = stage_meas(h, z, 3, valid, memspace, &dz, &dv);
  if (rc) return rc;
  bottom_normal_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, dz, mat_of(R, 4), dv, h->B, h->d_status, policy_of(h));
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

int uwvkf_indirect_set_tau(uwvkf_handle h, double tau) {
  REQUIRE(h, UWVKF_FILTER_INDIRECT_POSE);
  if (!(tau > 0)) return fail(UWVKF_ERR_INVALID_ARGUMENT, "orientation_error_tau must be positive");
  h->tau = tau;
  return UWVKF_OK;
}

int uwvkf_indirect_update_pose_reference(uwvkf_handle h, const double *pose_ref, int memspace) {
  REQUIRE(h, UWVKF_FILTER_INDIRECT_POSE);
  return copy_ctrl(h, h->d_ctrl0, pose_ref, 7, memspace);
}

int uwvkf_indirect_update_features(uwvkf_handle h, int nfeat, const double *px, const double *pxcov, const double *feature_pos,
                                   const double *marker_pose, const double *cov_marker, const double *camera,
                                   const double *cam_in_body, const uint8_t *valid, int memspace) {
  REQUIRE(h, UWVKF_FILTER_INDIRECT_POSE);
  if (nfeat < 1 || nfeat > 8) return fail(UWVKF_ERR_INVALID_ARGUMENT, "between 1 and 8 features per call");
  if (!pxcov || !feature_pos || !marker_pose || !cov_marker || !camera || !cam_in_body) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null argument");
  if (any_nan(pxcov, nfeat * 4)) return fail(UWVKF_ERR_NAN_MEASUREMENT, "measurement covariance contains NaN");
  const double *dz; const uint8_t *dv;
  int rc = stage_meas(h, px, nfeat * 2, valid, memspace, &dz, &dv);
  if (rc) return rc;
  FeatArgs A{};
  A.nfeat = nfeat;
  for (int i = 0; i < nfeat; ++i) {
    for (int k = 0; k < 4; ++k) A.pxcov[i][k] = pxcov[i * 4 + k];
    for (int k = 0; k < 3; ++k) A.fpos[i][k] = feature_pos[i * 3 + k];
  }
  std::memcpy(A.marker, marker_pose, sizeof A.marker);
  std::memcpy(A.covm, cov_marker, sizeof A.covm);
  std::memcpy(A.cam, camera, sizeof A.cam);
  std::memcpy(A.cib, cam_in_body, sizeof A.cib);
  indirect_features_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, dz, h->d_ctrl0, A, dv, h->B, h->d_status, policy_of(h));
  CK(cudaGetLastError());
  h->launches++;
  return UWVKF_OK;
}

int uwvkf_indirect_get_corrected_pose(uwvkf_handle h, double *out) {
  REQUIRE(h, UWVKF_FILTER_INDIRECT_POSE);
  if (!out) return fail(UWVKF_ERR_INVALID_ARGUMENT, "null output");
  int rc = ensure_tmp(h, (size_t)h->B * 7 * sizeof(double));
  if (rc) return rc;
  indirect_corrected_pose_kernel<<<grid_for(h->B), kThreads, 0, h->stream>>>(h->d_state, h->d_ctrl0, h->d_tmp, h->B);
  CK(cudaGetLastError());
  h->launches++;
  CK(cudaMemcpyAsync(out, h->d_tmp, (size_t)h->B * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return UWVKF_OK;
}

}  // extern "C"
