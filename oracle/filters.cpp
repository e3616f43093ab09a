// ORACLE (test infrastructure, NOT product code) -- parity unpinned.
//
// Line-by-line CPU restatement of the four filters of the reference on top of the ukfom recipe in
// ukfom.hpp.  Every function cites the reference lines it follows (paths relative to /root/reference).
// The arithmetic of the reference lives in dependencies that are absent here (slam/mtk,
// slam/pose_estimation, control/uwv_dynamic_model, Eigen, LAPACK -- manifest.xml:10-13, no versions
// pinned) and the reference ships no tests or golden vectors (test/suite.cpp:1-6), hence "parity
// unpinned": what pins this oracle is listed in DESIGN.md (linear-KF known answer, LAPACK dpotrf,
// manifold identities, high-precision replays).
//
// Exposed as a flat C API (ctypes) -- see oracle/oracle_api.h.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "dynamics.hpp"
#include "geo.hpp"
#include "oracle_api.h"
#include "ukfom.hpp"

using namespace orc;

namespace {

// ---- PoseState (src/PoseState.hpp:29-45): stored layout, 54 doubles --------------------------------
enum PoseIdx {
  P_POS = 0, P_ORI = 3, P_VEL = 7, P_ACC = 10, P_BG = 13, P_BA = 16, P_G = 19, P_INERTIA = 20,
  P_LIN = 29, P_QUAD = 38, P_WV = 47, P_WVB = 49, P_BADCP = 51, P_RHO = 53, P_DIM = 54, P_DOF = 53
};
// tangent offsets
enum PoseTan { T_POS = 0, T_ORI = 3, T_VEL = 6, T_ACC = 9, T_BG = 12, T_BA = 15, T_G = 18, T_INERTIA = 19,
               T_LIN = 28, T_QUAD = 37, T_WV = 46, T_WVB = 48, T_BADCP = 50, T_RHO = 52 };

Layout pose_layout() {
  return Layout{{{VECT, 3, 3}, {SO3, 3, 4}, {VECT, 3, 3}, {VECT, 3, 3}, {VECT, 3, 3}, {VECT, 3, 3}, {VECT, 1, 1},
                 {VECT, 9, 9}, {VECT, 9, 9}, {VECT, 9, 9}, {VECT, 2, 2}, {VECT, 2, 2}, {VECT, 2, 2}, {VECT, 1, 1}}};
}

// PoseUKF::PoseUKFParameter (src/PoseUKF.hpp:46-76) + projection origin, flat (see oracle_api.h)
struct PoseParams {
  double lat0, lon0;
  double gyro_bias_tau, gyro_bias_offset[3];
  double acc_bias_tau, acc_bias_offset[3];
  double inertia_tau, lin_damping_tau, quad_damping_tau;
  double water_velocity_tau, water_velocity_limits, water_velocity_scale;
  double adcp_bias_tau, atmospheric_pressure, water_density_tau;
  double imu_in_body[3];
};
static_assert(sizeof(PoseParams) == ORC_POSE_NPARAMS * sizeof(double), "flat layout");

Policy make_policy(const double *pol) {
  Policy p;
  if (pol) { p.mean_tol = pol[0]; p.mean_max_it = (int)pol[1]; }
  return p;
}

// src/PoseUKF.cpp:12-84  processModel
void pose_process_model(const double *s, double *o, const double w_meas[3], const TMerc &proj,
                        const double *offsets /*28*/, const PoseParams &fp, double dt) {
  std::memcpy(o, s, P_DIM * sizeof(double));
  // :25 position (+)= velocity * dt
  for (int k = 0; k < 3; ++k) o[P_POS + k] = s[P_POS + k] + dt * s[P_VEL + k];
  // :28-32 angular velocity
  double lat, lon;
  proj.navToWorld(s[P_POS + 0], s[P_POS + 1], lat, lon);
  double earth[3] = {kEarthW * std::cos(lat), 0., kEarthW * std::sin(lat)};
  double wb[3] = {w_meas[0] - s[P_BG + 0], w_meas[1] - s[P_BG + 1], w_meas[2] - s[P_BG + 2]};
  double wn[3];
  quat_rot(s + P_ORI, wb, wn);
  for (int k = 0; k < 3; ++k) wn[k] -= earth[k];
  so3_boxplus(s + P_ORI, wn, dt, o + P_ORI);
  // :35 velocity (+)= acceleration * dt
  for (int k = 0; k < 3; ++k) o[P_VEL + k] = s[P_VEL + k] + dt * s[P_ACC + k];
  // :37-43 bias first-order Markov pulls
  for (int k = 0; k < 3; ++k) {
    double d = (-1.0 / fp.gyro_bias_tau) * (s[P_BG + k] - fp.gyro_bias_offset[k]);
    o[P_BG + k] = s[P_BG + k] + dt * d;
    double e = (-1.0 / fp.acc_bias_tau) * (s[P_BA + k] - fp.acc_bias_offset[k]);
    o[P_BA + k] = s[P_BA + k] + dt * e;
  }
  // :45-55 inertia / damping pulls towards their initial values
  for (int k = 0; k < 9; ++k) {
    double d = (-1.0 / fp.inertia_tau) * (s[P_INERTIA + k] - offsets[k]);
    o[P_INERTIA + k] = s[P_INERTIA + k] + dt * d;
    double e = (-1.0 / fp.lin_damping_tau) * (s[P_LIN + k] - offsets[9 + k]);
    o[P_LIN + k] = s[P_LIN + k] + dt * e;
    double f = (-1.0 / fp.quad_damping_tau) * (s[P_QUAD + k] - offsets[18 + k]);
    o[P_QUAD + k] = s[P_QUAD + k] + dt * f;
  }
  // :64-74 water velocities and ADCP bias decay to zero
  for (int k = 0; k < 2; ++k) {
    o[P_WV + k] = s[P_WV + k] + dt * ((-1.0 / fp.water_velocity_tau) * s[P_WV + k]);
    o[P_WVB + k] = s[P_WVB + k] + dt * ((-1.0 / fp.water_velocity_tau) * s[P_WVB + k]);
    o[P_BADCP + k] = s[P_BADCP + k] + dt * ((-1.0 / fp.adcp_bias_tau) * s[P_BADCP + k]);
  }
  // :76-78 water density
  o[P_RHO] = s[P_RHO] + dt * ((-1.0 / fp.water_density_tau) * (s[P_RHO] - offsets[27]));
}

// src/PoseUKF.cpp:446-460  effective process noise
void pose_process_noise(const double *mu, const double *Q, const PoseParams &fp, double dt, std::vector<double> &Qe) {
  const int n = P_DOF;
  Qe.assign(Q, Q + n * n);
  double rot[9];
  quat_to_mat(mu + P_ORI, rot);
  // :451 rot * Q_ori * rot^T
  double t[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0; for (int k = 0; k < 3; ++k) s += rot[i * 3 + k] * Q[(T_ORI + k) * n + T_ORI + j];
    t[i * 3 + j] = s;
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0; for (int k = 0; k < 3; ++k) s += t[i * 3 + k] * rot[j * 3 + k];
    Qe[(T_ORI + i) * n + T_ORI + j] = s;
  }
  // :453-458
  double sv[3] = {mu[P_VEL], mu[P_VEL + 1], 10 * mu[P_VEL + 2]};
  double add = fp.water_velocity_scale * (sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2]) * dt;
  for (int k = 0; k < 2; ++k) {
    Qe[(T_WV + k) * n + T_WV + k] += add;
    Qe[(T_WVB + k) * n + T_WVB + k] += add;
  }
  // :460
  double dt2 = std::pow(dt, 2.);
  for (auto &v : Qe) v *= dt2;
}

// measurement models, src/PoseUKF.cpp:87-151
void meas_xy(const double *s, double *z) { z[0] = s[P_POS]; z[1] = s[P_POS + 1]; }            // :87-92
void meas_z(const double *s, double *z) { z[0] = s[P_POS + 2]; }                               // :100-105
void meas_pressure(const double *s, double *z, const double r[3], double p_atm) {              // :107-115
  double rn[3];
  quat_rot(s + P_ORI, r, rn);
  double zn = s[P_POS + 2] + rn[2];
  z[0] = p_atm - zn * s[P_G] * s[P_RHO];
}
void meas_velocity(const double *s, double *z) { quat_rot_inv(s + P_ORI, s + P_VEL, z); }      // :117-123
void meas_acceleration(const double *s, double *z) {                                            // :125-131
  double a[3] = {s[P_ACC], s[P_ACC + 1], s[P_ACC + 2] + s[P_G]};
  quat_rot_inv(s + P_ORI, a, z);
  for (int k = 0; k < 3; ++k) z[k] += s[P_BA + k];
}
void meas_water_currents(const double *s, double *z, double w) {                                // :133-151
  double below[3] = {s[P_VEL] - s[P_WVB], s[P_VEL + 1] - s[P_WVB + 1], s[P_VEL + 2] - 0};
  double around[3] = {s[P_VEL] - s[P_WV], s[P_VEL + 1] - s[P_WV + 1], s[P_VEL + 2] - 0};
  double b[3], a[3];
  quat_rot_inv(s + P_ORI, below, b);
  quat_rot_inv(s + P_ORI, around, a);
  z[0] = w * b[0] + (1 - w) * a[0] + s[P_BADCP];
  z[1] = w * b[1] + (1 - w) * a[1] + s[P_BADCP + 1];
}

// 3x3 state blocks <-> 6x6 model matrices: rows/cols {0,1,5}  (src/PoseUKF.cpp:160-171).
// The state blocks are MTK::matrix<3,3>, vectorised column-major (src/PoseState.hpp:37).
void state_blocks_into_params(const double *s, UwvParams &P) {
  const int map[3] = {0, 1, 5};
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    P.inertia[map[r] * 6 + map[c]] = s[P_INERTIA + c * 3 + r];
    P.damp_lin[map[r] * 6 + map[c]] = s[P_LIN + c * 3 + r];
    P.damp_quad[map[r] * 6 + map[c]] = s[P_QUAD + c * 3 + r];
  }
}
// src/PoseUKF.cpp:153-196  measurementEfforts
void meas_efforts(const double *s, double *z, const UwvParams &base, const double imu_in_body[3], const double wb[3]) {
  UwvParams P = base;
  state_blocks_into_params(s, P);
  double wv[3] = {s[P_WV], s[P_WV + 1], 0};
  double vb[3], wvb[3], c[3], cc[3], ab[3];
  quat_rot_inv(s + P_ORI, s + P_VEL, vb);
  cross3(wb, imu_in_body, c);
  quat_rot_inv(s + P_ORI, wv, wvb);
  for (int k = 0; k < 3; ++k) vb[k] = vb[k] - c[k] - wvb[k];
  double vel6[6] = {vb[0], vb[1], vb[2], wb[0], wb[1], wb[2]};
  quat_rot_inv(s + P_ORI, s + P_ACC, ab);
  cross3(wb, c, cc);
  double acc6[6] = {ab[0] - cc[0], ab[1] - cc[1], ab[2] - cc[2], 0, 0, 0};
  calc_efforts(P, acc6, vel6, s + P_ORI, z);
}
// src/PoseUKF.cpp:198-219  constrainVelocity
void meas_constrain_velocity(const double *s, double *z, const UwvParams &P, const double imu_in_body[3], const double wb[3],
                             const double water_velocity[3], const double q[4], const double acc_body[3]) {
  double vb[3], c[3], wvb[3];
  quat_rot_inv(q, s + P_VEL, vb);
  cross3(wb, imu_in_body, c);
  quat_rot_inv(q, water_velocity, wvb);
  for (int k = 0; k < 3; ++k) vb[k] = vb[k] - c[k] - wvb[k];
  double vel6[6] = {vb[0], vb[1], vb[2], wb[0], wb[1], wb[2]};
  double acc6[6] = {acc_body[0], acc_body[1], acc_body[2], 0, 0, 0};
  calc_efforts(P, acc6, vel6, q, z);
}

// PoseStateWithMarker (src/PoseUKF.cpp:225-228): filter_state (53 DOF), marker_position, marker_orientation -> 59 DOF, 61 stored
Layout pose_marker_layout() {
  Layout l = pose_layout();
  l.comps.push_back({VECT, 3, 3});
  l.comps.push_back({SO3, 3, 4});
  return l;
}
// src/PoseUKF.cpp:233-243  measurementVisualLandmark: the feature seen from the camera, as a point on S2.
// cam_in_imu = (t[3], q[4]); nav_in_cam = (imu_in_nav * cam_in_imu)^-1
void pose_landmark(const double *s /*61*/, double *z /*3*/, const double *feature_pos, const double *cam_in_imu) {
  double cam_t[3], cam_q[4], fp[3], d[3];
  quat_rot(s + P_ORI, cam_in_imu, cam_t);                       // imu_in_nav * cam_in_imu
  for (int k = 0; k < 3; ++k) cam_t[k] += s[P_POS + k];
  quat_mul(s + P_ORI, cam_in_imu + 3, cam_q);
  quat_rot(s + P_DIM + 3, feature_pos, fp);                     // marker_orientation * feature_pos + marker_position
  for (int k = 0; k < 3; ++k) d[k] = fp[k] + s[P_DIM + k] - cam_t[k];
  quat_rot_inv(cam_q, d, z);
  s2_normalize(z);                                               // MTK::S2<double>(vec) normalises
}

bool gate_any(double) { return true; }                 // ukfom::accept_any_mahalanobis_distance
bool gate_d2p95(double d2) { return !(d2 > 5.991); }   // src/PoseUKF.cpp:275-286

bool has_nan(const double *v, int n) { for (int i = 0; i < n; ++i) if (std::isnan(v[i])) return true; return false; }

// src/PoseUKF.cpp:693-699 getRotationRate
void pose_rotation_rate(const double *mu, const double w_meas[3], const TMerc &proj, double out[3]) {
  double lat, lon;
  proj.navToWorld(mu[P_POS], mu[P_POS + 1], lat, lon);
  double earth[3] = {kEarthW * std::cos(lat), 0., kEarthW * std::sin(lat)}, eb[3];
  quat_rot_inv(mu + P_ORI, earth, eb);
  for (int k = 0; k < 3; ++k) out[k] = w_meas[k] - mu[P_BG + k] - eb[k];
}

int pose_predict_impl(Ukf &ukf, double *mu, double *cov, const double *Q, const double *w, const PoseParams &fp, const TMerc &proj,
                      const double *offsets, double dt, int *mean_it) {
  ukf.reset(mu, cov);
  static thread_local std::vector<double> Qe;
  pose_process_noise(mu, Q, fp, dt, Qe);                                                   // :448-460
  ukf.predict([&](const double *x, double *o) { pose_process_model(x, o, w, proj, offsets, fp, dt); }, Qe.data());  // :462
  if (!(ukf.status() & ST_NOT_SPD)) {
    std::memcpy(mu, ukf.mu().data(), P_DIM * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), P_DOF * P_DOF * sizeof(double));
  }
  if (mean_it) *mean_it = ukf.last_mean_iterations();
  return (int)ukf.status();
}

int pose_update_impl(Ukf &ukf, int kind, double *mu, double *cov, const double *z_in, const double *R, const double *aux,
                     const double *w_meas, const PoseParams &fp, const TMerc &proj, double *d2_out) {
  int m = orc_pose_meas_dim(kind);
  if (m <= 0) return -1;
  if (has_nan(z_in, m) || has_nan(R, m * m)) return ST_NAN_MEAS;   // checkMeasurment [X4]
  ukf.reset(mu, cov);
  Layout zl = vect_layout(m);
  double z[6];
  std::memcpy(z, z_in, m * sizeof(double));
  switch (kind) {
    case ORC_POSE_XY:   // src/PoseUKF.cpp:506-512
      ukf.update(zl, z, [](const double *x, double *o) { meas_xy(x, o); }, R, gate_any);
      break;
    case ORC_POSE_Z:    // :498-504
      ukf.update(zl, z, [](const double *x, double *o) { meas_z(x, o); }, R, gate_any);
      break;
    case ORC_POSE_PRESSURE: {   // :559-565, aux = pressure_sensor_in_imu[3]
      double r[3] = {aux ? aux[0] : 0, aux ? aux[1] : 0, aux ? aux[2] : 0};
      ukf.update(zl, z, [&](const double *x, double *o) { meas_pressure(x, o, r, fp.atmospheric_pressure); }, R, gate_any);
      break;
    }
    case ORC_POSE_ACCELERATION:  // :484-490
      ukf.update(zl, z, [](const double *x, double *o) { meas_acceleration(x, o); }, R, gate_any);
      break;
    case ORC_POSE_VELOCITY:      // :476-482
      ukf.update(zl, z, [](const double *x, double *o) { meas_velocity(x, o); }, R, gate_any);
      break;
    case ORC_POSE_WATER_VELOCITY: {  // :604-611, aux = cell_weighting
      double wgt = aux[0];
      ukf.update(zl, z, [&](const double *x, double *o) { meas_water_currents(x, o, wgt); }, R, gate_d2p95);
      break;
    }
    case ORC_POSE_GEOGRAPHIC: {  // :567-579, z = (lat, lon), aux = gps_in_body[3]
      double px, py, lever[3], g[3] = {aux ? aux[0] : 0, aux ? aux[1] : 0, aux ? aux[2] : 0};
      proj.worldToNav(z_in[0], z_in[1], px, py);
      quat_rot(mu + P_ORI, g, lever);
      z[0] = px - lever[0]; z[1] = py - lever[1];
      ukf.update(zl, z, [](const double *x, double *o) { meas_xy(x, o); }, R, gate_d2p95);
      break;
    }
    case ORC_POSE_DELAYED_XY: {  // :514-527, aux = delayed_position[2]
      z[0] = z_in[0] + (mu[P_POS] - aux[0]);
      z[1] = z_in[1] + (mu[P_POS + 1] - aux[1]);
      ukf.update(zl, z, [](const double *x, double *o) { meas_xy(x, o); }, R, gate_any);
      break;
    }
    case ORC_POSE_BODY_EFFORTS:
    case ORC_POSE_BODY_EFFORTS_VELOCITY_ONLY: {  // :581-602, aux = UwvParams (flat)
      UwvParams P;
      std::memcpy(&P, aux, sizeof(P));
      double wb[3];
      pose_rotation_rate(mu, w_meas, proj, wb);
      if (kind == ORC_POSE_BODY_EFFORTS_VELOCITY_ONLY) {
        double wv[3] = {mu[P_WV], mu[P_WV + 1], 0.}, q[4], ab[3], c[3], cc[3];
        std::memcpy(q, mu + P_ORI, sizeof q);
        quat_rot_inv(q, mu + P_ACC, ab);
        cross3(wb, fp.imu_in_body, c);
        cross3(wb, c, cc);
        for (int k = 0; k < 3; ++k) ab[k] -= cc[k];
        ukf.update(zl, z, [&](const double *x, double *o) { meas_constrain_velocity(x, o, P, fp.imu_in_body, wb, wv, q, ab); }, R, gate_any);
      } else {
        ukf.update(zl, z, [&](const double *x, double *o) { meas_efforts(x, o, P, fp.imu_in_body, wb); }, R, gate_any);
      }
      break;
    }
    default: return -1;
  }
  if (d2_out) *d2_out = ukf.last_mahalanobis2();
  if (!(ukf.status() & (ST_NOT_SPD | ST_GATE_REJECTED))) {
    std::memcpy(mu, ukf.mu().data(), P_DIM * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), P_DOF * P_DOF * sizeof(double));
  }
  return (int)ukf.status();
}

}  // namespace

// =====================================================================================================
extern "C" {

int orc_pose_meas_dim(int kind) {
  switch (kind) {
    case ORC_POSE_XY: case ORC_POSE_GEOGRAPHIC: case ORC_POSE_DELAYED_XY: case ORC_POSE_WATER_VELOCITY: return 2;
    case ORC_POSE_Z: case ORC_POSE_PRESSURE: return 1;
    case ORC_POSE_ACCELERATION: case ORC_POSE_VELOCITY: return 3;
    case ORC_POSE_BODY_EFFORTS: case ORC_POSE_BODY_EFFORTS_VELOCITY_ONLY: return 6;
  }
  return -1;
}

// src/PoseUKF.cpp:613-654  integrateMeasurement(marker_corners, feature_positions, marker_pose, cov_marker_pose, camera_config,
// camera_in_IMU): augment with the marker pose, one S2 update per feature, keep the leading 53 x 53 block.
// px[nfeat][2], pxcov[nfeat][4], feature_pos[nfeat][3], marker_pose = t[3] q[4], cov_marker[36], camera = fx fy cx cy, cam_in_imu = t q
int orc_pose_update_features(double *mu, double *cov, int nfeat, const double *px, const double *pxcov, const double *feature_pos,
                             const double *marker_pose, const double *cov_marker, const double *camera, const double *cam_in_imu,
                             const double *policy) {
  const int NA = P_DOF + 6;
  std::vector<double> amu(P_DIM + 7), acov((size_t)NA * NA, 0.0);
  std::memcpy(amu.data(), mu, P_DIM * sizeof(double));                                  // :619-622
  std::memcpy(amu.data() + P_DIM, marker_pose, 7 * sizeof(double));
  for (int i = 0; i < P_DOF; ++i) for (int j = 0; j < P_DOF; ++j) acov[(size_t)i * NA + j] = cov[i * P_DOF + j];   // :623-625
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) acov[(size_t)(P_DOF + i) * NA + P_DOF + j] = cov_marker[i * 6 + j];
  Ukf ukf(pose_marker_layout(), amu.data(), acov.data(), make_policy(policy));
  const double fx = camera[0], fy = camera[1], cx = camera[2], cy = camera[3];
  const double fx2 = std::pow(fx, 2.), fy2 = std::pow(fy, 2.), fxy = fx * fy;           // :628-630
  Layout s2l{{{S2, 2, 3}}};
  int status = 0;
  for (int i = 0; i < nfeat; ++i) {                                                     // :633
    if (has_nan(px + 2 * i, 2) || has_nan(pxcov + 4 * i, 4)) return ST_NAN_MEAS;        // :635
    double z[3] = {(px[2 * i] - cx) / fx, (px[2 * i + 1] - cy) / fy, 1.0};              // :638-640
    s2_normalize(z);
    double R[4] = {pxcov[4 * i] / fx2, pxcov[4 * i + 1] / fxy, pxcov[4 * i + 2] / fxy, pxcov[4 * i + 3] / fy2};   // :642-644
    const double *fp = feature_pos + 3 * i;
    ukf.update(s2l, z, [&](const double *x, double *o) { pose_landmark(x, o, fp, cam_in_imu); }, R, gate_any);     // :646
    status |= (int)ukf.status();
  }
  if (!(status & ST_NOT_SPD)) {                                                         // :653
    std::memcpy(mu, ukf.mu().data(), P_DIM * sizeof(double));
    for (int i = 0; i < P_DOF; ++i) for (int j = 0; j < P_DOF; ++j) cov[i * P_DOF + j] = ukf.sigma()[(size_t)i * NA + j];
  }
  return status;
}

int orc_pose_predict(double *mu, double *cov, const double *Q, const double *rot_rate, const double *params,
                     const double *offsets, double dt, const double *policy, int *mean_it) {
  PoseParams fp;
  std::memcpy(&fp, params, sizeof fp);
  TMerc proj(fp.lat0, fp.lon0);
  Ukf ukf(pose_layout(), mu, cov, make_policy(policy));
  return pose_predict_impl(ukf, mu, cov, Q, rot_rate, fp, proj, offsets, dt, mean_it);
}

int orc_pose_update(int kind, double *mu, double *cov, const double *z, const double *R, const double *aux,
                    const double *rot_rate, const double *params, const double *policy, double *d2_out) {
  PoseParams fp;
  std::memcpy(&fp, params, sizeof fp);
  TMerc proj(fp.lat0, fp.lon0);
  Ukf ukf(pose_layout(), mu, cov, make_policy(policy));
  return pose_update_impl(ukf, kind, mu, cov, z, R, aux, rot_rate, fp, proj, d2_out);
}

void orc_pose_rotation_rate(const double *mu, const double *rot_rate, const double *params, double *out) {
  PoseParams fp;
  std::memcpy(&fp, params, sizeof fp);
  TMerc proj(fp.lat0, fp.lon0);
  pose_rotation_rate(mu, rot_rate, proj, out);
}

// src/PoseUKF.cpp:288-372 (first constructor): initial state / covariance / offsets from the config.
void orc_pose_init_from_config(const double *pos, const double *pos_cov, const double *quat, const double *rot_cov,
                               const double *cfg, const double *inertia6, const double *lin6, const double *quad6,
                               const double *imu_in_body_rot /*3x3 row-major*/, double *mu, double *cov) {
  const int n = P_DOF;
  std::memset(mu, 0, P_DIM * sizeof(double));
  std::memset(cov, 0, n * n * sizeof(double));
  const double *Rib = imu_in_body_rot;
  auto rot3 = [&](const double *v, double *o) { for (int i = 0; i < 3; ++i) o[i] = Rib[i * 3] * v[0] + Rib[i * 3 + 1] * v[1] + Rib[i * 3 + 2] * v[2]; };
  auto rdr = [&](const double *d, int off) {  // R diag(d^2) R^T into the block at tangent offset off
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      double s = 0; for (int k = 0; k < 3; ++k) s += Rib[i * 3 + k] * d[k] * d[k] * Rib[j * 3 + k];
      cov[(off + i) * n + off + j] = s;
    }
  };
  for (int k = 0; k < 3; ++k) mu[P_POS + k] = pos[k];
  for (int k = 0; k < 4; ++k) mu[P_ORI + k] = quat[k];
  rot3(cfg + ORC_CFG_GYRO_BIAS_OFFSET, mu + P_BG);                     // :298
  rot3(cfg + ORC_CFG_ACC_BIAS_OFFSET, mu + P_BA);                      // :299
  mu[P_G] = gravity_wgs84(cfg[ORC_CFG_LATITUDE], cfg[ORC_CFG_ALTITUDE]);  // :301
  const int map[3] = {0, 1, 5};
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {           // :303-314 (column-major blocks)
    mu[P_INERTIA + c * 3 + r] = inertia6[map[r] * 6 + map[c]];
    mu[P_LIN + c * 3 + r] = lin6[map[r] * 6 + map[c]];
    mu[P_QUAD + c * 3 + r] = quad6[map[r] * 6 + map[c]];
  }
  mu[P_RHO] = cfg[ORC_CFG_WATER_DENSITY];                              // :319
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {           // :324-327
    cov[(T_POS + i) * n + T_POS + j] = pos_cov[i * 3 + j];
    cov[(T_ORI + i) * n + T_ORI + j] = rot_cov[i * 3 + j];
  }
  for (int i = 0; i < 3; ++i) { cov[(T_VEL + i) * n + T_VEL + i] = 1.0; cov[(T_ACC + i) * n + T_ACC + i] = 10.0; }
  rdr(cfg + ORC_CFG_GYRO_BIAS_INSTABILITY, T_BG);                      // :328
  rdr(cfg + ORC_CFG_ACC_BIAS_INSTABILITY, T_BA);                       // :329
  cov[T_G * n + T_G] = std::pow(0.05, 2.);                             // :331
  for (int k = 0; k < 9; ++k) {                                        // :333-335
    cov[(T_INERTIA + k) * n + T_INERTIA + k] = std::pow(cfg[ORC_CFG_INERTIA_INSTABILITY + k], 2);
    cov[(T_LIN + k) * n + T_LIN + k] = std::pow(cfg[ORC_CFG_LIN_DAMPING_INSTABILITY + k], 2);
    cov[(T_QUAD + k) * n + T_QUAD + k] = std::pow(cfg[ORC_CFG_QUAD_DAMPING_INSTABILITY + k], 2);
  }
  for (int k = 0; k < 2; ++k) {                                        // :336-338
    cov[(T_WV + k) * n + T_WV + k] = std::pow(cfg[ORC_CFG_WV_LIMITS], 2);
    cov[(T_WVB + k) * n + T_WVB + k] = std::pow(cfg[ORC_CFG_WV_LIMITS], 2);
    cov[(T_BADCP + k) * n + T_BADCP + k] = std::pow(cfg[ORC_CFG_ADCP_BIAS_LIMITS], 2);
  }
  cov[T_RHO * n + T_RHO] = std::pow(cfg[ORC_CFG_WATER_DENSITY_LIMITS], 2.);  // :340
}

// src/PoseUKF.cpp:393-439  setProcessNoiseFromConfig
void orc_pose_process_noise_from_config(const double *cfg, double T, const double *imu_in_body_rot, double *Q) {
  const int n = P_DOF;
  std::memset(Q, 0, n * n * sizeof(double));
  const double *Rib = imu_in_body_rot;
  auto rdr = [&](const double *d, double scale, int off) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      double s = 0; for (int k = 0; k < 3; ++k) s += Rib[i * 3 + k] * (scale * d[k] * d[k]) * Rib[j * 3 + k];
      Q[(off + i) * n + off + j] = s;
    }
  };
  const double *jerk = cfg + ORC_CFG_MAX_JERK;
  for (int k = 0; k < 3; ++k) {
    double a = (1. / 6.) * 0.25 * jerk[k], b = 0.5 * 0.25 * jerk[k], c = 0.25 * jerk[k];
    Q[(T_POS + k) * n + T_POS + k] = 1.5 * (std::pow(T, 4.0) * a * a);   // :401
    Q[(T_VEL + k) * n + T_VEL + k] = 1.5 * (std::pow(T, 2.0) * b * b);   // :404
    Q[(T_ACC + k) * n + T_ACC + k] = c * c;                              // :407
  }
  rdr(cfg + ORC_CFG_GYRO_RANDOMWALK, 1.0, T_ORI);                                                  // :409
  rdr(cfg + ORC_CFG_GYRO_BIAS_INSTABILITY, 2. / (cfg[ORC_CFG_GYRO_BIAS_TAU] * T), T_BG);          // :410
  rdr(cfg + ORC_CFG_ACC_BIAS_INSTABILITY, 2. / (cfg[ORC_CFG_ACC_BIAS_TAU] * T), T_BA);            // :412
  Q[T_G * n + T_G] = 1.e-12;                                                                       // :415
  for (int k = 0; k < 9; ++k) {                                                                    // :417-422
    Q[(T_INERTIA + k) * n + T_INERTIA + k] = (2. / (cfg[ORC_CFG_INERTIA_TAU] * T)) * std::pow(cfg[ORC_CFG_INERTIA_INSTABILITY + k], 2);
    Q[(T_LIN + k) * n + T_LIN + k] = (2. / (cfg[ORC_CFG_LIN_DAMPING_TAU] * T)) * std::pow(cfg[ORC_CFG_LIN_DAMPING_INSTABILITY + k], 2);
    Q[(T_QUAD + k) * n + T_QUAD + k] = (2. / (cfg[ORC_CFG_QUAD_DAMPING_TAU] * T)) * std::pow(cfg[ORC_CFG_QUAD_DAMPING_INSTABILITY + k], 2);
  }
  for (int k = 0; k < 2; ++k) {                                                                    // :424-431
    Q[(T_WV + k) * n + T_WV + k] = (2. / (cfg[ORC_CFG_WV_TAU] * T)) * std::pow(cfg[ORC_CFG_WV_LIMITS], 2);
    Q[(T_WVB + k) * n + T_WVB + k] = (2. / (cfg[ORC_CFG_WV_TAU] * T)) * std::pow(cfg[ORC_CFG_WV_LIMITS], 2);
    Q[(T_BADCP + k) * n + T_BADCP + k] = (2. / (cfg[ORC_CFG_ADCP_BIAS_TAU] * T)) * std::pow(cfg[ORC_CFG_ADCP_BIAS_LIMITS], 2);
  }
  Q[T_RHO * n + T_RHO] = (2. / (cfg[ORC_CFG_WATER_DENSITY_TAU] * T)) * std::pow(cfg[ORC_CFG_WATER_DENSITY_LIMITS], 2.);  // :434
}

void orc_tmerc_world_to_nav(double lat0, double lon0, double lat, double lon, double *xy) {
  TMerc p(lat0, lon0);
  p.worldToNav(lat, lon, xy[0], xy[1]);
}
void orc_tmerc_nav_to_world(double lat0, double lon0, double x, double y, double *latlon) {
  TMerc p(lat0, lon0);
  p.navToWorld(x, y, latlon[0], latlon[1]);
}
double orc_gravity_wgs84(double lat, double alt) { return gravity_wgs84(lat, alt); }

// Batched predict + optional DVL update over B independent filters (OpenMP over filters): the timed CPU
// baseline ("reference path" on host cores).  Layouts: mu[B][54], cov[B][53*53], rot_rate[B][3], z[B][3].
double orc_pose_step_batch(int B, double *mu, double *cov, const double *Q, const double *rot_rate, const double *z_vel,
                           const double *R_vel, const double *params, const double *offsets, double dt, int nsteps,
                           int nthreads, const double *policy, int *status_out) {
  PoseParams fp;
  std::memcpy(&fp, params, sizeof fp);
  TMerc proj(fp.lat0, fp.lon0);
  Policy pol = make_policy(policy);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
  {
    Ukf ukf(pose_layout(), mu, cov, pol);  // one workspace per thread
#pragma omp for schedule(dynamic, 1)
    for (int f = 0; f < B; ++f) {
      int st = 0;
      for (int s = 0; s < nsteps; ++s) {
        st |= pose_predict_impl(ukf, mu + (size_t)f * P_DIM, cov + (size_t)f * P_DOF * P_DOF, Q, rot_rate + (size_t)f * 3, fp, proj,
                                offsets + (size_t)f * 28, dt, nullptr);
        if (z_vel)
          st |= pose_update_impl(ukf, ORC_POSE_VELOCITY, mu + (size_t)f * P_DIM, cov + (size_t)f * P_DOF * P_DOF,
                                 z_vel + (size_t)f * 3, R_vel, nullptr, rot_rate + (size_t)f * 3, fp, proj, nullptr);
      }
      if (status_out) status_out[f] = st;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- generic vector-state UKF (used by the linear-KF known-answer tests) ------------------------------
// x' = F x + c ; z = H x ; all-vector state of dimension n.
int orc_linear_predict(int n, double *mu, double *cov, const double *F, const double *c, const double *Q) {
  Ukf ukf(vect_layout(n), mu, cov);
  ukf.predict([&](const double *x, double *o) {
    for (int i = 0; i < n; ++i) { double s = c[i]; for (int j = 0; j < n; ++j) s += F[i * n + j] * x[j]; o[i] = s; }
  }, Q);
  std::memcpy(mu, ukf.mu().data(), n * sizeof(double));
  std::memcpy(cov, ukf.sigma().data(), n * n * sizeof(double));
  return (int)ukf.status();
}
int orc_linear_update(int n, int m, double *mu, double *cov, const double *H, const double *z, const double *R) {
  Ukf ukf(vect_layout(n), mu, cov);
  ukf.update(vect_layout(m), z, [&](const double *x, double *o) {
    for (int i = 0; i < m; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += H[i * n + j] * x[j]; o[i] = s; }
  }, R, gate_any);
  std::memcpy(mu, ukf.mu().data(), n * sizeof(double));
  std::memcpy(cov, ukf.sigma().data(), n * n * sizeof(double));
  return (int)ukf.status();
}
int orc_cholesky_lower(int n, const double *A, double *L) { return cholesky_lower(n, A, L) ? 0 : 1; }

// manifold primitives for the identity tests
void orc_so3_boxplus(const double *q, const double *v, double scale, double *out) { so3_boxplus(q, v, scale, out); }
void orc_so3_boxminus(const double *a, const double *b, double *out) { so3_boxminus(a, b, out); }
void orc_s2_boxplus(const double *x, const double *d, double scale, double *out) { s2_boxplus(x, d, scale, out); }
void orc_s2_boxminus(const double *y, const double *x, double *out) { s2_boxminus(y, x, out); }
void orc_quat_rot(const double *q, const double *v, double *out) { quat_rot(q, v, out); }

}  // extern "C"
