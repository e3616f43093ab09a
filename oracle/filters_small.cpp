// ORACLE (test infrastructure, NOT product code) -- parity unpinned.
//
// CPU restatement of VelocityUKF, BottomUKF and IndirectPoseUKF of the reference on the ukfom recipe of
// ukfom.hpp (literal: all 2n+1 sigma points).  Reference lines are cited per function
// (paths relative to /root/reference).
#include <cmath>
#include <cstring>

#include "dynamics.hpp"
#include "oracle_api.h"
#include "ukfom.hpp"

using namespace orc;

namespace {

Policy make_policy(const double *pol) {
  Policy p;
  if (pol) { p.mean_tol = pol[0]; p.mean_max_it = (int)pol[1]; }
  return p;
}
bool gate_any(double) { return true; }
bool has_nan(const double *v, int n) { for (int i = 0; i < n; ++i) if (std::isnan(v[i])) return true; return false; }

// ---- VelocityUKF ---------------------------------------------------------------------------------------
// VelocityState (src/VelocityUKF.hpp:24-27): velocity(3), z_position(1)
Layout velocity_layout() { return Layout{{{VECT, 3, 3}, {VECT, 1, 1}}}; }

// src/VelocityUKF.cpp:6-33  processMotionModel
void velocity_process_model(const double *s, double *o, const UwvParams &P, const double q[4], const double w[3],
                            const double tau[6], double dt) {
  double vel6[6] = {s[0], s[1], s[2], w[0], w[1], w[2]}, nv[6];
  sim_step(P, tau, q, vel6, dt, nv);                       // :18-21 sendEffort(body_efforts, model_state)
  double dv[3] = {nv[0] - s[0], nv[1] - s[1], nv[2] - s[2]};  // :24
  for (int k = 0; k < 3; ++k) o[k] = s[k] + dv[k];         // :26 boxplus(velocity_delta)
  double vn[3];
  quat_rot(q, o, vn);                                       // :30 (orientation * new_state.velocity).z()
  o[3] = s[3] + dt * vn[2];                                 // :31
}

// ---- BottomUKF -----------------------------------------------------------------------------------------
// BottomState (src/BottomUKF.hpp:18-21): distance (Scalar), normal (S2)
Layout bottom_layout() { return Layout{{{VECT, 1, 1}, {S2, 2, 3}}}; }

// ---- IndirectPoseUKF -----------------------------------------------------------------------------------
// PoseErrorState (src/IndirectPoseUKF.hpp:19-22) and FilterStateWithMarker (src/IndirectPoseUKF.cpp:26-30)
Layout pose_error_layout() { return Layout{{{VECT, 3, 3}, {SO3, 3, 4}}}; }
Layout pose_error_marker_layout() { return Layout{{{VECT, 3, 3}, {SO3, 3, 4}, {VECT, 3, 3}, {SO3, 3, 4}}}; }

// rigid transforms as (t[3], q[4]) = 7 doubles; compose c = a * b
void xf_mul(const double *a, const double *b, double *c) {
  double t[3], q[4];
  quat_rot(a + 3, b, t);
  for (int k = 0; k < 3; ++k) t[k] += a[k];
  quat_mul(a + 3, b + 3, q);
  std::memcpy(c, t, sizeof t);
  std::memcpy(c + 3, q, sizeof q);
}
// p_out = a^-1 * p
void xf_inv_apply(const double *a, const double *p, double *o) {
  double d[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
  quat_rot_inv(a + 3, d, o);
}

// src/IndirectPoseUKF.cpp:36-50  measurementVisualLandmark
void indirect_landmark(const double *s /*14*/, double *z /*3*/, const double *feature_pos, const double *cam_in_body,
                       const double *body_in_nav) {
  double bn_err[7], cam_in_nav[7];
  xf_mul(body_in_nav, s, bn_err);          // body_in_nav * pose_error ; pose_error = (position_error, orientation_error) = s[0..6]
  xf_mul(bn_err, cam_in_body, cam_in_nav);
  double fp[3];
  quat_rot(s + 10, feature_pos, fp);       // marker_orientation * feature_pos
  for (int k = 0; k < 3; ++k) fp[k] += s[7 + k];
  xf_inv_apply(cam_in_nav, fp, z);
  s2_normalize(z);                          // MTK::S2<double>(vec) normalises
}

}  // namespace

extern "C" {

// src/VelocityUKF.cpp:58-77  setupMotionModel: the private simulator starts at identity orientation, the current
// angular-velocity measurement and the filter's current velocity.  sim[10] = q(x,y,z,w), linear(3), angular(3)
void orc_vel_setup_motion_model(const double *mu, const double *gyro, double *sim) {
  sim[0] = sim[1] = sim[2] = 0; sim[3] = 1;
  for (int k = 0; k < 3; ++k) { sim[4 + k] = mu[k]; sim[7 + k] = gyro[k]; }
}

// src/VelocityUKF.cpp:114-130  predictionStepImpl
int orc_vel_predict(double *mu, double *cov, const double *Q, double *sim, const double *gyro, const double *efforts,
                    const double *uwv, double dt, const double *policy) {
  UwvParams P;
  std::memcpy(&P, uwv, sizeof P);
  Ukf ukf(velocity_layout(), mu, cov, make_policy(policy));
  double Qs[16];
  for (int i = 0; i < 16; ++i) Qs[i] = dt * Q[i];                                     // :123 delta * process_noise_cov
  const double *q = sim;                                                              // :121 model_state.orientation
  ukf.predict([&](const double *x, double *o) { velocity_process_model(x, o, P, q, gyro, efforts, dt); }, Qs);
  if (!(ukf.status() & ST_NOT_SPD)) {
    std::memcpy(mu, ukf.mu().data(), 4 * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), 16 * sizeof(double));
  }
  // :126-127 the private motion model is advanced with the same efforts (DYNAMIC: velocities only)
  double nv[6];
  sim_step(P, efforts, sim, sim + 4, dt, nv);
  std::memcpy(sim + 4, nv, sizeof nv);
  return (int)ukf.status();
}

// src/VelocityUKF.cpp:79-85 (DVL, kind 0, m=3) and :106-112 (pressure, kind 1, m=1)
int orc_vel_update(int kind, double *mu, double *cov, const double *z, const double *R, const double *policy) {
  const int m = kind == 0 ? 3 : 1;
  if (has_nan(z, m) || has_nan(R, m * m)) return ST_NAN_MEAS;
  Ukf ukf(velocity_layout(), mu, cov, make_policy(policy));
  if (kind == 0)
    ukf.update(vect_layout(3), z, [](const double *x, double *o) { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; }, R, gate_any);  // :35-40
  else
    ukf.update(vect_layout(1), z, [](const double *x, double *o) { o[0] = x[3]; }, R, gate_any);                           // :42-47
  if (!(ukf.status() & ST_NOT_SPD)) {
    std::memcpy(mu, ukf.mu().data(), 4 * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), 16 * sizeof(double));
  }
  return (int)ukf.status();
}

// src/BottomUKF.cpp:48-54  predictionStepImpl with processModel :5-16
int orc_bottom_predict(double *mu, double *cov, const double *Q, const double *velocity, double dt, const double *policy) {
  Ukf ukf(bottom_layout(), mu, cov, make_policy(policy));
  double nxy = std::sqrt(velocity[0] * velocity[0] + velocity[1] * velocity[1]);
  double scale = std::pow(nxy, 2.) * std::pow(dt, 2.), Qs[9];                          // :50
  for (int i = 0; i < 9; ++i) Qs[i] = scale * Q[i];
  const double vz = velocity[2];
  ukf.predict([&](const double *x, double *o) {
    std::memcpy(o, x, 4 * sizeof(double));
    o[0] = x[0] + dt * (-1. * vz);                                                     // :10-12
  }, Qs);
  if (!(ukf.status() & ST_NOT_SPD)) {
    std::memcpy(mu, ukf.mu().data(), 4 * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), 9 * sizeof(double));
  }
  return (int)ukf.status();
}

// src/BottomUKF.cpp:56-61 with measurementDistance :18-30
int orc_bottom_update_range(double *mu, double *cov, const double *z, const double *R, const double *unit_direction,
                            const double *origin, const double *policy) {
  if (has_nan(z, 1) || has_nan(R, 1)) return ST_NAN_MEAS;
  Ukf ukf(bottom_layout(), mu, cov, make_policy(policy));
  ukf.update(vect_layout(1), z, [&](const double *x, double *o) {
    double bottom[3] = {0. - origin[0], 0. - origin[1], -x[0] - origin[2]};
    const double *n = x + 1;
    double v = unit_direction[0] * n[0] + unit_direction[1] * n[1] + unit_direction[2] * n[2];
    o[0] = (v != 0.) ? (bottom[0] * n[0] + bottom[1] * n[1] + bottom[2] * n[2]) / v : 0.;
  }, R, gate_any);
  if (!(ukf.status() & ST_NOT_SPD)) {
    std::memcpy(mu, ukf.mu().data(), 4 * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), 9 * sizeof(double));
  }
  return (int)ukf.status();
}

// src/BottomUKF.cpp:63-67 with measurementNormal :32-37 (measurement lives on S2)
int orc_bottom_update_normal(double *mu, double *cov, const double *z, const double *R, const double *policy) {
  if (has_nan(z, 3) || has_nan(R, 4)) return ST_NAN_MEAS;
  Ukf ukf(bottom_layout(), mu, cov, make_policy(policy));
  Layout s2l{{{S2, 2, 3}}};
  ukf.update(s2l, z, [](const double *x, double *o) { o[0] = x[1]; o[1] = x[2]; o[2] = x[3]; }, R, gate_any);
  if (!(ukf.status() & ST_NOT_SPD)) {
    std::memcpy(mu, ukf.mu().data(), 4 * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), 9 * sizeof(double));
  }
  return (int)ukf.status();
}

// src/IndirectPoseUKF.cpp:80-92 predictionStepImpl with processModel :7-20
int orc_indirect_predict(double *mu, double *cov, const double *Q, double tau, double dt, const double *policy) {
  Ukf ukf(pose_error_layout(), mu, cov, make_policy(policy));
  double rot[9], Qe[36], t[9];
  quat_to_mat(mu + 3, rot);                                                           // :82
  std::memcpy(Qe, Q, sizeof Qe);
  const double f = 2. / (tau * dt);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {                          // :85-86
    double s = 0; for (int k = 0; k < 3; ++k) s += rot[i * 3 + k] * (f * Q[(3 + k) * 6 + 3 + j]);
    t[i * 3 + j] = s;
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0; for (int k = 0; k < 3; ++k) s += t[i * 3 + k] * rot[j * 3 + k];
    Qe[(3 + i) * 6 + 3 + j] = s;
  }
  const double dt2 = std::pow(dt, 2.);                                                // :88
  for (double &v : Qe) v *= dt2;
  ukf.predict([&](const double *x, double *o) {
    std::memcpy(o, x, 7 * sizeof(double));
    double lg[3];
    so3_log(x + 3, lg);                                                               // :15
    for (int k = 0; k < 3; ++k) lg[k] = (-1.0 / tau) * lg[k];                         // :16
    so3_boxplus(x + 3, lg, dt, o + 3);                                                // :17
  }, Qe);
  if (!(ukf.status() & ST_NOT_SPD)) {
    std::memcpy(mu, ukf.mu().data(), 7 * sizeof(double));
    std::memcpy(cov, ukf.sigma().data(), 36 * sizeof(double));
  }
  return (int)ukf.status();
}

// src/IndirectPoseUKF.cpp:94-135  integrateMeasurement(marker features...)
// px[nfeat][2], pxcov[nfeat][4], feature_pos[nfeat][3]; poses are (t[3], q[x,y,z,w]); camera = fx, fy, cx, cy
int orc_indirect_update_features(double *mu, double *cov, int nfeat, const double *px, const double *pxcov,
                                 const double *feature_pos, const double *marker_pose, const double *cov_marker,
                                 const double *camera, const double *cam_in_body, const double *pose_ref, const double *policy) {
  double amu[14], acov[144];
  std::memcpy(amu, mu, 7 * sizeof(double));                                           // :100-103
  std::memcpy(amu + 7, marker_pose, 7 * sizeof(double));
  std::memset(acov, 0, sizeof acov);
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {                          // :104-106
    acov[i * 12 + j] = cov[i * 6 + j];
    acov[(6 + i) * 12 + 6 + j] = cov_marker[i * 6 + j];
  }
  Ukf ukf(pose_error_marker_layout(), amu, acov, make_policy(policy));
  const double fx = camera[0], fy = camera[1], cx = camera[2], cy = camera[3];
  const double fx2 = std::pow(fx, 2.), fy2 = std::pow(fy, 2.), fxy = fx * fy;         // :109-111
  Layout s2l{{{S2, 2, 3}}};
  int status = 0;
  for (int i = 0; i < nfeat; ++i) {                                                   // :114
    if (has_nan(px + 2 * i, 2) || has_nan(pxcov + 4 * i, 4)) return ST_NAN_MEAS;      // :116
    double z[3] = {(px[2 * i] - cx) / fx, (px[2 * i + 1] - cy) / fy, 1.0};            // :119-121
    s2_normalize(z);
    double R[4] = {pxcov[4 * i] / fx2, pxcov[4 * i + 1] / fxy, pxcov[4 * i + 2] / fxy, pxcov[4 * i + 3] / fy2};  // :122-124
    const double *fp = feature_pos + 3 * i;
    ukf.update(s2l, z, [&](const double *x, double *o) { indirect_landmark(x, o, fp, cam_in_body, pose_ref); }, R, gate_any);  // :126
    status |= (int)ukf.status();
  }
  if (!(status & ST_NOT_SPD)) {                                                       // :134
    std::memcpy(mu, ukf.mu().data(), 7 * sizeof(double));
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) cov[i * 6 + j] = ukf.sigma()[i * 12 + j];
  }
  return status;
}

// src/IndirectPoseUKF.cpp:137-142  getCorrectedPose = pose_ref * pose_error
void orc_indirect_corrected_pose(const double *mu, const double *pose_ref, double *out) { xf_mul(pose_ref, mu, out); }

}  // extern "C"
