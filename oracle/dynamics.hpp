// ORACLE (test infrastructure, NOT product code) -- parity unpinned.
//
// Restatement of control/uwv_dynamic_model (un-vendored, no pinned version): DynamicModel::calcEfforts /
// calcAcceleration and ModelSimulation(DYNAMIC)::sendEffort.  Reference call sites:
// src/VelocityUKF.cpp:18-21 (sendEffort(efforts, state) per sigma point), :60-62 (ModelSimulation(DYNAMIC,
// 0.01, 1)), :126-127 (sendEffort(efforts) on the private simulator); src/PoseUKF.cpp:159-173,192,215
// (calcEfforts with inertia / damping taken from the sigma point).
// Frozen definition: Fossen 6-DOF, body frame, nu = [u v w p q r],
//   tau = M nu_dot + C(M, nu) + D_lin nu + D_quad (|nu| o nu) + g(q),
//   C = [ (M nu)_lin x nu_ang ; (M nu)_lin x nu_lin + (M nu)_ang x nu_ang ],
//   g(q) = [ q^-1 (0,0,W-B) ; (r_g W - r_b B) x q^-1 (0,0,1) ],
// DYNAMIC simulator: classic RK4 on the velocities over one sampling period, pose left untouched.
#pragma once
#include <cmath>

#include "manifold.hpp"

namespace orc {

struct UwvParams {
  double inertia[36];    // row-major 6x6
  double damp_lin[36];   // damping_matrices[0]
  double damp_quad[36];  // damping_matrices[1]
  double cog[3];         // distance_body2centerofgravity
  double cob[3];         // distance_body2centerofbuoyancy
  double weight, buoyancy;
};

inline void cross3(const double a[3], const double b[3], double o[3]) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
inline void matvec6(const double *M, const double v[6], double o[6]) {
  for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += M[i * 6 + j] * v[j]; o[i] = s; }
}

// the velocity/orientation dependent part of calcEfforts: C + D nu + g
inline void hydro_efforts(const UwvParams &P, const double vel[6], const double q[4], double out[6]) {
  double mv[6];
  matvec6(P.inertia, vel, mv);
  double c0[3], c1[3], c2[3];
  cross3(mv, vel + 3, c0);
  cross3(mv, vel, c1);
  cross3(mv + 3, vel + 3, c2);
  double dl[6], dq[6], av[6];
  for (int i = 0; i < 6; ++i) av[i] = std::fabs(vel[i]) * vel[i];
  matvec6(P.damp_lin, vel, dl);
  matvec6(P.damp_quad, av, dq);
  double up_f[3] = {0, 0, P.weight - P.buoyancy}, up[3] = {0, 0, 1}, gf[3], gu[3], lever[3], gt[3];
  quat_rot_inv(q, up_f, gf);
  quat_rot_inv(q, up, gu);
  for (int i = 0; i < 3; ++i) lever[i] = P.cog[i] * P.weight - P.cob[i] * P.buoyancy;
  cross3(lever, gu, gt);
  for (int i = 0; i < 3; ++i) {
    out[i] = c0[i] + dl[i] + dq[i] + gf[i];
    out[3 + i] = c1[i] + c2[i] + dl[3 + i] + dq[3 + i] + gt[i];
  }
}

// DynamicModel::calcEfforts(acceleration, velocity, orientation)
inline void calc_efforts(const UwvParams &P, const double acc[6], const double vel[6], const double q[4], double out[6]) {
  double ma[6], h[6];
  matvec6(P.inertia, acc, ma);
  hydro_efforts(P, vel, q, h);
  for (int i = 0; i < 6; ++i) out[i] = ma[i] + h[i];
}

// solve M x = b for the 6x6 inertia matrix (partial-pivot Gaussian elimination)
inline void solve6(const double *M, const double b[6], double x[6]) {
  double a[6][7];
  for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) a[i][j] = M[i * 6 + j]; a[i][6] = b[i]; }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
    if (p != c) for (int k = 0; k < 7; ++k) { double t = a[p][k]; a[p][k] = a[c][k]; a[c][k] = t; }
    for (int r = c + 1; r < 6; ++r) {
      double f = a[r][c] / a[c][c];
      for (int k = c; k < 7; ++k) a[r][k] -= f * a[c][k];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double s = a[r][6];
    for (int k = r + 1; k < 6; ++k) s -= a[r][k] * x[k];
    x[r] = s / a[r][r];
  }
}

// DynamicModel::calcAcceleration(efforts, velocity, orientation) = M^-1 (tau - C - D nu - g)
inline void calc_acceleration(const UwvParams &P, const double tau[6], const double vel[6], const double q[4], double acc[6]) {
  double h[6], rhs[6];
  hydro_efforts(P, vel, q, h);
  for (int i = 0; i < 6; ++i) rhs[i] = tau[i] - h[i];
  solve6(P.inertia, rhs, acc);
}

// ModelSimulation(DYNAMIC, dt, 1)::sendEffort(tau, state): one RK4 step of the velocities.
inline void sim_step(const UwvParams &P, const double tau[6], const double q[4], const double vel[6], double dt, double vel_out[6]) {
  double k1[6], k2[6], k3[6], k4[6], t[6];
  calc_acceleration(P, tau, vel, q, k1);
  for (int i = 0; i < 6; ++i) t[i] = vel[i] + 0.5 * dt * k1[i];
  calc_acceleration(P, tau, t, q, k2);
  for (int i = 0; i < 6; ++i) t[i] = vel[i] + 0.5 * dt * k2[i];
  calc_acceleration(P, tau, t, q, k3);
  for (int i = 0; i < 6; ++i) t[i] = vel[i] + dt * k3[i];
  calc_acceleration(P, tau, t, q, k4);
  for (int i = 0; i < 6; ++i) vel_out[i] = vel[i] + dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

}  // namespace orc
