// Small-state filters (VelocityUKF n=4, BottomUKF n=3, IndirectPoseUKF n=6 / augmented n=12): one THREAD per
// filter, the ukfom recipe executed literally over all 2n+1 sigma points -- at these sizes
// the whole filter lives in one thread's registers / local memory and there is nothing to gain from structure.
// Reference: src/VelocityUKF.cpp:6-130, src/BottomUKF.cpp:5-67, src/IndirectPoseUKF.cpp:7-142.
#pragma once
#include "dmath.cuh"

namespace uwvkf {
namespace small {

enum { ST_NOT_SPD = 1u, ST_NAN_MEAS = 2u, ST_GATE_REJECTED = 4u, ST_MEAN_NOT_CONVERGED = 8u };

struct Policy {
  double tol;
  int max_it;
};

// ---- S2 chart (MTK, Householder construction; same definition as oracle/manifold.hpp) -------------------
__device__ __forceinline__ void s2_reflect(const double *x, const double *in, double *out) {
  double u0 = x[0] - 1.0, u1 = x[1], u2 = x[2];
  double uu = u0 * u0 + u1 * u1 + u2 * u2;
  if (uu < kMtkTol * kMtkTol) { out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; return; }
  double f = 2.0 * (u0 * in[0] + u1 * in[1] + u2 * in[2]) / uu;
  out[0] = in[0] - f * u0; out[1] = in[1] - f * u1; out[2] = in[2] - f * u2;
}
__device__ __forceinline__ void s2_boxplus(const double *x, const double *d, double *o) {
  double c, s;
  cos_sinc_sqrt(d[0] * d[0] + d[1] * d[1], c, s);
  double loc[3] = {c, s * d[0], s * d[1]};
  double t[3];
  s2_reflect(x, loc, t);
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
__device__ __forceinline__ void s2_boxminus(const double *y, const double *x, double *o) {
  double loc[3];
  s2_reflect(x, y, loc);
  double nv = sqrt(loc[1] * loc[1] + loc[2] * loc[2]);
  if (nv < kMtkTol) { o[0] = 0; o[1] = 0; return; }
  double s = atan2(nv, loc[0]) / nv;
  o[0] = s * loc[1]; o[1] = s * loc[2];
}
__device__ __forceinline__ void so3_bp(const double *q, const double *d, double scale, double *o) {
  Quat r = so3_boxplus(Quat{q[0], q[1], q[2], q[3]}, Vec3{d[0], d[1], d[2]}, scale);
  o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
__device__ __forceinline__ void so3_bm(const double *a, const double *b, double *o) {
  Vec3 r = so3_boxminus(Quat{a[0], a[1], a[2], a[3]}, Quat{b[0], b[1], b[2], b[3]});
  o[0] = r.x; o[1] = r.y; o[2] = r.z;
}

// ---- manifolds: static DOF / DIM / boxplus(x, d, out) / boxminus(a, b, out = a (-) b) ------------------
template <int NV>
struct VectM {
  static constexpr int DOF = NV, DIM = NV;
  __device__ static void boxplus(const double *x, const double *d, double *o) { for (int k = 0; k < NV; ++k) o[k] = x[k] + d[k]; }
  __device__ static void boxminus(const double *a, const double *b, double *o) { for (int k = 0; k < NV; ++k) o[k] = a[k] - b[k]; }
};
struct S2M {
  static constexpr int DOF = 2, DIM = 3;
  __device__ static void boxplus(const double *x, const double *d, double *o) { s2_boxplus(x, d, o); }
  __device__ static void boxminus(const double *a, const double *b, double *o) { s2_boxminus(a, b, o); }
};
struct BottomM {  // distance (Scalar) + normal (S2), src/BottomUKF.hpp:18-21
  static constexpr int DOF = 3, DIM = 4;
  __device__ static void boxplus(const double *x, const double *d, double *o) { o[0] = x[0] + d[0]; s2_boxplus(x + 1, d + 1, o + 1); }
  __device__ static void boxminus(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; s2_boxminus(a + 1, b + 1, o + 1); }
};
struct PoseErrM {  // position_error (vect3) + orientation_error (SO3), src/IndirectPoseUKF.hpp:19-22
  static constexpr int DOF = 6, DIM = 7;
  __device__ static void boxplus(const double *x, const double *d, double *o) {
    for (int k = 0; k < 3; ++k) o[k] = x[k] + d[k];
    so3_bp(x + 3, d + 3, 1.0, o + 3);
  }
  __device__ static void boxminus(const double *a, const double *b, double *o) {
    for (int k = 0; k < 3; ++k) o[k] = a[k] - b[k];
    so3_bm(a + 3, b + 3, o + 3);
  }
};
struct PoseErrMarkerM {  // FilterStateWithMarker, src/IndirectPoseUKF.cpp:26-30
  static constexpr int DOF = 12, DIM = 14;
  __device__ static void boxplus(const double *x, const double *d, double *o) {
    PoseErrM::boxplus(x, d, o);
    PoseErrM::boxplus(x + 7, d + 6, o + 7);
  }
  __device__ static void boxminus(const double *a, const double *b, double *o) {
    PoseErrM::boxminus(a, b, o);
    PoseErrM::boxminus(a + 7, b + 7, o + 6);
  }
};

// lower Cholesky of the lower triangle of row-major S (n x n); L full row-major with zero upper part
template <int N>
__device__ bool chol_lower(const double *S, double *L) {
  bool ok = true;
  for (int i = 0; i < N * N; ++i) L[i] = 0.0;
  for (int j = 0; j < N; ++j) {
    double d = S[j * N + j];
    for (int k = 0; k < j; ++k) d -= L[j * N + k] * L[j * N + k];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    double ljj = sqrt(d);
    L[j * N + j] = ljj;
    for (int i = j + 1; i < N; ++i) {
      double s = S[i * N + j];
      for (int k = 0; k < j; ++k) s -= L[i * N + k] * L[j * N + k];
      L[i * N + j] = s / ljj;
    }
  }
  return ok;
}

template <int M>
__device__ void small_inverse(const double *S, double *Si) {
  if constexpr (M == 1) {
    Si[0] = 1.0 / S[0];
  } else if constexpr (M == 2) {
    double id = 1.0 / (S[0] * S[3] - S[1] * S[2]);
    Si[0] = S[3] * id; Si[1] = -S[1] * id; Si[2] = -S[2] * id; Si[3] = S[0] * id;
  } else {
    static_assert(M == 3, "measurement dimension");
    double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[5] * S[6] - S[3] * S[8], c02 = S[3] * S[7] - S[4] * S[6];
    double id = 1.0 / (S[0] * c00 + S[1] * c01 + S[2] * c02);
    Si[0] = c00 * id; Si[1] = (S[2] * S[7] - S[1] * S[8]) * id; Si[2] = (S[1] * S[5] - S[2] * S[4]) * id;
    Si[3] = c01 * id; Si[4] = (S[0] * S[8] - S[2] * S[6]) * id; Si[5] = (S[2] * S[3] - S[0] * S[5]) * id;
    Si[6] = c02 * id; Si[7] = (S[1] * S[6] - S[0] * S[7]) * id; Si[8] = (S[0] * S[4] - S[1] * S[3]) * id;
  }
}

// ukfom::ukf<state> for one filter held by one thread
template <class SM>
struct Ukf {
  static constexpr int N = SM::DOF, DIM = SM::DIM, NP = 2 * SM::DOF + 1;
  double mu[DIM];
  double S[N * N];
  double X[NP * DIM];
  Policy pol;
  unsigned status;

  __device__ bool sigma_points(const double *delta) {
    double L[N * N], d[N], t[N];
    if (!chol_lower<N>(S, L)) { status |= ST_NOT_SPD; return false; }
    for (int a = 0; a < N; ++a) d[a] = delta ? delta[a] : 0.0;
    SM::boxplus(mu, d, X);
    for (int j = 0, i = 1; j < N; ++j) {
      for (int a = 0; a < N; ++a) t[a] = d[a] + L[a * N + j];
      SM::boxplus(mu, t, X + (i++) * DIM);
      for (int a = 0; a < N; ++a) t[a] = d[a] - L[a * N + j];
      SM::boxplus(mu, t, X + (i++) * DIM);
    }
    return true;
  }
  template <class MM>
  __device__ void mean(const double *P, double *out) {
    double ref[MM::DIM], md[MM::DOF], d[MM::DOF];
    for (int k = 0; k < MM::DIM; ++k) ref[k] = P[k];
    int it = 0;
    double nrm;
    do {
      for (int a = 0; a < MM::DOF; ++a) md[a] = 0;
      for (int i = 0; i < NP; ++i) {
        MM::boxminus(P + i * MM::DIM, ref, d);
        for (int a = 0; a < MM::DOF; ++a) md[a] += d[a];
      }
      for (int a = 0; a < MM::DOF; ++a) md[a] /= double(NP);
      double nr[MM::DIM];
      MM::boxplus(ref, md, nr);
      for (int k = 0; k < MM::DIM; ++k) ref[k] = nr[k];
      nrm = 0;
      for (int a = 0; a < MM::DOF; ++a) nrm += md[a] * md[a];
      nrm = sqrt(nrm);
    } while (nrm > pol.tol && ++it < pol.max_it);
    if (it >= pol.max_it) status |= ST_MEAN_NOT_CONVERGED;
    for (int k = 0; k < MM::DIM; ++k) out[k] = ref[k];
  }
  template <class MM>
  __device__ void cov(const double *mean_, const double *P, double *c) {
    double d[MM::DOF];
    for (int i = 0; i < MM::DOF * MM::DOF; ++i) c[i] = 0;
    for (int i = 0; i < NP; ++i) {
      MM::boxminus(P + i * MM::DIM, mean_, d);
      for (int a = 0; a < MM::DOF; ++a)
        for (int b = 0; b < MM::DOF; ++b) c[a * MM::DOF + b] += d[a] * d[b];
    }
    for (int i = 0; i < MM::DOF * MM::DOF; ++i) c[i] *= 0.5;
  }
  // predict(g, Q): g(x_in, x_out)
  template <class G>
  __device__ bool predict(G g, const double *Q) {
    if (!sigma_points(nullptr)) return false;
    for (int i = 0; i < NP; ++i) {
      double y[DIM];
      g(X + i * DIM, y);
      for (int k = 0; k < DIM; ++k) X[i * DIM + k] = y[k];
    }
    double nm[DIM];
    mean<SM>(X, nm);
    for (int k = 0; k < DIM; ++k) mu[k] = nm[k];
    cov<SM>(mu, X, S);
    for (int i = 0; i < N * N; ++i) S[i] += Q[i];
    return true;
  }
  __device__ bool apply_delta(const double *delta) {
    if (!sigma_points(delta)) return false;
    double nm[DIM];
    mean<SM>(X, nm);
    for (int k = 0; k < DIM; ++k) mu[k] = nm[k];
    cov<SM>(mu, X, S);
    return true;
  }
  // update(z, h, R, accept_any): ZM = measurement manifold, h(x, z_out)
  template <class ZM, class H>
  __device__ bool update(const double *z, H h, const double *R) {
    constexpr int M = ZM::DOF, ZD = ZM::DIM;
    if (!sigma_points(nullptr)) return false;
    double Z[NP * ZD];
    for (int i = 0; i < NP; ++i) h(X + i * DIM, Z + i * ZD);
    double zbar[ZD], Szz[M * M], Si[M * M], Pxz[N * M], K[N * M], KS[N * M], nu[M], dx[N], dz[M], delta[N];
    mean<ZM>(Z, zbar);
    cov<ZM>(zbar, Z, Szz);
    for (int i = 0; i < M * M; ++i) Szz[i] += R[i];
    for (int i = 0; i < N * M; ++i) Pxz[i] = 0;
    for (int i = 0; i < NP; ++i) {
      SM::boxminus(X + i * DIM, mu, dx);
      ZM::boxminus(Z + i * ZD, zbar, dz);
      for (int a = 0; a < N; ++a)
        for (int b = 0; b < M; ++b) Pxz[a * M + b] += dx[a] * dz[b];
    }
    for (int i = 0; i < N * M; ++i) Pxz[i] *= 0.5;
    small_inverse<M>(Szz, Si);
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < M; ++b) {
        double s = 0;
        for (int c = 0; c < M; ++c) s += Pxz[a * M + c] * Si[c * M + b];
        K[a * M + b] = s;
      }
    ZM::boxminus(z, zbar, nu);
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < M; ++b) {
        double s = 0;
        for (int c = 0; c < M; ++c) s += K[a * M + c] * Szz[c * M + b];
        KS[a * M + b] = s;
      }
    double backup[N * N];
    for (int i = 0; i < N * N; ++i) backup[i] = S[i];
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        double s = 0;
        for (int c = 0; c < M; ++c) s += KS[a * M + c] * K[b * M + c];
        S[a * N + b] -= s;
      }
    for (int a = 0; a < N; ++a) {
      double s = 0;
      for (int c = 0; c < M; ++c) s += K[a * M + c] * nu[c];
      delta[a] = s;
    }
    if (!apply_delta(delta)) {
      for (int i = 0; i < N * N; ++i) S[i] = backup[i];
      return false;
    }
    return true;
  }
};

// ---- Fossen 6-DOF model (uwv_dynamic_model restatement) ------------------------------
struct UwvParams {
  double inertia[36], damp_lin[36], damp_quad[36];
  double cog[3], cob[3], weight, buoyancy;
  double inertia_inv[36];  // filled on the host
};
__device__ __forceinline__ void matvec6(const double *M, const double *v, double *o) {
  for (int i = 0; i < 6; ++i) {
    double s = 0;
    for (int j = 0; j < 6; ++j) s += M[i * 6 + j] * v[j];
    o[i] = s;
  }
}
__device__ __forceinline__ void cross3(const double *a, const double *b, double *o) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ inline void calc_acceleration(const UwvParams &P, const double *tau, const double *vel, const Quat &q, double *acc) {
  double mv[6], c0[3], c1[3], c2[3], av[6], dl[6], dq[6], rhs[6];
  matvec6(P.inertia, vel, mv);
  cross3(mv, vel + 3, c0);
  cross3(mv, vel, c1);
  cross3(mv + 3, vel + 3, c2);
  for (int i = 0; i < 6; ++i) av[i] = fabs(vel[i]) * vel[i];
  matvec6(P.damp_lin, vel, dl);
  matvec6(P.damp_quad, av, dq);
  Vec3 gf = qrot_inv(q, Vec3{0, 0, P.weight - P.buoyancy});
  Vec3 gu = qrot_inv(q, Vec3{0, 0, 1});
  double lever[3] = {P.cog[0] * P.weight - P.cob[0] * P.buoyancy, P.cog[1] * P.weight - P.cob[1] * P.buoyancy,
                     P.cog[2] * P.weight - P.cob[2] * P.buoyancy};
  double guv[3] = {gu.x, gu.y, gu.z}, gt[3], gfv[3] = {gf.x, gf.y, gf.z};
  cross3(lever, guv, gt);
  for (int i = 0; i < 3; ++i) {
    rhs[i] = tau[i] - (c0[i] + dl[i] + dq[i] + gfv[i]);
    rhs[3 + i] = tau[3 + i] - (c1[i] + c2[i] + dl[3 + i] + dq[3 + i] + gt[i]);
  }
  matvec6(P.inertia_inv, rhs, acc);
}
// ModelSimulation(DYNAMIC)::sendEffort: one RK4 step of the 6 velocities, pose untouched
__device__ inline void sim_step(const UwvParams &P, const double *tau, const Quat &q, const double *vel, double dt, double *out) {
  double k1[6], k2[6], k3[6], k4[6], t[6];
  calc_acceleration(P, tau, vel, q, k1);
  for (int i = 0; i < 6; ++i) t[i] = vel[i] + 0.5 * dt * k1[i];
  calc_acceleration(P, tau, t, q, k2);
  for (int i = 0; i < 6; ++i) t[i] = vel[i] + 0.5 * dt * k2[i];
  calc_acceleration(P, tau, t, q, k3);
  for (int i = 0; i < 6; ++i) t[i] = vel[i] + dt * k3[i];
  calc_acceleration(P, tau, t, q, k4);
  for (int i = 0; i < 6; ++i) out[i] = vel[i] + dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

// rigid transforms (t[3], q[4]) as 7 doubles
__device__ __forceinline__ void xf_mul(const double *a, const double *b, double *c) {
  Quat qa{a[3], a[4], a[5], a[6]};
  Vec3 t = qrot(qa, Vec3{b[0], b[1], b[2]});
  Quat q = qmul(qa, Quat{b[3], b[4], b[5], b[6]});
  c[0] = t.x + a[0]; c[1] = t.y + a[1]; c[2] = t.z + a[2];
  c[3] = q.x; c[4] = q.y; c[5] = q.z; c[6] = q.w;
}

template <class SM>
__device__ __forceinline__ void load_filter(Ukf<SM> &u, const double *state, long long f) {
  constexpr int W = SM::DIM + SM::DOF * SM::DOF;
  const double *p = state + f * W;
  for (int k = 0; k < SM::DIM; ++k) u.mu[k] = p[k];
  for (int k = 0; k < SM::DOF * SM::DOF; ++k) u.S[k] = p[SM::DIM + k];
}
template <class SM>
__device__ __forceinline__ void store_filter(const Ukf<SM> &u, double *state, long long f) {
  constexpr int W = SM::DIM + SM::DOF * SM::DOF;
  double *p = state + f * W;
  for (int k = 0; k < SM::DIM; ++k) p[k] = u.mu[k];
  for (int k = 0; k < SM::DOF * SM::DOF; ++k) p[SM::DIM + k] = u.S[k];
}

}  // namespace small
}  // namespace uwvkf
