// ORACLE (test infrastructure, NOT product code) -- parity unpinned.
//
// CPU restatement of ukfom::ukf<state> (slam/mtk, un-vendored, no pinned version), the unscented
// Kalman filter the reference drives at src/PoseUKF.cpp:462-465 (predict) and :479-481, :487, :501,
// :509, :524, :562, :576, :608 (update), src/VelocityUKF.cpp:82,109,122, src/BottomUKF.cpp:52,59,65,
// src/IndirectPoseUKF.cpp:90,126.  Recipe as frozen in SURVEY.md Appendix B:
//   * sigma points  X0 = mu (+) delta, X(2j+1) = mu (+) (delta + L.col(j)), X(2j+2) = mu (+) (delta - L.col(j)),
//     L = lower Cholesky factor of Sigma (LAPACK dpotrf 'L'), NOT scaled;
//   * mean = equal-weight iterated manifold mean started at X0;
//   * covariances = 1/2 * sum over ALL 2n+1 points of d d^T;
//   * update: S = cov(Z)+R, K = cov(X,Z) S^-1, gate on nu^T S^-1 nu, Sigma -= K S K^T, apply_delta(K nu);
//   * apply_delta re-draws the sigma points around mu (+) delta and recomputes mean and covariance.
// Everything is executed literally (all 2n+1 points, full n x n accumulations) -- this is the CPU
// "reference path" and the checker for the structured CUDA engine.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <functional>
#include <vector>

#include "manifold.hpp"

namespace orc {

struct Policy {
  double mean_tol = 1e-9;   // loop while |mean_delta| > mean_tol ...
  int mean_max_it = 100;    // ... and fewer than mean_max_it passes were made
};

enum Status : unsigned {
  ST_OK = 0,
  ST_NOT_SPD = 1u,         // Cholesky met a non-positive pivot (ukfom: stderr message + assert)
  ST_NAN_MEAS = 2u,        // checkMeasurment() would have thrown
  ST_GATE_REJECTED = 4u,   // innovation gate refused the update
  ST_MEAN_NOT_CONVERGED = 8u,
};

// lower Cholesky of the lower triangle of row-major A (n x n); L row-major with zero upper part.
inline bool cholesky_lower(int n, const double *A, double *L) {
  bool spd = true;
  for (int i = 0; i < n * n; ++i) L[i] = 0.0;
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    if (!(d > 0.0)) { spd = false; d = std::fabs(d) > 0 ? std::fabs(d) : 1.0; }
    double ljj = std::sqrt(d);
    L[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / ljj;
    }
  }
  return spd;
}

// Eigen's fixed-size inverse: closed form (cofactors) up to 4x4, partial-pivot LU above.
inline bool small_inverse(int m, const double *S, double *Si) {
  if (m == 1) { Si[0] = 1.0 / S[0]; return S[0] != 0; }
  if (m == 2) {
    double det = S[0] * S[3] - S[1] * S[2];
    double id = 1.0 / det;
    Si[0] = S[3] * id; Si[1] = -S[1] * id; Si[2] = -S[2] * id; Si[3] = S[0] * id;
    return det != 0;
  }
  if (m == 3) {
    double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[5] * S[6] - S[3] * S[8], c02 = S[3] * S[7] - S[4] * S[6];
    double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
    double id = 1.0 / det;
    Si[0] = c00 * id; Si[1] = (S[2] * S[7] - S[1] * S[8]) * id; Si[2] = (S[1] * S[5] - S[2] * S[4]) * id;
    Si[3] = c01 * id; Si[4] = (S[0] * S[8] - S[2] * S[6]) * id; Si[5] = (S[2] * S[3] - S[0] * S[5]) * id;
    Si[6] = c02 * id; Si[7] = (S[1] * S[6] - S[0] * S[7]) * id; Si[8] = (S[0] * S[4] - S[1] * S[3]) * id;
    return det != 0;
  }
  // general: Gauss-Jordan with partial pivoting
  std::vector<double> a(S, S + m * m);
  for (int i = 0; i < m * m; ++i) Si[i] = 0;
  for (int i = 0; i < m; ++i) Si[i * m + i] = 1;
  for (int c = 0; c < m; ++c) {
    int p = c;
    for (int r = c + 1; r < m; ++r) if (std::fabs(a[r * m + c]) > std::fabs(a[p * m + c])) p = r;
    if (a[p * m + c] == 0) return false;
    if (p != c) for (int k = 0; k < m; ++k) { std::swap(a[p * m + k], a[c * m + k]); std::swap(Si[p * m + k], Si[c * m + k]); }
    double ip = 1.0 / a[c * m + c];
    for (int k = 0; k < m; ++k) { a[c * m + k] *= ip; Si[c * m + k] *= ip; }
    for (int r = 0; r < m; ++r) if (r != c) {
      double f = a[r * m + c];
      if (f != 0) for (int k = 0; k < m; ++k) { a[r * m + k] -= f * a[c * m + k]; Si[r * m + k] -= f * Si[c * m + k]; }
    }
  }
  return true;
}

class Ukf {
 public:
  using Vec = std::vector<double>;
  using Model = std::function<void(const double *x, double *out)>;
  using Gate = std::function<bool(double)>;

  // Workspaces are allocated once per object; reset() re-arms it for another filter (the batched CPU
  // baseline keeps one object per thread so that the timed loop is free of heap traffic -- the reference's
  // Eigen fixed-size types live on the stack, only its std::vector<state> of sigma points is heap memory).
  Ukf(const Layout &lay, const double *mu, const double *sigma, Policy pol = Policy())
      : lay_(lay), n_(lay.dof()), dim_(lay.dim()), mu_(lay.dim()), sigma_(lay.dof() * lay.dof()), pol_(pol),
        X_((2 * lay.dof() + 1) * lay.dim()), Y_(lay.dim()), L_(lay.dof() * lay.dof()), d_(lay.dof()), t_(lay.dof()),
        md_(lay.dof()), ref_(lay.dim()), backup_(lay.dof() * lay.dof()), delta_(lay.dof()), Z_((2 * lay.dof() + 1) * 8) {
    reset(mu, sigma);
  }
  void reset(const double *mu, const double *sigma) {
    std::copy(mu, mu + dim_, mu_.begin());
    std::copy(sigma, sigma + n_ * n_, sigma_.begin());
    status_ = 0;
  }

  const Vec &mu() const { return mu_; }
  const Vec &sigma() const { return sigma_; }
  int n() const { return n_; }
  unsigned status() const { return status_; }
  int last_mean_iterations() const { return last_mean_it_; }
  double last_mahalanobis2() const { return last_d2_; }

  // ukfom::ukf::predict(g, Q).  Returns false (state untouched) when Sigma is not SPD.
  bool predict(const Model &g, const double *Q) {
    if (!sigma_points(nullptr)) return false;
    const int np = 2 * n_ + 1;
    for (int i = 0; i < np; ++i) {
      g(&X_[i * dim_], Y_.data());
      std::copy(Y_.begin(), Y_.end(), &X_[i * dim_]);
    }
    mean(lay_, X_.data(), np, mu_.data());
    cov(lay_, mu_.data(), X_.data(), np, sigma_.data());
    for (int i = 0; i < n_ * n_; ++i) sigma_[i] += Q[i];
    return true;
  }

  // ukfom::ukf::update(z, h, R, mtest); zl is the manifold of the measurement (dim <= 8).
  // Returns true when the update was applied (gate accepted and Sigma SPD).
  bool update(const Layout &zl, const double *z, const Model &h, const double *R, const Gate &gate) {
    const int m = zl.dof(), zd = zl.dim(), np = 2 * n_ + 1;
    if (!sigma_points(nullptr)) return false;
    for (int i = 0; i < np; ++i) h(&X_[i * dim_], &Z_[i * zd]);
    double zbar[8], S[64], Si[64], nu[8], dz[8];
    mean(zl, Z_.data(), np, zbar);
    cov(zl, zbar, Z_.data(), np, S);
    for (int i = 0; i < m * m; ++i) S[i] += R[i];
    // cross_cov_sigma_points(mu_, meanZ, X, Z): deviations of X from the PRIOR mean
    Vec Pxz(n_ * m, 0.0), K(n_ * m, 0.0), KS(n_ * m, 0.0);
    for (int i = 0; i < np; ++i) {
      lay_.boxminus(&X_[i * dim_], mu_.data(), d_.data());
      zl.boxminus(&Z_[i * zd], zbar, dz);
      for (int a = 0; a < n_; ++a) for (int b = 0; b < m; ++b) Pxz[a * m + b] += d_[a] * dz[b];
    }
    for (auto &v : Pxz) v *= 0.5;
    small_inverse(m, S, Si);
    for (int a = 0; a < n_; ++a) for (int b = 0; b < m; ++b) {
      double s = 0; for (int c = 0; c < m; ++c) s += Pxz[a * m + c] * Si[c * m + b];
      K[a * m + b] = s;
    }
    zl.boxminus(z, zbar, nu);
    double d2 = 0;
    for (int a = 0; a < m; ++a) for (int b = 0; b < m; ++b) d2 += nu[a] * Si[a * m + b] * nu[b];
    last_d2_ = d2;
    if (!gate(d2)) { status_ |= ST_GATE_REJECTED; return false; }
    // sigma_ -= K * S * K^T  (evaluated as (K*S)*K^T over the full matrix, like Eigen)
    for (int a = 0; a < n_; ++a) for (int b = 0; b < m; ++b) {
      double s = 0; for (int c = 0; c < m; ++c) s += K[a * m + c] * S[c * m + b];
      KS[a * m + b] = s;
    }
    backup_ = sigma_;
    for (int a = 0; a < n_; ++a) for (int b = 0; b < n_; ++b) {
      double s = 0; for (int c = 0; c < m; ++c) s += KS[a * m + c] * K[b * m + c];
      sigma_[a * n_ + b] -= s;
    }
    for (int a = 0; a < n_; ++a) { double s = 0; for (int c = 0; c < m; ++c) s += K[a * m + c] * nu[c]; delta_[a] = s; }
    if (!apply_delta(delta_.data())) { sigma_ = backup_; return false; }
    return true;
  }

  // ukfom::ukf::apply_delta(delta)
  bool apply_delta(const double *delta) {
    if (!sigma_points(delta)) return false;
    const int np = 2 * n_ + 1;
    mean(lay_, X_.data(), np, mu_.data());
    cov(lay_, mu_.data(), X_.data(), np, sigma_.data());
    return true;
  }

 private:
  // generate_sigma_points(mu_, delta, sigma_, X)
  bool sigma_points(const double *delta) {
    if (!cholesky_lower(n_, sigma_.data(), L_.data())) { status_ |= ST_NOT_SPD; return false; }
    for (int a = 0; a < n_; ++a) d_[a] = delta ? delta[a] : 0.0;
    lay_.boxplus(mu_.data(), d_.data(), 1.0, &X_[0]);
    for (int j = 0, i = 1; j < n_; ++j) {
      for (int a = 0; a < n_; ++a) t_[a] = d_[a] + L_[a * n_ + j];
      lay_.boxplus(mu_.data(), t_.data(), 1.0, &X_[(i++) * dim_]);
      for (int a = 0; a < n_; ++a) t_[a] = d_[a] - L_[a * n_ + j];
      lay_.boxplus(mu_.data(), t_.data(), 1.0, &X_[(i++) * dim_]);
    }
    return true;
  }

  // mean_sigma_points: X holds np points of l.dim() doubles; out may alias nothing in X
  void mean(const Layout &l, const double *X, int np, double *out) {
    const int nd = l.dof(), dm = l.dim();
    double *ref = ref_.data(), *md = md_.data(), *d = t_.data();
    std::copy(X, X + dm, ref);
    int it = 0;
    double nrm;
    do {
      for (int a = 0; a < nd; ++a) md[a] = 0;
      for (int i = 0; i < np; ++i) {
        l.boxminus(X + i * dm, ref, d);
        for (int a = 0; a < nd; ++a) md[a] += d[a];
      }
      for (int a = 0; a < nd; ++a) md[a] /= (double)np;
      l.boxplus(ref, md, 1.0, ref);
      nrm = 0; for (int a = 0; a < nd; ++a) nrm += md[a] * md[a]; nrm = std::sqrt(nrm);
    } while (nrm > pol_.mean_tol && ++it < pol_.mean_max_it);
    last_mean_it_ = it + 1;
    if (it >= pol_.mean_max_it) status_ |= ST_MEAN_NOT_CONVERGED;
    std::copy(ref, ref + dm, out);
  }

  // cov_sigma_points: c = 1/2 sum d d^T over all points (full matrix)
  void cov(const Layout &l, const double *mean, const double *X, int np, double *c) {
    const int nd = l.dof(), dm = l.dim();
    double *d = t_.data();
    for (int i = 0; i < nd * nd; ++i) c[i] = 0.0;
    for (int i = 0; i < np; ++i) {
      l.boxminus(X + i * dm, mean, d);
      for (int a = 0; a < nd; ++a) { double da = d[a]; double *row = &c[a * nd]; for (int b = 0; b < nd; ++b) row[b] += da * d[b]; }
    }
    for (int i = 0; i < nd * nd; ++i) c[i] *= 0.5;
  }

  Layout lay_;
  int n_, dim_;
  Vec mu_, sigma_;
  Policy pol_;
  Vec X_, Y_, L_, d_, t_, md_, ref_, backup_, delta_, Z_;
  unsigned status_ = 0;
  int last_mean_it_ = 0;
  double last_d2_ = 0;
};

}  // namespace orc
