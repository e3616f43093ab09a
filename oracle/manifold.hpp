// ORACLE (test infrastructure, NOT product code) -- parity unpinned.
//
// CPU restatement of the manifold toolkit the reference executes through its un-vendored
// dependency slam/mtk (MTK + ukfom; no version pinned: /root/reference/manifest.xml:11).
// The reference uses these charts at src/PoseState.hpp:15-27 (typedefs), src/PoseUKF.cpp:25-78
// (boxplus(vec, scale)), src/IndirectPoseUKF.cpp:15 (SO3::log), src/BottomUKF.hpp:15-16 (Scalar, S2).
// The formulas are the published MTK ones (Hertzberg et al. 2013, "Integrating generic sensor fusion
// algorithms with sound state representations through encapsulation of manifolds") as summarised in
// SURVEY.md Appendix C.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
// this directory.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

// MTK::tolerance<double>() (mtkmath.h) -- guards the small-angle branches of log().
constexpr double kMtkTol = 1e-11;

// MTK cos_sinc_sqrt(x2): returns cos(sqrt(x2)) and sin(sqrt(x2))/sqrt(x2); a 3-term alternating
// Taylor series below eps^(1/4), libm above.
inline void cos_sinc_sqrt(double x2, double &c, double &s) {
  const double taylor_0 = 2.220446049250313e-16;          // machine epsilon
  const double taylor_n = std::sqrt(std::sqrt(taylor_0)); // eps^(1/4)
  if (x2 >= taylor_n) {
    double x = std::sqrt(x2);
    c = std::cos(x);
    s = std::sin(x) / x;
    return;
  }
  static const double inv[] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
  double cosi = 1., sinc = 1.;
  double term = -1 / 2. * x2;
  for (int i = 0; i < 3; ++i) {
    cosi += term;
    term *= inv[2 * i];
    sinc += term;
    term *= -inv[2 * i + 1] * x2;
  }
  c = cosi;
  s = sinc;
}

// Quaternions are stored (x, y, z, w) like Eigen::Quaterniond::coeffs().
inline void quat_mul(const double a[4], const double b[4], double o[4]) {
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
inline void quat_conj(const double a[4], double o[4]) {
  o[0] = -a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = a[3];
}
// q * v  (Eigen QuaternionBase::_transformVector: v + w*t + u x t with t = 2 (u x v)).
inline void quat_rot(const double q[4], const double v[3], double o[3]) {
  double tx = 2 * (q[1] * v[2] - q[2] * v[1]);
  double ty = 2 * (q[2] * v[0] - q[0] * v[2]);
  double tz = 2 * (q[0] * v[1] - q[1] * v[0]);
  double ox = v[0] + q[3] * tx + (q[1] * tz - q[2] * ty);
  double oy = v[1] + q[3] * ty + (q[2] * tx - q[0] * tz);
  double oz = v[2] + q[3] * tz + (q[0] * ty - q[1] * tx);
  o[0] = ox; o[1] = oy; o[2] = oz;
}
// q^-1 * v for a unit quaternion (reference: state.orientation.inverse() * v, src/PoseUKF.cpp:122).
inline void quat_rot_inv(const double q[4], const double v[3], double o[3]) {
  double c[4];
  quat_conj(q, c);
  quat_rot(c, v, o);
}
// q.matrix() row-major 3x3 (used at src/PoseUKF.cpp:448, src/IndirectPoseUKF.cpp:82).
inline void quat_to_mat(const double q[4], double R[9]) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// MTK SO3::exp(v, scale): (w, vec) = (cos(|v| scale/2), sinc(|v| scale/2) * scale/2 * v).
inline void so3_exp(const double v[3], double scale, double q[4]) {
  double h = scale / 2;
  double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  double c, s;
  cos_sinc_sqrt(h * h * n2, c, s);
  double mult = s * h;
  q[0] = mult * v[0]; q[1] = mult * v[1]; q[2] = mult * v[2]; q[3] = c;
}
// MTK SO3::log(q) = MTK::log(w, vec, scale = 2, plus_minus_periodicity = true).
inline void so3_log(const double q[4], double out[3]) {
  double nv = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  if (nv < kMtkTol) nv = kMtkTol;
  double s = 2.0 / nv * std::atan(nv / q[3]);
  out[0] = s * q[0]; out[1] = s * q[1]; out[2] = s * q[2];
}
// q (+) (v, scale) = q * exp(v, scale)   (right multiplication, body-frame perturbation)
inline void so3_boxplus(const double q[4], const double v[3], double scale, double o[4]) {
  double d[4];
  so3_exp(v, scale, d);
  quat_mul(q, d, o);
}
// a (-) b = log(b^-1 * a)
inline void so3_boxminus(const double a[4], const double b[4], double o[3]) {
  double c[4], r[4];
  quat_conj(b, c);
  quat_mul(c, a, r);
  so3_log(r, o);
}

// ---- S2 (unit sphere) chart, MTK's Householder construction ---------------------------------------
// R_x = reflection that maps e_x to x (identity when x == e_x).  x (+) d = R_x [cos|d|, sinc|d| d];
// y (-) x = theta/|v| v  with  [w, v] = R_x y,  theta = atan2(|v|, w).
inline void s2_reflect(const double x[3], const double in[3], double out[3]) {
  // Householder vector u = x - e_x ; R = I - 2 u u^T / (u^T u)
  double u0 = x[0] - 1.0, u1 = x[1], u2 = x[2];
  double uu = u0 * u0 + u1 * u1 + u2 * u2;
  if (uu < kMtkTol * kMtkTol) { out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; return; }
  double f = 2.0 * (u0 * in[0] + u1 * in[1] + u2 * in[2]) / uu;
  out[0] = in[0] - f * u0; out[1] = in[1] - f * u1; out[2] = in[2] - f * u2;
}
inline void s2_boxplus(const double x[3], const double d[2], double scale, double o[3]) {
  double n2 = d[0] * d[0] + d[1] * d[1];
  double c, s;
  cos_sinc_sqrt(scale * scale * n2, c, s);
  double loc[3] = {c, s * scale * d[0], s * scale * d[1]};
  s2_reflect(x, loc, o);
}
inline void s2_boxminus(const double y[3], const double x[3], double o[2]) {
  double loc[3];
  s2_reflect(x, y, loc);
  double nv = std::sqrt(loc[1] * loc[1] + loc[2] * loc[2]);
  if (nv < kMtkTol) { o[0] = 0; o[1] = 0; return; }
  double s = std::atan2(nv, loc[0]) / nv;
  o[0] = s * loc[1]; o[1] = s * loc[2];
}
inline void s2_normalize(double v[3]) {
  double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  v[0] /= n; v[1] /= n; v[2] /= n;
}

// ---- compound manifolds (MTK_BUILD_MANIFOLD): ordered list of sub-manifolds --------------------------
enum CompType { VECT = 0, SO3 = 1, S2 = 2 };
struct Comp { CompType type; int dof; int dim; };

struct Layout {
  std::vector<Comp> comps;
  int dof() const { int n = 0; for (auto &c : comps) n += c.dof; return n; }
  int dim() const { int n = 0; for (auto &c : comps) n += c.dim; return n; }
  // out = x (+) scale*d     (out may alias x)
  void boxplus(const double *x, const double *d, double scale, double *out) const {
    int io = 0, id = 0;
    for (auto &c : comps) {
      if (c.type == VECT) {
        for (int k = 0; k < c.dim; ++k) out[io + k] = x[io + k] + scale * d[id + k];
      } else if (c.type == SO3) {
        double o[4];
        so3_boxplus(x + io, d + id, scale, o);
        std::memcpy(out + io, o, sizeof o);
      } else {
        double o[3];
        s2_boxplus(x + io, d + id, scale, o);
        std::memcpy(out + io, o, sizeof o);
      }
      io += c.dim; id += c.dof;
    }
  }
  // out = a (-) b
  void boxminus(const double *a, const double *b, double *out) const {
    int io = 0, id = 0;
    for (auto &c : comps) {
      if (c.type == VECT) {
        for (int k = 0; k < c.dim; ++k) out[id + k] = a[io + k] - b[io + k];
      } else if (c.type == SO3) {
        so3_boxminus(a + io, b + io, out + id);
      } else {
        s2_boxminus(a + io, b + io, out + id);
      }
      io += c.dim; id += c.dof;
    }
  }
};

inline Layout vect_layout(int n) { return Layout{{{VECT, n, n}}}; }

}  // namespace orc
