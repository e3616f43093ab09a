// Device math shared by the uwvkf kernels: warp primitives, MTK manifold charts (SO3), the local
// transverse-Mercator projection.  Formulas follow the reference's dependencies as frozen in
// SURVEY.md Appendix C/D (MTK SO3 chart: q (+) d = q * exp(d), a (-) b = log(b^-1 a), series branch of
// cos/sinc below eps^(1/4); pose_estimation::GeographicProjection = Krueger n-series TM on WGS-84).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uwvkf {

constexpr unsigned kFull = 0xffffffffu;
constexpr double kMtkTol = 1e-11;      // MTK::tolerance<double>()
constexpr double kEarthW = 7.292115e-5;  // pose_estimation::EARTHW

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(kFull, v, src); }

// ---- warp-wide sum on the FP64 tensor path ----------------------------------------------------------------
// mma.sync.m8n8k4.f64 (SASS DMMA) computes D[i][j] = sum_k A[i][k] B[k][j] with A[i][k] held by lane 4i+k, B[k][j] by lane
// 4j+k and D[i][2t], D[i][2t+1] by lane 4i+t.  With A = 1 and B = v the first DMMA hands every quad of lanes the eight
// quad sums of v (two per lane); one DADD and a second DMMA with B = 1 sum those over the quad.  3 instructions and no
// shuffle instead of 5 x (2 SHFL + DADD): the shuffles of the v2 kernel were 45 % of its shared-memory pipe traffic.
__device__ __forceinline__ double wsum(double v) {
  double d0, d1, e0, e1;
  const double one = 1.0, zero = 0.0;
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
      : "=d"(d0), "=d"(d1) : "d"(one), "d"(v), "d"(zero), "d"(zero));
  const double s = d0 + d1;
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
      : "=d"(e0), "=d"(e1) : "d"(s), "d"(one), "d"(zero), "d"(zero));
  (void)e1;
  return e0;
}


// 1/sqrt(d) for d > 0: hardware seed (rsqrt.approx.ftz.f64, ~2^-22) + two Newton steps -> full double accuracy (~1 ulp)
// without the special-case handling of rsqrt(); callers test d > 0 themselves.
__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double t = d * y;
  double e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  t = d * y;
  e = fma(-t, y, 1.0);
  y = fma(0.5 * y, e, y);
  return y;
}
// same result (seed error 2^-22 cubed per step), one third-order step: y (1 + e/2 + 3 e^2 / 8), e = 1 - d y^2
__device__ __forceinline__ double fast_rsqrt3(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}

// sin and cos of a SMALL argument (|x| <= 2^-4) by their Taylor polynomials (truncation < 3e-19 relative): what libm
// returns to an ulp, without its argument reduction.  Callers guard the range and fall back to sincos().
__device__ __forceinline__ void sincos_small(double x, double &sn, double &cs) {
  const double u = x * x;
  double ps = fma(u, 1.0 / 362880.0, -1.0 / 5040.0);
  ps = fma(ps, u, 1.0 / 120.0);
  ps = fma(ps, u, -1.0 / 6.0);
  sn = fma(x * u, ps, x);
  double pc = fma(u, -1.0 / 3628800.0, 1.0 / 40320.0);
  pc = fma(pc, u, -1.0 / 720.0);
  pc = fma(pc, u, 1.0 / 24.0);
  pc = fma(pc, u, -0.5);
  cs = fma(pc, u, 1.0);
}
// sinh and cosh of a small argument (|t| <= 2^-3), Taylor (truncation < 4e-20 relative); no cancellation near 0
__device__ __forceinline__ void sinhcosh_small(double t, double &sh, double &ch) {
  const double u = t * t;
  double ps = fma(u, 1.0 / 39916800.0, 1.0 / 362880.0);
  ps = fma(ps, u, 1.0 / 5040.0);
  ps = fma(ps, u, 1.0 / 120.0);
  ps = fma(ps, u, 1.0 / 6.0);
  sh = fma(t * u, ps, t);
  double pc = fma(u, 1.0 / 479001600.0, 1.0 / 3628800.0);
  pc = fma(pc, u, 1.0 / 40320.0);
  pc = fma(pc, u, 1.0 / 720.0);
  pc = fma(pc, u, 1.0 / 24.0);
  pc = fma(pc, u, 0.5);
  ch = fma(pc, u, 1.0);
}
// 1/x (normal range): seed + two Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

struct Quat {
  double x, y, z, w;
};
struct Vec3 {
  double x, y, z;
};

__device__ __forceinline__ Quat qmul(const Quat &a, const Quat &b) {
  Quat o;
  o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  o.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  o.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return o;
}
__device__ __forceinline__ Quat qconj(const Quat &a) { return Quat{-a.x, -a.y, -a.z, a.w}; }
// q * v  (v + w t + u x t, t = 2 u x v)
__device__ __forceinline__ Vec3 qrot(const Quat &q, const Vec3 &v) {
  double tx = 2 * (q.y * v.z - q.z * v.y);
  double ty = 2 * (q.z * v.x - q.x * v.z);
  double tz = 2 * (q.x * v.y - q.y * v.x);
  Vec3 o;
  o.x = v.x + q.w * tx + (q.y * tz - q.z * ty);
  o.y = v.y + q.w * ty + (q.z * tx - q.x * tz);
  o.z = v.z + q.w * tz + (q.x * ty - q.y * tx);
  return o;
}
__device__ __forceinline__ Vec3 qrot_inv(const Quat &q, const Vec3 &v) { return qrot(qconj(q), v); }
__device__ __forceinline__ void qmat(const Quat &q, double R[9]) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// MTK cos_sinc_sqrt: cos(sqrt(x2)), sin(sqrt(x2))/sqrt(x2)
__device__ __forceinline__ void cos_sinc_sqrt(double x2, double &c, double &s) {
  if (x2 >= 0.0625) {  // |x| >= 0.25: libm, as MTK does above eps^(1/4)
    double x = sqrt(x2), sn, cs;
    sincos(x, &sn, &cs);
    c = cs;
    s = sn / x;
  } else {
    // Below 1/16 ONE polynomial serves every lane.  MTK switches to its three-term series (cos: 1 - x2/2 + x2^2/24 - x2^3/720,
    // sinc: 1 - x2/6 + x2^2/120 - x2^3/5040) below eps^(1/4) = 2^-13 and calls cos / sin of sqrt(x2) above; the Taylor
    // polynomials here (truncation < 3e-20) return the libm doubles to 1 ulp above that bound, and below it their extra terms
    // (x2^4 / 40320 < 6e-21) vanish against the 1.1e-16 ulp of the result: the same doubles as the three-term series.  A
    // sigma-point set always mixes both ranges (the centre point has x2 = 0), so a second branch here would make every
    // boxplus of the kernel divergent.  Even / odd split (Estrin) in y = x2^2: half the dependent depth of a Horner chain.
    const double y = x2 * x2;
    double ce = fma(y, 1.0 / 479001600.0, 1.0 / 40320.0);       // x2^2, x2^4, x2^6 terms of cos: 1/24, 1/40320, 1/12!
    ce = fma(ce, y, 1.0 / 24.0);
    double co = fma(y, -1.0 / 87178291200.0, -1.0 / 3628800.0);  // x2^1, x2^3, x2^5, x2^7: -1/2, -1/720, -1/10!, -1/14!
    co = fma(co, y, -1.0 / 720.0);
    co = fma(co, y, -0.5);
    c = fma(ce, y, fma(co, x2, 1.0));
    double se = fma(y, 1.0 / 6227020800.0, 1.0 / 362880.0);      // sinc: 1/120, 1/9!, 1/13!
    se = fma(se, y, 1.0 / 120.0);
    double so = fma(y, -1.0 / 1307674368000.0, -1.0 / 39916800.0);   // -1/6, -1/5040, -1/11!, -1/15!
    so = fma(so, y, -1.0 / 5040.0);
    so = fma(so, y, -1.0 / 6.0);
    s = fma(se, y, fma(so, x2, 1.0));
  }
}
__device__ __forceinline__ Quat so3_exp(const Vec3 &v, double scale) {
  double h = 0.5 * scale;
  double n2 = v.x * v.x + v.y * v.y + v.z * v.z;
  double c, s;
  cos_sinc_sqrt(h * h * n2, c, s);
  double m = s * h;
  return Quat{m * v.x, m * v.y, m * v.z, c};
}
// MTK SO3::log: s = 2 atan(|u| / w) / |u|.  For |u|^2 <= 1e-2 w^2 (rotation below ~0.2 rad, every sigma-point
// deviation of a working filter) atan(x)/x is evaluated by its alternating series in t = x^2 -- truncation
// t^10/21 < 5e-22, i.e. the same double as sqrt + atan + two divisions to 1 ulp -- without sqrt, atan or |u|.
__device__ __forceinline__ Vec3 so3_log(const Quat &q) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
  double w2 = q.w * q.w;
  double s;
  if (n2 <= 1.0e-2 * w2) {
    double iw = fast_rcp(q.w);
    double t = n2 * iw * iw;
    // atan(x)/x = sum (-t)^k / (2k+1), even / odd split in u = t^2 (half the dependent depth of a Horner chain)
    const double u = t * t;
    double pe = fma(u, 1.0 / 17.0, 1.0 / 13.0);      // k = 0, 2, 4, 6, 8
    pe = fma(pe, u, 1.0 / 9.0);
    pe = fma(pe, u, 1.0 / 5.0);
    pe = fma(pe, u, 1.0);
    double po = fma(u, -1.0 / 19.0, -1.0 / 15.0);    // k = 1, 3, 5, 7, 9
    po = fma(po, u, -1.0 / 11.0);
    po = fma(po, u, -1.0 / 7.0);
    po = fma(po, u, -1.0 / 3.0);
    const double p = fma(po, t, pe);
    s = 2.0 * iw * p;
  } else {
    double nv = sqrt(n2);
    if (nv < kMtkTol) nv = kMtkTol;
    s = 2.0 / nv * atan(nv / q.w);
  }
  return Vec3{s * q.x, s * q.y, s * q.z};
}
__device__ __forceinline__ Quat so3_boxplus(const Quat &q, const Vec3 &v, double scale) { return qmul(q, so3_exp(v, scale)); }
__device__ __forceinline__ Vec3 so3_boxminus(const Quat &a, const Quat &b) { return so3_log(qmul(qconj(b), a)); }

// ---- local transverse Mercator (Krueger series, order n^6); constants are filled on the host -----
struct TMercConst {
  double lat0, lon0, A, Ainv, xi0, e;
  double sin_xi0, cos_xi0;
  double alpha[6], beta[6], delta[6];
};

// NWU (x north, y west) -> geodetic latitude (the only output the process model consumes,
// reference src/PoseUKF.cpp:29-30)
__device__ __forceinline__ double tmerc_nav_to_lat(const TMercConst &tm, double x, double y) {
  double xi = x * tm.Ainv + tm.xi0;
  double eta = -y * tm.Ainv;
  double s2, c2;
  sincos(2 * xi, &s2, &c2);
  double sh2 = sinh(2 * eta), ch2 = cosh(2 * eta);
  // sin/cos(2j xi), sinh/cosh(2j eta) by angle addition
  double sj = s2, cj = c2, shj = sh2, chj = ch2;
  double xip = xi, etap = eta;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    xip -= tm.beta[j] * sj * chj;
    etap -= tm.beta[j] * cj * shj;
    double sn = sj * c2 + cj * s2, cn = cj * c2 - sj * s2;
    double shn = shj * ch2 + chj * sh2, chn = chj * ch2 + shj * sh2;
    sj = sn; cj = cn; shj = shn; chj = chn;
  }
  double chi = asin(sin(xip) / cosh(etap));
  double sc, cc;
  sincos(2 * chi, &sc, &cc);
  double sk = sc, ck = cc, phi = chi;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    phi += tm.delta[j] * sk;
    double sn = sk * cc + ck * sc, cn = ck * cc - sk * sc;
    sk = sn; ck = cn;
  }
  return phi;
}

// Same projection, returning sin(lat) and cos(lat) directly (what reference src/PoseUKF.cpp:30 consumes):
// sin(chi) = sin(xi') / cosh(eta') gives sin/cos(2 chi) algebraically, and lat = chi + eps with the small series sum
// eps (|eps| < 4e-3) is folded in by the angle-addition formulas -- no asin and only one small-argument sincos.
__device__ __forceinline__ void tmerc_nav_to_sincos_lat(const TMercConst &tm, double x, double y, double &sin_lat, double &cos_lat) {
  const double dx = x * tm.Ainv;
  double xi = dx + tm.xi0;
  double eta = -y * tm.Ainv;
  // within ~400 km of the projection origin every transcendental argument is small: Taylor polynomials around the origin
  // (sin/cos(xi0) come from the host) replace sincos / exp; outside that range the generic routines are used
  const bool near = fabs(dx) <= 0.0625 && fabs(eta) <= 0.0625;
  double s1, c1, sh2, ch2;
  if (near) {
    double sd_, cd_;
    sincos_small(dx, sd_, cd_);
    s1 = fma(tm.sin_xi0, cd_, tm.cos_xi0 * sd_);
    c1 = fma(tm.cos_xi0, cd_, -(tm.sin_xi0 * sd_));
    sinhcosh_small(2 * eta, sh2, ch2);
  } else {
    sincos(xi, &s1, &c1);
    double e2 = exp(2 * eta), ie2 = fast_rcp(e2);
    sh2 = 0.5 * (e2 - ie2); ch2 = 0.5 * (e2 + ie2);
  }
  double s2 = 2 * s1 * c1, c2 = fma(-2 * s1, s1, 1.0);
  double sj = s2, cj = c2, shj = sh2, chj = ch2;
  double dxi = 0, deta = 0;     // xi' = xi - dxi, eta' = eta - deta
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    dxi = fma(tm.beta[j] * sj, chj, dxi);
    deta = fma(tm.beta[j] * cj, shj, deta);
    double sn = sj * c2 + cj * s2, cn = cj * c2 - sj * s2;
    double shn = shj * ch2 + chj * sh2, chn = chj * ch2 + shj * sh2;
    sj = sn; cj = cn; shj = shn; chj = chn;
  }
  double sd, cd, chp;
  if (near && fabs(dxi) <= 0.0625) {
    double shp;
    sincos_small(dxi, sd, cd);                   // |dxi| < 1e-3
    sinhcosh_small(eta - deta, shp, chp);        // cosh(eta')
  } else {
    sincos(dxi, &sd, &cd);
    double ep = exp(eta - deta);
    chp = 0.5 * (ep + fast_rcp(ep));
  }
  double sxp = s1 * cd - c1 * sd;              // sin(xi')
  double sch = sxp * fast_rcp(chp);            // sin(chi)
  double cch = sqrt(fma(-sch, sch, 1.0));      // cos(chi), |chi| < pi/2
  double sc = 2 * sch * cch, cc = fma(-2 * sch, sch, 1.0);
  double sk = sc, ck = cc, eps = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    eps = fma(tm.delta[j], sk, eps);
    double sn = sk * cc + ck * sc, cn = ck * cc - sk * sc;
    sk = sn; ck = cn;
  }
  double se, ce;
  if (fabs(eps) <= 0.0625) sincos_small(eps, se, ce);   // |eps| < 4e-3
  else sincos(eps, &se, &ce);
  sin_lat = sch * ce + cch * se;
  cos_lat = cch * ce - sch * se;
}

// The `near` path of tmerc_nav_to_sincos_lat as STRAIGHT-LINE code (same operations in the same order: bit-identical results
// where it applies); returns false when an argument leaves the range of the small-argument polynomials, and the caller then
// calls the generic routine.  Without branches the whole evaluation is one basic block, which lets the instruction scheduler
// interleave its long dependent chain with whatever the caller has around it (the panel Cholesky of the predict).
__device__ __forceinline__ bool tmerc_nav_to_sincos_lat_near(const TMercConst &tm, double x, double y, double &sin_lat, double &cos_lat) {
  const double dx = x * tm.Ainv;
  double eta = -y * tm.Ainv;
  bool ok = fabs(dx) <= 0.0625 && fabs(eta) <= 0.0625;
  double s1, c1, sh2, ch2;
  {
    double sd_, cd_;
    sincos_small(dx, sd_, cd_);
    s1 = fma(tm.sin_xi0, cd_, tm.cos_xi0 * sd_);
    c1 = fma(tm.cos_xi0, cd_, -(tm.sin_xi0 * sd_));
    sinhcosh_small(2 * eta, sh2, ch2);
  }
  double s2 = 2 * s1 * c1, c2 = fma(-2 * s1, s1, 1.0);
  double sj = s2, cj = c2, shj = sh2, chj = ch2;
  double dxi = 0, deta = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    dxi = fma(tm.beta[j] * sj, chj, dxi);
    deta = fma(tm.beta[j] * cj, shj, deta);
    double sn = sj * c2 + cj * s2, cn = cj * c2 - sj * s2;
    double shn = shj * ch2 + chj * sh2, chn = chj * ch2 + shj * sh2;
    sj = sn; cj = cn; shj = shn; chj = chn;
  }
  ok = ok && fabs(dxi) <= 0.0625;
  double sd, cd, chp, shp;
  sincos_small(dxi, sd, cd);
  sinhcosh_small(eta - deta, shp, chp);
  double sxp = s1 * cd - c1 * sd;
  double sch = sxp * fast_rcp(chp);
  double cch = sqrt(fma(-sch, sch, 1.0));
  double sc = 2 * sch * cch, cc = fma(-2 * sch, sch, 1.0);
  double sk = sc, ck = cc, eps = 0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    eps = fma(tm.delta[j], sk, eps);
    double sn = sk * cc + ck * sc, cn = ck * cc - sk * sc;
    sk = sn; ck = cn;
  }
  ok = ok && fabs(eps) <= 0.0625;
  double se, ce;
  sincos_small(eps, se, ce);
  sin_lat = sch * ce + cch * se;
  cos_lat = cch * ce - sch * se;
  return ok;
}

// The same evaluation cut into STAGES of a few dependent operations each.  The predict calls one stage per column of its
// panel Cholesky (from column 1 on): both are single long dependent chains, and side by side in the instruction stream they
// fill each other's latency (ptxas interleaves neighbouring independent chains, it does not pull a 700-cycle chain apart).
struct TMercStages {
  double eta, s1, c1, s2, c2, sh2, ch2, sj, cj, shj, chj, dxi, deta, sd, cd, chp, sch, cch, sc, cc, sk, ck, eps;
  bool ok;
  double sin_lat, cos_lat;
  template <int J>
  __device__ __forceinline__ void series1(const TMercConst &tm) {
    dxi = fma(tm.beta[J] * sj, chj, dxi);
    deta = fma(tm.beta[J] * cj, shj, deta);
    const double sn = sj * c2 + cj * s2, cn = cj * c2 - sj * s2;
    const double shn = shj * ch2 + chj * sh2, chn = chj * ch2 + shj * sh2;
    sj = sn; cj = cn; shj = shn; chj = chn;
  }
  template <int J>
  __device__ __forceinline__ void series2(const TMercConst &tm) {
    eps = fma(tm.delta[J], sk, eps);
    const double sn = sk * cc + ck * sc, cn = ck * cc - sk * sc;
    sk = sn; ck = cn;
  }
  template <int STAGE>
  __device__ __forceinline__ void step(const TMercConst &tm, double x, double y) {
    if constexpr (STAGE == 0) {
      const double dx = x * tm.Ainv;
      eta = -y * tm.Ainv;
      ok = fabs(dx) <= 0.0625 && fabs(eta) <= 0.0625;
      double sd_, cd_;
      sincos_small(dx, sd_, cd_);
      s1 = fma(tm.sin_xi0, cd_, tm.cos_xi0 * sd_);
      c1 = fma(tm.cos_xi0, cd_, -(tm.sin_xi0 * sd_));
      sinhcosh_small(2 * eta, sh2, ch2);
    } else if constexpr (STAGE == 1) {
      s2 = 2 * s1 * c1; c2 = fma(-2 * s1, s1, 1.0);
      sj = s2; cj = c2; shj = sh2; chj = ch2; dxi = 0; deta = 0;
      series1<0>(tm);
    } else if constexpr (STAGE == 2) {
      series1<1>(tm); series1<2>(tm);
    } else if constexpr (STAGE == 3) {
      series1<3>(tm); series1<4>(tm);
    } else if constexpr (STAGE == 4) {
      series1<5>(tm);
      ok = ok && fabs(dxi) <= 0.0625;
      double shp;
      sincos_small(dxi, sd, cd);
      sinhcosh_small(eta - deta, shp, chp);
    } else if constexpr (STAGE == 5) {
      const double sxp = s1 * cd - c1 * sd;
      sch = sxp * fast_rcp(chp);
    } else if constexpr (STAGE == 6) {
      const double x2 = fma(-sch, sch, 1.0);        // cos^2(chi) > 0
      const double y = fast_rsqrt3(x2);
      double r = x2 * y;
      r = fma(fma(-r, r, x2), 0.5 * y, r);           // one Newton step on the root: sqrt() to an ulp without its slow-path branch
      cch = r;
      sc = 2 * sch * cch; cc = fma(-2 * sch, sch, 1.0);
      sk = sc; ck = cc; eps = 0;
    } else if constexpr (STAGE == 7) {
      series2<0>(tm); series2<1>(tm);
    } else if constexpr (STAGE == 8) {
      series2<2>(tm); series2<3>(tm);
    } else if constexpr (STAGE == 9) {
      series2<4>(tm); series2<5>(tm);
      ok = ok && fabs(eps) <= 0.0625;
    } else if constexpr (STAGE == 10) {
      double se, ce;
      sincos_small(eps, se, ce);
      sin_lat = sch * ce + cch * se;
      cos_lat = cch * ce - sch * se;
    }
  }
};

// (lat, lon) -> NWU (x, y)   (reference src/PoseUKF.cpp:573 worldToNav)
__device__ __forceinline__ void tmerc_world_to_nav(const TMercConst &tm, double lat, double lon, double &x, double &y) {
  double tau = tan(lat);
  double sig = sinh(tm.e * atanh(tm.e * tau / sqrt(1 + tau * tau)));
  double taup = tau * sqrt(1 + sig * sig) - sig * sqrt(1 + tau * tau);
  double dl = lon - tm.lon0;
  double cdl = cos(dl);
  double xip = atan2(taup, cdl);
  double etap = asinh(sin(dl) / sqrt(taup * taup + cdl * cdl));
  double xi = xip, eta = etap;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double a2 = 2.0 * (j + 1);
    xi += tm.alpha[j] * sin(a2 * xip) * cosh(a2 * etap);
    eta += tm.alpha[j] * cos(a2 * xip) * sinh(a2 * etap);
  }
  x = tm.A * (xi - tm.xi0);
  y = -tm.A * eta;
}

}  // namespace uwvkf
