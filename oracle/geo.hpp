// ORACLE (test infrastructure, NOT product code) -- parity unpinned.
//
// Restatement of pose_estimation::GeographicProjection / GravitationalModel / EARTHW (slam/pose_estimation,
// un-vendored, no pinned version).  Call sites in the reference: navToWorld per sigma point at
// src/PoseUKF.cpp:29 and in getRotationRate :696; worldToNav :573; WGS_84 gravity :301; EARTHW :30,:697.
// Frozen definition: local transverse Mercator on WGS-84 centred at (lat0, lon0),
// k0 = 1, Krueger n-series to order n^6 (Karney 2011, eqs. 35/36) for the rectifying<->conformal step and
// for conformal->geodetic latitude; navigation axes NWU: x = northing, y = -easting.
#pragma once
#include <cmath>

namespace orc {

constexpr double kEarthW = 7.292115e-5;  // rad/s
constexpr double kWgsA = 6378137.0;
constexpr double kWgsF = 1.0 / 298.257223563;

struct TMerc {
  double lat0, lon0;
  double A;         // rectifying radius
  double xi0;       // xi of the origin latitude on the central meridian
  double e;         // first eccentricity
  double alpha[6], beta[6], delta[6];

  TMerc(double lat0_, double lon0_) : lat0(lat0_), lon0(lon0_) {
    const double n = kWgsF / (2 - kWgsF);
    const double n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
    e = std::sqrt(kWgsF * (2 - kWgsF));
    A = kWgsA / (1 + n) * (1 + n2 / 4 + n4 / 64 + n6 / 256);
    alpha[0] = n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800;
    alpha[1] = 13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360;
    alpha[2] = 61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440;
    alpha[3] = 49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600;
    alpha[4] = 34729 * n5 / 80640 - 3418889 * n6 / 1995840;
    alpha[5] = 212378941 * n6 / 319334400;
    beta[0] = n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800;
    beta[1] = n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720;
    beta[2] = 17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720;
    beta[3] = 4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600;
    beta[4] = 4583 * n5 / 161280 - 108847 * n6 / 3991680;
    beta[5] = 20648693 * n6 / 638668800;
    // conformal -> geodetic latitude:  phi = chi + sum delta_j sin(2 j chi)
    delta[0] = 2 * n - 2 * n2 / 3 - 2 * n3 + 116 * n4 / 45 + 26 * n5 / 45 - 2854 * n6 / 675;
    delta[1] = 7 * n2 / 3 - 8 * n3 / 5 - 227 * n4 / 45 + 2704 * n5 / 315 + 2323 * n6 / 945;
    delta[2] = 56 * n3 / 15 - 136 * n4 / 35 - 1262 * n5 / 105 + 73814 * n6 / 2835;
    delta[3] = 4279 * n4 / 630 - 332 * n5 / 35 - 399572 * n6 / 14175;
    delta[4] = 4174 * n5 / 315 - 144838 * n6 / 6237;
    delta[5] = 601676 * n6 / 22275;
    double chi0 = conformal_lat(lat0);
    double s = chi0;
    for (int j = 0; j < 6; ++j) s += alpha[j] * std::sin(2 * (j + 1) * chi0);
    xi0 = s;
  }

  // closed-form geodetic -> conformal latitude (Karney 2011 eqs. 7-9)
  double conformal_lat(double phi) const {
    double tau = std::tan(phi);
    double sig = std::sinh(e * std::atanh(e * tau / std::sqrt(1 + tau * tau)));
    double taup = tau * std::sqrt(1 + sig * sig) - sig * std::sqrt(1 + tau * tau);
    return std::atan(taup);
  }

  // (lat, lon) [rad] -> NWU (x north, y west) [m]
  void worldToNav(double lat, double lon, double &x, double &y) const {
    double tau = std::tan(lat);
    double sig = std::sinh(e * std::atanh(e * tau / std::sqrt(1 + tau * tau)));
    double taup = tau * std::sqrt(1 + sig * sig) - sig * std::sqrt(1 + tau * tau);
    double dl = lon - lon0;
    double xip = std::atan2(taup, std::cos(dl));
    double etap = std::asinh(std::sin(dl) / std::sqrt(taup * taup + std::cos(dl) * std::cos(dl)));
    double xi = xip, eta = etap;
    for (int j = 0; j < 6; ++j) {
      xi += alpha[j] * std::sin(2 * (j + 1) * xip) * std::cosh(2 * (j + 1) * etap);
      eta += alpha[j] * std::cos(2 * (j + 1) * xip) * std::sinh(2 * (j + 1) * etap);
    }
    x = A * (xi - xi0);
    y = -A * eta;
  }

  // NWU (x, y) [m] -> (lat, lon) [rad]
  void navToWorld(double x, double y, double &lat, double &lon) const {
    double xi = x / A + xi0;
    double eta = -y / A;
    double xip = xi, etap = eta;
    for (int j = 0; j < 6; ++j) {
      xip -= beta[j] * std::sin(2 * (j + 1) * xi) * std::cosh(2 * (j + 1) * eta);
      etap -= beta[j] * std::cos(2 * (j + 1) * xi) * std::sinh(2 * (j + 1) * eta);
    }
    double chi = std::asin(std::sin(xip) / std::cosh(etap));
    double phi = chi;
    for (int j = 0; j < 6; ++j) phi += delta[j] * std::sin(2 * (j + 1) * chi);
    lat = phi;
    lon = lon0 + std::atan2(std::sinh(etap), std::cos(xip));
  }
};

// pose_estimation::GravitationalModel::WGS_84(latitude, altitude): Somigliana normal gravity on the
// ellipsoid with a free-air (a/(a+h))^2 height scaling.
inline double gravity_wgs84(double lat, double alt) {
  const double g0 = 9.7803267714, k = 0.00193185138639, e = 0.0818191908426;
  double s2 = std::sin(lat) * std::sin(lat);
  double g = g0 * (1 + k * s2) / std::sqrt(1 - e * e * s2);
  double r = kWgsA / (kWgsA + alt);
  return g * r * r;
}

}  // namespace orc
