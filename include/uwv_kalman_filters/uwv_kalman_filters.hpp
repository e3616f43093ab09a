// Thin C++ host classes with the reference's names and call signatures on top of the uwvkf C ABI.
//
//   uwv_kalman_filters::PoseUKF          (reference src/PoseUKF.hpp:40-208)
//   uwv_kalman_filters::VelocityUKF      (reference src/VelocityUKF.hpp:33-66)
//   uwv_kalman_filters::BottomUKF        (reference src/BottomUKF.hpp:27-52)
//   uwv_kalman_filters::IndirectPoseUKF  (reference src/IndirectPoseUKF.hpp:29-85)
//
// Every object owns one uwvkf handle.  Constructed the reference way it holds ONE filter (batch = 1) and
// behaves like the reference object; the extra `batch` constructor argument turns the same object into B
// independent filters that take B-row inputs.  All arithmetic happens on the GPU: the methods only marshal
// arguments.  Errors surface as std::runtime_error exactly where the reference throws (NaN measurements via
// checkMeasurment, "Motion model is not initialized!").
//
// Eigen and Boost are not required: the small fixed-size value types below stand in for Eigen::Vector3d etc.
// (construct them from `const double*`, e.g. `Vector3d(v.data())` for an Eigen vector; matrices are row-major,
// so pass `Eigen::Matrix<double,N,N,Eigen::RowMajor>` data or a symmetric matrix).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../uwvkf.h"

namespace uwv_kalman_filters {

template <int R, int C = 1>
struct Mat {
  std::array<double, R * C> v{};
  Mat() = default;
  explicit Mat(const double *p) { std::memcpy(v.data(), p, sizeof(double) * R * C); }
  Mat(std::initializer_list<double> l) { int k = 0; for (double x : l) if (k < R * C) v[k++] = x; }
  double &operator()(int r, int c = 0) { return v[r * C + c]; }
  double operator()(int r, int c = 0) const { return v[r * C + c]; }
  double &operator[](int k) { return v[k]; }
  double operator[](int k) const { return v[k]; }
  const double *data() const { return v.data(); }
  double *data() { return v.data(); }
  static Mat Zero() { return Mat(); }
  static Mat Identity() { Mat m; for (int k = 0; k < (R < C ? R : C); ++k) m(k, k) = 1.0; return m; }
};
using Vector2d = Mat<2>;
using Vector3d = Mat<3>;
using Vector6d = Mat<6>;
using Matrix2d = Mat<2, 2>;
using Matrix3d = Mat<3, 3>;
using Matrix6d = Mat<6, 6>;
// unit quaternion, coefficient order (x, y, z, w) like Eigen::Quaterniond::coeffs()
struct Quaterniond {
  double x = 0, y = 0, z = 0, w = 1;
  static Quaterniond Identity() { return Quaterniond(); }
  Vector3d rotate(const Vector3d &p) const {
    double tx = 2 * (y * p[2] - z * p[1]), ty = 2 * (z * p[0] - x * p[2]), tz = 2 * (x * p[1] - y * p[0]);
    return Vector3d{p[0] + w * tx + (y * tz - z * ty), p[1] + w * ty + (z * tx - x * tz), p[2] + w * tz + (x * ty - y * tx)};
  }
  Matrix3d toRotationMatrix() const {
    Matrix3d R;
    for (int c = 0; c < 3; ++c) { Vector3d e; e[c] = 1; Vector3d r = rotate(e); for (int k = 0; k < 3; ++k) R(k, c) = r[k]; }
    return R;
  }
};
// rigid transform (Eigen::Affine3d stand-in): rotation + translation
struct Affine3d {
  Quaterniond rotation;
  Vector3d translation;
  static Affine3d Identity() { return Affine3d(); }
};

// pose_estimation::Measurement / MEASUREMENT(NAME, DIM): one distinct type per measurement name
#define UWVKF_MEASUREMENT(NAME, DIM)                                   \
  struct NAME {                                                        \
    typedef Mat<DIM> Mu;                                               \
    typedef Mat<DIM, DIM> Cov;                                         \
    Mu mu;                                                             \
    Cov cov;                                                           \
  };

// ---- configuration structs (reference src/PoseUKFConfig.hpp:20-194) -------------------------------------------
struct WaterVelocityParameters { double tau, limits; Vector3d measurement_std; double scale, cell_size, first_cell_blank, minimum_correlation, adcp_bias_tau, adcp_bias_limits; };
struct InertialNoiseParameters { Vector3d randomwalk, bias_offset, bias_instability; double bias_tau; };
struct DynamicModelNoiseParameters {
  Vector6d body_efforts_std;
  Mat<9> inertia_instability, lin_damping_instability, quad_damping_instability;
  double inertia_tau, lin_damping_tau, quad_damping_tau;
};
struct LocationConfiguration { double latitude, longitude, altitude; };
struct CameraConfiguration { double fx, fy, cx, cy; };
struct HydrostaticConfiguration { double water_density, water_density_limits, water_density_tau, atmospheric_pressure, pressure_std; };
struct PoseUKFConfig {
  InertialNoiseParameters acceleration, rotation_rate;
  DynamicModelNoiseParameters model_noise_parameters;
  WaterVelocityParameters water_velocity;
  LocationConfiguration location;
  HydrostaticConfiguration hydrostatics;
  Vector3d max_jerk;
  Vector6d max_effort;
  double dynamic_model_min_depth;
};
// uwv_dynamic_model::UWVParameters (fields used by the reference: src/PoseUKF.cpp:160-171,303-314)
struct UWVParameters {
  Matrix6d inertia_matrix;
  std::array<Matrix6d, 2> damping_matrices;  // [0] linear, [1] quadratic
  Vector3d distance_body2centerofgravity, distance_body2centerofbuoyancy;
  double weight = 0, buoyancy = 0;
  std::array<double, UWVKF_UWV_NPARAMS> flat() const {
    std::array<double, UWVKF_UWV_NPARAMS> p{};
    std::memcpy(&p[0], inertia_matrix.data(), 36 * sizeof(double));
    std::memcpy(&p[36], damping_matrices[0].data(), 36 * sizeof(double));
    std::memcpy(&p[72], damping_matrices[1].data(), 36 * sizeof(double));
    std::memcpy(&p[108], distance_body2centerofgravity.data(), 3 * sizeof(double));
    std::memcpy(&p[111], distance_body2centerofbuoyancy.data(), 3 * sizeof(double));
    p[114] = weight; p[115] = buoyancy;
    return p;
  }
};

namespace detail {
inline void check(int rc) { if (rc != UWVKF_OK) throw std::runtime_error(std::string("uwvkf: ") + uwvkf_last_error()); }
// pose_estimation::GravitationalModel::WGS_84 (reference src/PoseUKF.cpp:301)
inline double gravity_wgs84(double lat, double alt) {
  const double g0 = 9.7803267714, k = 0.00193185138639, e = 0.0818191908426, a = 6378137.0;
  double s2 = std::sin(lat) * std::sin(lat);
  return g0 * (1 + k * s2) / std::sqrt(1 - e * e * s2) * (a / (a + alt)) * (a / (a + alt));
}
}  // namespace detail

// pose_estimation::UnscentedKalmanFilter<State> stand-in: handle ownership and the inherited interface
template <int DIM, int DOF>
class FilterBase {
 public:
  typedef Mat<DIM> State;               // stored state vector (layouts: include/uwvkf.h)
  typedef Mat<DOF, DOF> Covariance;
  virtual ~FilterBase() { uwvkf_destroy(h_); }
  FilterBase(const FilterBase &) = delete;
  FilterBase &operator=(const FilterBase &) = delete;
  void predictionStep(double delta_t) { detail::check(uwvkf_predict(h_, delta_t)); }
  // Caller-side helper of the base class [pose_estimation, EXT-RECALLED: SURVEY.md 8f-4]: derive delta_t from consecutive sample
  // times (seconds).  The first call only latches the time; a sample older than the last one, or further ahead than
  // max_time_delta, is refused (false) and the filter is not stepped.
  bool predictionStepFromSampleTime(double sample_time, double max_time_delta = 2.0) {
    if (!have_time_) { last_sample_time_ = sample_time; have_time_ = true; return true; }
    const double delta = sample_time - last_sample_time_;
    if (!(delta >= 0.0) || delta > max_time_delta) return false;
    if (delta > 0.0) predictionStep(delta);
    last_sample_time_ = sample_time;
    return true;
  }
  bool getCurrentState(State &state, int64_t filter = 0) const {
    return uwvkf_get_state(h_, state.data(), nullptr, filter, 1, UWVKF_HOST) == UWVKF_OK;
  }
  bool getCurrentStateAndCovariance(State &state, Covariance &cov, int64_t filter = 0) const {
    return uwvkf_get_state(h_, state.data(), cov.data(), filter, 1, UWVKF_HOST) == UWVKF_OK;
  }
  void setProcessNoiseCovariance(const Covariance &Q) { detail::check(uwvkf_set_process_noise(h_, Q.data())); }
  std::vector<uint32_t> status() const {
    std::vector<uint32_t> s((size_t)batch_);
    detail::check(uwvkf_get_status(h_, s.data(), 0, batch_));
    return s;
  }
  int64_t batch() const { return batch_; }
  uwvkf_handle handle() const { return h_; }

 protected:
  FilterBase(int kind, int64_t batch, int device) : batch_(batch) { detail::check(uwvkf_create(kind, batch, device, &h_)); }
  void initializeFilter(const State &state, const Covariance &cov) {
    std::vector<double> mu((size_t)batch_ * DIM);
    for (int64_t f = 0; f < batch_; ++f) std::memcpy(&mu[f * DIM], state.data(), DIM * sizeof(double));
    detail::check(uwvkf_set_state(h_, mu.data(), cov.data(), 0, 0, batch_, UWVKF_HOST));
  }
  uwvkf_handle h_ = nullptr;
  int64_t batch_;
  double last_sample_time_ = 0.0;
  bool have_time_ = false;
};

// ==============================================================================================================
class PoseUKF : public FilterBase<UWVKF_POSE_DIM, UWVKF_POSE_DOF> {
 public:
  struct PoseUKFParameter {   // reference src/PoseUKF.hpp:46-76
    Vector3d imu_in_body, gyro_bias_offset;
    double gyro_bias_tau = 0;
    Vector3d acc_bias_offset;
    double acc_bias_tau = 0, inertia_tau = 0, lin_damping_tau = 0, quad_damping_tau = 0, water_velocity_tau = 0,
           water_velocity_limits = 0, water_velocity_scale = 0, adcp_bias_tau = 0, atmospheric_pressure = 0, water_density_tau = 0;
  };
  UWVKF_MEASUREMENT(GeographicPosition, 2)
  UWVKF_MEASUREMENT(XY_Position, 2)
  UWVKF_MEASUREMENT(Z_Position, 1)
  UWVKF_MEASUREMENT(Pressure, 1)
  UWVKF_MEASUREMENT(RotationRate, 3)
  UWVKF_MEASUREMENT(Acceleration, 3)
  UWVKF_MEASUREMENT(Velocity, 3)
  UWVKF_MEASUREMENT(WaterVelocityMeasurement, 2)
  UWVKF_MEASUREMENT(BodyEffortsMeasurement, 6)
  UWVKF_MEASUREMENT(VisualFeatureMeasurement, 2)

  // reference src/PoseUKF.cpp:288-372: state and covariance derived from the configuration
  PoseUKF(const Vector3d &imu_in_nwu_pos, const Matrix3d &imu_in_nwu_pos_cov, const Quaterniond &imu_in_nwu_rot,
          const Matrix3d &imu_in_nwu_rot_cov, const PoseUKFConfig &cfg, const UWVParameters &model_parameters,
          const Affine3d &imu_in_body = Affine3d::Identity(), int64_t batch = 1, int device = 0)
      : FilterBase(UWVKF_FILTER_POSE, batch, device) {
    const Matrix3d Rib = imu_in_body.rotation.toRotationMatrix();
    auto rot = [&](const Vector3d &p) { return imu_in_body.rotation.rotate(p); };
    State s;
    Covariance P;
    for (int k = 0; k < 3; ++k) s[k] = imu_in_nwu_pos[k];
    s[3] = imu_in_nwu_rot.x; s[4] = imu_in_nwu_rot.y; s[5] = imu_in_nwu_rot.z; s[6] = imu_in_nwu_rot.w;
    Vector3d bg = rot(cfg.rotation_rate.bias_offset), ba = rot(cfg.acceleration.bias_offset);
    for (int k = 0; k < 3; ++k) { s[13 + k] = bg[k]; s[16 + k] = ba[k]; }
    s[19] = detail::gravity_wgs84(cfg.location.latitude, cfg.location.altitude);
    const int map[3] = {0, 1, 5};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {   // 3x3 blocks are MTK::matrix<3,3>, column-major (src/PoseState.hpp:37)
        s[20 + c * 3 + r] = model_parameters.inertia_matrix(map[r], map[c]);
        s[29 + c * 3 + r] = model_parameters.damping_matrices[0](map[r], map[c]);
        s[38 + c * 3 + r] = model_parameters.damping_matrices[1](map[r], map[c]);
      }
    s[53] = cfg.hydrostatics.water_density;
    auto rdr = [&](const Vector3d &d, double scale, int off, Covariance &M) {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double a = 0;
          for (int k = 0; k < 3; ++k) a += Rib(i, k) * (scale * d[k] * d[k]) * Rib(j, k);
          M(off + i, off + j) = a;
        }
    };
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { P(i, j) = imu_in_nwu_pos_cov(i, j); P(3 + i, 3 + j) = imu_in_nwu_rot_cov(i, j); }
    for (int i = 0; i < 3; ++i) { P(6 + i, 6 + i) = 1.0; P(9 + i, 9 + i) = 10.0; }
    rdr(cfg.rotation_rate.bias_instability, 1.0, 12, P);
    rdr(cfg.acceleration.bias_instability, 1.0, 15, P);
    P(18, 18) = std::pow(0.05, 2.);
    for (int k = 0; k < 9; ++k) {
      P(19 + k, 19 + k) = std::pow(cfg.model_noise_parameters.inertia_instability[k], 2);
      P(28 + k, 28 + k) = std::pow(cfg.model_noise_parameters.lin_damping_instability[k], 2);
      P(37 + k, 37 + k) = std::pow(cfg.model_noise_parameters.quad_damping_instability[k], 2);
    }
    for (int k = 0; k < 2; ++k) {
      P(46 + k, 46 + k) = P(48 + k, 48 + k) = std::pow(cfg.water_velocity.limits, 2);
      P(50 + k, 50 + k) = std::pow(cfg.water_velocity.adcp_bias_limits, 2);
    }
    P(52, 52) = std::pow(cfg.hydrostatics.water_density_limits, 2.);
    PoseUKFParameter fp;
    fp.imu_in_body = imu_in_body.translation;
    fp.acc_bias_tau = cfg.acceleration.bias_tau; fp.acc_bias_offset = ba;
    fp.gyro_bias_tau = cfg.rotation_rate.bias_tau; fp.gyro_bias_offset = bg;
    fp.inertia_tau = cfg.model_noise_parameters.inertia_tau; fp.lin_damping_tau = cfg.model_noise_parameters.lin_damping_tau;
    fp.quad_damping_tau = cfg.model_noise_parameters.quad_damping_tau; fp.water_velocity_tau = cfg.water_velocity.tau;
    fp.water_velocity_limits = cfg.water_velocity.limits; fp.water_velocity_scale = cfg.water_velocity.scale;
    fp.adcp_bias_tau = cfg.water_velocity.adcp_bias_tau; fp.atmospheric_pressure = cfg.hydrostatics.atmospheric_pressure;
    fp.water_density_tau = cfg.hydrostatics.water_density_tau;
    setParameters(cfg.location, fp);
    setModelParameters(model_parameters);   // dynamic_model.reset(new DynamicModel()); setUWVParameters(...)  :353-354
    initializeFilter(s, P);
  }
  // reference src/PoseUKF.cpp:374-391: full state + covariance + persistent parameters
  PoseUKF(const State &initial_state, const Covariance &state_cov, const LocationConfiguration &location,
          const UWVParameters &model_parameters, const PoseUKFParameter &filter_parameter, int64_t batch = 1, int device = 0)
      : FilterBase(UWVKF_FILTER_POSE, batch, device) {
    setParameters(location, filter_parameter);
    setModelParameters(model_parameters);   // :382-383
    initializeFilter(initial_state, state_cov);
  }

  // reference src/PoseUKF.cpp:393-439
  void setProcessNoiseFromConfig(const PoseUKFConfig &cfg, double T, const Quaterniond &imu_in_body = Quaterniond::Identity()) {
    const Matrix3d Rib = imu_in_body.toRotationMatrix();
    Covariance Q;
    auto rdr = [&](const Vector3d &d, double scale, int off) {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double a = 0;
          for (int k = 0; k < 3; ++k) a += Rib(i, k) * (scale * d[k] * d[k]) * Rib(j, k);
          Q(off + i, off + j) = a;
        }
    };
    for (int k = 0; k < 3; ++k) {
      double j = cfg.max_jerk[k], a = (1. / 6.) * 0.25 * j, b = 0.5 * 0.25 * j, c = 0.25 * j;
      Q(k, k) = 1.5 * (std::pow(T, 4.0) * a * a);
      Q(6 + k, 6 + k) = 1.5 * (std::pow(T, 2.0) * b * b);
      Q(9 + k, 9 + k) = c * c;
    }
    rdr(cfg.rotation_rate.randomwalk, 1.0, 3);
    rdr(cfg.rotation_rate.bias_instability, 2. / (cfg.rotation_rate.bias_tau * T), 12);
    rdr(cfg.acceleration.bias_instability, 2. / (cfg.acceleration.bias_tau * T), 15);
    Q(18, 18) = 1.e-12;
    const DynamicModelNoiseParameters &m = cfg.model_noise_parameters;
    for (int k = 0; k < 9; ++k) {
      Q(19 + k, 19 + k) = (2. / (m.inertia_tau * T)) * std::pow(m.inertia_instability[k], 2);
      Q(28 + k, 28 + k) = (2. / (m.lin_damping_tau * T)) * std::pow(m.lin_damping_instability[k], 2);
      Q(37 + k, 37 + k) = (2. / (m.quad_damping_tau * T)) * std::pow(m.quad_damping_instability[k], 2);
    }
    for (int k = 0; k < 2; ++k) {
      Q(46 + k, 46 + k) = Q(48 + k, 48 + k) = (2. / (cfg.water_velocity.tau * T)) * std::pow(cfg.water_velocity.limits, 2);
      Q(50 + k, 50 + k) = (2. / (cfg.water_velocity.adcp_bias_tau * T)) * std::pow(cfg.water_velocity.adcp_bias_limits, 2);
    }
    Q(52, 52) = (2. / (cfg.hydrostatics.water_density_tau * T)) * std::pow(cfg.hydrostatics.water_density_limits, 2.);
    setProcessNoiseCovariance(Q);
  }

  // integrateMeasurement overloads, reference src/PoseUKF.hpp:137-177.  For batch > 1 use the *Batch forms.
  void integrateMeasurement(const GeographicPosition &m, const Vector3d &gps_in_body = Vector3d::Zero()) {
    update(UWVKF_POSE_GEOGRAPHIC, m.mu.data(), m.cov.data(), gps_in_body.data());
  }
  void integrateMeasurement(const XY_Position &m) { update(UWVKF_POSE_XY, m.mu.data(), m.cov.data(), nullptr); }
  void integrateDelayedPositionMeasurement(const XY_Position &m, const Vector2d &delayed_position) {
    update(UWVKF_POSE_DELAYED_XY, m.mu.data(), m.cov.data(), delayed_position.data());
  }
  void integrateDelayedPositionMeasurementWithStateAugmentation(const XY_Position &, const Vector2d &, const Matrix2d &) {
    // declared at src/PoseUKF.hpp:144 but its definition is commented out in the reference (src/PoseUKF.cpp:529-557)
    throw std::runtime_error("integrateDelayedPositionMeasurementWithStateAugmentation is not implemented in the reference");
  }
  void integrateMeasurement(const Z_Position &m) { update(UWVKF_POSE_Z, m.mu.data(), m.cov.data(), nullptr); }
  void integrateMeasurement(const Pressure &m, const Vector3d &pressure_sensor_in_imu = Vector3d::Zero()) {
    update(UWVKF_POSE_PRESSURE, m.mu.data(), m.cov.data(), pressure_sensor_in_imu.data());
  }
  void integrateMeasurement(const RotationRate &m) {
    for (double c : m.cov.v) if (std::isnan(c)) throw std::runtime_error("uwvkf: measurement covariance contains NaN");
    detail::check(uwvkf_pose_set_rotation_rate(h_, tile(m.mu.data(), 3).data(), UWVKF_HOST));
  }
  void integrateMeasurement(const Acceleration &m) { update(UWVKF_POSE_ACCELERATION, m.mu.data(), m.cov.data(), nullptr); }
  void integrateMeasurement(const Velocity &m) { update(UWVKF_POSE_VELOCITY, m.mu.data(), m.cov.data(), nullptr); }
  // reference src/PoseUKF.cpp:581-602 (forces and torques in the body frame; m = 6)
  void integrateMeasurement(const BodyEffortsMeasurement &m, bool only_affect_velocity = false) {
    update(only_affect_velocity ? UWVKF_POSE_BODY_EFFORTS_VELOCITY_ONLY : UWVKF_POSE_BODY_EFFORTS, m.mu.data(), m.cov.data(), nullptr);
  }
  // reference src/PoseUKF.cpp:613-654 (features of a visual marker in undistorted image coordinates)
  void integrateMeasurement(const std::vector<VisualFeatureMeasurement> &marker_features, const std::vector<Vector3d> &feature_positions,
                            const Affine3d &marker_pose, const Matrix6d &cov_marker_pose, const CameraConfiguration &camera_config,
                            const Affine3d &camera_in_IMU) {
    // the reference loops with `||` over both vectors (src/PoseUKF.cpp:633): unequal sizes read out of bounds there
    if (marker_features.size() != feature_positions.size()) throw std::runtime_error("uwvkf: marker_features and feature_positions must have equal size");
    const int n = (int)marker_features.size();
    std::vector<double> px((size_t)batch_ * n * 2), pc((size_t)n * 4), fp((size_t)n * 3);
    for (int i = 0; i < n; ++i) {
      std::memcpy(&pc[i * 4], marker_features[i].cov.data(), 4 * sizeof(double));
      std::memcpy(&fp[i * 3], feature_positions[i].data(), 3 * sizeof(double));
      for (int64_t f = 0; f < batch_; ++f) std::memcpy(&px[(f * n + i) * 2], marker_features[i].mu.data(), 2 * sizeof(double));
    }
    auto pack = [](const Affine3d &a, double *p) {
      p[0] = a.translation[0]; p[1] = a.translation[1]; p[2] = a.translation[2];
      p[3] = a.rotation.x; p[4] = a.rotation.y; p[5] = a.rotation.z; p[6] = a.rotation.w;
    };
    double mp[7], ci[7], cam[4] = {camera_config.fx, camera_config.fy, camera_config.cx, camera_config.cy};
    pack(marker_pose, mp);
    pack(camera_in_IMU, ci);
    detail::check(uwvkf_pose_update_features(h_, n, px.data(), pc.data(), fp.data(), mp, cov_marker_pose.data(), cam, ci, nullptr, UWVKF_HOST));
  }
  void integrateMeasurement(const WaterVelocityMeasurement &m, double cell_weighting) {
    update(UWVKF_POSE_WATER_VELOCITY, m.mu.data(), m.cov.data(), &cell_weighting);
  }
  // batched forms: z is [batch][m] row-major, R one shared m x m covariance, valid optional [batch]
  void integrateMeasurementBatch(int meas_kind, const double *z, const double *R, const double *aux = nullptr,
                                 const uint8_t *valid = nullptr) {
    detail::check(uwvkf_pose_update(h_, meas_kind, z, R, 0, aux, valid, UWVKF_HOST));
  }
  void integrateRotationRateBatch(const double *w) { detail::check(uwvkf_pose_set_rotation_rate(h_, w, UWVKF_HOST)); }
  // fused predictionStep + integrateMeasurement(kind): the hot path, one kernel
  void predictionStepAndIntegrate(double delta_t, const double *w, int meas_kind, const double *z, const double *R,
                                  const double *aux = nullptr, const uint8_t *valid = nullptr) {
    detail::check(uwvkf_pose_predict_update(h_, delta_t, w, meas_kind, z, R, 0, aux, valid, UWVKF_HOST));
  }
  // predictionStep + two updates (e.g. DVL + accelerometer, the 100 Hz step of a full PoseUKF) in one call
  void predictionStepAndIntegrate2(double delta_t, const double *w, int kind_a, const double *z_a, const double *R_a, int kind_b,
                                   const double *z_b, const double *R_b, const uint8_t *valid_a = nullptr,
                                   const uint8_t *valid_b = nullptr) {
    detail::check(uwvkf_pose_predict_update2(h_, delta_t, w, kind_a, z_a, R_a, valid_a, kind_b, z_b, R_b, valid_b, UWVKF_HOST));
  }
  // pipelined driver-loop step on HOST sample buffers (uwvkf_pose_step_host): rotation rate + predictionStep + integrateMeasurement,
  // navigation solution [batch][10] back; see include/uwvkf.h for when the buffers may be reused.  synchronize() drains the pipeline.
  void stepHost(double delta_t, const double *w_host, int meas_kind, const double *z_host, const double *R, double *nav_out_host,
                const double *aux = nullptr) {
    detail::check(uwvkf_pose_step_host(h_, delta_t, w_host, meas_kind, z_host, R, aux, nav_out_host));
  }
  void synchronize() { detail::check(uwvkf_synchronize(h_)); }
  void resetFilterWithExternalPose(const Affine3d &imu_in_nav) {   // reference src/PoseUKF.cpp:685-691
    double p[7] = {imu_in_nav.translation[0], imu_in_nav.translation[1], imu_in_nav.translation[2],
                   imu_in_nav.rotation.x, imu_in_nav.rotation.y, imu_in_nav.rotation.z, imu_in_nav.rotation.w};
    detail::check(uwvkf_pose_reset_with_external_pose(h_, tile(p, 7).data()));
  }
  RotationRate::Mu getRotationRate() {   // reference src/PoseUKF.cpp:693-699
    std::vector<double> out((size_t)batch_ * 3);
    detail::check(uwvkf_pose_get_rotation_rate(h_, out.data()));
    return RotationRate::Mu(out.data());
  }

 private:
  void setModelParameters(const UWVParameters &model_parameters) {
    detail::check(uwvkf_pose_set_model_parameters(h_, model_parameters.flat().data()));
  }
  void setParameters(const LocationConfiguration &loc, const PoseUKFParameter &fp) {
    double p[UWVKF_POSE_NPARAMS] = {0};
    p[UWVKF_PP_LAT0] = loc.latitude; p[UWVKF_PP_LON0] = loc.longitude;
    p[UWVKF_PP_GYRO_BIAS_TAU] = fp.gyro_bias_tau; p[UWVKF_PP_ACC_BIAS_TAU] = fp.acc_bias_tau;
    for (int k = 0; k < 3; ++k) {
      p[UWVKF_PP_GYRO_BIAS_OFFSET + k] = fp.gyro_bias_offset[k];
      p[UWVKF_PP_ACC_BIAS_OFFSET + k] = fp.acc_bias_offset[k];
      p[UWVKF_PP_IMU_IN_BODY + k] = fp.imu_in_body[k];
    }
    p[UWVKF_PP_INERTIA_TAU] = fp.inertia_tau; p[UWVKF_PP_LIN_DAMPING_TAU] = fp.lin_damping_tau;
    p[UWVKF_PP_QUAD_DAMPING_TAU] = fp.quad_damping_tau; p[UWVKF_PP_WV_TAU] = fp.water_velocity_tau;
    p[UWVKF_PP_WV_LIMITS] = fp.water_velocity_limits; p[UWVKF_PP_WV_SCALE] = fp.water_velocity_scale;
    p[UWVKF_PP_ADCP_BIAS_TAU] = fp.adcp_bias_tau; p[UWVKF_PP_ATM_PRESSURE] = fp.atmospheric_pressure;
    p[UWVKF_PP_WATER_DENSITY_TAU] = fp.water_density_tau;
    detail::check(uwvkf_pose_set_params(h_, p));
  }
  std::vector<double> tile(const double *v, int n) const {
    std::vector<double> t((size_t)batch_ * n);
    for (int64_t f = 0; f < batch_; ++f) std::memcpy(&t[f * n], v, n * sizeof(double));
    return t;
  }
  void update(int kind, const double *z, const double *R, const double *aux) {
    const int m = (kind == UWVKF_POSE_BODY_EFFORTS || kind == UWVKF_POSE_BODY_EFFORTS_VELOCITY_ONLY) ? 6 :
                  (kind == UWVKF_POSE_Z || kind == UWVKF_POSE_PRESSURE) ? 1 : ((kind == UWVKF_POSE_VELOCITY || kind == UWVKF_POSE_ACCELERATION) ? 3 : 2);
    detail::check(uwvkf_pose_update(h_, kind, tile(z, m).data(), R, 0, aux, nullptr, UWVKF_HOST));
  }
};

// ==============================================================================================================
class VelocityUKF : public FilterBase<4, 4> {
 public:
  UWVKF_MEASUREMENT(DVLMeasurement, 3)
  UWVKF_MEASUREMENT(GyroMeasurement, 3)
  UWVKF_MEASUREMENT(BodyEffortsMeasurement, 6)
  UWVKF_MEASUREMENT(PressureMeasurement, 1)
  VelocityUKF(const State &initial_state, const Covariance &state_cov, int64_t batch = 1, int device = 0)   // src/VelocityUKF.cpp:49-56
      : FilterBase(UWVKF_FILTER_VELOCITY, batch, device) { initializeFilter(initial_state, state_cov); }
  bool setupMotionModel(const UWVParameters &parameters) {                                                    // :58-77
    detail::check(uwvkf_velocity_setup_motion_model(h_, parameters.flat().data()));
    return true;
  }
  void integrateMeasurement(const DVLMeasurement &m) { detail::check(uwvkf_velocity_update(h_, UWVKF_VEL_DVL, tile(m.mu.data(), 3).data(), m.cov.data(), nullptr, UWVKF_HOST)); }
  void integrateMeasurement(const GyroMeasurement &m) { nanCheck(m.cov.v); detail::check(uwvkf_velocity_set_gyro(h_, tile(m.mu.data(), 3).data(), UWVKF_HOST)); }
  void integrateMeasurement(const BodyEffortsMeasurement &m) { nanCheck(m.cov.v); detail::check(uwvkf_velocity_set_efforts(h_, tile(m.mu.data(), 6).data(), UWVKF_HOST)); }
  void integrateMeasurement(const PressureMeasurement &m) { detail::check(uwvkf_velocity_update(h_, UWVKF_VEL_PRESSURE, tile(m.mu.data(), 1).data(), m.cov.data(), nullptr, UWVKF_HOST)); }

 private:
  template <class A> static void nanCheck(const A &a) { for (double c : a) if (std::isnan(c)) throw std::runtime_error("uwvkf: measurement covariance contains NaN"); }
  std::vector<double> tile(const double *v, int n) const {
    std::vector<double> t((size_t)batch_ * n);
    for (int64_t f = 0; f < batch_; ++f) std::memcpy(&t[f * n], v, n * sizeof(double));
    return t;
  }
};

// ==============================================================================================================
class BottomUKF : public FilterBase<4, 3> {
 public:
  UWVKF_MEASUREMENT(RangeMeasurement, 1)
  typedef Vector3d NormalType;   // unit vector on S2
  BottomUKF(const State &initial_state, const Covariance &state_cov, int64_t batch = 1, int device = 0)     // src/BottomUKF.cpp:40-46
      : FilterBase(UWVKF_FILTER_BOTTOM, batch, device) { initializeFilter(initial_state, state_cov); }
  void integrateMeasurement(const RangeMeasurement &m, const Vector3d &unit_direction, const Vector3d &origin) {   // :56-61
    std::vector<double> z((size_t)batch_, m.mu[0]);
    detail::check(uwvkf_bottom_update_range(h_, z.data(), m.cov.data(), unit_direction.data(), origin.data(), nullptr, UWVKF_HOST));
  }
  void integrateMeasurement(const NormalType &measurement, const Matrix2d &measurement_cov) {                       // :63-67
    std::vector<double> z((size_t)batch_ * 3);
    for (int64_t f = 0; f < batch_; ++f) std::memcpy(&z[f * 3], measurement.data(), 3 * sizeof(double));
    detail::check(uwvkf_bottom_update_normal(h_, z.data(), measurement_cov.data(), nullptr, UWVKF_HOST));
  }
  void setVelocity(const Vector3d &velocity) {                                                                     // :69-72
    std::vector<double> v((size_t)batch_ * 3);
    for (int64_t f = 0; f < batch_; ++f) std::memcpy(&v[f * 3], velocity.data(), 3 * sizeof(double));
    detail::check(uwvkf_bottom_set_velocity(h_, v.data(), UWVKF_HOST));
  }
};

// ==============================================================================================================
class IndirectPoseUKF : public FilterBase<7, 6> {
 public:
  UWVKF_MEASUREMENT(VisualFeatureMeasurement, 2)
  IndirectPoseUKF(const Vector3d &position_error_std, const Vector3d &orientation_error_std, double orientation_error_tau,
                  const Vector3d &initial_position_error = Vector3d::Zero(),
                  const Vector3d &initial_position_error_std = Vector3d{1.0, 1.0, 1.0}, int64_t batch = 1, int device = 0)
      : FilterBase(UWVKF_FILTER_INDIRECT_POSE, batch, device) {   // src/IndirectPoseUKF.cpp:53-78
    State s;
    for (int k = 0; k < 3; ++k) s[k] = initial_position_error[k];
    s[6] = 1.0;
    Covariance P, Q;
    for (int k = 0; k < 3; ++k) {
      P(k, k) = initial_position_error_std[k] * initial_position_error_std[k];
      P(3 + k, 3 + k) = orientation_error_std[k] * orientation_error_std[k];
      Q(k, k) = position_error_std[k] * position_error_std[k];
      Q(3 + k, 3 + k) = orientation_error_std[k] * orientation_error_std[k];
    }
    initializeFilter(s, P);
    setProcessNoiseCovariance(Q);
    detail::check(uwvkf_indirect_set_tau(h_, orientation_error_tau));
  }
  void updatePoseReference(const Affine3d &pose_ref) {   // :144-147
    std::vector<double> p((size_t)batch_ * 7);
    for (int64_t f = 0; f < batch_; ++f) pack(pose_ref, &p[f * 7]);
    detail::check(uwvkf_indirect_update_pose_reference(h_, p.data(), UWVKF_HOST));
  }
  void getCorrectedPose(Affine3d &pose_corrected) const {   // :137-142
    std::vector<double> o((size_t)batch_ * 7);
    detail::check(uwvkf_indirect_get_corrected_pose(h_, o.data()));
    pose_corrected.translation = Vector3d(o.data());
    pose_corrected.rotation.x = o[3]; pose_corrected.rotation.y = o[4]; pose_corrected.rotation.z = o[5]; pose_corrected.rotation.w = o[6];
  }
  void integrateMeasurement(const std::vector<VisualFeatureMeasurement> &marker_features, const std::vector<Vector3d> &feature_positions,
                            const Affine3d &marker_pose, const Matrix6d &cov_marker_pose, const CameraConfiguration &camera_config,
                            const Affine3d &camera_in_body) {   // :94-135
    // the reference loops with `||` over both vectors (src/IndirectPoseUKF.cpp:114): unequal sizes read out of bounds there
    if (marker_features.size() != feature_positions.size()) throw std::runtime_error("uwvkf: marker_features and feature_positions must have equal size");
    const int n = (int)marker_features.size();
    std::vector<double> px((size_t)batch_ * n * 2), pc((size_t)n * 4), fp((size_t)n * 3);
    for (int i = 0; i < n; ++i) {
      std::memcpy(&pc[i * 4], marker_features[i].cov.data(), 4 * sizeof(double));
      std::memcpy(&fp[i * 3], feature_positions[i].data(), 3 * sizeof(double));
      for (int64_t f = 0; f < batch_; ++f) std::memcpy(&px[(f * n + i) * 2], marker_features[i].mu.data(), 2 * sizeof(double));
    }
    double mp[7], cb[7], cam[4] = {camera_config.fx, camera_config.fy, camera_config.cx, camera_config.cy};
    pack(marker_pose, mp);
    pack(camera_in_body, cb);
    detail::check(uwvkf_indirect_update_features(h_, n, px.data(), pc.data(), fp.data(), mp, cov_marker_pose.data(), cam, cb, nullptr, UWVKF_HOST));
  }

 private:
  static void pack(const Affine3d &a, double *p) {
    p[0] = a.translation[0]; p[1] = a.translation[1]; p[2] = a.translation[2];
    p[3] = a.rotation.x; p[4] = a.rotation.y; p[5] = a.rotation.z; p[6] = a.rotation.w;
  }
};

}  // namespace uwv_kalman_filters
