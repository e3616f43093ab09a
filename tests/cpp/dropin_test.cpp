// Drives the C++ host classes (include/uwv_kalman_filters/uwv_kalman_filters.hpp) the way an application drives
// the reference classes: construct from the configuration, setProcessNoiseFromConfig, then per IMU sample
// integrateMeasurement(RotationRate) -> predictionStep(dt) -> integrateMeasurement(Velocity / Acceleration).
// Inputs come from stdin (written by tests/test_cpp_shim_gpu.py), results go to stdout with 17 digits.
#include <cstdio>
#include <iostream>
#include <vector>

#include "../../include/uwv_kalman_filters/uwv_kalman_filters.hpp"

using namespace uwv_kalman_filters;

static PoseUKFConfig synth_config() {   // SYNTH-1 (slam-uwv_kalman_filters_b200/synthetic.py::pose_config)
  PoseUKFConfig c{};
  for (int k = 0; k < 3; ++k) {
    c.acceleration.randomwalk[k] = 1e-3; c.acceleration.bias_instability[k] = 1e-4;
    c.rotation_rate.randomwalk[k] = 1e-5; c.rotation_rate.bias_instability[k] = 1e-6;
    c.max_jerk[k] = 4.0;
  }
  c.acceleration.bias_tau = c.rotation_rate.bias_tau = 3600.0;
  for (int k = 0; k < 9; ++k)
    c.model_noise_parameters.inertia_instability[k] = c.model_noise_parameters.lin_damping_instability[k] =
        c.model_noise_parameters.quad_damping_instability[k] = 5.0;
  c.model_noise_parameters.inertia_tau = c.model_noise_parameters.lin_damping_tau = c.model_noise_parameters.quad_damping_tau = 1e4;
  c.water_velocity.tau = 1800.0; c.water_velocity.limits = 0.3; c.water_velocity.scale = 1e-4;
  c.water_velocity.adcp_bias_tau = 3600.0; c.water_velocity.adcp_bias_limits = 0.05;
  c.location.latitude = 0.92677; c.location.longitude = 0.15446; c.location.altitude = 0.0;
  c.hydrostatics.water_density = 1025.0; c.hydrostatics.water_density_limits = 5.0; c.hydrostatics.water_density_tau = 3600.0;
  c.hydrostatics.atmospheric_pressure = 101325.0;
  return c;
}
static UWVParameters synth_model() {
  UWVParameters p;
  const double M[6] = {200, 250, 300, 20, 40, 40}, Dl[6] = {30, 60, 80, 5, 10, 10}, Dq[6] = {60, 120, 160, 10, 20, 20};
  for (int k = 0; k < 6; ++k) { p.inertia_matrix(k, k) = M[k]; p.damping_matrices[0](k, k) = Dl[k]; p.damping_matrices[1](k, k) = Dq[k]; }
  p.weight = p.buoyancy = 1500.0;
  return p;
}
template <class V> static void read(V &v, int n) { for (int k = 0; k < n; ++k) std::cin >> v[k]; }
template <class S, class C> static void dump(const char *tag, const S &s, int ns, const C &c, int nc) {
  std::printf("%s", tag);
  for (int k = 0; k < ns; ++k) std::printf(" %.17g", s[k]);
  for (int k = 0; k < nc; ++k) std::printf(" %.17g", c[k]);
  std::printf("\n");
}

int main() {
  int steps;
  std::cin >> steps;
  // ---- PoseUKF ------------------------------------------------------------------------------------------------
  PoseUKFConfig cfg = synth_config();
  Matrix3d rot_cov;
  rot_cov(0, 0) = 1e-4; rot_cov(1, 1) = 1e-4; rot_cov(2, 2) = 1e-2;
  PoseUKF pose(Vector3d::Zero(), Matrix3d::Identity(), Quaterniond::Identity(), rot_cov, cfg, synth_model());
  pose.setProcessNoiseFromConfig(cfg, 0.01);
  // the same filter twice more, driven through the batched entry points a sensor loop would use: predictionStepAndIntegrate2
  // (one call per IMU sample: prediction + DVL + accelerometer) and stepHost (pipelined host-buffer step)
  PoseUKF pose2(Vector3d::Zero(), Matrix3d::Identity(), Quaterniond::Identity(), rot_cov, cfg, synth_model());
  PoseUKF pose3(Vector3d::Zero(), Matrix3d::Identity(), Quaterniond::Identity(), rot_cov, cfg, synth_model());
  pose2.setProcessNoiseFromConfig(cfg, 0.01);
  pose3.setProcessNoiseFromConfig(cfg, 0.01);
  std::vector<std::vector<double>> keep_w(steps, std::vector<double>(3)), keep_v(steps, std::vector<double>(3));
  std::vector<double> nav(10 * 3);
  for (int s = 0; s < steps; ++s) {
    PoseUKF::RotationRate w; PoseUKF::Velocity v; PoseUKF::Acceleration a; PoseUKF::Pressure p;
    read(w.mu, 3); read(v.mu, 3); read(a.mu, 3); read(p.mu, 1);
    for (int k = 0; k < 3; ++k) { w.cov(k, k) = 1e-10; v.cov(k, k) = 1e-4; a.cov(k, k) = 1e-4; }
    p.cov(0, 0) = 1e4;
    pose.integrateMeasurement(w);
    pose.predictionStep(0.01);
    pose.integrateMeasurement(v);
    pose.integrateMeasurement(a);
    if ((s + 1) % 5 == 0) pose.integrateMeasurement(p, Vector3d{0.3, -0.1, 0.2});
    const double Rv[9] = {1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-4};
    for (int k = 0; k < 3; ++k) { keep_w[s][k] = w.mu[k]; keep_v[s][k] = v.mu[k]; }     // host buffers stay alive (stepHost is asynchronous)
    pose2.predictionStepAndIntegrate2(0.01, keep_w[s].data(), UWVKF_POSE_VELOCITY, keep_v[s].data(), Rv, UWVKF_POSE_ACCELERATION,
                                      a.mu.data(), Rv);
    if ((s + 1) % 5 == 0) pose2.integrateMeasurement(p, Vector3d{0.3, -0.1, 0.2});
    pose3.stepHost(0.01, keep_w[s].data(), UWVKF_POSE_VELOCITY, keep_v[s].data(), Rv, nav.data() + 10 * (s % 3));
    pose3.integrateMeasurement(a);
    if ((s + 1) % 5 == 0) pose3.integrateMeasurement(p, Vector3d{0.3, -0.1, 0.2});
  }
  pose3.synchronize();
  {
    PoseUKF::State s2, s3; PoseUKF::Covariance c2, c3;
    pose2.getCurrentStateAndCovariance(s2, c2);
    pose3.getCurrentStateAndCovariance(s3, c3);
    dump("POSE2", s2, 54, c2.v, 53 * 53);
    dump("POSE3", s3, 54, c3.v, 53 * 53);
  }
  {   // model-aided part: forces / torques in the body frame, both flavours (reference src/PoseUKF.cpp:581-602)
    PoseUKF::BodyEffortsMeasurement e;
    const double tau[6] = {12.0, -3.0, 1.5, 0.0, 0.0, 0.4};
    for (int k = 0; k < 6; ++k) { e.mu[k] = tau[k]; e.cov(k, k) = k < 3 ? 4.0 : 0.25; }
    pose.integrateMeasurement(e);
    pose.integrateMeasurement(e, true);
  }
  PoseUKF::State st; PoseUKF::Covariance cv;
  pose.getCurrentStateAndCovariance(st, cv);
  dump("POSE", st, 54, cv.v, 53 * 53);
  Vector3d rr = pose.getRotationRate();
  dump("ROTRATE", rr, 3, rr, 0);
  // NaN measurement -> std::runtime_error like checkMeasurment()
  bool threw = false;
  try { PoseUKF::Velocity bad; bad.mu[1] = NAN; bad.cov = Matrix3d::Identity(); pose.integrateMeasurement(bad); } catch (const std::runtime_error &) { threw = true; }
  std::printf("NAN_THROWS %d\n", (int)threw);
  // ---- VelocityUKF --------------------------------------------------------------------------------------------
  VelocityUKF::State vs{1.0, 0.0, 0.0, -10.0};
  VelocityUKF::Covariance vc;
  vc(0, 0) = vc(1, 1) = vc(2, 2) = 0.1; vc(3, 3) = 1.0;
  VelocityUKF vel(vs, vc);
  threw = false;
  try { vel.predictionStep(0.01); } catch (const std::runtime_error &e) { threw = std::string(e.what()).find("Motion model is not initialized!") != std::string::npos; }
  std::printf("MODEL_THROWS %d\n", (int)threw);
  vel.setupMotionModel(synth_model());
  for (int s = 0; s < steps; ++s) {
    VelocityUKF::GyroMeasurement g; VelocityUKF::BodyEffortsMeasurement e; VelocityUKF::DVLMeasurement d;
    read(g.mu, 3); read(e.mu, 6); read(d.mu, 3);
    for (int k = 0; k < 3; ++k) d.cov(k, k) = 1e-4;
    vel.integrateMeasurement(g);
    vel.integrateMeasurement(e);
    vel.predictionStep(0.01);
    if ((s + 1) % 5 == 0) vel.integrateMeasurement(d);
  }
  VelocityUKF::State vst; VelocityUKF::Covariance vcv;
  vel.getCurrentStateAndCovariance(vst, vcv);
  dump("VEL", vst, 4, vcv.v, 16);
  // ---- BottomUKF ----------------------------------------------------------------------------------------------
  BottomUKF::State bs{5.0, 0.0, 0.0, 1.0};
  BottomUKF::Covariance bc;
  bc(0, 0) = 1.0; bc(1, 1) = bc(2, 2) = 0.05;
  BottomUKF bottom(bs, bc);
  bottom.setVelocity(Vector3d{1.0, 0.2, -0.05});
  bottom.predictionStep(0.1);
  BottomUKF::RangeMeasurement r;
  r.mu[0] = 6.4; r.cov(0, 0) = 4e-4;
  bottom.integrateMeasurement(r, Vector3d{0.35355339059327373, 0.35355339059327373, -0.8660254037844387}, Vector3d{0.1, 0.0, -0.2});
  BottomUKF::State bst; BottomUKF::Covariance bcv;
  bottom.getCurrentStateAndCovariance(bst, bcv);
  dump("BOTTOM", bst, 4, bcv.v, 9);
  // ---- IndirectPoseUKF ----------------------------------------------------------------------------------------
  IndirectPoseUKF ind(Vector3d{0.01, 0.01, 0.01}, Vector3d{0.002, 0.002, 0.005}, 100.0, Vector3d::Zero(), Vector3d{0.5, 0.5, 0.5});
  Affine3d ref;
  ref.translation = Vector3d{0.2, 0.05, 0.0};
  ind.updatePoseReference(ref);
  ind.predictionStep(0.01);
  Affine3d corrected;
  ind.getCorrectedPose(corrected);
  IndirectPoseUKF::State ist; IndirectPoseUKF::Covariance icv;
  ind.getCurrentStateAndCovariance(ist, icv);
  dump("INDIRECT", ist, 7, icv.v, 36);
  dump("CORRECTED", corrected.translation, 3, corrected.translation, 0);
  return 0;
}
