/* ORACLE (test infrastructure, NOT product code) -- parity unpinned.
 * Flat C API of the CPU restatement (ctypes-friendly).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library. */
#ifndef UWVKF_ORACLE_API_H
#define UWVKF_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* PoseUKF measurement kinds (reference: src/PoseUKF.hpp:79-88, integrateMeasurement overloads :137-177) */
enum {
  ORC_POSE_XY = 0,             /* XY_Position, m=2                         src/PoseUKF.cpp:506-512 */
  ORC_POSE_Z = 1,              /* Z_Position, m=1                          :498-504 */
  ORC_POSE_PRESSURE = 2,       /* Pressure, m=1, aux = sensor_in_imu[3]    :559-565 */
  ORC_POSE_ACCELERATION = 3,   /* Acceleration, m=3                        :484-490 */
  ORC_POSE_VELOCITY = 4,       /* Velocity (DVL), m=3                      :476-482 */
  ORC_POSE_WATER_VELOCITY = 5, /* ADCP, m=2, aux = cell_weighting, gate    :604-611 */
  ORC_POSE_GEOGRAPHIC = 6,     /* lat/lon, m=2, aux = gps_in_body[3], gate :567-579 */
  ORC_POSE_DELAYED_XY = 7,     /* m=2, aux = delayed_position[2]           :514-527 */
  ORC_POSE_BODY_EFFORTS = 8,   /* m=6, aux = UWV parameters (116 doubles)  :581-602 */
  ORC_POSE_BODY_EFFORTS_VELOCITY_ONLY = 9
};

/* flat PoseUKFParameter (+ projection origin): 22 doubles */
#define ORC_POSE_NPARAMS 22
enum {
  ORC_PP_LAT0 = 0, ORC_PP_LON0 = 1, ORC_PP_GYRO_BIAS_TAU = 2, ORC_PP_GYRO_BIAS_OFFSET = 3, ORC_PP_ACC_BIAS_TAU = 6,
  ORC_PP_ACC_BIAS_OFFSET = 7, ORC_PP_INERTIA_TAU = 10, ORC_PP_LIN_DAMPING_TAU = 11, ORC_PP_QUAD_DAMPING_TAU = 12,
  ORC_PP_WV_TAU = 13, ORC_PP_WV_LIMITS = 14, ORC_PP_WV_SCALE = 15, ORC_PP_ADCP_BIAS_TAU = 16, ORC_PP_ATM_PRESSURE = 17,
  ORC_PP_WATER_DENSITY_TAU = 18, ORC_PP_IMU_IN_BODY = 19
};

/* flat PoseUKFConfig (src/PoseUKFConfig.hpp:159-194): 65 doubles */
#define ORC_CFG_N 65
enum {
  ORC_CFG_ACC_RANDOMWALK = 0, ORC_CFG_ACC_BIAS_OFFSET = 3, ORC_CFG_ACC_BIAS_INSTABILITY = 6, ORC_CFG_ACC_BIAS_TAU = 9,
  ORC_CFG_GYRO_RANDOMWALK = 10, ORC_CFG_GYRO_BIAS_OFFSET = 13, ORC_CFG_GYRO_BIAS_INSTABILITY = 16, ORC_CFG_GYRO_BIAS_TAU = 19,
  ORC_CFG_INERTIA_INSTABILITY = 20, ORC_CFG_LIN_DAMPING_INSTABILITY = 29, ORC_CFG_QUAD_DAMPING_INSTABILITY = 38,
  ORC_CFG_INERTIA_TAU = 47, ORC_CFG_LIN_DAMPING_TAU = 48, ORC_CFG_QUAD_DAMPING_TAU = 49,
  ORC_CFG_WV_TAU = 50, ORC_CFG_WV_LIMITS = 51, ORC_CFG_WV_SCALE = 52, ORC_CFG_ADCP_BIAS_TAU = 53, ORC_CFG_ADCP_BIAS_LIMITS = 54,
  ORC_CFG_LATITUDE = 55, ORC_CFG_LONGITUDE = 56, ORC_CFG_ALTITUDE = 57,
  ORC_CFG_WATER_DENSITY = 58, ORC_CFG_WATER_DENSITY_LIMITS = 59, ORC_CFG_WATER_DENSITY_TAU = 60, ORC_CFG_ATM_PRESSURE = 61,
  ORC_CFG_MAX_JERK = 62
};

int orc_pose_meas_dim(int kind);
/* mu[54] (p, q(x,y,z,w), v, a, bg, ba, g, inertia9, lin9, quad9, wv2, wvb2, badcp2, rho); cov[53*53] row-major.
 * policy = {mean_tol, mean_max_it} or NULL.  Returns the status bit mask (0 = ok). */
int orc_pose_predict(double *mu, double *cov, const double *Q, const double *rot_rate, const double *params,
                     const double *offsets28, double dt, const double *policy, int *mean_it);
int orc_pose_update(int kind, double *mu, double *cov, const double *z, const double *R, const double *aux,
                    const double *rot_rate, const double *params, const double *policy, double *d2_out);
/* visual-marker update on the marker-augmented state (src/PoseUKF.cpp:613-654); poses = t[3], q(x,y,z,w) */
int orc_pose_update_features(double *mu, double *cov, int nfeat, const double *px, const double *pxcov, const double *feature_pos,
                             const double *marker_pose, const double *cov_marker, const double *camera, const double *cam_in_imu,
                             const double *policy);
void orc_pose_rotation_rate(const double *mu, const double *rot_rate, const double *params, double *out);
void orc_pose_init_from_config(const double *pos, const double *pos_cov, const double *quat, const double *rot_cov,
                               const double *cfg, const double *inertia6, const double *lin6, const double *quad6,
                               const double *imu_in_body_rot, double *mu, double *cov);
void orc_pose_process_noise_from_config(const double *cfg, double imu_delta_t, const double *imu_in_body_rot, double *Q);
double orc_pose_step_batch(int B, double *mu, double *cov, const double *Q, const double *rot_rate, const double *z_vel,
                           const double *R_vel, const double *params, const double *offsets, double dt, int nsteps,
                           int nthreads, const double *policy, int *status_out);
int orc_num_threads(void);

/* VelocityUKF (src/VelocityUKF.cpp). mu[4], cov[16]; sim[10] = q(x,y,z,w), linear(3), angular(3) of the private motion
 * model; uwv[116] = inertia 6x6, lin damping 6x6, quad damping 6x6 (row-major), cog[3], cob[3], weight, buoyancy */
#define ORC_UWV_N 116
void orc_vel_setup_motion_model(const double *mu, const double *gyro, double *sim);
int orc_vel_predict(double *mu, double *cov, const double *Q, double *sim, const double *gyro, const double *efforts,
                    const double *uwv, double dt, const double *policy);
int orc_vel_update(int kind /*0 DVL (3), 1 pressure (1)*/, double *mu, double *cov, const double *z, const double *R,
                   const double *policy);
/* BottomUKF (src/BottomUKF.cpp). mu[4] = distance, normal(3); cov[9] */
int orc_bottom_predict(double *mu, double *cov, const double *Q, const double *velocity, double dt, const double *policy);
int orc_bottom_update_range(double *mu, double *cov, const double *z, const double *R, const double *unit_direction,
                            const double *origin, const double *policy);
int orc_bottom_update_normal(double *mu, double *cov, const double *z, const double *R, const double *policy);
/* IndirectPoseUKF (src/IndirectPoseUKF.cpp). mu[7] = position_error, q(x,y,z,w); cov[36]; poses = t[3], q[4] */
int orc_indirect_predict(double *mu, double *cov, const double *Q, double tau, double dt, const double *policy);
int orc_indirect_update_features(double *mu, double *cov, int nfeat, const double *px, const double *pxcov,
                                 const double *feature_pos, const double *marker_pose, const double *cov_marker,
                                 const double *camera, const double *cam_in_body, const double *pose_ref, const double *policy);
void orc_indirect_corrected_pose(const double *mu, const double *pose_ref, double *out);

void orc_tmerc_world_to_nav(double lat0, double lon0, double lat, double lon, double *xy);
void orc_tmerc_nav_to_world(double lat0, double lon0, double x, double y, double *latlon);
double orc_gravity_wgs84(double lat, double alt);

int orc_linear_predict(int n, double *mu, double *cov, const double *F, const double *c, const double *Q);
int orc_linear_update(int n, int m, double *mu, double *cov, const double *H, const double *z, const double *R);
int orc_cholesky_lower(int n, const double *A, double *L);
void orc_so3_boxplus(const double *q, const double *v, double scale, double *out);
void orc_so3_boxminus(const double *a, const double *b, double *out);
void orc_s2_boxplus(const double *x, const double *d, double scale, double *out);
void orc_s2_boxminus(const double *y, const double *x, double *out);
void orc_quat_rot(const double *q, const double *v, double *out);

#ifdef __cplusplus
}
#endif
#endif
