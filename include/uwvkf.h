/*
 * uwvkf -- C ABI of the B200-native batched manifold-UKF engine.
 *
 * Drop-in boundary for the filter classes of tomcreutz/slam-uwv_kalman_filters.  The reference has no FFI
 * layer: its operator API for the hot path is the C++ class interface (predictionStep(dt) /
 * integrateMeasurement(...) / getters) on top of ukfom::ukf<State>::predict(g, Q) / update(z, h, R, gate),
 * which take host callables (reference src/PoseUKF.cpp:462-465, 479-481).  Host callables cannot run on a
 * GPU, so the seam sits at the filter-class level: the process / measurement models are compiled into the
 * kernels and selected by enum.  Each entry point cites the reference interface it replaces.  One handle
 * owns a BATCH of independent filters of one class on one device; batch = 1 is the reference's object.
 *
 * Conventions: plain pointers and sizes only; every call returns UWVKF_OK (0) or a negative error code
 * and never throws; uwvkf_last_error() returns a thread-local message.  Calls are asynchronous on the
 * handle's stream unless they return data to the host.  A handle is not thread-safe; distinct handles
 * are independent.  There is NO CPU fallback: without a CUDA device uwvkf_create fails with
 * UWVKF_ERR_NO_DEVICE.
 *
 * Layouts (host side, row-major, per filter, filters contiguous):
 *   PoseUKF      mu[54]  = p(3) q(x,y,z,w) v(3) a(3) b_gyro(3) b_acc(3) g inertia(9,col-major 3x3)
 *                          lin_damping(9) quad_damping(9) water_velocity(2) water_velocity_below(2)
 *                          bias_adcp(2) water_density            (reference src/PoseState.hpp:29-45)
 *                cov[53*53], tangent order = declaration order, orientation = 3 DOF at 3..5
 *   VelocityUKF  mu[4] = v(3) z ; cov[4*4]                       (src/VelocityUKF.hpp:24-27)
 *   BottomUKF    mu[4] = distance, normal(3, unit) ; cov[3*3]    (src/BottomUKF.hpp:18-21)
 *   IndirectPoseUKF mu[7] = position_error(3) q(x,y,z,w) ; cov[6*6] (src/IndirectPoseUKF.hpp:19-22)
 */
#ifndef UWVKF_H
#define UWVKF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uwvkf_filter_s *uwvkf_handle;

enum uwvkf_filter_kind {
  UWVKF_FILTER_POSE = 0,          /* uwv_kalman_filters::PoseUKF          src/PoseUKF.hpp:40 */
  UWVKF_FILTER_VELOCITY = 1,      /* uwv_kalman_filters::VelocityUKF      src/VelocityUKF.hpp:33 */
  UWVKF_FILTER_BOTTOM = 2,        /* uwv_kalman_filters::BottomUKF        src/BottomUKF.hpp:27 */
  UWVKF_FILTER_INDIRECT_POSE = 3  /* uwv_kalman_filters::IndirectPoseUKF  src/IndirectPoseUKF.hpp:29 */
};

enum uwvkf_error {
  UWVKF_OK = 0,
  UWVKF_ERR_INVALID_ARGUMENT = -1,
  UWVKF_ERR_NO_DEVICE = -2,        /* no CUDA device / driver: the engine has no CPU path */
  UWVKF_ERR_CUDA = -3,
  UWVKF_ERR_NAN_MEASUREMENT = -4,  /* checkMeasurment() of the reference base class would have thrown */
  UWVKF_ERR_NOT_INITIALIZED = -5,  /* e.g. "Motion model is not initialized!" src/VelocityUKF.cpp:117-118 */
  UWVKF_ERR_UNSUPPORTED = -6,
  UWVKF_ERR_BUSY = -7              /* the sample ring is full */
};

enum uwvkf_memspace { UWVKF_HOST = 0, UWVKF_DEVICE = 1 };

/* per-filter asynchronous status bits (uwvkf_get_status) */
enum uwvkf_status_bits {
  UWVKF_ST_NOT_SPD = 1u,            /* Cholesky met a non-positive pivot: op skipped, filter left untouched */
  UWVKF_ST_NAN_MEASUREMENT = 2u,    /* device-side NaN check refused a measurement */
  UWVKF_ST_GATE_REJECTED = 4u,      /* innovation gate (d2p95) refused at least one update */
  UWVKF_ST_MEAN_NOT_CONVERGED = 8u  /* iterated manifold mean hit max_it */
};

/* PoseUKF measurement kinds: one per integrateMeasurement overload (src/PoseUKF.hpp:137-177) */
enum uwvkf_pose_meas {
  UWVKF_POSE_XY = 0,             /* XY_Position m=2                              src/PoseUKF.cpp:506-512 */
  UWVKF_POSE_Z = 1,              /* Z_Position m=1                               src/PoseUKF.cpp:498-504 */
  UWVKF_POSE_PRESSURE = 2,       /* Pressure m=1, aux = pressure_sensor_in_imu[3] src/PoseUKF.cpp:559-565 */
  UWVKF_POSE_ACCELERATION = 3,   /* Acceleration m=3                             src/PoseUKF.cpp:484-490 */
  UWVKF_POSE_VELOCITY = 4,       /* Velocity (DVL) m=3                           src/PoseUKF.cpp:476-482 */
  UWVKF_POSE_WATER_VELOCITY = 5, /* ADCP m=2, aux = cell_weighting, gate d2p95   src/PoseUKF.cpp:604-611 */
  UWVKF_POSE_GEOGRAPHIC = 6,     /* lat/lon [rad] m=2, aux = gps_in_body[3], d2p95 src/PoseUKF.cpp:567-579 */
  UWVKF_POSE_DELAYED_XY = 7,     /* m=2, aux = delayed_position[2]               src/PoseUKF.cpp:514-527 */
  UWVKF_POSE_BODY_EFFORTS = 8,   /* BodyEffortsMeasurement m=6 (forces, torques), measurementEfforts: the Fossen model with the
                                    inertia / damping blocks of the state                src/PoseUKF.cpp:153-196, 581-602 */
  UWVKF_POSE_BODY_EFFORTS_VELOCITY_ONLY = 9 /* same with only_affect_velocity = true: constrainVelocity, :198-219, 585-594 */
};

#define UWVKF_POSE_DIM 54
#define UWVKF_POSE_DOF 53
#define UWVKF_POSE_NPARAMS 22
#define UWVKF_POSE_NOFFSETS 28
#define UWVKF_UWV_NPARAMS 116

/* flat PoseUKF::PoseUKFParameter (src/PoseUKF.hpp:46-76) preceded by the projection origin that the
 * reference keeps in its GeographicProjection member (src/PoseUKF.cpp:356,390) */
enum uwvkf_pose_param_index {
  UWVKF_PP_LAT0 = 0, UWVKF_PP_LON0 = 1, UWVKF_PP_GYRO_BIAS_TAU = 2, UWVKF_PP_GYRO_BIAS_OFFSET = 3,
  UWVKF_PP_ACC_BIAS_TAU = 6, UWVKF_PP_ACC_BIAS_OFFSET = 7, UWVKF_PP_INERTIA_TAU = 10,
  UWVKF_PP_LIN_DAMPING_TAU = 11, UWVKF_PP_QUAD_DAMPING_TAU = 12, UWVKF_PP_WV_TAU = 13, UWVKF_PP_WV_LIMITS = 14,
  UWVKF_PP_WV_SCALE = 15, UWVKF_PP_ADCP_BIAS_TAU = 16, UWVKF_PP_ATM_PRESSURE = 17,
  UWVKF_PP_WATER_DENSITY_TAU = 18, UWVKF_PP_IMU_IN_BODY = 19
};

/* ---- lifetime ------------------------------------------------------------------------------------- */
/* Replaces the constructors (src/PoseUKF.hpp:100-115, src/VelocityUKF.hpp:42, src/BottomUKF.hpp:33,
 * src/IndirectPoseUKF.hpp:44-48): allocates `batch` filters of `filter_kind` on CUDA device `device`. */
int uwvkf_create(int filter_kind, int64_t batch, int device, uwvkf_handle *out);
int uwvkf_destroy(uwvkf_handle h);
/* All work of the handle is ordered on this cudaStream_t (default: the legacy default stream).  The old stream is drained
 * before the switch (staged inputs may still be in flight on it); uwvkf_create returns with all device memory initialised,
 * so a non-blocking user stream may be used right away. */
int uwvkf_set_stream(uwvkf_handle h, void *cuda_stream);
int uwvkf_synchronize(uwvkf_handle h);
const char *uwvkf_last_error(void);
const char *uwvkf_version(void);
int64_t uwvkf_batch(uwvkf_handle h);
/* kernels launched by this handle so far (bench.py's gpu_launches) */
int64_t uwvkf_kernel_launches(uwvkf_handle h);

/* ---- state ---------------------------------------------------------------------------------------- */
/* initializeFilter(state, cov) of the base class (call sites src/PoseUKF.cpp:344,378): uploads
 * mu[count][DIM] and the full covariances for filters [first, first+count): cov_stride = DOF*DOF for
 * cov[count][DOF*DOF], or 0 when one covariance cov[DOF*DOF] is shared by all `count` filters.
 * For PoseUKF it also latches the Markov offsets from the uploaded state (src/PoseUKF.cpp:346-349,385-388). */
int uwvkf_set_state(uwvkf_handle h, const double *mu, const double *cov, int64_t cov_stride, int64_t first,
                    int64_t count, int memspace);
/* getCurrentState / covariance getter of the base class (used at src/VelocityUKF.cpp:70). cov may be NULL. */
int uwvkf_get_state(uwvkf_handle h, double *mu, double *cov, int64_t first, int64_t count, int memspace);
/* setProcessNoiseCovariance (src/PoseUKF.cpp:438): Q[DOF*DOF], shared by the batch. */
int uwvkf_set_process_noise(uwvkf_handle h, const double *Q);
/* ukfom mean-iteration policy: loop while |mean_delta| > tol and fewer than max_it passes (default 1e-9, 100) */
int uwvkf_set_mean_policy(uwvkf_handle h, double tol, int max_it);
int uwvkf_get_status(uwvkf_handle h, uint32_t *status, int64_t first, int64_t count);
/* metrics: counts[4] = number of filters whose status has UWVKF_ST_NOT_SPD / NAN_MEASUREMENT / GATE_REJECTED /
 * MEAN_NOT_CONVERGED set (the reference only writes such events to stderr) */
int uwvkf_status_counts(uwvkf_handle h, int64_t *counts);
int uwvkf_clear_status(uwvkf_handle h);

/* ---- PoseUKF -------------------------------------------------------------------------------------- */
/* PoseUKFParameter + projection origin (src/PoseUKF.cpp:358-371, 356): params[UWVKF_POSE_NPARAMS], shared. */
int uwvkf_pose_set_params(uwvkf_handle h, const double *params);
/* model_parameters of the constructors (uwv_dynamic_model::UWVParameters -> dynamic_model, src/PoseUKF.cpp:353-354,382-383):
 * flat UWVKF_UWV_NPARAMS doubles = inertia 6x6, linear damping 6x6, quadratic damping 6x6 (row-major), cog[3], cob[3], weight,
 * buoyancy; shared by the batch.  Needed by the UWVKF_POSE_BODY_EFFORTS* updates only (else UWVKF_ERR_NOT_INITIALIZED). */
int uwvkf_pose_set_model_parameters(uwvkf_handle h, const double *uwv_params);
/* Markov offsets (inertia 9, lin_damping 9, quad_damping 9, water_density) per filter, host pointer;
 * overrides what uwvkf_set_state latched (src/PoseUKF.cpp:346-349). */
int uwvkf_pose_set_offsets(uwvkf_handle h, const double *offsets, int64_t first, int64_t count);
/* integrateMeasurement(const RotationRate&) (src/PoseUKF.cpp:492-496): stores the control input w[B][3].  NaN in a HOST
 * w returns UWVKF_ERR_NAN_MEASUREMENT (any batch size; the reference throws); a NaN in a DEVICE w is caught per filter by
 * the next prediction, which is skipped for that filter with UWVKF_ST_NAN_MEASUREMENT set. */
int uwvkf_pose_set_rotation_rate(uwvkf_handle h, const double *w, int memspace);
/* getRotationRate() (src/PoseUKF.cpp:693-699) for every filter: out[B][3], host. */
int uwvkf_pose_get_rotation_rate(uwvkf_handle h, double *out);
/* predictionStep(dt) -> PoseUKF::predictionStepImpl (src/PoseUKF.cpp:446-466) for the whole batch. */
int uwvkf_predict(uwvkf_handle h, double dt);
/* integrateMeasurement(<kind>) (src/PoseUKF.cpp:476-611).  z[B][m]; R[m*m] shared or R[B][m*m] when
 * R_per_filter != 0; aux: per-kind extra arguments shared by the batch (host pointer, may be NULL ->
 * zeros); valid[B] (may be NULL): filters with valid == 0 skip the update (sensor dropout).
 * memspace says where z, R (if per filter) and valid live.  NaN in a HOST z/R returns
 * UWVKF_ERR_NAN_MEASUREMENT like checkMeasurment() (every HOST input is scanned, whatever the batch size);
 * DEVICE inputs are checked per filter in the kernel (UWVKF_ST_NAN_MEASUREMENT, update skipped). */
int uwvkf_pose_update(uwvkf_handle h, int meas_kind, const double *z, const double *R, int R_per_filter,
                      const double *aux, const uint8_t *valid, int memspace);
/* Fused hot path: predictionStep(dt) immediately followed by integrateMeasurement(<kind>) in ONE kernel
 * (one HBM round trip of the filter state).  w[B][3] is the rotation-rate control for this step (NULL:
 * keep the stored one).  Results are identical to uwvkf_pose_set_rotation_rate + uwvkf_predict +
 * uwvkf_pose_update. */
int uwvkf_pose_predict_update(uwvkf_handle h, double dt, const double *w, int meas_kind, const double *z,
                              const double *R, int R_per_filter, const double *aux, const uint8_t *valid,
                              int memspace);
/* The 100 Hz step of a full PoseUKF in ONE call: predictionStep(dt) (dt == 0: none) followed by integrateMeasurement(kind_a)
 * and integrateMeasurement(kind_b) (src/PoseUKF.cpp:446-466, 476-490) -- e.g. DVL + accelerometer, BASELINE config 3.  It
 * launches the fused predict + kind_a kernel and then the kind_b kernel.  (A single two-update kernel -- the panel of the
 * covariance on chip between the three operations, the rest read and written once with both downdates -- exists for (VELOCITY,
 * ACCELERATION) behind the build switch UWVKF_FUSED2: bit-identical, but slower than the two launches on B200, see
 * DESIGN.md section 6b.)  w[B][3] (may be NULL: keep the stored rotation rate),
 * z_a / z_b [B][m], R_a / R_b [m*m] shared by the batch, valid_a / valid_b [B] optional dropout masks.  Bit-identical to
 * uwvkf_pose_set_rotation_rate + uwvkf_predict + uwvkf_pose_update(kind_a) + uwvkf_pose_update(kind_b). */
int uwvkf_pose_predict_update2(uwvkf_handle h, double dt, const double *w, int kind_a, const double *z_a, const double *R_a,
                               const uint8_t *valid_a, int kind_b, const double *z_b, const double *R_b, const uint8_t *valid_b,
                               int memspace);
/* integrateMeasurement(marker_features, feature_positions, marker_pose, cov_marker_pose, camera_config, camera_in_IMU)
 * (src/PoseUKF.cpp:613-654): the state is augmented with the marker pose (59 DOF, 119 sigma points), one S2 update per
 * feature, the leading 53 x 53 block is kept.  px[B][nfeat][2] pixel coordinates (memspace), pxcov[nfeat][4],
 * feature_pos[nfeat][3] (marker frame), marker_pose[7] = t(3) q(x,y,z,w) in the navigation frame, cov_marker[36],
 * camera = fx, fy, cx, cy, cam_in_imu[7]; 1 <= nfeat <= 8; valid[B] may be NULL. */
int uwvkf_pose_update_features(uwvkf_handle h, int nfeat, const double *px, const double *pxcov, const double *feature_pos,
                               const double *marker_pose, const double *cov_marker, const double *camera, const double *cam_in_imu,
                               const uint8_t *valid, int memspace);
/* Navigation solution of filters [first, first+count): out[count][10] = p(3) q(x,y,z,w) v(3) -- the leading
 * part of getCurrentState() that a caller reads every step.  Asynchronous for memspace == UWVKF_DEVICE. */
int uwvkf_pose_get_navigation(uwvkf_handle h, double *out, int64_t first, int64_t count, int memspace);
/* Pipelined HOST-buffer form of the fused hot path: the call a sensor-driver loop makes once per IMU sample.  One call =
 * integrateMeasurement(RotationRate) + predictionStep(dt) + integrateMeasurement(<kind>) for the whole batch
 * (src/PoseUKF.cpp:492-496, 446-466, 476-611) followed by a read-back of the navigation solution:
 *   w_host[B][3], z_host[B][m]  this step's samples in HOST memory (pinned memory makes the copies asynchronous),
 *   R[m*m], aux                 shared by the batch, as in uwvkf_pose_update,
 *   nav_out_host[B][10]         (may be NULL) receives p(3) q(x,y,z,w) v(3) of every filter AFTER this step.
 * The call only enqueues work and returns: uploads run on an internal copy stream into one of two device slots, the fused
 * kernel on the handle's stream, the download on a second copy stream -- so the upload of step i+1 and the download of step
 * i-1 overlap the kernel of step i.  Buffer rules: w_host / z_host of call i may be overwritten once call i+1 has returned;
 * nav_out_host of call i is complete once call i+2 has returned, or after uwvkf_synchronize.  The host never runs more than two
 * steps ahead of the device.  NaN samples are refused per filter on the device (UWVKF_ST_NAN_MEASUREMENT): a host scan would
 * serialise the pipeline. */
int uwvkf_pose_step_host(uwvkf_handle h, double dt, const double *w_host, int meas_kind, const double *z_host, const double *R,
                         const double *aux, double *nav_out_host);
/* Orders everything the internal copy streams of uwvkf_pose_step_host still have in flight before whatever is enqueued next on
 * the handle's stream (so that an event recorded there covers the downloads as well).  uwvkf_synchronize includes it. */
int uwvkf_join(uwvkf_handle h);
/* resetFilterWithExternalPose (src/PoseUKF.cpp:685-691): overwrite p and q of every filter, keep Sigma.
 * pose[B][7] = p(3) q(x,y,z,w), host. */
int uwvkf_pose_reset_with_external_pose(uwvkf_handle h, const double *pose);

/* ---- VelocityUKF (src/VelocityUKF.hpp:33-66) ------------------------------------------------------------ */
enum uwvkf_velocity_meas {
  UWVKF_VEL_DVL = 0,      /* DVLMeasurement m=3       src/VelocityUKF.cpp:79-85 */
  UWVKF_VEL_PRESSURE = 1  /* PressureMeasurement m=1  src/VelocityUKF.cpp:106-112 */
};
/* flat uwv_dynamic_model::UWVParameters: inertia 6x6, linear damping 6x6, quadratic damping 6x6 (row-major),
 * distance_body2centerofgravity[3], distance_body2centerofbuoyancy[3], weight, buoyancy */
/* setupMotionModel(parameters) (src/VelocityUKF.cpp:58-77); uwvkf_predict fails with UWVKF_ERR_NOT_INITIALIZED
 * ("Motion model is not initialized!", :117-118) until this was called. */
int uwvkf_velocity_setup_motion_model(uwvkf_handle h, const double *uwv_params);
/* integrateMeasurement(GyroMeasurement) (:87-98): w[B][3] */
int uwvkf_velocity_set_gyro(uwvkf_handle h, const double *w, int memspace);
/* integrateMeasurement(BodyEffortsMeasurement) (:100-104): tau[B][6] */
int uwvkf_velocity_set_efforts(uwvkf_handle h, const double *tau, int memspace);
/* integrateMeasurement(DVLMeasurement / PressureMeasurement): z[B][m], R[m*m] shared */
int uwvkf_velocity_update(uwvkf_handle h, int meas_kind, const double *z, const double *R, const uint8_t *valid, int memspace);
/* state of the private motion model (src/VelocityUKF.hpp:62): sim[B][10] = q(x,y,z,w), linear(3), angular(3) velocity */
int uwvkf_velocity_get_motion_model_state(uwvkf_handle h, double *sim);

/* ---- BottomUKF (src/BottomUKF.hpp:27-52) ---------------------------------------------------------------- */
/* setVelocity (src/BottomUKF.cpp:69-72): v[B][3] */
int uwvkf_bottom_set_velocity(uwvkf_handle h, const double *v, int memspace);
/* integrateMeasurement(RangeMeasurement, unit_direction, origin) (:56-61): z[B], R[1], beam geometry shared */
int uwvkf_bottom_update_range(uwvkf_handle h, const double *z, const double *R, const double *unit_direction,
                              const double *origin, const uint8_t *valid, int memspace);
/* integrateMeasurement(NormalType, measurement_cov) (:63-67): z[B][3] unit vectors, R[2*2] */
int uwvkf_bottom_update_normal(uwvkf_handle h, const double *z, const double *R, const uint8_t *valid, int memspace);

/* ---- IndirectPoseUKF (src/IndirectPoseUKF.hpp:29-85) ---------------------------------------------------- */
/* orientation_error_tau of the constructor (src/IndirectPoseUKF.cpp:53-58); the initial state / covariance /
 * process noise the constructor derives (:60-77) are set with uwvkf_set_state / uwvkf_set_process_noise */
int uwvkf_indirect_set_tau(uwvkf_handle h, double orientation_error_tau);
/* updatePoseReference (:144-147): pose_ref[B][7] = t(3) q(x,y,z,w), body in world */
int uwvkf_indirect_update_pose_reference(uwvkf_handle h, const double *pose_ref, int memspace);
/* integrateMeasurement(marker_features, feature_positions, marker_pose, cov_marker_pose, camera_config,
 * camera_in_body) (:94-135): px[B][nfeat][2] pixel coordinates, pxcov[nfeat][4], feature_pos[nfeat][3],
 * marker_pose[7], cov_marker[36], camera = fx, fy, cx, cy, cam_in_body[7]; 1 <= nfeat <= 8 */
int uwvkf_indirect_update_features(uwvkf_handle h, int nfeat, const double *px, const double *pxcov,
                                   const double *feature_pos, const double *marker_pose, const double *cov_marker,
                                   const double *camera, const double *cam_in_body, const uint8_t *valid, int memspace);
/* getCorrectedPose (:137-142): out[B][7], host */
int uwvkf_indirect_get_corrected_pose(uwvkf_handle h, double *out);

/* ---- PoseUKF sample stream: the caller-side loop ------------------------------------------------- */
/* The reference is driven by an application that, for every arriving sensor sample, calls predictionStepFromSampleTime(ts) and
 * integrateMeasurement(sample) (Rock task, not in the repository).  A uwvkf_stream is that loop for a batch: timestamped sample
 * sets are pushed from host memory into a ring of pinned staging buffers + HBM slots (uploaded asynchronously on a copy stream)
 * and uwvkf_stream_run replays them in time order on the handle's stream -- predictionStep(ts - filter time) fused with the
 * update into one kernel -- with event dependencies only, so uploads of later samples overlap the kernels of earlier ones and
 * the host never blocks on the GPU.  kind = a uwvkf_pose_meas vector kind (z[B][m], R[m*m] shared, aux, valid[B] or NULL), or
 * UWVKF_STREAM_ROTATION_RATE (z[B][3]: the control input of the following predictions, no prediction by itself).  The first
 * measurement only latches the filter time.  Samples must be pushed in non-decreasing time order. */
typedef struct uwvkf_stream_s *uwvkf_stream;
#define UWVKF_STREAM_ROTATION_RATE 100
int uwvkf_stream_create(uwvkf_handle h, int slots, uwvkf_stream *out);
int uwvkf_stream_destroy(uwvkf_stream s);
int uwvkf_stream_push(uwvkf_stream s, double timestamp, int kind, const double *z, const double *R, const double *aux,
                      const uint8_t *valid);
/* replays every queued sample set with timestamp <= until; asynchronous (uwvkf_synchronize(h) to wait) */
int uwvkf_stream_run(uwvkf_stream s, double until);
int uwvkf_stream_pending(uwvkf_stream s);
double uwvkf_stream_time(uwvkf_stream s);

/* ---- CUDA-graph replay of a fixed schedule ------------------------------------------------------------- */
/* Every step / update entry point called with memspace = UWVKF_DEVICE only enqueues work on the handle's stream, so a fixed
 * schedule (e.g. K steps of rotation rate + fused predict/DVL + acceleration reading consecutive rows of device-resident input
 * arrays) can be captured once and replayed with one launch: uwvkf_graph_begin(h); <calls>; uwvkf_graph_end(h, &g);
 * uwvkf_graph_launch(g) as often as needed (the inputs are read from the same device addresses at every replay).  The handle
 * needs a user stream (uwvkf_set_stream) and every kernel of the schedule must have been launched once before the capture.
 * Calls that return data to the host or take host pointers cannot be captured. */
typedef struct uwvkf_graph_s *uwvkf_graph;
int uwvkf_graph_begin(uwvkf_handle h);
int uwvkf_graph_end(uwvkf_handle h, uwvkf_graph *out);
int uwvkf_graph_launch(uwvkf_graph g);
int uwvkf_graph_destroy(uwvkf_graph g);

/* ---- device-math probe (tests only) ------------------------------------------------------------------- */
/* Evaluates one device function of the engine per input row on the GPU so that tests can pin the kernels' math against
 * the oracle / high-precision values: in[n][8] -> out[n][4] (host pointers, PoseUKF handle with parameters set).
 *   TMERC_SINCOS_LAT  in = x, y [m, NWU]           out = sin(lat), cos(lat) (fast path), lat (generic path)
 *   SO3_BOXPLUS       in = q(x,y,z,w), v(3), scale out = q (+) scale*v
 *   SO3_BOXMINUS      in = a(x,y,z,w), b(x,y,z,w)  out = a (-) b
 *   RSQRT             in = d                       out = 1/sqrt(d) (third-order), 1/sqrt(d) (two Newton steps), 1/d
 *   WARP_SUM          in = v                       out = sum of v over the 32 rows of the same warp (DMMA path, shuffle path) */
enum uwvkf_debug_what {
  UWVKF_DEBUG_TMERC_SINCOS_LAT = 0, UWVKF_DEBUG_SO3_BOXPLUS = 1, UWVKF_DEBUG_SO3_BOXMINUS = 2, UWVKF_DEBUG_RSQRT = 3,
  UWVKF_DEBUG_WARP_SUM = 4
};
int uwvkf_debug_eval(uwvkf_handle h, int what, const double *in, double *out, int n);
/* Per-phase latency table of the fused PoseUKF kernel (40 counters: clock cycles summed over the filters of warp 0 of every
 * CTA, [39] = number of such filters).  Only a library built with -DUWVKF_PHASE_TIMING (tools/phase_timing.py) fills it; the
 * product build returns UWVKF_ERR_INVALID_ARGUMENT.  No reference counterpart: profiling aid. */
int uwvkf_debug_phase_cycles(uwvkf_handle h, uint64_t *out, int reset);

/* ---- ensemble statistics (north_star: the only cross-GPU exchange) -------------------------------- */
/* Writes [count, sum mu (DIM), sum mu mu^T lower triangle row-major (DIM*(DIM+1)/2)] over this handle's filters into
 * out (device or host pointer, 1 + DIM + DIM*(DIM+1)/2 doubles).  The caller all-reduces it (NCCL sum). */
int uwvkf_ensemble_stats(uwvkf_handle h, double *out, int memspace);
int64_t uwvkf_ensemble_stats_size(uwvkf_handle h);

#ifdef __cplusplus
}
#endif
#endif /* UWVKF_H */
