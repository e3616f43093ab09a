"""ctypes binding of libuwvkf.so (include/uwvkf.h) and a batched mirror of the reference's PoseUKF class API."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST, DEVICE = 0, 1
FILTER_POSE, FILTER_VELOCITY, FILTER_BOTTOM, FILTER_INDIRECT_POSE = 0, 1, 2, 3
(POSE_XY, POSE_Z, POSE_PRESSURE, POSE_ACCELERATION, POSE_VELOCITY, POSE_WATER_VELOCITY, POSE_GEOGRAPHIC,
 POSE_DELAYED_XY, POSE_BODY_EFFORTS, POSE_BODY_EFFORTS_VELOCITY_ONLY) = range(10)
POSE_MEAS_DIM = {POSE_XY: 2, POSE_Z: 1, POSE_PRESSURE: 1, POSE_ACCELERATION: 3, POSE_VELOCITY: 3, POSE_WATER_VELOCITY: 2,
                 POSE_GEOGRAPHIC: 2, POSE_DELAYED_XY: 2, POSE_BODY_EFFORTS: 6, POSE_BODY_EFFORTS_VELOCITY_ONLY: 6}
ST_NOT_SPD, ST_NAN_MEASUREMENT, ST_GATE_REJECTED, ST_MEAN_NOT_CONVERGED = 1, 2, 4, 8
ERR_NAN_MEASUREMENT = -4

_dp = C.POINTER(C.c_double)


class UwvkfError(RuntimeError):
    """Raised for any non-zero return code; the reference throws std::runtime_error in the same places."""

    def __init__(self, code, msg):
        super().__init__(f"uwvkf error {code}: {msg}")
        self.code = code


def library_path():
    """In-tree CUDA library; UWVKF_LIBRARY selects another build of the same sources (kernel tuning experiments)."""
    return os.environ.get("UWVKF_LIBRARY") or os.path.join(_HERE, "libuwvkf.so")


_lib = None


def load_library():
    """Loads the in-tree CUDA library; there is deliberately no fallback when it is missing."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise UwvkfError(-2, f"{path} not built: run __graft_entry__.build() (nvcc, sm_100a); there is no CPU path")
        lib = C.CDLL(path)
        lib.uwvkf_last_error.restype = C.c_char_p
        lib.uwvkf_version.restype = C.c_char_p
        lib.uwvkf_batch.restype = C.c_int64
        lib.uwvkf_kernel_launches.restype = C.c_int64
        lib.uwvkf_ensemble_stats_size.restype = C.c_int64
        for name in ("uwvkf_batch", "uwvkf_kernel_launches", "uwvkf_ensemble_stats_size", "uwvkf_destroy", "uwvkf_synchronize",
                     "uwvkf_clear_status"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.uwvkf_create.argtypes = [C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
        lib.uwvkf_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]
        lib.uwvkf_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        lib.uwvkf_set_process_noise.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_set_mean_policy.argtypes = [C.c_void_p, C.c_double, C.c_int]
        lib.uwvkf_get_status.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        lib.uwvkf_status_counts.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_pose_set_params.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_pose_set_model_parameters.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_pose_set_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        lib.uwvkf_pose_set_rotation_rate.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_pose_get_rotation_rate.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_predict.argtypes = [C.c_void_p, C.c_double]
        lib.uwvkf_pose_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_pose_predict_update.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_pose_predict_update2.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_pose_update_features.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8 + [C.c_int]
        lib.uwvkf_pose_get_navigation.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        lib.uwvkf_pose_reset_with_external_pose.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_pose_step_host.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.uwvkf_join.argtypes = [C.c_void_p]
        lib.uwvkf_ensemble_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_stream_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.uwvkf_stream_destroy.argtypes = [C.c_void_p]
        lib.uwvkf_stream_push.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.uwvkf_stream_run.argtypes = [C.c_void_p, C.c_double]
        lib.uwvkf_stream_pending.argtypes = [C.c_void_p]
        lib.uwvkf_stream_time.argtypes = [C.c_void_p]
        lib.uwvkf_stream_time.restype = C.c_double
        lib.uwvkf_graph_begin.argtypes = [C.c_void_p]
        lib.uwvkf_graph_end.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.uwvkf_graph_launch.argtypes = [C.c_void_p]
        lib.uwvkf_graph_destroy.argtypes = [C.c_void_p]
        lib.uwvkf_debug_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        vp, ci = C.c_void_p, C.c_int
        lib.uwvkf_velocity_setup_motion_model.argtypes = [vp, vp]
        lib.uwvkf_velocity_set_gyro.argtypes = [vp, vp, ci]
        lib.uwvkf_velocity_set_efforts.argtypes = [vp, vp, ci]
        lib.uwvkf_velocity_update.argtypes = [vp, ci, vp, vp, vp, ci]
        lib.uwvkf_velocity_get_motion_model_state.argtypes = [vp, vp]
        lib.uwvkf_bottom_set_velocity.argtypes = [vp, vp, ci]
        lib.uwvkf_bottom_update_range.argtypes = [vp, vp, vp, vp, vp, vp, ci]
        lib.uwvkf_bottom_update_normal.argtypes = [vp, vp, vp, vp, ci]
        lib.uwvkf_indirect_set_tau.argtypes = [vp, C.c_double]
        lib.uwvkf_indirect_update_pose_reference.argtypes = [vp, vp, ci]
        lib.uwvkf_indirect_update_features.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, ci]
        lib.uwvkf_indirect_get_corrected_pose.argtypes = [vp, vp]
        _lib = lib
    return _lib


def _ptr(a):
    """numpy array -> host pointer, int -> raw (device) pointer, None -> NULL.

    The returned c_void_p keeps a reference to the array, so a converted temporary (`_ptr(_f64([0, 0, 1]))`) stays alive for
    as long as the pointer object does -- i.e. for the whole foreign call it is an argument of."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "contiguous numpy array or raw pointer expected"
    p = C.c_void_p(a.ctypes.data)
    p._keepalive = a
    return p


def _f64(a):
    if a is None or isinstance(a, int):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(valid):
    if valid is None or isinstance(valid, int):
        return valid
    return np.ascontiguousarray(valid, dtype=np.uint8)


class _FilterBatch:
    """Common part of the batched filter mirrors: handle lifetime, state access, status, ensemble statistics."""
    KIND, DIM, DOF = None, 0, 0

    def __init__(self, batch, device=0, stream=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.batch = int(batch)
        self._ck(self._lib.uwvkf_create(self.KIND, self.batch, device, C.byref(self._h)))
        if stream is not None:
            self._ck(self._lib.uwvkf_set_stream(self._h, C.c_void_p(stream)))

    def _ck(self, rc):
        if rc != 0:
            raise UwvkfError(rc, self._lib.uwvkf_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.uwvkf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def graph_begin(self):
        """Start capturing the DEVICE-memspace calls that follow into a CUDA graph (the handle needs a user stream)."""
        self._ck(self._lib.uwvkf_graph_begin(self._h))

    def graph_end(self):
        g = C.c_void_p()
        self._ck(self._lib.uwvkf_graph_end(self._h, C.byref(g)))
        return g

    def graph_launch(self, g):
        self._ck(self._lib.uwvkf_graph_launch(g))

    def graph_destroy(self, g):
        self._lib.uwvkf_graph_destroy(g)

    def synchronize(self):
        self._ck(self._lib.uwvkf_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(self._lib.uwvkf_kernel_launches(self._h))

    def set_mean_policy(self, tol, max_it):
        self._ck(self._lib.uwvkf_set_mean_policy(self._h, float(tol), int(max_it)))

    def set_process_noise(self, Q):
        """setProcessNoiseCovariance of the reference base class."""
        Q = _f64(Q)
        assert Q.shape == (self.DOF, self.DOF)
        self._ck(self._lib.uwvkf_set_process_noise(self._h, _ptr(Q)))

    def set_state(self, mu, cov, first=0, memspace=HOST, count=None):
        """initializeFilter(state, cov); cov is [count,DOF,DOF] or a single shared [DOF,DOF]."""
        mu, cov = _f64(mu), _f64(cov)
        if count is None:
            count = mu.shape[0]
        stride = self.DOF * self.DOF if (isinstance(cov, int) or cov.ndim == 3) else 0
        self._ck(self._lib.uwvkf_set_state(self._h, _ptr(mu), _ptr(cov), stride, first, count, memspace))

    def get_state(self, first=0, count=None, with_cov=True):
        count = self.batch - first if count is None else count
        mu = np.empty((count, self.DIM))
        cov = np.empty((count, self.DOF, self.DOF)) if with_cov else None
        self._ck(self._lib.uwvkf_get_state(self._h, _ptr(mu), _ptr(cov), first, count, HOST))
        return (mu, cov) if with_cov else mu

    def get_state_into(self, mu, cov=None, first=0, count=None, memspace=HOST):
        count = self.batch - first if count is None else count
        self._ck(self._lib.uwvkf_get_state(self._h, _ptr(mu), _ptr(cov), first, count, memspace))

    def get_status(self, first=0, count=None):
        count = self.batch - first if count is None else count
        st = np.empty(count, dtype=np.uint32)
        self._ck(self._lib.uwvkf_get_status(self._h, _ptr(st), first, count))
        return st

    def status_counts(self):
        """{bit name: number of filters with that status bit set} (device-side reduction)."""
        c = np.zeros(4, dtype=np.int64)
        self._ck(self._lib.uwvkf_status_counts(self._h, _ptr(c)))
        return dict(zip(("not_spd", "nan_measurement", "gate_rejected", "mean_not_converged"), (int(x) for x in c)))

    def clear_status(self):
        self._ck(self._lib.uwvkf_clear_status(self._h))

    def prediction_step(self, dt):
        """predictionStep(dt) of the reference base class -> predictionStepImpl of the filter."""
        self._ck(self._lib.uwvkf_predict(self._h, float(dt)))

    def ensemble_stats_size(self):
        return int(self._lib.uwvkf_ensemble_stats_size(self._h))

    def ensemble_stats(self, out=None, memspace=HOST):
        """[count, sum mu, sum mu mu^T (lower, row-major)] over this rank's filters; all-reduce (sum) across ranks."""
        if out is None:
            out = np.empty(self.ensemble_stats_size())
        self._ck(self._lib.uwvkf_ensemble_stats(self._h, _ptr(out), memspace))
        return out


class VelocityUKFBatch(_FilterBatch):
    """Batch of uwv_kalman_filters::VelocityUKF (reference src/VelocityUKF.hpp:33-66)."""
    KIND, DIM, DOF = FILTER_VELOCITY, 4, 4
    DVL, PRESSURE = 0, 1

    def setup_motion_model(self, uwv_params):
        """setupMotionModel(parameters) (src/VelocityUKF.cpp:58-77); uwv_params: flat UWVParameters (116)."""
        p = _f64(uwv_params)
        assert p.shape == (116,)
        self._ck(self._lib.uwvkf_velocity_setup_motion_model(self._h, _ptr(p)))

    def integrate_gyro(self, w, memspace=HOST):
        self._ck(self._lib.uwvkf_velocity_set_gyro(self._h, _ptr(_f64(w)), memspace))

    def integrate_body_efforts(self, tau, memspace=HOST):
        self._ck(self._lib.uwvkf_velocity_set_efforts(self._h, _ptr(_f64(tau)), memspace))

    def integrate_measurement(self, kind, z, R, valid=None, memspace=HOST):
        self._ck(self._lib.uwvkf_velocity_update(self._h, kind, _ptr(_f64(z)), _ptr(_f64(R)), _ptr(_u8(valid)), memspace))

    def get_motion_model_state(self):
        out = np.empty((self.batch, 10))
        self._ck(self._lib.uwvkf_velocity_get_motion_model_state(self._h, _ptr(out)))
        return out


class BottomUKFBatch(_FilterBatch):
    """Batch of uwv_kalman_filters::BottomUKF (reference src/BottomUKF.hpp:27-52)."""
    KIND, DIM, DOF = FILTER_BOTTOM, 4, 3

    def set_velocity(self, v, memspace=HOST):
        self._ck(self._lib.uwvkf_bottom_set_velocity(self._h, _ptr(_f64(v)), memspace))

    def integrate_range(self, z, R, unit_direction, origin, valid=None, memspace=HOST):
        self._ck(self._lib.uwvkf_bottom_update_range(self._h, _ptr(_f64(z)), _ptr(_f64(np.atleast_1d(R))), _ptr(_f64(unit_direction)),
                                                     _ptr(_f64(origin)), _ptr(_u8(valid)), memspace))

    def integrate_normal(self, z, R, valid=None, memspace=HOST):
        self._ck(self._lib.uwvkf_bottom_update_normal(self._h, _ptr(_f64(z)), _ptr(_f64(R)), _ptr(_u8(valid)), memspace))


class IndirectPoseUKFBatch(_FilterBatch):
    """Batch of uwv_kalman_filters::IndirectPoseUKF (reference src/IndirectPoseUKF.hpp:29-85)."""
    KIND, DIM, DOF = FILTER_INDIRECT_POSE, 7, 6

    def __init__(self, batch, position_error_std, orientation_error_std, orientation_error_tau,
                 initial_position_error=(0.0, 0.0, 0.0), initial_position_error_std=(1.0, 1.0, 1.0), device=0, stream=None):
        """Mirrors the constructor (src/IndirectPoseUKF.cpp:53-78): state, covariance and process noise from the stds."""
        super().__init__(batch, device, stream)
        pes, oes = np.asarray(position_error_std, float), np.asarray(orientation_error_std, float)
        mu = np.zeros(7)
        mu[0:3] = initial_position_error
        mu[6] = 1.0
        cov = np.diag(np.concatenate([np.asarray(initial_position_error_std, float) ** 2, oes ** 2]))
        self.set_state(np.tile(mu, (self.batch, 1)), cov)
        self.set_process_noise(np.diag(np.concatenate([pes ** 2, oes ** 2])))
        self._ck(self._lib.uwvkf_indirect_set_tau(self._h, float(orientation_error_tau)))

    def update_pose_reference(self, pose_ref, memspace=HOST):
        self._ck(self._lib.uwvkf_indirect_update_pose_reference(self._h, _ptr(_f64(pose_ref)), memspace))

    def integrate_features(self, px, pxcov, feature_pos, marker_pose, cov_marker, camera, cam_in_body, valid=None, memspace=HOST):
        """integrateMeasurement(marker_features, ...) (src/IndirectPoseUKF.cpp:94-135); px: [B, nfeat, 2]."""
        px = _f64(px)
        nfeat = np.asarray(feature_pos).shape[0]
        self._ck(self._lib.uwvkf_indirect_update_features(self._h, nfeat, _ptr(px), _ptr(_f64(pxcov)), _ptr(_f64(feature_pos)),
                                                          _ptr(_f64(marker_pose)), _ptr(_f64(cov_marker)), _ptr(_f64(camera)),
                                                          _ptr(_f64(cam_in_body)), _ptr(_u8(valid)), memspace))

    def get_corrected_pose(self):
        out = np.empty((self.batch, 7))
        self._ck(self._lib.uwvkf_indirect_get_corrected_pose(self._h, _ptr(out)))
        return out


STREAM_ROTATION_RATE = 100


class PoseSampleStream:
    """Caller-side loop for a PoseUKFBatch (include/uwvkf.h, sample stream): push timestamped sample sets, replay in time order."""

    def __init__(self, ukf, slots=64):
        self._ukf, self._lib = ukf, ukf._lib
        self._s = C.c_void_p()
        ukf._ck(self._lib.uwvkf_stream_create(ukf._h, int(slots), C.byref(self._s)))

    def push(self, timestamp, kind, z, R=None, aux=None, valid=None):
        self._ukf._ck(self._lib.uwvkf_stream_push(self._s, float(timestamp), int(kind), _ptr(_f64(z)), _ptr(_f64(R)), _ptr(_f64(aux)),
                                                  _ptr(_u8(valid))))

    def run(self, until=float("inf")):
        self._ukf._ck(self._lib.uwvkf_stream_run(self._s, float(until)))

    @property
    def pending(self):
        return int(self._lib.uwvkf_stream_pending(self._s))

    @property
    def time(self):
        return float(self._lib.uwvkf_stream_time(self._s))

    def close(self):
        if self._s.value:
            self._lib.uwvkf_stream_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PoseUKFBatch(_FilterBatch):
    """A batch of independent uwv_kalman_filters::PoseUKF instances on one GPU.

    Mirrors the reference class (src/PoseUKF.hpp:40-208): the second constructor (state, covariance,
    location, parameters; :113-115), predictionStep(dt), integrateMeasurement(<typed measurement>), and the
    getters.  Batch = 1 behaves like one reference object.  Array arguments are numpy arrays (host) or raw
    integer device pointers with memspace=DEVICE.
    """

    KIND, DIM, DOF = FILTER_POSE, 54, 53

    def __init__(self, batch, params, device=0, stream=None):
        super().__init__(batch, device, stream)
        self.set_params(params)

    def set_params(self, params):
        p = _f64(params)
        assert p.shape == (22,)
        self._ck(self._lib.uwvkf_pose_set_params(self._h, _ptr(p)))

    def set_model_parameters(self, uwv_params):
        """model_parameters of the constructors (flat uwv_dynamic_model::UWVParameters, 116): needed by the body-efforts updates."""
        p = _f64(uwv_params)
        assert p.shape == (116,)
        self._ck(self._lib.uwvkf_pose_set_model_parameters(self._h, _ptr(p)))

    def set_offsets(self, offsets, first=0):
        o = _f64(offsets)
        self._ck(self._lib.uwvkf_pose_set_offsets(self._h, _ptr(o), first, o.shape[0]))

    # -- getters
    def get_navigation(self, first=0, count=None):
        """p(3), q(x,y,z,w), v(3) of every filter: the part of getCurrentState() read every step."""
        count = self.batch - first if count is None else count
        out = np.empty((count, 10))
        self._ck(self._lib.uwvkf_pose_get_navigation(self._h, _ptr(out), first, count, HOST))
        return out

    def get_navigation_into(self, out, first=0, count=None, memspace=HOST):
        count = self.batch - first if count is None else count
        self._ck(self._lib.uwvkf_pose_get_navigation(self._h, _ptr(out), first, count, memspace))

    def get_rotation_rate(self):
        """getRotationRate() (reference src/PoseUKF.cpp:693-699) for every filter."""
        out = np.empty((self.batch, 3))
        self._ck(self._lib.uwvkf_pose_get_rotation_rate(self._h, _ptr(out)))
        return out

    # -- the filter API
    def integrate_rotation_rate(self, w, memspace=HOST):
        """integrateMeasurement(const RotationRate&) (src/PoseUKF.cpp:492-496)."""
        self._ck(self._lib.uwvkf_pose_set_rotation_rate(self._h, _ptr(_f64(w)), memspace))

    def integrate_measurement(self, kind, z, R, aux=None, valid=None, memspace=HOST, per_filter_R=None):
        """integrateMeasurement(<kind>) (src/PoseUKF.cpp:476-611). R: [m,m] shared or [B,m,m] (per_filter_R says which when R
        is a raw device pointer)."""
        z, R, aux = _f64(z), _f64(R), _f64(aux)
        per = int(per_filter_R) if per_filter_R is not None else (0 if isinstance(R, int) else int(R.ndim == 3))
        if valid is not None and not isinstance(valid, int):
            valid = np.ascontiguousarray(valid, dtype=np.uint8)
        self._ck(self._lib.uwvkf_pose_update(self._h, kind, _ptr(z), _ptr(R), per, _ptr(aux), _ptr(valid), memspace))

    def integrate_features(self, px, pxcov, feature_pos, marker_pose, cov_marker, camera, cam_in_imu, valid=None, memspace=HOST):
        """integrateMeasurement(marker_features, feature_positions, marker_pose, cov_marker_pose, camera_config, camera_in_IMU)
        (src/PoseUKF.cpp:613-654); px: [B, nfeat, 2] pixel coordinates, poses as t(3), q(x,y,z,w)."""
        nfeat = np.asarray(feature_pos).shape[0]
        self._ck(self._lib.uwvkf_pose_update_features(self._h, nfeat, _ptr(_f64(px)), _ptr(_f64(pxcov)), _ptr(_f64(feature_pos)),
                                                      _ptr(_f64(marker_pose)), _ptr(_f64(cov_marker)), _ptr(_f64(camera)),
                                                      _ptr(_f64(cam_in_imu)), _ptr(_u8(valid)), memspace))

    def predict_update(self, dt, w, kind, z, R, aux=None, valid=None, memspace=HOST, per_filter_R=None):
        """Fused predictionStep(dt) + integrateMeasurement(kind) in one kernel (one HBM round trip)."""
        w, z, R, aux = _f64(w), _f64(z), _f64(R), _f64(aux)
        per = int(per_filter_R) if per_filter_R is not None else (0 if isinstance(R, int) else int(R.ndim == 3))
        if valid is not None and not isinstance(valid, int):
            valid = np.ascontiguousarray(valid, dtype=np.uint8)
        self._ck(self._lib.uwvkf_pose_predict_update(self._h, float(dt), _ptr(w), kind, _ptr(z), _ptr(R), per, _ptr(aux),
                                                     _ptr(valid), memspace))

    def predict_update2(self, dt, w, kind_a, z_a, R_a, kind_b, z_b, R_b, valid_a=None, valid_b=None, memspace=HOST):
        """predictionStep(dt) + integrateMeasurement(kind_a) + integrateMeasurement(kind_b) in one kernel (DVL + accelerometer:
        the 100 Hz step of a full PoseUKF); dt = 0 skips the prediction."""
        w, z_a, R_a, z_b, R_b = _f64(w), _f64(z_a), _f64(R_a), _f64(z_b), _f64(R_b)
        self._ck(self._lib.uwvkf_pose_predict_update2(self._h, float(dt), _ptr(w), int(kind_a), _ptr(z_a), _ptr(R_a), _ptr(_u8(valid_a)),
                                                      int(kind_b), _ptr(z_b), _ptr(R_b), _ptr(_u8(valid_b)), memspace))

    def step_host(self, dt, w_host, kind, z_host, R, aux=None, nav_out=None):
        """Pipelined host-buffer step (uwvkf_pose_step_host): rotation rate + predictionStep(dt) + integrateMeasurement(kind)
        from HOST sample buffers, navigation solution back into nav_out [B, 10].  Arrays must be C-contiguous float64 (ideally
        pinned) or raw host addresses; see include/uwvkf.h for when they may be reused."""
        for a in (w_host, z_host, nav_out):
            assert a is None or isinstance(a, int) or (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"])
        R, aux = _f64(R), _f64(aux)
        self._ck(self._lib.uwvkf_pose_step_host(self._h, float(dt), _ptr(w_host), int(kind), _ptr(z_host), _ptr(R), _ptr(aux),
                                                _ptr(nav_out)))

    def join(self):
        self._ck(self._lib.uwvkf_join(self._h))

    DEBUG_TMERC_SINCOS_LAT, DEBUG_SO3_BOXPLUS, DEBUG_SO3_BOXMINUS, DEBUG_RSQRT, DEBUG_WARP_SUM = range(5)

    def debug_eval(self, what, rows):
        """Device-math probe (tests): evaluates one device function per input row, rows [n, 8] -> [n, 4]."""
        rows = _f64(rows)
        assert rows.ndim == 2 and rows.shape[1] == 8
        out = np.empty((rows.shape[0], 4))
        self._ck(self._lib.uwvkf_debug_eval(self._h, int(what), _ptr(rows), _ptr(out), rows.shape[0]))
        return out

    def debug_phase_cycles(self, reset=False):
        """Per-phase latency table of the fused kernel; only a -DUWVKF_PHASE_TIMING build (tools/phase_timing.py) fills it,
        the product library raises."""
        out = np.zeros(40, dtype=np.uint64)
        self._ck(self._lib.uwvkf_debug_phase_cycles(self._h, out.ctypes.data, int(bool(reset))))
        return out

    def reset_with_external_pose(self, pose7):
        """resetFilterWithExternalPose (src/PoseUKF.cpp:685-691)."""
        self._ck(self._lib.uwvkf_pose_reset_with_external_pose(self._h, _ptr(_f64(pose7))))
