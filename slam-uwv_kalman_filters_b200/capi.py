"""ctypes binding of libuwvkf.so (include/uwvkf.h) and a batched mirror of the reference's PoseUKF class API."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST, DEVICE = 0, 1
FILTER_POSE, FILTER_VELOCITY, FILTER_BOTTOM, FILTER_INDIRECT_POSE = 0, 1, 2, 3
(POSE_XY, POSE_Z, POSE_PRESSURE, POSE_ACCELERATION, POSE_VELOCITY, POSE_WATER_VELOCITY, POSE_GEOGRAPHIC,
 POSE_DELAYED_XY) = range(8)
POSE_MEAS_DIM = {POSE_XY: 2, POSE_Z: 1, POSE_PRESSURE: 1, POSE_ACCELERATION: 3, POSE_VELOCITY: 3, POSE_WATER_VELOCITY: 2,
                 POSE_GEOGRAPHIC: 2, POSE_DELAYED_XY: 2}
ST_NOT_SPD, ST_NAN_MEASUREMENT, ST_GATE_REJECTED, ST_MEAN_NOT_CONVERGED = 1, 2, 4, 8
ERR_NAN_MEASUREMENT = -4

_dp = C.POINTER(C.c_double)


class UwvkfError(RuntimeError):
    """Raised for any non-zero return code; the reference throws std::runtime_error in the same places."""

    def __init__(self, code, msg):
        super().__init__(f"uwvkf error {code}: {msg}")
        self.code = code


def library_path():
    return os.path.join(_HERE, "libuwvkf.so")


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
        lib.uwvkf_pose_set_params.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_pose_set_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        lib.uwvkf_pose_set_rotation_rate.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_pose_get_rotation_rate.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_predict.argtypes = [C.c_void_p, C.c_double]
        lib.uwvkf_pose_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_pose_predict_update.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.c_void_p, C.c_void_p, C.c_int]
        lib.uwvkf_pose_get_navigation.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        lib.uwvkf_pose_reset_with_external_pose.argtypes = [C.c_void_p, C.c_void_p]
        lib.uwvkf_ensemble_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def _ptr(a):
    """numpy array -> host pointer, int -> raw (device) pointer, None -> NULL."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "contiguous numpy array or raw pointer expected"
    return C.c_void_p(a.ctypes.data)


def _f64(a):
    if a is None or isinstance(a, int):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


class PoseUKFBatch:
    """A batch of independent uwv_kalman_filters::PoseUKF instances on one GPU.

    Mirrors the reference class (src/PoseUKF.hpp:40-208): the second constructor (state, covariance,
    location, parameters; :113-115), predictionStep(dt), integrateMeasurement(<typed measurement>), and the
    getters.  Batch = 1 behaves like one reference object.  Array arguments are numpy arrays (host) or raw
    integer device pointers with memspace=DEVICE.
    """

    DIM, DOF = 54, 53

    def __init__(self, batch, params, device=0, stream=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.batch = int(batch)
        self._ck(self._lib.uwvkf_create(FILTER_POSE, self.batch, device, C.byref(self._h)))
        if stream is not None:
            self._ck(self._lib.uwvkf_set_stream(self._h, C.c_void_p(stream)))
        self.set_params(params)

    # -- plumbing
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

    def synchronize(self):
        self._ck(self._lib.uwvkf_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(self._lib.uwvkf_kernel_launches(self._h))

    # -- configuration
    def set_params(self, params):
        p = _f64(params)
        assert p.shape == (22,)
        self._ck(self._lib.uwvkf_pose_set_params(self._h, _ptr(p)))

    def set_mean_policy(self, tol, max_it):
        self._ck(self._lib.uwvkf_set_mean_policy(self._h, float(tol), int(max_it)))

    def set_process_noise(self, Q):
        """setProcessNoiseCovariance (reference src/PoseUKF.cpp:438)."""
        Q = _f64(Q)
        assert Q.shape == (53, 53)
        self._ck(self._lib.uwvkf_set_process_noise(self._h, _ptr(Q)))

    def set_state(self, mu, cov, first=0, memspace=HOST, count=None):
        """initializeFilter(state, cov); cov is [count,53,53] or a single shared [53,53]."""
        mu, cov = _f64(mu), _f64(cov)
        if count is None:
            count = mu.shape[0]
        stride = 0
        if isinstance(cov, int) or cov.ndim == 3:
            stride = 53 * 53
        self._ck(self._lib.uwvkf_set_state(self._h, _ptr(mu), _ptr(cov), stride, first, count, memspace))

    def set_offsets(self, offsets, first=0):
        o = _f64(offsets)
        self._ck(self._lib.uwvkf_pose_set_offsets(self._h, _ptr(o), first, o.shape[0]))

    # -- getters
    def get_state(self, first=0, count=None, with_cov=True):
        count = self.batch - first if count is None else count
        mu = np.empty((count, 54))
        cov = np.empty((count, 53, 53)) if with_cov else None
        self._ck(self._lib.uwvkf_get_state(self._h, _ptr(mu), _ptr(cov), first, count, HOST))
        return (mu, cov) if with_cov else mu

    def get_state_into(self, mu, cov=None, first=0, count=None, memspace=HOST):
        count = self.batch - first if count is None else count
        self._ck(self._lib.uwvkf_get_state(self._h, _ptr(mu), _ptr(cov), first, count, memspace))

    def get_navigation(self, first=0, count=None):
        """p(3), q(x,y,z,w), v(3) of every filter: the part of getCurrentState() read every step."""
        count = self.batch - first if count is None else count
        out = np.empty((count, 10))
        self._ck(self._lib.uwvkf_pose_get_navigation(self._h, _ptr(out), first, count, HOST))
        return out

    def get_navigation_into(self, out, first=0, count=None, memspace=HOST):
        count = self.batch - first if count is None else count
        self._ck(self._lib.uwvkf_pose_get_navigation(self._h, _ptr(out), first, count, memspace))

    def get_status(self, first=0, count=None):
        count = self.batch - first if count is None else count
        st = np.empty(count, dtype=np.uint32)
        self._ck(self._lib.uwvkf_get_status(self._h, _ptr(st), first, count))
        return st

    def clear_status(self):
        self._ck(self._lib.uwvkf_clear_status(self._h))

    def get_rotation_rate(self):
        """getRotationRate() (reference src/PoseUKF.cpp:693-699) for every filter."""
        out = np.empty((self.batch, 3))
        self._ck(self._lib.uwvkf_pose_get_rotation_rate(self._h, _ptr(out)))
        return out

    # -- the filter API
    def integrate_rotation_rate(self, w, memspace=HOST):
        """integrateMeasurement(const RotationRate&) (src/PoseUKF.cpp:492-496)."""
        self._ck(self._lib.uwvkf_pose_set_rotation_rate(self._h, _ptr(_f64(w)), memspace))

    def prediction_step(self, dt):
        """predictionStep(dt) (src/PoseUKF.cpp:446-466)."""
        self._ck(self._lib.uwvkf_predict(self._h, float(dt)))

    def integrate_measurement(self, kind, z, R, aux=None, valid=None, memspace=HOST):
        """integrateMeasurement(<kind>) (src/PoseUKF.cpp:476-611). R: [m,m] shared or [B,m,m]."""
        z, R, aux = _f64(z), _f64(R), _f64(aux)
        per = 0 if isinstance(R, int) else int(R.ndim == 3)
        if valid is not None and not isinstance(valid, int):
            valid = np.ascontiguousarray(valid, dtype=np.uint8)
        self._ck(self._lib.uwvkf_pose_update(self._h, kind, _ptr(z), _ptr(R), per, _ptr(aux), _ptr(valid), memspace))

    def predict_update(self, dt, w, kind, z, R, aux=None, valid=None, memspace=HOST):
        """Fused predictionStep(dt) + integrateMeasurement(kind) in one kernel (one HBM round trip)."""
        w, z, R, aux = _f64(w), _f64(z), _f64(R), _f64(aux)
        per = 0 if isinstance(R, int) else int(R.ndim == 3)
        if valid is not None and not isinstance(valid, int):
            valid = np.ascontiguousarray(valid, dtype=np.uint8)
        self._ck(self._lib.uwvkf_pose_predict_update(self._h, float(dt), _ptr(w), kind, _ptr(z), _ptr(R), per, _ptr(aux),
                                                     _ptr(valid), memspace))

    def reset_with_external_pose(self, pose7):
        """resetFilterWithExternalPose (src/PoseUKF.cpp:685-691)."""
        self._ck(self._lib.uwvkf_pose_reset_with_external_pose(self._h, _ptr(_f64(pose7))))

    def ensemble_stats_size(self):
        return int(self._lib.uwvkf_ensemble_stats_size(self._h))

    def ensemble_stats(self, out=None, memspace=HOST):
        """[count, sum mu, sum mu mu^T (lower packed)] over this rank's filters; all-reduce (sum) across ranks."""
        if out is None:
            out = np.empty(self.ensemble_stats_size())
        self._ck(self._lib.uwvkf_ensemble_stats(self._h, _ptr(out), memspace))
        return out
