"""ctypes binding of the CPU oracle (oracle/libuwvkf_oracle.so) -- test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "libuwvkf_oracle.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

# measurement kinds (oracle_api.h)
XY, Z, PRESSURE, ACCELERATION, VELOCITY, WATER_VELOCITY, GEOGRAPHIC, DELAYED_XY, BODY_EFFORTS, BODY_EFFORTS_VEL = range(10)
MEAS_DIM = {XY: 2, Z: 1, PRESSURE: 1, ACCELERATION: 3, VELOCITY: 3, WATER_VELOCITY: 2, GEOGRAPHIC: 2, DELAYED_XY: 2,
            BODY_EFFORTS: 6, BODY_EFFORTS_VEL: 6}


def build():
    if not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_pose_step_batch.restype = C.c_double
        _lib.orc_gravity_wgs84.restype = C.c_double
        _lib.orc_gravity_wgs84.argtypes = [C.c_double, C.c_double]
        _lib.orc_tmerc_world_to_nav.argtypes = [C.c_double] * 4 + [dp]
        _lib.orc_tmerc_nav_to_world.argtypes = [C.c_double] * 4 + [dp]
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def pose_predict(mu, cov, Q, rot_rate, params, offsets, dt, policy=None):
    """One PoseUKF::predictionStep(dt). Returns (mu', cov', status, mean_iterations)."""
    mu, cov = _c(mu).copy(), _c(cov).copy()
    Q, rot_rate, params, offsets, policy = _c(Q), _c(rot_rate), _c(params), _c(offsets), _c(policy)
    it = C.c_int(0)
    st = lib().orc_pose_predict(_p(mu), _p(cov), _p(Q), _p(rot_rate), _p(params), _p(offsets), C.c_double(dt), _p(policy),
                                C.byref(it))
    return mu, cov, st, it.value


def pose_update(kind, mu, cov, z, R, aux, rot_rate, params, policy=None):
    """One PoseUKF::integrateMeasurement(kind). Returns (mu', cov', status, mahalanobis2)."""
    mu, cov = _c(mu).copy(), _c(cov).copy()
    z, R, aux, rot_rate, params, policy = _c(z), _c(R), _c(aux), _c(rot_rate), _c(params), _c(policy)
    d2 = C.c_double(0)
    st = lib().orc_pose_update(C.c_int(kind), _p(mu), _p(cov), _p(z), _p(R), _p(aux), _p(rot_rate), _p(params), _p(policy),
                               C.byref(d2))
    return mu, cov, st, d2.value


def pose_update_features(mu, cov, px, pxcov, feature_pos, marker_pose, cov_marker, camera, cam_in_imu, policy=None):
    """PoseUKF::integrateMeasurement(marker_corners, ...) on the marker-augmented state. Returns (mu', cov', status)."""
    mu, cov = _c(mu).copy(), _c(cov).copy()
    px, fp = _c(px), _c(feature_pos)
    st = lib().orc_pose_update_features(_p(mu), _p(cov), C.c_int(fp.shape[0]), _p(px), _p(_c(pxcov)), _p(fp), _p(_c(marker_pose)),
                                        _p(_c(cov_marker)), _p(_c(camera)), _p(_c(cam_in_imu)), _p(_c(policy)))
    return mu, cov, st


def pose_rotation_rate(mu, rot_rate, params):
    out = np.zeros(3)
    lib().orc_pose_rotation_rate(_p(_c(mu)), _p(_c(rot_rate)), _p(_c(params)), _p(out))
    return out


def pose_init_from_config(pos, pos_cov, quat, rot_cov, cfg, inertia6, lin6, quad6, imu_in_body_rot=None):
    mu, cov = np.zeros(54), np.zeros((53, 53))
    Rib = np.eye(3) if imu_in_body_rot is None else imu_in_body_rot
    lib().orc_pose_init_from_config(_p(_c(pos)), _p(_c(pos_cov)), _p(_c(quat)), _p(_c(rot_cov)), _p(_c(cfg)), _p(_c(inertia6)),
                                    _p(_c(lin6)), _p(_c(quad6)), _p(_c(Rib)), _p(mu), _p(cov))
    return mu, cov


def pose_process_noise_from_config(cfg, imu_dt, imu_in_body_rot=None):
    Q = np.zeros((53, 53))
    Rib = np.eye(3) if imu_in_body_rot is None else imu_in_body_rot
    lib().orc_pose_process_noise_from_config(_p(_c(cfg)), C.c_double(imu_dt), _p(_c(Rib)), _p(Q))
    return Q


def pose_step_batch(mu, cov, Q, rot_rate, z_vel, R_vel, params, offsets, dt, nsteps=1, nthreads=0, policy=None):
    """In-place batched predict (+ DVL update when z_vel is not None); returns (seconds, status[B])."""
    B = mu.shape[0]
    st = np.zeros(B, dtype=np.int32)
    t = lib().orc_pose_step_batch(C.c_int(B), _p(mu), _p(cov), _p(_c(Q)), _p(_c(rot_rate)), _p(_c(z_vel)), _p(_c(R_vel)),
                                  _p(_c(params)), _p(_c(offsets)), C.c_double(dt), C.c_int(nsteps), C.c_int(nthreads),
                                  _p(_c(policy)), st.ctypes.data_as(ip))
    return t, st


def num_threads():
    return lib().orc_num_threads()


def tmerc_world_to_nav(lat0, lon0, lat, lon):
    out = np.zeros(2)
    lib().orc_tmerc_world_to_nav(lat0, lon0, lat, lon, _p(out))
    return out


def tmerc_nav_to_world(lat0, lon0, x, y):
    out = np.zeros(2)
    lib().orc_tmerc_nav_to_world(lat0, lon0, x, y, _p(out))
    return out


def gravity_wgs84(lat, alt):
    return lib().orc_gravity_wgs84(lat, alt)


def linear_predict(mu, cov, F, c, Q):
    n = len(mu)
    mu, cov = _c(mu).copy(), _c(cov).copy()
    st = lib().orc_linear_predict(C.c_int(n), _p(mu), _p(cov), _p(_c(F)), _p(_c(c)), _p(_c(Q)))
    return mu, cov, st


def linear_update(mu, cov, H, z, R):
    n, m = len(mu), len(z)
    mu, cov = _c(mu).copy(), _c(cov).copy()
    st = lib().orc_linear_update(C.c_int(n), C.c_int(m), _p(mu), _p(cov), _p(_c(H)), _p(_c(z)), _p(_c(R)))
    return mu, cov, st


def cholesky_lower(A):
    n = A.shape[0]
    L = np.zeros((n, n))
    st = lib().orc_cholesky_lower(C.c_int(n), _p(_c(A)), _p(L))
    return L, st


def so3_boxplus(q, v, scale=1.0):
    out = np.zeros(4)
    lib().orc_so3_boxplus(_p(_c(q)), _p(_c(v)), C.c_double(scale), _p(out))
    return out


def so3_boxminus(a, b):
    out = np.zeros(3)
    lib().orc_so3_boxminus(_p(_c(a)), _p(_c(b)), _p(out))
    return out


def s2_boxplus(x, d, scale=1.0):
    out = np.zeros(3)
    lib().orc_s2_boxplus(_p(_c(x)), _p(_c(d)), C.c_double(scale), _p(out))
    return out


def s2_boxminus(y, x):
    out = np.zeros(2)
    lib().orc_s2_boxminus(_p(_c(y)), _p(_c(x)), _p(out))
    return out


def quat_rot(q, v):
    out = np.zeros(3)
    lib().orc_quat_rot(_p(_c(q)), _p(_c(v)), _p(out))
    return out


# ---- small filters ------------------------------------------------------------------------------------------
def vel_setup_motion_model(mu, gyro):
    sim = np.zeros(10)
    lib().orc_vel_setup_motion_model(_p(_c(mu)), _p(_c(gyro)), _p(sim))
    return sim


def vel_predict(mu, cov, Q, sim, gyro, efforts, uwv, dt, policy=None):
    mu, cov, sim = _c(mu).copy(), _c(cov).copy(), _c(sim).copy()
    st = lib().orc_vel_predict(_p(mu), _p(cov), _p(_c(Q)), _p(sim), _p(_c(gyro)), _p(_c(efforts)), _p(_c(uwv)), C.c_double(dt),
                               _p(_c(policy)))
    return mu, cov, sim, st


def vel_update(kind, mu, cov, z, R, policy=None):
    mu, cov = _c(mu).copy(), _c(cov).copy()
    st = lib().orc_vel_update(C.c_int(kind), _p(mu), _p(cov), _p(_c(z)), _p(_c(R)), _p(_c(policy)))
    return mu, cov, st


def bottom_predict(mu, cov, Q, velocity, dt, policy=None):
    mu, cov = _c(mu).copy(), _c(cov).copy()
    st = lib().orc_bottom_predict(_p(mu), _p(cov), _p(_c(Q)), _p(_c(velocity)), C.c_double(dt), _p(_c(policy)))
    return mu, cov, st


def bottom_update_range(mu, cov, z, R, unit_direction, origin, policy=None):
    mu, cov = _c(mu).copy(), _c(cov).copy()
    st = lib().orc_bottom_update_range(_p(mu), _p(cov), _p(_c(z)), _p(_c(R)), _p(_c(unit_direction)), _p(_c(origin)),
                                       _p(_c(policy)))
    return mu, cov, st


def bottom_update_normal(mu, cov, z, R, policy=None):
    mu, cov = _c(mu).copy(), _c(cov).copy()
    st = lib().orc_bottom_update_normal(_p(mu), _p(cov), _p(_c(z)), _p(_c(R)), _p(_c(policy)))
    return mu, cov, st


def indirect_predict(mu, cov, Q, tau, dt, policy=None):
    mu, cov = _c(mu).copy(), _c(cov).copy()
    st = lib().orc_indirect_predict(_p(mu), _p(cov), _p(_c(Q)), C.c_double(tau), C.c_double(dt), _p(_c(policy)))
    return mu, cov, st


def indirect_update_features(mu, cov, px, pxcov, feature_pos, marker_pose, cov_marker, camera, cam_in_body, pose_ref,
                             policy=None):
    mu, cov = _c(mu).copy(), _c(cov).copy()
    px = _c(px)
    st = lib().orc_indirect_update_features(_p(mu), _p(cov), C.c_int(px.shape[0]), _p(px), _p(_c(pxcov)), _p(_c(feature_pos)),
                                            _p(_c(marker_pose)), _p(_c(cov_marker)), _p(_c(camera)), _p(_c(cam_in_body)),
                                            _p(_c(pose_ref)), _p(_c(policy)))
    return mu, cov, st


def indirect_corrected_pose(mu, pose_ref):
    out = np.zeros(7)
    lib().orc_indirect_corrected_pose(_p(_c(mu)), _p(_c(pose_ref)), _p(out))
    return out
