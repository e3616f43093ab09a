"""Synthetic PoseUKF scenario (SURVEY.md Appendix F "SYNTH-1"): configuration, initial state and sensor streams.

The reference ships no test sequence (test/suite.cpp:1-6 is an empty stub), so this is the authored
workload for BASELINE.json's configs.  All randomness comes from a splitmix64 integer hash of
(seed, filter, step, channel) so that any process -- CPU oracle, GPU engine, any rank -- regenerates
bit-identical inputs for a given global filter index.
"""
import numpy as np

SEED = 20181018
LAT0, LON0, ALT0 = 0.92677, 0.15446, 0.0
EARTHW = 7.292115e-5
P_ATM = 101325.0
RHO = 1025.0

# flat PoseUKFConfig indices (mirrors include/uwvkf.h UWVKF_CFG_*)
CFG_N = 65
(CFG_ACC_RANDOMWALK, CFG_ACC_BIAS_OFFSET, CFG_ACC_BIAS_INSTABILITY, CFG_ACC_BIAS_TAU) = (0, 3, 6, 9)
(CFG_GYRO_RANDOMWALK, CFG_GYRO_BIAS_OFFSET, CFG_GYRO_BIAS_INSTABILITY, CFG_GYRO_BIAS_TAU) = (10, 13, 16, 19)
(CFG_INERTIA_INSTABILITY, CFG_LIN_DAMPING_INSTABILITY, CFG_QUAD_DAMPING_INSTABILITY) = (20, 29, 38)
(CFG_INERTIA_TAU, CFG_LIN_DAMPING_TAU, CFG_QUAD_DAMPING_TAU) = (47, 48, 49)
(CFG_WV_TAU, CFG_WV_LIMITS, CFG_WV_SCALE, CFG_ADCP_BIAS_TAU, CFG_ADCP_BIAS_LIMITS) = (50, 51, 52, 53, 54)
(CFG_LATITUDE, CFG_LONGITUDE, CFG_ALTITUDE) = (55, 56, 57)
(CFG_WATER_DENSITY, CFG_WATER_DENSITY_LIMITS, CFG_WATER_DENSITY_TAU, CFG_ATM_PRESSURE) = (58, 59, 60, 61)
CFG_MAX_JERK = 62

# flat PoseUKFParameter indices (mirrors UWVKF_PP_*)
PP_N = 22


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform(filt, step, chan, seed=SEED):
    """u in [0,1): hash of (seed, filter index, step, channel); broadcasts over numpy arrays."""
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) ^ (np.asarray(filt, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
               ^ (np.asarray(step, dtype=np.uint64) * np.uint64(0xBF58476D1CE4E5B9)) ^ np.asarray(chan, dtype=np.uint64))
        h = splitmix64(key)
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def noise(filt, step, chan, seed=SEED):
    """unit-variance Irwin-Hall(4) noise: sqrt(3) * (u1+u2+u3+u4 - 2)."""
    c = np.asarray(chan, dtype=np.uint64) * np.uint64(4)
    s = (uniform(filt, step, c, seed) + uniform(filt, step, c + np.uint64(1), seed)
         + uniform(filt, step, c + np.uint64(2), seed) + uniform(filt, step, c + np.uint64(3), seed))
    return np.sqrt(3.0) * (s - 2.0)


def gravity_wgs84(lat, alt):
    g0, k, e, a = 9.7803267714, 0.00193185138639, 0.0818191908426, 6378137.0
    s2 = np.sin(lat) ** 2
    return g0 * (1 + k * s2) / np.sqrt(1 - e * e * s2) * (a / (a + alt)) ** 2


def pose_config():
    c = np.zeros(CFG_N)
    c[CFG_ACC_RANDOMWALK:CFG_ACC_RANDOMWALK + 3] = 1e-3
    c[CFG_ACC_BIAS_INSTABILITY:CFG_ACC_BIAS_INSTABILITY + 3] = 1e-4
    c[CFG_ACC_BIAS_TAU] = 3600.0
    c[CFG_GYRO_RANDOMWALK:CFG_GYRO_RANDOMWALK + 3] = 1e-5
    c[CFG_GYRO_BIAS_INSTABILITY:CFG_GYRO_BIAS_INSTABILITY + 3] = 1e-6
    c[CFG_GYRO_BIAS_TAU] = 3600.0
    c[CFG_INERTIA_INSTABILITY:CFG_INERTIA_INSTABILITY + 9] = 5.0
    c[CFG_LIN_DAMPING_INSTABILITY:CFG_LIN_DAMPING_INSTABILITY + 9] = 5.0
    c[CFG_QUAD_DAMPING_INSTABILITY:CFG_QUAD_DAMPING_INSTABILITY + 9] = 5.0
    c[CFG_INERTIA_TAU] = c[CFG_LIN_DAMPING_TAU] = c[CFG_QUAD_DAMPING_TAU] = 1e4
    c[CFG_WV_TAU], c[CFG_WV_LIMITS], c[CFG_WV_SCALE] = 1800.0, 0.3, 1e-4
    c[CFG_ADCP_BIAS_TAU], c[CFG_ADCP_BIAS_LIMITS] = 3600.0, 0.05
    c[CFG_LATITUDE], c[CFG_LONGITUDE], c[CFG_ALTITUDE] = LAT0, LON0, ALT0
    c[CFG_WATER_DENSITY], c[CFG_WATER_DENSITY_LIMITS], c[CFG_WATER_DENSITY_TAU] = RHO, 5.0, 3600.0
    c[CFG_ATM_PRESSURE] = P_ATM
    c[CFG_MAX_JERK:CFG_MAX_JERK + 3] = 4.0
    return c


def uwv_matrices():
    M = np.diag([200.0, 250.0, 300.0, 20.0, 40.0, 40.0])
    Dl = np.diag([30.0, 60.0, 80.0, 5.0, 10.0, 10.0])
    Dq = np.diag([60.0, 120.0, 160.0, 10.0, 20.0, 20.0])
    return M, Dl, Dq


def pose_params(cfg):
    """PoseUKFParameter as filled by the first constructor (reference src/PoseUKF.cpp:358-371), IMU == body."""
    p = np.zeros(PP_N)
    p[0], p[1] = cfg[CFG_LATITUDE], cfg[CFG_LONGITUDE]
    p[2] = cfg[CFG_GYRO_BIAS_TAU]
    p[3:6] = cfg[CFG_GYRO_BIAS_OFFSET:CFG_GYRO_BIAS_OFFSET + 3]
    p[6] = cfg[CFG_ACC_BIAS_TAU]
    p[7:10] = cfg[CFG_ACC_BIAS_OFFSET:CFG_ACC_BIAS_OFFSET + 3]
    p[10], p[11], p[12] = cfg[CFG_INERTIA_TAU], cfg[CFG_LIN_DAMPING_TAU], cfg[CFG_QUAD_DAMPING_TAU]
    p[13], p[14], p[15] = cfg[CFG_WV_TAU], cfg[CFG_WV_LIMITS], cfg[CFG_WV_SCALE]
    p[16], p[17], p[18] = cfg[CFG_ADCP_BIAS_TAU], cfg[CFG_ATM_PRESSURE], cfg[CFG_WATER_DENSITY_TAU]
    return p


def pose_initial(cfg):
    """(mu0[54], cov0[53,53]) exactly as the first constructor builds them (src/PoseUKF.cpp:293-342), IMU == body."""
    M, Dl, Dq = uwv_matrices()
    mu = np.zeros(54)
    mu[6] = 1.0  # quaternion (x, y, z, w)
    mu[13:16] = cfg[CFG_GYRO_BIAS_OFFSET:CFG_GYRO_BIAS_OFFSET + 3]
    mu[16:19] = cfg[CFG_ACC_BIAS_OFFSET:CFG_ACC_BIAS_OFFSET + 3]
    mu[19] = gravity_wgs84(cfg[CFG_LATITUDE], cfg[CFG_ALTITUDE])
    idx = [0, 1, 5]
    for name, off in ((M, 20), (Dl, 29), (Dq, 38)):
        blk = name[np.ix_(idx, idx)]
        mu[off:off + 9] = blk.flatten(order="F")
    mu[53] = cfg[CFG_WATER_DENSITY]
    d = np.zeros(53)
    d[0:3] = 1.0
    d[3:6] = [1e-4, 1e-4, 1e-2]
    d[6:9] = 1.0
    d[9:12] = 10.0
    d[12:15] = cfg[CFG_GYRO_BIAS_INSTABILITY:CFG_GYRO_BIAS_INSTABILITY + 3] ** 2
    d[15:18] = cfg[CFG_ACC_BIAS_INSTABILITY:CFG_ACC_BIAS_INSTABILITY + 3] ** 2
    d[18] = 0.05 ** 2
    d[19:28] = cfg[CFG_INERTIA_INSTABILITY:CFG_INERTIA_INSTABILITY + 9] ** 2
    d[28:37] = cfg[CFG_LIN_DAMPING_INSTABILITY:CFG_LIN_DAMPING_INSTABILITY + 9] ** 2
    d[37:46] = cfg[CFG_QUAD_DAMPING_INSTABILITY:CFG_QUAD_DAMPING_INSTABILITY + 9] ** 2
    d[46:48] = cfg[CFG_WV_LIMITS] ** 2
    d[48:50] = cfg[CFG_WV_LIMITS] ** 2
    d[50:52] = cfg[CFG_ADCP_BIAS_LIMITS] ** 2
    d[52] = cfg[CFG_WATER_DENSITY_LIMITS] ** 2
    return mu, np.diag(d)


def pose_process_noise(cfg, T):
    """Q exactly as setProcessNoiseFromConfig builds it (src/PoseUKF.cpp:398-438), IMU == body."""
    j = cfg[CFG_MAX_JERK:CFG_MAX_JERK + 3]
    d = np.zeros(53)
    d[0:3] = 1.5 * (T ** 4.0 * ((1. / 6.) * 0.25 * j) ** 2)
    d[6:9] = 1.5 * (T ** 2.0 * (0.5 * 0.25 * j) ** 2)
    d[9:12] = (0.25 * j) ** 2
    d[3:6] = cfg[CFG_GYRO_RANDOMWALK:CFG_GYRO_RANDOMWALK + 3] ** 2
    d[12:15] = (2. / (cfg[CFG_GYRO_BIAS_TAU] * T)) * cfg[CFG_GYRO_BIAS_INSTABILITY:CFG_GYRO_BIAS_INSTABILITY + 3] ** 2
    d[15:18] = (2. / (cfg[CFG_ACC_BIAS_TAU] * T)) * cfg[CFG_ACC_BIAS_INSTABILITY:CFG_ACC_BIAS_INSTABILITY + 3] ** 2
    d[18] = 1e-12
    d[19:28] = (2. / (cfg[CFG_INERTIA_TAU] * T)) * cfg[CFG_INERTIA_INSTABILITY:CFG_INERTIA_INSTABILITY + 9] ** 2
    d[28:37] = (2. / (cfg[CFG_LIN_DAMPING_TAU] * T)) * cfg[CFG_LIN_DAMPING_INSTABILITY:CFG_LIN_DAMPING_INSTABILITY + 9] ** 2
    d[37:46] = (2. / (cfg[CFG_QUAD_DAMPING_TAU] * T)) * cfg[CFG_QUAD_DAMPING_INSTABILITY:CFG_QUAD_DAMPING_INSTABILITY + 9] ** 2
    d[46:48] = (2. / (cfg[CFG_WV_TAU] * T)) * cfg[CFG_WV_LIMITS] ** 2
    d[48:50] = (2. / (cfg[CFG_WV_TAU] * T)) * cfg[CFG_WV_LIMITS] ** 2
    d[50:52] = (2. / (cfg[CFG_ADCP_BIAS_TAU] * T)) * cfg[CFG_ADCP_BIAS_LIMITS] ** 2
    d[52] = (2. / (cfg[CFG_WATER_DENSITY_TAU] * T)) * cfg[CFG_WATER_DENSITY_LIMITS] ** 2
    return np.diag(d)


# ---- truth trajectory (analytic; identical for all filters) ---------------------------------------------
def truth(t):
    """Returns dict of truth quantities at time(s) t: body velocity, yaw, yaw rate, nav velocity/acceleration, z."""
    t = np.asarray(t, dtype=np.float64)
    vb = np.stack([1.0 + 0.2 * np.sin(0.2 * t), 0.1 * np.sin(0.1 * t), 0.05 * np.sin(0.05 * t)], -1)
    vb_dot = np.stack([0.04 * np.cos(0.2 * t), 0.01 * np.cos(0.1 * t), 0.0025 * np.cos(0.05 * t)], -1)
    yaw_rate = 0.05 * np.sin(0.02 * t)
    yaw = 2.5 * (1.0 - np.cos(0.02 * t))
    c, s = np.cos(yaw), np.sin(yaw)
    vn = np.stack([c * vb[..., 0] - s * vb[..., 1], s * vb[..., 0] + c * vb[..., 1], vb[..., 2]], -1)
    # a_nav = R (vb_dot + w x vb)
    wx = np.stack([-yaw_rate * vb[..., 1], yaw_rate * vb[..., 0], np.zeros_like(t)], -1)
    ab = vb_dot + wx
    an = np.stack([c * ab[..., 0] - s * ab[..., 1], s * ab[..., 0] + c * ab[..., 1], ab[..., 2]], -1)
    z = -10.0 + 1.0 * (1.0 - np.cos(0.05 * t))  # integral of 0.05 sin(0.05 t), start depth 10 m
    return dict(vb=vb, yaw=yaw, yaw_rate=yaw_rate, vn=vn, an=an, ab=ab, z=z, c=c, s=s)


def pose_inputs(filters, step, dt=0.01, g=None):
    """Sensor samples for global filter indices `filters` (array) at integer `step` (time = step*dt).

    Returns dict: gyro[B,3], acc[B,3], dvl[B,3], pressure[B,1], adcp[B,2].  Noise sigmas per Appendix F.
    """
    filters = np.asarray(filters, dtype=np.uint64)
    B = filters.shape[0]
    tr = truth(step * dt)
    if g is None:
        g = gravity_wgs84(LAT0, ALT0)
    c, s = tr["c"], tr["s"]
    ch = lambda k: np.stack([noise(filters, step, 3 * k + i) for i in range(3)], -1)
    earth_n = np.array([EARTHW * np.cos(LAT0), 0.0, EARTHW * np.sin(LAT0)])
    earth_b = np.array([c * earth_n[0] + s * earth_n[1], -s * earth_n[0] + c * earth_n[1], earth_n[2]])
    gyro = np.array([0.0, 0.0, tr["yaw_rate"]]) + earth_b + 1e-5 * ch(0)
    acc = tr["ab"] + np.array([0.0, 0.0, g]) + 1e-3 * ch(1)
    dvl = tr["vb"] + 0.01 * ch(2)
    pressure = (P_ATM - tr["z"] * g * RHO + 100.0 * noise(filters, step, 40)).reshape(B, 1)
    wv, wvb, w = np.array([0.1, -0.05, 0.0]), np.array([0.05, 0.0, 0.0]), 0.5
    rel = lambda cur: np.array([c * (tr["vn"][0] - cur[0]) + s * (tr["vn"][1] - cur[1]),
                                -s * (tr["vn"][0] - cur[0]) + c * (tr["vn"][1] - cur[1])])
    adcp_model = w * rel(wvb) + (1 - w) * rel(wv)
    adcp = adcp_model + 0.02 * np.stack([noise(filters, step, 50), noise(filters, step, 51)], -1)
    return dict(gyro=np.ascontiguousarray(gyro), acc=np.ascontiguousarray(acc), dvl=np.ascontiguousarray(dvl),
                pressure=pressure, adcp=np.ascontiguousarray(adcp))


R_DVL = 1e-4 * np.eye(3)
R_ACC = 1e-6 * np.eye(3) * 100.0  # accelerometer measurement noise (1e-2 m/s^2)^2
R_PRESSURE = np.array([[1e4]])
R_ADCP = 4e-4 * np.eye(2)


def pose_initial_batch(filters, cfg=None, spread=1.0, with_cov=True):
    """Per-filter initial states: mu0 (+) L0*eps with independent draws (Monte-Carlo ensemble). Returns (mu[B,54], cov[B,53,53])."""
    if cfg is None:
        cfg = pose_config()
    filters = np.asarray(filters, dtype=np.uint64)
    B = filters.shape[0]
    mu0, cov0 = pose_initial(cfg)
    sd = np.sqrt(np.diag(cov0))
    mu = np.tile(mu0, (B, 1))
    eps = np.stack([noise(filters, 0xFFFFFF, 100 + k) for k in range(53)], -1) * spread
    d = eps * sd
    # vector components (skip gravity/inertia/damping/density: they are the Markov offsets of the filter)
    mu[:, 0:3] += d[:, 0:3]
    mu[:, 7:10] += 0.1 * d[:, 6:9]
    mu[:, 10:13] += 0.01 * d[:, 9:12]
    # orientation: q = exp(d_ori)
    v = d[:, 3:6]
    ang = np.linalg.norm(v, axis=1)
    half = 0.5 * ang
    sc = np.where(ang > 1e-12, np.sin(half) / np.maximum(ang, 1e-300), 0.5)
    mu[:, 3:6] = v * sc[:, None]
    mu[:, 6] = np.cos(half)
    cov = np.tile(cov0, (B, 1, 1)) if with_cov else cov0
    return mu, cov


# ---- VelocityUKF / BottomUKF / IndirectPoseUKF scenarios (BASELINE.json configs 1, 2, 5) ---------------------
def uwv_params_flat():
    """flat uwv_dynamic_model::UWVParameters (116) of SYNTH-1: SIMPLE damping, W = B = 1500 N, r_g = r_b = 0."""
    M, Dl, Dq = uwv_matrices()
    p = np.zeros(116)
    p[0:36], p[36:72], p[72:108] = M.ravel(), Dl.ravel(), Dq.ravel()
    p[114] = p[115] = 1500.0
    return p


def velocity_inputs(filters, step, dt=0.01):
    """thrust / gyro / DVL / pressure samples for VelocityUKF (Appendix F)."""
    filters = np.asarray(filters, dtype=np.uint64)
    B = filters.shape[0]
    t = step * dt
    tau = np.array([10 + 20 * np.sin(0.31 * t), 5 * np.sin(0.63 * t), 2 * np.sin(0.2 * t), 0.0, 0.0, 1.0 * np.sin(0.1 * t)])
    efforts = np.tile(tau, (B, 1)) + 0.1 * np.stack([noise(filters, step, 60 + k) for k in range(6)], -1)
    gyro = np.tile(np.array([0.0, 0.0, 0.05 * np.sin(0.02 * t)]), (B, 1)) + 1e-4 * np.stack([noise(filters, step, 70 + k) for k in range(3)], -1)
    tr = truth(t)
    dvl = tr["vb"] + 0.01 * np.stack([noise(filters, step, 80 + k) for k in range(3)], -1)
    pressure = (tr["z"] + 0.05 * noise(filters, step, 90)).reshape(B, 1)
    return dict(efforts=np.ascontiguousarray(efforts), gyro=np.ascontiguousarray(gyro), dvl=np.ascontiguousarray(dvl),
                pressure=np.ascontiguousarray(pressure))


def velocity_initial_batch(filters):
    filters = np.asarray(filters, dtype=np.uint64)
    B = filters.shape[0]
    mu = np.zeros((B, 4))
    mu[:, 0] = 1.0
    mu[:, 0:3] += 0.1 * np.stack([noise(filters, 0xFFFFFE, k) for k in range(3)], -1)
    mu[:, 3] = -10.0 + 0.2 * noise(filters, 0xFFFFFE, 3)
    cov = np.tile(np.diag([0.1, 0.1, 0.1, 1.0]), (B, 1, 1))
    return mu, cov


# DVL beam geometry (Janus, 30 deg) and a sloped sea floor for BottomUKF
BEAMS = np.array([[np.sin(np.pi / 6) * np.cos(a), np.sin(np.pi / 6) * np.sin(a), -np.cos(np.pi / 6)] for a in
                  (np.pi / 4, 3 * np.pi / 4, 5 * np.pi / 4, 7 * np.pi / 4)])
BEAM_ORIGIN = np.array([0.1, 0.0, -0.2])


def bottom_initial_batch(filters):
    filters = np.asarray(filters, dtype=np.uint64)
    B = filters.shape[0]
    mu = np.zeros((B, 4))
    mu[:, 0] = 5.0 + 0.5 * noise(filters, 0xFFFFFD, 0)
    n = np.stack([0.05 * noise(filters, 0xFFFFFD, 1), 0.05 * noise(filters, 0xFFFFFD, 2), np.ones(B)], -1)
    mu[:, 1:4] = n / np.linalg.norm(n, axis=1, keepdims=True)
    cov = np.tile(np.diag([1.0, 0.05, 0.05]), (B, 1, 1))
    return mu, cov


def bottom_inputs(filters, step, dt=0.1):
    filters = np.asarray(filters, dtype=np.uint64)
    B = filters.shape[0]
    t = step * dt
    vel = np.tile(np.array([1.0, 0.2 * np.sin(0.1 * t), -0.05]), (B, 1))
    n_true = np.array([0.08, -0.03, 1.0])
    n_true /= np.linalg.norm(n_true)
    d_true = 5.5 + 0.05 * t
    ranges = np.stack([((np.array([0, 0, -d_true]) - BEAM_ORIGIN) @ n_true) / (BEAMS[k] @ n_true) + 0.02 * noise(filters, step, 110 + k)
                       for k in range(4)], -1)
    nm = n_true + 0.01 * np.stack([noise(filters, step, 120 + k) for k in range(3)], -1)
    nm /= np.linalg.norm(nm, axis=1, keepdims=True)
    return dict(velocity=np.ascontiguousarray(vel), ranges=np.ascontiguousarray(ranges), normal=np.ascontiguousarray(nm))


CAMERA = np.array([800.0, 800.0, 320.0, 240.0])
CAM_IN_BODY = np.array([0.3, 0.0, -0.1, 0.5, -0.5, 0.5, -0.5])      # camera z axis along body x
MARKER_POSE = np.array([4.0, 0.3, -0.2, 0.0, 0.0, 0.0, 1.0])
MARKER_FEATURES = 0.1 * np.array([[0.0, -1.0, -1.0], [0.0, 1.0, -1.0], [0.0, 1.0, 1.0], [0.0, -1.0, 1.0]])


def indirect_inputs(filters, step, dt=0.01):
    """pose reference (dead-reckoned body pose with a slowly growing offset) and marker corner pixels."""
    filters = np.asarray(filters, dtype=np.uint64)
    B = filters.shape[0]
    t = step * dt
    yaw = 0.05 * np.sin(0.3 * t)
    ref = np.tile(np.array([0.2 * t, 0.05 * np.sin(t), 0.0, 0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2)]), (B, 1))
    # true body pose = reference shifted by a constant error (what the filter should estimate)
    err_t = np.array([0.15, -0.1, 0.05])
    c, s = np.cos(yaw), np.sin(yaw)
    Rb = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    body_t = ref[0, 0:3] + Rb @ err_t
    Rc = np.array([[0, 0, 1.0], [-1.0, 0, 0], [0, -1.0, 0]])         # rotation of CAM_IN_BODY quaternion
    cam_t = body_t + Rb @ CAM_IN_BODY[0:3]
    Rcn = Rb @ Rc
    px = np.zeros((B, 4, 2))
    for k in range(4):
        pw = MARKER_POSE[0:3] + MARKER_FEATURES[k]
        pc = Rcn.T @ (pw - cam_t)
        uv = np.array([CAMERA[0] * pc[0] / pc[2] + CAMERA[2], CAMERA[1] * pc[1] / pc[2] + CAMERA[3]])
        px[:, k, :] = uv + 0.5 * np.stack([noise(filters, step, 130 + 2 * k), noise(filters, step, 131 + 2 * k)], -1)
    return dict(pose_ref=np.ascontiguousarray(ref), px=np.ascontiguousarray(px))
