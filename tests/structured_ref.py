"""numpy statement of the prefix-structured unscented transform used by the CUDA engine.

It is NOT the oracle: the oracle (oracle/) executes the ukfom recipe literally over all 2n+1 sigma points.
This file states the algebraically equivalent O(n^2) evaluation the kernels implement, so that the
equivalence can be tested on the CPU (tests/test_structured_equivalence.py) independent of any GPU.
"""
import numpy as np

N = 53
EARTHW = 7.292115e-5
MTK_TOL = 1e-11


def cos_sinc_sqrt(x2):
    if x2 >= (2.220446049250313e-16) ** 0.25:
        x = np.sqrt(x2)
        return np.cos(x), np.sin(x) / x
    inv = [1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.]
    c, s, term = 1., 1., -0.5 * x2
    for i in range(3):
        c += term
        term *= inv[2 * i]
        s += term
        term *= -inv[2 * i + 1] * x2
    return c, s


def qmul(a, b):
    return np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                     a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                     a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0],
                     a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])


def qconj(a):
    return np.array([-a[0], -a[1], -a[2], a[3]])


def qrot(q, v):
    u = q[:3]
    t = 2 * np.cross(u, v)
    return v + q[3] * t + np.cross(u, t)


def qmat(q):
    return np.stack([qrot(q, e) for e in np.eye(3)], 1)


def so3_exp(v, scale=1.0):
    h = scale / 2
    c, s = cos_sinc_sqrt(h * h * float(v @ v))
    return np.array([s * h * v[0], s * h * v[1], s * h * v[2], c])


def so3_log(q):
    nv = np.linalg.norm(q[:3])
    if nv < MTK_TOL:
        nv = MTK_TOL
    return 2.0 / nv * np.arctan(nv / q[3]) * q[:3]


def bplus(q, v, scale=1.0):
    return qmul(q, so3_exp(v, scale))


def bminus(a, b):
    return so3_log(qmul(qconj(b), a))


def iterated_mean_so3(qs, wts, ntot, tol, max_it):
    """equal-weight iterated mean over ntot logical points given distinct points qs with multiplicities wts."""
    r = qs[0].copy()
    it = 0
    while True:
        m = sum(w * bminus(q, r) for q, w in zip(qs, wts)) / ntot
        r = bplus(r, m)
        it += 1
        if not (np.linalg.norm(m) > tol and it < max_it):
            break
    return r


def chol_leading(S, k):
    return np.linalg.cholesky(S[:k, :k])


def tangent_to_stored(t):
    return t if t < 3 else t + 1  # orientation takes 4 stored slots at 3..6


def pose_predict(mu, S, Q, w_meas, par, offsets, dt, nav_lat, tol=1e-9, max_it=100):
    """Structured PoseUKF predict. nav_lat(x, y) -> latitude.  par: flat PoseUKFParameter (22)."""
    k = 15
    L = chol_leading(S, k)
    p, q, v, a, bg = mu[0:3], mu[3:7], mu[7:10], mu[10:13], mu[13:16]

    def prop(d):  # d = tangent perturbation of the first 15 coordinates
        pp, qq, bb = p + d[0:3], bplus(q, d[3:6]), bg + d[12:15]
        lat = nav_lat(pp[0], pp[1])
        earth = np.array([EARTHW * np.cos(lat), 0., EARTHW * np.sin(lat)])
        wn = qrot(qq, w_meas - bb) - earth
        return bplus(qq, wn, dt)

    z = np.zeros(k)
    pts = [prop(z)] + [prop(s * L[:, j]) for j in range(k) for s in (+1, -1)]
    wts = [1 + 2 * (N - k)] + [1] * (2 * k)
    r = iterated_mean_so3(pts, wts, 2 * N + 1, tol, max_it)
    dev = [bminus(x, r) for x in pts]
    Soo = 0.5 * sum(w * np.outer(d, d) for d, w in zip(dev, wts))
    A = np.stack([0.5 * (dev[1 + 2 * j] - dev[2 + 2 * j]) for j in range(k)], 1)  # 3 x k
    G = np.linalg.solve(L.T, A.T).T  # A L^-1
    T = G @ S[:k, :]  # 3 x 53
    # affine part: F = D + dt*E
    f = np.ones(N)
    f[12:15] = 1 - dt / par[2]
    f[15:18] = 1 - dt / par[6]
    f[19:28] = 1 - dt / par[10]
    f[28:37] = 1 - dt / par[11]
    f[37:46] = 1 - dt / par[12]
    f[46:50] = 1 - dt / par[13]
    f[50:52] = 1 - dt / par[16]
    f[52] = 1 - dt / par[18]
    F = np.diag(f)
    for i in range(3):
        F[i, 6 + i] = dt
        F[6 + i, 9 + i] = dt
    F[3:6, :] = 0
    Sn = F @ S @ F.T
    Tf = T @ F.T
    Sn[3:6, :] = Tf
    Sn[:, 3:6] = Tf.T
    Sn[3:6, 3:6] = Soo
    # effective process noise (reference src/PoseUKF.cpp:448-460)
    Qe = Q.copy()
    rot = qmat(q)
    Qe[3:6, 3:6] = rot @ Q[3:6, 3:6] @ rot.T
    sv = np.array([v[0], v[1], 10 * v[2]])
    add = par[15] * float(sv @ sv) * dt
    for i in (46, 47, 48, 49):
        Qe[i, i] += add
    Sn = Sn + dt ** 2 * Qe
    # mean
    mn = mu.copy()
    mn[0:3] = p + dt * v
    mn[3:7] = r
    mn[7:10] = v + dt * a
    mn[13:16] = bg + dt * ((-1.0 / par[2]) * (bg - par[3:6]))
    ba = mu[16:19]
    mn[16:19] = ba + dt * ((-1.0 / par[6]) * (ba - par[7:10]))
    for off, tau, o0 in ((20, par[10], 0), (29, par[11], 9), (38, par[12], 18)):
        x = mu[off:off + 9]
        mn[off:off + 9] = x + dt * ((-1.0 / tau) * (x - offsets[o0:o0 + 9]))
    mn[47:51] = mu[47:51] + dt * ((-1.0 / par[13]) * mu[47:51])
    mn[51:53] = mu[51:53] + dt * ((-1.0 / par[16]) * mu[51:53])
    mn[53] = mu[53] + dt * ((-1.0 / par[18]) * (mu[53] - offsets[27]))
    return mn, Sn


def apply_delta(mu, S, delta, tol=1e-9, max_it=100):
    k = 6
    L = chol_leading(S, k)
    q = mu[3:7]
    pts = [bplus(q, delta[3:6])] + [bplus(q, delta[3:6] + s * L[3:6, j]) for j in range(k) for s in (+1, -1)]
    wts = [1 + 2 * (N - k)] + [1] * (2 * k)
    r = iterated_mean_so3(pts, wts, 2 * N + 1, tol, max_it)
    dev = [bminus(x, r) for x in pts]
    Soo = 0.5 * sum(w * np.outer(d, d) for d, w in zip(dev, wts))
    A = np.stack([0.5 * (dev[1 + 2 * j] - dev[2 + 2 * j]) for j in range(k)], 1)
    G = np.linalg.solve(L.T, A.T).T
    T = G @ S[:k, :]
    Sn = S.copy()
    Sn[3:6, :] = T
    Sn[:, 3:6] = T.T
    Sn[3:6, 3:6] = Soo
    mn = mu.copy()
    mn[0:3] += delta[0:3]
    mn[3:7] = r
    mn[7:] += delta[6:]
    return mn, Sn


def pose_update(mu, S, z, R, h, k, gate=None, tol=1e-9, max_it=100):
    """Structured update for a vector measurement whose model h(mu, d) depends on the first k tangent coordinates.
    h(mu, d): evaluate the measurement at mu (+) d where d is a length-k perturbation of the leading coordinates."""
    L = chol_leading(S, k)
    zero = np.zeros(k)
    Z = [h(mu, zero)] + [h(mu, s * L[:, j]) for j in range(k) for s in (+1, -1)]
    wts = [1 + 2 * (N - k)] + [1] * (2 * k)
    ntot = 2 * N + 1
    zb = Z[0].copy()
    it = 0
    while True:
        m = sum(w * (x - zb) for x, w in zip(Z, wts)) / ntot
        zb = zb + m
        it += 1
        if not (np.linalg.norm(m) > tol and it < max_it):
            break
    Szz = 0.5 * sum(w * np.outer(x - zb, x - zb) for x, w in zip(Z, wts)) + R
    Bm = np.stack([0.5 * (Z[1 + 2 * j] - Z[2 + 2 * j]) for j in range(k)], 1)  # m x k
    G = np.linalg.solve(L.T, Bm.T).T  # m x k
    Pxz = (G @ S[:k, :]).T  # 53 x m
    Si = np.linalg.inv(Szz)
    K = Pxz @ Si
    nu = z - zb
    d2 = float(nu @ Si @ nu)
    if gate is not None and not gate(d2):
        return mu.copy(), S.copy(), False
    Sd = S - K @ Szz @ K.T
    mn, Sn = apply_delta(mu, Sd, K @ nu, tol, max_it)
    return mn, Sn, True


def h_velocity(mu, d):
    q = bplus(mu[3:7], d[3:6])
    v = mu[7:10] + d[6:9]
    return qrot(qconj(q), v)


def pose_update_tail(mu, S, z, R, h, k, tail, tail_jac, gate=None, tol=1e-9, max_it=100):
    """Structured update with an AFFINE TAIL: h reads the first k tangent coordinates arbitrarily
    and the tangent coordinates `tail` (all >= k) affinely once the prefix is fixed.  Sigma columns j >= k then move only
    the tail, Z_j(+-) = Z_0 +- C L[tail, j] with C = tail_jac(mu) (m x r), so their sums collapse to Schur complements:
      S_zz  = 1/2 sum_{centre, j<k} w (Z - zb)(Z - zb)^T + C (Sigma_TT - L_T L_T^T) C^T + R
      P_xz  = Sigma[:, :k] L11^-T (B - L_T^T C^T) + Sigma[:, T] C^T,        L_T = Sigma[T, :k] L11^-T.
    h(mu, d, dt): measurement at mu (+) (prefix perturbation d, tail perturbation dt)."""
    tail = list(tail)
    L = chol_leading(S, k)
    LT = np.linalg.solve(L, S[:k, tail]).T  # r x k  (rows `tail` of the full-height panel factor)
    zero, zt = np.zeros(k), np.zeros(len(tail))
    Z = [h(mu, zero, zt)] + [h(mu, s * L[:, j], s * LT[:, j]) for j in range(k) for s in (+1, -1)]
    wts = [1 + 2 * (N - k)] + [1] * (2 * k)
    ntot = 2 * N + 1
    zb = Z[0].copy()
    it = 0
    while True:
        m = sum(w * (x - zb) for x, w in zip(Z, wts)) / ntot
        zb = zb + m
        it += 1
        if not (np.linalg.norm(m) > tol and it < max_it):
            break
    C = tail_jac(mu)
    U = C @ LT  # m x k
    W = C @ S[np.ix_(tail, tail)] @ C.T - U @ U.T
    Szz = 0.5 * sum(w * np.outer(x - zb, x - zb) for x, w in zip(Z, wts)) + W + R
    Bm = np.stack([0.5 * (Z[1 + 2 * j] - Z[2 + 2 * j]) for j in range(k)], 1)  # m x k
    G = np.linalg.solve(L.T, (Bm - U).T).T
    Pxz = (G @ S[:k, :] + C @ S[tail, :]).T
    Si = np.linalg.inv(Szz)
    K = Pxz @ Si
    nu = z - zb
    d2 = float(nu @ Si @ nu)
    if gate is not None and not gate(d2):
        return mu.copy(), S.copy(), False
    Sd = S - K @ Szz @ K.T
    mn, Sn = apply_delta(mu, Sd, K @ nu, tol, max_it)
    return mn, Sn, True


def make_h_pressure(atm_pressure, sensor_in_imu):
    """reference src/PoseUKF.cpp:107-115: prefix 19 (z, q, gravity), tail = water density (52)."""
    def h(mu, d, dt):
        q = bplus(mu[3:7], d[3:6])
        zn = (mu[2] + d[2]) + qrot(q, sensor_in_imu)[2]
        return np.array([atm_pressure - zn * (mu[19] + d[18]) * (mu[53] + dt[0])])

    def jac(mu):
        zn = mu[2] + qrot(mu[3:7], sensor_in_imu)[2]
        return np.array([[-(zn * mu[19])]])
    return h, jac


def make_h_adcp(wgt):
    """reference src/PoseUKF.cpp:133-151: prefix 9 (q, v), tail = water velocities and ADCP bias (46..51)."""
    def h(mu, d, dt):
        q = bplus(mu[3:7], d[3:6])
        v = mu[7:10] + d[6:9]
        wv, wvb, b = mu[47:49] + dt[0:2], mu[49:51] + dt[2:4], mu[51:53] + dt[4:6]
        qi = qconj(q)
        below = qrot(qi, v - np.array([wvb[0], wvb[1], 0.0]))
        around = qrot(qi, v - np.array([wv[0], wv[1], 0.0]))
        return wgt * below[:2] + (1 - wgt) * around[:2] + b

    def jac(mu):
        Rt = qmat(mu[3:7]).T[:2, :2]
        return np.concatenate([-(1 - wgt) * Rt, -wgt * Rt, np.eye(2)], axis=1)
    return h, jac
