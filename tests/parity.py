"""Parity metrics used by the GPU tests (tolerances of BASELINE.json's north_star).

Three gates per comparison:
  * the north_star metrics: max |d mu| / max(1, |mu|_max) <= 1e-9 and max |d Sigma| / |Sigma|_max <= 1e-8 per step.  They are
    normalised by the LARGEST entry (water density ~1025, acceleration variance ~10), so on their own they would let a small
    block (gyro bias 1e-5, gravity variance 2.5e-3) be entirely wrong;
  * therefore also, per component, |d mu_i| <= TOL_STATE_COMP * max(|mu_i|, sigma_i) (sigma_i = sqrt(Sigma_ii) of the reference;
    unit-quaternion components are scaled by 1), and
  * the unit-free covariance error |d Sigma_ij| / sqrt(Sigma_ii Sigma_jj) <= TOL_COV_SCALED.
"""
import numpy as np

TOL_STATE = 1e-9        # max |d mu| / max(1, |mu|_max)  per step
TOL_COV = 1e-8          # max |d Sigma| / |Sigma|_max    per step
TOL_STATE_COMP = 1e-9   # max_i |d mu_i| / max(|mu_i|, sigma_i)  per step
TOL_COV_SCALED = 1e-8   # max_ij |d Sigma_ij| / sqrt(Sigma_ii Sigma_jj)  per step


def state_err(mu, ref):
    mu, ref = np.asarray(mu), np.asarray(ref)
    return float(np.abs(mu - ref).max() / max(1.0, np.abs(ref).max()))


def cov_err(cov, ref):
    return float(np.abs(cov - ref).max() / np.abs(ref).max())


def cov_err_scaled(cov, ref):
    """|dSigma_ij| / sqrt(Sigma_ii Sigma_jj): unit-free, every block counts the same."""
    sd = np.sqrt(np.diag(ref))
    sd = np.where(sd > 0, sd, 1.0)
    return float((np.abs(cov - ref) / np.outer(sd, sd)).max())


def stored_sigma(cov, quat_at=None):
    """Standard deviation per STORED state component from the tangent-space covariance: a unit quaternion stored at
    [quat_at, quat_at + 4) replaces the three tangent coordinates [quat_at, quat_at + 3) and is scaled by 1."""
    sd = np.sqrt(np.maximum(np.diag(cov), 0.0))
    if quat_at is None:
        return sd
    return np.concatenate([sd[:quat_at], np.ones(4), sd[quat_at + 3:]])


def state_err_comp(mu, ref, sd):
    """max_i |d mu_i| / max(|mu_i|, sd_i): every component against its own magnitude or uncertainty."""
    mu, ref = np.asarray(mu), np.asarray(ref)
    scale = np.maximum(np.abs(ref), sd)
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(mu - ref) / scale).max())


# where the unit quaternion sits in the stored state of each filter (None: plain vector state, S2 handled by the caller)
QUAT_AT = {54: 3, 7: 3}


def all_errs(mu, cov, ref_mu, ref_cov):
    """(state, cov, state per component, cov scaled) for one filter."""
    mu, ref_mu = np.asarray(mu), np.asarray(ref_mu)
    n, dof = ref_mu.shape[0], ref_cov.shape[0]
    if n == dof:
        sd = stored_sigma(ref_cov)
    elif n in QUAT_AT and n == dof + 1:
        sd = stored_sigma(ref_cov, QUAT_AT[n])
    else:   # BottomUKF: [distance | unit normal (3)] with tangent [distance | S2 (2)]
        sd = np.concatenate([np.sqrt(np.maximum(np.diag(ref_cov)[: n - 3], 0.0)), np.ones(3)])
    return (state_err(mu, ref_mu), cov_err(cov, ref_cov), state_err_comp(mu, ref_mu, sd), cov_err_scaled(cov, ref_cov))


def check_all(mu, cov, ref_mu, ref_cov, what="", scale=1.0):
    """Asserts the four gates for a batch ([B, ...] arrays); returns the worst (state, cov, state_comp, cov_scaled).
    scale > 1 loosens every gate by that factor (free-running comparisons over many steps)."""
    worst = np.zeros(4)
    for f in range(len(mu)):
        worst = np.maximum(worst, all_errs(mu[f], cov[f], ref_mu[f], ref_cov[f]))
    print(f"[parity] {what}: state {worst[0]:.2e}  cov {worst[1]:.2e}  state(comp) {worst[2]:.2e}  cov(scaled) {worst[3]:.2e}")
    tol = np.array([TOL_STATE, TOL_COV, TOL_STATE_COMP, TOL_COV_SCALED]) * scale
    names = ("state", "covariance", "per-component state", "scaled covariance")
    for k in range(4):
        assert worst[k] <= tol[k], f"{what}: {names[k]} error {worst[k]:.3e} > {tol[k]:.1e}"
    return tuple(float(x) for x in worst)
