"""Parity metrics used by the GPU tests (tolerances of BASELINE.json's north_star)."""
import numpy as np

TOL_STATE = 1e-9   # max |d mu| / max(1, |mu|_max)  per step
TOL_COV = 1e-8     # max |d Sigma| / |Sigma|_max    per step


def state_err(mu, ref):
    mu, ref = np.asarray(mu), np.asarray(ref)
    return float(np.abs(mu - ref).max() / max(1.0, np.abs(ref).max()))


def cov_err(cov, ref):
    return float(np.abs(cov - ref).max() / np.abs(ref).max())


def cov_err_scaled(cov, ref):
    """|dSigma_ij| / sqrt(Sigma_ii Sigma_jj): a far stricter, unit-free measure (reported)."""
    sd = np.sqrt(np.diag(ref))
    return float((np.abs(cov - ref) / np.outer(sd, sd)).max())
