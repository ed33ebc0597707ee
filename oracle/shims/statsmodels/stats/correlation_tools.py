"""`cov_nearest(method='clipped', threshold=1e-15)` restated from statsmodels' published behaviour:
covariance -> correlation, symmetric eigendecomposition, eigenvalues clipped from below at the
threshold (only if any is below it), diagonal renormalised to one, back to covariance.
PARITY UNPINNED: the reference's tests pin only the matrix shape at this boundary.
"""
import numpy as np


def cov_nearest(cov, method='clipped', threshold=1e-15, n_fact=100, return_all=False):
    if method != 'clipped':
        raise NotImplementedError(method)
    cov = np.asarray(cov)
    std = np.sqrt(np.diag(cov))
    corr = cov / np.outer(std, std)
    evals, evecs = np.linalg.eigh(corr)
    if np.any(evals < threshold):
        x_new = np.dot(evecs * np.maximum(evals, threshold), evecs.T)
        x_std = np.sqrt(np.diag(x_new))
        corr = x_new / x_std / x_std[:, None]
    cov_ = corr * np.outer(std, std)
    if return_all:
        return cov_, corr, std
    return cov_
