"""CPU restatement of the scoring code in the reference's application scripts.  TEST INFRASTRUCTURE.

Everything here works from the stored samples, like the scripts do (the engine's evaluator,
functionalmf_b200/metrics.py, never stores them).  Pinned by tests/golden/metrics_cases.npz, whose
expected values come from the reference's own expressions (oracle/make_golden_metrics.py evaluates
the statements of politics/benchmark.py, flutrends/benchmark.py and
examples/poisson_tensor_filtering.py in place).
"""
import numpy as np
from scipy.stats import norm, poisson


def ilogit(x):
    return 1.0 / (1.0 + np.exp(-x))


def surface(Ws, Vs, transform='identity', Rs=None):
    """[S, N, M, T] surface per saved sample.
    identity: flutrends/benchmark.py:49; ilogit: examples/binomial_tensor_filtering.py;
    nb_mean: politics/benchmark.py:155-156."""
    psi = np.einsum('znk,zmtk->znmt', Ws, Vs)
    if transform == 'identity':
        return psi
    if transform == 'ilogit':
        return ilogit(psi)
    if transform == 'nb_mean':
        Ps = ilogit(psi.clip(-10, 10))
        return Rs * Ps / (1 - Ps)
    raise ValueError(transform)


def split(Y, Y_train):
    """politics/benchmark.py:164-166, flutrends/benchmark.py:125-127."""
    is_missing = np.isnan(Y)
    is_held_out = (~is_missing) & np.isnan(Y_train)
    is_in_sample = (~is_missing) & (~is_held_out)
    return is_in_sample, is_held_out


def per_sample_scores(Y, mask, mu, loglik=None, nu2s=None):
    """politics/benchmark.py:168-178: the mean over samples of the per-sample RMSE, MAE and mean
    log-likelihood over the cells in ``mask``."""
    err = Y[None, mask] - mu[:, mask]
    out = dict(rmse=np.sqrt(np.mean(err ** 2, axis=-1)).mean(), mae=np.mean(np.abs(err), axis=-1).mean())
    if loglik == 'poisson':
        out['ll'] = poisson.logpmf(Y[None, mask], mu[:, mask]).mean(axis=-1).mean()
    elif loglik == 'gaussian':
        out['ll'] = norm.logpdf(Y[None, mask], mu[:, mask], np.sqrt(nu2s)[:, None]).mean(axis=-1).mean()
    return out


def mean_scores(Y, mask, mu_mean, loglik=None):
    """flutrends/benchmark.py:135-141 (RMSE / MAE of the posterior-mean surface) and
    examples/poisson_tensor_filtering.py:166-168 (NLL at the posterior mean)."""
    err = Y[mask] - mu_mean[mask]
    out = dict(rmse=np.sqrt(np.mean(err ** 2)), mae=np.mean(np.abs(err)))
    if loglik == 'poisson':
        out['nll'] = -np.sum(poisson.logpmf(Y[mask], mu_mean[mask]))
    return out


def coverage_at(truth, samples, interval, mask=None):
    """examples/poisson_tensor_filtering.py:20-23."""
    lower = np.percentile(samples, (100 - interval) / 2, axis=0)
    upper = np.percentile(samples, (100 - interval) / 2 + interval, axis=0)
    inside = (truth >= lower) & (truth <= upper)
    if mask is not None:
        inside = inside[mask]
    return np.mean(inside) * 100


def predictive_cdf(Y, mu, nu2s):
    """Mixture cdf of the Gaussian posterior predictive at Y: mean_s Phi((y - mu_s) / sqrt(nu2_s)).
    flutrends/benchmark.py:68-75 estimates the 2.5 / 97.5 % points of this mixture from
    100 x nsamples normal draws per cell; `Y < Y_lower` there is `cdf < 0.025` here."""
    sd = np.sqrt(np.asarray(nu2s, dtype=float)).reshape((-1,) + (1,) * (mu.ndim - 1))
    return norm.cdf((Y[None] - mu) / sd).mean(axis=0)


def predictive_coverage(Y, mask, mu, nu2s, interval=95):
    lo = (100 - interval) / 200
    F = predictive_cdf(Y, mu, nu2s)[mask]
    return 100 - ((F < lo) | (F > 1 - lo)).mean() * 100


def predictive_band_mc(mu, nu2s, ndraws=100):
    """flutrends/benchmark.py:66-75 as written (row-by-row Monte-Carlo band, global np.random)."""
    S, N, M, T = mu.shape
    Y_lower, Y_upper = np.zeros((N, M, T)), np.zeros((N, M, T))
    for i in range(N):
        for k in range(T):
            Y_samples_ik = np.random.normal(mu[:, i, 0, k], np.sqrt(nu2s[:, 0]), size=(ndraws, S))
            Y_upper[i, 0, k] = np.percentile(Y_samples_ik, 97.5)
            Y_lower[i, 0, k] = np.percentile(Y_samples_ik, 2.5)
    return Y_lower, Y_upper
