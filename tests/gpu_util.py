"""Shared helpers for the -m gpu parity tests (all calls go through the C ABI)."""
import numpy as np


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def band_rows_from_lower(ab):
    """LAPACK lower band ab[d, c] = Q[c+d, c]  ->  row-band R[j, kd-d] = Q[j, j-d]."""
    kd, n = ab.shape[0] - 1, ab.shape[1]
    out = np.zeros((n, kd + 1))
    for d in range(kd + 1):
        out[d:, kd - d] = ab[d, :n - d]
    return out


def band_rows_from_dense(Q, kd):
    n = Q.shape[0]
    out = np.zeros((n, kd + 1))
    for d in range(kd + 1):
        idx = np.arange(d, n)
        out[idx, kd - d] = Q[idx, idx - d]
    return out


def load_state(eng, st, gaussian=True):
    for k in ('W', 'V', 'Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
        eng.set(k, st[k])
    for k in ('lam2', 'lam2_a', 'sigma2'):
        eng.set(k, [st[k]])
    if gaussian:
        eng.set('nu2', [float(np.ravel(st['nu2'])[0])])


def inject_all(eng, noise):
    for k, v in noise.items():
        eng.inject(k, np.atleast_1d(v))
