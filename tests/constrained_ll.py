"""Black-box likelihood shared by the constrained-model fixtures and tests
(the Poisson log-likelihood of examples/poisson_tensor_filtering.py:26-37)."""
import warnings
import numpy as np
from scipy.stats import poisson


def rowcol_loglikelihood(Y, WV, W, V, row=None, col=None):
    if row is not None:
        Y = Y[row]
    if col is not None:
        Y = Y[:, col]
    if len(Y.shape) > len(WV.shape):
        WV = WV[..., None]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=RuntimeWarning)
        return np.nansum(poisson.logpmf(Y, WV))


class ReplayRng(object):
    """Replays recorded np.random.random / np.random.choice results in order."""

    def __init__(self, uniforms, choices):
        self.u = list(uniforms)
        self.c = list(choices)

    def random(self, size=None):
        return self.u.pop(0)

    def choice(self, a, size=None, replace=True):
        return self.c.pop(0)
