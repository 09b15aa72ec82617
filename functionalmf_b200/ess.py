"""Elliptical slice sampling (Murray, Adams & MacKay 2010), host side.

Restatement of functionalmf/elliptical_slice.py:59-124 for the case used by the non-conjugate
tensor-filtering model: the caller supplies a draw ``nu`` from the zero-mean Gaussian prior (here
it comes from the CUDA engine) and a black-box log-likelihood.  ``rng.rand()`` is consumed in the
reference's order: slice height, first angle, then one angle per shrink step.
"""
import math
import warnings
import numpy as np


def elliptical_slice(x, nu, log_like_fn, cur_log_like=None, ll_args=None, mu=None, rng=np.random):
    x = np.asarray(x, dtype=float)
    nu = np.reshape(nu, x.shape)
    mu = np.zeros(x.size) if mu is None else np.asarray(mu, dtype=float)
    if cur_log_like is None:
        cur_log_like = log_like_fn(x, ll_args)
    height = np.log(rng.rand()) + cur_log_like
    phi = rng.rand() * 2 * math.pi
    phi_min, phi_max = phi - 2 * math.pi, phi
    while True:
        prop = (x - mu) * np.cos(phi) + nu * np.sin(phi) + mu
        ll = log_like_fn(prop, ll_args)
        if ll >= height:
            return prop, ll
        if phi > 0:
            phi_max = phi
        elif phi < 0:
            phi_min = phi
        else:
            warnings.warn('elliptical slice shrunk to the current point and was still rejected')
            return prop, ll
        phi = rng.rand() * (phi_max - phi_min) + phi_min
