"""Generalized analytic slice sampling (GASS) under linear inequality constraints.

Host-side restatement of functionalmf/gass.py:13-130 (Tansey & Tosh) for the constrained
non-conjugate model.  Differences in form only: the elliptical proposal ``v`` is passed in
(in this package it comes from the CUDA engine's Cholesky draw) and the three sources of
randomness are explicit (``rng`` must offer ``random()`` and ``choice()`` like ``np.random``).

The state moves on the ellipse  x(theta) = (x - mu) cos(theta) + v sin(theta) + mu.  A row
(A_r, c_r) of ``Constraints`` demands  A_r x >= c_r, i.e.  a cos(theta) + b sin(theta) >= c
with a = A_r (x - mu), b = A_r v, c = c_r - A_r mu.  The feasible angles are intersected on a
fine grid, thinned to ``ngrid`` candidates, the black-box log-likelihood is evaluated on all of
them, and the new state is drawn uniformly from the candidates above the slice height.
"""
import numpy as np

FINE_GRID = 10000
EDGE_EPS = 1e-6


def feasible_angles(a, b, c, ngrid):
    """Grid of angles in [-pi, pi] satisfying every a cos + b sin >= c (gass.py:37-80)."""
    disc = a ** 2 + b ** 2 - c ** 2
    # the whole ellipse is feasible for a row when the discriminant is negative or a == -c
    active = (disc >= 0) & (a != -c)
    if not np.any(active):
        return np.linspace(-np.pi, np.pi, ngrid)
    root = np.sqrt(disc[active])
    denom = (a + c)[active]
    th1 = 2 * np.arctan((b[active] + root) / denom)
    th2 = 2 * np.arctan((b[active] - root) / denom)
    outside = a[active] ** 2 < c[active] ** 2       # feasible set is the complement of [min, max]
    grid = np.linspace(-np.pi, np.pi, FINE_GRID)
    for t1, t2 in zip(th1[outside], th2[outside]):
        grid = grid[(grid <= min(t1, t2)) | (grid >= max(t1, t2))]
    if np.any(~outside):
        lo = np.minimum(th1[~outside], th2[~outside]).max() + EDGE_EPS
        hi = np.maximum(th1[~outside], th2[~outside]).min() - EDGE_EPS
        grid = grid[(grid >= lo) & (grid <= hi)]
    return grid


def gass(x, v, loglikelihood, Constraints, mu=None, cur_ll=None, ll_args=None, ngrid=100, rng=np.random):
    """One GASS transition.  Returns (x_new, loglik_new).

    Order of random draws as in the reference: slice height, (proposal, drawn by the caller),
    grid thinning, final choice -- so a recorded tape replays identically."""
    if cur_ll is None:
        cur_ll = loglikelihood(x, ll_args)
    height = cur_ll + np.log(rng.random())
    if mu is None:
        mu = np.zeros_like(x)
    A, cvec = Constraints[:, :-1], Constraints[:, -1]
    assert Constraints.shape[1] == mu.shape[0] + 1
    assert np.all(A.dot(x) >= cvec), 'Invalid starting point!\n{}\nConstraints:\n{}'.format(x, (A.dot(x) - cvec).min())
    x0 = x - mu
    grid = feasible_angles(A.dot(x0), A.dot(v), cvec - A.dot(mu), ngrid)
    if len(grid) == 0:
        return x, cur_ll
    if len(grid) > ngrid:
        grid = rng.choice(grid, size=ngrid, replace=False)
    cand = x0[None] * np.cos(grid[:, None]) + v[None] * np.sin(grid[:, None]) + mu[None]
    cand_ll = loglikelihood(cand, ll_args)            # must support a batch of candidates
    keep = cand_ll >= height
    cand, cand_ll = cand[keep], cand_ll[keep]
    if len(cand) == 0:
        return x, cur_ll
    pick = rng.choice(len(cand))
    return cand[pick], cand_ll[pick]
