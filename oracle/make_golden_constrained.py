#!/usr/bin/env python
"""Fixtures for the constrained (GASS) model from the UNMODIFIED reference.  TEST INFRASTRUCTURE;
runs only in the build container (needs /root/reference).

The reference's per-row / per-column workers (factor.py:665-709, 759-845) are called in-process on
a context object (its multiprocessing=False branch is broken and its Pool branch needs real shared
memory), with a recording tape over np.random.random / normal / choice.  Writes
tests/golden/constrained_{plain,ep}.npz and tests/golden/gass_cases.npz.
"""
import os
import sys
import warnings
import numpy as np
from scipy.sparse import coo_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
import sksparse.cholmod as shim_chol           # noqa: E402
import functionalmf.factor as F                # noqa: E402
import functionalmf.gass as G                  # noqa: E402
from functionalmf.utils import bayes_grid_penalty   # noqa: E402
from constrained_ll import rowcol_loglikelihood     # noqa: E402


class Tape(object):
    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)
        self.u, self.z, self.c = [], [], []

    def random(self, size=None):
        v = self.rs.random_sample(size)
        self.u.append(v)
        return v

    def normal(self, loc=0.0, scale=1.0, size=None):
        z = self.rs.standard_normal(size)
        self.z.append(np.array(z, dtype=float))
        return loc + scale * z

    def choice(self, a, size=None, replace=True):
        v = self.rs.choice(a, size=size, replace=replace)
        self.c.append(np.array(v))
        return v

    def __enter__(self):
        self._o = (np.random.random, np.random.normal, np.random.choice)
        np.random.random, np.random.normal, np.random.choice = self.random, self.normal, self.choice
        return self

    def __exit__(self, *a):
        np.random.random, np.random.normal, np.random.choice = self._o


class Ctx(object):
    pass


def pack_choices(cs):
    """Ragged list of choice results -> (flat values, lengths); scalars have length 0."""
    flat, lens = [], []
    for c in cs:
        c = np.asarray(c, dtype=float)
        if c.ndim == 0:
            lens.append(0); flat.append(c.reshape(1))
        else:
            lens.append(len(c)); flat.append(c)
    return np.concatenate(flat) if flat else np.zeros(0), np.array(lens)


def make_model_case(name, with_ep, seed):
    rs = np.random.RandomState(seed)
    N, M, T, K, order = 6, 5, 8, 3, 1
    shim_chol.set_layout(K, T)
    W = rs.gamma(2.0, 0.5, size=(N, K)); W[np.triu_indices(K, k=1)] = 0
    V = rs.gamma(2.0, 0.5, size=(M, T, K))
    Mu = np.einsum('nk,mtk->nmt', W, V)
    Y = rs.poisson(Mu).astype(float)
    Y[rs.random_sample(Y.shape) < 0.1] = np.nan
    Constraints = np.concatenate([np.eye(T), np.zeros((T, 1))], axis=1)      # positive means
    Delta = bayes_grid_penalty(T, order)
    ctx = Ctx()
    ctx.W, ctx.V = W.copy(), V.copy()
    ctx.Tau2 = rs.gamma(2.0, 1.0, size=(M, Delta.shape[0])) + 0.05
    ctx.sigma2, ctx.lam2, ctx.stability = 0.8, 0.3, 1e-6
    ctx.Constraints_A, ctx.Constraints_C = Constraints[:, :-1], Constraints[:, -1:]
    ctx.Delta = coo_matrix(Delta)
    ctx.Row_constraints = None
    if with_ep:
        ctx.Mu_ep = Mu * (1 + 0.05 * rs.normal(size=Mu.shape))
        ctx.Sigma_ep = 0.5 + 0.1 * rs.random_sample(Mu.shape)
    else:
        ctx.Mu_ep, ctx.Sigma_ep = None, None
    ctx.nrows, ctx.ncols, ctx.ndepth, ctx.nembeds = N, M, T, K
    ctx.nconstraints = T
    ctx.gass_ngrid = 40
    ctx.loglikelihood = rowcol_loglikelihood
    setattr(F, '__worker_model', ctx)
    out = dict(dims=np.array([N, M, T, K, order, ctx.gass_ngrid]), Y=Y, W0=W, V0=V, Tau2=ctx.Tau2,
               scal=np.array([ctx.sigma2, ctx.lam2, ctx.stability]), Constraints=Constraints)
    if with_ep:
        out['Mu_ep'], out['Sigma_ep'] = ctx.Mu_ep, ctx.Sigma_ep
    with Tape(seed + 7) as tape:
        for i in range(N):
            F._resample_W_i((i, Y))
    zW = np.zeros((N, K))
    for i, z in enumerate(tape.z):
        zW[i, :len(z)] = z
    out['W_u'] = np.array(tape.u, dtype=float)
    out['W_z'] = zW
    out['W_c'], out['W_clen'] = pack_choices(tape.c)
    out['W1'] = ctx.W.copy()
    with Tape(seed + 8) as tape:
        for j in range(M):
            F._resample_V_j((j, Y))
    out['V_u'] = np.array(tape.u, dtype=float)
    out['V_z'] = np.stack(tape.z).reshape(M, T, K)          # t-major (shim permutation)
    out['V_c'], out['V_clen'] = pack_choices(tape.c)
    out['V1'] = ctx.V.copy()
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    moved_w = float(np.mean(np.any(out['W1'] != W, axis=1)))
    moved_v = float(np.mean(np.any((out['V1'] != V).reshape(M, -1), axis=1)))
    print('wrote', path, 'rows moved %.2f cols moved %.2f' % (moved_w, moved_v))


def make_gass_cases(seed=3):
    """Direct calls of the reference's gass() on small constrained problems (dense covariance)."""
    rs = np.random.RandomState(seed)
    out = {}
    ncase = 12
    for c in range(ncase):
        d = 4 + c % 3
        A = rs.normal(size=(d, d))
        Sigma = A @ A.T / d + 0.3 * np.eye(d)
        mu = rs.normal(size=d) * (c % 2)
        ncons = 3 + c % 4
        Cm = np.concatenate([rs.normal(size=(ncons, d)), np.zeros((ncons, 1))], axis=1)
        x = mu + np.abs(rs.normal(size=d))
        Cm[:, -1] = Cm[:, :-1].dot(x) - np.abs(rs.normal(size=ncons)) * (0.1 + 2 * (c % 3))   # x strictly feasible
        target = x + rs.normal(size=d)

        def ll(pts, args):
            return -0.5 * ((pts - args) ** 2).sum(axis=-1) * 3.0
        with Tape(seed + 100 + c) as tape:
            xn, lln = G.gass(x.copy(), Sigma, ll, Cm, mu=mu, ll_args=target, ngrid=25)
        # the proposal the reference drew: v = L z (covariance form, dense)
        v = np.linalg.cholesky(Sigma).dot(tape.z[0])
        pre = 'c%d_' % c
        out[pre + 'x'], out[pre + 'mu'], out[pre + 'v'], out[pre + 'C'], out[pre + 'target'] = x, mu, v, Cm, target
        out[pre + 'u'] = np.array(tape.u, dtype=float)
        out[pre + 'c'], out[pre + 'clen'] = pack_choices(tape.c)
        out[pre + 'xn'], out[pre + 'lln'] = xn, np.array([lln])
    out['ncase'] = np.array([ncase])
    path = os.path.join(ROOT, 'tests', 'golden', 'gass_cases.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


def make_ess_cases(seed=9):
    """Direct calls of the reference's elliptical_slice_ (elliptical_slice.py:59-124)."""
    from functionalmf.elliptical_slice import elliptical_slice_
    rs = np.random.RandomState(seed)
    out = {}
    ncase = 10
    orig = np.random.rand
    for c in range(ncase):
        d = 3 + c
        x = rs.normal(size=d)
        nu = rs.normal(size=d) * 1.5
        target = rs.normal(size=d)
        mu = rs.normal(size=d) * (c % 2)
        tape = []

        def rand(*a):
            v = rs.random_sample()
            tape.append(v)
            return v

        def ll(pts, args):
            return -2.0 * ((pts - args) ** 2).sum()
        np.random.rand = rand
        try:
            xn, lln = elliptical_slice_(x.copy(), nu, ll, ll_args=target, mu=mu)
        finally:
            np.random.rand = orig
        pre = 'e%d_' % c
        out[pre + 'x'], out[pre + 'nu'], out[pre + 'target'], out[pre + 'mu'] = x, nu, target, mu
        out[pre + 'u'] = np.array(tape)
        out[pre + 'xn'], out[pre + 'lln'] = xn, np.array([lln])
    out['ncase'] = np.array([ncase])
    path = os.path.join(ROOT, 'tests', 'golden', 'ess_cases.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    make_ess_cases()
    make_gass_cases()
    make_model_case('constrained_plain', False, 41)
    make_model_case('constrained_ep', True, 42)
