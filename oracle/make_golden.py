#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Runs only in the build container (it needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference package is imported from /root/reference with the three import
shims of ``oracle/shims`` (scikit-sparse, SharedArray, pypolyagamma are not
installable offline).  ``np.random.normal / gamma / random`` are replaced by a
recording tape that hands out *standard* variates from a seeded RandomState, so
every fixture stores (inputs, initial state, noise, state after every step of
``resample``, per-row / per-column Cholesky inputs and factors).  The fixtures
are what pins ``oracle/btf_oracle.py`` and, on the GPU box, the CUDA engine.
"""
import os
import sys
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)

warnings.filterwarnings('ignore', category=SyntaxWarning)
import sksparse.cholmod as shim_chol          # noqa: E402
import pypolyagamma as shim_pg                # noqa: E402
import functionalmf.factor as F               # noqa: E402


class Tape(object):
    """Recording replacement for the three np.random entry points the sweep uses."""

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)
        self.log = []          # (context, kind, array)
        self.context = 'init'
        self._orig = None

    def normal(self, loc=0.0, scale=1.0, size=None):
        shape = np.broadcast(np.asarray(loc), np.asarray(scale)).shape if size is None else size
        z = self.rs.standard_normal(shape if shape != () else None)
        self.log.append((self.context, 'normal', np.array(z, dtype=float, copy=True)))
        return loc + scale * z

    def gamma(self, shape, scale=1.0, size=None):
        oshape = np.broadcast(np.asarray(shape), np.asarray(scale)).shape if size is None else size
        if isinstance(oshape, int):
            oshape = (oshape,)
        g = self.rs.standard_gamma(shape, oshape if oshape != () else None)
        self.log.append((self.context, 'gamma', np.array(g, dtype=float, copy=True)))
        return g * scale

    def random(self, size=None):
        u = self.rs.random_sample(size)
        self.log.append((self.context, 'random', np.array(u, dtype=float, copy=True)))
        return u

    def __enter__(self):
        self._orig = (np.random.normal, np.random.gamma, np.random.random)
        np.random.normal, np.random.gamma, np.random.random = self.normal, self.gamma, self.random
        return self

    def __exit__(self, *a):
        np.random.normal, np.random.gamma, np.random.random = self._orig

    def take(self, context, kind):
        out = [a for (c, k, a) in self.log if c == context and k == kind]
        self.log = [(c, k, a) for (c, k, a) in self.log if not (c == context and k == kind)]
        return out


class CholRecorder(object):
    """Records np.linalg.cholesky inputs/outputs during the W step (factor.py:357)."""

    def __init__(self):
        self.items = []
        self.on = False
        self._orig = np.linalg.cholesky

    def __call__(self, a, *args, **kw):
        L = self._orig(a, *args, **kw)
        if self.on:
            self.items.append((np.array(a, dtype=float), np.array(L, dtype=float)))
        return L


def snapshot(model, out, tag):
    for name in ('W', 'V', 'Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
        out['%s/%s' % (tag, name)] = np.array(getattr(model, name), dtype=float, copy=True)
    for name in ('lam2', 'lam2_a', 'sigma2'):
        out['%s/%s' % (tag, name)] = np.array(getattr(model, name), dtype=float).reshape(-1)[:1].copy()
    out['%s/nu2' % tag] = np.array(model.nu2, dtype=float, copy=True)
    if hasattr(model, 'R'):
        out['%s/R' % tag] = np.array(model.R, dtype=float, copy=True)


def pad_rows(chunks, K):
    z = np.zeros((len(chunks), K))
    for i, c in enumerate(chunks):
        z[i, :len(c)] = c
    return z


def run_sweeps(model, data, tape, nsweeps, out, kind):
    """Drive resample() step by step (order: factor.py:306-311, 112-128, 494-511)."""
    N, M, T, K = model.nrows, model.ncols, model.ndepth, model.nembeds
    RD = model.Delta.shape[0]
    rec = CholRecorder()
    np.linalg.cholesky = rec
    try:
        for s in range(nsweeps):
            tag = 's%d' % s
            bdata = data
            if kind == 'negbin':
                d4 = data if data.ndim == 4 else data[..., None]
                missing = np.all(np.isnan(d4), axis=-1)
                tape.context = 'R'
                model._resample_R(d4)
                zs = tape.take('R', 'normal')
                us = tape.take('R', 'random')
                out[tag + '/noise/z_R'] = np.stack(zs)[..., 0]
                out[tag + '/noise/u_R'] = np.stack(us)[..., 0]
                out[tag + '/after_R/R'] = np.array(model.R, copy=True)
                out[tag + '/after_R/N'] = np.array(model.N, copy=True)
                Ysum = np.nansum(d4, axis=-1)
                Ysum[missing] = np.nan
                bdata = (Ysum, model.N)
            tape.context = 'nu2'
            if kind == 'gaussian':
                model._resample_nu2(bdata)
                out[tag + '/noise/g_nu2'] = np.array(tape.take('nu2', 'gamma')[0]).reshape(-1)[:1]
            else:
                shim_pg.RECORD = []
                model._resample_nu2(bdata)
                out[tag + '/noise/omega'] = shim_pg.RECORD[0].reshape(N, M, T)
                shim_pg.RECORD = None
            out[tag + '/after_nu2/nu2'] = np.array(model.nu2, dtype=float, copy=True)

            tape.context = 'sigma2'
            model._resample_sigma2()
            out[tag + '/noise/g_sigma2'] = np.array(tape.take('sigma2', 'gamma')[0]).reshape(-1)[:1]
            out[tag + '/after_sigma2/sigma2'] = np.array([model.sigma2], dtype=float)

            tape.context = 'Tau2'
            model._resample_Tau2()
            g = tape.take('Tau2', 'gamma')
            out[tag + '/noise/g_tau'] = np.stack(g).reshape(M, 4, RD)
            for name in ('Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
                out['%s/after_Tau2/%s' % (tag, name)] = np.array(getattr(model, name), copy=True)

            tape.context = 'lam2'
            model._resample_lam2()
            g = tape.take('lam2', 'gamma')
            out[tag + '/noise/g_lam'] = np.array([np.ravel(g[0])[0], np.ravel(g[1])[0]])
            out[tag + '/after_lam2/lam2'] = np.array([model.lam2], dtype=float).reshape(-1)[:1]
            out[tag + '/after_lam2/lam2_a'] = np.array([model.lam2_a], dtype=float).reshape(-1)[:1]

            tape.context = 'W'
            rec.items, rec.on = [], True
            model._resample_W(bdata)
            rec.on = False
            out[tag + '/noise/z_W'] = pad_rows(tape.take('W', 'normal'), K)
            Qs = np.zeros((N, K, K))
            Ls = np.zeros((N, K, K))
            assert len(rec.items) == N, (len(rec.items), N)
            for i, (Q, L) in enumerate(rec.items):
                d = Q.shape[0]
                Qs[i, :d, :d], Ls[i, :d, :d] = Q, L
            out[tag + '/diag/W_Q'] = Qs
            out[tag + '/diag/W_L'] = Ls
            out[tag + '/after_W/W'] = np.array(model.W, copy=True)

            tape.context = 'V'
            shim_chol.RECORD = []
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                model._resample_V(bdata)
            recs, shim_chol.RECORD = shim_chol.RECORD, None
            zv = tape.take('V', 'normal')
            out[tag + '/noise/z_V'] = np.stack(zv).reshape(M, T, K)      # t-major (shim permutation)
            assert len(recs) == M
            out[tag + '/diag/V_Q'] = np.stack([r[0] for r in recs])     # t-major dense (after jitter, if any)
            out[tag + '/diag/V_L'] = np.stack([r[1] for r in recs])
            out[tag + '/after_V/V'] = np.array(model.V, copy=True)
            snapshot(model, out, tag + '/end')
    finally:
        np.linalg.cholesky = rec._orig


def wiggly(rs, N, M, T, K, break_prob=0.3):
    """Ground truth in the spirit of examples/gaussian_tensor_filtering.py:28-44."""
    W = rs.normal(0, 1, size=(N, K))
    if N > 1:
        W[np.triu_indices(min(N, K), k=1, m=K)] = 0
    V = np.zeros((M, T, K))
    for j in range(M):
        x = rs.normal(0, 1, size=K)
        coef = rs.normal(0, 1)
        V[j, -1] = x
        for t in range(T - 2, -1, -1):
            V[j, t] = V[j, t + 1]
            if rs.random_sample() < break_prob:
                coef = rs.normal(0, 1)
                x = rs.normal(0, 1, size=K)
            V[j, t] += coef * x
    return W, V


def distinct_missing(Y, rs):
    """Give every column its own fully-missing cell so that consecutive columns
    never share an all-missing pattern (avoids the reference's stale likelihood
    cache) and there is at least one NaN (Q3)."""
    N, M, T = Y.shape[:3]
    for j in range(M):
        Y[(3 * j + 1) % N, j, (5 * j + 2) % T] = np.nan
    return Y


def make_case(name, kind, N, M, T, R, K, order, seed, nsweeps=2, nan_frac=0.0,
              rdims=(0, 1, 2), ctor_kwargs=None):
    rs = np.random.RandomState(seed)
    shim_chol.set_layout(K, T)
    out = {}
    ctor_kwargs = dict(ctor_kwargs or {})
    with Tape(seed + 1000) as tape:
        common = dict(nembeds=K, tf_order=order, sigma2_init=0.5, lam2_init=0.1, nthreads=1)
        common.update(ctor_kwargs)
        if kind == 'gaussian':
            model = F.GaussianBayesianTensorFiltering(N, M, T, nu2_init=1.0, **common)
        elif kind == 'binomial':
            model = F.BinomialBayesianTensorFiltering(N, M, T, **common)
        else:
            model = F.NegativeBinomialBayesianTensorFiltering(N, M, T, rdims=rdims, **common)
        tape.log = []
        Wt, Vt = wiggly(rs, N, M, T, K)
        Mu = np.einsum('nk,mtk->nmt', Wt, Vt)
        if kind == 'gaussian':
            Y = Mu[..., None] + rs.normal(0, 1.5, size=(N, M, T, R))
            if nan_frac > 0:
                Y[rs.random_sample(Y.shape) < nan_frac] = np.nan
            Y[:2, :2] = np.nan
            Y = distinct_missing(Y, rs)
            data = Y if R > 1 else Y[..., 0]
            out['data/Y'] = data
        elif kind == 'binomial':
            Mu = 2.0 * Mu / np.abs(Mu).max()
            Nt = np.full((N, M, T), 4.0)
            Nt[rs.random_sample(Nt.shape) < 0.3] = 7.0
            Ys = rs.binomial(Nt.astype(int), 1 / (1 + np.exp(-Mu))).astype(float)
            Ys = distinct_missing(Ys, rs)
            Ys[rs.random_sample(Ys.shape) < nan_frac] = np.nan
            Nt[np.isnan(Ys)] = np.nan
            data = (Ys, Nt)
            out['data/Y'], out['data/N'] = Ys, Nt
        else:
            Mu = 2.0 * Mu / np.abs(Mu).max()
            P = 1 / (1 + np.exp(-Mu))
            Rt = 3.0
            lamg = rs.gamma(Rt, (P / (1 - P))[..., None], size=(N, M, T, R))
            Y = rs.poisson(lamg).astype(float)
            Y[rs.random_sample(Y.shape) < nan_frac] = np.nan
            Y = distinct_missing(Y, rs)
            data = Y if R > 1 else Y[..., 0]
            out['data/Y'] = data
        snapshot(model, out, 'init')
        out['cfg/dims'] = np.array([N, M, T, R, K, order])
        out['cfg/rdims'] = np.array(sorted(rdims))
        out['cfg/Delta'] = model.Delta.toarray()
        for k in ('sigma2_a', 'sigma2_b', 'nu2_a', 'nu2_b', 'stability'):
            out['cfg/' + k] = np.array([getattr(model, k)], dtype=float)
        if kind == 'negbin':
            out['cfg/nb'] = np.array([model.nmetropolis, model.rpropstdev, model.rstdev], dtype=float)
        run_sweeps(model, data, tape, nsweeps, out, kind)
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, '%.1f KB' % (os.path.getsize(path) / 1024.0))


def make_delta_fixture():
    """bayes_grid_penalty(T, k) for a grid of (T, k)  (utils.py:83-90)."""
    from functionalmf.utils import bayes_grid_penalty
    out = {}
    for T in (4, 5, 9, 20, 33):
        for k in (0, 1, 2, 3):
            out['T%d_k%d' % (T, k)] = bayes_grid_penalty(T, k).toarray()
    path = os.path.join(ROOT, 'tests', 'golden', 'delta.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    make_delta_fixture()
    make_case('gauss_small', 'gaussian', N=7, M=5, T=9, R=1, K=3, order=2, seed=11)
    make_case('gauss_reps', 'gaussian', N=12, M=6, T=10, R=3, K=4, order=1, seed=12, nan_frac=0.25)
    make_case('gauss_p0', 'gaussian', N=6, M=4, T=8, R=2, K=2, order=0, seed=13, nan_frac=0.1)
    make_case('gauss_k8', 'gaussian', N=20, M=5, T=12, R=2, K=8, order=2, seed=14, nan_frac=0.2)
    make_case('gauss_square', 'gaussian', N=5, M=7, T=6, R=1, K=5, order=1, seed=15)
    make_case('binom_small', 'binomial', N=9, M=6, T=8, R=1, K=3, order=1, seed=21, nan_frac=0.05)
    make_case('negbin_all', 'negbin', N=6, M=5, T=7, R=2, K=3, order=2, seed=31, nan_frac=0.1,
              rdims=(0, 1, 2))
    make_case('negbin_rows', 'negbin', N=6, M=5, T=7, R=1, K=2, order=1, seed=32, nan_frac=0.1,
              rdims=(1, 2))
