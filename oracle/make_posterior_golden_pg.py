#!/usr/bin/env python
"""Posterior summaries of the UNMODIFIED reference on the shipped Binomial and Negative-Binomial example problems
(examples/binomial_tensor_filtering.py:14-25, 27-79; examples/negbinom_tensor_filtering.py:13-25, 44-95).
TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference):

    python oracle/make_posterior_golden_pg.py binom|negbin

Writes tests/golden/posterior_{binom,negbin}.npz: the data and per-chain batch means of Mu = einsum(W, V) (the logit
scale) and Mu^2 from free-running chains of the reference under the import shims (pypolyagamma -> oracle/pg.py,
CHOLMOD -> banded LAPACK Cholesky), from which tests/test_gpu_posterior.py derives posterior means / variances with
batch-means Monte-Carlo standard errors.

Data: the examples' generators (same shapes, nembeds, tf_order, trials / replicates, the held-out [:3, :3] block), plus
ONE extra missing cell per column at a column-specific position.  Why: the reference's V step caches the likelihood
part of the precision (`Xt`, `Q_likelihood`, factor.py:394-400) and rebuilds it only when the MISSING PATTERN differs
from the previous column's, although on the Polya-Gamma paths the weights 1/nu2 differ in every column (SURVEY.md
appendix D, Q2/Q3).  On the examples as shipped 2 of 12 columns are rebuilt and the other 10 are drawn from a
precision that belongs to another column - not a sampler of the model's posterior, and not something the engine
reproduces (it always uses the exact statistics).  A distinct pattern per column makes the reference rebuild every
column, so both sides sample the same, correct posterior and can be compared.  (The W step rebuilds every row whenever
the data contain any NaN, factor.py:320/349.)
"""
import os
import sys
import time
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')
import sksparse.cholmod as shim_chol                    # noqa: E402
from functionalmf.factor import BinomialBayesianTensorFiltering, NegativeBinomialBayesianTensorFiltering   # noqa: E402
from functionalmf.utils import ilogit                   # noqa: E402

N, M, T, K, ORDER = 11, 12, 20, 3, 2
NCHAINS, NBURN, NSAMPLES, NBATCH = 4, 1000, 3000, 10


def distinct_patterns(Y):
    """Hold out the examples' [:3, :3] block and one more cell per column, at a different place in every column."""
    Y = Y.copy()
    Y[:3, :3] = np.nan
    for j in range(M):
        Y[3 + (j % (N - 3)), j, (7 * j + 3) % T] = np.nan
    pats = [np.isnan(Y[:, j].reshape(N, -1)).tobytes() for j in range(M)]
    assert all(pats[j] != pats[j - 1] for j in range(1, M))
    return Y


def binom_data(seed=1):
    """examples/binomial_tensor_filtering.py:27-43, 59-74 (10 trials per cell)."""
    rs = np.random.RandomState(seed)
    W = rs.normal(0, 1, size=(N, K))
    W[np.triu_indices(K, k=1)] = 0
    V = np.zeros((M, T, K))
    for j in range(M):
        x = rs.normal(0, 1, size=K)
        coef = rs.normal(0, 0.1)
        V[j, -1] = x
        for t in range(T - 2, -1, -1):
            V[j, t] = V[j, t + 1]
            if rs.random_sample() < 0.3:
                coef = rs.normal(0, 0.1)
                x = rs.normal(0, 1, size=K)
            V[j, t] += coef * x
    Mu = np.einsum('nk,mtk->nmt', W, V)
    Nt = np.full((N, M, T), 10.0)
    Y = rs.binomial(10, ilogit(Mu)).astype(float)
    Y = distinct_patterns(Y[..., None])[..., 0]
    Nt[np.isnan(Y)] = np.nan
    return Y, Nt, Mu


def negbin_data(seed=42):
    """examples/negbinom_tensor_filtering.py:44-62, 80-90 (piecewise-constant rates, one replicate)."""
    rs = np.random.RandomState(seed)
    W = rs.gamma(1, 1, size=(N, K))
    W[np.triu_indices(K, k=1)] = 0
    V = np.zeros((M, T, K))
    for j in range(M):
        V[j, -1] = rs.gamma(1, 1, size=K)
        for t in range(T - 2, -1, -1):
            V[j, t] = V[j, t + 1]
            if rs.random_sample() < 0.2:
                V[j, t] += rs.gamma(1, 1, size=K)
    Mu = np.einsum('nk,mzk->nmz', W, V)
    Var = rs.gamma(1, scale=1, size=(N, 1, 1)) * Mu ** 2 + Mu
    P = 1 - Mu / Var
    R = Mu * (1 - P) / P
    Y = rs.poisson(rs.gamma(R[..., None], scale=(P / (1 - P))[..., None], size=(N, M, T, 1))).astype(float)
    return distinct_patterns(Y), Mu


def batch_moments(Ws, Vs, nbatch):
    Mu = np.einsum('znk,zmtk->znmt', Ws, Vs)
    b = Mu.reshape(nbatch, -1, *Mu.shape[1:])
    return b.mean(axis=1), (b ** 2).mean(axis=1)


def main(kind):
    shim_chol.set_layout(K, T)
    m1 = np.zeros((NCHAINS, NBATCH, N, M, T))
    m2 = np.zeros_like(m1)
    scal = np.zeros((NCHAINS, 2))
    t0 = time.time()
    if kind == 'binom':
        Y, Nt, Mu_true = binom_data()
        data, extra = (Y, Nt), dict(Y=Y, Nt=Nt)
    else:
        Y, Mu_true = negbin_data()
        data, extra = Y, dict(Y=Y)
    Rm = np.zeros((NCHAINS, N, 1, 1))
    for c in range(NCHAINS):
        np.random.seed(200 + c)
        if kind == 'binom':
            model = BinomialBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=ORDER, sigma2_init=0.5, nthreads=1,
                                                    lam2_init=0.1, pg_seed=300 + c)
        else:
            model = NegativeBinomialBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=ORDER, sigma2_init=0.5,
                                                            nthreads=1, lam2_init=0.1, rdims=(1, 2), pg_seed=300 + c)
        res = model.run_gibbs(data, nburn=NBURN, nthin=1, nsamples=NSAMPLES, verbose=False)
        m1[c], m2[c] = batch_moments(res['W'], res['V'], NBATCH)
        scal[c] = [np.median(res['sigma2']), np.median(res['lam2'])]
        if kind == 'negbin':
            Rm[c] = res['R'].mean(axis=0)
        print(kind, 'chain', c, 'done %.0fs' % (time.time() - t0), 'median sigma2/lam2', scal[c], flush=True)
    out = os.path.join(ROOT, 'tests', 'golden', 'posterior_%s.npz' % kind)
    np.savez_compressed(out, Mu_true=Mu_true, m1=m1.astype(np.float32), m2=m2.astype(np.float32), scal=scal, R_mean=Rm,
                        cfg=np.array([N, M, T, K, ORDER, NCHAINS, NBURN, NSAMPLES, NBATCH]), **extra)
    print('wrote', out, os.path.getsize(out) / 1024, 'KB')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'binom')
