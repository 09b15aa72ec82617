#!/usr/bin/env python
"""Posterior summaries of the UNMODIFIED reference on the shipped Gaussian example
(examples/gaussian_tensor_filtering.py: 11 x 12 x 20 x 1, nembeds=3, tf_order=2).
TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference).

Writes tests/golden/posterior_c1.npz: the data, and per-chain batch means of
Mu = einsum(W, V) and Mu^2 from free-running chains, from which the GPU test derives
posterior means / variances with batch-means Monte-Carlo standard errors.
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
from functionalmf.factor import GaussianBayesianTensorFiltering   # noqa: E402

N, M, T, K, ORDER = 11, 12, 20, 3, 2
NCHAINS, NBURN, NSAMPLES, NBATCH = 4, 500, 1500, 10


def example_data(seed=1):
    """Data generator of examples/gaussian_tensor_filtering.py:28-44, 62-70."""
    rs = np.random.RandomState(seed)
    W = rs.normal(0, 1, size=(N, K))
    W[np.triu_indices(K, k=1)] = 0
    V = np.zeros((M, T, K))
    for j in range(M):
        x = rs.normal(0, 1, size=K)
        coef = rs.normal(0, 1)
        V[j, -1] = x
        for t in range(T - 2, -1, -1):
            V[j, t] = V[j, t + 1]
            if rs.random_sample() < 0.3:
                coef = rs.normal(0, 1)
                x = rs.normal(0, 1, size=K)
            V[j, t] += coef * x
    Mu = np.einsum('nk,mtk->nmt', W, V)
    Y = rs.normal(Mu[..., None], 3.0, size=(N, M, T, 1))
    Y[:3, :3] = np.nan
    return Y, Mu


def batch_moments(Ws, Vs, nbatch):
    Mu = np.einsum('znk,zmtk->znmt', Ws, Vs)
    b = Mu.reshape(nbatch, -1, *Mu.shape[1:])
    return b.mean(axis=1), (b ** 2).mean(axis=1)


if __name__ == '__main__':
    Y, Mu_true = example_data()
    shim_chol.set_layout(K, T)
    m1 = np.zeros((NCHAINS, NBATCH, N, M, T))
    m2 = np.zeros_like(m1)
    scal = np.zeros((NCHAINS, 3))
    t0 = time.time()
    for c in range(NCHAINS):
        np.random.seed(100 + c)
        model = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=ORDER, sigma2_init=0.5, nthreads=1,
                                                lam2_init=0.1, nu2_init=1)
        res = model.run_gibbs(Y, nburn=NBURN, nthin=1, nsamples=NSAMPLES, verbose=False)
        m1[c], m2[c] = batch_moments(res['W'], res['V'], NBATCH)
        scal[c] = [np.median(res['nu2']), np.median(res['sigma2']), np.median(res['lam2'])]
        print('chain', c, 'done %.0fs' % (time.time() - t0), 'median nu2/sigma2/lam2', scal[c], flush=True)
    out = os.path.join(ROOT, 'tests', 'golden', 'posterior_c1.npz')
    np.savez_compressed(out, Y=Y, Mu_true=Mu_true, m1=m1.astype(np.float32), m2=m2.astype(np.float32), scal=scal,
                        cfg=np.array([N, M, T, K, ORDER, NCHAINS, NBURN, NSAMPLES, NBATCH]))
    print('wrote', out, os.path.getsize(out) / 1024, 'KB')
