#!/usr/bin/env python
"""Sweeps/s and per-phase times of the non-headline configurations (parity-test cases of
BASELINE.json): C1 (shipped Gaussian example), C3 (Binomial / Polya-Gamma), C4 (negative
binomial, GDELT-shaped), and a K=32 Gaussian shape (per-GPU slice of C5).  One GPU."""
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functionalmf_b200.engine import Engine      # noqa: E402
from functionalmf_b200 import _lib as L          # noqa: E402


def truth(rs, N, M, T, K):
    W = rs.normal(size=(N, K)); W[np.triu_indices(min(N, K), k=1, m=K)] = 0
    V = (rs.normal(size=(M, T, K)) * (rs.random_sample((M, T, 1)) < 0.3)).cumsum(axis=1) * 0.5
    return W, V


def run(name, eng, nsweeps, warm=5, extra=None):
    eng.init_state(127)
    eng.set('sigma2', [0.5]); eng.set('lam2', [0.1])
    if eng.likelihood == L.GAUSSIAN:
        eng.set('nu2', [1.0])
    eng.sweep(warm)
    l0 = eng.kernel_launches
    ms = eng.sweep_timed(nsweeps)
    launches = (eng.kernel_launches - l0) / float(nsweeps)
    ph = eng.time_phases(3)
    out = dict(config=name, sweeps_per_s=nsweeps / (ms * 1e-3), ms_per_sweep=ms / nsweeps,
               launches_per_sweep=launches, phases_ms=ph)
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)
    return out


def main():
    rs = np.random.RandomState(0)
    which = sys.argv[1:] or ['c1', 'c4', 'k32', 'c3', 'eval']
    if 'c1' in which:
        N, M, T, K = 11, 12, 20, 3
        W, V = truth(rs, N, M, T, K)
        Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + 3 * rs.normal(size=(N, M, T, 1))
        Y[:3, :3] = np.nan
        eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1)
        eng.set_data_gaussian(Y)
        run('C1 gaussian example 11x12x20x1 K3 p2 (CUDA graph replay)', eng, 2000)
        eng.close()
        eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1, use_graph=0)
        eng.set_data_gaussian(Y)
        run('C1 same, eager launches', eng, 2000)
        eng.close()
        # the shipped example end to end through the drop-in class: 2000 sweeps, every post-burn-in
        # sample copied back (examples/gaussian_tensor_filtering.py:49-51, 73)
        from functionalmf_b200 import GaussianBayesianTensorFiltering
        model = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=2, sigma2_init=0.5, nthreads=1,
                                                lam2_init=0.1, nu2_init=1, seed=1)
        model.run_gibbs(Y, nburn=10, nthin=1, nsamples=10, verbose=False)
        t0 = time.time()
        res = model.run_gibbs(Y, nburn=1000, nthin=1, nsamples=1000, verbose=False)
        dt = time.time() - t0
        print(json.dumps(dict(config='C1 run_gibbs(nburn=1000, nthin=1, nsamples=1000) through the Python class',
                              seconds=dt, sweeps_per_s=2000 / dt, samples=int(res['W'].shape[0]))), flush=True)
    if 'c4' in which:
        N, M, T, K = 19, 19, 228, 10
        W, V = truth(rs, N, M, T, K)
        Mu = np.einsum('nk,mtk->nmt', W, V)
        Mu = 2 * Mu / np.abs(Mu).max()
        P = 1 / (1 + np.exp(-Mu))
        Y = rs.poisson(rs.gamma(5.0, P / (1 - P))).astype(float)
        Y[rs.random_sample(Y.shape) < 0.1025] = np.nan
        eng = Engine(N, M, T, nembeds=K, tf_order=2, likelihood=L.NEGBINOMIAL, seed=1)
        eng.set_data_negbin(Y)
        run('C4 negative binomial 19x19x228 K10 p2, 30 MH steps, rdims=(0,1,2)', eng, 200)
        eng.close()
    if 'k32' in which:
        N, M, T, R, K = 2048, 256, 64, 2, 32
        W, V = truth(rs, N, M, T, K)
        Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + rs.normal(size=(N, M, T, R))
        Y[rs.random_sample(Y.shape) < 0.2] = np.nan
        eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1)
        eng.set_data_gaussian(Y)
        cells = N * M * T
        fl = 4.0 * cells * (K * (K + 1) // 2 + K)
        o = run('K32 gaussian 2048x256x64x2 K32 p2 (C5-like per-GPU slice)', eng, 20)
        st = o['phases_ms']['row_stats'] + o['phases_ms']['col_stats']
        print(json.dumps(dict(k32_stats_tflops=fl / (st * 1e-3) / 1e12)), flush=True)
        eng.close()
    if 'c3' in which:
        N, M, T, K = 4096, 1024, 32, 8
        W, V = truth(rs, N, M, T, K)
        Mu = np.einsum('nk,mtk->nmt', W, V)
        Mu = 3 * Mu / np.abs(Mu).max()
        Nt = np.full((N, M, T), 4.0)
        Ys = rs.binomial(4, 1 / (1 + np.exp(-Mu))).astype(float)
        miss = rs.random_sample(Ys.shape) < 0.01
        Ys[miss] = np.nan; Nt[miss] = np.nan
        eng = Engine(N, M, T, nembeds=K, tf_order=1, likelihood=L.BINOMIAL, seed=1)
        t0 = time.time()
        eng.set_data_binomial(Ys, Nt)
        cells = N * M * T
        fl = 4.0 * cells * (K * (K + 1) // 2 + K)
        o = run('C3 binomial 4096x1024x32 (4 trials) K8 p1, 1% NaN', eng, 10, extra=dict(set_data_s=time.time() - t0))
        st = o['phases_ms']['row_stats'] + o['phases_ms']['col_stats']
        print(json.dumps(dict(c3_stats_tflops=fl / (st * 1e-3) / 1e12,
                              c3_pg_draws_per_s=cells * 4 / (o['phases_ms']['nu2_or_pg'] * 1e-3))), flush=True)
        eng.close()

    if 'eval' in which:
        # held-out evaluator (eval_kernels.cu) at the C2 cell count: one scoring pass per saved sample
        N, M, T, K = 4096, 1024, 64, 16
        W, V = truth(rs, N, M, T, K)
        eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1)
        eng.set('W', W); eng.set('V', V); eng.set('nu2', [1.0])
        target = np.einsum('nk,mtk->nmt', W, V) + rs.normal(size=(N, M, T))
        cls = (rs.random_sample((N, M, T)) < 0.1).astype(np.uint8)
        cells = N * M * T
        for state, bpc in ((0, 9), (1, 9 + 2 * 32), (2, 9 + 2 * 40)):
            eng.eval_set(0, target, cls, nclasses=2, loglik=1, cell_state=state, auto_update=False, max_samples=64)
            for _ in range(3):
                eng.eval_update(0)
            eng.synchronize()
            t0 = time.time()
            for _ in range(20):
                eng.eval_update(0)
            eng.synchronize()
            dt = (time.time() - t0) / 20
            print(json.dumps(dict(config='held-out evaluator, 4096x1024x64 cells K16, cell_state=%d' % state,
                                  ms_per_sample=dt * 1e3, algorithmic_bytes_per_cell=bpc,
                                  gb_per_s=cells * bpc / dt / 1e9)), flush=True)
            eng.eval_clear(0)
        eng.close()


if __name__ == '__main__':
    main()
