#!/usr/bin/env python
"""Fixtures for the scoring code from the UNMODIFIED reference scripts.
TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference).

The scripts cannot be imported (they load data, fit models and plot at module level), so the
statements that score a chain are picked out of their syntax trees and evaluated IN PLACE on a
synthetic chain: the assignments that build the surface (politics/benchmark.py:155-156), the
held-out split (164-166), every expression printed by rmse / mae / log_likelihood (168-178), the
flutrends band loop and coverage / error expressions (flutrends/benchmark.py:49-52, 66-75,
125-141) and coverage_at (examples/poisson_tensor_filtering.py:20-23).  Nothing of the reference
is copied into the repository; only inputs and the numbers it produced are stored in
tests/golden/metrics_cases.npz.
"""
import ast
import os
import sys
import numpy as np
from scipy.stats import poisson

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'


def tree_of(path):
    with open(os.path.join(REF, path)) as f:
        return ast.parse(f.read())


def assigns(tree, names, nested=False):
    """assignments to the given names, in file order (top level only unless nested)"""
    out = []
    for node in (ast.walk(tree) if nested else tree.body):
        if isinstance(node, ast.Assign) and len(node.targets) == 1:
            t = node.targets[0]
            if isinstance(t, ast.Name) and t.id in names:
                out.append(node)
    out.sort(key=lambda n: n.lineno)
    return out


def run_nodes(nodes, ns):
    mod = ast.Module(body=list(nodes), type_ignores=[])
    exec(compile(mod, '<reference>', 'exec'), ns)


def funcdef(tree, name):
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return node
    raise KeyError(name)


def formatted_values(node, ns):
    """evaluate the argument of every '...'.format(arg) call below `node`, in file order"""
    calls = [c for c in ast.walk(node) if isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute)
             and c.func.attr == 'format' and isinstance(c.func.value, ast.Constant) and len(c.args) == 1]
    calls.sort(key=lambda c: (c.lineno, c.col_offset))
    out = []
    for c in calls:
        expr = ast.Expression(body=c.args[0])
        out.append((c.func.value.value, float(eval(compile(expr, '<reference>', 'eval'), ns))))
    return out


def ilogit(x):
    return 1.0 / (1.0 + np.exp(-x))


def synthetic_chain(rng, S, N, M, T, K, kind):
    Ws = rng.normal(0, 0.6, size=(S, N, K))
    Vs = rng.normal(0, 0.6, size=(S, M, T, K)).cumsum(axis=2) * 0.3
    base_W, base_V = rng.normal(0, 0.7, size=(N, K)), rng.normal(0, 0.5, size=(M, T, K))
    Ws = base_W[None] + 0.15 * Ws
    Vs = base_V[None] + 0.15 * Vs
    return Ws, Vs


def main():
    rng = np.random.RandomState(20260101)
    out = {}

    # ---------------- politics: NB mean, in-sample / held-out RMSE, MAE, Poisson LL
    S, N, M, T, K = 12, 9, 9, 21, 3
    Ws, Vs = synthetic_chain(rng, S, N, M, T, K, 'nb')
    Rs = np.exp(rng.normal(1.0, 0.3, size=(S, N, M, T)))
    Mu_true = np.exp(rng.normal(1.0, 0.5, size=(N, M, T)))
    Y = rng.poisson(Mu_true).astype(float)
    for i in range(N):
        Y[i, i] = np.nan
    Y_train = Y.copy()
    for (i, j) in [(0, 3), (5, 2), (7, 8), (4, 1)]:
        Y_train[i, j] = np.nan
    tree = tree_of('politics/benchmark.py')
    src = open(os.path.join(REF, 'politics/benchmark.py')).read()
    # the surface statements sit inside a commented-out block there (a ''' string); parse that block too
    blocks = [n.value.value for n in tree.body if isinstance(n, ast.Expr) and isinstance(n.value, ast.Constant)
              and isinstance(n.value.value, str) and 'Rs * Ps' in n.value.value]
    ns = dict(np=np, ilogit=ilogit, poisson=poisson, Ws=Ws, Vs=Vs, Rs=Rs, Y=Y, Y_train=Y_train)
    sub = ast.parse(blocks[0]) if blocks else tree
    run_nodes(assigns(sub, {'Ps', 'Mu_hat'})[:2], ns)
    run_nodes(assigns(tree, {'is_missing', 'is_held_out', 'is_in_sample'}), ns)
    ns['mu'] = ns['Mu_hat']
    vals = []
    for fn in ('rmse', 'mae', 'log_likelihood'):
        vals += formatted_values(funcdef(tree, fn), ns)
    out.update(pol_Ws=Ws, pol_Vs=Vs, pol_Rs=Rs, pol_Y=Y, pol_Y_train=Y_train,
               pol_labels=np.array([v[0] for v in vals]), pol_values=np.array([v[1] for v in vals]),
               pol_Mu_hat=ns['Mu_hat'])
    del src

    # ---------------- flutrends: Gaussian, posterior mean errors, MC predictive band, coverage
    S, N, M, T, K = 40, 7, 1, 30, 2
    Ws, Vs = synthetic_chain(rng, S, N, M, T, K, 'gauss')
    nu2s = np.exp(rng.normal(-1.0, 0.2, size=(S, 1)))
    Mu_true = np.einsum('nk,mtk->nmt', Ws.mean(0), Vs.mean(0))
    Y = Mu_true + rng.normal(0, np.sqrt(np.exp(-1.0)), size=Mu_true.shape)
    Y[2, 0, 5:8] = np.nan
    Y_train = Y.copy()
    Y_train[1, 0, 10:20] = np.nan
    Y_train[4, 0, 0:6] = np.nan
    tree = tree_of('flutrends/benchmark.py')
    ns = dict(np=np, Ws=Ws, Vs=Vs, nu2s=nu2s, Y=Y, Y_train=Y_train)
    run_nodes(assigns(tree, {'Mu_hat', 'Mu_hat_mean', 'Mu_hat_upper', 'Mu_hat_lower'}, nested=True)[:4], ns)
    # the band loop (the `for i in range(Y.shape[0])` statement following the Y_lower, Y_upper assignment)
    loops = [n for n in ast.walk(tree) if isinstance(n, ast.For) and isinstance(n.target, ast.Name) and n.target.id == 'i'
             and any(isinstance(c, ast.Name) and c.id == 'Y_samples_ik' for c in ast.walk(n))]
    init = [n for n in ast.walk(tree) if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Tuple)
            and [getattr(e, 'id', None) for e in n.targets[0].elts] == ['Y_lower', 'Y_upper']]
    ns['print'] = lambda *a, **k: None
    np.random.seed(7)
    run_nodes([init[0], loops[0]], ns)
    # the coverage prints at module level after the "Check posterior predictive coverage" comment, and rmse / mae
    run_nodes([n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], 'id', None)
               in ('is_missing', 'is_held_out', 'is_in_sample')], ns)
    cov_prints = [n for n in tree.body if isinstance(n, ast.Expr) and isinstance(n.value, ast.Call)
                  and getattr(n.value.func, 'id', None) == 'print' and 'coverage' in ast.dump(n)]
    vals = []
    for n in cov_prints[:2]:
        vals += formatted_values(n, ns)
    ns['mu'] = ns['Mu_hat_mean']
    for fn in ('rmse', 'mae'):
        vals += formatted_values(funcdef(tree, fn), ns)
    out.update(flu_Ws=Ws, flu_Vs=Vs, flu_nu2s=nu2s, flu_Y=Y, flu_Y_train=Y_train,
               flu_labels=np.array([v[0] for v in vals]), flu_values=np.array([v[1] for v in vals]),
               flu_Y_lower=ns['Y_lower'], flu_Y_upper=ns['Y_upper'], flu_Mu_hat_mean=ns['Mu_hat_mean'],
               flu_Mu_hat_lower=ns['Mu_hat_lower'], flu_Mu_hat_upper=ns['Mu_hat_upper'])

    # ---------------- coverage_at (examples/poisson_tensor_filtering.py:20-23)
    tree = tree_of('examples/poisson_tensor_filtering.py')
    ns = dict(np=np)
    run_nodes([funcdef(tree, 'coverage_at')], ns)
    S, N, M, T, K = 25, 6, 5, 8, 2
    Ws, Vs = synthetic_chain(rng, S, N, M, T, K, 'gauss')
    # dyadic factors: every product and 2-term sum is exact in any evaluation order, so the ties
    # planted below are ties for any implementation of the dot product
    Ws, Vs = np.round(Ws * 64) / 64, np.round(Vs * 64) / 64
    samples = np.einsum('znk,zmtk->znmt', Ws, Vs)
    truth = samples.mean(0) + rng.normal(0, 0.12, size=samples.shape[1:])
    truth[0, 0, 0] = samples[3, 0, 0, 0]        # a tie with one of the samples
    truth[1, 1, 1] = samples[:, 1, 1, 1].min()  # ties at the extremes
    truth[2, 2, 2] = samples[:, 2, 2, 2].max()
    ivals = np.array([50.0, 75.0, 90.0, 95.0, 100.0])
    cov = np.array([ns['coverage_at'](truth, samples, iv) for iv in ivals])
    out.update(cov_Ws=Ws, cov_Vs=Vs, cov_truth=truth, cov_intervals=ivals, cov_values=cov)

    path = os.path.join(ROOT, 'tests', 'golden', 'metrics_cases.npz')
    np.savez_compressed(path, **out)
    for k in ('pol', 'flu'):
        for lab, v in zip(out[k + '_labels'], out[k + '_values']):
            print(k, repr(str(lab)), v)
    print('coverage_at', dict(zip(ivals, cov)))
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
