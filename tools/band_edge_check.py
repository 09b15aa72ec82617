#!/usr/bin/env python
"""Small-T / small-M edge cases of the band solve: look-ahead kernel vs the scalar kernel (separate processes,
the kernel choice is read once per process), same injected noise."""
import os, subprocess, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [(6, 3, 2, 3, 0), (6, 3, 4, 3, 2), (9, 2, 4, 8, 2), (20, 4, 5, 16, 1), (12, 1, 5, 5, 3), (40, 3, 4, 32, 2), (10, 2, 3, 16, 1), (10, 2, 5, 16, 3)]


def run_case(N, M, T, K, order):
    from functionalmf_b200.engine import Engine
    from functionalmf_b200 import _lib as L
    rs = np.random.RandomState(N * 100 + T)
    W = rs.normal(size=(N, K)); V = rs.normal(size=(M, T, K))
    Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + rs.normal(size=(N, M, T, 2))
    Y[rs.random_sample(Y.shape) < 0.2] = np.nan
    eng = Engine(N, M, T, nembeds=K, tf_order=order, seed=3, use_graph=0)
    eng.set_data_gaussian(Y)
    RD = eng.RD
    eng.set('W', W); eng.set('V', V)
    for k in ('Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
        eng.set(k, rs.gamma(2.0, size=(M, RD)) + 0.1)
    for k, v in dict(lam2=0.7, lam2_a=1.3, sigma2=0.9, nu2=1.1).items():
        eng.set(k, [v])
    eng.set_sample_mask(L.SAMPLE_V)
    eng.inject('z_V', rs.normal(size=(M, T, K)))
    eng.sweep(1)
    out = eng.get('V')
    eng.close()
    return out


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        res = [run_case(*c).tolist() for c in CASES]
        json.dump(res, open(sys.argv[2], 'w'))
        sys.exit(0)
    outs = {}
    for mode, env in (('lookahead', {}), ('scalar', {'BTF_BAND_SCALAR': '1'})):
        f = '/tmp/band_edge_%s.json' % mode
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), 'child', f], check=True, env=e)
        outs[mode] = json.load(open(f))
    ok = True
    for c, a, b in zip(CASES, outs['lookahead'], outs['scalar']):
        a, b = np.array(a), np.array(b)
        err = float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        good = err < 1e-8 and np.all(np.isfinite(a))
        ok = ok and good
        print('case N,M,T,K,order =', c, 'max normwise diff %.2e' % err, 'OK' if good else 'MISMATCH')
    print('BAND_EDGE', 'PASS' if ok else 'FAIL')
