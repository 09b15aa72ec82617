#!/usr/bin/env python
"""Prior draws of the reference's elliptical-slice model (factor.py:567-590) with recorded noise.
TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference).

For NonconjugateBayesianTensorFiltering the ellipse of every W / V update is spanned by the current state and ONE draw from
the prior: `_pack_W` / `_pack_V` (factor.py:155-194) build the prior precision of the packed vector and
`sample_mvn_from_precision` (fast_mvn.py:33-47) draws from it.  This script runs exactly those calls of the unmodified
reference on small problems, with `np.random.normal` replaced by a tape, and writes tests/golden/ess_prior.npz:
the state (sigma2, lam2, Tau2), the standard normals in the engine's layouts (z_W [N, K] on the free entries, z_V
[M, T, K]) and the resulting prior draws unpacked with the reference's `_unpack_W` / `_unpack_V`.
tests/test_gpu_constrained.py feeds the same normals to the engine's batched prior step and compares.
"""
import os
import sys
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')
import sksparse.cholmod as shim_chol                    # noqa: E402
import functionalmf.factor as F                         # noqa: E402
from functionalmf.fast_mvn import sample_mvn_from_precision   # noqa: E402

CASES = [(9, 4, 10, 3, 2), (5, 3, 8, 6, 1), (12, 2, 7, 4, 0)]      # N, M, T, K, tf_order  (second: nrows < nembeds)


class NormalTape(object):
    def __init__(self, seed):
        self.rs, self.draws = np.random.RandomState(seed), []

    def normal(self, loc=0.0, scale=1.0, size=None):
        z = self.rs.standard_normal(size)
        self.draws.append(np.array(z, copy=True))
        return loc + scale * z

    def __enter__(self):
        self._orig = np.random.normal
        np.random.normal = self.normal
        return self

    def __exit__(self, *a):
        np.random.normal = self._orig


def main():
    out = {}
    shim_chol.set_layout(None)          # the packed systems are factorised in their natural (k-major) order
    for ci, (N, M, T, K, order) in enumerate(CASES):
        np.random.seed(40 + ci)
        model = F.NonconjugateBayesianTensorFiltering(N, M, T, lambda W, V, d: 0.0, nembeds=K, tf_order=order,
                                                      sigma2_init=0.7, lam2_init=0.3)
        rs = np.random.RandomState(ci)
        model.Tau2 = rs.gamma(2.0, 1.0, size=model.Tau2.shape) + 0.05
        # ---- W: prior N(0, sigma2) on the free (lower-triangular + dense) entries
        cur, Q = model._pack_W(model.W)
        with NormalTape(100 + ci) as tape:
            prior = sample_mvn_from_precision(Q, sparse=True, **model.linalg_opts)
        Wp, zW = np.zeros((N, K)), np.zeros((N, K))
        model._unpack_W(prior, Wp)
        model._unpack_W(tape.draws[0], zW)
        # ---- V: one banded MVN draw per column, packed k-major (V[j].T.flatten())
        cur, Q = model._pack_V(model.V)
        with NormalTape(200 + ci) as tape:
            prior = sample_mvn_from_precision(Q, sparse=True, **model.linalg_opts)
        Vp, zV = np.zeros((M, T, K)), np.zeros((M, T, K))
        model._unpack_V(prior, Vp)
        model._unpack_V(tape.draws[0], zV)
        p = 'c%d_' % ci
        out.update({p + 'cfg': np.array([N, M, T, K, order]), p + 'sigma2': np.array(model.sigma2), p + 'lam2': np.array(model.lam2),
                    p + 'Tau2': model.Tau2, p + 'z_W': zW, p + 'prior_W': Wp, p + 'z_V': zV, p + 'prior_V': Vp})
        print('case', ci, (N, M, T, K, order), 'max |prior_W| %.3f max |prior_V| %.3f' % (np.abs(Wp).max(), np.abs(Vp).max()))
    path = os.path.join(ROOT, 'tests', 'golden', 'ess_prior.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) / 1024, 'KB')


if __name__ == '__main__':
    main()
