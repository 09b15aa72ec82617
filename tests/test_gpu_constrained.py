"""Constrained non-conjugate model: the engine-assisted GASS W and V steps
against the reference's own per-row / per-column workers (fixtures of
oracle/make_golden_constrained.py) with identical noise, plus a short free-running chain."""
import os
import numpy as np
import pytest

from golden_util import GOLDEN
from constrained_ll import rowcol_loglikelihood, ReplayRng
from test_gass_host import unpack_choices

pytestmark = pytest.mark.gpu


def _model(z, with_ep):
    from functionalmf_b200 import ConstrainedNonconjugateBayesianTensorFiltering
    N, M, T, K, order, ngrid = [int(x) for x in z['dims']]
    ep = (z['Mu_ep'], z['Sigma_ep']) if with_ep else None
    sigma2, lam2, stab = [float(x) for x in z['scal']]
    m = ConstrainedNonconjugateBayesianTensorFiltering(
        N, M, T, rowcol_loglikelihood, z['Constraints'], ep_approx=ep, nembeds=K, tf_order=order,
        gass_ngrid=ngrid, sigma2_init=sigma2, lam2_init=lam2, Tau2_init=z['Tau2'], W_init=z['W0'], V_init=z['V0'],
        stability=stab, seed=3, nthreads=2, multiprocessing=True, sharedprefix='x')
    return m


@pytest.mark.parametrize('name,with_ep', [('constrained_plain', False), ('constrained_ep', True)])
def test_gass_steps_match_reference_workers(name, with_ep):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    m = _model(z, with_ep)
    assert np.allclose(m.W, z['W0']) and np.allclose(m.V, z['V0']) and np.allclose(m.Tau2, z['Tau2'])
    Y = z['Y']
    m._sync_ep()
    m._rng = ReplayRng(z['W_u'], unpack_choices(z['W_c'], z['W_clen']))
    m._resample_W(Y, z=z['W_z'])
    assert not m._rng.u and not m._rng.c
    assert np.allclose(m.W, z['W1'], rtol=1e-9, atol=1e-11)
    m.W[:] = z['W1']
    m._rng = ReplayRng(z['V_u'], unpack_choices(z['V_c'], z['V_clen']))
    m._resample_V(Y, z=z['V_z'])
    assert not m._rng.u and not m._rng.c
    assert np.allclose(m.V, z['V1'], rtol=1e-7, atol=1e-9)
    m.shutdown()


def test_constrained_chain_keeps_constraints():
    z = np.load(os.path.join(GOLDEN, 'constrained_plain.npz'))
    m = _model(z, False)
    np.random.seed(5)
    res = m.run_gibbs(z['Y'], nburn=3, nthin=1, nsamples=5, verbose=False)
    assert res['W'].shape[0] == 5 and res['V'].shape[0] == 5
    Mu = np.einsum('znk,zmtk->znmt', res['W'], res['V'])
    assert np.all(Mu >= -1e-9)                 # positivity constraints of the fixture
    assert np.all(np.isfinite(res['Tau2'])) and np.all(res['sigma2'] > 0)
    assert not np.allclose(res['V'][0], res['V'][-1])


def test_nonconjugate_ess_chain_runs_and_improves_fit():
    """factor.py:567-612: joint elliptical slice sampling with engine-drawn priors."""
    from functionalmf_b200 import NonconjugateBayesianTensorFiltering
    rs = np.random.RandomState(3)
    N, M, T, K = 8, 5, 7, 2
    Wt = rs.normal(size=(N, K)); Wt[np.triu_indices(K, k=1)] = 0
    Vt = rs.normal(size=(M, T, K)).cumsum(axis=1) * 0.4
    Y = np.einsum('nk,mtk->nmt', Wt, Vt) + 0.3 * rs.normal(size=(N, M, T))

    def loglik(W, V, data):
        return -0.5 * np.nansum((data - np.einsum('nk,mtk->nmt', W, V)) ** 2) / 0.09
    np.random.seed(1)
    m = NonconjugateBayesianTensorFiltering(N, M, T, loglik, nembeds=K, tf_order=1, sigma2_init=1.0, lam2_init=0.5,
                                            seed=4)
    assert np.all(m.W[np.triu_indices(K, k=1)] == 0)
    ll0 = m.logprob(Y)
    res = m.run_gibbs(Y, nburn=150, nthin=1, nsamples=20, verbose=False)
    assert res['W'].shape == (20, N, K) and res['V'].shape == (20, M, T, K)
    assert np.all(res['W'][:, np.triu_indices(K, k=1)[0], np.triu_indices(K, k=1)[1]] == 0)
    assert m.logprob(Y) > ll0 + 10          # slice sampling never decreases below the slice; the fit improves
    assert np.all(np.isfinite(res['Tau2'])) and np.all(res['sigma2'] > 0)


@pytest.mark.parametrize('case', [0, 1, 2])
def test_ess_prior_draws_match_reference_pack_path(case):
    """the ellipse of NonconjugateBayesianTensorFiltering is spanned by ONE prior draw per update.  The
    reference packs W / V into a vector, builds the packed prior precision (`_pack_W`, `_pack_V`, factor.py:155-194) and
    calls `sample_mvn_from_precision` (fast_mvn.py:33-47); the engine draws the same priors for all free W entries / all
    columns of V in one batched step (K2 / K3 with zero statistics).  Same standard normals in, same draws out: fixtures
    of oracle/make_golden_ess_prior.py (unmodified reference, recorded noise), including nrows < nembeds."""
    from functionalmf_b200 import NonconjugateBayesianTensorFiltering
    from functionalmf_b200 import _lib as L
    z = np.load(os.path.join(GOLDEN, 'ess_prior.npz'))
    p = 'c%d_' % case
    N, M, T, K, order = [int(x) for x in z[p + 'cfg']]
    m = NonconjugateBayesianTensorFiltering(N, M, T, lambda W, V, d: 0.0, nembeds=K, tf_order=order,
                                            sigma2_init=float(z[p + 'sigma2']), lam2_init=float(z[p + 'lam2']),
                                            Tau2_init=z[p + 'Tau2'], seed=9)
    m._sync_ep()
    free = m._free_mask()
    # W: x = sqrt(sigma2) z on the free entries, structural zeros elsewhere
    m._engine_step(L.SAMPLE_W, 'z_W', z[p + 'z_W'])
    Wp = m._engine.get('W')
    assert np.all(Wp[~free] == 0)
    assert np.max(np.abs(Wp[free] - z[p + 'prior_W'][free])) <= 1e-12 * np.max(np.abs(z[p + 'prior_W']))
    # V: x_j = chol(Delta^T diag(1 / (lam2 tau2_j)) Delta (x) I_K)^-T z_j; the reference factorises the same matrix in
    # k-major order, so the draws agree to kappa * eps
    m._engine_step(L.SAMPLE_V, 'z_V', z[p + 'z_V'])
    Vp = m._engine.get('V')
    Delta = np.asarray(m.Delta.todense())
    for j in range(M):
        P = Delta.T @ ((1.0 / (float(z[p + 'lam2']) * z[p + 'Tau2'][j]))[:, None] * Delta)
        tol = max(1e-10, 50 * np.linalg.cond(P) * np.finfo(float).eps)
        err = np.max(np.abs(Vp[j] - z[p + 'prior_V'][j])) / np.max(np.abs(z[p + 'prior_V'][j]))
        assert err < tol, (j, err, tol)
