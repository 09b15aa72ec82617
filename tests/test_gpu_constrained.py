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
