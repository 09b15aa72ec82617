"""Device-side held-out evaluation (btf_eval_*, functionalmf_b200/metrics.py) against the
reference's own scoring statements (tests/golden/metrics_cases.npz) and against oracle/metrics.py
on chains the engine itself produced.  Tolerance: 1e-10 relative on every error / likelihood sum,
exact on every coverage count."""
import os
import numpy as np
import pytest

from oracle import metrics as OM

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'metrics_cases.npz')
RTOL = 1e-10


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def _replay(model, ev, Ws, Vs, nu2s=None, Rs=None):
    for s in range(Ws.shape[0]):
        model.W[:] = Ws[s]
        model.V[:] = Vs[s]
        if nu2s is not None:
            model.nu2 = float(np.ravel(nu2s[s])[0])
        if Rs is not None:
            model.R = Rs[s]
        ev.update()


def test_politics_golden_replay(gold):
    """NB mean, per-sample RMSE / MAE / Poisson LL, in-sample vs held out (politics/benchmark.py:155-178)."""
    from functionalmf_b200 import NegativeBinomialBayesianTensorFiltering, HeldOutEvaluator
    Ws, Vs, Rs = gold['pol_Ws'], gold['pol_Vs'], gold['pol_Rs']
    S, N, K = Ws.shape
    M, T = Vs.shape[1:3]
    model = NegativeBinomialBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=2, sigma2_init=0.5,
                                                    lam2_init=0.1, rdims=(), seed=3)    # one R per cell
    ev = HeldOutEvaluator(model, gold['pol_Y'], train=gold['pol_Y_train'], transform='nb_mean',
                          loglik='poisson', max_samples=S)
    _replay(model, ev, Ws, Vs, Rs=Rs)
    r, m, ll = ev.rmse(), ev.mae(), ev.loglik()
    got = [r['in_sample'], r['held_out'], m['in_sample'], m['held_out'], ll['in_sample'], ll['held_out']]
    np.testing.assert_allclose(got, gold['pol_values'], rtol=RTOL)
    ins, out = OM.split(gold['pol_Y'], gold['pol_Y_train'])
    per = ev.per_sample()
    assert per.shape == (S, 2, 4)
    np.testing.assert_array_equal(per[:, 0, 0], ins.sum())
    np.testing.assert_array_equal(per[:, 1, 0], out.sum())
    # posterior-mean surface and its NLL (examples/poisson_tensor_filtering.py:166-168)
    mean = gold['pol_Mu_hat'].mean(axis=0)
    np.testing.assert_allclose(ev.posterior_mean()[ins | out], mean[ins | out], rtol=RTOL)
    want = OM.mean_scores(gold['pol_Y'], out, mean, 'poisson')
    assert ev.nll_of_mean()['held_out'] == pytest.approx(want['nll'], rel=RTOL)
    assert ev.rmse_of_mean()['held_out'] == pytest.approx(want['rmse'], rel=RTOL)
    ev.close()


def test_flutrends_golden_replay(gold):
    """Posterior-mean RMSE / MAE, percentile band of Mu and the predictive band
    (flutrends/benchmark.py:49-52, 66-75, 125-141); a single-column tensor like the flu data."""
    from functionalmf_b200 import GaussianBayesianTensorFiltering, HeldOutEvaluator
    Ws, Vs, nu2s = gold['flu_Ws'], gold['flu_Vs'], gold['flu_nu2s']
    Y, Y_train = gold['flu_Y'], gold['flu_Y_train']
    S, N, K = Ws.shape
    M, T = Vs.shape[1:3]
    model = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=2, sigma2_init=1, lam2_init=0.1,
                                            nu2_init=1, seed=4)
    ev = HeldOutEvaluator(model, Y, train=Y_train, loglik='gaussian', predictive=True, max_samples=S)
    _replay(model, ev, Ws, Vs, nu2s=nu2s)
    r, m = ev.rmse_of_mean(), ev.mae_of_mean()
    np.testing.assert_allclose([r['in_sample'], r['held_out'], m['in_sample'], m['held_out']],
                               gold['flu_values'][2:6], rtol=RTOL)
    ins, out = OM.split(Y, Y_train)
    mu = OM.surface(Ws, Vs)
    # band of the mean surface, percentiles exactly as np.percentile computes them
    lo, hi = gold['flu_Mu_hat_lower'], gold['flu_Mu_hat_upper']
    cov = ev.coverage(95)
    for name, mask in (('in_sample', ins), ('held_out', out)):
        want = ((Y[mask] >= lo[mask]) & (Y[mask] <= hi[mask])).mean() * 100
        assert cov[name] == pytest.approx(want, abs=1e-12)
    # predictive band: equal to the exact mixture, and within a few cells of the reference's Monte Carlo
    pc = ev.predictive_coverage(95)
    for name, mask, ref in (('in_sample', ins, gold['flu_values'][0]), ('held_out', out, gold['flu_values'][1])):
        assert pc[name] == pytest.approx(OM.predictive_coverage(Y, mask, mu, nu2s, 95), abs=1e-12)
        assert abs(pc[name] - ref) <= 100.0 * 3 / mask.sum()
    # Gaussian per-sample log-likelihood
    ll = ev.loglik()
    for name, mask in (('in_sample', ins), ('held_out', out)):
        want = OM.per_sample_scores(Y, mask, mu, 'gaussian', nu2s[:, 0])
        assert ll[name] == pytest.approx(want['ll'], rel=RTOL)
        assert ev.rmse()[name] == pytest.approx(want['rmse'], rel=RTOL)
    ev.close()


def test_coverage_at_golden_replay(gold):
    """coverage_at (examples/poisson_tensor_filtering.py:20-23) at 50/75/90/95/100 %, with targets
    that tie with a sample, with the smallest and with the largest sample."""
    from functionalmf_b200 import GaussianBayesianTensorFiltering, HeldOutEvaluator
    Ws, Vs, truth = gold['cov_Ws'], gold['cov_Vs'], gold['cov_truth']
    S, N, K = Ws.shape
    M, T = Vs.shape[1:3]
    model = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=1, sigma2_init=1, lam2_init=0.1,
                                            nu2_init=1, seed=5)
    ev = HeldOutEvaluator(model, truth, max_samples=S)
    _replay(model, ev, Ws, Vs)
    for iv, want in zip(gold['cov_intervals'], gold['cov_values']):
        assert ev.coverage(float(iv))['all'] == pytest.approx(want, abs=1e-12), iv
    # every prefix of the chain as well (different percentile positions and interpolation weights)
    samples = OM.surface(Ws, Vs)
    for n in (1, 2, 3, 7, 16):
        ev2 = HeldOutEvaluator(model, truth, max_samples=n)
        _replay(model, ev2, Ws[:n], Vs[:n])
        for iv in (0.0, 10.0, 50.0, 80.0, 95.0, 100.0):
            assert ev2.coverage(iv)['all'] == pytest.approx(OM.coverage_at(truth, samples[:n], iv), abs=1e-12), (n, iv)
        ev2.close()
    ev.close()


@pytest.mark.parametrize('K', [3, 12, 20])
def test_run_gibbs_gaussian_end_to_end(K):
    """Evaluators armed by run_gibbs score exactly the samples it returns."""
    from functionalmf_b200 import GaussianBayesianTensorFiltering, HeldOutEvaluator
    rng = np.random.RandomState(K)
    N, M, T = 70, 9, 33            # not multiples of the 64 x 256 tile
    Wt, Vt = rng.normal(size=(N, K)), rng.normal(size=(M, T, K)).cumsum(axis=1) * 0.2
    Mu = np.einsum('nk,mtk->nmt', Wt, Vt)
    Y = Mu + rng.normal(size=Mu.shape)
    Y[rng.random_sample(Y.shape) < 0.05] = np.nan
    Y_train = Y.copy()
    Y_train[rng.random_sample(Y.shape) < 0.2] = np.nan
    model = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=1, sigma2_init=0.5, lam2_init=0.1,
                                            nu2_init=1, seed=11)
    ev = HeldOutEvaluator(model, Y, train=Y_train, loglik='gaussian', predictive=True)
    ev_truth = HeldOutEvaluator(model, Mu, cell_state=True)                 # second slot: coverage of the truth
    ev_light = HeldOutEvaluator(model, Y, train=Y_train, cell_state=False)  # per-sample sums only
    res = model.run_gibbs(Y_train, nburn=15, nthin=2, nsamples=24, verbose=False)
    mu = OM.surface(res['W'], res['V'])
    ins, out = OM.split(Y, Y_train)
    for name, mask in (('in_sample', ins), ('held_out', out)):
        want = OM.per_sample_scores(Y, mask, mu, 'gaussian', res['nu2'][:, 0])
        for got in (ev, ev_light):
            assert got.rmse()[name] == pytest.approx(want['rmse'], rel=RTOL)
            assert got.mae()[name] == pytest.approx(want['mae'], rel=RTOL)
        assert ev.loglik()[name] == pytest.approx(want['ll'], rel=RTOL)
        wm = OM.mean_scores(Y, mask, mu.mean(axis=0))
        assert ev.rmse_of_mean()[name] == pytest.approx(wm['rmse'], rel=RTOL)
        assert ev.mae_of_mean()[name] == pytest.approx(wm['mae'], rel=RTOL)
        assert ev.predictive_coverage(90)[name] == pytest.approx(
            OM.predictive_coverage(Y, mask, mu, res['nu2'][:, 0], 90), abs=1e-12)
        for iv in (50, 95):
            assert ev.coverage(iv)[name] == pytest.approx(OM.coverage_at(np.nan_to_num(Y), mu, iv, mask), abs=1e-12)
    full = np.ones(Mu.shape, dtype=bool)
    for iv in (50, 75, 90, 95):
        assert ev_truth.coverage(iv)['all'] == pytest.approx(OM.coverage_at(Mu, mu, iv, full), abs=1e-12)
    with pytest.raises(Exception):
        ev_light.rmse_of_mean()             # no per-cell state was kept
    # a second chain re-arms the evaluators
    res2 = model.run_gibbs(Y_train, nburn=0, nthin=1, nsamples=5, verbose=False)
    mu2 = OM.surface(res2['W'], res2['V'])
    assert ev.per_sample().shape[0] == 5
    assert ev.rmse()['held_out'] == pytest.approx(OM.per_sample_scores(Y, out, mu2)['rmse'], rel=RTOL)
    for e in (ev, ev_truth, ev_light):
        e.close()


def test_run_gibbs_negbin_end_to_end():
    from functionalmf_b200 import NegativeBinomialBayesianTensorFiltering, HeldOutEvaluator
    rng = np.random.RandomState(8)
    N, M, T, K = 12, 11, 40, 4
    Wt, Vt = rng.normal(size=(N, K)) * 0.5, rng.normal(size=(M, T, K)).cumsum(axis=1) * 0.1
    P = OM.ilogit(np.einsum('nk,mtk->nmt', Wt, Vt))
    Y = rng.poisson(rng.gamma(5.0, P / (1 - P))).astype(float)
    for i in range(min(N, M)):
        Y[i, i] = np.nan
    Y_train = Y.copy()
    Y_train[2, 5] = np.nan
    Y_train[7, 1] = np.nan
    for rdims in ((0, 1, 2), (0, 1), ()):
        model = NegativeBinomialBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=2, sigma2_init=0.5,
                                                        lam2_init=0.1, rdims=rdims, seed=21)
        ev = HeldOutEvaluator(model, Y, train=Y_train, transform='nb_mean', loglik='poisson')
        res = model.run_gibbs(Y_train, nburn=5, nthin=1, nsamples=12, verbose=False)
        mu = OM.surface(res['W'], res['V'], 'nb_mean', res['R'])
        ins, out = OM.split(Y, Y_train)
        for name, mask in (('in_sample', ins), ('held_out', out)):
            want = OM.per_sample_scores(Y, mask, mu, 'poisson')
            assert ev.rmse()[name] == pytest.approx(want['rmse'], rel=RTOL)
            assert ev.mae()[name] == pytest.approx(want['mae'], rel=RTOL)
            assert ev.loglik()[name] == pytest.approx(want['ll'], rel=RTOL)
        ev.close()


def test_binomial_ilogit_transform():
    from functionalmf_b200 import BinomialBayesianTensorFiltering, HeldOutEvaluator
    rng = np.random.RandomState(9)
    N, M, T, K = 20, 6, 25, 3
    Wt, Vt = rng.normal(size=(N, K)), rng.normal(size=(M, T, K)).cumsum(axis=1) * 0.2
    Ptrue = OM.ilogit(np.einsum('nk,mtk->nmt', Wt, Vt))
    Nt = np.full(Ptrue.shape, 5.0)
    Ysucc = rng.binomial(5, Ptrue).astype(float)
    model = BinomialBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=1, sigma2_init=0.5, lam2_init=0.1, seed=2)
    ev = HeldOutEvaluator(model, Ptrue, transform='ilogit')
    res = model.run_gibbs((Ysucc, Nt), nburn=5, nthin=1, nsamples=10, verbose=False)
    mu = OM.surface(res['W'], res['V'], 'ilogit')
    full = np.ones(Ptrue.shape, dtype=bool)
    want = OM.per_sample_scores(Ptrue, full, mu)
    assert ev.rmse()['all'] == pytest.approx(want['rmse'], rel=RTOL)
    assert ev.mae_of_mean()['all'] == pytest.approx(OM.mean_scores(Ptrue, full, mu.mean(0))['mae'], rel=RTOL)
    assert ev.coverage(90)['all'] == pytest.approx(OM.coverage_at(Ptrue, mu, 90), abs=1e-12)
    ev.close()


def test_evaluator_errors():
    from functionalmf_b200 import GaussianBayesianTensorFiltering, HeldOutEvaluator, BTFError
    model = GaussianBayesianTensorFiltering(8, 4, 10, nembeds=2, tf_order=1, sigma2_init=1, lam2_init=0.1,
                                            nu2_init=1, seed=1)
    Y = np.zeros((8, 4, 10))
    with pytest.raises(ValueError):
        HeldOutEvaluator(model, Y[:, :, :5], max_samples=2)
    with pytest.raises(BTFError):
        HeldOutEvaluator(model, Y, transform='nb_mean', max_samples=2)      # not an NB engine
    evs = [HeldOutEvaluator(model, Y, max_samples=2) for _ in range(4)]
    with pytest.raises(RuntimeError):
        HeldOutEvaluator(model, Y, max_samples=2)                             # all four slots taken
    evs[0].update()
    evs[0].update()
    with pytest.raises(BTFError):
        evs[0].update()                                                       # max_samples reached
    with pytest.raises(BTFError):
        evs[1].coverage(95)                                                   # nothing scored yet
    for e in evs:
        e.close()
