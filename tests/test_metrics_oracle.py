"""oracle/metrics.py against the numbers the reference's own scoring statements produced
(tests/golden/metrics_cases.npz, oracle/make_golden_metrics.py), and the host side of the
evaluator / loaders.  CPU only."""
import os
import numpy as np
import pytest

from oracle import metrics as OM

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'metrics_cases.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def test_politics_scores(gold):
    Y, Y_train = gold['pol_Y'], gold['pol_Y_train']
    mu = OM.surface(gold['pol_Ws'], gold['pol_Vs'], 'nb_mean', gold['pol_Rs'])
    np.testing.assert_allclose(mu, gold['pol_Mu_hat'], rtol=1e-13)
    ins, out = OM.split(Y, Y_train)
    a, b = OM.per_sample_scores(Y, ins, mu, 'poisson'), OM.per_sample_scores(Y, out, mu, 'poisson')
    got = [a['rmse'], b['rmse'], a['mae'], b['mae'], a['ll'], b['ll']]
    labels = [str(x) for x in gold['pol_labels']]
    assert ['RMSE' in labels[0], 'Out' in labels[1], 'MAE' in labels[2], 'LL' in labels[5]] == [True] * 4
    np.testing.assert_allclose(got, gold['pol_values'], rtol=1e-13)


def test_flutrends_scores(gold):
    Y, Y_train = gold['flu_Y'], gold['flu_Y_train']
    mu = OM.surface(gold['flu_Ws'], gold['flu_Vs'])
    ins, out = OM.split(Y, Y_train)
    mean = mu.mean(axis=0)
    np.testing.assert_allclose(mean, gold['flu_Mu_hat_mean'], rtol=1e-13)
    a, b = OM.mean_scores(Y, ins, mean), OM.mean_scores(Y, out, mean)
    np.testing.assert_allclose([a['rmse'], b['rmse'], a['mae'], b['mae']], gold['flu_values'][2:6], rtol=1e-13)
    # the Monte-Carlo band as written, same seed -> same band
    np.random.seed(7)
    lo, hi = OM.predictive_band_mc(mu, gold['flu_nu2s'])
    np.testing.assert_array_equal(lo, gold['flu_Y_lower'])
    np.testing.assert_array_equal(hi, gold['flu_Y_upper'])
    for mask, ref in ((ins, gold['flu_values'][0]), (out, gold['flu_values'][1])):
        mc = 100 - ((Y[mask] < lo[mask]) | (Y[mask] > hi[mask])).mean() * 100
        assert mc == pytest.approx(ref, abs=1e-12)
        # the exact mixture band is the limit of the Monte-Carlo one: a few cells may flip
        exact = OM.predictive_coverage(Y, mask, mu, gold['flu_nu2s'], 95)
        assert abs(exact - ref) <= 100.0 * 3 / mask.sum()


def test_coverage_at(gold):
    samples = OM.surface(gold['cov_Ws'], gold['cov_Vs'])
    got = [OM.coverage_at(gold['cov_truth'], samples, iv) for iv in gold['cov_intervals']]
    np.testing.assert_allclose(got, gold['cov_values'], rtol=0, atol=1e-12)


def test_heldout_classes():
    from functionalmf_b200.metrics import heldout_classes, IGNORE
    Y = np.array([[1.0, np.nan, 3.0, 4.0]])
    T = np.array([[1.0, np.nan, np.nan, 4.0]])
    np.testing.assert_array_equal(heldout_classes(Y, T), [[0, IGNORE, 1, 0]])


def test_loaders_on_synthetic_files(tmp_path):
    from scipy.io import savemat
    from functionalmf_b200.datasets import load_politics, load_flu_states
    rng = np.random.RandomState(0)
    Y = rng.poisson(3.0, size=(5, 5, 12)).astype(float)
    Yt = Y.copy()
    Yt[1, 2] = np.nan
    np.save(tmp_path / 'cooperate.npy', Y)
    np.save(tmp_path / 'cooperate_train.npy', Yt)
    np.save(tmp_path / 'held_out.npy', np.array([[1, 2]]))
    d = load_politics(str(tmp_path))
    assert d['Y'].shape == (5, 5, 12) and (d['classes'][1, 2] == 1).all() and (d['classes'] == 1).sum() == 12
    assert d['held_out'].tolist() == [[1, 2]]
    weeks, series = 30, 60
    data = np.exp(rng.normal(size=(weeks, series)))
    data[:4, 7] = np.nan
    names = np.array([['s%d' % i] for i in range(series)], dtype=object)
    dates = np.array([['2004-%02d-01' % (1 + w % 12)] for w in range(weeks)], dtype=object)
    savemat(str(tmp_path / 'flu_US.mat'), dict(data=data, USnames=names, dates=dates))
    np.save(tmp_path / 'held_out_years.npy', np.array([[3, 10, 20], [6, 0, 4]]))
    f = load_flu_states(str(tmp_path))
    assert f['Y'].shape == (50, 1, weeks) and f['Y_train'].shape == (50, 1, weeks)
    np.testing.assert_allclose(f['Y'][3, 0], np.log(data[:, 4]))          # state 3 is column 4 of the file
    assert np.isnan(f['Y_train'][3, 0, 10:20]).all() and not np.isnan(f['Y_train'][3, 0, :10]).any()
    assert (f['classes'][3, 0, 10:20] == 1).all()
    assert (f['classes'][6, 0, :4] == 255).all()                          # missing in the file itself
    assert f['names'][0] == 's1'


@pytest.mark.skipif(not os.path.isdir('/root/reference/politics'), reason='reference checkout not present')
def test_loaders_on_reference_files():
    from functionalmf_b200.datasets import load_politics, load_flu_states
    d = load_politics('/root/reference/politics')
    assert d['Y'].shape == (19, 19, 228)
    assert (d['classes'] == 1).sum() == np.isnan(d['Y_train']).sum() - np.isnan(d['Y']).sum()
    pairs = {(int(i), int(j)) for i, j in d['held_out']}
    held = {(i, j) for i in range(19) for j in range(19) if (d['classes'][i, j] == 1).any()}
    assert held <= pairs
    f = load_flu_states('/root/reference/flutrends')
    assert f['Y'].shape == (50, 1, 370) and len(f['names']) == 50
    for i, j, k in f['held_out']:
        assert np.isnan(f['Y_train'][i, 0, j:k]).all()
    assert (f['classes'] == 1).sum() > 0
