"""Parity of the CUDA engine (through the C ABI) against the golden fixtures of the
unmodified reference and against the CPU oracle.  Run on the B200 box: pytest -m gpu."""
import os
import numpy as np
import pytest

from golden_util import Case, GAUSS_CASES, BINOM_CASES, NEGBIN_CASES, relerr, normerr, GOLDEN
from gpu_util import band_rows_from_dense, band_rows_from_lower, load_state, inject_all

pytestmark = pytest.mark.gpu

# north_star (a): 1e-10 relative in FP64 for statistics, rates, factors of well-conditioned
# systems; draws of ill-conditioned systems are bounded by c * kappa * eps.
TOL = 1e-10
EPS = np.finfo(float).eps


def _engine(c, likelihood=0, **kw):
    from functionalmf_b200.engine import Engine
    return Engine(c.N, c.M, c.T, nembeds=c.K, tf_order=c.order, likelihood=likelihood, seed=3, **kw)


def test_delta_matches_reference():
    from functionalmf_b200.engine import Engine
    z = np.load(os.path.join(GOLDEN, 'delta.npz'))
    for key in z.files:
        T, k = [int(s[1:]) for s in key.split('_')]
        if T < k + 2:
            continue
        eng = Engine(3, 2, T, nembeds=2, tf_order=k)
        assert np.array_equal(eng.get('Delta'), z[key]), key
        eng.close()


def _check_v_step(c, eng, tag, tol_scale=50.0):
    kd = (c.order + 1) * c.K
    Qref = c.z[tag + '/diag/V_Q']
    Lref = c.z[tag + '/diag/V_L']
    band = eng.diag('V_band')
    chol = eng.diag('V_chol')
    retries = eng.diag('V_retries')
    Vref = c.z[tag + '/after_V/V']
    V = eng.get('V')
    for j in range(c.M):
        assert normerr(band[j], band_rows_from_dense(Qref[j], kd)) < TOL, ('band', j)
        kappa = np.linalg.cond(Qref[j])
        tol = max(TOL, tol_scale * kappa * EPS)
        if retries[j] == 0:
            assert normerr(chol[j], band_rows_from_dense(Lref[j], kd)) < max(1e-9, tol), ('chol', j, kappa)
        assert normerr(V[j], Vref[j]) < tol, ('draw', j, kappa)
        # backward error of the conditional mean:  || Q m - b || small
    return V


@pytest.mark.parametrize('name', GAUSS_CASES)
def test_gaussian_golden(name):
    c = Case(name)
    eng = _engine(c)
    eng.set_data_gaussian(c.z['data/Y'])
    eng.enable_diag(True)
    prev = c.state('init')
    for s in range(c.nsweeps):
        tag = 's%d' % s
        load_state(eng, prev)
        inject_all(eng, c.noise(s))
        eng.sweep(1)
        assert relerr(eng.get_scalar('nu2'), c.scalar(tag + '/after_nu2/nu2')) < TOL
        assert relerr(eng.get_scalar('sigma2'), c.scalar(tag + '/after_sigma2/sigma2')) < TOL
        for nm in ('Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
            assert relerr(eng.get(nm), c.z['%s/after_Tau2/%s' % (tag, nm)]) < TOL, nm
        assert relerr(eng.get_scalar('lam2'), c.scalar(tag + '/after_lam2/lam2')) < TOL
        assert relerr(eng.get_scalar('lam2_a'), c.scalar(tag + '/after_lam2/lam2_a')) < TOL
        assert normerr(eng.diag('W_Q'), c.z[tag + '/diag/W_Q']) < TOL
        assert normerr(eng.diag('W_L'), c.z[tag + '/diag/W_L']) < 1e-9
        assert normerr(eng.get('W'), c.z[tag + '/after_W/W']) < 1e-9
        _check_v_step(c, eng, tag)
        prev = c.state(tag + '/end')
    eng.close()


@pytest.mark.parametrize('name', GAUSS_CASES)
def test_gaussian_residual_paths_agree(name):
    """nu2 rate from the direct pass == from the V-step by-product (second sweep)."""
    c = Case(name)
    rates = []
    for direct in (1, 0):
        eng = _engine(c, resid_direct=direct, use_graph=0)
        eng.set_data_gaussian(c.z['data/Y'])
        load_state(eng, c.state('init'))
        inject_all(eng, c.noise(0))
        eng.sweep(1)
        inject_all(eng, c.noise(1))
        eng.sweep(1)
        rates.append(eng.diag('nu2_rate'))
        eng.close()
    assert relerr(rates[0], rates[1]) < 1e-11


@pytest.mark.parametrize('name', BINOM_CASES)
def test_binomial_golden(name):
    c = Case(name)
    eng = _engine(c, likelihood=1)
    eng.set_data_binomial(c.z['data/Y'], c.z['data/N'])
    eng.enable_diag(True)
    prev = c.state('init')
    for s in range(c.nsweeps):
        tag = 's%d' % s
        load_state(eng, prev, gaussian=False)
        inject_all(eng, c.noise(s))
        eng.sweep(1)
        assert relerr(eng.get_scalar('sigma2'), c.scalar(tag + '/after_sigma2/sigma2')) < TOL
        assert relerr(eng.get('Tau2'), c.z[tag + '/after_Tau2/Tau2']) < TOL
        assert relerr(eng.get_scalar('lam2'), c.scalar(tag + '/after_lam2/lam2')) < TOL
        assert normerr(eng.diag('W_Q'), c.z[tag + '/diag/W_Q']) < TOL
        assert normerr(eng.get('W'), c.z[tag + '/after_W/W']) < 1e-9
        _check_v_step(c, eng, tag)
        prev = c.state(tag + '/end')
    eng.close()


@pytest.mark.parametrize('name', NEGBIN_CASES)
def test_negbin_golden(name):
    c = Case(name)
    rdims = [int(x) for x in c.z['cfg/rdims']]
    nmh, rprop, rstd = c.z['cfg/nb']
    mask = sum(1 << d for d in rdims)
    eng = _engine(c, likelihood=2, nmetropolis=int(nmh), rpropstdev=float(rprop), rstdev=float(rstd),
                  rdims_mask=mask)
    eng.set_data_negbin(c.z['data/Y'])
    eng.enable_diag(True)
    prev = c.state('init')
    for s in range(c.nsweeps):
        tag = 's%d' % s
        load_state(eng, prev, gaussian=False)
        eng.set('R', prev['R'])
        inject_all(eng, c.noise(s))
        eng.sweep(1)
        assert relerr(eng.get('R'), c.z[tag + '/after_R/R']) < TOL
        Nref = c.z[tag + '/after_R/N']
        Ngot = eng.get('Ntrials')
        obs = Nref > 0
        assert relerr(Ngot[obs], Nref[obs]) < TOL
        assert relerr(eng.get_scalar('sigma2'), c.scalar(tag + '/after_sigma2/sigma2')) < TOL
        assert normerr(eng.get('W'), c.z[tag + '/after_W/W']) < 1e-9
        _check_v_step(c, eng, tag)
        prev = c.state(tag + '/end')
    eng.close()


# ------------------------------------------------------------------ oracle comparisons at larger shapes
def _random_problem(rs, N, M, T, R, K, nan_frac=0.2):
    W = rs.normal(size=(N, K))
    W[np.triu_indices(min(N, K), k=1, m=K)] = 0
    V = rs.normal(size=(M, T, K)).cumsum(axis=1) * 0.3
    Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < nan_frac] = np.nan
    return W, V, Y


def _random_state(rs, N, M, T, K, RD, W, V):
    return dict(W=W + 0.05 * rs.normal(size=W.shape) * (W != 0), V=V + 0.05 * rs.normal(size=V.shape),
                Tau2=rs.gamma(2.0, 1.0, size=(M, RD)) + 0.05, Tau2_a=rs.gamma(2.0, 1.0, size=(M, RD)) + 0.05,
                Tau2_b=rs.gamma(2.0, 1.0, size=(M, RD)) + 0.05, Tau2_c=rs.gamma(2.0, 1.0, size=(M, RD)) + 0.05,
                lam2=0.7, lam2_a=1.3, sigma2=0.9, nu2=1.1)


def _random_noise(rs, N, M, T, K, RD):
    return dict(g_nu2=rs.gamma(50.0), g_sigma2=rs.gamma(20.0), g_tau=rs.gamma(2.0, size=(M, 4, RD)) + 1e-3,
                g_lam=rs.gamma(30.0, size=2), z_W=rs.normal(size=(N, K)), z_V=rs.normal(size=(M, T, K)))


SHAPES = [
    # N, M, T, R, K, order, extra engine options
    (300, 20, 16, 3, 16, 2, {}),
    (300, 20, 16, 3, 16, 2, dict(stats_splits_row=3, stats_splits_col=2)),
    (200, 6, 12, 2, 32, 2, {}),
    (140, 7, 9, 1, 32, 1, dict(stats_splits_row=2, stats_splits_col=3)),
    (19, 19, 40, 1, 10, 2, {}),
    (50, 3, 37, 1, 5, 2, {}),
    (130, 9, 11, 4, 8, 1, {}),
    (64, 5, 8, 2, 1, 0, {}),
    (40, 4, 6, 2, 17, 3, {}),
]


@pytest.mark.parametrize('shape', SHAPES)
def test_gaussian_sweep_vs_oracle(shape):
    from oracle import btf_oracle as O
    from functionalmf_b200.engine import Engine
    N, M, T, R, K, order, opts = shape
    rs = np.random.RandomState(1000 + N + K)
    W, V, Y = _random_problem(rs, N, M, T, R, K)
    eng = Engine(N, M, T, nembeds=K, tf_order=order, seed=5, **opts)
    Delta = eng.get('Delta')
    assert np.array_equal(Delta, O.delta_matrix(T, order))
    RD = Delta.shape[0]
    st = _random_state(rs, N, M, T, K, RD, W, V)
    noise = _random_noise(rs, N, M, T, K, RD)
    cfg = dict(K=K, order=order, Delta=Delta, nu2_a=0.1, nu2_b=0.1, sigma2_a=0.1, sigma2_b=0.1,
               stability=1e-6, force_psd=True, force_psd_eps=1e-6, force_psd_attempts=4, ref_compat=True)
    eng.set_data_gaussian(Y)
    eng.enable_diag(True)
    load_state(eng, st)
    inject_all(eng, noise)
    eng.sweep(1)
    # statistics (inputs of the solves): 1e-10
    cnt, S, _ = O.prereduce(Y)
    A, b = O.row_stats(st['V'], cnt.astype(float), S)
    rstat = eng.diag('row_stats')
    Lp = K * (K + 1) // 2
    il = np.tril_indices(K)
    assert normerr(rstat[:, :Lp], A[:, il[0], il[1]]) < TOL
    assert normerr(rstat[:, Lp:], b) < TOL
    want = O.gaussian_sweep(st, Y, noise, cfg)
    Ac, bc = O.col_stats(want['W'], cnt.astype(float), S)
    cstat = eng.diag('col_stats')
    assert normerr(cstat[:, :Lp], Ac.reshape(M * T, K, K)[:, il[0], il[1]]) < TOL
    assert normerr(cstat[:, Lp:], bc.reshape(M * T, K)) < TOL
    for k in ('nu2', 'sigma2', 'lam2', 'lam2_a'):
        assert relerr(eng.get_scalar(k), want[k]) < TOL, k
    for k in ('Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
        assert relerr(eng.get(k), want[k]) < TOL, k
    assert normerr(eng.get('W'), want['W']) < 1e-9
    # V: conditional mean and draw, kappa-scaled
    cw, sw = O.gaussian_weights(cnt, S, want['nu2'])
    Vn, dV = O.step_V(want['W'], st['V'], cw, sw, Delta, want['lam2'], want['Tau2'], noise['z_V'], order,
                      want_diag=True)
    band = eng.diag('V_band')
    mean = eng.diag('V_mean')
    Vg = eng.get('V')
    for j in range(M):
        assert normerr(band[j], band_rows_from_lower(dV['band'][j])) < TOL, j
        Qd = O.band_to_dense_lower(dV['band'][j])
        Qd = Qd + np.tril(Qd, -1).T
        kappa = np.linalg.cond(Qd)
        tol = max(TOL, 50 * kappa * EPS)
        assert normerr(mean[j], dV['mean'][j]) < tol, (j, kappa)
        assert normerr(Vg[j], Vn[j]) < tol, (j, kappa)
        # normwise backward error of the mean: ||Q m - rhs|| / (||Q|| ||m|| + ||rhs||)
        rhs = dV['b'][j].ravel()               # step_V was given the scaled weights
        m = mean[j].ravel()
        bwd = np.linalg.norm(Qd @ m - rhs) / (np.linalg.norm(Qd, 2) * np.linalg.norm(m) + np.linalg.norm(rhs))
        assert bwd < 1e-13, (j, bwd)
    eng.close()


def test_jitter_retry_matches_oracle():
    """A column whose precision is not numerically PD takes the eps, 10 eps, ... path
    of fast_mvn.py:62-68 on both sides."""
    from oracle import btf_oracle as O
    from functionalmf_b200.engine import Engine
    N, M, T, R, K, order = 8, 3, 10, 1, 2, 2
    rs = np.random.RandomState(5)
    W, V, Y = _random_problem(rs, N, M, T, R, K, nan_frac=0.0)
    Y[:, 1] = np.nan                       # column 1 has no data at all
    eng = Engine(N, M, T, nembeds=K, tf_order=order, seed=5)
    Delta = eng.get('Delta')
    RD = Delta.shape[0]
    st = _random_state(rs, N, M, T, K, RD, W, V)
    # negative anchor weight w0 = -5e-5: lambda_min(Q_1) ~ w0 / T = -5e-6, so the first jitter
    # (1e-6) still fails and the cumulative second one (1.1e-5) succeeds -> exactly 2 retries
    st['Tau2'][1] = 1.0
    st['Tau2'][1, 0] = -1.0 / (st['lam2'] * 5e-5)
    noise = _random_noise(rs, N, M, T, K, RD)
    eng.set_data_gaussian(Y)
    eng.enable_diag(True)
    load_state(eng, st)
    eng.set_sample_mask(16 | 32)           # W and V only, so Tau2 stays as set
    inject_all(eng, dict(z_W=noise['z_W'], z_V=noise['z_V']))
    eng.sweep(1)
    cnt, S, _ = O.prereduce(Y)
    cw, sw = O.gaussian_weights(cnt, S, st['nu2'])
    Wn = O.step_W(st['W'], st['V'], cw, sw, st['sigma2'], noise['z_W'])[0]
    Vn, dV = O.step_V(Wn, st['V'], cw, sw, Delta, st['lam2'], st['Tau2'], noise['z_V'], order)
    retries = eng.diag('V_retries')
    assert list(retries.astype(int)) == list(dV['retries'])
    assert retries[1] == 2
    assert normerr(eng.get('V'), Vn) < 1e-6
    eng.close()


def test_not_positive_definite_raises():
    from functionalmf_b200.engine import Engine
    from functionalmf_b200 import NotPositiveDefiniteError
    N, M, T, K = 6, 2, 8, 2
    rs = np.random.RandomState(2)
    W, V, Y = _random_problem(rs, N, M, T, 1, K, nan_frac=0.0)
    eng = Engine(N, M, T, nembeds=K, tf_order=1, seed=1, force_psd=0)
    st = _random_state(rs, N, M, T, K, eng.RD, W, V)
    st['Tau2'][:] = -1.0                   # negative prior precision: indefinite
    Y[:] = np.nan
    Y[0, 0, 0, 0] = 1.0
    eng.set_data_gaussian(Y)
    load_state(eng, st)
    eng.set_sample_mask(32)
    with pytest.raises(NotPositiveDefiniteError):
        eng.sweep(1)
    eng.close()


@pytest.mark.parametrize('shape', [(150, 6, 9, 8, 1), (260, 5, 8, 16, 2), (140, 4, 7, 32, 1), (70, 5, 6, 5, 2)])
def test_binomial_sweep_vs_oracle(shape):
    """Heteroskedastic (Polya-Gamma) path with injected omega: the f64-weight statistics kernels
    (compile-time K = 8 / 16 / 32 and the runtime-K variant) against the oracle."""
    from oracle import btf_oracle as O
    from functionalmf_b200.engine import Engine
    N, M, T, K, order = shape
    rs = np.random.RandomState(77 + K)
    W = rs.normal(size=(N, K)) * 0.5
    W[np.triu_indices(K, k=1)] = 0
    V = rs.normal(size=(M, T, K)).cumsum(axis=1) * 0.2
    Nt = rs.randint(1, 9, size=(N, M, T)).astype(float)
    Ys = rs.binomial(Nt.astype(int), 1 / (1 + np.exp(-np.einsum('nk,mtk->nmt', W, V)))).astype(float)
    miss = rs.random_sample(Ys.shape) < 0.1
    Ys[miss] = np.nan
    Nt[miss] = np.nan
    eng = Engine(N, M, T, nembeds=K, tf_order=order, likelihood=1, seed=9)
    Delta = eng.get('Delta')
    RD = Delta.shape[0]
    st = _random_state(rs, N, M, T, K, RD, W, V)
    noise = _random_noise(rs, N, M, T, K, RD)
    del noise['g_nu2']
    noise['omega'] = np.where(miss, 0.0, rs.gamma(2.0, 0.15, size=(N, M, T)) + 0.02)
    cfg = dict(K=K, order=order, Delta=Delta, nu2_a=0.1, nu2_b=0.1, sigma2_a=0.1, sigma2_b=0.1,
               stability=1e-6, force_psd=True, force_psd_eps=1e-6, force_psd_attempts=4, ref_compat=True)
    want = O.binomial_sweep(st, Ys, Nt, noise, cfg)
    eng.set_data_binomial(Ys, Nt)
    eng.enable_diag(True)
    load_state(eng, st, gaussian=False)
    inject_all(eng, noise)
    eng.sweep(1)
    cw, sw = O.binomial_weights(Ys, Nt, noise['omega'])
    A, b = O.row_stats(st['V'], cw, sw)
    Lp = K * (K + 1) // 2
    il = np.tril_indices(K)
    rstat = eng.diag('row_stats')
    assert normerr(rstat[:, :Lp], A[:, il[0], il[1]]) < TOL
    assert normerr(rstat[:, Lp:], b) < TOL
    Ac, bc = O.col_stats(want['W'], cw, sw)
    cstat = eng.diag('col_stats')
    assert normerr(cstat[:, :Lp], Ac.reshape(M * T, K, K)[:, il[0], il[1]]) < TOL
    assert normerr(cstat[:, Lp:], bc.reshape(M * T, K)) < TOL
    for k in ('sigma2', 'lam2', 'lam2_a'):
        assert relerr(eng.get_scalar(k), want[k]) < TOL, k
    assert relerr(eng.get('Tau2'), want['Tau2']) < TOL
    assert normerr(eng.get('W'), want['W']) < 1e-9
    Vn, dV = O.step_V(want['W'], st['V'], cw, sw, Delta, want['lam2'], want['Tau2'], noise['z_V'], order,
                      want_diag=True)
    Vg = eng.get('V')
    for j in range(M):
        Qd = O.band_to_dense_lower(dV['band'][j])
        Qd = Qd + np.tril(Qd, -1).T
        tol = max(TOL, 50 * np.linalg.cond(Qd) * EPS)
        assert normerr(Vg[j], Vn[j]) < tol, j
    # kappa and the trial counts the PG step sees
    kap = eng.get('kappa')
    assert np.allclose(kap[~miss], (Ys - Nt / 2)[~miss]) and np.all(kap[miss] == 0)
    eng.close()


def test_lam2_all_columns_mode_vs_oracle():
    """ref_compat=False: rate summed over all columns plus the 1/lam2_a prior term."""
    from oracle import btf_oracle as O
    from functionalmf_b200.engine import Engine
    N, M, T, R, K, order = 30, 7, 11, 2, 4, 2
    rs = np.random.RandomState(12)
    W, V, Y = _random_problem(rs, N, M, T, R, K)
    eng = Engine(N, M, T, nembeds=K, tf_order=order, seed=1, ref_compat_lam2=0)
    Delta = eng.get('Delta')
    RD = Delta.shape[0]
    st = _random_state(rs, N, M, T, K, RD, W, V)
    noise = _random_noise(rs, N, M, T, K, RD)
    eng.set_data_gaussian(Y)
    load_state(eng, st)
    inject_all(eng, noise)
    eng.set_sample_mask(4 | 8)           # Tau2 and lam2
    eng.sweep(1)
    Tau2 = O.step_tau2(st['V'], Delta, st['lam2'], st['Tau2_a'], st['Tau2_b'], st['Tau2_c'], noise['g_tau'], K)[0]
    lam2, lam2_a, rate, shape = O.step_lam2(st['V'], Delta, Tau2, st['lam2_a'], noise['g_lam'], K, ref_compat=False)
    assert relerr(eng.get_scalar('lam2'), lam2) < TOL and relerr(eng.get_scalar('lam2_a'), lam2_a) < TOL
    got = eng.diag('lam2_rate')
    assert relerr(got[0], rate) < TOL and relerr(got[1], shape) < TOL
    eng.close()
