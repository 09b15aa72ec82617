"""Integer-tensor-core statistics path (i8gemm.cu, stats_i8.cu).

* the tcgen05 int8 GEMM is bit-exact against numpy integer arithmetic (ragged shapes included);
* the product block of the sufficient statistics equals, BIT FOR BIT, the correctly rounded value
  of the fixed-point sum it is defined as (checked in Python big-integer arithmetic);
* a sweep on the integer path equals the same sweep on the FP64 DMMA path to rounding.
"""
import ctypes as C
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(A, B):
    from functionalmf_b200 import _lib as L
    lib = L.load()
    M, K = A.shape
    N = B.shape[0]
    D = np.zeros((M, N), dtype=np.int32)
    ms = lib.btf_i8gemm_test(0, C.c_void_p(A.ctypes.data), C.c_void_p(B.ctypes.data), C.c_void_p(D.ctypes.data), M, N, K, 1)
    assert ms >= 0
    return D


@pytest.mark.parametrize('shape', [(128, 256, 128), (1, 1, 128), (130, 257, 384), (200, 300, 1024), (1088, 520, 2048), (128, 256, 8192), (300, 100, 16384)])
def test_i8gemm_exact(shape):
    M, N, K = shape
    rs = np.random.RandomState(M + N + K)
    A = rs.randint(-64, 64, size=(M, K)).astype(np.int8)
    B = rs.randint(0, 4, size=(N, K)).astype(np.int8)
    np.testing.assert_array_equal(_gemm(A, B), A.astype(np.int32) @ B.astype(np.int32).T)
    # extreme digits and counts: the int32 accumulators hold them exactly
    A[:] = -64
    B[:] = 127
    np.testing.assert_array_equal(_gemm(A, B), np.full((M, N), -64 * 127 * K, dtype=np.int32))


GEMM_MODES = ('BTF_STATS_I8_FUSED', 'BTF_STATS_I8_NOFUSED')   # recombination in the GEMM epilogue / through the int32 planes


def _engine(N, M, T, R, K, env, gemm_mode=None):
    from functionalmf_b200.engine import Engine
    for k in ('BTF_STATS_FORCE_I8', 'BTF_STATS_NO_I8') + GEMM_MODES:
        os.environ.pop(k, None)
    os.environ[env] = '1'
    if gemm_mode:
        os.environ[gemm_mode] = '1'
    return Engine(N, M, T, nembeds=K, tf_order=2, seed=5, use_graph=0)


def _problem(N, M, T, R, K, seed):
    rs = np.random.RandomState(seed)
    W = rs.normal(size=(N, K))
    V = rs.normal(size=(M, T, K)).cumsum(axis=1) * 0.3
    Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < 0.25] = np.nan
    RD = None
    return rs, W, V, Y


def _load(eng, rs, W, V):
    RD = eng.RD
    M = V.shape[0]
    eng.set('W', W * 0.9)
    eng.set('V', V * 1.1)
    for k in ('Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
        eng.set(k, np.random.RandomState(3).gamma(2.0, size=(M, RD)) + 0.05)
    for k, v in dict(lam2=0.7, lam2_a=1.3, sigma2=0.9, nu2=1.1).items():
        eng.set(k, [v])


@pytest.mark.parametrize('gemm_mode', GEMM_MODES)
@pytest.mark.parametrize('shape', [(150, 9, 21, 3, 8), (300, 7, 40, 2, 16), (129, 5, 33, 4, 32)])
def test_product_block_is_the_exact_fixed_point_sum(shape, gemm_mode):
    N, M, T, R, K = shape
    rs, W, V, Y = _problem(N, M, T, R, K, 11)
    eng = _engine(N, M, T, R, K, 'BTF_STATS_FORCE_I8', gemm_mode)
    try:
        eng.set_data_gaussian(Y)
        _load(eng, rs, W, V)
        eng.enable_diag(True)
        from functionalmf_b200 import _lib as L
        eng.set_sample_mask(L.SAMPLE_W | L.SAMPLE_V)      # W and V steps only
        z = np.zeros(N * K)
        eng.inject('z_W', z)                  # W stays at its conditional mean: the V statistics use a known W
        eng.sweep(1)
        L = K * (K + 1) // 2
        cnt = (~np.isnan(Y)).sum(axis=-1).reshape(N, M * T)
        Ssum = np.nansum(Y, axis=-1).reshape(N, M * T)

        from oracle.fixed_point import product_block as exact     # the definition, in Python integers

        rows = eng.diag('row_stats')
        Vf = (V * 1.1).reshape(M * T, K)
        want = exact(Vf, cnt)
        np.testing.assert_array_equal(rows[:, :L], want)
        # linear block: plain FP64
        np.testing.assert_allclose(rows[:, L:], Ssum @ Vf, rtol=1e-12, atol=1e-12 * np.abs(Ssum @ Vf).max())
        cols = eng.diag('col_stats')
        Wn = eng.get('W')
        want = exact(Wn, cnt.T)
        np.testing.assert_array_equal(cols[:, :L], want)
        np.testing.assert_allclose(cols[:, L:], Ssum.T @ Wn, rtol=1e-12, atol=1e-12 * np.abs(Ssum.T @ Wn).max())
    finally:
        eng.close()
        for k in ('BTF_STATS_FORCE_I8',) + GEMM_MODES:
            os.environ.pop(k, None)


@pytest.mark.parametrize('gemm_mode', GEMM_MODES)
@pytest.mark.parametrize('shape', [(260, 12, 24, 3, 16), (140, 6, 30, 2, 8), (200, 5, 20, 3, 32)])
def test_sweep_matches_the_fp64_path(shape, gemm_mode):
    N, M, T, R, K = shape
    rs, W, V, Y = _problem(N, M, T, R, K, 21)
    res = {}
    for env in ('BTF_STATS_FORCE_I8', 'BTF_STATS_NO_I8'):
        eng = _engine(N, M, T, R, K, env, gemm_mode)
        try:
            eng.set_data_gaussian(Y)
            _load(eng, rs, W, V)
            eng.enable_diag(True)
            eng.sweep(2)
            res[env] = dict(W=eng.get('W'), V=eng.get('V'), Tau2=eng.get('Tau2'), nu2=eng.get_scalar('nu2'),
                            sigma2=eng.get_scalar('sigma2'), rows=eng.diag('row_stats'), cols=eng.diag('col_stats'))
        finally:
            eng.close()
            for k in (env,) + GEMM_MODES:
                os.environ.pop(k, None)
    a, b = res['BTF_STATS_FORCE_I8'], res['BTF_STATS_NO_I8']
    for k in ('rows', 'cols'):
        np.testing.assert_allclose(a[k], b[k], rtol=0, atol=1e-11 * np.abs(b[k]).max())
    for k in ('W', 'V', 'Tau2'):
        assert np.max(np.abs(a[k] - b[k])) <= 1e-8 * np.max(np.abs(b[k])), k
    assert a['nu2'] == pytest.approx(b['nu2'], rel=1e-9) and a['sigma2'] == pytest.approx(b['sigma2'], rel=1e-9)


def test_two_cta_gemm_split_k_and_fused_modes_are_exact():
    """300 rows x (64 x 128) columns at K = 16: the row contraction has 2 x 4 tiles of the 2-CTA GEMM, so it runs in its
    split-K mode (int32 atomics + recombination kernel); the column contraction has 128 tiles and runs with the fused
    epilogue.  Both must equal the fixed-point definition bit for bit (oracle.fixed_point.product_block_fast)."""
    from oracle.fixed_point import product_block_fast
    from functionalmf_b200 import _lib as L
    N, M, T, R, K = 300, 64, 128, 2, 16
    rs, W, V, Y = _problem(N, M, T, R, K, 31)
    eng = _engine(N, M, T, R, K, 'BTF_STATS_FORCE_I8')
    try:
        eng.set_data_gaussian(Y)
        _load(eng, rs, W, V)
        eng.enable_diag(True)
        eng.set_sample_mask(L.SAMPLE_W | L.SAMPLE_V)
        eng.inject('z_W', np.zeros(N * K))
        eng.sweep(1)
        Lp = K * (K + 1) // 2
        cnt = (~np.isnan(Y)).sum(axis=-1).reshape(N, M * T)
        np.testing.assert_array_equal(eng.diag('row_stats')[:, :Lp], product_block_fast((V * 1.1).reshape(M * T, K), cnt))
        np.testing.assert_array_equal(eng.diag('col_stats')[:, :Lp], product_block_fast(eng.get('W'), cnt.T))
    finally:
        eng.close()
        os.environ.pop('BTF_STATS_FORCE_I8', None)


def _exact_product_block(F, cnt):
    """sum_k cnt[m,k] F[k,k1] F[k,k2] in extended precision (packed lower triangle), as float64."""
    K = F.shape[1]
    il = np.tril_indices(K)
    Z = (F[:, il[0]].astype(np.longdouble) * F[:, il[1]].astype(np.longdouble))
    return (cnt.astype(np.longdouble) @ Z).astype(np.float64)


def _guard_case(no_guard):
    """Block-structured missingness + factor scales spanning 1e6: rows >= 200 observe only the
    columns whose V is 1e-6 times smaller, so their statistics sit 1e-12 below the column maxima that fix the
    fixed-point scale; same construction transposed for the column statistics."""
    from functionalmf_b200 import _lib as L
    N, M, T, R, K = 256, 8, 16, 2, 16
    rs = np.random.RandomState(4)
    V = rs.normal(size=(M, T, K)); V[4:] *= 1e-6
    W = rs.normal(size=(N, K)); W[128:] *= 1e-6
    Y = rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < 0.2] = np.nan
    Y[200:, :4] = np.nan                  # rows 200.. see only the small columns
    Y[:128, 6:] = np.nan                  # columns 6.. see only the small rows
    for k in ('BTF_STATS_FORCE_I8', 'BTF_STATS_NO_I8', 'BTF_I8_NO_GUARD') + GEMM_MODES:
        os.environ.pop(k, None)
    os.environ['BTF_STATS_FORCE_I8'] = '1'
    if no_guard:
        os.environ['BTF_I8_NO_GUARD'] = '1'
    from functionalmf_b200.engine import Engine
    eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=5, use_graph=0)
    try:
        eng.set_data_gaussian(Y)
        eng.init_state(127)
        eng.set('W', W); eng.set('V', V)
        for k, v in dict(lam2=0.7, sigma2=0.9, nu2=1.1).items():
            eng.set(k, [v])
        eng.enable_diag(True)
        cnt = (~np.isnan(Y)).sum(axis=-1).reshape(N, M * T)
        Lp = K * (K + 1) // 2
        dg = np.array([k * (k + 3) // 2 for k in range(K)])
        il = np.tril_indices(K)
        out = {}
        # columns first (V | rest with the W set above), then rows (W | rest with the V just drawn replaced by ours)
        eng.set_sample_mask(L.SAMPLE_V)
        eng.sweep(1)
        cols = eng.diag('col_stats')[:, :Lp]
        out['col_flagged'] = int(eng.diag('i8_guard')[1])
        want = _exact_product_block(W, cnt.T)
        out['col_rel'] = np.abs(cols[:, dg] / want[:, dg] - 1.0)
        sc = np.sqrt(want[:, dg][:, il[0]] * want[:, dg][:, il[1]])
        out['col_off'] = np.abs(cols - want) / sc
        eng.set('V', V)
        eng.set_sample_mask(L.SAMPLE_W)
        eng.sweep(1)
        rows = eng.diag('row_stats')[:, :Lp]
        out['row_flagged'] = int(eng.diag('i8_guard')[0])
        want = _exact_product_block(V.reshape(M * T, K), cnt)
        out['row_rel'] = np.abs(rows[:, dg] / want[:, dg] - 1.0)
        sc = np.sqrt(want[:, dg][:, il[0]] * want[:, dg][:, il[1]])
        out['row_off'] = np.abs(rows - want) / sc
        return out
    finally:
        eng.close()
        for k in ('BTF_STATS_FORCE_I8', 'BTF_I8_NO_GUARD'):
            os.environ.pop(k, None)


def test_elementwise_guard_keeps_badly_scaled_rows_accurate():
    """Element-wise (not normwise) accuracy of the integer path: every diagonal entry of every row / (j,t) statistic to
    1e-10 relative, every off-diagonal entry to 1e-10 of sqrt(d1 d2), although the fixed-point scale is set by column
    maxima 1e12 times larger.  The guard must list exactly the rows / columns that only see the small factors, and
    without it (BTF_I8_NO_GUARD=1) the same data must miss the bound - the test has teeth."""
    g = _guard_case(no_guard=False)
    assert g['row_flagged'] == 56 and g['col_flagged'] == 2 * 16          # rows 200..255; columns j = 6, 7 at every t
    assert g['row_rel'].max() < 1e-10 and g['col_rel'].max() < 1e-10
    assert g['row_off'].max() < 1e-10 and g['col_off'].max() < 1e-10
    bad = _guard_case(no_guard=True)
    assert bad['row_flagged'] == 0 and bad['col_flagged'] == 0
    assert bad['row_rel'].max() > 1e-8 and bad['col_rel'].max() > 1e-8
