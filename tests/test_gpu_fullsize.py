"""Full-size (BASELINE.json configs[1]: 4096 x 1024 x 64 x 3, K = 16) property checks that do not
need an oracle run: the row and the column contractions are two factorizations of the same sums,
    sum_i w_i^T A_i(V) w_i = sum_p v_p^T A~_p(W) v_p = sum cnt * Mu^2
    sum_i w_i . b_i(V)     = sum_p v_p . b~_p(W)     = sum S * Mu
("checksum of checksums" tying K1a to K1b), the pre-reduction totals match a direct reduction of
the raw tensor, and the streaming upload equals the one-shot upload."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _quad_packed(stats, X, K):
    """sum_rows x^T A x - style contraction for packed-lower statistics [rows, L+K]."""
    Lp = K * (K + 1) // 2
    il = np.tril_indices(K)
    wgt = np.where(il[0] == il[1], 1.0, 2.0)
    quad = (stats[:, :Lp] * (X[:, il[0]] * X[:, il[1]] * wgt)).sum()
    lin = (stats[:, Lp:] * X).sum()
    return quad, lin


def test_c2_row_and_column_statistics_agree():
    import torch
    from functionalmf_b200.engine import Engine
    N, M, T, R, K, order = 4096, 1024, 64, 3, 16, 2
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    eng = Engine(N, M, T, nembeds=K, tf_order=order, seed=11)
    V0 = (torch.randn(M, T, K, generator=g, device=dev, dtype=torch.float64) * 0.3).cumsum(1)
    n_obs, ss = 0, 0.0
    piece = 256
    for a in range(0, N, piece):
        W = torch.randn(piece, K, generator=g, device=dev, dtype=torch.float64)
        Y = (W @ V0.reshape(M * T, K).T).reshape(piece, M, T, 1) + \
            torch.randn(piece, M, T, R, generator=g, device=dev, dtype=torch.float64)
        Y[torch.rand(Y.shape, generator=g, device=dev) < 0.2] = float('nan')
        obs = ~torch.isnan(Y)
        n_obs += int(obs.sum().item())
        ss += float((torch.nan_to_num(Y) ** 2).sum().item())
        torch.cuda.synchronize()
        eng.set_data_gaussian_rows_device(Y.data_ptr(), a, piece, R, a == 0)
        del Y, W, obs
    # pre-reduction totals (K0) against the direct reduction of the raw tensor
    assert eng.get_scalar('n_obs') == n_obs
    assert abs(eng.get_scalar('ss_total') / ss - 1.0) < 1e-12
    eng.init_state(127)
    eng.set('sigma2', [0.5]); eng.set('lam2', [0.1]); eng.set('nu2', [1.0])
    W0 = eng.get('W')
    eng.set_sample_mask(32)            # V | rest only: column statistics of W0, V0 -> V1
    eng.sweep(1)
    V1 = eng.get('V')
    cstat = eng.diag('col_stats')
    eng.set_sample_mask(16)            # W | rest only: row statistics of V1
    eng.sweep(1)
    rstat = eng.diag('row_stats')
    assert np.all(np.isfinite(cstat)) and np.all(np.isfinite(rstat)) and np.all(np.isfinite(eng.get('W')))
    qr, lr = _quad_packed(rstat, W0, K)
    qc, lc = _quad_packed(cstat, V1.reshape(M * T, K), K)
    assert abs(qr / qc - 1.0) < 1e-10, (qr, qc)
    assert abs(lr / lc - 1.0) < 1e-10, (lr, lc)
    # a few free-running sweeps at full size stay finite and keep W lower triangular
    eng.set_sample_mask(127)
    eng.sweep(3)
    Wn = eng.get('W')
    assert np.all(np.isfinite(Wn)) and np.all(np.isfinite(eng.get('V')))
    assert np.all(Wn[np.triu_indices(K, k=1)] == 0)
    assert 0.5 < eng.get_scalar('nu2') < 2.0        # data were generated with unit noise
    eng.close()


def test_streaming_upload_equals_one_shot():
    from functionalmf_b200.engine import Engine
    import ctypes as C
    from functionalmf_b200 import _lib as L
    rs = np.random.RandomState(8)
    N, M, T, R, K = 300, 7, 9, 2, 4
    Y = rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < 0.3] = np.nan
    outs = []
    for pieces in (None, [0, 1, 130, 257, 300]):
        eng = Engine(N, M, T, nembeds=K, tf_order=1, seed=3)
        if pieces is None:
            eng.set_data_gaussian(Y)
        else:
            for a, b in zip(pieces[:-1], pieces[1:]):
                blk = np.ascontiguousarray(Y[a:b])
                L.check(eng.lib.btf_set_data_gaussian_rows(eng._h, C.c_void_p(blk.ctypes.data), a, b - a, R, int(a == 0)))
        eng.init_state(127)
        eng.sweep(2)
        outs.append((eng.get('W'), eng.get('V'), eng.get_scalar('nu2'), eng.get_scalar('n_obs')))
        eng.close()
    assert outs[0][3] == outs[1][3] == float((~np.isnan(Y)).sum())
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-12, atol=1e-13)
    assert np.allclose(outs[0][1], outs[1][1], rtol=1e-12, atol=1e-13)
    assert abs(outs[0][2] / outs[1][2] - 1) < 1e-12
