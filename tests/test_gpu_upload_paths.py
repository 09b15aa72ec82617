"""The three ways a Gaussian data tensor reaches the device - ordinary numpy memory (threaded copy through pinned bounce
buffers), a pinned host array, and the plain driver copy (BTF_UPLOAD_BOUNCE=0) - must leave the same pre-reduced data:
one sweep from the same seed gives bit-identical statistics and state.  Replaces the per-call data load of
genlasso.py:31-66 / factor.py:287-304 (the reference keeps Y on the host)."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _one_sweep(Y, N, M, T, K):
    from functionalmf_b200.engine import Engine
    eng = Engine(N, M, T, nembeds=K, tf_order=1, seed=3)
    eng.set_data_gaussian(Y)
    eng.init_state(127)
    eng.sweep(1)
    out = (eng.diag('row_stats').copy(), eng.diag('col_stats').copy(), eng.get('W').copy(), eng.get('V').copy(),
           eng.diag('nu2_rate').copy())
    eng.close()
    return out


def test_pageable_pinned_and_plain_uploads_agree(monkeypatch):
    from functionalmf_b200.engine import pinned_empty
    rs = np.random.RandomState(0)
    # rows of 19 * 7 * 2 doubles, more than one 64 MB piece would need ~ 250k rows: force several pieces with a ragged tail
    N, M, T, R, K = 301, 19, 7, 2, 4
    Y = rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < 0.25] = np.nan
    ref = _one_sweep(Y, N, M, T, K)                      # pageable -> bounce buffers
    P = pinned_empty(Y.shape)
    P[...] = Y
    pin = _one_sweep(P, N, M, T, K)                      # pinned -> direct asynchronous copies
    monkeypatch.setenv('BTF_UPLOAD_BOUNCE', '0')
    plain = _one_sweep(Y, N, M, T, K)                    # pageable -> driver copy
    monkeypatch.setenv('BTF_UPLOAD_BOUNCE', '1')
    monkeypatch.setenv('BTF_UPLOAD_THREADS', '3')
    thr = _one_sweep(Y, N, M, T, K)
    for other in (pin, plain, thr):
        for a, b in zip(ref, other):
            assert np.array_equal(a, b)


def test_bounce_upload_in_many_pieces(monkeypatch):
    """A tensor whose upload takes several bounce pieces (piece size is 64 MB) with a ragged last piece."""
    rs = np.random.RandomState(1)
    N, M, T, R, K = 1100, 64, 128, 2, 8                  # 144 MB, rows of 128 KB
    Y = rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < 0.1] = np.nan
    a = _one_sweep(Y, N, M, T, K)
    monkeypatch.setenv('BTF_UPLOAD_BOUNCE', '0')
    b = _one_sweep(Y, N, M, T, K)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
