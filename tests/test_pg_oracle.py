"""Moment checks of the CPU Polya-Gamma restatement (oracle/pg.py)."""
import numpy as np
import pytest
from oracle import pg


@pytest.mark.parametrize('b', [1, 3, 2.5, 0.4, 250.0])
@pytest.mark.parametrize('z', [0.0, 1.0, 4.0])
def test_pg_moments(b, z):
    rng = np.random.default_rng(int(b * 10 + z))
    n = 30000
    x = pg.pgdraw(np.full(n, float(b)), np.full(n, z), rng)
    m, v = float(pg.pg_mean(b, z)), float(pg.pg_var(b, z))
    assert abs(x.mean() - m) < 5 * np.sqrt(v / n)
    assert abs(x.var() / v - 1) < 0.08


def test_pg_zero_for_missing():
    rng = np.random.default_rng(0)
    x = pg.pgdraw(np.array([np.nan, 0.0, -2.0]), np.array([0.3, 0.3, 0.3]), rng)
    assert np.all(x == 0)
