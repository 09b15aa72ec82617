"""Host helpers kept from functionalmf/utils.py (metrics and the penalty matrix)."""
import numpy as np


def ilogit(x):
    """Inverse logit (utils.py:106-107)."""
    return 1 / (1 + np.exp(-x))


def mse(x, y):
    """Mean squared error ignoring NaNs (utils.py:109-110)."""
    return np.nanmean((x - y) ** 2)


def mae(x, y):
    """Mean absolute error ignoring NaNs (utils.py:112-113)."""
    return np.nanmean(np.abs(x - y))


def bayes_grid_penalty(ndepth, k, anchor=0):
    """Bayesian trend-filtering matrix Delta as scipy CSC (utils.py:83-90, 1-D grids only).

    Rows: the anchor e_anchor^T, then the difference operators of order 0..k built
    by alternately applying D^T and D to the first-difference matrix D.
    """
    from scipy.sparse import csc_matrix
    T = int(ndepth[0]) if hasattr(ndepth, '__len__') else int(ndepth)
    D = np.zeros((T - 1, T))
    i = np.arange(T - 1)
    D[i, i], D[i, i + 1] = -1.0, 1.0
    rows = [np.zeros((1, T))]
    rows[0][0, anchor] = 1.0
    for order in range(k + 1):
        cur = D
        for s in range(order):
            cur = D.T.dot(cur) if s % 2 == 0 else D.dot(cur)
        rows.append(cur)
    return csc_matrix(np.concatenate(rows, axis=0))
