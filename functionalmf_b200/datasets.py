"""Loaders for the data files the reference's application scripts read.

The files themselves are not part of this package: pass the directory that holds them
(``politics/`` or ``flutrends/`` of a functionalmf checkout).  Each loader returns the arrays
in the layout the scripts build before constructing the model, plus the held-out classes
``HeldOutEvaluator`` takes.
"""
import os
import numpy as np

from .metrics import heldout_classes


def load_politics(path):
    """GDELT G20 "intend to cooperate" counts (politics/benchmark.py:60-62, create_datasets.py:66-80).

    Returns dict(Y [19,19,228] with NaN on the diagonal pairs, Y_train (10 % of the nation pairs
    NaN), held_out [npairs, 2], classes uint8, dates, nations)."""
    Y = np.load(os.path.join(path, 'cooperate.npy'))
    Y_train = np.load(os.path.join(path, 'cooperate_train.npy'))
    out = dict(Y=Y, Y_train=Y_train, classes=heldout_classes(Y, Y_train))
    for key, fname in (('held_out', 'held_out.npy'), ('dates', 'dates.npy'), ('nations', 'nations.npy')):
        f = os.path.join(path, fname)
        if os.path.exists(f):
            out[key] = np.load(f)
    return out


def load_flu_states(path, log=True):
    """Google Flu Trends, the 50 state series (flutrends/create_datasets.py:15-39 and
    flutrends/benchmark.py:19-25): columns 1..50 of ``flu_US.mat``'s ``data``, the year blocks
    listed in ``held_out_years.npy`` ([state, week_start, week_end)) set to NaN for training,
    log-transformed, laid out [state, 1, week].

    Returns dict(Y, Y_train, held_out, classes, dates, names)."""
    from scipy.io import loadmat
    df = loadmat(os.path.join(path, 'flu_US.mat'))
    data = np.array(df['data'][:, 1:51], dtype=np.float64)
    held = np.load(os.path.join(path, 'held_out_years.npy'))
    train = data.copy()
    for i, j, k in held:
        train[j:k, i] = np.nan
    with np.errstate(divide='ignore', invalid='ignore'):
        Y = np.log(data.T[:, None]) if log else data.T[:, None].copy()
        Y_train = np.log(train.T[:, None]) if log else train.T[:, None].copy()
    out = dict(Y=Y, Y_train=Y_train, held_out=held, classes=heldout_classes(Y, Y_train))
    if 'dates' in df:
        out['dates'] = np.array([str(x[0][0]) for x in df['dates']])
    if 'USnames' in df:
        out['names'] = np.array([str(x[0][0]) for x in df['USnames'][1:51]])
    return out
