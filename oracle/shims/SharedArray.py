"""Dict-backed stand-in for the SharedArray POSIX-shm extension (absent offline).

TEST INFRASTRUCTURE ONLY.  Only needed so ``import functionalmf.factor`` succeeds
(factor.py:20); the conjugate Gibbs path never touches shared memory.
"""
import numpy as np

_STORE = {}


def _key(name):
    return name[len('shm://'):] if name.startswith('shm://') else name


def create(name, shape, dtype=float):
    arr = np.zeros(shape, dtype=dtype)
    _STORE[_key(name)] = arr
    return arr


def attach(name):
    return _STORE[_key(name)]


def delete(name):
    _STORE.pop(_key(name), None)
