"""Stand-in for the ``pypolyagamma`` package (absent offline; factor.py:431-432, 459).

TEST INFRASTRUCTURE ONLY.  ``PyPolyaGamma(seed).pgdrawv(n, z, out)`` fills ``out``
with PG(n, z) draws from the CPU restatement in ``oracle/pg.py``; non-finite or
non-positive ``n`` (the reference passes NaN for missing cells, factor.py:459)
yields 0.  ``RECORD`` (list or None) receives a copy of every ``out`` so golden
fixtures can replay identical omega on the CUDA side; ``REPLAY`` (list or None)
supplies pre-recorded draws instead of sampling.
"""
import os
import sys
import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(os.path.dirname(_here))
if _root not in sys.path:
    sys.path.insert(0, _root)
from oracle import pg as _pg     # noqa: E402

RECORD = None
REPLAY = None


class PyPolyaGamma(object):
    def __init__(self, seed=0, trunc=200):
        self._rng = np.random.default_rng(seed)

    def pgdraw(self, n, z):
        return float(_pg.pgdraw(np.array([n]), np.array([z]), self._rng)[0])

    def pgdrawv(self, n, z, out):
        if REPLAY is not None:
            out[:] = REPLAY.pop(0)
        else:
            out[:] = _pg.pgdraw(n, z, self._rng)
        if RECORD is not None:
            RECORD.append(np.array(out, copy=True))
