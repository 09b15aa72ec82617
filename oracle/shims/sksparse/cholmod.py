"""Stand-in for ``sksparse.cholmod`` so the unmodified reference can be imported.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (fast_mvn.py:38-47, factor.py:789-795) only relies on the
*semantics* ``L L^T = P Q P^T`` of CHOLMOD, never on its particular fill-reducing
permutation: the draw is ``solve_Lt(z)[argsort(P())] + solve_A(mu_part)``.
This shim therefore fixes the permutation to the k-major -> t-major shuffle
(``P[t*K + k] = k*T + t``) when a layout is registered with ``set_layout(K, T)``,
and factorises the permuted matrix with LAPACK's banded Cholesky
(``scipy.linalg.cholesky_banded``, half-bandwidth auto-detected).  With that
choice the reference consumes its standard-normal vector ``z`` in the same
t-major order as the CUDA engine, which is what makes 1e-10 parity of the draws
meaningful.  Without a registered layout the natural order is used.

``RECORD`` (a list, or None) receives ``(Q_permuted_dense, L_dense, P)`` for
every successful factorisation so golden fixtures can include the factors.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sps

_LAYOUT = None      # (K, T) or None
RECORD = None       # list or None


class CholmodError(Exception):
    pass


class CholmodNotPositiveDefiniteError(CholmodError):
    pass


def set_layout(K=None, T=None):
    """Register the (nembeds, ndepth) layout of the k-major systems to come."""
    global _LAYOUT
    _LAYOUT = None if K is None else (int(K), int(T))


def _perm(n):
    if _LAYOUT is not None:
        K, T = _LAYOUT
        if K * T == n:
            return np.arange(n).reshape(K, T).T.ravel()
    return np.arange(n)


class Factor(object):
    def __init__(self, Q):
        Qd = Q.toarray() if sps.issparse(Q) else np.asarray(Q, dtype=float)
        n = Qd.shape[0]
        self._P = _perm(n)
        Qp = Qd[np.ix_(self._P, self._P)]
        # half-bandwidth of the permuted matrix
        nz = np.nonzero(Qp)
        kd = int(np.max(np.abs(nz[0] - nz[1]))) if len(nz[0]) else 0
        ab = np.zeros((kd + 1, n))
        for d in range(kd + 1):
            ab[d, :n - d] = np.diagonal(Qp, -d)
        try:
            cb = sla.cholesky_banded(ab, lower=True, check_finite=False)
        except np.linalg.LinAlgError as exc:
            raise CholmodNotPositiveDefiniteError(str(exc))
        if not np.all(np.isfinite(cb[0])) or np.any(cb[0] <= 0):
            raise CholmodNotPositiveDefiniteError('non-positive pivot')
        self._kd, self._n, self._cb = kd, n, cb
        # upper-banded storage of L^T for solve_banded
        ub = np.zeros_like(cb)
        for d in range(kd + 1):
            ub[kd - d, d:] = cb[d, :n - d]
        self._ub = ub
        if RECORD is not None:
            RECORD.append((Qp.copy(), self._dense_L(), self._P.copy()))

    def _dense_L(self):
        L = np.zeros((self._n, self._n))
        for d in range(self._kd + 1):
            idx = np.arange(self._n - d)
            L[idx + d, idx] = self._cb[d, :self._n - d]
        return L

    def P(self):
        return self._P

    def L(self):
        return sps.csc_matrix(self._dense_L())

    def solve_Lt(self, b, use_LDLt_decomposition=True):
        return sla.solve_banded((0, self._kd), self._ub, np.asarray(b, dtype=float),
                                check_finite=False)

    def solve_L(self, b, use_LDLt_decomposition=True):
        return sla.solve_banded((self._kd, 0), self._cb, np.asarray(b, dtype=float),
                                check_finite=False)

    def solve_A(self, b):
        b = np.asarray(b, dtype=float)
        xp = sla.cho_solve_banded((self._cb, True), b[self._P], check_finite=False)
        x = np.empty_like(xp)
        x[self._P] = xp
        return x

    __call__ = solve_A


def cholesky(A, beta=0, mode='auto', ordering_method='default', use_long=None):
    return Factor(A)
