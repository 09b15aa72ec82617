"""Helpers to read the golden fixtures written by oracle/make_golden.py."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

GAUSS_CASES = ['gauss_small', 'gauss_reps', 'gauss_p0', 'gauss_k8', 'gauss_square']
BINOM_CASES = ['binom_small']
NEGBIN_CASES = ['negbin_all', 'negbin_rows']


class Case(object):
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + '.npz'))
        d = self.z['cfg/dims']
        self.N, self.M, self.T, self.R, self.K, self.order = [int(x) for x in d]
        self.Delta = self.z['cfg/Delta']
        self.nsweeps = len([k for k in self.z.files if k.endswith('/end/W')])

    def scalar(self, key):
        return float(np.ravel(self.z[key])[0])

    def state(self, tag):
        st = {}
        for name in ('W', 'V', 'Tau2', 'Tau2_a', 'Tau2_b', 'Tau2_c'):
            st[name] = self.z['%s/%s' % (tag, name)]
        for name in ('lam2', 'lam2_a', 'sigma2'):
            st[name] = self.scalar('%s/%s' % (tag, name))
        nu2 = self.z['%s/nu2' % tag]
        st['nu2'] = float(nu2) if nu2.ndim == 0 else nu2
        if ('%s/R' % tag) in self.z.files:
            st['R'] = self.z['%s/R' % tag]
        return st

    def noise(self, s):
        pre = 's%d/noise/' % s
        out = {}
        for k in self.z.files:
            if k.startswith(pre):
                v = self.z[k]
                name = k[len(pre):]
                out[name] = float(v[0]) if name in ('g_nu2', 'g_sigma2') else v
        return out

    def cfg(self):
        return dict(K=self.K, order=self.order, Delta=self.Delta,
                    nu2_a=self.scalar('cfg/nu2_a'), nu2_b=self.scalar('cfg/nu2_b'),
                    sigma2_a=self.scalar('cfg/sigma2_a'), sigma2_b=self.scalar('cfg/sigma2_b'),
                    stability=self.scalar('cfg/stability'), force_psd=True,
                    force_psd_eps=1e-6, force_psd_attempts=4, ref_compat=True)


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def normerr(a, b):
    """max |a-b| / max |b| (norm-wise relative error)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0
