"""Host GASS logic (functionalmf_b200/gass.py) against recorded transitions of the reference's
gass() (oracle/make_golden_constrained.py -> tests/golden/gass_cases.npz).  CPU only."""
import os
import numpy as np

from golden_util import GOLDEN
from constrained_ll import ReplayRng


def unpack_choices(flat, lens):
    out, pos = [], 0
    for n in lens:
        n = int(n)
        if n == 0:
            out.append(int(flat[pos])); pos += 1
        else:
            out.append(np.array(flat[pos:pos + n])); pos += n
    return out


def test_gass_matches_reference_transitions():
    from functionalmf_b200.gass import gass
    z = np.load(os.path.join(GOLDEN, 'gass_cases.npz'))
    moved = 0
    for c in range(int(z['ncase'][0])):
        pre = 'c%d_' % c
        target = z[pre + 'target']

        def ll(pts, args):
            return -0.5 * ((pts - args) ** 2).sum(axis=-1) * 3.0
        rng = ReplayRng(z[pre + 'u'], unpack_choices(z[pre + 'c'], z[pre + 'clen']))
        xn, lln = gass(z[pre + 'x'].copy(), z[pre + 'v'], ll, z[pre + 'C'], mu=z[pre + 'mu'], ll_args=target,
                       ngrid=25, rng=rng)
        assert np.allclose(xn, z[pre + 'xn'], rtol=1e-12, atol=1e-13), c
        assert abs(lln - float(z[pre + 'lln'][0])) <= 1e-10 * max(1.0, abs(lln))
        assert not rng.u and not rng.c, 'random draws consumed in a different order than the reference'
        assert np.all(z[pre + 'C'][:, :-1].dot(xn) >= z[pre + 'C'][:, -1] - 1e-9)
        moved += int(not np.allclose(xn, z[pre + 'x']))
    assert moved >= 8


def test_feasible_angles_cover_whole_ellipse_when_unconstrained():
    from functionalmf_b200.gass import feasible_angles
    a, b, c = np.array([1.0, 2.0]), np.array([0.1, -0.2]), np.array([-5.0, -9.0])
    g = feasible_angles(a, b, c, 17)
    assert len(g) == 17 and np.isclose(g[0], -np.pi) and np.isclose(g[-1], np.pi)


def test_elliptical_slice_matches_reference_transitions():
    """functionalmf_b200/ess.py against recorded calls of the reference's elliptical_slice_."""
    from functionalmf_b200.ess import elliptical_slice
    z = np.load(os.path.join(GOLDEN, 'ess_cases.npz'))

    class Replay(object):
        def __init__(self, u):
            self.u = list(u)

        def rand(self):
            return self.u.pop(0)

    for c in range(int(z['ncase'][0])):
        pre = 'e%d_' % c
        target = z[pre + 'target']

        def ll(pts, args):
            return -2.0 * ((pts - args) ** 2).sum()
        rng = Replay(z[pre + 'u'])
        xn, lln = elliptical_slice(z[pre + 'x'].copy(), z[pre + 'nu'], ll, ll_args=target, mu=z[pre + 'mu'], rng=rng)
        assert not rng.u
        assert np.allclose(xn, z[pre + 'xn'], rtol=1e-12, atol=1e-14), c
        assert abs(lln - float(z[pre + 'lln'][0])) < 1e-10 * max(1.0, abs(lln))
