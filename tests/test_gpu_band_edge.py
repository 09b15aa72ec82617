"""Edge shapes of the band solve (T down to tf_order + 2, i.e. fewer block rows than the window, one column,
K = 3 ... 32, orders 0 ... 3): the look-ahead kernel against the scalar kernel, same injected noise.  The kernel
choice is read once per process, so tools/band_edge_check.py runs both in child processes."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lookahead_equals_scalar_on_edge_shapes():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'band_edge_check.py')], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'BAND_EDGE PASS' in r.stdout, r.stdout[-2000:]
