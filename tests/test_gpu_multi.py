"""Sharded sweep == single-GPU sweep, run by the test suite whenever the box has >= 2 GPUs (skipped otherwise).

Launches tools/multi_gpu_check.py under torch.distributed.run on 2 devices, on the FP64 statistics path and on
the (forced) integer-tensor-core path with its column-sharded product block: W, V, Tau2, nu2, sigma2, lam2 and the
held-out scores after 3 sweeps equal the single-GPU engine's to 1e-8 normwise (measured 1e-13), and a model built
without seed= runs the same chain on every rank (seed agreement)."""
import os
import subprocess
import sys

import pytest

from gpu_util import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(not have_gpu() or _ngpu() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('force_i8', [False, True])
def test_sharded_equals_single_gpu(force_i8):
    env = dict(os.environ)
    env.pop('BTF_STATS_NO_I8', None)
    if force_i8:
        env['BTF_STATS_FORCE_I8'] = '1'
    else:
        env.pop('BTF_STATS_FORCE_I8', None)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29531' if force_i8 else '29530',
           os.path.join(ROOT, 'tools', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0, tail
    assert 'MULTI_GPU_CHECK PASS' in out.stdout, tail
    assert 'MISMATCH' not in out.stdout and 'RANKS DIFFER' not in out.stdout, tail


@pytest.mark.skipif(not have_gpu() or _ngpu() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('force_i8', [False, True])
def test_engines_on_two_devices_of_one_process(force_i8, monkeypatch):
    """One process, an engine on device 0 and then one on device 1 (kernel attributes such as the opt-in shared-memory
    size belong to a device's context and must be set on each): same seed and data -> identical chains."""
    import numpy as np
    from functionalmf_b200.engine import Engine
    if force_i8:
        monkeypatch.setenv('BTF_STATS_FORCE_I8', '1')
    else:
        monkeypatch.delenv('BTF_STATS_FORCE_I8', raising=False)
    rs = np.random.RandomState(4)
    N, M, T, R, K = 160, 24, 40, 2, 8          # T K doubles of band workspace: beyond the 48 KB default of a kernel
    Y = rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < 0.2] = np.nan
    out = []
    for dev in (0, 1):
        eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=9, device=dev)
        eng.set_data_gaussian(Y)
        eng.init_state(127)
        eng.sweep(3)
        out.append((eng.get('W').copy(), eng.get('V').copy(), eng.get('Tau2').copy()))
        eng.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)
