"""Debug aid: run the i8 GPU tests in this process, then compare the C2 row / column product blocks of the integer path
with the FP64 statistics kernels and print where they differ."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytest
if len(sys.argv) > 1 and sys.argv[1] == 'poison':
    pytest.main(['tests/test_gpu_i8.py', '-x', '-q', '-m', 'gpu', '-p', 'no:cacheprovider'])
import torch
from functionalmf_b200.engine import Engine
N, M, T, R, K = 4096, 1024, 64, 3, 16
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(5)
V0 = (torch.randn(M, T, K, generator=g, device=dev, dtype=torch.float64) * 0.3).cumsum(1)
pieces = []
for a in range(0, N, 256):
    W = torch.randn(256, K, generator=g, device=dev, dtype=torch.float64)
    Y = (W @ V0.reshape(M * T, K).T).reshape(256, M, T, 1) + torch.randn(256, M, T, R, generator=g, device=dev, dtype=torch.float64)
    Y[torch.rand(Y.shape, generator=g, device=dev) < 0.2] = float('nan')
    pieces.append(Y)
torch.cuda.synchronize()
Lp = K * (K + 1) // 2
res = {}
MODES = os.environ.get('DBG_MODES', 'i8,fp64,i8b,i8c').split(',')
USE_GRAPH = int(os.environ.get('DBG_GRAPH', '1'))
for mode in MODES:
    os.environ.pop('BTF_STATS_NO_I8', None)
    if mode == 'fp64':
        os.environ['BTF_STATS_NO_I8'] = '1'
    eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=11, use_graph=USE_GRAPH)
    for i, Y in enumerate(pieces):
        eng.set_data_gaussian_rows_device(Y.data_ptr(), i * 256, 256, R, i == 0)
    eng.init_state(127)
    eng.set('sigma2', [0.5]); eng.set('lam2', [0.1]); eng.set('nu2', [1.0])
    eng.set_sample_mask(32)
    try:
        eng.sweep(1)
    except Exception as exc:
        print(mode, 'V sweep:', str(exc)[:80])
    c = eng.diag('col_stats')[:, :Lp].copy()
    eng.set_sample_mask(16)
    try:
        eng.sweep(1)
    except Exception as exc:
        print(mode, 'W sweep:', str(exc)[:80])
    r = eng.diag('row_stats')[:, :Lp].copy()
    res[mode] = (r, c)
    eng.close()
ref = res['fp64'] if 'fp64' in res else res['i8']
for mode in [m for m in MODES if m != 'fp64']:
    for nm, a, b in (('rows', res[mode][0], ref[0]), ('cols', res[mode][1], ref[1])):
        err = np.abs(a - b) / np.abs(b).max()
        err[~np.isfinite(err)] = np.inf
        bad = np.argwhere(err > 1e-9)
        print(mode, nm, 'max err %.3e' % err.max(), 'bad entries', len(bad))
        if len(bad):
            rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1])
            print('   bad rows: n=%d min=%d max=%d  mod256 range [%d, %d]  first %s' % (len(rows), rows.min(), rows.max(), (rows % 256).min(), (rows % 256).max(), rows[:12]))
            print('   bad cols: n=%d %s' % (len(cols), cols[:40]))
            i, j = bad[0]
            print('   sample got %.6e want %.6e' % (a[i, j], b[i, j]))
