"""Band-solve timing over shapes.  With the instrumented library (tools/build_band_profile.sh, BTF_B200_LIB=...) the
look-ahead kernel also prints per-phase clock64 totals of column 0 (init | B | C1 | C2 (own ...) | backward | resid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functionalmf_b200.engine import Engine
rs = np.random.RandomState(0)
shapes = [(256, 1024, 64, 16), (256, 148, 64, 16), (256, 256, 128, 32), (256, 148, 128, 32), (64, 19, 228, 10)]
for (N, M, T, K) in shapes:
    W = rs.normal(size=(N, K)); V = rs.normal(size=(M, T, K)).cumsum(axis=1) * 0.3
    Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + rs.normal(size=(N, M, T, 1))
    eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1, use_graph=0)
    eng.set_data_gaussian(Y); eng.init_state(127); eng.set('sigma2', [0.5]); eng.set('lam2', [0.1]); eng.set('nu2', [1.0])
    eng.sweep(2)
    print('shape', N, M, T, K, flush=True)
    eng.sweep(1); eng.synchronize()
    ph = eng.time_phases(3)
    print('  band_solve %.3f ms (%.2f us per block step per wave-column)' % (ph['band_solve'], ph['band_solve'] * 1e3 / T), flush=True)
    eng.close()
