import numpy as np, sys
sys.path.insert(0, '/root/repo')
from functionalmf_b200.engine import Engine
rs = np.random.RandomState(0)
for (N, M, T, K) in [(4096, 19, 64, 16), (64, 19, 64, 16), (64, 148, 228, 16), (2048, 19, 228, 16), (64, 19, 228, 16), (64, 19, 228, 10)]:
    W = rs.normal(size=(N, K)); V = rs.normal(size=(M, T, K)).cumsum(axis=1) * 0.3
    Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + rs.normal(size=(N, M, T, 1))
    eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1, use_graph=0)
    eng.set_data_gaussian(Y); eng.init_state(127); eng.set('sigma2', [0.5]); eng.set('lam2', [0.1]); eng.set('nu2', [1.0])
    eng.sweep(2)
    print('shape', N, M, T, K, flush=True)
    eng.sweep(1); eng.synchronize()
    print(eng.time_phases(3), flush=True)
    eng.close()
