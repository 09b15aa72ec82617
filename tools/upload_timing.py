"""How should a caller's ordinary (pageable) numpy array reach the device?  Times Engine.set_data_gaussian at the C2 shape
from (a) pageable memory as is, (b) the same array page-locked in place around the call (cudaHostRegister + unregister
included in the time), (c) a pinned allocation (what bench.py's e2e leg uses)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from functionalmf_b200.engine import Engine, pinned_empty
from functionalmf_b200 import _lib as L

N, M, T, R, K = 4096, 1024, 64, 3, 16
rs = np.random.default_rng(0)
Y = rs.standard_normal((N, M, T, R))
Y[rs.random(Y.shape) < 0.2] = np.nan
lib = L.load()
eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1)
eng.set_data_gaussian(Y)            # warm-up: staging buffers, kernels
def timed(f):
    t0 = time.perf_counter(); f(); return time.perf_counter() - t0
for rep in range(2):
    a = timed(lambda: eng.set_data_gaussian(Y))
    def reg():
        rc = lib.btf_host_register(C.c_void_p(Y.ctypes.data), Y.nbytes)
        eng.set_data_gaussian(Y)
        if rc == 0: lib.btf_host_unregister(C.c_void_p(Y.ctypes.data))
        return rc
    b = timed(reg)
    print('pageable %.3f s (%.1f GB/s)   register+copy+unregister %.3f s (%.1f GB/s)' % (a, Y.nbytes / a / 1e9, b, Y.nbytes / b / 1e9), flush=True)
P = pinned_empty(Y.shape); P[...] = Y
c = timed(lambda: eng.set_data_gaussian(P))
print('pinned   %.3f s (%.1f GB/s)' % (c, Y.nbytes / c / 1e9))
