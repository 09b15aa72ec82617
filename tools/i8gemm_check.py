#!/usr/bin/env python
"""Correctness + throughput of the tcgen05 int8 GEMM (btf_i8gemm_test) against numpy."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functionalmf_b200 import _lib as L

lib = L.load()
rs = np.random.RandomState(0)
ok = True
for (M, N, K) in [(128, 256, 128), (128, 256, 512), (200, 300, 1024), (1152, 4096, 4096)]:
    A = rs.randint(-64, 64, size=(M, K)).astype(np.int8)
    B = rs.randint(0, 4, size=(N, K)).astype(np.int8)
    D = np.zeros((M, N), dtype=np.int32)
    ms = lib.btf_i8gemm_test(0, C.c_void_p(A.ctypes.data), C.c_void_p(B.ctypes.data), C.c_void_p(D.ctypes.data), M, N, K, 3)
    ref = A.astype(np.int32) @ B.astype(np.int32).T
    bad = int((D != ref).sum())
    ok = ok and bad == 0 and ms >= 0
    print('M %d N %d K %d: ms/launch %.4f  mismatches %d / %d  (TOPS %.1f)' % (M, N, K, ms / 3, bad, D.size, 2.0 * M * N * K / (ms / 3 * 1e-3) / 1e12 if ms > 0 else 0), flush=True)
    if bad:
        idx = np.argwhere(D != ref)[:5]
        print('  first mismatches', [(int(i), int(j), int(D[i, j]), int(ref[i, j])) for i, j in idx])
print('I8GEMM', 'PASS' if ok else 'FAIL')
