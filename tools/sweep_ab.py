"""A/B timing of the C2 sweep under different environment switches (one subprocess per configuration; data generated
on the device, so a configuration costs a few seconds).  usage: sweep_ab.py NAME:VAR=VAL,VAR=VAL ..."""
import os, subprocess, sys

CHILD = r'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(sys.argv[1]))))
import torch
from functionalmf_b200.engine import Engine
N, M, T, R, K = 4096, 1024, 64, 3, 16
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(5)
V0 = (torch.randn(M, T, K, generator=g, device=dev, dtype=torch.float64) * 0.3).cumsum(1)
eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=11)
for i, a in enumerate(range(0, N, 256)):
    W = torch.randn(256, K, generator=g, device=dev, dtype=torch.float64)
    Y = (W @ V0.reshape(M * T, K).T).reshape(256, M, T, 1) + torch.randn(256, M, T, R, generator=g, device=dev, dtype=torch.float64)
    Y[torch.rand(Y.shape, generator=g, device=dev) < 0.2] = float('nan')
    torch.cuda.synchronize()
    eng.set_data_gaussian_rows_device(Y.data_ptr(), a, 256, R, i == 0)
    del Y
eng.init_state(127)
eng.sweep(5)
best = min(eng.sweep_timed(20) / 20 for _ in range(3))
print('%.4f ms/sweep  %.1f sweeps/s' % (best, 1e3 / best))
'''

if __name__ == '__main__':
    for spec in sys.argv[1:]:
        name, _, envs = spec.partition(':')
        env = dict(os.environ)
        for kv in filter(None, envs.split(',')):
            k, _, v = kv.partition('=')
            env[k] = v
        r = subprocess.run([sys.executable, '-c', CHILD, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=600)
        print(name.ljust(14), r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
