#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize_oracle.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5
# narrow-shard shape on one GPU: 512 rows x C2 columns (what every rank of the 8-GPU C2 run sees for its W step)
python - <<'PY'
import numpy as np, os, sys, time
sys.path.insert(0, '.')
from functionalmf_b200.engine import Engine
for env in ('1', '0'):
    os.environ['BTF_I8_G2_SPLITK'] = env
    import subprocess
    code = '''
import numpy as np, sys
sys.path.insert(0, '.')
from functionalmf_b200.engine import Engine
rs = np.random.RandomState(0)
N, M, T, R, K = 512, 1024, 64, 3, 16
Y = rs.normal(size=(N, M, T, R)); Y[rs.random_sample(Y.shape) < 0.2] = np.nan
eng = Engine(N, M, T, nembeds=K, tf_order=2, seed=1)
eng.set_data_gaussian(Y); eng.init_state(127); eng.set('sigma2', [0.5]); eng.set('lam2', [0.1]); eng.set('nu2', [1.0])
eng.sweep(3)
ph = eng.time_phases(3)
print('splitk env', __import__('os').environ.get('BTF_I8_G2_SPLITK'), {k: round(v, 3) for k, v in ph.items() if 'row' in k or 'col' in k}, 'ms/sweep', round(eng.sweep_timed(10) / 10, 3))
'''
    print(subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300).stdout.strip())
PY
