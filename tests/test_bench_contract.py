"""The CPU-runnable parts of bench.py keep their contract (reference arm JSON line)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'c1',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'Gibbs sweeps/sec' and d['unit'] == 'sweeps/s'
    assert d['higher_is_better'] is True and d['value'] > 0
    # the unmodified reference from baseline/_ref when build() installed it there (kind "reference"), else the oracle port
    have_ref = os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'functionalmf', 'factor.py'))
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['config']['workload'] and 'model' not in d['config']
    assert d['e2e'] == {'value': d['value'], 'unit': 'sweeps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_workload_table_names_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    c2 = bench.WORKLOADS['c2']
    assert (c2['N'], c2['M'], c2['T'], c2['R'], c2['K'], c2['order'], c2['nan']) == (4096, 1024, 64, 3, 16, 2, 0.2)
    c5 = bench.WORKLOADS['c5']
    assert (c5['N'] * 8, c5['M'], c5['T'], c5['R'], c5['K']) == (65536, 8192, 128, 2, 32)
