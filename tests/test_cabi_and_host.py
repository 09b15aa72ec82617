"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares;
host-side logic (sharding, run_gibbs segment schedule, penalty matrix)."""
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'btf_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(btf_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    from functionalmf_b200 import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), 'libbtf_b200.so does not export %s' % n
    # and the ctypes table binds exactly the declared interface
    assert sorted(_lib.SIGNATURES) == names


def test_config_struct_matches_default():
    import ctypes as C
    from functionalmf_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config()
    lib.btf_config_default(C.byref(cfg))
    assert (cfg.nembeds, cfg.tf_order) == (5, 2)                    # factor.py:25
    assert (cfg.sigma2_a, cfg.sigma2_b, cfg.nu2_a, cfg.nu2_b) == (0.1, 0.1, 0.1, 0.1)
    assert cfg.stability == 1e-6 and cfg.force_psd == 1 and cfg.force_psd_attempts == 4
    assert cfg.force_psd_eps == 1e-6
    assert (cfg.nmetropolis, cfg.rpropstdev, cfg.rstdev, cfg.rdims_mask) == (30, 0.1, 1.0, 7)   # factor.py:467-470
    assert cfg.sample_mask == _lib.SAMPLE_ALL and cfg.world_size == 1 and cfg.use_graph == 1


def test_no_cpu_fallback_without_gpu():
    """Creating a model without a CUDA device must fail loudly, never fall back."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('a GPU is present')
    except ImportError:
        pass
    import functionalmf_b200 as F
    with pytest.raises((F.BTFError, F.BTFLibraryError)):
        F.GaussianBayesianTensorFiltering(4, 3, 5, nembeds=2)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'functionalmf_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f


def test_bayes_grid_penalty_matches_reference_fixture():
    from functionalmf_b200.utils import bayes_grid_penalty
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'delta.npz'))
    for key in z.files:
        T, k = [int(s[1:]) for s in key.split('_')]
        assert np.array_equal(bayes_grid_penalty(T, k).toarray(), z[key]), key


def test_partition_and_shard():
    from functionalmf_b200.distributed import partition, Shard
    for n, parts, align in [(4096, 8, 128), (4096, 3, 128), (100, 4, 128), (65536, 8, 128), (19, 2, 1), (5, 8, 1)]:
        b = partition(n, parts, align)
        assert b[0] == 0 and b[-1] == n and len(b) == parts + 1
        assert all(b[i] <= b[i + 1] for i in range(parts))
        assert all(x % align == 0 for x in b[1:-1])
    sh = [Shard(r, 4, 4096, 1024) for r in range(4)]
    rows = [s.rows for s in sh]
    assert rows[0][0] == 0 and rows[-1][1] == 4096
    assert all(rows[i][1] == rows[i + 1][0] for i in range(3))
    opts = sh[2].engine_options()
    assert opts['world_size'] == 4 and opts['rank'] == 2 and opts['row_end'] - opts['row_begin'] == 1024
    assert opts['col_end'] - opts['col_begin'] == 256


class _FakeEngine(object):
    """Records the sweeps and saves a run_gibbs call schedules."""

    def __init__(self):
        self.step = 0
        self.saved = {}

    def sweep(self, n):
        self.step += n

    def run_segment(self, nsweeps, first_save, nthin, sample_offset, **outs):
        for s in range(nsweeps):
            if s >= first_save and (s - first_save) % nthin == 0:
                self.saved[sample_offset + (s - first_save) // nthin] = self.step + s
        self.step += nsweeps


@pytest.mark.parametrize('nburn,nthin,nsamples,print_freq,verbose', [
    (10, 1, 5, 100, False), (10, 3, 4, 7, True), (0, 2, 6, 5, True), (7, 5, 3, 4, True), (3, 1, 0, 2, True),
    (1000, 1, 1000, 50, True)])
def test_run_gibbs_schedule_matches_reference(nburn, nthin, nsamples, print_freq, verbose, capsys):
    """Same sweeps, same saved steps and same 'Step' lines as genlasso.py:37-66."""
    from functionalmf_b200.factor import _BayesianModel

    class M(_BayesianModel):
        def __init__(self):
            self._engine = _FakeEngine()

        def _begin(self, data):
            pass

        def _end(self):
            pass

        def _alloc_results(self, n):
            return {}

        def _result_buffers(self, r):
            return {}

        def _finish_results(self, r):
            return r

    m = M()
    m.run_gibbs(None, nburn=nburn, nthin=nthin, nsamples=nsamples, verbose=verbose, print_freq=print_freq)
    nsteps = nburn + nthin * nsamples
    assert m._engine.step == nsteps                       # includes the trailing nthin-1 sweeps (Q12)
    want = {(s - nburn) // nthin: s for s in range(nsteps) if s >= nburn and (s - nburn) % nthin == 0}
    assert m._engine.saved == want
    out = capsys.readouterr().out
    lines = [l for l in out.splitlines() if l.startswith('\tStep')]
    if verbose:
        assert lines == ['\tStep {}'.format(s) for s in range(0, nsteps, print_freq)]
    else:
        assert lines == []
