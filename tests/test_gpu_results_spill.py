"""run_gibbs(results_dir=...) keeps the saved samples in memory-mapped .npy files: same
numbers as the in-memory run with the same seed, readable afterwards with np.load."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_results_dir_equals_in_memory(tmp_path):
    from functionalmf_b200 import GaussianBayesianTensorFiltering
    rs = np.random.RandomState(0)
    N, M, T, K = 14, 6, 9, 3
    Y = rs.normal(size=(N, M, T, 2))
    Y[rs.random_sample(Y.shape) < 0.2] = np.nan
    res = []
    for spill in (None, str(tmp_path / 'chain')):
        model = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=1, sigma2_init=0.5, lam2_init=0.1,
                                                nu2_init=1, seed=17)
        res.append(model.run_gibbs(Y, nburn=5, nthin=2, nsamples=7, verbose=False, results_dir=spill))
    a, b = res
    for k in ('W', 'V', 'Tau2', 'sigma2', 'lam2', 'nu2'):
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))
    assert isinstance(b['V'], np.memmap)
    for k in ('W', 'V', 'Tau2'):
        on_disk = np.load(os.path.join(str(tmp_path / 'chain'), k + '.npy'), mmap_mode='r')
        np.testing.assert_array_equal(on_disk, a[k])
