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


def test_checkpoint_resume_continues_the_same_chain(tmp_path):
    """save_checkpoint / load_checkpoint (state arrays, scalars, Philox seed and sweep counter): a chain of 4 + 6 sweeps with
    a checkpoint in between, restored into a FRESH model, ends bit for bit where an uninterrupted chain of 10 sweeps ends."""
    import numpy as np
    from functionalmf_b200 import GaussianBayesianTensorFiltering
    rs = np.random.RandomState(11)
    N, M, T, K = 40, 9, 12, 4
    Y = rs.normal(size=(N, M, T, 2))
    Y[rs.random_sample(Y.shape) < 0.2] = np.nan

    def model():
        return GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=2, seed=77, sigma2_init=0.5, lam2_init=0.1)
    a = model()
    for _ in range(10):
        a.resample(Y)
    b = model()
    for _ in range(4):
        b.resample(Y)
    ck = str(tmp_path / 'chain.npz')
    b.save_checkpoint(ck)
    c = model()
    c.resample(Y)                      # uploads the data (and moves the chain somewhere else)
    c.load_checkpoint(ck)
    for _ in range(6):
        c.resample(Y)
    for name in ('W', 'V', 'Tau2'):
        assert np.array_equal(getattr(a, name), getattr(c, name)), name
    assert a.sigma2 == c.sigma2 and a.lam2 == c.lam2 and a.nu2 == c.nu2
    import pytest
    d = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=2, seed=78)
    with pytest.raises(ValueError):
        d.load_checkpoint(ck)
