#!/usr/bin/env python
"""Sharded sweep == single-GPU sweep (same Philox keys).  Launch with
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functionalmf_b200.engine import Engine                      # noqa: E402
from functionalmf_b200.distributed import Shard, agree_unique_id  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok = True
    for (N, M, T, R, K, order) in [(1024, 64, 32, 3, 16, 2), (700, 37, 12, 2, 5, 1), (512, 16, 10, 1, 32, 2), (900, 19, 24, 3, 8, 3)]:
        rs = np.random.RandomState(7)
        W = rs.normal(size=(N, K)); W[np.triu_indices(K, k=1)] = 0
        V = rs.normal(size=(M, T, K)).cumsum(axis=1) * 0.3
        Y = np.einsum('nk,mtk->nmt', W, V)[..., None] + rs.normal(size=(N, M, T, R))
        Y[rs.random_sample(Y.shape) < 0.2] = np.nan
        sh = Shard(rank, world, N, M)
        eng = Engine(N, M, T, nembeds=K, tf_order=order, seed=99, device=local, **sh.engine_options())
        eng.nccl_init(agree_unique_id())
        RD = eng.RD
        st = dict(W=W * 0.9, V=V * 1.1, Tau2=rs.gamma(2.0, size=(M, RD)) + 0.05, Tau2_a=rs.gamma(2.0, size=(M, RD)) + 0.05,
                  Tau2_b=rs.gamma(2.0, size=(M, RD)) + 0.05, Tau2_c=rs.gamma(2.0, size=(M, RD)) + 0.05)
        sc = dict(lam2=0.7, lam2_a=1.3, sigma2=0.9, nu2=1.1)

        def load(e):
            for k, v in st.items():
                e.set(k, v)
            for k, v in sc.items():
                e.set(k, [v])

        r0, r1 = sh.rows
        eng.set_data_gaussian(Y[r0:r1])
        load(eng)
        eng.sweep(3)
        got = {k: eng.get(k) for k in ('W', 'V', 'Tau2')}
        got.update({k: eng.get_scalar(k) for k in ('nu2', 'sigma2', 'lam2')})
        # held-out evaluator: local sums add up to the single-GPU sums
        target = np.nanmean(Y, axis=-1) if R > 1 else Y[..., 0]
        cls = (np.arange(N * M * T).reshape(N, M, T) % 3 == 0).astype(np.uint8)
        eng.eval_set(0, target[r0:r1], cls[r0:r1], nclasses=2, loglik=1, cell_state=1, auto_update=False, max_samples=2)
        eng.eval_update(0)
        ev = torch.from_numpy(eng.eval_samples(0, 2)).cuda()
        dist.all_reduce(ev)
        ev = ev.cpu().numpy()
        ms = eng.sweep_timed(5)
        eng.close()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            ref = Engine(N, M, T, nembeds=K, tf_order=order, seed=99, device=local)
            ref.set_data_gaussian(Y)
            load(ref)
            ref.sweep(3)
            worst = 0.0
            for k in ('W', 'V', 'Tau2'):
                a, b = got[k], ref.get(k)
                worst = max(worst, float(np.max(np.abs(a - b)) / np.max(np.abs(b))))
            for k in ('nu2', 'sigma2', 'lam2'):
                worst = max(worst, abs(got[k] - ref.get_scalar(k)) / abs(ref.get_scalar(k)))
            ref.eval_set(0, target, cls, nclasses=2, loglik=1, cell_state=1, auto_update=False, max_samples=2)
            ref.eval_update(0)
            ev1 = ref.eval_samples(0, 2)
            worst = max(worst, float(np.max(np.abs(ev - ev1) / np.abs(ev1))))
            ms1 = ref.sweep_timed(5)
            ref.close()
            good = worst < 1e-8
            ok = ok and good
            print('shape %s world %d: max normwise diff vs single GPU %.3e %s | ms/sweep sharded %.3f single %.3f'
                  % ((N, M, T, R, K, order), world, worst, 'OK' if good else 'MISMATCH', ms / 5, ms1 / 5), flush=True)
        dist.barrier()
    ok = model_level_check(rank, world, local) and ok
    dist.destroy_process_group()
    if rank == 0:
        print('MULTI_GPU_CHECK', 'PASS' if ok else 'FAIL', flush=True)
        sys.exit(0 if ok else 1)


def model_level_check(rank, world, local):
    """a sharded model built WITHOUT seed= (every process draws its own from np.random) must still
    run one chain: the constructor agrees on rank 0's seed, so sigma2 / lam2 / nu2 / Tau2 / W / V are bit-identical on
    every rank after a few sweeps, and run_gibbs returns the same samples everywhere."""
    from functionalmf_b200 import GaussianBayesianTensorFiltering
    N, M, T, R, K = 640, 24, 16, 2, 8
    rs = np.random.RandomState(3)
    Y = rs.normal(size=(N, M, T, R))
    Y[rs.random_sample(Y.shape) < 0.2] = np.nan
    np.random.seed(1000 + rank)                      # different per-process numpy streams
    sh = Shard(rank, world, N, M)
    model = GaussianBayesianTensorFiltering(N, M, T, nembeds=K, tf_order=1, shard=sh, device=local)
    r0, r1 = sh.rows
    res = model.run_gibbs(Y[r0:r1], nburn=2, nthin=1, nsamples=3, verbose=False)
    from functionalmf_b200.distributed import state_digest
    digest = state_digest([res['W'], res['V'], res['Tau2'], res['sigma2'], res['lam2'], model.Tau2_a, model.nu2 * np.ones(1)])
    box = [None] * world
    dist.all_gather_object(box, digest)
    good = len(set(box)) == 1 and np.all(np.isfinite(res['V']))
    if rank == 0:
        print('model-level (no seed=) world %d: %s' % (world, 'identical on all ranks OK' if good else 'RANKS DIFFER'), flush=True)
    return good


if __name__ == '__main__':
    main()
