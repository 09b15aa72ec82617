"""world_size-2 gloo test of the host-side plumbing of the sharded sweep (CPU only):
both ranks agree on the (fake) NCCL unique id and on complementary shards."""
import os
import socket
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from functionalmf_b200.distributed import Shard, broadcast_bytes, env_rank_world
    assert env_rank_world() == (rank, world)
    uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
    sh = Shard(rank, world, 1000, 37, row_align=128)
    # emulate the exchange pattern of one sweep on CPU tensors: all-gather of the W row blocks,
    # sum of partial column statistics, all-gather of the V column blocks
    N, M, K = 1000, 37, 3
    rs = np.random.RandomState(5)
    W = rs.normal(size=(N, K))
    partial = rs.normal(size=(world, M, 4))
    r0, r1 = sh.rows
    Wloc = torch.zeros(N, K, dtype=torch.float64)
    Wloc[r0:r1] = torch.from_numpy(W[r0:r1])
    for r in range(world):
        b0, b1 = Shard(r, world, N, M, row_align=128).rows
        if b1 > b0:
            dist.broadcast(Wloc[b0:b1], src=r)
    stats = torch.from_numpy(partial[rank].copy())
    dist.all_reduce(stats)
    np.save(os.path.join(out_dir, 'r%d.npy' % rank),
            np.concatenate([np.frombuffer(uid, dtype=np.uint8).astype(float), [r0, r1, sh.cols[0], sh.cols[1]],
                            Wloc.numpy().ravel(), stats.numpy().ravel()]))
    dist.destroy_process_group()


def test_two_rank_plumbing(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / 'r0.npy')
    b = np.load(tmp_path / 'r1.npy')
    assert np.array_equal(a[:128], np.arange(128)) and np.array_equal(b[:128], a[:128])
    assert a[128] == 0 and a[129] == b[128] and b[129] == 1000          # contiguous row shards
    assert a[129] % 128 == 0
    assert a[130] == 0 and a[131] == b[130] and b[131] == 37            # contiguous column shards
    rs = np.random.RandomState(5)
    W = rs.normal(size=(1000, 3))
    partial = rs.normal(size=(2, 37, 4))
    assert np.array_equal(a[132:132 + 3000], W.ravel()) and np.array_equal(b[132:132 + 3000], W.ravel())
    assert np.allclose(a[132 + 3000:], partial.sum(0).ravel()) and np.allclose(a[132 + 3000:], b[132 + 3000:])


def _seed_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from functionalmf_b200.distributed import agree_seed, assert_same_on_all_ranks
    # every rank draws its own seed (what the constructor does without seed=): rank 0's must win everywhere
    mine = 1000 + 17 * rank
    got = agree_seed(mine)
    same = np.arange(6.0).reshape(2, 3)
    assert_same_on_all_ranks([same, np.array([np.nan, 1.0])], what='identical arrays')
    raised = False
    try:
        assert_same_on_all_ranks([same + (1e-300 if rank == 1 else 0.0)], what='arrays that differ in one bit')
    except ValueError:
        raised = True
    np.save(os.path.join(out_dir, 's%d.npy' % rank), np.array([got, float(raised)]))
    dist.destroy_process_group()


def test_two_rank_seed_agreement_and_state_check(tmp_path):
    """sharded models replicate sigma2 / nu2 / lam2 / Tau2, so the Philox seed and the initial
    state must be identical on every rank; the constructor enforces it with these two helpers."""
    world = 2
    port = _free_port()
    mp.spawn(_seed_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / 's0.npy')
    b = np.load(tmp_path / 's1.npy')
    assert a[0] == 1000 and b[0] == 1000          # rank 0's seed on both ranks
    assert a[1] == 1.0 and b[1] == 1.0            # the mismatch is reported on every rank


def _colshard_worker(rank, world, port, out_dir):
    """CPU emulation of the round-2 V-step decomposition with the same index arithmetic as
    nccl_shard.cu: ring exchange of the transposed count tiles (world rounds of one send + one receive), product block of
    the OWN columns over ALL rows, linear block as partial sums over the LOCAL rows reduced to the owner."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from functionalmf_b200.distributed import Shard
    N, M, T, K = 300, 7, 5, 3
    rs = np.random.RandomState(9)
    cnt = rs.randint(0, 4, size=(N, M * T)).astype(np.uint8)
    S = rs.normal(size=(N, M * T)) * (cnt > 0)
    W = rs.normal(size=(N, K))
    shards = [Shard(r, world, N, M, row_align=128) for r in range(world)]
    me = shards[rank]
    r0, r1 = me.rows
    c0, c1 = me.cols
    pad = lambda n: -(-n // 128) * 128
    nall_pad = pad(N)
    # local transposed counts [P][nloc_pad]
    srcT = np.zeros((M * T, pad(max(r1 - r0, 1))), dtype=np.uint8)
    srcT[:, :r1 - r0] = cnt[r0:r1].T
    ploc = (c1 - c0) * T
    cntT = np.zeros((max(ploc, 1), nall_pad), dtype=np.uint8)
    for r in range(world):                      # nccl_exchange_counts
        to, frm = (rank + r) % world, (rank - r + world) % world
        f0, f1 = shards[frm].rows
        if r == 0:
            block = srcT[c0 * T:c1 * T]
        else:
            t0, t1 = shards[to].cols
            send = torch.from_numpy(np.ascontiguousarray(srcT[t0 * T:t1 * T]))
            recv = torch.zeros((ploc, pad(max(f1 - f0, 1))), dtype=torch.uint8)
            reqs = []
            if send.numel():
                reqs.append(dist.isend(send, to))
            if recv.numel():
                reqs.append(dist.irecv(recv, frm))
            for q in reqs:
                q.wait()
            block = recv.numpy()
        if ploc and f1 > f0:
            cntT[:ploc, f0:f1] = block[:, :f1 - f0]
    # product block of the own columns over all rows; linear block: partial sums over local rows, summed across ranks
    il = np.tril_indices(K)
    Z = W[:, il[0]] * W[:, il[1]]
    prod = cntT[:ploc, :N].astype(float) @ Z
    part = torch.from_numpy(S[r0:r1].T @ W[r0:r1])          # [P][K] for ALL columns
    dist.all_reduce(part)                                      # (the engine reduce-scatters; every rank keeps its block)
    lin = part.numpy()[c0 * T:c1 * T]
    np.save(os.path.join(out_dir, 'c%d.npy' % rank), np.concatenate([[c0, c1], prod.ravel(), lin.ravel()]))
    dist.destroy_process_group()


def test_two_rank_column_sharded_statistics(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_colshard_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    N, M, T, K = 300, 7, 5, 3
    rs = np.random.RandomState(9)
    cnt = rs.randint(0, 4, size=(N, M * T)).astype(np.uint8)
    S = rs.normal(size=(N, M * T)) * (cnt > 0)
    W = rs.normal(size=(N, K))
    il = np.tril_indices(K)
    want_prod = cnt.T.astype(float) @ (W[:, il[0]] * W[:, il[1]])
    want_lin = S.T @ W
    seen = 0
    for r in range(world):
        a = np.load(tmp_path / ('c%d.npy' % r))
        c0, c1 = int(a[0]), int(a[1])
        ploc, L = (c1 - c0) * T, len(il[0])
        prod = a[2:2 + ploc * L].reshape(ploc, L)
        lin = a[2 + ploc * L:].reshape(ploc, K)
        assert np.allclose(prod, want_prod[c0 * T:c1 * T], rtol=1e-12, atol=1e-12)
        assert np.allclose(lin, want_lin[c0 * T:c1 * T], rtol=1e-12, atol=1e-12)
        seen += c1 - c0
    assert seen == M
