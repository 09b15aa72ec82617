"""Row / column sharding of the sweep across the GPUs of one node.

One process per GPU.  Rank g owns a contiguous block of data rows (and of W) and a
contiguous block of columns (the V conditionals it factorises); the exchange steps
(all-gather W, reduce-scatter of the column statistics, all-gather V, all-reduce of
the residual) run inside the C engine over NCCL.  ``torch.distributed`` (any
backend, gloo on CPU) is only used to agree on the NCCL unique id.
"""
import os


class Shard(object):
    def __init__(self, rank, world_size, nrows, ncols, row_align=128):
        self.rank, self.world_size = int(rank), int(world_size)
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.row_bounds = partition(nrows, world_size, row_align)
        self.col_bounds = partition(ncols, world_size, 1)

    @property
    def rows(self):
        return self.row_bounds[self.rank], self.row_bounds[self.rank + 1]

    @property
    def cols(self):
        return self.col_bounds[self.rank], self.col_bounds[self.rank + 1]

    def engine_options(self):
        r0, r1 = self.rows
        c0, c1 = self.cols
        return dict(row_begin=r0, row_end=r1, col_begin=c0, col_end=c1,
                    world_size=self.world_size, rank=self.rank)


def partition(n, parts, align=1):
    """Boundaries [b_0=0, ..., b_parts=n] of a near-even contiguous partition whose
    interior boundaries are multiples of ``align`` (the statistics kernels read the
    factor rows of a shard with 16-byte vector loads)."""
    n, parts, align = int(n), int(parts), max(1, int(align))
    blocks = (n + align - 1) // align
    bounds = [0]
    for r in range(1, parts):
        bounds.append(min(n, ((blocks * r) // parts) * align))
    bounds.append(n)
    return bounds


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def local_device():
    return int(os.environ.get('LOCAL_RANK', '0'))


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from ``src`` with torch.distributed (any backend)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def agree_unique_id():
    """Rank 0 creates the NCCL unique id (C ABI), everyone receives it."""
    import torch.distributed as dist
    from .engine import nccl_unique_id
    uid = nccl_unique_id() if dist.get_rank() == 0 else None
    return broadcast_bytes(uid, 0)


def agree_seed(seed, src=0):
    """Every rank returns rank ``src``'s seed (the replicated hyper-parameter steps need one Philox key)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError('a sharded model needs torch.distributed to be initialised (it agrees on the RNG seed '
                           'and the NCCL unique id)')
    return int.from_bytes(broadcast_bytes(int(seed).to_bytes(8, 'little'), src), 'little')


def state_digest(arrays):
    """Order-sensitive checksum of a list of float arrays (bit patterns, so NaN-safe)."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a, dtype=np.float64)
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def assert_same_on_all_ranks(arrays, what='state'):
    """Raise on every rank if the arrays differ between ranks."""
    import torch.distributed as dist
    mine = state_digest(arrays)
    box = [None] * dist.get_world_size()
    dist.all_gather_object(box, mine)
    if len(set(box)) != 1:
        raise ValueError('sharded model: %s differs between ranks (digests %s); pass identical *_init / *_true arrays '
                         'and seed on every rank' % (what, [b[:8] for b in box]))
