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
