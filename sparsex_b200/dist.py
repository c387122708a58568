"""One-process-per-GPU plumbing for the row-partitioned SpMV (torch.distributed is only the transport).

Rank r owns partition r of the reference's nnz-balanced split with spx.rt.nr_threads = world size
(SparseInternal.hpp:119-152).  Between repeated SpMVs the y pieces are exchanged into every rank's next x.
"""
import torch
import torch.distributed as dist


def rank_options(rank, world, local_rank=None):
    """Engine options that restrict a process to its own partition."""
    return {"spx.rt.nr_threads": world, "spx.b200.part_lo": rank, "spx.b200.part_hi": rank + 1,
            "spx.b200.device": rank if local_rank is None else local_rank}


def gather_row_ranges(row_lo, row_n, device):
    """All ranks' (first row, row count); consecutive and disjoint by construction of the split."""
    world = dist.get_world_size()
    mine = torch.tensor([row_lo, row_n], dtype=torch.int64, device=device)
    out = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(out, mine)
    ranges = [(int(t[0]), int(t[1])) for t in out]
    pos = 0
    for lo, cnt in ranges:
        if cnt and lo != pos:
            raise RuntimeError("row partitions are not consecutive: %r" % (ranges,))
        pos = lo + cnt if cnt else pos
    return ranges


class PieceExchange(object):
    """x_next[lo_r : lo_r + n_r] := y piece of rank r, for every r (an all-gather with unequal pieces)."""

    def __init__(self, x, ranges):
        self.x, self.ranges = x, ranges
        self.views = [x[lo:lo + cnt] for lo, cnt in ranges]
        self.uneven_ok = dist.get_backend() == "nccl"   # NCCL: one grouped launch of per-owner broadcasts
        if not self.uneven_ok:
            self.maxn = max(cnt for _, cnt in ranges)
            self.pad = torch.zeros(self.maxn, dtype=x.dtype, device=x.device)
            self.recv = [torch.zeros(self.maxn, dtype=x.dtype, device=x.device) for _ in ranges]

    def __call__(self, y_piece):
        if self.uneven_ok:
            dist.all_gather(self.views, y_piece)
            return self.x
        self.pad[:y_piece.numel()] = y_piece
        dist.all_gather(self.recv, self.pad)
        for v, r, (_, cnt) in zip(self.views, self.recv, self.ranges):
            v.copy_(r[:cnt])
        return self.x
