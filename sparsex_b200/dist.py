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


class WindowExchange(object):
    """Sends every rank only the part of the other ranks' y pieces that its rows read.

    window[r] = (first, last) zero-based column the partition of rank r touches (CSXB_P_COL_MIN/MAX).  For
    banded matrices this is a halo of a few thousand elements per neighbour instead of the whole vector; when
    a window covers (almost) everything the plain all-gather of PieceExchange is used.  All sends and receives
    of a step go out as one grouped NCCL launch (batch_isend_irecv).
    """

    def __init__(self, ranges, windows, rank):
        self.rank, self.plan_send, self.plan_recv = rank, [], []
        world = len(ranges)
        need = 0
        for q in range(world):          # owner
            qlo, qn = ranges[q]
            for r in range(world):      # reader
                if q == r or qn == 0:
                    continue
                wlo, whi = windows[r]
                lo, hi = max(qlo, wlo), min(qlo + qn, whi + 1)
                if hi <= lo:
                    continue
                if q == rank:
                    self.plan_send.append((r, lo, hi))
                if r == rank:
                    self.plan_recv.append((q, lo, hi))
                    need += hi - lo
        total_other = sum(n for i, (_, n) in enumerate(ranges) if i != rank)
        self.fraction = need / max(total_other, 1)

    def __call__(self, buf):
        """buf: full-length vector whose own rows were just written by the local SpMV."""
        ops = [dist.P2POp(dist.isend, buf[lo:hi], peer) for peer, lo, hi in self.plan_send]
        ops += [dist.P2POp(dist.irecv, buf[lo:hi], peer) for peer, lo, hi in self.plan_recv]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return buf


class SymHaloReduce(object):
    """CSX-Sym across devices: rank r's lower triangle also updates rows [halo_lo_r, halo_hi_r) that lower ranks
    own (csxb_info CSXB_SYM_HALO_LO/HI).  After the local SpMV every rank sends those slices of its y to their
    owners, which add them to their own rows — the cross-device form of the reference's local-buffer + map
    reduction (CsxSpmv.cpp:37-50, Vector.cpp:291-299).  One grouped NCCL launch plus one add per neighbour."""

    def __init__(self, ranges, halos, rank, like):
        self.rank, self.send, self.recv = rank, [], []
        for r, (hlo, hhi) in enumerate(halos):       # sender r
            for q, (qlo, qn) in enumerate(ranges):   # owner q
                if q == r or qn == 0:
                    continue
                lo, hi = max(hlo, qlo), min(hhi, qlo + qn)
                if hi <= lo:
                    continue
                if r == rank:
                    self.send.append((q, lo, hi))
                if q == rank:
                    self.recv.append((r, lo, hi, torch.zeros(hi - lo, dtype=like.dtype, device=like.device)))

    def __call__(self, y):
        ops = [dist.P2POp(dist.isend, y[lo:hi], peer) for peer, lo, hi in self.send]
        ops += [dist.P2POp(dist.irecv, tmp, peer) for peer, lo, hi, tmp in self.recv]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for _, lo, hi, tmp in self.recv:
            y[lo:hi] += tmp
        return y


def connect_peer_exchange(matrix, rank, world, device):
    """One process per GPU: creates the engine's peer-memory exchange for `matrix` (csxb_xchg_*, include/csx_b200.h),
    all-gathers the CUDA IPC handles, row ranges and column windows over torch.distributed (plumbing only) and
    connects.  Afterwards a step is `ex.spmv(alpha)`: kernel launches only, the halo rows travel inside the SpMV
    kernel over NVLink.  Returns (exchange, ranges, windows); exchange is None on every rank when any rank could
    not set it up (no peer access between the GPUs, IPC not permitted ...), with the reason in `last_peer_error`."""
    global last_peer_error
    import numpy as np
    from .engine import PeerExchange, lib
    L = lib()
    ex, err = None, ""
    handle = np.zeros(64, np.uint8)
    try:
        ex = PeerExchange(matrix, rank, world)
        handle = ex.handle()
    except Exception as e:  # noqa: BLE001 - reported through last_peer_error, every rank must reach the collectives
        ex, err = None, repr(e)
    mine = torch.from_numpy(handle).to(device)
    allh = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(allh, mine)
    info = torch.tensor([L.csxb_part_info(matrix._h, 0, 3), L.csxb_part_info(matrix._h, 0, 1),
                         L.csxb_part_info(matrix._h, 0, 11), L.csxb_part_info(matrix._h, 0, 12)], dtype=torch.int64, device=device)
    alli = [torch.zeros(4, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(alli, info)
    ranges = [(int(t[0]), int(t[1])) for t in alli]
    windows = [(int(t[2]), int(t[3])) for t in alli]
    ok = torch.tensor([1 if ex is not None else 0], dtype=torch.int64, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok[0]):
        try:
            ex.connect(np.stack([t.cpu().numpy() for t in allh]), ranges, windows)
        except Exception as e:  # noqa: BLE001
            err = repr(e)
            ok[0] = 0
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not int(ok[0]):
        if ex is not None:
            ex.close()
        last_peer_error = err or "another rank could not set up the peer exchange"
        return None, ranges, windows
    dist.barrier()
    return ex, ranges, windows


last_peer_error = ""
