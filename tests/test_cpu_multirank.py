"""world_size-2 (and 3) gloo tests of the one-process-per-GPU host logic: each rank tunes only its own partition
through the C-ABI, the partitions tile the row space, and the unequal y pieces are exchanged into the next x."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


def _worker(rank, world, port, sym, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import OracleMatrix
        from sparsex_b200 import CsxMatrix
        from sparsex_b200.dist import PieceExchange, gather_row_ranges, rank_options
        from tests.matrices import poisson2d, sym_block_banded
        rp, ci, va, n = sym_block_banded(700, b=16) if sym else poisson2d(70)
        opts = {k: v for k, v in rank_options(rank, world).items() if k in ("spx.rt.nr_threads",)}
        if sym:
            opts["spx.matrix.symmetric"] = "true"
        A = CsxMatrix.tune_csr(rp, ci, va, n, n, opts, part_lo=rank, part_hi=rank + 1)
        assert A.nparts == 1 and A.part_lo == rank and A.nparts_total == world
        P = A.partition(0)
        # the partition this rank encoded is the reference's partition `rank` of the world-way split
        O = OracleMatrix.from_csr(rp, ci, va, n, n).tune(opts)
        assert np.array_equal(P.ctl, O.parts[rank].ctl) and np.array_equal(P.values, O.parts[rank].values)
        owned = len(P.dvalues) if sym else P.nrows
        ranges = gather_row_ranges(P.row_start, owned, "cpu")
        assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
        # one SpMV step + exchange: every rank ends up with the full y as its next x
        rng = np.random.default_rng(5)
        x = torch.from_numpy(rng.uniform(-1, 1, n))
        rows = np.repeat(np.arange(n), np.diff(rp))
        y_full = np.zeros(n)
        np.add.at(y_full, rows, va * x.numpy()[ci])
        lo, cnt = ranges[rank]
        ex = PieceExchange(x, ranges)
        xn = ex(torch.from_numpy(y_full[lo:lo + cnt].copy()))
        assert np.array_equal(xn.numpy(), y_full)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,sym", [(2, False), (3, False), (2, True)])
def test_partition_per_rank_and_piece_exchange(world, sym):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world * 7 + (3 if sym else 0) + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, sym, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _sym_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparsex_b200 import CsxMatrix
        from sparsex_b200.dist import SymHaloReduce, gather_row_ranges
        from tests.matrices import sym_block_banded
        rp, ci, va, n = sym_block_banded(600, b=12)
        opts = {"spx.rt.nr_threads": world, "spx.matrix.symmetric": "true"}
        A = CsxMatrix.tune_csr(rp, ci, va, n, n, opts, part_lo=rank, part_hi=rank + 1)
        P = A.partition(0)
        lo, cnt = P.row_start, len(P.dvalues)
        ranges = gather_row_ranges(lo, cnt, "cpu")
        # what the device kernels produce for this rank, computed here from the lower triangle of its rows:
        # own rows get a_ij x_j (j <= i), every column j < i gets the transposed a_ij x_i
        x = np.random.default_rng(3).uniform(-1, 1, n)
        y = np.zeros(n)
        col_min = lo
        for i in range(lo, lo + cnt):
            for k in range(rp[i], rp[i + 1]):
                j = ci[k]
                if j > i:
                    continue
                y[i] += va[k] * x[j]
                if j < i:
                    y[j] += va[k] * x[i]
                    col_min = min(col_min, j)
        halo = torch.tensor([col_min, lo], dtype=torch.int64)
        allh = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allh, halo)
        halos = [(int(t[0]), int(t[1])) for t in allh]
        yt = torch.from_numpy(y)
        SymHaloReduce(ranges, halos, rank, yt)(yt)
        rows = np.repeat(np.arange(n), np.diff(rp))
        yref = np.zeros(n)
        np.add.at(yref, rows, va * x[ci])
        assert np.abs(yt.numpy()[lo:lo + cnt] - yref[lo:lo + cnt]).max() < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_symmetric_halo_reduce(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 11 + os.getpid() % 200
    procs = [ctx.Process(target=_sym_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _peer_fallback_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sparsex_b200.dist as sdist
        from sparsex_b200 import CsxMatrix
        from tests.matrices import poisson2d
        rp, ci, va, n = poisson2d(40)
        A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.rt.nr_threads": world}, part_lo=rank, part_hi=rank + 1)
        # no GPU here: the exchange cannot be created; every rank must learn that and agree on the fallback,
        # without any rank hanging in a collective
        ex, ranges, windows = sdist.connect_peer_exchange(A, rank, world, "cpu")
        assert ex is None and sdist.last_peer_error
        assert sum(cnt for _, cnt in ranges) == n and len(windows) == world
        lo, cnt = ranges[rank]
        assert windows[rank][0] <= lo and windows[rank][1] >= lo + cnt - 1
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_peer_exchange_setup_falls_back_consistently():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_peer_fallback_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
