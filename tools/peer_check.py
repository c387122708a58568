"""Multi-GPU check of the peer-memory exchange (csxb_xchg_*), one process per GPU under torchrun:
repeated y = alpha*A*x with the halo rows stored into the neighbours' vectors by the SpMV kernel, compared
step by step with a CSR product on the host.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/peer_check.py [workload]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsex_b200 import CsxMatrix  # noqa: E402
from sparsex_b200.dist import connect_peer_exchange  # noqa: E402
from tests.matrices import poisson2d, rmat, stencil27  # noqa: E402

CASES = {"poisson": (lambda: poisson2d(300), {}), "rmat": (lambda: rmat(14), {"spx.preproc.xform": "none"}),
         "stencil": (lambda: stencil27(40), {}),
         # tiles of 4 rows per thread (what partitions of 2^20 rows and more run): edge tiles split over four CTAs
         "poisson4": (lambda: poisson2d(300), {"spx.b200.rows_per_thread": 4}),
         "stencil4": (lambda: stencil27(40), {"spx.b200.rows_per_thread": 4}),
         # block units in the block tables of the gather kernel; delta units in the stream kernel
         "blocks": (lambda: stencil27(40), {"spx.preproc.xform": "br,bc"}),
         "blocks4": (lambda: stencil27(40), {"spx.preproc.xform": "br,bc", "spx.b200.rows_per_thread": 4})}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name in (sys.argv[1:] or list(CASES)):
        gen, opts = CASES[name]
        rp, ci, va, n = gen()
        o = dict(opts, **{"spx.rt.nr_threads": world})
        A = CsxMatrix.tune_csr(rp, ci, va, n, n, o, part_lo=rank, part_hi=rank + 1).upload(local)
        ex, ranges, windows = connect_peer_exchange(A, rank, world, "cuda")
        lo, cnt = ranges[rank]
        x = np.random.default_rng(7).uniform(-1, 1, n)
        ex.vector(0).copy_(torch.from_numpy(x))
        ex.vector(1).zero_()
        dist.barrier()
        rows = np.repeat(np.arange(n), np.diff(rp))
        cur, alpha, worst = x, 0.2, 0.0
        steps = 6
        for _ in range(steps):
            ex.spmv(alpha)
        torch.cuda.synchronize()
        dist.barrier()
        # replay on the host: every step's input is the previous reference (6 steps of a contraction: stable)
        covered = np.zeros(n, bool)
        for a, b in ranges:
            covered[a:a + b] = True
        for _ in range(steps):
            nxt = alpha * np.bincount(rows, weights=va * cur[ci], minlength=n)
            nxt[~covered] = 0.0
            cur = nxt
        got = ex.vector(steps & 1)[lo:lo + cnt].cpu().numpy()
        scale = np.abs(cur).max() + 1e-300
        worst = float(np.abs(got - cur[lo:lo + cnt]).max() / scale)
        err = ex.error()
        good = worst < 1e-10 and err == 0 and ex.steps() == steps
        ok &= good
        print("rank %d/%d %-8s rows [%d,+%d) window %s  max err/scale %.2e  flag-error %d  %s"
              % (rank, world, name, lo, cnt, windows[rank], worst, err, "ok" if good else "FAIL"), flush=True)
        dist.barrier()
        ex.close()
        A.close()
    # CSX-Sym across ranks: local SpMV, transposed contributions sent to their owners (SymHaloReduce, NCCL)
    from sparsex_b200 import lib
    from sparsex_b200.dist import SymHaloReduce, gather_row_ranges
    from tests.matrices import sym_block_banded
    rp, ci, va, n = sym_block_banded(30000, b=256)
    A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.rt.nr_threads": world, "spx.matrix.symmetric": "true"},
                           part_lo=rank, part_hi=rank + 1).upload(local)
    L = lib()
    lo, cnt = L.csxb_part_info(A._h, 0, 3), L.csxb_part_info(A._h, 0, 8)
    ranges = gather_row_ranges(lo, cnt, "cuda")
    hl = torch.tensor([L.csxb_info(A._h, 8), L.csxb_info(A._h, 9)], dtype=torch.int64, device="cuda")
    allh = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(allh, hl)
    x = np.random.default_rng(9).uniform(-1, 1, n)
    dx = torch.from_numpy(x).cuda()
    dy = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    A.spmv(0.5, dx, dy, overwrite=True)
    SymHaloReduce(ranges, [(int(t[0]), int(t[1])) for t in allh], rank, dy)(dy)
    torch.cuda.synchronize()
    rows = np.repeat(np.arange(n), np.diff(rp))
    ref = 0.5 * np.bincount(rows, weights=va * x[ci], minlength=n)
    bound = 0.5 * np.bincount(rows, weights=np.abs(va * x[ci]), minlength=n) + 1e-300
    worst = float(np.max(np.abs(dy[lo:lo + cnt].cpu().numpy() - ref[lo:lo + cnt]) / bound[lo:lo + cnt]))
    good = worst <= 1e-12
    ok &= good
    print("rank %d/%d sym      rows [%d,+%d) halo %s  max err/bound %.2e  %s" % (rank, world, lo, cnt, tuple(int(v) for v in hl), worst,
                                                                              "ok" if good else "FAIL"), flush=True)
    A.close()
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
