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
         "stencil": (lambda: stencil27(40), {})}


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
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
