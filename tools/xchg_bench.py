"""Single-GPU cost breakdown of a peer-exchange step (tuning aid): plain SpMV kernel vs the exchange variant of
the same kernel plus the end-of-step sync kernel, with no neighbour (world = 1) — isolates launch/kernel
overheads from NVLink effects.  Usage: python tools/xchg_bench.py [grid] [parts]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsex_b200 import CsxMatrix, PeerExchange  # noqa: E402
from tests.matrices import poisson2d, stencil27  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rpt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rp, ci, va, n = stencil27(-g) if g < 0 else poisson2d(g)
A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.rt.nr_threads": parts, "spx.b200.rows_info": "false", "spx.b200.rows_per_thread": rpt},
                       part_lo=which, part_hi=which + 1).upload(0)
print("traffic", A.traffic())
print("grid %d, partition 0 of %d, rows per thread %d" % (g, parts, rpt))
x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, n)).cuda()
y = torch.zeros_like(x)


def timeit(fn, reps=64):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


print("plain kernel      %.1f us" % timeit(lambda: A.spmv(0.1, x, y)))
ex = PeerExchange(A, 0, 1)
ex.vector(0).copy_(x)
print("exchange step     %.1f us (eager launches)" % timeit(lambda: ex.spmv(0.1)))
for G in (2, 16):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for _ in range(G):
                ex.spmv(0.1)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    print("exchange step     %.1f us (graph of %d steps)" % (timeit(gr.replay, 16) / G, G))
    # plain kernel in a graph, ping-pong
    gr2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr2, stream=side):
            for i in range(G):
                A.spmv(0.1, x if i % 2 == 0 else y, y if i % 2 == 0 else x)
    torch.cuda.synchronize()
    print("plain kernel      %.1f us (graph of %d launches)" % (timeit(gr2.replay, 16) / G, G))
