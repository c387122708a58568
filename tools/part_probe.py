"""Times one partition of an N-way split on one GPU under several settings of an environment variable (tuning aid).
Usage: python tools/part_probe.py workload nparts part ENVVAR v1,v2,..."""
import os
import subprocess
import sys

if len(sys.argv) > 6:   # child: one measurement
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import workload
    from sparsex_b200 import CsxMatrix
    W = workload(sys.argv[1]); N = int(sys.argv[2]); p = int(sys.argv[3])
    lo, cnt = W.split(N, device="cuda")[p]
    rp, ci, va = W.rows(lo, lo + cnt, device="cuda")
    o = dict(W.opts, **{"spx.rt.nr_threads": N, "spx.b200.rows_info": "false"})
    A = (CsxMatrix.tune_csr_slab(rp, ci, va, W.n, W.n, lo, p, o) if N > 1 else CsxMatrix.tune_csr(rp, ci, va, W.n, W.n, o)).upload(0, free_host=True)
    x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, W.n)).cuda()
    y = torch.zeros(W.n, dtype=torch.float64, device="cuda")
    for _ in range(5):
        A.spmv(1.0, x, y)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(16):
                A.spmv(1.0, x, y)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    tr = A.traffic()
    ms = e0.elapsed_time(e1) / 64
    print("%s=%s  %s part %d/%d rows %d  %.1f us  %.0f GB/s" % (sys.argv[4], os.environ.get(sys.argv[4], "-"), sys.argv[1], p, N, cnt, ms * 1e3, tr["total"] / ms / 1e6), flush=True)
else:
    for v in sys.argv[5].split(","):
        env = dict(os.environ)
        env[sys.argv[4]] = v
        subprocess.run([sys.executable, __file__] + sys.argv[1:6] + ["child"], env=env)
