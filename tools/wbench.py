"""Times the shipped SpMV on scaled-down versions of every BASELINE config (tuning aid).
Usage: python tools/wbench.py [names...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsex_b200 import CsxMatrix, lib  # noqa: E402

from tools.wbench_cases import CASES  # noqa: E402


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


_cache = {}
for name in (sys.argv[1:] or list(CASES)):
    name, _, forced = name.partition(":")   # "c3b:4" forces spx.b200.slice=4
    gen, opts = CASES[name]
    if forced:
        opts = dict(opts, **{"spx.b200.slice": forced})
    if name not in _cache:
        _cache.clear()
        rp, ci, va, n = gen()
        xh = np.random.default_rng(0).uniform(-1, 1, n)
        rows = np.repeat(np.arange(n), np.diff(rp))
        _cache[name] = (rp, ci, va, n, xh, np.bincount(rows, weights=va * xh[ci], minlength=n))
        del rows
    rp, ci, va, n, xh, yref = _cache[name]
    if forced:
        name += ":" + forced
    nnz = int(rp[-1])
    t0 = time.time()
    A = CsxMatrix.tune_csr(rp, ci, va, n, n, dict(opts, **{"spx.b200.rows_info": "false"}))
    t1 = time.time()
    log = lib().csxb_part_log(A._h, 0).decode()
    A.upload(0, free_host=True)
    t2 = time.time()
    tr = A.traffic()
    x = torch.from_numpy(xh).cuda()
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    A.spmv(1.0, x, y)
    torch.cuda.synchronize()
    err = np.abs(y.cpu().numpy() - yref).max() / np.abs(yref).max()
    ms = timeit(lambda: A.spmv(1.0, x, y))
    print("%-6s rows %9d nnz %10d  tune %.1fs upload %.1fs  [%s]  %.1f us  %.0f GB/s (%.0f%% of 6552)  %.0f GFLOP/s  bytes/nnz %.2f (ctl %.2f tables %.2f)  relerr %.1e"
          % (name, n, nnz, t1 - t0, t2 - t1, log.strip(), ms * 1e3, tr["total"] / ms / 1e6, tr["total"] / ms / 1e6 / 65.517,
             2 * nnz / ms / 1e6, tr["total"] / nnz, tr["ctl"] / nnz, tr["tables"] / nnz, err), flush=True)
    A.close()
    del x, y
