"""Host-buffer SpMV (csxb_spmv_host) on the 2-D Poisson matrix of config 2: time per call for several slab sizes and
(CSXB_HOST_TRACE=1) the device time stamps of every slab (tuning aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsex_b200 import CsxMatrix  # noqa: E402
from tests.matrices import poisson2d  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
slabs = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["524288"])]
rp, ci, va, n = poisson2d(g)[:4]
xh = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, n)).pin_memory()
yh = torch.zeros(n, dtype=torch.float64).pin_memory()
for sr in slabs:
    A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.b200.rows_info": "false", "spx.b200.slab_rows": sr}).upload(0, free_host=True)
    for _ in range(3):
        A.spmv_host(1.0, xh.numpy(), yh.numpy())
    t0 = time.perf_counter()
    for _ in range(8):
        A.spmv_host(1.0, xh.numpy(), yh.numpy())
    dt = (time.perf_counter() - t0) / 8
    print("slab_rows %8d: %.3f ms per call, %.1f GB/s per direction" % (sr, dt * 1e3, n * 8 / dt / 1e9), flush=True)
    A.close()
