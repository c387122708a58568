"""Host-buffer SpMV on a 27-point stencil: time per call for several caps on the resident CTAs of the slab kernels
(CSXB_HOST_SMEM, tuning aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsex_b200 import CsxMatrix  # noqa: E402
from tests.matrices import stencil27  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rp, ci, va, n = stencil27(g)[:4]
xh = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, n)).pin_memory()
yh = torch.zeros(n, dtype=torch.float64).pin_memory()
A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.b200.rows_info": "false"}).upload(0, free_host=True)
x = xh.cuda()
y = torch.zeros(n, dtype=torch.float64, device="cuda")
for _ in range(3):
    A.spmv(1.0, x, y)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    A.spmv(1.0, x, y)
torch.cuda.synchronize()
print("rows %d nnz %d: device-resident %.3f ms per SpMV" % (n, len(va), (time.perf_counter() - t0) / 20 * 1e3), flush=True)
yref = y.cpu().numpy().copy()
for smem in sys.argv[2].split(","):
    os.environ["CSXB_HOST_SMEM"] = smem
    for _ in range(3):
        A.spmv_host(1.0, xh.numpy(), yh.numpy())
    t0 = time.perf_counter()
    for _ in range(8):
        A.spmv_host(1.0, xh.numpy(), yh.numpy())
    dt = (time.perf_counter() - t0) / 8
    print("CSXB_HOST_SMEM %7s: %.3f ms per call, %.1f GB/s per direction, equal to the device path: %s"
          % (smem, dt * 1e3, n * 8 / dt / 1e9, bool(np.array_equal(yh.numpy(), yref))), flush=True)
