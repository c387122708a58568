"""Times the experimental phase-B kernel variants (csxb_debug_variant) on one tuned matrix.
Usage: python tools/kbench.py [grid]   (tuning aid; results go to profiles/ by hand)"""
import ctypes as C
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsex_b200 import CsxMatrix, lib  # noqa: E402
from tests.matrices import poisson2d, stencil27  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "c2"
rp, ci, va, n = poisson2d(4096) if kind == "c2" else stencil27(int(sys.argv[2]) if len(sys.argv) > 2 else 160)
t0 = time.time()
A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.b200.rows_info": "false"}).upload(0, free_host=True)
print("tune+upload %.1f s" % (time.time() - t0), A.traffic())
L = lib()
L.csxb_debug_variant.restype = C.c_int
L.csxb_debug_variant.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, n)).cuda()
y = torch.zeros(n, dtype=torch.float64, device="cuda")
yref = torch.zeros(n, dtype=torch.float64, device="cuda")
A.spmv(0.5, x, yref)
torch.cuda.synchronize()
total = A.traffic()["total"]


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timeit(lambda: A.spmv(0.5, x, y))
print("shipped kernel: %.1f us  %.0f GB/s" % (ms * 1e3, total / ms / 1e6))
s = torch.cuda.current_stream().cuda_stream
for v in (8,):
    y.zero_()
    rc = L.csxb_debug_variant(A._h, v, 0.5, x.data_ptr(), y.data_ptr(), s)
    if rc != 0:
        print("variant", v, "failed:", L.csxb_last_error().decode())
        continue
    torch.cuda.synchronize()
    err = float((y - yref).abs().max() / yref.abs().max())
    ms = timeit(lambda: L.csxb_debug_variant(A._h, v, 0.5, x.data_ptr(), y.data_ptr(), s))
    print("variant %d: %.1f us  %.0f GB/s  relerr %.1e" % (v, ms * 1e3, total / ms / 1e6, err))
# plain device copy of the same number of bytes for reference
a = torch.empty(total // 16, dtype=torch.float64, device="cuda")
b2 = torch.empty_like(a)
ms = timeit(lambda: b2.copy_(a))
print("torch copy of %d MB (read+write): %.1f us  %.0f GB/s" % (total // 1e6, ms * 1e3, total / ms / 1e6))
