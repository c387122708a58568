"""Measurements for the SURVEY.md section 8f rows built so far (run on a B200):
  * container: spx_mat_save / spx_mat_restore + upload vs tuning from CSR (C2 matrix)
  * single entries: csxb_get_entry / csxb_set_entry latency on the device copy
  * BLAS-1 on HBM-resident vectors: csxb_vec_axpby (24 B per element) and csxb_vec_dot (16 B per element) vs the HBM peak
Usage: python tools/next_rows_bench.py [grid]"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsex_b200 import CsxMatrix, lib  # noqa: E402
from tests.matrices import poisson2d  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rp, ci, va, n = poisson2d(g)
out = {"matrix": "poisson2d_%d" % g, "rows": n, "nnz": int(rp[-1])}
t0 = time.time()
A = CsxMatrix.tune_csr(rp, ci, va, n, n, {"spx.b200.rows_info": "false"})
out["tune_s"] = round(time.time() - t0, 3)
t0 = time.time()
A.upload(0)
out["upload_s"] = round(time.time() - t0, 3)
path = os.path.join(tempfile.mkdtemp(), "m.csxb")
t0 = time.time()
A.save(path)
out["save_s"] = round(time.time() - t0, 3)
out["container_MB"] = round(os.path.getsize(path) / 1e6, 1)
t0 = time.time()
B = CsxMatrix.load(path)
out["load_s"] = round(time.time() - t0, 3)
t0 = time.time()
B.upload(0)
out["restore_upload_s"] = round(time.time() - t0, 3)
x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, n)).cuda()
y1, y2 = torch.zeros_like(x), torch.zeros_like(x)
A.spmv(1.0, x, y1)
B.spmv(1.0, x, y2)
torch.cuda.synchronize()
out["restored_equals_tuned"] = bool(torch.equal(y1, y2))
# single entries
rows = np.repeat(np.arange(n), np.diff(rp))
idx = np.random.default_rng(1).integers(0, len(ci), 2000)
A.get_entry(int(rows[idx[0]]), int(ci[idx[0]]))   # builds the row index
t0 = time.time()
for k in idx:
    A.get_entry(int(rows[k]), int(ci[k]))
out["get_entry_us"] = round((time.time() - t0) / len(idx) * 1e6, 2)
t0 = time.time()
for k in idx:
    A.set_entry(int(rows[k]), int(ci[k]), 1.25)
out["set_entry_us"] = round((time.time() - t0) / len(idx) * 1e6, 2)
# BLAS-1
L = lib()
L.csxb_vec_axpby.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_void_p]
L.csxb_vec_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_void_p]
m = 1 << 26
a = torch.rand(m, dtype=torch.float64, device="cuda")
b = torch.rand(m, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    L.csxb_vec_axpby(c.data_ptr(), a.data_ptr(), b.data_ptr(), 1.0, 0.5, m, s)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    L.csxb_vec_axpby(c.data_ptr(), a.data_ptr(), b.data_ptr(), 1.0, 0.5, m, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
out["axpby_GBs"] = round(24.0 * m / ms / 1e6, 1)
assert torch.allclose(c, a + 0.5 * b)
r = C.c_double()
L.csxb_vec_dot(a.data_ptr(), b.data_ptr(), m, C.byref(r), s)
t0 = time.time()
for _ in range(20):
    L.csxb_vec_dot(a.data_ptr(), b.data_ptr(), m, C.byref(r), s)
ms = (time.time() - t0) / 20 * 1e3
out["dot_GBs"] = round(16.0 * m / ms / 1e6, 1)
ref = float(torch.dot(a, b))
out["dot_rel_err"] = abs(r.value - ref) / abs(ref)
try:
    out["hbm_peak_GBs"] = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps(out))
