"""Times several builds of the library (tools/variants.sh) on the same tuned matrices (tuning aid).

Usage: python tools/vbench.py case1,case2,... lib1,lib2,...     (cases of tools/wbench.py; libs: names under
sparsex_b200/variants/ without the lib_ prefix, or "default")

The parent generates and tunes every case once and stores the container (csxb_save), x and the CSR product under /tmp;
one child per library loads the containers, uploads, checks the result and times 64 SpMVs captured in a CUDA graph."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TMP = os.environ.get("VBENCH_TMP", "/tmp/vbench")


def child(cases):
    import torch
    from sparsex_b200 import CsxMatrix
    tag = os.environ.get("VBENCH_TAG", "?")
    for name in cases:
        A = CsxMatrix.load(os.path.join(TMP, name + ".csxb")).upload(0, free_host=True)
        xh = np.load(os.path.join(TMP, name + ".x.npy"))
        yref = np.load(os.path.join(TMP, name + ".y.npy"))
        x = torch.from_numpy(xh).cuda()
        y = torch.zeros(A.nrows, dtype=torch.float64, device="cuda")
        for _ in range(3):
            A.spmv(1.0, x, y)
        torch.cuda.synchronize()
        err = float(np.abs(y.cpu().numpy() - yref).max() / np.abs(yref).max())
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
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 64)
        tr = A.traffic()
        print("%-10s %-7s %8.1f us  %6.0f GB/s (%4.1f%% of 6552)  %6.0f GFLOP/s  relerr %.1e" %
              (tag, name, best * 1e3, tr["total"] / best / 1e6, tr["total"] / best / 1e6 / 65.517, 2 * A.nnz / best / 1e6, err), flush=True)
        A.close()
        del x, y


def main():
    cases = sys.argv[1].split(",")
    libs = sys.argv[2].split(",")
    os.makedirs(TMP, exist_ok=True)
    from tools.wbench_cases import CASES
    from sparsex_b200 import CsxMatrix
    for name in cases:
        t0 = time.time()
        base, _, rpt = name.partition("@")   # "c4s@1" forces spx.b200.rows_per_thread=1
        gen, opts = CASES[base]
        if rpt:
            opts = dict(opts, **{"spx.b200.rows_per_thread": rpt})
        rp, ci, va, n = gen()
        xh = np.random.default_rng(0).uniform(-1, 1, n)
        rows = np.repeat(np.arange(n), np.diff(rp))
        yref = np.bincount(rows, weights=va * xh[ci], minlength=n)
        del rows
        A = CsxMatrix.tune_csr(rp, ci, va, n, n, dict(opts, **{"spx.b200.rows_info": "false"}))
        A.save(os.path.join(TMP, name + ".csxb"))
        np.save(os.path.join(TMP, name + ".x.npy"), xh)
        np.save(os.path.join(TMP, name + ".y.npy"), yref)
        print("# %s: %d rows, %d nnz, [%s] prepared in %.1f s" % (name, n, int(rp[-1]), A.partition(0).log.strip() if False else "", time.time() - t0), flush=True)
        A.close()
        del rp, ci, va
    for lib in libs:
        env = dict(os.environ, VBENCH_TAG=lib)
        if lib != "default":
            env["SPARSEX_B200_LIB"] = os.path.join(ROOT, "sparsex_b200", "variants", "lib_%s.so" % lib)
        subprocess.run([sys.executable, __file__, "--child", ",".join(cases)], env=env)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2].split(","))
    else:
        main()
