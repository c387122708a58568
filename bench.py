#!/usr/bin/env python
"""bench.py — CSX SpMV throughput of the B200 engine (and of the CPU reference arm).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4|c5|small] [--impl reference]

A *step* is one SpMV  y = alpha*A*x  over the whole matrix (the hot path of
BASELINE.json: spx_matvec_mult after spx_mat_tune).  For N > 1 the rows are the
reference's nnz-balanced partitions with spx.rt.nr_threads = N, one partition
per rank, and every step ends with the exchange of the y pieces into the next
x (NCCL all-gather over NVLink) — strong scaling on a fixed matrix.

Metric: GFLOP/s = 2 * nnz * K / t  (reference convention, src/bench/SparsexModule.cpp:80).
Timing: W untimed steps, then exactly K steps between a barrier + synchronize on both sides,
CUDA events on the launching stream, max over ranks.  The matrix (values alone: 671 MB for c2)
is far larger than the 126 MB L2, so consecutive steps cannot be served from cache.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, generator kwargs, tuning options, cpu sample kwargs)
    "c2": ("c2_poisson2d_4096", dict(kind="poisson2d", g=4096), {}, dict(kind="poisson2d", g=2048)),
    "c3": ("c3_stencil27_256", dict(kind="stencil27", g=256), {}, dict(kind="stencil27", g=96)),
    "c3b": ("c3_stencil27_256_blocks", dict(kind="stencil27", g=256), {"spx.preproc.xform": "br,bc"},
            dict(kind="stencil27", g=96)),
    "c4": ("c4_sym_block_banded_30M", dict(kind="symbb", nb=10_000_000, b=1024), {"spx.matrix.symmetric": "true"},
           dict(kind="symbb", nb=500_000, b=1024)),
    "c5": ("c5_rmat_26", dict(kind="rmat", scale=26), {"spx.preproc.xform": "none"}, dict(kind="rmat", scale=20)),
    "c5s": ("c5_rmat_23_scaled_down", dict(kind="rmat", scale=23), {"spx.preproc.xform": "none"}, dict(kind="rmat", scale=20)),
    "c4s": ("c4_sym_block_banded_3M_scaled_down", dict(kind="symbb", nb=1_000_000, b=1024), {"spx.matrix.symmetric": "true"},
            dict(kind="symbb", nb=200_000, b=1024)),
    "c2q": ("c2_poisson2d_2048_quarter_size", dict(kind="poisson2d", g=2048), {}, dict(kind="poisson2d", g=1024)),
    "small": ("small_poisson2d_512", dict(kind="poisson2d", g=512), {}, dict(kind="poisson2d", g=256)),
}


def generate(kind, **kw):
    from tests import matrices as M
    if kind == "poisson2d":
        return M.poisson2d(kw["g"])
    if kind == "stencil27":
        return M.stencil27(kw["g"])
    if kind == "symbb":
        return M.sym_block_banded(kw["nb"], b=kw["b"])
    if kind == "rmat":
        return M.rmat(kw["scale"])
    raise ValueError(kind)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pin_to_gpu_numa_node(index):
    """One process per GPU: run (and first-touch the pinned host buffers) on the cores NVML reports as local to the
    GPU, so that host<->device copies of different ranks do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and w * 64 + b < ncpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "cpus %d-%d (%d)" % (allowed[0], allowed[-1], len(allowed))
    except Exception as ex:  # affinity is an optimisation only
        return "not set: %r" % (ex,)
    return "not set"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch from the committed ncu --set full capture, if one exists for this workload."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(workload)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------- CPU arm --
def cpu_reference_run(sample_kw, opts, steps, warmup, threads=None):
    """The reference's CPU path on this box's host cores: CSX encoded with spx.rt.nr_threads = T,
    one persistent thread per partition.  Uses oracle/_ref (the reference's own kernel templates
    compiled by gcc) when it was built, else the oracle's port of the same unit loops."""
    from oracle.pyoracle import OracleMatrix
    T = threads or (os.cpu_count() or 1)
    rp, ci, va, n = generate(**sample_kw)
    o = dict(opts)
    o["spx.rt.nr_threads"] = T
    o["oracle.undefined_sampling"] = "break"
    t0 = time.time()
    M = OracleMatrix.from_csr(rp, ci, va, n, n).tune(o)
    tune_s = time.time() - t0
    nnz = int(rp[-1])
    x = np.random.default_rng(2).uniform(-1, 1, n)
    kind = "port"
    runner = M.bench
    try:
        from oracle import refkernels
        if refkernels.available():
            runner = refkernels.Runner(M).bench
            kind = "reference"
    except Exception:
        pass
    if warmup:
        runner(0.5, x, warmup)
    secs = runner(0.5, x, steps)
    return {"value": 2.0 * nnz * steps / secs / 1e9, "unit": "GFLOP/s", "cores": T, "kind": kind,
            "sample": "%s, %d rows, %d nnz, %d SpMVs, tune %.1f s, %s" % (sample_kw, n, nnz, steps, tune_s, M.log[:80]),
            "ms_per_step": secs / steps * 1e3, "nnz": nnz}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--e2e-steps", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: exchange fused into the SpMV kernel over peer memory (default) or NCCL after it")
    ap.add_argument("--graph-steps", type=int, default=16, help="N > 1, peer exchange: steps per captured CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name, gen_kw, opts, sample_kw = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(sample_kw, opts, max(args.steps, 1), args.warmup)
        line = {"impl": "reference", "metric": "csx_spmv_gflops", "value": r["value"], "unit": "GFLOP/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": name, "note": "CPU arm runs a bounded sample of the workload: " + r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import sparsex_b200
    from sparsex_b200.engine import CsxMatrix, SpxVector

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    numa_note = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api = sparsex_b200.load_spx_api()
    api.spx_init()

    # ---- load + tune through the public API (spx_input_load_csr / spx_mat_tune) ----
    t0 = time.time()
    rp, ci, va, n = generate(**gen_kw)
    gen_s = time.time() - t0
    nnz = int(rp[-1])
    all_opts = dict(opts)
    all_opts["spx.rt.nr_threads"] = world
    all_opts["spx.b200.rows_info"] = "false"
    for k, v in all_opts.items():
        api.spx_option_set(k.encode(), str(v).encode())
    api.spx_option_set(b"spx.b200.device", str(local_rank).encode())
    api.spx_option_set(b"spx.b200.part_lo", str(rank).encode())
    api.spx_option_set(b"spx.b200.part_hi", str(rank + 1).encode())
    t0 = time.time()
    inp = api.spx_input_load_csr(rp.ctypes.data, ci.ctypes.data, va.ctypes.data, n, n)
    A = api.spx_mat_tune(inp)
    if not A:
        raise SystemExit("spx_mat_tune failed")
    tune_s = time.time() - t0
    api.spx_input_destroy(inp)
    eng = CsxMatrix(api.spx_mat_get_engine(A))
    eng._h_owned = False
    part = eng.partition(0) if False else None  # values were kept on the host; only the log is needed below
    enc_log = sparsex_b200.lib().csxb_part_log(eng._h, 0).decode()
    row_lo = sparsex_b200.lib().csxb_part_info(eng._h, 0, 3)
    row_n = sparsex_b200.lib().csxb_part_info(eng._h, 0, 1)
    if str(opts.get("spx.matrix.symmetric", "false")) == "true":   # CSX-Sym partitions own dvalues.size() rows
        row_n = sparsex_b200.lib().csxb_part_info(eng._h, 0, 8)
    traffic = eng.traffic()
    del rp, ci, va

    # ---- device-resident timed region -----------------------------------------
    alpha = 0.1 if world > 1 else 0.5
    rng = np.random.default_rng(2)
    x = torch.from_numpy(rng.uniform(-1, 1, n)).cuda()
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    exchange = None
    xbuf = [x, y]   # ping-pong: the SpMV writes the own rows of the other buffer, the exchange fills the halo
    peer = None
    sym = str(opts.get("spx.matrix.symmetric", "false")) == "true"
    peer_note = None
    if world > 1 and args.exchange == "peer" and not sym:
        # the engine's own exchange: halo rows are stored into the neighbours' vectors by the SpMV kernel itself
        import sparsex_b200.dist as sdist
        peer, ranges, windows = sdist.connect_peer_exchange(eng, rank, world, "cuda")
        if peer is None:
            peer_note = "peer-memory exchange unavailable (%s): NCCL exchange used instead" % sdist.last_peer_error
    if peer is not None:
        peer.vector(0).copy_(x)
        exchange_kind = ("fused into the SpMV kernel: rows other ranks read are stored into their vectors over NVLink (peer memory), "
                         "device-side flags order the steps; protocol %d (1: edge tiles first, no sync kernel), %d edge tiles" % peer.protocol())
    if world > 1 and peer is None:
        symred = None
        from sparsex_b200.dist import PieceExchange, WindowExchange, gather_row_ranges
        L = sparsex_b200.lib()
        ranges = gather_row_ranges(row_lo, row_n, "cuda")
        win = torch.tensor([L.csxb_part_info(eng._h, 0, 11), L.csxb_part_info(eng._h, 0, 12)], dtype=torch.int64, device="cuda")
        allwin = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(allwin, win)
        windows = [(int(t[0]), int(t[1])) for t in allwin]
        if sym:   # CSX-Sym: transposed contributions to rows of lower ranks are sent to their owners and added
            from sparsex_b200.dist import SymHaloReduce
            hl = torch.tensor([L.csxb_info(eng._h, 8), L.csxb_info(eng._h, 9)], dtype=torch.int64, device="cuda")
            allh = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(allh, hl)
            symred = SymHaloReduce(ranges, [(int(t[0]), int(t[1])) for t in allh], rank, y)
            windows = [(w[0], max(w[1], r[0] + r[1] - 1)) for w, r in zip(windows, ranges)]
        wx = WindowExchange(ranges, windows, rank)
        frac = torch.tensor([wx.fraction], dtype=torch.float64, device="cuda")
        dist.all_reduce(frac, op=dist.ReduceOp.MAX)
        if float(frac[0]) < 0.5:
            exchange, exchange_kind = wx, "halo exchange (grouped NCCL send/recv of the column windows, %.2g%% of the vector)" % (100 * float(frac[0]))
        else:
            pieces = [PieceExchange(xbuf[0], ranges), PieceExchange(xbuf[1], ranges)]
            exchange_kind = "NCCL all-gather of the y pieces"
        if sym:
            exchange_kind = "CSX-Sym halo reduction to the owners (grouped NCCL send/recv + add), then " + exchange_kind
    state = {"cur": 0}

    def step():
        if peer is not None:
            peer.spmv(alpha)
            return
        src, dst = xbuf[state["cur"]], xbuf[1 - state["cur"]]
        eng.spmv(alpha, src, dst, overwrite=True)
        if world > 1:
            if sym and symred is not None:
                symred(dst)
            if exchange is not None:
                exchange(dst)
            else:  # dst's own rows -> every rank's dst
                pieces[1 - state["cur"]](dst[row_lo:row_lo + row_n])
            state["cur"] = 1 - state["cur"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1 with the peer exchange: a step is kernel launches only, so G steps are captured into one CUDA graph
    # (the step counter and the buffer parity live on the device, the graph is replayable)
    graph, G = None, 1
    if peer is not None and args.graph_steps > 1:
        G = args.graph_steps
        while G > 1 and (args.steps % G or G % 2):
            G -= 1
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if G > 1:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for _ in range(G):
                    step()
        torch.cuda.current_stream().wait_stream(side)
        barrier()
        graph.replay()   # one untimed replay
        barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    if graph is not None:
        for i in range(args.steps // G):
            graph.replay()
    else:
        for i in range(args.steps):
            kev[i][0].record()
            step()
            kev[i][1].record()
    e1.record()
    barrier()
    # the timed region can be a few milliseconds (N = 8: 4 ms): keep the same load running, untimed, until the
    # clock sampler has seen it for at least 100 ms
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 0.1:
        if graph is not None:
            graph.replay()
        else:
            for _ in range(16):
                step()
        torch.cuda.synchronize()
    barrier()
    clocks = sampler.finish()
    ms = e0.elapsed_time(e1)
    kernel_ms = ms / args.steps if graph is not None else sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kernel_ms = float(t[0]), float(t[1])
    value = 2.0 * nnz * args.steps / (ms * 1e-3) / 1e9
    if peer is not None and peer.error():
        raise SystemExit("peer exchange: a wait for a neighbour timed out")
    kernel_only_ms = None
    if world > 1:   # this rank's SpMV kernel alone (no exchange, no neighbour): what the step time is made of
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            eng.spmv(alpha, x, y, overwrite=True)
        ka.record()
        for _ in range(32):
            eng.spmv(alpha, x, y, overwrite=True)
        kb.record()
        torch.cuda.synchronize()
        t = torch.tensor([ka.elapsed_time(kb) / 32], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_only_ms = float(t[0])

    # ---- end to end through spx_matvec_mult with host buffers ---------------------
    xh = torch.from_numpy(rng.uniform(-1, 1, n)).pin_memory()
    yh = torch.zeros(n, dtype=torch.float64).pin_memory()
    vx = api.spx_vec_create_from_buff(xh.data_ptr(), None, n, None, 43)
    vy = api.spx_vec_create_from_buff(yh.data_ptr(), None, n, None, 43)
    for _ in range(2):
        api.spx_matvec_mult(alpha, A, vx, vy)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        if api.spx_matvec_mult(alpha, A, vx, vy) != 0:
            raise SystemExit("spx_matvec_mult failed")
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    L_ = sparsex_b200.lib()
    cw_lo, cw_hi = L_.csxb_part_info(eng._h, 0, 11), L_.csxb_part_info(eng._h, 0, 12)
    hb = torch.tensor([8.0 * max(0, cw_hi - cw_lo + 1), 8.0 * row_n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(hb)
    e2e = {"value": 2.0 * nnz * args.e2e_steps / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(hb[0]),
           "d2h_bytes_per_step": int(hb[1]), "steps": args.e2e_steps,
           "host_affinity": numa_note,
           "api": "spx_matvec_mult on spx_vec_create_from_buff vectors (pinned host memory); every rank uploads the "
                  "columns its partition reads and downloads its rows, slab-pipelined (H2D, kernels, D2H overlap)"}
    # check the device-resident result against the host-buffer path on the same x
    peer_sync_kernel = peer is not None and peer.protocol()[0] == 0   # protocol 0 ends every step with a sync kernel
    if peer is not None:
        barrier()
        peer.close()
    x.copy_(xh, non_blocking=False)
    y.zero_()
    eng.spmv(alpha, x, y, overwrite=True)
    torch.cuda.synchronize()
    dev = y[row_lo:row_lo + row_n].cpu().numpy()
    hostp = yh.numpy()[row_lo:row_lo + row_n]
    self_check = float(np.max(np.abs(dev - hostp)) / (np.max(np.abs(hostp)) + 1e-300))

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = traffic["total"] / (kernel_ms * 1e-3) / 1e9
        line = {"metric": "csx_spmv_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": name, "rows": n, "nnz": nnz, "options": {k: str(v) for k, v in all_opts.items()},
                           "encoding_rank0": enc_log.strip(), "alpha": alpha,
                           "step": "y = alpha*A*x (spx_matvec_mult semantics)" + (
                               "; then " + exchange_kind + " into the next x" if world > 1 else ""),
                           "l2_policy": "inputs larger than L2: %.0f MB of values+ctl per GPU vs 126 MB L2" % (
                               (traffic["values"] + traffic["ctl"]) / 1e6),
                           "exchange_note": peer_note, "tune_s": round(tune_s, 2), "generate_s": round(gen_s, 2),
                           "self_check_rel": self_check},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(name), "peak_source": peak_src,
                             "kernel": "csx_spmv_kernel (rank 0 partition)" + (
                                 "; N > 1: step time of the captured graph (SpMV kernel incl. fused exchange + flag kernel)"
                                 if graph is not None else ""), "kernel_ms": kernel_ms, "kernel_only_ms": kernel_only_ms,
                             "algorithmic_bytes": traffic["total"],
                             "bytes": {k: traffic[k] for k in ("values", "ctl", "tables", "x", "y")},
                             "frac_of_8TBs_nominal": achieved / 8000.0},
                "e2e": e2e, "gpu_launches": (int(traffic["launches"]) + (1 if peer_sync_kernel else 0)) * args.steps, "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_reference_run(sample_kw, opts, 32, 2)
                line["cpu_baseline"] = {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"],
                                        "sample": r["sample"]}
            except Exception as ex:  # the CPU baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "port",
                                        "sample": "failed: %r" % (ex,)}
        print(json.dumps(line))
    api.spx_vec_destroy(vx)
    api.spx_vec_destroy(vy)
    eng._h = None  # owned by the spx matrix handle
    api.spx_mat_destroy(A)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
