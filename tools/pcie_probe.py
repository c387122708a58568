"""What the host link gives (tuning aid for csxb_spmv_host): 128 MiB up, down, both at once, whole and in 4 MiB pieces."""
import time
import torch

n = 16 * 1024 * 1024
hx = torch.empty(n, dtype=torch.float64).pin_memory()
hy = torch.empty(n, dtype=torch.float64).pin_memory()
hx.uniform_(-1, 1)
dx = torch.empty(n, dtype=torch.float64, device="cuda")
dy = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def up(piece):
    with torch.cuda.stream(s1):
        for o in range(0, n, piece):
            dx[o:o + piece].copy_(hx[o:o + piece], non_blocking=True)


def down(piece):
    with torch.cuda.stream(s2):
        for o in range(0, n, piece):
            hy[o:o + piece].copy_(dy[o:o + piece], non_blocking=True)


for piece in (n, n // 32, n // 128):
    mb = piece * 8 / 2**20
    a = run(lambda: up(piece)); b = run(lambda: down(piece)); c = run(lambda: (up(piece), down(piece)))
    print("pieces of %6.1f MiB: up %.2f ms (%.1f GB/s)  down %.2f ms (%.1f GB/s)  both at once %.2f ms (%.1f GB/s per direction)"
          % (mb, a, n * 8 / a / 1e6, b, n * 8 / b / 1e6, c, n * 8 / c / 1e6), flush=True)
