"""Synthetic matrix generators of BASELINE.json's configs, CSR, zero-based, int32."""
import numpy as np


def _csr_from_coo(r, c, v, n, m):
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), c.astype(np.int32), v.astype(np.float64)


def _stencil_csr(n, offsets_cols_vals, perturb, seed):
    """Rows are emitted in natural order with ascending columns: no sort needed."""
    k = len(offsets_cols_vals)
    cols = np.empty((n, k), np.int64)
    vals = np.empty((n, k), np.float64)
    mask = np.empty((n, k), bool)
    for j, (ok, col, val) in enumerate(offsets_cols_vals):
        mask[:, j] = ok
        cols[:, j] = col
        vals[:, j] = val
    rowptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    ci = cols[mask].astype(np.int32)
    va = vals[mask]
    if perturb:
        va = va * (1 + 1e-3 * np.random.default_rng(seed).random(va.size))
    return rowptr, ci, va, n


def poisson2d(g, perturb=True, seed=1):
    """5-point Laplacian on a g x g grid, natural ordering (config 2: g = 4096)."""
    n = g * g
    i = np.arange(n, dtype=np.int64)
    gx, gy = i % g, i // g
    spec = []
    for dy, dx, val in ((-1, 0, -1.0), (0, -1, -1.0), (0, 0, 4.0), (0, 1, -1.0), (1, 0, -1.0)):
        ok = (gx + dx >= 0) & (gx + dx < g) & (gy + dy >= 0) & (gy + dy < g)
        spec.append((ok, i + dy * g + dx, val))
    return _stencil_csr(n, spec, perturb, seed)


def stencil27(g, perturb=True, seed=1):
    """27-point stencil on a g^3 grid, natural ordering (config 3: g = 256)."""
    n = g ** 3
    i = np.arange(n, dtype=np.int64)
    gx, gy, gz = i % g, (i // g) % g, i // (g * g)
    spec = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = ((gx + dx >= 0) & (gx + dx < g) & (gy + dy >= 0) & (gy + dy < g) & (gz + dz >= 0) & (gz + dz < g))
                spec.append((ok, i + (dz * g + dy) * g + dx, 26.0 if (dx, dy, dz) == (0, 0, 0) else -1.0))
    return _stencil_csr(n, spec, perturb, seed)


def sym_block_banded(nb, b=16, bs=3, seed=3):
    """Symmetric block-banded matrix (config 4): nb block rows of bs x bs dense blocks, block (I,J) present
    for |I-J| in {0, 1, b}; symmetric values, diagonally dominant.  Full matrix is returned."""
    rng = np.random.default_rng(seed)
    n = nb * bs
    rows, cols, vals = [], [], []
    I = np.arange(nb, dtype=np.int64)
    for off in (0, 1, b):
        J = I - off
        ok = J >= 0
        Ib, Jb = I[ok], J[ok]
        blk = rng.uniform(-1, 1, (Ib.size, bs, bs))
        if off == 0:
            blk = 0.5 * (blk + blk.transpose(0, 2, 1))
            blk[:, np.arange(bs), np.arange(bs)] += 10.0
        a, bb = np.meshgrid(np.arange(bs), np.arange(bs), indexing="ij")
        rr = (Ib[:, None, None] * bs + a[None]).ravel()
        cc = (Jb[:, None, None] * bs + bb[None]).ravel()
        vv = blk.ravel()
        rows.append(rr); cols.append(cc); vals.append(vv)
        if off:
            rows.append(cc); cols.append(rr); vals.append(vv)
    r, c, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    rp, ci, va = _csr_from_coo(r, c, v, n, n)
    return rp, ci, va, n


def rmat(scale, edge_factor=16, abcd=(0.57, 0.19, 0.19, 0.05), seed=42):
    """R-MAT power-law matrix (config 5: scale 26), duplicates removed, values U(-1,1)."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    ne = n * edge_factor
    r = np.zeros(ne, np.int64)
    c = np.zeros(ne, np.int64)
    a, b, cc, _ = abcd
    for _ in range(scale):
        u = rng.random(ne)
        rbit = (u >= a + b).astype(np.int64)
        cbit = (((u >= a) & (u < a + b)) | (u >= a + b + cc)).astype(np.int64)
        r = (r << 1) | rbit
        c = (c << 1) | cbit
    key = np.unique(r * n + c)
    r, c = key // n, key % n
    v = rng.uniform(-1, 1, r.size)
    rp, ci, va = _csr_from_coo(r, c, v, n, n)
    return rp, ci, va, n


def random_structured(rng, n, m, symmetric=False):
    """Small random matrix seeded with runs of every substructure kind (parity stress input)."""
    S = set()

    def add(r, c):
        if 0 <= r < n and 0 <= c < m:
            S.add((int(r), int(c)))
    for _ in range(n * 2):
        add(rng.integers(n), rng.integers(m))
    for _ in range(n // 4 + 1):
        r, c, L, d, k = rng.integers(n), rng.integers(m), rng.integers(2, 40), rng.integers(1, 4), rng.integers(7)
        if k == 0:
            for i in range(L): add(r, c + i * d)
        elif k == 1:
            for i in range(L): add(r + i * d, c)
        elif k == 2:
            for i in range(L): add(r + i * d, c + i * d)
        elif k == 3:
            for i in range(L): add(r + i * d, c - i * d)
        else:
            br, bc = rng.integers(1, 9), rng.integers(1, 12)
            for i in range(br):
                for j in range(bc): add(r + i, c + j)
    if symmetric:
        S |= {(c, r) for (r, c) in S}
        S |= {(i, i) for i in range(n)}
        if n >= 2:  # CSX-Sym needs a sub-diagonal entry in the last row
            S |= {(n - 1, n - 2), (n - 2, n - 1)}
    a = np.array(sorted(S), dtype=np.int64)
    r, c = a[:, 0], a[:, 1]
    v = rng.standard_normal(r.size)
    if symmetric:
        lo, hi = np.minimum(r, c), np.maximum(r, c)
        key = lo * m + hi
        _, inv = np.unique(key, return_inverse=True)
        v = rng.standard_normal(inv.max() + 1)[inv]
    rp, ci, va = _csr_from_coo(r, c, v, n, m)
    return rp, ci, va
