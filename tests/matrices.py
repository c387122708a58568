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


# ---- row-slab generators: the same matrices, any row range on its own ------------------------------------------
# One process per GPU cannot hold a 30 M-row or a 2^26-row matrix per rank: every rank generates the rows of its own
# partition.  Values are a hash of the coordinates, so that any slab can be produced independently (and a symmetric
# matrix is symmetric whoever generates which half).

def _mix64(z):
    """splitmix64 finaliser on uint64 arrays."""
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _unit(z):
    """uint64 hash -> float64 in [0, 1)."""
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def stencil_rows(kind, g, lo, hi, perturb=True):
    """Rows [lo, hi) of poisson2d(g) ('p2') or stencil27(g) ('s27') with hash-perturbed values: (rowptr, colind, values)."""
    i = np.arange(lo, hi, dtype=np.int64)
    if kind == "p2":
        n = g * g
        gx, gy = i % g, i // g
        offs = [((gx + dx >= 0) & (gx + dx < g) & (gy + dy >= 0) & (gy + dy < g), dy * g + dx, 4.0 if (dx, dy) == (0, 0) else -1.0)
                for dy, dx in ((-1, 0), (0, -1), (0, 0), (0, 1), (1, 0))]
    else:
        n = g ** 3
        gx, gy, gz = i % g, (i // g) % g, i // (g * g)
        offs = []
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    ok = ((gx + dx >= 0) & (gx + dx < g) & (gy + dy >= 0) & (gy + dy < g) & (gz + dz >= 0) & (gz + dz < g))
                    offs.append((ok, (dz * g + dy) * g + dx, 26.0 if (dx, dy, dz) == (0, 0, 0) else -1.0))
    k = len(offs)
    mask = np.empty((hi - lo, k), bool)
    cols = np.empty((hi - lo, k), np.int64)
    vals = np.empty((hi - lo, k), np.float64)
    for j, (ok, off, val) in enumerate(offs):
        mask[:, j] = ok
        cols[:, j] = i + off
        vals[:, j] = val
    rowptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    ci = cols[mask]
    va = vals[mask]
    if perturb:
        rows = np.repeat(i, mask.sum(axis=1))
        va = va * (1 + 1e-3 * _unit(_mix64(rows.astype(np.uint64) * np.uint64(n) + ci.astype(np.uint64))))
    return rowptr, ci.astype(np.int32), va


def stencil_row_counts(kind, g):
    """Non-zeros of every row of poisson2d(g) ('p2') or stencil27(g) ('s27')."""
    if kind == "p2":
        e = np.full(g, 3, np.int64); e[[0, -1]] = 2       # neighbours along one axis incl. self
        gx = np.tile(e, g); gy = np.repeat(e, g)
        return gx + gy - 1                                  # cross: x line + y line - centre counted twice
    e = np.full(g, 3, np.int64); e[[0, -1]] = 2
    return (e[None, None, :] * e[None, :, None] * e[:, None, None]).reshape(-1)


def symbb_rows(nb, b, lo, hi, bs=3):
    """Rows [lo, hi) of the symmetric block-banded matrix (config 4): nb block rows of bs x bs dense blocks, block (I, J)
    present for |I - J| in {0, 1, b}; value(r, c) = value(c, r) from a hash of the pair, diagonal +10."""
    n = nb * bs
    r = np.arange(lo, hi, dtype=np.int64)
    I = r // bs
    offs = (-b, -1, 0, 1, b)
    k = len(offs) * bs
    cols = np.empty((hi - lo, k), np.int64)
    mask = np.empty((hi - lo, k), bool)
    for t, off in enumerate(offs):
        J = I + off
        ok = (J >= 0) & (J < nb)
        for q in range(bs):
            cols[:, t * bs + q] = J * bs + q
            mask[:, t * bs + q] = ok
    rowptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    ci = cols[mask]
    rows = np.repeat(r, mask.sum(axis=1))
    a, c = np.minimum(rows, ci).astype(np.uint64), np.maximum(rows, ci).astype(np.uint64)
    va = 2.0 * _unit(_mix64(a * np.uint64(n) + c)) - 1.0
    va[rows == ci] += 10.0
    return rowptr, ci.astype(np.int32), va


def symbb_row_counts(nb, b, bs=3):
    I = np.arange(nb, dtype=np.int64)
    per = sum(((I + off >= 0) & (I + off < nb)).astype(np.int64) for off in (-b, -1, 0, 1, b)) * bs
    return np.repeat(per, bs)


def split_rows(counts, nparts, lower_counts=None):
    """Row ranges [(first row, rows)] of the reference's nnz-balanced split (SparseInternal.hpp:131-144,
    SparsePartition.hpp:519-534) from the row lengths.  lower_counts (CSX-Sym, SparsePartition.hpp:1103-1108): per row the
    number of entries left of the diagonal; every row is assumed to hold its diagonal entry."""
    n = counts.size
    out = []
    if lower_counts is None:
        rowptr = np.concatenate([[0], np.cumsum(counts)])
        total, done, row_start = int(rowptr[-1]), 0, 0
        nonempty = np.nonzero(counts)[0]
        for i in range(nparts):
            limit = (total - done) // (nparts - i)
            pos_end = total
            if limit and done + limit < total:
                pos_end = int(rowptr[np.searchsorted(rowptr, done + limit, "left")])
            # rows up to the last non-empty row below pos_end (trailing empty rows go to the next partition)
            j = np.searchsorted(rowptr, pos_end, "left")        # first row whose start is >= pos_end
            k = np.searchsorted(nonempty, j, "left")            # non-empty rows before j
            last = int(nonempty[k - 1]) if k > 0 and pos_end > done else row_start - 1
            rows = max(0, last - row_start + 1)
            out.append((row_start, rows))
            row_start += rows
            done = pos_end
        return out
    s = lower_counts.astype(np.int64) + 1                      # lower entries + diagonal
    cum = np.concatenate([[0], np.cumsum(s)])
    total, done, row_start = int(cum[-1]), 0, 0
    has_lower = lower_counts > 0
    for i in range(nparts):
        limit = (total - done) // (nparts - i)
        if limit == 0 or i == nparts - 1:
            end = n
        else:
            # first row j > row_start with lower entries, preceded by a row with lower entries (or being row_start + 1 ...),
            # at which the elements taken so far reach the limit
            j = int(np.searchsorted(cum, done + limit, "left"))
            j = max(j, row_start + 1)
            while j < n and not (has_lower[j] and (j - 1 == row_start or has_lower[j - 1] or j - 1 < row_start)):
                j += 1
            end = min(j, n)
        out.append((row_start, end - row_start))
        done = int(cum[end])
        row_start = end
    return out


# R-MAT by row blocks (config 5 at full size): the rows are cut into 2^K blocks by their top K bits; a block holds its
# expected share of the edges, drawn with the row prefix fixed and the column bits conditional on it — the same
# distribution as rmat(), but any block can be generated on its own (torch: on the rank's GPU in a fraction of a second).
def _rmat_k(scale):
    return max(0, min(10, scale - 6))


def _rmat_block(scale, blk, edge_factor, abcd, seed, device):
    """Sorted unique keys row * n + col (int64 tensor) of row block `blk`."""
    import torch
    a, b, c, d = abcd
    K = _rmat_k(scale)
    n = 1 << scale
    p = 1.0
    for lvl in range(K):
        p *= (c + d) if (blk >> (K - 1 - lvl)) & 1 else (a + b)
    m = int(round(n * edge_factor * p))
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + blk)
    r = torch.zeros(m, dtype=torch.int64, device=device)
    cc = torch.zeros(m, dtype=torch.int64, device=device)
    for lvl in range(scale):
        u = torch.rand(m, generator=g, device=device, dtype=torch.float32)
        if lvl < K:
            rb = (blk >> (K - 1 - lvl)) & 1
            cbit = u < (d / (c + d) if rb else b / (a + b))
            r = (r << 1) | rb
        else:
            r = (r << 1) | (u >= a + b).to(torch.int64)
            cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        cc = (cc << 1) | cbit.to(torch.int64)
    return torch.unique(r * n + cc)   # sorted


def rmat_block_row_counts(scale, edge_factor=16, abcd=(0.57, 0.19, 0.19, 0.05), seed=42, device=None):
    import torch
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    n = 1 << scale
    bsz = n >> _rmat_k(scale)
    counts = torch.zeros(n, dtype=torch.int64, device=device)
    for blk in range(1 << _rmat_k(scale)):
        key = _rmat_block(scale, blk, edge_factor, abcd, seed, device)
        counts[blk * bsz:(blk + 1) * bsz] = torch.bincount(key // n - blk * bsz, minlength=bsz)
    return counts.cpu().numpy()


def rmat_block_rows(scale, lo, hi, edge_factor=16, abcd=(0.57, 0.19, 0.19, 0.05), seed=42, device=None):
    """Rows [lo, hi) of the block-wise R-MAT matrix: (rowptr, colind, values), values U(-1, 1) from a hash of the coordinates."""
    import torch
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    n = 1 << scale
    K = _rmat_k(scale)
    bsz = n >> K
    keys = []
    for blk in range(lo // bsz, (hi - 1) // bsz + 1 if hi > lo else 0):
        key = _rmat_block(scale, blk, edge_factor, abcd, seed, device)
        key = key[(key >= lo * n) & (key < hi * n)]
        keys.append(key.cpu().numpy())
    key = np.concatenate(keys) if keys else np.zeros(0, np.int64)
    rows, cols = key // n, key % n
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows - lo, minlength=hi - lo))]).astype(np.int32)
    va = 2.0 * _unit(_mix64(key.astype(np.uint64))) - 1.0
    return rowptr, cols.astype(np.int32), va
