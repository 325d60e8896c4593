"""Host-side (numpy, vectorised) generators of the benchmark matrices that have no device generator: every function
returns the CSR arrays of the global rows [r0, r1) with GLOBAL column indices, ready for MatCreateB200CSR, so that each
rank builds only the rows it owns (the pattern of the `Istart/Iend` loops in the reference's example programs).

  markov_rows        MatMarkovModel, src/eps/tutorials/ex5.c:138-170 (N = m(m+1)/2; transposed stochastic matrix)
  random_sparse_rows SURVEY.md §8d config C5: `nnz_row` column indices per row, uniform i.i.d. from a counter-based
                     generator indexed by (seed, row, slot), duplicates merged, values ~N(0,1) from the same stream
  laplacian_rows     ex1.c:37-48 / ex2.c:39-54 and the 7-point analogue (the device generator b2k_csr_laplacian builds
                     the same matrix in HBM; this one is for host-CSR entry points and small tests)
"""
import numpy as np


def _csr_from_coo(nrows, r, c, v, sum_duplicates=False):
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    if sum_duplicates and r.size:
        first = np.ones(r.size, bool)
        first[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
        idx = np.cumsum(first) - 1
        v = np.bincount(idx, weights=v, minlength=int(idx[-1]) + 1)
        r, c = r[first], c[first]
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    np.cumsum(rowptr, out=rowptr)
    return rowptr.astype(np.int32), c.astype(np.int32), np.ascontiguousarray(v, dtype=np.float64)


def _csr_from_rowslots(cols, vals, valid=None, sum_duplicates=False):
    """CSR from a fixed number of candidate entries per row (cols/vals/valid are nrows x w): every row is sorted by
    column on its own (argsort along axis 1 — no global sort), invalid slots dropped, duplicates optionally merged."""
    nrows, w = cols.shape
    big = np.iinfo(np.int64).max
    key = cols.astype(np.int64, copy=True)
    if valid is not None:
        key[~valid] = big
    order = np.argsort(key, axis=1, kind="stable")
    key = np.take_along_axis(key, order, axis=1)
    vals = np.take_along_axis(vals, order, axis=1)
    keep = key != big
    if sum_duplicates and w > 1:
        dup = np.zeros_like(keep)
        dup[:, 1:] = (key[:, 1:] == key[:, :-1]) & keep[:, 1:]
        if dup.any():
            # add every run of equal columns into its first slot (runs are short: walk right to left)
            for s in range(w - 1, 0, -1):
                d = dup[:, s]
                if d.any():
                    vals[d, s - 1] += vals[d, s]
            keep &= ~dup
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(keep.sum(axis=1), out=rowptr[1:])
    assert rowptr[-1] < 2 ** 31, "local nnz exceeds PetscInt (int32)"
    return rowptr.astype(np.int32), key[keep].astype(np.int32), np.ascontiguousarray(vals[keep], dtype=np.float64)


def markov_size(m):
    return m * (m + 1) // 2


def markov_rows(m, r0=0, r1=None):
    """rows [r0,r1) of MatMarkovModel(m): (rowptr, colidx, val) with global columns (ex5.c:138-170)"""
    N = markov_size(m)
    r1 = N if r1 is None else r1
    cst = 0.5 / (m - 1)
    # start[i-1] = 0-based index of the first row of the block with outer index i (1-based), block length m-i+1
    start = np.concatenate(([0], np.cumsum(m - np.arange(1, m + 1) + 1)))
    ix0 = np.arange(r0, r1, dtype=np.int64)                     # 0-based row = ix-1
    i = np.searchsorted(start, ix0, side="right")                # 1-based outer index
    jmax = m - i + 1
    j = ix0 - start[i - 1] + 1                                   # 1-based inner index
    inner = j != jmax
    pd = cst * (i + j - 1)
    pu = 0.5 - cst * (i + j - 3)
    cols = np.stack([ix0 + 1, ix0 + jmax, ix0 - 1, ix0 - jmax - 1], axis=1)              # north, east, south, west
    vals = np.stack([np.where(i == 1, 2 * pd, pd), np.where(j == 1, 2 * pd, pd), pu, pu], axis=1)
    valid = np.stack([inner, inner, j > 1, i > 1], axis=1)
    return _csr_from_rowslots(cols, vals, valid)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def random_sparse_rows(M, N, nnz_row=20, seed=20261017, r0=0, r1=None):
    """rows [r0,r1) of the M x N synthetic matrix of config C5; reproducible per (seed,row,slot) whatever the partition"""
    r1 = M if r1 is None else r1
    nloc = r1 - r0
    step = 1 << 20                                   # rows per chunk: bounds the temporaries; chunks run on a thread pool

    def chunk(a):
        b = min(a + step, r1)
        with np.errstate(over="ignore"):
            row = np.arange(a, b, dtype=np.uint64)[:, None]
            slot = np.arange(nnz_row, dtype=np.uint64)[None, :]
            key = _splitmix64(row * np.uint64(0x100000001B3) + slot + (np.uint64(seed) << np.uint64(20)))
            col = (_splitmix64(key) % np.uint64(N)).astype(np.int64)
            u1 = ((_splitmix64(key ^ np.uint64(0xA5A5A5A5A5A5A5A5)) >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
            u2 = ((_splitmix64(key ^ np.uint64(0x5A5A5A5A5A5A5A5A)) >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
        val = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)       # Box-Muller
        return _csr_from_rowslots(col, val, None, sum_duplicates=True)

    starts = list(range(r0, r1, step))
    if len(starts) > 1:
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1, len(starts)))) as pool:
            chunks = list(pool.map(chunk, starts))   # numpy releases the GIL inside its loops
    else:
        chunks = [chunk(a) for a in starts]
    if not chunks:
        return np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0)
    if len(chunks) == 1:
        return chunks[0]
    nnz = np.cumsum([0] + [int(c[0][-1]) for c in chunks])
    assert nnz[-1] < 2 ** 31, "local nnz exceeds PetscInt (int32)"
    rowptr = np.concatenate([chunks[0][0]] + [c[0][1:] + np.int32(o) for c, o in zip(chunks[1:], nnz[1:-1])])
    return rowptr, np.concatenate([c[1] for c in chunks]), np.concatenate([c[2] for c in chunks])


def laplacian_rows(dim, nx, ny=1, nz=1, r0=0, r1=None):
    """rows [r0,r1) of the dim-dimensional Laplacian stencil, grid nx (slowest) x ny x nz (fastest), natural ordering"""
    if dim < 3:
        nz = 1
    if dim < 2:
        ny = 1
    plane = ny * nz
    N = nx * plane
    r1 = N if r1 is None else r1
    g = np.arange(r0, r1, dtype=np.int64)
    i, rem = g // plane, g % plane
    j, k = rem // nz, rem % nz
    cand = [(g - plane, i > 0)]
    if dim >= 2:
        cand.append((g - nz, j > 0))
    if dim >= 3:
        cand.append((g - 1, k > 0))
    cand.append((g, np.ones(g.size, bool)))
    if dim >= 3:
        cand.append((g + 1, k < nz - 1))
    if dim >= 2:
        cand.append((g + nz, j < ny - 1))
    cand.append((g + plane, i < nx - 1))
    cols = np.stack([c for c, _ in cand], axis=1)
    valid = np.stack([m for _, m in cand], axis=1)
    vals = np.full(cols.shape, -1.0)
    vals[:, len(cand) // 2] = 2.0 * dim
    rowptr = np.zeros(g.size + 1, dtype=np.int32)
    np.cumsum(valid.sum(axis=1), out=rowptr[1:])
    return rowptr, cols[valid].astype(np.int32), np.ascontiguousarray(vals[valid])


def hash_uniform(idx, seed):
    """uniform [-1,1) value of global index `idx` (uint64 array) under `seed`: the splitmix64-finaliser hash of b2k_set_random
    (b2k_bv.cu b2k_hash_uniform) evaluated on the host, bit for bit — lets a caller reproduce a device-generated vector"""
    with np.errstate(over="ignore"):
        x = (np.asarray(idx, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed) * np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return 2.0 * ((x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)) - 1.0
