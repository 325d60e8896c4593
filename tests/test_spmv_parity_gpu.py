"""Element-wise parity of the DEFAULT hot SpMV kernel — k_spmv_sell_pipe<false> and <true> (b2k_spmv.cu), the bulk-copy
pipeline over the SELL-32 copy that PETSc's MatMult behind bvops.c:879 / stsolve.c:22 is replaced by — against independent
answers, with the kernel that ran ASSERTED through b2k_csr_last_kernel:

  * closed-form stencils  y = 2d x - (shifted copies of x)  evaluated by numpy in O(n) at the full sizes of BASELINE.json
    (4096^2 = 16.8 M rows, 512^3 = 134 M rows) with hash-random x;
  * a 512^3 slab with BOTH ghost planes (what a middle rank of an 8-GPU run multiplies): the <true> instantiation;
  * scipy on a 1 M-row random matrix with 20 draws per row (with and without ghost columns) and on the Markov model m = 1200;
  * small irregular matrices forced through the pipeline (b2k_spmv_set_pipe_min_chunks): widths > 8, empty rows, chunks made
    only of empty slices (zero-byte stages), odd slice tails, ghost columns.

Tolerance (written here, from the verdict): max|y - y_ref| <= 1e-13 * ||x||_inf * max nnz per row * max|a_ij|.
Nothing here is skipped for lack of memory on a B200: a full-size test that cannot allocate FAILS.
"""
import ctypes

import numpy as np
import pytest

from oracle import slepc_oracle as O

pytestmark = pytest.mark.gpu

K_CSR, K_SELL, K_PIPE, K_PIPE_GHOST = 1, 2, 3, 4


def _check(rc):
    from slepc_b200._b2k import check
    check(rc)


def _last_kernel(ctx, h):
    w = ctypes.c_int()
    _check(ctx.lib.b2k_csr_last_kernel(h, ctypes.byref(w)))
    return w.value


def _is_b200(ctx):
    fr, to = ctypes.c_size_t(), ctypes.c_size_t()
    _check(ctx.lib.b2k_mem_info(ctx.h, ctypes.byref(fr), ctypes.byref(to)))
    return to.value >= 150 * 2 ** 30, fr.value


def _need(ctx, gb):
    """a full-size case needs `gb` GB of HBM: on a B200 (>= 150 GB) that is an assertion, elsewhere a skip"""
    big, free = _is_b200(ctx)
    if free >= gb * 2 ** 30:
        return
    if big:
        pytest.fail(f"needs {gb} GB of free HBM but only {free / 2**30:.1f} GB are free on a B200: full-size parity must not be skipped")
    pytest.skip(f"needs {gb} GB of free HBM (not a B200)")


def _hash_x(ctx, n, row0=0, seed=7):
    """hash-random x in [-1,1) generated on the device (bit-identical to oracle hash_uniform), returned as (device, host)"""
    dx = ctx.empty(n)
    _check(ctx.lib.b2k_set_random(ctx.h, dx.ptr, n, row0, seed))
    return dx, dx.to_host()


def _stencil_ref(xfull, dim, shape, p0, p1):
    """rows of planes [p0,p1) of the d-dimensional Laplacian stencil applied to xfull (planes [p0-1 or 0, p1+1 or nx) given as
    a (planes, ny, nz) array whose first plane is global plane q0); returns the flat y of the owned planes"""
    nx, ny, nz = shape
    q0 = max(p0 - 1, 0)
    X = xfull.reshape(-1, ny, nz)
    a, b = p0 - q0, p1 - q0                      # owned planes inside X
    Y = (2.0 * dim) * X[a:b]
    lo = X[a - 1:b - 1] if a > 0 else np.concatenate([np.zeros((1, ny, nz)), X[a:b - 1]])
    Y -= lo
    hi = X[a + 1:b + 1] if b < X.shape[0] else np.concatenate([X[a + 1:b], np.zeros((1, ny, nz))])
    Y -= hi
    if dim >= 2:
        Y[:, 1:, :] -= X[a:b, :-1, :]
        Y[:, :-1, :] -= X[a:b, 1:, :]
    if dim >= 3:
        Y[:, :, 1:] -= X[a:b, :, :-1]
        Y[:, :, :-1] -= X[a:b, :, 1:]
    return Y.ravel()


def _run_laplacian(ctx, dim, shape, p0, p1, expect, seed=7):
    nx, ny, nz = shape
    plane = ny * nz
    row0, nrows = p0 * plane, (p1 - p0) * plane
    h = ctypes.c_void_p()
    glo, ghi = ctypes.c_int64(), ctypes.c_int64()
    _check(ctx.lib.b2k_csr_laplacian(ctx.h, dim, nx, ny, nz, row0, nrows, ctypes.byref(h), ctypes.byref(glo), ctypes.byref(ghi)))
    q0, q1 = max(p0 - 1, 0), min(p1 + 1, nx)
    dxall, xall = _hash_x(ctx, (q1 - q0) * plane, q0 * plane, seed)       # owned planes + the neighbour planes, one hash stream
    off = (p0 - q0) * plane
    dg = ctx.empty(max(glo.value + ghi.value, 1))
    if glo.value:
        _check(ctx.lib.b2k_d2d(ctx.h, dg.ptr, dxall.ptr, 8 * plane))
    if ghi.value:
        _check(ctx.lib.b2k_d2d(ctx.h, dg.at(glo.value), dxall.at(off + nrows), 8 * plane))
    dy = ctx.empty(nrows)
    _check(ctx.lib.b2k_csr_spmv(ctx.h, h, dxall.at(off), dg.ptr, dy.ptr))
    ctx.sync()
    kern = _last_kernel(ctx, h)
    y = dy.to_host()
    for d in (dy, dg, dxall):
        d.free()
    _check(ctx.lib.b2k_csr_destroy(ctx.h, h))
    ref = _stencil_ref(xall, dim, shape, p0, p1)
    tol = 1e-13 * np.abs(xall).max() * (2 * dim + 1) * (2.0 * dim)
    err = float(np.abs(y - ref).max())
    assert kern == expect, f"dispatch ran kernel {kern}, the test is meant to cover {expect}"
    assert err <= tol, (err, tol)
    return err


# ---- full-size closed-form checks (BASELINE.json configs[1], configs[2]) -------------------------------------------------
def test_pipe_full_size_2d_4096(ctx):
    _need(ctx, 4)
    _run_laplacian(ctx, 2, (4096, 4096, 1), 0, 4096, K_PIPE)


def test_pipe_full_size_3d_512(ctx):
    _need(ctx, 20)
    _run_laplacian(ctx, 3, (512, 512, 512), 0, 512, K_PIPE)


@pytest.mark.parametrize("p0,p1", [(64, 128), (192, 448)], ids=["slab_of_8", "slab_of_2_wide"])
def test_pipe_ghost_slab_3d_512_both_planes(ctx, p0, p1):
    """a middle slab of the 512^3 grid (what rank r of an N-GPU run owns): lower AND upper ghost plane, <GHOST=true>"""
    _need(ctx, 12)
    _run_laplacian(ctx, 3, (512, 512, 512), p0, p1, K_PIPE_GHOST)


@pytest.mark.parametrize("p0,p1", [(0, 2048), (2048, 4096), (1024, 3072)], ids=["first", "last", "middle"])
def test_pipe_ghost_slab_2d_4096(ctx, p0, p1):
    _run_laplacian(ctx, 2, (4096, 4096, 1), p0, p1, K_PIPE_GHOST)


# ---- scipy comparisons at sizes that reach the pipeline by themselves -----------------------------------------------------
def _spmv_host_csr(ctx, rp, ci, va, nrows, ncl, x, xg=None, sigma=None):
    h = ctypes.c_void_p()
    ng = 0 if xg is None else len(xg)
    rp = np.ascontiguousarray(rp, dtype=np.int32); ci = np.ascontiguousarray(ci, dtype=np.int32); va = np.ascontiguousarray(va, dtype=np.float64)
    _check(ctx.lib.b2k_csr_create(ctx.h, nrows, ncl, ng, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, ctypes.byref(h)))
    dx = ctx.to_device(np.ascontiguousarray(x))
    dg = ctx.to_device(np.ascontiguousarray(xg)) if ng else None
    dy = ctx.empty(max(nrows, 1))
    if sigma is None:
        _check(ctx.lib.b2k_csr_spmv(ctx.h, h, dx.ptr, dg.ptr if dg else None, dy.ptr))
    else:
        _check(ctx.lib.b2k_csr_spmv_shift(ctx.h, h, dx.ptr, dg.ptr if dg else None, dy.ptr, sigma))
    ctx.sync()
    kern = _last_kernel(ctx, h)
    y = dy.to_host()[:nrows]
    nb = ctypes.c_int64()
    _check(ctx.lib.b2k_csr_bytes(h, ctypes.byref(nb)))
    _check(ctx.lib.b2k_csr_destroy(ctx.h, h))
    for d in (dx, dg, dy):
        if d is not None:
            d.free()
    return y, kern, nb.value


@pytest.mark.parametrize("ghost", [False, True], ids=["owned_columns", "ghost_columns"])
def test_pipe_random_1m_rows_20_per_row_vs_scipy(ctx, ghost):
    import scipy.sparse as sp
    from slepc_b200 import matgen
    M, N = 1 << 20, 900001
    rp, ci, va = matgen.random_sparse_rows(M, N, 20, seed=5)
    A = sp.csr_matrix((va, ci, rp), shape=(M, N))
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, N)
    ncl = 600000 if ghost else N
    y, kern, nbytes = _spmv_host_csr(ctx, rp, ci, va, M, ncl, x[:ncl], x[ncl:] if ghost else None)
    ref = A @ x
    nnz_row = int(np.diff(rp).max())
    tol = 1e-13 * np.abs(x).max() * nnz_row * np.abs(va).max()
    assert kern == (K_PIPE_GHOST if ghost else K_PIPE)
    assert np.abs(y - ref).max() <= tol
    # the matrix is stored ONCE (SELL copy + row pointer), not CSR + SELL
    assert nbytes < 1.3 * 12 * A.nnz + 16 * M


def test_pipe_markov_m1200_vs_scipy(ctx):
    import scipy.sparse as sp
    from slepc_b200 import matgen
    m = 1200
    N = matgen.markov_size(m)
    rp, ci, va = matgen.markov_rows(m)
    A = sp.csr_matrix((va, ci, rp), shape=(N, N))
    x = np.random.default_rng(4).uniform(-1, 1, N)
    y, kern, _ = _spmv_host_csr(ctx, rp, ci, va, N, N, x)
    assert kern == K_PIPE
    assert np.abs(y - A @ x).max() <= 1e-13 * 5
    y, kern, _ = _spmv_host_csr(ctx, rp, ci, va, N, N, x, sigma=0.3)      # STSHIFT operator A - sigma I
    assert kern == K_PIPE
    assert np.abs(y - (A @ x - 0.3 * x)).max() <= 1e-13 * 5


# ---- small irregular matrices forced through the pipeline -------------------------------------------------------------------
def _irregular(kind, rng):
    import scipy.sparse as sp
    if kind == "wide_rows":                         # widths 9..40: beyond the width-specialised fast path (<= 8)
        n = 6000
        lens = rng.integers(9, 41, n)
    elif kind == "empty_rows":
        n = 7003                                    # odd slice tail: 7003 = 218 slices + 27 rows
        lens = rng.integers(0, 7, n)
        lens[rng.random(n) < 0.4] = 0
    elif kind == "empty_chunks":                    # 4096 leading rows and a 2048-row band without entries: whole chunks of empty slices
        n = 12000
        lens = rng.integers(1, 6, n)
        lens[:4096] = 0
        lens[8192:10240] = 0
    elif kind == "mixed_widths":                    # neighbouring slices of different widths, some pairs equal, some not
        n = 9000
        lens = np.repeat(rng.integers(1, 13, (n + 31) // 32), 32)[:n]
        lens[::7] = np.maximum(lens[::7] - 1, 0)
    elif kind == "one_slice_pair":
        n = 64
        lens = rng.integers(1, 5, n)
    else:
        raise ValueError(kind)
    ncols = n + 500
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(ncols, l, replace=False)) for l in lens]) if rp[-1] else np.zeros(0, np.int64)
    va = rng.standard_normal(int(rp[-1]))
    A = sp.csr_matrix((va, ci, rp), shape=(n, ncols))
    return A


@pytest.mark.parametrize("ghost", [False, True], ids=["noghost", "ghost"])
@pytest.mark.parametrize("kind", ["wide_rows", "empty_rows", "empty_chunks", "mixed_widths", "one_slice_pair"])
def test_pipe_forced_on_small_irregular(ctx, kind, ghost):
    rng = np.random.default_rng({"wide_rows": 11, "empty_rows": 12, "empty_chunks": 13, "mixed_widths": 14, "one_slice_pair": 15}[kind])
    A = _irregular(kind, rng)
    n, ncols = A.shape
    x = rng.standard_normal(ncols)
    ncl = n if ghost else ncols
    _check(ctx.lib.b2k_spmv_set_pipe_min_chunks(1))
    _check(ctx.lib.b2k_spmv_set_sell(2))            # SELL copy whatever the padding (the default keeps CSR-stream above 25 % padding)
    try:
        y, kern, _ = _spmv_host_csr(ctx, A.indptr, A.indices, A.data, n, ncl, x[:ncl], x[ncl:] if ghost else None)
        ys, kern_s, _ = _spmv_host_csr(ctx, A.indptr, A.indices, A.data, n, ncl, x[:ncl], x[ncl:] if ghost else None, sigma=-1.25)
    finally:
        _check(ctx.lib.b2k_spmv_set_pipe_min_chunks(-1))
        _check(ctx.lib.b2k_spmv_set_sell(1))
    ref = A @ x
    assert kern == (K_PIPE_GHOST if ghost else K_PIPE) and kern_s == kern
    tol = 1e-13 * np.abs(x).max() * max(int(np.diff(A.indptr).max()), 1) * max(np.abs(A.data).max(), 1.0)
    assert np.abs(y - ref).max() <= tol
    assert np.abs(ys - (ref + 1.25 * x[:n])).max() <= tol + 1e-15 * np.abs(x).max()


def test_small_matrix_keeps_the_small_kernel(ctx):
    """below the chunk threshold the plain SELL kernel runs (the pipeline needs >= 4 chunks per SM to fill its ring)"""
    A = O.laplacian_2d(64, 64).tocsr()
    x = np.random.default_rng(1).standard_normal(A.shape[1])
    y, kern, _ = _spmv_host_csr(ctx, A.indptr, A.indices, A.data, A.shape[0], A.shape[1], x)
    assert kern == K_SELL
    assert np.abs(y - A @ x).max() <= 1e-13 * 5 * 4


def test_csr_arrays_come_back_after_the_drop(ctx):
    """b2k_csr_arrays rebuilds the CSR (col,val) copy from the SELL copy bit for bit; b2k_csr_release_arrays drops it again"""
    A = O.markov_model(120).tocsr()
    A.sort_indices()
    n = A.shape[0]
    h = ctypes.c_void_p()
    rp = A.indptr.astype(np.int32); ci = A.indices.astype(np.int32); va = A.data.astype(np.float64)
    _check(ctx.lib.b2k_csr_create(ctx.h, n, n, 0, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, ctypes.byref(h)))
    b0, b1 = ctypes.c_int64(), ctypes.c_int64()
    _check(ctx.lib.b2k_csr_bytes(h, ctypes.byref(b0)))
    prp, pci, pva = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    _check(ctx.lib.b2k_csr_arrays(h, ctypes.byref(prp), ctypes.byref(pci), ctypes.byref(pva)))
    _check(ctx.lib.b2k_csr_bytes(h, ctypes.byref(b1)))
    assert b1.value == b0.value + 12 * A.nnz
    gci, gva = np.empty(A.nnz, np.int32), np.empty(A.nnz)
    _check(ctx.lib.b2k_d2h(ctx.h, gci.ctypes.data, pci, gci.nbytes))
    _check(ctx.lib.b2k_d2h(ctx.h, gva.ctypes.data, pva, gva.nbytes))
    assert np.array_equal(gci, ci) and np.array_equal(gva, va)
    _check(ctx.lib.b2k_csr_release_arrays(h))
    _check(ctx.lib.b2k_csr_bytes(h, ctypes.byref(b1)))
    assert b1.value == b0.value
    _check(ctx.lib.b2k_csr_destroy(ctx.h, h))


@pytest.mark.parametrize("k", [1, 5, 8, 16, 23, 40])
@pytest.mark.parametrize("kind", ["lap3", "random_ghost", "empty_rows"])
def test_spmm_block_of_vectors(ctx, kind, k):
    """b2k_csr_spmm (k_spmm_sell: Y = A X for k vectors, the matrix read once per 16 columns — BVMatMult in BV_MATMULT_MAT mode,
    svec.c:203-231) against scipy, leading dimensions larger than the row counts, ghost block included"""
    import scipy.sparse as sp
    rng = np.random.default_rng(17 + k)
    if kind == "lap3":
        A = O.laplacian_3d(23, 17, 19).tocsr()
        ncl = A.shape[1]
    elif kind == "random_ghost":
        A = sp.random(30011, 26003, density=12 / 26003, random_state=5, format="csr")
        ncl = 20000
    else:
        A = _irregular("empty_rows", rng)
        ncl = A.shape[1]
    A.sort_indices()
    n, nc = A.shape
    ng = nc - ncl
    ldx, ldg, ldy = ncl + 2, ng + 4, n + 6
    X = np.zeros((ldx, k), order="F"); X[:ncl] = rng.standard_normal((ncl, k))
    G = np.zeros((max(ldg, 1), k), order="F"); G[:ng] = rng.standard_normal((ng, k))
    h = ctypes.c_void_p()
    rp = A.indptr.astype(np.int32); ci = A.indices.astype(np.int32); va = A.data.astype(np.float64)
    _check(ctx.lib.b2k_spmv_set_sell(2))
    try:
        _check(ctx.lib.b2k_csr_create(ctx.h, n, ncl, ng, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, ctypes.byref(h)))
    finally:
        _check(ctx.lib.b2k_spmv_set_sell(1))
    dX, dG = ctx.to_device(X), ctx.to_device(G)
    dY = ctx.to_device(np.full((ldy, k), 7.0, order="F"))
    _check(ctx.lib.b2k_csr_spmm(ctx.h, h, dX.ptr, ldx, dG.ptr if ng else None, ldg, dY.ptr, ldy, k))
    ctx.sync()
    assert _last_kernel(ctx, h) == 5                                          # small matrix: k_spmm_sell
    Y = dY.to_host((ldy, k))
    ref = A @ np.vstack([X[:ncl], G[:ng]])
    tol = 1e-13 * max(int(np.diff(A.indptr).max()), 1) * max(np.abs(A.data).max(), 1.0) * 6
    assert np.abs(Y[:n] - ref).max() <= tol
    assert np.array_equal(Y[n:], np.full((ldy - n, k), 7.0))                 # rows past n untouched
    _check(ctx.lib.b2k_csr_destroy(ctx.h, h))
    for d in (dX, dG, dY):
        d.free()


@pytest.mark.parametrize("ghost", [False, True], ids=["owned_columns", "ghost_columns"])
@pytest.mark.parametrize("k", [3, 16])
def test_spmm_through_the_pipeline_1m_rows(ctx, k, ghost):
    """large matrices: b2k_csr_spmm runs the bulk-copy pipeline with the column loop inside the consumers (k_spmv_sell_pipe<.,true>):
    the matrix is read once for the k vectors; element-wise against scipy on the 1 M-row random matrix"""
    import scipy.sparse as sp
    from slepc_b200 import matgen
    M, N = 1 << 20, 900001
    rp, ci, va = matgen.random_sparse_rows(M, N, 20, seed=5)
    A = sp.csr_matrix((va, ci, rp), shape=(M, N))
    rng = np.random.default_rng(9)
    ncl = 600000 if ghost else N
    ng = N - ncl
    ldx, ldg, ldy = ncl + 2, ng + 2, M + 4
    X = np.zeros((ldx, k), order="F"); X[:ncl] = rng.uniform(-1, 1, (ncl, k))
    G = np.zeros((ldg, k), order="F"); G[:ng] = rng.uniform(-1, 1, (ng, k))
    h = ctypes.c_void_p()
    rp32, ci32 = np.ascontiguousarray(rp, dtype=np.int32), np.ascontiguousarray(ci, dtype=np.int32)
    _check(ctx.lib.b2k_csr_create(ctx.h, M, ncl, ng, rp32.ctypes.data, ci32.ctypes.data, va.ctypes.data, ctypes.byref(h)))
    dX, dG, dY = ctx.to_device(X), ctx.to_device(G), ctx.empty(ldy * k)
    _check(ctx.lib.b2k_csr_spmm(ctx.h, h, dX.ptr, ldx, dG.ptr if ng else None, ldg, dY.ptr, ldy, k))
    ctx.sync()
    assert _last_kernel(ctx, h) == 6
    Y = dY.to_host((ldy, k))
    ref = A @ np.vstack([X[:ncl], G[:ng]])
    assert np.abs(Y[:M] - ref).max() <= 1e-13 * int(np.diff(rp).max()) * np.abs(va).max()
    _check(ctx.lib.b2k_csr_destroy(ctx.h, h))
    for d in (dX, dG, dY):
        d.free()


# ---- local transpose built in HBM (b2k_csr_transpose_split): what MatMultTranspose / the explicit transpose of SVDSetUp multiply with ----
def _csr_to_host(ctx, h):
    nr, ncl, ng, nnz = (ctypes.c_int64() for _ in range(4))
    _check(ctx.lib.b2k_csr_info(h, ctypes.byref(nr), ctypes.byref(ncl), ctypes.byref(ng), ctypes.byref(nnz)))
    prp, pci, pv = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    _check(ctx.lib.b2k_csr_arrays(h, ctypes.byref(prp), ctypes.byref(pci), ctypes.byref(pv)))
    rp = np.empty(nr.value + 1, dtype=np.int32); ci = np.empty(max(nnz.value, 1), dtype=np.int32); va = np.empty(max(nnz.value, 1))
    _check(ctx.lib.b2k_d2h(ctx.h, rp.ctypes.data, prp, rp.nbytes))
    if nnz.value:
        _check(ctx.lib.b2k_d2h(ctx.h, ci.ctypes.data, pci, 4 * nnz.value))
        _check(ctx.lib.b2k_d2h(ctx.h, va.ctypes.data, pv, 8 * nnz.value))
    _check(ctx.lib.b2k_csr_release_arrays(h))
    return rp, ci[:nnz.value], va[:nnz.value], (nr.value, ncl.value)


@pytest.mark.parametrize("case", ["random_ghost", "random_noghost", "irregular_empty", "markov", "no_entries"])
def test_device_transpose_matches_scipy_exactly(ctx, case):
    """AT_own / AT_ghost from the device radix sort + gather against scipy's transpose: identical row pointers, identical column
    order inside every transposed row (ascending row of A: the summation order of y = A^T x is fixed), identical values; then
    y = A^T x through the product kernels"""
    import scipy.sparse as sp
    from slepc_b200 import matgen
    rng = np.random.default_rng(21)
    if case in ("random_ghost", "random_noghost"):
        M, N = 300007, 90001
        rp, ci, va = matgen.random_sparse_rows(M, N, 12, seed=9)
        A = sp.csr_matrix((va, ci, rp), shape=(M, N))
        ncl = 50000 if case == "random_ghost" else N
    elif case == "irregular_empty":
        A = _irregular("empty_chunks", rng).tocsr()
        ncl = A.shape[0]                              # columns beyond n are ghosts
    elif case == "markov":
        m = 300
        N = matgen.markov_size(m)
        rp, ci, va = matgen.markov_rows(m)
        A = sp.csr_matrix((va, ci, rp), shape=(N, N))
        ncl = N
    else:
        A = sp.csr_matrix((40, 70))
        ncl = 30
    A.sort_indices()
    M, N = A.shape
    ng = N - ncl
    h, hown, hgh = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32); va = np.ascontiguousarray(A.data)
    _check(ctx.lib.b2k_csr_create(ctx.h, M, ncl, ng, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, ctypes.byref(h)))
    _check(ctx.lib.b2k_csr_transpose_split(ctx.h, h, ctypes.byref(hown), ctypes.byref(hgh) if ng else None))
    AT = A.T.tocsr()
    AT.sort_indices()
    for hh, blk in ((hown, AT[:ncl]), (hgh, AT[ncl:])):
        if not hh:
            assert blk.shape[0] == 0
            continue
        trp, tci, tva, shape = _csr_to_host(ctx, hh)
        assert shape == (blk.shape[0], M)
        assert np.array_equal(trp, blk.indptr) and np.array_equal(tci, blk.indices) and np.array_equal(tva, blk.data)
    # y = A^T x through the transposed blocks (whatever kernel the dispatch picks)
    x = rng.standard_normal(M)
    dx = ctx.to_device(x)
    ref = AT @ x
    for hh, lo, hi in ((hown, 0, ncl), (hgh, ncl, N)):
        if not hh or hi == lo:
            continue
        dy = ctx.empty(hi - lo)
        _check(ctx.lib.b2k_csr_spmv(ctx.h, hh, dx.ptr, None, dy.ptr))
        ctx.sync()
        tol = 1e-13 * max(np.abs(x).max(), 1.0) * max(int(np.diff(AT.indptr).max()), 1) * max(np.abs(A.data).max() if A.nnz else 1.0, 1.0)
        assert np.abs(dy.to_host() - ref[lo:hi]).max() <= tol
        dy.free()
    dx.free()
    for hh in (hown, hgh, h):
        if hh:
            _check(ctx.lib.b2k_csr_destroy(ctx.h, hh))


# ---- local column numbering built in HBM (b2k_csr_create_global): what MatCreateB200CSR does with a row block on > 1 rank ----
@pytest.mark.parametrize("case", ["middle_block", "first_block", "all_owned", "no_offrange_entries", "empty"])
def test_global_columns_localized_on_the_device(ctx, case):
    """rows with GLOBAL column indices: the ghost list (sorted distinct off-range columns, MPIAIJ's garray) and the local
    numbering [owned | ghosts] come from the device select / sort / unique / lookup; checked against numpy and through a product"""
    import scipy.sparse as sp
    from slepc_b200 import matgen
    rng = np.random.default_rng(33)
    M, N = 200003, 70001
    if case == "empty":
        A = sp.csr_matrix((50, N))
        c0, c1 = 100, 900
    else:
        rp, ci, va = matgen.random_sparse_rows(M, N, 9, seed=4)
        A = sp.csr_matrix((va, ci, rp), shape=(M, N))
        c0, c1 = {"middle_block": (20000, 45000), "first_block": (0, 30000), "all_owned": (0, N), "no_offrange_entries": (0, N)}[case]
        if case == "no_offrange_entries":
            A = A[:, :40000].tocsr()                 # entries only in [0, 40000), owned range declared as [0, 50000) of 70001
            A = sp.csr_matrix((A.data, A.indices, A.indptr), shape=(M, N))
            c0, c1 = 0, 50000
    A.sort_indices()
    nr = A.shape[0]
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32); ci = np.ascontiguousarray(A.indices, dtype=np.int32); va = np.ascontiguousarray(A.data)
    h, ng, gp = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_void_p()
    _check(ctx.lib.b2k_csr_create_global(ctx.h, nr, N, c0, c1, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, ctypes.byref(h), ctypes.byref(ng), ctypes.byref(gp)))
    cols = np.unique(A.indices)
    ghosts_ref = cols[(cols < c0) | (cols >= c1)]
    assert ng.value == len(ghosts_ref)
    ghosts = np.ctypeslib.as_array(ctypes.cast(gp, ctypes.POINTER(ctypes.c_int)), shape=(ng.value,)).copy() if ng.value else np.zeros(0, np.int32)
    assert np.array_equal(ghosts, ghosts_ref)
    x = rng.standard_normal(N)
    dx = ctx.to_device(np.ascontiguousarray(x[c0:c1]))
    dg = ctx.to_device(np.ascontiguousarray(x[ghosts])) if ng.value else None
    dy = ctx.empty(max(nr, 1))
    _check(ctx.lib.b2k_csr_spmv(ctx.h, h, dx.ptr, dg.ptr if dg else None, dy.ptr))
    ctx.sync()
    y = dy.to_host()[:nr]
    tol = 1e-13 * np.abs(x).max() * max(int(np.diff(A.indptr).max()), 1) * max(np.abs(A.data).max() if A.nnz else 1.0, 1.0)
    assert np.abs(y - A @ x).max() <= tol
    # the local numbering itself: owned columns first, then the ghosts in ascending global order
    lrp, lci, lva, shape = _csr_to_host(ctx, h)
    loc2glob = np.concatenate([np.arange(c0, c1), ghosts_ref])
    assert np.array_equal(lrp, A.indptr) and np.array_equal(loc2glob[lci], A.indices) and np.array_equal(lva, A.data)
    for d in (dx, dg, dy):
        if d is not None:
            d.free()
    _check(ctx.lib.b2k_csr_destroy(ctx.h, h))
    if gp:
        libc = ctypes.CDLL(None)
        libc.free.argtypes = [ctypes.c_void_p]
        libc.free(gp)


def test_global_columns_out_of_range_is_an_error(ctx):
    from slepc_b200._b2k import B2KError
    rp = np.array([0, 2, 3], dtype=np.int32); ci = np.array([0, 7, 12], dtype=np.int32); va = np.ones(3)
    h, ng, gp = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_void_p()
    with pytest.raises(B2KError):
        _check(ctx.lib.b2k_csr_create_global(ctx.h, 2, 10, 2, 6, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, ctypes.byref(h), ctypes.byref(ng), ctypes.byref(gp)))
