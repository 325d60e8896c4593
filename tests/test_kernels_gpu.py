"""Parity of the sm_100a BV / SpMV kernels (through the C ABI, include/b2k.h) against numpy
restatements of the reference BLAS shapes (bvblas.c:24-278) on the same seeded inputs.
Tolerances: FP64, reduction order differs from BLAS → relative 1e-13 * sqrt(n)-ish bounds below."""
import ctypes

import numpy as np
import pytest

from oracle import slepc_oracle as O

pytestmark = pytest.mark.gpu


def _mk(ctx, n, m, ld, seed, misalign=0):
    rng = np.random.default_rng(seed)
    H = np.zeros((ld, m), order="F")
    H[:n, :] = rng.standard_normal((n, m))
    if misalign:
        buf = ctx.empty(ld * m + misalign)
        ctx.lib.b2k_h2d(ctx.h, buf.at(misalign), H.ctypes.data, H.nbytes)
        return H, buf, buf.at(misalign)
    d = ctx.to_device(H)
    return H, d, d.ptr


def _tol(n, k=1):
    return 64 * np.finfo(float).eps * max(1.0, np.sqrt(n)) * max(1, k) ** 0.5


SHAPES = [(1, 1), (2, 3), (7, 5), (1000, 16), (1001, 17), (100003, 33), (1 << 18, 64), (5, 0)]


@pytest.mark.parametrize("n,k", SHAPES)
@pytest.mark.parametrize("misalign", [0, 1])
def test_dotvec_and_gs_dot(ctx, n, k, misalign):
    from slepc_b200._b2k import check
    ld = n + (n % 2) + 6
    H, dV, pV = _mk(ctx, n, max(k, 1), ld, 1, misalign)
    rng = np.random.default_rng(2)
    w = rng.standard_normal(n)
    dw = ctx.to_device(w)
    out = ctx.empty(k + 1)
    check(ctx.lib.b2k_memset0(ctx.h, out.ptr, 8 * (k + 1)))
    check(ctx.lib.b2k_gs_dot(ctx.h, pV, ld, n, k, dw.ptr, out.ptr))
    got = out.to_host()
    ref = np.concatenate([H[:n, :k].T @ w, [w @ w]])
    scale = np.linalg.norm(w) * max(1.0, np.abs(H[:n, :k]).max() if k else 1.0)
    assert np.allclose(got, ref, rtol=0, atol=_tol(n) * scale * np.sqrt(n))
    if k:
        out2 = ctx.empty(k)
        check(ctx.lib.b2k_dotvec(ctx.h, pV, ld, n, k, dw.ptr, out2.ptr))
        assert np.array_equal(out2.to_host(), got[:k])          # same sweep, same order → bitwise
        # determinism run to run
        check(ctx.lib.b2k_dotvec(ctx.h, pV, ld, n, k, dw.ptr, out2.ptr))
        assert np.array_equal(out2.to_host(), got[:k])


@pytest.mark.parametrize("n,k", SHAPES)
@pytest.mark.parametrize("alpha,beta", [(-1.0, 1.0), (2.0, 0.0), (0.5, -0.25)])
def test_multvec(ctx, n, k, alpha, beta):
    from slepc_b200._b2k import check
    ld = n + (n % 2)
    H, dV, pV = _mk(ctx, n, max(k, 1), ld, 3)
    rng = np.random.default_rng(4)
    y = rng.standard_normal(n)
    if beta == 0.0:
        y[::3] = np.nan                                        # BLAS semantics: beta==0 ⇒ y not read
    q = rng.standard_normal(max(k, 1))
    dy, dq = ctx.to_device(y), ctx.to_device(q)
    check(ctx.lib.b2k_multvec(ctx.h, pV, ld, n, k, alpha, beta, dy.ptr, dq.ptr))
    got = dy.to_host()
    ref = alpha * (H[:n, :k] @ q[:k]) + (0.0 if beta == 0.0 else beta * y)
    assert np.allclose(got, ref, rtol=1e-13, atol=1e-12 * max(1, k))


@pytest.mark.parametrize("n,k", [(1000, 5), (100003, 33), (1 << 18, 64), (77, 1)])
def test_gs_update_dot_matches_cgs_pass(ctx, n, k):
    """One fused sweep == BVMultColumn(-1,1,j,c) followed by BVDotColumnInc (bvorthog.c:100,113)."""
    from slepc_b200._b2k import check
    ld = n + (n % 2)
    rng = np.random.default_rng(5)
    Q, _ = np.linalg.qr(rng.standard_normal((n, k)))
    H = np.zeros((ld, k), order="F")
    H[:n] = Q
    w = rng.standard_normal(n)
    dV, dw = ctx.to_device(H), ctx.to_device(w)
    c1, c2 = ctx.empty(k + 1), ctx.empty(k + 1)
    check(ctx.lib.b2k_gs_dot(ctx.h, dV.ptr, ld, n, k, dw.ptr, c1.ptr))
    check(ctx.lib.b2k_gs_update_dot(ctx.h, dV.ptr, ld, n, k, dw.ptr, c1.ptr, c2.ptr))
    c1h, c2h, wn = c1.to_host(), c2.to_host(), dw.to_host()
    cref = Q.T @ w
    wref = w - Q @ cref
    assert np.allclose(c1h[:k], cref, atol=1e-12 * np.sqrt(n))
    assert np.allclose(wn, wref, atol=1e-12 * np.sqrt(n))
    assert np.isclose(c2h[k], wref @ wref, rtol=1e-12)
    assert np.allclose(c2h[:k], Q.T @ wn, atol=1e-12 * np.sqrt(n))
    assert np.abs(c2h[:k]).max() < 1e-12 * np.sqrt(n)           # orthogonal after one pass here


@pytest.mark.parametrize("k", [1, 5, 8, 9, 16, 17, 31, 33, 48, 64])
@pytest.mark.parametrize("n", [1, 63, 128, 130, 128 * 7 + 1, 4096, 4097, 12345, 128 * 148 * 3 + 77, 1 << 20])
def test_gs_fused_single_sweep_equals_two_sweep(ctx, n, k):
    """The single-sweep kernels (1: register tile, 3: 2-D tensor-map TMA pipeline; 3 falls back to 1
    below 4096 rows) and the two-sweep path (0) compute the same w
    and c (different reduction order ⇒ compare to rounding, and each against the float64 reference)."""
    from slepc_b200._b2k import check
    ld = n + (n % 2) + 2
    rng = np.random.default_rng(100 + k)
    H = np.zeros((ld, k), order="F")
    H[:n] = rng.standard_normal((n, k)) / np.sqrt(n)
    w = rng.standard_normal(n)
    cin = rng.standard_normal(k + 1)
    dV, dc = ctx.to_device(H), ctx.to_device(cin)
    out = {}
    for mode in (0, 1, 3):
        check(ctx.lib.b2k_gs_set_fused(mode))
        dw = ctx.to_device(w)
        co = ctx.empty(k + 1)
        check(ctx.lib.b2k_gs_update_dot(ctx.h, dV.ptr, ld, n, k, dw.ptr, dc.ptr, co.ptr))
        out[mode] = (dw.to_host(), co.to_host())
        if mode >= 1:   # determinism of the single-sweep kernels
            dw2 = ctx.to_device(w)
            co2 = ctx.empty(k + 1)
            check(ctx.lib.b2k_gs_update_dot(ctx.h, dV.ptr, ld, n, k, dw2.ptr, dc.ptr, co2.ptr))
            assert np.array_equal(dw2.to_host(), out[mode][0]) and np.array_equal(co2.to_host(), out[mode][1])
    check(ctx.lib.b2k_gs_set_fused(3))
    wref = w - H[:n] @ cin[:k]
    cref = np.concatenate([H[:n].T @ wref, [wref @ wref]])
    for mode in (0, 1, 3):
        assert np.allclose(out[mode][0], wref, rtol=1e-13, atol=1e-13 * np.sqrt(k)), mode
        assert np.allclose(out[mode][1], cref, rtol=1e-12, atol=1e-12 * np.sqrt(n)), mode


@pytest.mark.parametrize("alpha,beta", [(-1.0, 1.0), (2.0, 0.0), (0.5, -0.25)])
@pytest.mark.parametrize("n,k", [(4096, 1), (4097, 7), (50001, 16), (50001, 17), (1 << 18, 40), (300007, 64)])
def test_multvec_and_update_norm_tensor_map_pipeline(ctx, n, k, alpha, beta):
    """b2k_multvec / b2k_gs_update_norm through the 2-D tensor-map kernel (mode 3): general alpha/beta, beta==0 must not
    read y (BLAS semantics), odd n (zero-filled tail rows), k below the box width (zero-filled columns)."""
    from slepc_b200._b2k import check
    ld = n + (n % 2)
    rng = np.random.default_rng(n + k)
    H = np.zeros((ld, k), order="F")
    H[:n] = rng.standard_normal((n, k))
    y = rng.standard_normal(n)
    q = rng.standard_normal(k)
    dV, dq = ctx.to_device(H), ctx.to_device(q)
    check(ctx.lib.b2k_gs_set_fused(3))
    try:
        yy = y.copy()
        if beta == 0.0:
            yy[::3] = np.nan
        dy = ctx.to_device(yy)
        check(ctx.lib.b2k_multvec(ctx.h, dV.ptr, ld, n, k, alpha, beta, dy.ptr, dq.ptr))
        ref = alpha * (H[:n] @ q) + (0.0 if beta == 0.0 else beta * y)
        assert np.allclose(dy.to_host(), ref, rtol=1e-13, atol=1e-12 * k)
        dw, nr = ctx.to_device(y), ctx.empty(1)
        check(ctx.lib.b2k_gs_update_norm(ctx.h, dV.ptr, ld, n, k, dw.ptr, dq.ptr, nr.ptr))
        wref = y - H[:n] @ q
        assert np.allclose(dw.to_host(), wref, rtol=1e-13, atol=1e-12 * k)
        assert np.isclose(nr.to_host()[0], wref @ wref, rtol=1e-12)
    finally:
        check(ctx.lib.b2k_gs_set_fused(3))


def test_level1(ctx):
    from slepc_b200._b2k import check
    n, k, ld = 100003, 5, 100004
    H, dX, pX = _mk(ctx, n, k, ld, 7)
    out = ctx.empty(1)
    check(ctx.lib.b2k_sumsq(ctx.h, pX, ld, n, k, out.ptr))
    assert np.isclose(out.to_host()[0], np.sum(H[:n] ** 2), rtol=1e-13)
    check(ctx.lib.b2k_sumsq(ctx.h, pX + 8 * ld, ld, n, 1, out.ptr))
    assert np.isclose(out.to_host()[0], np.sum(H[:n, 1] ** 2), rtol=1e-13)
    outk = ctx.empty(k)
    check(ctx.lib.b2k_colabssum(ctx.h, pX, ld, n, k, outk.ptr))
    assert np.allclose(outk.to_host(), np.abs(H[:n]).sum(0), rtol=1e-13)
    omax = ctx.empty(1)
    check(ctx.lib.b2k_rowabssum_max(ctx.h, pX, ld, n, k, omax.ptr))                      # NORM_INFINITY of the block
    assert np.isclose(omax.to_host()[0], np.abs(H[:n]).sum(1).max(), rtol=1e-14)
    check(ctx.lib.b2k_rowabssum_max(ctx.h, pX + 8 * 2 * ld, ld, 77, 1, omax.ptr))        # one column, short
    assert omax.to_host()[0] == np.abs(H[:77, 2]).max()
    check(ctx.lib.b2k_scale(ctx.h, pX, ld, n, k, -2.5))
    G = dX.to_host((ld, k))
    assert np.array_equal(G[:n], -2.5 * H[:n]) and np.array_equal(G[n:], H[n:])
    dY = ctx.empty(ld * k)
    check(ctx.lib.b2k_memset0(ctx.h, dY.ptr, 8 * ld * k))
    check(ctx.lib.b2k_copy(ctx.h, dY.ptr, ld, pX, ld, n, k))
    assert np.array_equal(dY.to_host((ld, k))[:n], G[:n])
    check(ctx.lib.b2k_axpby(ctx.h, dY.ptr, ld, pX, ld, n, k, 0.5, 2.0))
    assert np.allclose(dY.to_host((ld, k))[:n], 2.5 * G[:n], rtol=1e-15)
    check(ctx.lib.b2k_scale(ctx.h, pX, ld, n, 2, 0.0))
    G2 = dX.to_host((ld, k))
    assert not G2[:n, :2].any() and np.array_equal(G2[:, 2:], G[:, 2:])
    x = ctx.empty(n)
    check(ctx.lib.b2k_fill(ctx.h, x.ptr, n, 3.25))
    assert np.all(x.to_host() == 3.25)
    check(ctx.lib.b2k_set_random(ctx.h, x.ptr, n, 1000, 0x5EED + 3))
    assert np.array_equal(x.to_host(), O.hash_uniform(np.arange(1000, 1000 + n), 0x5EED + 3))   # bit exact
    ss = ctx.to_device(np.array([4.0]))
    check(ctx.lib.b2k_fill(ctx.h, x.ptr, n, 3.0))
    check(ctx.lib.b2k_scale_rsqrt(ctx.h, x.ptr, n, ss.ptr))
    assert np.all(x.to_host() == 1.5)


@pytest.mark.parametrize("n,kin,nout", [(1000, 5, 3), (4097, 64, 32), (100003, 33, 70), (50, 200, 7), (333, 700, 5),
                                        (4096, 64, 32), (100002, 48, 20), (130, 64, 64), (2, 3, 2), (33000, 17, 47)])
def test_mult(ctx, n, kin, nout):
    from slepc_b200._b2k import check
    ldx, ldy, ldq = n + 2, n + 4, kin + 3
    X, dX, pX = _mk(ctx, n, kin, ldx, 11)
    Y, dY, pY = _mk(ctx, n, nout, ldy, 12)
    rng = np.random.default_rng(13)
    Q = np.zeros((ldq, nout), order="F")
    Q[:kin] = rng.standard_normal((kin, nout))
    dQ = ctx.to_device(Q)
    check(ctx.lib.b2k_mult(ctx.h, pY, ldy, pX, ldx, n, nout, kin, 2.0, 0.5, dQ.ptr, ldq))
    got = dY.to_host((ldy, nout))
    ref = 0.5 * Y[:n] + 2.0 * (X[:n] @ Q[:kin])
    assert np.allclose(got[:n], ref, rtol=1e-12, atol=1e-11 * np.sqrt(kin))
    assert np.array_equal(got[n:], Y[n:])


@pytest.mark.parametrize("tma", [1, 0])                # TMA ring + FP64 tensor cores (k_vq_tma) / DFMA kernel (k_vq)
@pytest.mark.parametrize("trans", [0, 1])
@pytest.mark.parametrize("n,k,s,e", [(10, 5, 1, 3), (4099, 64, 0, 40), (100003, 48, 5, 48), (1000, 65, 0, 65),
                                     (4098, 64, 0, 40), (100002, 48, 5, 48), (126, 17, 3, 17), (70000, 64, 0, 64), (1000, 64, 20, 21)])
def test_mult_inplace(ctx, n, k, s, e, trans, tma):
    """BVMultInPlace semantics (bvblas.c:74-106): V(:,s:e) = V(:,0:k) Q(0:k,s:e), other columns untouched."""
    from slepc_b200._b2k import check
    check(ctx.lib.b2k_vq_set_tma(tma))
    ld, ldq = n + 2, k + 1
    V, dV, pV = _mk(ctx, n, k + 2, ld, 21)
    rng = np.random.default_rng(22)
    Q = np.zeros((ldq, ldq), order="F")
    Q[:k, :k] = rng.standard_normal((k, k))
    dQ = ctx.to_device(Q)
    check(ctx.lib.b2k_mult_inplace(ctx.h, pV, ld, n, k, s, e, dQ.ptr, ldq, trans))
    got = dV.to_host((ld, k + 2))
    Qe = Q.T if trans else Q
    ref = V.copy()
    ref[:n, s:e] = V[:n, :k] @ Qe[:k, s:e]
    assert np.allclose(got[:n], ref[:n], rtol=1e-12, atol=1e-11 * np.sqrt(k))
    assert np.array_equal(got[:, e:], V[:, e:]) and np.array_equal(got[:, :s], V[:, :s])
    check(ctx.lib.b2k_vq_set_tma(1))


def test_dot(ctx):
    from slepc_b200._b2k import check
    n, kx, ky = 50001, 6, 9
    X, dX, pX = _mk(ctx, n, kx, n + 1, 31)
    Y, dY, pY = _mk(ctx, n, ky, n + 3, 32)
    M = ctx.empty((ky + 2) * kx)
    check(ctx.lib.b2k_memset0(ctx.h, M.ptr, 8 * (ky + 2) * kx))
    check(ctx.lib.b2k_dot(ctx.h, pY, n + 3, pX, n + 1, n, ky, kx, M.ptr, ky + 2))
    got = M.to_host((ky + 2, kx))
    assert np.allclose(got[:ky], Y[:n].T @ X[:n], rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("n,ky,kx,same", [(4099, 5, 7, False), (50002, 64, 64, False), (50002, 64, 64, True), (131072, 33, 17, False),
                                          (70001, 24, 24, True), (20000, 40, 9, True), (30000, 100, 70, False), (30000, 130, 130, True),
                                          (1 << 20, 32, 32, True), (3000, 8, 8, True)])
def test_dot_one_sweep_tensor_core_gram(ctx, n, ky, kx, same):
    """b2k_dot = k_gram_tma (one sweep of Y and X, DMMA) against numpy and against the column-by-column sweeps it replaces
    (B2K_GRAM_TMA=0 is a process-wide switch, so the reference here is b2k_dotvec per column): odd n (zero-filled tail rows),
    column counts that are not multiples of 8, more than 64 columns (64 x 64 blocks through the scratch area), X == Y (the Gram
    matrix of BVOrthogonalize CHOL / SVQB, read once), leading dimension of M larger than ky."""
    from slepc_b200._b2k import check
    ld = n + (n % 2) + 2
    Y, dY, pY = _mk(ctx, n, ky, ld, 51)
    if same:
        X, pX, kx = Y, pY, min(kx, ky)
    else:
        X, dX, pX = _mk(ctx, n, kx, ld, 52)
    for ldm in (ky, ky + 3):
        M = ctx.empty(ldm * kx)
        check(ctx.lib.b2k_memset0(ctx.h, M.ptr, 8 * ldm * kx))
        check(ctx.lib.b2k_dot(ctx.h, pY, ld, pX, ld, n, ky, kx, M.ptr, ldm))
        got = M.to_host((ldm, kx))
        ref = Y[:n, :ky].T @ X[:n, :kx]
        assert np.allclose(got[:ky], ref, rtol=1e-12, atol=1e-11 * np.sqrt(n))
        assert np.array_equal(got[ky:], np.zeros((ldm - ky, kx)))
        M2 = ctx.empty(ldm * kx)
        check(ctx.lib.b2k_dot(ctx.h, pY, ld, pX, ld, n, ky, kx, M2.ptr, ldm))          # bit-reproducible run to run
        assert np.array_equal(M2.to_host((ldm, kx))[:ky], got[:ky])
        col = ctx.empty(ky)
        for j in (0, kx - 1):
            check(ctx.lib.b2k_dotvec(ctx.h, pY, ld, n, ky, pX + 8 * j * ld, col.ptr))
            assert np.allclose(col.to_host(), got[:ky, j], rtol=1e-12, atol=1e-11 * np.sqrt(n))
        for a in (M, M2, col):
            a.free()


def _spmv(ctx, A, x, xg=None, sell=1, sigma=None):
    from slepc_b200._b2k import check
    check(ctx.lib.b2k_spmv_set_sell(sell))
    A = A.tocsr()
    A.sort_indices()
    nloc = A.shape[1] if xg is None else len(x)
    h = ctypes.c_void_p()
    rp = A.indptr.astype(np.int32)
    ci = A.indices.astype(np.int32)
    va = A.data.astype(np.float64)
    check(ctx.lib.b2k_csr_create(ctx.h, A.shape[0], nloc, 0 if xg is None else len(xg), rp.ctypes.data, ci.ctypes.data,
                                 va.ctypes.data, ctypes.byref(h)))
    dx = ctx.to_device(x)
    dg = ctx.to_device(xg) if xg is not None else None
    dy = ctx.empty(A.shape[0])
    if sigma is None:
        check(ctx.lib.b2k_csr_spmv(ctx.h, h, dx.ptr, dg.ptr if dg else None, dy.ptr))
    else:
        check(ctx.lib.b2k_csr_spmv_shift(ctx.h, h, dx.ptr, dg.ptr if dg else None, dy.ptr, sigma))
    y = dy.to_host()
    check(ctx.lib.b2k_csr_destroy(ctx.h, h))
    check(ctx.lib.b2k_spmv_set_sell(1))
    return y


@pytest.mark.parametrize("sell", [0, 1, 2])          # CSR-stream kernel / SELL-32 when padding is small / SELL-32 always
@pytest.mark.parametrize("gen", ["lap1", "lap2", "lap3", "markov", "random", "longrow", "empty_rows"])
def test_spmv_csr(ctx, gen, sell):
    import scipy.sparse as sp
    rng = np.random.default_rng(41)
    if gen == "lap1":
        A = O.laplacian_1d(100003)
    elif gen == "lap2":
        A = O.laplacian_2d(301, 257)
    elif gen == "lap3":
        A = O.laplacian_3d(37, 41, 29)
    elif gen == "markov":
        A = O.markov_model(150)
    elif gen == "random":
        A = sp.random(20011, 15013, density=20 / 15013, random_state=7, format="csr")
    elif gen == "longrow":
        A = sp.random(300, 9000, density=0.01, random_state=8, format="lil")
        A[17, :] = rng.standard_normal(9000)            # 9000 nnz in one row > shared-memory capacity
        A = A.tocsr()
    else:
        A = sp.random(5000, 5000, density=0.0005, random_state=9, format="csr")   # many empty rows
    x = rng.standard_normal(A.shape[1])
    y = _spmv(ctx, A, x, sell=sell)
    ref = A @ x
    assert np.allclose(y, ref, rtol=1e-13, atol=1e-12)
    if A.shape[0] == A.shape[1]:                            # shifted operator of STSHIFT (shift.c:79)
        y = _spmv(ctx, A, x, sell=sell, sigma=0.75)
        assert np.allclose(y, ref - 0.75 * x, rtol=1e-13, atol=1e-12)


def test_spmv_ghost_columns(ctx):
    A = O.laplacian_2d(64, 50).tocsr()
    nloc = 32 * 50
    B = A[:nloc, :]                                         # rows of rank 0; columns >= nloc are ghosts
    rng = np.random.default_rng(43)
    x = rng.standard_normal(A.shape[1])
    used = B[:, :nloc + 50]
    for sell in (0, 2):
        y = _spmv(ctx, used, x[:nloc], x[nloc:nloc + 50], sell=sell)
        assert np.allclose(y, (A @ x)[:nloc], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("dim,dims", [(1, (1000, 1, 1)), (2, (40, 33, 1)), (3, (12, 9, 7))])
@pytest.mark.parametrize("part", [(0, 1), (1, 4), (3, 4)])
def test_laplacian_device_generator(ctx, dim, dims, part):
    """b2k_csr_laplacian == the reference stencils (ex1.c:37-48, ex2.c:39-54) row for row."""
    from slepc_b200._b2k import check
    nx, ny, nz = dims
    if nx % part[1]:
        nx = (nx // part[1]) * part[1]
    A = {1: lambda: O.laplacian_1d(nx), 2: lambda: O.laplacian_2d(nx, ny), 3: lambda: O.laplacian_3d(nx, ny, nz)}[dim]().tocsr()
    plane = ny * nz
    rows_per = (nx // part[1]) * plane
    row0 = part[0] * rows_per
    h = ctypes.c_void_p()
    glo, ghi = ctypes.c_int64(), ctypes.c_int64()
    check(ctx.lib.b2k_csr_laplacian(ctx.h, dim, nx, ny, nz, row0, rows_per, ctypes.byref(h), ctypes.byref(glo), ctypes.byref(ghi)))
    assert glo.value == (plane if row0 > 0 else 0) and ghi.value == (plane if row0 + rows_per < nx * plane else 0)
    rng = np.random.default_rng(44)
    x = rng.standard_normal(nx * plane)
    xl = x[row0:row0 + rows_per]
    xg = np.concatenate([x[row0 - glo.value:row0], x[row0 + rows_per:row0 + rows_per + ghi.value]])
    dx, dg, dy = ctx.to_device(xl), ctx.to_device(xg if len(xg) else np.zeros(1)), ctx.empty(rows_per)
    check(ctx.lib.b2k_csr_spmv(ctx.h, h, dx.ptr, dg.ptr, dy.ptr))
    nnz = ctypes.c_int64()
    check(ctx.lib.b2k_csr_info(h, None, None, None, ctypes.byref(nnz)))
    assert nnz.value == A[row0:row0 + rows_per].nnz
    assert np.array_equal(dy.to_host(), (A @ x)[row0:row0 + rows_per]) or np.allclose(dy.to_host(), (A @ x)[row0:row0 + rows_per], rtol=1e-14, atol=1e-14)
    check(ctx.lib.b2k_csr_destroy(ctx.h, h))


@pytest.mark.parametrize("n,k", [(1000, 7), (5000, 3), (50001, 17), (300007, 64), (20000, 70)])
def test_gs_update_norm_gated_device_side_dgks_decision(ctx, n, k):
    """b2k_gs_update_norm_gated: the speculative refinement sweep runs iff  nrm != 0 && nrm < eta*onrm  (bvorthog.c:180), evaluated
    on the device from two scalars — through the register-tile kernel (small n / k <= 4), the tensor-map kernel and the generic
    kernel (k > 64).  Closed gate: w untouched bit for bit; open gate: identical to b2k_gs_update_norm."""
    from slepc_b200._b2k import check
    ld = n + (n % 2)
    rng = np.random.default_rng(n + 3 * k)
    H = np.zeros((ld, k), order="F")
    H[:n] = rng.standard_normal((n, k))
    w = rng.standard_normal(n)
    q = rng.standard_normal(k)
    dV, dq = ctx.to_device(H), ctx.to_device(q)
    eta = 0.7071
    wref = w - H[:n] @ q
    for onrm2, nrm2, runs in [(4.0, 1.0, True), (4.0, 2.0001, False), (4.0, 0.0, False), (1.0, 1.0, False), (2.0, 0.99, True), (-1.0, 1.0, False),
                              (4.0, float(np.nextafter((eta * 2.0) ** 2, 0)), None)]:
        g = ctx.to_device(np.array([onrm2, nrm2]))
        dw, out = ctx.to_device(w), ctx.to_device(np.array([-7.0]))
        check(ctx.lib.b2k_gs_update_norm_gated(ctx.h, dV.ptr, ld, n, k, dw.ptr, dq.ptr, out.ptr, g.ptr, g.at(1), eta))
        got = dw.to_host()
        on, nr = np.sqrt(max(onrm2, 0.0)), np.sqrt(max(nrm2, 0.0))
        expect = bool(nr != 0.0 and abs(nr) < eta * abs(on))             # the host's own expression (bvorthog.c:180)
        if runs is not None:
            assert expect == runs
        if expect:
            assert np.allclose(got, wref, rtol=1e-13, atol=1e-12 * k)
            assert np.isclose(out.to_host()[0], wref @ wref, rtol=1e-12)
            dw2, out2 = ctx.to_device(w), ctx.empty(1)
            check(ctx.lib.b2k_gs_update_norm(ctx.h, dV.ptr, ld, n, k, dw2.ptr, dq.ptr, out2.ptr))
            assert np.array_equal(dw2.to_host(), got) and out2.to_host()[0] == out.to_host()[0]
        else:
            assert np.array_equal(got, w)
        for a in (g, dw, out):
            a.free()
