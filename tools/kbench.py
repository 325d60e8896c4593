"""Per-kernel achieved HBM GB/s on one B200 (CUDA events on the launching stream, inputs >> L2).
Algorithmic bytes per SURVEY.md §8(d).  Usage: python tools/kbench.py [n] [reps]"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slepc_b200 import _b2k  # noqa: E402
from slepc_b200._b2k import check  # noqa: E402


def timeit(ctx, fn, reps):
    for _ in range(3):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 4096
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    ctx = _b2k.Context(0)
    lib, h = ctx.lib, ctx.h
    m = 65
    ld = n
    V = ctx.empty(ld * m)
    for j in range(m):
        check(lib.b2k_set_random(h, V.at(j * ld), n, 0, 100 + j))
    check(lib.b2k_scale(h, V.ptr, ld, n, m, 1.0 / np.sqrt(n / 3.0)))
    w = ctx.empty(n)
    w2 = ctx.empty(n)
    check(lib.b2k_set_random(h, w.ptr, n, 0, 7))
    c = ctx.empty(m + 1)
    c2 = ctx.empty(m + 1)
    check(lib.b2k_memset0(h, c.ptr, 8 * (m + 1)))
    res = []

    def rec(name, ms, nbytes, **kw):
        r = dict(kernel=name, ms=round(ms, 4), GBs=round(nbytes / ms / 1e6, 1), **kw)
        res.append(r)
        print(json.dumps(r), flush=True)

    rec("copy(n x 32)", timeit(ctx, lambda: check(lib.b2k_copy(h, V.at(32 * ld), ld, V.ptr, ld, n, 32)), reps), 16 * n * 32)
    for k in (1, 8, 16, 24, 32, 48, 64):
        rec("gs_dot", timeit(ctx, lambda: check(lib.b2k_gs_dot(h, V.ptr, ld, n, k, w.ptr, c.ptr)), reps), 8 * n * (k + 1), k=k)
    check(lib.b2k_memset0(h, c.ptr, 8 * (m + 1)))
    for mode in (1, 3):
        check(lib.b2k_gs_set_fused(mode))
        for k in (1, 8, 16, 32, 48, 64):
            rec("multvec", timeit(ctx, lambda: check(lib.b2k_multvec(h, V.ptr, ld, n, k, -1.0, 1.0, w.ptr, c.ptr)), reps),
                8 * n * (k + 2), k=k, single_sweep=mode)
        for k in (16, 32, 48, 64):
            rec("gs_update_norm", timeit(ctx, lambda: check(lib.b2k_gs_update_norm(h, V.ptr, ld, n, k, w.ptr, c.ptr, c2.ptr)), reps),
                8 * n * (k + 2), k=k, single_sweep=mode)
    for mode in (0, 1, 3):
        check(lib.b2k_gs_set_fused(mode))
        for k in (1, 4, 8, 16, 24, 32, 48, 64):
            rec("gs_update_dot", timeit(ctx, lambda: check(lib.b2k_gs_update_dot(h, V.ptr, ld, n, k, w.ptr, c.ptr, c2.ptr)), reps),
                8 * n * (k + 2), k=k, single_sweep=mode, note="bytes of ONE read of V + w read/write")
    check(lib.b2k_gs_set_fused(3))
    rec("scale_rsqrt", timeit(ctx, lambda: check(lib.b2k_scale_rsqrt(h, w.ptr, n, c.ptr)), reps), 16 * n)
    rec("sumsq", timeit(ctx, lambda: check(lib.b2k_sumsq(h, w.ptr, n, n, 1, c2.ptr)), reps), 8 * n)
    # mult_inplace: k=64 -> 32 columns
    Q = np.zeros((64, 64), order="F")
    Q[:, :] = np.linalg.qr(np.random.default_rng(0).standard_normal((64, 64)))[0]
    dQ = ctx.to_device(Q)
    for (k, s, e) in ((64, 0, 32), (64, 0, 44), (32, 0, 16)):
        ms = timeit(ctx, lambda: check(lib.b2k_mult_inplace(h, V.ptr, ld, n, k, s, e, dQ.ptr, 64, 0)), max(2, reps // 3))
        rec("mult_inplace", ms, 8 * n * (k + e - s), k=k, nout=e - s, TFLOPs=round(2.0 * n * k * (e - s) / ms / 1e9, 2))
    # SpMV on stencils of ~n rows
    for dim, dims in ((2, (int(round(n ** 0.5)), int(round(n ** 0.5)), 1)), (3, (int(round(n ** (1 / 3))),) * 3), (1, (n, 1, 1))):
        nx, ny, nz = dims
        N = nx * ny * nz
        if N > n:
            continue
        A = ctypes.c_void_p()
        check(lib.b2k_csr_laplacian(h, dim, nx, ny, nz, 0, N, ctypes.byref(A), None, None))
        nnz = ctypes.c_int64()
        check(lib.b2k_csr_info(A, None, None, None, ctypes.byref(nnz)))
        ms = timeit(ctx, lambda: check(lib.b2k_csr_spmv(h, A, w.ptr, None, w2.ptr)), reps)
        rec("spmv_csr", ms, 12 * nnz.value + 4 * (N + 1) + 16 * N, stencil=f"lap{dim}d", rows=N, nnz=nnz.value)
        check(lib.b2k_csr_destroy(h, A))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
