#!/usr/bin/env python
"""Round-2 kernel timings on one B200 (CUDA events through the C ABI, inputs far larger than L2): the one-sweep Gram kernel
against the column-by-column sweeps it replaces (B2K_GRAM_TMA=0 in a second process), the SELL SpMM against k SpMVs, the
restart GEMM at the C3 shape.     python tools/kbench2.py > gpurun_out/kbench2.jsonl"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slepc_b200 import _b2k                     # noqa: E402
from slepc_b200._b2k import check               # noqa: E402


def timeit(ctx, fn, reps=5):
    for _ in range(2):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


def main():
    ctx = _b2k.Context(0)
    lib = ctx.lib
    n = 4096 * 4096
    ld = n
    V = ctx.empty(ld * 65)
    check(lib.b2k_set_random(ctx.h, V.ptr, ld * 65, 0, 3))
    M = ctx.empty(64 * 64)
    tag = "gram_tma" if os.environ.get("B2K_GRAM_TMA", "1") != "0" else "dotvec_per_column"
    for ky, kx, same in [(64, 64, True), (32, 32, True), (64, 32, False), (25, 25, True)]:
        X = V.ptr if same else V.at(32 * ld)
        kxx = kx
        ms = timeit(ctx, lambda: check(lib.b2k_dot(ctx.h, V.ptr, ld, X, ld, n, ky, kxx, M.ptr, ky)))
        byt = 8.0 * n * (ky if same else ky + kx)
        print(json.dumps(dict(kernel="b2k_dot", impl=tag, n=n, ky=ky, kx=kx, same_block=same, ms=round(ms, 4), GBs_one_sweep_bytes=round(byt / ms / 1e6, 1),
                              TFLOPs=round(2.0 * n * ky * kx / ms / 1e9, 2))), flush=True)
    if tag == "gram_tma":
        # SpMM vs k SpMVs on the 5-point 4096^2 and the 7-point 256^3 stencils
        for dim, g in [(2, 4096), (3, 256)]:
            h = ctypes.c_void_p()
            glo, ghi = ctypes.c_int64(), ctypes.c_int64()
            nn = g ** dim
            check(lib.b2k_csr_laplacian(ctx.h, dim, g, g, g, 0, nn, ctypes.byref(h), ctypes.byref(glo), ctypes.byref(ghi)))
            nnz = ctypes.c_int64()
            check(lib.b2k_csr_info(h, None, None, None, ctypes.byref(nnz)))
            for k in (8, 16, 32):
                Y = V.at(32 * ld)
                ms_mm = timeit(ctx, lambda: check(lib.b2k_csr_spmm(ctx.h, h, V.ptr, nn, None, 0, Y, nn, k)))

                def loop():
                    for j in range(k):
                        check(lib.b2k_csr_spmv(ctx.h, h, V.at(j * nn), None, V.at((32 + j) * nn)))
                ms_mv = timeit(ctx, loop)
                print(json.dumps(dict(kernel="b2k_csr_spmm", stencil=f"{dim}-D {g}^{dim}", k=k, ms_spmm=round(ms_mm, 4), ms_k_spmv=round(ms_mv, 4),
                                      speedup=round(ms_mv / ms_mm, 2), GBs_spmm_algorithmic=round((12.0 * nnz.value * ((k + 15) // 16 if k > 8 else 1) + 16.0 * nn * k) / ms_mm / 1e6, 1))), flush=True)
            check(lib.b2k_csr_destroy(ctx.h, h))
        # restart GEMM shapes
        Q = ctx.to_device(np.asfortranarray(np.random.default_rng(1).standard_normal((64, 64))))
        for k, s, e in [(64, 0, 32), (25, 0, 13), (32, 0, 16), (48, 0, 24)]:
            Qk = ctx.to_device(np.asfortranarray(np.random.default_rng(1).standard_normal((k, k)) / k))
            ms = timeit(ctx, lambda: check(lib.b2k_mult_inplace(ctx.h, V.ptr, ld, n, k, s, e, Qk.ptr, k, 0)))
            print(json.dumps(dict(kernel="b2k_mult_inplace", n=n, kin=k, nout=e - s, ms=round(ms, 4), GBs=round(8.0 * n * (k + e - s) / ms / 1e6, 1),
                                  TFLOPs=round(2.0 * n * k * (e - s) / ms / 1e9, 2))), flush=True)
            check(lib.b2k_set_random(ctx.h, V.ptr, ld * 65, 0, 3))


if __name__ == "__main__":
    main()
