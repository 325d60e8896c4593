#!/usr/bin/env python
"""Launch each round-2 hot kernel a few times at the benchmark shapes so that ONE `ncu --set full` run can capture them all:
    ncu --set full --clock-control none --import-source on -k regex:'k_gram_tma|k_vq_tma|k_spmv_sell_pipe|k_gs_tma|k_dotvec' \
        --launch-skip 0 -c 14 -o gpurun_out/r02_kernels python tools/ncu_targets.py
Order of the captured launches (one warm-up of each is skipped by the kernel filter count, see tools/jobs/r02g.sh):
  gram 64x64, gram 25x25, restart GEMM 64->32, 48->24, 25->13, SpMM 3-D k=16, SpMM 2-D k=8, SpMV 3-D, dotvec k=48, gs<DOT> k=48, gs<NRM> gated k=48"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slepc_b200 import _b2k                     # noqa: E402
from slepc_b200._b2k import check               # noqa: E402


def main():
    ctx = _b2k.Context(0)
    lib = ctx.lib
    n = 4096 * 4096
    ld = n
    V = ctx.empty(ld * 65)
    check(lib.b2k_set_random(ctx.h, V.ptr, ld * 65, 0, 3))
    M = ctx.empty(64 * 64)
    for ky in (64, 25):
        check(lib.b2k_dot(ctx.h, V.ptr, ld, V.ptr, ld, n, ky, ky, M.ptr, ky))
    for k, e in ((64, 32), (48, 24), (25, 13)):
        Qk = ctx.to_device(np.asfortranarray(np.random.default_rng(1).standard_normal((k, k)) / k))
        check(lib.b2k_mult_inplace(ctx.h, V.ptr, ld, n, k, 0, e, Qk.ptr, k, 0))
    check(lib.b2k_set_random(ctx.h, V.ptr, ld * 65, 0, 3))
    for dim, g, k in ((3, 256, 16), (2, 4096, 8)):
        h = ctypes.c_void_p()
        glo, ghi = ctypes.c_int64(), ctypes.c_int64()
        nn = g ** dim
        check(lib.b2k_csr_laplacian(ctx.h, dim, g, g, g, 0, nn, ctypes.byref(h), ctypes.byref(glo), ctypes.byref(ghi)))
        check(lib.b2k_csr_spmm(ctx.h, h, V.ptr, nn, None, 0, V.at(32 * ld), nn, k))
        if dim == 3:
            check(lib.b2k_csr_spmv(ctx.h, h, V.ptr, None, V.at(32 * ld)))
        ctx.sync()
        check(lib.b2k_csr_destroy(ctx.h, h))
    # the three sweeps of a Lanczos step at k = 48 (as bvb200.c schedules them)
    k = 48
    w = V.at(k * ld)
    c = ctx.empty(2 * 72)
    check(lib.b2k_gs_dot(ctx.h, V.ptr, ld, n, k, w, c.ptr))
    check(lib.b2k_gs_update_dot(ctx.h, V.ptr, ld, n, k, w, c.ptr, c.at(72)))
    check(lib.b2k_gs_update_norm(ctx.h, V.ptr, ld, n, k, w, c.at(72), c.ptr))
    ctx.sync()
    print("ncu targets done")


if __name__ == "__main__":
    main()
