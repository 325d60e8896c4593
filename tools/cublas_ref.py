#!/usr/bin/env python
"""What SLEPc's own CUDA BV back-end would execute on this GPU, operation by operation (column
"cuBLAS/cuSPARSE on B200"): the library kernels behind bvcuda.cu / sveccuda.cu and PETSc's AIJCUSPARSE MatMult, timed with
CUDA events on the C2 shapes (n = 4096^2 rows, k columns) and reported with the SAME algorithmic-byte formulas as bench.py.

  BVDotVec      cublasDgemv 'T' on the n x k panel        bvcuda.cu:204-264     -> torch.mv
  BVMultVec     cublasDgemv 'N'                           bvcuda.cu:45-60       -> torch.addmv
  BVNorm        cublasDnrm2                               bvcuda.cu:290-303     -> torch.linalg.vector_norm
  BVScale       cublasDscal                               bvcuda.cu:269-285     -> Tensor.mul_
  BVMultInPlace cublasDgemm into a workspace + cudaMemcpy2D back   bvcuda.cu:65-112   -> torch.mm + copy_
  MatMult       cusparseSpMV, CSR, fp64 / int32           PETSc aijcusparse     -> torch sparse_csr @ x

This is a LIBRARY baseline measured beside the product, never part of it (torch is not used by the product path).
    python tools/cublas_ref.py [--n 16777216] [--reps 10] > gpurun_out/cublas_ref.jsonl
"""
import argparse
import json

import torch


def timeit(fn, reps, dev):
    if dev == "cpu":
        import time
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def laplacian2d_csr(g, dev):
    """5-point stencil on a g x g grid (ex2.c:39-54) as torch sparse CSR with int32 indices (what PETSc hands to cuSPARSE)"""
    N = g * g
    r = torch.arange(N, device=dev, dtype=torch.int64)
    i, j = r // g, r % g
    cand = torch.stack([r - g, r - 1, r, r + 1, r + g], dim=1)
    valid = torch.stack([i > 0, j > 0, torch.ones_like(i, dtype=torch.bool), j < g - 1, i < g - 1], dim=1)
    vals = torch.full((N, 5), -1.0, dtype=torch.float64, device=dev)
    vals[:, 2] = 4.0
    crow = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    crow[1:] = torch.cumsum(valid.sum(dim=1), dim=0)
    A = torch.sparse_csr_tensor(crow.to(torch.int32), cand[valid].to(torch.int32), vals[valid], size=(N, N), dtype=torch.float64, device=dev)
    return A, int(crow[-1].item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096 * 4096)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--kmax", type=int, default=64)
    args = ap.parse_args()
    dev, n, reps = args.device, args.n, args.reps
    g = int(round(n ** 0.5))
    n = g * g
    out = []

    def rec(op, ms, nbytes, **kw):
        d = dict(op=op, ms=round(ms, 4), GBs=round(nbytes / ms / 1e6, 1), n=n, **kw)
        out.append(d)
        print(json.dumps(d), flush=True)

    torch.manual_seed(0)
    V = torch.randn(args.kmax, n, dtype=torch.float64, device=dev)      # row k of this array = column k of the column-major basis
    w = torch.randn(n, dtype=torch.float64, device=dev)
    for k in [x for x in (16, 32, 48, 64) if x <= args.kmax]:
        Vk = V[:k]
        c = torch.randn(k, dtype=torch.float64, device=dev) * 1e-3
        rec("BVDotVec gemv'T' (cublasDgemv)", timeit(lambda: torch.mv(Vk, w), reps, dev), 8.0 * n * (k + 1), k=k)
        rec("BVMultVec gemv'N' (cublasDgemv)", timeit(lambda: w.addmv_(Vk.t(), c, alpha=-1.0), reps, dev), 8.0 * n * (k + 2), k=k)
    rec("BVNorm (nrm2)", timeit(lambda: torch.linalg.vector_norm(w), reps, dev), 8.0 * n)
    rec("BVScale (scal)", timeit(lambda: w.mul_(1.0000001), reps, dev), 16.0 * n)
    for k, nout in [(args.kmax, args.kmax // 2), (args.kmax // 2, args.kmax // 4)]:
        Q = torch.randn(nout, k, dtype=torch.float64, device=dev)
        Vk = V[:k]

        def restart():
            W = torch.mm(Q, Vk)                     # workspace, bvcuda.cu:90-94
            V[:nout].copy_(W)                       # cudaMemcpy2D back, bvcuda.cu:104

        ms = timeit(restart, max(2, reps // 3), dev)
        rec("BVMultInPlace (cublasDgemm + copy back)", ms, 8.0 * n * (k + nout), k=k, nout=nout, TFLOPs=round(2.0 * n * k * nout / ms / 1e9, 2))
    A, nnz = laplacian2d_csr(g, dev)
    x = torch.randn(n, dtype=torch.float64, device=dev)
    rec("MatMult CSR (cusparseSpMV)", timeit(lambda: A @ x, reps, dev), 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n, nnz=nnz)
    # one Lanczos step at column k=48 as the reference schedules it (2 CGS passes: 2 x (gemv'T' + gemv'N'), SpMV, nrm2, scal)
    by = {(d["op"], d.get("k")): d["ms"] for d in out}
    if ("BVDotVec gemv'T' (cublasDgemv)", 48) in by:
        step = 2 * (by[("BVDotVec gemv'T' (cublasDgemv)", 48)] + by[("BVMultVec gemv'N' (cublasDgemv)", 48)]) + by[("MatMult CSR (cusparseSpMV)", None)] \
            + by[("BVNorm (nrm2)", None)] + by[("BVScale (scal)", None)]
        print(json.dumps({"op": "one Lanczos step at k=48, library kernels only (no host syncs, no micro-kernels)", "ms": round(step, 4), "n": n}), flush=True)


if __name__ == "__main__":
    main()
