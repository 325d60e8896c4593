#!/usr/bin/env python
"""bench.py — the hot path of EPSSolve (Krylov-Schur: SpMV + fused DGKS Gram-Schmidt + in-place restart) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3] [--tts c3|c3small|none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N … bench.py --gpus N …      (N > 1)

A "step" is ONE Krylov-Schur restart cycle (krylovschur.c:248-332): ~ncv-l Lanczos expansion steps (MatMult + DGKS
orthonormalisation), the projected solve on the host and the restart V <- V Q.

Legs of the b200 arm (one JSON line, rank 0):
  value / roofline   workload c2 (BASELINE.json configs[1]): 2-D 5-point Laplacian 4096x4096 per GPU (weak scaling: the grid is
                     (4096 N) x 4096, slab-partitioned), symmetric Krylov-Schur nev=20 ncv=64; K restart cycles, inputs resident
                     in HBM, CUDA events around every kernel launch of the timed region (per-class GB/s, dominant-kernel roofline)
  e2e                the same through the host API: host CSR + host start vector in, Ritz vectors out (copies in the timed region)
  time_to_solution   BASELINE.json's other metric, the north star: workload c3 (configs[2]) 7-point Laplacian 512^3 = 1.3e8 rows,
                     nev=10 (ncv=25), STRONG-scaled over the N GPUs, solved TO CONVERGENCE inside this run and checked before
                     anything is printed: nconv >= nev, residuals ||Ax-kx||/|k| < 5 tol, every value within 1e-10 (relative) of
                     the analytic spectrum.  The N=1 run leaves its time in /tmp so that the N>1 runs of the same lease can
                     print strong_efficiency = t1 / (N tN) next to their own time (the driver computes its own from the lines).
  latency_leg        the same global 1024^2 problem on every N (strong scaling in the launch/latency-bound regime), us per step
  multi_gpu_parity   N > 1: the row-partitioned correctness cases of tests/mgpu_cases.py (BV reductions over the NVLink
                     mailboxes vs ncclAllReduce, halo SpMV vs the closed-form stencil, Lanczos/Arnoldi/SVD solves vs numpy)
  library_gpu_baseline  N = 1: cuBLAS / cuSPARSE in the reference's schedule on the same GPU (baseline/libbase.cu)
  cpu_baseline       N = 1: the restated reference on the host cores, same protocol as `--impl reference`
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(dim=2, nx=4096, ny=4096, nz=1, nev=20, ncv=64, scaling="weak",
               name="2-D Laplacian 5-point 4096x4096 per GPU, Krylov-Schur HEP nev=20 ncv=64 tol=1e-8 (BASELINE configs[1])"),
    "c3": dict(dim=3, nx=512, ny=512, nz=512, nev=10, ncv=25, scaling="strong",
               name="3-D Laplacian 7-point 512^3 row-partitioned, Krylov-Schur HEP nev=10 ncv=25 tol=1e-8 (BASELINE configs[2])"),
}
TTS = {
    "c3": dict(dim=3, g=512, nev=10, ncv=25, name="C3: 3-D Laplacian 7-point 512^3 (134 217 728 rows), nev=10 ncv=25 tol=1e-8, strong-scaled"),
    "c3small": dict(dim=3, g=128, nev=10, ncv=25, name="3-D Laplacian 7-point 128^3 (reduced C3), nev=10 ncv=25 tol=1e-8, strong-scaled"),
}
METRIC = "EPSSolve Krylov-Schur hot-path throughput (matrix rows x Lanczos steps per second; s/restart-cycle and per-kernel GB/s in extra keys)"
UNIT = "row-steps/s"
KCLASS = ["dotvec", "multvec", "gs_fused", "spmv", "gemm_restart", "elementwise"]
TOL = 1e-8


def config_of(wl, world):
    """the SAME dict in both arms (the driver compares them)"""
    nx_total = wl["nx"] * (world if wl["scaling"] == "weak" else 1)
    rows_global = nx_total * wl["ny"] * (wl["nz"] if wl["dim"] == 3 else 1)
    return {"workload": wl["name"], "rows_global": rows_global, "rows_per_gpu": rows_global // world,
            "l2": "inputs larger than L2 (basis + matrix >= 1 GB per GPU vs 126 MB L2)", "step": "one Krylov-Schur restart cycle"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


# ------------------------------------------------------------------------------------------------------------
def cpu_sample(wl, steps, warmup):
    """The reference's CPU path as restated in oracle/ (same C host driver, host-memory BV type with BLAS gemv/gemm + OpenMP
    CSR SpMV) on the box's host cores: `warmup` untimed + `steps` timed restart cycles on a BOUNDED sample of the workload —
    a slab of the grid with the same stencil, the same row length and the same nev/ncv (the metric is per row, and the CPU
    path is bandwidth-bound with every array far larger than cache, so the slab's rows x steps / s is the workload's).
    The unmodified SLEPc needs PETSc+MPI, absent here."""
    from oracle import cpu_plugin as CP
    from slepc_b200 import slepc as SL
    from slepc_b200.slepc import S
    CP.load()
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    if CP.threads() < avail:                       # torchrun sets OMP_NUM_THREADS=1 for every rank: use all host cores anyway
        CP.set_threads(avail)
    cores = CP.threads()
    t0 = time.time()
    if wl["dim"] == 2:
        nx, ny, nz = wl["nx"] // 4, wl["ny"], 1    # 1024 x 4096 slab of the 4096 x 4096 grid: 4.2 M rows, 2.2 GB basis
    else:
        nx, ny, nz = 32, wl["ny"], wl["nz"]        # a 32-plane slab of the 512^3 grid: 8.4 M rows
    M = CP.mat_laplacian(wl["dim"], nx, ny, nz)
    rows = nx * ny * nz
    eps = SL.EPS(M, hermitian=True)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, wl["nev"], wl["ncv"], SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, TOL, 1000000)
    eps.cycles(1)                                   # first (unrestarted) cycle: page-faults the basis, never timed
    if warmup:
        eps.cycles(warmup)
    bv = eps.bv()
    m0 = bv.counters()[1]
    t1 = time.time()
    done = eps.cycles(steps)
    dt = time.time() - t1
    nsteps = bv.counters()[1] - m0
    value = rows * nsteps / dt
    triad = CP.stream_triad_gbs(1 << 26, 3)
    sample = (f"{done} restart cycles ({nsteps} Lanczos steps) on a {nx}x{ny}x{nz} slab of the workload's grid = {rows} rows "
              f"(same stencil, nev, ncv), after 1 unrestarted + {warmup} warm-up cycles; set-up {t1 - t0:.1f}s")
    cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "stream_triad_gbs": triad,
           "lanczos_steps_per_s": nsteps / dt, "seconds": dt,
           "note": "restated reference (oracle/oracle_cpu.c under the same C host driver); SLEPc+PETSc cannot be built here"}
    return cpu, done, dt


def run_reference(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cpu, done, dt = cpu_sample(wl, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(wl, args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
def build_host_csr_slab(lib, wl, nx_total, rank, world):
    """host (pinned) CSR of this rank's slab with GLOBAL column indices + the halo plan, for the e2e leg"""
    dim, ny, nz = wl["dim"], wl["ny"], wl["nz"] if wl["dim"] == 3 else 1
    if dim == 1:
        ny = 1
    plane = ny * nz
    base, rem = divmod(nx_total, world)
    p0 = rank * base + min(rank, rem)
    npl = base + (1 if rank < rem else 0)
    row0, nloc, N = p0 * plane, npl * plane, nx_total * plane
    g = np.arange(row0, row0 + nloc, dtype=np.int64)
    i, r = g // plane, g % plane
    j, k = r // nz, r % nz
    cand = [(g - plane, i > 0)]
    if dim >= 2:
        cand.append((g - nz, j > 0))
    if dim >= 3:
        cand.append((g - 1, k > 0))
    cand.append((g, np.ones(nloc, bool)))
    if dim >= 3:
        cand.append((g + 1, k < nz - 1))
    if dim >= 2:
        cand.append((g + nz, j < ny - 1))
    cand.append((g + plane, i < nx_total - 1))
    cols = np.stack([c for c, _ in cand], axis=1)
    valid = np.stack([m for _, m in cand], axis=1)
    center = len(cand) // 2
    vals = np.full(cols.shape, -1.0)
    vals[:, center] = 2.0 * dim
    rowptr = np.zeros(nloc + 1, dtype=np.int32)
    np.cumsum(valid.sum(axis=1), out=rowptr[1:])
    nnz = int(rowptr[-1])

    def pinned(arr, dtype):
        p = ctypes.c_void_p()
        assert lib.b2k_host_alloc(ctypes.byref(p), arr.size * np.dtype(dtype).itemsize) == 0
        out = np.frombuffer((ctypes.c_char * (arr.size * np.dtype(dtype).itemsize)).from_address(p.value), dtype=dtype)
        out[:] = arr
        return out, p

    rp_h, rp_p = pinned(rowptr, np.int32)
    ci_h, ci_p = pinned(cols[valid].astype(np.int32), np.int32)
    va_h, va_p = pinned(vals[valid], np.float64)
    del cols, valid, vals
    halo = dict(recvrank=[], recvcount=[], sendrank=[], sendcount=[], sendidx=[])
    if rank > 0:
        halo["recvrank"].append(rank - 1); halo["recvcount"].append(plane)
        halo["sendrank"].append(rank - 1); halo["sendcount"].append(plane); halo["sendidx"].extend(range(0, plane))
    if rank < world - 1:
        halo["recvrank"].append(rank + 1); halo["recvcount"].append(plane)
        halo["sendrank"].append(rank + 1); halo["sendcount"].append(plane); halo["sendidx"].extend(range(nloc - plane, nloc))
    return dict(N=N, row0=row0, nloc=nloc, nnz=nnz, rowptr=(rp_h, rp_p), colidx=(ci_h, ci_p), val=(va_h, va_p), halo=halo, keep=(rp_h, ci_h, va_h))


def library_baseline(wl):
    """cuBLAS/cuSPARSE in the reference's schedule on this GPU (baseline/libbase.cu), same shapes as the device-resident leg"""
    path = os.path.join(ROOT, "baseline", "libbase.so")
    if not os.path.exists(path):
        return {"unavailable": "baseline/libbase.so not built (make)"}
    try:
        lb = ctypes.CDLL(path)
        lb.libbase_last_error.restype = ctypes.c_char_p
        out = (ctypes.c_double * 16)()
        kmin = wl["ncv"] // 2                                     # the thick restart keeps about half of the basis
        kprobe = (kmin + wl["ncv"]) // 2
        rc = lb.libbase_run(ctypes.c_int(wl["dim"]), ctypes.c_int64(wl["nx"]), ctypes.c_int64(wl["ny"]), ctypes.c_int64(wl["nz"]),
                            ctypes.c_int(kmin), ctypes.c_int(wl["ncv"]), ctypes.c_int(kmin), ctypes.c_int(3), ctypes.c_int(1), ctypes.c_int(kprobe), out)
        if rc:
            return {"unavailable": f"libbase_run failed ({rc}): {lb.libbase_last_error().decode()}"}
        n, nnz = out[10], out[9]
        per = dict(gemv_T=out[3], gemv_N=out[4], spmv=out[5], nrm2=out[6], scal=out[7], gemm_plus_copyback=out[8])
        gbs = dict(gemv_T=8 * n * (kprobe + 1) / out[3] / 1e6, gemv_N=8 * n * (kprobe + 2) / out[4] / 1e6,
                   spmv=(12 * nnz + 4 * (n + 1) + 16 * n) / out[5] / 1e6, gemm_plus_copyback=8 * n * (wl["ncv"] + kmin) / out[8] / 1e6)
        return {"what": "cublasDgemv T/N x2 + cusparseSpMV + cublasDnrm2 + cublasDscal per Lanczos step, cublasDgemm + cudaMemcpy2D per restart; "
                        "device pointer mode, no host syncs, the reference's coefficient micro-kernels and blocking copies NOT counted (lower bound)",
                "ms_per_restart_cycle": out[0], "ms_per_lanczos_step": out[1], "lanczos_steps_per_cycle": int(out[2]),
                "row_steps_per_s": n * out[2] / (out[0] / 1e3), "per_op_ms_at_k": {"k": kprobe, **per}, "per_op_gbs_algorithmic": gbs}
    except Exception as e:                                       # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_b200(args, wl):
    rank, world, local = dist_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}"
    import torch
    import torch.distributed as dist
    from slepc_b200 import _b2k
    from slepc_b200 import slepc as SL
    from slepc_b200.slepc import S
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 arm has no CPU fallback (use --impl reference for the CPU baseline)")
    from slepc_b200 import dist as D
    lib = _b2k.load()
    D.init()                                       # B2KInitialize(local GPU) + NCCL communicator over the ranks
    ctx = S.B2KGetContext()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        S.B2KDeviceSynchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def launches():
        n = ctypes.c_uint64()
        lib.b2k_ctx_launches(ctx, ctypes.byref(n))
        return n.value

    def syncs():
        n = ctypes.c_uint64()
        lib.b2k_ctx_syncs(ctx, ctypes.byref(n))
        return n.value

    def hbm_used():
        fr, to = ctypes.c_size_t(), ctypes.c_size_t()
        lib.b2k_mem_info(ctx, ctypes.byref(fr), ctypes.byref(to))
        return to.value - fr.value

    cfg = config_of(wl, world)
    nx_total = wl["nx"] * (world if wl["scaling"] == "weak" else 1)
    rows_global = cfg["rows_global"]

    # ---------------- device-resident leg: K restart cycles, inputs already in HBM -----------------------------
    hbm0 = hbm_used()
    M = SL.Mat.laplacian(wl["dim"], nx_total, wl["ny"], wl["nz"])
    hbm_matrix = hbm_used() - hbm0
    eps = SL.EPS(M, hermitian=True)
    S.EPSSetDimensions(eps.h, wl["nev"], wl["ncv"], SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, TOL, 100000000)
    eps.cycles(args.warmup)
    hbm_solver = hbm_used() - hbm0
    bv = eps.bv()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    _b2k.check(lib.b2k_prof_enable(ctx, 1))
    m0, l0, g0, s0 = bv.counters()[1], launches(), bv.counters()[0], syncs()
    _b2k.check(lib.b2k_timer_start(ctx))
    done = 0
    for _ in range(args.steps):
        done += eps.cycles(1)
    ms = ctypes.c_double()
    _b2k.check(lib.b2k_timer_stop_ms(ctx, ctypes.byref(ms)))
    barrier()
    clocks = sampler.stop() if sampler else None
    steps = bv.counters()[1] - m0
    gs_passes = bv.counters()[0] - g0
    nl = launches() - l0
    nsync = syncs() - s0 - 1                      # the timer's own wait is not the path's
    t_ms = allmax(ms.value)
    prof = {}
    for cid, name in enumerate(KCLASS):
        n_, ms_, b_ = ctypes.c_uint64(), ctypes.c_double(), ctypes.c_double()
        _b2k.check(lib.b2k_prof_get(ctx, cid, ctypes.byref(n_), ctypes.byref(ms_), ctypes.byref(b_)))
        f_ = ctypes.c_double()
        _b2k.check(lib.b2k_prof_get_flops(ctx, cid, ctypes.byref(f_)))
        prof[name] = dict(launches=n_.value, ms=ms_.value, bytes=b_.value, flops=f_.value)
    _b2k.check(lib.b2k_prof_enable(ctx, 0))
    if done < args.steps:
        raise SystemExit(f"bench.py: the solve converged after {done} timed cycles (< --steps {args.steps}); lower --steps/--warmup")
    value = rows_global * steps / (t_ms / 1e3)
    peak, peak_src = measured_peaks()
    kernels = {}
    for name, p in prof.items():
        if p["launches"]:
            kernels[name] = dict(launches=p["launches"], ms_total=round(p["ms"], 3), avg_ms=p["ms"] / p["launches"],
                                 share_of_step=p["ms"] / ms.value, achieved_gbs=p["bytes"] / p["ms"] / 1e6,
                                 frac_of_peak=p["bytes"] / p["ms"] / 1e6 / peak, frac_of_nominal_8tbs=p["bytes"] / p["ms"] / 1e6 / 8000.0)
            if p["flops"]:                       # level-3 (restart GEMM): at the FP64 ridge, so the FP64 pipe is the second roofline
                tf = p["flops"] / p["ms"] / 1e9
                clk = (clocks or {}).get("sm_mhz") or 1965.0
                kernels[name].update(fp64_tflops=tf, frac_of_fp64_peak_at_sampled_clock=tf / (148 * 64 * 2 * clk * 1e6 / 1e12),
                                     fp64_note="FP64 peak = 148 SMs x 64 FMA/clk x SM clock sampled under load (DMMA = DFMA rate, tools/dmma_probe.cu)")
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"])
    gs_bytes = sum(prof[k]["bytes"] for k in ("dotvec", "multvec", "gs_fused"))
    gs_ms = sum(prof[k]["ms"] for k in ("dotvec", "multvec", "gs_fused"))
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["achieved_gbs"] / peak, "frac_of_nominal_8tbs": kernels[dom]["achieved_gbs"] / 8000.0,
                "traffic": None, "peak_source": peak_src,
                "launches": kernels[dom]["launches"], "avg_launch_ms": kernels[dom]["avg_ms"],
                "note": "achieved = algorithmic bytes (SURVEY.md §8d, DESIGN.md) / CUDA-event time of every launch of the class in the timed region"}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if dom in tr:
            roofline["traffic"] = tr[dom]["ratio"] * prof[dom]["bytes"] / prof[dom]["launches"]
            roofline["traffic_source"] = f"ncu dram bytes / algorithmic bytes = {tr[dom]['ratio']} for {tr[dom]['kernel']} ({tr[dom]['report']}) x this run's bytes per launch"
    except Exception:
        pass
    eps.destroy()
    M.destroy()

    # ---------------- e2e leg: host CSR + host start vector in, Ritz vectors out -------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            h = build_host_csr_slab(lib, wl, nx_total, rank, world)
            def pinned_f64(nelem):                      # the step's host buffers are pinned, like the CSR arrays
                p = ctypes.c_void_p()
                assert lib.b2k_host_alloc(ctypes.byref(p), 8 * nelem) == 0
                return np.frombuffer((ctypes.c_char * (8 * nelem)).from_address(p.value), dtype=np.float64), p
            v0, v0_p = pinned_f64(h["nloc"])
            v0[:] = np.sin(0.37 * np.arange(h["row0"], h["row0"] + h["nloc"]) + 0.1) + 0.5
            out, out_p = pinned_f64(h["nloc"])
            hb0, db0 = ctypes.c_uint64(), ctypes.c_uint64()
            lib.b2k_ctx_copy_bytes(ctx, ctypes.byref(hb0), ctypes.byref(db0))
            barrier()
            t0 = time.perf_counter()
            ph = {}

            def phase(name, t_prev):
                S.B2KDeviceSynchronize()
                t = time.perf_counter()
                ph[name] = t - t_prev
                return t
            A = SL.Mat()
            S.MatCreateB200CSR(h["N"], h["N"], h["row0"], h["row0"] + h["nloc"], h["rowptr"][1], h["colidx"][1], h["val"][1],
                               h["row0"], h["row0"] + h["nloc"], A.ref)
            tp = phase("matrix_upload_and_sell", t0)
            if world > 1:
                hl = h["halo"]
                i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
                rr, rc, sr, sc, si = i32(hl["recvrank"]), i32(hl["recvcount"]), i32(hl["sendrank"]), i32(hl["sendcount"]), i32(hl["sendidx"])
                pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
                S.MatB200CSRSetHalo(A.h, len(rr), pp(rr), pp(rc), len(sr), pp(sr), pp(sc), pp(si))
            tp = phase("halo_setup", tp)
            e2 = SL.EPS(A, hermitian=True)
            S.EPSSetDimensions(e2.h, wl["nev"], wl["ncv"], SL.PETSC_DETERMINE)
            S.EPSSetTolerances(e2.h, TOL, 100000000)
            x0, _ = A.create_vecs()
            x0.set_values(v0)
            S.EPSSetInitialSpace(e2.h, 1, (ctypes.c_void_p * 1)(x0.h))
            ke = 0
            for it in range(args.steps):
                ke += e2.cycles(1)
                if it == 0:
                    tp = phase("solver_setup_and_first_unrestarted_cycle", tp)
            tp = phase("restart_cycles", tp)
            bv2 = e2.bv()
            steps2 = bv2.counters()[1]
            checksum = 0.0
            for j in range(wl["nev"]):                 # device→host read of the result: the nev leading Ritz vectors
                S.BVGetColumnHost(bv2.h, j, out.ctypes.data_as(ctypes.c_void_p))
                checksum += float(out[0])
            tp = phase("ritz_vectors_to_host", tp)
            barrier()
            dt = allmax(time.perf_counter() - t0)
            hb1, db1 = ctypes.c_uint64(), ctypes.c_uint64()
            lib.b2k_ctx_copy_bytes(ctx, ctypes.byref(hb1), ctypes.byref(db1))
            e2e = {"value": rows_global * steps2 / dt, "unit": UNIT, "h2d_bytes_per_step": (hb1.value - hb0.value) / max(ke, 1),
                   "d2h_bytes_per_step": (db1.value - db0.value) / max(ke, 1), "seconds": dt, "steps": ke, "lanczos_steps": steps2,
                   "phases_s_rank0": {k: round(v, 4) for k, v in ph.items()},
                   "what": "MatCreateB200CSR(host CSR, pinned) + EPSSetInitialSpace(host vector) + K restart cycles (incl. the first, "
                           "unrestarted one) + BVGetColumnHost of the nev leading Ritz vectors; wall clock, max over ranks"}
            e2.destroy()
            x0.destroy()
            A.destroy()
            del v0, out
            for p in (h["rowptr"][1], h["colidx"][1], h["val"][1], v0_p, out_p):
                lib.b2k_host_free(p)
            del h
        except Exception as ex:                     # noqa: BLE001 — a failed optional leg must not cost the headline line
            e2e = {"error": (type(ex).__name__ + ": " + str(ex))[:400]}

    # ---------------- latency leg: the SAME global 1024^2 problem on every N, solved to convergence ----------------
    lat = None
    if not args.no_latency:
        try:
            g = 1024
            Mt = SL.Mat.laplacian(2, g, g)
            et = SL.EPS(Mt, hermitian=True)
            S.EPSSetDimensions(et.h, 20, 64, SL.PETSC_DETERMINE)
            S.EPSSetTolerances(et.h, TOL, SL.PETSC_CURRENT)
            barrier()
            s0, l0 = syncs(), launches()
            t0 = time.perf_counter()
            et.solve()
            barrier()
            dt = allmax(time.perf_counter() - t0)
            nst = et.bv().counters()[1]
            lat = {"host_syncs_per_lanczos_step": (syncs() - s0) / max(nst, 1), "kernel_launches_per_lanczos_step": (launches() - l0) / max(nst, 1),"workload": "2-D Laplacian 1024x1024 (global, split over the ranks), nev=20 ncv=64: launch/latency-bound regime",
                   "seconds": dt, "restarts": et.its, "nconv": et.nconv, "lanczos_steps": nst, "us_per_lanczos_step": 1e6 * dt / max(nst, 1)}
            lat["max_rel_residual"] = max(et.error(i) for i in range(et.nconv)) if et.nconv else None
            et.destroy()
            Mt.destroy()
        except Exception as ex:                     # noqa: BLE001 — a failed optional leg must not cost the headline line
            lat = {"error": (type(ex).__name__ + ": " + str(ex))[:400]}

    # ---------------- time-to-solution of the north-star problem, strong-scaled, checked before printing ----------------
    tts = None
    if args.tts != "none":
        try:
            tw = TTS[args.tts]
            g, dim = tw["g"], tw["dim"]
            t0 = time.perf_counter()
            Mt = SL.Mat.laplacian(dim, g, g, g)
            et = SL.EPS(Mt, hermitian=True)
            S.EPSSetDimensions(et.h, tw["nev"], tw["ncv"], SL.PETSC_DETERMINE)
            S.EPSSetTolerances(et.h, TOL, SL.PETSC_CURRENT)
            barrier()
            t_build = time.perf_counter() - t0
            hbm_tts = hbm_used() - hbm0
            l0 = launches()
            t0 = time.perf_counter()
            et.solve()
            barrier()
            dt = allmax(time.perf_counter() - t0)
            hbm_tts = max(hbm_tts, hbm_used() - hbm0)
            nconv = et.nconv
            vals = [et.eigenvalue(i)[0] for i in range(nconv)]
            errs = [et.error(i) for i in range(nconv)]
            th = 2 - 2 * np.cos(np.arange(max(1, g - 40), g + 1) * np.pi / (g + 1))       # the top 40 1-D values are enough
            analytic = (th[:, None, None] + th[None, :, None] + th[None, None, :]).ravel()
            dist_an = float(max(np.min(np.abs(analytic - x)) / abs(x) for x in vals)) if vals else None
            nst = et.bv().counters()[1]
            checks = {"nconv_ge_nev": nconv >= tw["nev"], "residuals_lt_5tol": bool(errs) and max(errs) < 5 * TOL,
                      "values_within_1e-10_of_analytic": dist_an is not None and dist_an < 1e-10, "reason_converged": et.reason > 0}
            if not all(checks.values()):
                raise SystemExit(f"bench.py: the {tw['name']} solve on {world} GPU(s) FAILED its checks {checks}: nconv={nconv} "
                                 f"max residual={max(errs) if errs else None} distance to analytic={dist_an}; nothing is printed for a wrong answer")
            tts = {"workload": tw["name"], "n_gpus": world, "seconds": dt, "seconds_build": t_build, "its": et.its, "nconv": nconv,
                   "lanczos_steps": nst, "ms_per_lanczos_step": 1e3 * dt / max(nst, 1), "max_rel_residual": max(errs),
                   "max_rel_dist_to_analytic": dist_an, "values": vals[:tw["nev"]], "checks": checks, "kernel_launches": launches() - l0,
                   "hbm_bytes_per_gpu": hbm_tts,
                   "timed": "EPSSolve wall clock (first start-vector op to convergence), barrier + device sync on both sides, max over ranks; matrix generation reported separately"}
            mark = os.path.join(tempfile.gettempdir(), f"b2k_tts_{args.tts}_n1.json")
            if rank == 0:
                if world == 1:
                    try:
                        json.dump({"seconds": dt, "when": time.time()}, open(mark, "w"))
                    except Exception:
                        pass
                elif os.path.exists(mark):
                    try:
                        t1 = json.load(open(mark))
                        if time.time() - t1["when"] < 6 * 3600:
                            tts["strong_efficiency"] = t1["seconds"] / (world * dt)
                            tts["strong_efficiency_from"] = f"N=1 run of this lease ({t1['seconds']:.2f} s, {mark})"
                    except Exception:
                        pass
            et.destroy()
            Mt.destroy()
        except Exception as ex:                     # noqa: BLE001 — a failed optional leg must not cost the headline line
            tts = {"error": (type(ex).__name__ + ": " + str(ex))[:400]}

    # ---------------- baselines on the same box (rank 0, N=1 only) -------------------------------------------------
    cpu = None
    libgpu = None
    if world == 1 and not args.no_lib:
        libgpu = library_baseline(wl)
        if "ms_per_lanczos_step" in libgpu:
            libgpu["this_build_ms_per_lanczos_step"] = t_ms / max(steps, 1)
            libgpu["speedup_vs_library"] = libgpu["ms_per_lanczos_step"] / (t_ms / max(steps, 1))
    if world == 1 and not args.no_cpu:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
                                "--workload", args.workload], capture_output=True, text=True, timeout=900)
            for l in r.stdout.splitlines():
                if l.startswith("{"):
                    cpu = json.loads(l)["cpu_baseline"]
            if cpu is None:
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + r.stderr[-300:]}
        except Exception as e:                                   # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    parity = None

    def emit():
        if rank != 0:
            return
        halo = "none (1 GPU)"
        if world > 1:
            halo = ("k-vector reductions fused into the reduction kernel over NVLink peer memory (k_reduce_partials_xg); " if D.P2P
                    else "k-vector reductions: ncclAllReduce; ")
            halo += "halo: " + ("pushed over NVLink peer memory (k_halo_push/k_halo_wait)" if (D.P2P and os.environ.get("B2K_HALO_P2P", "1") != "0")
                                else "ncclSend/ncclRecv")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "lanczos_steps": steps, "lanczos_steps_per_s": steps / (t_ms / 1e3), "gs_passes_per_step": gs_passes / max(steps, 1),
            "seconds_per_restart_cycle": t_ms / 1e3 / args.steps, "host_syncs_per_lanczos_step": nsync / max(steps, 1),
            "kernel_launches_per_lanczos_step": nl / max(steps, 1),
            "hbm_bytes_per_gpu": {"matrix": hbm_matrix, "matrix_plus_basis_and_scratch": hbm_solver,
                                  "note": "cudaMemGetInfo deltas; the matrix is stored once (SELL-32 copy + row pointer)"},
            "kernels": kernels, "gs_sweeps_gbs": (gs_bytes / gs_ms / 1e6) if gs_ms else None,
            "roofline": roofline, "cpu_baseline": cpu, "library_gpu_baseline": libgpu, "e2e": e2e, "time_to_solution": tts,
            "latency_leg": lat, "multi_gpu_parity": parity, "gpu_launches": nl, "clocks": clocks, "collectives": halo,
        }
        print(json.dumps(line), flush=True)

    # ---------------- multi-GPU correctness inside the scaling run (last: a watchdog prints the line if a case hangs) ----------
    if world > 1 and not args.no_parity:
        import threading
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import mgpu_cases

        def watchdog():
            nonlocal parity
            parity = {"cases": len(mgpu_cases.CASES), "passed": 0, "failed": ["a row-partitioned correctness case did not finish within 300 s"], "ranks": world}
            emit()
            os._exit(3)
        timer = threading.Timer(300.0, watchdog)
        timer.daemon = True
        timer.start()
        parity = mgpu_cases.run_all(rank, world)
        timer.cancel()
        parity["transport"] = "reductions: NVLink mailboxes" if D.P2P else "reductions: ncclAllReduce"
    emit()
    D.finalize()
    if parity is not None and parity["passed"] != parity["cases"]:
        raise SystemExit(f"bench.py: multi-GPU parity failed: {parity['failed']}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--tts", default="c3", choices=sorted(TTS) + ["none"],
                    help="time-to-solution leg: c3 = 512^3 to convergence (~5 min on 1 GPU, ~45 s on 8), c3small = 128^3, none = skip")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-lib", action="store_true")
    ap.add_argument("--no-tts", action="store_true", help="same as --tts none")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.no_tts:
        args.tts = "none"
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        if args.warmup < 3:
            args.warmup = 3                        # timing rule: at least 3 warm-up steps
        run_b200(args, wl)


if __name__ == "__main__":
    main()
