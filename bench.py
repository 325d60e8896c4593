#!/usr/bin/env python
"""bench.py — the hot path of EPSSolve (Krylov-Schur: SpMV + fused DGKS Gram-Schmidt + in-place restart) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N … bench.py --gpus N …      (N > 1)

A "step" is ONE Krylov-Schur restart cycle (krylovschur.c:248-332): ~ncv-l Lanczos expansion steps (MatMult + DGKS
orthonormalisation), the projected solve on the host and the restart V <- V Q.
Workload c2 (default, BASELINE.json configs[1]): 2-D 5-point Laplacian 4096x4096 per GPU (weak scaling: the grid is
(4096 N) x 4096, slab-partitioned), symmetric Krylov-Schur nev=20 ncv=64, tol 1e-8, synthetic.
Workload c3 (configs[2], strong scaling): 7-point Laplacian 512^3 split over the N GPUs, nev=10 (ncv=25).
metric: rows x Lanczos-steps per second over the whole job (extensive, like tokens/s); seconds per restart cycle,
per-kernel achieved GB/s (CUDA events around every launch, live in the timed region) and the CPU baseline ride along.
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
METRIC = "EPSSolve Krylov-Schur hot-path throughput (matrix rows x Lanczos steps per second; s/restart-cycle and per-kernel GB/s in extra keys)"
UNIT = "row-steps/s"
KCLASS = ["dotvec", "multvec", "gs_fused", "spmv", "gemm_restart", "elementwise"]


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
def run_reference(args, wl):
    """CPU arm: the reference's CPU path as restated in oracle/ (same C host driver, host-memory BV type with BLAS
    gemv/gemm + OpenMP CSR SpMV), on the box's host cores.  The unmodified SLEPc needs PETSc+MPI, absent here."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
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
    K = max(1, min(args.steps, 4))
    W = max(0, min(args.warmup, 1))
    t0 = time.time()
    if wl["dim"] == 2:
        nx, ny, nz = wl["nx"], wl["ny"], 1
    else:
        nx, ny, nz = 128, wl["ny"], wl["nz"]       # bounded sample: a 128-plane slab of the 512^3 grid
    M = CP.mat_laplacian(wl["dim"], nx, ny, nz)
    rows = nx * ny * nz
    eps = SL.EPS(M, hermitian=True)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, wl["nev"], wl["ncv"], SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, 1e-8, 1000000)
    eps.cycles(1)                                   # first (unrestarted) cycle: page-faults the basis, never timed
    if W:
        eps.cycles(W)
    bv = eps.bv()
    m0 = bv.counters()[1]
    t1 = time.time()
    done = eps.cycles(K)
    dt = time.time() - t1
    steps = bv.counters()[1] - m0
    value = rows * steps / dt
    triad = CP.stream_triad_gbs(1 << 26, 3)
    sample = (f"{done} restart cycles ({steps} Lanczos steps) at {nx}x{ny}x{nz} = {rows} rows after {1 + W} untimed cycles; "
              f"steps/warmup requested {args.steps}/{args.warmup}, capped to {K}/{W} to bound CPU time; setup {t1 - t0:.1f}s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": 1 + W,
        "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "l2": "inputs (>= 1 GB basis + matrix) far larger than cache"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "stream_triad_gbs": triad, "lanczos_steps_per_s": steps / dt,
                         "note": "restated reference (oracle/oracle_cpu.c under the same C host driver); SLEPc+PETSc cannot be built here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nx_total = wl["nx"] * (world if wl["scaling"] == "weak" else 1)
    rows_global = nx_total * wl["ny"] * (wl["nz"] if wl["dim"] == 3 else 1)

    # ---------------- device-resident leg: K restart cycles, inputs already in HBM -----------------------------
    M = SL.Mat.laplacian(wl["dim"], nx_total, wl["ny"], wl["nz"])
    eps = SL.EPS(M, hermitian=True)
    S.EPSSetDimensions(eps.h, wl["nev"], wl["ncv"], SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, 1e-8, 100000000)
    eps.cycles(args.warmup)
    bv = eps.bv()

    def launches():
        n = ctypes.c_uint64()
        lib.b2k_ctx_launches(ctx, ctypes.byref(n))
        return n.value

    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    _b2k.check(lib.b2k_prof_enable(ctx, 1))
    m0, l0, g0 = bv.counters()[1], launches(), bv.counters()[0]
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
    t_ms = allmax(ms.value)
    prof = {}
    for cid, name in enumerate(KCLASS):
        n_, ms_, b_ = ctypes.c_uint64(), ctypes.c_double(), ctypes.c_double()
        _b2k.check(lib.b2k_prof_get(ctx, cid, ctypes.byref(n_), ctypes.byref(ms_), ctypes.byref(b_)))
        prof[name] = dict(launches=n_.value, ms=ms_.value, bytes=b_.value)
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
        h = build_host_csr_slab(lib, wl, nx_total, rank, world)
        v0 = np.empty(h["nloc"])
        from slepc_b200.slepc import _set_vec_rstart
        v0[:] = np.sin(0.37 * np.arange(h["row0"], h["row0"] + h["nloc"]) + 0.1) + 0.5
        out = np.empty(h["nloc"])
        hb0, db0 = ctypes.c_uint64(), ctypes.c_uint64()
        lib.b2k_ctx_copy_bytes(ctx, ctypes.byref(hb0), ctypes.byref(db0))
        barrier()
        t0 = time.perf_counter()
        A = SL.Mat()
        S.MatCreateB200CSR(h["N"], h["N"], h["row0"], h["row0"] + h["nloc"], h["rowptr"][1], h["colidx"][1], h["val"][1],
                           h["row0"], h["row0"] + h["nloc"], A.ref)
        if world > 1:
            hl = h["halo"]
            i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
            rr, rc, sr, sc, si = i32(hl["recvrank"]), i32(hl["recvcount"]), i32(hl["sendrank"]), i32(hl["sendcount"]), i32(hl["sendidx"])
            pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            S.MatB200CSRSetHalo(A.h, len(rr), pp(rr), pp(rc), len(sr), pp(sr), pp(sc), pp(si))
        e2 = SL.EPS(A, hermitian=True)
        S.EPSSetDimensions(e2.h, wl["nev"], wl["ncv"], SL.PETSC_DETERMINE)
        S.EPSSetTolerances(e2.h, 1e-8, 100000000)
        x0, _ = A.create_vecs()
        x0.set_values(v0)
        S.EPSSetInitialSpace(e2.h, 1, (ctypes.c_void_p * 1)(x0.h))
        ke = 0
        for _ in range(args.steps):
            ke += e2.cycles(1)
        bv2 = e2.bv()
        steps2 = bv2.counters()[1]
        checksum = 0.0
        for j in range(wl["nev"]):                 # device→host read of the result: the nev leading Ritz vectors
            S.BVGetColumnHost(bv2.h, j, out.ctypes.data_as(ctypes.c_void_p))
            checksum += float(out[0])
        barrier()
        dt = allmax(time.perf_counter() - t0)
        hb1, db1 = ctypes.c_uint64(), ctypes.c_uint64()
        lib.b2k_ctx_copy_bytes(ctx, ctypes.byref(hb1), ctypes.byref(db1))
        e2e = {"value": rows_global * steps2 / dt, "unit": UNIT, "h2d_bytes_per_step": (hb1.value - hb0.value) / max(ke, 1),
               "d2h_bytes_per_step": (db1.value - db0.value) / max(ke, 1), "seconds": dt, "steps": ke, "lanczos_steps": steps2,
               "what": "MatCreateB200CSR(host CSR, pinned) + EPSSetInitialSpace(host vector) + K restart cycles (incl. the first, "
                       "unrestarted one) + BVGetColumnHost of the nev leading Ritz vectors; wall clock, max over ranks"}
        e2.destroy()
        A.destroy()

    # ---------------- time-to-solution of a complete solve (BASELINE.json's other metric), reduced grid ----------------
    tts = None
    if not args.no_tts:
        g = 1024                                   # the SAME global problem for every N (rows split over the ranks): a known, bounded solve
        Mt = SL.Mat.laplacian(2, g, g) if wl["dim"] == 2 else SL.Mat.laplacian(3, 128, 128, 128)
        et = SL.EPS(Mt, hermitian=True)
        S.EPSSetDimensions(et.h, wl["nev"], wl["ncv"], SL.PETSC_DETERMINE)
        S.EPSSetTolerances(et.h, 1e-8, SL.PETSC_CURRENT)
        barrier()
        t0 = time.perf_counter()
        et.solve()
        barrier()
        dt = allmax(time.perf_counter() - t0)
        tts = {"workload": "2-D Laplacian 1024x1024 (global, split over the ranks)" if wl["dim"] == 2 else "3-D Laplacian 128^3 (global, split over the ranks)",
               "nev": wl["nev"], "ncv": wl["ncv"], "seconds": dt, "restarts": et.its, "nconv": et.nconv,
               "max_rel_residual": max(et.error(i) for i in range(et.nconv)) if et.nconv else None,
               "full_size": "profiles/r01_tts_1gpu.jsonl, r01_tts_8gpu.jsonl: C2 4096x4096 converges 20 pairs in 455.8 s (4171 restarts, 116955 MatMults) on one B200; "
                            "C3 512^3 in 318.2 s on one and 43.1 s on eight B200"}
        et.destroy()
        Mt.destroy()

    # ---------------- CPU baseline on the same box (rank 0, N=1 only) --------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                "--workload", args.workload], capture_output=True, text=True, timeout=900)
            for l in r.stdout.splitlines():
                if l.startswith("{"):
                    cpu = json.loads(l)["cpu_baseline"]
            if cpu is None:
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + r.stderr[-300:]}
        except Exception as e:                                   # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "rows_global": rows_global, "rows_per_gpu": rows_global // world,
                       "l2": "inputs larger than L2 (basis 8.7 GB + matrix 1.1 GB per GPU vs 126 MB L2)", "step": "one Krylov-Schur restart cycle"},
            "lanczos_steps": steps, "lanczos_steps_per_s": steps / (t_ms / 1e3), "gs_passes_per_step": gs_passes / max(steps, 1),
            "seconds_per_restart_cycle": t_ms / 1e3 / args.steps,
            "kernels": kernels, "gs_sweeps_gbs": (gs_bytes / gs_ms / 1e6) if gs_ms else None,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "time_to_solution": tts, "gpu_launches": nl, "clocks": clocks,
            "collectives": ("none (1 GPU)" if world == 1 else
                            ("k-vector reductions fused into the reduction kernel over NVLink peer memory (k_reduce_partials_xg); halo: ncclSend/Recv"
                             if D.P2P else "k-vector reductions: ncclAllReduce; halo: ncclSend/Recv")),
        }
        print(json.dumps(line), flush=True)
    D.finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tts", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        if args.warmup < 3:
            args.warmup = 3                        # timing rule: at least 3 warm-up steps
        run_b200(args, wl)


if __name__ == "__main__":
    main()
