"""ORACLE / CPU BASELINE loader — TEST INFRASTRUCTURE, NOT PRODUCT.

Loads oracle/_build/liboraclecpu.so (oracle_cpu.c: host-memory BV type "oraclecpu" + host CSR Mat
"oraclecsr", OpenMP + host BLAS) and registers it with the C host driver through the public plugin
entry BVRegister().  Used by tests/ (host logic on CPU, world_size-2 gloo), __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs only.
"""
import ctypes
import os

import numpy as np

from slepc_b200 import slepc as SL
from slepc_b200.slepc import S, c_int, c_vp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboraclecpu.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    SL.load()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `make` (or __graft_entry__.build())")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    lib.OracleCPURegister.restype = c_int
    lib.OracleCPUGetMaxThreads.restype = c_int
    lib.OracleCPUSetThreads.argtypes = [c_int]
    lib.OracleCPUStreamTriad.restype = ctypes.c_double
    lib.OracleCPUStreamTriad.argtypes = [ctypes.c_int64, c_int]
    lib.MatCreateOracleCSR.restype = c_int
    lib.MatCreateOracleCSR.argtypes = [c_int] * 6 + [c_vp] * 3 + [c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]
    lib.MatCreateOracleLaplacian.restype = c_int
    lib.MatCreateOracleLaplacian.argtypes = [c_int, c_int, c_int, c_int, c_vp]
    SL.chk(lib.OracleCPURegister())
    _lib = lib
    return lib


def threads():
    return load().OracleCPUGetMaxThreads()


def set_threads(n):
    """OpenMP threads of the CPU plug-in (torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU baseline must not inherit it)"""
    load().OracleCPUSetThreads(int(n))


def stream_triad_gbs(n=1 << 26, reps=5):
    return load().OracleCPUStreamTriad(n, reps)


def split_rows(N, size):
    """PetscLayout-style contiguous split: list of (start, end) per rank"""
    base, rem = divmod(N, size)
    out, s = [], 0
    for r in range(size):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def halo_plan(A_loc, cranges, rank):
    """For the local rows A_loc (scipy CSR, GLOBAL column indices) and the column ownership ranges
    `cranges` = [(start,end)] per rank: ghosts (sorted global indices outside the own range), the local
    column numbering [owned | ghosts] and, per peer in increasing rank order, what to receive.
    Returns (colidx_local, ghosts, recvrank, recvcount, needed_from) where needed_from[p] are the GLOBAL
    indices this rank needs from peer p (to be exchanged so that p learns what to send)."""
    cs, ce = cranges[rank]
    ci = A_loc.indices.astype(np.int64)
    off = (ci < cs) | (ci >= ce)
    ghosts = np.unique(ci[off])
    loc = np.where(off, (ce - cs) + np.searchsorted(ghosts, ci), ci - cs).astype(np.int32)
    starts = np.array([c[0] for c in cranges] + [cranges[-1][1]])
    owner = np.searchsorted(starts, ghosts, side="right") - 1
    recvrank, recvcount, needed = [], [], {}
    for p in np.unique(owner):
        g = ghosts[owner == p]
        recvrank.append(int(p))
        recvcount.append(len(g))
        needed[int(p)] = g
    return loc, ghosts, recvrank, recvcount, needed


def mat_csr(A, rank=0, size=1, cranges=None, exchange=None):
    """host "oraclecsr" Mat for rows of the scipy matrix A owned by `rank` (PetscLayout split).
    `exchange(obj)` must all-gather a python object over the ranks (world_size>1 only)."""
    load()
    A = A.tocsr()
    A.sort_indices()
    M, N = A.shape
    rranges = split_rows(M, size)
    cranges = cranges or split_rows(N, size)
    rs, re = rranges[rank]
    cs, ce = cranges[rank]
    loc = A[rs:re]
    rp = np.ascontiguousarray(loc.indptr, dtype=np.int32)
    val = np.ascontiguousarray(loc.data, dtype=np.float64)
    if size == 1:
        cl = np.ascontiguousarray(loc.indices, dtype=np.int32)
        nghost, rr, rc, sr, sc, si = 0, [], [], [], [], []
    else:
        cl, ghosts, rr, rc, needed = halo_plan(loc, cranges, rank)
        allneeded = exchange(needed)                       # list over ranks of {peer: global indices}
        sr, sc, si = [], [], []
        for p in range(size):
            g = allneeded[p].get(rank)
            if p != rank and g is not None and len(g):
                sr.append(p)
                sc.append(len(g))
                si.extend((np.asarray(g) - cs).tolist())
        nghost = len(ghosts)
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    rr, rc, sr, sc, si = i32(rr), i32(rc), i32(sr), i32(sc), i32(si)
    m = SL.Mat()
    p = lambda a: a.ctypes.data_as(c_vp)
    SL.chk(_lib.MatCreateOracleCSR(M, N, rs, re, cs, ce, p(rp), p(cl), p(val), nghost, len(rr), p(rr), p(rc), len(sr), p(sr), p(sc), p(si),
                                   m.ref))
    return m


def mat_laplacian(dim, nx, ny=1, nz=1):
    """host CSR Laplacian built in C (single rank) — the CPU baseline's operator"""
    lib = load()
    m = SL.Mat()
    SL.chk(lib.MatCreateOracleLaplacian(dim, nx, ny, nz, m.ref))
    return m


def bv_cpu(n, m, N=None, rstart=0):
    load()
    return SL.BV.create(n, m, N=N, bvtype="oraclecpu", mem=SL.MEM_HOST, rstart=rstart)


def use_cpu_bv(solver):
    """make an EPS/SVD object use the CPU BV type (the reference's `-bv_type` switch)"""
    load()
    if isinstance(solver, SL.EPS):
        S.BVSetType(solver.bv().h, b"oraclecpu")
    else:
        v, u = solver.bvs()
        S.BVSetType(v.h, b"oraclecpu")
        S.BVSetType(u.h, b"oraclecpu")
