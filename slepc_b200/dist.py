"""One process per GPU: rendezvous of the row-partition communicator (include/b2k.h b2k_comm_*, NCCL over NVLink) and
the halo plan of a row-partitioned Mat "b200csr".  torch.distributed is used ONLY as the launcher-side plumbing — to
broadcast the NCCL unique id and to all-gather the (python) lists of ghost columns once at set-up; every collective of the
solve itself is issued from C on the compute stream (slepc_b200/host/sys.c → b2k_comm.cu)."""
import ctypes
import os

import numpy as np

from . import _b2k
from . import slepc as SL
from .slepc import S


P2P = False      # set by init(): True when the Gram-Schmidt reductions run over peer-memory mailboxes instead of NCCL


def env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def init(backend="nccl"):
    """B2KInitialize on LOCAL_RANK's GPU and, when WORLD_SIZE > 1, the NCCL communicator over all ranks.
    Returns (rank, world)."""
    import torch
    rank, world, local = env()
    torch.cuda.set_device(local)
    lib = _b2k.load()
    SL.initialize(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        idbuf = (ctypes.c_char * 128)()
        if rank == 0:
            _b2k.check(lib.b2k_comm_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        S.B2KCommInitNCCL(rank, world, ctypes.c_char_p(bytes(t.cpu().numpy().tobytes())))
        if world <= 8 and os.environ.get("B2K_COMM_P2P", "1") != "0":
            global P2P
            P2P = enable_p2p(rank, world)
    return rank, world


def enable_p2p(rank, world):
    """Peer-memory (NVLink) mailboxes for the k-vector reductions of the Gram-Schmidt sweeps: every rank exports a CUDA IPC
    handle of its mailbox, the handles are all-gathered here, every rank maps its peers (include/b2k.h b2k_comm_p2p_*).
    All ranks take the same decision (a rank that cannot export or map vetoes it for everyone): NCCL stays the fallback."""
    import torch
    import torch.distributed as dist
    h = (ctypes.c_char * 64)()
    ok = 1
    try:
        S.B2KCommP2PHandle(h)
    except Exception:
        ok = 0
    t = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).cuda()
    allh = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allh, t)
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return False
    blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
    try:
        S.B2KCommP2POpen(ctypes.c_char_p(blob))
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        S.B2KCommDisableP2P()
        return False
    return True


def finalize():
    import torch.distributed as dist
    _, world, _ = env()
    if world > 1 and dist.is_initialized():
        dist.barrier()
        S.B2KCommReset()
        dist.destroy_process_group()


def split(N, world, rank):
    """PetscLayout-style contiguous ownership range of `rank` (bvbasic.c:129-134)"""
    base, rem = divmod(N, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def setup_halo(A, ncols_global):
    """Build and install the halo plan of a row-partitioned b200csr matrix whose columns (= the vector it multiplies) are
    split like `split(ncols_global, world, rank)`: every rank learns which of its entries each peer needs
    (what PETSc's MatSetUpMultiply_MPIAIJ / VecScatterCreate do).  Collective."""
    import torch.distributed as dist
    rank, world, _ = env()
    if world == 1:
        return
    ng = ctypes.c_int()
    gp = ctypes.c_void_p()
    S.MatB200CSRGetGhosts(A.h, ctypes.byref(ng), ctypes.byref(gp))
    ghosts = np.ctypeslib.as_array(ctypes.cast(gp, ctypes.POINTER(ctypes.c_int)), shape=(ng.value,)).copy() if ng.value else np.empty(0, np.int32)
    starts = np.array([split(ncols_global, world, r)[0] for r in range(world)] + [ncols_global])
    owner = np.searchsorted(starts, ghosts, side="right") - 1
    needed = {int(p): np.ascontiguousarray(ghosts[owner == p], dtype=np.int32) for p in np.unique(owner)}
    allneeded = [None] * world
    dist.all_gather_object(allneeded, needed)
    c0 = int(starts[rank])
    rr = sorted(needed)
    rc = [len(needed[p]) for p in rr]
    sr, sc, si = [], [], [np.empty(0, np.int32)]
    for p in range(world):
        g = allneeded[p].get(rank)
        if p != rank and g is not None and len(g):
            sr.append(p)
            sc.append(len(g))
            si.append(np.asarray(g, dtype=np.int64) - c0)
    si = np.concatenate(si)
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    rr, rc, sr, sc, si = i32(rr), i32(rc), i32(sr), i32(sc), i32(si)
    pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    S.MatB200CSRSetHalo(A.h, len(rr), pp(rr), pp(rc), len(sr), pp(sr), pp(sc), pp(si))
