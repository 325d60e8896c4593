"""ctypes binding of libb200krylov.so (include/b2k.h): the sm_100a kernels behind a plain C ABI.

There is deliberately NO fallback: if the library is missing, or no GPU is visible when a context is
requested, this raises.  (The oracle under oracle/ is test infrastructure and is never imported here.)
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200krylov.so")

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p
c_sz = ctypes.c_size_t

# name -> argtypes (restype is always int unless listed in _SPECIAL)
SIGNATURES = {
    "b2k_device_count": [ctypes.POINTER(c_int)],
    "b2k_ctx_create": [c_int, ctypes.POINTER(c_vp)],
    "b2k_ctx_destroy": [c_vp],
    "b2k_ctx_sync": [c_vp],
    "b2k_ctx_sm_count": [c_vp],
    "b2k_ctx_launches": [c_vp, ctypes.POINTER(c_u64)],
    "b2k_ctx_syncs": [c_vp, ctypes.POINTER(c_u64)],
    "b2k_ctx_copy_bytes": [c_vp, ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)],
    "b2k_prof_enable": [c_vp, c_int],
    "b2k_prof_get": [c_vp, c_int, ctypes.POINTER(c_u64), ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)],
    "b2k_prof_get_flops": [c_vp, c_int, ctypes.POINTER(c_dbl)],
    "b2k_malloc": [c_vp, ctypes.POINTER(c_vp), c_sz],
    "b2k_free": [c_vp, c_vp],
    "b2k_memset0": [c_vp, c_vp, c_sz],
    "b2k_h2d": [c_vp, c_vp, c_vp, c_sz],
    "b2k_d2h": [c_vp, c_vp, c_vp, c_sz],
    "b2k_h2d_async": [c_vp, c_vp, c_vp, c_sz],
    "b2k_d2h_async": [c_vp, c_vp, c_vp, c_sz],
    "b2k_d2d": [c_vp, c_vp, c_vp, c_sz],
    "b2k_host_alloc": [ctypes.POINTER(c_vp), c_sz],
    "b2k_host_free": [c_vp],
    "b2k_mem_info": [c_vp, ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)],
    "b2k_timer_start": [c_vp],
    "b2k_timer_stop_ms": [c_vp, ctypes.POINTER(c_dbl)],
    "b2k_dotvec": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp],
    "b2k_multvec": [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl, c_dbl, c_vp, c_vp],
    "b2k_sumsq": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp],
    "b2k_colabssum": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp],
    "b2k_rowabssum_max": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp],
    "b2k_scale": [c_vp, c_vp, c_i64, c_i64, c_int, c_dbl],
    "b2k_copy": [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_int],
    "b2k_axpby": [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_dbl, c_dbl],
    "b2k_mult": [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_dbl, c_dbl, c_vp, c_int],
    "b2k_mult_inplace": [c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_vp, c_int, c_int],
    "b2k_dot": [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_int],
    "b2k_set_random": [c_vp, c_vp, c_i64, c_i64, c_u64],
    "b2k_fill": [c_vp, c_vp, c_i64, c_dbl],
    "b2k_pointwise_mult": [c_vp, c_vp, c_vp, c_vp, c_i64],
    "b2k_csr_get_diagonal": [c_vp, c_vp, c_i64, c_vp],
    "b2k_gs_dot": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp],
    "b2k_gs_update_dot": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp],
    "b2k_gs_update_norm": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp],
    "b2k_gs_update_norm_gated": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl],
    "b2k_scale_rsqrt": [c_vp, c_vp, c_i64, c_vp],
    "b2k_scale_rsqrt_gated": [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_dbl],
    "b2k_gs_set_fused": [c_int],
    "b2k_spmv_set_sell": [c_int],
    "b2k_vq_set_tma": [c_int],
    "b2k_csr_create": [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)],
    "b2k_csr_create_global": [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_vp)],
    "b2k_csr_adopt": [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)],
    "b2k_csr_destroy": [c_vp, c_vp],
    "b2k_csr_info": [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)],
    "b2k_csr_arrays": [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)],
    "b2k_csr_release_arrays": [c_vp],
    "b2k_csr_transpose_split": [c_vp, c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)],
    "b2k_csr_bytes": [c_vp, ctypes.POINTER(c_i64)],
    "b2k_csr_last_kernel": [c_vp, ctypes.POINTER(c_int)],
    "b2k_spmv_set_pipe_min_chunks": [c_int],
    "b2k_csr_spmm": [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int],
    "b2k_csr_spmv": [c_vp, c_vp, c_vp, c_vp, c_vp],
    "b2k_csr_spmv_shift": [c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl],
    "b2k_csr_laplacian": [c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, ctypes.POINTER(c_vp),
                          ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)],
    "b2k_tsqr_plan": [c_vp, c_i64, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)],
    "b2k_tsqr_forward": [c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp],
    "b2k_tsqr_backward": [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp],
    "b2k_gather": [c_vp, c_vp, c_vp, c_vp, c_i64],
    "b2k_scatter_add": [c_vp, c_vp, c_vp, c_vp, c_i64],
    "b2k_comm_unique_id": [c_vp],
    "b2k_comm_create": [c_vp, c_int, c_int, c_vp, ctypes.POINTER(c_vp)],
    "b2k_comm_destroy": [c_vp],
    "b2k_comm_rank": [c_vp, ctypes.POINTER(c_int), ctypes.POINTER(c_int)],
    "b2k_comm_allreduce_sum": [c_vp, c_vp, c_int],
    "b2k_comm_allreduce_max": [c_vp, c_vp, c_int],
    "b2k_comm_sendrecv": [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_int],
    "b2k_comm_group_start": [c_vp],
    "b2k_comm_group_end": [c_vp],
    "b2k_comm_allgather": [c_vp, c_vp, c_vp, c_i64],
    "b2k_comm_reduce_scatter_sum": [c_vp, c_vp, c_vp, c_i64],
    "b2k_comm_barrier": [c_vp],
    "b2k_comm_p2p_handle": [c_vp, c_vp],
    "b2k_comm_p2p_open": [c_vp, c_vp],
    "b2k_comm_p2p_close": [c_vp],
    "b2k_comm_p2p_resume": [c_vp],
    "b2k_comm_p2p_enabled": [c_vp],
    "b2k_comm_reduce_scope": [c_vp, c_int, ctypes.POINTER(c_int)],
    "b2k_comm_p2p_error": [c_vp, ctypes.POINTER(c_int)],
    "b2k_halo_create": [c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)],
    "b2k_halo_exchange": [c_vp, c_vp, ctypes.POINTER(c_vp)],
    "b2k_halo_error": [c_vp, ctypes.POINTER(c_int)],
    "b2k_halo_destroy": [c_vp],
}
_SPECIAL = {
    "b2k_last_error": ([], ctypes.c_char_p),
    "b2k_version": ([], c_int),
    "b2k_ctx_stream": ([c_vp], c_vp),
}

_lib = None


class B2KError(RuntimeError):
    pass


def load():
    """dlopen libb200krylov.so and attach prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2KError(f"{LIB_PATH} not found: run `make` (or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    for name, (args, res) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise B2KError(f"b2k error {rc}: {load().b2k_last_error().decode()}")


class Context:
    """One b2k_ctx (device + stream + scratch)."""

    def __init__(self, device=0):
        lib = load()
        h = c_vp()
        check(lib.b2k_ctx_create(device, ctypes.byref(h)))
        self.h = h
        self.lib = lib
        self.device = device

    def sync(self):
        check(self.lib.b2k_ctx_sync(self.h))

    @property
    def launches(self):
        n = c_u64()
        check(self.lib.b2k_ctx_launches(self.h, ctypes.byref(n)))
        return n.value

    @property
    def sm_count(self):
        return self.lib.b2k_ctx_sm_count(self.h)

    def close(self):
        if self.h:
            self.lib.b2k_ctx_destroy(self.h)
            self.h = None

    # ---- small helpers used by tests and bench (raw HBM buffers, no torch) ------------------
    def empty(self, nelem, dtype=np.float64):
        return DeviceArray(self, int(nelem), np.dtype(dtype))

    def to_device(self, a):
        a = np.ascontiguousarray(a) if not a.flags.f_contiguous else np.asfortranarray(a)
        d = DeviceArray(self, a.size, a.dtype)
        check(self.lib.b2k_h2d(self.h, d.ptr, a.ctypes.data, a.nbytes))
        return d

    def timer_start(self):
        check(self.lib.b2k_timer_start(self.h))

    def timer_stop_ms(self):
        ms = c_dbl()
        check(self.lib.b2k_timer_stop_ms(self.h, ctypes.byref(ms)))
        return ms.value


class DeviceArray:
    def __init__(self, ctx, nelem, dtype):
        self.ctx, self.nelem, self.dtype = ctx, nelem, dtype
        p = c_vp()
        check(ctx.lib.b2k_malloc(ctx.h, ctypes.byref(p), nelem * dtype.itemsize))
        self.ptr = p.value

    def at(self, elem_offset):
        return self.ptr + int(elem_offset) * self.dtype.itemsize

    def to_host(self, shape=None, order="F"):
        out = np.empty(self.nelem, dtype=self.dtype)
        check(self.ctx.lib.b2k_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out if shape is None else out.reshape(shape, order=order)

    def free(self):
        if self.ptr:
            self.ctx.lib.b2k_free(self.ctx.h, self.ptr)
            self.ptr = None
