"""CPU-side checks of the drop-in boundary: both shared libraries load without a GPU, export every symbol
that include/b2k.h and include/b2kslepc.h declare, and the product path fails LOUDLY (no CPU fallback)
when no CUDA device is present.  No compute calls here."""
import ctypes
import os
import subprocess

import pytest

from slepc_b200 import _b2k
from slepc_b200 import slepc as SL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernel_library_exports_every_declared_symbol():
    lib = _b2k.load()
    protos = SL.parse_header(os.path.join(ROOT, "include", "b2k.h"))
    assert len(protos) > 55
    missing = [n for n in protos if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes table used by the tests covers the same set
    declared = set(protos)
    bound = set(_b2k.SIGNATURES) | set(_b2k._SPECIAL)
    assert declared == bound, (declared - bound, bound - declared)


def test_host_library_exports_every_declared_symbol():
    lib = SL.load()
    protos = SL.parse_header(os.path.join(ROOT, "include", "b2kslepc.h"))
    assert len(protos) > 140
    missing = [n for n in protos if not hasattr(lib, n)]
    assert not missing, missing


def test_kernels_are_sm100a_sass():
    """the shipped library carries sm_100a code (and the TMA bulk-copy kernel carries UBLKCP)"""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _b2k.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run([cuobjdump, "-sass", _b2k.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    assert "UTMALDG" in sass          # cp.async.bulk.tensor.2d of the tensor-map pipeline (k_gs_tma)


def _gpu_present():
    lib = _b2k.load()
    n = ctypes.c_int()
    lib.b2k_device_count(ctypes.byref(n))
    return n.value > 0


def test_no_cpu_fallback_without_gpu():
    if _gpu_present():
        pytest.skip("a GPU is present")
    with pytest.raises(_b2k.B2KError, match="no CUDA device"):
        _b2k.Context(0)
    with pytest.raises(SL.SlepcError):
        SL.initialize(0)
    # the product's BV and Mat types refuse to exist without a device context
    with pytest.raises(SL.SlepcError, match="no CPU fallback"):
        SL.BV.create(10, 3, bvtype="b200")
    with pytest.raises(SL.SlepcError, match="no CPU fallback"):
        SL.Mat.laplacian(1, 10)


def test_product_does_not_reference_the_oracle():
    """nothing under slepc_b200/ imports, links or dlopens oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "slepc_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboraclecpu" not in txt and "oracle_cpu" not in txt, f
                assert "import oracle" not in txt and "from oracle" not in txt, f
    ldd = subprocess.run(["ldd", SL.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
