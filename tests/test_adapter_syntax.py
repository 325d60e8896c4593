"""The PETSc-typed plug-in sources (adapters/petsc/*.c: BV type "b200", the MatShell operator, the MPI glue) are TYPE-CHECKED
against the reference's own SLEPc headers — include/slepc/private/bvimpl.h with its 35-slot struct _BVOps and struct _p_BV —
using the stand-in PETSc declarations of adapters/petsc/stub/ (no PETSc in this image, so nothing is linked or run).
Needs the reference tree: skipped on the GPU box, where /root/reference is not mounted."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/include"
AD = os.path.join(ROOT, "adapters", "petsc")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("src", ["bvb200_petsc.c", "matb200_petsc.c", "b2k_petsc.c"])
def test_adapter_type_checks_against_reference_headers(src):
    r = subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Wno-unused-function", "-Werror", "-fsyntax-only", "-I" + os.path.join(AD, "stub"),
                        "-I" + REF, "-I" + os.path.join(ROOT, "include"), "-I" + AD, os.path.join(AD, src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_adapter_decides_every_bvops_slot():
    """every slot of the reference's struct _BVOps is assigned in BVCreate_B200 (a function or an explained NULL)"""
    hdr = open(os.path.join(REF, "slepc", "private", "bvimpl.h")).read()
    body = hdr[hdr.index("struct _BVOps {"):]
    body = body[:body.index("};")]
    slots = re.findall(r"\(\*(\w+)\)\(", body)
    assert len(slots) == 35
    src = open(os.path.join(AD, "bvb200_petsc.c")).read()
    missing = [s for s in slots if not re.search(r"bv->ops->%s\s*=" % s, src)]
    assert not missing, missing
