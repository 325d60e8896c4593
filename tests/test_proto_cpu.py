"""CPU models of the two peer-memory protocols of the multi-GPU path: threads stand in for the GPUs, shared memory for the
NVLink-mapped mailboxes, C11 atomics for the flags (tests/proto/*.c).  They check the protocol rules themselves — buffer
parities, acknowledgement distance, absence of deadlock — for up to 8 ranks and free-running (unsynchronised) ranks, which
the 2-GPU hardware runs cannot cover.  The CUDA kernels are k_reduce_partials_xg (b2k_bv.cu) and k_halo_push / k_halo_wait
(b2k_halo.cu)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def bins(tmp_path_factory):
    out = tmp_path_factory.mktemp("proto")
    for name in ("halo_proto", "reduce_proto"):
        subprocess.run(["gcc", "-O2", "-pthread", os.path.join(HERE, "proto", name + ".c"), "-o", str(out / name)], check=True, capture_output=True)
    return out


@pytest.mark.parametrize("ranks,n", [(2, 3000), (3, 3000), (8, 1500)])
def test_reduction_mailbox_protocol(bins, ranks, n):
    r = subprocess.run([str(bins / "reduce_proto"), str(ranks), str(n)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("ranks,n,seed,barrier_every", [(2, 2000, 1, 0), (4, 2000, 2, 0), (8, 1000, 3, 0), (8, 1000, 4, 7), (3, 2000, 5, 1)])
def test_halo_push_protocol(bins, ranks, n, seed, barrier_every):
    r = subprocess.run([str(bins / "halo_proto"), str(ranks), str(n), str(seed), str(barrier_every)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
