"""Row-partitioned solves on 2 (or more) B200s over NCCL: needs >= 2 visible GPUs, skipped otherwise
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).  The same host logic is covered on CPU by
tests/test_dist_cpu.py (gloo)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def run_case(case, world, tmp_path, extra_env=None, tag=""):
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = tmp_path / f"{case}_{world}{tag}.json"
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "mgpu_worker.py"), case, str(out)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-3000:] for l in logs)
    return json.load(open(out))


@pytest.mark.parametrize("world", [2, 4])
def test_bv_orthonormalize(world, tmp_path):
    r = run_case("bv", world, tmp_path)
    assert r["orth"] < 1e-13 and r["span"] < 1e-10 and r["dn"] < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_eps_laplacian_device_generator(world, tmp_path):
    r = run_case("lap", world, tmp_path)
    assert r["nconv"] >= 6
    an = np.array(r["analytic"])
    for x in r["lam"][:6]:
        assert np.min(np.abs(an - x)) < 1e-10 * abs(x)
    assert max(r["errs"][:6]) < 5e-8


@pytest.mark.parametrize("world", [2, 4])
def test_peer_memory_reductions_match_nccl(world, tmp_path):
    """k_reduce_partials_xg (sum over the GPUs inside the reduction kernel, NVLink mailboxes) against k_reduce_partials +
    ncclAllReduce on the same data: both are fixed-order sums, so norms and vectors agree to rounding of a k-term sum"""
    a = run_case("bv", world, tmp_path, {"B2K_COMM_P2P": "1"}, "_p2p")
    b = run_case("bv", world, tmp_path, {"B2K_COMM_P2P": "0"}, "_nccl")
    assert not b["p2p"]
    if not a["p2p"]:
        pytest.skip("CUDA IPC rendezvous not available on this box: the NCCL path ran twice")
    assert np.allclose(a["norms"], b["norms"], rtol=1e-14, atol=0)
    assert np.allclose(a["q_checksum"], b["q_checksum"], rtol=0, atol=1e-14)
    assert a["orth"] < 1e-13 and b["orth"] < 1e-13


@pytest.mark.skipif(os.environ.get("B2K_TEST_EXPERIMENTAL") != "1", reason="opt-in kernel, not measured yet: set B2K_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("case,world", [("lap", 2), ("lap", 4), ("markov", 2), ("svd", 2)])
def test_peer_memory_halo_matches_nccl_halo(case, world, tmp_path):
    """B2K_HALO_P2P=1 (b2k_halo.cu: the neighbours' entries of x pushed over NVLink into double-buffered ghost arrays, flags instead
    of ncclSend/ncclRecv) against the default halo: contiguous planes (lap), packed general plan (markov), two row layouts (svd)"""
    a = run_case(case, world, tmp_path, {"B2K_HALO_P2P": "1"}, "_p2phalo")
    b = run_case(case, world, tmp_path, {"B2K_HALO_P2P": "0"}, "_ncclhalo")
    if not a["p2p"]:
        pytest.skip("CUDA IPC rendezvous not available on this box")
    key = "sigma" if case == "svd" else "lam"
    assert a["nconv"] == b["nconv"]
    assert np.allclose(a[key][:a["nconv"]], b[key][:b["nconv"]], rtol=1e-12, atol=0)
    assert max(a["errs"][:4]) < 5e-8


def test_eps_markov_general_halo(tmp_path):
    r = run_case("markov", 2, tmp_path)
    assert r["nconv"] >= 4
    assert np.allclose(r["lam"][:4], r["ref"][:4], rtol=1e-9, atol=0)
    assert max(r["errs"][:4]) < 5e-8


def test_svd_two_row_layouts(tmp_path):
    r = run_case("svd", 2, tmp_path)
    assert r["nconv"] >= 5
    assert np.allclose(r["sigma"][:5], r["ref"][:5], rtol=1e-10, atol=0)
    assert max(r["errs"][:5]) < 5e-8
    assert r["nconv_impl"] >= 5 and np.allclose(r["sigma_impl"][:5], r["ref"][:5], rtol=1e-10, atol=0)
    assert max(r["errs_impl"][:5]) < 5e-8
