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
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            tails = []
            for q in procs:
                try:
                    tails.append((q.communicate(timeout=10)[0] or "")[-2000:])
                except Exception:                      # noqa: BLE001
                    tails.append("<no output>")
            pytest.fail(f"{case} on {world} ranks hung (240 s); output of the ranks:\n" + "\n----\n".join(tails))
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-3000:] for l in logs)
    return json.load(open(out))


def _verify(case, r):
    sys.path.insert(0, HERE)
    import mgpu_cases
    ok, msg = mgpu_cases.verify(case, r)
    assert ok, msg


@pytest.mark.parametrize("world", [2, 4, 8])
def test_bv_orthonormalize(world, tmp_path):
    _verify("bv", run_case("bv", world, tmp_path))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_eps_laplacian_device_generator(world, tmp_path):
    _verify("lap", run_case("lap", world, tmp_path))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_spmv_slab_halo_closed_form(world, tmp_path):
    """MatMult of the slab-partitioned 192^3 Laplacian element-wise against the closed-form stencil: ghost planes, both halo
    transports (whatever is the default), k_spmv_sell_pipe<true> on the middle ranks"""
    _verify("spmv_slab", run_case("spmv_slab", world, tmp_path))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_reductions_match_nccl(world, tmp_path):
    """k_reduce_partials_xg (sum over the GPUs inside the reduction kernel, NVLink mailboxes) against k_reduce_partials +
    ncclAllReduce on the same data in the same processes: norms and vectors agree to rounding of a k-term sum"""
    r = run_case("bv_transports", world, tmp_path)
    if not r["p2p"]:
        pytest.skip("CUDA IPC rendezvous not available on this box: the NCCL path ran twice")
    _verify("bv_transports", r)


@pytest.mark.parametrize("case,world", [("lap", 2), ("lap", 4), ("markov", 2), ("svd", 2), ("spmv_slab", 2), ("spmv_slab", 4), ("markov", 4)])
def test_halo_transports_agree(case, world, tmp_path):
    """peer-memory halo (b2k_halo.cu: the neighbours' entries of x pushed over NVLink into double-buffered ghost arrays, flags
    instead of ncclSend/ncclRecv; the default) against the NCCL halo (B2K_HALO_P2P=0): contiguous planes (lap, spmv_slab),
    packed general plan (markov), two row layouts (svd)"""
    a = run_case(case, world, tmp_path, {"B2K_HALO_P2P": "1"}, "_p2phalo")
    b = run_case(case, world, tmp_path, {"B2K_HALO_P2P": "0"}, "_ncclhalo")
    if not a["p2p"]:
        pytest.skip("CUDA IPC rendezvous not available on this box")
    _verify(case, a)
    _verify(case, b)
    if case == "spmv_slab":
        return
    key = "sigma" if case == "svd" else "lam"
    assert a["nconv"] == b["nconv"]
    assert np.allclose(a[key][:a["nconv"]], b[key][:b["nconv"]], rtol=1e-12, atol=0)


@pytest.mark.parametrize("world", [2, 4])
def test_eps_markov_general_halo(world, tmp_path):
    _verify("markov", run_case("markov", world, tmp_path))


@pytest.mark.parametrize("world", [2, 4])
def test_svd_two_row_layouts(world, tmp_path):
    _verify("svd", run_case("svd", world, tmp_path))
