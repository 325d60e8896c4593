"""N>1 path on CPU: world_size-2 (and 3) gloo runs of the C host driver over the callback communicator, rows of the
operator and of the basis split across ranks.  The arithmetic is the oracle's host BV/Mat plug-in; what is
under test is the distributed host logic the GPU path shares: globally reduced dot/norm results, local restarts,
the halo plan, redundant DS solves giving identical decisions on every rank."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_case(case, world, tmp_path):
    out = tmp_path / f"{case}_{world}.json"
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1", GLOO_SOCKET_IFNAME="lo")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), case, str(out)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return json.load(open(out))


@pytest.mark.parametrize("world", [2, 3])
def test_bv_orthonormalize_split_rows(world, tmp_path):
    r = run_case("bv", world, tmp_path)
    assert r["orth"] < 1e-13 and r["dq"] < 1e-12 and r["dn"] < 1e-12


@pytest.mark.parametrize("world", [2, 3])
def test_bv_test10_split_reductions_split_rows(world, tmp_path):
    """bv/tests/test10.c on 2 ranks (its own test list) and on 3: the reference prints `0.` for the difference"""
    r = run_case("bvsplit", world, tmp_path)
    assert r["diff"] < 1e-14 and r["err"] < 1e-13, r


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["bvchol", "bvsvqb", "bvtsqr", "bvtsqrchol"])
def test_bv_block_orthogonalize_split_rows(case, world, tmp_path):
    if world == 3 and case in ("bvchol", "bvsvqb"):
        pytest.skip("covered on 2 ranks")
    r = run_case(case, world, tmp_path)
    assert r["orth"] < 1e-13 and r["dq"] < 1e-12 and r["resid"] < 1e-12


@pytest.mark.parametrize("world", [2, 3])
def test_eps_hep_split_rows(world, tmp_path):
    r = run_case("hep", world, tmp_path)
    assert r["nconv"] == r["ref_nconv"] >= 4
    lam = np.array([x[0] for x in r["lam"]])
    assert np.allclose(np.sort(lam), np.sort(r["ref"]), rtol=1e-10, atol=0)
    assert max(r["errs"][:4]) < 5e-8 and max(r["resid"]) < 5e-8 * 8


def test_eps_nhep_split_rows(tmp_path):
    r = run_case("nhep", 2, tmp_path)
    assert r["nconv"] == r["ref_nconv"] >= 4
    lam = np.array([x[0] for x in r["lam"]])
    assert np.allclose(lam[:4], r["ref"][:4], rtol=1e-9, atol=0)
    assert abs(lam[0] - 1.0) < 1e-9


@pytest.mark.parametrize("case", ["svd", "svd_oneside"])
def test_svd_split_rows(case, tmp_path):
    r = run_case(case, 2, tmp_path)
    assert r["nconv"] == r["ref_nconv"] >= 4
    assert np.allclose(r["sigma"][:4], r["ref"][:4], rtol=1e-10, atol=0)
    assert max(r["errs"][:4]) < 5e-8
