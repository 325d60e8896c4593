"""Kernel variants selected by an environment switch, each run in a subprocess with the switch both ways:
  * B2K_VQ_NARROW (default 1) — k_vq_tma<.,.,32,6>: 32-column TMA boxes in a 6-stage ring and a <4,1> warp tiling for restart
    GEMMs with kin, nout <= 32 (C3: 25 -> 13 columns); 0 = the 64-column boxes.  Measured on B200 (round 2): 5410 vs 3189 GB/s."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

VQ_NARROW = r'''
import sys, time, ctypes, numpy as np
sys.path.insert(0, %r)
from slepc_b200 import _b2k
from slepc_b200._b2k import check
ctx = _b2k.Context(0)
rng = np.random.default_rng(5)
for (n, k, s, e) in [(4099, 25, 0, 13), (100003, 32, 0, 16), (70000, 32, 3, 32), (4096, 8, 0, 8), (50000, 17, 5, 9), (1 << 22, 25, 0, 13)]:
    ld = n + (n %% 2)
    H = np.asfortranarray(rng.standard_normal((ld, k))); H[n:] = 0.0      # column-major, as the basis lies in HBM
    Q = np.asfortranarray(rng.standard_normal((k, k)))
    dV, dQ = ctx.to_device(H), ctx.to_device(Q)
    check(ctx.lib.b2k_mult_inplace(ctx.h, dV.ptr, ld, n, k, s, e, dQ.ptr, k, 0))
    G = dV.to_host((ld, k))
    ref = H.copy(); ref[:n, s:e] = H[:n] @ Q[:, s:e]
    assert np.allclose(G[:n], ref[:n], rtol=1e-12, atol=1e-12), (n, k, s, e, np.abs(G[:n] - ref[:n]).max())
    assert np.array_equal(G[n:], H[n:])
    if n >= 1 << 22:
        ctx.sync(); ctx.timer_start()
        for _ in range(10):
            check(ctx.lib.b2k_mult_inplace(ctx.h, dV.ptr, ld, n, k, s, e, dQ.ptr, k, 0))
        ms = ctx.timer_stop_ms() / 10
        print("k_vq_tma narrow=%%s n=%%d %%d->%%d: %%.3f ms, %%.0f GB/s" %% (__import__("os").environ.get("B2K_VQ_NARROW"), n, k, e - s, ms, 8.0 * n * (k + e - s) / ms / 1e6))
print("ok")
''' % ROOT


@pytest.mark.parametrize("narrow", ["1", "0"])
def test_vq_tma_narrow_boxes(narrow):
    r = subprocess.run([sys.executable, "-c", VQ_NARROW], env=dict(os.environ, B2K_VQ_NARROW=narrow), capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
