"""slepc_b200 — B200-native Krylov hot path of SLEPc (Krylov-Schur / thick-restart Lanczos expansion,
Gram-Schmidt and restart) behind SLEPc's BV / Mat plugin API.

  csrc/   sm_100a CUDA kernels + the C ABI of include/b2k.h            → lib/libb200krylov.so
  host/   C host side mirroring SLEPc's BV/DS/ST/EPS/SVD interface     → lib/libb2kslepc.so
  _b2k.py, slepc.py   ctypes bindings used by tests/, bench.py and __graft_entry__.py
"""
