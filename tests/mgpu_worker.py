"""Worker of tests/test_multi_gpu.py: one process per GPU (NCCL), the product's BV type "b200" and Mat type "b200csr"
with rows split across the ranks (cases in tests/mgpu_cases.py).  Rank 0 writes the results as JSON to argv[2]."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from slepc_b200 import dist as D                          # noqa: E402
import mgpu_cases                                         # noqa: E402


def main():
    case, outpath = sys.argv[1], sys.argv[2]
    rank, world = D.init()
    if case not in mgpu_cases.CASES:
        raise SystemExit(f"unknown case {case}")
    res = mgpu_cases.CASES[case](rank, world)
    res.update(case=case, world=world)
    res.setdefault("p2p", bool(D.P2P))
    if rank == 0:
        json.dump(res, open(outpath, "w"))
    D.finalize()


if __name__ == "__main__":
    main()
