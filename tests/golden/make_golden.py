#!/usr/bin/env python
"""Extracts the known-answer numbers of the reference's own test outputs for the Krylov hot path into
tests/golden/reference_outputs.json.  Run in the container that has /root/reference; the JSON is what travels.

    python tests/golden/make_golden.py

For every output file: the numeric tokens of each line that contains any (in file order), with the command-line
arguments of the test that produced it (from the /*TEST*/ block of the source) recorded by hand below."""
import json
import os
import re

REF = "/root/reference/src"
FILES = {
    # output file: (source, args that matter)
    "sys/classes/bv/tests/output/test1_1_bv_type-svec.out": ("sys/classes/bv/tests/test1.c", "n=10 k=5 l=3 -bv_type svec -verbose"),
    "sys/classes/bv/tests/output/test2_1.out": ("sys/classes/bv/tests/test2.c", "n=20 k=8, CGS/IFNEEDED default"),
    "sys/classes/bv/tests/output/test4_1.out": ("sys/classes/bv/tests/test4.c", "n=18 k=12 l=3; second BV k=8 l=2"),
    "sys/classes/bv/tests/output/test8_1.out": ("sys/classes/bv/tests/test8.c", "n=20 k=7 selected columns"),
    "sys/classes/bv/tests/output/test13_1.out": ("sys/classes/bv/tests/test13.c", "n=10 k=5"),
    "eps/tests/output/test4_1.out": ("eps/tests/test4.c", "1-D Laplacian n=30 nev=4 -eps_type krylovschur"),
    "eps/tests/output/test2_1.out": ("eps/tests/test2.c", "1-D Laplacian n=30 nev=4: largest / smallest / interior(target 2.1, sinvert)"),
    "eps/tests/output/test1_1.out": ("eps/tests/test1.c", "GHEP: 2-D Laplacian 18x18, B = diag(2/log(i+2)), nev=4 (B-orthonormal eigenvectors)"),
    "eps/tests/output/test6_1.out": ("eps/tests/test6.c", "diagonal matrix 1..30, nev=4"),
    "eps/tutorials/output/ex13_1.out": ("eps/tutorials/ex13.c", "GHEP sinvert: 2-D Laplacian 10x10, B = 4 I, nev=4 ncv=22 tol=1e-5"),
    "eps/tutorials/output/ex2_1.out": ("eps/tutorials/ex2.c", "2-D Laplacian n=72 nev=4"),
    "eps/tutorials/output/ex5_1.out": ("eps/tutorials/ex5.c", "Markov m=15 nev=4 -eps_largest_real"),
    "svd/tests/output/test3_1.out": ("svd/tests/test3.c", "Grcar-like 35x30 nsv=4 trlanczos"),
    "svd/tutorials/output/ex8_1.out": ("svd/tutorials/ex8.c", "Grcar n=30: sigma_1, sigma_n, condition number"),
    "svd/tests/output/test8_1.out": ("svd/tests/test8.c", "Grcar n=30 nsv=3 ncv=12 tol=1e-6, solved twice (ncv, ncv+2)"),
    "svd/tests/output/test9_1.out": ("svd/tests/test9.c", "Grcar n=30 then n=60 on the same SVD object, nsv=3"),
    "svd/tests/output/test14_1.out": ("svd/tests/test14.c", "two 20x22 bidiagonal matrices on the same SVD object, nsv=3"),
    "svd/tests/output/test4_1.out": ("svd/tests/test4.c", "rectangular bidiagonal 20x22 (more columns than rows), suffix 1_trlanczos: ncv=12 restart 0.6"),
}
NUM = re.compile(r"(?<![\w.])[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?(?![\w])")


def main():
    out = {}
    for rel, (src, args) in FILES.items():
        rows = []
        for line in open(os.path.join(REF, rel)):
            s = line.strip()
            if not s or "MPI process" in s or s.startswith("type:"):
                continue
            toks = NUM.findall(s.replace("100*eps", ""))
            if toks:
                rows.append([float(t) for t in toks])
        out[rel] = {"source": src, "args": args, "rows": rows}
    # the one output kept verbatim (a golden OUTPUT, not source): bv/tests/test1.c -verbose, compared line by line with
    # examples/bv_test1.c in tests/test_z_examples.py
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(REF, "sys/classes/bv/tests/output/test1_1_bv_type-svec.out")) as f, open(os.path.join(here, "bv_test1_1_svec.out"), "w") as g:
        g.write(f.read())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_outputs.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, {k: len(v["rows"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
