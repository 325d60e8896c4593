"""The reference's tutorial / test programs rewritten against include/b2kslepc.h (examples/*.c): what a user of the
reference would compile after switching.  Each program's output is compared LINE BY LINE with the reference's own output
file (src/eps/tutorials/output/ex2_1.out, ex3_1.out, ex5_1.out, src/eps/tests/output/test4_1.out, bv/tests/output/test1_1, test2_1, src/svd/tutorials/output/ex8_1.out, src/svd/tests/output/test3_1.out — copied below as golden text).
  * not gpu: the programs are compiled against the CPU oracle plug-in through tests/ex_cpu_shim.h (host logic + format),
    and the product binaries are checked to FAIL LOUDLY without a GPU (no CPU fallback);
  * gpu: the product binaries (BV type b200, Mat type b200csr) run on cuda:0.
(The file sorts last on purpose: the kernel and solver parity tests run before the example programs.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")

EX2_OUT = """
2-D Laplacian Eigenproblem, N=5184 (72x72 grid)

 Solution method: krylovschur

 Number of requested eigenvalues: 4
 All requested eigenvalues computed up to the required tolerance:
     7.99630, 7.99074, 7.98519, 7.98150

"""
EX3_OUT = EX2_OUT.replace("Eigenproblem,", "Eigenproblem (matrix-free version),")
EX5_OUT = """
Markov Model, N=120 (m=15)

 Solution method: krylovschur

 Number of requested eigenvalues: 4
 All requested eigenvalues computed up to the required tolerance:
     1.00000, 0.97137, 0.90423, 0.85714

"""
TEST3_OUT = """
SVD of a rectangular Grcar matrix, 35x30

 All requested singular values computed up to the required tolerance:
     3.22175, 3.21797, 3.16825, 3.15128

Level of orthogonality below the tolerance
"""
BV_TEST1_OUT = open(os.path.join(ROOT, "tests", "golden", "bv_test1_1_svec.out")).read()   # bv/tests/output/test1_1_bv_type-svec.out


def bv_test1_expected(bvtype):
    """the reference's output for -bv_type svec with the two lines that name the implementation adapted: the type, and how
    BVMatMult runs (this build's ops.matmult is a column loop = the reference's BV_MATMULT_VECS wording, bvfunc.c:589-595)"""
    return BV_TEST1_OUT.replace("type: svec", "type: " + bvtype).replace("doing matmult as a single matrix-matrix product",
                                                                           "doing matmult as matrix-vector products")


TEST4_OUT = """
1-D Laplacian Eigenproblem, n=30

 Number of requested eigenvalues: 4
 All requested eigenvalues computed up to the required tolerance:
     3.98974, 3.95906, 3.90828, 3.83792

"""
BV_TEST2_OUT = """Test BV orthogonalization with 8 columns of length 20.
Level of orthogonality < 100*eps
Level of orthogonality < 100*eps
Residual ||X-QR|| < 100*eps
Norm of ones(n,1) after orthogonalizing against X: 2.50931
"""
EX8_OUT = """
Estimate the condition number of a Grcar matrix, n=30

 Computed singular values: sigma_1=3.2215, sigma_n=0.9551
 Estimated condition number: sigma_1/sigma_n=3.3731

"""
CASES = [
    ("svd_ex8", [], EX8_OUT),
    ("eps_test4", [], TEST4_OUT),
    ("bv_test2", [], BV_TEST2_OUT),
    ("bv_test2", ["-bv_orthog_type", "mgs"], BV_TEST2_OUT),
    ("ex2", ["-n", "72", "-eps_nev", "4", "-eps_ncv", "20", "-terse"], EX2_OUT),
    ("ex3", ["-n", "72", "-eps_nev", "4", "-eps_ncv", "20", "-terse"], EX3_OUT),
    ("ex5", ["-m", "15", "-eps_nev", "4", "-eps_largest_real", "-terse"], EX5_OUT),
    ("svd_test3", ["-svd_nsv", "4"], TEST3_OUT),
    ("svd_test3", ["-svd_nsv", "4", "-svd_trlanczos_locking", "0"], TEST3_OUT),
    ("svd_test3", ["-svd_nsv", "4", "-svd_trlanczos_oneside"], TEST3_OUT),
    ("svd_test3", ["-svd_nsv", "4", "-svd_trlanczos_oneside", "-bv_orthog_type", "mgs"], TEST3_OUT),
    ("svd_test3", ["-svd_nsv", "4", "-svd_trlanczos_oneside", "-bv_orthog_refine", "always"], TEST3_OUT),
]
IDS = [c[0] + "".join(a for a in c[1] if a.startswith("-") and a not in ("-terse",)).replace("-", "_") for c in CASES]


def ensure_built():
    if not all(os.path.exists(os.path.join(BIN, n)) for n in ("ex2", "ex3", "ex5", "svd_test3", "bv_test1", "bv_test2", "eps_test4", "svd_ex8")):
        subprocess.run(["make", "-C", ROOT, "all"], check=True, capture_output=True)


@pytest.fixture(scope="module")
def cpu_bins(tmp_path_factory):
    """the same sources compiled against the CPU oracle plug-in (tests/ex_cpu_shim.h)"""
    ensure_built()
    out = tmp_path_factory.mktemp("ex_cpu")
    lib = os.path.join(ROOT, "slepc_b200", "lib")
    orc = os.path.join(ROOT, "oracle", "_build")
    for name in ("ex2", "ex3", "ex5", "svd_test3", "bv_test1", "bv_test2", "eps_test4", "svd_ex8"):
        subprocess.run(["gcc", "-O1", "-std=gnu11", "-Wno-unused-function", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "examples"),
                        "-include", os.path.join(ROOT, "tests", "ex_cpu_shim.h"), "-o", str(out / name), os.path.join(ROOT, "examples", name + ".c"),
                        "-L" + orc, "-loraclecpu", "-L" + lib, "-lb2kslepc", "-lb200krylov",
                        "-Wl,-rpath," + orc, "-Wl,-rpath," + lib, "-lm"], check=True, capture_output=True)
    return out


@pytest.mark.parametrize("name,args,gold", CASES, ids=IDS)
def test_example_matches_reference_output_cpu_oracle(cpu_bins, name, args, gold):
    r = subprocess.run([str(cpu_bins / name)] + args, capture_output=True, text=True, timeout=120, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stderr
    assert r.stdout == gold


def test_bv_test1_verbose_matches_reference_output_cpu_oracle(cpu_bins):
    """bv/tests/test1.c -verbose: every BVView / MatView / VecView line of output/test1_1_bv_type-svec.out (the reference filters
    negative zeros with sed -e 's/-0[.]/0./g', test1.c:188)"""
    r = subprocess.run([str(cpu_bins / "bv_test1"), "-verbose"], capture_output=True, text=True, timeout=120, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stderr
    assert r.stdout.replace("-0.\n", "0.\n") == bv_test1_expected("oraclecpu")


def test_product_binaries_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    ensure_built()
    r = subprocess.run([os.path.join(BIN, "ex2"), "-n", "8"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,args,gold", CASES, ids=IDS)
def test_example_matches_reference_output_gpu(name, args, gold):
    ensure_built()
    r = subprocess.run([os.path.join(BIN, name)] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == gold


@pytest.mark.gpu
def test_bv_test1_verbose_matches_reference_output_gpu():
    ensure_built()
    r = subprocess.run([os.path.join(BIN, "bv_test1"), "-verbose"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.replace("-0.\n", "0.\n") == bv_test1_expected("b200")
