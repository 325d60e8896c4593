"""The slepc4py-shaped Python binding (slepc_b200/slepc4py_compat.py) driven the way the reference's demos drive slepc4py
(src/binding/slepc4py/demo/ex1.py: 1-D Laplacian HEP with a monitor; ex4.py-style SVD), on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_demo_ex1_1d_laplacian_hep():
    from slepc_b200.slepc4py_compat import PETSc, SLEPc
    PETSc.init(0)
    n = 30
    A = PETSc.Mat(); A.create()
    A.setSizes([n, n])
    A.setFromOptions()
    rstart, rend = A.getOwnershipRange()
    if rstart == 0:
        A[0, :2] = [2, -1]
        rstart += 1
    if rend == n:
        A[n - 1, -2:] = [-1, 2]
        rend -= 1
    for i in range(rstart, rend):
        A[i, i - 1:i + 2] = [-1, 2, -1]
    A.assemble()
    E = SLEPc.EPS(); E.create()
    E.setOperators(A)
    E.setProblemType(SLEPc.EPS.ProblemType.HEP)
    history = []

    def monitor(eps, its, nconv, eig, err):
        if nconv < len(err):
            history.append(err[nconv])
    E.setMonitor(monitor)
    E.setDimensions(4)
    E.setFromOptions()
    E.solve()
    assert E.getType() == "krylovschur"
    nev, ncv, mpd = E.getDimensions()
    assert nev == 4
    nconv = E.getConverged()
    assert nconv >= 4 and len(history) == E.getIterationNumber()
    v, _ = A.createVecs()
    lam = []
    for i in range(nconv):
        k = E.getEigenpair(i, v)
        assert k.imag == 0.0
        assert E.computeError(i) < 5e-8
        lam.append(k.real)
    # the reference's golden for this problem (ex1f_1.out / eps test4_1.out)
    assert [f"{x:.5f}" for x in lam[:4]] == ["3.98974", "3.95906", "3.90828", "3.83792"]
    assert abs(v.norm() - 1.0) < 1e-12
    E.destroy(); A.destroy()


def test_generalized_sinvert_and_svd_through_the_binding():
    import scipy.sparse as sp
    from oracle import slepc_oracle as O
    from slepc_b200.slepc4py_compat import PETSc, SLEPc
    PETSc.init(0)
    n = 10
    Asp = O.laplacian_2d(n).tocsr()
    A = PETSc.Mat().createAIJ(Asp.shape, csr=(Asp.indptr, Asp.indices, Asp.data))
    B = PETSc.Mat.fromScipy(sp.identity(n * n, format="csr") * 4.0)
    E = SLEPc.EPS().create()
    E.setOperators(A, B)
    E.setProblemType(SLEPc.EPS.ProblemType.GHEP)
    E.setDimensions(4, 22)
    E.setTolerances(1e-5)
    st = E.getST()
    st.setType(SLEPc.ST.Type.SINVERT)
    E.setTarget(0.0)
    E.setWhichEigenpairs(SLEPc.EPS.Which.TARGET_MAGNITUDE)
    E.solve()
    assert [f"{E.getEigenvalue(i).real:.5f}" for i in range(4)] == ["0.04051", "0.09963", "0.09963", "0.15875"]   # ex13_1.out
    E.destroy()
    # SVD of a rectangular sparse matrix (the demo ex4.py pattern)
    M = sp.random(400, 150, density=0.05, random_state=3, format="csr")
    Ms = PETSc.Mat.fromScipy(M)
    Sv = SLEPc.SVD().create()
    Sv.setOperators(Ms)
    Sv.setDimensions(3)
    Sv.solve()
    assert Sv.getConverged() >= 3
    ref = np.linalg.svd(M.toarray(), compute_uv=False)
    v, u = Ms.createVecs()
    for i in range(3):
        sigma = Sv.getSingularTriplet(i, u, v)
        assert abs(sigma - ref[i]) < 1e-10 * ref[i]
        assert Sv.computeError(i) < 5e-8
    Sv.destroy()
