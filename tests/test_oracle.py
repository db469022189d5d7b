"""CPU tier: pin the oracle to the independent known answers of SURVEY §4 and
to the committed golden fixtures (the reference ships no golden vectors)."""
import json
import os

import numpy
import pytest
from scipy.sparse import linalg

import oracle
from pyfvm_b200 import meshes

from . import problems

fl = oracle.form_language
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def poisson_c1():
    v, c = problems.rectangle(401, 201)
    mesh = meshes.Mesh(v, c)
    A, b = oracle.discretize_linear(problems.poisson(fl), mesh)
    return mesh, A, b


def test_c1_counts(poisson_c1):
    mesh, A, b = poisson_c1
    assert len(mesh.points) == 80601
    assert len(mesh.cells_points) == 160000
    assert mesh.idx_hierarchy.shape == (2, 3, 160000)
    assert A.nnz == 561801  # n + 2E
    assert A.indices.dtype == numpy.int32 and A.indptr.dtype == numpy.int32
    assert A.has_canonical_format
    assert mesh.is_boundary_point.sum() == 1200
    assert abs(mesh.control_volumes.sum() - 2.0) < 1e-14
    assert (A.data == 0.0).sum() == 163596  # explicit zeros are kept


def test_c1_interior_row_is_5_point_stencil(poisson_c1):
    mesh, A, b = poisson_c1
    i = 401 * 100 + 200
    row = A[i].toarray()[0]
    assert row[i] == pytest.approx(4.0, abs=1e-12)
    for j in (i - 1, i + 1, i - 401, i + 401):
        assert row[j] == pytest.approx(-1.0, abs=1e-12)
    assert abs(row.sum()) < 2e-15
    # hypotenuse neighbours are stored as explicit zeros
    lo, hi = A.indptr[i], A.indptr[i + 1]
    assert hi - lo in (7, 9) or hi - lo == 5 + 2 * 2
    assert (A.data[lo:hi] == 0.0).sum() == (hi - lo) - 5


def test_c1_dirichlet_rows(poisson_c1):
    mesh, A, b = poisson_c1
    for i in (0, 400, 401 * 200, 401 * 57):
        assert mesh.is_boundary_point[i]
        lo, hi = A.indptr[i], A.indptr[i + 1]
        cols, vals = A.indices[lo:hi], A.data[lo:hi]
        assert hi - lo >= 3  # pattern unchanged
        assert vals[cols == i] == 1.0
        assert (vals[cols != i] == 0.0).all()
        assert b[i] == 0.0
    interior = ~mesh.is_boundary_point
    assert numpy.allclose(b[interior], mesh.control_volumes[interior], rtol=0, atol=1e-18)


def test_c1_solution_max(poisson_c1):
    mesh, A, b = poisson_c1
    u = linalg.spsolve(A, b)
    assert u.max() == pytest.approx(0.113871, abs=2e-6)


def test_c2_bratu_newton():
    v, c = problems.rectangle(101, 51)
    mesh = meshes.Mesh(v, c)
    f, jac = oracle.discretize(problems.bratu(fl), mesh)
    norms = []

    def feval(u):
        out = f.eval(u)
        norms.append(numpy.sqrt(out @ out))
        return out

    def solver(u0, rhs):
        return linalg.spsolve(jac.get_linear_operator(u0), rhs)

    u = oracle.newton(feval, solver, numpy.zeros(len(v)), verbose=False)
    assert len(v) == 5151
    assert jac.get_linear_operator(u).nnz == 35451
    assert len(norms) == 4 and norms[-1] < 1e-10  # 3 Newton steps from 0
    assert norms[0] == pytest.approx(5.57193e-2, rel=1e-5)
    assert u.max() == pytest.approx(0.28395, abs=1e-5)


def test_cube_5tet_known_answers():
    p, c = problems.cube(9)
    mesh = meshes.Mesh(p, c)
    assert (len(p), len(c)) == (729, 2560)
    assert mesh.idx_hierarchy.shape == (2, 3, 4, 2560)
    assert abs(mesh.control_volumes.sum() - 1.0) < 1e-14
    A, b = oracle.discretize_linear(problems.poisson(fl), mesh)
    E = 3 * 81 * 8 + 3 * 64 * 9
    assert E == 3672 and A.nnz == 729 + 2 * E == 8073
    i = (4 * 9 + 4) * 9 + 4
    h = 1.0 / 8
    row = A[i].toarray()[0]
    assert row[i] == pytest.approx(6 * h, abs=1e-13)
    for j in (i - 1, i + 1, i - 9, i + 9, i - 81, i + 81):
        assert row[j] == pytest.approx(-h, abs=1e-13)


def test_jittered_tets_keep_volume():
    p, c = problems.cube(7)
    p2, c2 = meshes.jitter(p, c, scale=0.2, seed=0)
    mesh = meshes.Mesh(p2, c2)
    assert (mesh.ce_ratios < 0).any()
    assert abs(mesh.control_volumes.sum() - 1.0) < 1e-13


def test_appendix_c_counts():
    for nx, ny in ((101, 51), (33, 17)):
        v, c = problems.rectangle(nx, ny)
        mesh = meshes.Mesh(v, c)
        A, _ = oracle.discretize_linear(problems.poisson(fl), mesh)
        E = (nx - 1) * ny + nx * (ny - 1) + (nx - 1) * (ny - 1)
        assert len(c) == 2 * (nx - 1) * (ny - 1)
        assert A.nnz == nx * ny + 2 * E
        assert mesh.is_boundary_point.sum() == 2 * (nx + ny) - 4


def test_split():
    import sympy

    a, b, x = sympy.symbols("a b x")
    aff, lin, non = oracle.split(3 * a + 2 * b * x + x + sympy.sin(a) * sympy.exp(b), [a, b])
    assert aff == x
    assert lin[0] == 3 and lin[1] == 2 * x
    assert sympy.simplify(non - sympy.sin(a) * sympy.exp(b)) == 0


def test_scipy_keeps_explicit_zeros_and_sums_duplicates():
    from scipy import sparse

    m = sparse.coo_matrix(([1.0, -1.0, 2.0, 3.0], ([0, 0, 1, 1], [1, 1, 0, 0])), shape=(2, 2)).tocsr()
    assert m.nnz == 2 and m.data.tolist() == [0.0, 5.0]
    assert m.indices.dtype == numpy.int32 and m.has_canonical_format


@pytest.mark.parametrize("name", ["poisson_33x17", "convdiff_33x17", "bratu_21x11", "poisson_cube5"])
def test_golden(name):
    """Oracle output is stable against the committed fixtures (generated by
    tests/golden/make_golden.py from this very oracle, cross-checked by the
    known-answer tests above)."""
    from .golden import make_golden

    got = make_golden.CASES[name]()
    ref = numpy.load(os.path.join(GOLD, name + ".npz"))
    for k in ref.files:
        if ref[k].dtype.kind in "iu":
            assert numpy.array_equal(got[k], ref[k]), k
        else:
            assert numpy.allclose(got[k], ref[k], rtol=1e-12, atol=1e-15), k
