"""GPU tier (run with ``-m gpu`` on a B200): the CUDA path through the C-ABI
against the CPU oracle on the same seeded inputs.  Pattern arrays must be
byte-identical; values within 1e-13 of the slot's absolute contribution sum."""
import os

import numpy
import pytest
from scipy.sparse import linalg

from . import parity, problems

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _both():
    import oracle
    import pyfvm_b200

    return oracle, pyfvm_b200


def _mesh(kind, *args, jitter=False):
    from pyfvm_b200 import meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    if jitter:
        p, c = meshes.jitter(p, c, scale=0.2, seed=0)
    return meshes.Mesh(p, c)


LINEAR_CASES = [
    ("poisson", "rect", (401, 201), False),  # BASELINE config 1 (README example)
    ("poisson", "rect", (33, 17), True),
    ("convection_diffusion", "rect", (201, 101), False),  # config 4's operator
    ("convection_diffusion", "rect", (64, 47), True),
    ("reaction_x", "rect", (97, 55), True),
    ("subdomain_problem", "rect", (81, 41), False),
    ("poisson", "cube", (9,), False),
    ("poisson", "cube", (21,), True),  # negative ce_ratios
    ("reaction_x", "cube", (12,), True),
    ("poisson", "rect", (2, 2), False),  # smallest mesh: 2 cells, all boundary
]


@pytest.mark.parametrize("name,kind,args,jit", LINEAR_CASES)
def test_linear_assembly_parity(name, kind, args, jit):
    oracle, eng = _both()
    mesh = _mesh(kind, *args, jitter=jit)
    scales = {}
    A_ref, b_ref = oracle.discretize_linear(
        getattr(problems, name)(oracle.form_language), mesh, scales=scales
    )
    A, b = eng.discretize_linear(getattr(problems, name)(eng.form_language), mesh)
    parity.assert_matrix_close(A, A_ref, scales["matrix"])
    parity.assert_vector_close(b, b_ref, scales["rhs"])


NONLINEAR_CASES = [
    ("bratu", "rect", (101, 51), False),  # BASELINE config 2
    ("bratu", "cube", (11,), True),
    ("nonlinear_conv", "rect", (57, 33), True),
    ("nonlinear_conv", "cube", (8,), False),
]


@pytest.mark.parametrize("name,kind,args,jit", NONLINEAR_CASES)
def test_residual_and_jacobian_parity(name, kind, args, jit):
    oracle, eng = _both()
    mesh = _mesh(kind, *args, jitter=jit)
    n = len(mesh.points)
    f_ref, jac_ref = oracle.discretize(getattr(problems, name)(oracle.form_language), mesh)
    f, jac = eng.discretize(getattr(problems, name)(eng.form_language), mesh)
    rng = numpy.random.default_rng(0)
    for u in (numpy.zeros(n), 0.3 * rng.standard_normal(n)):
        s = {}
        F_ref = f_ref.eval(u, scales=s)
        parity.assert_vector_close(f.eval(u), F_ref, s["vec"])
        J_ref = jac_ref.get_linear_operator(u, scales=s)
        parity.assert_matrix_close(jac.get_linear_operator(u), J_ref, s["matrix"])


def test_readme_poisson_solution():
    """[README:25-54] end to end through the ``pyfvm`` shim."""
    import pyfvm
    from pyfvm.form_language import Boundary, dS, dV, integrate, n_dot_grad
    from pyfvm_b200 import meshes

    class Poisson:
        def apply(self, u):
            return integrate(lambda x: -n_dot_grad(u(x)), dS) - integrate(lambda x: 1.0, dV)

        def dirichlet(self, u):
            return [(lambda x: u(x) - 0.0, Boundary())]

    vertices, cells = meshes.rectangle_tri(
        numpy.linspace(0.0, 2.0, 401), numpy.linspace(0.0, 1.0, 201)
    )
    mesh = meshes.Mesh(vertices, cells)
    matrix, rhs = pyfvm.discretize_linear(Poisson(), mesh)
    assert matrix.nnz == 561801 and (matrix.data == 0.0).sum() == 163596
    u = linalg.spsolve(matrix, rhs)
    assert u.max() == pytest.approx(0.113871, abs=2e-6)


def test_readme_bratu_newton():
    """[README:86-124]: ``pyfvm.newton`` with per-step f.eval + Jacobian refresh."""
    import pyfvm
    from pyfvm.form_language import Boundary, dS, dV, integrate, n_dot_grad
    from pyfvm_b200 import meshes
    from sympy import exp

    class Bratu:
        def apply(self, u):
            return integrate(lambda x: -n_dot_grad(u(x)), dS) - integrate(
                lambda x: 2.0 * exp(u(x)), dV
            )

        def dirichlet(self, u):
            return [(u, Boundary())]

    vertices, cells = meshes.rectangle_tri(
        numpy.linspace(0.0, 2.0, 101), numpy.linspace(0.0, 1.0, 51)
    )
    mesh = meshes.Mesh(vertices, cells)
    f, jacobian = pyfvm.discretize(Bratu(), mesh)
    calls = []

    def jacobian_solver(u0, rhs):
        jac = jacobian.get_linear_operator(u0)
        calls.append(jac.nnz)
        return linalg.spsolve(jac, rhs)

    u = pyfvm.newton(f.eval, jacobian_solver, numpy.zeros(len(vertices)), verbose=False)
    assert calls == [35451] * 3
    assert u.max() == pytest.approx(0.28395, abs=1e-5)


@pytest.mark.parametrize("name", ["poisson_33x17", "convdiff_33x17", "poisson_cube5"])
def test_linear_golden(name):
    """Against the committed fixtures (no oracle at run time).  The fixtures are REGRESSION
    SNAPSHOTS generated from the repo's own oracle (tests/golden/make_golden.py): they pin
    the engine to what the oracle produced when they were written, not to the reference
    (which ships no vectors; parity is unpinned)."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes

    ref = numpy.load(os.path.join(GOLD, name + ".npz"))
    if name.endswith("cube5"):
        mesh = meshes.Mesh(*problems.cube(5))
    else:
        mesh = meshes.Mesh(*problems.rectangle(33, 17))
    prob = problems.poisson if name.startswith("poisson") else problems.convection_diffusion
    A, b = eng.discretize_linear(prob(eng.form_language), mesh)
    assert numpy.array_equal(A.indptr, ref["indptr"]) and numpy.array_equal(A.indices, ref["indices"])
    scale = numpy.abs(ref["data"]).max()
    assert numpy.abs(A.data - ref["data"]).max() <= 1e-13 * scale
    assert numpy.abs(b - ref["rhs"]).max() <= 1e-13 * max(1.0, numpy.abs(ref["rhs"]).max())


def test_bratu_golden():
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes

    ref = numpy.load(os.path.join(GOLD, "bratu_21x11.npz"))
    mesh = meshes.Mesh(*problems.rectangle(21, 11))
    f, jac = eng.discretize(problems.bratu(eng.form_language), mesh)
    F = f.eval(ref["u"])
    J = jac.get_linear_operator(ref["u"])
    assert numpy.array_equal(J.indptr, ref["indptr"]) and numpy.array_equal(J.indices, ref["indices"])
    assert numpy.abs(F - ref["F"]).max() <= 1e-13 * numpy.abs(ref["F"]).max()
    assert numpy.abs(J.data - ref["data"]).max() <= 1e-13 * numpy.abs(ref["data"]).max()


def test_determinism_and_tile_independence():
    """Two runs are bit-identical, and so are runs with different tile sizes
    (the fold order of a row never depends on tile boundaries)."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import engine, problem

    mesh = _mesh("cube", 13, jitter=True)
    P = problems.reaction_x(eng.form_language)
    lp = problem.LinearProblem(P, mesh)
    A1, b1 = lp.assemble()
    A2, b2 = lp.assemble()
    assert A1.data.tobytes() == A2.data.tobytes() and b1.tobytes() == b2.tobytes()
    for q in (512, 1024, 9000):
        dm = engine.DeviceMesh(lp.eng, mesh, need_el=True, tile_quantum=q)
        data, vec = lp.eng.alloc(8 * lp.nnz), lp.eng.alloc(8 * lp.n)
        dm.launch(lp.kernels[0], dirid=lp.dirid.ptr, data=data.ptr, vec=vec.ptr)
        d, v = numpy.empty(lp.nnz), numpy.empty(lp.n)
        data.download(d), vec.download(v)
        assert d.tobytes() == A1.data.tobytes() and v.tobytes() == b1.tobytes(), q


def test_large_mesh_properties():
    """At a size the oracle does not finish quickly: size-independent properties.
    Poisson rows sum to zero off the boundary, the matrix is symmetric there,
    rhs = control volumes, and the residual of the linear problem written as a
    nonlinear one satisfies F(u) = A u - b."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes

    mesh = meshes.Mesh(*problems.rectangle(1201, 801))
    n = len(mesh.points)
    A, b = eng.discretize_linear(problems.convection_diffusion(eng.form_language), mesh)
    assert A.nnz == n + 2 * (1200 * 801 + 1201 * 800 + 1200 * 800)
    interior = ~mesh.is_boundary_point
    assert numpy.array_equal(b[~interior], numpy.zeros((~interior).sum()))
    assert numpy.allclose(b[interior], mesh.control_volumes[interior], rtol=1e-15, atol=0)
    f, jac = eng.discretize(problems.convection_diffusion(eng.form_language), mesh)
    u = numpy.random.default_rng(0).standard_normal(n)
    F = f.eval(u)
    lin = A @ u - b
    scale = abs(A) @ abs(u) + abs(b)
    assert (numpy.abs(F - lin) <= 1e-13 * scale).all()
    J = jac.get_linear_operator(u)
    assert J.indices.tobytes() == A.indices.tobytes()
    assert numpy.abs(J.data - A.data).max() <= 1e-13 * numpy.abs(A.data).max()


def test_norm2_matches_numpy():
    from pyfvm_b200.engine import Engine

    eng = Engine.get()
    x = numpy.random.default_rng(1).standard_normal(1_000_003)
    d = eng.to_device(x)
    assert eng.norm2(d.ptr, x.size) == pytest.approx(numpy.linalg.norm(x), rel=1e-14)
