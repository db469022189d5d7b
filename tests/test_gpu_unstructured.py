"""GPU tier: meshes without lexicographic vertex numbering [README:56-64: "anything
else that provides vertices and cells works as well"].  A randomly renumbered mesh
puts most columns of a tile outside its staged vertex table, so the kernel's cold
global gather (``fvm_gather_column``) runs -- asserted through ``uncovered_slots`` --
and the result must still match the oracle on the same renumbered mesh."""
import numpy
import pytest

from . import parity, problems

pytestmark = pytest.mark.gpu


def _mesh(kind, args, seed, window, jitter=True):
    from pyfvm_b200 import meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    if jitter:
        p, c = meshes.jitter(p, c, scale=0.15, seed=seed)
    p, c = problems.shuffle_numbering(p, c, seed=seed, window=window)
    return meshes.Mesh(p, c)


CASES = [
    ("convection_diffusion", "rect", (61, 43), None),
    ("reaction_x", "rect", (40, 37), 97),  # block shuffle: many short halo ranges
    ("poisson", "cube", (13,), None),
    ("reaction_x", "cube", (11,), 50),
]


@pytest.mark.parametrize("name,kind,args,window", CASES)
def test_linear_shuffled_numbering(name, kind, args, window):
    import oracle
    import pyfvm_b200 as eng
    from pyfvm_b200 import problem

    mesh = _mesh(kind, args, 5, window)
    s = {}
    A_ref, b_ref = oracle.discretize_linear(
        getattr(problems, name)(oracle.form_language), mesh, scales=s
    )
    lp = problem.LinearProblem(getattr(problems, name)(eng.form_language), mesh)
    A, b = lp.assemble()
    if window is None:
        assert lp.dmesh.stats()["uncovered_slots"] > 0, "the cold gather path did not run"
    parity.assert_matrix_close(A, A_ref, s["matrix"])
    parity.assert_vector_close(b, b_ref, s["rhs"])


NL_CASES = [
    ("bratu", "rect", (57, 41), None),
    ("nonlinear_conv", "rect", (45, 31), 64),
    ("nonlinear_conv", "cube", (10,), None),
]


@pytest.mark.parametrize("name,kind,args,window", NL_CASES)
def test_nonlinear_shuffled_numbering(name, kind, args, window):
    import oracle
    import pyfvm_b200 as eng

    mesh = _mesh(kind, args, 9, window)
    n = len(mesh.points)
    f_ref, jac_ref = oracle.discretize(getattr(problems, name)(oracle.form_language), mesh)
    f, jac = eng.discretize(getattr(problems, name)(eng.form_language), mesh)
    if window is None:
        assert f._s.dmesh.stats()["uncovered_slots"] > 0, "the cold gather path did not run"
    u = 0.3 * numpy.random.default_rng(1).standard_normal(n)
    s = {}
    F_ref = f_ref.eval(u, scales=s)
    parity.assert_vector_close(f.eval(u), F_ref, s["vec"])
    J_ref = jac_ref.get_linear_operator(u, scales=s)
    parity.assert_matrix_close(jac.get_linear_operator(u), J_ref, s["matrix"])


def test_spmv_shuffled_numbering():
    """y = A x with cold-gathered x entries against scipy."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import problem, solvers

    mesh = _mesh("rect", (50, 33), 3, None)
    lp = problem.LinearProblem(problems.convection_diffusion(eng.form_language), mesh)
    A, _ = lp.assemble()
    assert lp.dmesh.stats()["uncovered_slots"] > 0
    x = numpy.random.default_rng(2).standard_normal(lp.n)
    sysm = solvers.DeviceSystem(lp.dmesh, lp.data.ptr)
    xd, yd = lp.eng.to_device(x), lp.eng.alloc(8 * lp.n)
    sysm.matvec(xd.ptr, yd.ptr)
    y = numpy.empty(lp.n)
    yd.download(y)
    ref = A @ x
    assert (numpy.abs(y - ref) <= 1e-13 * (abs(A) @ abs(x))).all()
