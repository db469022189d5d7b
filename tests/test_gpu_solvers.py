"""GPU tier: SpMV on the assembled system and the device Krylov / Newton built on
it against scipy on the host."""
import numpy
import pytest
from scipy.sparse import linalg

from . import problems

pytestmark = pytest.mark.gpu


def _mesh(kind, *args, jit=True):
    from pyfvm_b200 import meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    if jit:
        p, c = meshes.jitter(p, c, scale=0.2, seed=0)
    return meshes.Mesh(p, c)


@pytest.mark.parametrize(
    "name,kind,args", [("convection_diffusion", "rect", (97, 61)), ("reaction_x", "cube", (12,)),
                       ("poisson", "rect", (2, 2))]
)
def test_spmv_matches_scipy(name, kind, args):
    import pyfvm_b200 as eng
    from pyfvm_b200 import problem, solvers

    mesh = _mesh(kind, *args, jit=args != (2, 2))
    lp = problem.LinearProblem(getattr(problems, name)(eng.form_language), mesh)
    A, _ = lp.assemble()
    sysm = solvers.DeviceSystem(lp.dmesh, lp.data.ptr)
    rng = numpy.random.default_rng(0)
    x = rng.standard_normal(lp.n)
    xd, yd = lp.eng.to_device(x), lp.eng.alloc(8 * lp.n)
    sysm.matvec(xd.ptr, yd.ptr)
    y = numpy.empty(lp.n)
    yd.download(y)
    ref = A @ x
    bound = 1e-13 * (abs(A) @ abs(x))
    assert (numpy.abs(y - ref) <= bound + 1e-300).all()
    # deterministic
    sysm.matvec(xd.ptr, yd.ptr)
    y2 = numpy.empty(lp.n)
    yd.download(y2)
    assert y.tobytes() == y2.tobytes()
    d = numpy.empty(lp.n)
    sysm.diagonal().download(d)
    assert numpy.array_equal(d, A.diagonal())


def test_device_bicgstab_solves_readme_poisson():
    """[README:25-54] with the solve on the device instead of spsolve."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, problem, solvers

    mesh = meshes.Mesh(*problems.rectangle(201, 101))
    lp = problem.LinearProblem(problems.poisson(eng.form_language), mesh)
    u, iters, relres = solvers.solve_linear(lp, tol=1e-11)
    A, b = lp.assemble()
    u_ref = linalg.spsolve(A, b)
    assert relres <= 1e-11 and iters > 0
    assert numpy.abs(u - u_ref).max() <= 1e-8 * numpy.abs(u_ref).max()
    assert u.max() == pytest.approx(0.11387, abs=2e-5)


def test_newton_with_device_jacobian_solver_and_device_newton():
    """[README:86-124]: pyfvm.newton with a device jacobian_solver, and the fully
    device-resident Newton, against the host spsolve version."""
    import pyfvm
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, solvers

    mesh = meshes.Mesh(*problems.rectangle(101, 51))
    f, jacobian = pyfvm.discretize(problems.bratu(eng.form_language), mesh)
    n = len(mesh.points)

    def host_solver(u0, rhs):
        return linalg.spsolve(jacobian.get_linear_operator(u0), rhs)

    u_host = pyfvm.newton(f.eval, host_solver, numpy.zeros(n), verbose=False)
    u_cb = pyfvm.newton(f.eval, solvers.jacobian_solver(jacobian), numpy.zeros(n), verbose=False)
    u_dev, steps = solvers.newton_device(f, jacobian, numpy.zeros(n))
    assert steps == 3
    assert u_host.max() == pytest.approx(0.28395, abs=1e-5)
    assert numpy.abs(u_cb - u_host).max() <= 1e-8
    assert numpy.abs(u_dev - u_host).max() <= 1e-8


def test_graph_bicgstab_matches_host_driven_loop():
    """fvm_bicgstab (device scalars, CUDA-graph iterations) against the host-driven loop."""
    import time

    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, problem, solvers

    mesh = meshes.Mesh(*problems.rectangle(151, 77))
    lp = problem.LinearProblem(problems.convection_diffusion(eng.form_language), mesh)
    lp.assemble_device()
    sysm = solvers.DeviceSystem(lp.dmesh, lp.data.ptr)
    xg, xh = lp.eng.alloc(8 * lp.n), lp.eng.alloc(8 * lp.n)
    t0 = time.time()
    ig, rg = sysm.bicgstab(lp.vec.ptr, xg.ptr, tol=1e-11, x0_is_zero=True)
    t1 = time.time()
    ih, rh = sysm.bicgstab(lp.vec.ptr, xh.ptr, tol=1e-11, x0_is_zero=True, graph=False)
    t2 = time.time()
    a, b = numpy.empty(lp.n), numpy.empty(lp.n)
    xg.download(a), xh.download(b)
    assert rg <= 1e-11 and rh <= 1e-11
    assert abs(ig - ih) <= max(3, ih // 10)  # same algorithm, operations fused differently
    assert numpy.abs(a - b).max() <= 1e-8 * numpy.abs(b).max()
    print(f"graph: {ig} its {t1 - t0:.4f} s; host-driven: {ih} its {t2 - t1:.4f} s")


@pytest.mark.parametrize("name,kind,args", [("poisson", "rect", (201, 101)), ("robin", "rect", (121, 77)),
                                            ("poisson", "cube", (21,)), ("neumann_only", "rect", (64, 48))])
def test_device_cg_solves_symmetric_systems(name, kind, args):
    """fvm_cg (Jacobi-preconditioned CG, device scalars, graph replay) on the symmetric
    systems of the README's family -- Dirichlet rows keep their columns, so the matrix as
    assembled is NOT symmetric; the Jacobi start keeps those rows out of the Krylov space."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import problem, solvers

    mesh = _mesh(kind, *args)
    lp = problem.LinearProblem(getattr(problems, name)(eng.form_language), mesh)
    u, iters, relres = solvers.solve_linear(lp, tol=1e-11, method="cg")
    A, b = lp.assemble()
    u_ref = linalg.spsolve(A.tocsc(), b)
    assert relres <= 1e-11 and iters > 0
    assert numpy.abs(u - u_ref).max() <= 1e-8 * numpy.abs(u_ref).max()
    # same answer as BiCGStab with about half its SpMVs or fewer
    u_b, iters_b, _ = solvers.solve_linear(lp, tol=1e-11)
    assert numpy.abs(u - u_b).max() <= 1e-8 * numpy.abs(u_ref).max()
    assert iters <= 2 * iters_b + 10
    # deterministic
    u2, iters2, _ = solvers.solve_linear(lp, tol=1e-11, method="cg")
    assert iters2 == iters and u2.tobytes() == u.tobytes()


def test_device_cg_negative_definite_and_breakdown():
    """The negated Poisson system is negative definite: CG's formulas do not care.  An
    indefinite one (-Laplace - 1000) must be refused, not silently mis-solved."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, problem, solvers

    fl = eng.form_language

    class Negated:
        def apply(self, u):
            return fl.integrate(lambda x: fl.n_dot_grad(u(x)), fl.dS) + fl.integrate(lambda x: 1.0, fl.dV)

        def dirichlet(self, u):
            return [(lambda x: u(x) - 0.25, fl.Boundary())]

    class Indefinite:
        def apply(self, u):
            return fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS) - fl.integrate(
                lambda x: 1000.0 * u(x) + 1.0, fl.dV)

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    mesh = meshes.Mesh(*problems.rectangle(101, 101))
    lp = problem.LinearProblem(Negated(), mesh)
    u, iters, relres = solvers.solve_linear(lp, tol=1e-11, method="cg")
    A, b = lp.assemble()
    u_ref = linalg.spsolve(A.tocsc(), b)
    assert relres <= 1e-11 and numpy.abs(u - u_ref).max() <= 1e-8 * numpy.abs(u_ref).max()
    with pytest.raises(ArithmeticError, match="CG "):
        solvers.solve_linear(problem.LinearProblem(Indefinite(), mesh), tol=1e-11, method="cg")
    with pytest.raises(ValueError, match="unknown linear solver"):
        solvers.solve_linear(lp, method="gmres")


def test_device_newton_with_cg():
    """[README:86-124] Bratu: the Jacobian is symmetric definite on the lower solution
    branch; the device Newton with CG gives the same iterates as with BiCGStab."""
    import pyfvm
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, solvers

    mesh = meshes.Mesh(*problems.rectangle(101, 51))
    f, jacobian = pyfvm.discretize(problems.bratu(eng.form_language), mesh)
    n = len(mesh.points)
    u_b, steps_b = solvers.newton_device(f, jacobian, numpy.zeros(n))
    u_c, steps_c = solvers.newton_device(f, jacobian, numpy.zeros(n), linear_solver="cg")
    assert steps_c == steps_b == 3
    assert numpy.abs(u_c - u_b).max() <= 1e-8
    u_cb = pyfvm.newton(f.eval, solvers.jacobian_solver(jacobian, method="cg"), numpy.zeros(n), verbose=False)
    assert numpy.abs(u_cb - u_b).max() <= 1e-8


@pytest.mark.parametrize("name,kind,args", [("poisson", "rect", (201, 101)), ("robin", "rect", (121, 77)),
                                            ("poisson", "cube", (21,))])
def test_chebyshev_preconditioned_cg(name, kind, args):
    """fvm_cg_poly: the Chebyshev polynomial of D^-1 A as preconditioner -- the same solution
    in 2.5-4 times fewer iterations (reductions) for about the same number of SpMVs."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import problem, solvers

    mesh = _mesh(kind, *args)
    lp = problem.LinearProblem(getattr(problems, name)(eng.form_language), mesh)
    A, b = lp.assemble()
    u_ref = linalg.spsolve(A.tocsc(), b)
    u1, it1, _ = solvers.solve_linear(lp, tol=1e-11, method="cg")
    sysm = solvers.DeviceSystem(lp.dmesh, lp.data.ptr)
    lmax = sysm._lambda_max()
    assert lmax == pytest.approx((abs(A) @ numpy.ones(lp.n) / abs(A.diagonal())).max(), rel=1e-12)
    for degree in (2, 4):
        u, it, relres = solvers.solve_linear(lp, tol=1e-11, method=f"cg-cheb{degree}")
        assert relres <= 1e-11
        assert numpy.abs(u - u_ref).max() <= 1e-8 * numpy.abs(u_ref).max()
        assert it * degree <= 1.5 * it1 + 8 and it <= 0.75 * it1 + 2  # fewer, fatter iterations
        u2, it2, _ = solvers.solve_linear(lp, tol=1e-11, method=f"cg-cheb{degree}")
        assert it2 == it and u2.tobytes() == u.tobytes()
