"""Convergence-order tests on manufactured solutions -- the strategy of the
reference's own (absent) test suite: solve on a sequence of refined
meshzoo-style meshes, measure the control-volume-weighted 2-norm and the max
norm of the error, expect second order.  They pin the *discretisation* (signs,
coefficients, Dirichlet handling of the lowering) to analytic answers, for the
CPU oracle (CPU tier) and for the engine (GPU tier)."""
import numpy
import pytest
import sympy
from scipy.sparse import linalg

from . import problems

PI = sympy.pi


def _exact(dim):
    def u(x):
        r = sympy.sin(PI * x[0]) * sympy.sin(PI * x[1])
        return r * sympy.sin(PI * x[2]) if dim == 3 else r

    return u


def manufactured(fl, dim, kind):
    """-lap(u) [+ a.grad(u)] [+ u] = f with u = prod sin(pi x_k) and u = 0 on the boundary."""
    uex = _exact(dim)
    a = sympy.Matrix([2, 1, 0])

    def source(x):
        ue = uex(x)
        f = dim * PI**2 * ue
        if kind == "convection":
            f = f + sum(a[k] * sympy.diff(uex(sympy.symbols("s0 s1 s2")), sympy.symbols("s0 s1 s2")[k]).subs(
                dict(zip(sympy.symbols("s0 s1 s2"), [x[0], x[1], x[2]]))) for k in range(2))
        if kind == "reaction":
            f = f + ue
        return f

    class P:
        def apply(self, u):
            edge = lambda x: -fl.n_dot_grad(u(x))
            if kind == "convection":
                edge = lambda x: -fl.n_dot_grad(u(x)) + fl.dot(a.T, fl.n) * u(x)
            out = fl.integrate(edge, fl.dS) - fl.integrate(source, fl.dV)
            if kind == "reaction":
                out = out + fl.integrate(lambda x: u(x), fl.dV)
            return out

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    return P()


def manufactured_robin(fl):
    """-lap(u) = f on the unit square with the ROBIN condition du/dn + u = g on the whole
    boundary (a dGamma term, no Dirichlet rows); u = sin(pi x) sin(pi y), for which
    g = -pi (sin(pi x) + sin(pi y)) on every side."""

    class P:
        def apply(self, u):
            return (
                fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS)
                + fl.integrate(
                    lambda x: u(x) + PI * (sympy.sin(PI * x[0]) + sympy.sin(PI * x[1])), fl.dGamma
                )
                - fl.integrate(lambda x: 2 * PI**2 * _exact(2)(x), fl.dV)
            )

    return P()


def _mesh(dim, n):
    from pyfvm_b200 import meshes

    xs = numpy.linspace(0.0, 1.0, n)
    if dim == 2:
        return meshes.Mesh(*meshes.rectangle_tri(xs, xs))
    return meshes.Mesh(*meshes.cube_tetra(xs, xs, xs))


def _orders(discretize_linear, fl, dim, kind, sizes):
    errs2, errsi, hs = [], [], []
    for n in sizes:
        mesh = _mesh(dim, n)
        prob = manufactured_robin(fl) if kind == "robin" else manufactured(fl, dim, kind)
        A, b = discretize_linear(prob, mesh)
        u = linalg.spsolve(A.tocsc(), b)
        X = mesh.node_coords.T
        exact = numpy.prod(numpy.sin(numpy.pi * X[:dim]), axis=0)
        e = u - exact
        errs2.append(numpy.sqrt(numpy.sum(mesh.control_volumes * e * e)))
        errsi.append(numpy.abs(e).max())
        hs.append(1.0 / (n - 1))
    o2 = numpy.log(numpy.array(errs2[:-1]) / errs2[1:]) / numpy.log(numpy.array(hs[:-1]) / hs[1:])
    oi = numpy.log(numpy.array(errsi[:-1]) / errsi[1:]) / numpy.log(numpy.array(hs[:-1]) / hs[1:])
    return o2, oi, errs2


@pytest.mark.parametrize("kind", ["poisson", "convection", "reaction", "robin"])
def test_oracle_second_order_2d(kind):
    import oracle

    o2, oi, errs = _orders(oracle.discretize_linear, oracle.form_language, 2, kind, [9, 17, 33])
    assert (o2 > 1.85).all() and (oi > 1.8).all(), (o2, oi)
    assert errs[-1] < 2e-3


def test_oracle_second_order_3d():
    import oracle

    o2, oi, _ = _orders(oracle.discretize_linear, oracle.form_language, 3, "poisson", [5, 9, 17])
    assert (o2 > 1.8).all() and (oi > 1.7).all(), (o2, oi)


@pytest.mark.gpu
@pytest.mark.parametrize(
    "dim,kind,sizes",
    [(2, "poisson", [17, 33, 65, 129]), (2, "convection", [17, 33, 65, 129]),
     (2, "reaction", [17, 33, 65]), (3, "poisson", [5, 9, 17, 33]), (2, "robin", [17, 33, 65])],
)
def test_engine_second_order(dim, kind, sizes):
    import pyfvm_b200 as eng

    o2, oi, _ = _orders(eng.discretize_linear, eng.form_language, dim, kind, sizes)
    assert (o2 > 1.85).all() and (oi > 1.75).all(), (o2, oi)
