"""Problem classes shared by the oracle and the GPU tests.  Each factory takes
the form-language module to use (``oracle.form_language`` for the CPU oracle,
``pyfvm_b200.form_language`` for the engine), so both sides see the same
user-level problem written against their own front end."""
import numpy
import sympy


def poisson(fl):
    """README Poisson example [README:35-40]."""

    class Poisson:
        def apply(self, u):
            return fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS) - fl.integrate(
                lambda x: 1.0, fl.dV
            )

        def dirichlet(self, u):
            return [(lambda x: u(x) - 0.0, fl.Boundary())]

    return Poisson()


def bratu(fl):
    """README Bratu example [README:95-102]."""

    class Bratu:
        def apply(self, u):
            return fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS) - fl.integrate(
                lambda x: 2.0 * sympy.exp(u(x)), fl.dV
            )

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    return Bratu()


def convection_diffusion(fl):
    """BASELINE config 4: -n.grad(u) + (a.n) u with a = (2, 1, 0), source 1."""

    class ConvectionDiffusion:
        def apply(self, u):
            a = sympy.Matrix([2, 1, 0])
            return fl.integrate(
                lambda x: -fl.n_dot_grad(u(x)) + fl.dot(a.T, fl.n) * u(x), fl.dS
            ) - fl.integrate(lambda x: 1.0, fl.dV)

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    return ConvectionDiffusion()


def reaction_x(fl):
    """Variable coefficients: uses x in the vertex term, edge_length in the edge
    term and an inhomogeneous Dirichlet condition."""

    class Reaction:
        def apply(self, u):
            return (
                fl.integrate(lambda x: -(1.0 + x[0]) * fl.n_dot_grad(u(x)), fl.dS)
                + fl.integrate(lambda x: (2.0 + sympy.sin(x[1])) * u(x), fl.dV)
                - fl.integrate(lambda x: sympy.cos(x[0] * x[1]), fl.dV)
            )

        def dirichlet(self, u):
            return [(lambda x: u(x) - sympy.sin(x[0]), fl.Boundary())]

    return Reaction()


def nonlinear_conv(fl):
    """Nonlinear in the edge term (u^2 flux) and the vertex term; nonlinear Dirichlet."""

    class Burgersish:
        def apply(self, u):
            a = sympy.Matrix([1, 2, 0])
            return (
                fl.integrate(
                    lambda x: -0.1 * fl.n_dot_grad(u(x)) + fl.dot(a, fl.n) * u(x) ** 2, fl.dS
                )
                + fl.integrate(lambda x: u(x) ** 3 - sympy.exp(-x[0]), fl.dV)
            )

        def dirichlet(self, u):
            return [(lambda x: u(x) - 0.25 * x[1], fl.Boundary())]

    return Burgersish()


class LeftHalf:
    """A user subdomain: x < 1 (used for cell- and vertex-restricted terms)."""

    is_boundary_only = False

    def is_inside(self, x):
        return x[0] < 1.0


class LeftBoundary:
    is_boundary_only = True

    def is_inside(self, x):
        return x[0] < 1.0e-12


def subdomain_problem(fl):
    """Edge and vertex terms restricted to a subdomain + two Dirichlet sets."""
    left = LeftHalf()
    lb = LeftBoundary()

    class Sub:
        def apply(self, u):
            return (
                fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS)
                + fl.integrate(lambda x: -2.0 * fl.n_dot_grad(u(x)), fl.dS, subdomains=[left])
                + fl.integrate(lambda x: 3.0 * u(x), fl.dV, subdomains=[left])
                - fl.integrate(lambda x: 1.0, fl.dV)
            )

        def dirichlet(self, u):
            return [(u, fl.Boundary()), (lambda x: u(x) - 1.0, lb)]

    return Sub()


def robin(fl):
    """Poisson with a Robin condition on the whole boundary (a dGamma term, SURVEY §8f
    rank 3) and a Dirichlet condition on the left side."""
    lb = LeftBoundary()

    class Robin:
        def apply(self, u):
            return (
                fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS)
                + fl.integrate(lambda x: 2.0 * u(x) - sympy.cos(x[0]), fl.dGamma)
                - fl.integrate(lambda x: 1.0 + x[1], fl.dV)
            )

        def dirichlet(self, u):
            return [(lambda x: u(x) - 1.0, lb)]

    return Robin()


def neumann_only(fl):
    """Pure Neumann data (an affine dGamma term) restricted to a boundary subdomain, plus
    a reaction term that keeps the system regular; no Dirichlet rows at all."""
    lb = LeftBoundary()

    class Neumann:
        def apply(self, u):
            return (
                fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS)
                + fl.integrate(lambda x: u(x), fl.dV)
                - fl.integrate(lambda x: 3.0 + x[1], fl.dGamma, subdomains=[lb])
            )

    return Neumann()


def radiation(fl):
    """Nonlinear boundary term (u^4-type radiation) for discretize / Newton."""

    class Radiation:
        def apply(self, u):
            return (
                fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS)
                + fl.integrate(lambda x: 0.5 * u(x) ** 4 - sympy.exp(-x[0]), fl.dGamma)
                + fl.integrate(lambda x: u(x) - 1.0, fl.dV)
            )

    return Radiation()


def rectangle(nx, ny, lx=2.0, ly=1.0):
    from pyfvm_b200 import meshes

    return meshes.rectangle_tri(numpy.linspace(0.0, lx, nx), numpy.linspace(0.0, ly, ny))


def cube(n):
    from pyfvm_b200 import meshes

    xs = numpy.linspace(0.0, 1.0, n)
    return meshes.cube_tetra(xs, xs, xs)


def shuffle_numbering(points, cells, seed=0, window=None):
    """The same mesh with its vertices renumbered: a random permutation of all
    vertices (``window=None``) or of blocks of ``window`` consecutive vertices.
    Real meshio / gmsh meshes [README:56-64] have no lexicographic numbering; this is
    what makes a tile's columns fall outside its staged vertex table."""
    n = len(points)
    rng = numpy.random.default_rng(seed)
    if window is None:
        old_of_new = rng.permutation(n)
    else:
        old_of_new = numpy.concatenate(
            [lo + rng.permutation(min(window, n - lo)) for lo in range(0, n, window)]
        )
    new_of_old = numpy.empty(n, dtype=numpy.int64)
    new_of_old[old_of_new] = numpy.arange(n)
    return numpy.ascontiguousarray(points[old_of_new]), new_of_old[cells]
