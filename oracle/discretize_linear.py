"""Oracle restatement of ``pyfvm.discretize_linear`` [README:49] (SURVEY §2,
App. A.2-A.3).  Test infrastructure only; see ``oracle/__init__.py``."""
import numpy
import sympy

from . import form_language as fl
from .edge_integral import DiscretizeEdgeIntegral
from .linear_fvm_problem import get_linear_fvm_problem

a2a = [{"ImmutableMatrix": numpy.array}, "numpy"]


def split(expr, variables):
    """Affine, linear and nonlinear part of ``expr`` w.r.t. ``variables``."""
    if isinstance(expr, float):
        return expr, 0, 0
    input_is_list = True
    if not isinstance(variables, list):
        input_is_list = False
        variables = [variables]
    expr = expr.expand()
    affine = expr
    for var in variables:
        affine = affine.coeff(var, n=0)
    linear = []
    for var in variables:
        d = sympy.diff(expr, var)
        for var2 in variables:
            d = d.coeff(var2, n=0)
        linear.append(d)
    nonlinear = expr - affine
    for var, coeff in zip(variables, linear):
        nonlinear -= var * coeff
    nonlinear = sympy.simplify(nonlinear)
    if not input_is_list:
        assert len(linear) == 1
        linear = linear[0]
    return affine, linear, nonlinear


def _coords_first(X):
    # (..., 3) -> (3, ...): the lambdified kernels index components first
    return numpy.moveaxis(X, -1, 0)


class EdgeLinearKernel:
    def __init__(self, linear, affine, subdomains):
        self.linear = linear
        self.affine = affine
        self.subdomains = subdomains

    def eval(self, mesh, cell_mask):
        edge_ce_ratio = mesh.ce_ratios[..., cell_mask]
        edge_length = numpy.sqrt(mesh.ei_dot_ei[..., cell_mask])
        nec = mesh.idx_hierarchy[..., cell_mask]
        X = mesh.node_coords[nec]
        x0, x1 = _coords_first(X[0]), _coords_first(X[1])
        val = self.linear(x0, x1, edge_ce_ratio, edge_length)
        rhs = self.affine(x0, x1, edge_ce_ratio, edge_length)
        # constants come back as python scalars: broadcast
        ones = numpy.ones(nec.shape[1:])
        rhs = [rhs[i] * ones for i in (0, 1)]
        val = [[val[i][j] * ones for j in (0, 1)] for i in (0, 1)]
        return val, rhs, nec


class VertexLinearKernel:
    def __init__(self, mesh, linear, affine, subdomains):
        self.mesh = mesh
        self.linear = linear
        self.affine = affine
        self.subdomains = subdomains

    def eval(self, vertex_mask):
        control_volumes = self.mesh.control_volumes[vertex_mask]
        X = self.mesh.node_coords[vertex_mask].T
        ones = numpy.ones(len(control_volumes))
        return (
            self.linear(control_volumes, X) * ones,
            self.affine(control_volumes, X) * ones,
        )


class BoundaryLinearKernel:
    """``dGamma`` term: evaluated per vertex like a ``dV`` term, weighted with the vertex's
    share of the domain boundary (``mesh.surface_areas``, 0 off the boundary)."""

    def __init__(self, mesh, linear, affine, subdomains):
        self.mesh = mesh
        self.linear = linear
        self.affine = affine
        self.subdomains = subdomains

    def eval(self, vertex_mask):
        surface_areas = self.mesh.surface_areas[vertex_mask]
        X = self.mesh.node_coords[vertex_mask].T
        ones = numpy.ones(len(surface_areas))
        return (
            self.linear(surface_areas, X) * ones,
            self.affine(surface_areas, X) * ones,
        )


class DirichletLinearKernel:
    def __init__(self, mesh, coeff, rhs, subdomain):
        self.mesh = mesh
        self.coeff = coeff
        self.rhs = rhs
        self.subdomain = subdomain

    def eval(self, vertex_mask):
        X = self.mesh.node_coords[vertex_mask].T
        zero = numpy.zeros(int(numpy.sum(vertex_mask)))
        return self.coeff(X) + zero, self.rhs(X) + zero


def _sd_list(subdomains):
    return list(subdomains) if subdomains else [None]


def discretize_linear(obj, mesh, scales=None):
    """``pyfvm.discretize_linear(problem, mesh) -> (matrix, rhs)`` [README:49]."""
    return get_linear_fvm_problem(mesh, *lower(obj, mesh), scales=scales)


def lower(obj, mesh):
    """Symbolic half (O(1) in the mesh size): problem -> lambdified kernels.
    Split out so the numeric half can be timed on its own (bench.py)."""
    u = sympy.Function("u")
    res = obj.apply(u)
    edge_kernels, vertex_kernels, dirichlet_kernels = [], [], []
    for integral in res.integrals:
        if isinstance(integral.measure, fl.ControlVolumeSurface):
            x0 = sympy.DeferredVector("x0")
            x1 = sympy.DeferredVector("x1")
            el = sympy.Symbol("edge_length")
            er = sympy.Symbol("edge_ce_ratio")
            expr, index_vars = DiscretizeEdgeIntegral(x0, x1, el, er).generate(
                integral.integrand, [u]
            )
            expr = sympy.simplify(expr)
            uk0, uk1 = index_vars[0]
            affine0, linear0, nonlinear = split(expr, [uk0, uk1])
            assert nonlinear == 0
            # turn the edge around: the row of the other endpoint
            swap = {uk0: uk1, uk1: uk0}
            for k in range(3):
                swap[x0[k]] = x1[k]
                swap[x1[k]] = x0[k]
            expr_turned = expr.subs(swap, simultaneous=True)
            affine1, linear1, nonlinear = split(expr_turned, [uk0, uk1])
            assert nonlinear == 0
            linear = [[linear0[0], linear0[1]], [linear1[0], linear1[1]]]
            affine = [affine0, affine1]
            l_eval = sympy.lambdify((x0, x1, er, el), linear, modules=a2a)
            a_eval = sympy.lambdify((x0, x1, er, el), affine, modules=a2a)
            edge_kernels.append(
                EdgeLinearKernel(l_eval, a_eval, _sd_list(integral.subdomains))
            )
        elif isinstance(integral.measure, fl.ControlVolume):
            x = sympy.DeferredVector("x")
            fx = integral.integrand(x)
            uk0 = sympy.Symbol("uk0")
            try:
                expr = fx.subs(u(x), uk0)
            except AttributeError:  # 'float' object has no attribute 'subs'
                expr = fx
            control_volume = sympy.Symbol("control_volume")
            expr *= control_volume
            affine, linear, nonlinear = split(expr, uk0)
            assert nonlinear == 0
            l_eval = sympy.lambdify((control_volume, x), linear, modules=a2a)
            a_eval = sympy.lambdify((control_volume, x), affine, modules=a2a)
            vertex_kernels.append(
                VertexLinearKernel(mesh, l_eval, a_eval, _sd_list(integral.subdomains))
            )
        elif isinstance(integral.measure, fl.CellSurface):
            # dGamma: like dV with surface_area in place of control_volume
            x = sympy.DeferredVector("x")
            fx = integral.integrand(x)
            uk0 = sympy.Symbol("uk0")
            try:
                expr = fx.subs(u(x), uk0)
            except AttributeError:
                expr = fx
            surface_area = sympy.Symbol("surface_area")
            expr *= surface_area
            affine, linear, nonlinear = split(expr, uk0)
            assert nonlinear == 0
            l_eval = sympy.lambdify((surface_area, x), linear, modules=a2a)
            a_eval = sympy.lambdify((surface_area, x), affine, modules=a2a)
            vertex_kernels.append(
                BoundaryLinearKernel(mesh, l_eval, a_eval, _sd_list(integral.subdomains))
            )
        else:
            raise NotImplementedError(f"unknown measure {integral.measure}")

    dirichlet = getattr(obj, "dirichlet", None)
    if callable(dirichlet):
        u = sympy.Function("u")
        x = sympy.DeferredVector("x")
        for f, subdomain in dirichlet(u):
            uk0 = sympy.Symbol("uk0")
            try:
                expr = f(x).subs(u(x), uk0)
            except AttributeError:
                expr = f(x)
            affine, coeff, nonlinear = split(expr, uk0)
            assert nonlinear == 0
            rhs = -affine
            coeff_eval = sympy.lambdify((x,), coeff, modules=a2a)
            rhs_eval = sympy.lambdify((x,), rhs, modules=a2a)
            dirichlet_kernels.append(
                DirichletLinearKernel(mesh, coeff_eval, rhs_eval, subdomain)
            )
    return edge_kernels, vertex_kernels, dirichlet_kernels
