"""Symbolic front end: problem object -> per-half-edge / per-vertex expressions.

This is the host-side half of ``pyfvm.discretize_linear`` [README:49] and
``pyfvm.discretize`` [README:110] ("the Jacobian is computed symbolically"
[README:126]).  It is O(1) in the mesh size.  The numeric half (the hot path)
lives in ``csrc/``; ``codegen.py`` prints what this module produces as device
functors.

Half-edge view: an edge record (i0 -> i1) contributes
``expr(uk0, uk1, x0, x1)`` to row i0 and the *turned* expression
``expr(uk1, uk0, x1, x0)`` to row i1.  The device therefore needs only the
row-0 expressions and evaluates them per half-edge (row, col) with
``x0 = x[row], x1 = x[col], uk0 = u[row], uk1 = u[col]``.
"""
from dataclasses import dataclass, field

import sympy

from . import form_language as fl

# --- the symbols every functor is written in -------------------------------
UK0 = sympy.Symbol("uk0")
UK1 = sympy.Symbol("uk1")
CE = sympy.Symbol("edge_ce_ratio")
EL = sympy.Symbol("edge_length")
CV = sympy.Symbol("control_volume")
SA = sympy.Symbol("surface_area")
X0 = [sympy.Symbol(f"x0_{k}") for k in range(3)]
X1 = [sympy.Symbol(f"x1_{k}") for k in range(3)]
XV = [sympy.Symbol(f"x_{k}") for k in range(3)]

_x = sympy.DeferredVector("x")
_XSYM = [_x[k] for k in range(3)]
_NSYM = [fl.n[k] for k in range(3)]


def split(expr, variables):
    """Affine part, linear coefficients and nonlinear remainder of ``expr``
    with respect to ``variables``."""
    expr = sympy.sympify(expr)
    single = not isinstance(variables, (list, tuple))
    if single:
        variables = [variables]
    expr = expr.expand()
    affine = expr
    for var in variables:
        affine = affine.coeff(var, 0)
    linear = []
    for var in variables:
        d = sympy.diff(expr, var)
        for var2 in variables:
            d = d.coeff(var2, 0)
        linear.append(d)
    nonlinear = expr - affine
    for var, c in zip(variables, linear):
        nonlinear -= var * c
    # (expr is expanded: a linear integrand cancels term by term; simplify -- ~15 ms per
    # call -- is only needed to recognise a remainder that does not)
    nonlinear = sympy.expand(nonlinear)
    if nonlinear != 0:
        nonlinear = sympy.simplify(nonlinear)
    return affine, (linear[0] if single else linear), nonlinear


def _at_point(expr, u, pt, uk):
    """``expr`` with ``u(x) -> uk`` and ``x -> pt`` (no normal allowed)."""
    out = expr.subs(u(_x), uk)
    return out.subs(dict(zip(_XSYM, pt)), simultaneous=True)


def lower_edge_integrand(integrand, u, dim):
    """Discretise a ``dS`` integrand on the edge x0 -> x1, row of vertex 0.

    ``n_dot_grad(g) -> (g(x1) - g(x0))/|e|``; bare ``u(x) -> (uk0+uk1)/2``;
    ``x -> (x0+x1)/2``; ``n -> (x1-x0)/|e|``; the whole expression is
    multiplied by ``edge_ce_ratio*edge_length`` (the dual-face measure) and
    simplified."""
    expr = sympy.sympify(integrand(_x))
    # normal derivatives first (their arguments are evaluated at the endpoints)
    for node in list(expr.atoms(fl.n_dot_grad)):
        (g,) = node.args
        repl = (_at_point(g, u, X1, UK1) - _at_point(g, u, X0, UK0)) / EL
        expr = expr.xreplace({node: repl})
    for node in list(expr.atoms(fl.n_dot)):
        repl = sum(a * nk for a, nk in zip(node.args, _NSYM))
        expr = expr.xreplace({node: repl})
    half = sympy.Rational(1, 2)
    expr = expr.subs(u(_x), half * (UK0 + UK1))
    sub = {}
    for k in range(3):
        sub[_XSYM[k]] = half * (X0[k] + X1[k])
        sub[_NSYM[k]] = (X1[k] - X0[k]) / EL
    expr = expr.subs(sub, simultaneous=True)
    expr = CE * EL * expr
    expr = _flatten_dim(expr, dim)
    return _simplify_cached(expr)


_SIMPLIFIED = {}


def _simplify_cached(expr):
    """``sympy.simplify`` dominates the lowering (~0.1 s per dS integrand); the same
    problem is usually discretised many times (every mesh, every Newton restart),
    so the result is memoised on the expression's structure."""
    key = sympy.srepr(expr)
    out = _SIMPLIFIED.get(key)
    if out is None:
        out = _SIMPLIFIED[key] = sympy.simplify(expr)
    return out


def _flatten_dim(expr, dim):
    """Meshes stored with ``dim`` < 3 coordinate columns have the missing
    components identically 0.0; dropping them symbolically gives the same
    floating-point values (adding/multiplying exact zeros)."""
    expr = sympy.sympify(expr)
    if dim >= 3:
        return expr
    zero = {s: 0 for k in range(dim, 3) for s in (X0[k], X1[k], XV[k])}
    return expr.subs(zero)


def lower_vertex_integrand(integrand, u, dim):
    """``dV``: ``integrand(x)`` with ``u(x) -> uk0``, times ``control_volume``."""
    fx = sympy.sympify(integrand(_x))
    expr = fx.subs(u(_x), UK0).subs(dict(zip(_XSYM, XV)), simultaneous=True)
    return _flatten_dim(expr * CV, dim)


def lower_face_integrand(integrand, u, dim):
    """``dGamma``: a boundary integral is evaluated per boundary vertex like a ``dV`` term,
    weighted with the vertex's share of the domain boundary (``mesh.surface_areas``)
    instead of its control volume: ``integrand(x)`` with ``u(x) -> uk0``, times
    ``surface_area`` (0 off the boundary).  Neumann and Robin conditions."""
    fx = sympy.sympify(integrand(_x))
    if fx.atoms(fl.n_dot_grad) or fx.atoms(fl.n_dot):
        raise NotImplementedError("n / n_dot_grad inside a dGamma integrand (give the flux itself)")
    expr = fx.subs(u(_x), UK0).subs(dict(zip(_XSYM, XV)), simultaneous=True)
    return _flatten_dim(expr * SA, dim)


def lower_dirichlet(f, u, dim):
    fx = sympy.sympify(f(_x))
    expr = fx.subs(u(_x), UK0).subs(dict(zip(_XSYM, XV)), simultaneous=True)
    return _flatten_dim(expr, dim)


@dataclass
class EdgeTerm:
    """One (dS integral, subdomain) instance. ``outputs`` is a dict of
    expressions: linear -> diag/off/aff; nonlinear -> val/diag/off."""

    outputs: dict
    subdomain: object = None


@dataclass
class VertexTerm:
    outputs: dict
    subdomain: object = None


@dataclass
class DirichletTerm:
    outputs: dict
    subdomain: object = None


@dataclass
class Lowered:
    kind: str  # "linear" | "nonlinear"
    dim: int
    edge: list = field(default_factory=list)
    vertex: list = field(default_factory=list)
    dirichlet: list = field(default_factory=list)


def _subdomain_list(subdomains):
    # one kernel application per subdomain; none given = everywhere
    return list(subdomains) if subdomains else [None]


def _vertex_expr(integral, u, dim):
    """Per-vertex expression of a ``dV`` (control volume) or ``dGamma`` (boundary surface)
    integral; both end up in the vertex functor."""
    if isinstance(integral.measure, fl.CellSurface):
        return lower_face_integrand(integral.integrand, u, dim)
    return lower_vertex_integrand(integral.integrand, u, dim)


def lower_linear(obj, dim):
    """Front end of ``discretize_linear`` [README:49]: every integral is split
    into a linear coefficient per unknown and an affine part; a nonlinear
    remainder is an error."""
    u = sympy.Function("u")
    res = obj.apply(u)
    out = Lowered("linear", dim)
    for integral in res.integrals:
        if isinstance(integral.measure, fl.ControlVolumeSurface):
            expr = lower_edge_integrand(integral.integrand, u, dim)
            affine, linear, nonlinear = split(expr, [UK0, UK1])
            assert nonlinear == 0, f"dS integrand is not linear in u: {nonlinear}"
            for sd in _subdomain_list(integral.subdomains):
                out.edge.append(
                    EdgeTerm({"diag": linear[0], "off": linear[1], "aff": affine}, sd)
                )
        else:
            expr = _vertex_expr(integral, u, dim)
            affine, linear, nonlinear = split(expr, UK0)
            assert nonlinear == 0, f"dV / dGamma integrand is not linear in u: {nonlinear}"
            for sd in _subdomain_list(integral.subdomains):
                out.vertex.append(VertexTerm({"lin": linear, "aff": affine}, sd))
    dirichlet = getattr(obj, "dirichlet", None)
    if callable(dirichlet):
        for f, sd in dirichlet(u):
            expr = lower_dirichlet(f, u, dim)
            affine, coeff, nonlinear = split(expr, UK0)
            assert nonlinear == 0, f"Dirichlet condition is not linear in u: {nonlinear}"
            out.dirichlet.append(DirichletTerm({"coeff": coeff, "rhs": -affine}, sd))
    return out


def lower_nonlinear(obj, dim):
    """Front end of ``discretize`` [README:110]: residual expressions and their
    symbolic derivatives (the Jacobian, [README:126])."""
    u = sympy.Function("u")
    res = obj.apply(u)
    out = Lowered("nonlinear", dim)
    for integral in res.integrals:
        if isinstance(integral.measure, fl.ControlVolumeSurface):
            expr = lower_edge_integrand(integral.integrand, u, dim)
            outs = {
                "val": expr,
                "diag": sympy.diff(expr, UK0),
                "off": sympy.diff(expr, UK1),
            }
            for sd in _subdomain_list(integral.subdomains):
                out.edge.append(EdgeTerm(outs, sd))
        else:
            expr = _vertex_expr(integral, u, dim)
            outs = {"val": expr, "lin": sympy.diff(expr, UK0)}
            for sd in _subdomain_list(integral.subdomains):
                out.vertex.append(VertexTerm(outs, sd))
    dirichlet = getattr(obj, "dirichlet", None)
    if callable(dirichlet):
        for f, sd in dirichlet(u):
            expr = lower_dirichlet(f, u, dim)
            out.dirichlet.append(
                DirichletTerm({"val": expr, "coeff": sympy.diff(expr, UK0)}, sd)
            )
    return out
