"""Form language of the drop-in: ``integrate``, ``dS``, ``dV``, ``dGamma``,
``n_dot_grad``, ``n_dot``, ``n``, ``dot``, ``Subdomain``, ``Boundary``.

Mirrors the names the reference exports to users
(``from pyfvm.form_language import Boundary, dS, dV, integrate, n_dot_grad``
[README:32]; star-import [README:88]).  Everything here is host-side symbolic
bookkeeping; nothing touches the device.

A problem's ``apply(u)`` returns an :class:`IntegralSum`; the lowering in
``pyfvm_b200.lowering`` turns every term into a device functor.
"""
import numpy
import sympy

__all__ = [
    "integrate",
    "dS",
    "dV",
    "dGamma",
    "n_dot_grad",
    "n_dot",
    "n",
    "dot",
    "Subdomain",
    "Boundary",
    "FvmProblem",
]


class Measure:
    pass


class ControlVolume(Measure):
    """``dV``: the control volume of a vertex -> vertex functor."""


class ControlVolumeSurface(Measure):
    """``dS``: the surface of a control volume -> edge functor."""


class CellSurface(Measure):
    """``dGamma``: the domain boundary (Neumann / Robin terms) -> evaluated per boundary
    vertex with ``mesh.surface_areas`` as the weight (vertex functor)."""


dV = ControlVolume()
dS = ControlVolumeSurface()
dGamma = CellSurface()


class Subdomain:
    """User subdomains define ``is_inside(x)`` on a ``(dim, k)`` coordinate
    array and may set ``is_boundary_only``."""

    is_boundary_only = False


class Boundary(Subdomain):
    """All boundary vertices (``u = 0 on Boundary()`` in [README:40, 102])."""

    is_boundary_only = True

    def is_inside(self, x):
        return numpy.ones(x.shape[1], dtype=bool)


class FvmProblem:
    """Optional base class for user problems (duck typing is enough)."""


class n_dot_grad(sympy.Function):
    """Normal derivative across the dual face: ``(f(x1) - f(x0)) / |e|``."""


class n_dot(sympy.Function):
    """``n_dot(a0, a1, a2)`` = ``n . a`` with the edge direction as normal."""


# The dual-face unit normal n = (x1 - x0)/|e|; components are the symbols
# "n[0]", "n[1]", "n[2]" and are substituted by the edge lowering.
n = sympy.DeferredVector("n")


def dot(a, b):
    """Euclidean product of two 3-vectors given as anything indexable
    (``sympy.Matrix``, its transpose, list, ``n``, ``x``)."""

    def comp(v, k):
        if isinstance(v, sympy.MatrixBase):
            return v[k] if k < v.shape[0] * v.shape[1] else 0
        if isinstance(v, (list, tuple, numpy.ndarray)):
            return v[k] if k < len(v) else 0
        return v[k]

    return sum(comp(a, k) * comp(b, k) for k in range(3))


class Integral:
    def __init__(self, integrand, measure, subdomains):
        assert isinstance(measure, Measure)
        if subdomains is None:
            subdomains = set()
        elif not isinstance(subdomains, set):
            try:
                subdomains = set(subdomains)
            except TypeError:
                subdomains = {subdomains}
        self.integrand = integrand
        self.measure = measure
        self.subdomains = subdomains


def _scaled(integrand, a):
    return lambda x: a * integrand(x)


class IntegralSum:
    """Formal sum of integrals supporting ``+``, ``-``, unary ``-`` and scalar
    multiplication [README:37, 97-99]."""

    def __init__(self, integrals):
        self.integrals = list(integrals)

    def __add__(self, other):
        assert isinstance(other, IntegralSum)
        return IntegralSum(self.integrals + other.integrals)

    def __neg__(self):
        return IntegralSum(
            Integral(_scaled(i.integrand, -1), i.measure, i.subdomains)
            for i in self.integrals
        )

    def __sub__(self, other):
        assert isinstance(other, IntegralSum)
        return self + (-other)

    def __mul__(self, a):
        assert not isinstance(a, IntegralSum)
        return IntegralSum(
            Integral(_scaled(i.integrand, a), i.measure, i.subdomains)
            for i in self.integrals
        )

    __rmul__ = __mul__


def integrate(integrand, measure, subdomains=None):
    """One-term :class:`IntegralSum` of ``integrand`` (a callable of ``x``
    returning a sympy expression or a plain float [README:37]) over
    ``measure``, optionally restricted to ``subdomains``."""
    return IntegralSum([Integral(integrand, measure, subdomains)])
