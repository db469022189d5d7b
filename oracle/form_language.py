"""Oracle restatement of ``pyfvm.form_language`` (names per [README:32, 88]).
Test infrastructure only; see ``oracle/__init__.py``."""
import numpy
import sympy


class Measure:
    pass


class ControlVolume(Measure):
    pass


class ControlVolumeSurface(Measure):
    pass


class CellSurface(Measure):
    pass


dV = ControlVolume()
dS = ControlVolumeSurface()
dGamma = CellSurface()


class Subdomain:
    is_boundary_only = False


class Boundary(Subdomain):
    is_boundary_only = True

    def is_inside(self, x):
        return numpy.ones(x.shape[1], dtype=bool)


class n_dot_grad(sympy.Function):
    pass


class n_dot(sympy.Function):
    pass


n = sympy.DeferredVector("n")


def dot(a, b):
    def comp(v, k):
        try:
            if isinstance(v, sympy.MatrixBase) and k >= len(v):
                return 0
            return v[k]
        except IndexError:
            return 0

    return comp(a, 0) * comp(b, 0) + comp(a, 1) * comp(b, 1) + comp(a, 2) * comp(b, 2)


class Integral:
    def __init__(self, integrand, measure, subdomains):
        self.integrand = integrand
        self.measure = measure
        self.subdomains = subdomains


class IntegralSum:
    def __init__(self, integrand, measure, subdomains):
        self.integrals = [Integral(integrand, measure, subdomains)]

    def __add__(self, other):
        out = IntegralSum.__new__(IntegralSum)
        out.integrals = self.integrals + other.integrals
        return out

    def __sub__(self, other):
        assert isinstance(other, IntegralSum)
        out = IntegralSum.__new__(IntegralSum)
        out.integrals = self.integrals + [
            Integral((lambda f: lambda x: -f(x))(i.integrand), i.measure, i.subdomains)
            for i in other.integrals
        ]
        return out

    def __neg__(self):
        out = IntegralSum.__new__(IntegralSum)
        out.integrals = [
            Integral((lambda f: lambda x: -f(x))(i.integrand), i.measure, i.subdomains)
            for i in self.integrals
        ]
        return out

    def __mul__(self, a):
        assert isinstance(a, (int, float))
        out = IntegralSum.__new__(IntegralSum)
        out.integrals = [
            Integral((lambda f: lambda x: a * f(x))(i.integrand), i.measure, i.subdomains)
            for i in self.integrals
        ]
        return out

    __rmul__ = __mul__


def integrate(integrand, measure, subdomains=None):
    assert isinstance(measure, Measure)
    if subdomains is None:
        subdomains = set()
    elif not isinstance(subdomains, set):
        try:
            subdomains = set(subdomains)
        except TypeError:
            subdomains = {subdomains}
    return IntegralSum(integrand, measure, subdomains)
