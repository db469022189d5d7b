"""Oracle restatement of ``pyfvm.discretize`` [README:110], ``FvmProblem.eval``
[README:121, 136] and ``Jacobian.get_linear_operator`` [README:116]
(SURVEY §2, App. A).  Test infrastructure only; see ``oracle/__init__.py``."""
import numpy
import sympy
from scipy import sparse

from . import form_language as fl
from .discretize_linear import _coords_first, _sd_list, a2a
from .edge_integral import DiscretizeEdgeIntegral
from .linear_fvm_problem import apply_dirichlet_rows


class EdgeKernel:
    def __init__(self, val, subdomains):
        self.val = val
        self.subdomains = subdomains

    def eval(self, u, mesh, cell_mask):
        nec = mesh.idx_hierarchy[..., cell_mask]
        X = mesh.node_coords[nec]
        x0, x1 = _coords_first(X[0]), _coords_first(X[1])
        edge_ce_ratio = mesh.ce_ratios[..., cell_mask]
        edge_length = numpy.sqrt(mesh.ei_dot_ei[..., cell_mask])
        zero = numpy.zeros(nec.shape[1:])
        out = self.val(u[nec[0]], u[nec[1]], x0, x1, edge_ce_ratio, edge_length)
        # out is a (nested) list of arrays or scalars: broadcast each leaf
        return _broadcast(out, zero)


def _broadcast(out, zero):
    if isinstance(out, (list, tuple)):
        return [_broadcast(o, zero) for o in out]
    return out + zero


class VertexKernel:
    def __init__(self, val, subdomains):
        self.val = val
        self.subdomains = subdomains

    def eval(self, u, mesh, vertex_mask):
        control_volumes = mesh.control_volumes[vertex_mask]
        X = mesh.node_coords[vertex_mask].T
        zero = numpy.zeros(len(control_volumes))
        return self.val(u, control_volumes, X) + zero


class BoundaryKernel:
    """``dGamma`` term (see ``discretize_linear.BoundaryLinearKernel``)."""

    def __init__(self, val, subdomains):
        self.val = val
        self.subdomains = subdomains

    def eval(self, u, mesh, vertex_mask):
        surface_areas = mesh.surface_areas[vertex_mask]
        X = mesh.node_coords[vertex_mask].T
        zero = numpy.zeros(len(surface_areas))
        return self.val(u, surface_areas, X) + zero


class DirichletKernel:
    def __init__(self, val, subdomain):
        self.val = val
        self.subdomain = subdomain

    def eval(self, u, mesh, vertex_mask):
        assert len(u) == int(numpy.sum(vertex_mask))
        X = mesh.node_coords[vertex_mask].T
        zero = numpy.zeros(len(u))
        return self.val(u, X) + zero


class FvmProblem:
    def __init__(self, mesh, edge_kernels, vertex_kernels, dirichlets):
        self.mesh = mesh
        self.edge_kernels = edge_kernels
        self.vertex_kernels = vertex_kernels
        self.dirichlets = dirichlets

    def eval(self, u, scales=None):
        """``scales`` (test-only): dict receiving sum |contribution| per row."""
        out = numpy.zeros_like(u)
        mag = numpy.zeros_like(u) if scales is not None else None
        for edge_kernel in self.edge_kernels:
            for subdomain in edge_kernel.subdomains:
                cell_mask = self.mesh.get_cell_mask(subdomain)
                nec = self.mesh.idx_hierarchy[..., cell_mask]
                vals = numpy.array(edge_kernel.eval(u, self.mesh, cell_mask))
                numpy.add.at(out, nec, vals)
                if mag is not None:
                    numpy.add.at(mag, nec, numpy.abs(vals))
        for vertex_kernel in self.vertex_kernels:
            for subdomain in vertex_kernel.subdomains:
                vertex_mask = self.mesh.get_vertex_mask(subdomain)
                vals = vertex_kernel.eval(u[vertex_mask], self.mesh, vertex_mask)
                out[vertex_mask] += vals
                if mag is not None:
                    mag[vertex_mask] += numpy.abs(vals)
        if scales is not None:
            scales["vec"] = mag
        for dirichlet in self.dirichlets:
            vertex_mask = self.mesh.get_vertex_mask(dirichlet.subdomain)
            out[vertex_mask] = dirichlet.eval(u[vertex_mask], self.mesh, vertex_mask)
        return out


class Jacobian:
    def __init__(self, mesh, edge_kernels, vertex_kernels, dirichlets):
        self.mesh = mesh
        self.edge_kernels = edge_kernels
        self.vertex_kernels = vertex_kernels
        self.dirichlets = dirichlets

    def get_linear_operator(self, u, scales=None):
        mesh = self.mesh
        n = len(mesh.node_coords)
        V, I, J = [], [], []
        for edge_kernel in self.edge_kernels:
            for subdomain in edge_kernel.subdomains:
                cell_mask = mesh.get_cell_mask(subdomain)
                v_mtx = edge_kernel.eval(u, mesh, cell_mask)
                nec = mesh.idx_hierarchy[..., cell_mask]
                V += [v_mtx[0][0].flatten(), v_mtx[0][1].flatten()]
                V += [v_mtx[1][0].flatten(), v_mtx[1][1].flatten()]
                I += [nec[0].flatten(), nec[0].flatten(), nec[1].flatten(), nec[1].flatten()]
                J += [nec[0].flatten(), nec[1].flatten(), nec[0].flatten(), nec[1].flatten()]
        for vertex_kernel in self.vertex_kernels:
            for subdomain in vertex_kernel.subdomains:
                vertex_mask = mesh.get_vertex_mask(subdomain)
                vals = vertex_kernel.eval(u[vertex_mask], mesh, vertex_mask)
                verts = numpy.arange(n)[vertex_mask]
                V.append(vals)
                I.append(verts)
                J.append(verts)
        V, I, J = numpy.concatenate(V), numpy.concatenate(I), numpy.concatenate(J)
        matrix = sparse.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()
        if scales is not None:  # test-only
            scales["matrix"] = sparse.coo_matrix((numpy.abs(V), (I, J)), shape=(n, n)).tocsr()
        apply_dirichlet_rows(
            matrix,
            mesh,
            self.dirichlets,
            lambda d, mask: (d.eval(u[mask], mesh, mask), None),
        )
        return matrix


def discretize(obj, mesh):
    u = sympy.Function("u")
    res = obj.apply(u)
    edge_kernels, vertex_kernels = [], []
    jac_edge_kernels, jac_vertex_kernels = [], []
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
            swap = {uk0: uk1, uk1: uk0}
            for k in range(3):
                swap[x0[k]] = x1[k]
                swap[x1[k]] = x0[k]
            expr_turned = expr.subs(swap, simultaneous=True)
            args = (uk0, uk1, x0, x1, er, el)
            sds = _sd_list(integral.subdomains)
            edge_kernels.append(
                EdgeKernel(sympy.lambdify(args, [expr, expr_turned], modules=a2a), sds)
            )
            lin0 = [sympy.diff(expr, var) for var in (uk0, uk1)]
            lin1 = [sympy.diff(expr_turned, var) for var in (uk0, uk1)]
            jac_edge_kernels.append(
                EdgeKernel(sympy.lambdify(args, [lin0, lin1], modules=a2a), sds)
            )
        elif isinstance(integral.measure, fl.ControlVolume):
            x = sympy.DeferredVector("x")
            fx = integral.integrand(x)
            uk0 = sympy.Symbol("uk0")
            try:
                expr = fx.subs(u(x), uk0)
            except AttributeError:
                expr = sympy.sympify(fx)
            control_volume = sympy.Symbol("control_volume")
            expr *= control_volume
            sds = _sd_list(integral.subdomains)
            args = (uk0, control_volume, x)
            vertex_kernels.append(VertexKernel(sympy.lambdify(args, expr, modules=a2a), sds))
            jac_vertex_kernels.append(
                VertexKernel(sympy.lambdify(args, sympy.diff(expr, uk0), modules=a2a), sds)
            )
        elif isinstance(integral.measure, fl.CellSurface):
            x = sympy.DeferredVector("x")
            fx = integral.integrand(x)
            uk0 = sympy.Symbol("uk0")
            try:
                expr = fx.subs(u(x), uk0)
            except AttributeError:
                expr = sympy.sympify(fx)
            surface_area = sympy.Symbol("surface_area")
            expr *= surface_area
            sds = _sd_list(integral.subdomains)
            args = (uk0, surface_area, x)
            vertex_kernels.append(BoundaryKernel(sympy.lambdify(args, expr, modules=a2a), sds))
            jac_vertex_kernels.append(
                BoundaryKernel(sympy.lambdify(args, sympy.diff(expr, uk0), modules=a2a), sds)
            )
        else:
            raise NotImplementedError(f"unknown measure {integral.measure}")

    dirichlets, jac_dirichlets = [], []
    dirichlet = getattr(obj, "dirichlet", None)
    if callable(dirichlet):
        x = sympy.DeferredVector("x")
        for f, subdomain in dirichlet(u):
            uk0 = sympy.Symbol("uk0")
            try:
                expr = f(x).subs(u(x), uk0)
            except AttributeError:
                expr = sympy.sympify(f(x))
            dirichlets.append(
                DirichletKernel(sympy.lambdify((uk0, x), expr, modules=a2a), subdomain)
            )
            jac_dirichlets.append(
                DirichletKernel(
                    sympy.lambdify((uk0, x), sympy.diff(expr, uk0), modules=a2a), subdomain
                )
            )
    residual = FvmProblem(mesh, edge_kernels, vertex_kernels, dirichlets)
    jacobian = Jacobian(mesh, jac_edge_kernels, jac_vertex_kernels, jac_dirichlets)
    return residual, jacobian
