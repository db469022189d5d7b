"""Oracle restatement of pyfvm's linear assembler (``get_linear_fvm_problem``,
``_get_VIJ``): COO triplets -> ``tocsr()`` (duplicates
summed, explicit zeros kept) -> Dirichlet row replacement.
Test infrastructure only; see ``oracle/__init__.py``."""
import numpy
from scipy import sparse


def _get_VIJ(mesh, edge_kernels, vertex_kernels, rhs_abs=None):
    V, I, J = [], [], []
    n = len(mesh.node_coords)
    rhs_V = numpy.zeros(n)
    for edge_kernel in edge_kernels:
        for subdomain in edge_kernel.subdomains:
            cell_mask = mesh.get_cell_mask(subdomain)
            v_mtx, v_rhs, nec = edge_kernel.eval(mesh, cell_mask)
            # diagonal entries, then off-diagonal entries
            V += [v_mtx[0][0].flatten(), v_mtx[0][1].flatten()]
            V += [v_mtx[1][0].flatten(), v_mtx[1][1].flatten()]
            I += [nec[0].flatten(), nec[0].flatten(), nec[1].flatten(), nec[1].flatten()]
            J += [nec[0].flatten(), nec[1].flatten(), nec[0].flatten(), nec[1].flatten()]
            numpy.subtract.at(rhs_V, nec[0], v_rhs[0])
            numpy.subtract.at(rhs_V, nec[1], v_rhs[1])
            if rhs_abs is not None:  # test-only: magnitude of what was summed
                numpy.add.at(rhs_abs, nec[0], numpy.abs(v_rhs[0]))
                numpy.add.at(rhs_abs, nec[1], numpy.abs(v_rhs[1]))
    for vertex_kernel in vertex_kernels:
        for subdomain in vertex_kernel.subdomains:
            vertex_mask = mesh.get_vertex_mask(subdomain)
            vals_matrix, vals_rhs = vertex_kernel.eval(vertex_mask)
            verts = numpy.arange(n)[vertex_mask]
            V.append(vals_matrix)
            I.append(verts)
            J.append(verts)
            rhs_V[verts] -= vals_rhs
            if rhs_abs is not None:
                rhs_abs[verts] += numpy.abs(vals_rhs)
    return numpy.concatenate(V), numpy.concatenate(I), numpy.concatenate(J), rhs_V


def apply_dirichlet_rows(matrix, mesh, dirichlets, evaluate):
    """Zero every Dirichlet row's stored slice, then write the diagonal back
    (``diagonal()`` read before zeroing, overwritten at masked rows, ``setdiag``).
    ``evaluate(dirichlet, vertex_mask) -> (coeff, rhs_or_None)``."""
    d = matrix.diagonal()
    rhs_updates = []
    for dirichlet in dirichlets:
        vertex_mask = mesh.get_vertex_mask(dirichlet.subdomain)
        for i in numpy.where(vertex_mask)[0]:
            matrix.data[matrix.indptr[i] : matrix.indptr[i + 1]] = 0.0
        coeff, rhs = evaluate(dirichlet, vertex_mask)
        d[vertex_mask] = coeff
        rhs_updates.append((vertex_mask, rhs))
    matrix.setdiag(d)
    return rhs_updates


def get_linear_fvm_problem(mesh, edge_kernels, vertex_kernels, dirichlets, scales=None):
    """``scales`` (test-only): a dict that receives the per-slot / per-row sums
    of |contribution| the parity tolerance is relative to."""
    n = len(mesh.node_coords)
    rhs_abs = numpy.zeros(n) if scales is not None else None
    V, I, J, rhs = _get_VIJ(mesh, edge_kernels, vertex_kernels, rhs_abs)
    matrix = sparse.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()
    if scales is not None:
        scales["matrix"] = sparse.coo_matrix((numpy.abs(V), (I, J)), shape=(n, n)).tocsr()
        scales["rhs"] = rhs_abs
    updates = apply_dirichlet_rows(
        matrix, mesh, dirichlets, lambda d, mask: d.eval(mask)
    )
    for vertex_mask, vals in updates:
        rhs[vertex_mask] = vals
    return matrix, rhs
