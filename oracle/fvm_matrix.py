"""Oracle restatement of ``pyfvm.fvm_matrix.get_fvm_matrix`` (user-supplied
numeric edge kernels assembled by the same COO path as the linear assembler): COO blocks
00, 01, 10, 11 per edge kernel and subdomain, diagonal blocks per vertex / face kernel,
``tocsr()`` duplicate summation, Dirichlet rows zeroed and their diagonal set.
Test infrastructure only; see ``oracle/__init__.py``."""
import numpy
from scipy import sparse


class EdgeMatrixKernel:
    def __init__(self):
        self.subdomains = [None]


def get_fvm_matrix(mesh, edge_kernels=None, vertex_kernels=None, face_kernels=None,
                   dirichlets=None, scales=None):
    V, I, J = [], [], []
    n = len(mesh.node_coords)
    for kernel in edge_kernels or []:
        for subdomain in kernel.subdomains:
            cell_mask = mesh.get_cell_mask(subdomain)
            v_mtx = kernel.eval(mesh, cell_mask)
            nec = mesh.idx_hierarchy[..., cell_mask]
            V += [v_mtx[0][0].flatten(), v_mtx[0][1].flatten()]
            V += [v_mtx[1][0].flatten(), v_mtx[1][1].flatten()]
            I += [nec[0].flatten(), nec[0].flatten(), nec[1].flatten(), nec[1].flatten()]
            J += [nec[0].flatten(), nec[1].flatten(), nec[0].flatten(), nec[1].flatten()]
    for kernel in list(vertex_kernels or []) + list(face_kernels or []):
        for subdomain in kernel.subdomains:
            vertex_mask = mesh.get_vertex_mask(subdomain)
            vals = kernel.eval(mesh, vertex_mask)
            verts = numpy.arange(n)[vertex_mask]
            V.append(vals + numpy.zeros(len(verts)))
            I.append(verts)
            J.append(verts)
    V, I, J = numpy.concatenate(V), numpy.concatenate(I), numpy.concatenate(J)
    matrix = sparse.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()
    if scales is not None:
        scales["matrix"] = sparse.coo_matrix((numpy.abs(V), (I, J)), shape=(n, n)).tocsr()
    d = matrix.diagonal()
    for dirichlet in dirichlets or []:
        vertex_mask = mesh.get_vertex_mask(dirichlet.subdomain)
        for i in numpy.where(vertex_mask)[0]:
            matrix.data[matrix.indptr[i] : matrix.indptr[i + 1]] = 0.0
        d[vertex_mask] = dirichlet.eval(mesh, vertex_mask)
    if dirichlets:
        matrix.setdiag(d)
    return matrix
