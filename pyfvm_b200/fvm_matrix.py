"""``pyfvm.fvm_matrix``: matrices from USER-WRITTEN numeric kernels instead of symbolic
integrands (SURVEY §2 "pyfvm.fvm_matrix / EdgeMatrixKernel", §8f rank 3; the reference's
Ginzburg-Landau-type examples reach it through README:79 "more examples").

A user edge kernel is an object with ``subdomains`` (list; ``[None]`` = everywhere) and
``eval(mesh, cell_mask)`` returning the 2x2 block of every record as a numpy array of
shape ``(2, 2) + mesh.idx_hierarchy[..., cell_mask].shape[1:]``; vertex kernels return
one diagonal value per masked vertex, Dirichlet kernels the diagonal coefficient of the
rows they replace.  The user's ``eval`` is arbitrary Python and runs on the host; what
the reference then does with the blocks -- COO append, ``tocsr()`` duplicate summation,
Dirichlet row replacement -- runs on the device through the same tile kernel as the
symbolic path: the blocks are permuted into the half-edge streams
(``fvm_mesh_set_halfedge_values``) and folded per slot / per row in the same order.

Real-valued kernels only (the tile kernel is fp64).
"""
import numpy

from . import codegen
from .engine import DeviceMesh, Engine, _mesh_arrays, _cell_dim
from .problem import _KEEP_BIT


class EdgeMatrixKernel:
    """Base class of user edge kernels: set ``self.subdomains`` and define
    ``eval(self, mesh, cell_mask)``."""

    def __init__(self):
        self.subdomains = [None]

    def eval(self, mesh, cell_mask):  # pragma: no cover - interface
        raise NotImplementedError


class VertexMatrixKernel:
    """User vertex kernel: ``eval(self, mesh, vertex_mask)`` -> diagonal values."""

    def __init__(self):
        self.subdomains = [None]


class DirichletMatrixKernel:
    """User Dirichlet kernel: rows of ``subdomain`` become ``coeff * e_row``;
    ``eval(self, mesh, vertex_mask)`` -> coefficients."""

    def __init__(self, subdomain):
        self.subdomain = subdomain


_NUMERIC_PRELUDE = """// pyfvm_b200.fvm_matrix: user-evaluated blocks ride the geometry streams
#define FVM_MODE 2
#define FVM_DIM {dim}
#define FVM_FLAGS {flags}u
#define FVM_CONSUMER_WARPS {consumer_warps}
#define FVM_STAGES {stages}
#define FVM_PRODUCERS {producers}
#define FVM_EDGE_FACTORED 0
#define FVM_EDGE_USES_X 0
#define FVM_EDGE_USES_U 0
#define FVM_VERT_USES_X 0
#define FVM_VERT_USES_U 0
#define FVM_HAS_VERTEX {has_vertex}

__device__ __forceinline__ void fvm_edge(const double uk0, const double uk1,
    const double x0_0, const double x0_1, const double x0_2,
    const double x1_0, const double x1_1, const double x1_2,
    const double edge_ce_ratio, const double edge_length, const unsigned mask,
    double& D, double& O, double& C) {{
  D = edge_ce_ratio;  // block[side][side] of the record, in half-edge order
  O = edge_length;    // block[side][other]
}}

__device__ __forceinline__ void fvm_vertex(const double uk0, const double control_volume,
    const double surface_area, const double x_0, const double x_1, const double x_2,
    const unsigned mask, double& L, double& C) {{
  L = control_volume;  // summed user vertex-kernel values
}}

__device__ __forceinline__ void fvm_dirichlet(const int id, const double uk0,
    const double control_volume, const double surface_area,
    const double x_0, const double x_1, const double x_2,
    double& coeff, double& val) {{
  coeff = surface_area;  // user Dirichlet coefficient of the row
}}
"""


def _numeric_functors(dim, cell_dim, has_vertex, has_dirichlet, masked):
    tune = codegen.tuning(cell_dim)
    flags = codegen.F_EDGE_EL | codegen.F_EDGE_D
    flags |= codegen.F_CV if has_vertex else 0
    flags |= (codegen.F_DIR | codegen.F_SA) if has_dirichlet else 0
    src = _NUMERIC_PRELUDE.format(
        dim=dim, flags=flags, consumer_warps=tune["consumer_warps"], stages=tune["stages"],
        producers=tune["producers"], has_vertex=int(has_vertex),
    )
    return codegen.Functors(
        source=src, mode=codegen.MODE_JACOBIAN, flags=flags,
        consumer_warps=tune["consumer_warps"], stages=tune["stages"],
        ctas_per_sm=tune["ctas_per_sm"], producers=tune["producers"], factored=False,
        uses_el=True, uses_x=False, edge_masked=masked, vert_masked=False,
        has_dirichlet=has_dirichlet, needs_u=False,
    )


def get_fvm_matrix(mesh, edge_kernels=None, vertex_kernels=None, face_kernels=None,
                   dirichlets=None, device=None):
    """``pyfvm.get_fvm_matrix(mesh, edge_kernels, vertex_kernels, face_kernels, dirichlets)
    -> scipy CSR``.  ``face_kernels`` behave like vertex kernels (per boundary vertex)."""
    from scipy import sparse

    eng = Engine.get(device)
    coords, dim = _mesh_arrays(mesh)
    n = coords.shape[0]
    idx = numpy.asarray(mesh.idx_hierarchy)
    rshape = idx.shape[1:]
    R = int(numpy.prod(rshape))
    blocks = numpy.zeros((2, 2) + rshape)
    rec_mask = numpy.zeros(rshape, dtype=numpy.uint8)
    everywhere = False
    for kernel in edge_kernels or []:
        for subdomain in kernel.subdomains:
            cell_mask = mesh.get_cell_mask(subdomain)
            vals = numpy.asarray(kernel.eval(mesh, cell_mask))
            if numpy.iscomplexobj(vals):
                raise NotImplementedError("complex-valued edge kernels (the tile kernel is fp64)")
            blocks[..., cell_mask] += vals
            rec_mask[..., cell_mask] |= _KEEP_BIT
            everywhere = everywhere or subdomain is None
    vdiag = numpy.zeros(n)
    has_vertex = False
    for kernel in list(vertex_kernels or []) + list(face_kernels or []):
        for subdomain in kernel.subdomains:
            vertex_mask = mesh.get_vertex_mask(subdomain)
            vdiag[vertex_mask] += kernel.eval(mesh, vertex_mask)
            has_vertex = True
    dcoeff = numpy.zeros(n)
    dirid = numpy.zeros(n, dtype=numpy.uint8)
    for d in dirichlets or []:
        vertex_mask = mesh.get_vertex_mask(d.subdomain)
        dcoeff[vertex_mask] = d.eval(mesh, vertex_mask)
        dirid[vertex_mask] = 1
    has_dir = bool(dirichlets)

    masked = not everywhere
    dm = DeviceMesh(
        eng, mesh, rec_mask=numpy.ascontiguousarray(rec_mask.reshape(-1)) if masked else None,
        need_el=True,
    )
    # side-major source values: the row-of-idx0 view of every record, then the row-of-idx1 view
    dm.set_halfedge_values(
        numpy.concatenate([blocks[0, 0].reshape(-1), blocks[1, 1].reshape(-1)]),
        numpy.concatenate([blocks[0, 1].reshape(-1), blocks[1, 0].reshape(-1)]),
    )
    if has_vertex:
        dm.update_geometry(cv=vdiag)  # the vertex functor reads this stream as L
    if has_dir:
        dm.set_surface_areas(dcoeff)  # ... and the Dirichlet functor this one as coeff
    fun = _numeric_functors(dim, _cell_dim(mesh), has_vertex, has_dir, masked)
    kernel = eng.kernel(fun)
    data = eng.alloc(8 * dm.nnz + 16)
    dir_dev = eng.to_device(dirid, pad=16) if has_dir else None
    dm.launch(kernel, dirid=None if dir_dev is None else dir_dev.ptr, data=data.ptr)
    values = numpy.empty(dm.nnz)
    data.download(values)
    indptr, indices = dm.pattern()
    m = sparse.csr_matrix((values, indices, indptr), shape=(n, n), copy=False)
    m.has_sorted_indices = True
    m.has_canonical_format = True
    return m
