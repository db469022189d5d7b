"""Numeric half of the drop-in: what ``discretize_linear`` / ``discretize``
hand back, backed by the device engine.

Reference surface reproduced here:
* ``pyfvm.discretize_linear(problem, mesh) -> (matrix, rhs)``        [README:49]
* ``pyfvm.discretize(problem, mesh) -> (f, jacobian)``               [README:110]
* ``f.eval(u) -> ndarray``                                           [README:121, 136]
* ``jacobian.get_linear_operator(u0) -> scipy CSR``                  [README:116]
"""
import concurrent.futures

import numpy
from scipy import sparse

from . import codegen, lowering
from .engine import Engine, device_mesh, _mesh_arrays, _cell_dim

_KEEP_BIT = 0x80  # record is covered by at least one edge term without subdomain


def _edge_bits(mesh, lowered):
    """Per-record subdomain bits of the edge terms (None: all terms everywhere)."""
    bits, rec_mask, nbit = [], None, 0
    if all(t.subdomain is None for t in lowered.edge):
        return [None] * len(lowered.edge), None
    if hasattr(mesh, "device_records"):
        # cells live in HBM: the per-record bits are built there too
        terms = []
        for t in lowered.edge:
            if t.subdomain is None:
                bits.append(None)
                terms.append((None, _KEEP_BIT))
            else:
                if nbit >= 7:
                    raise NotImplementedError("more than 7 subdomain-restricted dS terms")
                bits.append(nbit)
                terms.append((t.subdomain, 1 << nbit))
                nbit += 1
        return bits, mesh.device_record_mask(Engine.get(), terms)
    idx = numpy.asarray(mesh.idx_hierarchy)
    shape = idx.shape[1:]
    rec_mask = numpy.zeros(shape, dtype=numpy.uint8)
    for t in lowered.edge:
        if t.subdomain is None:
            bits.append(None)
            rec_mask |= _KEEP_BIT
        else:
            if nbit >= 7:
                raise NotImplementedError("more than 7 subdomain-restricted dS terms")
            cell_mask = numpy.asarray(mesh.get_cell_mask(t.subdomain), dtype=bool)
            rec_mask[..., cell_mask] |= 1 << nbit
            bits.append(nbit)
            nbit += 1
    return bits, numpy.ascontiguousarray(rec_mask.reshape(-1))


def _vertex_term_bits(lowered):
    """Bit number of every vertex term's subdomain (None: everywhere) -- what the code
    generator needs; the masks themselves (`_vertex_bits`) are only needed at launch."""
    bits, nbit = [], 0
    for t in lowered.vertex:
        if t.subdomain is None:
            bits.append(None)
            continue
        if nbit >= 8:
            raise NotImplementedError("more than 8 subdomain-restricted dV terms")
        bits.append(nbit)
        nbit += 1
    return bits


def _vertex_bits(mesh, lowered, n):
    bits, vmask, nbit = [], None, 0
    for t in lowered.vertex:
        if t.subdomain is None:
            bits.append(None)
            continue
        if vmask is None:
            vmask = numpy.zeros(n, dtype=numpy.uint8)
        if nbit >= 8:
            raise NotImplementedError("more than 8 subdomain-restricted dV terms")
        vmask[mesh.get_vertex_mask(t.subdomain)] |= 1 << nbit
        bits.append(nbit)
        nbit += 1
    return bits, vmask


def _dirichlet_ids(mesh, lowered, n):
    """1-based index of the last Dirichlet term covering each vertex (later
    conditions overwrite earlier ones, as in the reference's loop)."""
    if not lowered.dirichlet:
        return None
    if len(lowered.dirichlet) > 255:
        raise NotImplementedError("more than 255 Dirichlet conditions")
    ids = numpy.zeros(n, dtype=numpy.uint8)
    for k, t in enumerate(lowered.dirichlet):
        ids[mesh.get_vertex_mask(t.subdomain)] = k + 1
    return ids


class _DeviceProblem:
    """State shared by the linear problem, the residual and the Jacobian of one
    (problem, mesh) pair: device mesh, vertex masks, compiled kernels, buffers."""

    def __init__(self, obj, mesh, kind, device=None, fmad=False, row_range=None):
        self.eng = eng = Engine.get(device)
        coords, dim = _mesh_arrays(mesh)
        self.n = n = coords.shape[0]
        lower = lowering.lower_linear if kind == "linear" else lowering.lower_nonlinear
        self.lowered = lowered = lower(obj, dim)
        ebits, rec_mask = _edge_bits(mesh, lowered)
        vbits = _vertex_term_bits(lowered)
        modes = (
            [codegen.MODE_LINEAR]
            if kind == "linear"
            else [codegen.MODE_RESIDUAL, codegen.MODE_JACOBIAN]
        )
        tune = codegen.tuning(_cell_dim(mesh))
        self.functors = {m: codegen.generate(lowered, m, ebits, vbits, tune) for m in modes}
        need_el = any(f.uses_el for f in self.functors.values())
        # row_range = (first owned vertex, owned rows): row-partitioned multi-GPU runs.
        # The device mesh (upload + pattern build: one long GIL-free library call) is made on
        # a helper thread while this one evaluates the user's subdomain predicates on the
        # vertices (numpy, ~15 ms per 16 M vertices); their uploads wait for the join --
        # they share the engine's bounce buffers with the mesh upload.
        # (Device-resident meshes answer those predicates with library calls of their own on
        # the same context: no overlap for them.)
        if hasattr(mesh, "device_records"):
            self.dmesh = device_mesh(eng, mesh, rec_mask=rec_mask, need_el=need_el, row_range=row_range)
            _, vmask = _vertex_bits(mesh, lowered, n)
            dirid = _dirichlet_ids(mesh, lowered, n)
        else:
            with concurrent.futures.ThreadPoolExecutor(1) as pool:
                fut = pool.submit(
                    device_mesh, eng, mesh, rec_mask=rec_mask, need_el=need_el, row_range=row_range
                )
                try:
                    _, vmask = _vertex_bits(mesh, lowered, n)
                    dirid = _dirichlet_ids(mesh, lowered, n)
                finally:
                    concurrent.futures.wait([fut])
                self.dmesh = fut.result()
        if any(f.uses_sa for f in self.functors.values()) and not self.dmesh.has_sa:
            if not hasattr(mesh, "surface_areas"):
                raise NotImplementedError(
                    "dGamma terms need mesh.surface_areas (the per-vertex boundary measure)"
                )
            self.dmesh.set_surface_areas(mesh.surface_areas)
        self.n_rows = self.dmesh.n_rows
        self.kernels = {m: eng.kernel(f, fmad=fmad) for m, f in self.functors.items()}
        # (+16: the kernel's 16-byte-rounded bulk copies may read past the last element)
        self.vmask = eng.to_device(vmask, pad=16) if vmask is not None else None
        self.dirid = eng.to_device(dirid, pad=16) if dirid is not None else None
        self.nnz = self.dmesh.nnz
        self.data = eng.alloc(8 * self.nnz + 16)
        self.vec = eng.alloc(8 * n + 16)
        self.u = eng.alloc(8 * n + 16)
        self._mesh_args = (mesh, rec_mask, need_el, row_range)

    def _p(self, buf):
        return None if buf is None else buf.ptr

    def launch(self, mode, u_ptr=None, data_ptr=None, vec_ptr=None, halo=None):
        self.dmesh.launch(
            self.kernels[mode],
            vmask=self._p(self.vmask),
            dirid=self._p(self.dirid),
            u=u_ptr,
            data=data_ptr,
            vec=vec_ptr,
            halo=halo,
        )

    def csr(self, data_host):
        """scipy CSR on the cached pattern (sorted, duplicate-free, int32)."""
        indptr, indices = self.dmesh.pattern()
        m = sparse.csr_matrix(
            (data_host, indices, indptr), shape=(self.n_rows, self.n), copy=False
        )
        m.has_sorted_indices = True
        m.has_canonical_format = True
        return m

    def torch_pattern(self, device):
        """(crow_indices, col_indices) as int32 torch tensors on ``device`` (cached)."""
        import torch

        if getattr(self, "_tpat", None) is None or self._tpat[0].device != device:
            indptr, indices = self.dmesh.pattern()
            self._tpat = (torch.from_numpy(indptr).to(device), torch.from_numpy(indices).to(device))
        return self._tpat

    def padded_u(self, ptr):
        """A caller's device array of an ODD number of doubles has no slack for the
        kernel's even-count bulk copies: it goes through the padded internal buffer."""
        if self.n % 2 == 0:
            return ptr
        self.eng._ck(self.eng.lib.fvm_d2d(self.eng.ctx, self.u.ptr, ptr, 8 * self.n))
        return self.u.ptr

    def upload_u(self, u):
        u = numpy.ascontiguousarray(u, dtype=numpy.float64)
        assert u.shape == (self.n,), "u must have one value per vertex"
        self.u.upload(u)
        return self.u.ptr


class LinearProblem(_DeviceProblem):
    """``discretize_linear``'s numeric half: one tile-kernel launch writes the CSR
    values and the right-hand side (Dirichlet rows replaced in the same pass)."""

    def __init__(self, obj, mesh, **kw):
        super().__init__(obj, mesh, "linear", **kw)

    def update_geometry(self, **kw):
        """Re-upload geometry of the same connectivity.  The cached device mesh is shared
        by every problem built on that mesh object, so the first update moves this
        problem onto a private copy (one more pattern build) instead of changing the
        geometry under the others."""
        if self.dmesh.shared:
            from .engine import DeviceMesh

            mesh, rec_mask, need_el, row_range = self._mesh_args
            self.dmesh = DeviceMesh(
                self.eng, mesh, rec_mask=rec_mask, need_el=need_el, row_range=row_range
            )
        self.dmesh.update_geometry(**kw)

    def assemble_device(self):
        """Launch only; results stay in ``self.data`` / ``self.vec`` on the device."""
        self.launch(codegen.MODE_LINEAR, data_ptr=self.data.ptr, vec_ptr=self.vec.ptr)

    def assemble_torch(self, device=None):
        """``(matrix, rhs)`` as a ``torch.sparse_csr_tensor`` and a torch vector on the
        GPU: the device-resident hand-off."""
        import torch

        device = device or torch.device("cuda", self.eng.device)
        values = torch.empty(self.nnz, dtype=torch.float64, device=device)
        rhs = torch.empty(self.n_rows, dtype=torch.float64, device=device)
        torch.cuda.current_stream(device).synchronize()
        self.launch(codegen.MODE_LINEAR, data_ptr=values.data_ptr(), vec_ptr=rhs.data_ptr())
        self.eng.sync()
        crow, col = self.torch_pattern(device)
        return torch.sparse_csr_tensor(crow, col, values, size=(self.n_rows, self.n)), rhs

    def assemble(self, pinned=False):
        self.assemble_device()
        if pinned:
            data = self.eng.pinned_empty(self.nnz, numpy.float64)
            rhs = self.eng.pinned_empty(self.n_rows, numpy.float64)
        else:
            data = self.eng.host_result(self.nnz, numpy.float64)
            rhs = self.eng.host_result(self.n_rows, numpy.float64)
        self.data.download(data)
        self.vec.download(rhs)
        return self.csr(data), rhs


def _cuda_array(x):
    """(device pointer, element count) of an object exposing ``__cuda_array_interface__``
    (a torch / cupy array living on the GPU), or None for host arrays."""
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is None:
        return None
    assert cai["typestr"] in ("<f8", "=f8") and not cai.get("strides"), "contiguous float64 expected"
    n = 1
    for d in cai["shape"]:
        n *= int(d)
    return int(cai["data"][0]), n


def _device_like(x, n):
    """A new float64 device array of length n of the same library as ``x``."""
    import torch  # the only __cuda_array_interface__ producer in this image

    assert isinstance(x, torch.Tensor), "device outputs are returned as torch tensors"
    return torch.empty(n, dtype=torch.float64, device=x.device)


def _sync_producer(x):
    """Work queued by the array's own library must be finished before the engine's
    stream reads it (the engine runs on its own stream)."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            torch.cuda.current_stream(x.device).synchronize()
    except ImportError:
        pass


class FvmProblem:
    """The residual ``f`` returned by ``discretize`` [README:110]."""

    def __init__(self, state):
        self._s = state

    def eval(self, u):
        """``F(u)`` [README:121, 136].  Host array in, host array out -- or, for an
        array that already lives on the GPU (``__cuda_array_interface__``, e.g. a
        torch CUDA tensor), device in, device out: nothing crosses PCIe."""
        s = self._s
        dev = _cuda_array(u)
        if dev is not None:
            ptr, n = dev
            assert n == s.n and ptr % 16 == 0, "u must hold one 16-byte aligned double per vertex"
            _sync_producer(u)
            out = _device_like(u, s.n_rows)
            self.eval_device(s.padded_u(ptr), out.data_ptr())
            s.eng.sync()
            return out
        self.eval_device(s.upload_u(u))
        out = s.eng.host_result(s.n_rows, numpy.float64)
        s.vec.download(out)
        return out

    def eval_device(self, u_ptr, out_ptr=None):
        """Device-resident fast path: ``u_ptr`` / ``out_ptr`` are device addresses."""
        s = self._s
        s.launch(codegen.MODE_RESIDUAL, u_ptr=u_ptr, vec_ptr=out_ptr or s.vec.ptr)


class Jacobian:
    """The Jacobian returned by ``discretize``; values are refreshed in place on
    the fixed sparsity pattern (no COO rebuild, no sort)."""

    def __init__(self, state):
        self._s = state

    def get_linear_operator(self, u):
        """CSR Jacobian at ``u`` usable by ``spsolve`` [README:116-117]."""
        s = self._s
        self.values_device(s.upload_u(u))
        data = s.eng.host_result(s.nnz, numpy.float64)
        s.data.download(data)
        return s.csr(data)

    def values_device(self, u_ptr, out_ptr=None):
        s = self._s
        s.launch(codegen.MODE_JACOBIAN, u_ptr=u_ptr, data_ptr=out_ptr or s.data.ptr)

    def get_linear_operator_torch(self, u):
        """The Jacobian at ``u`` (a torch CUDA tensor) as a ``torch.sparse_csr_tensor`` on
        the same device: pattern and values never leave HBM."""
        import torch

        s = self._s
        ptr, n = _cuda_array(u)
        assert n == s.n and ptr % 16 == 0
        _sync_producer(u)
        values = torch.empty(s.nnz, dtype=torch.float64, device=u.device)
        self.values_device(s.padded_u(ptr), values.data_ptr())
        s.eng.sync()
        crow, col = s.torch_pattern(u.device)
        return torch.sparse_csr_tensor(crow, col, values, size=(s.n_rows, s.n))


def discretize_linear(obj, mesh, device=None, fmad=False):
    """``pyfvm.discretize_linear(problem, mesh) -> (matrix, rhs)`` [README:49]."""
    return LinearProblem(obj, mesh, device=device, fmad=fmad).assemble()


def discretize(obj, mesh, device=None, fmad=False):
    """``pyfvm.discretize(problem, mesh) -> (f, jacobian)`` [README:110]."""
    state = _DeviceProblem(obj, mesh, "nonlinear", device=device, fmad=fmad)
    return FvmProblem(state), Jacobian(state)
