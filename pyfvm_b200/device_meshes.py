"""Meshes whose per-record arrays only ever exist in HBM (SURVEY §8f rank 1:
the step *before* the hot path, moved to the device).

meshplex derives ``idx_hierarchy``, ``ce_ratios``, ``ei_dot_ei`` and
``control_volumes`` with numpy on the host; at BASELINE config 5
those arrays are ~38 GB and their PCIe upload would dwarf the assembly.  Here

* ``device_rectangle_tri`` / ``device_cube_tetra`` generate the points and cells
  of the meshzoo grids [README:44-46] on the device (``fvm_gen_*``),
* ``DeviceSimplexMesh`` wraps any (points, cells) pair: ``fvm_geometry`` computes
  the per-record endpoints, ce_ratios and edge lengths in HBM, the control
  volumes come out of the pattern build (deterministic row sums), and -- for
  triangles -- the boundary vertices are read off the pattern.

The object exposes the part of the meshplex attribute contract that stays cheap
on the host (``points``, ``node_coords``, ``is_boundary_point``,
``get_vertex_mask``); ``engine.DeviceMesh`` takes the rest from
``device_records()``.
"""
import numpy

from .engine import Engine


class DeviceSimplexMesh:
    def __init__(self, points, cells_dev, n_cells, cell_dim, index_bytes=4, boundary=None,
                 eng=None, points_dev=None):
        self.eng = eng or Engine.get()
        self.points = numpy.ascontiguousarray(points, dtype=numpy.float64)
        nc = numpy.zeros((len(self.points), 3))
        nc[:, : self.points.shape[1]] = self.points
        self.node_coords = nc
        self.cell_dim = cell_dim
        self.n_cells = int(n_cells)
        self.n_records = (3 if cell_dim == 2 else 12) * self.n_cells
        self.index_bytes = index_bytes
        self._cells = cells_dev
        self._points_dev = points_dev
        self._boundary = boundary
        self._rec = None

    @classmethod
    def from_host(cls, points, cells, boundary=None, eng=None):
        """Any simplex mesh given as host arrays (points (n, 2|3), cells (c, 3|4))."""
        eng = eng or Engine.get()
        cells = numpy.ascontiguousarray(cells, dtype=numpy.int32)
        return cls(points, eng.to_device(cells), len(cells), cells.shape[1] - 1, 4,
                   boundary=boundary, eng=eng)

    # -- device side ------------------------------------------------------------
    def device_records(self, eng, need_el=True):
        """Per-record endpoints / ce_ratios / edge lengths, computed in HBM."""
        if self._rec is None:
            assert eng is self.eng
            R, ib = self.n_records, self.index_bytes
            pd = self._points_dev or eng.to_device(self.points)
            bufs = {
                "idx0": eng.alloc(ib * R), "idx1": eng.alloc(ib * R),
                "ce": eng.alloc(8 * R), "el": eng.alloc(8 * R),
            }
            eng._ck(
                eng.lib.fvm_geometry(
                    eng.ctx, self.cell_dim, self.points.shape[1], self.n_cells, self._cells.ptr, ib,
                    pd.ptr, bufs["idx0"].ptr, bufs["idx1"].ptr, bufs["ce"].ptr, bufs["el"].ptr,
                )
            )
            eng.sync()
            self._bufs = bufs
            self._rec = {k: b.ptr for k, b in bufs.items()}
            self._rec.update(n_records=R, index_bytes=ib)
        return self._rec

    def release_device_records(self):
        """Free the record arrays (the device mesh keeps its own half-edge streams)."""
        if self._rec is not None:
            for b in self._bufs.values():
                b.free()
            self._rec = self._bufs = None

    def records_to_host(self):
        """(idx0, idx1, ce_ratios, edge_length) as host arrays -- tests / inspection."""
        rec = self.device_records(self.eng)
        it = numpy.int32 if self.index_bytes == 4 else numpy.int64
        out = []
        for key, dt in (("idx0", it), ("idx1", it), ("ce", numpy.float64), ("el", numpy.float64)):
            a = numpy.empty(self.n_records, dtype=dt)
            self._bufs[key].download(a)
            out.append(a)
        return out

    # -- host side of the attribute contract ---------------------------------------
    @property
    def is_boundary_point(self):
        """meshplex's ``is_boundary_point``: vertices of facets that belong to exactly one
        cell, found on the device by a sort of the facet keys (``fvm_boundary_vertices``)
        unless the flags were supplied (structured grids know them analytically)."""
        if self._boundary is None:
            eng, n = self.eng, len(self.points)
            out = eng.alloc(n + 16)
            eng._ck(
                eng.lib.fvm_boundary_vertices(
                    eng.ctx, self.cell_dim, self.n_cells, self._cells.ptr, self.index_bytes, n, out.ptr
                )
            )
            flags = numpy.empty(n, dtype=numpy.uint8)
            out.download(flags)
            out.free()
            self._boundary = flags.astype(bool)
        return self._boundary

    def device_record_mask(self, eng, terms):
        """Per-record subdomain bits of the dS terms, built in HBM (``fvm_record_mask``).
        ``terms``: list of ``(subdomain or None, bit value)``; a record's cell is inside a
        subdomain when all its vertices are (meshplex's ``get_cell_mask``).  Returns a
        device buffer of ``n_records`` bytes carrying a host-side ``digest`` (cache key)."""
        import hashlib

        assert eng is self.eng
        R = self.n_records
        mask = eng.alloc(R + 64)
        eng._ck(eng.lib.fvm_dev_memset(eng.ctx, mask.ptr, 0, R + 64))
        h = hashlib.sha256()
        for subdomain, bit in terms:
            inside_dev = None
            if subdomain is not None:
                if getattr(subdomain, "is_boundary_only", False):
                    inside = numpy.zeros(len(self.points), dtype=numpy.uint8)  # no cell is "boundary only"
                else:
                    inside = numpy.ascontiguousarray(
                        numpy.asarray(subdomain.is_inside(self.node_coords.T), dtype=bool).astype(numpy.uint8)
                    )
                h.update(inside.tobytes())
                inside_dev = eng.to_device(inside, pad=16)
            h.update(bytes([bit]))
            eng._ck(
                eng.lib.fvm_record_mask(
                    eng.ctx, self.cell_dim, self.n_cells, self._cells.ptr, self.index_bytes,
                    None if inside_dev is None else inside_dev.ptr, bit, mask.ptr,
                )
            )
        eng.sync()
        mask.digest = h.hexdigest()
        return mask

    def get_vertex_mask(self, subdomain=None):
        if subdomain is None:
            return numpy.s_[:]
        mask = numpy.asarray(subdomain.is_inside(self.node_coords.T), dtype=bool)
        if getattr(subdomain, "is_boundary_only", False):
            mask = mask & self.is_boundary_point
        return mask


def _grid_boundary(shape):
    """Vertices on the faces of a tensor grid (x fastest)."""
    n = int(numpy.prod(shape))
    lid = numpy.arange(n)
    on = numpy.zeros(n, dtype=bool)
    stride = 1
    for m in shape:
        k = (lid // stride) % m
        on |= (k == 0) | (k == m - 1)
        stride *= m
    return on


def device_rectangle_tri(xs, ys, parity=0, eng=None, boundary=None):
    """``meshzoo.rectangle_tri(xs, ys)`` [README:44-46] generated on the device."""
    eng = eng or Engine.get()
    xs = numpy.ascontiguousarray(xs, dtype=numpy.float64)
    ys = numpy.ascontiguousarray(ys, dtype=numpy.float64)
    nx, ny = len(xs), len(ys)
    n, nc = nx * ny, 2 * (nx - 1) * (ny - 1)
    pts, cells = eng.alloc(16 * n), eng.alloc(4 * 3 * nc)
    eng._ck(
        eng.lib.fvm_gen_rectangle_tri(
            eng.ctx, nx, ny, parity, xs.ctypes.data, ys.ctypes.data, 4, pts.ptr, cells.ptr
        )
    )
    points = numpy.empty((n, 2))
    pts.download(points)
    if boundary is None:
        boundary = _grid_boundary((nx, ny))
    return DeviceSimplexMesh(points, cells, nc, 2, 4, boundary=boundary, eng=eng, points_dev=pts)


def device_cube_tetra(xs, ys, zs, parity=0, eng=None, boundary=None):
    """``meshzoo.cube_tetra``-style 5-tetrahedra grid generated on the device."""
    eng = eng or Engine.get()
    xs, ys, zs = (numpy.ascontiguousarray(a, dtype=numpy.float64) for a in (xs, ys, zs))
    nx, ny, nz = len(xs), len(ys), len(zs)
    n, nc = nx * ny * nz, 5 * (nx - 1) * (ny - 1) * (nz - 1)
    pts, cells = eng.alloc(24 * n), eng.alloc(4 * 4 * nc)
    eng._ck(
        eng.lib.fvm_gen_cube_tetra(
            eng.ctx, nx, ny, nz, parity, xs.ctypes.data, ys.ctypes.data, zs.ctypes.data, 4,
            pts.ptr, cells.ptr,
        )
    )
    points = numpy.empty((n, 3))
    pts.download(points)
    if boundary is None:
        boundary = _grid_boundary((nx, ny, nz))
    return DeviceSimplexMesh(points, cells, nc, 3, 4, boundary=boundary, eng=eng, points_dev=pts)
