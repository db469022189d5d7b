"""Minimal host-side mesh provider: the meshplex attribute contract the path
reads plus the two meshzoo generators the configs use
(SURVEY App. C).  meshplex/meshzoo are third-party inputs of the hot path, not
part of it, and are not installable here; any object exposing the same
attributes (e.g. a real ``meshplex.Mesh`` [README:47]) can be passed to
``discretize_linear`` / ``discretize`` instead.

Geometry (numpy, once per mesh, outside the timed path):
* triangles: ``idx_hierarchy`` (2, 3, cells), local edge k opposite local vertex
  k; ``ce_ratios[k] = 1/2 cot(angle at vertex k)``;
* tetrahedra: ``idx_hierarchy`` (2, 3, 4, cells), the 3 edges of face k (opposite
  vertex k) -> 12 records per cell; ``ce_ratios[j, k] = 1/4 cot(alpha_jk) s_k``
  with ``s_k`` the signed distance of the circumcentre from face k;
* ``control_volumes``: scatter of ``ei_dot_ei * ce_ratios / (2 dim)`` over both
  endpoints; boundary vertices: vertices of facets that belong to one cell.
"""
import numpy


class Mesh:
    """``Mesh(points, cells)`` [README:47]: simplex mesh with the lazily
    computed attributes ``idx_hierarchy``, ``ce_ratios``, ``ei_dot_ei``,
    ``control_volumes``, ``is_boundary_point`` and the subdomain mask getters."""

    def __init__(self, points, cells):
        points = numpy.asarray(points, dtype=numpy.float64)
        cells = numpy.asarray(cells)
        assert points.ndim == 2 and cells.ndim == 2
        self.points = points
        self.cells_points = cells
        self.cell_dim = cells.shape[1] - 1
        assert self.cell_dim in (2, 3), "triangle or tetrahedral cells only"
        nc = numpy.zeros((len(points), 3))
        nc[:, : points.shape[1]] = points
        self.node_coords = nc  # upstream carries 3 columns
        self._geom = None
        self._boundary = None
        self._cv = None
        self._sa = None

    # -- index hierarchy ----------------------------------------------------
    @property
    def idx_hierarchy(self):
        if getattr(self, "_idx", None) is None:
            c = self.cells_points.T  # (dim+1, cells)
            if self.cell_dim == 2:
                loc = numpy.array([[1, 2], [2, 0], [0, 1]]).T  # (2, 3)
                self._idx = c[loc]
            else:
                # face k = the three vertices other than k; its edge j is
                # opposite the j-th vertex of the face
                faces = numpy.array([[1, 2, 3], [2, 3, 0], [3, 0, 1], [0, 1, 2]])
                e = numpy.array([[1, 2], [2, 0], [0, 1]])
                loc = faces[:, e]  # (4 faces, 3 edges, 2 ends)
                loc = loc.transpose(2, 1, 0)  # (2, 3, 4)
                self._idx = c[loc]
        return self._idx

    def _compute_geometry(self):
        if self._geom is not None:
            return
        idx = self.idx_hierarchy
        P = self.node_coords
        if self.cell_dim == 2:
            # component-wise (no (cells, 3) temporaries): e_k = P[idx1[k]] - P[idx0[k]]
            ncomp = 3 if numpy.any(P[:, 2]) else 2
            comp = [numpy.ascontiguousarray(P[:, k]) for k in range(ncomp)]
            e = [[comp[k][idx[1, j]] - comp[k][idx[0, j]] for k in range(ncomp)] for j in range(3)]
            ei_dot_ei = numpy.stack([sum(c * c for c in e[j]) for j in range(3)])
            # cot at vertex k = -(e_{k+1} . e_{k+2}) / (2 area)
            if ncomp == 2:
                area2 = numpy.abs(e[0][0] * e[1][1] - e[0][1] * e[1][0])
            else:
                cx = e[0][1] * e[1][2] - e[0][2] * e[1][1]
                cy = e[0][2] * e[1][0] - e[0][0] * e[1][2]
                cz = e[0][0] * e[1][1] - e[0][1] * e[1][0]
                area2 = numpy.sqrt(cx * cx + cy * cy + cz * cz)
            ce = numpy.stack(
                [
                    -0.5 * sum(a * b for a, b in zip(e[(j + 1) % 3], e[(j + 2) % 3])) / area2
                    for j in range(3)
                ]
            )
            self._geom = (ei_dot_ei, ce)
            return
        e = P[idx[1]] - P[idx[0]]  # (3, 4, cells, 3)
        ei_dot_ei = numpy.einsum("...k,...k->...", e, e)
        c = self.cells_points.T
        ce = numpy.empty(ei_dot_ei.shape)
        p = [P[c[k]] for k in range(4)]
        # circumcentre from |x - p0|^2 = |x - pk|^2
        A = numpy.stack([p[1] - p[0], p[2] - p[0], p[3] - p[0]], axis=1)
        b = 0.5 * numpy.einsum("cij,cij->ci", A, A)
        cc = p[0] + numpy.linalg.solve(A, b[..., None])[..., 0]
        faces = [[1, 2, 3], [2, 3, 0], [3, 0, 1], [0, 1, 2]]
        for k in range(4):
            f = faces[k]
            a0, a1, a2 = p[f[0]], p[f[1]], p[f[2]]
            nrm = numpy.cross(a1 - a0, a2 - a0)
            nlen = numpy.sqrt(numpy.einsum("ck,ck->c", nrm, nrm))
            nrm = nrm / nlen[:, None]
            # orient towards vertex k
            sgn = numpy.sign(numpy.einsum("ck,ck->c", p[k] - a0, nrm))
            s_k = numpy.einsum("ck,ck->c", cc - a0, nrm) * sgn
            fe = e[:, k]  # (3, cells, 3): edge j of face k
            fe1 = numpy.roll(fe, -1, axis=0)
            fe2 = numpy.roll(fe, -2, axis=0)
            dotp = numpy.einsum("jck,jck->jc", fe1, fe2)
            cot = -dotp / nlen[None, :]  # nlen = 2*face area
            ce[:, k] = 0.25 * cot * s_k[None, :]
        self._geom = (ei_dot_ei, ce)

    @property
    def ei_dot_ei(self):
        self._compute_geometry()
        return self._geom[0]

    @property
    def ce_ratios(self):
        self._compute_geometry()
        return self._geom[1]

    @property
    def control_volumes(self):
        if self._cv is None:
            idx = self.idx_hierarchy
            share = self.ei_dot_ei * self.ce_ratios / (2.0 * self.cell_dim)
            cv = numpy.zeros(len(self.points))
            numpy.add.at(cv, idx[0].ravel(), share.ravel())
            numpy.add.at(cv, idx[1].ravel(), share.ravel())
            self._cv = cv
        return self._cv

    @property
    def is_boundary_point(self):
        if self._boundary is None:
            c = self.cells_points
            d = self.cell_dim
            # facets: drop one vertex; boundary facets occur once
            facets = numpy.concatenate(
                [numpy.delete(c, k, axis=1) for k in range(d + 1)], axis=0
            )
            facets = numpy.sort(facets, axis=1)
            n = len(self.points)
            assert float(n) ** d < 2.0**63, "facet keys overflow int64"
            key = facets[:, 0].astype(numpy.int64)
            for k in range(1, d):
                key = key * n + facets[:, k]
            # facets that occur exactly once, decoded back to their vertices
            ks = numpy.sort(key)
            once = numpy.ones(len(ks), dtype=bool)
            once[1:] &= ks[1:] != ks[:-1]
            once[:-1] &= ks[:-1] != ks[1:]
            kb = ks[once]
            mask = numpy.zeros(n, dtype=bool)
            for k in range(d):
                mask[kb % n] = True
                kb = kb // n
            self._boundary = mask
        return self._boundary

    @property
    def surface_areas(self):
        """Per-vertex measure of the domain boundary (meshplex's ``surface_areas``): every
        boundary facet is split among its vertices -- a boundary edge of a triangle mesh
        gives half its length to each endpoint, a boundary triangle of a tetrahedral mesh
        its 2-D control-volume shares (``|e|^2 cot / 8`` per edge and endpoint).  The weight
        of ``dGamma`` integrals; 0 for interior vertices."""
        if self._sa is None:
            c, d, n = self.cells_points, self.cell_dim, len(self.points)
            facets = numpy.concatenate([numpy.delete(c, k, axis=1) for k in range(d + 1)], axis=0)
            srt = numpy.sort(facets, axis=1)
            key = srt[:, 0].astype(numpy.int64)
            for k in range(1, d):
                key = key * n + srt[:, k]
            _, inverse, counts = numpy.unique(key, return_inverse=True, return_counts=True)
            bf = facets[counts[inverse] == 1]  # facets that belong to exactly one cell
            P = self.node_coords
            sa = numpy.zeros(n)
            if d == 2:
                half = 0.5 * numpy.sqrt(((P[bf[:, 1]] - P[bf[:, 0]]) ** 2).sum(axis=1))
                numpy.add.at(sa, bf[:, 0], half)
                numpy.add.at(sa, bf[:, 1], half)
            else:
                p = [P[bf[:, k]] for k in range(3)]
                e = [p[(k + 2) % 3] - p[(k + 1) % 3] for k in range(3)]  # edge k opposite vertex k
                area2 = numpy.sqrt((numpy.cross(e[0], e[1]) ** 2).sum(axis=1))  # 2 * face area
                for k in range(3):
                    cot = -(e[(k + 1) % 3] * e[(k + 2) % 3]).sum(axis=1) / area2
                    share = (e[k] ** 2).sum(axis=1) * cot / 8.0
                    numpy.add.at(sa, bf[:, (k + 1) % 3], share)
                    numpy.add.at(sa, bf[:, (k + 2) % 3], share)
            self._sa = sa
        return self._sa

    # -- output ---------------------------------------------------------------
    def write(self, filename, point_data=None):
        """``mesh.write("out.vtk", point_data={"x": x})`` [README:53, 123]: legacy ASCII
        VTK unstructured grid (meshplex delegates to meshio; this stand-in writes
        the one format the README uses)."""
        if not str(filename).endswith(".vtk"):
            raise ValueError("only legacy .vtk output is supported by this mesh stand-in")
        pts, cells = self.node_coords, self.cells_points
        k = cells.shape[1]
        with open(filename, "w") as f:
            f.write("# vtk DataFile Version 3.0\npyfvm_b200 mesh\nASCII\nDATASET UNSTRUCTURED_GRID\n")
            f.write(f"POINTS {len(pts)} double\n")
            numpy.savetxt(f, pts, fmt="%.17g")
            f.write(f"CELLS {len(cells)} {len(cells) * (k + 1)}\n")
            numpy.savetxt(f, numpy.column_stack([numpy.full(len(cells), k), cells]), fmt="%d")
            f.write(f"CELL_TYPES {len(cells)}\n")
            numpy.savetxt(f, numpy.full(len(cells), 5 if k == 3 else 10), fmt="%d")
            if point_data:
                f.write(f"POINT_DATA {len(pts)}\n")
                for name, arr in point_data.items():
                    arr = numpy.asarray(arr, dtype=numpy.float64)
                    assert arr.shape == (len(pts),), "point_data arrays hold one value per vertex"
                    f.write(f"SCALARS {name} double 1\nLOOKUP_TABLE default\n")
                    numpy.savetxt(f, arr, fmt="%.17g")

    # -- subdomain masks ------------------------------------------------------
    def get_vertex_mask(self, subdomain=None):
        if subdomain is None:
            return numpy.s_[:]
        mask = numpy.asarray(subdomain.is_inside(self.node_coords.T), dtype=bool)
        if getattr(subdomain, "is_boundary_only", False):
            mask = mask & self.is_boundary_point
        return mask

    def get_cell_mask(self, subdomain=None):
        if subdomain is None:
            return numpy.s_[:]
        if getattr(subdomain, "is_boundary_only", False):
            return numpy.zeros(len(self.cells_points), dtype=bool)
        inside = numpy.asarray(subdomain.is_inside(self.node_coords.T), dtype=bool)
        return numpy.all(inside[self.cells_points], axis=1)


def rectangle_tri(xs, ys, parity=0):
    """``meshzoo.rectangle_tri(xs, ys) -> (vertices, cells)`` [README:44-46]:
    tensor grid, x fastest, every quad split into two triangles with the
    diagonal direction alternating ("zigzag").  ``parity`` shifts the zigzag so
    that a slab ``ys[g0:g1]`` of a larger grid reproduces the global cells."""
    xs = numpy.asarray(xs, dtype=numpy.float64)
    ys = numpy.asarray(ys, dtype=numpy.float64)
    nx, ny = len(xs), len(ys)
    X, Y = numpy.meshgrid(xs, ys, indexing="xy")
    points = numpy.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = numpy.meshgrid(numpy.arange(nx - 1), numpy.arange(ny - 1), indexing="xy")
    ix, iy = ix.ravel(), iy.ravel()
    a = iy * nx + ix
    b = a + 1
    c = a + nx
    d = c + 1
    even = (ix + iy + parity) % 2 == 0
    dt = numpy.int64
    t1 = numpy.where(even[:, None], numpy.stack([a, b, d], 1), numpy.stack([a, b, c], 1))
    t2 = numpy.where(even[:, None], numpy.stack([a, d, c], 1), numpy.stack([b, d, c], 1))
    cells = numpy.concatenate([t1, t2], axis=0).astype(dt)
    return points, cells


def cube_tetra(xs, ys, zs, parity=0):
    """``meshzoo.cube_tetra``-style tensor grid, every hexahedron split into 5
    tetrahedra with alternating orientation (conforming).
    ``parity`` shifts the alternation so that a slab ``zs[g0:g1]`` of a larger
    grid reproduces the global cells."""
    xs, ys, zs = (numpy.asarray(a, dtype=numpy.float64) for a in (xs, ys, zs))
    nx, ny, nz = len(xs), len(ys), len(zs)
    Z, Y, X = numpy.meshgrid(zs, ys, xs, indexing="ij")
    points = numpy.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = numpy.meshgrid(
        numpy.arange(nz - 1), numpy.arange(ny - 1), numpy.arange(nx - 1), indexing="ij"
    )
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()

    def vid(dx, dy, dz):
        return ((iz + dz) * ny + (iy + dy)) * nx + (ix + dx)

    v = {(dx, dy, dz): vid(dx, dy, dz) for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)}
    even = (ix + iy + iz + parity) % 2 == 0
    # even cubes: corner tets at 000,110,101,011 cut off; centre tet joins the rest
    split_even = [
        [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)],
        [(1, 1, 0), (0, 1, 0), (1, 0, 0), (1, 1, 1)],
        [(1, 0, 1), (0, 0, 1), (1, 1, 1), (1, 0, 0)],
        [(0, 1, 1), (1, 1, 1), (0, 0, 1), (0, 1, 0)],
        [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)],
    ]
    # odd cubes: mirrored in x
    split_odd = [[(1 - a, b, c) for (a, b, c) in tet] for tet in split_even]
    cells = []
    for te, to in zip(split_even, split_odd):
        ce_ = numpy.stack([v[p] for p in te], axis=1)
        co_ = numpy.stack([v[p] for p in to], axis=1)
        cells.append(numpy.where(even[:, None], ce_, co_))
    cells = numpy.concatenate(cells, axis=0).astype(numpy.int64)
    return points, cells


def jitter(points, cells, scale=0.2, seed=0):
    """Robustness variant: interior vertices moved by
    ``U(-scale*h, scale*h)``; creates negative ce_ratios."""
    mesh = Mesh(points, cells)
    rng = numpy.random.default_rng(seed)
    e = mesh.ei_dot_ei
    h = numpy.sqrt(e.min())
    out = numpy.array(points, dtype=numpy.float64)
    interior = ~mesh.is_boundary_point
    out[interior] += rng.uniform(-scale * h, scale * h, size=out[interior].shape)
    return out, cells
