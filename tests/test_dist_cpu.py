"""CPU tier for the row-partitioned path (world_size 2 and 3, gloo): partition
bookkeeping, the halo exchange, and -- with the oracle standing in for the
device on every rank -- that owned rows assembled from a local piece are the
global rows, bit for bit."""
import multiprocessing as mp
import os
import socket

import numpy
import pytest

from . import dist_workers, problems


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind,n", [("rect", 23), ("cube", 7)])
@pytest.mark.parametrize("world", [2, 3])
def test_partition_bookkeeping(kind, n, world):
    from pyfvm_b200 import dist as fd

    mesh = dist_workers._mesh(kind, n)
    nglob = len(mesh.points)
    idx = numpy.asarray(mesh.idx_hierarchy).reshape(2, -1)
    parts = [fd.Partition(mesh, r, world) for r in range(world)]
    assert sum(p.n_owned for p in parts) == nglob
    for p in parts:
        lo, hi = p.bounds[p.rank], p.bounds[p.rank + 1]
        g0 = p.n_ghost_lo
        assert numpy.array_equal(p.verts[g0 : g0 + p.n_owned], numpy.arange(lo, hi))
        assert (numpy.diff(p.verts) > 0).all()  # monotone local numbering
        # every record touching an owned vertex is present, in the global order
        touch = ((idx[0] >= lo) & (idx[0] < hi)) | ((idx[1] >= lo) & (idx[1] < hi))
        assert numpy.array_equal(p.verts[p.mesh.idx_hierarchy[0]], idx[0][touch])
        assert numpy.array_equal(p.verts[p.mesh.idx_hierarchy[1]], idx[1][touch])
        # both sides of every halo link agree on what travels, in the same order
        for q, (start, count) in p.recv.items():
            want = p.verts[start : start + count]
            have = parts[q].verts[parts[q].send[p.rank]]
            assert numpy.array_equal(want, have)
        assert set(p.recv) == set(p.send)


@pytest.mark.parametrize("kind,n,world", [("rect", 23, 2), ("cube", 7, 2), ("rect", 31, 3)])
def test_gloo_halo_and_owned_rows(kind, n, world, tmp_path):
    import oracle

    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [
        ctx.Process(target=dist_workers.gloo_halo_worker, args=(r, world, port, kind, n, str(tmp_path)))
        for r in range(world)
    ]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]

    mesh = dist_workers._mesh(kind, n)
    u = numpy.random.default_rng(7).standard_normal(len(mesh.points))
    f, jac = oracle.discretize(problems.bratu(oracle.form_language), mesh)
    F, J = f.eval(u), jac.get_linear_operator(u)
    A, b = oracle.discretize_linear(problems.convection_diffusion(oracle.form_language), mesh)
    for r in range(world):
        z = numpy.load(os.path.join(tmp_path, f"rank{r}.npz"))
        lo, hi = int(z["lo"]), int(z["hi"])
        assert z["F"].tobytes() == F[lo:hi].tobytes()
        assert z["b"].tobytes() == b[lo:hi].tobytes()
        for name, M in (("J", J), ("A", A)):
            blk = M[lo:hi]
            assert numpy.array_equal(z[f"{name}_indptr"], blk.indptr)
            assert numpy.array_equal(z[f"{name}_indices"], blk.indices)
            assert z[f"{name}_data"].tobytes() == blk.data.tobytes()
    normsq = numpy.load(os.path.join(tmp_path, "normsq.npy"))[0]
    assert normsq == pytest.approx(float(F @ F), rel=1e-14)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_partition_matches_general_partition(dim, world):
    """Slabs generated locally (the global mesh never exists) hold exactly the
    records, control volumes, boundary flags and halo links of the general
    partition of the global mesh."""
    from pyfvm_b200 import dist as fd
    from pyfvm_b200 import meshes

    if dim == 2:
        axes = [numpy.linspace(0, 1, 7), numpy.linspace(0, 2, 9)]
        gm = meshes.Mesh(*meshes.rectangle_tri(*axes))
    else:
        axes = [numpy.linspace(0, 1, 7), numpy.linspace(0, 2, 6), numpy.linspace(0, 1, 9)]
        gm = meshes.Mesh(*meshes.cube_tetra(*axes))

    def records(part):
        idx = part.mesh.idx_hierarchy.reshape(2, -1)
        ce = numpy.asarray(part.mesh.ce_ratios).reshape(-1)
        g0, g1 = part.verts[idx[0]], part.verts[idx[1]]
        lo, hi = part.bounds[part.rank], part.bounds[part.rank + 1]
        keep = ((g0 >= lo) & (g0 < hi)) | ((g1 >= lo) & (g1 < hi))
        a, c = numpy.stack([g0[keep], g1[keep]], 1), ce[keep]
        o = numpy.lexsort((c, a[:, 1], a[:, 0]))
        return a[o], c[o]

    for r in range(world):
        sp = fd.SlabPartition(axes, r, world)
        gp = fd.Partition(gm, r, world, bounds=sp.bounds)
        (a1, c1), (a2, c2) = records(sp), records(gp)
        assert numpy.array_equal(a1, a2)
        assert numpy.allclose(c1, c2, rtol=1e-12, atol=1e-15)
        g0, lo, hi = sp.n_ghost_lo, sp.bounds[r], sp.bounds[r + 1]
        assert numpy.allclose(
            sp.mesh.control_volumes[g0 : g0 + sp.n_owned], gm.control_volumes[lo:hi], rtol=1e-13
        )
        assert numpy.array_equal(
            sp.mesh.is_boundary_point[g0 : g0 + sp.n_owned], gm.is_boundary_point[lo:hi]
        )
        for q, (start, count) in sp.recv.items():
            sq = fd.SlabPartition(axes, q, world)
            assert numpy.array_equal(sp.verts[start : start + count], sq.verts[sq.send[r]])
