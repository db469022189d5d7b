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
