"""GPU tier: edge cases of the path -- vertices in no cell, terms that leave
records out of the pattern, int32 index input, very high vertex degree, FMA
build, geometry re-upload."""
import numpy
import pytest
import sympy

from . import parity, problems

pytestmark = pytest.mark.gpu


def _both():
    import oracle
    import pyfvm_b200

    return oracle, pyfvm_b200


def _compare_linear(prob_factory, mesh, **kw):
    oracle, eng = _both()
    s = {}
    A_ref, b_ref = oracle.discretize_linear(prob_factory(oracle.form_language), mesh, scales=s)
    A, b = eng.discretize_linear(prob_factory(eng.form_language), mesh, **kw)
    parity.assert_matrix_close(A, A_ref, s["matrix"])
    parity.assert_vector_close(b, b_ref, s["rhs"])
    return A, b


def reaction_only_left(fl):
    """The only dS term is restricted to a subdomain: records of the other cells
    leave the sparsity pattern (as the reference's masked COO blocks do)."""
    left = problems.LeftHalf()

    class P:
        def apply(self, u):
            return (
                fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS, subdomains=[left])
                + fl.integrate(lambda x: u(x), fl.dV)
                - fl.integrate(lambda x: 1.0 + x[0], fl.dV)
            )

        def dirichlet(self, u):
            return []

    return P()


def test_isolated_vertices_keep_their_diagonal():
    from pyfvm_b200 import meshes

    p, c = problems.rectangle(9, 7)
    extra = numpy.array([[5.0, 5.0], [6.0, 5.0], [7.0, 5.5]])
    mesh = meshes.Mesh(numpy.concatenate([p[:10], extra[:1], p[10:], extra[1:]]), c + (c >= 10))
    assert (mesh.control_volumes == 0.0).sum() == 3
    A, b = _compare_linear(problems.poisson, mesh)
    assert A.nnz == A.shape[0] + 2 * (8 * 7 + 9 * 6 + 8 * 6)  # diagonal kept for every vertex
    _compare_linear(problems.reaction_x, mesh)


def test_restricted_edge_term_shrinks_the_pattern():
    from pyfvm_b200 import meshes

    mesh = meshes.Mesh(*meshes.jitter(*problems.rectangle(33, 17), scale=0.15, seed=2))
    A, _ = _compare_linear(reaction_only_left, mesh)
    full, _ = _compare_linear(problems.poisson, mesh)
    assert A.nnz < full.nnz


def test_int32_index_input():
    from pyfvm_b200 import meshes

    mesh = meshes.Mesh(*problems.rectangle(21, 13))
    mesh._idx = numpy.ascontiguousarray(mesh.idx_hierarchy.astype(numpy.int32))
    _compare_linear(problems.convection_diffusion, mesh)


def _fan(k):
    """k triangles around a central vertex: a row with 2k half-edges."""
    ang = numpy.linspace(0.0, 2.0 * numpy.pi, k, endpoint=False)
    pts = numpy.concatenate([[[0.0, 0.0]], numpy.stack([numpy.cos(ang), numpy.sin(ang)], 1)])
    ring = numpy.arange(1, k + 1)
    cells = numpy.stack([numpy.zeros(k, dtype=numpy.int64), ring, numpy.roll(ring, -1)], 1)
    return pts, cells


def test_high_degree_vertex():
    from pyfvm_b200 import engine, meshes

    mesh = meshes.Mesh(*_fan(1000))  # centre row: 2000 half-edges, 1001 slots
    _compare_linear(problems.poisson, mesh)
    _compare_linear(problems.reaction_x, mesh)
    wider = meshes.Mesh(*_fan(1900))  # 3800 half-edges: the tile quantum shrinks, results hold
    _compare_linear(problems.reaction_x, wider)
    too_big = meshes.Mesh(*_fan(2100))
    with pytest.raises(engine.FvmError, match="4030"):
        _both()[1].discretize_linear(problems.poisson(_both()[1].form_language), too_big)


def test_fma_build_stays_within_tolerance():
    from pyfvm_b200 import meshes

    mesh = meshes.Mesh(*meshes.jitter(*problems.cube(9), scale=0.2, seed=3))
    _compare_linear(problems.reaction_x, mesh, fmad=True)


def test_geometry_reupload_matches_fresh_mesh():
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, problem

    p, c = problems.rectangle(25, 14)
    mesh = meshes.Mesh(p, c)
    lp = problem.LinearProblem(problems.convection_diffusion(eng.form_language), mesh)
    A0, b0 = lp.assemble()
    p2, _ = meshes.jitter(p, c, scale=0.1, seed=4)
    moved = meshes.Mesh(p2, c)
    lp.update_geometry(
        ce=numpy.asarray(moved.ce_ratios).reshape(-1), coords=p2, cv=moved.control_volumes
    )
    A1, b1 = lp.assemble()
    A2, b2 = eng.discretize_linear(problems.convection_diffusion(eng.form_language), moved)
    assert A1.indices.tobytes() == A2.indices.tobytes()
    assert A1.data.tobytes() == A2.data.tobytes() and b1.tobytes() == b2.tobytes()
    assert A1.data.tobytes() != A0.data.tobytes()


def test_newton_on_device_generated_mesh():
    """pyfvm.newton [README:113-121] unchanged on top of a device-generated mesh."""
    import pyfvm
    from pyfvm_b200 import device_meshes as dm
    from scipy.sparse import linalg

    d = dm.device_rectangle_tri(numpy.linspace(0.0, 2.0, 101), numpy.linspace(0.0, 1.0, 51))
    import pyfvm_b200 as eng

    f, jacobian = pyfvm.discretize(problems.bratu(eng.form_language), d)

    def jacobian_solver(u0, rhs):
        return linalg.spsolve(jacobian.get_linear_operator(u0), rhs)

    u = pyfvm.newton(f.eval, jacobian_solver, numpy.zeros(len(d.points)), verbose=False)
    assert u.max() == pytest.approx(0.28395, abs=1e-5)


def test_piecewise_sqrt_pi_functors():
    """Codegen gotchas end to end: Piecewise, sqrt, pi, x**4, u**3."""
    from pyfvm_b200 import meshes

    from .test_codegen_compiles import piecewise_problem

    oracle, eng = _both()
    mesh = meshes.Mesh(*meshes.jitter(*problems.rectangle(29, 17, lx=1.0), scale=0.15, seed=6))
    f_ref, j_ref = oracle.discretize(piecewise_problem(oracle.form_language), mesh)
    f, j = eng.discretize(piecewise_problem(eng.form_language), mesh)
    u = 0.4 * numpy.random.default_rng(2).standard_normal(len(mesh.points))
    s = {}
    F_ref = f_ref.eval(u, scales=s)
    parity.assert_vector_close(f.eval(u), F_ref, s["vec"])
    J_ref = j_ref.get_linear_operator(u, scales=s)
    parity.assert_matrix_close(j.get_linear_operator(u), J_ref, s["matrix"])


def test_device_resident_handoff_torch():
    """__cuda_array_interface__ in, torch tensors / sparse CSR out: same bits as the
    host path, nothing crosses PCIe (SURVEY §8b proposal, §8f rank 4)."""
    import torch

    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, problem

    mesh = meshes.Mesh(*meshes.jitter(*problems.rectangle(37, 23), scale=0.1, seed=8))
    f, jac = eng.discretize(problems.nonlinear_conv(eng.form_language), mesh)
    u = 0.3 * numpy.random.default_rng(4).standard_normal(len(mesh.points))
    ut = torch.from_numpy(u).cuda()
    Ft = f.eval(ut)
    assert isinstance(Ft, torch.Tensor) and Ft.is_cuda
    assert Ft.cpu().numpy().tobytes() == f.eval(u).tobytes()
    Jt = jac.get_linear_operator_torch(ut)
    J = jac.get_linear_operator(u)
    assert Jt.values().cpu().numpy().tobytes() == J.data.tobytes()
    assert numpy.array_equal(Jt.col_indices().cpu().numpy(), J.indices)
    assert numpy.array_equal(Jt.crow_indices().cpu().numpy(), J.indptr)
    lp = problem.LinearProblem(problems.convection_diffusion(eng.form_language), mesh)
    At, bt = lp.assemble_torch()
    A, b = lp.assemble()
    assert At.values().cpu().numpy().tobytes() == A.data.tobytes()
    assert bt.cpu().numpy().tobytes() == b.tobytes()
    x = torch.from_numpy(u).cuda()
    assert numpy.allclose((At @ x).cpu().numpy(), A @ u, rtol=1e-12, atol=1e-12)


def _moved(mesh, p2):
    """Move a mesh's points the way meshplex does: derived arrays are recomputed (new
    array objects), the mesh object stays the same."""
    mesh.points = p2
    nc = numpy.zeros((len(p2), 3))
    nc[:, : p2.shape[1]] = p2
    mesh.node_coords = nc
    mesh._geom = mesh._cv = None


def test_mesh_cache_revalidates_a_mutated_mesh():
    """A second discretize on the SAME mesh object after its points moved must see the
    new geometry (the reference re-reads the mesh on every call)."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes

    p, c = problems.rectangle(31, 19)
    mesh = meshes.Mesh(p, c)
    P = problems.convection_diffusion(eng.form_language)
    A0, b0 = eng.discretize_linear(P, mesh)
    p2, _ = meshes.jitter(p, c, scale=0.1, seed=5)
    _moved(mesh, p2)
    A1, b1 = eng.discretize_linear(P, mesh)
    A2, b2 = eng.discretize_linear(P, meshes.Mesh(p2, c))
    assert A1.data.tobytes() == A2.data.tobytes() and b1.tobytes() == b2.tobytes()
    assert A1.data.tobytes() != A0.data.tobytes()


def test_update_geometry_is_private_to_its_problem():
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes, problem

    p, c = problems.rectangle(23, 15)
    mesh = meshes.Mesh(p, c)
    P = problems.convection_diffusion(eng.form_language)
    lp_a = problem.LinearProblem(P, mesh)
    lp_b = problem.LinearProblem(P, mesh)
    assert lp_a.dmesh is lp_b.dmesh  # shared through the mesh cache
    A0, _ = lp_b.assemble()
    moved = meshes.Mesh(meshes.jitter(p, c, scale=0.1, seed=6)[0], c)
    lp_a.update_geometry(ce=numpy.asarray(moved.ce_ratios).reshape(-1), coords=moved.points,
                         cv=moved.control_volumes)
    assert lp_a.dmesh is not lp_b.dmesh
    A1, _ = lp_b.assemble()
    assert A1.data.tobytes() == A0.data.tobytes()
    with pytest.raises(ValueError):
        lp_a.update_geometry(ce=numpy.zeros(7))


def test_out_of_range_endpoint_is_refused():
    import pyfvm_b200 as eng
    from pyfvm_b200 import engine, meshes

    mesh = meshes.Mesh(*problems.rectangle(9, 7))
    mesh.ce_ratios, mesh.control_volumes, mesh.is_boundary_point  # geometry of the intact mesh
    idx = numpy.array(mesh.idx_hierarchy)
    idx[1, 0, 3] = len(mesh.points) + 5
    mesh._idx = idx
    with pytest.raises(engine.FvmError, match="outside"):
        eng.discretize_linear(problems.poisson(eng.form_language), mesh)


def test_odd_length_device_u():
    """f.eval on a torch tensor with an odd number of vertices (no slack after the last
    element) goes through the padded internal buffer."""
    import torch

    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes

    mesh = meshes.Mesh(*problems.rectangle(21, 13))  # 273 vertices
    assert len(mesh.points) % 2 == 1
    f, jac = eng.discretize(problems.bratu(eng.form_language), mesh)
    u = 0.2 * numpy.random.default_rng(0).standard_normal(len(mesh.points))
    F_host = f.eval(u)
    F_dev = f.eval(torch.from_numpy(u).cuda())
    assert F_dev.cpu().numpy().tobytes() == F_host.tobytes()


def test_cubin_cache_second_process_skips_nvrtc(tmp_path):
    """The NVRTC output is persisted: a second process with the same functor
    set loads the cubin from $FVM_CACHE_DIR and never calls nvrtcCompileProgram."""
    import os
    import subprocess
    import sys

    code = (
        "import numpy, pyfvm_b200 as e\n"
        "from pyfvm_b200 import meshes, engine\n"
        "from tests import problems\n"
        "m = meshes.Mesh(*problems.rectangle(17, 11))\n"
        "A, b = e.discretize_linear(problems.reaction_x(e.form_language), m)\n"
        "f, j = e.discretize(problems.bratu(e.form_language), m)\n"
        "f.eval(numpy.zeros(len(m.points))); j.get_linear_operator(numpy.zeros(len(m.points)))\n"
        "print('STATS', *engine.Engine.get().kernel_cache_stats(), float(abs(A).sum()))\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FVM_CACHE_DIR=str(tmp_path / "cubins"), PYTHONPATH=root)
    env.pop("FVM_KERNEL_SRC_DIR", None)
    outs = []
    for _ in range(2):
        r = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("STATS")][0].split())
    assert int(outs[0][1]) == 3 and int(outs[0][2]) == 0  # first process: three kernels compiled
    assert int(outs[1][1]) == 0 and int(outs[1][2]) == 3  # second: none compiled, three loaded
    assert outs[0][3] == outs[1][3]  # and the same numbers come out
    assert len(os.listdir(tmp_path / "cubins")) == 3


@pytest.mark.parametrize("kind,args,shuffle", [("rect", (63, 41), False), ("cube", (11,), False),
                                               ("rect", (37, 29), True), ("cube", (9,), True)])
def test_chunked_pattern_build_is_bit_identical(kind, args, shuffle, monkeypatch):
    """The pattern build in row chunks (what lets config 5's 4.8 G half-edges be built on one
    GPU) must produce byte-identical layouts and results to the one-chunk build."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import engine, meshes, problem

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    p, c = meshes.jitter(p, c, scale=0.15, seed=6)
    if shuffle:
        p, c = problems.shuffle_numbering(p, c, seed=6)
    mesh = meshes.Mesh(p, c)
    P = problems.subdomain_problem(eng.form_language) if kind == "rect" else problems.reaction_x(eng.form_language)
    out = []
    for chunk in (None, "4000", "700"):
        if chunk is None:
            monkeypatch.delenv("FVM_BUILD_CHUNK_HE", raising=False)
        else:
            monkeypatch.setenv("FVM_BUILD_CHUNK_HE", chunk)
        engine._mesh_cache.clear()
        lp = problem.LinearProblem(P, mesh)
        st = lp.dmesh.stats()
        assert (st["build_chunks"] > 1) == (chunk is not None), st
        tiles, meta, ce = lp.dmesh.layout()
        A, b = lp.assemble()
        out.append((st["build_chunks"], tiles.tobytes(), meta.tobytes(), ce.tobytes(),
                    A.indptr.tobytes(), A.indices.tobytes(), A.data.tobytes(), b.tobytes(),
                    lp.dmesh.control_volumes().tobytes()))
    monkeypatch.delenv("FVM_BUILD_CHUNK_HE", raising=False)
    engine._mesh_cache.clear()
    assert out[2][0] > out[1][0] > 1
    for o in out[1:]:
        assert o[1:] == out[0][1:]


def test_chunked_build_on_device_mesh_derives_control_volumes(monkeypatch):
    """Device-generated mesh (control volumes derived per chunk from chunk-sized edge-length
    buffers, edge lengths not kept): same residual bits chunked and unchunked."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import device_meshes as dm
    from pyfvm_b200 import engine

    xs = numpy.linspace(0.0, 1.0, 12)
    u = 0.2 * numpy.random.default_rng(3).standard_normal(12**3)
    res = []
    for chunk in (None, "5000"):
        if chunk is None:
            monkeypatch.delenv("FVM_BUILD_CHUNK_HE", raising=False)
        else:
            monkeypatch.setenv("FVM_BUILD_CHUNK_HE", chunk)
        engine._mesh_cache.clear()
        d = dm.device_cube_tetra(xs, xs, xs)
        f, j = eng.discretize(problems.bratu(eng.form_language), d)
        assert not f._s.dmesh.has_el  # Bratu reads no edge_length: the stream is not kept
        res.append((f._s.dmesh.stats()["build_chunks"], f.eval(u).tobytes(),
                    j.get_linear_operator(u).data.tobytes(), f._s.dmesh.control_volumes().tobytes()))
    monkeypatch.delenv("FVM_BUILD_CHUNK_HE", raising=False)
    engine._mesh_cache.clear()
    assert res[0][0] == 1 and res[1][0] > 1 and res[0][1:] == res[1][1:]


def test_device_block_cache_recycles_poisoned_blocks(monkeypatch):
    """The block cache: a dropped mesh's arrays and the build temporaries are
    handed out again.  With FVM_CACHE_POISON=1 every recycled block comes back full of NaN bit
    patterns, so the rebuilt mesh + assembly must not depend on fresh memory being zero:
    same bits as the first build, and the oracle's values."""
    import gc

    import oracle
    import pyfvm_b200 as eng
    from pyfvm_b200 import engine, meshes

    e = engine.Engine.get()
    monkeypatch.setenv("FVM_CACHE_POISON", "1")
    p, c = problems.rectangle(143, 97)
    p, c = meshes.jitter(p, c, scale=0.2, seed=4)
    mesh = meshes.Mesh(p, c)
    prob = problems.convection_diffusion

    def run():
        engine._mesh_cache.clear()  # a fresh device mesh (and pattern build) every time
        gc.collect()
        A, b = eng.discretize_linear(prob(eng.form_language), mesh)
        return A.indptr.copy(), A.indices.copy(), A.data.copy(), b.copy()

    first = run()
    gc.collect()
    idle_bytes, idle_blocks, _ = e.device_cache_stats()
    assert idle_blocks > 0 and idle_bytes > 0  # the first build's blocks are waiting
    second = run()
    gc.collect()
    for a, b in zip(first, second):
        assert a.tobytes() == b.tobytes()
    A_ref, b_ref = oracle.discretize_linear(prob(oracle.form_language), mesh)
    assert numpy.array_equal(first[0], A_ref.indptr) and numpy.array_equal(first[1], A_ref.indices)
    assert numpy.abs(second[2] - A_ref.data).max() <= 1e-13 * numpy.abs(A_ref.data).max()
    assert numpy.abs(second[3] - b_ref).max() <= 1e-13 * max(1.0, numpy.abs(b_ref).max())
    # residual / Jacobian / SpMV on recycled blocks too
    f, jac = eng.discretize(problems.bratu(eng.form_language), mesh)
    u = 0.1 * numpy.random.default_rng(0).standard_normal(len(p))
    fo, jo = oracle.discretize(problems.bratu(oracle.form_language), mesh)
    assert numpy.abs(f.eval(u) - fo.eval(u)).max() <= 1e-12
    J, Jo = jac.get_linear_operator(u), jo.get_linear_operator(u)
    assert numpy.abs(J.data - Jo.data).max() <= 1e-13 * numpy.abs(Jo.data).max()
    e.device_cache_flush()
    assert e.device_cache_stats()[0] == 0


@pytest.mark.parametrize("shuffle", [False, True])
def test_bank_steering_changes_no_bits(shuffle, monkeypatch):
    """A triangle mesh's vertex tables are re-laid (bank steering) on the mesh's
    second launch: the first (unsteered) and later (steered) launches, and a mesh that is
    never steered (FVM_BANK_STEERING=0), must produce the same bits in every mode."""
    import gc

    import pyfvm_b200 as eng
    from pyfvm_b200 import engine, meshes, solvers

    p, c = problems.rectangle(211, 157)
    p, c = meshes.jitter(p, c, scale=0.2, seed=2)
    if shuffle:
        p, c = problems.shuffle_numbering(p, c, seed=3)
    mesh = meshes.Mesh(p, c)
    u = 0.1 * numpy.random.default_rng(1).standard_normal(len(p))

    def run():
        engine._mesh_cache.clear()
        gc.collect()
        out = []
        lp = eng.problem.LinearProblem(problems.convection_diffusion(eng.form_language), mesh)
        for _ in range(3):  # launch 1: natural tables; launches 2, 3: steered (if enabled)
            A, b = lp.assemble()
            out.append((A.data.copy(), b.copy()))
        xtab = lp.dmesh.stats()["max_xtab"]
        f, jac = eng.discretize(problems.bratu(eng.form_language), mesh)
        F = f.eval(u).copy()
        J = jac.get_linear_operator(u).data.copy()
        sysm = solvers.DeviceSystem(lp.dmesh, lp.data.ptr)
        xd, yd = lp.eng.to_device(u, pad=16), lp.eng.alloc(8 * lp.n + 16)
        sysm.matvec(xd.ptr, yd.ptr)
        y = numpy.empty(lp.n)
        yd.download(y)
        return out, F, J, y, xtab

    monkeypatch.setenv("FVM_BANK_STEERING", "1")
    on = run()
    monkeypatch.setenv("FVM_BANK_STEERING", "0")
    off = run()
    for (d0, b0) in on[0] + off[0]:
        assert d0.tobytes() == on[0][0][0].tobytes() and b0.tobytes() == on[0][0][1].tobytes()
    for a, b in zip(on[1:4], off[1:4]):
        assert a.tobytes() == b.tobytes()
    if not shuffle:
        assert on[4] >= off[4]  # the steered tables stage a few extra vertices


def test_error_in_a_subdomain_predicate_during_the_threaded_mesh_build():
    """The device mesh is built on a helper thread while the main thread evaluates the
    user's subdomain predicates: an exception there must surface as itself (after the build
    thread was joined) and leave the engine usable."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import meshes

    fl = eng.form_language

    class Broken(fl.Subdomain if hasattr(fl, "Subdomain") else object):
        is_boundary_only = True

        def is_inside(self, x):
            raise ValueError("predicate exploded")

    class P:
        def apply(self, u):
            return fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS) - fl.integrate(lambda x: 1.0, fl.dV)

        def dirichlet(self, u):
            return [(u, Broken())]

    mesh = meshes.Mesh(*problems.rectangle(61, 47))
    with pytest.raises(ValueError, match="predicate exploded"):
        eng.discretize_linear(P(), mesh)
    A, b = eng.discretize_linear(problems.poisson(fl), mesh)  # same mesh object, engine intact
    A2, b2 = eng.discretize_linear(problems.poisson(fl), meshes.Mesh(*problems.rectangle(61, 47)))
    assert A.data.tobytes() == A2.data.tobytes() and b.tobytes() == b2.tobytes()
