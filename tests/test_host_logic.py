"""CPU tier for the host side of the engine: the symbolic lowering
(``pyfvm_b200.lowering``) evaluated with numpy in the kernel's half-edge semantics
must reproduce the oracle; the form language and the code generator behave as
the drop-in needs."""
import numpy
import pytest
import sympy
from scipy import sparse

from . import problems


def _lam(expr, args):
    f = sympy.lambdify(args, sympy.sympify(expr), modules="numpy")
    return lambda *a: numpy.broadcast_to(f(*a), numpy.shape(a[0])).astype(float)


def emulate_linear(lowered, mesh):
    """What the tile kernel computes, in numpy: every record is seen as two
    half-edges (row, col); the row-0 expressions are evaluated with x0 = x[row],
    x1 = x[col]; D -> (row,row), O -> (row,col), C -> -rhs[row]."""
    from pyfvm_b200 import lowering as lo

    idx = numpy.asarray(mesh.idx_hierarchy).reshape(2, -1)
    ce = numpy.asarray(mesh.ce_ratios).reshape(-1)
    el = numpy.sqrt(numpy.asarray(mesh.ei_dot_ei)).reshape(-1)
    X = mesh.node_coords
    n = len(X)
    rows, cols, vals = [], [], []
    rhs = numpy.zeros(n)
    eargs = lo.X0 + lo.X1 + [lo.CE, lo.EL]
    for t in lowered.edge:
        keep = numpy.ones(len(ce), dtype=bool)
        if t.subdomain is not None:
            cm = numpy.asarray(mesh.get_cell_mask(t.subdomain))
            keep = numpy.broadcast_to(cm, numpy.asarray(mesh.ce_ratios).shape).reshape(-1)
        for a, b in ((idx[0], idx[1]), (idx[1], idx[0])):
            a, b, c, l = a[keep], b[keep], ce[keep], el[keep]
            args = [X[a, k] for k in range(3)] + [X[b, k] for k in range(3)] + [c, l]
            rows += [a, a]
            cols += [a, b]
            vals += [_lam(t.outputs["diag"], eargs)(*args), _lam(t.outputs["off"], eargs)(*args)]
            numpy.subtract.at(rhs, a, _lam(t.outputs["aff"], eargs)(*args))
    vargs = lo.XV + [lo.CV, lo.SA]
    allv = numpy.arange(n)
    for t in lowered.vertex:
        m = allv if t.subdomain is None else allv[mesh.get_vertex_mask(t.subdomain)]
        args = [X[m, k] for k in range(3)] + [numpy.asarray(mesh.control_volumes)[m],
                                              numpy.asarray(mesh.surface_areas)[m]]
        rows.append(m)
        cols.append(m)
        vals.append(_lam(t.outputs["lin"], vargs)(*args))
        rhs[m] -= _lam(t.outputs["aff"], vargs)(*args)
    # every row owns its diagonal
    rows.append(allv), cols.append(allv), vals.append(numpy.zeros(n))
    A = sparse.coo_matrix(
        (numpy.concatenate(vals), (numpy.concatenate(rows), numpy.concatenate(cols))), shape=(n, n)
    ).tocsr()
    for t in lowered.dirichlet:
        m = allv[mesh.get_vertex_mask(t.subdomain)]
        args = [X[m, k] for k in range(3)]
        for i, cf, r in zip(m, _lam(t.outputs["coeff"], lo.XV)(*args), _lam(t.outputs["rhs"], lo.XV)(*args)):
            A.data[A.indptr[i] : A.indptr[i + 1]] = 0.0
            A[i, i] = cf
            rhs[i] = r
    return A, rhs


@pytest.mark.parametrize(
    "name,kind,args",
    [("poisson", "rect", (9, 6)), ("convection_diffusion", "rect", (8, 7)), ("reaction_x", "rect", (7, 9)),
     ("subdomain_problem", "rect", (11, 7)), ("reaction_x", "cube", (4,)),
     ("robin", "rect", (9, 7)), ("robin", "cube", (5,)), ("neumann_only", "rect", (8, 6))],
)
def test_lowering_in_half_edge_semantics_matches_oracle(name, kind, args):
    import oracle
    from pyfvm_b200 import form_language, lowering, meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    p, c = meshes.jitter(p, c, scale=0.15, seed=1)
    mesh = meshes.Mesh(p, c)
    A_ref, b_ref = oracle.discretize_linear(getattr(problems, name)(oracle.form_language), mesh)
    low = lowering.lower_linear(getattr(problems, name)(form_language), 3)
    A, b = emulate_linear(low, mesh)
    assert numpy.array_equal(A.indptr, A_ref.indptr) and numpy.array_equal(A.indices, A_ref.indices)
    scale = numpy.abs(A_ref.data).max()
    assert numpy.abs(A.data - A_ref.data).max() <= 1e-12 * scale
    assert numpy.abs(b - b_ref).max() <= 1e-12 * max(1.0, numpy.abs(b_ref).max())


def test_form_language_algebra_and_exports():
    import pyfvm.form_language as shim
    from pyfvm_b200 import form_language as fl

    for name in ("Boundary", "dS", "dV", "integrate", "n_dot_grad"):  # [README:32]
        assert hasattr(shim, name) and hasattr(fl, name)
    u = sympy.Function("u")
    s = fl.integrate(lambda x: u(x), fl.dS) - 2 * fl.integrate(lambda x: 1.0, fl.dV)
    s = -s + fl.integrate(lambda x: x[0], fl.dV, subdomains=[fl.Boundary()])
    assert len(s.integrals) == 3
    assert fl.Boundary().is_boundary_only and fl.Boundary().is_inside(numpy.zeros((3, 4))).all()


def test_lowering_known_forms():
    """App. A.2: Poisson edge block [[er, -er], [-er, er]], affine 0; Bratu derivatives."""
    from pyfvm_b200 import form_language, lowering as lo

    low = lo.lower_linear(problems.poisson(form_language), 2)
    (e,) = low.edge
    assert sympy.simplify(e.outputs["diag"] - lo.CE) == 0
    assert sympy.simplify(e.outputs["off"] + lo.CE) == 0 and e.outputs["aff"] == 0
    (v,) = low.vertex
    assert v.outputs["lin"] == 0 and sympy.simplify(v.outputs["aff"] + lo.CV) == 0
    (d,) = low.dirichlet
    assert d.outputs["coeff"] == 1 and d.outputs["rhs"] == 0
    non = lo.lower_nonlinear(problems.bratu(form_language), 2)
    assert sympy.simplify(non.vertex[0].outputs["lin"] + 2 * lo.CV * sympy.exp(lo.UK0)) == 0
    with pytest.raises(AssertionError):
        lo.lower_linear(problems.bratu(form_language), 2)  # not linear in u


def test_codegen_flags_and_factoring():
    from pyfvm_b200 import codegen as cg
    from pyfvm_b200 import form_language, lowering as lo

    low = lo.lower_linear(problems.convection_diffusion(form_language), 2)
    f = cg.generate(low, cg.MODE_LINEAR, [None], [None])
    assert f.factored and f.uses_x and not f.uses_el and f.has_dirichlet
    assert f.flags & cg.F_EDGE_D and not f.flags & cg.F_EDGE_C and not f.flags & cg.F_U
    low = lo.lower_linear(problems.reaction_x(form_language), 2)
    f = cg.generate(low, cg.MODE_LINEAR, [None], [None, None])
    assert "x0_0" in f.source and "t0" in f.source or "x1_0" in f.source  # cse names never clash with x0_k
    non = lo.lower_nonlinear(problems.bratu(form_language), 3)
    r = cg.generate(non, cg.MODE_RESIDUAL, [None], [None])
    j = cg.generate(non, cg.MODE_JACOBIAN, [None], [None])
    assert r.flags & cg.F_EDGE_C and not r.flags & cg.F_EDGE_D and r.needs_u
    assert j.flags & cg.F_EDGE_D and not j.flags & cg.F_EDGE_C
    sub = lo.lower_linear(problems.subdomain_problem(form_language), 2)
    ebits = [None if t.subdomain is None else 0 for t in sub.edge]
    f = cg.generate(sub, cg.MODE_LINEAR, ebits, [None if t.subdomain is None else 0 for t in sub.vertex])
    assert f.edge_masked and not f.factored and f.vert_masked  # per-half-edge masks forbid factoring


def test_vtk_output(tmp_path):
    """mesh.write("out.vtk", point_data={"x": x}) [README:53]."""
    from pyfvm_b200 import meshes

    mesh = meshes.Mesh(*problems.rectangle(5, 4))
    x = numpy.arange(20, dtype=float)
    out = tmp_path / "out.vtk"
    mesh.write(str(out), point_data={"x": x})
    text = out.read_text()
    assert "POINTS 20 double" in text and "CELLS 24 96" in text and "SCALARS x double 1" in text


def test_mesh_fingerprint_tracks_mutation():
    """The device-mesh cache key (no GPU needed): same object + same arrays -> same
    fingerprint; moved points or flipped cells -> different."""
    from pyfvm_b200 import engine, meshes
    from tests import problems

    p, c = problems.rectangle(17, 11)
    mesh = meshes.Mesh(p, c)
    fp0 = engine.mesh_fingerprint(mesh)
    assert engine.mesh_fingerprint(mesh) == fp0
    p2 = p.copy()
    p2[40] += 1e-3
    mesh.points = p2
    mesh.node_coords = numpy.concatenate([p2, numpy.zeros((len(p2), 1))], axis=1)
    mesh._geom = mesh._cv = None
    assert engine.mesh_fingerprint(mesh) != fp0
    other = meshes.Mesh(p, c[::-1].copy())
    assert engine.mesh_fingerprint(other) != fp0


def test_result_pool_hands_out_a_buffer_only_when_nobody_holds_it():
    """engine.host_result (no GPU: the page-locked allocation is replaced by a ctypes
    buffer): a size gets a pooled buffer from its second request on, a buffer is reused
    only after every array / scipy matrix over it is gone, small results bypass the pool."""
    import ctypes

    from pyfvm_b200 import engine

    eng = object.__new__(engine.Engine)
    made = []

    def fake(nbytes):
        raw = (ctypes.c_char * nbytes)()
        made.append(nbytes)
        return raw

    eng._new_result_buffer = fake
    n = (engine.Engine.POOL_MIN_BYTES // 8) + 1000

    def addr(a):
        return a.__array_interface__["data"][0]

    small = eng.host_result(10, numpy.float64)
    assert small.shape == (10,) and not made
    first = eng.host_result(n, numpy.float64)  # one-shot callers never pay for page-locking
    assert not made and first.shape == (n,) and first.dtype == numpy.float64
    a = eng.host_result(n, numpy.float64)
    assert len(made) == 1 and made[0] >= 8 * n and made[0] % (1 << 20) == 0
    b = eng.host_result(n, numpy.float64)  # a is alive: must not alias it
    assert len(made) == 2 and addr(b) != addr(a)
    pa = addr(a)
    a[:] = 1.0
    # a scipy matrix over the array keeps the buffer busy, and scipy must not copy it
    # (it does copy slices and reinterpreting views, hence one frombuffer array per result)
    m = sparse.csr_matrix((a, numpy.arange(n, dtype=numpy.int32), numpy.arange(n + 1, dtype=numpy.int32)),
                          shape=(n, n), copy=False)
    assert addr(m.data) == pa
    del a
    c = eng.host_result(n, numpy.float64)
    assert len(made) == 3 and addr(c) not in (pa, addr(b))
    del m
    d = eng.host_result(n, numpy.int64)  # free again: reused, any dtype of the same byte size
    assert len(made) == 3 and addr(d) == pa and d.dtype == numpy.int64
    # pool cap: beyond it plain arrays
    eng.POOL_MAX_BYTES = sum(made)
    e = eng.host_result(n, numpy.float64)
    assert len(made) == 3 and addr(e) not in (pa, addr(b), addr(c))


def test_chebyshev_preconditioner_recurrence_on_the_oracle_matrix():
    """The algorithm of fvm_cg / fvm_cg_poly restated in numpy on the oracle's 3-D Poisson
    matrix (CPU tier for the device solver's arithmetic): the Jacobi start x0 = b/diag keeps
    pyfvm's Dirichlet rows (columns not eliminated) out of the Krylov space, the Chebyshev
    polynomial on [lmax/30, lmax] with Gershgorin's lmax cuts the iterations 2.5-4 times for
    about the same number of SpMVs, and every variant reaches spsolve's solution."""
    import oracle
    from pyfvm_b200 import meshes
    from scipy.sparse import linalg

    mesh = meshes.Mesh(*problems.cube(26))
    A, b = oracle.discretize_linear(problems.poisson(oracle.form_language), mesh)
    A = A.tocsr()
    d = A.diagonal()
    n = A.shape[0]
    assert abs(A - A.T).max() > 0  # Dirichlet rows: the assembled matrix is NOT symmetric
    lmax = (abs(A) @ numpy.ones(n) / abs(d)).max()
    assert lmax == pytest.approx(2.0, abs=1e-12)  # diagonally dominant M-matrix

    def pcg(prec, degree):
        x = b / d
        r = b - A @ x
        z = prec(r)
        p, rz, it, spmv = z.copy(), r @ z, 0, 1
        while r @ r > 1e-22 * (b @ b):
            q = A @ p
            alpha = rz / (p @ q)
            assert alpha > 0  # the device's breakdown test
            x += alpha * p
            r -= alpha * q
            z = prec(r)
            rz, rz_old = r @ z, rz
            p = z + (rz / rz_old) * p
            it += 1
            spmv += degree
            assert it < 500
        return x, it, spmv

    def chebyshev(k):
        lmin = lmax / 30.0
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta

        def prec(r):
            rho = 1.0 / sigma
            dd = (r / d) / theta
            z = dd.copy()
            for _ in range(1, k):
                rho_new = 1.0 / (2.0 * sigma - rho)
                dd = rho_new * rho * dd + (2.0 * rho_new / delta) * ((r - A @ z) / d)
                z = z + dd
                rho = rho_new
            return z

        return prec

    x_ref = linalg.spsolve(A.tocsc(), b)
    x1, it1, mv1 = pcg(lambda r: r / d, 1)
    assert numpy.abs(x1 - x_ref).max() <= 1e-9 * numpy.abs(x_ref).max()
    for k in (3, 4):
        xk, itk, mvk = pcg(chebyshev(k), k)
        assert numpy.abs(xk - x_ref).max() <= 1e-9 * numpy.abs(x_ref).max()
        assert itk <= 0.5 * it1 + 2 and mvk <= 1.3 * mv1 + k, (k, itk, it1, mvk, mv1)
