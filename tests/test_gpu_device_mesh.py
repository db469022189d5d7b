"""GPU tier: mesh generation and geometry on the device
against the host mesh provider (``pyfvm_b200.meshes``, the meshplex stand-in the
oracle consumes), then the hot path on top of it against the oracle."""
import numpy
import pytest

from . import parity, problems

pytestmark = pytest.mark.gpu


def _host_records(mesh):
    idx = numpy.asarray(mesh.idx_hierarchy)
    return (
        idx[0].reshape(-1), idx[1].reshape(-1),
        numpy.asarray(mesh.ce_ratios).reshape(-1),
        numpy.sqrt(numpy.asarray(mesh.ei_dot_ei)).reshape(-1),
    )


def _check_geometry(dmesh, hmesh):
    i0, i1, ce, el = dmesh.records_to_host()
    h0, h1, hce, hel = _host_records(hmesh)
    assert numpy.array_equal(i0, h0) and numpy.array_equal(i1, h1)  # bit-exact index work
    # fp64 geometry: 1e-13 relative to the largest ce of the mesh (hypotenuse /
    # sliver entries are ~0 by cancellation), edge lengths 1e-14 relative
    assert numpy.abs(ce - hce).max() <= 1e-13 * numpy.abs(hce).max()
    assert numpy.abs(el - hel).max() <= 1e-14 * hel.max()


@pytest.mark.parametrize("nx,ny,par", [(2, 2, 0), (17, 9, 0), (40, 23, 1)])
def test_device_rectangle_tri_matches_host(nx, ny, par):
    from pyfvm_b200 import device_meshes as dm
    from pyfvm_b200 import meshes

    xs, ys = numpy.linspace(0.0, 2.0, nx), numpy.linspace(-1.0, 1.0, ny)
    d = dm.device_rectangle_tri(xs, ys, parity=par)
    p, c = meshes.rectangle_tri(xs, ys, parity=par)
    assert d.points.tobytes() == p.tobytes()
    _check_geometry(d, meshes.Mesh(p, c))
    if par == 0:
        assert numpy.array_equal(d.is_boundary_point, meshes.Mesh(p, c).is_boundary_point)


@pytest.mark.parametrize("n,par", [(2, 0), (6, 0), (9, 1)])
def test_device_cube_tetra_matches_host(n, par):
    from pyfvm_b200 import device_meshes as dm
    from pyfvm_b200 import meshes

    xs, ys, zs = numpy.linspace(0, 1, n), numpy.linspace(0, 2, n + 1), numpy.linspace(0, 1, n + 2)
    d = dm.device_cube_tetra(xs, ys, zs, parity=par)
    p, c = meshes.cube_tetra(xs, ys, zs, parity=par)
    assert d.points.tobytes() == p.tobytes()
    _check_geometry(d, meshes.Mesh(p, c))
    if par == 0:
        assert numpy.array_equal(d.is_boundary_point, meshes.Mesh(p, c).is_boundary_point)


@pytest.mark.parametrize("kind,args", [("rect", (31, 19)), ("cube", (9,))])
def test_general_mesh_geometry_control_volumes_and_boundary(kind, args):
    """Jittered (non-Delaunay in places: negative ce_ratios) meshes from host arrays."""
    from pyfvm_b200 import device_meshes as dm
    from pyfvm_b200 import engine, meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    p, c = meshes.jitter(p, c, scale=0.2, seed=1)
    h = meshes.Mesh(p, c)
    d = dm.DeviceSimplexMesh.from_host(p, c, boundary=h.is_boundary_point)
    _check_geometry(d, h)
    dmesh = engine.DeviceMesh(d.eng, d)
    cv = dmesh.control_volumes()
    assert numpy.abs(cv - h.control_volumes).max() <= 1e-13 * h.control_volumes.max()
    assert cv.sum() == pytest.approx(h.control_volumes.sum(), rel=1e-13)
    if kind == "rect":
        assert numpy.array_equal(dmesh.boundary_vertices_tri(), h.is_boundary_point)


@pytest.mark.parametrize(
    "name,kind,n", [("convection_diffusion", "rect", 45), ("poisson", "cube", 10), ("reaction_x", "rect", 33)]
)
def test_assembly_on_device_generated_mesh(name, kind, n):
    """discretize_linear on a device-generated mesh vs the oracle on the host mesh."""
    import oracle
    import pyfvm_b200 as eng
    from pyfvm_b200 import device_meshes as dm
    from pyfvm_b200 import meshes

    xs = numpy.linspace(0.0, 1.0, n)
    if kind == "rect":
        d = dm.device_rectangle_tri(xs, xs)
        h = meshes.Mesh(*meshes.rectangle_tri(xs, xs))
    else:
        d = dm.device_cube_tetra(xs, xs, xs)
        h = meshes.Mesh(*meshes.cube_tetra(xs, xs, xs))
    s = {}
    A_ref, b_ref = oracle.discretize_linear(getattr(problems, name)(oracle.form_language), h, scales=s)
    A, b = eng.discretize_linear(getattr(problems, name)(eng.form_language), d)
    # the device geometry differs from numpy's in the last bits of each ce_ratio:
    # allow 4x the assembly tolerance
    parity.assert_pattern_equal(A, A_ref)
    assert (numpy.abs(A.data - A_ref.data) <= 4 * parity.RTOL * s["matrix"].data).all()
    bound = 4 * parity.RTOL * numpy.maximum(s["rhs"], numpy.abs(b_ref))
    assert (numpy.abs(b - b_ref) <= bound).all()


def test_bratu_residual_on_device_slab():
    """Nonlinear path on a device-generated slab (world 1) vs the oracle."""
    import oracle
    import pyfvm_b200 as eng
    from pyfvm_b200 import dist as fd
    from pyfvm_b200 import meshes

    xs = numpy.linspace(0.0, 1.0, 11)
    part = fd.SlabPartition([xs, xs, xs], 0, 1, device=True)
    h = meshes.Mesh(*meshes.cube_tetra(xs, xs, xs))
    f_ref, j_ref = oracle.discretize(problems.bratu(oracle.form_language), h)
    f, j = eng.discretize(problems.bratu(eng.form_language), part.mesh)
    u = 0.2 * numpy.random.default_rng(5).standard_normal(len(h.points))
    s = {}
    F_ref = f_ref.eval(u, scales=s)
    F = f.eval(u)
    assert (numpy.abs(F - F_ref) <= 4 * parity.RTOL * numpy.maximum(s["vec"], numpy.abs(F_ref))).all()
    J_ref = j_ref.get_linear_operator(u, scales=s)
    J = j.get_linear_operator(u)
    parity.assert_pattern_equal(J, J_ref)
    assert (numpy.abs(J.data - J_ref.data) <= 4 * parity.RTOL * s["matrix"].data).all()


@pytest.mark.parametrize("kind,args,shuffle", [("rect", (37, 23), False), ("cube", (9,), False),
                                               ("cube", (8,), True), ("rect", (20, 31), True)])
def test_boundary_vertices_on_device(kind, args, shuffle):
    """meshplex's is_boundary_point from a facet sort on the device (2-D and 3-D, any
    numbering) equals the host provider's."""
    from pyfvm_b200 import device_meshes as dm
    from pyfvm_b200 import meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    p, c = meshes.jitter(p, c, scale=0.2, seed=2)
    if shuffle:
        p, c = problems.shuffle_numbering(p, c, seed=8)
    h = meshes.Mesh(p, c)
    d = dm.DeviceSimplexMesh.from_host(p, c)  # no flags supplied
    assert numpy.array_equal(d.is_boundary_point, h.is_boundary_point)


@pytest.mark.parametrize("kind,args", [("rect", (41, 27)), ("cube", (9,))])
def test_subdomain_restricted_terms_on_device_mesh(kind, args):
    """dS terms restricted to a cell subdomain on a mesh whose cells live in HBM: the
    per-record bits are built on the device (fvm_record_mask)."""
    import oracle
    import pyfvm_b200 as eng
    from pyfvm_b200 import device_meshes as dm
    from pyfvm_b200 import meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    if kind == "cube":
        p = p * numpy.array([2.0, 1.0, 1.0])  # LeftHalf is x < 1
    p, c = meshes.jitter(p, c, scale=0.15, seed=3)
    h = meshes.Mesh(p, c)
    d = dm.DeviceSimplexMesh.from_host(p, c)
    s = {}
    A_ref, b_ref = oracle.discretize_linear(problems.subdomain_problem(oracle.form_language), h, scales=s)
    A, b = eng.discretize_linear(problems.subdomain_problem(eng.form_language), d)
    parity.assert_pattern_equal(A, A_ref)
    assert (numpy.abs(A.data - A_ref.data) <= 4 * parity.RTOL * s["matrix"].data).all()
    bound = 4 * parity.RTOL * numpy.maximum(s["rhs"], numpy.abs(b_ref))
    assert (numpy.abs(b - b_ref) <= bound).all()
