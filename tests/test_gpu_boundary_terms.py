"""GPU tier, SURVEY §8f rank 3: ``dGamma`` boundary integrals (Neumann / Robin terms, per
boundary vertex with ``mesh.surface_areas``) and user-written numeric kernels
(``pyfvm.fvm_matrix.EdgeMatrixKernel``) through the same tile kernel, against the oracle."""
import numpy
import pytest

from . import parity, problems

pytestmark = pytest.mark.gpu


def _mesh(kind, args, jitter=True, shuffle=False):
    from pyfvm_b200 import meshes

    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    if jitter:
        p, c = meshes.jitter(p, c, scale=0.15, seed=4)
    if shuffle:
        p, c = problems.shuffle_numbering(p, c, seed=4)
    return meshes.Mesh(p, c)


@pytest.mark.parametrize(
    "name,kind,args,shuffle",
    [("robin", "rect", (65, 41), False), ("robin", "cube", (12,), False),
     ("neumann_only", "rect", (50, 33), False), ("robin", "rect", (37, 29), True)],
)
def test_dgamma_linear(name, kind, args, shuffle):
    import oracle
    import pyfvm_b200 as eng

    mesh = _mesh(kind, args, shuffle=shuffle)
    s = {}
    A_ref, b_ref = oracle.discretize_linear(getattr(problems, name)(oracle.form_language), mesh, scales=s)
    A, b = eng.discretize_linear(getattr(problems, name)(eng.form_language), mesh)
    parity.assert_matrix_close(A, A_ref, s["matrix"])
    parity.assert_vector_close(b, b_ref, s["rhs"])


@pytest.mark.parametrize("kind,args", [("rect", (45, 31)), ("cube", (9,))])
def test_dgamma_nonlinear(kind, args):
    import oracle
    import pyfvm_b200 as eng

    mesh = _mesh(kind, args)
    n = len(mesh.points)
    f_ref, jac_ref = oracle.discretize(problems.radiation(oracle.form_language), mesh)
    f, jac = eng.discretize(problems.radiation(eng.form_language), mesh)
    u = 0.5 + 0.2 * numpy.random.default_rng(2).standard_normal(n)
    s = {}
    F_ref = f_ref.eval(u, scales=s)
    parity.assert_vector_close(f.eval(u), F_ref, s["vec"])
    J_ref = jac_ref.get_linear_operator(u, scales=s)
    parity.assert_matrix_close(jac.get_linear_operator(u), J_ref, s["matrix"])


def test_robin_newton_solution():
    """The nonlinear boundary term through ``pyfvm.newton`` with a host ``spsolve``: the
    engine and the oracle must walk the same Newton iterates."""
    import oracle
    import pyfvm_b200 as eng
    from scipy.sparse import linalg

    mesh = _mesh("rect", (41, 27), jitter=False)
    n = len(mesh.points)

    def solve(mod, fl):
        f, jac = mod.discretize(problems.radiation(fl), mesh)
        return mod.newton(
            f.eval, lambda u0, rhs: linalg.spsolve(jac.get_linear_operator(u0), rhs),
            numpy.ones(n), verbose=False,
        )

    u_gpu, u_cpu = solve(eng, eng.form_language), solve(oracle, oracle.form_language)
    assert numpy.abs(u_gpu - u_cpu).max() <= 1e-12


def _kernels(mod):
    """An anisotropic-diffusion edge kernel, a lumped-mass vertex kernel and a Dirichlet
    kernel written the way pyfvm.fvm_matrix users write them: plain numpy on mesh arrays."""

    class Diffusion(mod.EdgeMatrixKernel):
        def __init__(self, subdomains=None, k=1.0):
            super().__init__()
            self.subdomains = subdomains or [None]
            self.k = k

        def eval(self, mesh, cell_mask):
            ce = mesh.ce_ratios[..., cell_mask]
            X = mesh.node_coords[mesh.idx_hierarchy[..., cell_mask]]
            a0 = self.k * (1.0 + X[0][..., 0] ** 2)  # coefficient seen from either endpoint
            a1 = self.k * (1.0 + X[1][..., 0] ** 2)
            return numpy.array([[ce * a0, -ce * a0], [-ce * a1, ce * a1]])

    class Mass:
        subdomains = [None]

        def eval(self, mesh, vertex_mask):
            return 3.0 * mesh.control_volumes[vertex_mask]

    class Fixed:
        def __init__(self, subdomain):
            self.subdomain = subdomain

        def eval(self, mesh, vertex_mask):
            return 2.0 + mesh.node_coords[vertex_mask][:, 1]

    return Diffusion, Mass, Fixed


@pytest.mark.parametrize("kind,args", [("rect", (39, 27)), ("cube", (8,))])
def test_fvm_matrix_user_kernels(kind, args):
    import oracle
    import pyfvm_b200 as eng

    mesh = _mesh(kind, args)
    out = {}
    for mod in (oracle, eng):
        Diffusion, Mass, Fixed = _kernels(mod)
        kw = {"scales": out} if mod is oracle else {}
        out[mod.__name__] = mod.get_fvm_matrix(
            mesh,
            edge_kernels=[Diffusion(), Diffusion(subdomains=[problems.LeftHalf()], k=0.5)],
            vertex_kernels=[Mass()],
            dirichlets=[Fixed(eng.form_language.Boundary() if mod is eng else oracle.form_language.Boundary())],
            **kw,
        )
    parity.assert_matrix_close(out["pyfvm_b200"], out["oracle"], out["matrix"])


def test_fvm_matrix_restricted_kernel_shrinks_the_pattern():
    import oracle
    import pyfvm_b200 as eng

    mesh = _mesh("rect", (25, 17))
    out = {}
    for mod in (oracle, eng):
        Diffusion, _, _ = _kernels(mod)
        kw = {"scales": out} if mod is oracle else {}
        Diffusion, Mass, _ = _kernels(mod)
        # (a vertex kernel everywhere: the engine stores the diagonal slot of EVERY row,
        # scipy only those something contributes to)
        out[mod.__name__] = mod.get_fvm_matrix(
            mesh, edge_kernels=[Diffusion(subdomains=[problems.LeftHalf()])],
            vertex_kernels=[Mass()], **kw
        )
    parity.assert_matrix_close(out["pyfvm_b200"], out["oracle"], out["matrix"])
