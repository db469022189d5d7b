cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# memcheck on a small but complete pass through the path (one tool per call)
cat > /tmp/san.py <<'PY'
import sys, numpy
sys.path.insert(0, '.')
import pyfvm_b200 as e
from pyfvm_b200 import meshes, solvers, problem, device_meshes as dm
from tests import problems
for kind, args, shuffle in (("rect", (21, 13), False), ("cube", (7,), False), ("rect", (19, 14), True)):
    p, c = (problems.rectangle if kind == "rect" else problems.cube)(*args)
    p, c = meshes.jitter(p, c, scale=0.15, seed=1)
    if shuffle:
        p, c = problems.shuffle_numbering(p, c, seed=2)
    m = meshes.Mesh(p, c)
    n = len(p)
    for name in ("reaction_x", "subdomain_problem" if kind == "rect" else "poisson", "robin"):
        A, b = e.discretize_linear(getattr(problems, name)(e.form_language), m)
    f, j = e.discretize(problems.nonlinear_conv(e.form_language), m)
    u = 0.1 * numpy.random.default_rng(0).standard_normal(n)
    f.eval(u); J = j.get_linear_operator(u)
    lp = problem.LinearProblem(problems.convection_diffusion(e.form_language), m)
    lp.assemble()
    x, it, rr = solvers.solve_linear(lp, tol=1e-8)
    d = dm.DeviceSimplexMesh.from_host(p, c)
    d.is_boundary_point
    e.discretize_linear(problems.poisson(e.form_language), d)
print("SANITIZER_RUN_DONE")
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/r2_19_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_19_memcheck.log
tail -15 gpurun_out/r2_19_memcheck.log
