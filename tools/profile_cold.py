import sys, time, cProfile, pstats, numpy
sys.path.insert(0, '/root/repo')
import bench, pyfvm_b200 as pe
from pyfvm_b200 import meshes, engine
xs = numpy.linspace(0, 1, 4001)
mesh = meshes.Mesh(*meshes.rectangle_tri(xs, xs))
mesh.ce_ratios, mesh.control_volumes, mesh.is_boundary_point
A, b = pe.discretize_linear(bench.problem_c4(pe.form_language), mesh)   # compile kernels
engine._mesh_cache.clear()
pr = cProfile.Profile(); pr.enable()
t = time.time()
A, b = pe.discretize_linear(bench.problem_c4(pe.form_language), mesh)
print("cold", time.time() - t)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
