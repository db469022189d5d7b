"""Measure every BASELINE config once on one GPU and print a markdown table
(profiles/r01_configs.md): device time per launch next to the CPU oracle.

  python tools/config_table.py [--skip-cpu]
"""
import argparse
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import oracle  # noqa: E402
import pyfvm_b200 as pe  # noqa: E402
from pyfvm_b200 import codegen, meshes, problem  # noqa: E402
from tests import problems  # noqa: E402


def timed(eng, fn, steps=50, warm=5):
    for _ in range(warm):
        fn()
    eng.timer_begin()
    for _ in range(steps):
        fn()
    return eng.timer_end() / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    peak, _ = bench.measured_peak()
    rows = []

    def linear_case(label, mesh, prob_f, cpu=True):
        lp = problem.LinearProblem(prob_f(pe.form_language), mesh)
        ms = timed(lp.eng, lp.assemble_device)
        R = lp.dmesh.R
        fun = lp.functors[codegen.MODE_LINEAR]
        alg = bench.algorithmic_bytes(R, lp.nnz, lp.n, lp.dmesh.dim, fun.uses_x, fun.uses_el, "assembly")
        cpu_s = None
        if cpu and not args.skip_cpu:
            k = oracle.lower_linear(prob_f(oracle.form_language), mesh)
            t = time.time()
            oracle.get_linear_fvm_problem(mesh, *k)
            cpu_s = time.time() - t
        rows.append((label, "assembly", lp.n, R, lp.nnz, ms, alg / ms / 1e6, alg / ms / 1e6 / peak,
                     lp.dmesh.build_ms, cpu_s))

    def nonlinear_case(label, mesh, prob_f, cpu=True):
        f, jac = pe.discretize(prob_f(pe.form_language), mesh)
        s = f._s
        u = 0.1 * numpy.random.default_rng(0).standard_normal(s.n)
        s.upload_u(u)
        R = s.dmesh.R
        for mode, name, fn in (
            (codegen.MODE_RESIDUAL, "f.eval", lambda: f.eval_device(s.u.ptr)),
            (codegen.MODE_JACOBIAN, "Jacobian refresh", lambda: jac.values_device(s.u.ptr)),
        ):
            ms = timed(s.eng, fn)
            fun = s.functors[mode]
            alg = bench.algorithmic_bytes(R, s.nnz, s.n, s.dmesh.dim, fun.uses_x, fun.uses_el,
                                          "residual" if mode == codegen.MODE_RESIDUAL else "assembly")
            cpu_s = None
            if cpu and not args.skip_cpu:
                fo, jo = oracle.discretize(prob_f(oracle.form_language), mesh)
                t = time.time()
                fo.eval(u) if mode == codegen.MODE_RESIDUAL else jo.get_linear_operator(u)
                cpu_s = time.time() - t
            rows.append((label, name, s.n, R, s.nnz, ms, alg / ms / 1e6, alg / ms / 1e6 / peak,
                         s.dmesh.build_ms, cpu_s))

    linear_case("C1 Poisson 401x201", meshes.Mesh(*problems.rectangle(401, 201)), problems.poisson)
    nonlinear_case("C2 Bratu 101x51", meshes.Mesh(*problems.rectangle(101, 51)), problems.bratu)
    linear_case("C3 Poisson cube 60^3", meshes.Mesh(*problems.cube(60)), problems.poisson)
    from pyfvm_b200 import device_meshes as dm

    xs = numpy.linspace(0.0, 1.0, 4001)
    linear_case("C4 conv-diff 4001^2 (device mesh)", dm.device_rectangle_tri(xs, xs), bench.problem_c4, cpu=False)
    xs = numpy.linspace(0.0, 1.0, 200)
    nonlinear_case("C5 family Bratu cube 200^3 (device mesh)", dm.device_cube_tetra(xs, xs, xs),
                   problems.bratu, cpu=False)

    print("| config | launch | vertices | contributions | nnz | device ms | alg GB/s | frac of measured peak "
          "| pattern build ms | CPU oracle s (1 thread) | speed-up |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for lab, what, n, R, nnz, ms, gbs, frac, bms, cpu in rows:
        cs = "-" if cpu is None else f"{cpu:.3f}"
        sp = "-" if cpu is None else f"{cpu / (ms * 1e-3):.0f}x"
        print(f"| {lab} | {what} | {n} | {R} | {nnz} | {ms:.4f} | {gbs:.0f} | {frac:.2f} | {bms:.1f} | {cs} | {sp} |")


if __name__ == "__main__":
    main()
