"""Development tool: time the tile kernel on one mesh for several modes and tuning
variants (same box, back to back: kernel changes are judged by same-box A/B only).

  python tools/sweep_tile.py --size 4001 --modes 0,1,2,3 --quantum 3432 \
      --variants "row_runs=0;row_runs=1;fold_ahead=0,consumer_warps=8"
  python tools/sweep_tile.py --cube 200 --modes 1,2 --variants "fold_ahead=0;fold_ahead=5"

A variant is a comma-separated list of codegen.tuning overrides; "" is the default.
Mode 3 = SpMV on the assembled system (needs mode 0 in the same run order-independent:
it assembles once itself).  Prints one line per (quantum, mode, variant): ms per launch,
algorithmic GB/s, fraction of the measured HBM peak.  Not part of the product or tests."""
import argparse
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from pyfvm_b200 import codegen, engine, form_language, lowering, meshes, solvers  # noqa: E402
from tests import problems  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4001)
    ap.add_argument("--cube", type=int, default=0, help="cube_tetra mesh (generated on the device)")
    ap.add_argument("--quantum", default="0")
    ap.add_argument("--modes", default="0,1")
    ap.add_argument("--variants", default="")
    ap.add_argument("--fmad", type=int, default=0)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--steering", default="1", help="FVM_BANK_STEERING values to build meshes with")
    args = ap.parse_args()
    ints = lambda s: [int(x) for x in s.split(",")]
    variants = []
    for v in args.variants.split(";"):
        variants.append({k: int(x) for k, x in (kv.split("=") for kv in v.split(",") if kv)})
    peak, _ = bench.measured_peak()

    eng = engine.Engine.get()
    if args.cube:
        from pyfvm_b200 import device_meshes

        xs = numpy.linspace(0.0, 1.0, args.cube)
        mesh = device_meshes.device_cube_tetra(xs, xs, xs)
        lin, nonlin, dim, cdim = problems.poisson, problems.bratu, 3, 3
    else:
        xs = numpy.linspace(0.0, 1.0, args.size)
        mesh = meshes.Mesh(*meshes.rectangle_tri(xs, xs))
        lin = nonlin = bench.problem_c4
        dim, cdim = 2, 2
    n = len(mesh.points)
    dirid = numpy.zeros(n, dtype=numpy.uint8)
    dirid[mesh.is_boundary_point] = 1
    d_dir = eng.to_device(dirid, pad=16)
    u = eng.to_device(0.1 * numpy.random.default_rng(0).standard_normal(n), pad=16)
    low = {0: lowering.lower_linear(lin(form_language), dim)}
    low[1] = low[2] = lowering.lower_nonlinear(nonlin(form_language), dim)
    # every (quantum, steering) mesh is built first; then, per mode, every (mesh, variant)
    # pair is timed round-robin `reps` times: clocks and box state drift over a sweep (and from
    # process to process), so alternatives are only ever compared on interleaved samples
    built = []
    for q, steer in [(q, s) for q in ints(args.quantum) for s in ints(args.steering)]:
        os.environ["FVM_BANK_STEERING"] = str(steer)  # read when the mesh is tiled
        dm = engine.DeviceMesh(eng, mesh, need_el=False, tile_quantum=q)
        data, vec = eng.alloc(8 * dm.nnz + 16), eng.alloc(8 * n + 16)
        # (the vertex tables are bank-steered on a mesh's second launch, with the switch as it
        # is set then: do those launches now)
        k0 = eng.kernel(codegen.generate(low[0], 0, [None], [None], codegen.tuning(cdim)))
        for _ in range(2):
            dm.launch(k0, dirid=d_dir.ptr, u=u.ptr, data=data.ptr, vec=vec.ptr)
        eng.sync()
        st = dm.stats()
        print(f"steering={steer}", st, f"tiles={dm.n_tiles} build_ms={dm.build_ms:.1f}", flush=True)
        built.append((f"q={st['quantum']} steer={steer}", dm, data, vec))
    for mode in ints(args.modes):
        runs = []
        for label, dm, data, vec in built:
            R = dm.n_halfedges / 2.0
            for tv in variants:
                tune = codegen.tuning(cdim)
                tune.update(tv)
                try:
                    if mode == 3:
                        if tv:
                            continue  # the SpMV kernel has a fixed prelude
                        k0 = eng.kernel(codegen.generate(low[0], 0, [None], [None], codegen.tuning(cdim)))
                        dm.launch(k0, dirid=d_dir.ptr, u=u.ptr, data=data.ptr, vec=vec.ptr)
                        sysm = solvers.DeviceSystem(dm, data.ptr)
                        y = eng.alloc(8 * n + 16)
                        run = lambda sysm=sysm, y=y: sysm.matvec(u.ptr, y.ptr)
                        alg = 12 * dm.nnz + 20 * n
                        info = {}
                    else:
                        lw = low[mode]
                        fun = codegen.generate(
                            lw, mode, [None] * len(lw.edge), [None] * len(lw.vertex), tune
                        )
                        k = eng.kernel(fun, fmad=bool(args.fmad))
                        run = lambda k=k, dm=dm, data=data, vec=vec: dm.launch(
                            k, dirid=d_dir.ptr, u=u.ptr, data=data.ptr, vec=vec.ptr)
                        alg = bench.algorithmic_bytes(
                            R, dm.nnz, n, dim, fun.uses_x, fun.uses_el,
                            "residual" if mode == 1 else "assembly",
                        )
                        info = k.info(dm)
                    for _ in range(5):
                        run()
                    eng.sync()
                except engine.FvmError as e:
                    print(f"{label} mode={mode} {tv}: FAILED {str(e)[:200]}", flush=True)
                    continue
                runs.append((label, tv, run, alg, info, []))
        for _ in range(args.reps):
            for label, tv, run, alg, info, ms in runs:
                eng.timer_begin()
                for _ in range(args.steps):
                    run()
                ms.append(eng.timer_end() / args.steps)
        for label, tv, run, alg, info, ms in runs:
            med, best = float(numpy.median(ms)), min(ms)
            print(
                f"{label} mode={mode} {tv or 'default'}: median {med:.4f} ms (min {best:.4f})  "
                f"{alg / med / 1e6:.0f} GB/s alg  frac={alg / med / 1e6 / peak:.3f}  "
                f"regs={info.get('regs')} grid={info.get('grid')} smem={info.get('smem')}",
                flush=True,
            )


if __name__ == "__main__":
    main()
