"""Development tool: sweep the tile kernel's launch shape on one mesh.

  python tools/sweep_tile.py --size 4001 --quantum 1536,2048,3072 --warps 4,8 --stages 2,3

Prints one line per combination: ms per assembly launch, algorithmic GB/s and
registers.  Not part of the product or the tests."""
import argparse
import itertools
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from pyfvm_b200 import codegen, engine, form_language, lowering, meshes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4001)
    ap.add_argument("--cube", type=int, default=0, help="use a cube_tetra mesh with this many points per side")
    ap.add_argument("--quantum", default="3072")
    ap.add_argument("--warps", default="8")
    ap.add_argument("--stages", default="2")
    ap.add_argument("--ctas", default="0")
    ap.add_argument("--factor", default="1")
    ap.add_argument("--fmad", type=int, default=0)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    ints = lambda s: [int(x) for x in s.split(",")]

    eng = engine.Engine.get()
    if args.cube:
        xs = numpy.linspace(0.0, 1.0, args.cube)
        mesh = meshes.Mesh(*meshes.cube_tetra(xs, xs, xs))
        from tests import problems

        prob = problems.poisson(form_language) if args.mode == 0 else problems.bratu(form_language)
    else:
        xs = numpy.linspace(0.0, 1.0, args.size)
        mesh = meshes.Mesh(*meshes.rectangle_tri(xs, xs))
        prob = bench.problem_c4(form_language)
    dim = 3 if args.cube else 2
    lowered = (lowering.lower_linear if args.mode == 0 else lowering.lower_nonlinear)(prob, dim)
    n = len(mesh.points)
    dirid = numpy.zeros(n, dtype=numpy.uint8)
    dirid[mesh.is_boundary_point] = 1
    d_dir = eng.to_device(dirid)
    u = eng.to_device(numpy.random.default_rng(0).standard_normal(n))
    for q in ints(args.quantum):
        dm = engine.DeviceMesh(eng, mesh, need_el=False, tile_quantum=q)
        print(dm.stats(), flush=True)
        data = eng.alloc(8 * dm.nnz)
        vec = eng.alloc(8 * n)
        for w, st, ct, fc in itertools.product(
            ints(args.warps), ints(args.stages), ints(args.ctas), ints(args.factor)
        ):
            fun = codegen.generate(
                lowered, args.mode, [None] * len(lowered.edge), [None] * len(lowered.vertex),
                tune={"consumer_warps": w, "stages": st, "ctas_per_sm": ct, "factor_ce": fc},
            )
            try:
                k = eng.kernel(fun, fmad=bool(args.fmad))
                for _ in range(5):
                    dm.launch(k, dirid=d_dir.ptr, u=u.ptr, data=data.ptr, vec=vec.ptr)
                eng.timer_begin()
                for _ in range(args.steps):
                    dm.launch(k, dirid=d_dir.ptr, u=u.ptr, data=data.ptr, vec=vec.ptr)
                ms = eng.timer_end() / args.steps
            except engine.FvmError as e:
                print(f"q={q} warps={w} stages={st} ctas={ct} factor={fc}: FAILED {str(e)[:100]}")
                continue
            alg = bench.algorithmic_bytes(
                dm.R, dm.nnz, n, dim, fun.uses_x, fun.uses_el, "assembly" if args.mode != 1 else "residual"
            )
            print(
                f"q={q} warps={w} stages={st} ctas={ct} factor={fc} fmad={args.fmad}: {ms:.4f} ms  "
                f"{alg / ms / 1e6:.0f} GB/s alg  {dm.R / ms / 1e6:.1f} Gcontrib/s  "
                f"{k.info(dm)} tiles={dm.n_tiles}",
                flush=True,
            )
        del dm, data, vec


if __name__ == "__main__":
    main()
