"""Development tool: compile the tile kernel for a problem OFFLINE (no GPU) exactly as
NVRTC would see it, print registers / spills, and keep the SASS + an opcode histogram.

  python tools/offline_sass.py --problem c4 --mode 1 [--cell-dim 2] [--out /tmp/k] [-D FVM_X=1]

--problem: c4 (convection-diffusion), poisson, bratu; --mode: 0 linear, 1 residual,
2 Jacobian, 3 SpMV.  Used to check instruction counts before spending GPU time and to
produce the opcode histograms under profiles/."""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from pyfvm_b200 import codegen, form_language, lowering  # noqa: E402
from tests import problems  # noqa: E402

SPMV_PRELUDE = None


def source(problem, mode, cell_dim, tune):
    csrc = os.path.join(ROOT, "pyfvm_b200", "csrc")
    if mode == 3:
        import re as _re

        cu = open(os.path.join(csrc, "fvm_b200.cu")).read()
        m = _re.search(r"static const char\* prelude =\n(.*?);\n", cu, _re.S)
        prelude = "".join(eval(l.split("//")[0].strip()) for l in m.group(1).splitlines() if l.strip())
    else:
        prob = {
            "c4": bench.problem_c4, "poisson": problems.poisson, "bratu": problems.bratu,
            "reaction_x": problems.reaction_x, "nonlinear_conv": problems.nonlinear_conv,
        }[problem](form_language)
        dim = 3 if cell_dim == 3 else 2
        low = (lowering.lower_linear if mode == 0 else lowering.lower_nonlinear)(prob, dim)
        t = codegen.tuning(cell_dim)
        t.update(tune)
        fun = codegen.generate(low, mode, [None] * len(low.edge), [None] * len(low.vertex), t)
        prelude = fun.source
    return (
        open(os.path.join(csrc, "fvm_tile_args.h")).read() + "\n" + prelude + "\n"
        + open(os.path.join(csrc, "fvm_tile_kernel.cuh")).read()
    )


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="c4")
    ap.add_argument("--mode", type=int, default=1)
    ap.add_argument("--cell-dim", type=int, default=2)
    ap.add_argument("--out", default="/tmp/fvm_offline")
    ap.add_argument("-D", action="append", default=[], help="tuning override name=value")
    ap.add_argument("--hist", action="store_true", help="print the opcode histogram")
    args = ap.parse_args()
    tune = {k: int(v) for k, v in (d.split("=") for d in args.D)}
    src = source(args.problem, args.mode, args.cell_dim, tune)
    cu, cubin = args.out + ".cu", args.out + ".cubin"
    open(cu, "w").write(src)
    r = subprocess.run(
        ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "--fmad=false", "-Xptxas", "-v", "-cubin", "-o", cubin, cu],
        capture_output=True, text=True,
    )
    if r.returncode:
        print(r.stderr)
        sys.exit(1)
    for l in r.stderr.splitlines():
        if "registers" in l or "spill" in l:
            print(l.strip())
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    open(args.out + ".sass", "w").write(sass)
    ops = collections.Counter()
    for l in sass.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    print("instructions:", sum(ops.values()))
    if args.hist:
        for k, v in ops.most_common():
            print(f"  {k:12s} {v}")
    print("sass:", args.out + ".sass")


if __name__ == "__main__":
    main()
