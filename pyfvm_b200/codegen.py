"""sympy -> CUDA device functors (the back end that replaces ``sympy.lambdify``
in the reference's discretize step; SURVEY §2 "back end replaced", App. A.6).

``generate(lowered, mode, ...)`` prints the prelude (``#define`` switches + the
three functors ``fvm_edge`` / ``fvm_vertex`` / ``fvm_dirichlet``) that is
prepended to ``csrc/fvm_tile_kernel.cuh`` and compiled by NVRTC for sm_100a.
"""
import os
from dataclasses import dataclass

import sympy
from sympy.printing.c import C99CodePrinter

from . import lowering as lo

MODE_LINEAR, MODE_RESIDUAL, MODE_JACOBIAN = 0, 1, 2

# FVM_F_* bits of csrc/fvm_tile_args.h (= FVM_KF_* of include/fvm_b200.h)
F_EDGE_EL, F_EDGE_MASK, F_X, F_U, F_CV, F_VMASK, F_DIR = 1, 2, 4, 8, 16, 32, 64
F_EDGE_D, F_EDGE_C, F_SA = 128, 256, 512


def tuning(cell_dim=2):
    """Launch-shape knobs of the tile kernel (env overrides are for experiments).

    None of them changes a result bit: the fold order of a slot / row is fixed
    by the mesh alone."""
    return {
        "consumer_warps": int(os.environ.get("FVM_CONSUMER_WARPS", "4")),
        "stages": int(os.environ.get("FVM_STAGES", "2")),
        # producer warps per CTA: the stage fill is split into four roles (half-edge streams,
        # row streams, coordinate table, u table) issued in parallel
        "producers": int(os.environ.get("FVM_PRODUCERS", "1")),
        "ctas_per_sm": int(os.environ.get("FVM_CTAS_PER_SM", "0")),
        # fold the ce_ratios of a slot first and evaluate ce*g(slot) once
        "factor_ce": int(os.environ.get("FVM_FACTOR_CE", "1")),
        # resident CTAs per SM the register allocation must allow (__launch_bounds__)
        "min_ctas": int(os.environ.get("FVM_MIN_CTAS", "0")),
        # multi-pair slots: pairs loaded before the in-order adds start (-1: template default)
        "fold_ahead": int(os.environ.get("FVM_FOLD_AHEAD", "-1")),
        # row pass: equal runs of rows per warp (triangles: ~104 rows per tile, 26 per warp)
        # or 32-row blocks dealt to the warps (tetrahedra: ~17 rows per tile, one warp)
        # slot pass: 32-slot chunks a warp keeps in flight together; row pass: shares loaded
        # ahead of the in-order adds (-1: template defaults)
        "ilp": int(os.environ.get("FVM_ILP", "-1")),
        "setmaxnreg": int(os.environ.get("FVM_SETMAXNREG", "-1")),
        # math warps (of consumer_warps) that only fold rows; the others only fold slots
        "row_warps": int(os.environ.get("FVM_ROW_WARPS", "-1")),
        # tiles ahead of the shared-memory fill whose streams are prefetched into L2
        "l2_ahead": int(os.environ.get("FVM_L2_AHEAD", "-1")),
        "row_ahead": int(os.environ.get("FVM_ROW_AHEAD", "-1")),
        # (2: decided per tile by its row count -- the template's default)
        "row_runs": int(os.environ.get("FVM_ROW_RUNS", "2")),
    }

# functor outputs per mode: (edge D, O, C), (vertex L, C), (dirichlet coeff, val)
_EDGE_KEYS = {
    MODE_LINEAR: ("diag", "off", "aff"),
    MODE_RESIDUAL: (None, None, "val"),
    MODE_JACOBIAN: ("diag", "off", None),
}
_VERT_KEYS = {
    MODE_LINEAR: ("lin", "aff"),
    MODE_RESIDUAL: (None, "val"),
    MODE_JACOBIAN: ("lin", None),
}
_DIR_KEYS = {
    MODE_LINEAR: ("coeff", "rhs"),
    MODE_RESIDUAL: (None, "val"),
    MODE_JACOBIAN: ("coeff", None),
}


class _DevicePrinter(C99CodePrinter):
    """C99 printer with numpy-compatible small powers: ``x**2`` is printed as
    ``x*x`` (what ``numpy.power``'s fast path computes) instead of ``pow``."""

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= int(e) <= 3 and (b.is_Symbol or b.is_Number):
            s = self._print(b)
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and int(e) == -2 and b.is_Symbol:
            s = self._print(b)
            return f"(1.0/({s}*{s}))"
        return super()._print_Pow(expr)


_printer = _DevicePrinter()


def _ccode(e):
    return _printer.doprint(sympy.sympify(e))


def _emit_block(lines, targets, exprs, indent="    ", assigned=None):
    """``target (+)= expr`` statements for one term, sharing subexpressions.
    cse temporaries are named ``t<i>`` so they cannot collide with the edge
    endpoint symbols ``x0_k``/``x1_k``.  ``assigned``: set of
    targets already written; the first unconditional write of a target is a
    plain assignment (the caller zero-initialises, ``0.0 + x`` is not free in
    fp64)."""
    live = [(t, sympy.sympify(e)) for t, e in zip(targets, exprs) if t is not None]
    live = [(t, e) for t, e in live if e != 0]
    if not live:
        return
    repl, red = sympy.cse(
        [e for _, e in live], symbols=sympy.numbered_symbols("t"), order="none"
    )
    lines.append(indent + "{")
    for sym, sub in repl:
        lines.append(f"{indent}  const double {sym} = {_ccode(sub)};")
    for (t, _), e in zip(live, red):
        op = "+="
        if assigned is not None and t not in assigned:
            op = "="
            assigned.add(t)
        lines.append(f"{indent}  {t} {op} {_ccode(e)};")
    lines.append(indent + "}")


def _emit_masked(lines, bit, targets, exprs, assigned):
    block = []
    # a subdomain-restricted term is conditional: it must accumulate
    _emit_block(block, targets, exprs, indent="  ", assigned=assigned if bit is None else None)
    if bit is not None:
        assigned.update(t for t in targets if t is not None)
    if block and bit is not None:
        lines.append(f"  if (mask & {1 << bit}u)")
    lines.extend(block)


def _is_linear_in_ce(e):
    """``e == edge_ce_ratio * g`` with ``g`` free of ``edge_ce_ratio``."""
    e = sympy.sympify(e)
    if e == 0:
        return True
    if e.subs(lo.CE, 0) != 0:
        return False
    return lo.CE not in sympy.diff(e, lo.CE).free_symbols


def _uses(exprs, symbols):
    fs = set()
    for e in exprs:
        fs |= sympy.sympify(e).free_symbols
    return bool(fs & set(symbols))


@dataclass
class Functors:
    source: str  # prelude to prepend to the tile-kernel template
    mode: int
    flags: int  # FVM_F_* bits: which streams the kernel stages / the host must supply
    consumer_warps: int
    stages: int
    ctas_per_sm: int
    producers: int
    factored: bool
    uses_el: bool
    uses_x: bool
    edge_masked: bool
    vert_masked: bool
    has_dirichlet: bool
    needs_u: bool
    uses_sa: bool = False


def generate(lowered, mode, edge_bits, vert_bits, tune=None):
    """Print the functor prelude for ``mode``.

    ``edge_bits[i]`` / ``vert_bits[i]`` is the subdomain mask bit of edge /
    vertex term ``i`` or ``None`` when the term applies everywhere."""
    ek, vk, dk = _EDGE_KEYS[mode], _VERT_KEYS[mode], _DIR_KEYS[mode]
    edge_exprs = [[t.outputs[k] if k else 0 for k in ek] for t in lowered.edge]
    vert_exprs = [[t.outputs[k] if k else 0 for k in vk] for t in lowered.vertex]
    dir_exprs = [[t.outputs[k] if k else 0 for k in dk] for t in lowered.dirichlet]

    flat_e = [e for ex in edge_exprs for e in ex]
    flat_v = [e for ex in vert_exprs + dir_exprs for e in ex]
    xv = lo.X0 + lo.X1
    tune = {**tuning(), **(tune or {})}
    edge_x, vert_x = _uses(flat_e, xv), _uses(flat_v, lo.XV)
    edge_u, vert_u = _uses(flat_e, [lo.UK0, lo.UK1]), _uses(flat_v, [lo.UK0])
    flags = 0
    flags |= F_EDGE_EL if _uses(flat_e, [lo.EL]) else 0
    flags |= F_EDGE_MASK if any(b is not None for b in edge_bits) else 0
    flags |= F_X if (edge_x or vert_x) else 0
    flags |= F_U if (edge_u or vert_u) else 0
    flags |= F_CV if _uses(flat_v, [lo.CV]) else 0
    flags |= F_SA if _uses(flat_v, [lo.SA]) else 0
    flags |= F_VMASK if any(b is not None for b in vert_bits) else 0
    flags |= F_DIR if lowered.dirichlet else 0
    # which row sums the edge functor feeds in this mode (their per-slot shares
    # are parked in shared memory between the slot pass and the row pass)
    flags |= F_EDGE_D if any(sympy.sympify(ex[0]) != 0 for ex in edge_exprs) else 0
    flags |= F_EDGE_C if any(sympy.sympify(ex[2]) != 0 for ex in edge_exprs) else 0
    # Every dS integrand is multiplied by edge_ce_ratio (App. A.2), so the edge
    # outputs are ce * g(x0, x1, u0, u1).  When nothing else varies per half-edge
    # (no edge_length, no subdomain bits) the kernel sums the ce_ratios of a slot
    # and evaluates the functor once with that sum.
    factored = bool(
        tune.get("factor_ce", 1)
        and flat_e
        and not (flags & (F_EDGE_EL | F_EDGE_MASK))
        and all(_is_linear_in_ce(e) for e in flat_e)
    )
    defs = {
        "FVM_MODE": mode,
        "FVM_DIM": lowered.dim,
        "FVM_FLAGS": f"{flags}u",
        "FVM_CONSUMER_WARPS": tune["consumer_warps"],
        "FVM_STAGES": tune["stages"],
        "FVM_PRODUCERS": tune["producers"],
        "FVM_EDGE_FACTORED": int(factored),
        "FVM_MIN_CTAS": tune["min_ctas"],
        "FVM_ROW_RUNS": tune["row_runs"],
        "FVM_EDGE_USES_X": int(edge_x),
        "FVM_EDGE_USES_U": int(edge_u),
        "FVM_VERT_USES_X": int(vert_x),
        "FVM_VERT_USES_U": int(vert_u),
        "FVM_HAS_VERTEX": int(bool(lowered.vertex)),
    }
    bad = set()
    for e in flat_e + flat_v:
        bad |= {f.func.__name__ for f in sympy.sympify(e).atoms(sympy.core.function.AppliedUndef)}
    if bad:
        raise ValueError(f"undefined functions cannot be lowered to device code: {sorted(bad)}")

    for key, macro in (("fold_ahead", "FVM_FOLD_AHEAD"), ("ilp", "FVM_ILP"), ("row_ahead", "FVM_ROW_AHEAD"),
                       ("setmaxnreg", "FVM_SETMAXNREG"), ("l2_ahead", "FVM_L2_AHEAD"),
                       ("row_warps", "FVM_ROW_WARPS")):
        if tune.get(key, -1) >= 0:
            defs[macro] = tune[key]
    L = [f"// generated by pyfvm_b200.codegen (mode {mode}); do not edit"]
    L += [f"#define {k} {v}" for k, v in defs.items()]
    L.append("")
    L.append(
        "__device__ __forceinline__ void fvm_edge(const double uk0, const double uk1,\n"
        "    const double x0_0, const double x0_1, const double x0_2,\n"
        "    const double x1_0, const double x1_1, const double x1_2,\n"
        "    const double edge_ce_ratio, const double edge_length, const unsigned mask,\n"
        "    double& D, double& O, double& C) {"
    )
    done = set()
    for ex, bit in sorted(zip(edge_exprs, edge_bits), key=lambda eb: eb[1] is not None):
        _emit_masked(L, bit, ["D", "O", "C"], ex, done)
    L.append("}")
    L.append("")
    L.append(
        "__device__ __forceinline__ void fvm_vertex(const double uk0, const double control_volume,\n"
        "    const double surface_area, const double x_0, const double x_1, const double x_2,\n"
        "    const unsigned mask,\n"
        "    double& L, double& C) {"
    )
    done = set()
    for ex, bit in sorted(zip(vert_exprs, vert_bits), key=lambda eb: eb[1] is not None):
        _emit_masked(L, bit, ["L", "C"], ex, done)
    L.append("}")
    L.append("")
    L.append(
        "__device__ __forceinline__ void fvm_dirichlet(const int id, const double uk0,\n"
        "    const double control_volume, const double surface_area,\n"
        "    const double x_0, const double x_1, const double x_2,\n"
        "    double& coeff, double& val) {"
    )
    if dir_exprs:
        L.append("  switch (id) {")
        for k, ex in enumerate(dir_exprs):
            L.append(f"  case {k + 1}:")
            _emit_block(L, ["coeff", "val"], ex, indent="    ", assigned=set())
            L.append("    break;")
        L.append("  default: break;")
        L.append("  }")
    L.append("}")
    L.append("")
    return Functors(
        source="\n".join(L),
        mode=mode,
        flags=flags,
        consumer_warps=tune["consumer_warps"],
        stages=tune["stages"],
        ctas_per_sm=tune["ctas_per_sm"],
        producers=tune["producers"],
        factored=factored,
        uses_el=bool(flags & F_EDGE_EL),
        uses_x=bool(flags & F_X),
        edge_masked=bool(flags & F_EDGE_MASK),
        vert_masked=bool(flags & F_VMASK),
        has_dirichlet=bool(flags & F_DIR),
        needs_u=bool(flags & F_U),
        uses_sa=bool(flags & F_SA),
    )
