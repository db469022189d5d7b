#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json metric).

A "step" is one matrix+rhs assembly pass (one tile-kernel launch) over the
convection-diffusion problem of BASELINE config 4 on a synthetic
``rectangle_tri`` mesh (default 4001 x 4001 points = 16 008 001 vertices,
96 000 000 edge contributions).  One edge contribution = one column of
``idx_hierarchy``.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

What one run reports (one JSON line on rank 0):

* ``value``     : device-resident assembly, K launches timed with CUDA events on the
                  ctx stream; ``sustained`` repeats it for >= 1 s (power-capped clocks);
* ``e2e``       : THE API CALL ``pyfvm.discretize_linear(problem, mesh)`` with host numpy
                  arrays in and scipy CSR + numpy rhs out, on a mesh object the engine has
                  not seen (every input array crosses PCIe, pattern built on the device,
                  pattern + values + rhs come back).  ``e2e_warm`` is the same call on the
                  same mesh object (device mesh cached and revalidated), ``e2e_pipelined``
                  the hand-pipelined geometry re-upload loop of round 1;
* ``residual``  : ``f.eval`` (config 4: "linear assembly + residual eval"), device-resident
                  and through the API with host arrays; ``jacobian`` likewise;
* ``c5``        : the config-5 family (Bratu on a cube_tetra grid, 256^3 by default,
                  generated in HBM): halo push + residual + Jacobian refresh, at every N;
* ``dist_parity`` (N > 1): N-rank results == 1-rank results, bit for bit;
* ``small_configs`` (N = 1): configs 1-3 through the API, first call and warm, next to the
                  CPU oracle; ``cpu_baseline``: the oracle on a bounded sample.

N > 1 (under torchrun): the vertex rows are partitioned into N contiguous slabs, one
rank per GPU; linear assembly needs no collective (strong scaling: total work fixed);
the residual / Jacobian legs push ``u`` halos over NVLink peer memory; times are the
max over ranks.

``--impl reference`` times the CPU oracle (the restated reference algorithm with the
reference's own numpy/scipy calls; the reference itself is a README and cannot be
installed) on bounded samples of the same workload, plus -- memory
permitting -- one step at the full size.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "edge contributions assembled/s (CSR matrix+rhs)"
L2_BYTES = 126e6
UNIT = "contributions/s"


def problem_c4(fl):
    """BASELINE config 4: -n.grad(u) + (a.n) u, a = (2, 1, 0), source 1, u = 0 on the boundary."""
    import sympy

    class ConvectionDiffusion:
        def apply(self, u):
            a = sympy.Matrix([2, 1, 0])
            return fl.integrate(
                lambda x: -fl.n_dot_grad(u(x)) + fl.dot(a.T, fl.n) * u(x), fl.dS
            ) - fl.integrate(lambda x: 1.0, fl.dV)

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    return ConvectionDiffusion()


def problem_bratu(fl):
    """BASELINE configs 2 / 5 [README:95-102]: -n.grad(u) on dS, 2 exp(u) on dV, u = 0 on the boundary."""
    import sympy

    class Bratu:
        def apply(self, u):
            return fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS) - fl.integrate(
                lambda x: 2.0 * sympy.exp(u(x)), fl.dV
            )

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    return Bratu()


def problem_poisson(fl):
    """BASELINE configs 1 / 3 [README:35-40]."""

    class Poisson:
        def apply(self, u):
            return fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS) - fl.integrate(
                lambda x: 1.0, fl.dV
            )

        def dirichlet(self, u):
            return [(lambda x: u(x) - 0.0, fl.Boundary())]

    return Poisson()


def workload_config(size, world, flush_l2):
    """``config`` of the JSON line: the same dict in both arms (ours / reference)."""
    R = 3 * 2 * (size - 1) ** 2
    return {
        "workload": f"convection-diffusion (BASELINE config 4), rectangle_tri {size}x{size}, "
        f"{size**2} vertices, {R} edge contributions, linear assembly (CSR values + rhs)",
        "partition": "1 GPU" if world == 1 else f"{world} contiguous row slabs, no data-path collective",
        "l2": "L2 flushed (256 MB memset) before every timed launch, launches timed one by one"
        if flush_l2
        else "inputs larger than L2 (no flush needed)",
    }


def algorithmic_bytes(R, nnz, n, dim, uses_x, uses_el, mode):
    """Compulsory HBM traffic of one launch."""
    b = 16 * R + (8 * R if uses_el else 0) + (8 * dim * n if uses_x else 0)
    if mode == "assembly":
        return b + 8 * nnz + 17 * n
    return b + 25 * n  # residual: u, cv, F, mask byte


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self):
        """Block until nvidia-smi printed its first sample (it needs ~0.3 s to start)."""
        if self.proc is None:
            return
        for _ in range(60):
            if self.lines:
                return
            time.sleep(0.05)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.wait_first()
        time.sleep(0.06)
        self.proc.terminate()
        self.thread.join(timeout=2)
        inside = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.06]
        rows = inside or [l for _, l in self.lines[-1:]]
        sm, mx, pw, reasons = [], [], [], set()
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(self.NAMES, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(numpy.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(pw) if pw else None,
            "reasons": sorted(reasons),
            "samples": len(sm) if inside else 0,
        }


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(size, quantum):
    """DRAM bytes per assembly launch from the committed ncu capture of THIS kernel
    (profiles/r02_traffic.json), if it was taken on this workload and tile size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        if int(t["size"]) == size and int(t["tile_quantum"]) == quantum:
            return int(t["assembly_dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        pass
    return None


class Dist:
    """torch.distributed plumbing (NCCL) -- barriers and max-over-ranks only."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, values):
        if self.dist is None:
            return list(values)
        import torch

        t = torch.tensor(list(values), device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def build_local_problem(size, rank, world):
    """This rank's slab of the size x size grid (generated locally; the ghost grid
    rows hold every cell that touches an owned vertex) and its linear problem."""
    import pyfvm_b200 as eng
    from pyfvm_b200 import dist as fd
    from pyfvm_b200 import problem

    xs = numpy.linspace(0.0, 1.0, size)
    part = fd.SlabPartition([xs, xs], rank, world)
    lp = problem.LinearProblem(problem_c4(eng.form_language), part.mesh, row_range=part.row_range)
    return part, lp


def timed_launches(eng, D, step, steps, warmup, flush_buf=None):
    """ms per step on the ctx stream (CUDA events); L2 flushed before every launch and
    launches timed one by one when ``flush_buf`` is given."""
    for _ in range(warmup):
        step()
    eng.sync()
    D.barrier()
    if flush_buf is not None:
        ms = 0.0
        # (measured: re-aligning the ranks after every flush changes nothing; off by default)
        align = D.world > 1 and os.environ.get("FVM_BENCH_ALIGN", "0") != "0"
        for _ in range(steps):
            eng._ck(eng.lib.fvm_dev_memset(eng.ctx, flush_buf.ptr, 0, flush_buf.nbytes))
            if align:
                # the flush takes a different time on every rank; without re-aligning the ranks a
                # step that waits for its neighbours' halo would be charged their skew
                eng.sync()
                D.barrier()
            eng.timer_begin()
            step()
            ms += eng.timer_end()
    else:
        eng.timer_begin()
        for _ in range(steps):
            step()
        ms = eng.timer_end()
    D.barrier()
    return ms / steps


def leg_c5(args, D, N):
    """BASELINE config 5 family: Bratu on a cube_tetra grid of N^3 points, vertex rows
    partitioned into z-slabs across the GPUs, every slab generated and its geometry
    computed in HBM.  A step = one Newton step's device work: halo push of u + residual
    f.eval + Jacobian value refresh on the fixed pattern."""
    import pyfvm_b200 as pe
    from pyfvm_b200 import dist as fd

    xs = numpy.linspace(0.0, 1.0, N)
    t_setup = time.time()
    part = fd.SlabPartition([xs, xs, xs], D.rank, D.world, device=True)
    dp = fd.DistributedFvm(problem_bratu(pe.form_language), part, halo="peer")
    eng, s, dm = dp.eng, dp.s, dp.s.dmesh
    setup_s = time.time() - t_setup
    cells = 5 * (N - 1) ** 3
    R_global = 12 * cells
    R_local = dm.n_halfedges / 2.0
    n_rows, nnz = dm.n_rows, dm.nnz
    rng = numpy.random.default_rng(D.rank)
    for _ in (0, 1):  # both u buffers get owned values (Newton's update writes them in place)
        dp.set_owned(0.1 * rng.standard_normal(part.n_owned))
        dp.residual_device(dp.exchange())
    launches0 = eng.launches
    from pyfvm_b200 import codegen

    ms_res = timed_launches(eng, D, lambda: dp.step_device(codegen.MODE_RESIDUAL), args.steps, args.warmup)
    u_ptr = dp.exchange()
    ms_jac = timed_launches(eng, D, lambda: dp.jacobian_device(u_ptr), args.steps, args.warmup)

    def newton_device_work():
        p = dp.exchange()
        dp.residual_device(p)
        dp.jacobian_device(p)

    ms_step = timed_launches(eng, D, newton_device_work, args.steps, args.warmup)
    launches = eng.launches - launches0
    dp.halo.check_timeout()
    ms_res, ms_jac, ms_step = D.max([ms_res, ms_jac, ms_step])
    peak, _ = measured_peak()
    b_res = algorithmic_bytes(R_local, nnz, n_rows, 3, False, False, "residual")
    b_jac = algorithmic_bytes(R_local, nnz, n_rows, 3, False, False, "assembly")
    out = {
        "workload": f"Bratu (BASELINE config 5 family), cube_tetra {N}^3 points, {N**3} vertices, "
        f"{cells} cells, {R_global} edge contributions; step = halo push + f.eval + Jacobian refresh",
        "partition": f"{D.world} z-slab(s), generated and measured in HBM; u halo = one grid layer per "
        "side pushed over NVLink peer memory (CUDA IPC), waited for inside the tile kernel",
        "ms_per_step": ms_step,
        "contributions_per_s": 2 * R_global / (ms_step * 1e-3),
        "residual": {
            "ms": ms_res,
            "contributions_per_s": R_global / (ms_res * 1e-3),
            "achieved_gbs_per_gpu": b_res / (ms_res * 1e-3) / 1e9,
            "frac": b_res / (ms_res * 1e-3) / 1e9 / peak,
        },
        "jacobian": {
            "ms": ms_jac,
            "contributions_per_s": R_global / (ms_jac * 1e-3),
            "achieved_gbs_per_gpu": b_jac / (ms_jac * 1e-3) / 1e9,
            "frac": b_jac / (ms_jac * 1e-3) / 1e9 / peak,
        },
        "frac_step": (b_res + b_jac) / (ms_step * 1e-3) / 1e9 / peak,
        "tile_quantum": dm.stats()["quantum"],
        "gpu_launches": launches,
        "pattern_build_ms": dm.build_ms,
        "setup_s": setup_s,
    }
    if args.solve:
        # the whole nonlinear solve on the partition: pyfvm.newton's loop with the device
        # BiCGStab (Jacobi) as jacobian_solver; nothing per-record or per-slot leaves HBM
        from pyfvm_b200 import solvers

        solvers.DeviceSystem(dm, s.data.ptr, _partitioned=True)  # NVRTC-compiles the SpMV kernel
        out["newton_solve"] = {}
        for method in args.linear_solver.split(","):
            D.barrier()
            ts = time.time()
            u_own, steps, lin_its = solvers.newton_distributed(
                dp, numpy.zeros(part.n_owned), tol=args.newton_tol, lin_tol=1e-8, verbose=(D.rank == 0),
                linear_solver=method,
            )
            D.barrier()
            umax = D.max([float(u_own.max()) if len(u_own) else 0.0])[0]
            out["newton_solve"][method] = {
                "seconds": time.time() - ts,
                "newton_steps": steps,
                "linear_iterations": lin_its,
                "spmv_per_iteration": 2 if method == "bicgstab" else (int(method[7:]) if method.startswith("cg-cheb") else 1),
                "tol": args.newton_tol,
                "max_u": umax,
                "what": f"pyfvm.newton from u = 0 with a Jacobi-{method} jacobian_solver, residual / "
                "Jacobian refresh / SpMV / updates in HBM across the ranks",
            }
    D.barrier()
    dp.halo.close()
    return out


def leg_small_configs():
    """BASELINE configs 1-3 through the drop-in API with host arrays: wall clock of the
    first call in this process (sympy lowering, NVRTC or cubin-cache load, upload, pattern
    build, launch, download) and of a repeated call, next to the CPU oracle's numeric
    path.  These meshes are launch-latency-sized: no roofline claim is made on them."""
    import oracle
    import pyfvm_b200 as pe
    from pyfvm_b200 import meshes
    from scipy.sparse import linalg

    out = []

    def wall(fn):
        t = time.time()
        r = fn()
        return time.time() - t, r

    # config 1: README Poisson, 401 x 201
    m1 = meshes.Mesh(*meshes.rectangle_tri(numpy.linspace(0, 2, 401), numpy.linspace(0, 1, 201)))
    m1.ce_ratios, m1.control_volumes, m1.is_boundary_point
    first, _ = wall(lambda: pe.discretize_linear(problem_poisson(pe.form_language), m1))
    warm, _ = wall(lambda: pe.discretize_linear(problem_poisson(pe.form_language), m1))
    k1 = oracle.lower_linear(problem_poisson(oracle.form_language), m1)
    cpu, _ = wall(lambda: oracle.get_linear_fvm_problem(m1, *k1))
    out.append({"config": "1: Poisson rectangle_tri 401x201, discretize_linear", "contributions": 480000,
                "first_call_s": first, "warm_call_s": warm, "cpu_oracle_numeric_s": cpu})

    # config 2: README Bratu, 101 x 51, the whole pyfvm.newton with a host spsolve
    m2 = meshes.Mesh(*meshes.rectangle_tri(numpy.linspace(0, 2, 101), numpy.linspace(0, 1, 51)))
    m2.ce_ratios, m2.control_volumes, m2.is_boundary_point

    def newton_with(mod, fl):
        f, jac = mod.discretize(problem_bratu(fl), m2)

        def solver(u0, rhs):
            return linalg.spsolve(jac.get_linear_operator(u0), rhs)

        return mod.newton(f.eval, solver, numpy.zeros(len(m2.points)), verbose=False)

    first, u_gpu = wall(lambda: newton_with(pe, pe.form_language))
    warm, _ = wall(lambda: newton_with(pe, pe.form_language))
    cpu, u_cpu = wall(lambda: newton_with(oracle, oracle.form_language))
    out.append({"config": "2: Bratu rectangle_tri 101x51, discretize + pyfvm.newton (host spsolve)",
                "contributions": 30000, "first_call_s": first, "warm_call_s": warm,
                "cpu_oracle_s": cpu, "max_abs_diff_u": float(numpy.abs(u_gpu - u_cpu).max())})

    # config 3: Poisson on cube_tetra 60^3
    xs = numpy.linspace(0.0, 1.0, 60)
    m3 = meshes.Mesh(*meshes.cube_tetra(xs, xs, xs))
    m3.ce_ratios, m3.control_volumes, m3.is_boundary_point
    first, _ = wall(lambda: pe.discretize_linear(problem_poisson(pe.form_language), m3))
    warm, _ = wall(lambda: pe.discretize_linear(problem_poisson(pe.form_language), m3))
    k3 = oracle.lower_linear(problem_poisson(oracle.form_language), m3)
    cpu, _ = wall(lambda: oracle.get_linear_fvm_problem(m3, *k3))
    out.append({"config": "3: Poisson cube_tetra 60^3 (1 026 895 cells), discretize_linear",
                "contributions": 12 * 1026895, "first_call_s": first, "warm_call_s": warm,
                "cpu_oracle_numeric_s": cpu})
    return out


def run_ours(args):
    D = Dist()
    rank, world, local = D.rank, D.world, D.local
    from pyfvm_b200 import codegen
    from pyfvm_b200 import engine as pengine
    import pyfvm_b200 as pe

    t_setup = time.time()
    part, lp = build_local_problem(args.size, rank, world)
    mesh = part.mesh
    eng = lp.eng
    dm = lp.dmesh
    n_rows, nnz = dm.n_rows, dm.nnz
    R_global = 3 * 2 * (args.size - 1) ** 2
    # contributions this rank processes: half-edges / 2 (cut records count once per side)
    R_local = dm.n_halfedges / 2.0
    setup_s = time.time() - t_setup
    quantum = dm.stats()["quantum"]
    peak, peak_src = measured_peak()

    fun = lp.functors[codegen.MODE_LINEAR]
    alg = algorithmic_bytes(R_local, nnz, n_rows, dm.dim, True, fun.uses_el, "assembly")
    # ---- device-resident timing on the ctx stream (CUDA events).  Working set well
    # above L2: K launches back to back.  Otherwise (many GPUs on a fixed mesh): L2 is
    # flushed before every launch (a 256 MB memset) and each launch is timed alone.
    flush_l2 = alg < 4 * L2_BYTES
    flush_buf = eng.alloc(256 << 20) if flush_l2 else None
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    for _ in range(args.warmup):
        lp.assemble_device()
    launches0 = eng.launches
    t0 = time.time()
    ms_step = timed_launches(eng, D, lp.assemble_device, args.steps, 0, flush_buf)
    launches = eng.launches - launches0  # tile-kernel launches inside the timed region
    # ---- the same launch for >= 1 s: what the 1000 W power cap sustains (the K-step figure
    # above is a burst of a few milliseconds); the clock samples cover both regions
    sus_steps = max(args.steps, int(1.2 / max(ms_step * 1e-3, 1e-5))) if not flush_l2 else 4 * args.steps
    ms_sus = timed_launches(eng, D, lp.assemble_device, sus_steps, 0, flush_buf)
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms_step, ms_sus = D.max([ms_step, ms_sus])

    # ---- END TO END = the API call a pyfvm user makes, host arrays in and out, on a mesh
    # object the engine has not seen: sympy lowering (simplify memoised), upload of the
    # int64 index hierarchy + geometry from pageable numpy arrays, on-device pattern
    # build, one assembly launch, indptr / indices / values / rhs back into numpy.
    from pyfvm_b200 import dist as fd

    def api_call():
        if world == 1:
            return pe.discretize_linear(problem_c4(pe.form_language), mesh)
        return fd.discretize_linear(problem_c4(pe.form_language), part)

    e2e_n = max(2, min(args.steps, 4))
    cold_s, first_s = [], None
    A_api = b_api = None
    for i in range(3 + e2e_n):
        # three untimed calls (W >= 3): the very first one of the process also loads the cubin
        # (or compiles), pays one-off page faults, and -- from the second request of a size
        # on -- the engine page-locks its result buffers; its wall time is reported separately
        A_api = b_api = None  # a caller's loop overwrites its previous result
        pengine._mesh_cache.clear()
        gc.collect()
        eng.sync()
        D.barrier()
        tc = time.time()
        A_api, b_api = api_call()
        dt = D.max([time.time() - tc])[0]
        if i == 0:
            first_s = dt
        if i >= 3:
            cold_s.append(dt)
    e2e_s = sum(cold_s) / len(cold_s)
    # the same call again on the SAME mesh object: the device mesh is found in the cache and
    # revalidated against the mesh arrays (identity + sampled content)
    warm_s = []
    for i in range(1 + e2e_n):
        # (one untimed call: a mesh's second launch bank-steers its vertex tables, once)
        A_api = b_api = None
        D.barrier()
        tc = time.time()
        A_api, b_api = api_call()
        if i >= 1:
            warm_s.append(D.max([time.time() - tc])[0])
    e2e_warm_s = sum(warm_s) / len(warm_s)
    idx_bytes = 2 * 8 * dm.R
    h2d_cold = idx_bytes + 8 * dm.R + 8 * dm.dim * part.n_local + 8 * part.n_local
    d2h_cold = 4 * (n_rows + 1) + 4 * nnz + 8 * nnz + 8 * n_rows
    del A_api, b_api

    # ---- round 1's hand-pipelined loop, kept for comparison: per step ce_ratios host ->
    # device (pinned) + permute + assembly + values / rhs device -> host (pinned, copy stream)
    pip = None
    if not args.no_pipelined:
        ce_host = eng.pinned_empty(dm.R, numpy.float64)
        ce_host[:] = numpy.asarray(mesh.ce_ratios).reshape(-1)
        data_host = eng.pinned_empty(nnz, numpy.float64)
        rhs_host = eng.pinned_empty(n_rows, numpy.float64)
        pn = 4
        for i in range(2 + pn):
            if i == 2:
                eng.copy_sync()
                eng.sync()
                D.barrier()
                te = time.time()
            lp.update_geometry(ce=ce_host)
            eng.copy_fence()
            lp.assemble_device()
            lp.data.download_async(data_host)
            lp.vec.download_async(rhs_host)
        eng.copy_sync()
        pip_s = D.max([(time.time() - te) / pn])[0]
        pip = {
            "value": R_global / pip_s,
            "unit": UNIT,
            "h2d_bytes_per_step": int(8 * dm.R),
            "d2h_bytes_per_step": int(8 * nnz + 8 * n_rows),
            "what": "NOT the pyfvm API: ce_ratios re-uploaded from pinned memory onto a cached "
            "pattern, assembly, values + rhs to pinned memory on the copy stream (overlaps the next upload)",
        }
        del ce_host, data_host, rhs_host

    # ---- residual evaluation on the same mesh (config 4: "linear assembly + residual
    # eval").  N > 1: rows partitioned the same way, ghost values of u pushed over
    # NVLink peer memory before every evaluation (included in the time).
    res = jac_leg = None
    if not args.no_residual:
        dp = fd.DistributedFvm(problem_c4(pe.form_language), part, halo="peer") if world > 1 else None
        if world == 1:
            f, jac = pe.discretize(problem_c4(pe.form_language), mesh)
            s = f._s
            s.upload_u(numpy.random.default_rng(0).standard_normal(part.n_local))
            step = lambda: f.eval_device(s.u.ptr)
            jstep = lambda: jac.values_device(s.u.ptr)
        else:
            rng = numpy.random.default_rng(rank)
            s = dp.s

            def step():
                # the owned part of the next u is already on the device (Newton's
                # update would have written it); push / wait / evaluate: one host call
                dp.step_device(codegen.MODE_RESIDUAL)

            for k in (0, 1):
                dp.set_owned(rng.standard_normal(part.n_owned))
                step()
            u_fixed = dp.exchange(codegen.MODE_JACOBIAN)
            jstep = lambda: dp.jacobian_device(u_fixed)
        rms = D.max([timed_launches(eng, D, step, args.steps, args.warmup, flush_buf)])[0]
        jms = D.max([timed_launches(eng, D, jstep, args.steps, args.warmup, flush_buf)])[0]
        if dp is not None:
            dp.halo.check_timeout()
        fr = s.functors[codegen.MODE_RESIDUAL]
        fj = s.functors[codegen.MODE_JACOBIAN]
        rb = algorithmic_bytes(R_local, nnz, n_rows, dm.dim, fr.uses_x, fr.uses_el, "residual")
        jb = algorithmic_bytes(R_local, nnz, n_rows, dm.dim, fj.uses_x, fj.uses_el, "assembly")
        res = {
            "ms_per_eval": rms,
            "algorithmic_bytes": int(rb),
            "contributions_per_s": R_global / (rms * 1e-3),
            "achieved_gbs_per_gpu": rb / (rms * 1e-3) / 1e9,
            "frac": rb / (rms * 1e-3) / 1e9 / peak,
            "frac_of_8TBs_nominal": rb / (rms * 1e-3) / 8e12,
        }
        jac_leg = {
            "ms_per_refresh": jms,
            "algorithmic_bytes": int(jb),
            "contributions_per_s": R_global / (jms * 1e-3),
            "achieved_gbs_per_gpu": jb / (jms * 1e-3) / 1e9,
            "frac": jb / (jms * 1e-3) / 1e9 / peak,
        }
        if world > 1:
            res["halo"] = "CUDA-IPC peer push over NVLink + epoch flags, inside the timed region"
        if world == 1:
            # through the plugin objects with host arrays, as scipy's newton_krylov
            # [README:136] and the README's jacobian_solver [README:116] call them
            u_host = numpy.random.default_rng(0).standard_normal(part.n_local)
            # three untimed calls: the engine page-locks a result buffer on the second request
            # of a size (engine.host_result), which must not land in the timed ones
            for i in range(3 + e2e_n):
                if i == 3:
                    tf = time.time()
                f.eval(u_host)
            res["e2e_host_arrays"] = {
                "value": R_global / ((time.time() - tf) / e2e_n),
                "unit": UNIT,
                "h2d_bytes_per_step": 8 * part.n_local,
                "d2h_bytes_per_step": 8 * n_rows,
                "what": "f.eval(u) with pageable numpy arrays in and out",
            }
            J = None
            for i in range(3 + e2e_n):
                if i == 3:
                    tf = time.time()
                J = None  # a Newton loop drops the previous operator before asking for the next
                J = jac.get_linear_operator(u_host)
            jac_leg["e2e_host_arrays"] = {
                "value": R_global / ((time.time() - tf) / e2e_n),
                "unit": UNIT,
                "h2d_bytes_per_step": 8 * part.n_local,
                "d2h_bytes_per_step": 8 * nnz,
                "what": "jacobian.get_linear_operator(u): numpy u in, scipy CSR out (values downloaded, pattern cached)",
            }
            del J
        if dp is not None:
            D.barrier()
            dp.halo.close()

    # ---- y = A x on the assembled system (the Krylov building block after the path)
    spmv = None
    if world == 1:
        from pyfvm_b200 import solvers

        lp.assemble_device()
        sysm = solvers.DeviceSystem(dm, lp.data.ptr)
        xv = eng.to_device(numpy.random.default_rng(1).standard_normal(part.n_local), pad=16)
        yv = eng.alloc(8 * n_rows + 16)
        sms = timed_launches(eng, D, lambda: sysm.matvec(xv.ptr, yv.ptr), args.steps, args.warmup)
        sb = 12 * nnz + 20 * n_rows  # values + column ids, x, y, row pointers
        spmv = {
            "ms": sms,
            "algorithmic_bytes": int(sb),
            "achieved_gbs": sb / (sms * 1e-3) / 1e9,
            "frac": sb / (sms * 1e-3) / 1e9 / peak,
            "gflops": 2 * nnz / (sms * 1e-3) / 1e9,
        }
        del sysm, xv, yv

    out = {
        "metric": METRIC,
        "value": R_global / (ms_step * 1e-3),
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.size, world, flush_l2),
        "tile_quantum": quantum,
        "gpu_launches": launches,
        "clocks": clocks,
        "sustained": {
            "launches": sus_steps,
            "ms_per_step": ms_sus,
            "value": R_global / (ms_sus * 1e-3),
            "frac": alg / (ms_sus * 1e-3) / 1e9 / peak,
            "what": "the same launch repeated for >= 1 s (power-capped clocks); `value` is the K-step burst",
        },
        "e2e": {
            "value": R_global / e2e_s,
            "unit": UNIT,
            "seconds_per_step": e2e_s,
            "seconds_each_call": [round(t, 4) for t in cold_s],
            "first_call_of_the_process_s": first_s,
            "h2d_bytes_per_step": int(h2d_cold),
            "d2h_bytes_per_step": int(d2h_cold),
            "what": "pyfvm.discretize_linear(problem, mesh) -- the API call -- on a mesh object the engine "
            "has not seen, pageable numpy in, scipy CSR + numpy rhs out: lowering, upload of the int64 index "
            "hierarchy + geometry, on-device pattern build, assembly, pattern + values + rhs download"
            "; three untimed calls first (first_call_of_the_process_s: the very first one, incl. cubin "
            "load and one-off page faults)"
            + ("" if world == 1 else " (every rank calls dist.discretize_linear on its slab; max over ranks)"),
        },
        "e2e_warm": {
            "value": R_global / e2e_warm_s,
            "unit": UNIT,
            "seconds_per_step": e2e_warm_s,
            "seconds_each_call": [round(t, 4) for t in warm_s],
            "h2d_bytes_per_step": 0,
            "d2h_bytes_per_step": int(8 * nnz + 8 * n_rows),
            "what": "the same API call again on the same mesh object: device mesh cached and revalidated, "
            "values + rhs downloaded into the engine's page-locked result buffers (numpy views)",
        },
        "e2e_pipelined": pip,
        "roofline": {
            "bound": "hbm",
            "achieved": alg / (ms_step * 1e-3) / 1e9,
            "peak": peak,
            "unit": "GB/s",
            "frac": alg / (ms_step * 1e-3) / 1e9 / peak,
            "traffic": measured_traffic(args.size, quantum) if world == 1 else None,
            "kernel": "fvm_tile_kernel (mode linear)",
            "algorithmic_bytes_per_launch": int(alg),
            "peak_source": peak_src,
            "frac_of_8TBs_nominal": alg / (ms_step * 1e-3) / 8e12,
        },
        "residual": res,
        "jacobian": jac_leg,
        "spmv": spmv,
        "pattern_build_ms": dm.build_ms,
        "setup_s": setup_s,
    }
    # free the config-4 objects before the 3-D leg builds its mesh
    del lp, dm, mesh, part
    if not args.no_residual:
        del s
        if world == 1:
            del f, jac
    pengine._mesh_cache.clear()
    gc.collect()
    if args.c5_cube > 0:
        out["c5"] = leg_c5(args, D, args.c5_cube)
    if world > 1 and not args.no_parity:
        out["dist_parity"] = fd.parity_selfcheck(verbose=(rank == 0 and args.verbose))
    if rank == 0 and world == 1 and not args.no_small:
        out["small_configs"] = leg_small_configs()
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(args.cpu_size)
    if rank == 0:
        print(json.dumps(out), flush=True)
    D.close()


def run_c5(args):
    """``--workload c5``: the config-5 family alone (see ``leg_c5``), one JSON line."""
    D = Dist()
    sampler = ClockSampler(D.local)
    sampler.start()
    sampler.wait_first()
    t0 = time.time()
    leg = leg_c5(args, D, args.cube)
    clocks = sampler.stop(t0, time.time())
    peak, peak_src = measured_peak()
    out = {
        "metric": "edge contributions processed/s (residual f.eval + Jacobian value refresh)",
        "value": leg["contributions_per_s"],
        "unit": UNIT,
        "n_gpus": D.world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": leg["ms_per_step"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": leg["workload"], "partition": leg["partition"],
                   "l2": "inputs larger than L2 (no flush needed)"},
        "gpu_launches": leg["gpu_launches"],
        "clocks": clocks,
        "roofline": {"bound": "hbm", "frac": leg["frac_step"], "peak": peak, "unit": "GB/s",
                     "peak_source": peak_src, "kernel": "fvm_tile_kernel (modes residual + jacobian), per GPU",
                     "traffic": None},
    }
    out.update({k: leg[k] for k in ("residual", "jacobian", "tile_quantum", "pattern_build_ms", "setup_s")})
    if "newton_solve" in leg:
        out["newton_solve"] = leg["newton_solve"]
    if D.rank == 0:
        print(json.dumps(out), flush=True)
    D.close()


def cpu_assembly_seconds(size, repeat=1):
    """Numeric half of the oracle's ``discretize_linear`` (COO build -> tocsr ->
    Dirichlet) on a size x size sample of the workload; lowering excluded."""
    import oracle
    from pyfvm_b200 import meshes

    xs = numpy.linspace(0.0, 1.0, size)
    v, c = meshes.rectangle_tri(xs, xs)
    mesh = meshes.Mesh(v, c)
    mesh.ce_ratios, mesh.control_volumes, mesh.is_boundary_point  # geometry is an input
    kernels = oracle.lower_linear(problem_c4(oracle.form_language), mesh)
    times = []
    for _ in range(repeat):
        t = time.time()
        oracle.get_linear_fvm_problem(mesh, *kernels)
        times.append(time.time() - t)
    return 3 * 2 * (size - 1) ** 2, times


def cpu_baseline(size):
    R, times = cpu_assembly_seconds(size, repeat=1)
    return {
        "value": R / times[0],
        "unit": UNIT,
        "cores": 1,
        "host_cores": len(os.sched_getaffinity(0)),
        "kind": "port",
        "sample": f"oracle (restated reference: numpy ufunc.at + scipy coo->csr) on a {size}x{size} "
        f"rectangle_tri sample of the workload ({R} contributions, {times[0]:.1f} s, 1 thread: the "
        "reference path is single-threaded by construction); symbolic lowering excluded",
    }


def run_reference(args):
    """The reference arm: the CPU oracle ("port": the reference is a README and cannot be
    installed) on this arm's own workload, metric and unit.  Each of the W + K steps is a
    bounded sample (a smaller rectangle_tri grid of the same operator, sized so the whole
    run ends in about two minutes); one extra step at the FULL size is timed when the host
    has the memory for it (the 4001^2 COO intermediate is ~25 GB)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 100.0 * 2.0e6 / (args.warmup + args.steps)
    size = int(min(args.cpu_size, max(101, (per_step / 6.0) ** 0.5 + 1)))
    R, times = cpu_assembly_seconds(size, repeat=args.warmup + args.steps)
    t = times[args.warmup:]
    sec = sum(t) / len(t)
    val = R / sec
    full = None
    if not args.no_full_size:
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except ImportError:
            avail = 0
        need = 420.0 * 6 * (args.size - 1) ** 2  # ~420 B per contribution at the peak (measured)
        if avail > 1.3 * need:
            Rf, tf = cpu_assembly_seconds(args.size, repeat=1)
            full = {"size": args.size, "contributions": Rf, "seconds": tf[0], "value": Rf / tf[0],
                    "what": "ONE oracle assembly at the full size of the workload (same config as the GPU arm)"}
        else:
            full = {"skipped": f"host has {avail / 1e9:.0f} GB available, the full-size COO path needs ~{need / 1e9:.0f} GB"}
    base = {
        "value": val,
        "unit": UNIT,
        "cores": 1,
        "host_cores": len(os.sched_getaffinity(0)),
        "kind": "port",
        "sample": f"each step = oracle assembly of a {size}x{size} rectangle_tri sample of the workload "
        f"({R} contributions, same operator); per-contribution rate is size-independent "
        "(see full_size_step)",
    }
    print(
        json.dumps(
            {
                "impl": "reference",
                "metric": METRIC,
                "value": val,
                "unit": UNIT,
                "n_gpus": args.gpus,
                "steps": args.steps,
                "warmup": args.warmup,
                "ms_per_step": sec * 1e3,
                "higher_is_better": True,
                "scaling": "strong",
                "vs_baseline": None,
                "dtype": "f64",
                "data": "synthetic",
                "config": workload_config(args.size, args.gpus, False),
                "cpu_baseline": base,
                "full_size_step": full,
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
        ),
        flush=True,
    )


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4001, help="grid points per side")
    ap.add_argument("--cpu-size", type=int, default=2001, help="side of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-small", action="store_true", help="skip the configs 1-3 API lines")
    ap.add_argument("--no-pipelined", action="store_true", help="skip round 1's pipelined loop")
    ap.add_argument("--no-residual", action="store_true", help="skip the residual / Jacobian legs")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the bit-identity self-check")
    ap.add_argument("--no-full-size", action="store_true", help="reference arm: skip the full-size step")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: BASELINE config 4 (default, the headline); c5: Bratu on cube_tetra slabs")
    ap.add_argument("--c5-cube", type=int, default=256,
                    help="c4 run: points per side of the config-5-family leg (0: skip)")
    ap.add_argument("--cube", type=int, default=200, help="points per side of the c5 cube")
    ap.add_argument("--solve", action="store_true", help="c5: also run the full Newton solve")
    ap.add_argument("--linear-solver", default="bicgstab,cg,cg-cheb3",
                    help="c5 --solve: comma list of the device Krylov solvers to time "
                    "(bicgstab, cg, cg-cheb<degree>)")
    ap.add_argument("--newton-tol", type=float, default=1e-10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
