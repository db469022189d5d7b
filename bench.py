#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json metric).

A "step" is one matrix+rhs assembly pass (one tile-kernel launch) over the
convection-diffusion problem of BASELINE config 4 on a synthetic
``rectangle_tri`` mesh (default 4001 x 4001 points = 16 008 001 vertices,
96 000 000 edge contributions).  One edge contribution = one column of
``idx_hierarchy``.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (under torchrun): the vertex rows are partitioned into N contiguous
slabs, one rank per GPU; linear assembly needs no collective, so the ranks run
their slabs independently (strong scaling: total work fixed) and the time is
the max over ranks.

``--impl reference`` times the CPU oracle (the restated reference algorithm with
the reference's own numpy/scipy calls; the reference itself is a README and
cannot be installed) on a bounded sample of the same workload.
"""
import argparse
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

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        for _ in range(40):  # nvidia-smi needs a moment to print its first sample
            if self.lines:
                break
            time.sleep(0.05)
        time.sleep(0.06)
        self.proc.terminate()
        self.thread.join(timeout=2)
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.06] or [l for _, l in self.lines[-1:]]
        sm, mx, reasons = [], [], set()
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(self.NAMES, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(numpy.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def measured_traffic(size, quantum):
    """DRAM bytes per assembly launch from the committed ncu capture
    (profiles/r01_traffic.json), if it was taken on this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        if int(t["size"]) == size and int(t["tile_quantum"]) == quantum:
            return int(t["assembly_dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        pass
    return None


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    from pyfvm_b200 import codegen

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

    sampler = ClockSampler(local)
    fun = lp.functors[codegen.MODE_LINEAR]
    alg = algorithmic_bytes(R_local, nnz, n_rows, dm.dim, True, fun.uses_el, "assembly")
    # ---- device-resident timing on the ctx stream (CUDA events).  Working set well
    # above L2: K launches back to back.  Otherwise (many GPUs on a fixed mesh): L2 is
    # flushed before every launch (a 256 MB memset) and each launch is timed alone.
    flush_l2 = alg < 4 * L2_BYTES
    flush_buf = eng.alloc(256 << 20) if flush_l2 else None
    sampler.start()
    for _ in range(args.warmup):
        lp.assemble_device()
    eng.sync()
    barrier()
    launches0 = eng.launches
    t0 = time.time()
    if flush_l2:
        ms = 0.0
        for _ in range(args.steps):
            eng._ck(eng.lib.fvm_dev_memset(eng.ctx, flush_buf.ptr, 0, flush_buf.nbytes))
            eng.timer_begin()
            lp.assemble_device()
            ms += eng.timer_end()
    else:
        eng.timer_begin()
        for _ in range(args.steps):
            lp.assemble_device()
        ms = eng.timer_end()
    t1 = time.time()
    launches = eng.launches - launches0
    barrier()
    clocks = sampler.stop(t0, t1)

    # ---- end to end through the plugin objects with HOST buffers.  Per step the
    # per-record geometry goes host -> device (pinned), is permuted into half-edge
    # order and assembled, and the CSR values + rhs come back to pinned host
    # arrays.  The device -> host copy runs on the copy stream, so it overlaps the
    # next step's upload (PCIe is full duplex); the fence keeps a buffer from being
    # rewritten while it is still being read.
    ce_host = eng.pinned_empty(dm.R, numpy.float64)
    ce_host[:] = numpy.asarray(mesh.ce_ratios).reshape(-1)
    data_host = eng.pinned_empty(nnz, numpy.float64)
    rhs_host = eng.pinned_empty(n_rows, numpy.float64)
    e2e_steps = max(3, min(args.steps, 10))
    for i in range(2 + e2e_steps):
        if i == 2:
            eng.copy_sync()
            eng.sync()
            barrier()
            te = time.time()
        lp.update_geometry(ce=ce_host)
        eng.copy_fence()
        lp.assemble_device()
        lp.data.download_async(data_host)
        lp.vec.download_async(rhs_host)
    eng.copy_sync()
    e2e_s = (time.time() - te) / e2e_steps
    barrier()

    peak, peak_src = measured_peak()
    # ---- cold call: pyfvm.discretize_linear(problem, mesh) on a mesh the engine has
    # not seen (upload of the whole mesh, on-device pattern build, assembly, pattern +
    # values back to the host); the functor set is already compiled
    cold = None
    if world == 1 and not args.no_cold:
        import pyfvm_b200 as pe
        from pyfvm_b200 import engine as pengine

        pengine._mesh_cache.clear()
        eng.sync()
        tc = time.time()
        A_cold, b_cold = pe.discretize_linear(problem_c4(pe.form_language), mesh)
        cold_s = time.time() - tc
        cold = {
            "seconds": cold_s,
            "value": R_global / cold_s,
            "unit": UNIT,
            "what": "pyfvm.discretize_linear(problem, mesh) on an uncached mesh: sympy lowering (simplify memoised per expression), "
            "int64 index + geometry upload from pageable numpy arrays, on-device pattern build, "
            "assembly, indptr/indices/values/rhs back to the host",
        }
        del A_cold, b_cold
        pengine._mesh_cache.clear()

    # ---- residual evaluation on the same mesh (config 4: "linear assembly + residual
    # eval").  N > 1: rows partitioned the same way, ghost values of u pushed over
    # NVLink peer memory before every evaluation (included in the time).
    from pyfvm_b200 import dist as fd
    import pyfvm_b200 as pe

    res = None
    u_glob = None
    if not args.no_residual:
        dp = fd.DistributedFvm(problem_c4(pe.form_language), part, halo="peer") if world > 1 else None
        if world == 1:
            f, jac = pe.discretize(problem_c4(pe.form_language), mesh)
            s = f._s
            s.upload_u(numpy.random.default_rng(0).standard_normal(part.n_local))
            step = lambda: f.eval_device(s.u.ptr)
        else:
            rng = numpy.random.default_rng(rank)
            s = dp.s

            def step():
                # the owned part of the next u is already on the device (Newton's
                # update would have written it); push / wait / evaluate
                dp.residual_device(dp.exchange())

            for k in (0, 1):
                dp.set_owned(rng.standard_normal(part.n_owned))
                step()
        for _ in range(args.warmup):
            step()
        eng.sync()
        barrier()
        if flush_l2:
            rms = 0.0
            for _ in range(args.steps):
                eng._ck(eng.lib.fvm_dev_memset(eng.ctx, flush_buf.ptr, 0, flush_buf.nbytes))
                eng.timer_begin()
                step()
                rms += eng.timer_end()
            rms /= args.steps
        else:
            eng.timer_begin()
            for _ in range(args.steps):
                step()
            rms = eng.timer_end() / args.steps
        barrier()
        if dp is not None:
            dp.halo.check_timeout()
        fun = s.functors[codegen.MODE_RESIDUAL]
        rb = algorithmic_bytes(R_local, nnz, n_rows, dm.dim, True, fun.uses_el, "residual")
        res = {"ms_per_eval": rms, "algorithmic_bytes": int(rb)}
        if world == 1:
            # f.eval(u) through the plugin object with host arrays, as scipy's
            # newton_krylov [README:136] calls it: u up, F down, every evaluation
            u_host = numpy.random.default_rng(0).standard_normal(part.n_local)
            for i in range(2 + e2e_steps):
                if i == 2:
                    tf = time.time()
                f.eval(u_host)
            res["e2e_host_arrays"] = {
                "value": R_global / ((time.time() - tf) / e2e_steps),
                "unit": UNIT,
                "h2d_bytes_per_step": 8 * part.n_local,
                "d2h_bytes_per_step": 8 * n_rows,
                "what": "f.eval(u) with pageable numpy arrays in and out",
            }

    # ---- y = A x on the assembled system (the Krylov building block after the path)
    spmv = None
    if world == 1:
        from pyfvm_b200 import solvers

        lp.assemble_device()
        sysm = solvers.DeviceSystem(dm, lp.data.ptr)
        xv = eng.to_device(numpy.random.default_rng(1).standard_normal(part.n_local))
        yv = eng.alloc(8 * n_rows)
        for _ in range(args.warmup):
            sysm.matvec(xv.ptr, yv.ptr)
        eng.timer_begin()
        for _ in range(args.steps):
            sysm.matvec(xv.ptr, yv.ptr)
        sms = eng.timer_end() / args.steps
        sb = 12 * nnz + 20 * n_rows  # values + column ids, x, y, row pointers
        spmv = {
            "ms": sms,
            "algorithmic_bytes": int(sb),
            "achieved_gbs": sb / (sms * 1e-3) / 1e9,
            "frac": sb / (sms * 1e-3) / 1e9 / peak,
            "gflops": 2 * nnz / (sms * 1e-3) / 1e9,
        }

    # ---- reduce over ranks: max time, summed work
    ms_step = ms / args.steps
    if dist is not None:
        import torch

        rms_local = res["ms_per_eval"] if res else 0.0
        t = torch.tensor([ms_step, e2e_s, rms_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_s, rms_max = t.tolist()
        if res:
            res["ms_per_eval"] = rms_max
            res["halo"] = "CUDA-IPC peer push over NVLink + epoch flags, inside the timed region"
    if res:
        rms = res["ms_per_eval"]
        res["contributions_per_s"] = R_global / (rms * 1e-3)
        res["achieved_gbs_per_gpu"] = res["algorithmic_bytes"] / (rms * 1e-3) / 1e9
        res["frac"] = res["achieved_gbs_per_gpu"] / peak

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
        "config": {
            "workload": f"convection-diffusion (BASELINE config 4), rectangle_tri {args.size}x{args.size}, "
            f"{args.size**2} vertices, {R_global} edge contributions, linear assembly (CSR values + rhs)",
            "partition": "1 GPU" if world == 1 else f"{world} contiguous row slabs, no data-path collective",
            "l2": "L2 flushed (256 MB memset) before every timed launch, launches timed one by one"
            if flush_l2
            else "inputs larger than L2 (no flush needed)",
            "tile_quantum": quantum,
        },
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {
            "value": R_global / e2e_s,
            "unit": UNIT,
            "h2d_bytes_per_step": int(8 * dm.R),
            "d2h_bytes_per_step": int(8 * nnz + 8 * n_rows),
            "what": "per step: ce_ratios host->device (pinned) + permute to half-edge order + "
            "assembly + CSR values and rhs device->host (pinned, copy stream: overlaps the next "
            "step's upload); pattern cached per mesh",
        },
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
        "spmv": spmv,
        "e2e_cold": cold,
        "pattern_build_ms": dm.build_ms,
        "setup_s": setup_s,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(args.cpu_size)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


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


def run_c5(args):
    """BASELINE config 5: Bratu on a cube_tetra grid, vertex rows partitioned into
    z-slabs across the GPUs, every slab generated and its geometry computed in HBM.
    A step = one Newton step's device work: halo push of u + residual f.eval +
    Jacobian value refresh on the fixed pattern."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    import pyfvm_b200 as pe
    from pyfvm_b200 import codegen
    from pyfvm_b200 import dist as fd

    N = args.cube
    xs = numpy.linspace(0.0, 1.0, N)
    t_setup = time.time()
    part = fd.SlabPartition([xs, xs, xs], rank, world, device=True)
    dp = fd.DistributedFvm(problem_bratu(pe.form_language), part, halo="peer")
    eng, s, dm = dp.eng, dp.s, dp.s.dmesh
    setup_s = time.time() - t_setup
    cells = 5 * (N - 1) ** 3
    R_global = 12 * cells
    R_local = dm.n_halfedges / 2.0
    n_rows, nnz = dm.n_rows, dm.nnz
    rng = numpy.random.default_rng(rank)
    for _ in (0, 1):  # both u buffers get owned values (Newton's update writes them in place)
        dp.set_owned(0.1 * rng.standard_normal(part.n_owned))
        dp.residual_device(dp.exchange())

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        eng.sync()
        barrier()
        eng.timer_begin()
        for _ in range(args.steps):
            fn()
        ms = eng.timer_end() / args.steps
        barrier()
        return ms

    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.time()
    launches0 = eng.launches
    ms_res = timed(lambda: dp.residual_device(dp.exchange()))
    u_ptr = dp.exchange()
    ms_jac = timed(lambda: dp.jacobian_device(u_ptr))

    def newton_device_work():
        p = dp.exchange()
        dp.residual_device(p)
        dp.jacobian_device(p)

    ms_step = timed(newton_device_work)
    launches = eng.launches - launches0
    clocks = sampler.stop(t0, time.time())
    dp.halo.check_timeout()
    if dist is not None:
        import torch

        t = torch.tensor([ms_res, ms_jac, ms_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_jac, ms_step = t.tolist()
    peak, peak_src = measured_peak()
    b_res = algorithmic_bytes(R_local, nnz, n_rows, 3, False, False, "residual")
    b_jac = algorithmic_bytes(R_local, nnz, n_rows, 3, False, False, "assembly")
    out = {
        "metric": "edge contributions processed/s (residual f.eval + Jacobian value refresh)",
        "value": 2 * R_global / (ms_step * 1e-3),
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
        "config": {
            "workload": f"Bratu (BASELINE config 5 family), cube_tetra {N}^3 points, {N**3} vertices, "
            f"{cells} cells, {R_global} edge contributions; step = halo push + f.eval + Jacobian refresh",
            "partition": f"{world} z-slab(s), generated and measured in HBM; u halo = one grid layer per "
            "side pushed over NVLink peer memory",
            "l2": "inputs larger than L2 (no flush needed)" if b_res > 4 * 126e6 else "working set may fit in L2",
            "tile_quantum": dm.stats()["quantum"],
        },
        "gpu_launches": launches,
        "clocks": clocks,
        "residual": {
            "ms": ms_res,
            "contributions_per_s": R_global / (ms_res * 1e-3),
            "algorithmic_bytes_per_gpu": int(b_res),
            "achieved_gbs_per_gpu": b_res / (ms_res * 1e-3) / 1e9,
            "frac": b_res / (ms_res * 1e-3) / 1e9 / peak,
        },
        "jacobian": {
            "ms": ms_jac,
            "contributions_per_s": R_global / (ms_jac * 1e-3),
            "algorithmic_bytes_per_gpu": int(b_jac),
            "achieved_gbs_per_gpu": b_jac / (ms_jac * 1e-3) / 1e9,
            "frac": b_jac / (ms_jac * 1e-3) / 1e9 / peak,
        },
        "roofline": {
            "bound": "hbm",
            "achieved": (b_res + b_jac) / (ms_step * 1e-3) / 1e9,
            "peak": peak,
            "unit": "GB/s",
            "frac": (b_res + b_jac) / (ms_step * 1e-3) / 1e9 / peak,
            "traffic": None,
            "kernel": "fvm_tile_kernel (modes residual + jacobian), per GPU",
            "peak_source": peak_src,
        },
        "pattern_build_ms": dm.build_ms,
        "setup_s": setup_s,
    }
    if args.solve:
        # the whole nonlinear solve on the partition: pyfvm.newton's loop with the device
        # BiCGStab (Jacobi) as jacobian_solver; nothing per-record or per-slot leaves HBM
        from pyfvm_b200 import solvers

        solvers.DeviceSystem(dm, s.data.ptr, _partitioned=True)  # NVRTC-compiles the SpMV kernel
        barrier()
        ts = time.time()
        u_own, steps, lin_its = solvers.newton_distributed(
            dp, numpy.zeros(part.n_owned), tol=args.newton_tol, lin_tol=1e-8, verbose=(rank == 0)
        )
        barrier()
        umax = float(u_own.max()) if len(u_own) else 0.0
        if dist is not None:
            import torch

            t = torch.tensor([umax], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            umax = float(t.item())
        out["newton_solve"] = {
            "seconds": time.time() - ts,
            "newton_steps": steps,
            "bicgstab_iterations": lin_its,
            "tol": args.newton_tol,
            "max_u": umax,
            "what": "pyfvm.newton from u = 0 with a Jacobi-BiCGStab jacobian_solver, residual / "
            "Jacobian refresh / SpMV / updates in HBM across the ranks",
        }
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dp.halo.close()
        dist.destroy_process_group()


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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the whole K+W run should end within ~2 minutes at the
    # ~2 M contributions/s this path reaches on one core
    per_step = 120.0 * 2.0e6 / (args.warmup + args.steps)
    size = int(min(args.cpu_size, max(101, (per_step / 6.0) ** 0.5 + 1)))
    R, times = cpu_assembly_seconds(size, repeat=args.warmup + args.steps)
    t = times[args.warmup:]
    sec = sum(t) / len(t)
    val = R / sec
    base = {
        "value": val,
        "unit": UNIT,
        "cores": 1,
        "host_cores": len(os.sched_getaffinity(0)),
        "kind": "port",
        "sample": f"each step = oracle assembly of a {size}x{size} rectangle_tri sample ({R} contributions)",
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
                "config": {
                    "workload": f"convection-diffusion (BASELINE config 4), bounded sample rectangle_tri "
                    f"{size}x{size}; CPU oracle = restated reference algorithm (reference is README-only)"
                },
                "cpu_baseline": base,
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
    ap.add_argument("--no-cold", action="store_true", help="skip the cold discretize_linear call")
    ap.add_argument("--no-residual", action="store_true", help="skip the residual leg")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: BASELINE config 4 (default, the headline); c5: Bratu on cube_tetra slabs")
    ap.add_argument("--cube", type=int, default=200, help="points per side of the c5 cube")
    ap.add_argument("--solve", action="store_true", help="c5: also run the full Newton solve")
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
