"""Worker bodies of the multi-process tests (importable by ``spawn``)."""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _mesh(kind, n, shuffled=False):
    from pyfvm_b200 import meshes
    from tests import problems

    p, c = (problems.rectangle(n, n - 3) if kind == "rect" else problems.cube(n))
    p, c = meshes.jitter(p, c, scale=0.2, seed=0)
    if shuffled:  # no lexicographic numbering: ghost columns are reached by the cold gather
        p, c = problems.shuffle_numbering(p, c, seed=11)
    return meshes.Mesh(p, c)


def gloo_halo_worker(rank, world, port, kind, n, out_dir):
    """CPU / gloo: partition, halo exchange of u, and the oracle run per rank on
    the local piece; writes what the parent compares against the global oracle."""
    import torch
    import torch.distributed as dist

    import oracle
    from pyfvm_b200 import dist as fd
    from tests import problems

    dist.init_process_group(
        "gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world
    )
    try:
        mesh = _mesh(kind, n)
        nglob = len(mesh.points)
        part = fd.Partition(mesh, rank, world)
        u_glob = numpy.random.default_rng(7).standard_normal(nglob)
        # owned values only; ghosts arrive through the exchange
        u_loc = torch.full((part.n_local,), float("nan"), dtype=torch.float64)
        g0 = part.n_ghost_lo
        u_loc[g0 : g0 + part.n_owned] = torch.from_numpy(part.owned(u_glob))
        fd.TorchHalo(part).exchange(u_loc)
        u_loc = u_loc.numpy()
        assert numpy.array_equal(u_loc, u_glob[part.verts]), "halo exchange delivered wrong values"

        # the oracle on the local piece: owned rows must equal the global rows
        f, jac = oracle.discretize(problems.bratu(oracle.form_language), part.mesh)
        F = f.eval(u_loc)[g0 : g0 + part.n_owned]
        J = jac.get_linear_operator(u_loc)[g0 : g0 + part.n_owned]
        A, b = oracle.discretize_linear(problems.convection_diffusion(oracle.form_language), part.mesh)
        A = A[g0 : g0 + part.n_owned]
        numpy.savez(
            os.path.join(out_dir, f"rank{rank}.npz"),
            F=F, J_data=J.data, J_indices=part.verts[J.indices], J_indptr=J.indptr,
            A_data=A.data, A_indices=part.verts[A.indices], A_indptr=A.indptr,
            b=b[g0 : g0 + part.n_owned], lo=part.bounds[rank], hi=part.bounds[rank + 1],
        )
        # one all-reduce of a double: the distributed ||F||^2
        t = torch.tensor([float(numpy.dot(F, F))], dtype=torch.float64)
        dist.all_reduce(t)
        if rank == 0:
            numpy.save(os.path.join(out_dir, "normsq.npy"), t.numpy())
    finally:
        dist.destroy_process_group()


def gpu_worker():
    """Launched by torchrun on a multi-GPU box: every rank runs its partition on its
    GPU with both halo back ends; rank 0 compares the gathered result with the
    single-GPU engine bit for bit."""
    import torch
    import torch.distributed as dist

    import pyfvm_b200 as eng
    from pyfvm_b200 import dist as fd
    from tests import problems

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    try:
        cases = (("rect", 67, "nonlinear_conv", False), ("cube", 15, "bratu", False),
                 ("rect", 45, "nonlinear_conv", True), ("cube", 11, "bratu", True))
        for kind, n, pname, shuffled in cases:
            mesh = _mesh(kind, n, shuffled)
            nglob = len(mesh.points)
            part = fd.Partition(mesh, rank, world)
            prob = getattr(problems, pname)(eng.form_language)
            rng = numpy.random.default_rng(3)
            us = [0.3 * rng.standard_normal(nglob) for _ in range(3)]
            results = {}
            for halo in ("peer", "peer-unfused", "torch"):
                dp = fd.DistributedFvm(prob, part, halo=halo.split("-")[0], fused=halo == "peer")
                if shuffled and dp.s.dmesh.stats()["uncovered_slots"] == 0:
                    print(f"rank {rank}: shuffled mesh but no uncovered slots", flush=True)
                    ok = False
                Fs, Js = [], []
                for u in us:  # several epochs: exercises the double buffering
                    Fs.append(dp.eval(part.owned(u)))
                    Js.append(dp.jacobian_values(part.owned(u)))
                results[halo] = (Fs, Js)
                if halo.startswith("peer"):
                    dist.barrier()
                    dp.halo.close()
            A_loc, b_loc = fd.discretize_linear(
                problems.convection_diffusion(eng.form_language), part
            )
            gathered = [None] * world
            payload = {
                h: ([f.tobytes() for f in r[0]], [j.data.tobytes() for j in r[1]],
                    [j.indices.tobytes() for j in r[1]])
                for h, r in results.items()
            }
            payload["lin"] = (A_loc.data.tobytes(), A_loc.indices.tobytes(), b_loc.tobytes())
            dist.all_gather_object(gathered, payload)
            if rank == 0:
                f1, j1 = eng.discretize(prob, mesh)
                A1, b1 = eng.discretize_linear(problems.convection_diffusion(eng.form_language), mesh)
                for h in ("peer", "peer-unfused", "torch"):
                    for k, u in enumerate(us):
                        F = b"".join(g[h][0][k] for g in gathered)
                        Jd = b"".join(g[h][1][k] for g in gathered)
                        Ji = b"".join(g[h][2][k] for g in gathered)
                        J1 = j1.get_linear_operator(u)
                        same = (
                            F == f1.eval(u).tobytes()
                            and Jd == J1.data.tobytes()
                            and Ji == J1.indices.tobytes()
                        )
                        print(f"{kind} {pname} shuffled={shuffled} halo={h} u{k}: "
                              f"bit-identical={same}", flush=True)
                        ok &= same
                lin = (
                    b"".join(g["lin"][0] for g in gathered) == A1.data.tobytes()
                    and b"".join(g["lin"][1] for g in gathered) == A1.indices.tobytes()
                    and b"".join(g["lin"][2] for g in gathered) == b1.tobytes()
                )
                print(f"{kind} linear assembly: bit-identical={lin}", flush=True)
                ok &= lin
        # distributed Newton (Bratu) with the device BiCGStab vs the single-GPU device Newton
        from pyfvm_b200 import solvers

        mesh = _mesh("cube", 13)
        part = fd.Partition(mesh, rank, world)
        prob = problems.bratu(eng.form_language)
        dp = fd.DistributedFvm(prob, part, halo="peer")
        u_own, steps, its = solvers.newton_distributed(dp, numpy.zeros(part.n_owned))
        pieces = [None] * world
        dist.all_gather_object(pieces, u_own)
        dist.barrier()
        dp.halo.close()
        if rank == 0:
            f1, j1 = eng.discretize(prob, mesh)
            u1, steps1 = solvers.newton_device(f1, j1, numpy.zeros(len(mesh.points)))
            err = numpy.abs(numpy.concatenate(pieces) - u1).max()
            good = steps == steps1 and err <= 1e-8
            print(f"distributed Newton: steps={steps} (1 GPU: {steps1}), BiCGStab its={its}, "
                  f"max |u - u_1gpu| = {err:.2e}: ok={good}", flush=True)
            ok &= good
        # the device-resident distributed BiCGStab (CUDA graph + peer-memory all-reduce) against
        # the host-driven loop (NCCL all-reduces) on the same Jacobian system
        dp = fd.DistributedFvm(prob, part, halo="peer")
        dp.set_owned(part.owned(0.1 * numpy.random.default_rng(9).standard_normal(len(mesh.points))))
        up = dp.exchange()
        dp.residual_device(up)
        dp.jacobian_device(up)
        sysm = solvers.DistributedSystem(dp)
        xs = []
        for graph in (True, False):
            xd = dp.eng.alloc(8 * max(1, part.n_owned) + 16)
            its_k, rel = sysm.bicgstab(dp.F.ptr, xd.ptr, tol=1e-10, x0_is_zero=True, graph=graph)
            x = numpy.empty(part.n_owned)
            xd.download(x)
            xs.append((its_k, rel, x))
        dk = float(numpy.abs(xs[0][2] - xs[1][2]).max()) if part.n_owned else 0.0
        scale = float(numpy.abs(xs[1][2]).max()) if part.n_owned else 1.0
        good = xs[0][1] <= 1e-10 and xs[1][1] <= 1e-10 and dk <= 1e-7 * max(scale, 1e-30)
        if rank == 0:
            print(f"distributed BiCGStab: graph its={xs[0][0]} relres={xs[0][1]:.2e}; host-driven "
                  f"its={xs[1][0]} relres={xs[1][1]:.2e}; max diff {dk:.2e}: ok={good}", flush=True)
        ok &= good
        # row-partitioned CG (fvm_cg with the halo / all-reduce tables) on the same symmetric
        # Jacobian system: same solution as BiCGStab, and the same bits on a second run
        xc = []
        for _ in range(2):
            xd = dp.eng.alloc(8 * max(1, part.n_owned) + 16)
            its_c, rel_c = sysm.cg(dp.F.ptr, xd.ptr, tol=1e-10)
            x = numpy.empty(part.n_owned)
            xd.download(x)
            xc.append(x)
        dc = float(numpy.abs(xc[0] - xs[1][2]).max()) if part.n_owned else 0.0
        good = rel_c <= 1e-10 and dc <= 1e-7 * max(scale, 1e-30) and xc[0].tobytes() == xc[1].tobytes()
        if rank == 0:
            print(f"distributed CG: its={its_c} relres={rel_c:.2e}; max diff to BiCGStab {dc:.2e}; "
                  f"repeatable bits: ok={good}", flush=True)
        ok &= good
        # the Chebyshev-preconditioned variant (3 halo exchanges per iteration, an odd number:
        # the buffer parity bookkeeping of the captured graph is exercised)
        xd = dp.eng.alloc(8 * max(1, part.n_owned) + 16)
        its_p, rel_p = sysm.cg(dp.F.ptr, xd.ptr, tol=1e-10, degree=3)
        xp = numpy.empty(part.n_owned)
        xd.download(xp)
        dpp = float(numpy.abs(xp - xs[1][2]).max()) if part.n_owned else 0.0
        good = rel_p <= 1e-10 and dpp <= 1e-7 * max(scale, 1e-30) and its_p < its_c
        if rank == 0:
            print(f"distributed CG + Chebyshev(3): its={its_p} (Jacobi: {its_c}) relres={rel_p:.2e}; "
                  f"max diff to BiCGStab {dpp:.2e}: ok={good}", flush=True)
        ok &= good
        dist.barrier()
        sysm.close()
        dp.halo.close()
        # distributed Newton again with CG as the linear solver
        dp = fd.DistributedFvm(prob, part, halo="peer")
        u_cg, steps_cg, its_cg = solvers.newton_distributed(dp, numpy.zeros(part.n_owned), linear_solver="cg")
        dcg = float(numpy.abs(u_cg - u_own).max()) if part.n_owned else 0.0
        good = steps_cg == steps and dcg <= 1e-8
        if rank == 0:
            print(f"distributed Newton with CG: steps={steps_cg}, CG its={its_cg}, "
                  f"max |u - u_bicgstab| = {dcg:.2e}: ok={good}", flush=True)
        ok &= good
        dist.barrier()
        dp.halo.close()
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # any rank's failure fails the run
        ok = bool(flag.item())
    finally:
        dist.destroy_process_group()
    if rank == 0:
        print("DIST_GPU_OK" if ok else "DIST_GPU_FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    gpu_worker()
