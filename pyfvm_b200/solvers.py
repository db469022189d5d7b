"""The step *after* the hot path, kept in HBM: SpMV on the
assembled system and a Jacobi-preconditioned BiCGStab, so that the
``jacobian_solver(u0, rhs)`` callback of ``pyfvm.newton`` [README:113-121] -- a
host ``spsolve`` in the README -- and the whole Newton loop can run without the
matrix ever crossing PCIe.

The reference leaves the solver to the user (scipy ``spsolve`` [README:51, 117],
pyamg [README:67-77]); nothing here replaces reference code.  ``y = A x`` runs
the assembly's own tile kernel in SpMV mode (``fvm_spmv``): same tiles, x gathered
from the staged vertex tables, rows folded in slot order -- deterministic.
"""
import ctypes
from ctypes import byref, c_double, c_void_p  # noqa: F401

import numpy

from . import codegen


class DeviceSystem:
    """``A`` = (pattern of a device mesh, CSR values in HBM) for square systems."""

    def __init__(self, dmesh, values_ptr, _partitioned=False):
        assert _partitioned or (dmesh.n_rows == dmesh.n and dmesh.row_base == 0), (
            "square single-GPU systems; row partitions use DistributedSystem"
        )
        self.dmesh, self.eng, self.values = dmesh, dmesh.eng, values_ptr
        self.n = dmesh.n_rows  # length of the vectors the solver works on (owned rows)
        eng = self.eng
        if getattr(eng, "_spmv_kernel", None) is None:
            h = c_void_p()
            eng._ck(eng.lib.fvm_spmv_kernel_create(eng.ctx, byref(h)))
            eng._spmv_kernel = h
        self._diag = None
        self._work = None

    def matvec(self, x_ptr, y_ptr):
        e = self.eng
        e._ck(e.lib.fvm_spmv(self.dmesh.handle, e._spmv_kernel, self.values, x_ptr, y_ptr))
        e.launches += 1

    def diagonal(self):
        """Device buffer with the diagonal of the current values."""
        if self._diag is None:
            self._diag = self.eng.alloc(8 * self.n)
        self.eng._ck(self.eng.lib.fvm_csr_diagonal(self.dmesh.handle, self.values, self._diag.ptr))
        return self._diag

    # -- vector helpers ---------------------------------------------------------
    def _dot(self, x, y):
        out = c_double()
        self.eng._ck(self.eng.lib.fvm_dot(self.eng.ctx, x, y, self.n, byref(out)))
        return out.value

    def _norm(self, x):
        return self.eng.norm2(x, self.n)

    def _dot2(self, x1, y1, x2, y2):
        """(x1.y1, x2.y2) with one host synchronisation."""
        out = (c_double * 2)()
        self.eng._ck(self.eng.lib.fvm_dot2(self.eng.ctx, x1, y1, x2, y2, self.n, out))
        return out[0], out[1]

    def _axpby(self, a, x, b, y, z):
        self.eng._ck(self.eng.lib.fvm_axpby(self.eng.ctx, a, x, b, y, z, self.n))

    def bicgstab_graph(self, b_ptr, x_ptr, tol=1e-10, maxiter=None, check_every=16):
        """``A x = b`` from x0 = 0 with ``fvm_bicgstab``: all scalars stay in HBM and one
        iteration is a CUDA graph replayed ``check_every`` times between host looks at
        the residual (single-GPU systems)."""
        e, n = self.eng, self.n
        maxiter = maxiter or 20 * int(numpy.sqrt(n)) + 200
        if getattr(self, "_gwork", None) is None:
            self._gwork = e.alloc(8 * (9 * ((n + 1) & ~1) + 16))
        its, rel = ctypes.c_int(), c_double()
        rc = e.lib.fvm_bicgstab(
            self.dmesh.handle, e._spmv_kernel, self.values, b_ptr, x_ptr, self._gwork.ptr, tol,
            maxiter, check_every, byref(its), byref(rel),
        )
        if rc != 0:
            msg = e.lib.fvm_last_error().decode()
            if "BiCGStab" in msg:
                raise ArithmeticError(f"{msg}: relative residual {rel.value:.3e}")
            e._ck(rc)
        e.launches += 2 * its.value
        return its.value, rel.value

    def _dist_args(self):
        return None  # one GPU: no halo, no all-reduce

    def _after_graph_solve(self):
        pass

    def _lambda_max(self):
        """Gershgorin's bound on the spectrum of D^-1 A (this GPU's rows)."""
        out = c_double()
        self.eng._ck(self.eng.lib.fvm_csr_gershgorin(self.dmesh.handle, self.values, byref(out)))
        return out.value

    def cg(self, b_ptr, x_ptr, tol=1e-10, maxiter=None, check_every=16, degree=1):
        """``A x = b`` for a symmetric definite ``A`` (Dirichlet rows as pyfvm writes them
        are fine) with ``fvm_cg``: Jacobi-preconditioned CG from x0 = b / diag(A), one SpMV
        and two reductions per iteration, scalars in HBM, CUDA graph replayed between host
        looks at the residual.  On a row partition the halo pushes and the all-reduces go
        over peer memory.  Returns (iterations, relative residual); ``ArithmeticError`` on
        breakdown (an indefinite or unsymmetric matrix) or no convergence.

        ``degree > 1``: the preconditioner is the Chebyshev polynomial of that degree in the
        Jacobi-scaled operator (``fvm_cg_poly``): ``degree - 1`` more SpMVs per iteration, no
        more reductions -- about the same number of SpMVs in 3-4 times fewer iterations,
        which pays where an iteration is bound by its rank-wide synchronisations."""
        e, n = self.eng, self.n
        maxiter = maxiter or self._default_maxiter()
        if getattr(self, "_gwork", None) is None:
            self._gwork = e.alloc(8 * (9 * ((n + 1) & ~1) + 16) + 64)
        its, rel = ctypes.c_int(), c_double()
        lmax = self._lambda_max() if degree > 1 else 0.0  # (before the epochs are synchronised)
        args = self._dist_args()
        rc = e.lib.fvm_cg_poly(
            self.dmesh.handle, e._spmv_kernel, self.values, b_ptr, x_ptr, self._gwork.ptr, tol,
            maxiter, check_every, degree, lmax, None if args is None else ctypes.byref(args),
            byref(its), byref(rel),
        )
        msg = e.lib.fvm_last_error().decode() if rc else ""
        self._after_graph_solve()
        if rc != 0:
            if msg.startswith("CG "):
                raise ArithmeticError(f"{msg}: relative residual {rel.value:.3e}")
            raise RuntimeError(msg)
        e.launches += degree * (its.value + 1)
        return its.value, rel.value

    def _default_maxiter(self):
        return 20 * int(numpy.sqrt(self.n)) + 200

    def solve(self, b_ptr, x_ptr, method="bicgstab", tol=1e-10, maxiter=None):
        """``A x = b`` from the solver's own initial guess; ``method``: "bicgstab" (any
        matrix), "cg" (symmetric definite, Jacobi) or "cg-cheb<k>" (CG with the Chebyshev
        polynomial preconditioner of degree k)."""
        if method == "cg":
            return self.cg(b_ptr, x_ptr, tol=tol, maxiter=maxiter)
        if method.startswith("cg-cheb") and method[7:].isdigit():
            return self.cg(b_ptr, x_ptr, tol=tol, maxiter=maxiter, degree=int(method[7:]))
        if method != "bicgstab":
            raise ValueError(f"unknown linear solver {method!r} (bicgstab, cg, cg-cheb<degree>)")
        return self.bicgstab(b_ptr, x_ptr, tol=tol, maxiter=maxiter, x0_is_zero=True)

    def bicgstab(self, b_ptr, x_ptr, tol=1e-10, maxiter=None, x0_is_zero=False, graph=True):
        """Solve ``A x = b`` (device pointers), right-preconditioned with the
        diagonal.  Stops at ``||r|| <= tol * ||b||``.  Returns (iterations,
        relative residual); raises ``ArithmeticError`` on breakdown or no
        convergence.  From x0 = 0 on one GPU the device-resident, graph-replayed
        ``fvm_bicgstab`` does the work; otherwise this host-driven loop."""
        if graph and x0_is_zero and type(self) is DeviceSystem:
            return self.bicgstab_graph(b_ptr, x_ptr, tol=tol, maxiter=maxiter)
        e, lib, n = self.eng, self.eng.lib, self.n
        maxiter = maxiter or 20 * int(numpy.sqrt(n)) + 200
        if self._work is None:
            self._work = [e.alloc(8 * n + 16) for _ in range(8)]
        r, r0, p, v, s, t, ph, sh = (w.ptr for w in self._work)
        d = self.diagonal().ptr
        bnorm = self._norm(b_ptr)
        if bnorm == 0.0:
            e._ck(lib.fvm_dev_memset(e.ctx, x_ptr, 0, 8 * n))
            return 0, 0.0
        if x0_is_zero:
            e._ck(lib.fvm_dev_memset(e.ctx, x_ptr, 0, 8 * n))
            self._axpby(1.0, b_ptr, 0.0, b_ptr, r)
        else:
            self.matvec(x_ptr, r)
            self._axpby(1.0, b_ptr, -1.0, r, r)
        self._axpby(1.0, r, 0.0, r, r0)
        e._ck(lib.fvm_dev_memset(e.ctx, p, 0, 8 * n))
        e._ck(lib.fvm_dev_memset(e.ctx, v, 0, 8 * n))
        rho = alpha = omega = 1.0
        # three reductions per iteration: r0.v | (t.s, t.t) | (r0.r, r.r)
        rho_new, rr = self._dot2(r0, r, r, r)
        res = numpy.sqrt(rr)
        for it in range(1, maxiter + 1):
            if res <= tol * bnorm:
                return it - 1, res / bnorm
            if rho_new == 0.0 or omega == 0.0:
                raise ArithmeticError("BiCGStab breakdown")
            beta = (rho_new / rho) * (alpha / omega)
            self._axpby(1.0, p, -omega, v, p)
            self._axpby(1.0, r, beta, p, p)
            e._ck(lib.fvm_vdiv(e.ctx, p, d, ph, n))
            self.matvec(ph, v)
            r0v = self._dot(r0, v)
            if r0v == 0.0:
                raise ArithmeticError("BiCGStab breakdown")
            alpha = rho_new / r0v
            self._axpby(1.0, r, -alpha, v, s)
            e._ck(lib.fvm_vdiv(e.ctx, s, d, sh, n))
            self.matvec(sh, t)
            ts, tt = self._dot2(t, s, t, t)
            # t = 0 means s = 0: the half step already solved the system
            omega = ts / tt if tt != 0.0 else 0.0
            self._axpby(1.0, x_ptr, alpha, ph, x_ptr)
            self._axpby(1.0, x_ptr, omega, sh, x_ptr)
            self._axpby(1.0, s, -omega, t, r)
            rho = rho_new
            rho_new, rr = self._dot2(r0, r, r, r)
            res = numpy.sqrt(rr)
        if res <= tol * bnorm:
            return maxiter, res / bnorm
        raise ArithmeticError(f"BiCGStab did not converge: relative residual {res / bnorm:.3e}")


def solve_linear(lp, tol=1e-10, maxiter=None, method="bicgstab"):
    """``discretize_linear`` + solve, all on the device: assemble, BiCGStab (or CG for a
    symmetric definite system, ``method="cg"``), and only the solution comes back (host
    array)."""
    lp.assemble_device()
    sysm = DeviceSystem(lp.dmesh, lp.data.ptr)
    x = lp.eng.alloc(8 * lp.n)
    iters, relres = sysm.solve(lp.vec.ptr, x.ptr, method=method, tol=tol, maxiter=maxiter)
    out = numpy.empty(lp.n)
    x.download(out)
    return out, iters, relres


def jacobian_solver(jacobian, tol=1e-10, maxiter=None, method="bicgstab"):
    """A ``jacobian_solver(u0, rhs) -> du`` callback for ``pyfvm.newton`` [README:113-117]
    that refreshes the Jacobian values and solves on the device."""
    s = jacobian._s
    sysm = DeviceSystem(s.dmesh, s.data.ptr)
    b, x = s.eng.alloc(8 * s.n), s.eng.alloc(8 * s.n)

    def solver(u0, rhs):
        jacobian.values_device(s.upload_u(u0))
        b.upload(numpy.ascontiguousarray(rhs, dtype=numpy.float64))
        sysm.solve(b.ptr, x.ptr, method=method, tol=tol, maxiter=maxiter)
        du = numpy.empty(s.n)
        x.download(du)
        return du

    return solver


def newton_device(f, jacobian, u0, tol=1e-10, max_iter=20, lin_tol=1e-10, verbose=False,
                  linear_solver="bicgstab"):
    """``pyfvm.newton`` [README:121] with everything in HBM: residual, norm,
    Jacobian refresh, BiCGStab (``linear_solver="cg"``: CG, for symmetric definite
    Jacobians such as Bratu's [README:96-110]) and the update; one double per step crosses
    PCIe.  Same stopping rule and iteration cap as the reference's newton."""
    s = f._s
    e, n = s.eng, s.n
    sysm = DeviceSystem(s.dmesh, s.data.ptr)
    du = e.alloc(8 * n)
    u_ptr = s.upload_u(u0)
    steps = 0
    while True:
        f.eval_device(u_ptr)  # F in s.vec
        nrm = e.norm2(s.vec.ptr, n)
        if verbose:
            print(f"||F(u)|| = {nrm:e}")
        if nrm < tol:
            break
        if steps >= max_iter:
            raise RuntimeError(f"Newton did not converge in {max_iter} steps (||F|| = {nrm:e})")
        jacobian.values_device(u_ptr)
        # J du = -F: solve for +F and subtract
        sysm.solve(s.vec.ptr, du.ptr, method=linear_solver, tol=lin_tol)
        e._ck(e.lib.fvm_axpy(e.ctx, -1.0, du.ptr, u_ptr, n))
        steps += 1
    out = numpy.empty(n)
    s.u.download(out)
    return out, steps


class DistributedSystem(DeviceSystem):
    """The row-partitioned system of a :class:`pyfvm_b200.dist.DistributedFvm`: every
    rank holds its owned rows; vectors are the owned parts.  ``matvec`` pushes the
    owned boundary entries of x to the neighbours over NVLink peer memory and the
    SpMV kernel waits for theirs itself (``fvm_spmv_halo``); dot products are the
    deterministic local sums plus one all-reduce of a double."""

    def __init__(self, dp, values_ptr=None, group=None):
        import torch

        from .dist import PeerHalo

        super().__init__(dp.s.dmesh, values_ptr or dp.s.data.ptr, _partitioned=True)
        self.dp, self.part, self.group = dp, dp.part, group
        self.vhalo = PeerHalo(dp.part, self.eng, group=group)  # its own buffers / flags
        self._reduce = None  # peer-memory all-reduce, mapped on first use
        self._red = torch.zeros(1, dtype=torch.float64, device=f"cuda:{self.eng.device}")
        self._red2 = torch.zeros(2, dtype=torch.float64, device=f"cuda:{self.eng.device}")

    def _allreduce(self, v):
        if self.part.world == 1:
            return v
        import torch.distributed as dist

        self._red[0] = v
        dist.all_reduce(self._red, group=self.group)
        return float(self._red.item())

    def _dot(self, x, y):
        return self._allreduce(super()._dot(x, y))

    def _norm(self, x):
        return numpy.sqrt(self._allreduce(self.eng.norm2(x, self.n) ** 2))

    def _dot2(self, x1, y1, x2, y2):
        a, b = super()._dot2(x1, y1, x2, y2)
        if self.part.world == 1:
            return a, b
        import torch
        import torch.distributed as dist

        self._red2[0], self._red2[1] = a, b
        dist.all_reduce(self._red2, group=self.group)
        a, b = self._red2.tolist()
        return a, b

    def _default_maxiter(self):
        return 20 * int(numpy.sqrt(max(1, self.part.n_global))) + 200

    def _lambda_max(self):
        v = super()._lambda_max()
        if self.part.world == 1:
            return v
        import torch.distributed as dist

        self._red[0] = v
        dist.all_reduce(self._red, op=dist.ReduceOp.MAX, group=self.group)
        return float(self._red.item())

    def _dist_args(self):
        """The halo and all-reduce tables of this rank for the graph-replayed solvers; the
        device-resident halo epoch is brought in step with the host's first."""
        from .dist import PeerReduce
        from .engine import DistArgs

        e, h = self.eng, self.vhalo
        if self._reduce is None:
            self._reduce = PeerReduce(self.part, e, group=self.group)
        rd = self._reduce
        if h.epoch & 1:  # start every solve on an even epoch: the graph's buffer order is fixed
            h.epoch += 1
        ep = numpy.array([h.epoch], dtype=numpy.uint64)
        e._ck(e.lib.fvm_h2d(e.ctx, h.epoch_dev_ptr, ep.ctypes.data, 8))
        return DistArgs(
            halo_peers_dev=None if h.peer_tab is None else h.peer_tab.ptr,
            n_halo_peers=len(h.peer), max_send=h.max_send, push_counters_dev=h.push_counters.ptr,
            halo_flags_dev=None if h.flag_ptrs is None else h.flag_ptrs.ptr, n_halo_flags=h.n_flags,
            halo_timeout_dev=h.timeout_ptr, xbuf=(c_void_p * 2)(h.u_ptr(0), h.u_ptr(1)),
            ghost_lo=self.part.n_ghost_lo, halo_epoch_dev=h.epoch_dev_ptr,
            reduce_peers_dev=rd.table.ptr, world=self.part.world, rank=self.part.rank,
            reduce_epoch_dev=rd.epoch_ptr, reduce_timeout_dev=rd.timeout_ptr,
        )

    def _after_graph_solve(self):
        e, h = self.eng, self.vhalo
        ep = numpy.zeros(1, dtype=numpy.uint64)
        e._ck(e.lib.fvm_d2h(e.ctx, ep.ctypes.data, h.epoch_dev_ptr, 8))
        h.epoch = int(ep[0])  # the exchanges the graph replays made

    def bicgstab_graph(self, b_ptr, x_ptr, tol=1e-10, maxiter=None, check_every=16):
        """``A x = b`` from x0 = 0 on the row partition without the host in the loop
        (``fvm_bicgstab_dist``): the iteration is a CUDA graph, its two halo exchanges push
        over NVLink with device-resident epochs, its three reductions are peer-memory
        all-reduces; the host looks at the residual every ``check_every`` iterations."""
        e, n = self.eng, self.n
        maxiter = maxiter or self._default_maxiter()
        if getattr(self, "_gwork", None) is None:
            self._gwork = e.alloc(8 * (9 * ((n + 1) & ~1) + 16) + 64)
        args = self._dist_args()
        its, rel = ctypes.c_int(), c_double()
        rc = e.lib.fvm_bicgstab_dist(
            self.dmesh.handle, e._spmv_kernel, self.values, b_ptr, x_ptr, self._gwork.ptr, tol,
            maxiter, check_every, ctypes.byref(args), byref(its), byref(rel),
        )
        msg = e.lib.fvm_last_error().decode() if rc else ""
        self._after_graph_solve()
        if rc != 0:
            if "BiCGStab" in msg:
                raise ArithmeticError(f"{msg}: relative residual {rel.value:.3e}")
            raise RuntimeError(msg)
        e.launches += 2 * its.value
        return its.value, rel.value

    def bicgstab(self, b_ptr, x_ptr, tol=1e-10, maxiter=None, x0_is_zero=False, graph=True):
        if graph and x0_is_zero and int(__import__("os").environ.get("FVM_DIST_GRAPH", "1")):
            return self.bicgstab_graph(b_ptr, x_ptr, tol=tol, maxiter=maxiter)
        return super().bicgstab(b_ptr, x_ptr, tol=tol, maxiter=maxiter, x0_is_zero=x0_is_zero,
                                graph=False)

    def matvec(self, x_ptr, y_ptr):
        e, h = self.eng, self.vhalo
        dst = h.u_ptr(h.next_buffer()) + 8 * self.part.n_ghost_lo
        self._axpby(1.0, x_ptr, 0.0, x_ptr, dst)  # owned part into the exchange buffer
        xl = h.exchange(wait=False)
        fw = h.fused_wait()
        if fw is None:
            e._ck(e.lib.fvm_spmv(self.dmesh.handle, e._spmv_kernel, self.values, xl, y_ptr))
        else:
            e._ck(e.lib.fvm_spmv_halo(self.dmesh.handle, e._spmv_kernel, self.values, xl, y_ptr, *fw))
        e.launches += 1

    def close(self):
        self.vhalo.close()
        if self._reduce is not None:
            self._reduce.close()


def newton_distributed(dp, u0_owned, tol=1e-10, max_iter=20, lin_tol=1e-10, verbose=False, group=None,
                       linear_solver="bicgstab"):
    """``pyfvm.newton`` [README:121] on a row partition (one rank per GPU): residual,
    Jacobian refresh, BiCGStab and update stay in HBM; per step a handful of
    doubles cross NCCL (norms / dot products) and the ghost values cross NVLink.
    Returns (owned part of the solution, steps, total BiCGStab iterations)."""
    e, part = dp.eng, dp.part
    n = part.n_owned
    sysm = DistributedSystem(dp, group=group)
    du = e.alloc(8 * max(1, n))
    off = 8 * part.n_ghost_lo
    assert dp.kind == "peer", "newton_distributed drives the peer-memory halo"
    dp.set_owned(u0_owned)
    u_ptr = dp.exchange()
    steps = lin_its = 0
    try:
        while True:
            dp.residual_device(u_ptr)
            nrm = sysm._norm(dp.F.ptr)
            if verbose and part.rank == 0:
                print(f"||F(u)|| = {nrm:e}", flush=True)
            if nrm < tol:
                break
            if steps >= max_iter:
                raise RuntimeError(f"Newton did not converge in {max_iter} steps (||F|| = {nrm:e})")
            dp.jacobian_device(u_ptr)
            its, _ = sysm.solve(dp.F.ptr, du.ptr, method=linear_solver, tol=lin_tol)
            lin_its += its
            dp.halo.check_timeout()  # a halo wait that gave up must not feed the next step
            sysm.vhalo.check_timeout()
            nxt = dp.halo.u_ptr(dp.halo.next_buffer()) + off
            sysm._axpby(1.0, u_ptr + off, -1.0, du.ptr, nxt)  # u_new = u - du, owned part
            u_ptr = dp.exchange()
            steps += 1
        out = numpy.empty(n)
        e._ck(e.lib.fvm_d2h(e.ctx, out.ctypes.data, u_ptr + off, out.nbytes))
        dp.halo.check_timeout()
        sysm.vhalo.check_timeout()
    finally:
        if part.world > 1:
            import torch.distributed as dist

            dist.barrier(group=group)
        sysm.close()
    return out, steps, lin_its
