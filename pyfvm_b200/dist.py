"""Row-partitioned multi-GPU path: one process per GPU, vertices
split into contiguous ranges, every rank assembles complete owned rows.

* ``Partition(mesh, rank, world)`` cuts a global mesh: the records (columns of
  ``idx_hierarchy``) with at least one owned endpoint, the local vertex set
  ``ghost-low | owned | ghost-high`` (monotone in the global numbering, so the
  fold order of every owned row equals the single-GPU run: results are
  bit-identical), and the halo plan.  ``Partition.mesh`` exposes the meshplex
  attribute contract for the local piece, with the GLOBAL
  control volumes and boundary flags.
* Linear assembly needs no communication at all.
* Residual / Jacobian need ``u`` at the ghost vertices, once per evaluation:
  - ``TorchHalo``: grouped ``isend``/``irecv`` through ``torch.distributed``
    (NCCL on GPUs, gloo in the CPU tests); ghosts of one owner are contiguous in
    the local numbering, so the receive side needs no unpack;
  - ``PeerHalo``: CUDA-IPC peer memory.  Every rank PUSHES the owned values a
    neighbour needs straight into that neighbour's ghost slots (P2P stores over
    NVLink) and releases an epoch flag; the neighbour's stream waits on the flag
    (``fvm_halo_push`` / ``fvm_halo_wait``).  ``u`` is double-buffered by epoch
    parity so a fast rank never overwrites values a slow one still reads.

The reference is single-process; there is no reference code for this file.
"""
import ctypes
import os
from ctypes import byref, c_void_p

import numpy

from . import codegen
from .engine import Engine


def row_bounds(n, world):
    """Contiguous, balanced vertex ranges: rank r owns [b[r], b[r+1])."""
    return [(n * r) // world for r in range(world + 1)]


class LocalMesh:
    """The meshplex attribute contract for one rank's piece of a mesh.  Records
    are kept as a flat list: ``idx_hierarchy`` has shape (2, R_local) and "cell
    masks" are per record."""

    def __init__(self, points, idx, ce, ei_dot_ei, cv, boundary, rec_global, cell_of_record,
                 parent, verts=None):
        self._verts = verts
        self.points = points
        nc = numpy.zeros((len(points), 3))
        nc[:, : points.shape[1]] = points
        self.node_coords = nc
        self.idx_hierarchy = idx
        self.ce_ratios = ce
        self.ei_dot_ei = ei_dot_ei
        self.control_volumes = cv
        self.is_boundary_point = boundary
        self._rec_global = rec_global
        self._cell_of_record = cell_of_record
        self._parent = parent
        self.cell_dim = getattr(parent, "cell_dim", 3 if numpy.asarray(parent.idx_hierarchy).ndim == 4 else 2)

    @property
    def surface_areas(self):
        """The GLOBAL boundary measure of the local vertices (dGamma terms)."""
        return numpy.asarray(self._parent.surface_areas)[self._verts]

    def get_vertex_mask(self, subdomain=None):
        if subdomain is None:
            return numpy.s_[:]
        mask = numpy.asarray(subdomain.is_inside(self.node_coords.T), dtype=bool)
        if getattr(subdomain, "is_boundary_only", False):
            mask = mask & self.is_boundary_point
        return mask

    def get_cell_mask(self, subdomain=None):
        """Per local record: is the record's cell inside ``subdomain``?"""
        if subdomain is None:
            return numpy.s_[:]
        cells = numpy.asarray(self._parent.get_cell_mask(subdomain))
        return cells[self._cell_of_record]


class Partition:
    """Rank ``rank``'s share of ``mesh`` (a global mesh every rank can see)."""

    def __init__(self, mesh, rank, world, bounds=None):
        idx = numpy.asarray(mesh.idx_hierarchy)
        ncell = idx.shape[-1]
        i0 = idx[0].reshape(-1)
        i1 = idx[1].reshape(-1)
        pts = numpy.asarray(mesh.points if hasattr(mesh, "points") else mesh.node_coords)
        n = len(pts)
        self.rank, self.world, self.n_global = rank, world, n
        self.bounds = list(bounds) if bounds is not None else row_bounds(n, world)
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        own0 = (i0 >= lo) & (i0 < hi)
        own1 = (i1 >= lo) & (i1 < hi)
        rec = numpy.nonzero(own0 | own1)[0]  # ascending: the global record order is kept
        verts = numpy.unique(numpy.concatenate([i0[rec], i1[rec], numpy.arange(lo, hi)]))
        self.verts = verts  # local -> global vertex id, ascending
        self.n_local = len(verts)
        self.n_ghost_lo = int(numpy.searchsorted(verts, lo))
        self.n_owned = hi - lo
        self.row_range = (self.n_ghost_lo, self.n_owned)
        l0 = numpy.searchsorted(verts, i0[rec])
        l1 = numpy.searchsorted(verts, i1[rec])
        # flat record id -> cell id (the last axis of idx_hierarchy is the cell)
        cell_of_record = rec % ncell
        self.mesh = LocalMesh(
            points=numpy.ascontiguousarray(pts[verts]),
            idx=numpy.stack([l0, l1]).astype(numpy.int64),
            ce=numpy.asarray(mesh.ce_ratios).reshape(-1)[rec],
            ei_dot_ei=numpy.asarray(mesh.ei_dot_ei).reshape(-1)[rec],
            cv=numpy.asarray(mesh.control_volumes)[verts],
            boundary=numpy.asarray(mesh.is_boundary_point)[verts],
            rec_global=rec,
            cell_of_record=cell_of_record,
            parent=mesh,
            verts=verts,
        )
        # ---- halo plan -------------------------------------------------------
        # receive: ghosts sorted by global id are grouped by owner (owners are
        # contiguous ranges), so each owner's ghosts form one local slice
        ghost = numpy.concatenate([verts[: self.n_ghost_lo], verts[self.n_ghost_lo + self.n_owned:]])
        ghost_local = numpy.concatenate(
            [numpy.arange(self.n_ghost_lo), numpy.arange(self.n_ghost_lo + self.n_owned, self.n_local)]
        )
        owner = numpy.searchsorted(self.bounds, ghost, side="right") - 1
        self.recv = {}  # peer -> (first local ghost slot, count)
        for p in numpy.unique(owner):
            sel = ghost_local[owner == p]
            assert sel[-1] - sel[0] + 1 == len(sel)
            self.recv[int(p)] = (int(sel[0]), len(sel))
        # send: my vertices that appear in a record whose other endpoint peer p owns
        # (= exactly p's ghosts owned by me), ascending -> same order on both sides
        self.send = {}  # peer -> local indices (int32) of the owned values p needs
        a = numpy.concatenate([i0[rec][own0[rec] & ~own1[rec]], i1[rec][own1[rec] & ~own0[rec]]])
        b = numpy.concatenate([i1[rec][own0[rec] & ~own1[rec]], i0[rec][own1[rec] & ~own0[rec]]])
        if len(a):
            bowner = numpy.searchsorted(self.bounds, b, side="right") - 1
            for p in numpy.unique(bowner):
                mine = numpy.unique(a[bowner == p])
                self.send[int(p)] = numpy.searchsorted(verts, mine).astype(numpy.int32)
        self.neighbours = sorted(set(self.recv) | set(self.send))

    def owned(self, x_global):
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        return numpy.ascontiguousarray(x_global[lo:hi])


class SlabPartition:
    """Rank ``rank``'s slab of a structured ``rectangle_tri`` (2-D) or
    ``cube_tetra`` (3-D) grid, generated locally: the global mesh is never
    materialised, so the largest configurations only ever exist in pieces.

    The slowest grid axis is cut; a slab carries one ghost grid layer on each
    interior side, which holds every cell that touches an owned vertex.  Same
    fields as :class:`Partition` (local numbering = global numbering shifted by
    the first local vertex: monotone)."""

    def __init__(self, axes, rank, world, device=False):
        """``device=True``: points, cells and geometry of the slab are generated in
        HBM (``pyfvm_b200.device_meshes``); nothing per-record exists on the host."""
        from . import meshes

        axes = [numpy.asarray(a, dtype=numpy.float64) for a in axes]
        dim = len(axes)
        assert dim in (2, 3)
        nslow = len(axes[-1])
        layer = int(numpy.prod([len(a) for a in axes[:-1]]))  # vertices per grid layer
        cuts = [(nslow * r) // world for r in range(world + 1)]
        assert all(cuts[r + 1] > cuts[r] for r in range(world)), "more ranks than grid layers"
        j0, j1 = cuts[rank], cuts[rank + 1]
        g0, g1 = max(j0 - 1, 0), min(j1 + 1, nslow)
        self.rank, self.world = rank, world
        self.n_global = layer * nslow
        self.bounds = [c * layer for c in cuts]
        # ghost layers are not domain boundary; the true boundary is known analytically
        nloc = layer * (g1 - g0)
        lid = numpy.arange(nloc)
        on = numpy.zeros(nloc, dtype=bool)
        stride = 1
        for a in axes[:-1]:
            k = (lid // stride) % len(a)
            on |= (k == 0) | (k == len(a) - 1)
            stride *= len(a)
        ks = lid // layer + g0
        on |= (ks == 0) | (ks == nslow - 1)
        if device:
            from . import device_meshes as dm

            if dim == 2:
                mesh = dm.device_rectangle_tri(axes[0], axes[1][g0:g1], parity=g0, boundary=on)
            else:
                mesh = dm.device_cube_tetra(axes[0], axes[1], axes[2][g0:g1], parity=g0, boundary=on)
        else:
            if dim == 2:
                pts, cells = meshes.rectangle_tri(axes[0], axes[1][g0:g1], parity=g0)
            else:
                pts, cells = meshes.cube_tetra(axes[0], axes[1], axes[2][g0:g1], parity=g0)
            mesh = meshes.Mesh(pts, cells)
            mesh._boundary = on
        self.mesh = mesh
        self.n_local = nloc
        self.n_ghost_lo = (j0 - g0) * layer
        self.n_owned = (j1 - j0) * layer
        self.row_range = (self.n_ghost_lo, self.n_owned)
        self.verts = numpy.arange(g0 * layer, g1 * layer)
        self.vert_offset = g0 * layer  # local id + offset = global id
        self.recv, self.send = {}, {}
        if j0 > g0:  # ghost layer below: owned by rank - 1; it needs my first layer
            self.recv[rank - 1] = (0, layer)
            self.send[rank - 1] = numpy.arange(self.n_ghost_lo, self.n_ghost_lo + layer, dtype=numpy.int32)
        if g1 > j1:
            self.recv[rank + 1] = (self.n_ghost_lo + self.n_owned, layer)
            first = self.n_ghost_lo + self.n_owned - layer
            self.send[rank + 1] = numpy.arange(first, first + layer, dtype=numpy.int32)
        self.neighbours = sorted(self.recv)

    def owned(self, x_global):
        return numpy.ascontiguousarray(x_global[self.bounds[self.rank] : self.bounds[self.rank + 1]])


# ---------------------------------------------------------------------------
# halo exchange back ends
# ---------------------------------------------------------------------------
class TorchHalo:
    """``u_local`` is a torch tensor (CPU with gloo, CUDA with NCCL); one grouped
    isend/irecv round per evaluation."""

    def __init__(self, part, group=None):
        import torch

        self.part, self.group = part, group
        self._idx = {p: torch.from_numpy(ix.astype(numpy.int64)) for p, ix in part.send.items()}

    def exchange(self, u_local):
        import torch
        import torch.distributed as dist

        ops, keep = [], []
        for p, ix in sorted(self._idx.items()):
            if ix.device != u_local.device:
                ix = self._idx[p] = ix.to(u_local.device)
            buf = u_local.index_select(0, ix)  # pack
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, p, group=self.group))
        for p, (start, count) in sorted(self.part.recv.items()):
            ops.append(dist.P2POp(dist.irecv, u_local[start : start + count], p, group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return u_local


class PeerHalo:
    """CUDA-IPC peer memory: push into the neighbour's ghost slots + epoch flag.

    One device allocation per rank, exported once:
    ``[u buffer 0 | u buffer 1 | flags (one uint64 per rank) | timeout flag]``.
    """

    TIMEOUT_S = 20.0

    def __init__(self, part, eng, group=None):
        import torch.distributed as dist

        self.part, self.eng = part, eng
        lib = eng.lib
        nl = part.n_local
        self.ustride = (8 * nl + 255) & ~255
        self.flag_off = 2 * self.ustride
        nbytes = self.flag_off + 8 * part.world + 64
        self.epoch_dev_ptr = None  # set below: [.. | flags | timeout (4) | pad | epoch (8)]
        self.buf = eng.alloc(nbytes)
        eng._ck(lib.fvm_dev_memset(eng.ctx, self.buf.ptr, 0, nbytes))
        eng.sync()
        handle = ctypes.create_string_buffer(64)
        eng._ck(lib.fvm_ipc_export(eng.ctx, self.buf.ptr, handle))
        mine = {
            "handle": handle.raw,
            "ustride": self.ustride,
            "flag_off": self.flag_off,
            "recv": part.recv,  # owner -> (first ghost slot, count) in MY local numbering
        }
        infos = [mine]
        if part.world > 1:
            infos = [None] * part.world
            dist.all_gather_object(infos, mine, group=group)
        self.peer = {}
        for p in part.neighbours:
            base = c_void_p()
            eng._ck(lib.fvm_ipc_open(eng.ctx, infos[p]["handle"], byref(base)))
            start, count = infos[p]["recv"][part.rank]
            assert count == len(part.send[p]), "halo plans of the two sides disagree"
            self.peer[p] = {
                "base": base.value,
                "ustride": infos[p]["ustride"],
                "ghost_byte_off": 8 * start,
                "flag": base.value + infos[p]["flag_off"] + 8 * part.rank,
                "send": eng.to_device(part.send[p]),
                "n": count,
            }
        # device array of pointers to the flags my neighbours write
        flags = numpy.array(
            [self.buf.ptr + self.flag_off + 8 * p for p in sorted(part.recv)], dtype=numpy.uint64
        )
        self.flag_ptrs = eng.to_device(flags) if len(flags) else None
        self.n_flags = len(flags)
        self.timeout_ptr = self.buf.ptr + self.flag_off + 8 * part.world
        self.epoch_dev_ptr = self.timeout_ptr + 16  # device copy of the epoch (graph replays)
        self.epoch = 0
        # the whole push as one launch (fvm_halo_push_all): a device table of the peers
        PEER = numpy.dtype([("idx", "<u8"), ("dst0", "<u8"), ("dst1", "<u8"), ("flag", "<u8"), ("n", "<i8")])
        tab = numpy.zeros(len(self.peer), dtype=PEER)
        for i, (p, d) in enumerate(sorted(self.peer.items())):
            first = d["base"] + d["ghost_byte_off"]
            tab[i] = (d["send"].ptr, first, first + d["ustride"], d["flag"], d["n"])
        self.peer_tab = eng.to_device(tab) if len(tab) else None
        self.push_counters = eng.alloc(4 * max(1, len(tab)))
        eng._ck(lib.fvm_dev_memset(eng.ctx, self.push_counters.ptr, 0, 4 * max(1, len(tab))))
        self.max_send = max([d["n"] for d in self.peer.values()], default=0)
        self.single_launch = int(os.environ.get("FVM_HALO_SINGLE_LAUNCH", "1"))
        if part.world > 1:
            dist.barrier(group=group)  # every mapping exists before the first push

    def u_ptr(self, k):
        return self.buf.ptr + k * self.ustride

    def next_buffer(self):
        """Buffer index the next exchange will use (owned part to be filled first)."""
        return (self.epoch + 1) & 1

    def fused_wait(self):
        """Argument tuple that makes the next tile-kernel launch wait for this epoch's
        pushes inside the kernel (``fvm_assemble_halo``); None without neighbours."""
        if not self.n_flags:
            return None
        return (self.flag_ptrs.ptr, self.n_flags, self.epoch, self.timeout_ptr)

    def exchange(self, wait=True):
        """Push my owned boundary values of buffer ``next_buffer()`` to the
        neighbours and -- ``wait=True`` -- wait for theirs with a stream-ordered
        kernel; with ``wait=False`` the consumer launch waits itself
        (``fused_wait()``).  Returns the device address of the local ``u``.
        Everything is stream-ordered on the ctx stream."""
        eng, lib = self.eng, self.eng.lib
        self.epoch += 1
        k = self.epoch & 1
        src = self.u_ptr(k)
        if self.single_launch and self.peer_tab is not None:
            eng._ck(lib.fvm_halo_push_all(eng.ctx, self.peer_tab.ptr, len(self.peer), self.max_send,
                                          src, k, self.epoch, self.push_counters.ptr))
        else:
            for p, d in sorted(self.peer.items()):
                dst = d["base"] + k * d["ustride"] + d["ghost_byte_off"]
                eng._ck(lib.fvm_halo_push(eng.ctx, src, d["send"].ptr, d["n"], dst, d["flag"], self.epoch))
        if self.n_flags and wait:
            eng._ck(
                lib.fvm_halo_wait(
                    eng.ctx, self.flag_ptrs.ptr, self.n_flags, self.epoch, self.TIMEOUT_S,
                    self.timeout_ptr,
                )
            )
        return src

    def check_timeout(self):
        flag = numpy.zeros(1, dtype=numpy.int32)
        self.eng._ck(self.eng.lib.fvm_d2h(self.eng.ctx, flag.ctypes.data, self.timeout_ptr, 4))
        if flag[0]:
            raise RuntimeError("halo exchange timed out waiting for a neighbour rank")

    def close(self):
        self.eng.sync()
        for d in self.peer.values():
            self.eng.lib.fvm_ipc_close(self.eng.ctx, d["base"])
        self.peer = {}


class PeerReduce:
    """All-reduce of two doubles over peer memory (``k_allreduce2_peer``): every rank owns
    an inbox ``[2 parities][world][2]`` and a flag word per rank, exported with CUDA IPC;
    the device table lists every rank's (inbox, flags) as seen from this rank."""

    def __init__(self, part, eng, group=None):
        import torch.distributed as dist

        self.part, self.eng = part, eng
        lib, world = eng.lib, part.world
        self.inbox_bytes = 2 * world * 2 * 8
        nbytes = self.inbox_bytes + 8 * world + 64
        self.buf = eng.alloc(nbytes)
        eng._ck(lib.fvm_dev_memset(eng.ctx, self.buf.ptr, 0, nbytes))
        eng.sync()
        self.flags_ptr = self.buf.ptr + self.inbox_bytes
        self.epoch_ptr = self.flags_ptr + 8 * world
        self.timeout_ptr = self.epoch_ptr + 16
        handle = ctypes.create_string_buffer(64)
        eng._ck(lib.fvm_ipc_export(eng.ctx, self.buf.ptr, handle))
        handles = [handle.raw]
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, handle.raw, group=group)
        self._opened = []
        tab = numpy.zeros((world, 2), dtype=numpy.uint64)
        for r in range(world):
            if r == part.rank:
                base = self.buf.ptr
            else:
                b = c_void_p()
                eng._ck(lib.fvm_ipc_open(eng.ctx, handles[r], byref(b)))
                base = b.value
                self._opened.append(base)
            tab[r] = (base, base + self.inbox_bytes)
        self.table = eng.to_device(tab)
        if world > 1:
            dist.barrier(group=group)

    def close(self):
        self.eng.sync()
        for base in self._opened:
            self.eng.lib.fvm_ipc_close(self.eng.ctx, base)
        self._opened = []


# ---------------------------------------------------------------------------
# distributed problems
# ---------------------------------------------------------------------------
def discretize_linear(obj, part, device=None, fmad=False):
    """Owned rows of ``pyfvm.discretize_linear`` on this rank: a CSR block of shape
    (n_owned, n_global) with GLOBAL column ids, and the owned rhs.  No collective."""
    from scipy import sparse

    from .problem import LinearProblem

    lp = LinearProblem(obj, part.mesh, device=device, fmad=fmad, row_range=part.row_range)
    A_local, rhs = lp.assemble()
    A = sparse.csr_matrix(
        (A_local.data, global_columns(part, lp.dmesh), A_local.indptr),
        shape=(part.n_owned, part.n_global), copy=False,
    )
    return A, rhs


def global_columns(part, dmesh):
    """The pattern's column ids in the GLOBAL numbering (int32), cached with the device
    mesh.  A slab's local numbering is the global one shifted by its first vertex."""
    cols = getattr(dmesh, "_global_columns", None)
    if cols is None:
        _, indices = dmesh.pattern()
        off = getattr(part, "vert_offset", None)
        if off is not None:
            cols = indices + numpy.int32(off)
        else:
            cols = part.verts[indices].astype(numpy.int32)
        dmesh._global_columns = cols
    return cols


class DistributedFvm:
    """``pyfvm.discretize`` on one rank of a row partition: ``eval(u_owned)`` and
    ``jacobian_values(u_owned)`` exchange the ghost values first."""

    def __init__(self, obj, part, halo="peer", device=None, fmad=False, group=None, fused=True):
        """``halo``: "peer" (CUDA-IPC pushes) or "torch" (NCCL send/recv).  ``fused``
        (peer only): the tile kernel itself waits for the neighbours' pushes, and only
        where a tile touches ghost vertices, instead of a wait kernel in front of it."""
        from .problem import _DeviceProblem

        self.part = part
        import os

        fused = int(os.environ.get("FVM_FUSED_HALO", "1" if fused else "0"))  # experiments
        self.fused = bool(fused) and halo == "peer"
        self.s = _DeviceProblem(
            obj, part.mesh, "nonlinear", device=device, fmad=fmad, row_range=part.row_range
        )
        self.eng = self.s.eng
        self.kind = halo
        if halo == "peer":
            self.halo = PeerHalo(part, self.eng, group=group)
        elif halo == "torch":
            import torch

            self.halo = TorchHalo(part, group=group)
            self.u_t = torch.zeros(part.n_local, dtype=torch.float64, device=f"cuda:{self.eng.device}")
        else:
            raise ValueError("halo must be 'peer' or 'torch'")
        self.F = self.eng.alloc(8 * max(1, part.n_owned))

    # -- device-resident steps ------------------------------------------------
    def set_owned(self, u_owned):
        """Host -> device: the owned part of the next ``u``."""
        u_owned = numpy.ascontiguousarray(u_owned, dtype=numpy.float64)
        assert u_owned.shape == (self.part.n_owned,)
        off = 8 * self.part.n_ghost_lo
        if self.kind == "peer":
            dst = self.halo.u_ptr(self.halo.next_buffer()) + off
            self.eng._ck(self.eng.lib.fvm_h2d(self.eng.ctx, dst, u_owned.ctypes.data, u_owned.nbytes))
        else:
            import torch

            self.u_t[self.part.n_ghost_lo : self.part.n_ghost_lo + self.part.n_owned] = (
                torch.from_numpy(u_owned).to(self.u_t.device)
            )

    def exchange(self, mode=None):
        """Ghost values in place; returns the device address of the local ``u``.
        ``mode``: the launch this exchange feeds.  A launch that never reads ``u`` (the
        Jacobian of a u-independent operator) gets no exchange at all: every epoch that
        is pushed must also be waited for, or the double buffering of the peer halo
        would let a fast rank overwrite values a slow rank still reads."""
        if mode is not None and not self.s.functors[mode].needs_u:
            if self.kind == "peer":
                return self.halo.u_ptr(self.halo.next_buffer())
            return self.u_t.data_ptr()
        if self.kind == "peer":
            return self.halo.exchange(wait=not self.fused)
        import torch

        self.halo.exchange(self.u_t)
        torch.cuda.current_stream().synchronize()  # NCCL stream -> ctx stream hand-off
        return self.u_t.data_ptr()

    def _halo_arg(self, mode):
        if not self.fused or not self.s.functors[mode].needs_u:
            return None
        return self.halo.fused_wait()

    def step_device(self, mode):
        """Exchange + launch in one host call (``fvm_halo_step``): pushes the owned boundary
        values of the NEXT u buffer and evaluates ``mode`` on it.  Falls back to
        ``exchange()`` + launch when the fused peer halo does not apply."""
        s, h = self.s, self.halo if self.kind == "peer" else None
        if (h is None or not self.fused or not h.n_flags or h.peer_tab is None
                or not s.functors[mode].needs_u or mode == codegen.MODE_LINEAR):
            u_ptr = self.exchange(mode)
            (self.residual_device if mode == codegen.MODE_RESIDUAL else self.jacobian_device)(u_ptr)
            return u_ptr
        h.epoch += 1
        k = h.epoch & 1
        src = h.u_ptr(k)
        eng = self.eng
        res = mode == codegen.MODE_RESIDUAL
        eng._ck(eng.lib.fvm_halo_step(
            s.dmesh.handle, s.kernels[mode].handle, s._p(s.vmask), s._p(s.dirid), src,
            None if res else s.data.ptr, self.F.ptr if res else None,
            h.flag_ptrs.ptr, h.n_flags, h.epoch, h.timeout_ptr,
            h.peer_tab.ptr, len(h.peer), h.max_send, k, h.push_counters.ptr,
        ))
        eng.launches += 1
        return src

    def residual_device(self, u_ptr):
        self.s.launch(codegen.MODE_RESIDUAL, u_ptr=u_ptr, vec_ptr=self.F.ptr,
                      halo=self._halo_arg(codegen.MODE_RESIDUAL))

    def jacobian_device(self, u_ptr):
        self.s.launch(codegen.MODE_JACOBIAN, u_ptr=u_ptr, data_ptr=self.s.data.ptr,
                      halo=self._halo_arg(codegen.MODE_JACOBIAN))

    # -- host API ---------------------------------------------------------------
    def eval(self, u_owned):
        """Owned rows of ``f.eval(u)`` [README:121]."""
        self.set_owned(u_owned)
        self.step_device(codegen.MODE_RESIDUAL)
        out = self.eng.host_result(self.part.n_owned, numpy.float64)
        self.F.download(out)
        if self.kind == "peer":
            self.halo.check_timeout()
        return out

    def jacobian_values(self, u_owned):
        """Owned rows of ``jacobian.get_linear_operator(u)`` [README:116] as a CSR
        block (n_owned, n_global) with global column ids."""
        from scipy import sparse

        self.set_owned(u_owned)
        self.step_device(codegen.MODE_JACOBIAN)
        data = self.eng.host_result(self.s.nnz, numpy.float64)
        self.s.data.download(data)
        if self.kind == "peer":
            self.halo.check_timeout()
        indptr, _ = self.s.dmesh.pattern()
        return sparse.csr_matrix(
            (data, global_columns(self.part, self.s.dmesh), indptr),
            shape=(self.part.n_owned, self.part.n_global), copy=False,
        )

    def norm2_owned_sq(self):
        """Sum of squares of the owned residual (deterministic); the caller adds
        the ranks' values with one all-reduce of a double."""
        return self.eng.norm2(self.F.ptr, self.part.n_owned) ** 2


def parity_selfcheck(group=None, verbose=False):
    """N-rank result == 1-rank result, bit for bit, on small meshes (structured, jittered
    and randomly renumbered): linear assembly (no collective), residual and Jacobian
    values through the fused peer-memory halo, several epochs.  Every rank partitions the
    same global mesh; rank 0 gathers the owned rows and compares them with its own
    single-GPU engine.  Returns True on every rank iff everything matched.  Used by
    ``bench.py --gpus N`` (``"dist_parity"``) and the multi-GPU tests."""
    import torch
    import torch.distributed as dist

    from . import form_language as fl
    from . import meshes
    from .problem import discretize as discretize_1gpu
    from .problem import discretize_linear as discretize_linear_1gpu

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    import sympy

    class ConvDiff:
        def apply(self, u):
            a = sympy.Matrix([2, 1, 0])
            return fl.integrate(
                lambda x: -fl.n_dot_grad(u(x)) + fl.dot(a.T, fl.n) * u(x), fl.dS
            ) - fl.integrate(lambda x: 1.0, fl.dV)

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    class Bratu:
        def apply(self, u):
            return fl.integrate(lambda x: -fl.n_dot_grad(u(x)), fl.dS) - fl.integrate(
                lambda x: 2.0 * sympy.exp(u(x)), fl.dV
            )

        def dirichlet(self, u):
            return [(u, fl.Boundary())]

    def renumbered(p, c, seed):
        perm = numpy.random.default_rng(seed).permutation(len(p))
        inv = numpy.empty(len(p), dtype=numpy.int64)
        inv[perm] = numpy.arange(len(p))
        return numpy.ascontiguousarray(p[perm]), inv[c]

    xs = numpy.linspace(0.0, 1.0, 97)
    cases = []
    p, c = meshes.jitter(*meshes.rectangle_tri(xs, xs[:71]), scale=0.2, seed=0)
    cases.append(("rect 97x71 jittered", p, c))
    cases.append(("rect 97x71 renumbered", *renumbered(p, c, 1)))
    zs = numpy.linspace(0.0, 1.0, 17)
    p, c = meshes.jitter(*meshes.cube_tetra(zs, zs, zs), scale=0.15, seed=2)
    cases.append(("cube 17^3 jittered", p, c))
    ok = True
    for name, p, c in cases:
        mesh = meshes.Mesh(p, c)
        n = len(p)
        part = Partition(mesh, rank, world)
        rng = numpy.random.default_rng(5)
        us = [0.3 * rng.standard_normal(n) for _ in range(3)]
        dp = DistributedFvm(Bratu(), part, halo="peer", group=group)
        Fs, Js = [], []
        for u in us:  # three epochs: both u buffers and a buffer reuse
            Fs.append(dp.eval(part.owned(u)).tobytes())
            Js.append(dp.jacobian_values(part.owned(u)).data.tobytes())
        dist.barrier(group=group)
        dp.halo.close()
        A, b = discretize_linear(ConvDiff(), part)
        mine = (Fs, Js, A.data.tobytes(), A.indices.tobytes(), b.tobytes())
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        if rank == 0:
            f1, j1 = discretize_1gpu(Bratu(), mesh)
            A1, b1 = discretize_linear_1gpu(ConvDiff(), mesh)
            same = (
                b"".join(g[2] for g in gathered) == A1.data.tobytes()
                and b"".join(g[3] for g in gathered) == A1.indices.tobytes()
                and b"".join(g[4] for g in gathered) == b1.tobytes()
            )
            for k, u in enumerate(us):
                same = same and b"".join(g[0][k] for g in gathered) == f1.eval(u).tobytes()
                same = same and (
                    b"".join(g[1][k] for g in gathered) == j1.get_linear_operator(u).data.tobytes()
                )
            if verbose:
                print(f"dist parity, {name}, {world} ranks: bit-identical={same}", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())
