"""ctypes binding of ``libfvm_b200.so`` (C-ABI in ``include/fvm_b200.h``) and
the per-mesh device state.

There is no CPU fallback: if the CUDA library is missing or no B200 is
visible, every entry point raises.
"""
import ctypes
import hashlib
import os
import weakref
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64
from ctypes import c_size_t, c_uint, c_uint8, c_void_p

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfvm_b200.so")



class FvmError(RuntimeError):
    pass


class MeshDesc(ctypes.Structure):
    # mirrors fvm_mesh_desc in include/fvm_b200.h
    _fields_ = [
        ("dim", c_int32),
        ("n_vertices", c_int32),
        ("row_base", c_int32),
        ("n_rows", c_int32),
        ("n_records", c_int64),
        ("index_bytes", c_int32),
        ("on_device", c_int32),
        ("idx0", c_void_p),
        ("idx1", c_void_p),
        ("ce_ratios", c_void_p),
        ("edge_len", c_void_p),
        ("rec_mask", c_void_p),
        ("coords", c_void_p),
        ("control_volumes", c_void_p),
        ("tile_quantum", c_int32),
        ("cell_dim", c_int32),
        ("drop_edge_len", c_int32),
        ("coord_cols", c_int32),
    ]


class DistArgs(ctypes.Structure):
    # mirrors fvm_dist_args in include/fvm_b200.h
    _fields_ = [
        ("halo_peers_dev", c_void_p),
        ("n_halo_peers", c_int32),
        ("max_send", c_int64),
        ("push_counters_dev", c_void_p),
        ("halo_flags_dev", c_void_p),
        ("n_halo_flags", c_int32),
        ("halo_timeout_dev", c_void_p),
        ("xbuf", c_void_p * 2),
        ("ghost_lo", c_int64),
        ("halo_epoch_dev", c_void_p),
        ("reduce_peers_dev", c_void_p),
        ("world", c_int32),
        ("rank", c_int32),
        ("reduce_epoch_dev", c_void_p),
        ("reduce_timeout_dev", c_void_p),
    ]


class KernelOpts(ctypes.Structure):
    # mirrors fvm_kernel_opts in include/fvm_b200.h
    _fields_ = [
        ("mode", c_int32),
        ("flags", c_uint),
        ("fmad", c_int32),
        ("consumer_warps", c_int32),
        ("stages", c_int32),
        ("ctas_per_sm", c_int32),
        ("producers", c_int32),
    ]


# every symbol include/fvm_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fvm_version": (c_int, []),
    "fvm_last_error": (c_char_p, []),
    "fvm_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "fvm_ctx_destroy": (c_int, [c_void_p]),
    "fvm_ctx_sync": (c_int, [c_void_p]),
    "fvm_ctx_set_stream": (c_int, [c_void_p, c_void_p, c_int]),
    "fvm_ctx_info": (c_int, [c_void_p, c_char_p, c_size_t, POINTER(c_int), POINTER(c_size_t)]),
    "fvm_dev_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "fvm_dev_free": (c_int, [c_void_p, c_void_p]),
    "fvm_dev_memset": (c_int, [c_void_p, c_void_p, c_int, c_size_t]),
    "fvm_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "fvm_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "fvm_d2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "fvm_d2h_async": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "fvm_copy_fence": (c_int, [c_void_p]),
    "fvm_copy_sync": (c_int, [c_void_p]),
    "fvm_host_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "fvm_host_free": (c_int, [c_void_p, c_void_p]),
    "fvm_timer_begin": (c_int, [c_void_p]),
    "fvm_timer_end": (c_int, [c_void_p, POINTER(c_float)]),
    "fvm_mesh_create": (c_int, [c_void_p, POINTER(MeshDesc), POINTER(c_void_p)]),
    "fvm_mesh_destroy": (c_int, [c_void_p]),
    "fvm_mesh_info": (
        c_int,
        [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int32), POINTER(c_float)],
    ),
    "fvm_mesh_stats": (c_int, [c_void_p, POINTER(c_int64), c_int]),
    "fvm_mesh_debug_copy": (c_int, [c_void_p, c_int, c_void_p, c_size_t]),
    "fvm_mesh_get_pattern": (c_int, [c_void_p, c_void_p, c_void_p]),
    "fvm_mesh_pattern_dev": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p)]),
    "fvm_mesh_update_geometry": (
        c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int],
    ),
    "fvm_mesh_set_surface_areas": (c_int, [c_void_p, c_void_p, c_int32, c_int]),
    "fvm_mesh_set_halfedge_values": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int]),
    "fvm_gen_rectangle_tri": (
        c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    ),
    "fvm_gen_cube_tetra": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    ),
    "fvm_geometry": (
        c_int,
        [c_void_p, c_int, c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "fvm_mesh_boundary_tri": (c_int, [c_void_p, c_void_p]),
    "fvm_boundary_vertices": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int32, c_void_p]),
    "fvm_record_mask": (
        c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_uint8, c_void_p],
    ),
    "fvm_kernel_create": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_void_p)]),
    "fvm_kernel_launch_shape": (
        c_int,
        [c_void_p, c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)],
    ),
    "fvm_kernel_destroy": (c_int, [c_void_p]),
    "fvm_kernel_cache_stats": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "fvm_kernel_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "fvm_kernel_source": (c_int, [c_char_p, c_char_p, c_size_t, POINTER(c_size_t)]),
    "fvm_assemble": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "fvm_spmv_kernel_create": (c_int, [c_void_p, POINTER(c_void_p)]),
    "fvm_spmv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fvm_spmv_halo": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, ctypes.c_uint64, c_void_p],
    ),
    "fvm_csr_diagonal": (c_int, [c_void_p, c_void_p, c_void_p]),
    "fvm_dot": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double)]),
    "fvm_dot2": (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double)],
    ),
    "fvm_axpby": (c_int, [c_void_p, c_double, c_void_p, c_double, c_void_p, c_void_p, c_int64]),
    "fvm_vdiv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
    "fvm_bicgstab": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int, c_int,
         POINTER(c_int), POINTER(c_double)],
    ),
    "fvm_bicgstab_dist": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int, c_int,
         c_void_p, POINTER(c_int), POINTER(c_double)],
    ),
    "fvm_device_cache_stats": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "fvm_device_cache_flush": (c_int, [c_void_p]),
    "fvm_cg_poly": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int, c_int,
         c_int, c_double, c_void_p, POINTER(c_int), POINTER(c_double)],
    ),
    "fvm_csr_gershgorin": (c_int, [c_void_p, c_void_p, POINTER(c_double)]),
    "fvm_cg": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int, c_int,
         c_void_p, POINTER(c_int), POINTER(c_double)],
    ),
    "fvm_norm2": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_double)]),
    "fvm_gather_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "fvm_ipc_export": (c_int, [c_void_p, c_void_p, c_char_p]),
    "fvm_ipc_open": (c_int, [c_void_p, c_char_p, POINTER(c_void_p)]),
    "fvm_ipc_close": (c_int, [c_void_p, c_void_p]),
    "fvm_halo_push": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, ctypes.c_uint64],
    ),
    "fvm_halo_push_all": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, ctypes.c_uint64, c_void_p],
    ),
    "fvm_halo_step": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
         ctypes.c_uint64, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p],
    ),
    "fvm_halo_wait": (
        c_int,
        [c_void_p, c_void_p, c_int, ctypes.c_uint64, c_double, c_void_p],
    ),
    "fvm_assemble_halo": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
         ctypes.c_uint64, c_void_p],
    ),
    "fvm_axpy": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_int64]),
}

_lib = None


def load_library():
    """Load the C-ABI library and bind every declared symbol (no GPU needed)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FvmError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C pyfvm_b200/csrc` (there is no CPU fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class DeviceBuffer:
    """Device allocation owned by python; ``ptr`` is the raw device address."""

    def __init__(self, eng, nbytes):
        self.eng = eng
        self.nbytes = int(nbytes)
        p = c_void_p()
        eng._ck(eng.lib.fvm_dev_alloc(eng.ctx, self.nbytes, byref(p)))
        self.ptr = p.value
        self._fin = weakref.finalize(self, eng.lib.fvm_dev_free, eng.ctx, c_void_p(self.ptr))

    def free(self):
        self._fin()

    def upload(self, arr):
        arr = numpy.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.eng._ck(self.eng.lib.fvm_h2d(self.eng.ctx, self.ptr, _ptr(arr), arr.nbytes))
        return self

    def download(self, out):
        assert out.flags["C_CONTIGUOUS"] and out.nbytes <= self.nbytes
        self.eng._ck(self.eng.lib.fvm_d2h(self.eng.ctx, _ptr(out), self.ptr, out.nbytes))
        return out

    def download_async(self, out_pinned):
        """Copy to a pinned host array on the copy stream (``Engine.copy_sync``
        completes it; ``Engine.copy_fence`` before the buffer is rewritten)."""
        assert out_pinned.flags["C_CONTIGUOUS"] and out_pinned.nbytes <= self.nbytes
        self.eng._ck(
            self.eng.lib.fvm_d2h_async(self.eng.ctx, _ptr(out_pinned), self.ptr, out_pinned.nbytes)
        )
        return out_pinned


class Engine:
    """One device context (= one CUDA stream on one GPU) plus the compiled
    kernel cache.  ``Engine.get(device)`` returns the per-process singleton."""

    _instances = {}

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def __init__(self, device):
        self.lib = load_library()
        self.device = device
        ctx = c_void_p()
        rc = self.lib.fvm_ctx_create(device, byref(ctx))
        if rc != 0:
            raise FvmError(
                "cannot create a B200 context (no CPU fallback): "
                + self.lib.fvm_last_error().decode()
            )
        self.ctx = ctx
        self._kernels = {}
        self.launches = 0  # tile-kernel launches issued (bench's gpu_launches)
        name = ctypes.create_string_buffer(256)
        sm = c_int()
        mem = c_size_t()
        self._ck(self.lib.fvm_ctx_info(ctx, name, 256, byref(sm), byref(mem)))
        self.name, self.sm_count, self.mem_bytes = name.value.decode(), sm.value, mem.value

    def _ck(self, rc):
        if rc != 0:
            raise FvmError(self.lib.fvm_last_error().decode())

    # -- memory ---------------------------------------------------------------
    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr, pad=0):
        """Upload ``arr``; ``pad`` spare bytes after it (the tile kernel's bulk copies are
        rounded up to 16 bytes and may read that far past the last element)."""
        arr = numpy.ascontiguousarray(arr)
        return self.alloc(arr.nbytes + pad).upload(arr)

    def pinned_raw(self, nbytes):
        """Page-locked host allocation as a ctypes buffer (freed when it is collected)."""
        nbytes = max(1, int(nbytes))
        p = c_void_p()
        self._ck(self.lib.fvm_host_alloc(self.ctx, nbytes, byref(p)))
        raw = (ctypes.c_char * nbytes).from_address(p.value)
        weakref.finalize(raw, self.lib.fvm_host_free, self.ctx, c_void_p(p.value))
        return raw

    def pinned_empty(self, shape, dtype):
        """numpy array in page-locked host memory (fast, async-capable copies)."""
        dtype = numpy.dtype(dtype)
        n = int(numpy.prod(shape)) if not numpy.isscalar(shape) else int(shape)
        raw = self.pinned_raw(n * dtype.itemsize)
        return numpy.frombuffer(raw, dtype=dtype, count=n).reshape(shape)

    # -- host arrays for results ------------------------------------------------------
    # A fresh numpy array of 1 GB costs ~45 ms of first-touch page faults on top of the
    # device -> host copy (measured: 12 GB/s into fresh pageable memory against 50 GB/s
    # into pinned memory).  Large results are therefore handed out as views of page-locked
    # buffers from a small pool; a buffer goes back into circulation when the last array
    # (and scipy matrix) referring to it is gone.  Page-locking itself is slow (~0.25 s per
    # GB), so a size is only given a pinned buffer from its SECOND request on: a one-shot
    # discretize_linear never pays for it, a Newton loop pays once.  FVM_PINNED_RESULTS=0
    # turns the pool off.
    POOL_MIN_BYTES = 8 << 20
    POOL_MAX_BYTES = int(os.environ.get("FVM_PINNED_POOL_BYTES", str(6 << 30)))

    def host_result(self, n, dtype):
        """Uninitialised host array of ``n`` entries for a result download."""
        import sys

        dtype = numpy.dtype(dtype)
        nbytes = int(n) * dtype.itemsize
        if nbytes < self.POOL_MIN_BYTES or os.environ.get("FVM_PINNED_RESULTS", "1") == "0":
            return numpy.empty(n, dtype=dtype)
        pool = self.__dict__.setdefault("_result_pool", [])
        for raw in pool:
            # referenced only by the pool (+ the loop variable + getrefcount's argument): every
            # array handed out (and whatever views / scipy matrices hang off it) keeps its
            # buffer's count up
            if nbytes <= len(raw) <= nbytes + nbytes // 2 and sys.getrefcount(raw) <= 3:
                return numpy.frombuffer(raw, dtype=dtype, count=n)
        seen = self.__dict__.setdefault("_result_sizes", {})
        seen[nbytes] = seen.get(nbytes, 0) + 1
        if seen[nbytes] < 2 or sum(len(r) for r in pool) + nbytes > self.POOL_MAX_BYTES:
            return numpy.empty(n, dtype=dtype)
        try:
            raw = self._new_result_buffer((nbytes + (1 << 20) - 1) & ~((1 << 20) - 1))
        except FvmError:  # no page-locked memory to be had: plain pageable array
            return numpy.empty(n, dtype=dtype)
        pool.append(raw)
        return numpy.frombuffer(raw, dtype=dtype, count=n)

    def _new_result_buffer(self, nbytes):
        return self.pinned_raw(nbytes)

    def device_cache_stats(self):
        """(idle bytes, idle blocks, live blocks) of this device's block cache."""
        a, b, c = c_int64(), c_int64(), c_int64()
        self._ck(self.lib.fvm_device_cache_stats(self.ctx, byref(a), byref(b), byref(c)))
        return a.value, b.value, c.value

    def device_cache_flush(self):
        """Give the cached idle device blocks back to the driver."""
        self._ck(self.lib.fvm_device_cache_flush(self.ctx))

    def sync(self):
        self._ck(self.lib.fvm_ctx_sync(self.ctx))

    def copy_fence(self):
        self._ck(self.lib.fvm_copy_fence(self.ctx))

    def copy_sync(self):
        self._ck(self.lib.fvm_copy_sync(self.ctx))

    def timer_begin(self):
        self._ck(self.lib.fvm_timer_begin(self.ctx))

    def timer_end(self):
        ms = c_float()
        self._ck(self.lib.fvm_timer_end(self.ctx, byref(ms)))
        return ms.value

    # -- kernels ----------------------------------------------------------------
    def kernel(self, functors, fmad=False):
        """Compile (NVRTC, sm_100a) or fetch the cached tile kernel of a functor set."""
        key = (hashlib.sha256(functors.source.encode()).hexdigest(), functors.mode, bool(fmad))
        k = self._kernels.get(key)
        if k is None:
            h = c_void_p()
            opts = KernelOpts(
                mode=functors.mode,
                flags=functors.flags,
                fmad=int(fmad),
                consumer_warps=functors.consumer_warps,
                stages=functors.stages,
                ctas_per_sm=functors.ctas_per_sm,
                producers=functors.producers,
            )
            self._ck(
                self.lib.fvm_kernel_create(
                    self.ctx, functors.source.encode(), ctypes.cast(ctypes.pointer(opts), c_void_p), byref(h)
                )
            )
            k = Kernel(self, h, functors)
            self._kernels[key] = k
        return k

    def kernel_cache_stats(self):
        """(NVRTC compilations, cubins loaded from the on-disk cache) of this process."""
        a, b = c_int(), c_int()
        self.lib.fvm_kernel_cache_stats(byref(a), byref(b))
        return a.value, b.value

    def kernel_source(self, functors):
        need = c_size_t()
        self.lib.fvm_kernel_source(functors.source.encode(), None, 0, byref(need))
        buf = ctypes.create_string_buffer(need.value)
        self.lib.fvm_kernel_source(functors.source.encode(), buf, need.value, byref(need))
        return buf.value.decode()

    def norm2(self, x_dev, n):
        out = c_double()
        self._ck(self.lib.fvm_norm2(self.ctx, x_dev, n, byref(out)))
        return out.value


class Kernel:
    def __init__(self, eng, handle, functors):
        self.eng, self.handle, self.functors = eng, handle, functors

    def info(self, dmesh=None):
        r, s, t = c_int(), c_int(), c_int()
        self.eng._ck(self.eng.lib.fvm_kernel_info(self.handle, byref(r), byref(s), byref(t)))
        out = {"regs": r.value, "static_smem": s.value, "max_threads": t.value}
        if dmesh is not None:
            g, b, m = c_int(), c_int(), c_int()
            self.eng._ck(
                self.eng.lib.fvm_kernel_launch_shape(
                    self.handle, dmesh.handle, byref(g), byref(b), byref(m)
                )
            )
            out.update(grid=g.value, block=b.value, smem=m.value)
        return out


def _mesh_arrays(mesh):
    """Vertex coordinates of a mesh object and the number of
    coordinate columns the functors see."""
    coords = numpy.asarray(
        mesh.node_coords if hasattr(mesh, "node_coords") else mesh.points, dtype=numpy.float64
    )
    dim = coords.shape[1]
    if dim == 3:
        planar = getattr(mesh, "_fvm_planar", None)
        if planar is None:
            planar = not numpy.any(coords[:, 2])
            try:
                mesh._fvm_planar = planar  # one pass over the coordinates per mesh object
            except AttributeError:
                pass
        if planar:
            dim = 2  # planar mesh carried with a zero third column
    return coords, dim


def _cell_dim(mesh):
    if hasattr(mesh, "cell_dim"):
        return int(mesh.cell_dim)
    idx = numpy.asarray(mesh.idx_hierarchy)
    return 3 if idx.ndim == 4 else 2  # (2, 3, cells) / (2, 3, 4, cells); flat lists: triangles


class DeviceMesh:
    """Device mirror of one mesh (+ optional per-record subdomain bits): the
    CSR pattern, the half-edge streams and the vertex arrays.  Built once and
    cached per mesh object, like meshplex caches its own derived arrays."""

    def __init__(self, eng, mesh, rec_mask=None, need_el=True, row_range=None, tile_quantum=0):
        tile_quantum = tile_quantum or int(os.environ.get("FVM_TILE_QUANTUM", "0"))
        self.eng = eng
        coords, dim = _mesh_arrays(mesh)
        self.dim = dim
        self.n = coords.shape[0]
        row_base, n_rows = (0, self.n) if row_range is None else row_range
        self.row_base, self.n_rows = row_base, n_rows
        if rec_mask is not None and not isinstance(rec_mask, DeviceBuffer):
            rec_mask = numpy.ascontiguousarray(rec_mask, dtype=numpy.uint8)
        if hasattr(mesh, "device_records"):
            # per-record arrays already live in HBM (pyfvm_b200.device_meshes)
            rec = mesh.device_records(eng, need_el=True)
            self.R = rec["n_records"]
            keep = [rec]  # keeps the device buffers alive until the mesh is built
            xy = eng.to_device(numpy.ascontiguousarray(coords[:, :dim]))
            mask_dev = rec_mask  # built in HBM (DeviceSimplexMesh.device_record_mask) ...
            if rec_mask is not None and not isinstance(rec_mask, DeviceBuffer):
                assert rec_mask.shape == (self.R,)
                mask_dev = eng.to_device(rec_mask)  # ... or handed in as a host array
            d = MeshDesc(
                dim=dim, n_vertices=self.n, row_base=row_base, n_rows=n_rows, n_records=self.R,
                index_bytes=rec["index_bytes"], on_device=1, idx0=rec["idx0"], idx1=rec["idx1"],
                ce_ratios=rec["ce"], edge_len=rec["el"],
                rec_mask=None if mask_dev is None else mask_dev.ptr, coords=xy.ptr,
                control_volumes=None, tile_quantum=tile_quantum, cell_dim=_cell_dim(mesh),
                drop_edge_len=0 if need_el else 1,  # kept only when a functor reads edge_length
            )
            el = True if need_el else None
        else:
            idx = numpy.asarray(mesh.idx_hierarchy)
            idx0 = numpy.ascontiguousarray(idx[0].reshape(-1))
            idx1 = numpy.ascontiguousarray(idx[1].reshape(-1))
            if idx0.dtype not in (numpy.int32, numpy.int64):
                idx0, idx1 = idx0.astype(numpy.int64), idx1.astype(numpy.int64)
            ce = numpy.ascontiguousarray(
                numpy.asarray(mesh.ce_ratios, dtype=numpy.float64).reshape(-1)
            )
            el = None
            if need_el:
                el = numpy.sqrt(numpy.asarray(mesh.ei_dot_ei, dtype=numpy.float64)).reshape(-1)
                el = numpy.ascontiguousarray(el)
            xy = numpy.ascontiguousarray(coords)  # all columns: the device takes the first dim
            cv = numpy.ascontiguousarray(numpy.asarray(mesh.control_volumes, dtype=numpy.float64))
            assert rec_mask is None or rec_mask.shape == idx0.shape
            self.R = idx0.shape[0]
            d = MeshDesc(
                dim=dim, n_vertices=self.n, row_base=row_base, n_rows=n_rows, n_records=self.R,
                index_bytes=idx0.dtype.itemsize, on_device=0, idx0=_ptr(idx0), idx1=_ptr(idx1),
                ce_ratios=_ptr(ce), edge_len=_ptr(el), rec_mask=_ptr(rec_mask), coords=_ptr(xy),
                control_volumes=_ptr(cv), tile_quantum=tile_quantum, cell_dim=_cell_dim(mesh),
                coord_cols=xy.shape[1],
            )
        h = c_void_p()
        eng._ck(eng.lib.fvm_mesh_create(eng.ctx, byref(d), byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, eng.lib.fvm_mesh_destroy, h)
        self.has_el = el is not None
        if hasattr(mesh, "release_device_records"):
            mesh.release_device_records()  # the half-edge streams replace them
        nnz, nhe, ntiles, ms = c_int64(), c_int64(), c_int32(), c_float()
        eng._ck(eng.lib.fvm_mesh_info(h, byref(nnz), byref(nhe), byref(ntiles), byref(ms)))
        self.nnz, self.n_halfedges, self.n_tiles, self.build_ms = (
            nnz.value,
            nhe.value,
            ntiles.value,
            ms.value,
        )
        self._pattern = None
        self.has_sa = False
        self.shared = False  # True while other problems may hold this object (mesh cache)

    TILE_DTYPE = numpy.dtype(
        [("h0", "<i8"), ("r0", "<i4"), ("nr", "<i4"), ("s0", "<i4"), ("nsl", "<i4"),
         ("nhe", "<i4"), ("own_start", "<i4"), ("own_count", "<i4"), ("nhalo", "<i4"),
         ("hstart", "<i4", (8,)), ("hcount", "<u2", (8,)), ("uncovered", "<i4"),
         ("pad", "<i4", (9,))]
    )

    def stats(self):
        """Layout diagnostics of the device mesh (see ``fvm_mesh_stats``)."""
        out = (c_int64 * 10)()
        self.eng._ck(self.eng.lib.fvm_mesh_stats(self.handle, out, 10))
        keys = ["max_he", "max_slots", "max_rows", "max_xtab", "uncovered_slots", "quantum",
                "tiles", "halfedges", "padded_halfedges", "build_chunks"]
        return dict(zip(keys, [int(x) for x in out]))

    def layout(self):
        """Host copies of the device layout (tests / inspection): tile descriptors,
        packed slot metadata, ce_ratios in (padded) half-edge order."""
        tiles = numpy.empty(self.n_tiles, dtype=self.TILE_DTYPE)
        meta = numpy.empty(self.nnz - self.n_rows, dtype=numpy.uint32)
        ce = numpy.empty(self.stats()["padded_halfedges"], dtype=numpy.float64)
        for which, arr in ((0, tiles), (1, meta), (2, ce)):
            self.eng._ck(
                self.eng.lib.fvm_mesh_debug_copy(self.handle, which, _ptr(arr), arr.nbytes)
            )
        return tiles, meta, ce

    def control_volumes(self):
        """Host copy of the device control volumes (owned rows are meaningful)."""
        cv = numpy.empty(self.n, dtype=numpy.float64)
        self.eng._ck(self.eng.lib.fvm_mesh_debug_copy(self.handle, 3, _ptr(cv), cv.nbytes))
        return cv

    def boundary_vertices_tri(self):
        """Triangle meshes: boundary flags of the owned vertices, derived on the
        device from the pattern (edges that belong to exactly one cell)."""
        out = self.eng.alloc(self.n)
        self.eng._ck(self.eng.lib.fvm_dev_memset(self.eng.ctx, out.ptr, 0, self.n))
        self.eng._ck(self.eng.lib.fvm_mesh_boundary_tri(self.handle, out.ptr))
        flags = numpy.empty(self.n, dtype=numpy.uint8)
        out.download(flags)
        return flags.astype(bool)

    def pattern(self):
        """Host copy of ``(indptr, indices)`` (int32), fetched once."""
        if self._pattern is None:
            indptr = self.eng.host_result(self.n_rows + 1, numpy.int32)
            indices = self.eng.host_result(self.nnz, numpy.int32)
            self.eng._ck(
                self.eng.lib.fvm_mesh_get_pattern(self.handle, _ptr(indptr), _ptr(indices))
            )
            self._pattern = (indptr, indices)
        return self._pattern

    def set_halfedge_values(self, diag_vals, off_vals):
        """Replace the ce_ratio / edge_length streams with user-evaluated edge blocks (2R
        entries each, side-major; see ``fvm_mesh_set_halfedge_values``)."""
        d = numpy.ascontiguousarray(diag_vals, dtype=numpy.float64).reshape(-1)
        o = numpy.ascontiguousarray(off_vals, dtype=numpy.float64).reshape(-1)
        assert d.size == o.size == 2 * self.R
        assert not self.shared, "the streams of a cached device mesh must not be overwritten"
        self.eng._ck(
            self.eng.lib.fvm_mesh_set_halfedge_values(self.handle, self.R, _ptr(d), _ptr(o), 0)
        )

    def set_surface_areas(self, sa):
        """Per-vertex boundary measure (``mesh.surface_areas``): the weight of dGamma terms."""
        sa = numpy.ascontiguousarray(sa, dtype=numpy.float64)
        assert sa.shape == (self.n,)
        self.eng._ck(self.eng.lib.fvm_mesh_set_surface_areas(self.handle, _ptr(sa), self.n, 0))
        self.has_sa = True

    def update_geometry(self, ce=None, el=None, coords=None, cv=None):
        """Re-upload per-record / per-vertex geometry of the same connectivity
        (host arrays; the device permutes records into half-edge order)."""
        arrs = [None if a is None else numpy.ascontiguousarray(a, dtype=numpy.float64)
                for a in (ce, el, coords, cv)]
        for a, want in zip(arrs, (self.R, self.R, self.n * self.dim, self.n)):
            if a is not None and a.size != want:
                raise ValueError(f"update_geometry: array of {a.size} entries, expected {want}")
        self.eng._ck(
            self.eng.lib.fvm_mesh_update_geometry(
                self.handle, self.R, self.n, *[_ptr(a) for a in arrs], 0
            )
        )

    def launch(self, kernel, vmask=None, dirid=None, u=None, data=None, vec=None, halo=None):
        """One tile-kernel launch; all arguments are device addresses (int) or None.
        ``halo = (flag pointer array, count, epoch, timeout flag)``: the launch waits for
        the neighbours' pushes inside the kernel (``fvm_assemble_halo``)."""
        if halo is not None:
            flags, n, epoch, tout = halo
            self.eng._ck(
                self.eng.lib.fvm_assemble_halo(
                    self.handle, kernel.handle, vmask, dirid, u, data, vec, flags, n, epoch, tout
                )
            )
        else:
            self.eng._ck(
                self.eng.lib.fvm_assemble(self.handle, kernel.handle, vmask, dirid, u, data, vec)
            )
        self.eng.launches += 1


_mesh_cache = {}


def _array_sig(a):
    """Identity + a sampled content hash of one mesh array (cheap: <= 4096 samples)."""
    a = numpy.asarray(a)
    flat = a.reshape(-1) if a.flags["C_CONTIGUOUS"] else a.ravel()
    step = max(1, flat.size // 4096)
    sample = numpy.ascontiguousarray(flat[::step])
    return (
        a.__array_interface__["data"][0], a.shape, a.dtype.str,
        hashlib.blake2b(sample.tobytes(), digest_size=8).hexdigest(),
    )


def mesh_fingerprint(mesh):
    """What a cached :class:`DeviceMesh` was built from.  meshplex meshes are mutable
    (points moved, edges flipped: its derived arrays are then NEW array objects), and the
    reference re-reads the mesh on every call; a cached device mesh is only reused while
    the arrays it mirrors are the same objects with the same (sampled) content."""
    coords = mesh.node_coords if hasattr(mesh, "node_coords") else mesh.points
    sig = [_array_sig(coords)]
    if hasattr(mesh, "device_records"):
        sig.append(("device", id(getattr(mesh, "_cells", None)), getattr(mesh, "n_records", 0)))
    else:
        sig += [_array_sig(mesh.idx_hierarchy), _array_sig(mesh.ce_ratios),
                _array_sig(mesh.control_volumes)]
    return tuple(sig)


def device_mesh(eng, mesh, rec_mask=None, need_el=False, row_range=None, cache=True):
    """:class:`DeviceMesh` of ``mesh``, cached per mesh object (key: mesh identity,
    subdomain bits, owned row range) and revalidated against ``mesh_fingerprint`` on
    every hit.  Mesh objects that cannot be weak-referenced are never cached (their
    ``id`` could be reused by another mesh)."""
    if rec_mask is None:
        mkey = None
    elif isinstance(rec_mask, DeviceBuffer):
        mkey = rec_mask.digest
    else:
        mkey = hashlib.sha256(rec_mask.tobytes()).hexdigest()
    key = (id(mesh), eng.device, mkey, row_range)
    fp = None
    if cache:
        try:
            weakref.ref(mesh)
            fp = mesh_fingerprint(mesh)
        except TypeError:
            cache = False
    hit = _mesh_cache.get(key) if cache else None
    if hit is not None and (hit[0] != fp or (need_el and not hit[1].has_el)):
        _mesh_cache.pop(key, None)  # stale geometry / connectivity, or no edge lengths
        hit[1].shared = False
        hit = None
    if hit is not None:
        return hit[1]
    dm = DeviceMesh(eng, mesh, rec_mask=rec_mask, need_el=need_el, row_range=row_range)
    dm.shared = bool(cache)
    if cache:
        _mesh_cache[key] = (fp, dm)
        weakref.finalize(mesh, _mesh_cache.pop, key, None)
    return dm
