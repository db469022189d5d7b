// Hand-written sm_100a assembly template for the finite-volume hot path.
//
// One template serves the three numeric entry points of the reference:
//   FVM_MODE_LINEAR   : matrix values + rhs of pyfvm.discretize_linear  [README:49]
//   FVM_MODE_RESIDUAL : f.eval(u)                                       [README:121,136]
//   FVM_MODE_JACOBIAN : jacobian.get_linear_operator(u0) values         [README:116]
// plus FVM_MODE_SPMV : y = A x on the same pattern and tiles (the Krylov solver a
// jacobian_solver callback [README:113-117] can run without leaving HBM); it uses no
// functor: slot value times x of the slot's column, row fold as in the other modes.
//
// It is compiled at run time by NVRTC after a generated prelude that defines
//   FVM_MODE, FVM_DIM, FVM_FLAGS (FVM_F_* bits), FVM_CONSUMER_WARPS, FVM_STAGES,
//   FVM_EDGE_USES_X / _U, FVM_VERT_USES_X / _U, FVM_HAS_VERTEX, FVM_EDGE_FACTORED
// and the three sympy-printed device functors
//   fvm_edge(uk0, uk1, x0_0..2, x1_0..2, edge_ce_ratio, edge_length, mask, &D, &O, &C)
//   fvm_vertex(uk0, control_volume, surface_area, x_0..2, mask, &L, &C)
//   fvm_dirichlet(id, uk0, control_volume, surface_area, x_0..2, &coeff, &val)
// (edge functor outputs per half-edge (row, col): D -> slot (row,row),
//  O -> slot (row,col), C -> rhs/residual of row).
//
// Persistent, warp-specialised CTAs (grid = resident CTAs of the whole GPU);
// CTA b walks tiles b, b+grid, ... through a ring of FVM_STAGES shared-memory
// stages guarded by full/empty mbarriers:
//   producer warp  : reads the tile descriptor and issues TMA bulk copies
//                    (cp.async.bulk, SASS UBLKCP) of every contiguous stream the
//                    tile needs -- ce_ratio / edge_length / mask per half-edge,
//                    offset + table index per slot, row pointer / cv / Dirichlet id
//                    per row, and the tile-local vertex table (own rows + the
//                    halo ranges holding the neighbour columns) of coordinates
//                    and u -- one or more tiles ahead of the math;
//   consumer warps : sweep the tile twice.
//                    Pass A is SLOT-parallel: the tile's off-diagonal CSR slots are cut into
//                    32-slot chunks dealt round-robin to the warps; lane l takes
//                    slot c+l, folds the slot's ce_ratios in half-edge order out of
//                    shared memory (consecutive lanes read consecutive addresses),
//                    evaluates the edge functor with x/u of its row and column
//                    from the staged vertex table (slot_meta carries the row and
//                    the table index), stores the off-diagonal value straight to
//                    the CSR array (consecutive lanes -> consecutive slots:
//                    coalesced) and parks the slot's diagonal / rhs share in
//                    shared memory.  A named barrier over the math warps follows.
//                    Pass B is ROW-parallel: the rows are dealt to the warps 32 at a
//                    time (the first warp rotates with the tile); lane l folds the
//                    parked shares of its row in slot order, adds the vertex term,
//                    applies the Dirichlet row replacement in the same pass and
//                    stores the diagonal value and the rhs / residual entry.
// Row-partitioned runs (fvm_assemble_halo): tiles are walked interior-first and the
// producer acquires the neighbours' epoch flags (ld.acquire.sys) only before the u copies
// of a tile whose vertex table reaches into ghost vertices -- the halo wait is part
// of this kernel.
// Deterministic and atomic-free: a slot is the in-order sum of its half-edges, a
// row the in-order sum of its slots -- never a function of tile boundaries, the
// warp count or the grid size, so a row-partitioned multi-GPU run is bit-identical
// to the single-GPU run.

#define FVM_EDGE_USES_EL ((FVM_FLAGS & FVM_F_EDGE_EL) != 0)
#define FVM_EDGE_MASKED ((FVM_FLAGS & FVM_F_EDGE_MASK) != 0)
#define FVM_STAGE_X ((FVM_FLAGS & FVM_F_X) != 0)
#define FVM_STAGE_U ((FVM_FLAGS & FVM_F_U) != 0)
#define FVM_STAGE_CV ((FVM_FLAGS & FVM_F_CV) != 0)
#define FVM_STAGE_SA ((FVM_FLAGS & FVM_F_SA) != 0)
#define FVM_VERT_MASKED ((FVM_FLAGS & FVM_F_VMASK) != 0)
#define FVM_HAS_DIRICHLET ((FVM_FLAGS & FVM_F_DIR) != 0)

#define FVM_ACC_D ((FVM_FLAGS & FVM_F_EDGE_D) != 0)
#define FVM_ACC_C ((FVM_FLAGS & FVM_F_EDGE_C) != 0)

#define FVM_XS (FVM_DIM == 2 ? 2 : 4)       // doubles per vertex in the coordinate arrays
#ifndef FVM_PRODUCERS
#define FVM_PRODUCERS 1  // producer warps (one elected lane each); 1, 2 or 4
#endif
#define FVM_NCT (32 * FVM_CONSUMER_WARPS)            // consumer threads
#define FVM_THREADS (FVM_NCT + 32 * FVM_PRODUCERS)   // + the producer warps
// experiment: register reallocation between the warpgroups (setmaxnreg; the producer
// warpgroup keeps 40 registers, the math warpgroup takes 88).  ptxas fits the math path
// into the launch's 64 registers anyway, so it is off by default.
#ifndef FVM_SETMAXNREG
#define FVM_SETMAXNREG 0
#endif
#define FVM_FULL 0xffffffffu
#ifndef FVM_ROW_RUNS
// pass B row dealing: 0 = 32-row blocks round-robin over the warps, 1 = equal runs of rows
// per warp, 2 = decided per tile (runs when the tile has >= 64 rows: ~104-row triangle
// tiles give every warp 26 rows; ~17-row tetrahedral tiles keep one warp's lanes busy)
#define FVM_ROW_RUNS 2
#endif
#ifndef FVM_ROW_WARPS
// 0: every math warp runs the slot pass, a named barrier, then the row pass of the tile.
// 1: the LAST math warp only runs row passes and the others only slot passes: the slot
// warps hand a tile's parked shares over through an mbarrier and move on to the next tile
// (no barrier stall, the row pass of tile i overlaps the slot pass of tile i + 1).
#define FVM_ROW_WARPS 0
#endif
#define FVM_SLOT_WARPS (FVM_CONSUMER_WARPS - FVM_ROW_WARPS)
#define FVM_NST (32 * FVM_SLOT_WARPS)  // threads of the slot pass
#ifndef FVM_L2_AHEAD
#define FVM_L2_AHEAD 0  // experiment (measured: hurts): tiles ahead whose streams are prefetched into L2
#endif
#ifndef FVM_ILP
#define FVM_ILP 1  // 32-slot chunks a warp keeps in flight together in the slot pass (2: measured, no gain)
#endif
#ifndef FVM_ROW_AHEAD
#define FVM_ROW_AHEAD 6  // parked shares of a row loaded before the in-order adds start
#endif
#ifndef FVM_FOLD_AHEAD
// pairs of a multi-pair slot whose loads are issued before the in-order adds start
// (tetrahedra: 4 or 6 pairs per slot; the loads are independent, the adds are not)
#define FVM_FOLD_AHEAD 0  // (measured on 200^3: 0 beats 3 and 5 -- the selects cost more)
#endif

static_assert(!(FVM_EDGE_USES_X || FVM_VERT_USES_X) || FVM_STAGE_X, "FVM_F_X missing");
static_assert(!(FVM_EDGE_USES_U || FVM_VERT_USES_U) || FVM_STAGE_U, "FVM_F_U missing");
static_assert(FVM_MODE != FVM_MODE_RESIDUAL || !FVM_ACC_D, "residual mode has no D output");
static_assert(FVM_MODE != FVM_MODE_JACOBIAN || !FVM_ACC_C, "Jacobian mode has no C output");
static_assert(FVM_MODE != FVM_MODE_SPMV || (FVM_ACC_C && !FVM_ACC_D && FVM_STAGE_U),
              "SpMV mode parks one share per slot and gathers x");
static_assert(!(FVM_EDGE_FACTORED && (FVM_EDGE_USES_EL || FVM_EDGE_MASKED)),
              "ce factoring needs a slot-invariant functor");

__device__ __forceinline__ unsigned fvm_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void fvm_mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fvm_smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fvm_mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fvm_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void fvm_mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fvm_smem_u32(bar)) : "memory");
}

// TMA 1-D bulk copy global -> shared (SASS: UBLKCP); 16B-aligned, size % 16 == 0
__device__ __forceinline__ void fvm_bulk_g2s(void* dst, const void* src, unsigned bytes,
                                             void* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(fvm_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(fvm_smem_u32(bar))
      : "memory");
}

// try_wait suspends the warp in hardware for up to the time hint (ns) before it
// reports failure, so a waiting warp does not burn issue slots
__device__ __forceinline__ void fvm_mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FVM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra FVM_DONE;\n"
      "bra FVM_WAIT;\n"
      "FVM_DONE:\n"
      "}\n" ::"r"(fvm_smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}

// One staged stream: `nbytes` bytes at global address g (any alignment) land in
// a 16B-aligned shared buffer at byte offset `shift`.
struct FvmStage {
  const unsigned char* src;  // 16B-aligned global start
  unsigned shift;            // offset of the first element inside the buffer
  unsigned bytes;            // copy size, multiple of 16
};

__device__ __forceinline__ FvmStage fvm_stage(const void* g, unsigned nbytes) {
  FvmStage s;
  const unsigned long long a = (unsigned long long)g;
  s.shift = (unsigned)(a & 15ull);
  s.src = (const unsigned char*)(a - s.shift);
  s.bytes = nbytes ? fvm_round16(s.shift + nbytes) : 0u;
  return s;
}

// Byte offset of element `index` (elements of `elem_bytes`) inside its 16-byte line.
// Every base pointer of FvmTileArgs is 16-byte aligned (the host checks), so the math
// warps derive a stream's shift from the index alone -- no 64-bit pointer arithmetic.
__device__ __forceinline__ unsigned fvm_shift_at(long long index, unsigned elem_bytes) {
  return ((unsigned)index * elem_bytes) & 15u;
}

// Fused halo wait: spin (with back-off) until every neighbour has released this epoch
// into its flag word -- the ghost values it pushed over NVLink are then in our HBM --
// and order the bulk copies (async proxy) after the acquire.  Gives up after ~20 s.
__device__ __noinline__ void fvm_halo_acquire(const FvmTileArgs& a) {
  unsigned long long t0, epoch = a.halo_epoch;
  if (a.halo_epoch_ptr)
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(epoch) : "l"(a.halo_epoch_ptr) : "memory");
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < a.n_halo_flags; ++i) {
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.halo_flags[i]) : "memory");
      if (v >= epoch) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) {
        *a.halo_timeout = 1;
        break;
      }
      __nanosleep(100);
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");
}

// L2 prefetch of a contiguous stream (cp.async.bulk.prefetch.L2): the DRAM -> L2 leg of a
// tile's streams is started FVM_L2_AHEAD tiles before the shared-memory fill, so the fill
// of the 2-stage ring waits for an L2 hit, not for DRAM under load.
__device__ __forceinline__ void fvm_l2_prefetch(const void* g, unsigned nbytes) {
  const unsigned long long a = (unsigned long long)g;
  const unsigned shift = (unsigned)(a & 15ull);
  const unsigned bytes = fvm_round16(shift + nbytes);
  if (nbytes)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a - shift), "r"(bytes) : "memory");
}

// The stage fill is split into four ROLES so that several producer warps can issue it in
// parallel (the serial issue time of ~20 bulk copies per tile -- 8 halo ranges for each of
// the two vertex tables on tetrahedra -- is on the critical path of the 2-stage ring):
//   role 0 (HE)  : ce_ratio / edge_length / mask per half-edge (SpMV: CSR values), slot_meta
//   role 1 (ROW) : the tile descriptor, row pointers, cv, Dirichlet ids, diagonal offsets, vmask
//   role 2 (XY)  : vertex table of coordinates (own rows + halo ranges)
//   role 3 (U)   : vertex table of u (own rows + halo ranges), after the fused halo wait
// Producer warp p of P takes the roles r with r % P == p; it adds the bytes of its roles
// to the stage's full barrier with ONE arrive.expect_tx (the barrier counts P arrivals).
// `rd` points at the descriptor in the warp's own descriptor ring (shared memory).
#define FVM_ROLE_MINE(r, p) (((r) % FVM_PRODUCERS) == (p))

template <int P_ID>
__device__ __forceinline__ void fvm_produce(const FvmTileArgs& a, const FvmSmemLayout& L,
                                            unsigned char* stage, void* full,
                                            const FvmTileDesc* rd, bool wait_empty, void* empty,
                                            unsigned empty_parity, bool& halo_ready) {
  const int d_r0 = rd->r0, d_nr = rd->nr, d_s0 = rd->s0, d_nsl = rd->nsl, d_nhe = rd->nhe;
  const long long d_h0 = rd->h0;
  const int d_own_start = rd->own_start, d_own_count = rd->own_count, nhalo = rd->nhalo;
  const int v0 = a.row_base + d_r0;
  unsigned total = 0;
  // ---- sizes first (nothing below touches the stage before the empty wait)
#if FVM_MODE != FVM_MODE_SPMV
  const FvmStage g_ce = fvm_stage(a.he_ce + d_h0, (unsigned)d_nhe * 8u);
#else
  const FvmStage g_ce = fvm_stage(a.mat + d_s0, (unsigned)d_nsl * 8u);  // CSR values instead
#endif
  const FvmStage g_meta = fvm_stage(a.slot_meta + (d_s0 - d_r0), (unsigned)(d_nsl - d_nr) * 4u);
  const FvmStage g_rp = fvm_stage(a.indptr + d_r0, (unsigned)(d_nr + 1) * 4u);
  if (FVM_ROLE_MINE(0, P_ID)) total += g_ce.bytes + g_meta.bytes;
  if (FVM_ROLE_MINE(1, P_ID)) total += g_rp.bytes;
  // vertex table: own rows, then the halo ranges (even starts, even counts)
  unsigned n_tab = (unsigned)d_own_count;
#pragma unroll 1
  for (int i = 0; i < nhalo; ++i) n_tab += rd->hcount[i];
#if FVM_EDGE_USES_EL
  const FvmStage g_el = fvm_stage(a.he_el + d_h0, (unsigned)d_nhe * 8u);
  if (FVM_ROLE_MINE(0, P_ID)) total += g_el.bytes;
#endif
#if FVM_EDGE_MASKED
  const FvmStage g_mask = fvm_stage(a.he_mask + d_h0, (unsigned)d_nhe);
  if (FVM_ROLE_MINE(0, P_ID)) total += g_mask.bytes;
#endif
#if FVM_STAGE_CV
  const FvmStage g_cv = fvm_stage(a.cv + v0, (unsigned)d_nr * 8u);
  if (FVM_ROLE_MINE(1, P_ID)) total += g_cv.bytes;
#endif
#if FVM_STAGE_SA
  const FvmStage g_sa = fvm_stage(a.sa + v0, (unsigned)d_nr * 8u);
  if (FVM_ROLE_MINE(1, P_ID)) total += g_sa.bytes;
#endif
#if FVM_STAGE_X
  if (FVM_ROLE_MINE(2, P_ID)) total += n_tab * 8u * FVM_XS;
#endif
#if FVM_STAGE_U
  if (FVM_ROLE_MINE(3, P_ID)) total += n_tab * 8u;
#endif
#if FVM_HAS_DIRICHLET
  const FvmStage g_dir = fvm_stage(a.dir_id + v0, (unsigned)d_nr);
  if (FVM_ROLE_MINE(1, P_ID)) total += g_dir.bytes;
#endif
#if FVM_MODE != FVM_MODE_RESIDUAL
  const FvmStage g_dg = fvm_stage(a.row_diag + d_r0, (unsigned)d_nr * 2u);
  if (FVM_ROLE_MINE(1, P_ID)) total += g_dg.bytes;
#endif
#if FVM_VERT_MASKED
  const FvmStage g_vm = fvm_stage(a.vmask + v0, (unsigned)d_nr);
  if (FVM_ROLE_MINE(1, P_ID)) total += g_vm.bytes;
#endif
#if FVM_STAGE_U
  if (FVM_ROLE_MINE(3, P_ID) && a.n_halo_flags > 0 && !halo_ready) {
    // does this tile's vertex table reach into the ghost vertices?
    int tmin = d_own_start, tmax = d_own_start + d_own_count;
    if (nhalo > 0) {
      tmin = min(tmin, rd->hstart[0]);
      tmax = max(tmax, rd->hstart[nhalo - 1] + (int)rd->hcount[nhalo - 1]);
    }
    // (a tile with columns outside its table gathers them from global memory: any of
    // them may be a ghost vertex)
    if (tmin < a.ghost_lo_end || tmax > a.ghost_hi_begin || rd->uncovered) {
      fvm_halo_acquire(a);
      halo_ready = true;
    }
  }
#endif
  if (wait_empty) fvm_mbar_wait(empty, empty_parity);  // the math warps released the stage
  if (FVM_ROLE_MINE(1, P_ID)) {
    const int4* src = (const int4*)rd;
    int4* dst = (int4*)(stage + L.desc);
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = src[i];  // .. uncovered; the padding is not read
  }
  fvm_mbar_expect_tx(full, total);  // release: the descriptor store above is visible
  if (FVM_ROLE_MINE(0, P_ID)) {
#if FVM_MODE != FVM_MODE_SPMV
    if (g_ce.bytes) fvm_bulk_g2s(stage + L.ce, g_ce.src, g_ce.bytes, full);
#else
    if (g_ce.bytes) fvm_bulk_g2s(stage + L.mat, g_ce.src, g_ce.bytes, full);
#endif
    if (g_meta.bytes) fvm_bulk_g2s(stage + L.meta, g_meta.src, g_meta.bytes, full);
#if FVM_EDGE_USES_EL
    if (g_el.bytes) fvm_bulk_g2s(stage + L.el, g_el.src, g_el.bytes, full);
#endif
#if FVM_EDGE_MASKED
    if (g_mask.bytes) fvm_bulk_g2s(stage + L.mask, g_mask.src, g_mask.bytes, full);
#endif
  }
  if (FVM_ROLE_MINE(1, P_ID)) {
    fvm_bulk_g2s(stage + L.rowptr, g_rp.src, g_rp.bytes, full);
#if FVM_STAGE_CV
    if (g_cv.bytes) fvm_bulk_g2s(stage + L.cv, g_cv.src, g_cv.bytes, full);
#endif
#if FVM_STAGE_SA
    if (g_sa.bytes) fvm_bulk_g2s(stage + L.sa, g_sa.src, g_sa.bytes, full);
#endif
#if FVM_HAS_DIRICHLET
    if (g_dir.bytes) fvm_bulk_g2s(stage + L.dir, g_dir.src, g_dir.bytes, full);
#endif
#if FVM_MODE != FVM_MODE_RESIDUAL
    if (g_dg.bytes) fvm_bulk_g2s(stage + L.dg, g_dg.src, g_dg.bytes, full);
#endif
#if FVM_VERT_MASKED
    if (g_vm.bytes) fvm_bulk_g2s(stage + L.vmask, g_vm.src, g_vm.bytes, full);
#endif
  }
#if FVM_STAGE_X
  if (FVM_ROLE_MINE(2, P_ID)) {
    unsigned char* dst = stage + L.xy;
    const unsigned vb = 8u * FVM_XS;  // bytes per vertex
    if (d_own_count)
      fvm_bulk_g2s(dst, a.coords + (size_t)d_own_start * FVM_XS, (unsigned)d_own_count * vb, full);
    dst += (unsigned)d_own_count * vb;
#pragma unroll 1
    for (int i = 0; i < nhalo; ++i) {
      const unsigned cnt = rd->hcount[i];
      fvm_bulk_g2s(dst, a.coords + (size_t)rd->hstart[i] * FVM_XS, cnt * vb, full);
      dst += cnt * vb;
    }
  }
#endif
#if FVM_STAGE_U
  if (FVM_ROLE_MINE(3, P_ID)) {
    unsigned char* dst = stage + L.u;
    if (d_own_count) fvm_bulk_g2s(dst, a.u + d_own_start, (unsigned)d_own_count * 8u, full);
    dst += (unsigned)d_own_count * 8u;
#pragma unroll 1
    for (int i = 0; i < nhalo; ++i) {
      const unsigned cnt = rd->hcount[i];
      fvm_bulk_g2s(dst, a.u + rd->hstart[i], cnt * 8u, full);
      dst += cnt * 8u;
    }
  }
#endif
}

template <int P_ID>
__device__ __forceinline__ void fvm_prefetch_tile(const FvmTileArgs& a, const FvmTileDesc* rd) {
  const int d_r0 = rd->r0, d_nr = rd->nr, d_s0 = rd->s0, d_nsl = rd->nsl, d_nhe = rd->nhe;
  const long long d_h0 = rd->h0;
  const int v0 = a.row_base + d_r0;
  if (FVM_ROLE_MINE(0, P_ID)) {
#if FVM_MODE != FVM_MODE_SPMV
    fvm_l2_prefetch(a.he_ce + d_h0, (unsigned)d_nhe * 8u);
#else
    fvm_l2_prefetch(a.mat + d_s0, (unsigned)d_nsl * 8u);
#endif
    fvm_l2_prefetch(a.slot_meta + (d_s0 - d_r0), (unsigned)(d_nsl - d_nr) * 4u);
#if FVM_EDGE_USES_EL
    fvm_l2_prefetch(a.he_el + d_h0, (unsigned)d_nhe * 8u);
#endif
#if FVM_EDGE_MASKED
    fvm_l2_prefetch(a.he_mask + d_h0, (unsigned)d_nhe);
#endif
  }
  if (FVM_ROLE_MINE(1, P_ID)) {
    fvm_l2_prefetch(a.indptr + d_r0, (unsigned)(d_nr + 1) * 4u);
#if FVM_STAGE_CV
    fvm_l2_prefetch(a.cv + v0, (unsigned)d_nr * 8u);
#endif
#if FVM_HAS_DIRICHLET
    fvm_l2_prefetch(a.dir_id + v0, (unsigned)d_nr);
#endif
#if FVM_MODE != FVM_MODE_RESIDUAL
    fvm_l2_prefetch(a.row_diag + d_r0, (unsigned)d_nr * 2u);
#endif
#if FVM_VERT_MASKED
    fvm_l2_prefetch(a.vmask + v0, (unsigned)d_nr);
#endif
  }
  // vertex tables: the tile's own rows and its LAST halo range (the highest vertices: on a
  // lexicographic numbering the only part no earlier tile has pulled into L2 yet)
#if FVM_STAGE_X
  if (FVM_ROLE_MINE(2, P_ID)) {
    fvm_l2_prefetch(a.coords + (size_t)rd->own_start * FVM_XS, (unsigned)rd->own_count * 8u * FVM_XS);
    const int nh = rd->nhalo;
    if (nh > 0)
      fvm_l2_prefetch(a.coords + (size_t)rd->hstart[nh - 1] * FVM_XS,
                      (unsigned)rd->hcount[nh - 1] * 8u * FVM_XS);
  }
#endif
#if FVM_STAGE_U
  if (FVM_ROLE_MINE(3, P_ID) && a.n_halo_flags == 0) {  // (ghost values may still be in flight)
    fvm_l2_prefetch(a.u + rd->own_start, (unsigned)rd->own_count * 8u);
    const int nh = rd->nhalo;
    if (nh > 0) fvm_l2_prefetch(a.u + rd->hstart[nh - 1], (unsigned)rd->hcount[nh - 1] * 8u);
  }
#endif
}

// One producer warp: walks the CTA's tiles, keeps FVM_NDESC tile descriptors in flight
// in its own shared-memory ring (fetched by bulk copies, so the ~1 us DRAM latency of a
// descriptor is never waited for) and fills its roles of every stage.
template <int P_ID>
__device__ __forceinline__ void fvm_producer_warp(const FvmTileArgs& a, const FvmSmemLayout& L,
                                                  unsigned char* sm, unsigned long long* full,
                                                  unsigned long long* empty) {
  unsigned char* ring = sm + L.ring + P_ID * (FVM_NDESC * 128);
  unsigned long long* dbar = (unsigned long long*)(sm + L.dbar) + P_ID * FVM_NDESC;
  // tile t of the walk is tile (t + tile_rot) mod ntiles of the mesh
  auto rot = [&](int i) { return i + a.tile_rot < a.ntiles ? i + a.tile_rot : i + a.tile_rot - a.ntiles; };
  auto fetch = [&](int slot, int tile) {
    fvm_mbar_expect_tx(dbar + slot, 128u);
    fvm_bulk_g2s(ring + slot * 128, a.tiles + rot(tile), 128u, dbar + slot);
  };
#pragma unroll
  for (int j = 0; j < FVM_NDESC; ++j) {
    const long long tj = (long long)blockIdx.x + (long long)j * gridDim.x;
    if (tj < a.ntiles) fetch(j, (int)tj);
  }
  bool halo_ready = false;
  int it = 0;
  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
    const int ds = it % FVM_NDESC, st = it % FVM_STAGES;
    fvm_mbar_wait(dbar + ds, (it / FVM_NDESC) & 1);
    fvm_produce<P_ID>(a, L, sm + L.stage0 + st * L.stage_bytes, full + st,
                      (const FvmTileDesc*)(ring + ds * 128), it >= FVM_STAGES, empty + st,
                      ((it / FVM_STAGES) - 1) & 1, halo_ready);
    // every read of the ring slot above fed an instruction already issued: refill it
    const long long tn = (long long)t + (long long)FVM_NDESC * gridDim.x;
    if (tn < a.ntiles) fetch(ds, (int)tn);
#if FVM_L2_AHEAD > 0
    // start the DRAM -> L2 leg of the tile FVM_L2_AHEAD steps ahead (its descriptor was
    // fetched FVM_NDESC - FVM_L2_AHEAD steps ago)
    if ((long long)t + (long long)FVM_L2_AHEAD * gridDim.x < a.ntiles) {
      const int ps = (it + FVM_L2_AHEAD) % FVM_NDESC;
      fvm_mbar_wait(dbar + ps, ((it + FVM_L2_AHEAD) / FVM_NDESC) & 1);
      fvm_prefetch_tile<P_ID>(a, (const FvmTileDesc*)(ring + ps * 128));
    }
#endif
  }
}

// Columns a tile's halo ranges do not cover (irregular numbering) are gathered
// from global memory.  Kept out of line so the hot loop carries a branch, not
// the predicated body.
__device__ __noinline__ double4 fvm_gather_column(const int* colidx, const double* coords,
                                                  const double* u, int slot) {
  double4 r = make_double4(0.0, 0.0, 0.0, 0.0);  // x1_0, x1_1, x1_2, u1
  const int col = __ldg(colidx + slot);
#if FVM_EDGE_USES_X
  const double* p1 = coords + (size_t)col * FVM_XS;
  const double2 q = __ldg((const double2*)p1);
  r.x = q.x;
  r.y = q.y;
#if FVM_DIM == 3
  r.z = __ldg(p1 + 2);
#endif
#endif
#if FVM_EDGE_USES_U
  // u may hold ghost values a peer GPU stored during this launch (ordered by the
  // producer's flag acquire + the stage barrier): a coherent load, never ld.global.nc
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(r.w) : "l"(u + col) : "memory");
#endif
  return r;
}

// 4 CTAs per SM is the resident count the shared-memory footprint allows on the BASELINE
// configurations; the register allocation is told so (measured: ptxas' unconstrained choice
// of 56 registers spills in the slot pass and is 3 % slower).  FVM_MIN_CTAS > 0 overrides.
#if defined(FVM_MIN_CTAS) && FVM_MIN_CTAS > 0
extern "C" __global__ void __launch_bounds__(FVM_THREADS, FVM_MIN_CTAS)
#else
extern "C" __global__ void __launch_bounds__(FVM_THREADS, 4)
#endif
    fvm_tile_kernel(const FvmTileArgs a) {
  extern __shared__ __align__(16) unsigned char sm[];
  const FvmSmemLayout L = a.L;  // computed by the host (fvm_smem_layout)
  unsigned long long* full = (unsigned long long*)(sm + L.bar);
  unsigned long long* empty = full + FVM_STAGES;
  unsigned long long* parked = empty + FVM_STAGES;  // slot warps -> row warp (FVM_ROW_WARPS)

  if (threadIdx.x == 0) {
    for (int i = 0; i < FVM_STAGES; ++i) {
      fvm_mbar_init(full + i, FVM_PRODUCERS);        // every producer warp's expect_tx arrival
      fvm_mbar_init(empty + i, FVM_CONSUMER_WARPS);  // one arrival per consumer warp
      fvm_mbar_init(parked + i, FVM_SLOT_WARPS);     // one arrival per slot warp
    }
    for (int i = 0; i < FVM_PRODUCERS * FVM_NDESC; ++i)
      fvm_mbar_init((unsigned long long*)(sm + L.dbar) + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x < 32 * FVM_PRODUCERS) {
    // ===================== producer warps (one elected lane each) =============
#if FVM_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
#endif
    if ((threadIdx.x & 31) == 0) {
      const int p = threadIdx.x >> 5;
      if (p == 0) fvm_producer_warp<0>(a, L, sm, full, empty);
#if FVM_PRODUCERS > 1
      else if (p == 1) fvm_producer_warp<1>(a, L, sm, full, empty);
#endif
#if FVM_PRODUCERS > 2
      else if (p == 2) fvm_producer_warp<2>(a, L, sm, full, empty);
      else fvm_producer_warp<3>(a, L, sm, full, empty);
#endif
    }
    return;
  }
#if FVM_SETMAXNREG
  asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
#endif

  // ========================= consumer warps ==================================
  const int lane = threadIdx.x & 31;
  const int cw = (threadIdx.x >> 5) - FVM_PRODUCERS;  // consumer warp id

  int it = 0;
  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
    const int st = it % FVM_STAGES;
    unsigned char* stage = sm + L.stage0 + st * L.stage_bytes;
    fvm_mbar_wait(full + st, (it / FVM_STAGES) & 1);
    const FvmTileDesc* sd = (const FvmTileDesc*)(stage + L.desc);
    const int d_nr = sd->nr, d_s0 = sd->s0, d_nsl = sd->nsl, d_nhe = sd->nhe, d_r0 = sd->r0;
    const int v0 = a.row_base + d_r0;  // first vertex of the tile
    // pass B row dealing (FVM_ROW_RUNS): blocks of 32 rows (lanes fill first: a 17-row
    // tetrahedral tile keeps one warp busy instead of four warps with 4 lanes each; the warp
    // that takes the first block rotates with the tile) or equal runs of rows per warp
#if FVM_ROW_WARPS > 0
    const bool row_role = cw >= FVM_SLOT_WARPS;  // this warp folds rows only
    const int rw = cw - FVM_SLOT_WARPS;
    const int g_first = (d_nr * rw) / FVM_ROW_WARPS, g_end = (d_nr * (rw + 1)) / FVM_ROW_WARPS;
    const int g_step = 32;
#else
    int g_first = ((cw + it) % FVM_CONSUMER_WARPS) * 32, g_end = d_nr, g_step = FVM_NCT;
    if (FVM_ROW_RUNS == 1 || (FVM_ROW_RUNS == 2 && d_nr >= 64)) {
      g_first = (d_nr * cw) / FVM_CONSUMER_WARPS;
      g_end = (d_nr * (cw + 1)) / FVM_CONSUMER_WARPS;
      g_step = 32;
    }
#endif

#if FVM_MODE != FVM_MODE_SPMV
    // a tile starts at a slot start: even position of the padded stream, 16-byte aligned
    const double2* s_ce2 = (const double2*)(stage + L.ce);
#else
    const double* s_mat = (const double*)(stage + L.mat + fvm_shift_at(d_s0, 8u));
#endif
    const int d_nk = d_nsl - d_nr;  // off-diagonal slots of the tile
    const unsigned* s_meta =
        (const unsigned*)(stage + L.meta + fvm_shift_at(d_s0 - d_r0, 4u));
    const int xoff = v0 - sd->own_start;  // table index of the tile's first row
    const int* s_rp = (const int*)(stage + L.rowptr + fvm_shift_at(d_r0, 4u));
#if FVM_EDGE_USES_EL
    const double* s_el = (const double*)(stage + L.el);
#endif
#if FVM_EDGE_MASKED
    const unsigned char* s_mask = stage + L.mask + fvm_shift_at(sd->h0, 1u);
#endif
#if FVM_STAGE_CV
    const double* s_cv = (const double*)(stage + L.cv + fvm_shift_at(v0, 8u));
#endif
#if FVM_STAGE_SA
    const double* s_sa = (const double*)(stage + L.sa + fvm_shift_at(v0, 8u));
#endif
#if FVM_STAGE_X
    const double* s_xy = (const double*)(stage + L.xy);
#endif
#if FVM_STAGE_U
    const double* s_u = (const double*)(stage + L.u);
#endif
#if FVM_HAS_DIRICHLET
    const unsigned char* s_dir = stage + L.dir + fvm_shift_at(v0, 1u);
#endif
#if FVM_VERT_MASKED
    const unsigned char* s_vm = stage + L.vmask + fvm_shift_at(v0, 1u);
#endif
#if FVM_ACC_D && FVM_MODE != FVM_MODE_SPMV
    double* s_accd = (double*)(stage + L.accd);  // lives and dies with the stage
#endif
#if FVM_ACC_C && FVM_MODE != FVM_MODE_SPMV
    double* s_accc = (double*)(stage + L.accc);
#endif
#if FVM_MODE != FVM_MODE_RESIDUAL
    const unsigned short* s_dg =
        (const unsigned short*)(stage + L.dg + fvm_shift_at(d_r0, 2u));
#endif

    {
      // ---------------- pass A: one lane per off-diagonal CSR slot -------------
      // 32-slot chunks; a warp takes FVM_ILP consecutive chunks per trip and keeps their
      // (independent) load -> functor chains in flight together: with 4 math warps per
      // scheduler the pass is bound by the latency of one chain, not by issue slots.
      // Lanes past the tile's last slot run on an all-zero metadata word (first pair, row 0,
      // table entry 0: valid addresses, one pair) and store nothing.
#if FVM_ROW_WARPS > 0
      if (!row_role)
#endif
#pragma unroll 1
      for (int kb = cw * (32 * FVM_ILP) + lane; kb < d_nk + lane; kb += FVM_NST * FVM_ILP) {
        unsigned meta[FVM_ILP];
        int kk[FVM_ILP];
        bool act[FVM_ILP];
#pragma unroll
        for (int j = 0; j < FVM_ILP; ++j) {
          kk[j] = kb + 32 * j;
          act[j] = kk[j] < d_nk;
          meta[j] = act[j] ? s_meta[kk[j]] : 0u;  // 0: pair 0, one pair, row 0, table entry 0
        }
        int g[FVM_ILP], sl[FVM_ILP];
        double x00[FVM_ILP], x01[FVM_ILP], x02[FVM_ILP], u0[FVM_ILP];
        double x10[FVM_ILP], x11[FVM_ILP], x12[FVM_ILP], u1[FVM_ILP];
        double ce[FVM_ILP];
#if FVM_MODE != FVM_MODE_SPMV
        int lo2[FVM_ILP], n2[FVM_ILP];
#endif
        // ---- loads of all chains first
#pragma unroll
        for (int j = 0; j < FVM_ILP; ++j) {
          const unsigned m = meta[j];
          g[j] = (int)((m >> FVM_META_ROW_SHIFT) & FVM_META_ROW_MASK);  // row of the slot
          // CSR slot (tile-relative): one diagonal slot per earlier row, plus this row's
          sl[j] = kk[j] + g[j] + (int)((m >> FVM_META_AFTER_BIT) & 1u);
          x00[j] = x01[j] = x02[j] = u0[j] = 0.0;
          x10[j] = x11[j] = x12[j] = u1[j] = 0.0;
          ce[j] = 0.0;
#if FVM_MODE != FVM_MODE_SPMV
          lo2[j] = (int)(m & FVM_META_PAIR_MASK);  // first pair of the slot
          n2[j] = 1;                               // pairs of the slot
          if ((m >> FVM_META_MULTI_BIT) & 1u)
            n2[j] = ((kk[j] + 1 < d_nk) ? (int)(s_meta[kk[j] + 1] & FVM_META_PAIR_MASK)
                                        : (d_nhe >> 1)) - lo2[j];
#if FVM_EDGE_FACTORED
          {
            const double2 p = s_ce2[lo2[j]];
            ce[j] = p.x + p.y;  // (the first pair of the fold; padding entries are +0.0)
          }
#endif
#endif
#if FVM_EDGE_USES_X
          {
            const double2 q0 = *(const double2*)(s_xy + FVM_XS * (xoff + g[j]));
            x00[j] = q0.x;
            x01[j] = q0.y;
#if FVM_DIM == 3
            x02[j] = s_xy[FVM_XS * (xoff + g[j]) + 2];
#endif
          }
#endif
#if FVM_EDGE_USES_U
          u0[j] = s_u[xoff + g[j]];
#endif
#if FVM_EDGE_USES_X || FVM_EDGE_USES_U
          const unsigned xl = m >> FVM_META_XL_SHIFT;
          if (xl != FVM_XLOC_NONE) {
#if FVM_EDGE_USES_X
            const double2 q = *(const double2*)(s_xy + FVM_XS * xl);
            x10[j] = q.x;
            x11[j] = q.y;
#if FVM_DIM == 3
            x12[j] = s_xy[FVM_XS * xl + 2];
#endif
#endif
#if FVM_EDGE_USES_U
            u1[j] = s_u[xl];
#endif
          } else {  // column outside the staged halo ranges: gather it (cold path)
            const double4 gq = fvm_gather_column(a.colidx, a.coords, a.u, d_s0 + sl[j]);
            x10[j] = gq.x;
            x11[j] = gq.y;
            x12[j] = gq.z;
            u1[j] = gq.w;
          }
#endif
        }
        // ---- functor + stores
#pragma unroll
        for (int j = 0; j < FVM_ILP; ++j) {
          double aD = 0.0, aO = 0.0, aC = 0.0;
#if FVM_MODE == FVM_MODE_SPMV
          aC = s_mat[sl[j]] * u1[j];
#elif FVM_EDGE_FACTORED
          // every edge term is ce * g(slot): fold the slot's ce_ratios in half-edge order
          // (aligned 16-byte loads), evaluate g once
          if (n2[j] > 1) {  // triangles: irregular slots only; tetrahedra: always (4 or 6)
            const double2* cp = s_ce2 + lo2[j];
            int i = 1;
#if FVM_FOLD_AHEAD > 1
            {
              double2 q[FVM_FOLD_AHEAD];
#pragma unroll
              for (int t = 0; t < FVM_FOLD_AHEAD; ++t)
                if (1 + t < n2[j]) q[t] = cp[1 + t];
#pragma unroll
              for (int t = 0; t < FVM_FOLD_AHEAD; ++t)
                if (1 + t < n2[j]) {
                  ce[j] += q[t].x;  // still in half-edge order
                  ce[j] += q[t].y;
                }
              i = 1 + FVM_FOLD_AHEAD;
            }
#endif
#pragma unroll 1
            for (; i < n2[j]; ++i) {
              const double2 p = cp[i];
              ce[j] += p.x;
              ce[j] += p.y;
            }
          }
          fvm_edge(u0[j], u1[j], x00[j], x01[j], x02[j], x10[j], x11[j], x12[j], ce[j], 0.0, 0xffu,
                   aD, aO, aC);
#else
          // the edge functor is evaluated per half-edge (edge_length / subdomain bits
          // vary inside a slot); a padding entry has ce = 0 and mask 0: it adds zeros
          {
            const double* s_ce = (const double*)s_ce2;
#pragma unroll 1
            for (int h = 2 * lo2[j]; h < 2 * (lo2[j] + n2[j]); ++h) {
              const double c1 = s_ce[h];
              double el = 0.0;
              unsigned m = 0xffu;
#if FVM_EDGE_USES_EL
              el = s_el[h];
#endif
#if FVM_EDGE_MASKED
              m = s_mask[h];
#endif
              double D = 0.0, O = 0.0, C = 0.0;
              fvm_edge(u0[j], u1[j], x00[j], x01[j], x02[j], x10[j], x11[j], x12[j], c1, el, m, D,
                       O, C);
              aD += D;
              aO += O;
              aC += C;
            }
          }
#endif
          if (act[j]) {
#if FVM_MODE == FVM_MODE_LINEAR || FVM_MODE == FVM_MODE_JACOBIAN
            a.data[d_s0 + sl[j]] = aO;  // consecutive lanes, (nearly) consecutive slots: coalesced
#endif
#if FVM_MODE == FVM_MODE_SPMV
            // SpMV: the product is parked in place, over the slot's own (now dead) CSR value
            ((double*)s_mat)[sl[j]] = aC;
#else
            // parked at k + g: a row's shares stay contiguous and the row stride is the CSR
            // row length (4 + 1 / 8 + 1 on the zigzag triangle grid: odd, so the row pass
            // below reads conflict-free; the compact stride 4 / 8 would 4-way conflict).
            // (Parking in place over the slot's first ce pair was measured: the row pass then
            // gathers through slot_meta with 4-way conflicts, 30 % slower overall.)
#if FVM_ACC_D
            s_accd[kk[j] + g[j]] = aD;
#endif
#if FVM_ACC_C
            s_accc[kk[j] + g[j]] = aC;
#endif
#endif
          }
        }
      }
#if FVM_ROW_WARPS > 0
      if (!row_role) {
        // hand the tile's parked shares (and the CSR stores) over to the row warp
        __syncwarp();
        if (lane == 0) fvm_mbar_arrive(parked + st);
      } else {
        fvm_mbar_wait(parked + st, (it / FVM_STAGES) & 1);
      }
      if (row_role)
#else
      // the parked shares (and pass A's CSR stores) are visible to all math warps
      asm volatile("bar.sync 1, %0;" ::"n"(FVM_NCT) : "memory");
#endif
      // ---------------- pass B: one lane per matrix row ------------------------
      for (int g = g_first + lane; g < g_end; g += g_step) {
        const int b = s_rp[g] - d_s0, e = s_rp[g + 1] - d_s0;  // CSR slots of the row
        double sD = 0.0, sC = 0.0;
#if FVM_MODE == FVM_MODE_SPMV
        // the off-diagonal products sit in the row's own CSR positions (the diagonal value
        // is untouched: it is multiplied in below)
        const int dpos = b + (int)s_dg[g];
        int k = b;
#if FVM_ROW_AHEAD > 0
        {
          double v[FVM_ROW_AHEAD];
#pragma unroll
          for (int t = 0; t < FVM_ROW_AHEAD; ++t) {
            v[t] = 0.0;  // (past the row's end / the diagonal: +0.0, exact)
            if (b + t < e && b + t != dpos) v[t] = s_mat[b + t];
          }
#pragma unroll
          for (int t = 0; t < FVM_ROW_AHEAD; ++t) sC += v[t];
          k = b + FVM_ROW_AHEAD;
        }
#endif
#pragma unroll 1
        for (; k < e; ++k)
          if (k != dpos) sC += s_mat[k];
#elif FVM_ACC_D || FVM_ACC_C
        int k = b;
#if FVM_ROW_AHEAD > 0
        {
          // the first shares of the row are loaded together, then added in slot order (a
          // share past the row's end reads as +0.0: exact); the loop below takes the rest
          double vD[FVM_ROW_AHEAD], vC[FVM_ROW_AHEAD];
#pragma unroll
          for (int t = 0; t < FVM_ROW_AHEAD; ++t) {
            vD[t] = vC[t] = 0.0;
            if (b + t < e - 1) {
#if FVM_ACC_D
              vD[t] = s_accd[b + t];
#endif
#if FVM_ACC_C
              vC[t] = s_accc[b + t];
#endif
            }
          }
#pragma unroll
          for (int t = 0; t < FVM_ROW_AHEAD; ++t) {
            sD += vD[t];
            sC += vC[t];
          }
          k = b + FVM_ROW_AHEAD;
        }
#endif
#pragma unroll 1
        for (; k < e - 1; ++k) {  // its (remaining) off-diagonal slots, in slot order
#if FVM_ACC_D
          sD += s_accd[k];
#endif
#if FVM_ACC_C
          sC += s_accc[k];
#endif
        }
#endif
        double x00 = 0.0, x01 = 0.0, x02 = 0.0, u0 = 0.0, cvv = 0.0, sav = 0.0;
#if FVM_STAGE_X && (FVM_VERT_USES_X)
        {
          const double2 q0 = *(const double2*)(s_xy + FVM_XS * (xoff + g));
          x00 = q0.x;
          x01 = q0.y;
#if FVM_DIM == 3
          x02 = s_xy[FVM_XS * (xoff + g) + 2];
#endif
        }
#endif
#if FVM_STAGE_U && (FVM_VERT_USES_U || FVM_MODE == FVM_MODE_SPMV)
        u0 = s_u[xoff + g];
#endif
#if FVM_STAGE_CV
        cvv = s_cv[g];
#endif
#if FVM_STAGE_SA
        sav = s_sa[g];
#endif
#if FVM_MODE == FVM_MODE_SPMV
        sC += s_mat[dpos] * u0;  // the diagonal entry closes the row
#endif
#if FVM_HAS_VERTEX
        {
          unsigned vm = 0xffu;
#if FVM_VERT_MASKED
          vm = s_vm[g];
#endif
          double VL = 0.0, VC = 0.0;
          fvm_vertex(u0, cvv, sav, x00, x01, x02, vm, VL, VC);
          sD += VL;
          sC += VC;
        }
#endif
#if FVM_MODE == FVM_MODE_LINEAR
        double out_vec = -sC;  // rhs = -(sum of affine parts)
#else
        double out_vec = sC;
#endif
#if FVM_HAS_DIRICHLET
        const int dir = s_dir[g];
        if (dir) {
          double dc = 0.0, dv = 0.0;
          fvm_dirichlet(dir, u0, cvv, sav, x00, x01, x02, dc, dv);
          sD = dc;       // row := coeff * e_row
          out_vec = dv;  // rhs := -affine (linear) / F := condition(u) (residual)
#if FVM_MODE == FVM_MODE_LINEAR || FVM_MODE == FVM_MODE_JACOBIAN
          // Dirichlet rows keep their pattern but store explicit zeros off the
          // diagonal (ordered after pass A's stores by the barrier above)
#pragma unroll 1
          for (int s = b; s < e; ++s) a.data[d_s0 + s] = 0.0;
#endif
        }
#endif
#if FVM_MODE == FVM_MODE_LINEAR || FVM_MODE == FVM_MODE_JACOBIAN
        a.data[d_s0 + b + (int)s_dg[g]] = sD;
#endif
#if FVM_MODE != FVM_MODE_JACOBIAN
        a.vec[d_r0 + g] = out_vec;
#endif
      }
    }
    // ---- release the stage to the producer (all lanes of this warp are done)
#if FVM_MODE == FVM_MODE_SPMV
    // (the math warps wrote into a stream the next bulk copy -- async proxy -- overwrites)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    __syncwarp();
    if (lane == 0) fvm_mbar_arrive(empty + st);
  }
}
