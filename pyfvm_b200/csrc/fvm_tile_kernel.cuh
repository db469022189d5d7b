// Hand-written sm_100a assembly template for the finite-volume hot path.
//
// One template serves the three numeric entry points of the reference:
//   FVM_MODE_LINEAR   : matrix values + rhs of pyfvm.discretize_linear  [README:49]
//   FVM_MODE_RESIDUAL : f.eval(u)                                       [README:121,136]
//   FVM_MODE_JACOBIAN : jacobian.get_linear_operator(u0) values         [README:116]
//
// It is compiled at run time by NVRTC after a generated prelude that defines
//   FVM_MODE, FVM_DIM, FVM_FLAGS (FVM_F_* bits), FVM_CONSUMER_WARPS, FVM_STAGES, FVM_GROUP,
//   FVM_EDGE_USES_X / _U, FVM_VERT_USES_X / _U, FVM_HAS_VERTEX, FVM_EDGE_FACTORED
// and the three sympy-printed device functors
//   fvm_edge(uk0, uk1, x0_0..2, x1_0..2, edge_ce_ratio, edge_length, mask, &D, &O, &C)
//   fvm_vertex(uk0, control_volume, x_0..2, mask, &L, &C)
//   fvm_dirichlet(id, uk0, x_0..2, &coeff, &val)
// (edge functor outputs per half-edge (row, col): D -> slot (row,row),
//  O -> slot (row,col), C -> rhs/residual of row).
//
// Persistent, warp-specialised CTAs (grid = resident CTAs of the whole GPU);
// CTA b walks tiles b, b+grid, ... through a ring of FVM_STAGES shared-memory
// stages guarded by full/empty mbarriers:
//   producer warp  : reads the 32-byte tile descriptor and issues TMA bulk copies
//                    (cp.async.bulk, SASS UBLKCP) of every contiguous stream the
//                    tile needs -- ce_ratio / edge_length / mask per half-edge,
//                    column + offset per slot, row pointer / coords / cv / u /
//                    Dirichlet id per row -- one or more tiles ahead of the math;
//   consumer warps : a group of FVM_GROUP lanes owns one matrix row: lane j folds
//                    the row's CSR slots j, j+G, ... -- each slot's contributions in
//                    half-edge order out of shared memory (deterministic,
//                    atomic-free), gathering x/u only for columns outside the
//                    tile -- and keeps the row's diagonal / rhs sums in registers;
//                    a fixed xor-shuffle tree joins the lanes; lane 0 adds the
//                    vertex term and applies the Dirichlet row replacement in the
//                    same pass.  CSR values leave through a double-buffered
//                    shared staging array (FVM_GROUP < 4) so that the global
//                    stores are coalesced, or directly (FVM_GROUP >= 4).
// The fold order of a slot/row depends only on the row's own half-edge order and
// FVM_GROUP (fixed per cell type), never on tile boundaries or the grid size, so
// a row-partitioned multi-GPU run is bit-identical to the single-GPU run.

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define FVM_EDGE_USES_EL ((FVM_FLAGS & FVM_F_EDGE_EL) != 0)
#define FVM_EDGE_MASKED ((FVM_FLAGS & FVM_F_EDGE_MASK) != 0)
#define FVM_STAGE_X ((FVM_FLAGS & FVM_F_X) != 0)
#define FVM_STAGE_U ((FVM_FLAGS & FVM_F_U) != 0)
#define FVM_STAGE_CV ((FVM_FLAGS & FVM_F_CV) != 0)
#define FVM_VERT_MASKED ((FVM_FLAGS & FVM_F_VMASK) != 0)
#define FVM_HAS_DIRICHLET ((FVM_FLAGS & FVM_F_DIR) != 0)

#define FVM_NCT (32 * FVM_CONSUMER_WARPS)  // consumer threads
#define FVM_STAGE_OUT (FVM_GROUP < 4 && FVM_MODE != FVM_MODE_RESIDUAL)
#define FVM_THREADS (FVM_NCT + 32)         // + one producer warp

static_assert(!(FVM_EDGE_USES_X || FVM_VERT_USES_X) || FVM_STAGE_X, "FVM_F_X missing");
static_assert(!(FVM_EDGE_USES_U || FVM_VERT_USES_U) || FVM_STAGE_U, "FVM_F_U missing");
static_assert(FVM_GROUP >= 1 && FVM_GROUP <= 32 && (FVM_GROUP & (FVM_GROUP - 1)) == 0,
              "FVM_GROUP must be a power of two");
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

__device__ __forceinline__ void fvm_mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FVM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FVM_DONE;\n"
      "bra FVM_WAIT;\n"
      "FVM_DONE:\n"
      "}\n" ::"r"(fvm_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// same, for the producer: back off between polls so the spin does not take
// issue slots from the consumer warps of this SM sub-partition
__device__ __forceinline__ void fvm_mbar_wait_backoff(void* bar, unsigned parity) {
  const unsigned addr = fvm_smem_u32(bar);
  for (;;) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(256);
  }
}

// barrier among the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void fvm_consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(FVM_NCT) : "memory");
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

__device__ __forceinline__ unsigned fvm_shift(const void* g) {
  return (unsigned)((unsigned long long)g & 15ull);
}

__device__ __forceinline__ void fvm_produce(const FvmTileArgs& a, const FvmSmemLayout& L,
                                            unsigned char* stage, void* full,
                                            const FvmTileDesc& d) {
  const int v0 = a.row_base + d.r0;
  *(FvmTileDesc*)(stage + L.desc) = d;
  unsigned total = 0;
  const FvmStage g_ce = fvm_stage(a.he_ce + d.h0, (unsigned)d.nhe * 8u);
  const FvmStage g_col = fvm_stage(a.colidx + d.s0, (unsigned)d.nsl * 4u);
  const FvmStage g_off = fvm_stage(a.slot_off + d.s0, (unsigned)d.nsl * 2u);
  const FvmStage g_rp = fvm_stage(a.indptr + d.r0, (unsigned)(d.nr + 1) * 4u);
  total += g_ce.bytes + g_col.bytes + g_off.bytes + g_rp.bytes;
#if FVM_EDGE_USES_EL
  const FvmStage g_el = fvm_stage(a.he_el + d.h0, (unsigned)d.nhe * 8u);
  total += g_el.bytes;
#endif
#if FVM_EDGE_MASKED
  const FvmStage g_mask = fvm_stage(a.he_mask + d.h0, (unsigned)d.nhe);
  total += g_mask.bytes;
#endif
#if FVM_STAGE_CV
  const FvmStage g_cv = fvm_stage(a.cv + v0, (unsigned)d.nr * 8u);
  total += g_cv.bytes;
#endif
#if FVM_STAGE_X
  const FvmStage g_xy = fvm_stage(a.coords + (size_t)v0 * FVM_DIM, (unsigned)d.nr * 8u * FVM_DIM);
  total += g_xy.bytes;
#endif
#if FVM_STAGE_U
  const FvmStage g_u = fvm_stage(a.u + v0, (unsigned)d.nr * 8u);
  total += g_u.bytes;
#endif
#if FVM_HAS_DIRICHLET
  const FvmStage g_dir = fvm_stage(a.dir_id + v0, (unsigned)d.nr);
  total += g_dir.bytes;
#endif
#if FVM_VERT_MASKED
  const FvmStage g_vm = fvm_stage(a.vmask + v0, (unsigned)d.nr);
  total += g_vm.bytes;
#endif
  fvm_mbar_expect_tx(full, total);  // release: the descriptor store above is visible
  if (g_ce.bytes) fvm_bulk_g2s(stage + L.ce, g_ce.src, g_ce.bytes, full);
  if (g_col.bytes) fvm_bulk_g2s(stage + L.col, g_col.src, g_col.bytes, full);
  if (g_off.bytes) fvm_bulk_g2s(stage + L.off, g_off.src, g_off.bytes, full);
  fvm_bulk_g2s(stage + L.rowptr, g_rp.src, g_rp.bytes, full);
#if FVM_EDGE_USES_EL
  if (g_el.bytes) fvm_bulk_g2s(stage + L.el, g_el.src, g_el.bytes, full);
#endif
#if FVM_EDGE_MASKED
  if (g_mask.bytes) fvm_bulk_g2s(stage + L.mask, g_mask.src, g_mask.bytes, full);
#endif
#if FVM_STAGE_CV
  if (g_cv.bytes) fvm_bulk_g2s(stage + L.cv, g_cv.src, g_cv.bytes, full);
#endif
#if FVM_STAGE_X
  if (g_xy.bytes) fvm_bulk_g2s(stage + L.xy, g_xy.src, g_xy.bytes, full);
#endif
#if FVM_STAGE_U
  if (g_u.bytes) fvm_bulk_g2s(stage + L.u, g_u.src, g_u.bytes, full);
#endif
#if FVM_HAS_DIRICHLET
  if (g_dir.bytes) fvm_bulk_g2s(stage + L.dir, g_dir.src, g_dir.bytes, full);
#endif
#if FVM_VERT_MASKED
  if (g_vm.bytes) fvm_bulk_g2s(stage + L.vmask, g_vm.src, g_vm.bytes, full);
#endif
}

__device__ __forceinline__ FvmTileDesc fvm_load_desc(const FvmTileDesc* g) {
  FvmTileDesc d;
  const int4 d0 = __ldg((const int4*)g), d1 = __ldg((const int4*)g + 1);
  ((int4*)&d)[0] = d0;
  ((int4*)&d)[1] = d1;
  return d;
}

extern "C" __global__ void __launch_bounds__(FVM_THREADS) fvm_tile_kernel(const FvmTileArgs a) {
  extern __shared__ __align__(16) unsigned char sm[];
  const FvmSmemLayout L = a.L;  // computed by the host (fvm_smem_layout)
  unsigned long long* full = (unsigned long long*)(sm + L.bar);
  unsigned long long* empty = full + FVM_STAGES;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FVM_STAGES; ++i) {
      fvm_mbar_init(full + i, 1);                    // the producer's expect_tx arrival
      fvm_mbar_init(empty + i, FVM_CONSUMER_WARPS);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x < 32) {
    // ===================== producer warp (one elected lane) ===================
    if (threadIdx.x == 0) {
      int t = blockIdx.x;
      if (t < a.ntiles) {
        FvmTileDesc d = fvm_load_desc(a.tiles + t);
        for (int it = 0; t < a.ntiles; ++it) {
          const int st = it % FVM_STAGES;
          const int tn = t + gridDim.x;
          FvmTileDesc dn = d;
          if (tn < a.ntiles) dn = fvm_load_desc(a.tiles + tn);  // next descriptor in flight
          if (it >= FVM_STAGES) fvm_mbar_wait_backoff(empty + st, ((it / FVM_STAGES) - 1) & 1);
          fvm_produce(a, L, sm + L.stage0 + st * L.stage_bytes, full + st, d);
          d = dn;
          t = tn;
        }
      }
    }
    return;
  }

  // ========================= consumer warps ==================================
  const int ct = threadIdx.x - 32;     // consumer thread id
  const int gl = ct & (FVM_GROUP - 1);  // lane within the row group
  const int gi = ct / FVM_GROUP;        // row group id
  constexpr int NG = FVM_NCT / FVM_GROUP;

  int it = 0;
  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
    const int st = it % FVM_STAGES;
    unsigned char* stage = sm + L.stage0 + st * L.stage_bytes;
    fvm_mbar_wait(full + st, (it / FVM_STAGES) & 1);
    const FvmTileDesc d = *(const FvmTileDesc*)(stage + L.desc);
    const int v0 = a.row_base + d.r0;  // first vertex of the tile

    const double* s_ce = (const double*)(stage + L.ce + fvm_shift(a.he_ce + d.h0));
    const int* s_col = (const int*)(stage + L.col + fvm_shift(a.colidx + d.s0));
    const unsigned short* s_off =
        (const unsigned short*)(stage + L.off + fvm_shift(a.slot_off + d.s0));
    const int* s_rp = (const int*)(stage + L.rowptr + fvm_shift(a.indptr + d.r0));
#if FVM_EDGE_USES_EL
    const double* s_el = (const double*)(stage + L.el + fvm_shift(a.he_el + d.h0));
#endif
#if FVM_EDGE_MASKED
    const unsigned char* s_mask = stage + L.mask + fvm_shift(a.he_mask + d.h0);
#endif
#if FVM_STAGE_CV
    const double* s_cv = (const double*)(stage + L.cv + fvm_shift(a.cv + v0));
#endif
#if FVM_STAGE_X
    const double* s_xy = (const double*)(stage + L.xy + fvm_shift(a.coords + (size_t)v0 * FVM_DIM));
#endif
#if FVM_STAGE_U
    const double* s_u = (const double*)(stage + L.u + fvm_shift(a.u + v0));
#endif
#if FVM_HAS_DIRICHLET
    const unsigned char* s_dir = stage + L.dir + fvm_shift(a.dir_id + v0);
#endif
#if FVM_VERT_MASKED
    const unsigned char* s_vm = stage + L.vmask + fvm_shift(a.vmask + v0);
#endif
#if FVM_STAGE_OUT
    double* out = (double*)(sm + L.out + (unsigned)(it & 1) * L.out_bytes);
#endif

#ifdef FVM_DEBUG_SKIP_COMPUTE  // experiment: stream the tiles, do no math (results are garbage)
    if (false)
#endif
    for (int g0 = 0; g0 < d.nr; g0 += NG) {  // trip count uniform over the consumer warps
      const int g = g0 + gi;                 // row of this lane group (tile-local)
      const bool live = g < d.nr;
      int b = 0, e = 0;
      if (live) {
        b = s_rp[g] - d.s0;
        e = s_rp[g + 1] - d.s0;
      }
      double x00 = 0.0, x01 = 0.0, x02 = 0.0, u0 = 0.0;
      int dir = 0;
      if (live) {
#if FVM_STAGE_X
#if FVM_DIM == 2
        const double2 q0 = *(const double2*)(s_xy + 2 * g);
        x00 = q0.x;
        x01 = q0.y;
#else
        x00 = s_xy[3 * g];
        x01 = s_xy[3 * g + 1];
        x02 = s_xy[3 * g + 2];
#endif
#endif
#if FVM_STAGE_U
        u0 = s_u[g];
#endif
#if FVM_HAS_DIRICHLET
        dir = s_dir[g];
#endif
      }
      double sD = 0.0, sC = 0.0;  // this lane's share of the row's diagonal / rhs sums
      int ds = -1;                // the row's diagonal slot, seen by exactly one lane
      for (int s = b + gl; s < e; s += FVM_GROUP) {
        const int lo = s_off[s];
        const int hi = (s + 1 < d.nsl) ? (int)s_off[s + 1] : d.nhe;
        if (lo == hi) {  // the diagonal slot owns no half-edge of its own
          ds = s;
          continue;
        }
        const int col = s_col[s];
        const unsigned cl = (unsigned)(col - v0);  // tile-local column, if inside
        double x10 = 0.0, x11 = 0.0, x12 = 0.0, u1 = 0.0;
#if FVM_EDGE_USES_X
        if (cl < (unsigned)d.nr) {
#if FVM_DIM == 2
          const double2 q = *(const double2*)(s_xy + 2 * cl);
          x10 = q.x;
          x11 = q.y;
#else
          x10 = s_xy[3 * cl];
          x11 = s_xy[3 * cl + 1];
          x12 = s_xy[3 * cl + 2];
#endif
        } else {
          const double* p1 = a.coords + (size_t)col * FVM_DIM;
#if FVM_DIM == 2
          const double2 q = __ldg((const double2*)p1);
          x10 = q.x;
          x11 = q.y;
#else
          x10 = __ldg(p1);
          x11 = __ldg(p1 + 1);
          x12 = __ldg(p1 + 2);
#endif
        }
#endif
#if FVM_EDGE_USES_U
        u1 = (cl < (unsigned)d.nr) ? s_u[cl] : __ldg(a.u + col);
#endif
        double aD = 0.0, aO = 0.0, aC = 0.0;
#if FVM_EDGE_FACTORED
        // every edge term is ce * g(slot): fold the ce_ratios, evaluate g once
        double ce = s_ce[lo];
        if (hi - lo > 1) {
          ce += s_ce[lo + 1];
#pragma unroll 1
          for (int k = lo + 2; k < hi; ++k) ce += s_ce[k];
        }
        fvm_edge(u0, u1, x00, x01, x02, x10, x11, x12, ce, 0.0, 0xffu, aD, aO, aC);
#else
#pragma unroll 1
        for (int k = lo; k < hi; ++k) {
          const double ce = s_ce[k];
          double el = 0.0;
          unsigned m = 0xffu;
#if FVM_EDGE_USES_EL
          el = s_el[k];
#endif
#if FVM_EDGE_MASKED
          m = s_mask[k];
#endif
          double D = 0.0, O = 0.0, C = 0.0;
          fvm_edge(u0, u1, x00, x01, x02, x10, x11, x12, ce, el, m, D, O, C);
          aD += D;
          aO += O;
          aC += C;
        }
#endif
        sD += aD;
        sC += aC;
#if FVM_MODE != FVM_MODE_RESIDUAL
        // Dirichlet rows keep their pattern but store explicit zeros off the diagonal
        if (dir) aO = 0.0;
#if FVM_STAGE_OUT
        out[s] = aO;
#else
        a.data[d.s0 + s] = aO;
#endif
#endif
      }
#if FVM_GROUP > 1
      // join the lanes of the group: fixed xor tree (deterministic)
#pragma unroll
      for (int w = FVM_GROUP / 2; w > 0; w >>= 1) {
        sD += __shfl_xor_sync(0xffffffffu, sD, w);
        sC += __shfl_xor_sync(0xffffffffu, sC, w);
        ds = max(ds, __shfl_xor_sync(0xffffffffu, ds, w));
      }
#endif
      if (live && gl == 0) {
        double x2 = x02, cvv = 0.0;
#if FVM_STAGE_CV
        cvv = s_cv[g];
#endif
#if FVM_HAS_VERTEX
        {
          unsigned vm = 0xffu;
#if FVM_VERT_MASKED
          vm = s_vm[g];
#endif
          double VL = 0.0, VC = 0.0;
          fvm_vertex(u0, cvv, x00, x01, x2, vm, VL, VC);
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
        if (dir) {
          double dc = 0.0, dv = 0.0;
          fvm_dirichlet(dir, u0, x00, x01, x2, dc, dv);
          sD = dc;       // row := coeff * e_row
          out_vec = dv;  // rhs := -affine (linear) / F := condition(u) (residual)
        }
#endif
#if FVM_MODE != FVM_MODE_RESIDUAL
#if FVM_STAGE_OUT
        out[ds] = sD;
#else
        a.data[d.s0 + ds] = sD;
#endif
#endif
#if FVM_MODE != FVM_MODE_JACOBIAN
        a.vec[d.r0 + g] = out_vec;
#endif
      }
    }
    // ---- release the stage to the producer (all lanes of this warp are done)
    __syncwarp();
    if ((threadIdx.x & 31) == 0) fvm_mbar_arrive(empty + st);
#if FVM_STAGE_OUT
    // ---- coalesced copy-out of the tile's CSR values
    fvm_consumer_sync();
    for (int s = ct; s < d.nsl; s += FVM_NCT) a.data[d.s0 + s] = out[s];
#endif
  }
}
