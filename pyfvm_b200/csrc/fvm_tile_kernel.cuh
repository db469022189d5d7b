// Hand-written sm_100a assembly template for the finite-volume hot path.
//
// One template serves the three numeric entry points of the reference:
//   FVM_MODE_LINEAR   : matrix values + rhs of pyfvm.discretize_linear  [README:49]
//   FVM_MODE_RESIDUAL : f.eval(u)                                       [README:121,136]
//   FVM_MODE_JACOBIAN : jacobian.get_linear_operator(u0) values         [README:116]
//
// It is compiled at run time by NVRTC after a generated prelude that defines
//   FVM_MODE, FVM_DIM, FVM_THREADS,
//   FVM_EDGE_USES_X / _EL / _U, FVM_VERT_USES_X / _U / _CV, FVM_EDGE_MASKED,
//   FVM_VERT_MASKED, FVM_HAS_DIRICHLET, FVM_HAS_VERTEX
// and the three sympy-printed device functors
//   fvm_edge(uk0, uk1, x0_0..2, x1_0..2, edge_ce_ratio, edge_length, mask, &D, &O, &C)
//   fvm_vertex(uk0, control_volume, x_0..2, mask, &L, &C)
//   fvm_dirichlet(id, uk0, x_0..2, &coeff, &val)
// (edge functor outputs per half-edge (row, col): D -> slot (row,row),
//  O -> slot (row,col), C -> rhs/residual of row).
//
// Data movement per CTA (= one tile of consecutive rows):
//   1. one elected thread issues TMA bulk copies (cp.async.bulk, UBLKCP) of the
//      tile's contiguous ce_ratio / edge_length / mask streams into shared
//      memory, completion tracked by an mbarrier;
//   2. meanwhile all threads paint slot -> row into shared memory;
//   3. slot-parallel phase: thread s folds the contributions of CSR slot s from
//      shared memory in half-edge order (deterministic, atomic-free), gathers
//      x/u of its column once, and stores data[s] coalesced;
//   4. row-parallel phase: thread r folds the per-slot partials of its row into
//      the diagonal slot and rhs/F, adds the vertex term and applies the
//      Dirichlet row replacement in the same pass.
// The fold order of a slot/row depends only on the row's own half-edge order,
// never on tile boundaries, so a row-partitioned multi-GPU run is bit-identical
// to the single-GPU run.

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

__device__ __forceinline__ unsigned fvm_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void fvm_mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fvm_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fvm_mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fvm_smem_u32(bar)),
               "r"(bytes)
               : "memory");
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

extern "C" __global__ void __launch_bounds__(FVM_THREADS) fvm_tile_kernel(const FvmTileArgs a) {
  extern __shared__ __align__(16) unsigned char fvm_smem[];
  const int t = blockIdx.x;
  const int r0 = a.tile_row[t];
  const int r1 = a.tile_row[t + 1];
  if (r0 >= r1) return;  // empty tile (a row heavier than the tile quantum precedes it)
  const long long h0 = a.heptr[r0];
  const long long h1 = a.heptr[r1];
  const int s0 = a.indptr[r0];
  const int nsl = a.indptr[r1] - s0;
  const int nhe = (int)(h1 - h0);
  const int nr = r1 - r0;

  const FvmSmemLayout L =
      fvm_smem_layout(a.max_he, a.max_slots, a.max_rows, FVM_EDGE_USES_EL, FVM_EDGE_MASKED);
  void* bar = fvm_smem + L.bar;
  const double* sm_ce = (const double*)(fvm_smem + L.ce);
  const double* sm_el = (const double*)(fvm_smem + L.el);
  const unsigned char* sm_mask = fvm_smem + L.mask;
  double* pd = (double*)(fvm_smem + L.pd);
  double* pc = (double*)(fvm_smem + L.pc);
  unsigned short* srow = (unsigned short*)(fvm_smem + L.srow);
  unsigned short* dpos = (unsigned short*)(fvm_smem + L.dpos);

  // ---- 1. TMA bulk load of the tile's half-edge streams -------------------
  const long long ha = h0 & ~1LL;  // 16B-aligned start of the fp64 streams
  const int sh = (int)(h0 - ha);   // 0 or 1: shift of half-edge 0 inside smem
#if FVM_EDGE_MASKED
  const long long hb = h0 & ~15LL;
  const int shm = (int)(h0 - hb);
#endif
  if (threadIdx.x == 0) fvm_mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0 && nhe > 0) {
    const unsigned bytes = fvm_round16((unsigned)(h1 - ha) * 8u);
    unsigned total = bytes;
#if FVM_EDGE_USES_EL
    total += bytes;
#endif
#if FVM_EDGE_MASKED
    const unsigned mbytes = fvm_round16((unsigned)(h1 - hb));
    total += mbytes;
#endif
    fvm_mbar_expect_tx(bar, total);
    fvm_bulk_g2s((void*)sm_ce, a.he_ce + ha, bytes, bar);
#if FVM_EDGE_USES_EL
    fvm_bulk_g2s((void*)sm_el, a.he_el + ha, bytes, bar);
#endif
#if FVM_EDGE_MASKED
    fvm_bulk_g2s((void*)sm_mask, a.he_mask + hb, mbytes, bar);
#endif
  }

  // ---- 2. paint slot -> row while the copy is in flight -------------------
  for (int r = threadIdx.x; r < nr; r += FVM_THREADS) {
    const int b = a.indptr[r0 + r] - s0;
    const int e = a.indptr[r0 + r + 1] - s0;
    for (int s = b; s < e; ++s) srow[s] = (unsigned short)r;
  }
  __syncthreads();
  if (nhe > 0) fvm_mbar_wait(bar, 0);

  // ---- 3. slot-parallel fold ----------------------------------------------
  for (int s = threadIdx.x; s < nsl; s += FVM_THREADS) {
    const int lo = a.slot_off[s0 + s];
    const int hi = (s + 1 < nsl) ? (int)a.slot_off[s0 + s + 1] : nhe;
    const int r = srow[s];
    if (lo == hi) {  // the diagonal slot owns no half-edge of its own
      dpos[r] = (unsigned short)s;
      pd[s] = 0.0;
      pc[s] = 0.0;
      continue;
    }
    const int row = a.row_base + r0 + r;
    const int col = a.colidx[s0 + s];
    double x00 = 0.0, x01 = 0.0, x02 = 0.0, x10 = 0.0, x11 = 0.0, x12 = 0.0;
    double u0 = 0.0, u1 = 0.0;
#if FVM_EDGE_USES_X
    {
      const double* p0 = a.coords + (size_t)row * FVM_DIM;
      const double* p1 = a.coords + (size_t)col * FVM_DIM;
      x00 = __ldg(p0);
      x01 = __ldg(p0 + 1);
      x10 = __ldg(p1);
      x11 = __ldg(p1 + 1);
#if FVM_DIM == 3
      x02 = __ldg(p0 + 2);
      x12 = __ldg(p1 + 2);
#endif
    }
#endif
#if FVM_EDGE_USES_U
    u0 = __ldg(a.u + row);
    u1 = __ldg(a.u + col);
#endif
    double aD = 0.0, aO = 0.0, aC = 0.0;
    for (int k = lo; k < hi; ++k) {
      const double ce = sm_ce[sh + k];
      double el = 0.0;
      unsigned m = 0xffu;
#if FVM_EDGE_USES_EL
      el = sm_el[sh + k];
#endif
#if FVM_EDGE_MASKED
      m = sm_mask[shm + k];
#endif
      double D = 0.0, O = 0.0, C = 0.0;
      fvm_edge(u0, u1, x00, x01, x02, x10, x11, x12, ce, el, m, D, O, C);
      aD += D;
      aO += O;
      aC += C;
    }
    pd[s] = aD;
    pc[s] = aC;
#if FVM_MODE != FVM_MODE_RESIDUAL
    {
#if FVM_HAS_DIRICHLET
      // Dirichlet rows keep their pattern but store explicit zeros off the diagonal
      if (a.dir_id[row]) aO = 0.0;
#endif
      a.data[s0 + s] = aO;
    }
#endif
  }
  __syncthreads();

  // ---- 4. row-parallel fold: diagonal slot, rhs / F, vertex term, Dirichlet
  for (int r = threadIdx.x; r < nr; r += FVM_THREADS) {
    const int b = a.indptr[r0 + r] - s0;
    const int e = a.indptr[r0 + r + 1] - s0;
    double sD = 0.0, sC = 0.0;
    for (int s = b; s < e; ++s) {
      sD += pd[s];
      sC += pc[s];
    }
    const int row = a.row_base + r0 + r;
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, uu = 0.0, cvv = 0.0;
#if FVM_VERT_USES_X
    {
      const double* p = a.coords + (size_t)row * FVM_DIM;
      x0 = __ldg(p);
      x1 = __ldg(p + 1);
#if FVM_DIM == 3
      x2 = __ldg(p + 2);
#endif
    }
#endif
#if FVM_VERT_USES_U
    uu = __ldg(a.u + row);
#endif
#if FVM_VERT_USES_CV
    cvv = __ldg(a.cv + row);
#endif
#if FVM_HAS_VERTEX
    {
      unsigned vm = 0xffu;
#if FVM_VERT_MASKED
      vm = a.vmask[row];
#endif
      double VL = 0.0, VC = 0.0;
      fvm_vertex(uu, cvv, x0, x1, x2, vm, VL, VC);
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
    {
      const int id = a.dir_id[row];
      if (id) {
        double dc = 0.0, dv = 0.0;
        fvm_dirichlet(id, uu, x0, x1, x2, dc, dv);
        sD = dc;       // row := coeff * e_row
        out_vec = dv;  // rhs := -affine (linear) / F := condition(u) (residual)
      }
    }
#endif
#if FVM_MODE != FVM_MODE_RESIDUAL
    a.data[s0 + dpos[r]] = sD;
#endif
#if FVM_MODE != FVM_MODE_JACOBIAN
    a.vec[r0 + r] = out_vec;
#endif
  }
}
