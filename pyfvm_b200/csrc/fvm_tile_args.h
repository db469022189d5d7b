// Kernel-parameter block shared by the host library (fvm_b200.cu) and the
// NVRTC-compiled tile kernels (fvm_tile_kernel.cuh).  Plain C types only.
#ifndef FVM_TILE_ARGS_H
#define FVM_TILE_ARGS_H

// sympy.ccode prints pi as M_PI, which NVRTC does not define; this
// header precedes the generated functors in the translation unit
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// A "tile" is a run of consecutive matrix rows [r0, r0+nr) whose half-edges
// and CSR slots are contiguous in HBM:
//   half-edges [h0, h0+nhe) -> he_ce / he_el / he_mask streams
//   slots      [s0, s0+nsl) -> colidx / slot_off / data
//   vertices   [row_base+r0, +nr) -> cv / dir_id / vmask slices
// Per-vertex data that slots GATHER (coordinates, u) is staged as a tile-local
// table: the tile's own vertices (plus a margin of a few vertices either side),
// then up to FVM_NHALO "halo" ranges holding the columns outside the tile (a
// structured triangle mesh needs 2, a structured tetrahedral mesh 8).  slot_meta
// carries each slot's index into that table, so the math warps never issue a
// global load.  Ranges start at even vertices and have even length: every bulk
// copy is 16-byte aligned for 8-byte elements.
//
// The half-edge streams are PADDED: every slot starts at an even position and holds an
// even number of entries (an odd slot gets one entry with ce = +0.0, mask 0), so a slot's
// ce_ratios are read with aligned 16-byte loads ("pairs") only.
//
// slot_meta has one entry per OFF-DIAGONAL CSR slot (the diagonal slot of a row owns
// no half-edge; the row pass writes it), packing everything the slot-parallel pass
// needs:
//   bits  0..10  first pair (16-byte unit) of the slot, relative to the tile  (< 2048)
//   bit  11      the slot holds more than one pair (then the next entry's pair index,
//                or the tile's end, closes it; interior triangle-mesh slots hold one)
//   bits 12..19  row of the slot, relative to the tile               (< 256)
//   bit  20      the slot sits after its row's diagonal slot
//   bits 21..31  index of the slot's column in the tile's vertex table
//                (FVM_XLOC_NONE: not covered -> cold global gather)
// Entry k of a tile (k counted from the tile's first off-diagonal slot) is CSR slot
// s0 + k + row + after.  The pattern build sizes the tiles so that the ranges hold.
#define FVM_NHALO 8
struct FvmTileDesc {  // 128 bytes, one per tile
  long long h0;
  int r0, nr;
  int s0, nsl;
  int nhe;
  int own_start, own_count;  // vertices [own_start, +own_count) = rows of the tile (+ margin)
  int nhalo;                 // halo ranges in use
  int hstart[FVM_NHALO];     // first vertex of every halo range (ascending, disjoint)
  unsigned short hcount[FVM_NHALO];
  int uncovered;             // != 0: some slot's column is outside the table (cold gather)
  int pad[9];
};

#define FVM_META_PAIR_MASK 0x7ffu
#define FVM_META_MULTI_BIT 11
#define FVM_META_ROW_SHIFT 12
#define FVM_META_ROW_MASK 0xffu
#define FVM_META_AFTER_BIT 20
#define FVM_META_XL_SHIFT 21
#define FVM_XLOC_NONE 0x7ffu  // slot_meta column field: column not in the tile's table
#define FVM_TILE_MAX_HE 4095    // half-edges per tile
#define FVM_TILE_MAX_ROWS 255   // rows per tile
#define FVM_TILE_MAX_XTAB 2046  // vertex-table entries per tile

// Dynamic shared-memory carve-up of one CTA, identical on host and device:
//   [full/empty mbarriers][descriptor-ring mbarriers][descriptor rings][stage 0][stage 1]...
// The stream offsets (desc .. vmask) are relative to the start of a stage.
// Every staged stream gets 16 spare bytes: its bulk copy starts at the 16B
// boundary below the first element and is rounded up to 16B.
#define FVM_NDESC 3  // tile descriptors a producer warp keeps in flight (its descriptor ring)
struct FvmSmemLayout {
  unsigned bar, stage0, stage_bytes;
  unsigned dbar, ring;  // per producer warp: FVM_NDESC mbarriers / 128-byte descriptor slots
  // within a stage (accd / accc: per-slot diagonal / rhs contributions waiting for
  // the row fold -- SpMV parks its products in place, over `mat`; dg: offset of the
  // diagonal slot inside every row)
  unsigned desc, ce, el, mask, meta, rowptr, cv, sa, xy, u, dir, vmask, accd, accc, dg, mat;
  unsigned total;
};

struct FvmTileArgs {
  const FvmTileDesc* tiles;        // [ntiles]
  const int* indptr;               // [nrows+1] CSR row pointer
  const int* colidx;               // [nnz] column (local vertex id) of every slot
  const unsigned* slot_meta;       // [nnz - nrows] packed half-edge | row | after | column
  const unsigned short* row_diag;  // [nrows] offset of the diagonal slot inside its row
  const double* he_ce;             // [H] edge_ce_ratio in half-edge order
  const double* he_el;             // [H] edge_length in half-edge order (or null)
  const unsigned char* he_mask;    // [H] edge-term subdomain bits (or null)
  const double* coords;            // [nverts*xs] vertex coordinates, xs = 2 (dim 2) or 4 (dim 3)
  const double* cv;                // [nverts] control volumes
  const double* sa;                // [nverts] boundary surface areas (dGamma terms; or null)
  const unsigned char* vmask;      // [nverts] vertex-term subdomain bits (or null)
  const unsigned char* dir_id;     // [nverts] 0 = free, k = Dirichlet term k-1 wins (or null)
  const double* u;                 // [nverts] state (residual / Jacobian modes), x (SpMV)
  const double* mat;               // [nnz] CSR values to multiply with (SpMV mode)
  double* data;                    // [nnz] CSR values out (assembly / Jacobian modes)
  double* vec;                     // [nrows] rhs (assembly) or F (residual) out
  // fused halo wait (row-partitioned runs): a tile whose vertex table reaches into the
  // ghost vertices [0, ghost_lo_end) or [ghost_hi_begin, nverts) makes the producer
  // acquire every neighbour's epoch flag before it copies u; tiles are walked starting
  // at tile_rot (interior first), so the NVLink transfer hides behind the interior math
  const unsigned long long* const* halo_flags;  // [n_halo_flags] flag words (or null)
  unsigned long long halo_epoch;
  // non-null: the epoch to wait for is read from device memory (a launch replayed from a
  // CUDA graph cannot carry a changing kernel parameter)
  const unsigned long long* halo_epoch_ptr;
  int* halo_timeout;               // set to 1 when a wait gives up (never hangs)
  int n_halo_flags;
  int ghost_lo_end, ghost_hi_begin;
  int tile_rot;
  int row_base;                    // local vertex id of row 0 (row partitioning)
  int ntiles;
  int max_he;                      // largest tile: half-edges
  int max_slots;                   //               slots
  int max_rows;                    //               rows
  int max_xtab;                    //               vertex-table entries (own + halos)
  struct FvmSmemLayout L;          // shared-memory carve-up (host: fvm_smem_layout)
};

#define FVM_MODE_LINEAR 0    // matrix values + rhs (discretize_linear)
#define FVM_MODE_RESIDUAL 1  // F(u) (f.eval)
#define FVM_MODE_JACOBIAN 2  // matrix values at u (jacobian.get_linear_operator)
#define FVM_MODE_SPMV 3      // y = A x on the mesh's pattern (Krylov solvers around the path)

// What a functor set reads; decides which streams are staged in shared memory.
#define FVM_F_EDGE_EL 1u    // edge functor reads edge_length
#define FVM_F_EDGE_MASK 2u  // edge terms carry cell-subdomain bits
#define FVM_F_X 4u          // any functor reads vertex coordinates
#define FVM_F_U 8u          // any functor reads the state u
#define FVM_F_CV 16u        // vertex functor reads control volumes
#define FVM_F_VMASK 32u     // vertex terms carry vertex-subdomain bits
#define FVM_F_DIR 64u       // Dirichlet conditions present
#define FVM_F_EDGE_D 128u   // edge functor writes D (diagonal coefficient) in this mode
#define FVM_F_EDGE_C 256u   // edge functor writes C (rhs / residual part) in this mode
#define FVM_F_SA 512u       // vertex functor reads boundary surface areas (dGamma terms)

#ifdef __CUDACC__
#define FVM_HD __host__ __device__
#else
#define FVM_HD
#endif

FVM_HD inline unsigned fvm_round16(unsigned x) { return (x + 15u) & ~15u; }

FVM_HD inline unsigned fvm_coord_stride(int dim) { return dim == 2 ? 2u : 4u; }

// max_slots: off-diagonal slots of the largest tile (slot_meta and the parked shares
// have one entry per off-diagonal slot)
FVM_HD inline FvmSmemLayout fvm_smem_layout(int max_he, int max_slots, int max_rows, int max_xtab,
                                            int dim, unsigned flags, int mode, int stages,
                                            int max_csr_slots, int producers) {
  FvmSmemLayout L;
  const unsigned he = (unsigned)max_he, sl = (unsigned)max_slots, nr = (unsigned)max_rows;
  unsigned o = 0;  // within a stage
  const unsigned xt = (unsigned)max_xtab;
  L.desc = o;   o += 128;
  // (the half-edge streams of a tile start 16-byte aligned: no slack)
  L.ce = o;     o += (mode != FVM_MODE_SPMV) ? fvm_round16(he * 8u) : 0u;
  L.el = o;     o += (flags & FVM_F_EDGE_EL) ? fvm_round16(he * 8u) : 0u;
  L.mask = o;   o += (flags & FVM_F_EDGE_MASK) ? fvm_round16(he) + 16u : 0u;
  L.meta = o;   o += fvm_round16(sl * 4u) + 16u;
  L.rowptr = o; o += fvm_round16((nr + 1u) * 4u) + 16u;
  L.cv = o;     o += (flags & FVM_F_CV) ? fvm_round16(nr * 8u) + 16u : 0u;
  L.sa = o;     o += (flags & FVM_F_SA) ? fvm_round16(nr * 8u) + 16u : 0u;
  L.xy = o;     o += (flags & FVM_F_X) ? fvm_round16(xt * 8u * fvm_coord_stride(dim)) : 0u;
  L.u = o;      o += (flags & FVM_F_U) ? fvm_round16(xt * 8u) : 0u;
  L.dir = o;    o += (flags & FVM_F_DIR) ? fvm_round16(nr) + 16u : 0u;
  L.vmask = o;  o += (flags & FVM_F_VMASK) ? fvm_round16(nr) + 16u : 0u;
  const unsigned csl = (unsigned)max_csr_slots;  // parked shares keep the CSR row stride
  L.accd = o;   o += ((flags & FVM_F_EDGE_D) && mode != FVM_MODE_SPMV) ? fvm_round16(csl * 8u) : 0u;
  L.accc = o;   o += ((flags & FVM_F_EDGE_C) && mode != FVM_MODE_SPMV) ? fvm_round16(csl * 8u) : 0u;
  L.dg = o;     o += (mode != FVM_MODE_RESIDUAL) ? fvm_round16(nr * 2u) + 16u : 0u;
  L.mat = o;    o += (mode == FVM_MODE_SPMV) ? fvm_round16((unsigned)max_csr_slots * 8u) + 16u : 0u;
  L.stage_bytes = o;
  L.bar = 0;  // 3 x 8B mbarriers per stage, then the descriptor-ring mbarriers and slots
  L.dbar = fvm_round16(24u * (unsigned)stages);
  L.ring = fvm_round16(L.dbar + 8u * FVM_NDESC * (unsigned)producers);
  L.stage0 = L.ring + 128u * FVM_NDESC * (unsigned)producers;
  o = L.stage0 + L.stage_bytes * (unsigned)stages;
  L.total = o;
  return L;
}

#endif
