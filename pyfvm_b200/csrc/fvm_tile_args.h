// Kernel-parameter block shared by the host library (fvm_b200.cu) and the
// NVRTC-compiled tile kernels (fvm_tile_kernel.cuh).  Plain C types only.
#ifndef FVM_TILE_ARGS_H
#define FVM_TILE_ARGS_H

// A "tile" is a run of consecutive matrix rows [tile_row[t], tile_row[t+1])
// whose half-edges and CSR slots are contiguous in HBM:
//   half-edges [heptr[r0], heptr[r1])   -> he_ce / he_el / he_mask streams
//   slots      [indptr[r0], indptr[r1]) -> colidx / slot_off / data
struct FvmTileArgs {
  const int* tile_row;             // [ntiles+1] first row of every tile
  const long long* heptr;          // [nrows+1] first half-edge of every row
  const int* indptr;               // [nrows+1] CSR row pointer
  const int* colidx;               // [nnz] column (local vertex id) of every slot
  const unsigned short* slot_off;  // [nnz] first half-edge of the slot, relative to its tile
  const double* he_ce;             // [H] edge_ce_ratio in half-edge order
  const double* he_el;             // [H] edge_length in half-edge order (or null)
  const unsigned char* he_mask;    // [H] edge-term subdomain bits (or null)
  const double* coords;            // [nverts*dim] vertex coordinates
  const double* cv;                // [nverts] control volumes
  const unsigned char* vmask;      // [nverts] vertex-term subdomain bits (or null)
  const unsigned char* dir_id;     // [nverts] 0 = free, k = Dirichlet term k-1 wins (or null)
  const double* u;                 // [nverts] state (residual / Jacobian modes)
  double* data;                    // [nnz] CSR values out (assembly / Jacobian modes)
  double* vec;                     // [nrows] rhs (assembly) or F (residual) out
  int row_base;                    // local vertex id of row 0 (row partitioning)
  int ntiles;
  int max_he;                      // largest tile: half-edges
  int max_slots;                   //               slots
  int max_rows;                    //               rows
};

#define FVM_MODE_LINEAR 0    // matrix values + rhs (discretize_linear)
#define FVM_MODE_RESIDUAL 1  // F(u) (f.eval)
#define FVM_MODE_JACOBIAN 2  // matrix values at u (jacobian.get_linear_operator)


// Dynamic shared-memory carve-up of one CTA, identical on host and device.
struct FvmSmemLayout {
  unsigned bar, ce, el, mask, pd, pc, srow, dpos, total;
};

#ifdef __CUDACC__
#define FVM_HD __host__ __device__
#else
#define FVM_HD
#endif

FVM_HD inline unsigned fvm_round16(unsigned x) { return (x + 15u) & ~15u; }

FVM_HD inline FvmSmemLayout fvm_smem_layout(int max_he, int max_slots, int max_rows,
                                            int uses_el, int masked) {
  FvmSmemLayout L;
  unsigned o = 0;
  L.bar = o;  o += 16;
  // +2 doubles: the bulk copy starts at an even half-edge and ends 16B-rounded
  L.ce = o;   o += fvm_round16((unsigned)(max_he + 2) * 8u);
  L.el = o;   o += uses_el ? fvm_round16((unsigned)(max_he + 2) * 8u) : 0u;
  L.mask = o; o += masked ? fvm_round16((unsigned)max_he + 32u) : 0u;
  L.pd = o;   o += fvm_round16((unsigned)max_slots * 8u);
  L.pc = o;   o += fvm_round16((unsigned)max_slots * 8u);
  L.srow = o; o += fvm_round16((unsigned)max_slots * 2u);
  L.dpos = o; o += fvm_round16((unsigned)max_rows * 2u);
  L.total = o;
  return L;
}

#endif
