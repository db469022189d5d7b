/* fvm_b200 -- C-ABI of the B200-native finite-volume assembly engine.
 *
 * The reference (meshpro/pyfvm) is pure Python and has no FFI; this header is
 * the boundary a maintainer would bind (ctypes stub in INTEGRATION.md) to
 * replace the numeric half of
 *   pyfvm.discretize_linear(problem, mesh) -> (matrix, rhs)   [README.md:49]
 *   f.eval(u)                                                 [README.md:121,136]
 *   jacobian.get_linear_operator(u0)                          [README.md:116]
 *   the norm inside pyfvm.newton                              [README.md:121]
 * Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - every call returns 0 on success, nonzero on error; fvm_last_error()
 *     returns the message of the last failing call on this thread;
 *   - handles are opaque pointers owned by the library until *_destroy;
 *   - one ctx = one device + one CUDA stream; calls on a ctx are not
 *     re-entrant; use one ctx per rank (one process per GPU);
 *   - pointers named *_dev are device pointers valid on the ctx's device
 *     (from fvm_dev_alloc or any other allocator, e.g. torch's data_ptr());
 *     all other pointers are host pointers.
 */
#ifndef FVM_B200_H
#define FVM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fvm_ctx fvm_ctx;
typedef struct fvm_mesh fvm_mesh;
typedef struct fvm_kernel fvm_kernel;

/* Kernel modes: which numeric entry point of the reference a compiled functor
 * set implements. */
#define FVM_LINEAR 0   /* discretize_linear: CSR values + rhs   [README.md:49]  */
#define FVM_RESIDUAL 1 /* f.eval(u)                             [README.md:121] */
#define FVM_JACOBIAN 2 /* jacobian.get_linear_operator(u0)      [README.md:116] */
#define FVM_SPMV 3     /* y = A x on the mesh's pattern (fvm_spmv)               */

/* fvm_kernel_create flags: what the functor set reads (must equal the
 * FVM_FLAGS the generated prelude defines) */
#define FVM_KF_EDGE_EL 1u    /* edge functor reads edge_length               */
#define FVM_KF_EDGE_MASK 2u  /* edge terms restricted to cell subdomains     */
#define FVM_KF_X 4u          /* some functor reads vertex coordinates        */
#define FVM_KF_U 8u          /* some functor reads the state u               */
#define FVM_KF_CV 16u        /* vertex functor reads control volumes         */
#define FVM_KF_VMASK 32u     /* vertex terms restricted to vertex subdomains */
#define FVM_KF_DIR 64u       /* Dirichlet conditions present                 */
#define FVM_KF_EDGE_D 128u   /* edge functor writes a diagonal coefficient   */
#define FVM_KF_EDGE_C 256u   /* edge functor writes an rhs / residual part   */
#define FVM_KF_SA 512u       /* vertex functor reads boundary surface areas  */

int fvm_version(void);
const char* fvm_last_error(void);

/* ---- context --------------------------------------------------------------*/
int fvm_ctx_create(int device, fvm_ctx** out);
int fvm_ctx_destroy(fvm_ctx* ctx);
int fvm_ctx_sync(fvm_ctx* ctx);
/* run every later call of this ctx on a caller-owned CUDA stream (cudaStream_t passed
 * as void*, e.g. torch.cuda.current_stream().cuda_stream) so that the library's
 * kernels are stream-ordered with the caller's NCCL collectives; use_external = 0
 * returns to the ctx's own stream.  Synchronises the old stream first. */
int fvm_ctx_set_stream(fvm_ctx* ctx, void* cuda_stream, int use_external);
/* name, SM count, bytes of global memory of the ctx's device */
int fvm_ctx_info(fvm_ctx* ctx, char* name, size_t name_len, int* sm_count, size_t* mem_bytes);

/* ---- device / pinned memory (so a host binding needs no other GPU library) */
int fvm_dev_alloc(fvm_ctx* ctx, size_t bytes, void** out_dev);
int fvm_dev_free(fvm_ctx* ctx, void* ptr_dev);
/* Device memory released through this library (fvm_dev_free, fvm_mesh_destroy, build
 * temporaries) is kept in a per-device block cache and handed out again to requests of
 * nearly the same size (FVM_DEVICE_CACHE_BYTES, default 24 GB; 0 turns it off): the repeated
 * discretize_linear(problem, mesh) calls of [README.md:49] otherwise pay cudaMalloc/cudaFree
 * of several GB each time.  Blocks exported with fvm_ipc_export are never recycled. */
int fvm_device_cache_stats(fvm_ctx* ctx, int64_t* idle_bytes, int64_t* idle_blocks,
                           int64_t* live_blocks);
/* give every idle block back to the driver (e.g. before another library needs the memory) */
int fvm_device_cache_flush(fvm_ctx* ctx);
int fvm_dev_memset(fvm_ctx* ctx, void* ptr_dev, int value, size_t bytes);
int fvm_h2d(fvm_ctx* ctx, void* dst_dev, const void* src, size_t bytes);
int fvm_d2h(fvm_ctx* ctx, void* dst, const void* src_dev, size_t bytes);
/* device -> device copy, stream-ordered on the ctx stream */
int fvm_d2d(fvm_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
/* device -> pinned host copy on the ctx's copy stream, ordered after the work issued
 * so far on the ctx stream; it overlaps whatever the ctx stream does next (e.g.
 * the next step's host -> device upload: PCIe is full duplex) */
int fvm_d2h_async(fvm_ctx* ctx, void* dst_pinned, const void* src_dev, size_t bytes);
/* later work on the ctx stream waits for the copies issued so far (call it before
 * overwriting a buffer an async copy is still reading) */
int fvm_copy_fence(fvm_ctx* ctx);
/* host waits for the copies issued so far */
int fvm_copy_sync(fvm_ctx* ctx);
int fvm_host_alloc(fvm_ctx* ctx, size_t bytes, void** out); /* pinned */
int fvm_host_free(fvm_ctx* ctx, void* ptr);

/* ---- timing on the ctx stream (CUDA events) -------------------------------*/
int fvm_timer_begin(fvm_ctx* ctx);
int fvm_timer_end(fvm_ctx* ctx, float* out_ms); /* records, syncs, returns ms */

/* ---- mesh + sparsity pattern -----------------------------------------------
 * Replaces what the reference re-derives on every call from
 * mesh.idx_hierarchy / ce_ratios / ei_dot_ei / node_coords / control_volumes
 * and scipy's coo_matrix(...).tocsr() sort [README.md:49].  Built once per mesh
 * on the device: radix sort of the half-edge (row, col) pairs -> CSR
 * indptr/indices + the half-edge -> slot permutation, with the per-record
 * geometry permuted into half-edge order. */
typedef struct fvm_mesh_desc {
  int32_t dim;          /* coordinate columns stored (2 or 3)                  */
  int32_t n_vertices;   /* local vertices (owned rows + ghosts)                */
  int32_t row_base;     /* first owned vertex id; rows = [row_base, +n_rows)   */
  int32_t n_rows;       /* owned rows (== n_vertices on one GPU)               */
  int64_t n_records;    /* columns of idx_hierarchy (edge contributions)       */
  int32_t index_bytes;  /* 4 or 8: width of idx0/idx1 entries                  */
  int32_t on_device;    /* 0: arrays below are host pointers; 1: device        */
  const void* idx0;     /* [n_records] first endpoint  (idx_hierarchy[0].ravel) */
  const void* idx1;     /* [n_records] second endpoint (idx_hierarchy[1].ravel) */
  const double* ce_ratios;  /* [n_records]                                     */
  const double* edge_len;   /* [n_records] sqrt(ei_dot_ei), may be NULL         */
  const uint8_t* rec_mask;  /* [n_records] edge-term subdomain bits, may be NULL
                               (records with mask 0 are left out of the pattern) */
  const double* coords;     /* [n_vertices*coord_cols] row-major (first dim columns used) */
  const double* control_volumes; /* [n_vertices]; NULL: derived on the device for the
                               owned rows (needs edge_len), as sum |e|^2 ce / (2 cell_dim) */
  int32_t tile_quantum; /* 0 = default; half-edges+slots per tile              */
  int32_t cell_dim;     /* 2 triangles, 3 tetrahedra (default tile size; scale of the
                           control volumes derived when control_volumes == NULL)       */
  int32_t drop_edge_len; /* 1: edge_len only serves to derive the control volumes; its
                           half-edge stream is freed after the build (functors that read
                           edge_length then fail at launch) */
  int32_t coord_cols;   /* columns per vertex of the coords array, >= dim (0 = dim): a
                           planar mesh carried with 3 columns needs no sliced host copy */
} fvm_mesh_desc;

int fvm_mesh_create(fvm_ctx* ctx, const fvm_mesh_desc* desc, fvm_mesh** out);
int fvm_mesh_destroy(fvm_mesh* mesh);
/* nnz, number of half-edges kept, number of tiles, milliseconds spent in the
 * on-device pattern build */
int fvm_mesh_info(fvm_mesh* mesh, int64_t* nnz, int64_t* n_halfedges, int32_t* n_tiles,
                  float* build_ms);
/* layout diagnostics, out[0..n): largest tile (half-edges, slots, rows, vertex-table
 * entries), slots whose column lies outside the staged vertex tables (they are
 * gathered from global memory instead), tile quantum, tiles, half-edges, length of the
 * padded half-edge streams (every slot starts even and holds an even number of entries),
 * row chunks the pattern was built in (> 1: the half-edges did not fit next to their sort
 * buffers; FVM_BUILD_CHUNK_HE forces a chunk size) */
int fvm_mesh_stats(fvm_mesh* mesh, int64_t* out, int n);
/* copy an internal layout array to the host for inspection / tests:
 * 0 = tile descriptors (128 B each), 1 = packed slot metadata (4 B per slot),
 * 2 = ce_ratios in half-edge order (8 B per half-edge), 3 = control volumes */
int fvm_mesh_debug_copy(fvm_mesh* mesh, int which, void* out, size_t bytes);
/* CSR pattern to host: indptr[n_rows+1], indices[nnz] (local vertex ids) */
int fvm_mesh_get_pattern(fvm_mesh* mesh, int32_t* indptr, int32_t* indices);
/* device views of the same arrays (owned by the mesh) */
int fvm_mesh_pattern_dev(fvm_mesh* mesh, const int32_t** indptr_dev, const int32_t** indices_dev);
/* re-upload per-record geometry (same connectivity), permuted on the device.
 * n_records / n_vertices are the lengths of the arrays handed in (ce_ratios, edge_len:
 * n_records; coords: n_vertices*dim; control_volumes: n_vertices); the call fails when they
 * differ from the mesh's.  Any of the four arrays may be NULL (left as is). */
int fvm_mesh_update_geometry(fvm_mesh* mesh, int64_t n_records, int32_t n_vertices,
                             const double* ce_ratios, const double* edge_len, const double* coords,
                             const double* control_volumes, int on_device);

/* user-supplied numeric edge blocks (the reference's pyfvm.fvm_matrix.EdgeMatrixKernel:
 * eval(mesh, cell_mask) returns a (2, 2, records) array instead of a symbolic integrand).
 * diag_vals / off_vals hold 2 * n_records entries, side-major: entry h < n_records is the
 * row-of-idx0 view of record h (block[0][0], block[0][1]), entry n_records + h the
 * row-of-idx1 view (block[1][1], block[1][0]).  They REPLACE the mesh's ce_ratio /
 * edge_length half-edge streams; a functor set with D = edge_ce_ratio, O = edge_length then
 * assembles them through the same tile kernel.  The mesh must have been created with
 * edge_len. */
int fvm_mesh_set_halfedge_values(fvm_mesh* mesh, int64_t n_records, const double* diag_vals,
                                 const double* off_vals, int on_device);
/* per-vertex measure of the domain boundary (meshplex's surface_areas: the part of the
 * boundary faces that belongs to a vertex; 0 off the boundary) -- the weight of dGamma
 * integrals (Neumann / Robin terms), which the reference evaluates per boundary vertex like
 * a dV term with surface_area in place of control_volume [README.md:79 "more examples"].
 * Only needed when a functor set carries FVM_KF_SA. */
int fvm_mesh_set_surface_areas(fvm_mesh* mesh, const double* surface_areas, int32_t n_vertices,
                               int on_device);

/* ---- mesh generation and geometry on the device (the step before the
 * path).  meshplex derives idx_hierarchy / ce_ratios / ei_dot_ei / control_volumes
 * with numpy on the host; for the largest meshes the per-record arrays only ever
 * exist in HBM.  Record order = meshplex's idx_hierarchy flattened: triangles
 * record = k*cells + c, tetrahedra record = (j*4 + k)*cells + c.
 * Generators reproduce meshzoo.rectangle_tri / cube_tetra [README.md:44-46] (x fastest;
 * `parity` shifts the zigzag so a slab of a larger grid reproduces the global cells);
 * xs/ys/zs are HOST axis arrays, outputs are caller-allocated device arrays:
 * points [n*2 | n*3], cells [2(nx-1)(ny-1)*3 | 5(nx-1)(ny-1)(nz-1)*4]. */
int fvm_gen_rectangle_tri(fvm_ctx* ctx, int nx, int ny, int parity, const double* xs,
                          const double* ys, int index_bytes, double* points_dev, void* cells_dev);
int fvm_gen_cube_tetra(fvm_ctx* ctx, int nx, int ny, int nz, int parity, const double* xs,
                       const double* ys, const double* zs, int index_bytes, double* points_dev,
                       void* cells_dev);
/* per-record endpoints, ce_ratios and edge lengths of n_cells simplices (device in,
 * device out; R = 3*n_cells or 12*n_cells records; el_dev may be NULL) */
int fvm_geometry(fvm_ctx* ctx, int cell_dim, int coord_dim, int64_t n_cells, const void* cells_dev,
                 int index_bytes, const double* points_dev, void* idx0_dev, void* idx1_dev,
                 double* ce_dev, double* el_dev);
/* any simplex mesh: out_dev[v] = 1 (n_vertices bytes, zeroed first) for the vertices of
 * facets that belong to exactly one cell -- meshplex's is_boundary_point -- from a sort of
 * the facet keys on the device */
int fvm_boundary_vertices(fvm_ctx* ctx, int cell_dim, int64_t n_cells, const void* cells_dev,
                          int index_bytes, int32_t n_vertices, uint8_t* out_dev);
/* subdomain bits of a cell-restricted dS term for a mesh whose cells live on the device:
 * rec_mask_dev[r] |= bit for every record r (record order of fvm_geometry) whose cell has
 * all its vertices flagged in vertex_inside_dev (NULL: every record) */
int fvm_record_mask(fvm_ctx* ctx, int cell_dim, int64_t n_cells, const void* cells_dev,
                    int index_bytes, const uint8_t* vertex_inside_dev, uint8_t bit,
                    uint8_t* rec_mask_dev);
/* triangle meshes: out_dev[v] = 1 for owned vertices on a boundary edge (an edge that
 * belongs to exactly one cell), from the pattern's half-edge counts */
int fvm_mesh_boundary_tri(fvm_mesh* mesh, uint8_t* out_dev);

/* ---- functor compilation (NVRTC -> sm_100a cubin) ---------------------------
 * Replaces sympy.lambdify in the reference's discretize step: functor_src is
 * the sympy.ccode-printed prelude (see pyfvm_b200/codegen.py); it is prepended
 * to the built-in tile-kernel template. */
typedef struct fvm_kernel_opts {
  int32_t mode;           /* FVM_LINEAR / FVM_RESIDUAL / FVM_JACOBIAN               */
  uint32_t flags;         /* FVM_KF_* bits; must equal the prelude's FVM_FLAGS      */
  int32_t fmad;           /* 0: --fmad=false (numpy-like rounding), 1: allow FMA    */
  int32_t consumer_warps; /* math warps per CTA (1..16); = FVM_CONSUMER_WARPS           */
  int32_t stages;         /* shared-memory ring depth; must equal FVM_STAGES        */
  int32_t ctas_per_sm;    /* persistent CTAs per SM, 0 = as many as fit             */
  int32_t producers;      /* producer warps per CTA (1, 2 or 4); = FVM_PRODUCERS    */
} fvm_kernel_opts;
int fvm_kernel_create(fvm_ctx* ctx, const char* functor_src, const fvm_kernel_opts* opts,
                      fvm_kernel** out);
int fvm_kernel_destroy(fvm_kernel* k);
/* Compiled cubins are kept under $FVM_CACHE_DIR (default ~/.cache/fvm_b200; "off"
 * disables), keyed by SHA-256 of the full translation unit + compile options + NVRTC
 * version: a functor set is compiled once per machine, not once per process.  Counters of
 * this process: NVRTC compilations run, cubins taken from the cache. */
int fvm_kernel_cache_stats(int* nvrtc_compiles, int* cubin_cache_hits);
/* registers/thread, static smem, max threads/block of the compiled entry */
int fvm_kernel_info(fvm_kernel* k, int* regs, int* static_smem, int* max_threads);
/* launch shape the kernel would use on this mesh: grid, block, dynamic smem bytes */
int fvm_kernel_launch_shape(fvm_kernel* k, fvm_mesh* mesh, int* grid, int* block, int* smem);
/* the full source handed to NVRTC (template + prelude), for offline inspection */
int fvm_kernel_source(const char* functor_src, char* out, size_t out_len, size_t* needed);

/* ---- the hot path -------------------------------------------------------------
 * One launch of the tile kernel over all rows of the mesh.
 *   mode FVM_LINEAR   : data_dev[nnz], vec_dev[n_rows] (rhs) written
 *   mode FVM_RESIDUAL : vec_dev[n_rows] (F) written; u_dev[n_vertices] read
 *   mode FVM_JACOBIAN : data_dev[nnz] written;       u_dev[n_vertices] read
 * vmask_dev / dirid_dev: per-vertex subdomain bits of the vertex terms and
 * 1-based index of the winning Dirichlet term (0 = free row); may be NULL when
 * the kernel was generated without them.
 * A triangle mesh's second launch (or its first residual / Jacobian / SpMV launch) first
 * re-lays the tiles' vertex tables for shared-memory banks (a few ms, once; results are
 * bit-identical; FVM_BANK_STEERING=0 turns it off).  FVM_TRACE=1 prints the host timeline
 * of fvm_mesh_create and of that step on stderr. */
int fvm_assemble(fvm_mesh* mesh, fvm_kernel* k, const uint8_t* vmask_dev,
                 const uint8_t* dirid_dev, const double* u_dev, double* data_dev,
                 double* vec_dev);

/* ---- the step after the path: y = A x on the assembled system, so
 * that a jacobian_solver callback [README.md:113-117] can run a Krylov method without
 * leaving HBM.  Same tiles, same staged vertex tables as the assembly (no global
 * gather of x), deterministic.  values_dev: CSR values in the mesh's pattern order
 * (what fvm_assemble wrote); x_dev[n_vertices], y_dev[n_rows]. */
int fvm_spmv_kernel_create(fvm_ctx* ctx, fvm_kernel** out);
int fvm_spmv(fvm_mesh* mesh, fvm_kernel* spmv_kernel, const double* values_dev, const double* x_dev,
             double* y_dev);
/* row-partitioned y = A x: x_dev holds the local vector (ghost | owned | ghost); the
 * kernel waits for the neighbours' pushes of the ghost entries itself (see
 * fvm_assemble_halo) */
int fvm_spmv_halo(fvm_mesh* mesh, fvm_kernel* spmv_kernel, const double* values_dev,
                  const double* x_dev, double* y_dev, const uint64_t* const* flags_dev,
                  int n_flags, uint64_t epoch, int32_t* timed_out_dev);
/* diag_dev[r] = A[r, r] (Jacobi preconditioner) */
int fvm_csr_diagonal(fvm_mesh* mesh, const double* values_dev, double* diag_dev);
/* deterministic dot product (fixed-shape two-stage reduction), z = a x + b y, z = x / d */
int fvm_dot(fvm_ctx* ctx, const double* x_dev, const double* y_dev, int64_t n, double* out);
/* out2[0] = x1 . y1, out2[1] = x2 . y2 in one pass and one host synchronisation */
int fvm_dot2(fvm_ctx* ctx, const double* x1_dev, const double* y1_dev, const double* x2_dev,
             const double* y2_dev, int64_t n, double* out2);
int fvm_axpby(fvm_ctx* ctx, double a, const double* x_dev, double b, const double* y_dev,
              double* z_dev, int64_t n);
int fvm_vdiv(fvm_ctx* ctx, const double* x_dev, const double* d_dev, double* z_dev, int64_t n);

/* Jacobi-preconditioned BiCGStab for A x = b (x0 = 0) with every scalar resident in HBM:
 * one iteration (2 SpMV launches of the tile kernel, 3 fused vector kernels, 3 two-stage
 * reductions, 3 scalar steps) is captured once in a CUDA graph and replayed; the host
 * reads the residual every check_every iterations only.  work_dev: 9 * ((n+1)&~1) + 16
 * doubles.  Stops at ||r|| <= tol ||b||; fails on breakdown / maxiter. */
int fvm_bicgstab(fvm_mesh* mesh, fvm_kernel* spmv_kernel, const double* values_dev,
                 const double* b_dev, double* x_dev, double* work_dev, double tol, int maxiter,
                 int check_every, int* iters_out, double* relres_out);

/* The same solver on a ROW PARTITION (one process per GPU), still without the host in the
 * loop: y = A x pushes the owned boundary entries of x into the neighbours' ghost slots
 * (fvm_halo_push_all) and waits for theirs inside the SpMV kernel, with the epochs held in
 * device memory so that the captured iteration can be replayed; the three reductions of an
 * iteration are all-reduces of two doubles over peer memory (every rank stores its partial
 * sums into every rank's inbox, releases an epoch flag, and adds the world's values in rank
 * order: identical bits, hence identical control flow, on every rank).  All vectors hold the
 * owned rows.  The caller maps the peers' buffers (fvm_ipc_*) and fills fvm_dist_args. */
typedef struct fvm_dist_args {
  /* halo of x (same layout as fvm_halo_push_all / fvm_spmv_halo) */
  const void* halo_peers_dev;      /* push table, n_halo_peers records                     */
  int32_t n_halo_peers;
  int64_t max_send;
  uint32_t* push_counters_dev;     /* n_halo_peers zeroed uint32                            */
  const uint64_t* const* halo_flags_dev; /* flag words my neighbours release               */
  int32_t n_halo_flags;
  int32_t* halo_timeout_dev;
  double* xbuf[2];                 /* my two local x buffers (ghost | owned | ghost)        */
  int64_t ghost_lo;                /* entries before the owned part                         */
  uint64_t* halo_epoch_dev;        /* device-resident epoch of this halo (in/out)           */
  /* all-reduce over peer memory */
  const void* reduce_peers_dev;    /* world records { double* inbox; uint64_t* flags; }:
                                      rank r's inbox [2][world][2] and flag words [world]   */
  int32_t world, rank;
  uint64_t* reduce_epoch_dev;      /* device-resident epoch of the reductions (in/out)      */
  int32_t* reduce_timeout_dev;
} fvm_dist_args;
int fvm_bicgstab_dist(fvm_mesh* mesh, fvm_kernel* spmv_kernel, const double* values_dev,
                      const double* b_dev, double* x_dev, double* work_dev, double tol, int maxiter,
                      int check_every, const fvm_dist_args* dist, int* iters_out,
                      double* relres_out);
/* Jacobi-preconditioned CG for the symmetric case (SURVEY 8f rank 2 names CG next to
 * BiCGStab; the README's Poisson and Bratu systems [README.md:36-51, 96-121] are symmetric
 * definite away from their Dirichlet rows): one SpMV and two reductions per iteration,
 * device-resident scalars, two iterations per replayed CUDA graph.  Starts from the Jacobi
 * guess x0 = b / diag(A), which solves pyfvm's Dirichlet rows (identity rows whose columns are
 * NOT eliminated) exactly and so keeps them out of the Krylov space.  dist == NULL: one GPU;
 * otherwise the row partition of fvm_bicgstab_dist.  work_dev: 7 * ((n + 1) & ~1) + 16
 * doubles.  Fails with "CG breakdown" when p.Ap and r.(r/d) differ in sign. */
int fvm_cg(fvm_mesh* mesh, fvm_kernel* spmv_kernel, const double* values_dev, const double* b_dev,
           double* x_dev, double* work_dev, double tol, int maxiter, int check_every,
           const fvm_dist_args* dist_or_null, int* iters_out, double* relres_out);
/* The same with a stronger preconditioner (SURVEY 8f rank 2 asks for more than Jacobi; the
 * README points at AMG [README.md:67-77], which is out of scope): z = p(D^-1 A) D^-1 r with
 * the Chebyshev polynomial of degree `degree` on [lambda_max / 30, lambda_max] -- degree - 1
 * more SpMVs per iteration and no more reductions, i.e. fewer, fatter iterations where an
 * iteration is bound by its rank-wide synchronisations.  lambda_max: an upper bound of the
 * spectrum of D^-1 A, e.g. fvm_csr_gershgorin (max over all ranks).  degree 1 = fvm_cg. */
int fvm_cg_poly(fvm_mesh* mesh, fvm_kernel* spmv_kernel, const double* values_dev,
                const double* b_dev, double* x_dev, double* work_dev, double tol, int maxiter,
                int check_every, int degree, double lambda_max, const fvm_dist_args* dist_or_null,
                int* iters_out, double* relres_out);
/* max_i sum_j |a_ij| / |a_ii| over this mesh's rows (Gershgorin's bound for D^-1 A) */
int fvm_csr_gershgorin(fvm_mesh* mesh, const double* values_dev, double* out);

/* ---- small device helpers used around the path --------------------------------*/
/* deterministic ||x||_2 (fixed-shape two-stage reduction) [README.md:121 newton] */
int fvm_norm2(fvm_ctx* ctx, const double* x_dev, int64_t n, double* out);
/* dst[i] = src[idx[i]]: halo pack for the row-partitioned multi-GPU path */
int fvm_gather_f64(fvm_ctx* ctx, const double* src_dev, const int32_t* idx_dev, int64_t n,
                   double* dst_dev);
/* ---- halo exchange over peer memory (one 8-GPU NVLink/NVSwitch box) ----------
 * The row-partitioned residual / Jacobian need u at ghost vertices.  Each rank
 * exports its local u arrays and a flag word per peer with CUDA IPC; a peer maps
 * them once (fvm_ipc_open) and then, every evaluation, PUSHES the owned values its
 * neighbour needs straight into that neighbour's ghost slots (P2P stores over
 * NVLink) and releases the epoch into the neighbour's flag; the neighbour's
 * fvm_halo_wait (a stream-ordered spin kernel with a timeout) acquires it before
 * the tile kernel runs.  No host round trip, no staging buffer, no NCCL call. */
int fvm_ipc_export(fvm_ctx* ctx, void* ptr_dev, unsigned char* handle64);
int fvm_ipc_open(fvm_ctx* ctx, const unsigned char* handle64, void** out_dev);
int fvm_ipc_close(fvm_ctx* ctx, void* ptr_dev);
/* peer_dst_dev[i] = src_dev[send_index_dev[i]], i < n; then (stream-ordered,
 * system-scope release) *peer_flag_dev = epoch.  peer_flag_dev may be NULL. */
int fvm_halo_push(fvm_ctx* ctx, const double* src_dev, const int32_t* send_index_dev, int64_t n,
                  double* peer_dst_dev, uint64_t* peer_flag_dev, uint64_t epoch);
/* the whole exchange in ONE launch: peers_dev is a device array of n_peers records
 *   { const int32_t* send_index_dev; double* peer_dst_dev[2]; uint64_t* peer_flag_dev; int64_t n; }
 * (40 bytes each; peer_dst_dev[b] = the peer's ghost slots in its u buffer b); values are
 * read from src_dev (this rank's u buffer `buffer`), and the last block of a peer to finish
 * releases `epoch` into that peer's flag.  max_n = the longest send list; counters_dev:
 * n_peers zero-initialised uint32 (re-armed by the kernel). */
int fvm_halo_push_all(fvm_ctx* ctx, const void* peers_dev, int n_peers, int64_t max_n,
                      const double* src_dev, int buffer, uint64_t epoch, uint32_t* counters_dev);
/* stream-ordered wait until every flags_dev[i] (device array of n_flags device
 * pointers) has reached epoch; after timeout_s seconds sets *timed_out_dev = 1 and
 * returns instead of hanging. */
int fvm_halo_wait(fvm_ctx* ctx, const uint64_t* const* flags_dev, int n_flags, uint64_t epoch,
                  double timeout_s, int32_t* timed_out_dev);
/* fvm_assemble (residual / Jacobian modes) with the halo wait FUSED into the tile
 * kernel: tiles are walked interior-first, and only a tile whose vertex table reaches
 * into the ghost vertices makes the kernel's producer acquire every flags_dev[i]
 * (>= epoch) before it copies u -- the NVLink transfer of the neighbours' pushes
 * overlaps the interior math instead of a separate wait kernel.  Results are the
 * same bits as fvm_halo_wait + fvm_assemble.  *timed_out_dev is set after ~20 s. */
int fvm_assemble_halo(fvm_mesh* mesh, fvm_kernel* k, const uint8_t* vmask_dev,
                      const uint8_t* dirid_dev, const double* u_dev, double* data_dev,
                      double* vec_dev, const uint64_t* const* flags_dev, int n_flags, uint64_t epoch,
                      int32_t* timed_out_dev);
/* fvm_halo_push_all + fvm_assemble_halo in one call: push this rank's boundary values of
 * u_dev (its u buffer `buffer`) to every neighbour, then launch the tile kernel that waits
 * for theirs only where a tile needs them. */
int fvm_halo_step(fvm_mesh* mesh, fvm_kernel* k, const uint8_t* vmask_dev, const uint8_t* dirid_dev,
                  const double* u_dev, double* data_dev, double* vec_dev,
                  const uint64_t* const* flags_dev, int n_flags, uint64_t epoch,
                  int32_t* timed_out_dev, const void* peers_dev, int n_peers, int64_t max_n,
                  int buffer, uint32_t* counters_dev);
/* y += a*x (Newton update on the device) */
int fvm_axpy(fvm_ctx* ctx, double a, const double* x_dev, double* y_dev, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
