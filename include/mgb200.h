/* mgb200.h -- C ABI of the B200-native geometric-multigrid V-cycle engine (libmgb200.so).
 *
 * This is the drop-in boundary for the hot path of nikhilTkur/Multigrid_dolfinx: the V-cycle in
 * multigrid.py (V_cycle_scheme, multigrid.py:231-268) and what it calls.  The reference has no FFI
 * of its own (it is pure Python over scipy); each entry point below names the reference code it
 * replaces.  Signatures use plain pointers and sizes only (no torch / numpy types).
 *
 * Conventions
 *   - every function returns MGB_OK (0) or a negative MGB_ERR_* code; the message is available
 *     from mgb_last_error().  No C++ exception crosses this boundary.
 *   - host arrays passed to the mgb_set_* calls are copied during the call and stay caller-owned.
 *   - vector arguments carry a memory-kind flag: MGB_MEM_HOST pointers are staged through
 *     cudaMemcpyAsync inside the call (the call returns after the result is back on the host);
 *     MGB_MEM_DEVICE pointers are borrowed for the duration of the call.
 *   - one handle = one device + one stream; a handle is NOT thread-safe, distinct handles are
 *     independent.  Calls with device pointers only enqueue work on the handle's stream
 *     (mgb_get_stream) and do not synchronise unless they return host data.
 *   - levels are keyed by the reference's integer level (coarsest_level .. finest_level,
 *     multigrid.py:13-14; cells per dimension = c * 2^level, Multigrid_prototype.py:63).
 *   - all floating point is IEEE fp64; column indices are int32 (PETSc 32-bit build); row pointers
 *     may be handed over as int32 or int64.
 *   - there is NO CPU fallback: without a CUDA device mgb_create fails with MGB_ERR_CUDA.
 */
#ifndef MGB200_H
#define MGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_VERSION 100

#define MGB_OK 0
#define MGB_ERR_INVALID (-1)     /* bad argument                                  */
#define MGB_ERR_CUDA (-2)        /* CUDA runtime / driver error, or no device     */
#define MGB_ERR_STATE (-3)       /* call out of order (e.g. vcycle before finalize) */
#define MGB_ERR_NOMEM (-4)
#define MGB_ERR_SINGULAR (-5)    /* zero diagonal / singular coarsest matrix      */
#define MGB_ERR_UNSUPPORTED (-6)
#define MGB_ERR_COMM (-7)        /* NCCL error                                    */

typedef struct mgb_handle mgb_handle;

/* restriction modes.  INJECTION is what the reference V-cycle executes
 * (Restriction2D_direct, multigrid.py:123-132, called at multigrid.py:251-252); FULL_WEIGHTING is
 * the reference's unused Restriction2D (= 2^-d P^T, multigrid.py:135-198); TRANSPOSE is R = P^T
 * (north-star wording); EXPLICIT takes any CSR. */
enum { MGB_R_INJECTION = 0, MGB_R_FULL_WEIGHTING = 1, MGB_R_TRANSPOSE = 2, MGB_R_EXPLICIT = 3 };

/* smoothers.  JACOBI_RJ is the reference formula with the precomputed iteration matrix
 * (jacobiRelaxation, multigrid.py:223-228 + getJacobiMatrices, multigrid.py:48-56);
 * JACOBI_A is the algebraically equal single-matrix form v + w D^-1 (f - A v);
 * GS_LEVEL is natural-order forward Gauss-Seidel executed by level sets (no reference text);
 * GS_MULTICOLOR is Gauss-Seidel in greedy-colour order (different sweep order, documented option). */
enum { MGB_SM_JACOBI_RJ = 0, MGB_SM_JACOBI_A = 1, MGB_SM_GS_LEVEL = 2, MGB_SM_GS_MULTICOLOR = 3 };

enum { MGB_MEM_HOST = 0, MGB_MEM_DEVICE = 1 };

/* artefacts that must be bit-exact (mgb_get_artifact) */
enum {
    MGB_ART_RJ_INDPTR = 0,     /* int32[n+1]  R_omega row pointers  (multigrid.py:52-55)          */
    MGB_ART_RJ_INDICES = 1,    /* int32[nnz]                                                     */
    MGB_ART_RJ_VALUES = 2,     /* double[nnz] fl(fl(1/a_ii) * a_ij)                              */
    MGB_ART_DINV = 3,          /* double[n]   fl(1/a_ii)            (multigrid.py:53)             */
    MGB_ART_LEVEL_OF_ROW = 4,  /* int32[n]    Gauss-Seidel dependency level of each row           */
    MGB_ART_LEVEL_ORDER = 5,   /* int32[n]    execution order (stable sort by level)              */
    MGB_ART_LEVEL_OFFSETS = 6, /* int32[nlev+1]                                                   */
    MGB_ART_COLOUR_OF_ROW = 7, /* int32[n]    first-fit greedy colour                             */
    MGB_ART_COLOUR_ORDER = 8,  /* int32[n]                                                        */
    MGB_ART_COLOUR_OFFSETS = 9,/* int32[ncol+1]                                                   */
    MGB_ART_R_INDPTR = 10,     /* int32[n_c+1] restriction CSR built from P (FULL_WEIGHTING/TRANSPOSE) */
    MGB_ART_R_INDICES = 11,
    MGB_ART_R_VALUES = 12,
    MGB_ART_COARSE_INVERSE = 13,/* double[n_c*n_c] row-major dense inverse of the coarsest matrix  */
    MGB_ART_A_INDPTR = 14,     /* the level matrix as held on the device (checks the device-side generator) */
    MGB_ART_A_INDICES = 15,
    MGB_ART_A_VALUES = 16,
    MGB_ART_P_INDPTR = 17,     /* interpolation rows of this level (fine side)                     */
    MGB_ART_P_INDICES = 18,
    MGB_ART_P_VALUES = 19,
    MGB_ART_INJECTION = 20     /* int32[n_c] injection list of this level (fine side)              */
};
/* The lossless coding mgb_finalize built for an operator of the level (compare with mgb_host_code_operator):
 * kind = MGB_ART_CODE(op, part), op: 0 A, 1 R_omega, 2 P, 3 R;
 * part 0: int32[4] {mode, dictionary entries / patterns in use, table entries, number of codes};
 * part 1: uint8 codes (per row in modes 3 and 4, per stored entry in modes 1 and 2); part 2: table, 16-byte entries
 * {double value; int32 col_minus_row (mode 4: col minus the row's anchor column); int32 0}; part 3 (modes 3, 4): int32
 * {first entry, length} x 256.  MGB_ART_CODE_ANCHOR(op) (mode 4): int32[nrows], every row's anchor = its first stored column
 * (0 for an empty row). */
#define MGB_ART_CODE(op, part) (32 + 4 * (op) + (part))
#define MGB_ART_CODE_ANCHOR(op) (48 + (op))

/* per-level device buffers (mgb_level_buffer) */
enum { MGB_BUF_V = 0, MGB_BUF_F = 1, MGB_BUF_R = 2 };

/* kernel kinds reported by the profiler (mgb_profile_get) */
enum {
    MGB_K_JACOBI = 0, MGB_K_RESIDUAL = 1, MGB_K_RESTRICT = 2, MGB_K_PROLONG_ADD = 3, MGB_K_COARSE = 4,
    MGB_K_INIT_GUESS = 5, MGB_K_GS = 6, MGB_K_NORM = 7, MGB_K_SPMV = 8, MGB_K_HALO = 9, MGB_K_COPY = 10,
    MGB_K_JACOBI2 = 11,   /* two Jacobi sweeps in one launch (k_hotrow2) */
    MGB_K_COUNT = 12
};

typedef struct {
    int32_t kind;        /* MGB_K_*                                            */
    int32_t level;
    int64_t launches;
    double total_ms;     /* CUDA-event time summed over launches                */
    double bytes;        /* algorithmic bytes of ONE launch (CSR form) */
    double moved_bytes;  /* bytes the chosen kernel streams per launch: equal to `bytes` unless the operator is
                            dictionary-coded (option "compress"): 1 (or 5) instead of 12 bytes per stored entry, or 1 byte per row */
} mgb_profile_record;

/* ---- lifetime ---------------------------------------------------------------------------- */
int mgb_version(void);
/* replaces the module-global hierarchy container set by initialize_problem (multigrid.py:10-45) */
int mgb_create(mgb_handle** out, int device_id);
int mgb_destroy(mgb_handle* h);
/* h == NULL: message of the last failed mgb_create in this thread */
const char* mgb_last_error(const mgb_handle* h);

/* ---- hierarchy upload -------------------------------------------------------------------- */
/* A_l exactly as exported from PETSc into scipy (Multigrid_prototype.py:95-99: A_sp_dict[l]).
 * indptr_bytes is 4 (int32) or 8 (int64).  Explicit zeros are kept. */
int mgb_set_level(mgb_handle* h, int level, int64_t n, int64_t nnz, const void* indptr, int indptr_bytes,
                  const int32_t* indices, const double* values);
/* Transfer pair between coarse_level and coarse_level+1.
 * P (n_fine x n_coarse) is the matrix of Interpolation2D (multigrid.py:59-120); row entries are summed
 * in stored order.  r_mode INJECTION needs inj[n_coarse] = fine dof of each coarse dof
 * (Restriction2D_direct, multigrid.py:128-131); EXPLICIT needs the R CSR (n_coarse x n_fine);
 * FULL_WEIGHTING / TRANSPOSE build R from P at finalize.  dim_for_fw: 2 or 3 (scale 2^-dim). */
int mgb_set_transfer(mgb_handle* h, int coarse_level, int64_t n_fine, int64_t n_coarse,
                     int64_t p_nnz, const void* p_indptr, int p_indptr_bytes, const int32_t* p_indices, const double* p_values,
                     int r_mode, int dim_for_fw, const int32_t* inj,
                     int64_t r_nnz, const void* r_indptr, int r_indptr_bytes, const int32_t* r_indices, const double* r_values);
/* ---- row-sharded hierarchies: one process (and one handle) per GPU ------------------------------------ */
/* The reference is single-rank; these calls have no counterpart there.  Rank 0 creates an NCCL
 * unique id (>= 128 bytes), the host side broadcasts it (torch.distributed, MPI, ...), every rank calls
 * mgb_dist_init before uploading levels. */
int mgb_dist_unique_id(void* out, int capacity);
int mgb_dist_init(mgb_handle* h, int rank, int world, const void* unique_id, int id_bytes);
/* This rank's row block of A_l: n_owned rows, columns renumbered [owned 0..n_owned) | ghost n_owned..n_owned+n_ghost).
 * Vectors of that level hold n_owned + n_ghost entries on this rank.  Row entry order is untouched, so results
 * are bit-identical to the single-GPU engine. */
int mgb_set_level_local(mgb_handle* h, int level, int64_t n_owned, int64_t n_ghost, int64_t nnz, const void* indptr, int indptr_bytes,
                        const int32_t* indices, const double* values);
/* Halo plan of a sharded level: for every neighbour rank the owned local indices to send (concatenated) and the
 * number of ghost entries received from it (ghost slots are ordered neighbour after neighbour). */
int mgb_set_halo(mgb_handle* h, int level, int npeers, const int32_t* peer_ranks, const int32_t* send_counts,
                 const int32_t* send_indices, const int32_t* recv_counts);
/* Peer-memory halo exchange (NVLink, CUDA IPC) instead of ncclSend/ncclRecv: every rank exports a blob per sharded
 * level (call with blob == NULL to get its size), the host side hands each neighbour's blob to mgb_p2p_import.
 * Once all neighbours of a level are imported its halo exchange is two small kernels: a push that gathers and stores
 * straight into the neighbours' memory, and a pull that waits for their arrival flags and fills the ghost section. */
int mgb_p2p_export(mgb_handle* h, int level, void* blob, int capacity, int* size);
int mgb_p2p_import(mgb_handle* h, int level, int peer_rank, const void* blob, int size);
/* Declares `level` (and everything coarser) to live on rank 0 only.  offsets[world+1]: the slice of that level's
 * right-hand side each rank produces when restricting from level+1.  On rank 0 the level must already be set in
 * full (mgb_set_level); on other ranks it must not be set at all. */
int mgb_set_gather_level(mgb_handle* h, int level, int64_t n_global, const int64_t* offsets);

/* ---- device-side input generation (stand-in for the host assembly of Multigrid_prototype.py:62-118) ------ */
/* Generates rows [row_begin, row_end) of the synthetic P1 Poisson matrix (unit square / cube, lexicographic DOFs,
 * dolfinx-shaped pattern with stored zeros, Dirichlet rows) straight into device CSR -- bit-identical to
 * multigrid_dolfinx_b200.problems.stencil_p1.  Columns are numbered [owned | ghost_lo..row_begin | row_end..ghost_hi).
 * R_omega / D^-1 (multigrid.py:48-56) are then also built on the device at mgb_finalize. */
int mgb_synth_poisson_level(mgb_handle* h, int level, int dim, int cells_per_dim, int64_t row_begin, int64_t row_end,
                            int64_t ghost_lo, int64_t ghost_hi);
/* Interpolation rows (matrix of Interpolation2D, multigrid.py:59-120; tensor product in 3-D) of the generated level
 * coarse_level+1 and the injection list (multigrid.py:128-131) of the coarse rows [inj_coarse_begin, inj_coarse_end). */
int mgb_synth_poisson_transfer(mgb_handle* h, int coarse_level, int64_t inj_coarse_begin, int64_t inj_coarse_end);

/* mu1, mu2, omega (multigrid.py:19-21), smoother = MGB_SM_* */
int mgb_set_params(mgb_handle* h, double omega, int mu1, int mu2, int smoother);
/* named numeric options (defaults in brackets), see DESIGN.md.  Results are bit-identical under every one of them except
 * "rj_order" and "kernel_family" = 2 (a different, still deterministic summation order).
 *  before mgb_finalize:  "rj_order" [1] (0 = R_omega rows as stored in A, 1 = reversed = the order scipy's DIA*CSR product leaves),
 *   "compress" [3] (lossless operator codings, each verified on the device before use: 0 off, 1 one byte per stored entry + a
 *   dictionary, 2 + one byte per ROW where whole rows repeat, 3 + anchored row patterns for rectangular operators),
 *   "stage_x" [3] (kernel for row-pattern-coded operators: 3 k_hotrow, 1 k_rowwin, 0 k_rowstream), "stream_cfg" [3] (0: register-
 *   staged tile kernel only), "code_cfg", "hot_cfg", "anch_cfg", "win_cfg" (kernel shapes), "kernel_family" [0 auto],
 *   "lanes_per_row", "tile_iter", "stream_auto", "fuse_halo" [1] (row-sharded levels: halo exchange inside the kernels),
 *   "device_setup" [0] (1: P^T, level sets, colourings and Gauss-Seidel operators of host-assembled levels are built on the
 *   device too -- generated levels always are), "s2_min_rows" [2^20] (smallest level that runs two Jacobi sweeps per launch);
 *  any time:  "use_graph" [1], "pdl" [-1: programmatic dependent launch for the coded kernels; 1 also the CSR stream kernels; 0
 *   never], "fuse_restrict" [1], "reuse_g" [1], "hot_inj" [1], "hot_pf" [262144 rows of L2 prefetch distance], "fuse_sweeps" [0]
 *   (1: pairs of Jacobi sweeps in ONE launch on unsharded hot-row levels, k_hotrow2; measured no faster), "s2_slack", "s2_tiles", "anch_tiles" [8: consecutive tiles per CTA of the anchored-pattern kernel, k_anchloop], "coarse_refine" [0],
 *   "tail_rows" [0], "tail_cluster" [0], "gs_cluster" [2], "p2p_enable" [1], "overlap_halo" [0], "overlap_waves", "win_prefetch". */
int mgb_set_option(mgb_handle* h, const char* key, double value);
/* builds R_omega and D^-1 (getJacobiMatrices, multigrid.py:48-56), R from P, the dense inverse of the
 * coarsest matrix (replaces spsolve, multigrid.py:239), level sets / colours when a GS smoother is
 * selected, uploads everything and allocates the per-level vectors. */
int mgb_finalize(mgb_handle* h);
/* the argument / state checks of mgb_finalize alone: no device work, no communication.  Row-sharded callers run it on
 * every rank and agree on the outcome before the (collective) mgb_finalize. */
int mgb_precheck(mgb_handle* h);

/* ---- the hot path ------------------------------------------------------------------------ */
/* ncycles x V_cycle_scheme(A_jacobi_sp_dict[top_level], v, f) (multigrid.py:231-268), v updated in place.
 * resnorm_hist (HOST pointer, nullable): ||f - A v||_2 after each cycle (the quantity multigrid.py:291 forms). */
int mgb_vcycle(mgb_handle* h, int top_level, double* v, const double* f, int mem, int ncycles, double* resnorm_hist);
/* one cycle with the reference's test=True outputs (multigrid.py:262-266): f2h = restricted residual,
 * v2h = coarse-level result, err_h = interpolated correction (all of top_level / top_level-1 size). */
int mgb_vcycle_debug(mgb_handle* h, int top_level, double* v, const double* f, int mem,
                     double* f2h, double* v2h, double* err_h);
/* same as mgb_vcycle but v and f already live in the engine's own level buffers (mgb_level_buffer):
 * no copy of any kind inside the call; resnorm_hist (host, nullable) forces a sync at the end. */
int mgb_vcycle_resident(mgb_handle* h, int top_level, int ncycles, double* resnorm_hist);

/* ---- the caller of the path: FullMultiGrid (multigrid.py:271-307) kept on the device --------------------------- */
/* b_dict[level] (Multigrid_prototype.py:110): the re-discretised right-hand side FMG starts each level from. */
int mgb_set_rhs(mgb_handle* h, int level, const double* b, int mem);
/* optional mass matrix of a level: the stopping rule then uses sqrt(r^T M r), the L2(Omega) norm the reference
 * assembles with dolfinx (res_calculator, multigrid.py:203-208); without it the l2 norm is used. */
int mgb_set_mass_matrix(mgb_handle* h, int level, int64_t n, int64_t nnz, const void* indptr, int indptr_bytes,
                        const int32_t* indices, const double* values);
/* nested iteration: coarsest solve, interpolate up, mu0 V-cycles per intermediate level (multigrid.py:305-306), V-cycles on
 * the finest level until the residual norm <= tol (multigrid.py:296) or max_cycles (the reference has no cap).
 * v_out (nullable) receives the finest-level solution; resnorm_hist[0..min(cycles, hist_capacity)) the norms. */
int mgb_fmg(mgb_handle* h, int mu0, double tol, int max_cycles, double* v_out, int mem, int* cycles_done,
            double* resnorm_hist, int hist_capacity);
/* optional nodal vector of the exact solution on the finest level: mgb_fmg then also records, after EVERY finest-level cycle,
 * the norm of (v - u_exact) in the same norm as the residual -- the reference's error_per_V_cycle_finest list
 * (err_calculator, multigrid.py:213-218, appended per cycle at multigrid.py:292-293).  mgb_fmg_error_history returns the
 * list of the last mgb_fmg run (count = number of finest-level cycles, 0 when no exact solution was set). */
int mgb_set_exact_solution(mgb_handle* h, int level, const double* u_exact, int mem);
int mgb_fmg_error_history(mgb_handle* h, double* errnorm_hist, int capacity, int* count);

/* Restriction of a transfer that is already set (mgb_set_transfer) or generated (mgb_synth_poisson_transfer), before
 * mgb_finalize: MGB_R_INJECTION (Restriction2D_direct, multigrid.py:123-132; needs the injection list the transfer came with),
 * MGB_R_FULL_WEIGHTING (2^-dim_for_fw P^T, Restriction2D, multigrid.py:135-198) or MGB_R_TRANSPOSE (P^T).  The transposed
 * operator is formed at mgb_finalize: on the device (mgb_devsetup.cu) for generated levels and under the option
 * "device_setup" = 1, else on the host; both give the same arrays bit for bit (MGB_ART_R_*). */
int mgb_set_restriction(mgb_handle* h, int coarse_level, int r_mode, int dim_for_fw);

/* Fused halo exchange: whether the kernels of `level` exchange the ghost rows themselves.  Decided per rank at
 * mgb_finalize from what that rank's shard looks like -- and it MUST be the same on every rank (a rank that sends without a
 * neighbour that waits, or the reverse, dead-locks), so the caller agrees on it after mgb_finalize: query every rank, switch the
 * level off everywhere unless all ranks qualified (multigrid_dolfinx_b200/dist.py does this with one all-gather).
 * mgb_set_halo_fused can only switch a level OFF. */
int mgb_halo_fused(mgb_handle* h, int level, int* fused);
int mgb_set_halo_fused(mgb_handle* h, int level, int on);

/* Caller numbering.  The reference hands over operators and vectors in dolfinx's DOF numbering (Multigrid_prototype.py:68-74
 * records it as coordinate dicts), which is not lexicographic; the lossless row-pattern codings need a banded numbering.
 * new_index[i] = position of dof i in the numbering the engine should work in (a permutation of 0..n-1; the drop-in module
 * passes the lexicographic lattice index recovered from the coordinate dicts).  mgb_finalize renumbers every operator once --
 * rows moved, columns relabelled, the ENTRY ORDER inside every row kept, so every row sum adds the same products in the same
 * order and the iterates stay bit-identical -- and every vector crossing the ABI (v, f, right-hand sides, per-operator calls) is
 * permuted on the device on the way in and out.  Artefacts (mgb_get_artifact) and mgb_level_buffer show the engine's numbering.
 * Call after mgb_set_level of that level, before mgb_finalize; single-device hierarchies only. */
int mgb_set_numbering(mgb_handle* h, int level, int64_t n, const int64_t* new_index);

/* ---- per-operator entry points (parity tests, profiling) -------------------------------------- */
int mgb_spmv(mgb_handle* h, int level, const double* x, double* y, int mem);                       /* A.dot(x), multigrid.py:244 */
int mgb_residual(mgb_handle* h, int level, const double* v, const double* f, double* r, int mem); /* f - A v, multigrid.py:244  */
int mgb_smooth(mgb_handle* h, int level, double* v, const double* f, int nsweeps, int mem);       /* jacobiRelaxation, multigrid.py:223-228 (or GS) */
int mgb_restrict(mgb_handle* h, int fine_level, const double* r_fine, double* f_coarse, int mem); /* multigrid.py:251-252 */
int mgb_prolong_add(mgb_handle* h, int fine_level, const double* e_coarse, double* v_fine, int mem); /* multigrid.py:258-260 */
int mgb_coarse_solve(mgb_handle* h, const double* f, double* u, int mem);                          /* multigrid.py:238-241 */
int mgb_norm2(mgb_handle* h, int64_t n, const double* x, int mem, double* out_host);

/* ---- introspection ----------------------------------------------------------------------- */
int mgb_get_artifact(mgb_handle* h, int level, int kind, void* out, int64_t capacity_bytes, int64_t* size_bytes);
int mgb_level_buffer(mgb_handle* h, int level, int which, void** device_ptr, int64_t* n);
int mgb_get_stream(mgb_handle* h, void** cuda_stream);
int mgb_synchronize(mgb_handle* h);
int mgb_launch_count(mgb_handle* h, int64_t* kernels_launched);
/* event-timed, non-graph execution of everything between begin and end; records are per (kind, level) */
int mgb_profile_begin(mgb_handle* h);
int mgb_profile_end(mgb_handle* h);
int mgb_profile_get(mgb_handle* h, mgb_profile_record* out, int capacity, int* count);
/* algorithmic bytes of one V-cycle at top_level with the current parameters */
int mgb_vcycle_bytes(mgb_handle* h, int top_level, double* bytes);
/* the same sum with the bytes the chosen kernels actually stream (dictionary-coded operators move fewer) */
int mgb_vcycle_bytes_moved(mgb_handle* h, int top_level, double* bytes);
/* human-readable description of the kernel variant chosen for each (level, operator) */
int mgb_describe(mgb_handle* h, char* out, int64_t capacity);

/* ---- host-side setup routines (no device needed; what mgb_finalize runs internally) ------------ */
/* R_omega / D^-1 from A (multigrid.py:48-56).  Call with rj_indices == NULL to query rj_nnz only. */
int mgb_host_build_rj(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values, int reversed,
                      int64_t* rj_nnz, int32_t* rj_indptr, int32_t* rj_indices, double* rj_values, double* dinv);
int mgb_host_level_sets(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                        int32_t* level_of_row, int32_t* order, int64_t* nlevels, int32_t* offsets, int64_t offsets_capacity);
int mgb_host_colouring(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                       int32_t* colour_of_row, int32_t* order, int64_t* ncolours, int32_t* offsets, int64_t offsets_capacity);
int mgb_host_dense_inverse(int64_t n, const int64_t* indptr, const int32_t* indices, const double* values, double* inv_row_major);
/* Row tiles of the stream kernels (what mgb_finalize cuts every operator into): tile t = rows [tiles[t], tiles[t+1]) with
 * at most row_cap rows and indptr[tiles[t+1]] - (indptr[tiles[t]] & ~7) <= cap stored entries, never straddling a breakpoint
 * (sorted row indices), every tile start a multiple of row_align unless it is a breakpoint.  tiles: int32[*ntiles + 1];
 * break_tile: int32[nbreaks] tile index at each breakpoint.  MGB_ERR_UNSUPPORTED: a single row exceeds cap (or cannot be
 * aligned) -- the engine then falls back to the warp-per-row kernel. */
int mgb_host_make_tiles(int64_t n, const int64_t* indptr, int64_t cap, int64_t row_cap, int nbreaks, const int32_t* breaks,
                        int64_t row_align, int32_t* tiles, int64_t tiles_capacity, int64_t* ntiles, int32_t* break_tile);
/* The lossless operator coding of DESIGN.md 4.1 as a host routine: the definition of what mgb_finalize builds (and verifies
 * entry by entry) on the device.  *mode_out: 0 none, 1 pair codes, 2 value codes, 3 row patterns (tried only if allow_patterns
 * >= 1), 4 anchored row patterns -- columns measured from each row's first stored column (tried after mode 3 if allow_patterns
 * >= 2; the anchors themselves are indices[indptr[i]], 0 for an empty row).
 * codes: uint8[nrows] (modes 3, 4) or uint8[nnz] (modes 1, 2).  table: 16-byte entries {double value; int32 col_minus_row (mode 4:
 * col minus anchor); int32 0}, 256 of them (modes 1, 2) or *table_entries <= 2048 (modes 3, 4: every pattern padded to a multiple
 * of 8 entries).  pattern_head (modes 3, 4): int32 {first entry, length} x 256.  *ndict_out: dictionary entries / patterns in use.
 * codes, table, pattern_head may be NULL to query the mode and the counts only. */
int mgb_host_code_operator(int64_t nrows, int64_t ncols, const int64_t* indptr, const int32_t* indices, const double* values,
                           int allow_patterns, int* mode_out, int* ndict_out, uint8_t* codes, void* table, int* table_entries,
                           int32_t* pattern_head);

#ifdef __cplusplus
}
#endif
#endif /* MGB200_H */
