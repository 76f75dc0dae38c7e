// Hierarchy set-up on the device, part 2 (mgb_devsetup.cu): the transposed restriction, the Gauss-Seidel level sets and
// colourings and the reordered Gauss-Seidel operator, built from the CSR arrays already resident in HBM.  Every function
// reproduces its host definition in mgb_setup.cpp (transpose_scaled, level_sets, greedy_colouring, split_offdiag +
// permute_rows) BIT FOR BIT -- tests/test_gpu_parity.py compares the artefacts of both builds on ragged inputs -- and exists so
// that levels generated on the device (config 5: 135 M rows, no host copy) get the same operators as host-assembled ones.
// Not part of the C ABI.  All functions return cudaSuccess or the first CUDA error; arrays returned through int32_t** /
// double** are cudaMalloc'ed here and owned by the caller.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace mgb {
namespace dev {

// R = scale * P^T, columns of every row ascending, duplicates in P's storage order (= mgb_setup.cpp::transpose_scaled).
// P is nrows x ncols with nnz stored entries.  R's arrays carry the engine's padding (rowptr + 8, cols / vals + 16, zeroed).
cudaError_t transpose_scaled(cudaStream_t s, int64_t nrows, int64_t ncols, int64_t nnz, const int32_t* rp, const int32_t* cols,
                             const double* vals, double scale, int32_t** t_rp, int32_t** t_cols, double** t_vals);

// Lower part of the symmetrised nonzero graph {i ~ j : i != j, j < n, a_ij != 0.0 or a_ji != 0.0}: for every row the neighbours
// with a smaller index (with repetitions, order unspecified -- only the SET enters the level sets and colours).
cudaError_t lower_sym_graph(cudaStream_t s, int n, const int32_t* rp, const int32_t* cols, const double* vals,
                            int32_t** lp, int32_t** lx);

// lev[i] = 0 if i has no neighbour j < i, else 1 + max lev[j]: relaxation passes over all rows until a pass changes nothing
// (values only grow and never exceed the answer, so in-place updates are safe).  passes: how many it took.
cudaError_t level_sets(cudaStream_t s, int n, const int32_t* lp, const int32_t* lx, int32_t** lev, int* passes);

// First-fit greedy colouring in natural row order, as a fixed point: col[i] = smallest colour no neighbour j < i holds.  The
// fixed point is unique (induction over i) and a row of dependency level k is final after pass k + 1.  overflow != 0: a row
// needed more than 128 colours (not supported on the device).
cudaError_t colouring(cudaStream_t s, int n, const int32_t* lp, const int32_t* lx, int32_t** col, int* passes, int* overflow);

// Stable sort of the rows by key (level or colour): order (device, n entries), offsets (host, nkeys + 1 entries).
cudaError_t order_from_keys(cudaStream_t s, int n, const int32_t* keys, int32_t** order, std::vector<int32_t>& offsets);

// Gauss-Seidel operator: rows of A in `order`, diagonal and explicit zeros dropped, entry order kept; diag[p] = the LAST stored
// diagonal entry of row order[p] (as mgb_setup.cpp::split_offdiag).  bad != 0: a zero or missing diagonal.
cudaError_t gs_operator(cudaStream_t s, int n, const int32_t* order, const int32_t* rp, const int32_t* cols, const double* vals,
                        int32_t** g_rp, int32_t** g_cols, double** g_vals, double** g_diag, int64_t* g_nnz, int* bad);

// ELL copy of the Gauss-Seidel operator (W x n, entry-slot major, zero filled) for the pipelined level-scheduled kernel.
cudaError_t gs_ell(cudaStream_t s, int n, int W, const int32_t* g_rp, const int32_t* g_cols, const double* g_vals,
                   int32_t** e_cols, double** e_vals);

}  // namespace dev
}  // namespace mgb
