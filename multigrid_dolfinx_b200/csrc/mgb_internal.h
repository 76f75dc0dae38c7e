// Internal declarations shared by the host-side setup code (mgb_setup.cpp) and the CUDA engine
// (mgb_engine.cu).  Nothing here is part of the C ABI (see include/mgb200.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mgb {

struct HostCsr {
    int64_t nrows = 0, ncols = 0;
    std::vector<int64_t> ip;   // nrows + 1
    std::vector<int32_t> ix;
    std::vector<double> ax;
    int64_t nnz() const { return ip.empty() ? 0 : ip.back(); }
    bool empty() const { return ip.empty(); }
};

// Copies a caller CSR (int32 or int64 row pointers) into a HostCsr.  Returns "" or an error message.
std::string import_csr(HostCsr& out, int64_t nrows, int64_t ncols, int64_t nnz, const void* indptr, int indptr_bytes,
                       const int32_t* indices, const double* values);

// getJacobiMatrices (multigrid.py:48-56): dinv_i = fl(1/a_ii); R_omega = off-diagonal entries with
// a_ij != 0.0, value fl(dinv_i * a_ij).  reversed: each row's entries in reverse stored order (the
// order scipy's DIA*CSR product leaves them in).  Returns false if a diagonal entry is zero/missing.
bool build_rj(const HostCsr& A, bool reversed, HostCsr& RJ, std::vector<double>& dinv);

// Off-diagonal nonzero part of A with unscaled values + the diagonal (Gauss-Seidel operator split).
bool split_offdiag(const HostCsr& A, HostCsr& G, std::vector<double>& diag);

// R = scale * P^T with sorted columns (what scipy's P.T.tocsr() + sort_indices gives).
void transpose_scaled(const HostCsr& P, double scale, HostCsr& R);

// Gauss-Seidel artefacts on the symmetrised nonzero graph {i~j : a_ij != 0 or a_ji != 0, i != j}.
//   level_sets: lev[i] = 0 if i has no neighbour j < i, else 1 + max lev[j]
//   colouring : first-fit greedy in natural row order
// order = stable sort of rows by level / colour; offsets has (count + 1) entries.
void level_sets(const HostCsr& A, std::vector<int32_t>& lev, std::vector<int32_t>& order, std::vector<int32_t>& offsets);
void greedy_colouring(const HostCsr& A, std::vector<int32_t>& col, std::vector<int32_t>& order, std::vector<int32_t>& offsets);

// rows of M permuted: out.row[p] = M.row[order[p]] (columns and in-row entry order untouched)
void permute_rows(const HostCsr& M, const std::vector<int32_t>& order, HostCsr& out);

// Dense inverse (row-major n*n) of a sparse matrix by Gauss-Jordan with partial pivoting.
// Returns false if singular.
bool dense_inverse(const HostCsr& A, std::vector<double>& inv);

// Row tiles for the tile kernel: tile t covers rows [tiles[t], tiles[t+1]); every tile satisfies
//   rowptr[tiles[t+1]] - (rowptr[tiles[t]] & ~7) <= cap   and   rows <= row_cap,
// and no tile straddles a breakpoint (breaks = sorted row indices, may be empty).
// row_align > 1 (only without breakpoints): every tile start is a multiple of row_align.
// Returns false if a single row exceeds cap.
bool make_tiles(const std::vector<int64_t>& ip, int64_t cap, int64_t row_cap, const std::vector<int32_t>& breaks,
                std::vector<int32_t>& tiles, std::vector<int32_t>* break_tile_index, int64_t row_align = 1);

}  // namespace mgb
