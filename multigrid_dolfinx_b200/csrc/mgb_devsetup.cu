// Hierarchy set-up on the device, part 2: see mgb_devsetup.h.  Host definitions: mgb_setup.cpp (transpose_scaled, lower_sym_graph,
// level_sets, greedy_colouring, order_from_keys, split_offdiag, permute_rows); reference slot: the reference builds none of these
// (its restriction is P^T formed by scipy, multigrid.py:135-198, and it has no Gauss-Seidel), DESIGN.md section 5.
#include "mgb_devsetup.h"

#include <cub/cub.cuh>

#include <algorithm>

namespace mgb {
namespace dev {

namespace {

#define DCU(call)                            \
    do {                                     \
        cudaError_t e_ = (call);             \
        if (e_ != cudaSuccess) return e_;    \
    } while (0)

constexpr int TPB = 256;
inline int blocks_for(int64_t n) { return (int)((n + TPB - 1) / TPB); }

template <class T>
cudaError_t zalloc(cudaStream_t s, T** p, size_t count)
{
    *p = nullptr;
    DCU(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return cudaMemsetAsync(*p, 0, std::max<size_t>(count, 1) * sizeof(T), s);
}

// counts (n + 1 entries, the last one 0) -> row pointers, in place
cudaError_t exclusive_scan(cudaStream_t s, int32_t* cnt, int64_t n_plus_1)
{
    void* tmp = nullptr; size_t bytes = 0;
    DCU(cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, cnt, (int)n_plus_1, s));
    DCU(cudaMalloc(&tmp, std::max<size_t>(bytes, 1)));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, cnt, (int)n_plus_1, s);
    cudaStreamSynchronize(s);
    cudaFree(tmp);
    return e;
}

// ---- transpose ------------------------------------------------------------------------------------------------------------
__global__ void k_tr_count(int64_t nnz, const int32_t* __restrict__ cols, int32_t* __restrict__ cnt)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) atomicAdd(&cnt[cols[k]], 1);
}

// entry d of R comes from stored entry k = src[d] of P: (row of P that holds k, value * scale)
__global__ void k_tr_gather(int64_t nnz, int nrows, const int32_t* __restrict__ rp, const double* __restrict__ vals, double scale,
                            const int32_t* __restrict__ src, int32_t* __restrict__ t_cols, double* __restrict__ t_vals)
{
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nnz) return;
    const int k = src[d];
    int lo = 0, hi = nrows;                                 // the row i of P with rp[i] <= k < rp[i + 1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (rp[mid] <= k) lo = mid; else hi = mid; }
    t_cols[d] = lo;
    t_vals[d] = __dmul_rn(vals[k], scale);
}

__global__ void k_iota(int n, int32_t* a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

// ---- symmetrised lower graph ----------------------------------------------------------------------------------------------
__global__ void k_lsg_count(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
                            int32_t* __restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        const int j = cols[k];
        if (j == i || j >= n || vals[k] == 0.0) continue;   // (j >= n: a ghost column of a row block, never a dependency)
        atomicAdd(&cnt[max(i, j)], 1);
    }
}

__global__ void k_lsg_fill(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, const double* __restrict__ vals,
                           const int32_t* __restrict__ lp, int32_t* __restrict__ pos, int32_t* __restrict__ lx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        const int j = cols[k];
        if (j == i || j >= n || vals[k] == 0.0) continue;
        const int hi = max(i, j), lo = min(i, j);
        lx[lp[hi] + atomicAdd(&pos[hi], 1)] = lo;
    }
}

// ---- level sets / colouring: relaxation passes -------------------------------------------------------------------------
__global__ void k_levset_pass(int n, const int32_t* __restrict__ lp, const int32_t* __restrict__ lx, int32_t* lev, int* changed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int m = -1;
    for (int k = lp[i]; k < lp[i + 1]; ++k) m = max(m, ((volatile int32_t*)lev)[lx[k]]);
    if (m + 1 != lev[i]) { lev[i] = m + 1; *changed = 1; }
}

__global__ void k_colour_pass(int n, const int32_t* __restrict__ lp, const int32_t* __restrict__ lx, int32_t* col, int* changed,
                              int* overflow)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long u0 = 0, u1 = 0;
    for (int k = lp[i]; k < lp[i + 1]; ++k) {
        const int c = ((volatile int32_t*)col)[lx[k]];
        if (c < 64) u0 |= 1ull << c;
        else if (c < 128) u1 |= 1ull << (c - 64);
    }
    int c;
    if (~u0) c = __ffsll((long long)~u0) - 1;
    else if (~u1) c = 64 + __ffsll((long long)~u1) - 1;
    else { c = 127; *overflow = 1; }
    if (c != col[i]) { col[i] = c; *changed = 1; }
}

template <class Launch>
cudaError_t relax_until_fixed(cudaStream_t s, int n, Launch&& pass, int* passes)
{
    int* flag = nullptr;                                  // [0]: scratch for the passes nobody looks at, [1]: the batch's last pass
    DCU(zalloc(s, &flag, 2));
    cudaError_t rc = cudaSuccess;
    int done = 0;
    const int batch = 16;
    for (int64_t guard = 0; guard <= (int64_t)n + batch; guard += batch) {
        if ((rc = cudaMemsetAsync(flag, 0, 2 * sizeof(int), s)) != cudaSuccess) break;
        for (int b = 0; b < batch; ++b) pass(flag + (b == batch - 1));
        int h = 1;
        if ((rc = cudaMemcpyAsync(&h, flag + 1, sizeof(int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        if ((rc = cudaStreamSynchronize(s)) != cudaSuccess) break;
        done += batch;
        if (!h) break;
    }
    cudaFree(flag);
    if (passes) *passes = done;
    return rc;
}

__global__ void k_hist(int n, const int32_t* __restrict__ keys, int32_t* __restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cnt[keys[i] + 1], 1);
}

// ---- Gauss-Seidel operator ---------------------------------------------------------------------------------------------
__global__ void k_gs_count(int n, const int32_t* __restrict__ order, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols,
                           const double* __restrict__ vals, int32_t* __restrict__ cnt, double* __restrict__ diag, int* bad)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = order[p];
    double d = 0.0; bool have = false; int c = 0;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        if (cols[k] == i) { d = vals[k]; have = true; }     // the last stored diagonal wins (host and oracle do the same)
        else if (vals[k] != 0.0) ++c;
    }
    if (!have || d == 0.0) *bad = 1;
    diag[p] = d;
    cnt[p] = c;
}

__global__ void k_gs_fill(int n, const int32_t* __restrict__ order, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols,
                          const double* __restrict__ vals, const int32_t* __restrict__ g_rp, int32_t* __restrict__ g_cols,
                          double* __restrict__ g_vals)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = order[p];
    int o = g_rp[p];
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        if (cols[k] == i || vals[k] == 0.0) continue;
        g_cols[o] = cols[k]; g_vals[o] = vals[k]; ++o;
    }
}

__global__ void k_gs_ell(int n, const int32_t* __restrict__ g_rp, const int32_t* __restrict__ g_cols, const double* __restrict__ g_vals,
                         int32_t* __restrict__ e_cols, double* __restrict__ e_vals)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int b = g_rp[p], e = g_rp[p + 1];
    for (int k = b; k < e; ++k) {
        e_cols[(size_t)(k - b) * n + p] = g_cols[k];
        e_vals[(size_t)(k - b) * n + p] = g_vals[k];
    }
}

}  // namespace

cudaError_t transpose_scaled(cudaStream_t s, int64_t nrows, int64_t ncols, int64_t nnz, const int32_t* rp, const int32_t* cols,
                             const double* vals, double scale, int32_t** t_rp, int32_t** t_cols, double** t_vals)
{
    // Stable sort of the stored entries by column: entries of one column stay in storage order, i.e. rows of P ascending and
    // duplicates in the order they are stored -- exactly the order the sequential transpose leaves (one pass over the rows of P
    // appending to the rows of R).  Row lengths of R do not matter (a dense column of P costs no more than a sparse one).
    DCU(zalloc(s, t_rp, (size_t)ncols + 1 + 8));
    DCU(zalloc(s, t_cols, (size_t)nnz + 16));
    DCU(zalloc(s, t_vals, (size_t)nnz + 16));
    if (nnz == 0 || nrows == 0 || ncols == 0) return cudaStreamSynchronize(s);
    int32_t *iota = nullptr, *keys = nullptr, *src = nullptr;
    void* tmp = nullptr; size_t bytes = 0;
    cudaError_t rc = cudaSuccess;
    auto step = [&](cudaError_t e) { if (rc == cudaSuccess) rc = e; return rc == cudaSuccess; };
    k_tr_count<<<blocks_for(nnz), TPB, 0, s>>>(nnz, cols, *t_rp);
    step(exclusive_scan(s, *t_rp, ncols + 1));
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < ncols) ++end_bit;
    if (step(zalloc(s, &iota, (size_t)nnz)) && step(zalloc(s, &keys, (size_t)nnz)) && step(zalloc(s, &src, (size_t)nnz))) {
        k_iota<<<blocks_for(nnz), TPB, 0, s>>>((int)nnz, iota);
        if (step(cub::DeviceRadixSort::SortPairs(tmp, bytes, cols, keys, iota, src, (int)nnz, 0, end_bit, s)) &&
            step(cudaMalloc(&tmp, std::max<size_t>(bytes, 1))) &&
            step(cub::DeviceRadixSort::SortPairs(tmp, bytes, cols, keys, iota, src, (int)nnz, 0, end_bit, s)))
            k_tr_gather<<<blocks_for(nnz), TPB, 0, s>>>(nnz, (int)nrows, rp, vals, scale, src, *t_cols, *t_vals);
    }
    step(cudaGetLastError());
    step(cudaStreamSynchronize(s));
    cudaFree(iota); cudaFree(keys); cudaFree(src); cudaFree(tmp);
    return rc;
}

cudaError_t lower_sym_graph(cudaStream_t s, int n, const int32_t* rp, const int32_t* cols, const double* vals, int32_t** lp, int32_t** lx)
{
    int32_t* pos = nullptr;
    *lx = nullptr;
    DCU(zalloc(s, lp, (size_t)n + 1));
    if (n > 0) k_lsg_count<<<blocks_for(n), TPB, 0, s>>>(n, rp, cols, vals, *lp);
    DCU(exclusive_scan(s, *lp, (int64_t)n + 1));
    int32_t total = 0;
    DCU(cudaMemcpyAsync(&total, *lp + n, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    DCU(cudaStreamSynchronize(s));
    DCU(zalloc(s, lx, (size_t)total));
    DCU(zalloc(s, &pos, (size_t)n));
    if (n > 0) k_lsg_fill<<<blocks_for(n), TPB, 0, s>>>(n, rp, cols, vals, *lp, pos, *lx);
    cudaError_t rc = cudaGetLastError();
    if (rc == cudaSuccess) rc = cudaStreamSynchronize(s);
    cudaFree(pos);
    return rc;
}

cudaError_t level_sets(cudaStream_t s, int n, const int32_t* lp, const int32_t* lx, int32_t** lev, int* passes)
{
    DCU(zalloc(s, lev, (size_t)n));
    if (passes) *passes = 0;
    if (n == 0) return cudaSuccess;
    int32_t* l = *lev;
    DCU(relax_until_fixed(s, n, [&](int* flag) { k_levset_pass<<<blocks_for(n), TPB, 0, s>>>(n, lp, lx, l, flag); }, passes));
    return cudaGetLastError();
}

cudaError_t colouring(cudaStream_t s, int n, const int32_t* lp, const int32_t* lx, int32_t** col, int* passes, int* overflow)
{
    DCU(zalloc(s, col, (size_t)n));
    if (passes) *passes = 0;
    *overflow = 0;
    if (n == 0) return cudaSuccess;
    int* ovf = nullptr;
    DCU(zalloc(s, &ovf, 1));
    int32_t* c = *col;
    cudaError_t rc = relax_until_fixed(s, n, [&](int* flag) { k_colour_pass<<<blocks_for(n), TPB, 0, s>>>(n, lp, lx, c, flag, ovf); }, passes);
    if (rc == cudaSuccess) rc = cudaGetLastError();
    if (rc == cudaSuccess) rc = cudaMemcpyAsync(overflow, ovf, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (rc == cudaSuccess) rc = cudaStreamSynchronize(s);
    cudaFree(ovf);
    return rc;
}

cudaError_t order_from_keys(cudaStream_t s, int n, const int32_t* keys, int32_t** order, std::vector<int32_t>& offsets)
{
    offsets.assign(1, 0);
    DCU(zalloc(s, order, (size_t)n));
    if (n == 0) return cudaSuccess;
    int32_t *iota = nullptr, *sorted = nullptr, *cnt = nullptr;
    void* tmp = nullptr; size_t bytes = 0;
    cudaError_t rc = cudaSuccess;
    auto step = [&](cudaError_t e) { if (rc == cudaSuccess) rc = e; return rc == cudaSuccess; };
    if (step(zalloc(s, &iota, (size_t)n)) && step(zalloc(s, &sorted, (size_t)n))) {
        k_iota<<<blocks_for(n), TPB, 0, s>>>(n, iota);
        // least-significant-digit radix sort: stable, i.e. rows of one level / colour stay in ascending order
        if (step(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys, sorted, iota, *order, n, 0, 32, s)) &&
            step(cudaMalloc(&tmp, std::max<size_t>(bytes, 1))))
            step(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys, sorted, iota, *order, n, 0, 32, s));
        int32_t last = 0;
        if (step(cudaMemcpyAsync(&last, sorted + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s)) && step(cudaStreamSynchronize(s))) {
            const size_t nk = (size_t)last + 1;
            if (step(zalloc(s, &cnt, nk + 1))) {
                k_hist<<<blocks_for(n), TPB, 0, s>>>(n, keys, cnt);
                offsets.assign(nk + 1, 0);
                if (step(cudaMemcpyAsync(offsets.data(), cnt, (nk + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s)) &&
                    step(cudaStreamSynchronize(s)))
                    for (size_t c = 0; c < nk; ++c) offsets[c + 1] += offsets[c];
            }
        }
    }
    step(cudaGetLastError());
    cudaFree(iota); cudaFree(sorted); cudaFree(cnt); cudaFree(tmp);
    return rc;
}

cudaError_t gs_operator(cudaStream_t s, int n, const int32_t* order, const int32_t* rp, const int32_t* cols, const double* vals,
                        int32_t** g_rp, int32_t** g_cols, double** g_vals, double** g_diag, int64_t* g_nnz, int* bad)
{
    int* dbad = nullptr;
    *g_cols = nullptr; *g_vals = nullptr; *bad = 0;
    DCU(zalloc(s, g_rp, (size_t)n + 1 + 8));
    DCU(zalloc(s, g_diag, (size_t)n));
    DCU(zalloc(s, &dbad, 1));
    if (n > 0) k_gs_count<<<blocks_for(n), TPB, 0, s>>>(n, order, rp, cols, vals, *g_rp, *g_diag, dbad);
    cudaError_t rc = exclusive_scan(s, *g_rp, (int64_t)n + 1);
    int32_t total = 0;
    if (rc == cudaSuccess) rc = cudaMemcpyAsync(&total, *g_rp + n, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    if (rc == cudaSuccess) rc = cudaMemcpyAsync(bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (rc == cudaSuccess) rc = cudaStreamSynchronize(s);
    cudaFree(dbad);
    DCU(rc);
    *g_nnz = total;
    DCU(zalloc(s, g_cols, (size_t)total + 16));
    DCU(zalloc(s, g_vals, (size_t)total + 16));
    if (n > 0) k_gs_fill<<<blocks_for(n), TPB, 0, s>>>(n, order, rp, cols, vals, *g_rp, *g_cols, *g_vals);
    DCU(cudaGetLastError());
    return cudaStreamSynchronize(s);
}

cudaError_t gs_ell(cudaStream_t s, int n, int W, const int32_t* g_rp, const int32_t* g_cols, const double* g_vals,
                   int32_t** e_cols, double** e_vals)
{
    DCU(zalloc(s, e_cols, (size_t)W * n));
    DCU(zalloc(s, e_vals, (size_t)W * n));
    if (n > 0) k_gs_ell<<<blocks_for(n), TPB, 0, s>>>(n, g_rp, g_cols, g_vals, *e_cols, *e_vals);
    DCU(cudaGetLastError());
    return cudaStreamSynchronize(s);
}

}  // namespace dev
}  // namespace mgb
