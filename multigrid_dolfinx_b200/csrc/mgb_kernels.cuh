// Device kernels of the V-cycle engine (sm_100a).  All of them are HBM-bound fp64 CSR row-sum
// kernels with a fused epilogue; see DESIGN.md for the byte counts and the roofline of each.
//
// Numerics contract: a row sum is ONE accumulator fed in stored entry order with separately rounded
// multiply and add (__dmul_rn / __dadd_rn: no FMA contraction) -- exactly what scipy's csr_matvec
// does on the reference's CPU path (multigrid.py:226, :244).  The "tile" family keeps that contract;
// the "sub-warp" family (shuffle tree) trades it for a different, still deterministic, order.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mgb {
namespace cg = cooperative_groups;

// ---- streaming loads: matrix arrays are read exactly once per launch.  sm_100 has 256-bit global loads;
// .L1::no_allocate keeps the stream out of L1 and .L2::evict_first keeps it from flushing x out of L2
// (SASS: LDG.E.NA.EFL2.256.CONSTANT) ------------------------------------------------------------------
struct I8 { int v[8]; };
struct D4 { double v[4]; };
__device__ __forceinline__ I8 ld_stream_i8(const int32_t* p)
{
    I8 r;
    asm("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ D4 ld_stream_d4(const double* p)
{
    D4 r;
    asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
        : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}

// ---- epilogues --------------------------------------------------------------------------------------
// An epilogue turns the row sum s of row i into the output.  load(i) fetches the per-row operands and is
// issued BEFORE the row sum is available (so their DRAM latency overlaps the matrix stream);
// store(i, s, pre) finishes the row.
// y = A x                                                   (multigrid.py:244, A.dot(v))
struct EpiStore {
    double* y;
    struct Pre {};
    __device__ __forceinline__ Pre load(int) const { return Pre{}; }
    __device__ __forceinline__ void store(int i, double s, const Pre&) const { y[i] = s; }
};
// r = f - A v                                               (multigrid.py:244)
struct EpiResidual {
    const double* f; double* r;
    struct Pre { double f; };
    __device__ __forceinline__ Pre load(int i) const { return Pre{f[i]}; }
    __device__ __forceinline__ void store(int i, double s, const Pre& p) const { r[i] = __dsub_rn(p.f, s); }
};
// weighted Jacobi, reference form (multigrid.py:226): out = ((1-w)*v + g) - w*s, g = w*(dinv*f)
struct EpiJacobiRJ {
    const double* v; const double* g; double* out; double om1, om;
    struct Pre { double v, g; };
    __device__ __forceinline__ Pre load(int i) const { return Pre{v[i], g[i]}; }
    __device__ __forceinline__ void store(int i, double s, const Pre& p) const
    {
        out[i] = __dsub_rn(__dadd_rn(__dmul_rn(om1, p.v), p.g), __dmul_rn(om, s));
    }
};
// same, first sweep of a relaxation call: also produces g (multigrid.py:226 recomputes w*(Dinv f) per sweep;
// the product is identical every time, so it is formed once and kept)
struct EpiJacobiRJFirst {
    const double* v; const double* dinv; const double* f; double* g; double* out; double om1, om;
    struct Pre { double v, g; };
    __device__ __forceinline__ Pre load(int i) const { return Pre{v[i], __dmul_rn(om, __dmul_rn(dinv[i], f[i]))}; }
    __device__ __forceinline__ void store(int i, double s, const Pre& p) const
    {
        g[i] = p.g;
        out[i] = __dsub_rn(__dadd_rn(__dmul_rn(om1, p.v), p.g), __dmul_rn(om, s));
    }
};
// single-matrix Jacobi: out = v + w*(dinv*(f - s)), s = (A v)_i
struct EpiJacobiA {
    const double* v; const double* dinv; const double* f; double* out; double om;
    struct Pre { double v, dinv, f; };
    __device__ __forceinline__ Pre load(int i) const { return Pre{v[i], dinv[i], f[i]}; }
    __device__ __forceinline__ void store(int i, double s, const Pre& p) const
    {
        out[i] = __dadd_rn(p.v, __dmul_rn(om, __dmul_rn(p.dinv, __dsub_rn(p.f, s))));
    }
};
// v = v + P e   (multigrid.py:258-260); err (nullable) receives P e (the test=True output, multigrid.py:265)
struct EpiProlongAdd {
    double* v; double* err;
    struct Pre { double v; };
    __device__ __forceinline__ Pre load(int i) const { return Pre{v[i]}; }
    __device__ __forceinline__ void store(int i, double s, const Pre& p) const
    {
        if (err) err[i] = s;
        v[i] = __dadd_rn(p.v, s);
    }
};
// Gauss-Seidel row update on a row-permuted off-diagonal operator: v[order[p]] = (f - s) / d
struct EpiGaussSeidel {
    const int32_t* order; const double* diag; const double* f; double* v;
    struct Pre { int i; double f, d; };
    __device__ __forceinline__ Pre load(int p) const { const int i = order[p]; return Pre{i, f[i], diag[p]}; }
    __device__ __forceinline__ void store(int, double s, const Pre& p) const { v[p.i] = __ddiv_rn(__dsub_rn(p.f, s), p.d); }
};

// ---- tile family ------------------------------------------------------------------------------------
// One CTA per row tile.  Phase 1 streams the tile's (cols, vals) with 256-bit loads (coalesced; ITER
// independent groups of 8 entries = 96 bytes per thread in flight), gathers x through L1/L2 and parks the
// products in shared memory.  Phase 2 is thread-per-row: each row is summed sequentially in stored order.
// The per-row epilogue operands and row pointers of the first RPT rows of every thread are fetched
// together with the matrix stream.  Shared index i is padded to i + (i >> 4) so that the stride-4 /
// stride-8 row starts of phase 2 fall into distinct 8-byte banks.
constexpr int TILE_ENT = 8;     // entries per thread per iteration
constexpr int TILE_RPT = 2;     // rows per thread with prefetched epilogue operands

template <int ITER, int THREADS>
struct TileCfg {
    static constexpr int CAP = TILE_ENT * THREADS * ITER;          // entries staged per tile
    static constexpr int SMEM_DOUBLES = CAP + CAP / 16 + 8;
};

__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

// NCX: x is read-only for the whole launch -> gather it through the non-coherent path (__ldg).
// NCX = false is used when the epilogue writes into x itself (in-place Gauss-Seidel colours).
template <bool NCX>
__device__ __forceinline__ double ld_x(const double* x, int c)
{
    if constexpr (NCX) return __ldg(x + c);
    else return x[c];
}

template <int ITER, int THREADS, bool NCX, class Epi>
__global__ void __launch_bounds__(THREADS)
k_tile(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, const double* __restrict__ vals,
       const int32_t* __restrict__ tile_rows, int tile_base, const double* x, Epi epi)
{
    using Cfg = TileCfg<ITER, THREADS>;
    __shared__ double prod[Cfg::SMEM_DOUBLES];
    const int t = tile_base + blockIdx.x;
    const int row0 = tile_rows[t], row1 = tile_rows[t + 1];
    const int nz0 = rowptr[row0], nz1 = rowptr[row1];
    const int nz0a = nz0 & ~(TILE_ENT - 1);

    I8 c[ITER];
    D4 va[ITER], vb[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int k = nz0a + TILE_ENT * (threadIdx.x + it * THREADS);
        if (k < nz1) {
            c[it] = ld_stream_i8(cols + k);
            va[it] = ld_stream_d4(vals + k);
            vb[it] = ld_stream_d4(vals + k + 4);
        }
    }
    // epilogue operands of this thread's rows: independent of the stream above, issued right behind it
    int ra[TILE_RPT], rb[TILE_RPT];
    typename Epi::Pre pre[TILE_RPT];
#pragma unroll
    for (int j = 0; j < TILE_RPT; ++j) {
        const int r = row0 + threadIdx.x + j * THREADS;
        if (r < row1) {
            ra[j] = rowptr[r]; rb[j] = rowptr[r + 1];
            pre[j] = epi.load(r);
        }
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int k = nz0a + TILE_ENT * (threadIdx.x + it * THREADS);
        if (k < nz1) {
            double xv[TILE_ENT];
#pragma unroll
            for (int e = 0; e < TILE_ENT; ++e) xv[e] = ld_x<NCX>(x, c[it].v[e]);
            const int b = pad16(k - nz0a);          // k - nz0a is a multiple of 8: the 8 entries share one pad offset
#pragma unroll
            for (int e = 0; e < 4; ++e) prod[b + e] = __dmul_rn(va[it].v[e], xv[e]);
#pragma unroll
            for (int e = 0; e < 4; ++e) prod[b + 4 + e] = __dmul_rn(vb[it].v[e], xv[4 + e]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TILE_RPT; ++j) {
        const int r = row0 + threadIdx.x + j * THREADS;
        if (r < row1) {
            double s = 0.0;
            for (int k = ra[j] - nz0a; k < rb[j] - nz0a; ++k) s = __dadd_rn(s, prod[pad16(k)]);
            epi.store(r, s, pre[j]);
        }
    }
    for (int r = row0 + threadIdx.x + TILE_RPT * THREADS; r < row1; r += THREADS) {     // tiles of many short rows
        const int a = rowptr[r] - nz0a, b = rowptr[r + 1] - nz0a;
        const typename Epi::Pre p = epi.load(r);
        double s = 0.0;
        for (int k = a; k < b; ++k) s = __dadd_rn(s, prod[pad16(k)]);
        epi.store(r, s, p);
    }
}

// ---- sub-warp family --------------------------------------------------------------------------------
// LPR lanes cooperate on one row (LPR = 32: warp per row), partial sums combined with a shuffle tree.
template <int LPR, bool NCX, class Epi>
__global__ void __launch_bounds__(256)
k_subwarp(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, const double* __restrict__ vals,
          int row_begin, int row_end, const double* x, Epi epi)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = gid % LPR;
    int row = row_begin + gid / LPR;
    const bool active = row < row_end;
    if (!active) row = row_end - 1;
    const int a = rowptr[row], b = rowptr[row + 1];
    typename Epi::Pre pre;
    if (active && lane == 0) pre = epi.load(row);
    double s = 0.0;
    for (int k = a + lane; k < b; k += LPR) s = __dadd_rn(s, __dmul_rn(vals[k], ld_x<NCX>(x, cols[k])));
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (active && lane == 0) epi.store(row, s, pre);
}

// ---- small kernels ----------------------------------------------------------------------------------
// zero initial guess + one Jacobi sweep collapses to v = g = w*(dinv*f)   (multigrid.py:253 + :226)
__global__ void k_init_guess(int n, const double* __restrict__ dinv, const double* __restrict__ f, double om,
                             double* __restrict__ g, double* __restrict__ v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double gi = __dmul_rn(om, __dmul_rn(dinv[i], f[i]));
        g[i] = gi;
        v[i] = gi;
    }
}

// injection: out[i] = r[inj[i]]   (Restriction2D_direct, multigrid.py:128-131)
__global__ void k_gather(int nc, const int32_t* __restrict__ inj, const double* __restrict__ r, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nc) out[i] = __ldg(r + inj[i]);
}

// residual evaluated only at the injected rows: out[i] = f[g] - (A v)[g], g = inj[i]   (multigrid.py:244 + :128-131)
template <int LPR>
__global__ void __launch_bounds__(256)
k_residual_injected(int nc, const int32_t* __restrict__ inj, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
                    const double* __restrict__ vals, const double* __restrict__ f, const double* __restrict__ v, double* __restrict__ out)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = gid % LPR;
    int i = gid / LPR;
    const bool active = i < nc;
    if (!active) i = nc - 1;
    const int row = inj[i];
    const int a = rowptr[row], len = rowptr[row + 1] - a;
    const int maxlen = __reduce_max_sync(0xffffffffu, len);      // warp-uniform trip count (shuffles inside)
    const int base = (threadIdx.x & 31) & ~(LPR - 1);
    // sequential order is kept: every lane of the group accumulates the same ordered sum; the lanes
    // only share the loads
    double s = 0.0;
    for (int o = 0; o < maxlen; o += LPR) {
        const int k = a + o + lane;
        const double p = (o + lane < len) ? __dmul_rn(vals[k], __ldg(v + cols[k])) : 0.0;
#pragma unroll
        for (int l = 0; l < LPR; ++l) {
            const double q = __shfl_sync(0xffffffffu, p, base + l);
            if (o + l < len) s = __dadd_rn(s, q);
        }
    }
    if (active && lane == 0) out[i] = __dsub_rn(f[row], s);
}

// dense coarsest apply: y = M x (+ y0), warp per row, M row-major n x n     (replaces spsolve, multigrid.py:239)
__global__ void __launch_bounds__(256)
k_dense_gemv(int n, const double* __restrict__ M, const double* __restrict__ x, const double* __restrict__ y0, double* __restrict__ y)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double* row = M + (size_t)warp * n;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int j = lane;
    for (; j + 96 < n; j += 128) {
        s0 = fma(row[j], x[j], s0);
        s1 = fma(row[j + 32], x[j + 32], s1);
        s2 = fma(row[j + 64], x[j + 64], s2);
        s3 = fma(row[j + 96], x[j + 96], s3);
    }
    for (; j < n; j += 32) s0 = fma(row[j], x[j], s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[warp] = y0 ? y0[warp] + s : s;
}

// ||x||_2: fixed grid, fixed trees -> bitwise reproducible
__global__ void __launch_bounds__(256)
k_sumsq_partial(int64_t n, const double* __restrict__ x, double* __restrict__ partial)
{
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fma(x[i], x[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void k_sumsq_final(int nb, const double* __restrict__ partial, double* __restrict__ out)
{
    __shared__ double sh[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        *out = sqrt(t);
    }
}

// level-scheduled forward Gauss-Seidel: one cooperative launch per sweep, grid-wide barrier between
// dependency levels.  G = off-diagonal nonzeros of A with rows in level-major order; a row only reads
// x entries finalised in earlier levels (new) or in later levels (old), so the result equals the
// sequential natural-order sweep.  x is read with ld.cg (L2) because other CTAs write it.
__global__ void __launch_bounds__(256)
k_gs_levels(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, const double* __restrict__ vals,
            const int32_t* __restrict__ order, const double* __restrict__ diag, const double* __restrict__ f,
            double* v, const int32_t* __restrict__ lvl_off, int nlev)
{
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    for (int L = 0; L < nlev; ++L) {
        const int p0 = lvl_off[L], p1 = lvl_off[L + 1];
        for (int p = p0 + tid; p < p1; p += nth) {
            const int a = rowptr[p], b = rowptr[p + 1];
            double s = 0.0;
            for (int k = a; k < b; ++k) s = __dadd_rn(s, __dmul_rn(vals[k], __ldcg(v + cols[k])));
            const int i = order[p];
            __stcg(v + i, __ddiv_rn(__dsub_rn(f[i], s), diag[p]));
        }
        grid.sync();
    }
}

}  // namespace mgb
