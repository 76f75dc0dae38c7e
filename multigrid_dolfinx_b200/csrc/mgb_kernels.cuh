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
// An epilogue turns the row sum s of row i into the output.  CONTIG epilogues read NOPS per-row operands
// from contiguous f64 arrays (operand(j)[i]); the kernels fetch them BEFORE the row sum is available (the
// tile kernel into registers, the stream kernel with the same bulk copies as the matrix), so their DRAM
// latency overlaps the matrix stream.  store(i, s, o) finishes the row.
// y = A x                                                   (multigrid.py:244, A.dot(v))
struct EpiStore {
    static constexpr int NOPS = 0; static constexpr bool CONTIG = true;
    double* y;
    __host__ __device__ __forceinline__ const double* operand(int) const { return nullptr; }
    __device__ __forceinline__ double store(int i, double s, const double*) const { y[i] = s; return s; }
};
// r = f - A v                                               (multigrid.py:244)
struct EpiResidual {
    static constexpr int NOPS = 1; static constexpr bool CONTIG = true;
    const double* f; double* r;
    __host__ __device__ __forceinline__ const double* operand(int) const { return f; }
    __device__ __forceinline__ double store(int i, double s, const double* o) const { const double t = __dsub_rn(o[0], s); r[i] = t; return t; }
};
// L2 eviction priorities for the two-sweep kernel (k_hotrow2): 0 none, 1 keep (evict_last), 2 stream (evict_first)
__device__ __forceinline__ unsigned long long l2_policy(int kind)
{
    unsigned long long p = 0;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <int KIND>
__device__ __forceinline__ void st_l2(double* q, double v, unsigned long long pol)
{
    if constexpr (KIND == 0) *q = v;
    else asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(q), "d"(v), "l"(pol) : "memory");
}
template <int KIND>
__device__ __forceinline__ double ld_stream_l2(const double* q, unsigned long long pol)
{
    double v;
    if constexpr (KIND == 0) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(q));
    else asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(q), "l"(pol));
    return v;
}

// weighted Jacobi, reference form (multigrid.py:226): out = ((1-w)*v + g) - w*s, g = w*(dinv*f)
struct EpiJacobiRJ {
    static constexpr int NOPS = 2; static constexpr bool CONTIG = true; static constexpr int XOP = 0;
    const double* v; const double* g; double* out; double om1, om;
    __host__ __device__ __forceinline__ const double* operand(int j) const { return j == 0 ? v : g; }
    __device__ __forceinline__ double store(int i, double s, const double* o) const
    {
        const double t = __dsub_rn(__dadd_rn(__dmul_rn(om1, o[0]), o[1]), __dmul_rn(om, s));
        out[i] = t;
        return t;
    }
    template <int KOUT>                                      // same, the result stored with an L2 eviction priority
    __device__ __forceinline__ void store_p(int i, double s, const double* o, unsigned long long pout, unsigned long long) const
    {
        st_l2<KOUT>(out + i, __dsub_rn(__dadd_rn(__dmul_rn(om1, o[0]), o[1]), __dmul_rn(om, s)), pout);
    }
};
// same, first sweep of a relaxation call: also produces g (multigrid.py:226 recomputes w*(Dinv f) per sweep;
// the product is identical every time, so it is formed once and kept)
struct EpiJacobiRJFirst {
    static constexpr int NOPS = 3; static constexpr bool CONTIG = true; static constexpr int XOP = 0;
    const double* v; const double* dinv; const double* f; double* g; double* out; double om1, om;
    __host__ __device__ __forceinline__ const double* operand(int j) const { return j == 0 ? v : (j == 1 ? dinv : f); }
    __device__ __forceinline__ double store(int i, double s, const double* o) const
    {
        const double gi = __dmul_rn(om, __dmul_rn(o[1], o[2]));
        g[i] = gi;
        const double t = __dsub_rn(__dadd_rn(__dmul_rn(om1, o[0]), gi), __dmul_rn(om, s));
        out[i] = t;
        return t;
    }
    template <int KOUT>                                      // (g is read again by the second sweep: always kept)
    __device__ __forceinline__ void store_p(int i, double s, const double* o, unsigned long long pout, unsigned long long pkeep) const
    {
        const double gi = __dmul_rn(om, __dmul_rn(o[1], o[2]));
        if constexpr (KOUT == 0) g[i] = gi; else st_l2<1>(g + i, gi, pkeep);
        st_l2<KOUT>(out + i, __dsub_rn(__dadd_rn(__dmul_rn(om1, o[0]), gi), __dmul_rn(om, s)), pout);
    }
};
// single-matrix Jacobi: out = v + w*(dinv*(f - s)), s = (A v)_i
struct EpiJacobiA {
    static constexpr int NOPS = 3; static constexpr bool CONTIG = true; static constexpr int XOP = 0;
    const double* v; const double* dinv; const double* f; double* out; double om;
    __host__ __device__ __forceinline__ const double* operand(int j) const { return j == 0 ? v : (j == 1 ? dinv : f); }
    __device__ __forceinline__ double store(int i, double s, const double* o) const
    {
        const double t = __dadd_rn(o[0], __dmul_rn(om, __dmul_rn(o[1], __dsub_rn(o[2], s))));
        out[i] = t;
        return t;
    }
};
// v = v + P e   (multigrid.py:258-260); err (nullable) receives P e (the test=True output, multigrid.py:265)
struct EpiProlongAdd {
    static constexpr int NOPS = 1; static constexpr bool CONTIG = true;
    double* v; double* err;
    __host__ __device__ __forceinline__ const double* operand(int) const { return v; }
    __device__ __forceinline__ double store(int i, double s, const double* o) const
    {
        if (err) err[i] = s;
        const double t = __dadd_rn(o[0], s);
        v[i] = t;
        return t;
    }
};
// Gauss-Seidel row update on a row-permuted off-diagonal operator: v[order[p]] = (f - s) / d
struct EpiGaussSeidel {
    static constexpr int NOPS = 0; static constexpr bool CONTIG = false;
    const int32_t* order; const double* diag; const double* f; double* v;
    struct Pre { int i; double f, d; };
    __device__ __forceinline__ Pre load(int p) const { const int i = order[p]; return Pre{i, f[i], diag[p]}; }
    __device__ __forceinline__ void store(int, double s, const Pre& p) const { v[p.i] = __ddiv_rn(__dsub_rn(p.f, s), p.d); }
};

// fused residual + injection: out[cmap[i]] = f[i] - s for the fine rows that have a coarse image
// (multigrid.py:244 followed by Restriction2D_direct, multigrid.py:128-131); cmap[i] = coarse dof or -1
struct EpiResidualInject {
    static constexpr int NOPS = 1; static constexpr bool CONTIG = true; static constexpr int NIOPS = 1;
    const double* f; const int32_t* cmap; double* out;
    __host__ __device__ __forceinline__ const double* operand(int) const { return f; }
    __device__ __forceinline__ const int32_t* ioperand() const { return cmap; }
    __device__ __forceinline__ double store_i(int, double s, const double* o, int c) const { const double t = __dsub_rn(o[0], s); if (c >= 0) out[c] = t; return t; }
};

template <class Epi, class = void> struct EpiNI { static constexpr int value = 0; };
template <class Epi> struct EpiNI<Epi, decltype((void)Epi::NIOPS)> { static constexpr int value = Epi::NIOPS; };

template <class Epi>
struct EpiOperands { double o[Epi::NOPS > 0 ? Epi::NOPS : 1]; int io; };

template <class Epi>
__device__ __forceinline__ auto epi_load(const Epi& e, int r)
{
    if constexpr (Epi::CONTIG) {
        EpiOperands<Epi> p;
#pragma unroll
        for (int j = 0; j < Epi::NOPS; ++j) p.o[j] = e.operand(j)[r];
        if constexpr (EpiNI<Epi>::value > 0) p.io = e.ioperand()[r];
        return p;
    } else {
        return e.load(r);
    }
}
template <class Epi, class P>
__device__ __forceinline__ void epi_store(const Epi& e, int r, double s, const P& p)
{
    if constexpr (!Epi::CONTIG) e.store(r, s, p);
    else if constexpr (EpiNI<Epi>::value > 0) e.store_i(r, s, p.o, p.io);
    else e.store(r, s, p.o);
}

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

// x is gathered with ordinary (coherent, L1-cached) loads, never through the non-coherent path (__ldg / ld.global.nc): under
// programmatic dependent launch a kernel is resident while its predecessors still run, and .nc requires the location to be
// read-only for the kernel's whole lifetime -- x is what the predecessor writes.  With ping-pong iterates a stale line left
// in L1 by the sweep before the predecessor was actually observed (k_hotinj on small levels, three kernels co-resident).
// NCX = false additionally marks launches whose epilogue writes into x itself (in-place Gauss-Seidel colours).
template <bool NCX>
__device__ __forceinline__ double ld_x(const double* x, int c)
{
    return x[c];
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
    decltype(epi_load(epi, 0)) pre[TILE_RPT];
#pragma unroll
    for (int j = 0; j < TILE_RPT; ++j) {
        const int r = row0 + threadIdx.x + j * THREADS;
        if (r < row1) {
            ra[j] = rowptr[r]; rb[j] = rowptr[r + 1];
            pre[j] = epi_load(epi, r);
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
            epi_store(epi, r, s, pre[j]);
        }
    }
    for (int r = row0 + threadIdx.x + TILE_RPT * THREADS; r < row1; r += THREADS) {     // tiles of many short rows
        const int a = rowptr[r] - nz0a, b = rowptr[r + 1] - nz0a;
        const auto p = epi_load(epi, r);
        double s = 0.0;
        for (int k = a; k < b; ++k) s = __dadd_rn(s, prod[pad16(k)]);
        epi_store(epi, r, s, p);
    }
}

// ---- stream family: TMA bulk-copy pipeline ------------------------------------------------------------
// Persistent CTAs (grid = SMs x CTAs/SM).  One producer lane walks this CTA's tiles and, per tile, issues
// 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) of the tile's cols / vals /
// row-pointer slice and of the epilogue's operand slices into a ring of STAGES shared-memory stages, up to
// STAGES tiles ahead of the consumers; matrix bytes carry an L2 evict-first policy.  THREADS consumer
// threads wait on the stage's "full" mbarrier, gather x, write products to a padded product buffer
// (phase A), sum each row sequentially in stored order and run the epilogue (phase B), then hand the stage
// back through its "empty" mbarrier.  Bytes in flight per SM = (STAGES-1) x stage size, independent of
// what the consumers are doing -- that is what the register-staged tile kernel cannot do.
// Requirements (checked on the host): tile starts are multiples of 4 rows, tiles hold <= EPT*THREADS entries
// and a quarter as many rows, operand arrays are 16-byte aligned and padded by >= 2 doubles.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    const uint32_t addr = smem_u32(b);
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000LL) __trap();      // ~2 s: a lost copy must fault, never hang the GPU
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

template <int THREADS, int EPT, int NOPS, int NIOPS = 0>
struct StreamCfg {
    static constexpr int CAP = EPT * THREADS;            // entries per tile
    static constexpr int ROWCAP = CAP / 4;               // rows per tile
    static constexpr int HDR_BYTES = 128;
    static constexpr int COLS_BYTES = CAP * 4;
    static constexpr int VALS_BYTES = CAP * 8;
    static constexpr int RP_BYTES = ((ROWCAP + 4) * 4 + 127) / 128 * 128;
    static constexpr int OP_BYTES = ROWCAP * 8;
    static constexpr int IOP_BYTES = ROWCAP * 4;
    static constexpr int STAGE_BYTES = HDR_BYTES + COLS_BYTES + VALS_BYTES + RP_BYTES + NOPS * OP_BYTES + NIOPS * IOP_BYTES;
    static constexpr int PROD_BYTES = ((CAP + CAP / 16 + 8) * 8 + 127) / 128 * 128;
    static constexpr int BAR_BYTES = 128;
    static constexpr int smem_bytes(int stages) { return BAR_BYTES + PROD_BYTES + stages * STAGE_BYTES; }
};

template <int THREADS, int EPT, int STAGES, bool NCX, class Epi>
__global__ void __launch_bounds__(THREADS + 32)
k_stream(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, const double* __restrict__ vals,
         const int4* __restrict__ desc, int ntiles, int tpc, const double* x, Epi epi)
{
    // tpc == 0: persistent CTAs, tiles dealt round-robin (tile = blockIdx.x + i * gridDim.x).
    // tpc  > 0: every CTA owns tpc consecutive tiles and then retires, so that a concurrently launched
    //           high-priority kernel (the NCCL halo exchange) finds a free slot within a few microseconds.
    static_assert(Epi::CONTIG, "stream kernel needs contiguous epilogue operands");
    static_assert(STAGES <= 8, "barrier block holds 8 stages");
    constexpr int NIOPS = EpiNI<Epi>::value;
    using Cfg = StreamCfg<THREADS, EPT, Epi::NOPS, NIOPS>;
    static_assert(EPT % 2 == 0, "entries are handled in pairs");
    constexpr int NJ = EPT / 2;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    double* prod = reinterpret_cast<double*>(smem + Cfg::BAR_BYTES);
    unsigned char* stage0 = smem + Cfg::BAR_BYTES + Cfg::PROD_BYTES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = tpc > 0 ? (int)blockIdx.x * tpc : (int)blockIdx.x;
    const int step = tpc > 0 ? 1 : (int)gridDim.x;
    const int my_tiles = tpc > 0 ? min(tpc, ntiles - first) : (ntiles - first + step - 1) / step;

    // Programmatic dependent launch: let the next kernel of the stream start its prologue while this one drains, and
    // -- when this kernel was itself launched that way -- do everything that does not depend on the predecessor's
    // output (barrier init above, the matrix part of the first STAGES tiles below) BEFORE waiting for it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (tid >= THREADS) {                                   // ---- producer warp (one lane works)
        if (tid == THREADS) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            auto issue_matrix = [&](int i, int4& d) {
                const int s = i % STAGES;
                d = __ldg(desc + first + (size_t)i * step);                          // {row0, nrows, nz0a, nent}
                unsigned char* st = stage0 + (size_t)s * Cfg::STAGE_BYTES;
                *reinterpret_cast<int4*>(st) = d;                                    // published by the arrive below (release)
                const uint32_t b_cols = (uint32_t)d.w * 4u, b_vals = (uint32_t)d.w * 8u;
                const uint32_t b_rp = (uint32_t)((d.y + 1 + 3) & ~3) * 4u, b_op = (uint32_t)((d.y + 1) & ~1) * 8u;
                const uint32_t b_iop = (uint32_t)((d.y + 3) & ~3) * 4u;
                mbar_expect_tx(full + s, b_cols + b_vals + b_rp + (uint32_t)Epi::NOPS * b_op + (uint32_t)NIOPS * b_iop);
                unsigned char* p = st + Cfg::HDR_BYTES;
                bulk_g2s_hint(p, cols + d.z, b_cols, full + s, pol);  p += Cfg::COLS_BYTES;
                bulk_g2s_hint(p, vals + d.z, b_vals, full + s, pol);  p += Cfg::VALS_BYTES;
                bulk_g2s(p, rowptr + d.x, b_rp, full + s);
            };
            auto issue_operands = [&](int i, const int4& d) {                        // slices of vectors the predecessor may have written
                const int s = i % STAGES;
                const uint32_t b_op = (uint32_t)((d.y + 1) & ~1) * 8u, b_iop = (uint32_t)((d.y + 3) & ~3) * 4u;
                unsigned char* p = stage0 + (size_t)s * Cfg::STAGE_BYTES + Cfg::HDR_BYTES + Cfg::COLS_BYTES + Cfg::VALS_BYTES + Cfg::RP_BYTES;
#pragma unroll
                for (int j = 0; j < Epi::NOPS; ++j) { bulk_g2s(p, epi.operand(j) + d.x, b_op, full + s); p += Cfg::OP_BYTES; }
                if constexpr (NIOPS > 0) bulk_g2s(p, epi.ioperand() + d.x, b_iop, full + s);
            };
            int4 dpre[STAGES];
            const int npre = my_tiles < STAGES ? my_tiles : STAGES;
            for (int i = 0; i < npre; ++i) issue_matrix(i, dpre[i]);
            asm volatile("griddepcontrol.wait;" ::: "memory");                       // predecessor's results are visible from here on
            for (int i = 0; i < npre; ++i) issue_operands(i, dpre[i]);
            for (int i = npre; i < my_tiles; ++i) {
                const int s = i % STAGES;
                mbar_wait(empty + s, (uint32_t)((i / STAGES - 1) & 1));
                int4 d;
                issue_matrix(i, d);
                issue_operands(i, d);
            }
        }
        return;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");      // consumers gather x and store results: strictly after the predecessor

    for (int i = 0; i < my_tiles; ++i) {                     // ---- consumers
        const int s = i % STAGES;
        mbar_wait(full + s, (uint32_t)((i / STAGES) & 1));
        const unsigned char* st = stage0 + (size_t)s * Cfg::STAGE_BYTES;
        const int4 d = *reinterpret_cast<const int4*>(st);
        const int32_t* scols = reinterpret_cast<const int32_t*>(st + Cfg::HDR_BYTES);
        const double* svals = reinterpret_cast<const double*>(st + Cfg::HDR_BYTES + Cfg::COLS_BYTES);
        const int32_t* srp = reinterpret_cast<const int32_t*>(st + Cfg::HDR_BYTES + Cfg::COLS_BYTES + Cfg::VALS_BYTES);
        const double* sops = reinterpret_cast<const double*>(st + Cfg::HDR_BYTES + Cfg::COLS_BYTES + Cfg::VALS_BYTES + Cfg::RP_BYTES);
        // phase A: entries (2*tid, 2*tid+1) + j*2*THREADS: 16-byte / 8-byte shared loads contiguous across the warp
        int2 c[NJ];
        double2 v[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int e = j * 2 * THREADS + 2 * tid;
            if (e < d.w) {
                c[j] = *reinterpret_cast<const int2*>(scols + e);
                v[j] = *reinterpret_cast<const double2*>(svals + e);
            }
        }
        double xa[NJ], xb[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int e = j * 2 * THREADS + 2 * tid;
            if (e < d.w) { xa[j] = ld_x<NCX>(x, c[j].x); xb[j] = ld_x<NCX>(x, c[j].y); }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int e = j * 2 * THREADS + 2 * tid;
            if (e < d.w) {
                const int b = pad16(e);
                prod[b] = __dmul_rn(v[j].x, xa[j]);
                prod[b + 1] = __dmul_rn(v[j].y, xb[j]);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
        // phase B: thread per row, sequential sum in stored order
        for (int r = tid; r < d.y; r += THREADS) {
            const int a = srp[r] - d.z, b = srp[r + 1] - d.z;
            double sum = 0.0;
            for (int k = a; k < b; ++k) sum = __dadd_rn(sum, prod[pad16(k)]);
            double o[Epi::NOPS > 0 ? Epi::NOPS : 1];
#pragma unroll
            for (int j = 0; j < Epi::NOPS; ++j) o[j] = sops[j * Cfg::ROWCAP + r];
            if constexpr (NIOPS > 0) epi.store_i(d.x + r, sum, o, reinterpret_cast<const int32_t*>(sops + Epi::NOPS * Cfg::ROWCAP)[r]);
            else epi.store(d.x + r, sum, o);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
        if (tid == 0) mbar_arrive(empty + s);
    }
}

// ---- row-stream family: dictionary-coded operators ------------------------------------------------------
// On the uniform meshes the reference works on, an operator's stored entries repeat: a handful of distinct values
// and -- for square operators in a banded numbering -- a handful of distinct column offsets (col - row).  Such an
// operator is kept a second time as ONE BYTE per stored entry, an index into a dictionary of <= 256
// (col - row, value) pairs (MODE 1), or as the int32 columns plus one byte per value (MODE 2: transfer operators,
// unstructured numberings).  The coding is lossless and verified entry by entry on the device when it is built
// (mgb_code.cuh); row sums are formed exactly as before -- one accumulator, stored order, separately rounded multiply
// and add -- so results stay bit-identical while the matrix stream shrinks from 12 to 1 (or 5) bytes per entry.
//
// Same producer / consumer structure as k_stream (one lane issues bulk copies of the tile's codes, columns,
// row-pointer slice and epilogue operand slices into a ring of stages), but the consumers are thread-per-row: with
// <= ~16 entries per row there is nothing to balance, consecutive rows gather consecutive x entries (coalesced), and
// no product buffer or mid-tile barrier is needed.  Every consumer warp arrives on the stage's "empty" barrier
// itself, so warps drift apart freely.
struct DictEnt { double val; int delta; int pad; };

template <int THREADS, int RPT, int EPR, int NOPS, int NIOPS, int MODE>
struct RowCfg {
    static constexpr int ROWCAP = THREADS * RPT;         // rows per tile
    static constexpr int ENTCAP = ROWCAP * EPR;          // stored entries per tile (from a 16-entry aligned start)
    static constexpr int HDR_BYTES = 128;
    static constexpr int CODE_BYTES = MODE == 3 ? ROWCAP : ENTCAP;       // MODE 3: one code per ROW
    static constexpr int COLS_BYTES = MODE == 2 ? ENTCAP * 4 : 0;
    static constexpr int RP_BYTES = MODE == 3 ? 0 : ((ROWCAP + 4) * 4 + 127) / 128 * 128;
    static constexpr int OP_BYTES = ROWCAP * 8;
    static constexpr int IOP_BYTES = ROWCAP * 4;
    static constexpr int STAGE_BYTES = HDR_BYTES + CODE_BYTES + COLS_BYTES + RP_BYTES + NOPS * OP_BYTES + NIOPS * IOP_BYTES;
    static constexpr int DICT_BYTES = 256 * (int)sizeof(DictEnt);        // MODE 3: 256 pattern heads (int2) + the pattern entries
    static constexpr int PHEAD_BYTES = 256 * 8;
    static constexpr int BAR_BYTES = 128;
    static constexpr int smem_bytes(int stages, int pent_bytes = 0)
    {
        return BAR_BYTES + (MODE == 3 ? PHEAD_BYTES + pent_bytes : DICT_BYTES) + stages * STAGE_BYTES;
    }
    static_assert(ENTCAP % 128 == 0 && ROWCAP % 32 == 0, "stage sections must stay 128-byte aligned");
};

// W entries of one row, starting at stage-local entry k0 (< b): all W gathers are issued unconditionally (entries past
// the row's end repeat its last entry -- a cache hit -- so that no load waits behind a predicate), then the products
// are added to the running sum in stored order.
template <int W, int MODE>
__device__ __forceinline__ double coded_chunk(double sum, int k0, int b, int row, const unsigned char* scodes, const int32_t* scols,
                                              const DictEnt* sdict, const double* x)
{
    double xv[W], vv[W];
    int col[W];
#pragma unroll
    for (int e = 0; e < W; ++e) {                          // shared-memory look-ups first ...
        const int k = min(k0 + e, b - 1);
        const DictEnt de = sdict[scodes[k]];
        if constexpr (MODE == 1) col[e] = row + de.delta; else col[e] = scols[k];
        vv[e] = de.val;
    }
#pragma unroll
    for (int e = 0; e < W; ++e) xv[e] = ld_x<true>(x, col[e]);   // ... then the gathers back to back
#pragma unroll
    for (int e = 0; e < W; ++e)
        if (k0 + e < b) sum = __dadd_rn(sum, __dmul_rn(vv[e], xv[e]));
    return sum;
}

// MINB: resident CTAs per SM the register budget is sized for (4 x 288 threads -> 56 registers; 5 -> 40, enough when
// only RPT * JW = 8 gathers are kept in flight)
template <int THREADS, int RPT, int EPR, int STAGES, int MODE, int JW, int MINB, class Epi>
__global__ void __launch_bounds__(THREADS + 32, MINB)
k_rowstream(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, const unsigned char* __restrict__ codes,
            const DictEnt* __restrict__ dict, const int2* __restrict__ phead, int npent,
            const int4* __restrict__ desc, int ntiles, int tpc, const double* x, Epi epi)
{
    static_assert(Epi::CONTIG, "row-stream kernel needs contiguous epilogue operands");
    static_assert(STAGES <= 8, "barrier block holds 8 stages");
    static_assert(MODE >= 1 && MODE <= 3, "MODE 1: pair codes, MODE 2: value codes + columns, MODE 3: row-pattern codes");
    static_assert(JW == 4 || JW == 8, "first-chunk width");
    constexpr int NIOPS = EpiNI<Epi>::value;
    using Cfg = RowCfg<THREADS, RPT, EPR, Epi::NOPS, NIOPS, MODE>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    // MODE 1/2: [barriers | 256 dictionary entries | stages]; MODE 3: [barriers | 256 pattern heads | npent pattern entries | stages]
    DictEnt* sdict = reinterpret_cast<DictEnt*>(smem + Cfg::BAR_BYTES + (MODE == 3 ? Cfg::PHEAD_BYTES : 0));
    const int2* sphead = reinterpret_cast<const int2*>(smem + Cfg::BAR_BYTES);
    const int ndict = MODE == 3 ? npent : 256;               // npent is a multiple of 8: the stages stay 128-byte aligned
    unsigned char* stage0 = smem + Cfg::BAR_BYTES + (MODE == 3 ? Cfg::PHEAD_BYTES + npent * (int)sizeof(DictEnt) : Cfg::DICT_BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int k = tid; k < ndict; k += THREADS + 32) sdict[k] = dict[k];
    if constexpr (MODE == 3)
        for (int k = tid; k < 256; k += THREADS + 32) const_cast<int2*>(sphead)[k] = phead[k];
    __syncthreads();
    const int first = tpc > 0 ? (int)blockIdx.x * tpc : (int)blockIdx.x;
    const int step = tpc > 0 ? 1 : (int)gridDim.x;
    const int my_tiles = tpc > 0 ? min(tpc, ntiles - first) : (ntiles - first + step - 1) / step;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (tid >= THREADS) {                                   // ---- producer warp (one lane works)
        if (tid == THREADS) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            auto issue_matrix = [&](int i, const int4& d) {                          // d = {row0, nrows, nz0a, nent}
                const int s = i % STAGES;
                unsigned char* st = stage0 + (size_t)s * Cfg::STAGE_BYTES;
                *reinterpret_cast<int4*>(st) = d;                                    // published by the arrive below (release)
                // MODE 3: one code per row, rows [d.x, d.x + d.y), d.x a multiple of 16; no row pointers are needed
                const uint32_t b_code = MODE == 3 ? (uint32_t)((d.y + 15) & ~15) : (uint32_t)d.w, b_cols = MODE == 2 ? (uint32_t)d.w * 4u : 0u;
                const uint32_t b_rp = MODE == 3 ? 0u : (uint32_t)((d.y + 1 + 3) & ~3) * 4u, b_op = (uint32_t)((d.y + 1) & ~1) * 8u;
                const uint32_t b_iop = (uint32_t)((d.y + 3) & ~3) * 4u;
                mbar_expect_tx(full + s, b_code + b_cols + b_rp + (uint32_t)Epi::NOPS * b_op + (uint32_t)NIOPS * b_iop);
                unsigned char* p = st + Cfg::HDR_BYTES;
                if (b_code) bulk_g2s_hint(p, codes + (MODE == 3 ? d.x : d.z), b_code, full + s, pol);
                p += Cfg::CODE_BYTES;
                if constexpr (MODE == 2) { if (b_cols) bulk_g2s_hint(p, cols + d.z, b_cols, full + s, pol);  p += Cfg::COLS_BYTES; }
                if constexpr (MODE != 3) bulk_g2s(p, rowptr + d.x, b_rp, full + s);
            };
            auto issue_operands = [&](int i, const int4& d) {                        // slices of vectors the predecessor may have written
                const int s = i % STAGES;
                const uint32_t b_op = (uint32_t)((d.y + 1) & ~1) * 8u, b_iop = (uint32_t)((d.y + 3) & ~3) * 4u;
                unsigned char* p = stage0 + (size_t)s * Cfg::STAGE_BYTES + Cfg::HDR_BYTES + Cfg::CODE_BYTES + Cfg::COLS_BYTES + Cfg::RP_BYTES;
#pragma unroll
                for (int j = 0; j < Epi::NOPS; ++j) { bulk_g2s(p, epi.operand(j) + d.x, b_op, full + s); p += Cfg::OP_BYTES; }
                if constexpr (NIOPS > 0) bulk_g2s(p, epi.ioperand() + d.x, b_iop, full + s);
            };
            int4 dpre[STAGES];
            const int npre = my_tiles < STAGES ? my_tiles : STAGES;
            for (int i = 0; i < npre; ++i) dpre[i] = __ldg(desc + first + (size_t)i * step);
            int4 dn = make_int4(0, 0, 0, 0);                                         // descriptor of the next tile, fetched one tile ahead
            if (npre < my_tiles) dn = __ldg(desc + first + (size_t)npre * step);
            for (int i = 0; i < npre; ++i) issue_matrix(i, dpre[i]);
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int i = 0; i < npre; ++i) issue_operands(i, dpre[i]);
            for (int i = npre; i < my_tiles; ++i) {
                const int s = i % STAGES;
                const int4 d = dn;
                if (i + 1 < my_tiles) dn = __ldg(desc + first + (size_t)(i + 1) * step);   // in flight while the stage drains
                mbar_wait(empty + s, (uint32_t)((i / STAGES - 1) & 1));
                issue_matrix(i, d);
                issue_operands(i, d);
            }
        }
        return;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");

    for (int i = 0; i < my_tiles; ++i) {                     // ---- consumers: thread per row, RPT rows per thread
        const int s = i % STAGES;
        mbar_wait(full + s, (uint32_t)((i / STAGES) & 1));
        const unsigned char* st = stage0 + (size_t)s * Cfg::STAGE_BYTES;
        const int4 d = *reinterpret_cast<const int4*>(st);
        const unsigned char* scodes = st + Cfg::HDR_BYTES;
        const int32_t* scols = reinterpret_cast<const int32_t*>(st + Cfg::HDR_BYTES + Cfg::CODE_BYTES);
        const int32_t* srp = reinterpret_cast<const int32_t*>(st + Cfg::HDR_BYTES + Cfg::CODE_BYTES + Cfg::COLS_BYTES);
        const double* sops = reinterpret_cast<const double*>(st + Cfg::HDR_BYTES + Cfg::CODE_BYTES + Cfg::COLS_BYTES + Cfg::RP_BYTES);
        if constexpr (MODE == 3) {
            // row-pattern codes: the row's whole entry list {(col - row, value)} comes from the pattern table; every
            // pattern is padded to a multiple of 8 entries with copies of its last entry ((0, 0.0) for an empty row),
            // so the gathers need neither a clamp nor a predicate.  Threads past the tile's end redo its last row.
            const DictEnt* pe[RPT];
            int len[RPT];
            double xv[RPT][JW];
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int rr = min(tid + j * THREADS, d.y - 1);
                const int2 ph = sphead[scodes[rr]];
                pe[j] = sdict + ph.x; len[j] = ph.y;
                const int row = d.x + rr;
#pragma unroll
                for (int e = 0; e < JW; ++e) xv[j][e] = ld_x<true>(x, row + pe[j][e].delta);
            }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int r = tid + j * THREADS;
                if (r < d.y) {
                    double o[Epi::NOPS > 0 ? Epi::NOPS : 1];
#pragma unroll
                    for (int q = 0; q < Epi::NOPS; ++q) o[q] = sops[q * Cfg::ROWCAP + r];
                    double sum = 0.0;                        // one accumulator, stored order
#pragma unroll
                    for (int e = 0; e < JW; ++e)
                        if (e < len[j]) sum = __dadd_rn(sum, __dmul_rn(pe[j][e].val, xv[j][e]));
                    for (int e0 = JW; e0 < len[j]; e0 += JW) {   // rows longer than JW entries
                        double xw[JW];
#pragma unroll
                        for (int e = 0; e < JW; ++e) xw[e] = ld_x<true>(x, d.x + r + pe[j][e0 + e].delta);
#pragma unroll
                        for (int e = 0; e < JW; ++e)
                            if (e0 + e < len[j]) sum = __dadd_rn(sum, __dmul_rn(pe[j][e0 + e].val, xw[e]));
                    }
                    if constexpr (NIOPS > 0) epi.store_i(d.x + r, sum, o, reinterpret_cast<const int32_t*>(sops + Epi::NOPS * Cfg::ROWCAP)[r]);
                    else epi.store(d.x + r, sum, o);
                }
            }
        } else {
        int a[RPT], b[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = tid + j * THREADS;
            const bool on = r < d.y;
            a[j] = on ? srp[r] - d.z : 0;
            b[j] = on ? srp[r + 1] - d.z : 0;
        }
        // the first JW entries of all RPT rows: RPT * JW gathers issued back to back, unconditionally (entries past a
        // row's end repeat its last entry, an empty row reads x[0]), so that none of them waits behind a predicate
        double xv[RPT][JW];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int row = d.x + tid + j * THREADS;
            const bool any = b[j] > a[j];
#pragma unroll
            for (int e = 0; e < JW; ++e) {
                const int k = any ? min(a[j] + e, b[j] - 1) : 0;
                int col;
                if constexpr (MODE == 1) col = row + sdict[scodes[k]].delta; else col = scols[k];
                xv[j][e] = ld_x<true>(x, any ? col : 0);
            }
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = tid + j * THREADS;
            if (r < d.y) {
                double o[Epi::NOPS > 0 ? Epi::NOPS : 1];
#pragma unroll
                for (int q = 0; q < Epi::NOPS; ++q) o[q] = sops[q * Cfg::ROWCAP + r];
                double sum = 0.0;                            // one accumulator, stored order
#pragma unroll
                for (int e = 0; e < JW; ++e)
                    if (a[j] + e < b[j]) sum = __dadd_rn(sum, __dmul_rn(sdict[scodes[a[j] + e]].val, xv[j][e]));
                for (int k0 = a[j] + JW; k0 < b[j];) {       // rows longer than JW entries
                    if (b[j] - k0 <= 4) { sum = coded_chunk<4, MODE>(sum, k0, b[j], d.x + r, scodes, scols, sdict, x); k0 += 4; }
                    else { sum = coded_chunk<8, MODE>(sum, k0, b[j], d.x + r, scodes, scols, sdict, x); k0 += 8; }
                }
                if constexpr (NIOPS > 0) epi.store_i(d.x + r, sum, o, reinterpret_cast<const int32_t*>(sops + Epi::NOPS * Cfg::ROWCAP)[r]);
                else epi.store(d.x + r, sum, o);
            }
        }
        }   // MODE 1 / 2
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + s);         // one arrival per consumer warp releases the stage
    }
}

// ---- row-window family: row patterns with x staged in shared memory ("stage_x") -------------------------------
// For a row-pattern-coded operator (MODE 3) in a banded numbering the distinct column offsets (col - row) over all patterns
// fall into a few WINDOWS of neighbouring offsets (3-D P1 Laplacian, N^3 nodes: {-N^2}, {-N .. +N}, {+N^2}).  A tile of
// consecutive rows [r0, r0 + nr) then needs, per window g, ONE contiguous slice x[r0 + gmin_g, r0 + nr + gmin_g + gspan_g):
// the producer lane fetches it with a 1-D bulk copy next to the code and operand slices, and the consumers read x from
// shared memory only -- no global gather, no L2 latency on the consumers' critical path.  The pattern table of this
// family holds, instead of (col - row), the SLOT of the entry's x value in the stage's x area for row 0 of the tile
// (window base + (col - row) - gmin); adding the row gives the slot.  Where offset 0 lies inside a window, the operand
// of the epilogue that aliases x (the old iterate of a Jacobi sweep) is read from that window too and not copied a
// second time.
// What the kernel is bound by is instruction issue (ncu: profiles/r2_ncu_rowwin_*.json), so the common case is stripped
// down: the most frequent ("hot") pattern travels as KERNEL PARAMETERS -- its values become constant-bank operands of the
// multiplies, its slots uniform-register offsets of the shared loads, its length a template parameter -- and a warp whose
// rows all carry it runs a straight-line body of (load, multiply, add) per entry.  Any other warp takes the table path.
// DRAM latency is covered by L2 prefetches (cp.async.bulk.prefetch.L2) PF tiles ahead instead of by more stages.
// Numerics are unchanged: one accumulator per row, stored order, separately rounded multiply and add.
constexpr int WIN_MAX = 8;               // windows per operator
constexpr int WIN_SPAN = 1536;           // neighbouring offsets are merged while a window spans at most this many columns
constexpr int WIN_GSHIFT = 20;           // pattern-table entry of this family: (window << WIN_GSHIFT) | (offset - gmin)
constexpr int WIN_HOT = 16;              // the hot pattern holds at most this many entries
struct WinPlan {
    int ng;                              // windows
    int xlen;                            // x holds xlen (even) readable entries
    int xdoubles;                        // doubles in a stage's x area = sum over windows of (ROWCAP + span, rounded up to even)
    int vslot;                           // slot of offset 0 (x[row] itself) or -1 when no window contains it
    int hot, hotlen;                     // most frequent pattern and its length
    int gmin[WIN_MAX];                   // even; slice of window g for a tile = x[row0 + gmin, row0 + nrows + gmin + gspan)
    int gspan[WIN_MAX];
    int goff[WIN_MAX];                   // first double of window g inside the x area
    int hs[WIN_HOT];                     // hot pattern: slot of every entry (filled at launch) ...
    double hv[WIN_HOT];                  // ... and its value
};

template <int THREADS, int RPT, int NOPS, int NIOPS>
struct WinCfg {
    static constexpr int ROWCAP = THREADS * RPT;
    static constexpr int HDR_BYTES = 128;
    static constexpr int CODE_BYTES = ROWCAP;
    static constexpr int OP_BYTES = ROWCAP * 8;
    static constexpr int IOP_BYTES = ROWCAP * 4;
    static constexpr int PHEAD_BYTES = 256 * 8;
    static constexpr int BAR_BYTES = 128;
    // nops: operands that are really copied (the one aliasing x is not, when a window holds offset 0)
    __host__ __device__ static constexpr int x_off(int nops) { return HDR_BYTES + CODE_BYTES + nops * OP_BYTES + NIOPS * IOP_BYTES; }
    static int stage_bytes(int nops, int xdoubles) { return (x_off(nops) + xdoubles * 8 + 127) / 128 * 128; }
    static int smem_bytes(int stages, int npent, int nops, int xdoubles) { return BAR_BYTES + PHEAD_BYTES + npent * (int)sizeof(DictEnt) + stages * stage_bytes(nops, xdoubles); }
    static_assert(ROWCAP % 32 == 0, "stage sections must stay 128-byte aligned");
};

template <class Epi, class = void> struct EpiXop { static constexpr int value = -1; };
template <class Epi> struct EpiXop<Epi, decltype((void)Epi::XOP)> { static constexpr int value = Epi::XOP; };

// consumers' wait: first probe inline, then a spin of (potentially blocking) try_waits bounded by an iteration count --
// no clock reads on the fast path (the 2 s clock bound of mbar_wait costs six instructions per probe)
__device__ __forceinline__ void mbar_wait_light(uint64_t* b, uint32_t parity)
{
    const uint32_t addr = smem_u32(b);
    uint32_t ok = 0;
    for (uint32_t it = 0; ; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) break;
        if (it > (1u << 24)) __trap();                      // a lost copy must fault, never hang the GPU
    }
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// HOTN: exact length of the hot pattern served by the straight-line body (0: table path only)
template <int THREADS, int RPT, int STAGES, int HOTN, int MINB, class Epi>
__global__ void __launch_bounds__(THREADS + 32, MINB)
k_rowwin(const unsigned char* __restrict__ rcodes, const DictEnt* __restrict__ pent, const int2* __restrict__ phead, int npent,
         const __grid_constant__ WinPlan W, const int4* __restrict__ desc, int ntiles, int tpc, int pf, const double* x, Epi epi)
{
    static_assert(Epi::CONTIG, "row-window kernel needs contiguous epilogue operands");
    static_assert(STAGES <= 8, "barrier block holds 8 stages");
    static_assert(HOTN >= 0 && HOTN <= WIN_HOT, "hot pattern length");
    constexpr int NIOPS = EpiNI<Epi>::value;
    constexpr int XOP = EpiXop<Epi>::value;
    constexpr int RB = HOTN > 8 ? 1 : 2;                     // rows whose x values are fetched back to back
    static_assert(RPT % RB == 0, "rows per thread");
    using Cfg = WinCfg<THREADS, RPT, Epi::NOPS, NIOPS>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    int2* sphead = reinterpret_cast<int2*>(smem + Cfg::BAR_BYTES);
    DictEnt* spent = reinterpret_cast<DictEnt*>(smem + Cfg::BAR_BYTES + Cfg::PHEAD_BYTES);
    unsigned char* stage0 = smem + Cfg::BAR_BYTES + Cfg::PHEAD_BYTES + npent * (int)sizeof(DictEnt);
    const bool alias = XOP >= 0 && W.vslot >= 0 && epi.operand(XOP < 0 ? 0 : XOP) == x;      // operand XOP comes out of a window
    const int nops = Epi::NOPS - (alias ? 1 : 0);
    const int x_off = Cfg::x_off(nops);
    const int stage_bytes = (x_off + W.xdoubles * 8 + 127) / 128 * 128;
    const int tid = threadIdx.x;
    __shared__ int s_gmin[WIN_MAX], s_gspan[WIN_MAX], s_goff[WIN_MAX];      // (dynamic indexing of a kernel parameter would go through local memory)
    constexpr int DCACHE = 512;
    __shared__ int2 s_desc[DCACHE];                          // {row0, nrows} of this CTA's first DCACHE tiles: the producer never waits for a descriptor
    const int first = tpc > 0 ? (int)blockIdx.x * tpc : (int)blockIdx.x;
    const int step = tpc > 0 ? 1 : (int)gridDim.x;
    const int my_tiles = tpc > 0 ? min(tpc, ntiles - first) : (ntiles - first + step - 1) / step;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int g = 0; g < WIN_MAX; ++g) { s_gmin[g] = W.gmin[g]; s_gspan[g] = W.gspan[g]; s_goff[g] = W.goff[g]; }
    }
    for (int k = tid; k < min(my_tiles, DCACHE); k += THREADS + 32) { const int4 q = __ldg(desc + first + (size_t)k * step); s_desc[k] = make_int2(q.x, q.y); }
    __syncthreads();
    for (int k = tid; k < npent; k += THREADS + 32) {        // (window, offset) -> slot in the x area
        DictEnt d = pent[k];
        d.delta = s_goff[d.delta >> WIN_GSHIFT] + (d.delta & ((1 << WIN_GSHIFT) - 1));
        spent[k] = d;
    }
    for (int k = tid; k < 256; k += THREADS + 32) sphead[k] = phead[k];
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (tid >= THREADS) {                                   // ---- producer warp (one lane works)
        if (tid == THREADS) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            const int ng = W.ng;
            // slice of window g for tile d: x[lo, hi), landing (lo - start) doubles into the window's area (start < 0 on the first tiles)
            auto slice = [&](const int4& d, int g, int& lo, int& hi, int& start) {
                start = d.x + s_gmin[g];
                const int len = (d.y + s_gspan[g] + 1) & ~1;
                lo = max(start, 0);
                hi = min(start + len, W.xlen);
            };
            auto issue_matrix = [&](int i, const int4& d) {                          // d = {row0, nrows, -, -}
                const int s = i % STAGES;
                unsigned char* st = stage0 + (size_t)s * stage_bytes;
                *reinterpret_cast<int4*>(st) = d;                                    // published by the arrive below (release)
                const uint32_t b_code = (uint32_t)((d.y + 15) & ~15), b_op = (uint32_t)((d.y + 1) & ~1) * 8u, b_iop = (uint32_t)((d.y + 3) & ~3) * 4u;
                uint32_t b_x = 0;
                for (int g = 0; g < ng; ++g) { int lo, hi, st0; slice(d, g, lo, hi, st0); if (hi > lo) b_x += (uint32_t)(hi - lo) * 8u; }
                mbar_expect_tx(full + s, b_code + (uint32_t)nops * b_op + (uint32_t)NIOPS * b_iop + b_x);
                bulk_g2s_hint(st + Cfg::HDR_BYTES, rcodes + d.x, b_code, full + s, pol);
            };
            auto issue_operands = [&](int i, const int4& d) {                        // everything the predecessor may have written
                const int s = i % STAGES;
                const uint32_t b_op = (uint32_t)((d.y + 1) & ~1) * 8u, b_iop = (uint32_t)((d.y + 3) & ~3) * 4u;
                unsigned char* st = stage0 + (size_t)s * stage_bytes;
                unsigned char* p = st + Cfg::HDR_BYTES + Cfg::CODE_BYTES;
#pragma unroll
                for (int j = 0; j < Epi::NOPS; ++j) {
                    if (alias && j == XOP) continue;
                    bulk_g2s(p, epi.operand(j) + d.x, b_op, full + s); p += Cfg::OP_BYTES;
                }
                if constexpr (NIOPS > 0) bulk_g2s(p, epi.ioperand() + d.x, b_iop, full + s);
                for (int g = 0; g < ng; ++g) {
                    int lo, hi, st0;
                    slice(d, g, lo, hi, st0);
                    if (hi > lo) bulk_g2s(st + x_off + ((size_t)s_goff[g] + (size_t)(lo - st0)) * 8, x + lo, (uint32_t)(hi - lo) * 8u, full + s);
                }
            };
            auto tile_desc = [&](int i) {
                if (i < DCACHE) { const int2 q = s_desc[i]; return make_int4(q.x, q.y, 0, 0); }
                return __ldg(desc + first + (size_t)i * step);
            };
            // L2 prefetch of what a tile streams from DRAM for the first time: codes, operand slices, the window of the largest offsets
            auto prefetch = [&](int i) {
                if (i >= my_tiles) return;
                const int4 d = tile_desc(i);
                const uint32_t b_code = (uint32_t)((d.y + 15) & ~15), b_op = (uint32_t)((d.y + 1) & ~1) * 8u, b_iop = (uint32_t)((d.y + 3) & ~3) * 4u;
                bulk_prefetch_l2(rcodes + d.x, b_code);
#pragma unroll
                for (int j = 0; j < Epi::NOPS; ++j) {
                    if (alias && j == XOP) continue;
                    bulk_prefetch_l2(epi.operand(j) + d.x, b_op);
                }
                if constexpr (NIOPS > 0) bulk_prefetch_l2(epi.ioperand() + d.x, b_iop);
                int lo, hi, st0;
                slice(d, ng - 1, lo, hi, st0);
                if (hi > lo) bulk_prefetch_l2(x + lo, (uint32_t)(hi - lo) * 8u);
            };
            const int npre = my_tiles < STAGES ? my_tiles : STAGES;
            for (int i = 0; i < npre; ++i) issue_matrix(i, tile_desc(i));
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int i = 0; i < npre; ++i) issue_operands(i, tile_desc(i));
            for (int i = npre; i < npre + pf; ++i) prefetch(i);
            for (int i = npre; i < my_tiles; ++i) {
                const int s = i % STAGES;
                const int4 d = tile_desc(i);
                if (pf > 0) prefetch(i + pf);
                mbar_wait(empty + s, (uint32_t)((i / STAGES - 1) & 1));
                issue_matrix(i, d);
                issue_operands(i, d);
            }
        }
        return;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");

    for (int i = 0; i < my_tiles; ++i) {                     // ---- consumers: thread per row, x from shared memory
        const int s = i % STAGES;
        mbar_wait_light(full + s, (uint32_t)((i / STAGES) & 1));
        const unsigned char* st = stage0 + (size_t)s * stage_bytes;
        const int nrows = reinterpret_cast<const int4*>(st)->y, row0 = reinterpret_cast<const int4*>(st)->x;
        const unsigned char* scodes = st + Cfg::HDR_BYTES;
        const double* sops = reinterpret_cast<const double*>(st + Cfg::HDR_BYTES + Cfg::CODE_BYTES);
        const double* xs = reinterpret_cast<const double*>(st + x_off);
#pragma unroll
        for (int jb = 0; jb < RPT; jb += RB) {
            int code[RB];
            bool allhot = HOTN > 0;
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int r = tid + (jb + j) * THREADS;
                code[j] = r < nrows ? (int)scodes[r] : -2;
                allhot = allhot && code[j] == W.hot;
            }
            if (HOTN > 0 && __all_sync(0xffffffffu, allhot)) {
                // ---- every row of this warp carries the hot pattern: straight-line body, constants from the parameter bank
                double xv[RB][HOTN > 0 ? HOTN : 1];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const double* xr = xs + tid + (jb + j) * THREADS;
#pragma unroll
                    for (int e = 0; e < HOTN; ++e) xv[j][e] = xr[W.hs[e]];
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int r = tid + (jb + j) * THREADS;
                    double o[Epi::NOPS > 0 ? Epi::NOPS : 1];
                    {
                        int q = 0;
#pragma unroll
                        for (int k = 0; k < Epi::NOPS; ++k) {
                            if (alias && k == XOP) o[k] = xs[W.vslot + r];
                            else { o[k] = sops[q * Cfg::ROWCAP + r]; ++q; }
                        }
                    }
                    double sum = 0.0;                        // one accumulator, stored order
#pragma unroll
                    for (int e = 0; e < HOTN; ++e) sum = __dadd_rn(sum, __dmul_rn(W.hv[e], xv[j][e]));
                    if constexpr (NIOPS > 0) epi.store_i(row0 + r, sum, o, reinterpret_cast<const int32_t*>(sops + nops * Cfg::ROWCAP)[r]);
                    else epi.store(row0 + r, sum, o);
                }
            } else {
                // ---- table path: the row's entries (value, slot) come from the pattern table in shared memory
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int r = tid + (jb + j) * THREADS;
                    if (code[j] < 0) continue;
                    double o[Epi::NOPS > 0 ? Epi::NOPS : 1];
                    {
                        int q = 0;
#pragma unroll
                        for (int k = 0; k < Epi::NOPS; ++k) {
                            if (alias && k == XOP) o[k] = xs[W.vslot + r];
                            else { o[k] = sops[q * Cfg::ROWCAP + r]; ++q; }
                        }
                    }
                    const int2 ph = sphead[code[j]];
                    const DictEnt* pe = spent + ph.x;
                    const double* xr = xs + r;
                    double sum = 0.0;                        // one accumulator, stored order
                    int e = 0;
                    for (; e + 2 <= ph.y; e += 2) {
                        const DictEnt d0 = pe[e], d1 = pe[e + 1];
                        const double x0 = xr[d0.delta], x1 = xr[d1.delta];
                        sum = __dadd_rn(sum, __dmul_rn(d0.val, x0));
                        sum = __dadd_rn(sum, __dmul_rn(d1.val, x1));
                    }
                    if (e < ph.y) { const DictEnt d0 = pe[e]; sum = __dadd_rn(sum, __dmul_rn(d0.val, xr[d0.delta])); }
                    if constexpr (NIOPS > 0) epi.store_i(row0 + r, sum, o, reinterpret_cast<const int32_t*>(sops + nops * Cfg::ROWCAP)[r]);
                    else epi.store(row0 + r, sum, o);
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + s);
    }
}

// the most frequent ("hot") row pattern of a pattern-coded operator, found from a sample of the codes at set-up
struct HotPlan {
    int hot, hotlen;                     // pattern number and length
    int hd[WIN_HOT];                     // its column offsets (col - row) ...
    double hv[WIN_HOT];                  // ... and values
};

// ---- halo exchange fused into the kernels that write / read an iterate (row-sharded hierarchies) ---------------
// Every kernel that writes an iterate (Jacobi sweep, prolongation + correction, zero-guess sweep) also stores the rows its
// neighbours hold as ghosts straight into the neighbours' copy of that vector (peer memory over NVLink), and the last CTA to
// finish such rows publishes a new epoch in the neighbour's arrival flag.  Every kernel that reads ghost entries makes the CTAs
// whose rows reference them wait until the neighbours' flags have reached this rank's own epoch count (all ranks run the same
// kernel sequence, so the counts agree).  No pack / push / pull launch remains.  Boundary tiles run FIRST in every kernel, so a
// flag is published a whole sweep before the neighbour's next kernel asks for it: the wait is over before it starts.
// Safety of writing into the neighbour's vector directly: iterates ping-pong between two buffers; this rank writes buffer Y of
// the neighbour during sweep s, the neighbour last read Y's ghosts in sweep s-1, and this rank can only be in sweep s after it
// has seen the neighbour's flag for the END of the neighbour's boundary rows of sweep s-1.
struct HaloFuse {
    int wait_n;                               // consumer: neighbours whose flag must have reached wepoch[p] (0: no ghost is read)
    const unsigned long long* wflag[2];       //   this rank's arrival flags, written by the neighbours
    const unsigned long long* wepoch;         //   device: [p] = pushes exchanged with neighbour p so far (level of x)
    int int_b0, int_b1;                       //   rows [int_b0, int_b1) reference no ghost entry
    int send_n;                               // producer: neighbours that hold some of this rank's rows as ghosts
    int send_a[2], send_cnt[2];               //   owned rows [a, a + cnt) are neighbour p's ghosts ...
    double* dst[2];                           //   ... at dst[p][row - a] in the neighbour's copy of the vector this kernel writes
    unsigned long long* rflag[2];             //   the neighbour's arrival flag for this rank
    unsigned long long* sepoch;               //   device: [p] pushes so far (level of the output), [8 + p] CTA arrival counters
    int send_ctas[2];                         //   CTAs whose tile intersects the range: the last one to arrive publishes
    int nlo, nhi;                             // tiles at the low / high end that wait or send: they run first
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// tile this CTA works on: boundary tiles (nlo at the low end, nhi at the high end) first, then the interior in order
__device__ __forceinline__ int halo_tile(const HaloFuse& hf, int b, int ntiles)
{
    if (hf.nlo + hf.nhi == 0 || b < hf.nlo) return b;
    if (b < hf.nlo + hf.nhi) return ntiles - hf.nhi + (b - hf.nlo);
    return b - hf.nhi;
}
// CTA-uniform: rows [row0, rend) read a ghost entry of x -> wait for the neighbours (one polling thread per neighbour; its
// acquire also drops stale L1 lines of this SM), then release the CTA
__device__ __forceinline__ void halo_wait(const HaloFuse& hf, int row0, int rend)
{
    if (hf.wait_n == 0 || (row0 >= hf.int_b0 && rend <= hf.int_b1)) return;
    if ((int)threadIdx.x < hf.wait_n) {
        const int p = threadIdx.x;
        const unsigned long long want = ld_relaxed_gpu(hf.wepoch + p);
        const long long t0 = clock64();
        unsigned long long got;
        do {                                                 // relaxed polls, one acquire fence at the end (an acquire drops the SM's L1 lines)
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(hf.wflag[p]) : "memory");
            if (got < want && clock64() - t0 > 60000000000LL) __trap();   // ~30 s: a lost neighbour must fault, never hang the GPU
        } while (got < want);
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
    __syncthreads();
}
// CTA-uniform: bit p set when rows [row0, rend) intersect the rows neighbour p holds as ghosts (0 for almost every CTA, which
// then skips all per-row work of the exchange)
__device__ __forceinline__ int halo_sends(const HaloFuse& hf, int row0, int rend)
{
    int m = 0;
#pragma unroll
    for (int p = 0; p < 2; ++p)
        if (p < hf.send_n && row0 < hf.send_a[p] + hf.send_cnt[p] && rend > hf.send_a[p]) m |= 1 << p;
    return m;
}
// the value just written to row r of the iterate also goes to the neighbours that hold row r as a ghost
__device__ __forceinline__ void halo_send(const HaloFuse& hf, int r, double val)
{
#pragma unroll
    for (int p = 0; p < 2; ++p)
        if (p < hf.send_n && r >= hf.send_a[p] && r < hf.send_a[p] + hf.send_cnt[p]) hf.dst[p][r - hf.send_a[p]] = val;
}
// CTA-uniform, after the CTA's rows [row0, rend) have been stored: if they intersect a send range, fence, count this CTA in,
// and -- last CTA of the range -- publish the next epoch in the neighbour's flag
__device__ __forceinline__ void halo_publish(const HaloFuse& hf, int row0, int rend)
{
    if (!halo_sends(hf, row0, rend)) return;
    __syncthreads();                                         // every thread's peer stores are ordered before thread 0's fence (cumulativity)
    if (threadIdx.x == 0) {
        __threadfence_system();
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            if (!(p < hf.send_n && row0 < hf.send_a[p] + hf.send_cnt[p] && rend > hf.send_a[p])) continue;
            const unsigned long long prev = atomicAdd(hf.sepoch + 8 + p, 1ULL);
            if (prev == (unsigned long long)(hf.send_ctas[p] - 1)) {
                hf.sepoch[8 + p] = 0;
                const unsigned long long e = ld_relaxed_gpu(hf.sepoch + p) + 1;
                hf.sepoch[p] = e;
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(hf.rflag[p]), "l"(e) : "memory");
            }
        }
    }
}

// ---- hot-row family: row patterns, speculative loads at the hot pattern's offsets ------------------------------
// ncu on k_rowstream / k_rowwin (profiles/r2_ncu_*): half of the issued instructions were consumers spinning on a
// stage that had not landed, and a thread-per-row kernel that first loads its code and only then its x values pays two
// DRAM latencies per row.  This family removes the dependency instead of hiding it:
//   * every pattern that is a SUBSEQUENCE of the most frequent ("hot") pattern -- same (col - row, value) pairs, same
//     order, some missing: the rows next to a Dirichlet boundary -- is described by a bit mask over the hot entries
//     (pmask[code]); any other pattern carries HOT_SLOW and takes the table walk;
//   * a thread issues, back to back and before anything has arrived, its code byte, ALL x values at the hot offsets
//     (speculatively: a masked-off value is loaded and ignored; in the first / last tiles the index is clamped into x)
//     and the epilogue operands.  One memory latency per row, no shared memory, no barrier, no descriptor;
//   * a CTA prefetches into L2 (cp.async.bulk.prefetch.L2) what the tile `pf` tiles ahead will read from DRAM for the
//     first time (codes, operand slices, the leading x slice), so the demand loads of that tile find L2, not DRAM.
// The row sum is formed as everywhere else: one accumulator, stored order, separately rounded multiply and add; a masked
// entry is skipped by predication (never multiplied by zero), so results are bit-identical to the CSR kernels.
constexpr uint32_t HOT_SLOW = 0x80000000u;
struct HotArgs {
    int dmin, dmax;                      // smallest / largest hot offset (tiles that may leave [0, xlen) take the clamped body)
    int hd[WIN_HOT];                     // hot pattern: col - row of every entry ...
    double hv[WIN_HOT];                  // ... and its value
};

__device__ __forceinline__ int ld_stream_u8(const unsigned char* p)
{
    unsigned short v;
    asm("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=h"(v) : "l"(p));
    return (int)v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p)
{
    double v;      // volatile: stays behind griddepcontrol.wait (the predecessor kernel may have written the operand)
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));      // (not .nc: an epilogue may update the operand in place)
    return v;
}

// Tiles: with desc == nullptr the rows [rb, re) are cut into tiles of T = THREADS * RPT consecutive rows (no descriptor is
// read: the tile's rows follow from blockIdx); with a descriptor list, tile t holds the rows [desc[t].x, desc[t].x + desc[t].y),
// desc[t].y <= T (tile subsets: the interior / boundary split of a sharded operator, the tiles holding injected rows).
// Thread t owns rows t, t + THREADS, ... of its tile.
// HALO: compiled with the fused halo exchange (HaloFuse); the single-GPU instantiation carries none of its tests -- the kernel
// issues at ~65 % of its slots (profiles/r2_ncu_full_k_hotrow_cfg5.json), so per-CTA bookkeeping is not free.
template <int HOTN, int THREADS, int RPT, int MINB, bool HALO, class Epi>
__global__ void __launch_bounds__(THREADS, MINB)
k_hotrow(const unsigned char* __restrict__ rcodes, const uint32_t* __restrict__ pmask, const int2* __restrict__ phead,
         const DictEnt* __restrict__ pent, const __grid_constant__ HotArgs H, const __grid_constant__ HaloFuse hf,
         const int4* __restrict__ desc, int ntiles, int rb, int re, int xlen, int pf, int pf_last, const double* x, Epi epi)
{
    static_assert(Epi::CONTIG, "hot-row kernel needs contiguous epilogue operands");
    static_assert(HOTN >= 1 && HOTN <= WIN_HOT, "hot pattern length");
    constexpr int T = THREADS * RPT;
    constexpr int NOPS = Epi::NOPS, NIOPS = EpiNI<Epi>::value, XOP = EpiXop<Epi>::value;
    const int tid = threadIdx.x;
    bool alias = false;                                      // the operand that is x itself (old iterate of a Jacobi sweep)
    if constexpr (XOP >= 0) alias = epi.operand(XOP) == x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int tile = (int)blockIdx.x;
    if constexpr (HALO) { if (!desc) tile = halo_tile(hf, (int)blockIdx.x, ntiles); }       // (sharded: boundary tiles first)
    int row0 = rb + tile * T, rend = re;                     // this tile: rows [row0, min(row0 + T, rend))
    if (desc) { const int4 d = __ldg(desc + blockIdx.x); row0 = d.x; rend = d.x + d.y; }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int sends = 0;
    if constexpr (HALO) {
        halo_wait(hf, row0, min(row0 + T, rend));            // rows that read ghost entries: the neighbours' rows have arrived
        sends = halo_sends(hf, row0, min(row0 + T, rend));
    }
    int code[RPT];
    double xv[RPT][HOTN];
    double o[RPT][NOPS > 0 ? NOPS : 1];
    int io[RPT];
    const bool fast = row0 + H.dmin >= 0 && (long long)row0 + T + H.dmax <= (long long)xlen && row0 + T <= rend;   // CTA-uniform
    if (fast) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = row0 + tid + j * THREADS;
            code[j] = ld_stream_u8(rcodes + r);
            const double* xr = x + r;
#pragma unroll
            for (int e = 0; e < HOTN; ++e) xv[j][e] = xr[H.hd[e]];             // (coherent loads: see ld_x)
#pragma unroll
            for (int k = 0; k < NOPS; ++k) o[j][k] = (alias && k == XOP) ? xr[0] : ld_stream_f64(epi.operand(k) + r);
            if constexpr (NIOPS > 0) io[j] = __ldg(epi.ioperand() + r);
        }
    } else {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = min(row0 + tid + j * THREADS, rend - 1);        // threads past the end redo the last row (and do not store)
            code[j] = ld_stream_u8(rcodes + r);
#pragma unroll
            for (int e = 0; e < HOTN; ++e) xv[j][e] = x[min(max(r + H.hd[e], 0), xlen - 1)];
#pragma unroll
            for (int k = 0; k < NOPS; ++k) o[j][k] = epi.operand(k)[r];
            if constexpr (NIOPS > 0) io[j] = epi.ioperand()[r];
        }
    }
    if (pf > 0 && tid < 3 + NOPS && tile + pf <= pf_last) {  // L2 prefetch for the tile pf tiles ahead (linear tiling only; the host
        const long long p0 = (long long)row0 + (long long)pf * T;         // bounds pf_last so that every slice stays inside its array)
        const void* ptr = rcodes + p0;
        uint32_t bytes = T;
        if (tid == 1) { ptr = x + p0 + (H.dmax & ~1); bytes = T * 8; }
#pragma unroll
        for (int k = 0; k < NOPS; ++k)
            if (tid == 2 + k) { ptr = (alias && k == XOP) ? nullptr : (const void*)(epi.operand(k) + p0); bytes = T * 8; }
        if constexpr (NIOPS > 0) { if (tid == 2 + NOPS) { ptr = epi.ioperand() + p0; bytes = T * 4; } }
        else { if (tid == 2 + NOPS) ptr = nullptr; }
        if (ptr) bulk_prefetch_l2(ptr, bytes);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int r = row0 + tid + j * THREADS;
        const uint32_t m = __ldg(pmask + code[j]);
        double sum = 0.0;                                    // one accumulator, stored order
        if (__all_sync(0xffffffffu, m == (1u << HOTN) - 1u)) {
            // every row of the warp carries the complete hot pattern (the common case away from the boundary): straight-line body,
            // no predicate, no select
#pragma unroll
            for (int e = 0; e < HOTN; ++e) sum = __dadd_rn(sum, __dmul_rn(H.hv[e], xv[j][e]));
        } else if (!(m & HOT_SLOW)) {
#pragma unroll
            for (int e = 0; e < HOTN; ++e)
                if ((m >> e) & 1u) sum = __dadd_rn(sum, __dmul_rn(H.hv[e], xv[j][e]));
        } else if (r < rend) {                               // any other pattern: walk its entry list
            const int2 ph = __ldg(phead + code[j]);
            for (int e = 0; e < ph.y; ++e) {
                const double val = __ldg(&pent[ph.x + e].val);
                const int dl = __ldg(&pent[ph.x + e].delta);
                sum = __dadd_rn(sum, __dmul_rn(val, x[r + dl]));
            }
        }
        if (fast || r < rend) {
            double t;
            if constexpr (NIOPS > 0) t = epi.store_i(r, sum, o[j], io[j]);
            else t = epi.store(r, sum, o[j]);
            if constexpr (HALO) { if (sends) halo_send(hf, r, t); }
        }
    }
    if constexpr (HALO) { if (sends) halo_publish(hf, row0, min(row0 + T, rend)); }
}

// ---- two Jacobi sweeps in one launch (k_hotrow2) -----------------------------------------------------------------------------
// A sweep of the 513^3 level streams 25 B per row and every byte comes from / goes to DRAM: the iterate (1.08 GB) is far larger than
// L2, so sweep s + 1 finds nothing of what sweep s wrote.  Rows are swept in storage order, however, and row i of sweep s + 1 only
// needs the rows i + dmin .. i + dmax of sweep s: the second sweep can follow the first at a distance of `lag` tiles INSIDE ONE
// KERNEL, while the intermediate iterate y (and g) of those tiles are still in L2.  CTA 2i sums tile i of the first sweep (x -> y),
// CTA 2i + 1 tile i - lag of the second (y -> out, which may be x itself: every reader of that part of x is among the tiles waited
// for).
// Dependencies: a CTA that has stored a first-sweep tile fences and counts itself into the tile's GROUP counter (64 tiles per
// group); before its second-sweep tile a CTA waits until the groups covering the tiles [t - reach, t + reach] are complete (one
// polling thread per group, relaxed loads, __syncthreads).  The counters are zeroed by a one-block kernel after every launch
// (k_s2_reset: a counter of CTAs that finished, to let the last one do it, would cost every CTA an atomic round trip), so the
// pair replays inside a CUDA graph.  No deadlock: the second sweep trails by at least reach + 64 tiles, so a CTA only ever waits
// for first-sweep tiles of CTAs with a SMALLER block index -- already dispatched, and first-sweep CTAs wait for nothing; a wait
// that still does not end traps.
// What it saves: y and g of the second sweep are L2 hits, i.e. 33 instead of 50 B per row from DRAM for the pair.  Arithmetic per
// row is that of k_hotrow: results are bit-identical to two separate sweeps.
constexpr int S2_GROUP = 64;

// one tile of the hot-row body without the fused exchange: rows [tile * T, min(tile * T + T, rend))  (see k_hotrow)
__device__ __forceinline__ double ld_l2_f64(const double* p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// XL2: x is read past L1 (at L2, the point of coherence).  POL: L2 eviction priorities -- 0 none; 1 first-sweep role (the result
// and g are needed again `lag` tiles later: evict_last); 2 second-sweep role (the result and g are not needed again: evict_first)
template <int HOTN, int THREADS, int RPT, bool XL2, int POL, class Epi>
__device__ __forceinline__ void hot_tile(const unsigned char* __restrict__ rcodes, const uint32_t* __restrict__ pmask,
                                         const int2* __restrict__ phead, const DictEnt* __restrict__ pent, const HotArgs& H, int tile,
                                         int rend, int xlen, int pf, int pf_last, const double* x, const Epi& epi)
{
    constexpr int T = THREADS * RPT;
    constexpr int NOPS = Epi::NOPS, XOP = EpiXop<Epi>::value;
    const int tid = threadIdx.x;
    bool alias = false;
    if constexpr (XOP >= 0) alias = epi.operand(XOP) == x;
    const int row0 = tile * T;
    const unsigned long long pkeep = l2_policy(POL ? 1 : 0), pout = l2_policy(POL);   // (operands other than the old iterate: g / dinv / f)
    int code[RPT];
    double xv[RPT][HOTN];
    double o[RPT][NOPS > 0 ? NOPS : 1];
    const bool fast = row0 + H.dmin >= 0 && (long long)row0 + T + H.dmax <= (long long)xlen && row0 + T <= rend;   // CTA-uniform
    if (fast) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = row0 + tid + j * THREADS;
            code[j] = ld_stream_u8(rcodes + r);
            const double* xr = x + r;
#pragma unroll
            for (int e = 0; e < HOTN; ++e) xv[j][e] = XL2 ? ld_l2_f64(xr + H.hd[e]) : xr[H.hd[e]];
#pragma unroll
            for (int k = 0; k < NOPS; ++k) o[j][k] = (alias && k == XOP) ? (XL2 ? ld_l2_f64(xr) : xr[0]) : ld_stream_l2<POL>(epi.operand(k) + r, pout);
        }
    } else {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = min(row0 + tid + j * THREADS, rend - 1);        // threads past the end redo the last row (and do not store)
            code[j] = ld_stream_u8(rcodes + r);
#pragma unroll
            for (int e = 0; e < HOTN; ++e) { const double* q = x + min(max(r + H.hd[e], 0), xlen - 1); xv[j][e] = XL2 ? ld_l2_f64(q) : *q; }
#pragma unroll
            for (int k = 0; k < NOPS; ++k) o[j][k] = XL2 ? ld_l2_f64(epi.operand(k) + r) : epi.operand(k)[r];
        }
    }
    if (pf > 0 && tid < 2 + NOPS && tile + pf <= pf_last) {  // L2 prefetch for the tile pf tiles ahead
        const long long p0 = (long long)row0 + (long long)pf * T;
        const void* ptr = rcodes + p0;
        uint32_t bytes = T;
        if (tid == 1) { ptr = x + p0 + (H.dmax & ~1); bytes = T * 8; }
#pragma unroll
        for (int k = 0; k < NOPS; ++k)
            if (tid == 2 + k) { ptr = (alias && k == XOP) ? nullptr : (const void*)(epi.operand(k) + p0); bytes = T * 8; }
        if (ptr) bulk_prefetch_l2(ptr, bytes);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int r = row0 + tid + j * THREADS;
        const uint32_t m = __ldg(pmask + code[j]);
        double sum = 0.0;                                    // one accumulator, stored order
        if (__all_sync(0xffffffffu, m == (1u << HOTN) - 1u)) {
#pragma unroll
            for (int e = 0; e < HOTN; ++e) sum = __dadd_rn(sum, __dmul_rn(H.hv[e], xv[j][e]));
        } else if (!(m & HOT_SLOW)) {
#pragma unroll
            for (int e = 0; e < HOTN; ++e)
                if ((m >> e) & 1u) sum = __dadd_rn(sum, __dmul_rn(H.hv[e], xv[j][e]));
        } else if (r < rend) {                               // any other pattern: walk its entry list
            const int2 ph = __ldg(phead + code[j]);
            for (int e = 0; e < ph.y; ++e) {
                const double val = __ldg(&pent[ph.x + e].val);
                const int dl = __ldg(&pent[ph.x + e].delta);
                sum = __dadd_rn(sum, __dmul_rn(val, XL2 ? ld_l2_f64(x + r + dl) : x[r + dl]));
            }
        }
        if (fast || r < rend) {
            if constexpr (POL == 0) epi.store(r, sum, o[j]);
            else epi.template store_p<POL>(r, sum, o[j], pout, pkeep);
        }
    }
}

// cnt[g]: first-sweep tiles of group g finished in this launch (zeroed again by k_s2_reset after every launch).
// kt: tiles per CTA, a power of two <= S2_GROUP -- a CTA lives ~1.4 us per tile, a release or a poll costs about as much, so both
// are paid once per kt tiles (tiles of one CTA are consecutive, the other resident CTAs hide the latency as separate CTAs would).
template <int HOTN, int THREADS, int RPT, int MINB, bool XL2, bool POL, class Epi1, class Epi2>
__global__ void __launch_bounds__(THREADS, MINB)
k_hotrow2(const unsigned char* __restrict__ rcodes, const uint32_t* __restrict__ pmask, const int2* __restrict__ phead,
          const DictEnt* __restrict__ pent, const __grid_constant__ HotArgs H, int ntiles, int rend, int xlen, int pf, int pf_last,
          const double* x, Epi1 epi1, const double* y, Epi2 epi2, unsigned long long* cnt, int kt, int lagc, int reach)
{
    static_assert(Epi1::CONTIG && Epi2::CONTIG, "hot-row kernel needs contiguous epilogue operands");
    const int tid = threadIdx.x, b = (int)blockIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // roles alternate with the block index: even blocks sum first-sweep tiles, odd blocks second-sweep tiles lagc chunks behind --
    // a CTA that did both in turn would sit through two memory latencies, a fence and a poll with nothing in flight
    if (!(b & 1)) {
        const int t0 = (b >> 1) * kt, t1 = min(t0 + kt, ntiles);
        if (t0 < ntiles) {
#pragma unroll 1
            for (int t = t0; t < t1; ++t)
                hot_tile<HOTN, THREADS, RPT, false, POL ? 1 : 0, Epi1>(rcodes, pmask, phead, pent, H, t, rend, xlen, pf, pf_last, x, epi1);
            __syncthreads();                                 // every thread's stores are ordered before thread 0's release (cumulativity)
            // release without an acquire: MEMBAR + REDG (a fence would also invalidate this SM's L1 under every CTA resident on it)
            if (tid == 0)
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(cnt + t0 / S2_GROUP), "l"((unsigned long long)(t1 - t0)) : "memory");
        }
    } else {
        const int t0 = ((b >> 1) - lagc) * kt, t1 = min(t0 + kt, ntiles);
        if (t0 >= 0 && t0 < ntiles) {
            const int g0 = max(t0 - reach, 0) / S2_GROUP, g1 = min(t1 - 1 + reach, ntiles - 1) / S2_GROUP;
            if (tid <= g1 - g0) {
                const int g = g0 + tid;
                const unsigned long long want = (unsigned long long)min(S2_GROUP, ntiles - g * S2_GROUP);
                const long long c0 = clock64();
                while (ld_relaxed_gpu(cnt + g) < want)
                    if (clock64() - c0 > 8000000000LL) __trap();     // ~4 s: a lost tile must fault, never hang the GPU
            }
            // No acquire fence: it would invalidate the SM's whole L1 (CCTL.IVALL) once per CTA, and no line of y can be stale in
            // it -- y is only ever loaded by second-sweep CTAs, each after ITS wait, and every line a tile touches lies inside the
            // tiles waited for, i.e. was complete at L2 (the writers' release) before any SM could have fetched it.  Ordering: the
            // loads below are issued after the barrier, the barrier after the polls returned.  XL2 reads y past L1 altogether.
            __syncthreads();
#pragma unroll 1
            for (int t = t0; t < t1; ++t)
                hot_tile<HOTN, THREADS, RPT, XL2, POL ? 2 : 0, Epi2>(rcodes, pmask, phead, pent, H, t, rend, xlen, 0, 0, y, epi2);
        }
    }
}

__global__ void k_s2_reset(int n, unsigned long long* cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = 0ULL;
}

// Fused residual + injection on a row-pattern-coded operator, one thread per COARSE row i: out[i] = f[g_i] - (A v)[g_i],
// g_i = inj[i] the fine row with the same coordinate (multigrid.py:244 restricted to the rows Restriction2D_direct keeps,
// multigrid.py:128-131).  Only the injected rows are summed -- 1/4 (2-D) or 1/8 (3-D) of the fine rows -- with the same
// speculative loads as k_hotrow (every index clamped into x: the rows are scattered, there is no tile-wide range test).
// The inj slice of the tile pf tiles ahead is prefetched into L2.
// x, f and out are deliberately NOT __restrict__: a const __restrict__ pointer turns its loads into invariant loads, which the
// compiler may hoist above griddepcontrol.wait -- i.e. read the iterate before the predecessor kernel has written it (found
// the hard way; tools/check_pdl_order.py lists the loads scheduled before the wait).
template <int HOTN, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_hotinj(const unsigned char* __restrict__ rcodes, const uint32_t* __restrict__ pmask, const int2* __restrict__ phead,
         const DictEnt* __restrict__ pent, const __grid_constant__ HotArgs H, const __grid_constant__ HaloFuse hf,
         const int32_t* __restrict__ inj, int mono, int nc, int xlen, int pf, const double* x, const double* f, double* out)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int i = (int)blockIdx.x * THREADS + (int)threadIdx.x;
    const int r = __ldg(inj + min(i, nc - 1));               // (static data: may be read before the predecessor has finished)
    if (pf > 0 && threadIdx.x == 0) {
        const long long p0 = ((long long)blockIdx.x + pf) * THREADS;
        if (p0 + THREADS <= (long long)nc) bulk_prefetch_l2(inj + p0, THREADS * 4);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    {   // rows this CTA sums: [first, last] of its slice of the injection list when that list is ascending (mono), else unknown
        int ra = 0, rb = 0x7fffffff;
        if (mono) { ra = __ldg(inj + min((int)blockIdx.x * THREADS, nc - 1)); rb = __ldg(inj + min((int)blockIdx.x * THREADS + THREADS - 1, nc - 1)) + 1; }
        halo_wait(hf, ra, rb);
    }
    const int code = ld_stream_u8(rcodes + r);
    double xv[HOTN];
#pragma unroll
    for (int e = 0; e < HOTN; ++e) xv[e] = ld_x<true>(x, min(max(r + H.hd[e], 0), xlen - 1));
    const double fr = ld_x<true>(f, r);
    const uint32_t m = __ldg(pmask + code);
    double sum = 0.0;                                        // one accumulator, stored order
    if (!(m & HOT_SLOW)) {
#pragma unroll
        for (int e = 0; e < HOTN; ++e)
            if ((m >> e) & 1u) sum = __dadd_rn(sum, __dmul_rn(H.hv[e], xv[e]));
    } else {
        const int2 ph = __ldg(phead + code);
        for (int e = 0; e < ph.y; ++e) {
            const double val = __ldg(&pent[ph.x + e].val);
            const int dl = __ldg(&pent[ph.x + e].delta);
            sum = __dadd_rn(sum, __dmul_rn(val, ld_x<true>(x, r + dl)));
        }
    }
    if (i < nc) out[i] = __dsub_rn(fr, sum);
}

// ---- anchored row patterns (mode 4): rectangular operators ------------------------------------------------------
// The prolongation of a uniform mesh repeats too, but its columns do not follow the row index: row i of P touches coarse
// nodes around (roughly) i / 2^d.  Measured from the row's FIRST stored column, however, the list of (col - first, value) is one
// of a handful (8 parity classes for the trilinear P).  Such an operator is kept as 4 + 1 bytes per ROW -- the anchor column
// and the pattern number -- instead of 12 bytes per entry plus a row pointer; the coding is found and verified on the device
// (try_patterns, anchored).  Thread per row, RPT rows per thread: (anchor, code, operands) are loaded for all rows first, then
// the pattern entries (a table of <= 2048 entries, L1-resident) and the x values, then the sums -- one accumulator, stored
// order, separately rounded multiply and add, as everywhere.  Patterns are padded to a multiple of 8 entries with copies of
// their last entry, so the JW loads of a chunk are issued without a bounds test.  Tiles as in k_hotrow.
__device__ __forceinline__ int ld_stream_i32(const int32_t* p)
{
    int v;
    asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// What bounds this kernel is the dependent chain (anchor, code) -> pattern entry -> x value: three memory latencies per row.
// So (1) anchors and codes are static data and are requested BEFORE griddepcontrol.wait, overlapping the predecessor's tail;
// (2) the pattern table sits in shared memory (copied by the CTA while those requests are in flight), which turns two of the
// three latencies into shared-memory look-ups; (3) a thread keeps RPT rows in flight, holding only their x values in registers
// (the entry values are looked up again when the row is summed).
// HALO: compiled with the fused halo exchange; the single-GPU instantiation carries none of its tests (as k_hotrow).
template <int THREADS, int RPT, int JW, int MINB, bool HALO, class Epi>
__global__ void __launch_bounds__(THREADS, MINB)
k_anchrow(const unsigned char* __restrict__ rcodes, const int32_t* __restrict__ anchor, const int2* __restrict__ phead,
          const DictEnt* __restrict__ pent, int ndict, int npent, const __grid_constant__ HaloFuse hf, const int4* __restrict__ desc,
          int ntiles, int rb, int re, int pf, const double* x, Epi epi)
{
    static_assert(Epi::CONTIG, "anchored-pattern kernel needs contiguous epilogue operands");
    static_assert(JW == 4 || JW == 8, "chunk width");
    constexpr int T = THREADS * RPT;
    constexpr int NOPS = Epi::NOPS, NIOPS = EpiNI<Epi>::value;
    extern __shared__ __align__(16) unsigned char smem_anch[];
    double* sval = reinterpret_cast<double*>(smem_anch);                       // [npent]
    int2* sphead = reinterpret_cast<int2*>(smem_anch + (size_t)npent * 8);      // [256]
    int* sdelta = reinterpret_cast<int*>(smem_anch + (size_t)npent * 8 + 256 * 8);   // [npent]
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int tile = (int)blockIdx.x;
    if constexpr (HALO) { if (!desc) tile = halo_tile(hf, (int)blockIdx.x, ntiles); }
    int row0 = rb + tile * T, rend = re;
    if (desc) { const int4 d = __ldg(desc + blockIdx.x); row0 = d.x; rend = d.x + d.y; }
    int code[RPT], anc[RPT];
    const bool full = row0 + T <= rend;                      // CTA-uniform
#pragma unroll
    for (int j = 0; j < RPT; ++j) {                          // static data: requested before the predecessor has finished
        const int r = full ? row0 + tid + j * THREADS : min(row0 + tid + j * THREADS, rend - 1);
        code[j] = ld_stream_u8(rcodes + r);
        anc[j] = ld_stream_i32(anchor + r);
    }
    for (int k = tid; k < npent; k += THREADS) { const DictEnt d = pent[k]; sval[k] = d.val; sdelta[k] = d.delta; }
    for (int k = tid; k < ndict; k += THREADS) sphead[k] = phead[k];
    asm volatile("griddepcontrol.wait;" ::: "memory");
    double o[RPT][NOPS > 0 ? NOPS : 1];
    int io[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int r = full ? row0 + tid + j * THREADS : min(row0 + tid + j * THREADS, rend - 1);
#pragma unroll
        for (int k = 0; k < NOPS; ++k) o[j][k] = ld_stream_f64(epi.operand(k) + r);
        if constexpr (NIOPS > 0) io[j] = epi.ioperand()[r];
    }
    if (pf > 0 && tid < 32 && tile + pf < ntiles) {          // L2 prefetch for the tile pf tiles ahead
        long long p0 = (long long)row0 + (long long)pf * T;
        int pn = T;
        if (desc) { const int4 d = __ldg(desc + blockIdx.x + pf); p0 = d.x; pn = (d.y + 15) & ~15; }
        if (p0 + pn <= (long long)re) {
            if (tid == 0) bulk_prefetch_l2(rcodes + (p0 & ~15LL), pn);
            if (tid == 1) bulk_prefetch_l2(anchor + (p0 & ~3LL), pn * 4);
            if (tid >= 2 && tid < 2 + NOPS) bulk_prefetch_l2(epi.operand(tid - 2) + (p0 & ~1LL), pn * 8);
            if constexpr (NIOPS > 0) if (tid == 2 + NOPS) bulk_prefetch_l2(epi.ioperand() + (p0 & ~3LL), pn * 4);
        }
    }
    __syncthreads();                                         // the table is in shared memory
    int sends = 0;
    if constexpr (HALO) {
        halo_wait(hf, row0, min(row0 + T, rend));            // rows whose columns reach into the ghost section of x
        sends = halo_sends(hf, row0, min(row0 + T, rend));
    }
    // Entries are fetched in two halves of JW / 2: the second half only by warps in which some row is longer than the first
    // (warp-uniform test).  Rows of a prolongation alternate between short and long patterns along a mesh line, and whole lines
    // are short: on the trilinear P three warps in four never need entries 4..7.
    // (Measured and dropped: the pattern table as a kernel parameter instead of shared memory -- divergent constant-bank loads
    // made the 513^3 prolongation 1.05 -> 1.31 ms, profiles/r2_variants_hot_513i.jsonl.)
    constexpr int HW = JW / 2;
    int2 ph[RPT];
    double xv[RPT][JW];
    bool more[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        ph[j] = sphead[code[j]];
        const int* sd = sdelta + ph[j].x;
        const double* xa = x + anc[j];                       // (one wide multiply-add per gather instead of an add and one)
#pragma unroll
        for (int e = 0; e < HW; ++e) xv[j][e] = xa[sd[e]];                    // (coherent load: see ld_x; padded entries repeat the last one)
        more[j] = __any_sync(0xffffffffu, ph[j].y > HW);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        if (more[j]) {
            const int* sd = sdelta + ph[j].x;
            const double* xa = x + anc[j];
#pragma unroll
            for (int e = HW; e < JW; ++e) xv[j][e] = xa[sd[e]];
        }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int r = row0 + tid + j * THREADS;
        double sum = 0.0;                                    // one accumulator, stored order
        const double* sv = sval + ph[j].x;
#pragma unroll
        for (int e = 0; e < HW; ++e)
            if (e < ph[j].y) sum = __dadd_rn(sum, __dmul_rn(sv[e], xv[j][e]));
        if (more[j]) {
#pragma unroll
            for (int e = HW; e < JW; ++e)
                if (e < ph[j].y) sum = __dadd_rn(sum, __dmul_rn(sv[e], xv[j][e]));
        }
        for (int e0 = JW; e0 < ph[j].y; e0 += JW) {          // rows longer than JW entries
            double xw[JW];
#pragma unroll
            for (int e = 0; e < JW; ++e) xw[e] = x[anc[j] + sdelta[ph[j].x + e0 + e]];
#pragma unroll
            for (int e = 0; e < JW; ++e)
                if (e0 + e < ph[j].y) sum = __dadd_rn(sum, __dmul_rn(sval[ph[j].x + e0 + e], xw[e]));
        }
        if (full || r < rend) {
            double t;
            if constexpr (NIOPS > 0) t = epi.store_i(r, sum, o[j], io[j]);
            else t = epi.store(r, sum, o[j]);
            if constexpr (HALO) { if (sends) halo_send(hf, r, t); }
        }
    }
    if constexpr (HALO) { if (sends) halo_publish(hf, row0, min(row0 + T, rend)); }
}

// Several consecutive tiles per CTA, the next tile's anchors and codes requested while the current tile is being summed.  The ncu
// source view of k_anchrow (profiles/r2_ncu_k_anchrow_plain_cfg5_*.json) shows three exposed waits per tile -- (anchor, code) ->
// shared-memory table -> gathered x values -- that account for 37 % of the stall samples at 20 resident warps per SM; here the
// first of them is paid once per CTA instead of once per tile, and so are the table copy and its barrier.  Linear tiling of all
// rows, no fused exchange (single GPU / unsharded levels); arithmetic per row as in k_anchrow, hence the same bits.
template <int THREADS, int RPT, int JW, int MINB, class Epi>
__global__ void __launch_bounds__(THREADS, MINB)
k_anchloop(const unsigned char* __restrict__ rcodes, const int32_t* __restrict__ anchor, const int2* __restrict__ phead,
           const DictEnt* __restrict__ pent, int ndict, int npent, int ntiles, int re, int kt, int pf, const double* x, Epi epi)
{
    static_assert(Epi::CONTIG, "anchored-pattern kernel needs contiguous epilogue operands");
    static_assert(JW == 4 || JW == 8, "chunk width");
    constexpr int T = THREADS * RPT;
    constexpr int NOPS = Epi::NOPS, NIOPS = EpiNI<Epi>::value;
    constexpr int HW = JW / 2;
    extern __shared__ __align__(16) unsigned char smem_anch[];
    double* sval = reinterpret_cast<double*>(smem_anch);                       // [npent]
    int2* sphead = reinterpret_cast<int2*>(smem_anch + (size_t)npent * 8);      // [256]
    int* sdelta = reinterpret_cast<int*>(smem_anch + (size_t)npent * 8 + 256 * 8);   // [npent]
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int t0 = (int)blockIdx.x * kt, t1 = min(t0 + kt, ntiles);
    int code[RPT], anc[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {                          // static data: requested before the predecessor has finished
        const int r = min(t0 * T + tid + j * THREADS, re - 1);
        code[j] = ld_stream_u8(rcodes + r);
        anc[j] = ld_stream_i32(anchor + r);
    }
    for (int k = tid; k < npent; k += THREADS) { const DictEnt d = pent[k]; sval[k] = d.val; sdelta[k] = d.delta; }
    for (int k = tid; k < ndict; k += THREADS) sphead[k] = phead[k];
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();                                         // the table is in shared memory
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
        const int row0 = t * T;
        const bool full = row0 + T <= re;                    // CTA-uniform
        double o[RPT][NOPS > 0 ? NOPS : 1];
        int io[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = full ? row0 + tid + j * THREADS : min(row0 + tid + j * THREADS, re - 1);
#pragma unroll
            for (int k = 0; k < NOPS; ++k) o[j][k] = ld_stream_f64(epi.operand(k) + r);
            if constexpr (NIOPS > 0) io[j] = epi.ioperand()[r];
        }
        int2 ph[RPT];
        double xv[RPT][JW];
        bool more[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            ph[j] = sphead[code[j]];
            const int* sd = sdelta + ph[j].x;
#pragma unroll
            for (int e = 0; e < HW; ++e) xv[j][e] = x[anc[j] + sd[e]];        // (coherent load; padded entries repeat the last one)
            more[j] = __any_sync(0xffffffffu, ph[j].y > HW);
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            if (more[j]) {
                const int* sd = sdelta + ph[j].x;
#pragma unroll
                for (int e = HW; e < JW; ++e) xv[j][e] = x[anc[j] + sd[e]];
            }
        }
        int canc[RPT];                                       // (rows longer than JW entries still need this tile's anchors)
#pragma unroll
        for (int j = 0; j < RPT; ++j) canc[j] = anc[j];
        if (t + 1 < t1) {                                    // the next tile's anchors and codes travel while this tile is summed
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int r = min(row0 + T + tid + j * THREADS, re - 1);
                code[j] = ld_stream_u8(rcodes + r);
                anc[j] = ld_stream_i32(anchor + r);
            }
        }
        if (pf > 0 && tid < 2 + NOPS + (NIOPS > 0 ? 1 : 0) && t + pf < ntiles) {     // L2 prefetch for the tile pf tiles ahead
            const long long p0 = (long long)row0 + (long long)pf * T;
            if (p0 + T <= (long long)re) {
                if (tid == 0) bulk_prefetch_l2(rcodes + p0, T);
                if (tid == 1) bulk_prefetch_l2(anchor + p0, T * 4);
                if (tid >= 2 && tid < 2 + NOPS) bulk_prefetch_l2(epi.operand(tid - 2) + p0, T * 8);
                if constexpr (NIOPS > 0) if (tid == 2 + NOPS) bulk_prefetch_l2(epi.ioperand() + p0, T * 4);
            }
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int r = row0 + tid + j * THREADS;
            double sum = 0.0;                                // one accumulator, stored order
            const double* sv = sval + ph[j].x;
#pragma unroll
            for (int e = 0; e < HW; ++e)
                if (e < ph[j].y) sum = __dadd_rn(sum, __dmul_rn(sv[e], xv[j][e]));
            if (more[j]) {
#pragma unroll
                for (int e = HW; e < JW; ++e)
                    if (e < ph[j].y) sum = __dadd_rn(sum, __dmul_rn(sv[e], xv[j][e]));
            }
            for (int e0 = JW; e0 < ph[j].y; e0 += JW) {      // rows longer than JW entries
                double xw[JW];
#pragma unroll
                for (int e = 0; e < JW; ++e) xw[e] = x[canc[j] + sdelta[ph[j].x + e0 + e]];
#pragma unroll
                for (int e = 0; e < JW; ++e)
                    if (e0 + e < ph[j].y) sum = __dadd_rn(sum, __dmul_rn(sval[ph[j].x + e0 + e], xw[e]));
            }
            if (full || r < re) {
                if constexpr (NIOPS > 0) epi.store_i(r, sum, o[j], io[j]);
                else epi.store(r, sum, o[j]);
            }
        }
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
    decltype(epi_load(epi, 0)) pre;
    if (active && lane == 0) pre = epi_load(epi, row);
    double s = 0.0;
    for (int k = a + lane; k < b; k += LPR) s = __dadd_rn(s, __dmul_rn(vals[k], ld_x<NCX>(x, cols[k])));
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (active && lane == 0) epi_store(epi, row, s, pre);
}

// ---- long-row fallback --------------------------------------------------------------------------------
// Rows longer than a tile (thousands of entries) cannot be staged by the tile / stream kernels.  One warp per row:
// the lanes form 32 products at a time, then every lane adds them IN STORED ORDER (the products are passed round with
// shuffles), so the numerics contract of the tile family -- one accumulator, stored order, no FMA -- still holds.
template <bool NCX, class Epi>
__global__ void __launch_bounds__(256)
k_seqrow(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, const double* __restrict__ vals,
         int row_begin, int row_end, const double* x, Epi epi)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int row = row_begin + warp;
    if (row >= row_end) return;                       // whole warp leaves together
    const int a = rowptr[row], b = rowptr[row + 1];
    decltype(epi_load(epi, 0)) pre;
    if (lane == 0) pre = epi_load(epi, row);
    double s = 0.0;
    for (int k0 = a; k0 < b; k0 += 32) {
        const int k = k0 + lane;
        const double p = k < b ? __dmul_rn(vals[k], ld_x<NCX>(x, cols[k])) : 0.0;
        const int m = min(32, b - k0);
        for (int l = 0; l < m; ++l) s = __dadd_rn(s, __shfl_sync(0xffffffffu, p, l));
    }
    if (lane == 0) epi_store(epi, row, s, pre);
}

// ---- small kernels ----------------------------------------------------------------------------------
// zero initial guess + one Jacobi sweep collapses to v = g = w*(dinv*f)   (multigrid.py:253 + :226)
// (sharded: the rows the neighbours hold as ghosts go to them as well, see HaloFuse)
__global__ void k_init_guess(int n, const double* __restrict__ dinv, const double* __restrict__ f, double om,
                             double* __restrict__ g, double* __restrict__ v, const __grid_constant__ HaloFuse hf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double gi = __dmul_rn(om, __dmul_rn(dinv[i], f[i]));
        g[i] = gi;
        v[i] = gi;
        if (hf.send_n) halo_send(hf, i, gi);
    }
    if (hf.send_n) halo_publish(hf, (int)(blockIdx.x * blockDim.x), min((int)((blockIdx.x + 1) * blockDim.x), n));
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
// (one definition of the row's dot product: the stand-alone kernel and the coarse-tail kernel must give the same bits)
__device__ __forceinline__ double dense_row_dot(int n, const double* __restrict__ row, const double* x, int lane)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int j = lane;
    for (; j + 96 < n; j += 128) {
        s0 = fma(row[j], __ldcg(x + j), s0);
        s1 = fma(row[j + 32], __ldcg(x + j + 32), s1);
        s2 = fma(row[j + 64], __ldcg(x + j + 64), s2);
        s3 = fma(row[j + 96], __ldcg(x + j + 96), s3);
    }
    for (; j < n; j += 32) s0 = fma(row[j], __ldcg(x + j), s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
__global__ void __launch_bounds__(256)
k_dense_gemv(int n, const double* __restrict__ M, const double* x, const double* y0, double* y)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double s = dense_row_dot(n, M + (size_t)warp * n, x, lane);
    if (lane == 0) y[warp] = y0 ? y0[warp] + s : s;
}

// ---- coarse tail: every level below a size threshold in ONE cooperative launch ----------------------------------------
// Below ~3e5 rows a kernel is all launch latency: the cfg2 cycle spent a third of its time in ~30 launches of 8-12 us on levels
// whose data fit L2 many times over (profiles/r1_*), and on 8 GPUs those levels run on rank 0 alone while 7 GPUs wait.  This
// kernel runs the whole sub-cycle of those levels -- zero-guess sweep, sweeps, residual + injection, ..., dense coarsest apply,
// prolongation + correction, sweeps (multigrid.py:238-261, recursion unrolled) -- with a grid-wide barrier between phases
// instead of a launch.  Arithmetic is the same as in the per-level kernels: one accumulator per row, stored entry order,
// separately rounded multiply and add; the operators are read in whichever form the level holds (CSR, row patterns, anchored
// row patterns).  Vectors are read with L2-only loads (ld.global.cg): a buffer is rewritten by other SMs two phases after it
// was read, and a grid barrier does not drop L1 lines.
struct TailOp {
    int mode;                                // 0 CSR, 3 row patterns, 4 anchored row patterns
    const int32_t* rowptr; const int32_t* cols; const double* vals;
    const unsigned char* codes; const int32_t* anchor; const int2* phead; const DictEnt* pent;
};
struct TailLevel {
    int n, nc;                               // rows; rows of the next coarser level (0 on the coarsest)
    TailOp A, RJ, P;                         // level matrix, smoother matrix, prolongation from the next coarser level
    const int32_t* inj;                      // fine row of every coarse row
    const double* dinv;
    double *f, *g, *a, *b;                   // right-hand side, w*(dinv*f), the iterate's two buffers (a holds the zero-guess sweep)
};
constexpr int TAIL_MAX_LEVELS = 6;
struct TailPlan {
    int nlev, mu1, mu2;                      // lev[0] = coarsest ... lev[nlev - 1] = top of the tail
    double om, om1;
    const double* coarse_inv;                // dense inverse of the coarsest matrix, row-major
    TailLevel lev[TAIL_MAX_LEVELS];
};

__device__ __forceinline__ double tail_rowsum(const TailOp& D, int r, const double* x)
{
    double sum = 0.0;                        // one accumulator, stored order
    if (D.mode == 3 || D.mode == 4) {
        const int2 ph = __ldg(D.phead + D.codes[r]);
        const int base = D.mode == 3 ? r : __ldg(D.anchor + r);
        for (int e = 0; e < ph.y; ++e) {
            const double val = __ldg(&D.pent[ph.x + e].val);
            const int dl = __ldg(&D.pent[ph.x + e].delta);
            sum = __dadd_rn(sum, __dmul_rn(val, __ldcg(x + base + dl)));
        }
    } else {
        const int a = __ldg(D.rowptr + r), b = __ldg(D.rowptr + r + 1);
        for (int k = a; k < b; ++k) sum = __dadd_rn(sum, __dmul_rn(__ldg(D.vals + k), __ldcg(x + __ldg(D.cols + k))));
    }
    return sum;
}

// CLUSTER: the grid is ONE thread-block cluster and the phases are separated by the hardware cluster barrier (release / acquire
// at cluster scope; vectors are read past L1 anyway) instead of the cooperative grid barrier, which costs ~6 us per phase.
template <bool CLUSTER>
__global__ void __launch_bounds__(512)
k_tail(const __grid_constant__ TailPlan T)
{
    auto phase_sync = [] {
        if constexpr (CLUSTER) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        else cg::this_grid().sync();
    };
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
    const double om = T.om, om1 = T.om1;
    auto sweep = [&](const TailLevel& L, const double* cur, double* oth) {     // multigrid.py:226
        for (int r = gt; r < L.n; r += gs) {
            const double s = tail_rowsum(L.RJ, r, cur);
            oth[r] = __dsub_rn(__dadd_rn(__dmul_rn(om1, __ldcg(cur + r)), __ldcg(L.g + r)), __dmul_rn(om, s));
        }
    };
    for (int k = T.nlev - 1; k >= 1; --k) {                  // ---- down: multigrid.py:243-253
        const TailLevel& L = T.lev[k];
        double* cur = L.a; double* oth = L.b;
        for (int r = gt; r < L.n; r += gs) {                 // zero guess + first sweep: v = g = w*(dinv*f)
            const double gi = __dmul_rn(om, __dmul_rn(__ldg(L.dinv + r), __ldcg(L.f + r)));
            L.g[r] = gi; cur[r] = gi;
        }
        phase_sync();
        for (int s = 1; s < T.mu1; ++s) { sweep(L, cur, oth); phase_sync(); double* t = cur; cur = oth; oth = t; }
        double* fc = T.lev[k - 1].f;
        for (int i = gt; i < L.nc; i += gs) {                // residual at the injected rows only (multigrid.py:244 + :128-131)
            const int r = __ldg(L.inj + i);
            fc[i] = __dsub_rn(__ldcg(L.f + r), tail_rowsum(L.A, r, cur));
        }
        phase_sync();
    }
    {                                                        // ---- coarsest: u = A^-1 f (multigrid.py:238-241)
        const TailLevel& C = T.lev[0];
        const int lane = threadIdx.x & 31, nw = gs >> 5;
        for (int row = gt >> 5; row < C.n; row += nw) {
            const double s = dense_row_dot(C.n, T.coarse_inv + (size_t)row * C.n, C.f, lane);
            if (lane == 0) C.a[row] = s;
        }
        phase_sync();
    }
    for (int k = 1; k < T.nlev; ++k) {                       // ---- up: multigrid.py:258-261
        const TailLevel& L = T.lev[k];
        const TailLevel& C = T.lev[k - 1];
        const double* e = (k == 1 || (((T.mu1 - 1) + T.mu2) & 1) == 0) ? C.a : C.b;     // where the coarser level's iterate ended
        double* cur = ((T.mu1 - 1) & 1) ? L.b : L.a;
        double* oth = ((T.mu1 - 1) & 1) ? L.a : L.b;
        for (int r = gt; r < L.n; r += gs) cur[r] = __dadd_rn(__ldcg(cur + r), tail_rowsum(L.P, r, e));
        phase_sync();
        for (int s = 0; s < T.mu2; ++s) { sweep(L, cur, oth); phase_sync(); double* t = cur; cur = oth; oth = t; }
    }
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
// x . y with the same fixed trees (mass-matrix norm sqrt(r^T M r) of the FMG stopping rule, multigrid.py:203-208)
__global__ void __launch_bounds__(256)
k_dot_partial(int64_t n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ partial)
{
    __shared__ double sh[8];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fma(x[i], y[i], s);
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
__global__ void k_sumsq_final(int nb, const double* __restrict__ partial, double* __restrict__ out, int take_root)
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
        *out = take_root ? sqrt(fmax(t, 0.0)) : t;          // (r^T M r can round below zero for a tiny r: never a NaN norm)
    }
}
// d = a - b   (error of the iterate against the exact solution, err_calculator multigrid.py:213-218)
__global__ void k_diff(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ d)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __dsub_rn(a[i], b[i]);
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

// Same sweep, but the whole level schedule runs inside ONE thread-block cluster: the barrier between dependency
// levels is the hardware cluster barrier (barrier.cluster arrive.release / wait.acquire, ~0.2 us) instead of a grid-wide
// software barrier (~2 us).  Used when the widest level fits a cluster (2-D problems: <= N rows per level).
__global__ void __launch_bounds__(512)
k_gs_levels_cluster(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols, const double* __restrict__ vals,
                    const int32_t* __restrict__ order, const double* __restrict__ diag, const double* __restrict__ f,
                    double* v, const int32_t* __restrict__ lvl_off, int nlev)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = (int)cluster.block_rank() * blockDim.x + threadIdx.x;
    const int nth = (int)cluster.num_blocks() * blockDim.x;
    for (int L = 0; L < nlev; ++L) {
        const int p0 = lvl_off[L], p1 = lvl_off[L + 1];
        for (int p = p0 + tid; p < p1; p += nth) {
            const int a = rowptr[p], b = rowptr[p + 1];
            double s = 0.0;
            for (int k = a; k < b; ++k) s = __dadd_rn(s, __dmul_rn(vals[k], __ldcg(v + cols[k])));
            const int i = order[p];
            __stcg(v + i, __ddiv_rn(__dsub_rn(f[i], s), diag[p]));
        }
        cluster.sync();
    }
}

// Level-scheduled Gauss-Seidel, pipelined.  The per-level critical path of k_gs_levels[_cluster] is a chain of
// dependent loads (row pointers -> entries -> x gathers -> order -> f) plus the barrier.  Here the operator is
// stored in ELL form (W <= 8 entries per row, level-major, column-major inside the array so that a level's rows are
// coalesced) and everything that does not depend on the sweep itself -- entries, row index, f, diagonal -- is
// fetched one / two dependency levels AHEAD, into registers that alternate between two sets (the loop is unrolled
// by two, so no register copy waits on a load).  What remains per level: the x gathers (L2) and the cluster barrier.
// The level offsets live in shared memory.  Entries beyond a row's length are (col 0, value 0).
struct GsRow { int i; double f, d; int c[8]; double a[8]; };

template <int W>
__device__ __forceinline__ void gs_fetch(GsRow& r, int p, bool on, int iord, int n, const int32_t* __restrict__ ecols,
                                         const double* __restrict__ evals, const double* __restrict__ diag, const double* __restrict__ f)
{
    if (!on) return;
    r.i = iord;
    r.f = f[iord];
    r.d = diag[p];
#pragma unroll
    for (int k = 0; k < W; ++k) { r.c[k] = ecols[(size_t)k * n + p]; r.a[k] = evals[(size_t)k * n + p]; }
}

template <int W>
__device__ __forceinline__ void gs_apply(const GsRow& r, bool on, double* v)
{
    if (!on) return;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < W; ++k) s = __dadd_rn(s, __dmul_rn(r.a[k], __ldcg(v + r.c[k])));
    __stcg(v + r.i, __ddiv_rn(__dsub_rn(r.f, s), r.d));
}

template <int W>
__global__ void __launch_bounds__(512)
k_gs_levels_ell(int n, const int32_t* __restrict__ ecols, const double* __restrict__ evals, const int32_t* __restrict__ order,
                const double* __restrict__ diag, const double* __restrict__ f, double* v, const int32_t* __restrict__ lvl_off, int nlev)
{
    extern __shared__ int s_off[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = (int)cluster.block_rank() * blockDim.x + threadIdx.x;
    const int nth = (int)cluster.num_blocks() * blockDim.x;
    for (int k = threadIdx.x; k <= nlev; k += blockDim.x) s_off[k] = lvl_off[k];
    __syncthreads();
    auto row_of = [&](int L, int& p) { p = (L < nlev) ? s_off[L] + tid : 0; return L < nlev && p < s_off[L + 1]; };
    GsRow ra, rb;
    int p0, p1, p2;
    bool on0 = row_of(0, p0), on1 = row_of(1, p1), on2;
    int i1 = on1 ? order[p1] : 0, i2;
    gs_fetch<W>(ra, p0, on0, on0 ? order[p0] : 0, n, ecols, evals, diag, f);
    for (int L = 0; L < nlev; L += 2) {
        // ---- level L from set a; fetch level L+1 into set b; row index of level L+2
        on2 = row_of(L + 2, p2);
        i2 = on2 ? order[p2] : 0;
        gs_fetch<W>(rb, p1, on1, i1, n, ecols, evals, diag, f);
        gs_apply<W>(ra, on0, v);
        for (int p = s_off[L] + tid + nth; p < s_off[L + 1]; p += nth) {       // levels wider than the cluster: plain path
            GsRow t; gs_fetch<W>(t, p, true, order[p], n, ecols, evals, diag, f); gs_apply<W>(t, true, v);
        }
        cluster.sync();
        if (L + 1 >= nlev) break;
        // ---- level L+1 from set b; fetch level L+2 into set a; row index of level L+3
        on0 = on2; p0 = p2;
        const int i0 = i2;
        on1 = row_of(L + 3, p1);
        i1 = on1 ? order[p1] : 0;
        gs_fetch<W>(ra, p0, on0, i0, n, ecols, evals, diag, f);
        gs_apply<W>(rb, true && (s_off[L + 1] + tid < s_off[L + 2]), v);
        for (int p = s_off[L + 1] + tid + nth; p < s_off[L + 2]; p += nth) {
            GsRow t; gs_fetch<W>(t, p, true, order[p], n, ecols, evals, diag, f); gs_apply<W>(t, true, v);
        }
        cluster.sync();
    }
}

}  // namespace mgb
