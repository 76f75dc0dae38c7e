// Micro-benchmark + bit-exactness check of the two-sweeps-per-launch kernel (k_hotrow2, mgb_kernels.cuh) against two launches of
// k_hotrow, on the weighted-Jacobi sweep of
// the 3-D 7-point smoother matrix (N^3 nodes, lexicographic numbering, Dirichlet boundary rows empty, rows next to the
// boundary carrying sub-patterns) -- the dominant kernel of BASELINE config 5 -- against a plain table-walk kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/ubench/hotrow2 tools/ubench/hotrow2.cu
//   tools/ubench/hotrow2 [N] > gpurun_out/r2_hotrow2.jsonl
// Persistent (grid-sized, tile-walking) variants were measured with an earlier version of the kernel and dropped: no better without
// prefetch (680 us against 681) and far worse with a per-warp prefetch (990 us); results in profiles/r2_hotrow_ubench_513.jsonl.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../multigrid_dolfinx_b200/csrc/mgb_kernels.cuh"

using namespace mgb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// code of row (z, y, x): 64 = Dirichlet row (empty); else 6 bits, bit e set when neighbour e (stored order -N^2, -N, -1, +1, +N, +N^2)
// is an interior node; code 65: an artificial pattern that is NOT a sub-pattern of the hot one (exercises the table walk)
__global__ void k_make_codes(int N, unsigned char* codes)
{
    const long long n = (long long)N * N * N;
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int x = (int)(r % N), y = (int)((r / N) % N), z = (int)(r / ((long long)N * N));
    auto bnd = [&](int a) { return a == 0 || a == N - 1; };
    if (bnd(x) || bnd(y) || bnd(z)) { codes[r] = 64; return; }
    int m = 0;
    if (!bnd(z - 1)) m |= 1;
    if (!bnd(y - 1)) m |= 2;
    if (!bnd(x - 1)) m |= 4;
    if (!bnd(x + 1)) m |= 8;
    if (!bnd(y + 1)) m |= 16;
    if (!bnd(z + 1)) m |= 32;
    if (z == N / 2 && y == N / 3 && (x % 7) == 3) { codes[r] = 65; return; }
    codes[r] = (unsigned char)m;
}

__global__ void k_fill(long long n, double* a, unsigned seed)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ULL + seed;
    h ^= h >> 31; h *= 0xff51afd7ed558ccdULL; h ^= h >> 29;
    a[i] = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// reference: thread per row, table walk, same numerics contract
template <class Epi>
__global__ void k_ref(long long n, const unsigned char* codes, const int2* phead, const DictEnt* pent, const double* x, Epi epi)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int2 ph = phead[codes[r]];
    double sum = 0.0;
    for (int e = 0; e < ph.y; ++e) sum = __dadd_rn(sum, __dmul_rn(pent[ph.x + e].val, x[r + pent[ph.x + e].delta]));
    double o[Epi::NOPS];
    for (int k = 0; k < Epi::NOPS; ++k) o[k] = epi.operand(k)[r];
    epi.store((int)r, sum, o);
}

__global__ void k_cmp(long long n, const double* a, const double* b, unsigned long long* bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && __double_as_longlong(a[i]) != __double_as_longlong(b[i])) atomicAdd(bad, 1ULL);
}

struct Problem {
    int N; long long n; int xlen;
    unsigned char* codes; uint32_t* pmask; int2* phead; DictEnt* pent; HotArgs H;
    double *x, *g, *out, *ref;
    unsigned long long* bad;
};

template <int THREADS, int RPT, int MINB, bool XL2, bool POL = false>
void run_pair(Problem& P, int pf_rows, int slack, int kt, const char* name, int reps)
{
    constexpr int T = THREADS * RPT;
    const int pf = pf_rows / T;
    const int ntiles = (int)((P.n + T - 1) / T);
    const int pf_last = (int)(std::min<long long>(P.n / T, (P.xlen - (P.H.dmax & ~1)) / T) - 1);
    auto k1 = k_hotrow<6, THREADS, RPT, MINB, false, EpiJacobiRJ>;
    auto k2 = k_hotrow2<6, THREADS, RPT, MINB, XL2, POL, EpiJacobiRJ, EpiJacobiRJ>;
    const int reach = (int)((std::max(-P.H.dmin, P.H.dmax) + T - 1) / T) + 1;
    const int groups = (ntiles + S2_GROUP - 1) / S2_GROUP;
    const int nchunks = (ntiles + kt - 1) / kt;
    const int lagc = (reach + S2_GROUP + std::max(slack, 0) + kt - 1) / kt + 1;
    const int lag = lagc * kt;
    const int grid2 = 2 * (nchunks + lagc);
    static unsigned long long* s2 = nullptr;
    if (!s2) { CK(cudaMalloc(&s2, (groups + 2) * 8)); CK(cudaMemset(s2, 0, (groups + 2) * 8)); }
    auto pair = [&](double* v, double* tmp) {               // two sweeps, result back in v
        EpiJacobiRJ e1{v, P.g, tmp, 1.0 - 2.0 / 3.0, 2.0 / 3.0}, e2{tmp, P.g, v, 1.0 - 2.0 / 3.0, 2.0 / 3.0};
        if (slack < 0) {
            k1<<<ntiles, THREADS>>>(P.codes, P.pmask, P.phead, P.pent, P.H, HaloFuse{}, nullptr, ntiles, 0, (int)P.n, P.xlen, pf, pf_last, v, e1);
            k1<<<ntiles, THREADS>>>(P.codes, P.pmask, P.phead, P.pent, P.H, HaloFuse{}, nullptr, ntiles, 0, (int)P.n, P.xlen, pf, pf_last, tmp, e2);
        } else {
            k2<<<grid2, THREADS>>>(P.codes, P.pmask, P.phead, P.pent, P.H, ntiles, (int)P.n, P.xlen, pf, pf_last, v, e1, tmp, e2, s2, kt, lagc, reach);
            k_s2_reset<<<(groups + 255) / 256, 256>>>(groups, s2);
        }
    };
    // correctness: x -> (out) -> x against ref2 = two reference sweeps
    CK(cudaMemset(P.out, 0xFF, sizeof(double) * P.n));
    pair(P.x, P.out);
    CK(cudaMemset(P.bad, 0, 8));
    k_cmp<<<(int)((P.n + 255) / 256), 256>>>(P.n, P.x, P.ref, P.bad);
    unsigned long long bad = 0;
    CK(cudaMemcpy(&bad, P.bad, 8, cudaMemcpyDeviceToHost));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) pair(P.x, P.out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) pair(P.x, P.out);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double us = ms * 1e3 / reps;
    printf("{\"variant\": \"%s\", \"N\": %d, \"threads\": %d, \"rpt\": %d, \"pf_rows\": %d, \"slack\": %d, \"kt\": %d, \"lag\": %d, \"us_per_pair\": %.2f, \"GBs_at_50B\": %.1f, \"mismatches\": %llu}\n",
           name, P.N, THREADS, RPT, pf_rows, slack, kt, lag, us, 50.0 * (double)P.n / us / 1e3, bad);
    fflush(stdout);
    k_fill<<<(int)((P.n + 255) / 256), 256>>>(P.n, P.x, 1u);
    CK(cudaDeviceSynchronize());
}

int main(int argc, char** argv)
{
    Problem P{};
    P.N = argc > 1 ? atoi(argv[1]) : 513;
    const int N = P.N;
    P.n = (long long)N * N * N;
    P.xlen = (int)(P.n + 16);
    CK(cudaMalloc(&P.codes, P.n + 4096));
    CK(cudaMalloc(&P.x, sizeof(double) * (P.n + 16))); CK(cudaMalloc(&P.g, sizeof(double) * (P.n + 16)));
    CK(cudaMalloc(&P.out, sizeof(double) * (P.n + 16))); CK(cudaMalloc(&P.ref, sizeof(double) * (P.n + 16)));
    CK(cudaMalloc(&P.bad, 8));
    CK(cudaMemset(P.codes, 64, P.n + 4096));
    CK(cudaMemset(P.x, 0, sizeof(double) * (P.n + 16))); CK(cudaMemset(P.out, 0, sizeof(double) * (P.n + 16)));
    const int gridn = (int)((P.n + 255) / 256);
    k_make_codes<<<gridn, 256>>>(N, P.codes);
    k_fill<<<gridn, 256>>>(P.n, P.x, 1u);
    k_fill<<<gridn, 256>>>(P.n, P.g, 2u);
    // pattern table: codes 0..63 sub-patterns, 64 empty, 65 foreign
    const int hd[6] = {-N * N, -N, -1, 1, N, N * N};
    const double w = (1.0 / 6.0) * -1.0;
    std::vector<int2> phead(256, make_int2(0, 0));
    std::vector<DictEnt> pent;
    std::vector<uint32_t> pmask(256, 0);
    for (int c = 0; c < 64; ++c) {
        phead[c].x = (int)pent.size();
        int len = 0;
        for (int e = 0; e < 6; ++e) if ((c >> e) & 1) { pent.push_back(DictEnt{w, hd[e], 0}); ++len; }
        phead[c].y = len;
        pmask[c] = (uint32_t)c;
    }
    phead[65] = make_int2((int)pent.size(), 3);
    pent.push_back(DictEnt{0.25, -2, 0}); pent.push_back(DictEnt{-0.125, 0, 0}); pent.push_back(DictEnt{0.5, 2 * N, 0});
    pmask[65] = HOT_SLOW;
    CK(cudaMalloc(&P.phead, 256 * sizeof(int2))); CK(cudaMemcpy(P.phead, phead.data(), 256 * sizeof(int2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&P.pent, pent.size() * sizeof(DictEnt))); CK(cudaMemcpy(P.pent, pent.data(), pent.size() * sizeof(DictEnt), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&P.pmask, 256 * 4)); CK(cudaMemcpy(P.pmask, pmask.data(), 256 * 4, cudaMemcpyHostToDevice));
    P.H.dmin = hd[0]; P.H.dmax = hd[5];
    for (int e = 0; e < WIN_HOT; ++e) { P.H.hd[e] = e < 6 ? hd[e] : 0; P.H.hv[e] = e < 6 ? w : 0.0; }
    {   // reference: two table-walk sweeps x -> out -> ref
        EpiJacobiRJ e1{P.x, P.g, P.out, 1.0 - 2.0 / 3.0, 2.0 / 3.0}, e2{P.out, P.g, P.ref, 1.0 - 2.0 / 3.0, 2.0 / 3.0};
        k_ref<EpiJacobiRJ><<<gridn, 256>>>(P.n, P.codes, P.phead, P.pent, P.x, e1);
        k_ref<EpiJacobiRJ><<<gridn, 256>>>(P.n, P.codes, P.phead, P.pent, P.out, e2);
        CK(cudaDeviceSynchronize());
    }
    const int reps = N >= 400 ? 10 : 40;
    if (argc > 2 && atoi(argv[2]) == 1) {                    // short list for ncu
        run_pair<128, 2, 8, false>(P, 1 << 18, -1, 1, "two_launches", 1);
        run_pair<128, 2, 8, false>(P, 1 << 18, 8192, 8, "k_hotrow2", 1);
        run_pair<128, 2, 8, false>(P, 1 << 18, 16384, 4, "k_hotrow2", 1);
        run_pair<128, 2, 8, false>(P, 1 << 18, 2048, 1, "k_hotrow2", 1);
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 3) {                    // L2 eviction priorities
        run_pair<128, 2, 8, false>(P, 1 << 18, -1, 1, "two_launches", reps);
        for (int kt : {1, 2, 4, 8})
            for (int slack : {1024, 2048, 4096, 8192, 16384}) {
                run_pair<128, 2, 8, false, true>(P, 1 << 18, slack, kt, "k_hotrow2_pol", reps);
            }
        run_pair<128, 2, 8, false, true>(P, 0, 4096, 4, "k_hotrow2_pol_nopf", reps);
        run_pair<128, 2, 8, false, true>(P, 1 << 16, 4096, 4, "k_hotrow2_pol_pf64k", reps);
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 5) {                    // around the sweet spot
        run_pair<128, 2, 8, false>(P, 1 << 18, -1, 1, "two_launches", reps);
        for (int slack : {2560, 3072, 3584, 4096, 4608, 5120, 6144})
            run_pair<128, 2, 8, false, true>(P, 1 << 18, slack, 4, "k_hotrow2_pol", reps);
        for (int slack : {1536, 2048, 2560, 3072})
            run_pair<128, 2, 8, false, true>(P, 1 << 18, slack, 2, "k_hotrow2_pol", reps);
        for (int slack : {4096, 5120, 6144, 7168})
            run_pair<128, 2, 8, false, true>(P, 1 << 18, slack, 8, "k_hotrow2_pol", reps);
        for (int slack : {3072, 4096, 6144}) {
            run_pair<256, 2, 4, false, true>(P, 1 << 18, slack, 2, "k_hotrow2_pol_t256r2", reps);
            run_pair<256, 1, 8, false, true>(P, 1 << 18, slack, 4, "k_hotrow2_pol_t256r1", reps);
            run_pair<128, 1, 16, false, true>(P, 1 << 18, slack, 8, "k_hotrow2_pol_t128r1", reps);
        }
        run_pair<128, 2, 8, false>(P, 1 << 18, -1, 1, "two_launches", reps);
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 4) {                    // short list for ncu
        run_pair<128, 2, 8, false, true>(P, 1 << 18, 8192, 8, "k_hotrow2_pol", 1);
        run_pair<128, 2, 8, false, true>(P, 1 << 18, 4096, 4, "k_hotrow2_pol", 1);
        run_pair<128, 2, 8, false, true>(P, 1 << 18, 2048, 2, "k_hotrow2_pol", 1);
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 2) {
        for (int kt : {2, 4, 8})
            for (int slack : {4096, 8192, 12288, 16384, 24576, 32768})
                run_pair<128, 2, 8, false>(P, 1 << 18, slack, kt, "k_hotrow2", reps);
        return 0;
    }
    run_pair<128, 2, 8, false>(P, 1 << 18, -1, 1, "two_launches", reps);
    for (int kt : {1, 4, 8, 16, 32, 64})
        for (int slack : {512, 2048, 8192})
            run_pair<128, 2, 8, false>(P, 1 << 18, slack, kt, "k_hotrow2", reps);
    run_pair<128, 2, 8, true>(P, 1 << 18, 2048, 16, "k_hotrow2_xl2", reps);
    run_pair<128, 2, 8, false>(P, 0, 2048, 16, "k_hotrow2_nopf", reps);
    run_pair<256, 1, 8, false>(P, 1 << 18, 2048, 16, "k_hotrow2_t256r1", reps);
    run_pair<128, 2, 8, false>(P, 1 << 18, 0, 16, "k_hotrow2_slack0", reps);
    run_pair<128, 2, 8, false>(P, 1 << 18, -1, 1, "two_launches", reps);
    printf("{\"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
