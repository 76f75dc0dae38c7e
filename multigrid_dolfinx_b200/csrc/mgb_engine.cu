// The device-resident level hierarchy and the V-cycle (drop-in for V_cycle_scheme, multigrid.py:231-268)
// behind the C ABI of include/mgb200.h.  One handle = one device + one stream; the whole cycle is a
// fixed sequence of launches, captured once per top level into a CUDA graph and replayed.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mgb200.h"
#include "mgb_internal.h"
#include "mgb_devsetup.h"
#include "mgb_kernels.cuh"

using namespace mgb;

#include "mgb_types.cuh"

namespace {

int fail(mgb_handle* h, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(h, MGB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(call)                 \
    do {                          \
        int rc_ = (call);         \
        if (rc_ != MGB_OK) return rc_; \
    } while (0)

template <class T>
int dev_alloc(mgb_handle* h, T** p, size_t count)
{
    *p = nullptr;
    CU(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return MGB_OK;
}

template <class T>
int dev_upload(mgb_handle* h, T** p, const T* src, size_t count, size_t pad = 0)
{
    TRY(dev_alloc(h, p, count + pad));
    if (pad) CU(cudaMemsetAsync(*p + count, 0, pad * sizeof(T), h->stream));
    if (count) CU(cudaMemcpyAsync(*p, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return MGB_OK;
}

void free_csr(DevCsr& D)
{
    cudaFree(D.cd.codes); cudaFree(D.cd.dict); cudaFree(D.cd.phead); cudaFree(D.cd.dict_win);
    cudaFree(D.rowptr); cudaFree(D.cols); cudaFree(D.vals); cudaFree(D.tiles); cudaFree(D.sdesc); cudaFree(D.sdesc_bnd);
    D = DevCsr();
}

int tile_cap(int iter) { return TILE_ENT * THREADS * iter; }

// TMA stream kernel configurations (option "stream_cfg"): consumer threads, entries per thread, stages
struct StreamChoice { int threads, ept, stages; };
StreamChoice stream_choice(int cfg)
{
    switch (cfg) {
        case 1: return {256, 8, 2};      // 2048-entry tiles, 2 CTAs/SM
        case 2: return {512, 4, 2};      // 2048-entry tiles, 2 CTAs/SM, twice the consumer threads
        case 3: return {256, 4, 2};      // 1024-entry tiles, 4 CTAs/SM (592 persistent CTAs) -- the default, fastest measured
        case 4: return {256, 4, 3};      // 1024-entry tiles, 3 CTAs/SM
        case 5: return {128, 8, 2};      // 1024-entry tiles, half the consumer threads
        case 7: return {128, 4, 2};      // 512-entry tiles, ~10 CTAs/SM
        default: return {256, 8, 3};     // 2048-entry tiles, 1 CTA/SM, deep ring (consumer-bound: slowest)
    }
}

// Row-stream kernel configurations for dictionary-coded operators (option "code_cfg"): consumer threads, rows per
// thread, entries per row the stage has room for, stages
struct CodeChoice { int threads, rpt, epr, stages; };
CodeChoice code_choice(int cfg)
{
    switch (cfg) {
        // measured and removed (profiles/r1_variant_sweep_cfg2_coded.jsonl): 256-row tiles with one row per thread (32.9 us per
        // fine-level sweep against 28.2), 1024-row tiles with four rows per thread (29.1), a register budget for 5 CTAs per SM (28.1)
        case 3: return {256, 2, 8, 3};       // 512-row tiles, deeper ring (27.8 us, but the prolongation loses: 43 against 32 us)
        default: return {256, 2, 8, 2};      // 512-row tiles of <= 4096 entries, two rows per thread, two stages -- the default
    }
}

// Row-window kernel configurations (option "win_cfg"): rows per thread (x 256 consumer threads = rows per tile), stages
struct WinChoice { int rpt, stages; };
WinChoice win_choice(int cfg)
{
    switch (cfg) {
        case 2: return {4, 2};           // 1024-row tiles, two stages, registers for 2 CTAs per SM
        case 3: return {2, 3};           // 512-row tiles, three stages, registers for 3 CTAs per SM
        case 4: return {4, 3};           // 1024-row tiles, three stages, registers for 2 CTAs per SM
        default: return {2, 2};          // 512-row tiles, two stages, registers for 4 CTAs per SM
    }
}

// Hot-row kernel configurations (option "hot_cfg"): threads per CTA, rows per thread (their product = rows per tile).
// Measured on the 513^3 smoother sweep (tools/ubench/hotrow.cu, profiles/r2_hotrow_ubench_513.jsonl): 256-row tiles win
// (128-row tiles are bound by the CTA launch rate, 512-row and larger tiles by the gap a retiring CTA leaves), and
// 128 threads x 2 rows beats 256 x 1 and 64 x 4.
struct HotChoice { int threads, rpt; };
HotChoice hot_choice(int cfg)
{
    switch (cfg) {
        case 2: return {256, 1};
        case 3: return {128, 1};
        case 4: return {256, 2};
        default: return {128, 2};
    }
}
bool hot_len_supported(int hotlen) { return hotlen == 4 || hotlen == 6 || hotlen == 7 || hotlen == 15; }

int try_encode(mgb_handle* h, DevCsr& D);
int try_patterns(mgb_handle* h, DevCsr& D, const std::vector<int64_t>& ip, bool anchored);
void free_coded(Coded& c);

// Kernel family, row tiles and stream descriptors of an operator whose arrays are already on the device.
// Needs only the row pointers on the host (ip).  interior (optional): row range [b0, b1) that references no ghost
// column -> the stream tiles are additionally split into boundary-low | interior | boundary-high.
int finish_csr(mgb_handle* h, DevCsr& D, const std::vector<int64_t>& ip, const std::vector<int32_t>& breaks = {},
               const int64_t* interior = nullptr)
{
    const int64_t n = D.nrows, nnz = D.nnz;
    int mx = 0;
    for (int64_t i = 0; i < n; ++i) mx = std::max<int>(mx, (int)(ip[i + 1] - ip[i]));
    D.max_row = mx;
    int family = h->opt_family;
    if (family == 0) family = 1;
    int iter = h->opt_iter ? h->opt_iter : 1;
    if (family == 1) {
        std::vector<int32_t> tiles;
        bool ok = make_tiles(ip, tile_cap(iter), 4 * THREADS, breaks, tiles, &D.break_tile);
        if (!ok && !h->opt_iter) { iter = 2; ok = make_tiles(ip, tile_cap(iter), 4 * THREADS, breaks, tiles, &D.break_tile); }
        if (ok) {
            D.ntiles = (int)tiles.size() - 1;
            TRY(dev_upload(h, &D.tiles, tiles.data(), tiles.size()));
        } else {
            family = 3;                                  // a row longer than a tile: warp per row, still summed in stored order
        }
    }
    D.family = family; D.iter = iter;
    if (family == 1 && h->stream_cfg > 0 && breaks.empty()) {     // descriptors for the TMA stream kernels
        // one configuration for all operators by default; "stream_auto" switches long-row operators (> 16 entries per
        // row on average, i.e. P2) to 2048-entry tiles -- measured slower, hence off
        D.scfg = h->stream_cfg;
        if (h->stream_auto && n > 0 && (double)nnz / (double)n > 16.0) D.scfg = 1;
        std::vector<int32_t> st, sbreaks, sbt;
        if (interior) {                          // interior row range [b0, b1), shrunk to multiples of 16 rows
            const int64_t b0 = (interior[0] + 15) & ~(int64_t)15, b1 = interior[1] == n ? n : (interior[1] & ~(int64_t)15);
            if (b1 > b0) sbreaks = {(int32_t)b0, (int32_t)b1};
        }
        // dictionary-coded copy (mgb_code.cuh) when the operator's entries are repetitive: row tiles for k_rowstream,
        // entries counted from a 16-entry aligned start (16-byte bulk copies of one-byte codes)
        bool tiled = false;
        int64_t align_mask = 7;
        TRY(try_patterns(h, D, ip, false));                  // whole rows repeat (uniform mesh, banded numbering): one byte per ROW
        if (!D.cd.mode) TRY(try_patterns(h, D, ip, true));   // ... or repeat when measured from their first column (P): 4 + 1 bytes per row
        if (!D.cd.mode) TRY(try_encode(h, D));               // else one byte per stored entry
        if (D.cd.mode) {
            D.ccfg = h->code_cfg;
            const CodeChoice cc = code_choice(D.ccfg);
            int64_t rowcap = (int64_t)cc.threads * cc.rpt;
            D.wcfg = 0; D.hcfg = 0;
            if (D.cd.mode == 3 && h->stage_x == 3 && D.cd.hot_ok && hot_len_supported(D.cd.hotplan.hotlen)) {   // hot-row kernel: its own tile size
                D.hcfg = std::max(1, h->hot_cfg);
                rowcap = (int64_t)hot_choice(D.hcfg).threads * hot_choice(D.hcfg).rpt;
            } else if (D.cd.mode == 4) {                                                                 // anchored patterns: their own tile size
                D.hcfg = 1;
                rowcap = (h->anch_cfg == 1 || h->anch_cfg == 0) ? 512 : 256;
            } else if (D.cd.mode == 3 && h->stage_x == 1 && D.cd.win.ng > 0 && D.cd.dict_win) {          // row-window kernel: its own tile size
                D.wcfg = std::max(1, h->win_cfg);
                rowcap = 256 * (int64_t)win_choice(D.wcfg).rpt;
            }
            if (D.cd.mode >= 3) tiled = make_tiles(ip, (int64_t)1 << 40, rowcap, sbreaks, st, &sbt, 16);    // rows only; 16-byte aligned code slices
            else tiled = make_tiles(ip, rowcap * cc.epr - 8, rowcap, sbreaks, st, &sbt, 4);
            if (tiled) align_mask = 15; else free_coded(D.cd);
        }
        if (!tiled) {
            const StreamChoice sc = stream_choice(D.scfg);
            const int cap = sc.threads * sc.ept;
            tiled = make_tiles(ip, cap, cap / 4, sbreaks, st, &sbt, 4);
        }
        if (tiled) {
            std::vector<int4> desc(st.size() - 1);
            for (size_t t = 0; t + 1 < st.size(); ++t) {
                const int64_t r0 = st[t], r1 = st[t + 1];
                const int64_t z0 = ip[r0] & ~align_mask, z1 = ip[r1];
                if (D.cd.mode >= 3) desc[t] = make_int4((int)r0, (int)(r1 - r0), 0, (int)(ip[r1] - ip[r0]));   // (entry count: accounting only)
                else desc[t] = make_int4((int)r0, (int)(r1 - r0), (int)z0, (int)((z1 - z0 + align_mask) & ~align_mask));
            }
            D.sntiles = (int)desc.size();
            if (D.sntiles > 0) TRY(dev_upload(h, &D.sdesc, desc.data(), desc.size()));
            if (interior) {
                D.split = true;
                D.int_r0 = interior[0]; D.int_r1 = interior[1];
                D.t_int0 = sbreaks.empty() ? 0 : sbt[0];
                D.t_int1 = sbreaks.empty() ? 0 : sbt[1];
                std::vector<int4> bnd(desc.begin(), desc.begin() + D.t_int0);
                bnd.insert(bnd.end(), desc.begin() + D.t_int1, desc.end());
                D.n_bnd = (int)bnd.size();
                if (D.n_bnd > 0) TRY(dev_upload(h, &D.sdesc_bnd, bnd.data(), bnd.size()));
            }
            D.sdesc_host.swap(desc);
        }
    }
    if (family == 3) D.break_tile.assign(breaks.begin(), breaks.end());
    if (family == 2) {
        int lpr = h->opt_lpr;
        if (!lpr) {
            const double avg = n ? (double)nnz / (double)n : 0.0;
            lpr = 1;
            while (lpr < 32 && lpr < avg) lpr *= 2;
            if (mx > tile_cap(2)) lpr = 32;
        }
        D.lpr = lpr;
        // breakpoints for the sub-warp family are plain row indices
        D.break_tile.assign(breaks.begin(), breaks.end());
    }
    return MGB_OK;
}

__global__ void k_not_ascending(int n, const int32_t* __restrict__ a, int* __restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < n && a[i + 1] <= a[i]) *bad = 1;
}

__global__ void k_row_refs_ghost(int n, const int32_t* __restrict__ rp, const int32_t* __restrict__ cols, int n_owned_cols,
                                 unsigned char* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char f = 0;
    for (int k = rp[i]; k < rp[i + 1]; ++k) f |= (cols[k] >= n_owned_cols);
    flag[i] = f;
}

// Longest run of rows that reference no ghost column (columns >= n_owned_cols).  out[0..1] = [b0, b1).
int interior_rows(mgb_handle* h, const DevCsr& D, int64_t n_owned_cols, int64_t* out)
{
    const int n = (int)D.nrows;
    out[0] = out[1] = 0;
    if (n == 0) return MGB_OK;
    unsigned char* d = nullptr;
    TRY(dev_alloc(h, &d, (size_t)n));
    k_row_refs_ghost<<<(n + 255) / 256, 256, 0, h->stream>>>(n, D.rowptr, D.cols, (int)n_owned_cols, d);
    std::vector<unsigned char> f((size_t)n);
    CU(cudaMemcpyAsync(f.data(), d, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d);
    int64_t best0 = 0, best1 = 0, run0 = 0;
    for (int64_t i = 0; i <= n; ++i) {
        if (i == n || f[(size_t)i]) {
            if (i - run0 > best1 - best0) { best0 = run0; best1 = i; }
            run0 = i + 1;
        }
    }
    out[0] = best0; out[1] = best1;
    return MGB_OK;
}

// Upload a host CSR and choose the kernel family / tile shape for it.  x_owned >= 0: the operator reads a sharded
// vector with x_owned owned entries followed by ghosts -> also split its tiles into interior / boundary.
int upload_csr(mgb_handle* h, const HostCsr& M, DevCsr& D, const std::vector<int32_t>& breaks = {}, int64_t x_owned = -1)
{
    const int64_t n = M.nrows, nnz = M.nnz();
    if (n >= (int64_t)2147483000 || nnz >= (int64_t)2147483000)
        return fail(h, MGB_ERR_UNSUPPORTED, "operator with %lld rows / %lld entries exceeds the int32 row-pointer range of one device shard",
                    (long long)n, (long long)nnz);
    D.nrows = n; D.ncols = M.ncols; D.nnz = nnz;
    std::vector<int32_t> rp((size_t)n + 1);
    for (int64_t i = 0; i <= n; ++i) rp[i] = (int32_t)M.ip[i];
    TRY(dev_upload(h, &D.rowptr, rp.data(), rp.size(), 8));
    TRY(dev_upload(h, &D.cols, M.ix.data(), (size_t)nnz, 16));   // padding: the last 8-wide group may read past nnz
    TRY(dev_upload(h, &D.vals, M.ax.data(), (size_t)nnz, 16));
    int64_t in[2];
    if (x_owned >= 0) TRY(interior_rows(h, D, x_owned, in));
    return finish_csr(h, D, M.ip, breaks, x_owned >= 0 ? in : nullptr);
}

// Row pointers of a device-resident operator back on the host (int64), for finish_csr.
int fetch_rowptr(mgb_handle* h, const DevCsr& D, std::vector<int64_t>& ip)
{
    std::vector<int32_t> rp((size_t)D.nrows + 1);
    CU(cudaMemcpyAsync(rp.data(), D.rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    ip.assign(rp.begin(), rp.end());
    return MGB_OK;
}

#include "mgb_synth.cuh"

// ---- launch bookkeeping -------------------------------------------------------------------------------
// bytes: algorithmic bytes of the launch (CSR form); moved: what the chosen kernel actually streams
// (smaller for dictionary-coded operators); < 0: same as bytes
template <class F>
int launch(mgb_handle* h, int kind, int level, double bytes, F&& fn, double moved = -1.0)
{
    ProfEvent pe{kind, level, bytes, moved < 0.0 ? bytes : moved, nullptr, nullptr};
    if (h->prof) {
        CU(cudaEventCreate(&pe.e0)); CU(cudaEventCreate(&pe.e1));
        CU(cudaEventRecord(pe.e0, h->stream));
    }
    fn();
    CU(cudaGetLastError());
    if (h->prof) {
        CU(cudaEventRecord(pe.e1, h->stream));
        h->prof_events.push_back(pe);
    }
    h->launches++;
    return MGB_OK;
}

template <class Epi, bool NCX>
void launch_tile(mgb_handle* h, const DevCsr& D, int t0, int t1, const double* x, const Epi& epi)
{
    const int nt = t1 - t0;
    if (nt <= 0) return;
    switch (D.iter) {
        case 1: k_tile<1, THREADS, NCX, Epi><<<nt, THREADS, 0, h->stream>>>(D.rowptr, D.cols, D.vals, D.tiles, t0, x, epi); break;
        default: k_tile<2, THREADS, NCX, Epi><<<nt, THREADS, 0, h->stream>>>(D.rowptr, D.cols, D.vals, D.tiles, t0, x, epi); break;
    }
}

template <int T, int E, int S, class Epi>
void launch_stream_cfg(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi, bool chunked)
{
    auto kern = k_stream<T, E, S, true, Epi>;
    constexpr int smem = StreamCfg<T, E, Epi::NOPS, EpiNI<Epi>::value>::smem_bytes(S);
    int occ = 1;
    {   // per instantiation and device; handles are independent, so first use is guarded
        static std::mutex mu;
        static std::map<int, int> occ_of;
        std::lock_guard<std::mutex> lock(mu);
        auto it = occ_of.find(h->device);
        if (it == occ_of.end()) {
            int o = 0;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, T + 32, smem);
            it = occ_of.emplace(h->device, std::max(o, 1)).first;
        }
        occ = it->second;
    }
    int grid = std::min(ntiles, h->sm_count * occ), tpc = 0;
    if (chunked) {                       // CTAs that retire after tpc tiles: `overlap_waves` waves of them (see k_stream)
        tpc = std::max(1, ntiles / (h->sm_count * occ * std::max(1, h->overlap_waves)));
        grid = (ntiles + tpc - 1) / tpc;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T + 32); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see k_stream: overlap this kernel's prologue and first
    at[0].val.programmaticStreamSerializationAllowed = 1;                // matrix tiles with the tail of the previous kernel
    cfg.attrs = at; cfg.numAttrs = h->pdl == 1 ? 1 : 0;        // CSR stream kernel: only when forced (measured slower)
    cudaLaunchKernelEx(&cfg, kern, (const int32_t*)D.rowptr, (const int32_t*)D.cols, (const double*)D.vals, desc, ntiles, tpc, x, epi);
}

template <int T, int RPT, int EPR, int S, int MODE, int JW, int MINB, class Epi>
void launch_rowstream_cfg(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi, bool chunked)
{
    auto kern = k_rowstream<T, RPT, EPR, S, MODE, JW, MINB, Epi>;
    const int npent = MODE == 3 ? D.cd.npent : 0;            // the pattern table's size decides the shared-memory footprint
    const int smem = RowCfg<T, RPT, EPR, Epi::NOPS, EpiNI<Epi>::value, MODE>::smem_bytes(S, npent * (int)sizeof(DictEnt));
    int occ = 1;
    {   // per-instantiation cache keyed by (device, bytes): the attribute is per device, and handles are independent, so first use is guarded
        static std::mutex mu;
        static std::map<std::pair<int, int>, int> occ_by_smem;
        static std::map<int, int> smem_attr;
        std::lock_guard<std::mutex> lock(mu);
        if (smem > smem_attr[h->device]) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); smem_attr[h->device] = smem; }
        auto it = occ_by_smem.find({h->device, smem});
        if (it == occ_by_smem.end()) {
            int o = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, T + 32, smem);
            it = occ_by_smem.emplace(std::make_pair(h->device, smem), std::max(o, 1)).first;
        }
        occ = it->second;
    }
    int grid = std::min(ntiles, h->sm_count * occ), tpc = 0;
    if (chunked) {
        tpc = std::max(1, ntiles / (h->sm_count * occ * std::max(1, h->overlap_waves)));
        grid = (ntiles + tpc - 1) / tpc;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T + 32); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl != 0 ? 1 : 0;        // on by default for the coded kernels (measured: cfg2 0.280 -> 0.264 ms)
    cudaLaunchKernelEx(&cfg, kern, (const int32_t*)D.rowptr, (const int32_t*)D.cols, (const unsigned char*)D.cd.codes,
                       (const DictEnt*)D.cd.dict, (const int2*)D.cd.phead, npent, desc, ntiles, tpc, x, epi);
}

// Row patterns with x staged in shared memory (k_rowwin).  The window plan of the operator is completed here for the
// tile size of the chosen configuration (where each window's slice sits inside a stage, the hot pattern's slots).
template <int RPT, int S, int MINB, int HOTN, class Epi>
void launch_rowwin_cfg(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi, bool chunked)
{
    constexpr int T = 256;
    using Cfg = WinCfg<T, RPT, Epi::NOPS, EpiNI<Epi>::value>;
    auto kern = k_rowwin<T, RPT, S, HOTN, MINB, Epi>;
    WinPlan W = D.cd.win;
    int xd = 0;
    for (int g = 0; g < W.ng; ++g) { W.goff[g] = xd; xd += (Cfg::ROWCAP + W.gspan[g] + 1) & ~1; }
    W.xdoubles = xd;
    const int gv = W.vslot;                                  // window holding offset 0 (try_patterns) -> slot
    W.vslot = W.goff[gv] - W.gmin[gv];
    for (int e = 0; e < WIN_HOT; ++e) W.hs[e] = W.goff[W.hs[e] >> WIN_GSHIFT] + (W.hs[e] & ((1 << WIN_GSHIFT) - 1));
    constexpr int XOP = EpiXop<Epi>::value;
    bool alias = false;
    if constexpr (XOP >= 0) alias = epi.operand(XOP) == x;
    const int smem = Cfg::smem_bytes(S, D.cd.npent, Epi::NOPS - (alias ? 1 : 0), xd);
    int occ = 1;
    {
        static std::mutex mu;
        static std::map<std::pair<int, int>, int> occ_by_smem;      // (device, bytes)
        static std::map<int, int> smem_attr;                        // device -> largest size configured
        std::lock_guard<std::mutex> lock(mu);
        if (smem > smem_attr[h->device]) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); smem_attr[h->device] = smem; }
        auto it = occ_by_smem.find({h->device, smem});
        if (it == occ_by_smem.end()) {
            int o = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, T + 32, smem);
            it = occ_by_smem.emplace(std::make_pair(h->device, smem), std::max(o, 1)).first;
        }
        occ = it->second;
    }
    int grid = std::min(ntiles, h->sm_count * occ), tpc = 0;
    if (chunked) {
        tpc = std::max(1, ntiles / (h->sm_count * occ * std::max(1, h->overlap_waves)));
        grid = (ntiles + tpc - 1) / tpc;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T + 32); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl != 0 ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, (const unsigned char*)D.cd.codes, (const DictEnt*)D.cd.dict_win, (const int2*)D.cd.phead, D.cd.npent, W,
                       desc, ntiles, tpc, h->win_prefetch, x, epi);
}

template <int HOTN, class Epi>
void launch_rowwin_hot(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi, bool chunked)
{
    switch (D.wcfg) {                    // win_choice()
        case 2: launch_rowwin_cfg<4, 2, 2, HOTN, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
        case 3: launch_rowwin_cfg<2, 3, 3, HOTN, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
        case 4: launch_rowwin_cfg<4, 3, 2, HOTN, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
        default: launch_rowwin_cfg<2, 2, 4, HOTN, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
    }
}

// The straight-line body exists for the pattern lengths of the operators this engine is built for: the smoother matrix of
// the P1 Laplacian (4 couplings per row in 2-D, 6 in 3-D) and the level matrix as exported (7 / 15 stored entries).
template <class Epi>
void launch_rowwin(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi, bool chunked)
{
    const int hl = D.cd.win.hotlen;
    if constexpr (std::is_same<Epi, EpiJacobiRJ>::value || std::is_same<Epi, EpiJacobiRJFirst>::value) {
        if (hl == 4) return launch_rowwin_hot<4, Epi>(h, D, desc, ntiles, x, epi, chunked);
        if (hl == 6) return launch_rowwin_hot<6, Epi>(h, D, desc, ntiles, x, epi, chunked);
    } else if constexpr (!std::is_same<Epi, EpiProlongAdd>::value) {
        if (hl == 7) return launch_rowwin_hot<7, Epi>(h, D, desc, ntiles, x, epi, chunked);
        if (hl == 15) return launch_rowwin_hot<15, Epi>(h, D, desc, ntiles, x, epi, chunked);
    }
    launch_rowwin_hot<0, Epi>(h, D, desc, ntiles, x, epi, chunked);
}

// ---- fused halo exchange (HaloFuse, mgb_kernels.cuh): plan of one launch ---------------------------------------------
// which epilogue / pattern-length combinations the hot-row kernel is instantiated for (launch_hotrow)
template <class Epi>
bool hot_len_ok(int hl)
{
    if constexpr (std::is_same<Epi, EpiJacobiRJ>::value || std::is_same<Epi, EpiJacobiRJFirst>::value) return hl == 4 || hl == 6;
    else if constexpr (std::is_same<Epi, EpiProlongAdd>::value) return false;
    else return hl == 7 || hl == 15;
}
// true when the kernel launch_stream will pick for (D, Epi) takes a HaloFuse (hot-row and anchored-pattern kernels)
template <class Epi>
bool kernel_takes_hf(const DevCsr& D)
{
    if constexpr (!Epi::CONTIG) return false;
    return D.cd.mode == 4 || (D.cd.mode == 3 && D.hcfg > 0 && hot_len_ok<Epi>(D.cd.hotplan.hotlen));
}
// XL: level whose vector the kernel reads through operator D (its ghost entries must have arrived); OL / out: level and buffer
// the kernel writes an iterate to (its boundary rows go to the neighbours).  Either may be null.  Outside a cycle, or on a level
// whose exchange is not fused, the corresponding half stays empty (and the caller uses the push / pull kernels).
HaloFuse no_hf()
{
    HaloFuse hf{};
    hf.int_b0 = 0; hf.int_b1 = INT_MAX;
    return hf;
}
HaloFuse make_hf(mgb_handle* h, const Level* XL, const DevCsr* D, const Level* OL, const double* out)
{
    HaloFuse hf{};
    hf.int_b0 = 0; hf.int_b1 = INT_MAX;
    if (!h->in_cycle) return hf;
    if (XL && XL->fuse_ok) {
        const P2PPlan& pl = XL->p2p;
        hf.wait_n = pl.npeers;
        for (int p = 0; p < pl.npeers && p < 2; ++p) hf.wflag[p] = pl.flags2 + pl.peer_rank[p];
        hf.wepoch = XL->fuse_counters;
        if (D && D->split) { hf.int_b0 = (int)D->int_r0; hf.int_b1 = (int)D->int_r1; }
        else { hf.int_b0 = 0; hf.int_b1 = 0; }              // (no row classification: every CTA waits)
    }
    if (OL && OL->fuse_ok && (out == OL->v || out == OL->vtmp)) {
        const P2PPlan& pl = OL->p2p;
        const int which = out == OL->v ? 0 : 1;
        hf.send_n = pl.npeers;
        for (int p = 0; p < pl.npeers && p < 2; ++p) {
            hf.send_a[p] = OL->send_a[p]; hf.send_cnt[p] = OL->send_cnt[(size_t)p];
            hf.dst[p] = pl.rvec[p][which]; hf.rflag[p] = pl.rflag2[p];
        }
        hf.sepoch = OL->fuse_counters;
    }
    return hf;
}
// per-launch part: CTAs per send range and the boundary-first tile order, for tiles of `rows` rows over n rows
void finish_hf(HaloFuse& hf, int rows, int64_t n)
{
    const int64_t ntiles = (n + rows - 1) / rows;
    int64_t lo_end = 0, hi_start = n;
    if (hf.wait_n) { lo_end = std::max<int64_t>(lo_end, std::min<int64_t>(hf.int_b0, n)); hi_start = std::min<int64_t>(hi_start, std::max<int64_t>(hf.int_b1, 0)); }
    for (int p = 0; p < hf.send_n; ++p) {
        const int64_t a = hf.send_a[p], b = a + hf.send_cnt[p];
        hf.send_ctas[p] = (int)((b - 1) / rows - a / rows + 1);
        if (a < n / 2) lo_end = std::max(lo_end, b); else hi_start = std::min(hi_start, a);
    }
    hf.nlo = hf.nhi = 0;
    if (hf.wait_n || hf.send_n) {
        const int64_t nlo = (lo_end + rows - 1) / rows, nhi = hi_start >= n ? 0 : ntiles - hi_start / rows;
        if (nlo + nhi <= ntiles) { hf.nlo = (int)nlo; hf.nhi = (int)nhi; }
    }
}

// Hot-row kernel (k_hotrow).  Registers per thread are budgeted from what a thread keeps in flight (RPT rows x (HOTN x values +
// operands)), and the resident CTAs per SM follow from that budget.
template <int HOTN, int T, int RPT, int NOPS>
constexpr int hot_minb()
{
    int regs = 28 + 2 * RPT * (HOTN + NOPS + 1);
    regs = (regs + 7) / 8 * 8;
    int b = 65536 / (T * regs);
    const int cap = 2048 / T > 16 ? 16 : 2048 / T;
    return b < 1 ? 1 : (b > cap ? cap : b);
}

template <int HOTN, int T, int RPT, class Epi>
void launch_hotrow_cfg(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi)
{
    constexpr int MINB = hot_minb<HOTN, T, RPT, Epi::NOPS>();
    constexpr int ROWS = T * RPT;
    const bool linear = desc == D.sdesc && ntiles == D.sntiles;              // all rows: tiles follow from blockIdx, no descriptor is read
    const int grid = linear ? (int)((D.nrows + ROWS - 1) / ROWS) : ntiles;
    if (grid <= 0) return;
    // L2 prefetch distance in tiles (linear tiling only); none while borrowed user pointers are in play (bulk prefetches need
    // 16-byte aligned operands).  pf_last: last tile whose slices -- whole tile of codes / operands, x shifted by the largest hot
    // offset -- lie inside their arrays.
    int pf = linear && h->allow_stream && h->hot_pf > 0 && D.cd.hot.dmax >= 0 ? std::max(1, h->hot_pf / ROWS) : 0;
    const int64_t pf_last = std::min<int64_t>(D.nrows / ROWS, (D.ncols - (D.cd.hot.dmax & ~1)) / ROWS) - 1;
    if (pf_last < pf) pf = 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl != 0 ? 1 : 0;
    HaloFuse hf = linear ? h->hf_cur : no_hf();              // (tile subsets never run inside a fused exchange)
    finish_hf(hf, ROWS, D.nrows);
    const bool halo = hf.wait_n > 0 || hf.send_n > 0;       // single GPU: the instantiation without any exchange code
    cudaLaunchKernelEx(&cfg, halo ? k_hotrow<HOTN, T, RPT, MINB, true, Epi> : k_hotrow<HOTN, T, RPT, MINB, false, Epi>,
                       (const unsigned char*)D.cd.codes, (const uint32_t*)D.cd.pmask, (const int2*)D.cd.phead, (const DictEnt*)D.cd.dict,
                       D.cd.hot, hf, linear ? (const int4*)nullptr : desc, grid, 0, (int)D.nrows, (int)D.ncols, pf, (int)pf_last, x, epi);
}

template <int HOTN, class Epi>
void launch_hotrow_len(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi)
{
    switch (D.hcfg) {                    // hot_choice()
        case 2: launch_hotrow_cfg<HOTN, 256, 1, Epi>(h, D, desc, ntiles, x, epi); break;
        case 3: launch_hotrow_cfg<HOTN, 128, 1, Epi>(h, D, desc, ntiles, x, epi); break;
        case 4: launch_hotrow_cfg<HOTN, 256, 2, Epi>(h, D, desc, ntiles, x, epi); break;
        default: launch_hotrow_cfg<HOTN, 128, 2, Epi>(h, D, desc, ntiles, x, epi); break;
    }
}

// The hot-row body exists for the pattern lengths of the operators this engine is built for: the smoother matrix of the P1
// Laplacian (4 couplings per row in 2-D, 6 in 3-D) and the level matrix as exported (7 / 15 stored entries).  false: not
// instantiated for this epilogue / length -- the caller falls back to the row-stream kernel (same tiles).
template <class Epi>
bool launch_hotrow(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi)
{
    const int hl = D.cd.hotplan.hotlen;
    if constexpr (std::is_same<Epi, EpiJacobiRJ>::value || std::is_same<Epi, EpiJacobiRJFirst>::value) {
        if (hl == 4) { launch_hotrow_len<4, Epi>(h, D, desc, ntiles, x, epi); return true; }
        if (hl == 6) { launch_hotrow_len<6, Epi>(h, D, desc, ntiles, x, epi); return true; }
    } else if constexpr (!std::is_same<Epi, EpiProlongAdd>::value) {
        if (hl == 7) { launch_hotrow_len<7, Epi>(h, D, desc, ntiles, x, epi); return true; }
        if (hl == 15) { launch_hotrow_len<15, Epi>(h, D, desc, ntiles, x, epi); return true; }
    }
    return false;
}

// Two Jacobi sweeps in one launch (k_hotrow2): x -> y -> out on an unsharded hot-row operator, linear tiling.  reach: tiles a row's
// hot offsets can lie away from its own tile.  false: shape not covered (the caller runs two launches).
struct Sweep2Plan { int rows, ntiles, reach, groups, ngrid_groups_max; };
bool sweep2_plan(const mgb_handle* h, const DevCsr& D, Sweep2Plan& p)
{
    if (D.cd.mode != 3 || D.hcfg <= 0 || !D.sdesc) return false;
    const int hl = D.cd.hotplan.hotlen;
    if (hl != 4 && hl != 6) return false;
    const HotChoice hc = hot_choice(D.hcfg);
    p.rows = hc.threads * hc.rpt;
    p.ntiles = (int)((D.nrows + p.rows - 1) / p.rows);
    const int64_t far = std::max<int64_t>(std::abs((int64_t)D.cd.hot.dmin), std::abs((int64_t)D.cd.hot.dmax));
    p.reach = (int)((far + p.rows - 1) / p.rows) + 1;
    p.groups = (p.ntiles + S2_GROUP - 1) / S2_GROUP;
    return 2 * p.reach / S2_GROUP + 2 <= hc.threads && D.nrows == D.ncols;
}

template <int HOTN, int T, int RPT, class Epi1>
void launch_hotrow2_cfg(mgb_handle* h, const DevCsr& D, const Sweep2Plan& p, unsigned long long* s2, const double* x, const Epi1& epi1,
                        const double* y, const EpiJacobiRJ& epi2)
{
    constexpr int MINB = hot_minb<HOTN, T, RPT, Epi1::NOPS>();
    constexpr int ROWS = T * RPT;
    int pf = h->allow_stream && h->hot_pf > 0 && D.cd.hot.dmax >= 0 ? std::max(1, h->hot_pf / ROWS) : 0;
    const int64_t pf_last = std::min<int64_t>(D.nrows / ROWS, (D.ncols - (D.cd.hot.dmax & ~1)) / ROWS) - 1;
    if (pf_last < pf) pf = 0;
    int kt = 1;
    while (kt < S2_GROUP && kt < h->s2_tiles) kt *= 2;                          // tiles per CTA: a power of two <= S2_GROUP
    const int nchunks = (p.ntiles + kt - 1) / kt;
    const int lagc = (p.reach + S2_GROUP + h->s2_slack + kt - 1) / kt + 1;      // chunks the second sweep trails the first by
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (nchunks + lagc)); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl != 0 ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_hotrow2<HOTN, T, RPT, MINB, false, true, Epi1, EpiJacobiRJ>,
                       (const unsigned char*)D.cd.codes, (const uint32_t*)D.cd.pmask, (const int2*)D.cd.phead, (const DictEnt*)D.cd.dict,
                       D.cd.hot, p.ntiles, (int)D.nrows, (int)D.ncols, pf, (int)pf_last, x, epi1, y, epi2, s2, kt, lagc, p.reach);
    k_s2_reset<<<(p.groups + 255) / 256, 256, 0, h->stream>>>(p.groups, s2);    // (stream-ordered after the whole pair)
}

template <class Epi1>
void launch_hotrow2(mgb_handle* h, const DevCsr& D, const Sweep2Plan& p, unsigned long long* s2, const double* x, const Epi1& epi1,
                    const double* y, const EpiJacobiRJ& epi2)
{
    const int hl = D.cd.hotplan.hotlen;
    switch (D.hcfg) {                    // hot_choice()
        case 2: if (hl == 4) launch_hotrow2_cfg<4, 256, 1>(h, D, p, s2, x, epi1, y, epi2); else launch_hotrow2_cfg<6, 256, 1>(h, D, p, s2, x, epi1, y, epi2); break;
        case 3: if (hl == 4) launch_hotrow2_cfg<4, 128, 1>(h, D, p, s2, x, epi1, y, epi2); else launch_hotrow2_cfg<6, 128, 1>(h, D, p, s2, x, epi1, y, epi2); break;
        case 4: if (hl == 4) launch_hotrow2_cfg<4, 256, 2>(h, D, p, s2, x, epi1, y, epi2); else launch_hotrow2_cfg<6, 256, 2>(h, D, p, s2, x, epi1, y, epi2); break;
        default: if (hl == 4) launch_hotrow2_cfg<4, 128, 2>(h, D, p, s2, x, epi1, y, epi2); else launch_hotrow2_cfg<6, 128, 2>(h, D, p, s2, x, epi1, y, epi2); break;
    }
}

// Anchored row patterns (k_anchrow).  Configurations (option "anch_cfg"): threads x rows per thread.
template <int T, int RPT, int JW, int MINB, class Epi>
void launch_anchrow_cfg(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi)
{
    constexpr int ROWS = T * RPT;
    auto kern_halo = k_anchrow<T, RPT, JW, MINB, true, Epi>;
    auto kern_plain = k_anchrow<T, RPT, JW, MINB, false, Epi>;       // single GPU: the instantiation without any exchange code
    const bool linear = desc == D.sdesc && ntiles == D.sntiles;
    const int grid = linear ? (int)((D.nrows + ROWS - 1) / ROWS) : ntiles;
    if (grid <= 0) return;
    const int pf = h->allow_stream && h->hot_pf > 0 ? std::max(1, h->hot_pf / ROWS) : 0;
    const int smem = D.cd.npent * 12 + 256 * 8;
    if (smem > 48 * 1024) {
        static std::mutex mu;
        static std::map<int, int> attr;
        std::lock_guard<std::mutex> lock(mu);
        if (smem > attr[h->device]) {
            cudaFuncSetAttribute(kern_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            attr[h->device] = smem;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl != 0 ? 1 : 0;
    HaloFuse hf = linear ? h->hf_cur : no_hf();
    finish_hf(hf, ROWS, D.nrows);
    const bool halo = hf.wait_n > 0 || hf.send_n > 0;
    const int kt = std::abs(h->anch_tiles);                  // (negative: also on grids too small to fill the GPU that way -- tests)
    if (!halo && linear && kt > 1 && (h->anch_tiles < 0 || grid >= 4 * kt * h->sm_count)) {    // several tiles per CTA, next anchors in flight
        auto kern_loop = k_anchloop<T, RPT, JW, (MINB > 1 ? MINB - 1 : 1), Epi>;       // (the next tile's anchors cost registers: one CTA fewer per SM instead of spills)
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cfg.gridDim = dim3((grid + kt - 1) / kt);
        cudaLaunchKernelEx(&cfg, kern_loop, (const unsigned char*)D.cd.codes, (const int32_t*)D.cd.anchor, (const int2*)D.cd.phead,
                           (const DictEnt*)D.cd.dict, D.cd.ndict, D.cd.npent, grid, (int)D.nrows, kt, pf, x, epi);
        return;
    }
    cudaLaunchKernelEx(&cfg, halo ? kern_halo : kern_plain, (const unsigned char*)D.cd.codes, (const int32_t*)D.cd.anchor, (const int2*)D.cd.phead, (const DictEnt*)D.cd.dict,
                       D.cd.ndict, D.cd.npent, hf, linear ? (const int4*)nullptr : desc, grid, 0, (int)D.nrows, pf, x, epi);
}

template <class Epi>
void launch_anchrow(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi)
{
    if (D.max_row <= 4) {
        switch (h->anch_cfg) {
            case 2: return launch_anchrow_cfg<128, 2, 4, 8, Epi>(h, D, desc, ntiles, x, epi);
            case 3: return launch_anchrow_cfg<64, 4, 4, 12, Epi>(h, D, desc, ntiles, x, epi);
            default: return launch_anchrow_cfg<128, 4, 4, 6, Epi>(h, D, desc, ntiles, x, epi);
        }
    }
    switch (h->anch_cfg) {
        case 2: return launch_anchrow_cfg<128, 2, 8, 5, Epi>(h, D, desc, ntiles, x, epi);
        case 3: return launch_anchrow_cfg<64, 4, 8, 10, Epi>(h, D, desc, ntiles, x, epi);
        case 4: return launch_anchrow_cfg<256, 1, 8, 5, Epi>(h, D, desc, ntiles, x, epi);
        default: return launch_anchrow_cfg<128, 4, 8, 5, Epi>(h, D, desc, ntiles, x, epi);
    }
}

// MODE: the coding (pair / value codes); JW: gathers issued up front per row (4 when no row is longer, else 8)
template <int MODE, int JW, class Epi>
void launch_rowstream(mgb_handle* h, const DevCsr& D, const int4* desc, int ntiles, const double* x, const Epi& epi, bool chunked)
{
    switch (D.ccfg) {                    // code_choice()
        case 3: launch_rowstream_cfg<256, 2, 8, 3, MODE, JW, 4, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
        default: launch_rowstream_cfg<256, 2, 8, 2, MODE, JW, 4, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
    }
}

template <class Epi>
void launch_stream(mgb_handle* h, const DevCsr& D, const double* x, const Epi& epi, const int4* desc = nullptr, int ntiles = 0,
                   bool chunked = false)
{
    if (!desc) { desc = D.sdesc; ntiles = D.sntiles; }
    if (ntiles <= 0) return;
    if constexpr (Epi::CONTIG) {
        if (D.cd.mode == 4) return launch_anchrow<Epi>(h, D, desc, ntiles, x, epi);
        if (D.cd.mode == 3 && D.hcfg > 0 && launch_hotrow<Epi>(h, D, desc, ntiles, x, epi)) return;
        if (D.cd.mode == 3 && D.wcfg > 0) {              // x staged in shared memory (tiles were cut for this configuration)
            return launch_rowwin<Epi>(h, D, desc, ntiles, x, epi, chunked);
        }
        if (D.cd.mode == 3 && D.max_row <= 4) return launch_rowstream<3, 4, Epi>(h, D, desc, ntiles, x, epi, chunked);
        if (D.cd.mode == 3) return launch_rowstream<3, 8, Epi>(h, D, desc, ntiles, x, epi, chunked);
        if (D.cd.mode == 1 && D.max_row <= 4) return launch_rowstream<1, 4, Epi>(h, D, desc, ntiles, x, epi, chunked);
        if (D.cd.mode == 1) return launch_rowstream<1, 8, Epi>(h, D, desc, ntiles, x, epi, chunked);
        if (D.cd.mode == 2 && D.max_row <= 4) return launch_rowstream<2, 4, Epi>(h, D, desc, ntiles, x, epi, chunked);
        if (D.cd.mode == 2) return launch_rowstream<2, 8, Epi>(h, D, desc, ntiles, x, epi, chunked);
        switch (D.scfg) {
            case 1: launch_stream_cfg<256, 8, 2, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
            case 2: launch_stream_cfg<512, 4, 2, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
            case 3: launch_stream_cfg<256, 4, 2, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
            case 4: launch_stream_cfg<256, 4, 3, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
            case 5: launch_stream_cfg<128, 8, 2, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
            case 7: launch_stream_cfg<128, 4, 2, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
            default: launch_stream_cfg<256, 8, 3, Epi>(h, D, desc, ntiles, x, epi, chunked); break;
        }
    }
}

template <class Epi, bool NCX>
void launch_subwarp(mgb_handle* h, const DevCsr& D, int r0, int r1, const double* x, const Epi& epi)
{
    const int64_t rows = r1 - r0;
    if (rows <= 0) return;
    const int64_t threads = rows * D.lpr;
    const int grid = (int)((threads + 255) / 256);
    switch (D.lpr) {
        case 1: k_subwarp<1, NCX, Epi><<<grid, 256, 0, h->stream>>>(D.rowptr, D.cols, D.vals, r0, r1, x, epi); break;
        case 2: k_subwarp<2, NCX, Epi><<<grid, 256, 0, h->stream>>>(D.rowptr, D.cols, D.vals, r0, r1, x, epi); break;
        case 4: k_subwarp<4, NCX, Epi><<<grid, 256, 0, h->stream>>>(D.rowptr, D.cols, D.vals, r0, r1, x, epi); break;
        case 8: k_subwarp<8, NCX, Epi><<<grid, 256, 0, h->stream>>>(D.rowptr, D.cols, D.vals, r0, r1, x, epi); break;
        case 16: k_subwarp<16, NCX, Epi><<<grid, 256, 0, h->stream>>>(D.rowptr, D.cols, D.vals, r0, r1, x, epi); break;
        default: k_subwarp<32, NCX, Epi><<<grid, 256, 0, h->stream>>>(D.rowptr, D.cols, D.vals, r0, r1, x, epi); break;
    }
}

// bytes a dictionary-coded operator does NOT move per pass, relative to its CSR form (12 per stored entry)
// (mode 3: one byte per row replaces the stored entries AND the row pointers)
double coded_saving(const DevCsr& D)
{
    switch (D.cd.mode) {
        case 1: return 11.0 * (double)D.nnz;
        case 2: return 7.0 * (double)D.nnz;
        case 3: return 12.0 * (double)D.nnz + 3.0 * (double)D.nrows;
        case 4: return 12.0 * (double)D.nnz - 1.0 * (double)D.nrows;      // anchor (4) + code (1) instead of the entries and the row pointer (4)
        default: return 0.0;
    }
}

// all rows of D (group < 0) or the rows of one breakpoint group (colour)
template <class Epi, bool NCX = true>
int row_sums(mgb_handle* h, int kind, int level, double bytes, const DevCsr& D, const double* x, const Epi& epi, int group = -1)
{
    if (D.nrows == 0) return MGB_OK;
    const bool streamed = Epi::CONTIG && NCX && D.family == 1 && group < 0 && D.sdesc && h->stream_cfg > 0 && h->allow_stream;
    return launch(h, kind, level, bytes, [&] {
        if (streamed) {
            launch_stream<Epi>(h, D, x, epi);
        } else if (D.family == 1) {
            int t0 = 0, t1 = D.ntiles;
            if (group >= 0) { t0 = D.break_tile[group]; t1 = D.break_tile[group + 1]; }
            launch_tile<Epi, NCX>(h, D, t0, t1, x, epi);
        } else {
            int r0 = 0, r1 = (int)D.nrows;
            if (group >= 0) { r0 = D.break_tile[group]; r1 = D.break_tile[group + 1]; }
            if (D.family == 3) {
                const int64_t rows = r1 - r0;
                if (rows > 0) k_seqrow<NCX, Epi><<<(int)((rows * 32 + 255) / 256), 256, 0, h->stream>>>(D.rowptr, D.cols, D.vals, r0, r1, x, epi);
            } else {
                launch_subwarp<Epi, NCX>(h, D, r0, r1, x, epi);
            }
        }
    }, streamed ? bytes - coded_saving(D) : -1.0);
}

// ---- algorithmic byte counts (SURVEY 8d / DESIGN.md) ---------------------------------------------------
double bytes_rowsum(const DevCsr& D, double vec_terms) { return 12.0 * (double)D.nnz + 4.0 * (double)D.nrows + 8.0 * vec_terms; }

Level* find_level(mgb_handle* h, int level)
{
    auto it = h->levels.find(level);
    return it == h->levels.end() ? nullptr : &it->second;
}

#include "mgb_dist.cuh"
#include "mgb_code.cuh"

// ---- smoothers ---------------------------------------------------------------------------------------
int gs_sweep(mgb_handle* h, Level& L, double* v, const double* f)
{
    const DevCsr& G = L.G;
    const double nb = 12.0 * (double)G.nnz + 8.0 * (double)L.n /*rowptr+order*/ + 8.0 * 3.0 * (double)L.n;
    if (h->smoother == MGB_SM_GS_LEVEL && h->gs_cluster >= 2 && L.gs_W > 0 && L.gs_groups + 1 <= 12000) {
        // rows of <= 8 entries: ELL copy of the operator, everything but the x gathers prefetched a level ahead
        return launch(h, MGB_K_GS, L.level, nb, [&] {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = (size_t)(L.gs_groups + 1) * sizeof(int); cfg.stream = h->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const int n = (int)L.n;
            const int32_t* ec = L.gs_ecols; const double* ev = L.gs_evals; const int32_t* ord = L.gs_order;
            const double* dg = L.gs_diag; const int32_t* off = L.gs_off; const int nlev = L.gs_groups;
            if (L.gs_W <= 4) cudaLaunchKernelEx(&cfg, k_gs_levels_ell<4>, n, ec, ev, ord, dg, f, v, off, nlev);
            else if (L.gs_W <= 6) cudaLaunchKernelEx(&cfg, k_gs_levels_ell<6>, n, ec, ev, ord, dg, f, v, off, nlev);
            else cudaLaunchKernelEx(&cfg, k_gs_levels_ell<8>, n, ec, ev, ord, dg, f, v, off, nlev);
        });
    }
    if (h->smoother == MGB_SM_GS_LEVEL && h->gs_cluster && L.gs_max_width <= 8 * 512 * 2) {
        // narrow levels: one 8-CTA cluster, hardware cluster barrier between dependency levels
        return launch(h, MGB_K_GS, L.level, nb, [&] {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_gs_levels_cluster, (const int32_t*)G.rowptr, (const int32_t*)G.cols, (const double*)G.vals,
                               (const int32_t*)L.gs_order, (const double*)L.gs_diag, f, v, (const int32_t*)L.gs_off, L.gs_groups);
        });
    }
    if (h->smoother == MGB_SM_GS_LEVEL) {
        return launch(h, MGB_K_GS, L.level, nb, [&] {
            int blocks = std::max(1, std::min(h->gs_coop_blocks_per_sm * h->sm_count, (L.gs_max_width + 255) / 256));
            const int32_t* rp = G.rowptr; const int32_t* cl = G.cols; const double* vl = G.vals;
            const int32_t* ord = L.gs_order; const double* dg = L.gs_diag; const int32_t* off = L.gs_off;
            int nlev = L.gs_groups;
            void* args[] = {(void*)&rp, (void*)&cl, (void*)&vl, (void*)&ord, (void*)&dg, (void*)&f, (void*)&v, (void*)&off, (void*)&nlev};
            cudaLaunchCooperativeKernel((void*)k_gs_levels, dim3(blocks), dim3(256), args, 0, h->stream);
        });
    }
    // multicolour: one launch per colour, in place (rows of one colour never read each other)
    for (int c = 0; c < L.gs_groups; ++c) {
        const double share = L.n ? (double)(L.col_off[c + 1] - L.col_off[c]) / (double)L.n : 0.0;
        EpiGaussSeidel epi{L.gs_order, L.gs_diag, f, v};
        TRY((row_sums<EpiGaussSeidel, false>(h, MGB_K_GS, L.level, nb * share, G, v, epi, c)));
    }
    return MGB_OK;
}

// nsweeps relaxation sweeps on level L.  v: current iterate, o: scratch of the same size (Jacobi
// ping-pong); on return v points at the result.  g_valid: L.g already holds w*(dinv*f) for this f.
int smooth(mgb_handle* h, Level& L, double*& v, double*& o, const double* f, int nsweeps, bool& g_valid)
{
    const double n = (double)L.n;
    for (int s = 0; s < nsweeps; ++s) {
        if (h->smoother == MGB_SM_JACOBI_RJ && s + 2 <= nsweeps && L.s2 && h->fuse_sweeps && h->allow_stream && h->stream_cfg > 0) {
            // two sweeps in one launch, the second trailing the first through L2 (k_hotrow2): v -> o -> v, no swap
            Sweep2Plan p;
            if (sweep2_plan(h, L.RJ, p) && p.groups == L.s2_groups) {
                const EpiJacobiRJ epi2{o, L.g, v, 1 - h->omega, h->omega};
                const double b1 = bytes_rowsum(L.RJ, g_valid ? 3 * n : 5 * n), b2 = bytes_rowsum(L.RJ, 3 * n);
                const double moved = b1 + b2 - 2.0 * coded_saving(L.RJ) - 16.0 * n;        // the second sweep's y and g are L2 hits
                if (!g_valid) {
                    const EpiJacobiRJFirst epi1{v, L.dinv, f, L.g, o, 1 - h->omega, h->omega};
                    TRY(launch(h, MGB_K_JACOBI2, L.level, b1 + b2, [&] { launch_hotrow2(h, L.RJ, p, L.s2, v, epi1, o, epi2); }, moved));
                } else {
                    const EpiJacobiRJ epi1{v, L.g, o, 1 - h->omega, h->omega};
                    TRY(launch(h, MGB_K_JACOBI2, L.level, b1 + b2, [&] { launch_hotrow2(h, L.RJ, p, L.s2, v, epi1, o, epi2); }, moved));
                }
                g_valid = true;
                ++s;
                continue;
            }
        }
        if (h->smoother == MGB_SM_JACOBI_RJ) {     // (the ghost entries of the iterate are exchanged inside row_sums_halo)
            if (!g_valid) {
                EpiJacobiRJFirst epi{v, L.dinv, f, L.g, o, 1 - h->omega, h->omega};
                TRY(row_sums_halo(h, MGB_K_JACOBI, L.level, bytes_rowsum(L.RJ, 5 * n), L.RJ, L, v, epi, nullptr, 0, 0, -1.0, &L, o));
                g_valid = true;
            } else {
                EpiJacobiRJ epi{v, L.g, o, 1 - h->omega, h->omega};
                TRY(row_sums_halo(h, MGB_K_JACOBI, L.level, bytes_rowsum(L.RJ, 3 * n), L.RJ, L, v, epi, nullptr, 0, 0, -1.0, &L, o));
            }
            std::swap(v, o);
        } else if (h->smoother == MGB_SM_JACOBI_A) {
            EpiJacobiA epi{v, L.dinv, f, o, h->omega};
            TRY(row_sums_halo(h, MGB_K_JACOBI, L.level, bytes_rowsum(L.A, 4 * n), L.A, L, v, epi));
            std::swap(v, o);
        } else {
            // row-sharded level: Gauss-Seidel inside every row block, the neighbours' unknowns taken from before the sweep
            // (block-Jacobi between GPUs) -- their current values are fetched once per sweep
            TRY(exchange(h, L, v));
            TRY(gs_sweep(h, L, v, f));
        }
    }
    return MGB_OK;
}

int residual(mgb_handle* h, Level& L, const double* v, const double* f, double* r)
{
    EpiResidual epi{f, r};
    return row_sums_halo(h, MGB_K_RESIDUAL, L.level, bytes_rowsum(L.A, 3.0 * (double)L.n), L.A, L, const_cast<double*>(v), epi);
}

// restriction from fine level L to its coarse neighbour: f_c = R r  (multigrid.py:251-252)
int restrict_to(mgb_handle* h, Level& L, const double* r_fine, double* f_coarse)
{
    const int64_t nc = L.n_coarse;
    if (L.r_mode == MGB_R_INJECTION) {
        return launch(h, MGB_K_RESTRICT, L.level, 20.0 * (double)nc, [&] {
            if (nc > 0) k_gather<<<(int)((nc + 255) / 256), 256, 0, h->stream>>>((int)nc, L.inj, r_fine, f_coarse);
        });
    }
    EpiStore epi{f_coarse};
    return row_sums_halo(h, MGB_K_RESTRICT, L.level, bytes_rowsum(L.R, (double)L.n + (double)nc), L.R, L, const_cast<double*>(r_fine), epi);
}

// fused residual + injection: f_c[i] = f[g_i] - (A v)[g_i]   (multigrid.py:244 + :128-131)
int residual_injected(mgb_handle* h, Level& L, const double* v, const double* f, double* f_coarse)
{
    const int64_t nc = L.n_coarse;
    const double avg = L.A.nrows ? (double)L.A.nnz / (double)L.A.nrows : 0.0;
    const int hl = L.A.cd.hotplan.hotlen;
    if (L.A.cd.mode == 3 && L.A.hcfg > 0 && (hl == 7 || hl == 15) && L.inj && nc > 0 && h->hot_inj && h->stream_cfg > 0 && h->allow_stream) {
        // thread per coarse row on the pattern-coded level matrix (k_hotinj): only the injected rows are summed
        const bool fused = h->in_cycle && L.fuse_ok && (v == L.v || v == L.vtmp);    // ghosts of an iterate are kept valid by the kernels that wrote it
        if (!fused) TRY(exchange(h, L, const_cast<double*>(v)));
        const HaloFuse hf = fused ? make_hf(h, &L, &L.A, nullptr, nullptr) : no_hf();
        const double nb = (12.0 * avg + 4.0 + 8.0 + 4.0 + 8.0 + 8.0) * (double)nc + 8.0 * (double)L.n;
        const double moved = 8.0 * (double)L.n + 21.0 * (double)nc;      // all of v (every line is touched), inj + code + f + result per coarse row
        return launch(h, MGB_K_RESIDUAL, L.level, nb, [&] {
            constexpr int T = 128;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)((nc + T - 1) / T)); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = h->pdl != 0 ? 1 : 0;
            const DevCsr& A = L.A;
            const int pf = h->hot_pf > 0 ? std::max(1, h->hot_pf / 8 / T) : 0;
            if (hl == 7) cudaLaunchKernelEx(&cfg, k_hotinj<7, T, 10>, (const unsigned char*)A.cd.codes, (const uint32_t*)A.cd.pmask, (const int2*)A.cd.phead,
                                            (const DictEnt*)A.cd.dict, A.cd.hot, hf, (const int32_t*)L.inj, L.inj_mono ? 1 : 0, (int)nc, (int)A.ncols, pf, v, f, f_coarse);
            else cudaLaunchKernelEx(&cfg, k_hotinj<15, T, 8>, (const unsigned char*)A.cd.codes, (const uint32_t*)A.cd.pmask, (const int2*)A.cd.phead,
                                    (const DictEnt*)A.cd.dict, A.cd.hot, hf, (const int32_t*)L.inj, L.inj_mono ? 1 : 0, (int)nc, (int)A.ncols, pf, v, f, f_coarse);
        }, moved);
    }
    if (L.inj_desc && (L.inj_fraction < 0.8 || L.A.cd.mode) && h->stream_cfg > 0 && h->allow_stream) {
        // stream only the tiles of A that hold injected rows; every row of such a tile is summed, injected ones are stored
        const double nb = L.inj_fraction * (12.0 * (double)L.A.nnz + (4.0 + 8.0 + 4.0) * (double)L.n) + 8.0 * (double)L.n + 8.0 * (double)nc;
        EpiResidualInject epi{f, L.cmap, f_coarse};
        return row_sums_halo(h, MGB_K_RESIDUAL, L.level, nb, L.A, L, const_cast<double*>(v), epi, L.inj_desc, L.inj_n_int, L.inj_n_bnd,
                             nb - L.inj_fraction * coded_saving(L.A));
    }
    TRY(exchange(h, L, const_cast<double*>(v)));
    return launch(h, MGB_K_RESIDUAL, L.level, (12.0 * avg + 4.0 + 8.0 + 4.0 + 8.0 + 8.0) * (double)nc + 8.0 * (double)L.n, [&] {
        if (nc <= 0) return;
        if (avg <= 8.0) k_residual_injected<8><<<(int)((nc * 8 + 255) / 256), 256, 0, h->stream>>>((int)nc, L.inj, L.A.rowptr, L.A.cols, L.A.vals, f, v, f_coarse);
        else k_residual_injected<16><<<(int)((nc * 16 + 255) / 256), 256, 0, h->stream>>>((int)nc, L.inj, L.A.rowptr, L.A.cols, L.A.vals, f, v, f_coarse);
    });
}

int prolong_add(mgb_handle* h, Level& L, const double* e_coarse, double* v_fine, double* err)
{
    Level* C = find_level(h, L.level - 1);
    EpiProlongAdd epi{v_fine, err};
    const double nb = bytes_rowsum(L.P, (double)L.n_coarse + 2.0 * (double)L.n);
    if (C && !C->gathered && !C->stub)          // (gathered level: everybody already has all of e)
        return row_sums_halo(h, MGB_K_PROLONG_ADD, L.level, nb, L.P, *C, const_cast<double*>(e_coarse), epi, nullptr, 0, 0, -1.0, &L, v_fine);
    // the corrected iterate's boundary rows still go to the neighbours when this level's exchange is fused
    h->hf_cur = kernel_takes_hf<EpiProlongAdd>(L.P) ? make_hf(h, nullptr, nullptr, &L, v_fine) : no_hf();
    const bool pushed = h->hf_cur.send_n > 0;
    int rc = row_sums(h, MGB_K_PROLONG_ADD, L.level, nb, L.P, e_coarse, epi);
    h->hf_cur = no_hf();
    if (rc == MGB_OK && !pushed && h->in_cycle && L.fuse_ok) rc = fail(h, MGB_ERR_STATE, "level %d: fused exchange without a fused prolongation", L.level);
    return rc;
}

int coarse_apply(mgb_handle* h, Level& C, const double* f, double* u)
{
    const int n = (int)C.n;
    const double nb = 8.0 * (double)n * (double)n + 16.0 * (double)n;
    TRY(launch(h, MGB_K_COARSE, C.level, nb, [&] {
        k_dense_gemv<<<(n * 32 + 255) / 256, 256, 0, h->stream>>>(n, h->coarse_inv, f, nullptr, u);
    }));
    if (h->coarse_refine) {          // u += Ainv (f - A u): one step of iterative refinement
        TRY(residual(h, C, u, f, C.r));
        TRY(launch(h, MGB_K_COARSE, C.level, nb, [&] {
            k_dense_gemv<<<(n * 32 + 255) / 256, 256, 0, h->stream>>>(n, h->coarse_inv, C.r, u, C.vtmp);
        }));
        CU(cudaMemcpyAsync(u, C.vtmp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
    }
    return MGB_OK;
}

int norm2_device(mgb_handle* h, int64_t n, const double* x, double* out_dev, int level)
{
    int rc = MGB_OK;
    TRY(launch(h, MGB_K_NORM, level, 8.0 * (double)n, [&] {
        k_sumsq_partial<<<h->norm_blocks, 256, 0, h->stream>>>(n, x, h->d_partial);
        k_sumsq_final<<<1, 1024, 0, h->stream>>>(h->norm_blocks, h->d_partial, out_dev, h->dist ? 0 : 1);
        if (h->dist) {          // sum of squares over all row blocks, then the root
            ncclResult_t r = g_nccl.AllReduce(out_dev, out_dev, 1, ncclDouble, ncclSum, h->comm, h->stream);
            if (r != ncclSuccess) rc = fail(h, MGB_ERR_COMM, "allreduce: %s", g_nccl.GetErrorString(r));
            k_sqrt_inplace<<<1, 1, 0, h->stream>>>(out_dev);
        }
    }));
    return rc;
}

// ---- coarse tail (k_tail): all levels of at most tail_rows rows in one cooperative launch ------------------------------
TailOp tail_op(const DevCsr& D)
{
    TailOp o{};
    o.mode = (D.cd.mode == 3 || D.cd.mode == 4) ? D.cd.mode : 0;
    o.rowptr = D.rowptr; o.cols = D.cols; o.vals = D.vals;
    o.codes = D.cd.codes; o.anchor = D.cd.anchor; o.phead = D.cd.phead; o.pent = D.cd.dict;
    return o;
}
// top level of the tail for a cycle whose top level is `top` (-1: no tail).  The tail covers the coarsest level and every
// level above it up to the last one that is small, lives entirely on this device and uses the reference configuration
// (weighted Jacobi in the reference form, injection fused with the residual).
int tail_top_for(mgb_handle* h, int top, bool debug)
{
    if (h->tail_rows <= 0 || debug || h->smoother != MGB_SM_JACOBI_RJ || h->mu1 < 1 || h->coarse_refine || !h->fuse_restrict) return -1;
    if (h->levels[h->coarsest].stub || !h->coarse_inv) return -1;
    int t = -1;
    for (int l = h->coarsest + 1; l < top && l - h->coarsest < TAIL_MAX_LEVELS; ++l) {
        const Level& L = h->levels[l];
        if (L.n > h->tail_rows || L.stub || L.n_ghost > 0 || !L.has_transfer || L.r_mode != MGB_R_INJECTION || !L.inj) break;
        if (h->dist && l > h->gather_level) break;
        if (!L.A.present() || !L.RJ.present() || !L.P.present() || !L.dinv) break;
        t = l;
    }
    return t;
}
int launch_tail(mgb_handle* h, int tail_top)
{
    TailPlan T{};
    T.nlev = tail_top - h->coarsest + 1; T.mu1 = h->mu1; T.mu2 = h->mu2; T.om = h->omega; T.om1 = 1 - h->omega;
    T.coarse_inv = h->coarse_inv;
    const bool swap = (((h->mu1 - 1) + h->mu2) & 1) != 0;    // the iterate must END in L.v
    double bytes = 0.0;
    for (int k = 0; k < T.nlev; ++k) {
        Level& L = h->levels[h->coarsest + k];
        TailLevel& t = T.lev[k];
        t.n = (int)L.n; t.nc = k > 0 ? (int)h->levels[h->coarsest + k - 1].n : 0;
        t.f = L.f; t.g = L.g; t.dinv = L.dinv; t.inj = L.inj;
        if (k == 0) { t.a = L.v; t.b = L.vtmp; bytes += 8.0 * (double)L.n * (double)L.n; continue; }
        t.a = swap ? L.vtmp : L.v; t.b = swap ? L.v : L.vtmp;
        t.A = tail_op(L.A); t.RJ = tail_op(L.RJ); t.P = tail_op(L.P);
        bytes += (double)(h->mu1 - 1 + h->mu2) * bytes_rowsum(L.RJ, 3.0 * (double)L.n) + 32.0 * (double)L.n + bytes_rowsum(L.P, 3.0 * (double)L.n);
    }
    if (h->tail_cluster) {                                   // one 8-CTA cluster, hardware cluster barrier between the phases
        return launch(h, MGB_K_COARSE, tail_top, bytes, [&] {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_tail<true>, T);
        });
    }
    return launch(h, MGB_K_COARSE, tail_top, bytes, [&] {
        static std::mutex mu;
        static std::map<int, int> grid_of;                  // device -> co-resident CTAs
        int grid = 0;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = grid_of.find(h->device);
            if (it == grid_of.end()) {
                int o = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tail<false>, 512, 0);
                it = grid_of.emplace(h->device, std::max(1, std::min(o, 1)) * h->sm_count).first;
            }
            grid = it->second;
        }
        void* args[] = {(void*)&T};
        cudaLaunchCooperativeKernel((void*)k_tail<false>, dim3(grid), dim3(512), args, 0, h->stream);
    });
}

// ---- one V-cycle, enqueued on the stream (multigrid.py:231-268 unrolled into a down and an up sweep) ----
// v and f of the top level live in levels[top].v / .f.  dbg: copy the test=True outputs into L.r (err_h);
// f2h / v2h are left in the coarse level's f / v buffers.
// top_g_valid: levels[top].g already holds w*(dinv*f) for the f in place (a previous cycle of the same call formed it: the
// product is the same every cycle, multigrid.py:226), so the top level's first sweep need not form and store it again.
int enqueue_cycle(mgb_handle* h, int top, bool debug, double** v2h_ptr, bool top_g_valid = false)
{
    if (top == h->coarsest) {          // multigrid.py:238-241
        Level& C = h->levels[top];
        if (C.stub) return fail(h, MGB_ERR_UNSUPPORTED, "level %d lives on rank 0 only", top);
        TRY(coarse_apply(h, C, C.f, C.vtmp));
        CU(cudaMemcpyAsync(C.v, C.vtmp, sizeof(double) * (size_t)C.n, cudaMemcpyDeviceToDevice, h->stream));
        return MGB_OK;
    }
    if (h->dist && top <= h->gather_level) return fail(h, MGB_ERR_UNSUPPORTED, "in a row-sharded hierarchy the top level must be a sharded level");
    std::map<int, double*> cur, oth;
    std::map<int, bool> gv;
    const bool jacobi = h->smoother == MGB_SM_JACOBI_RJ || h->smoother == MGB_SM_JACOBI_A;
    // ranks other than 0 stop at the gathered level: everything below it runs on rank 0 only
    const int bottom = h->coarsest;            // (on those ranks the gathered level IS the coarsest level they hold)
    const int tail_top = tail_top_for(h, top, debug);      // levels <= tail_top run in one cooperative launch (-1: none)
    for (int l = top; l > bottom; --l) {
        Level& L = h->levels[l];
        Level& C = h->levels[l - 1];
        if (l == tail_top) {                                // its right-hand side is in place: the whole sub-cycle, one launch
            TRY(launch_tail(h, tail_top));
            cur[l] = L.v;
            break;
        }
        double* v = L.v; double* o = L.vtmp;
        bool g_valid = l == top && top_g_valid && h->smoother == MGB_SM_JACOBI_RJ;
        int sweeps = h->mu1;
        if (l != top) {                                 // zero initial guess (multigrid.py:253)
            if (jacobi && h->mu1 > 0) {
                HaloFuse hf = make_hf(h, nullptr, nullptr, &L, v);         // sharded: the neighbours' ghosts of v are filled as well
                finish_hf(hf, 256, L.n);
                TRY(launch(h, MGB_K_INIT_GUESS, l, 32.0 * (double)L.n, [&] {
                    k_init_guess<<<(int)((L.n + 255) / 256), 256, 0, h->stream>>>((int)L.n, L.dinv, L.f, h->omega, L.g, v, hf);
                }));
                g_valid = (h->smoother == MGB_SM_JACOBI_RJ);
                sweeps = h->mu1 - 1;
            } else {
                CU(cudaMemsetAsync(v, 0, sizeof(double) * (size_t)L.n, h->stream));
            }
        }
        TRY(smooth(h, L, v, o, L.f, sweeps, g_valid));                    // multigrid.py:243
        double* fc = C.f + (C.gathered || C.stub ? C.my_off : 0);         // gathered level: this rank fills its slice
        if (L.r_mode == MGB_R_INJECTION && h->fuse_restrict && !debug) {
            TRY(residual_injected(h, L, v, L.f, fc));                      // multigrid.py:244 + :251
        } else {
            TRY(residual(h, L, v, L.f, L.r));                              // multigrid.py:244
            TRY(restrict_to(h, L, L.r, fc));                               // multigrid.py:251-252
        }
        if (h->dist && l - 1 == h->gather_level) TRY(gather_to_root(h, C, C.f));
        cur[l] = v; oth[l] = o; gv[l] = g_valid;
    }
    Level& C0 = h->levels[h->coarsest];
    if (!C0.stub && tail_top < 0) {
        TRY(coarse_apply(h, C0, C0.f, C0.v));                              // multigrid.py:238-241
        cur[h->coarsest] = C0.v;
    }
    for (int l = (tail_top < 0 ? bottom : tail_top) + 1; l <= top; ++l) {
        Level& L = h->levels[l];
        double* v = cur[l]; double* o = oth[l];
        bool g_valid = gv[l];
        double* err = (debug && l == top) ? L.r : nullptr;
        if (h->dist && l - 1 == h->gather_level) {                         // coarse-grid correction computed on rank 0
            Level& C = h->levels[l - 1];
            TRY(bcast_from_root(h, C, h->rank == 0 ? cur[l - 1] : C.v, C.v));
            cur[l - 1] = C.v;
        }
        TRY(prolong_add(h, L, cur[l - 1], v, err));                        // multigrid.py:258-260
        TRY(smooth(h, L, v, o, L.f, h->mu2, g_valid));                     // multigrid.py:261
        cur[l] = v;
    }
    Level& T = h->levels[top];
    if (cur[top] != T.v)                                  // (with the ghost section: it is valid in cur[top] when the exchange is fused)
        CU(cudaMemcpyAsync(T.v, cur[top], sizeof(double) * (size_t)(T.n + T.n_ghost), cudaMemcpyDeviceToDevice, h->stream));
    if (v2h_ptr) *v2h_ptr = cur[top - 1];
    return MGB_OK;
}

void drop_graphs(mgb_handle* h)
{
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
    h->graphs.clear();
    h->graph_kernels.clear();
}

int run_cycle(mgb_handle* h, int top, bool top_g_valid = false)
{
    const bool graphable = h->use_graph && !h->prof && h->smoother != MGB_SM_GS_LEVEL;
    if (h->smoother != MGB_SM_JACOBI_RJ || !h->reuse_g) top_g_valid = false;
    if (!graphable) return enqueue_cycle(h, top, false, nullptr, top_g_valid);
    const int key = top * 2 + (top_g_valid ? 1 : 0);          // two graphs per top level: first cycle of a call / the following ones
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        cudaGraph_t graph = nullptr;
        const int64_t before = h->launches;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
        int rc = enqueue_cycle(h, top, false, nullptr, top_g_valid);
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        const int64_t captured = h->launches - before;
        h->launches = before;
        if (rc != MGB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(h, MGB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, h->dist ? cudaGraphInstantiateFlagUseNodePriority : 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(h, MGB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        h->graphs[key] = exec;
        h->graph_kernels[key] = captured;
        it = h->graphs.find(key);
    }
    CU(cudaGraphLaunch(it->second, h->stream));
    h->launches += h->graph_kernels[key];
    return MGB_OK;
}

int ensure_hist(mgb_handle* h, int n)
{
    if (n <= h->hist_cap) return MGB_OK;
    cudaFree(h->d_hist);
    TRY(dev_alloc(h, &h->d_hist, (size_t)n));
    h->hist_cap = n;
    return MGB_OK;
}

// perm (device, optional): the level's numbering map, new index of every dof as the caller numbers them (mgb_set_numbering).
// The engine's buffers hold the lexicographic numbering; vectors are permuted on the way in and out.
__global__ void k_perm_scatter(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ src, double* __restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[perm[i]] = src[i];
}
__global__ void k_perm_gather(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ src, double* __restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}
int perm_scratch(mgb_handle* h, int64_t n)
{
    if (n <= h->perm_tmp_cap) return MGB_OK;
    cudaFree(h->perm_tmp); h->perm_tmp = nullptr; h->perm_tmp_cap = 0;
    TRY(dev_alloc(h, &h->perm_tmp, (size_t)n));
    h->perm_tmp_cap = n;
    return MGB_OK;
}
int copy_in(mgb_handle* h, double* dst, const double* src, int64_t n, int mem, const int32_t* perm = nullptr)
{
    if (dst == src || n == 0) return MGB_OK;
    if (perm) {
        const double* from = src;
        if (mem == MGB_MEM_HOST) {
            TRY(perm_scratch(h, n));
            CU(cudaMemcpyAsync(h->perm_tmp, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
            from = h->perm_tmp;
        }
        k_perm_scatter<<<(int)((n + 255) / 256), 256, 0, h->stream>>>(n, perm, from, dst);
        CU(cudaGetLastError());
        return MGB_OK;
    }
    CU(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, mem == MGB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
    return MGB_OK;
}
int copy_out(mgb_handle* h, double* dst, const double* src, int64_t n, int mem, const int32_t* perm = nullptr)
{
    if (dst == src || n == 0) return MGB_OK;
    if (perm) {
        double* to = dst;
        if (mem == MGB_MEM_HOST) { TRY(perm_scratch(h, n)); to = h->perm_tmp; }
        k_perm_gather<<<(int)((n + 255) / 256), 256, 0, h->stream>>>(n, perm, src, to);
        CU(cudaGetLastError());
        if (mem == MGB_MEM_HOST) CU(cudaMemcpyAsync(dst, to, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        return MGB_OK;
    }
    CU(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, mem == MGB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
    return MGB_OK;
}

// per-operator calls on borrowed device pointers: no padding / alignment guarantee -> no bulk-copy kernel
struct BorrowGuard {
    mgb_handle* h; bool saved;
    BorrowGuard(mgb_handle* h_, int mem) : h(h_), saved(h_->allow_stream) { if (mem == MGB_MEM_DEVICE) h->allow_stream = false; }
    ~BorrowGuard() { h->allow_stream = saved; }
};

int check_ready(mgb_handle* h, int level, Level** L)
{
    if (!h) return MGB_ERR_INVALID;
    if (!h->finalized) return fail(h, MGB_ERR_STATE, "mgb_finalize has not been called");
    *L = find_level(h, level);
    if (!*L) return fail(h, MGB_ERR_INVALID, "unknown level %d", level);
    CU(cudaSetDevice(h->device));
    return MGB_OK;
}

int cycles_on_buffers(mgb_handle* h, int top, int ncycles, double* resnorm_hist)
{
    Level& T = h->levels[top];
    if (resnorm_hist) TRY(ensure_hist(h, ncycles));
    // fused halo exchange: from here on every kernel that writes an iterate also fills the neighbours' ghosts of it; the iterate
    // the caller handed over gets its ghosts from one explicit exchange
    if (T.fuse_ok) TRY(exchange(h, T, T.v));
    struct InCycle { mgb_handle* h; ~InCycle() { h->in_cycle = false; h->hf_cur = no_hf(); } } guard{h};
    h->in_cycle = true;
    for (int c = 0; c < ncycles; ++c) {
        TRY(run_cycle(h, top, c > 0));          // (f is untouched between the cycles of one call)
        if (resnorm_hist) {                    // the residual the reference's driver forms after each cycle (multigrid.py:291)
            TRY(residual(h, T, T.v, T.f, T.r));
            TRY(norm2_device(h, T.n, T.r, h->d_hist + c, top));
        }
    }
    if (resnorm_hist) {
        CU(cudaMemcpyAsync(resnorm_hist, h->d_hist, sizeof(double) * (size_t)ncycles, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return MGB_OK;
}

// ---- caller numbering (mgb_set_numbering) -------------------------------------------------------------------------------
// rows of M renumbered by pr (new index of every old row), columns by pc; the ENTRY ORDER inside every row is kept, so each
// row sum adds the same products in the same order as before: results are bit-identical per row (nullptr: identity).
void permute_csr(HostCsr& M, const std::vector<int64_t>* pr, const std::vector<int64_t>* pc)
{
    if (M.empty() || (!pr && !pc)) return;
    const int64_t n = M.nrows;
    HostCsr N;
    N.nrows = M.nrows; N.ncols = M.ncols;
    N.ip.assign((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) N.ip[(size_t)(pr ? (*pr)[(size_t)i] : i) + 1] = M.ip[(size_t)i + 1] - M.ip[(size_t)i];
    for (int64_t i = 0; i < n; ++i) N.ip[(size_t)i + 1] += N.ip[(size_t)i];
    N.ix.resize(M.ix.size()); N.ax.resize(M.ax.size());
    for (int64_t i = 0; i < n; ++i) {
        int64_t o = N.ip[(size_t)(pr ? (*pr)[(size_t)i] : i)];
        for (int64_t k = M.ip[(size_t)i]; k < M.ip[(size_t)i + 1]; ++k, ++o) {
            N.ix[(size_t)o] = pc ? (int32_t)(*pc)[(size_t)M.ix[(size_t)k]] : M.ix[(size_t)k];
            N.ax[(size_t)o] = M.ax[(size_t)k];
        }
    }
    M = std::move(N);
}

}  // namespace
extern "C" int mgb_set_numbering(mgb_handle* h, int level, int64_t n, const int64_t* new_index)
{
    if (!h || !new_index) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    if (h->dist) return fail(h, MGB_ERR_UNSUPPORTED, "a caller numbering is not supported on row-sharded hierarchies");
    Level* L = find_level(h, level);
    if (!L || L->n != n) return fail(h, MGB_ERR_STATE, "set level %d (with %lld rows) before its numbering", level, (long long)n);
    std::vector<char> seen((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (new_index[i] < 0 || new_index[i] >= n || seen[(size_t)new_index[i]]) return fail(h, MGB_ERR_INVALID, "level %d: the numbering is not a permutation", level);
        seen[(size_t)new_index[i]] = 1;
    }
    L->perm_host.assign(new_index, new_index + n);
    h->numbered = true;
    return MGB_OK;
}
namespace {

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

int mgb_version(void) { return MGB_VERSION; }

const char* mgb_last_error(const mgb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mgb_create(mgb_handle** out, int device_id)
{
    mgb_handle* h = nullptr;
    if (!out) return fail(h, MGB_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(h, MGB_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device_id < 0 || device_id >= count) return fail(h, MGB_ERR_INVALID, "device %d out of range (0..%d)", device_id, count - 1);
    e = cudaSetDevice(device_id);
    if (e != cudaSuccess) return fail(h, MGB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    mgb_handle* H = new mgb_handle();
    H->device = device_id;
    e = cudaStreamCreateWithFlags(&H->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete H; return fail(h, MGB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_id);
    H->sm_count = prop.multiProcessorCount;
    H->norm_blocks = 4 * H->sm_count;
    cudaMalloc((void**)&H->d_partial, sizeof(double) * (size_t)H->norm_blocks);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&H->gs_coop_blocks_per_sm, k_gs_levels, 256, 0);
    *out = H;
    return MGB_OK;
}

int mgb_destroy(mgb_handle* h)
{
    if (!h) return MGB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drop_graphs(h);
    for (auto& kv : h->levels) {
        Level& L = kv.second;
        free_csr(L.A); free_csr(L.RJ); free_csr(L.P); free_csr(L.R); free_csr(L.G); free_csr(L.M); cudaFree(L.b); cudaFree(L.uex);
        cudaFree(L.dinv); cudaFree(L.inj); cudaFree(L.cmap); cudaFree(L.perm); cudaFree(L.inj_desc); cudaFree(L.send_idx); cudaFree(L.send_buf); cudaFree(L.p2p_counters);
        for (void* q : L.p2p_opened) cudaIpcCloseMemHandle(q);
        if (!L.vec_in_arena) { cudaFree(L.v); cudaFree(L.vtmp); }
        cudaFree(L.p2p_arena); cudaFree(L.fuse_counters); cudaFree(L.f); cudaFree(L.r); cudaFree(L.g);
        cudaFree(L.gs_order); cudaFree(L.gs_off); cudaFree(L.gs_diag); cudaFree(L.gs_ecols); cudaFree(L.gs_evals); cudaFree(L.s2);
    }
    cudaFree(h->coarse_inv); cudaFree(h->d_partial); cudaFree(h->d_hist); cudaFree(h->perm_tmp);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->comm_stream) { cudaStreamDestroy(h->comm_stream); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
    for (auto& pe : h->prof_events) { cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1); }
    cudaStreamDestroy(h->stream);
    delete h;
    return MGB_OK;
}

int mgb_set_level(mgb_handle* h, int level, int64_t n, int64_t nnz, const void* indptr, int indptr_bytes,
                  const int32_t* indices, const double* values)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    // refused before a single entry is read: one device shard addresses its entries with int32 row pointers
    if (n >= (int64_t)2147483000 || nnz >= (int64_t)2147483000)
        return fail(h, MGB_ERR_UNSUPPORTED, "level %d: %lld rows / %lld entries exceed the int32 row-pointer range of one device shard (row-shard it: mgb_set_level_local)",
                    level, (long long)n, (long long)nnz);
    Level& L = h->levels[level];
    L.level = level; L.n = n;
    std::string e = import_csr(L.A_host, n, n, nnz, indptr, indptr_bytes, indices, values);
    if (!e.empty()) { h->levels.erase(level); return fail(h, MGB_ERR_INVALID, "level %d: %s", level, e.c_str()); }
    return MGB_OK;
}

int mgb_synth_poisson_level(mgb_handle* h, int level, int dim, int cells_per_dim, int64_t row_begin, int64_t row_end,
                            int64_t ghost_lo, int64_t ghost_hi)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    if ((dim != 2 && dim != 3) || cells_per_dim < 2) return fail(h, MGB_ERR_INVALID, "dim must be 2 or 3, cells per dim >= 2");
    const SynthGeom g = make_geom(dim, cells_per_dim);
    if (row_begin < 0 || row_end > g.n || row_begin > row_end || ghost_lo < 0 || ghost_lo > row_begin || ghost_hi < row_end || ghost_hi > g.n)
        return fail(h, MGB_ERR_INVALID, "bad row / ghost ranges");
    const int64_t band = g.lin[g.noff - 1];
    if ((row_begin - ghost_lo < band && ghost_lo > 0) || (ghost_hi - row_end < band && ghost_hi < g.n))
        return fail(h, MGB_ERR_INVALID, "ghost ranges must cover the matrix bandwidth (%lld rows) or reach the ends", (long long)band);
    CU(cudaSetDevice(h->device));
    Level& L = h->levels[level];
    if (L.A.present()) return fail(h, MGB_ERR_STATE, "level %d already set", level);
    const int64_t n = row_end - row_begin;
    if (n >= 2147483000LL) return fail(h, MGB_ERR_UNSUPPORTED, "more than 2^31 rows in one shard");
    L.level = level; L.n = n; L.n_ghost = (row_begin - ghost_lo) + (ghost_hi - row_end);
    L.device_born = true; L.syn_dim = dim; L.syn_m = cells_per_dim;
    L.row_begin = row_begin; L.row_end = row_end; L.ghost_lo = ghost_lo; L.ghost_hi = ghost_hi;
    const int nr = (int)n;
    TRY(dev_alloc(h, &L.A.rowptr, (size_t)n + 1 + 8));
    CU(cudaMemsetAsync(L.A.rowptr, 0, ((size_t)n + 1 + 8) * sizeof(int32_t), h->stream));
    if (nr > 0) k_synth_a_count<<<(nr + 255) / 256, 256, 0, h->stream>>>(g, row_begin, nr, L.A.rowptr);
    TRY(scan_counts(h, L.A.rowptr, n + 1));
    TRY(alloc_csr_from_counts(h, L.A, n, n + L.n_ghost));
    const ColMap cm{row_begin, row_end, ghost_lo};
    if (nr > 0) k_synth_a_fill<<<(nr + 255) / 256, 256, 0, h->stream>>>(g, row_begin, nr, L.A.rowptr, L.A.cols, L.A.vals, cm);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    return MGB_OK;
}

int mgb_synth_poisson_transfer(mgb_handle* h, int coarse_level, int64_t inj_coarse_begin, int64_t inj_coarse_end)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    Level* F = find_level(h, coarse_level + 1);
    Level* C = find_level(h, coarse_level);
    if (!F || !C || !F->device_born) return fail(h, MGB_ERR_STATE, "generate level %d (mgb_synth_poisson_level) and set level %d first", coarse_level + 1, coarse_level);
    if (F->syn_m % 2) return fail(h, MGB_ERR_INVALID, "fine level has an odd number of cells per dimension");
    CU(cudaSetDevice(h->device));
    const int dim = F->syn_dim, Nf = F->syn_m + 1, Nc = F->syn_m / 2 + 1;
    int64_t nc_glob = 1; for (int d = 0; d < dim; ++d) nc_glob *= Nc;
    ColMap cm{0, nc_glob, 0};                       // coarse level held in full (single GPU, or gathered on rank 0)
    int64_t expect_rows = (C->gathered || C->stub) ? C->my_cnt : C->n;
    if (C->device_born) cm = ColMap{C->row_begin, C->row_end, C->ghost_lo};
    else if (C->n != nc_glob) return fail(h, MGB_ERR_INVALID, "coarse level %d has %lld rows, the nested grid has %lld", coarse_level, (long long)C->n, (long long)nc_glob);
    const int64_t ncr = inj_coarse_end - inj_coarse_begin;
    if (ncr != expect_rows) return fail(h, MGB_ERR_INVALID, "this rank must produce %lld coarse rows, got %lld", (long long)expect_rows, (long long)ncr);
    const int nr = (int)F->n;
    TRY(dev_alloc(h, &F->P.rowptr, (size_t)F->n + 1 + 8));
    CU(cudaMemsetAsync(F->P.rowptr, 0, ((size_t)F->n + 1 + 8) * sizeof(int32_t), h->stream));
    if (nr > 0) k_synth_p_count<<<(nr + 255) / 256, 256, 0, h->stream>>>(dim, Nf, F->row_begin, nr, F->P.rowptr);
    TRY(scan_counts(h, F->P.rowptr, F->n + 1));
    TRY(alloc_csr_from_counts(h, F->P, F->n, C->n + C->n_ghost));
    if (nr > 0) k_synth_p_fill<<<(nr + 255) / 256, 256, 0, h->stream>>>(dim, Nf, Nc, F->row_begin, nr, F->P.rowptr, F->P.cols, F->P.vals, cm);
    TRY(dev_alloc(h, &F->inj, (size_t)ncr + 16));
    if (ncr > 0) k_synth_inj<<<(int)((ncr + 255) / 256), 256, 0, h->stream>>>(dim, Nf, Nc, inj_coarse_begin, (int)ncr, F->row_begin, F->inj);
    CU(cudaGetLastError());
    F->inj_host.resize((size_t)ncr);
    CU(cudaMemcpyAsync(F->inj_host.data(), F->inj, sizeof(int32_t) * (size_t)ncr, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < ncr; ++i)
        if (F->inj_host[i] < 0 || F->inj_host[i] >= F->n) return fail(h, MGB_ERR_INVALID, "row blocks are not aligned with the injection map");
    F->r_mode = MGB_R_INJECTION; F->n_coarse = ncr; F->has_transfer = true;
    return MGB_OK;
}

int mgb_dist_unique_id(void* out, int capacity)
{
    mgb_handle* h = nullptr;
    if (!out || capacity < (int)sizeof(ncclUniqueId)) return fail(h, MGB_ERR_INVALID, "unique id buffer needs %d bytes", (int)sizeof(ncclUniqueId));
    if (const char* e = load_nccl()) return fail(h, MGB_ERR_COMM, "%s", e);
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(h, MGB_ERR_COMM, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    std::memcpy(out, &id, sizeof id);
    return MGB_OK;
}

int mgb_dist_init(mgb_handle* h, int rank, int world, const void* unique_id, int id_bytes)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized || h->dist) return fail(h, MGB_ERR_STATE, "mgb_dist_init must be called once, before the hierarchy is finalized");
    if (world < 1 || rank < 0 || rank >= world || !unique_id || id_bytes < (int)sizeof(ncclUniqueId)) return fail(h, MGB_ERR_INVALID, "bad rank/world/id");
    if (const char* e = load_nccl()) return fail(h, MGB_ERR_COMM, "%s", e);
    CU(cudaSetDevice(h->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof id);
    NC(g_nccl.CommInitRank(&h->comm, world, id, rank));
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_hi));   // exchange kernels jump the queue
    CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->dist = true; h->rank = rank; h->world = world;
    return MGB_OK;
}

int mgb_set_level_local(mgb_handle* h, int level, int64_t n_owned, int64_t n_ghost, int64_t nnz, const void* indptr, int indptr_bytes,
                        const int32_t* indices, const double* values)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    if (n_ghost < 0) return fail(h, MGB_ERR_INVALID, "negative ghost count");
    Level& L = h->levels[level];
    L.level = level; L.n = n_owned; L.n_ghost = n_ghost;
    std::string e = import_csr(L.A_host, n_owned, n_owned + n_ghost, nnz, indptr, indptr_bytes, indices, values);
    if (!e.empty()) { h->levels.erase(level); return fail(h, MGB_ERR_INVALID, "level %d: %s", level, e.c_str()); }
    return MGB_OK;
}

int mgb_set_halo(mgb_handle* h, int level, int npeers, const int32_t* peer_ranks, const int32_t* send_counts,
                 const int32_t* send_indices, const int32_t* recv_counts)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    if (!h->dist) return fail(h, MGB_ERR_STATE, "mgb_dist_init first");
    Level* L = find_level(h, level);
    if (!L) return fail(h, MGB_ERR_STATE, "set level %d before its halo plan", level);
    if (npeers < 0 || (npeers > 0 && (!peer_ranks || !send_counts || !recv_counts))) return fail(h, MGB_ERR_INVALID, "bad halo plan");
    CU(cudaSetDevice(h->device));
    L->peers.assign(peer_ranks, peer_ranks + npeers);
    L->send_cnt.assign(send_counts, send_counts + npeers);
    L->recv_cnt.assign(recv_counts, recv_counts + npeers);
    int64_t st = 0, rt = 0;
    for (int p = 0; p < npeers; ++p) {
        if (peer_ranks[p] < 0 || peer_ranks[p] >= h->world || peer_ranks[p] == h->rank) return fail(h, MGB_ERR_INVALID, "bad peer rank %d", peer_ranks[p]);
        st += send_counts[p]; rt += recv_counts[p];
    }
    if (rt != L->n_ghost) return fail(h, MGB_ERR_INVALID, "level %d: receive counts (%lld) do not add up to the ghost count (%lld)", level, (long long)rt, (long long)L->n_ghost);
    for (int64_t k = 0; k < st; ++k)
        if (send_indices[k] < 0 || send_indices[k] >= L->n) return fail(h, MGB_ERR_INVALID, "send index out of the owned range");
    L->send_total = st;
    L->send_a[0] = L->send_a[1] = -1;                       // fused exchange: the send list of a neighbour must be one run of rows
    if (npeers <= 2) {
        int64_t off = 0;
        for (int p = 0; p < npeers; ++p) {
            bool run = send_counts[p] > 0;
            for (int k = 1; k < send_counts[p] && run; ++k) run = send_indices[off + k] == send_indices[off] + k;
            if (run) L->send_a[p] = send_indices[off];
            off += send_counts[p];
        }
    }
    cudaFree(L->send_idx); cudaFree(L->send_buf);
    TRY(dev_upload(h, &L->send_idx, send_indices, (size_t)st));
    TRY(dev_alloc(h, &L->send_buf, (size_t)st));
    return MGB_OK;
}

int mgb_p2p_export(mgb_handle* h, int level, void* blob, int capacity, int* size)
{
    if (!h || !size) return MGB_ERR_INVALID;
    *size = (int)sizeof(P2PBlob);
    if (!blob) return MGB_OK;
    if (capacity < (int)sizeof(P2PBlob)) return fail(h, MGB_ERR_INVALID, "blob needs %d bytes", (int)sizeof(P2PBlob));
    Level* L = find_level(h, level);
    if (!L || !h->dist) return fail(h, MGB_ERR_STATE, "level %d has no halo plan", level);
    if ((int)L->peers.size() > P2P_MAX_PEERS || h->world > P2P_FLAG_SLOTS) return fail(h, MGB_ERR_UNSUPPORTED, "too many neighbours for the peer-memory exchange");
    CU(cudaSetDevice(h->device));
    // arena: [flags][fused flags][stage 0][stage 1][pad to 256 B][v][vtmp] -- the iterate buffers live here so that the
    // neighbours can store their boundary rows straight into this rank's ghost sections (HaloFuse)
    const size_t np = (size_t)L->n + (size_t)L->n_ghost + 16;
    const size_t off_stage = 2 * P2P_FLAG_SLOTS * sizeof(unsigned long long);
    const size_t off_v0 = (off_stage + 2 * ((size_t)L->n_ghost + 2) * sizeof(double) + 255) / 256 * 256;
    const size_t off_v1 = off_v0 + (np * sizeof(double) + 255) / 256 * 256;
    if (!L->p2p_arena) {
        const size_t bytes = off_v1 + np * sizeof(double);
        CU(cudaMalloc(&L->p2p_arena, bytes));
        CU(cudaMemset(L->p2p_arena, 0, bytes));
        TRY(dev_alloc(h, &L->p2p_counters, 64));
        CU(cudaMemset(L->p2p_counters, 0, 64 * sizeof(unsigned long long)));
        TRY(dev_alloc(h, &L->fuse_counters, 16));
        CU(cudaMemset(L->fuse_counters, 0, 16 * sizeof(unsigned long long)));
        L->v = (double*)((char*)L->p2p_arena + off_v0);
        L->vtmp = (double*)((char*)L->p2p_arena + off_v1);
        L->vec_in_arena = true;
    }
    P2PBlob b{};
    CU(cudaIpcGetMemHandle(&b.handle, L->p2p_arena));
    b.n_ghost = L->n_ghost;
    b.n_owned = L->n;
    b.off_vec[0] = (long long)off_v0; b.off_vec[1] = (long long)off_v1;
    b.npeers = (int)L->peers.size();
    int ro = 0;
    for (int p = 0; p < b.npeers; ++p) { b.peer_rank[p] = L->peers[p]; b.recv_off[p] = ro; ro += L->recv_cnt[p]; }
    b.recv_off[b.npeers] = ro;
    std::memcpy(blob, &b, sizeof b);
    return MGB_OK;
}

int mgb_p2p_import(mgb_handle* h, int level, int peer_rank, const void* blob, int size)
{
    if (!h || !blob || size < (int)sizeof(P2PBlob)) return MGB_ERR_INVALID;
    Level* L = find_level(h, level);
    if (!L || !L->p2p_arena) return fail(h, MGB_ERR_STATE, "mgb_p2p_export level %d first", level);
    CU(cudaSetDevice(h->device));
    P2PBlob b;
    std::memcpy(&b, blob, sizeof b);
    int mine = -1;                                     // this rank's slot in the neighbour's ghost section
    for (int j = 0; j < b.npeers; ++j) if (b.peer_rank[j] == h->rank) mine = j;
    int p = -1;
    for (size_t j = 0; j < L->peers.size(); ++j) if (L->peers[j] == peer_rank) p = (int)j;
    if (mine < 0 || p < 0) return fail(h, MGB_ERR_INVALID, "ranks %d and %d do not list each other as neighbours on level %d", h->rank, peer_rank, level);
    if (b.recv_off[mine + 1] - b.recv_off[mine] != L->send_cnt[p]) return fail(h, MGB_ERR_INVALID, "send/receive counts of ranks %d and %d differ on level %d", h->rank, peer_rank, level);
    void* base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess));
    L->p2p_opened.push_back(base);
    unsigned long long* rflags = (unsigned long long*)base;
    double* rstage = (double*)((char*)base + 2 * P2P_FLAG_SLOTS * sizeof(unsigned long long));
    L->p2p.rflag[p] = rflags + h->rank;
    L->p2p.rflag2[p] = rflags + P2P_FLAG_SLOTS + h->rank;
    for (int q = 0; q < 2; ++q) L->p2p.rvec[p][q] = (double*)((char*)base + b.off_vec[q]) + b.n_owned + b.recv_off[mine];
    L->p2p.rstage[p][0] = rstage + b.recv_off[mine];
    L->p2p.rstage[p][1] = rstage + (size_t)b.n_ghost + b.recv_off[mine];
    L->p2p.peer_rank[p] = peer_rank;
    if (++L->p2p_imported == (int)L->peers.size()) {  // plan complete
        P2PPlan& pl = L->p2p;
        pl.npeers = (int)L->peers.size();
        int so = 0, ro = 0;
        for (int j = 0; j < pl.npeers; ++j) { pl.send_off[j] = so; pl.recv_off[j] = ro; so += L->send_cnt[j]; ro += L->recv_cnt[j]; }
        pl.send_off[pl.npeers] = so; pl.recv_off[pl.npeers] = ro;
        pl.counters = L->p2p_counters;
        pl.flags = (unsigned long long*)L->p2p_arena;
        pl.flags2 = (unsigned long long*)L->p2p_arena + P2P_FLAG_SLOTS;
        pl.stage = (double*)((char*)L->p2p_arena + 2 * P2P_FLAG_SLOTS * sizeof(unsigned long long));
        pl.n_ghost = (int)L->n_ghost;
        L->p2p_ready = true;
    }
    return MGB_OK;
}

int mgb_halo_fused(mgb_handle* h, int level, int* fused)
{
    if (!h || !fused) return MGB_ERR_INVALID;
    Level* L = find_level(h, level);
    if (!L) return fail(h, MGB_ERR_INVALID, "no level %d", level);
    *fused = L->fuse_ok ? 1 : 0;
    return MGB_OK;
}

int mgb_set_halo_fused(mgb_handle* h, int level, int on)
{
    if (!h) return MGB_ERR_INVALID;
    Level* L = find_level(h, level);
    if (!L) return fail(h, MGB_ERR_INVALID, "no level %d", level);
    if (on && !L->fuse_ok) return fail(h, MGB_ERR_STATE, "level %d: the fused exchange can only be switched off (it did not qualify at mgb_finalize)", level);
    if (!on && L->fuse_ok) { L->fuse_ok = false; drop_graphs(h); }
    return MGB_OK;
}

int mgb_set_gather_level(mgb_handle* h, int level, int64_t n_global, const int64_t* offsets)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    if (!h->dist) return fail(h, MGB_ERR_STATE, "mgb_dist_init first");
    if (!offsets || offsets[0] != 0 || offsets[h->world] != n_global) return fail(h, MGB_ERR_INVALID, "bad gather offsets");
    Level* L = find_level(h, level);
    if (h->rank == 0) {
        if (!L || L->n != n_global || L->n_ghost != 0) return fail(h, MGB_ERR_STATE, "rank 0 must hold level %d completely before it is declared gathered", level);
    } else {
        if (L) return fail(h, MGB_ERR_STATE, "level %d must not be set on ranks other than 0", level);
        L = &h->levels[level];
        L->level = level; L->n = n_global; L->stub = true;
    }
    L->gathered = true;
    L->gather_off.assign(offsets, offsets + h->world + 1);
    L->my_off = offsets[h->rank]; L->my_cnt = offsets[h->rank + 1] - offsets[h->rank];
    h->gather_level = level;
    return MGB_OK;
}

int mgb_set_transfer(mgb_handle* h, int coarse_level, int64_t n_fine, int64_t n_coarse,
                     int64_t p_nnz, const void* p_indptr, int p_indptr_bytes, const int32_t* p_indices, const double* p_values,
                     int r_mode, int dim_for_fw, const int32_t* inj,
                     int64_t r_nnz, const void* r_indptr, int r_indptr_bytes, const int32_t* r_indices, const double* r_values)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    Level* F = find_level(h, coarse_level + 1);
    Level* C = find_level(h, coarse_level);
    if (!F || !C) return fail(h, MGB_ERR_STATE, "set levels %d and %d before their transfer", coarse_level, coarse_level + 1);
    // sharded hierarchies: n_fine = owned fine rows; n_coarse = coarse rows this rank produces (its slice of a gathered
    // level); P has one column per owned-or-ghost coarse entry, R / inj address owned-or-ghost fine entries
    const int64_t nc_rows = (C->gathered || C->stub) ? C->my_cnt : C->n;
    if (F->n != n_fine || nc_rows != n_coarse) return fail(h, MGB_ERR_INVALID, "transfer sizes (%lld, %lld) do not match the levels (%lld, %lld)",
                                                           (long long)n_fine, (long long)n_coarse, (long long)F->n, (long long)nc_rows);
    const int64_t p_cols = C->n + C->n_ghost, r_cols = F->n + F->n_ghost;
    if (r_mode < MGB_R_INJECTION || r_mode > MGB_R_EXPLICIT) return fail(h, MGB_ERR_INVALID, "bad r_mode %d", r_mode);
    std::string e = import_csr(F->P_host, n_fine, p_cols, p_nnz, p_indptr, p_indptr_bytes, p_indices, p_values);
    if (!e.empty()) return fail(h, MGB_ERR_INVALID, "P: %s", e.c_str());
    F->r_mode = r_mode; F->dim_fw = dim_for_fw; F->n_coarse = n_coarse;
    if (r_mode == MGB_R_INJECTION) {
        if (!inj) return fail(h, MGB_ERR_INVALID, "injection list is null");
        F->inj_host.assign(inj, inj + n_coarse);
        for (int64_t i = 0; i < n_coarse; ++i)
            if (inj[i] < 0 || inj[i] >= n_fine) return fail(h, MGB_ERR_INVALID, "injection index out of range (or not owned by this rank) at %lld", (long long)i);
    } else if (r_mode == MGB_R_EXPLICIT) {
        e = import_csr(F->R_host, n_coarse, r_cols, r_nnz, r_indptr, r_indptr_bytes, r_indices, r_values);
        if (!e.empty()) return fail(h, MGB_ERR_INVALID, "R: %s", e.c_str());
    } else if (dim_for_fw != 2 && dim_for_fw != 3 && r_mode == MGB_R_FULL_WEIGHTING) {
        return fail(h, MGB_ERR_INVALID, "dim_for_fw must be 2 or 3");
    }
    F->has_transfer = true;
    return MGB_OK;
}

int mgb_set_restriction(mgb_handle* h, int coarse_level, int r_mode, int dim_for_fw)
{
    if (!h) return MGB_ERR_INVALID;
    if (h->finalized) return fail(h, MGB_ERR_STATE, "hierarchy already finalized");
    Level* F = find_level(h, coarse_level + 1);
    if (!F || !F->has_transfer) return fail(h, MGB_ERR_STATE, "set or generate the transfer between levels %d and %d first", coarse_level, coarse_level + 1);
    if (r_mode != MGB_R_INJECTION && r_mode != MGB_R_FULL_WEIGHTING && r_mode != MGB_R_TRANSPOSE)
        return fail(h, MGB_ERR_INVALID, "r_mode must be injection, full weighting or transpose (explicit rows come with mgb_set_transfer)");
    if (r_mode == MGB_R_INJECTION && F->inj_host.empty() && F->n_coarse > 0)
        return fail(h, MGB_ERR_INVALID, "the transfer was set without an injection list");
    if (r_mode == MGB_R_FULL_WEIGHTING && dim_for_fw != 2 && dim_for_fw != 3) return fail(h, MGB_ERR_INVALID, "dim_for_fw must be 2 or 3");
    F->r_mode = r_mode;
    if (r_mode == MGB_R_FULL_WEIGHTING) F->dim_fw = dim_for_fw;
    return MGB_OK;
}

int mgb_set_params(mgb_handle* h, double omega, int mu1, int mu2, int smoother)
{
    if (!h) return MGB_ERR_INVALID;
    if (mu1 < 0 || mu2 < 0) return fail(h, MGB_ERR_INVALID, "negative sweep count");
    if (smoother < MGB_SM_JACOBI_RJ || smoother > MGB_SM_GS_MULTICOLOR) return fail(h, MGB_ERR_INVALID, "bad smoother %d", smoother);
    if (h->finalized && smoother != h->smoother) {
        const bool was_gs = h->smoother >= MGB_SM_GS_LEVEL, is_gs = smoother >= MGB_SM_GS_LEVEL;
        if (is_gs || was_gs) return fail(h, MGB_ERR_STATE, "Gauss-Seidel operators are built at finalize: choose the smoother before mgb_finalize");
    }
    h->omega = omega; h->mu1 = mu1; h->mu2 = mu2; h->smoother = smoother;
    drop_graphs(h);
    return MGB_OK;
}

int mgb_set_option(mgb_handle* h, const char* key, double value)
{
    if (!h || !key) return MGB_ERR_INVALID;
    const std::string k(key);
    const int iv = (int)value;
    const bool pre = !h->finalized;
    if (k == "use_graph") { h->use_graph = iv; drop_graphs(h); }
    else if (k == "coarse_refine") { h->coarse_refine = iv; drop_graphs(h); }
    else if (k == "fuse_restrict") { h->fuse_restrict = iv; drop_graphs(h); }
    else if (k == "rj_order" && pre) h->rj_reversed = iv;
    else if (k == "kernel_family" && pre) h->opt_family = iv;
    else if (k == "lanes_per_row" && pre) h->opt_lpr = iv;
    else if (k == "tile_iter" && pre) h->opt_iter = iv;
    else if (k == "stream_cfg" && pre) h->stream_cfg = iv;
    else if (k == "stream_auto" && pre) h->stream_auto = iv;
    else if (k == "compress" && pre) h->compress = iv;
    else if (k == "code_cfg" && pre) h->code_cfg = iv;
    else if (k == "stage_x" && pre) h->stage_x = iv;
    else if (k == "win_cfg" && pre) h->win_cfg = iv;
    else if (k == "hot_cfg" && pre) h->hot_cfg = iv;
    else if (k == "anch_cfg" && pre) h->anch_cfg = iv;
    else if (k == "anch_tiles") { h->anch_tiles = iv == 0 ? 1 : iv; drop_graphs(h); }
    else if (k == "reuse_g") { h->reuse_g = iv; drop_graphs(h); }
    else if (k == "tail_rows") { h->tail_rows = iv; drop_graphs(h); }
    else if (k == "tail_cluster") { h->tail_cluster = iv; drop_graphs(h); }
    else if (k == "hot_inj") { h->hot_inj = iv; drop_graphs(h); }
    else if (k == "fuse_halo" && pre) h->fuse_halo = iv;
    else if (k == "device_setup" && pre) h->device_setup = iv;
    else if (k == "hot_pf") { h->hot_pf = iv; drop_graphs(h); }
    else if (k == "fuse_sweeps") { h->fuse_sweeps = iv; drop_graphs(h); }
    else if (k == "s2_slack") { h->s2_slack = std::min(std::max(0, iv), 1 << 20); drop_graphs(h); }
    else if (k == "s2_min_rows" && pre) h->s2_min_rows = iv;
    else if (k == "s2_tiles") { h->s2_tiles = std::max(1, iv); drop_graphs(h); }
    else if (k == "win_prefetch") { h->win_prefetch = iv; drop_graphs(h); }
    else if (k == "gs_cluster") h->gs_cluster = iv;
    else if (k == "pdl") { h->pdl = iv; drop_graphs(h); }
    else if (k == "p2p_enable") { h->p2p_enable = iv; drop_graphs(h); }
    else if (k == "overlap_halo") { h->overlap = iv; drop_graphs(h); }
    else if (k == "overlap_waves") { h->overlap_waves = iv; drop_graphs(h); }
    else return fail(h, pre ? MGB_ERR_INVALID : MGB_ERR_STATE, "option '%s' unknown or not settable %s finalize", key, pre ? "before" : "after");
    return MGB_OK;
}

// Gauss-Seidel set-up from the level matrix in HBM (mgb_devsetup.cu): symmetrised lower graph, level sets and / or first-fit
// colouring, stable order, reordered off-diagonal operator, ELL copy -- the same artefacts, bit for bit, as the host path in
// mgb_finalize (level_sets / greedy_colouring / split_offdiag / permute_rows of mgb_setup.cpp).  Above 32 M rows only the
// ordering the chosen smoother runs in is built (the other one's artefact is then empty).
extern "C++" {
struct DevTemps {
    std::vector<void*> p;
    template <class T> T* keep(T* q) { p.push_back((void*)q); return q; }
    void release(void* q) { p.erase(std::remove(p.begin(), p.end(), q), p.end()); }
    ~DevTemps() { for (void* q : p) cudaFree(q); }
};

static int build_gs_device(mgb_handle* h, Level& L)
{
    const int n = (int)L.n;
    const bool lvl = h->smoother == MGB_SM_GS_LEVEL, both = L.n <= ((int64_t)32 << 20);
    cudaStream_t s = h->stream;
    DevTemps tmp;
    int32_t *lp = nullptr, *lx = nullptr, *lev = nullptr, *col = nullptr, *lev_order = nullptr, *col_order = nullptr;
    CU(dev::lower_sym_graph(s, n, L.A.rowptr, L.A.cols, L.A.vals, &lp, &lx));
    tmp.keep(lp); tmp.keep(lx);
    auto download = [&](std::vector<int32_t>& dst, const int32_t* src) -> int {
        dst.resize((size_t)n);
        if (n) CU(cudaMemcpyAsync(dst.data(), src, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return MGB_OK;
    };
    if (lvl || both) {
        CU(dev::level_sets(s, n, lp, lx, &lev, &L.gs_setup_passes[0]));
        tmp.keep(lev);
        CU(dev::order_from_keys(s, n, lev, &lev_order, L.lev_off));
        tmp.keep(lev_order);
        TRY(download(L.lev_of_row, lev)); TRY(download(L.lev_order, lev_order));
    }
    if (!lvl || both) {
        int overflow = 0;
        CU(dev::colouring(s, n, lp, lx, &col, &L.gs_setup_passes[1], &overflow));
        tmp.keep(col);
        if (overflow) return fail(h, MGB_ERR_UNSUPPORTED, "level %d: more than 128 colours (the device colouring holds the used colours in two words)", L.level);
        CU(dev::order_from_keys(s, n, col, &col_order, L.col_off));
        tmp.keep(col_order);
        TRY(download(L.col_of_row, col)); TRY(download(L.col_order, col_order));
    }
    int32_t* order = lvl ? lev_order : col_order;
    const std::vector<int32_t>& off = lvl ? L.lev_off : L.col_off;
    int64_t gnnz = 0; int bad = 0;
    CU(dev::gs_operator(s, n, order, L.A.rowptr, L.A.cols, L.A.vals, &L.G.rowptr, &L.G.cols, &L.G.vals, &L.gs_diag, &gnnz, &bad));
    if (bad) return fail(h, MGB_ERR_SINGULAR, "level %d: zero or missing diagonal entry", L.level);
    L.G.nrows = n; L.G.ncols = L.A.ncols; L.G.nnz = gnnz;
    std::vector<int64_t> ip;
    TRY(fetch_rowptr(h, L.G, ip));
    L.gs_groups = (int)off.size() - 1;
    for (int g = 0; g < L.gs_groups; ++g) L.gs_max_width = std::max(L.gs_max_width, off[g + 1] - off[g]);
    if (lvl) {
        const int sf = h->opt_family; h->opt_family = 2;      // rows are read directly by k_gs_levels
        int rc = finish_csr(h, L.G, ip); h->opt_family = sf; TRY(rc);
    } else {
        TRY(finish_csr(h, L.G, ip, off));
    }
    if (lvl && L.G.max_row <= 8 && n > 0) {                  // ELL copy for the pipelined kernel
        const int W = L.G.max_row <= 4 ? 4 : (L.G.max_row <= 6 ? 6 : 8);
        CU(dev::gs_ell(s, n, W, L.G.rowptr, L.G.cols, L.G.vals, &L.gs_ecols, &L.gs_evals));
        L.gs_W = W;
    }
    tmp.release(order);
    L.gs_order = order;
    TRY(dev_upload(h, &L.gs_off, off.data(), off.size()));
    return MGB_OK;
}
}  // extern "C++"

// everything mgb_finalize can reject without touching the device or the network (also exported as mgb_precheck so
// that a row-sharded setup can agree on success BEFORE entering the collective part of mgb_finalize)
static int validate_hierarchy(mgb_handle* h)
{
    if (h->finalized) return fail(h, MGB_ERR_STATE, "already finalized");
    if (h->levels.empty()) return fail(h, MGB_ERR_STATE, "no levels set");
    h->coarsest = h->levels.begin()->first;
    h->finest = h->levels.rbegin()->first;
    for (int l = h->coarsest; l <= h->finest; ++l) {
        Level* L = find_level(h, l);
        if (!L) return fail(h, MGB_ERR_STATE, "level %d missing (levels must be contiguous)", l);
        if (l > h->coarsest && !L->has_transfer) return fail(h, MGB_ERR_STATE, "transfer between levels %d and %d missing", l - 1, l);
        if (h->dist && l > h->gather_level && L->has_transfer && (L->r_mode == MGB_R_FULL_WEIGHTING || L->r_mode == MGB_R_TRANSPOSE))
            return fail(h, MGB_ERR_UNSUPPORTED, "row-sharded levels need the restriction rows explicitly (MGB_R_EXPLICIT) or injection");
        if (L->device_born && L->has_transfer && L->r_mode == MGB_R_EXPLICIT)
            return fail(h, MGB_ERR_UNSUPPORTED, "generated levels restrict by injection or by the transposed prolongation (mgb_set_restriction)");
    }
    if (h->opt_iter && h->opt_iter != 1 && h->opt_iter != 2) return fail(h, MGB_ERR_INVALID, "tile_iter must be 1 or 2");
    if (h->opt_lpr && (h->opt_lpr & (h->opt_lpr - 1) || h->opt_lpr > 32)) return fail(h, MGB_ERR_INVALID, "lanes_per_row must be a power of two <= 32");
    if (h->dist && h->world > 1 && h->gather_level == INT_MIN)
        return fail(h, MGB_ERR_STATE, "row-sharded hierarchy without a gathered level (mgb_set_gather_level)");
    return MGB_OK;
}

int mgb_precheck(mgb_handle* h)
{
    if (!h) return MGB_ERR_INVALID;
    return validate_hierarchy(h);
}

int mgb_finalize(mgb_handle* h)
{
    if (!h) return MGB_ERR_INVALID;
    TRY(validate_hierarchy(h));
    CU(cudaSetDevice(h->device));
    if (h->numbered) {                      // caller numbering: renumber every operator once, here (entry order inside rows kept)
        // the coarsest level keeps the caller's numbering: its dense inverse is computed by elimination, whose rounding follows the
        // row order -- renumbering it would change the coarse solve in the last bits (and there is nothing to gain on ~1000 rows)
        h->levels[h->coarsest].perm_host.clear();
        for (auto& kv : h->levels) {
            Level& L = kv.second;
            const std::vector<int64_t>* pf = L.perm_host.empty() ? nullptr : &L.perm_host;
            Level* C = find_level(h, kv.first - 1);
            const std::vector<int64_t>* pc = (C && !C->perm_host.empty()) ? &C->perm_host : nullptr;
            if (L.device_born) return fail(h, MGB_ERR_UNSUPPORTED, "a caller numbering needs host-assembled levels");
            permute_csr(L.A_host, pf, pf);
            if (L.has_transfer) {
                permute_csr(L.P_host, pf, pc);
                permute_csr(L.R_host, pc, pf);
                if (!L.inj_host.empty()) {
                    std::vector<int32_t> inj(L.inj_host.size());
                    for (size_t i = 0; i < inj.size(); ++i)
                        inj[(size_t)(pc ? (*pc)[i] : (int64_t)i)] = (int32_t)(pf ? (*pf)[(size_t)L.inj_host[i]] : L.inj_host[i]);
                    L.inj_host.swap(inj);
                }
            }
            if (pf) {
                std::vector<int32_t> p32(pf->begin(), pf->end());
                TRY(dev_upload(h, &L.perm, p32.data(), p32.size()));
            }
        }
    }
    for (auto& kv : h->levels) {
        Level& L = kv.second;
        const size_t n = (size_t)L.n;
        if (L.stub) {                       // gathered level on a rank other than 0: vectors only
            const size_t np = n + 16;
            TRY(dev_alloc(h, &L.v, np)); TRY(dev_alloc(h, &L.f, np));
            for (double* p : {L.v, L.f}) CU(cudaMemsetAsync(p, 0, np * sizeof(double), h->stream));
            continue;
        }
        const int64_t xo_self = (h->dist && L.n_ghost > 0) ? L.n : -1;       // operators reading this level's vectors
        if (L.device_born) {
            std::vector<int64_t> ip;
            int64_t in[2];
            TRY(fetch_rowptr(h, L.A, ip));
            if (xo_self >= 0) TRY(interior_rows(h, L.A, xo_self, in));
            TRY(finish_csr(h, L.A, ip, {}, xo_self >= 0 ? in : nullptr));
            TRY(build_rj_device(h, L));                                        // multigrid.py:48-56 on the device
        } else {
            TRY(upload_csr(h, L.A_host, L.A, {}, xo_self));
            HostCsr RJ; std::vector<double> dinv;
            if (!build_rj(L.A_host, h->rj_reversed != 0, RJ, dinv))            // multigrid.py:48-56
                return fail(h, MGB_ERR_SINGULAR, "level %d: zero or missing diagonal entry", kv.first);
            TRY(upload_csr(h, RJ, L.RJ, {}, xo_self));
            TRY(dev_upload(h, &L.dinv, dinv.data(), n, 16));
        }
        {   // counters of the two-sweep kernel (k_hotrow2): unsharded hot-row smoother matrices of at least s2_min_rows rows
            Sweep2Plan p;
            if (L.n_ghost == 0 && !(h->dist && !L.peers.empty()) && L.n >= h->s2_min_rows && sweep2_plan(h, L.RJ, p)) {
                const size_t cnt = (size_t)p.groups + 2;
                TRY(dev_alloc(h, &L.s2, cnt));
                CU(cudaMemsetAsync(L.s2, 0, cnt * sizeof(unsigned long long), h->stream));
                L.s2_groups = p.groups;
            }
        }
        const bool on_device = L.device_born || h->device_setup;       // set-up part 2 from the arrays in HBM (mgb_devsetup.cu)
        if (h->smoother >= MGB_SM_GS_LEVEL && kv.first > h->coarsest && on_device) {
            TRY(build_gs_device(h, L));
        } else if (h->smoother >= MGB_SM_GS_LEVEL && kv.first > h->coarsest) {
            HostCsr G0, G; std::vector<double> diag, dperm(n);
            if (!split_offdiag(L.A_host, G0, diag)) return fail(h, MGB_ERR_SINGULAR, "level %d: zero or missing diagonal entry", kv.first);
            level_sets(L.A_host, L.lev_of_row, L.lev_order, L.lev_off);
            greedy_colouring(L.A_host, L.col_of_row, L.col_order, L.col_off);
            const bool lvl = h->smoother == MGB_SM_GS_LEVEL;
            const std::vector<int32_t>& order = lvl ? L.lev_order : L.col_order;
            const std::vector<int32_t>& off = lvl ? L.lev_off : L.col_off;
            permute_rows(G0, order, G);
            for (size_t p = 0; p < n; ++p) dperm[p] = diag[order[p]];
            L.gs_groups = (int)off.size() - 1;
            for (int g = 0; g < L.gs_groups; ++g) L.gs_max_width = std::max(L.gs_max_width, off[g + 1] - off[g]);
            if (lvl) {
                const int sf = h->opt_family; h->opt_family = 2;      // rows are read directly by k_gs_levels
                int rc = upload_csr(h, G, L.G); h->opt_family = sf; TRY(rc);
            } else {
                TRY(upload_csr(h, G, L.G, off));
            }
            if (lvl && L.G.max_row <= 8 && n > 0) {            // ELL copy for the pipelined kernel
                const int W = L.G.max_row <= 4 ? 4 : (L.G.max_row <= 6 ? 6 : 8);
                std::vector<int32_t> ec((size_t)W * n, 0);
                std::vector<double> ev((size_t)W * n, 0.0);
                for (size_t p = 0; p < n; ++p)
                    for (int64_t k = G.ip[p]; k < G.ip[p + 1]; ++k) {
                        ec[(size_t)(k - G.ip[p]) * n + p] = G.ix[k];
                        ev[(size_t)(k - G.ip[p]) * n + p] = G.ax[k];
                    }
                TRY(dev_upload(h, &L.gs_ecols, ec.data(), ec.size()));
                TRY(dev_upload(h, &L.gs_evals, ev.data(), ev.size()));
                L.gs_W = W;
            }
            TRY(dev_upload(h, &L.gs_order, order.data(), n));
            TRY(dev_upload(h, &L.gs_off, off.data(), off.size()));
            TRY(dev_upload(h, &L.gs_diag, dperm.data(), n));
        }
        if (L.has_transfer) {
            Level& Cl = h->levels[kv.first - 1];
            const int64_t xo_coarse = (h->dist && Cl.n_ghost > 0) ? Cl.n : -1;     // P reads the coarse level's vectors
            if (L.device_born) {
                std::vector<int64_t> ip;
                int64_t in[2];
                TRY(fetch_rowptr(h, L.P, ip));
                if (xo_coarse >= 0) TRY(interior_rows(h, L.P, xo_coarse, in));
                TRY(finish_csr(h, L.P, ip, {}, xo_coarse >= 0 ? in : nullptr));
            } else {
                TRY(upload_csr(h, L.P_host, L.P, {}, xo_coarse));
            }
            if (L.r_mode == MGB_R_INJECTION) {
                if (!L.device_born) TRY(dev_upload(h, &L.inj, L.inj_host.data(), L.inj_host.size()));
                L.inj_mono = false;
                if (L.inj && L.n_coarse > 0) {                // ascending injection list: a CTA of k_hotinj knows the row range it touches
                    int* bad = nullptr;
                    TRY(dev_alloc(h, &bad, 1));
                    CU(cudaMemsetAsync(bad, 0, sizeof(int), h->stream));
                    k_not_ascending<<<(int)((L.n_coarse + 255) / 256), 256, 0, h->stream>>>((int)L.n_coarse, L.inj, bad);
                    int hb = 1;
                    CU(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
                    CU(cudaStreamSynchronize(h->stream));
                    cudaFree(bad);
                    L.inj_mono = hb == 0;
                }
                if (!L.A.sdesc_host.empty()) {
                    std::vector<int32_t> cmap(n + 16, -1);
                    for (size_t c = 0; c < L.inj_host.size(); ++c) cmap[(size_t)L.inj_host[c]] = (int32_t)c;
                    std::vector<int4> keep, keep_bnd;          // [interior tiles | boundary tiles]
                    int64_t ent = 0;
                    for (size_t t = 0; t < L.A.sdesc_host.size(); ++t) {
                        const int4& d = L.A.sdesc_host[t];
                        bool any = false;
                        for (int r = d.x; r < d.x + d.y && !any; ++r) any = cmap[(size_t)r] >= 0;
                        if (!any) continue;
                        ent += d.w;
                        const bool interior = !L.A.split || ((int)t >= L.A.t_int0 && (int)t < L.A.t_int1);
                        (interior ? keep : keep_bnd).push_back(d);
                    }
                    L.inj_n_int = (int)keep.size(); L.inj_n_bnd = (int)keep_bnd.size();
                    keep.insert(keep.end(), keep_bnd.begin(), keep_bnd.end());
                    TRY(dev_upload(h, &L.cmap, cmap.data(), cmap.size()));
                    if (!keep.empty()) TRY(dev_upload(h, &L.inj_desc, keep.data(), keep.size()));
                    L.inj_ntiles = (int)keep.size();
                    L.inj_fraction = L.A.nnz ? (double)ent / (double)L.A.nnz : 1.0;
                }
            } else {
                const double scale = L.r_mode == MGB_R_FULL_WEIGHTING ? std::ldexp(1.0, -L.dim_fw) : 1.0;   // multigrid.py:135-198: 2^-d P^T
                if (on_device && L.r_mode != MGB_R_EXPLICIT) {
                    CU(dev::transpose_scaled(h->stream, L.P.nrows, L.P.ncols, L.P.nnz, L.P.rowptr, L.P.cols, L.P.vals, scale,
                                             &L.R.rowptr, &L.R.cols, &L.R.vals));
                    L.R.nrows = L.P.ncols; L.R.ncols = L.P.nrows; L.R.nnz = L.P.nnz;
                    std::vector<int64_t> ip;
                    int64_t in[2];
                    TRY(fetch_rowptr(h, L.R, ip));
                    if (xo_self >= 0) TRY(interior_rows(h, L.R, xo_self, in));
                    TRY(finish_csr(h, L.R, ip, {}, xo_self >= 0 ? in : nullptr));
                } else {
                    if (L.r_mode != MGB_R_EXPLICIT) transpose_scaled(L.P_host, scale, L.R_host);
                    TRY(upload_csr(h, L.R_host, L.R, {}, xo_self));
                }
            }
        }
        const size_t np = n + (size_t)L.n_ghost + 16;   // [owned | ghost | tail padding for 16-byte bulk copies]
        if (!L.vec_in_arena) { TRY(dev_alloc(h, &L.v, np)); TRY(dev_alloc(h, &L.vtmp, np)); }     // (else: inside the exported arena)
        TRY(dev_alloc(h, &L.f, np));
        TRY(dev_alloc(h, &L.r, np)); TRY(dev_alloc(h, &L.g, np));
        for (double* p : {L.v, L.vtmp, L.f, L.r, L.g}) CU(cudaMemsetAsync(p, 0, np * sizeof(double), h->stream));
    }
    if (!h->levels[h->coarsest].stub) {   // dense inverse of the coarsest matrix: replaces spsolve (multigrid.py:239)
        Level& C = h->levels[h->coarsest];
        if (C.n > 8192) return fail(h, MGB_ERR_UNSUPPORTED, "coarsest level has %lld rows; the dense coarse solver is limited to 8192 (add levels)", (long long)C.n);
        if (!dense_inverse(C.A_host, h->coarse_inv_host)) return fail(h, MGB_ERR_SINGULAR, "coarsest matrix is singular");
        TRY(dev_upload(h, &h->coarse_inv, h->coarse_inv_host.data(), h->coarse_inv_host.size()));
    }
    CU(cudaStreamSynchronize(h->stream));
    for (auto& kv : h->levels) {      // host copies are no longer needed (artefacts are read back from the device)
        Level& L = kv.second;
        L.A_host = HostCsr(); L.P_host = HostCsr(); L.R_host = HostCsr();
        std::vector<int32_t>().swap(L.inj_host);
    }
    // levels whose halo exchange is fused into the kernels (HaloFuse): slab-like partitions (<= 2 neighbours, each send list one
    // run of rows), peer memory mapped, and every kernel that writes an iterate on the level able to send (hot-row smoother,
    // anchored-pattern prolongation, the zero-guess sweep)
    for (auto& kv : h->levels) {
        Level& L = kv.second;
        L.fuse_ok = false;
        if (!h->fuse_halo || !h->dist || L.stub || L.gathered || L.peers.empty() || L.peers.size() > 2) continue;
        if (!L.p2p_ready || !h->p2p_enable || !L.vec_in_arena || !L.fuse_counters) continue;
        if (h->smoother != MGB_SM_JACOBI_RJ || h->mu1 < 1 || h->stream_cfg <= 0) continue;
        bool ok = true;
        for (size_t p = 0; p < L.peers.size(); ++p) ok = ok && L.send_a[p] >= 0 && L.send_cnt[p] > 0;
        ok = ok && kernel_takes_hf<EpiJacobiRJ>(L.RJ) && kernel_takes_hf<EpiResidual>(L.A);
        if (L.has_transfer) ok = ok && L.P.cd.mode == 4;
        L.fuse_ok = ok;
    }
    h->finalized = true;
    {   // one eager cycle on zero data: sets kernel attributes outside any graph capture and faults early if a kernel is broken
        const int64_t before = h->launches;
        TRY(enqueue_cycle(h, h->finest, false, nullptr));
        CU(cudaStreamSynchronize(h->stream));
        h->launches = before;
    }
    return MGB_OK;
}

int mgb_vcycle(mgb_handle* h, int top_level, double* v, const double* f, int mem, int ncycles, double* resnorm_hist)
{
    Level* T = nullptr;
    TRY(check_ready(h, top_level, &T));
    if (!v || !f || ncycles < 0) return fail(h, MGB_ERR_INVALID, "null vector or negative cycle count");
    TRY(copy_in(h, T->f, f, T->n, mem, T->perm));
    TRY(copy_in(h, T->v, v, T->n, mem, T->perm));
    TRY(cycles_on_buffers(h, top_level, ncycles, resnorm_hist));
    TRY(copy_out(h, v, T->v, T->n, mem, T->perm));
    if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream));
    return MGB_OK;
}

int mgb_set_rhs(mgb_handle* h, int level, const double* b, int mem)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    if (!b) return fail(h, MGB_ERR_INVALID, "null right-hand side");
    if (!L->b) TRY(dev_alloc(h, &L->b, (size_t)L->n + 16));
    TRY(copy_in(h, L->b, b, L->n, mem, L->perm));
    if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream));
    return MGB_OK;
}

int mgb_set_mass_matrix(mgb_handle* h, int level, int64_t n, int64_t nnz, const void* indptr, int indptr_bytes,
                        const int32_t* indices, const double* values)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    if (n != L->n) return fail(h, MGB_ERR_INVALID, "mass matrix has %lld rows, level %d has %lld", (long long)n, level, (long long)L->n);
    HostCsr M;
    std::string e = import_csr(M, n, n, nnz, indptr, indptr_bytes, indices, values);
    if (!e.empty()) return fail(h, MGB_ERR_INVALID, "mass matrix: %s", e.c_str());
    if (!L->perm_host.empty()) permute_csr(M, &L->perm_host, &L->perm_host);
    free_csr(L->M);
    TRY(upload_csr(h, M, L->M));
    CU(cudaStreamSynchronize(h->stream));
    return MGB_OK;
}

// FullMultiGrid (multigrid.py:271-307) on the device: nested iteration from the re-discretised right-hand sides,
// mu0 V-cycles per intermediate level, V-cycles on the finest level until the residual norm is <= tol (capped).
// Norm: sqrt(r^T M r) if a mass matrix was set for the finest level (the reference's L2(Omega) norm, multigrid.py:203-208),
// else the l2 norm.  Only the per-cycle scalar travels to the host.
int mgb_fmg(mgb_handle* h, int mu0, double tol, int max_cycles, double* v_out, int mem, int* cycles_done,
            double* resnorm_hist, int hist_capacity)
{
    Level* T = nullptr;
    if (!h) return MGB_ERR_INVALID;
    TRY(check_ready(h, h->finest, &T));
    if (h->dist) return fail(h, MGB_ERR_UNSUPPORTED, "the FMG driver is single-GPU in this version");
    if (mu0 < 0 || max_cycles < 0) return fail(h, MGB_ERR_INVALID, "negative cycle count");
    for (int l = h->coarsest; l <= h->finest; ++l)
        if (!h->levels[l].b) return fail(h, MGB_ERR_STATE, "mgb_set_rhs missing for level %d", l);
    Level& C0 = h->levels[h->coarsest];
    CU(cudaMemcpyAsync(C0.f, C0.b, sizeof(double) * (size_t)C0.n, cudaMemcpyDeviceToDevice, h->stream));
    TRY(coarse_apply(h, C0, C0.f, C0.v));                                         // multigrid.py:274-277
    int done = 0;
    h->fmg_err.clear();
    TRY(ensure_hist(h, 2));
    CU(cudaMemsetAsync(h->d_hist, 0, 2 * sizeof(double), h->stream));
    for (int l = h->coarsest + 1; l <= h->finest; ++l) {
        Level& L = h->levels[l];
        Level& C = h->levels[l - 1];
        CU(cudaMemsetAsync(L.v, 0, sizeof(double) * (size_t)L.n, h->stream));
        TRY(prolong_add(h, L, C.v, L.v, nullptr));                                // v_h = Interpolation(v_2h), multigrid.py:283-284
        CU(cudaMemcpyAsync(L.f, L.b, sizeof(double) * (size_t)L.n, cudaMemcpyDeviceToDevice, h->stream));
        if (l < h->finest) {
            for (int c = 0; c < mu0; ++c) TRY(run_cycle(h, l, c > 0));            // multigrid.py:305-306
            continue;
        }
        while (done < max_cycles) {                                               // multigrid.py:288-302 (with a cap)
            TRY(run_cycle(h, l, done > 0));
            ++done;
            // norm of L.r into d_hist[slot]: sqrt(r^T M r) with a mass matrix (multigrid.py:203-208), else the l2 norm
            auto norm_of_r = [&](int slot) -> int {
                if (L.M.present()) {
                    EpiStore epi{L.vtmp};
                    TRY(row_sums(h, MGB_K_SPMV, l, bytes_rowsum(L.M, 2.0 * (double)L.n), L.M, L.r, epi));
                    return launch(h, MGB_K_NORM, l, 16.0 * (double)L.n, [&] {
                        k_dot_partial<<<h->norm_blocks, 256, 0, h->stream>>>(L.n, L.r, L.vtmp, h->d_partial);
                        k_sumsq_final<<<1, 1024, 0, h->stream>>>(h->norm_blocks, h->d_partial, h->d_hist + slot, 1);
                    });
                }
                return norm2_device(h, L.n, L.r, h->d_hist + slot, l);
            };
            TRY(residual(h, L, L.v, L.f, L.r));                                   // multigrid.py:291
            TRY(norm_of_r(0));
            if (L.uex) {                                                          // multigrid.py:292-293: the error, every cycle
                TRY(launch(h, MGB_K_NORM, l, 24.0 * (double)L.n, [&] {
                    k_diff<<<(int)((L.n + 255) / 256), 256, 0, h->stream>>>(L.n, L.v, L.uex, L.r);
                }));
                TRY(norm_of_r(1));
            }
            double nrm2[2] = {0.0, 0.0};
            CU(cudaMemcpyAsync(nrm2, h->d_hist, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            const double nrm = nrm2[0];
            if (L.uex) h->fmg_err.push_back(nrm2[1]);
            if (resnorm_hist && done <= hist_capacity) resnorm_hist[done - 1] = nrm;
            if (nrm <= tol) break;                                                // multigrid.py:296
        }
    }
    if (cycles_done) *cycles_done = done;
    if (v_out) {
        TRY(copy_out(h, v_out, T->v, T->n, mem, T->perm));
        if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream));
    }
    return MGB_OK;
}

int mgb_set_exact_solution(mgb_handle* h, int level, const double* u_exact, int mem)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    if (!u_exact) { cudaFree(L->uex); L->uex = nullptr; return MGB_OK; }      // (null: forget it)
    if (!L->uex) TRY(dev_alloc(h, &L->uex, (size_t)L->n + 16));
    TRY(copy_in(h, L->uex, u_exact, L->n, mem, L->perm));
    if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream));
    return MGB_OK;
}

int mgb_fmg_error_history(mgb_handle* h, double* errnorm_hist, int capacity, int* count)
{
    if (!h || !count) return MGB_ERR_INVALID;
    *count = (int)h->fmg_err.size();
    if (errnorm_hist) for (int i = 0; i < *count && i < capacity; ++i) errnorm_hist[i] = h->fmg_err[(size_t)i];
    return MGB_OK;
}

int mgb_vcycle_resident(mgb_handle* h, int top_level, int ncycles, double* resnorm_hist)
{
    Level* T = nullptr;
    TRY(check_ready(h, top_level, &T));
    if (ncycles < 0) return fail(h, MGB_ERR_INVALID, "negative cycle count");
    return cycles_on_buffers(h, top_level, ncycles, resnorm_hist);
}

int mgb_vcycle_debug(mgb_handle* h, int top_level, double* v, const double* f, int mem, double* f2h, double* v2h, double* err_h)
{
    Level* T = nullptr;
    TRY(check_ready(h, top_level, &T));
    if (top_level == h->coarsest) return fail(h, MGB_ERR_INVALID, "the debug outputs need a level above the coarsest");
    if (!v || !f) return fail(h, MGB_ERR_INVALID, "null vector");
    Level& C = h->levels[top_level - 1];
    TRY(copy_in(h, T->f, f, T->n, mem, T->perm));
    TRY(copy_in(h, T->v, v, T->n, mem, T->perm));
    double* v2 = nullptr;
    // f2h must be saved before the coarse recursion overwrites nothing -- C.f is only written by the restriction
    TRY(enqueue_cycle(h, top_level, true, &v2));
    TRY(copy_out(h, v, T->v, T->n, mem, T->perm));
    if (f2h) TRY(copy_out(h, f2h, C.f, C.n, mem, C.perm));
    if (v2h) TRY(copy_out(h, v2h, v2, C.n, mem, C.perm));
    if (err_h) TRY(copy_out(h, err_h, T->r, T->n, mem, T->perm));
    if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream));
    return MGB_OK;
}

int mgb_spmv(mgb_handle* h, int level, const double* x, double* y, int mem)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    const bool stage = mem == MGB_MEM_HOST || h->numbered;      // (a caller numbering: vectors pass through the level buffers)
    BorrowGuard guard(h, stage ? MGB_MEM_HOST : mem);
    const double* xd = x; double* yd = y;
    if (stage) { TRY(copy_in(h, L->v, x, L->n, mem, L->perm)); xd = L->v; yd = L->r; }
    EpiStore epi{yd};
    TRY(row_sums(h, MGB_K_SPMV, level, bytes_rowsum(L->A, 2.0 * (double)L->n), L->A, xd, epi));
    if (stage) { TRY(copy_out(h, y, yd, L->n, mem, L->perm)); if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream)); }
    return MGB_OK;
}

int mgb_residual(mgb_handle* h, int level, const double* v, const double* f, double* r, int mem)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    const bool stage = mem == MGB_MEM_HOST || h->numbered;      // (a caller numbering: vectors pass through the level buffers)
    BorrowGuard guard(h, stage ? MGB_MEM_HOST : mem);
    const double *vd = v, *fd = f; double* rd = r;
    if (stage) { TRY(copy_in(h, L->v, v, L->n, mem, L->perm)); TRY(copy_in(h, L->f, f, L->n, mem, L->perm)); vd = L->v; fd = L->f; rd = L->r; }
    TRY(residual(h, *L, vd, fd, rd));
    if (stage) { TRY(copy_out(h, r, rd, L->n, mem, L->perm)); if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream)); }
    return MGB_OK;
}

int mgb_smooth(mgb_handle* h, int level, double* v, const double* f, int nsweeps, int mem)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    const bool stage = mem == MGB_MEM_HOST || h->numbered;      // (a caller numbering: vectors pass through the level buffers)
    BorrowGuard guard(h, stage ? MGB_MEM_HOST : mem);
    if (level == h->coarsest && h->smoother >= MGB_SM_GS_LEVEL) return fail(h, MGB_ERR_STATE, "no Gauss-Seidel operator on the coarsest level");
    double* vd = v; const double* fd = f;
    if (stage) { TRY(copy_in(h, L->v, v, L->n, mem, L->perm)); TRY(copy_in(h, L->f, f, L->n, mem, L->perm)); vd = L->v; fd = L->f; }
    double* cur = vd; double* oth = L->vtmp;
    bool g_valid = false;
    TRY(smooth(h, *L, cur, oth, fd, nsweeps, g_valid));
    if (cur != vd) CU(cudaMemcpyAsync(vd, cur, sizeof(double) * (size_t)L->n, cudaMemcpyDeviceToDevice, h->stream));
    if (stage) { TRY(copy_out(h, v, vd, L->n, mem, L->perm)); if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream)); }
    return MGB_OK;
}

int mgb_restrict(mgb_handle* h, int fine_level, const double* r_fine, double* f_coarse, int mem)
{
    Level* L = nullptr;
    TRY(check_ready(h, fine_level, &L));
    const bool stage = mem == MGB_MEM_HOST || h->numbered;      // (a caller numbering: vectors pass through the level buffers)
    BorrowGuard guard(h, stage ? MGB_MEM_HOST : mem);
    if (!L->has_transfer) return fail(h, MGB_ERR_INVALID, "level %d has no coarser neighbour", fine_level);
    Level& C = h->levels[fine_level - 1];
    const double* rd = r_fine; double* fd = f_coarse;
    if (stage) { TRY(copy_in(h, L->r, r_fine, L->n, mem, L->perm)); rd = L->r; fd = C.f; }
    TRY(restrict_to(h, *L, rd, fd));
    if (stage) { TRY(copy_out(h, f_coarse, fd, C.n, mem, C.perm)); if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream)); }
    return MGB_OK;
}

int mgb_prolong_add(mgb_handle* h, int fine_level, const double* e_coarse, double* v_fine, int mem)
{
    Level* L = nullptr;
    TRY(check_ready(h, fine_level, &L));
    const bool stage = mem == MGB_MEM_HOST || h->numbered;      // (a caller numbering: vectors pass through the level buffers)
    BorrowGuard guard(h, stage ? MGB_MEM_HOST : mem);
    if (!L->has_transfer) return fail(h, MGB_ERR_INVALID, "level %d has no coarser neighbour", fine_level);
    Level& C = h->levels[fine_level - 1];
    const double* ed = e_coarse; double* vd = v_fine;
    if (stage) { TRY(copy_in(h, C.v, e_coarse, C.n, mem, C.perm)); TRY(copy_in(h, L->v, v_fine, L->n, mem, L->perm)); ed = C.v; vd = L->v; }
    TRY(prolong_add(h, *L, ed, vd, nullptr));
    if (stage) { TRY(copy_out(h, v_fine, vd, L->n, mem, L->perm)); if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream)); }
    return MGB_OK;
}

int mgb_coarse_solve(mgb_handle* h, const double* f, double* u, int mem)
{
    Level* C = nullptr;
    if (!h) return MGB_ERR_INVALID;
    TRY(check_ready(h, h->coarsest, &C));
    const bool stage = mem == MGB_MEM_HOST || h->numbered;      // (a caller numbering: vectors pass through the level buffers)
    BorrowGuard guard(h, stage ? MGB_MEM_HOST : mem);
    const double* fd = f; double* ud = u;
    if (stage) { TRY(copy_in(h, C->f, f, C->n, mem, C->perm)); fd = C->f; ud = C->v; }
    TRY(coarse_apply(h, *C, fd, ud));
    if (stage) { TRY(copy_out(h, u, ud, C->n, mem, C->perm)); if (mem == MGB_MEM_HOST) CU(cudaStreamSynchronize(h->stream)); }
    return MGB_OK;
}

int mgb_norm2(mgb_handle* h, int64_t n, const double* x, int mem, double* out_host)
{
    if (!h || !x || !out_host || n < 0) return MGB_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    const double* xd = x;
    double* tmp = nullptr;
    if (mem == MGB_MEM_HOST) { TRY(dev_alloc(h, &tmp, (size_t)n)); TRY(copy_in(h, tmp, x, n, mem)); xd = tmp; }
    TRY(ensure_hist(h, 1));
    TRY(norm2_device(h, n, xd, h->d_hist, -1));
    CU(cudaMemcpyAsync(out_host, h->d_hist, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(tmp);
    return MGB_OK;
}

int mgb_get_artifact(mgb_handle* h, int level, int kind, void* out, int64_t capacity_bytes, int64_t* size_bytes)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    const void* src = nullptr; bool on_device = true; int64_t bytes = 0;
    int32_t code_info[4] = {0, 0, 0, 0};
    auto host_vec = [&](const std::vector<int32_t>& v) { src = v.data(); bytes = (int64_t)v.size() * 4; on_device = false; };
    switch (kind) {
        case MGB_ART_RJ_INDPTR: src = L->RJ.rowptr; bytes = (L->RJ.nrows + 1) * 4; break;
        case MGB_ART_RJ_INDICES: src = L->RJ.cols; bytes = L->RJ.nnz * 4; break;
        case MGB_ART_RJ_VALUES: src = L->RJ.vals; bytes = L->RJ.nnz * 8; break;
        case MGB_ART_DINV: src = L->dinv; bytes = L->n * 8; break;
        case MGB_ART_LEVEL_OF_ROW: host_vec(L->lev_of_row); break;
        case MGB_ART_LEVEL_ORDER: host_vec(L->lev_order); break;
        case MGB_ART_LEVEL_OFFSETS: host_vec(L->lev_off); break;
        case MGB_ART_COLOUR_OF_ROW: host_vec(L->col_of_row); break;
        case MGB_ART_COLOUR_ORDER: host_vec(L->col_order); break;
        case MGB_ART_COLOUR_OFFSETS: host_vec(L->col_off); break;
        case MGB_ART_R_INDPTR: src = L->R.rowptr; bytes = L->R.present() ? (L->R.nrows + 1) * 4 : 0; break;
        case MGB_ART_R_INDICES: src = L->R.cols; bytes = L->R.nnz * 4; break;
        case MGB_ART_R_VALUES: src = L->R.vals; bytes = L->R.nnz * 8; break;
        case MGB_ART_A_INDPTR: src = L->A.rowptr; bytes = L->A.present() ? (L->A.nrows + 1) * 4 : 0; break;
        case MGB_ART_A_INDICES: src = L->A.cols; bytes = L->A.nnz * 4; break;
        case MGB_ART_A_VALUES: src = L->A.vals; bytes = L->A.nnz * 8; break;
        case MGB_ART_P_INDPTR: src = L->P.rowptr; bytes = L->P.present() ? (L->P.nrows + 1) * 4 : 0; break;
        case MGB_ART_P_INDICES: src = L->P.cols; bytes = L->P.nnz * 4; break;
        case MGB_ART_P_VALUES: src = L->P.vals; bytes = L->P.nnz * 8; break;
        case MGB_ART_INJECTION: src = L->inj; bytes = L->inj ? L->n_coarse * 4 : 0; break;
        case MGB_ART_COARSE_INVERSE:
            src = h->coarse_inv_host.data(); bytes = (int64_t)h->coarse_inv_host.size() * 8; on_device = false; break;
        default: {
            if (kind >= MGB_ART_CODE_ANCHOR(0) && kind <= MGB_ART_CODE_ANCHOR(3)) {         // mode 4: every row's anchor column
                const int op = kind - MGB_ART_CODE_ANCHOR(0);
                const DevCsr& D = op == 0 ? L->A : (op == 1 ? L->RJ : (op == 2 ? L->P : L->R));
                src = D.cd.anchor; bytes = D.cd.mode == 4 ? D.nrows * 4 : 0;
                break;
            }
            if (kind < MGB_ART_CODE(0, 0) || kind > MGB_ART_CODE(3, 3)) return fail(h, MGB_ERR_INVALID, "unknown artefact kind %d", kind);
            const int op = (kind - MGB_ART_CODE(0, 0)) / 4, part = (kind - MGB_ART_CODE(0, 0)) % 4;
            const DevCsr& D = op == 0 ? L->A : (op == 1 ? L->RJ : (op == 2 ? L->P : L->R));
            const Coded& cd = D.cd;
            const int64_t ncodes = cd.mode >= 3 ? D.nrows : (cd.mode ? D.nnz : 0);
            const int64_t tab = cd.mode >= 3 ? cd.npent : (cd.mode ? 256 : 0);
            code_info[0] = cd.mode; code_info[1] = cd.ndict; code_info[2] = (int32_t)tab; code_info[3] = (int32_t)ncodes;
            if (part == 0) { src = code_info; bytes = sizeof code_info; on_device = false; }
            else if (part == 1) { src = cd.codes; bytes = ncodes; }
            else if (part == 2) { src = cd.dict; bytes = tab * (int64_t)sizeof(DictEnt); }
            else { src = cd.phead; bytes = cd.mode >= 3 ? 256 * (int64_t)sizeof(int2) : 0; }
            break;
        }
    }
    if (size_bytes) *size_bytes = bytes;
    if (!out) return MGB_OK;
    if (capacity_bytes < bytes) return fail(h, MGB_ERR_INVALID, "artefact needs %lld bytes, buffer has %lld", (long long)bytes, (long long)capacity_bytes);
    if (bytes == 0) return MGB_OK;
    if (on_device) {
        CU(cudaMemcpyAsync(out, src, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    } else {
        std::memcpy(out, src, (size_t)bytes);
    }
    return MGB_OK;
}

int mgb_level_buffer(mgb_handle* h, int level, int which, void** device_ptr, int64_t* n)
{
    Level* L = nullptr;
    TRY(check_ready(h, level, &L));
    if (!device_ptr) return MGB_ERR_INVALID;
    switch (which) {
        case MGB_BUF_V: *device_ptr = L->v; break;
        case MGB_BUF_F: *device_ptr = L->f; break;
        case MGB_BUF_R: *device_ptr = L->r; break;
        default: return fail(h, MGB_ERR_INVALID, "unknown buffer %d", which);
    }
    if (n) *n = L->n;
    return MGB_OK;
}

int mgb_get_stream(mgb_handle* h, void** cuda_stream)
{
    if (!h || !cuda_stream) return MGB_ERR_INVALID;
    *cuda_stream = (void*)h->stream;
    return MGB_OK;
}

int mgb_synchronize(mgb_handle* h)
{
    if (!h) return MGB_ERR_INVALID;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return MGB_OK;
}

int mgb_launch_count(mgb_handle* h, int64_t* kernels_launched)
{
    if (!h || !kernels_launched) return MGB_ERR_INVALID;
    *kernels_launched = h->launches;
    return MGB_OK;
}

int mgb_profile_begin(mgb_handle* h)
{
    if (!h) return MGB_ERR_INVALID;
    CU(cudaStreamSynchronize(h->stream));
    h->prof = true;
    h->prof_records.clear();
    return MGB_OK;
}

int mgb_profile_end(mgb_handle* h)
{
    if (!h) return MGB_ERR_INVALID;
    CU(cudaStreamSynchronize(h->stream));
    for (auto& pe : h->prof_events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pe.e0, pe.e1);
        mgb_profile_record& r = h->prof_records[{pe.kind, pe.level}];
        r.kind = pe.kind; r.level = pe.level; r.launches++; r.total_ms += ms; r.bytes = pe.bytes; r.moved_bytes = pe.moved;
        cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1);
    }
    h->prof_events.clear();
    h->prof = false;
    return MGB_OK;
}

int mgb_profile_get(mgb_handle* h, mgb_profile_record* out, int capacity, int* count)
{
    if (!h || !count) return MGB_ERR_INVALID;
    *count = (int)h->prof_records.size();
    if (!out) return MGB_OK;
    int i = 0;
    for (auto& kv : h->prof_records) { if (i >= capacity) break; out[i++] = kv.second; }
    return MGB_OK;
}

static int vcycle_bytes_impl(mgb_handle* h, int top_level, double* bytes, bool moved);
int mgb_vcycle_bytes(mgb_handle* h, int top_level, double* bytes) { return vcycle_bytes_impl(h, top_level, bytes, false); }
int mgb_vcycle_bytes_moved(mgb_handle* h, int top_level, double* bytes) { return vcycle_bytes_impl(h, top_level, bytes, true); }

static int vcycle_bytes_impl(mgb_handle* h, int top_level, double* bytes, bool moved)
{
    Level* T = nullptr;
    if (!h) return MGB_ERR_INVALID;
    TRY(check_ready(h, top_level, &T));
    if (!bytes) return MGB_ERR_INVALID;
    auto bytes_rowsum = [&](const DevCsr& D, double vec_terms) {
        return ::bytes_rowsum(D, vec_terms) - (moved && h->stream_cfg > 0 ? coded_saving(D) : 0.0);
    };
    double b = 0.0;
    for (int l = h->coarsest + 1; l <= top_level; ++l) {
        Level& L = h->levels[l];
        const double n = (double)L.n, nc = (double)L.n_coarse;
        const int sweeps = h->mu1 + h->mu2;
        if (h->smoother == MGB_SM_JACOBI_RJ) {
            b += sweeps * bytes_rowsum(L.RJ, 3 * n);
            if (moved && L.s2 && h->fuse_sweeps && h->stream_cfg > 0)      // k_hotrow2: the second sweep of a pair reads y and g from L2
                b -= 16.0 * n * ((l == top_level ? h->mu1 / 2 : std::max(h->mu1 - 1, 0) / 2) + h->mu2 / 2);
        } else if (h->smoother == MGB_SM_JACOBI_A) b += sweeps * bytes_rowsum(L.A, 4 * n);
        else b += sweeps * (12.0 * (double)L.G.nnz + 8.0 * n + 24.0 * n);
        b += bytes_rowsum(L.A, 3 * n);
        b += L.r_mode == MGB_R_INJECTION ? 20.0 * nc : bytes_rowsum(L.R, n + nc);
        b += bytes_rowsum(L.P, nc + 2 * n);
    }
    const double n0 = (double)h->levels[h->coarsest].n;
    b += 8.0 * n0 * n0 + 16.0 * n0;
    *bytes = b;
    return MGB_OK;
}

int mgb_describe(mgb_handle* h, char* out, int64_t capacity)
{
    if (!h || !out || capacity <= 0) return MGB_ERR_INVALID;
    std::string s;
    char buf[384];
    auto one = [&](const char* name, const DevCsr& D) {
        if (!D.present()) return;
        if (D.family == 1 && D.sdesc && D.cd.mode == 4) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d anchrow(coded mode=4: first column + one of %d row patterns, %d table entries; 256-row tiles) tiles=%d\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, D.cd.ndict, D.cd.npent, D.sntiles);
        else if (D.family == 1 && D.sdesc && D.cd.mode == 3 && D.hcfg > 0) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d hotrow(coded mode=3: %d row patterns, %d table entries, hot pattern %d of %d entries, %d patterns take the table walk; cfg=%d, %d x %d rows) tiles=%d\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, D.cd.ndict, D.cd.npent, D.cd.hotplan.hot, D.cd.hotplan.hotlen, D.cd.hot_slow, D.hcfg, hot_choice(D.hcfg).threads, hot_choice(D.hcfg).rpt, D.sntiles);
        else if (D.family == 1 && D.sdesc && D.cd.mode == 3 && D.wcfg > 0) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d rowwin(coded mode=3: %d row patterns, %d table entries, %d x-windows, hot pattern %d; cfg=%d, 256 x %d rows, %d stages) tiles=%d\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, D.cd.ndict, D.cd.npent, D.cd.win.ng, D.cd.win.hot, D.wcfg, win_choice(D.wcfg).rpt, win_choice(D.wcfg).stages, D.sntiles);
        else if (D.family == 1 && D.sdesc && D.cd.mode == 3) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d rowstream(coded mode=3: %d row patterns, %d table entries; cfg=%d, %d x %d rows, %d stages) tiles=%d\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, D.cd.ndict, D.cd.npent, D.ccfg, code_choice(D.ccfg).threads, code_choice(D.ccfg).rpt, code_choice(D.ccfg).stages, D.sntiles);
        else if (D.family == 1 && D.sdesc && D.cd.mode) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d rowstream(coded mode=%d: %d dictionary entries, %d values, %d offsets; cfg=%d, %d x %d rows, %d stages) tiles=%d\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, D.cd.mode, D.cd.ndict, D.cd.nvals, D.cd.ndeltas, D.ccfg, code_choice(D.ccfg).threads, code_choice(D.ccfg).rpt, code_choice(D.ccfg).stages, D.sntiles);
        else if (D.family == 1 && D.sdesc) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d stream(cfg=%d, %d x %d entries, %d stages) tiles=%d\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, D.scfg, stream_choice(D.scfg).threads, stream_choice(D.scfg).ept, stream_choice(D.scfg).stages, D.sntiles);
        else if (D.family == 1) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d tile(cap=%d) tiles=%d\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, tile_cap(D.iter), D.ntiles);
        else if (D.family == 3) snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d warp-per-row (sequential order)\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row);
        else snprintf(buf, sizeof buf, "  %-3s rows=%lld nnz=%lld max_row=%d subwarp(lanes=%d)\n", name, (long long)D.nrows, (long long)D.nnz, D.max_row, D.lpr);
        s += buf;
    };
    for (auto& kv : h->levels) {
        snprintf(buf, sizeof buf, "level %d n=%lld%s\n", kv.first, (long long)kv.second.n, kv.second.fuse_ok ? " (halo exchange fused into the kernels)" : "");
        s += buf;
        one("A", kv.second.A); one("RJ", kv.second.RJ); one("P", kv.second.P); one("R", kv.second.R); one("G", kv.second.G);
    }
    snprintf(out, (size_t)capacity, "%s", s.c_str());
    return MGB_OK;
}

}  // extern "C"
