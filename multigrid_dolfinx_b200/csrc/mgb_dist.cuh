// Multi-GPU data path of the engine (included by mgb_engine.cu inside its anonymous namespace): NCCL loaded at run
// time, halo exchange through NVLink peer memory (push / pull kernels) or ncclSend/ncclRecv, gather / broadcast of the
// first rank-0-only level, and row sums with the exchange folded in (row_sums_halo).
#pragma once

// ---- NCCL, loaded at run time (libnccl.so.2 of the process, i.e. the one torch already loaded) ---------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

const char* load_nccl()
{
    if (g_nccl.lib) return nullptr;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "libnccl.so.2 not found";
#define NCCL_SYM(name) *(void**)(&g_nccl.name) = dlsym(lib, "nccl" #name); if (!g_nccl.name) return "symbol nccl" #name " missing";
    NCCL_SYM(GetUniqueId) NCCL_SYM(CommInitRank) NCCL_SYM(CommDestroy) NCCL_SYM(Send) NCCL_SYM(Recv) NCCL_SYM(Broadcast)
    NCCL_SYM(AllReduce) NCCL_SYM(GroupStart) NCCL_SYM(GroupEnd) NCCL_SYM(GetErrorString)
#undef NCCL_SYM
    g_nccl.lib = lib;
    return nullptr;
}

#define NC(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess) return fail(h, MGB_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

__global__ void k_pack(int n, const int32_t* __restrict__ idx, const double* __restrict__ v, double* __restrict__ buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = v[idx[i]];
}
__global__ void k_sqrt_inplace(double* x) { *x = sqrt(*x); }

// Halo exchange of one level vector: pack the owned entries the neighbours need, one grouped
// ncclSend/ncclRecv per neighbour; the ghost section of `vec` (behind the n owned entries) is the receive buffer.
// ---- halo exchange through peer memory ------------------------------------------------------------------
// push: gather the owned entries each neighbour needs and store them straight into that neighbour's staging copy
// (NVLink peer stores), then -- last block only -- publish the new epoch in the neighbour's arrival flag.
// pull: wait until every neighbour's flag shows the epoch, then copy the staging copy into the ghost section of
// the vector.  Epochs live in device memory, so the pair replays correctly inside a CUDA graph.  Two staging
// copies alternate: a neighbour can only push epoch e after it has received this rank's epoch e-1, which this rank
// sent after finishing its pull of epoch e-2 -- so copy (e mod 2) is never overwritten while it is still read.
__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_p2p_push(P2PPlan pl, const int32_t* __restrict__ send_idx, const double* __restrict__ vec)
{
    const int total = pl.send_off[pl.npeers];
    const int par = (int)((pl.counters[0] + 1) & 1);          // all neighbours of a level share one epoch count
    const int stride = gridDim.x * blockDim.x;
    for (int k0 = blockIdx.x * blockDim.x + threadIdx.x; k0 < total; k0 += 4 * stride) {
        double val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int k = k0 + u * stride; if (k < total) val[u] = vec[send_idx[k]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + u * stride;
            if (k < total) {
                int p = 0;
                while (k >= pl.send_off[p + 1]) ++p;
                pl.rstage[p][par][k - pl.send_off[p]] = val[u];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long prev = atomicAdd(&pl.counters[32], 1ULL);
        if (prev == gridDim.x - 1) {                  // every block's stores are fenced: publish
            pl.counters[32] = 0;
            for (int p = 0; p < pl.npeers; ++p) {
                const unsigned long long e = pl.counters[p] + 1;
                pl.counters[p] = e;
                st_flag(pl.rflag[p], e);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_p2p_pull(P2PPlan pl, double* __restrict__ vec, int n_owned)
{
    if (threadIdx.x < pl.npeers) {
        const int p = threadIdx.x;
        const unsigned long long want = pl.counters[16 + p] + 1;
        const long long t0 = clock64();
        while (ld_flag(pl.flags + pl.peer_rank[p]) < want)
            if (clock64() - t0 > 60000000000LL) __trap();         // ~30 s (ranks may enter a cycle seconds apart: rank 0 alone sets up the
                                                                  // coarse levels): a lost neighbour must fault, never hang the GPU
    }
    __syncthreads();
    const int par = (int)((pl.counters[16] + 1) & 1);
    const double* src = pl.stage + (size_t)par * pl.n_ghost;
    const int stride = gridDim.x * blockDim.x;
    for (int k0 = blockIdx.x * blockDim.x + threadIdx.x; k0 < pl.n_ghost; k0 += 4 * stride) {
        double val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int k = k0 + u * stride; if (k < pl.n_ghost) val[u] = __ldcg(src + k); }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int k = k0 + u * stride; if (k < pl.n_ghost) vec[n_owned + k] = val[u]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long prev = atomicAdd(&pl.counters[33], 1ULL);
        if (prev == gridDim.x - 1) {
            pl.counters[33] = 0;
            for (int p = 0; p < pl.npeers; ++p) pl.counters[16 + p] += 1;
        }
    }
}

void p2p_push(mgb_handle* h, Level& L, const double* vec, cudaStream_t st)
{
    const int wide = 4 * h->sm_count;              // NVLink stores and the local unpack need many SMs to reach bandwidth
    const int gb = std::max(1, std::min(wide, ((int)L.send_total + 1023) / 1024));
    k_p2p_push<<<gb, 256, 0, st>>>(L.p2p, L.send_idx, vec);
}
void p2p_pull(mgb_handle* h, Level& L, double* vec, cudaStream_t st)
{
    const int wide = 4 * h->sm_count;
    const int gp = std::max(1, std::min(wide, ((int)L.n_ghost + 1023) / 1024));
    k_p2p_pull<<<gp, 256, 0, st>>>(L.p2p, vec, (int)L.n);
}

int exchange_on(mgb_handle* h, Level& L, double* vec, cudaStream_t st)
{
    if (L.p2p_ready && h->p2p_enable) {
        p2p_push(h, L, vec, st);
        p2p_pull(h, L, vec, st);
        h->launches += 1;
        return MGB_OK;
    }
    if (L.send_total > 0) k_pack<<<(int)((L.send_total + 255) / 256), 256, 0, st>>>((int)L.send_total, L.send_idx, vec, L.send_buf);
    ncclResult_t r = g_nccl.GroupStart();
    int64_t so = 0, ro = 0;
    for (size_t p = 0; p < L.peers.size() && r == ncclSuccess; ++p) {
        if (L.send_cnt[p] > 0) r = g_nccl.Send(L.send_buf + so, (size_t)L.send_cnt[p], ncclDouble, L.peers[p], h->comm, st);
        if (r == ncclSuccess && L.recv_cnt[p] > 0) r = g_nccl.Recv(vec + L.n + ro, (size_t)L.recv_cnt[p], ncclDouble, L.peers[p], h->comm, st);
        so += L.send_cnt[p]; ro += L.recv_cnt[p];
    }
    ncclResult_t e = g_nccl.GroupEnd();
    if (r == ncclSuccess) r = e;
    if (r != ncclSuccess) return fail(h, MGB_ERR_COMM, "halo exchange on level %d: %s", L.level, g_nccl.GetErrorString(r));
    return MGB_OK;
}

int exchange(mgb_handle* h, Level& L, double* vec)
{
    if (!h->dist || L.peers.empty()) return MGB_OK;
    int rc = MGB_OK;
    TRY(launch(h, MGB_K_HALO, L.level, 16.0 * (double)L.send_total, [&] { rc = exchange_on(h, L, vec, h->stream); }));
    return rc;
}

// Row sums of a sharded operator whose input vector x lives on level XL: exchange XL's ghosts, overlapped with the
// interior tiles when the operator has an interior / boundary split.
// OL / out: when the kernel writes an iterate of level OL into buffer out, and OL's exchange is fused, the rows OL's neighbours
// hold as ghosts are stored into their copies by the kernel itself (HaloFuse).
template <class Epi>
int row_sums_halo(mgb_handle* h, int kind, int level, double bytes, const DevCsr& D, Level& XL, double* x, const Epi& epi,
                  const int4* sub_desc = nullptr, int sub_int = 0, int sub_bnd = 0, double sub_moved = -1.0,
                  Level* OL = nullptr, const double* out = nullptr)
{
    // (only an ITERATE of XL has its ghosts kept valid by the kernels: any other vector, e.g. a residual about to be restricted,
    // still needs the ordinary exchange)
    const bool xf = XL.fuse_ok && (x == XL.v || x == XL.vtmp);
    if (h->in_cycle && !sub_desc && kernel_takes_hf<Epi>(D) && D.sdesc && h->stream_cfg > 0 && h->allow_stream && (xf || (OL && OL->fuse_ok))) {
        // fused exchange: the kernel waits for the ghosts of x itself and / or sends its boundary rows itself
        if (h->dist && !XL.peers.empty() && !xf) TRY(exchange(h, XL, x));
        h->hf_cur = make_hf(h, xf ? &XL : nullptr, &D, OL, out);
        const int rc = row_sums(h, kind, level, bytes, D, x, epi);
        h->hf_cur = no_hf();
        return rc;
    }
    if (OL && OL->fuse_ok && h->in_cycle) return fail(h, MGB_ERR_STATE, "level %d: fused exchange, but this kernel cannot send its boundary rows", OL->level);
    // bytes actually streamed (profile records): the tile subset's own figure, else CSR bytes minus the coding's saving
    const bool streamable = Epi::CONTIG && D.family == 1 && D.sdesc && h->stream_cfg > 0 && h->allow_stream;
    const double moved = sub_desc ? sub_moved : (streamable ? bytes - coded_saving(D) : -1.0);
    const bool need = h->dist && !XL.peers.empty();
    // overlap mode 2 (peer-memory exchange only): push -> interior tiles -> pull -> boundary tiles, all on one stream.
    // The neighbours' data travels while the interior rows (which read no ghost entry) are being summed, so the pull
    // finds its flags already set; no second stream and no co-residency with another kernel is needed.
    if (need && h->overlap == 2 && !h->prof && XL.p2p_ready && h->p2p_enable && D.split && D.sdesc && h->stream_cfg > 0 &&
        h->allow_stream && D.family == 1) {
        TRY(launch(h, kind, level, bytes, [&] {
            p2p_push(h, XL, x, h->stream);
            if (sub_desc) launch_stream<Epi>(h, D, x, epi, sub_desc, sub_int);
            else launch_stream<Epi>(h, D, x, epi, D.sdesc + D.t_int0, D.t_int1 - D.t_int0);
            p2p_pull(h, XL, x, h->stream);
            if (sub_desc) launch_stream<Epi>(h, D, x, epi, sub_desc + sub_int, sub_bnd);
            else launch_stream<Epi>(h, D, x, epi, D.sdesc_bnd, D.n_bnd);
        }, moved));
        h->launches += 3;
        return MGB_OK;
    }
    const bool can_overlap = need && h->overlap == 1 && !h->prof && D.split && D.sdesc && h->stream_cfg > 0 && h->allow_stream && D.family == 1;
    if (!can_overlap) {
        if (need) TRY(exchange(h, XL, x));
        if (sub_desc) {
            return launch(h, kind, level, bytes, [&] { launch_stream<Epi>(h, D, x, epi, sub_desc, sub_int + sub_bnd); }, moved);
        }
        return row_sums(h, kind, level, bytes, D, x, epi);
    }
    int rc = MGB_OK;
    TRY(launch(h, kind, level, bytes, [&] {
        cudaEventRecord(h->ev_fork, h->stream);
        cudaStreamWaitEvent(h->comm_stream, h->ev_fork, 0);
        rc = exchange_on(h, XL, x, h->comm_stream);
        cudaEventRecord(h->ev_join, h->comm_stream);
        if (sub_desc) launch_stream<Epi>(h, D, x, epi, sub_desc, sub_int, true);
        else launch_stream<Epi>(h, D, x, epi, D.sdesc + D.t_int0, D.t_int1 - D.t_int0, true);
        cudaStreamWaitEvent(h->stream, h->ev_join, 0);
        if (sub_desc) launch_stream<Epi>(h, D, x, epi, sub_desc + sub_int, sub_bnd);
        else launch_stream<Epi>(h, D, x, epi, D.sdesc_bnd, D.n_bnd);
    }));
    h->launches += 2;
    return rc;
}

// slices of the gathered level's right-hand side -> rank 0
int gather_to_root(mgb_handle* h, Level& C, double* f)
{
    int rc = MGB_OK;
    TRY(launch(h, MGB_K_HALO, C.level, 8.0 * (double)C.n, [&] {
        ncclResult_t r = g_nccl.GroupStart();
        if (h->rank == 0) {
            for (int q = 1; q < h->world && r == ncclSuccess; ++q) {
                const int64_t cnt = C.gather_off[q + 1] - C.gather_off[q];
                if (cnt > 0) r = g_nccl.Recv(f + C.gather_off[q], (size_t)cnt, ncclDouble, q, h->comm, h->stream);
            }
        } else if (C.my_cnt > 0) {
            r = g_nccl.Send(f + C.my_off, (size_t)C.my_cnt, ncclDouble, 0, h->comm, h->stream);
        }
        ncclResult_t e = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = e;
        if (r != ncclSuccess) rc = fail(h, MGB_ERR_COMM, "gather to rank 0: %s", g_nccl.GetErrorString(r));
    }));
    return rc;
}

// coarse-grid correction of the gathered level: rank 0 -> everybody
int bcast_from_root(mgb_handle* h, Level& C, const double* src, double* dst)
{
    int rc = MGB_OK;
    TRY(launch(h, MGB_K_HALO, C.level, 8.0 * (double)C.n, [&] {
        ncclResult_t r = g_nccl.Broadcast(src, dst, (size_t)C.n, ncclDouble, 0, h->comm, h->stream);
        if (r != ncclSuccess) rc = fail(h, MGB_ERR_COMM, "broadcast from rank 0: %s", g_nccl.GetErrorString(r));
    }));
    return rc;
}

