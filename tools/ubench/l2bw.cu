// Micro-benchmark: how fast can the SMs pull bytes out of L2 (buffer that fits the 126 MB L2) and out of HBM
// (buffer far larger than L2), with plain 128-bit loads and with 1-D bulk copies (cp.async.bulk -> shared memory).
// The row-window kernels move ~1.3-2x more bytes L2 -> SM than HBM -> L2, so this ratio bounds them.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/l2bw tools/ubench/l2bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_ldg(const int4* __restrict__ p, size_t n16, int reps, int4* sink)
{
    int4 acc = make_int4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n16; i += 4 * stride) {
            int4 a = __ldcg(p + i), b = __ldcg(p + i + stride), c = __ldcg(p + i + 2 * stride), d = __ldcg(p + i + 3 * stride);
            acc.x ^= a.x ^ b.x ^ c.x ^ d.x; acc.y ^= a.y ^ b.y ^ c.y ^ d.y; acc.z ^= a.z ^ b.z ^ c.z ^ d.z; acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
        }
    if (acc.x == 0x12345678) *sink = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one producer thread per CTA issues bulk copies of CHUNK bytes into a ring of STAGES; nobody reads the data (pure transport)
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(128) k_bulk(const unsigned char* __restrict__ p, size_t bytes, int reps)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    unsigned char* buf = sm + 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = bytes / CHUNK;
    size_t it = 0;
    for (int r = 0; r < reps; ++r)
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
            const int s = (int)(it % STAGES);
            if (it >= STAGES) {
                const uint32_t par = (uint32_t)((it / STAGES - 1) & 1);
                uint32_t ok = 0; long long t0 = clock64();
                while (!ok && clock64() - t0 < 2000000000LL) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(s32(bar + s)), "r"(par) : "memory");
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + s)), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(buf + (size_t)s * CHUNK)), "l"(p + c * CHUNK), "r"(CHUNK), "r"(s32(bar + s)) : "memory");
        }
    const size_t total = it;
    for (size_t j = (total > STAGES ? total - STAGES : 0); j < total; ++j) {       // drain
        const int s = (int)(j % STAGES);
        const uint32_t par = (uint32_t)((j / STAGES) & 1);
        uint32_t ok = 0; long long t0 = clock64();
        while (!ok && clock64() - t0 < 2000000000LL) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(s32(bar + s)), "r"(par) : "memory");
    }
}

static float time_it(void (*fn)(void*), void* ctx)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    fn(ctx); cudaDeviceSynchronize();
    cudaEventRecord(a); fn(ctx); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

struct Ctx { const unsigned char* p; size_t bytes; int reps; int grid; int4* sink; };

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    const size_t big = (size_t)4 << 30;
    unsigned char* d; cudaMalloc(&d, big); cudaMemset(d, 1, big);
    int4* sink; cudaMalloc(&sink, 64);
    printf("{\"device\": \"%s\", \"sms\": %d, \"rows\": [\n", pr.name, sms);
    const size_t sizes[] = {(size_t)24 << 20, (size_t)48 << 20, (size_t)96 << 20, big};
    bool first = true;
    for (size_t sz : sizes) {
        const int reps = sz == big ? 1 : (int)(((size_t)4 << 30) / sz);
        for (int cps : {2, 4, 8}) {
            Ctx c{d, sz, reps, sms * cps, sink};
            float ms = time_it([](void* v) { Ctx* c = (Ctx*)v; k_ldg<<<c->grid, 256>>>((const int4*)c->p, c->bytes / 16, c->reps, c->sink); }, &c);
            printf("%s{\"kind\": \"ldg128\", \"MB\": %zu, \"ctas_per_sm\": %d, \"GBs\": %.1f}", first ? "" : ",\n", sz >> 20, cps, (double)sz * reps / ms / 1e6);
            first = false;
        }
        for (int cps : {2, 4}) {
            Ctx c{d, sz, reps, sms * cps, sink};
            cudaFuncSetAttribute(k_bulk<4096, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 4 * 4096);
            float ms = time_it([](void* v) { Ctx* c = (Ctx*)v; k_bulk<4096, 4><<<c->grid, 128, 128 + 4 * 4096>>>(c->p, c->bytes, c->reps); }, &c);
            printf(",\n{\"kind\": \"bulk4k_x4\", \"MB\": %zu, \"ctas_per_sm\": %d, \"GBs\": %.1f}", sz >> 20, cps, (double)sz * reps / ms / 1e6);
            cudaFuncSetAttribute(k_bulk<16384, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 3 * 16384);
            ms = time_it([](void* v) { Ctx* c = (Ctx*)v; k_bulk<16384, 3><<<c->grid, 128, 128 + 3 * 16384>>>(c->p, c->bytes, c->reps); }, &c);
            printf(",\n{\"kind\": \"bulk16k_x3\", \"MB\": %zu, \"ctas_per_sm\": %d, \"GBs\": %.1f}", sz >> 20, cps, (double)sz * reps / ms / 1e6);
        }
    }
    printf("\n], \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
