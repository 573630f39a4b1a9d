// Dev probe: how fast can SM-issued loads pull 1 KB feature cells (NHWC fp32, 256 channels) out of MAPPED PINNED HOST
// memory, dense and with the touched-cell bitmap of a config-2 step (~59 % of the cells, 10 x 10-cell RoI rectangles),
// against cudaMemcpyAsync of the whole buffer.  Answers VERDICT r1 item 6c (ship only the touched cells over PCIe).
// usage: zc_probe [cells_per_warp_in_flight]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float4 ld_nc(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// one warp per bitmap word (32 cells); U cells in flight per warp (2 x 16-byte loads per lane and cell)
template <int U>
__global__ void k_fetch(const float4* __restrict__ src, float4* __restrict__ dst, const unsigned* __restrict__ bits, int n_words) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int wd = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wd < n_words; wd += warps) {
        unsigned m = bits[wd];
        while (m) {
            long long cell[U];
            float4 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (m) { const int k = __ffs(m) - 1; m &= m - 1; cell[u] = (long long)wd * 32 + k; }
                else cell[u] = -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (cell[u] >= 0) { a[u] = ld_nc(src + cell[u] * 64 + lane); b[u] = ld_nc(src + cell[u] * 64 + 32 + lane); }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (cell[u] >= 0) { dst[cell[u] * 64 + lane] = a[u]; dst[cell[u] * 64 + 32 + lane] = b[u]; }
        }
    }
}

// TMA variant: one elected thread per CTA moves the marked cells of a bitmap word with bulk async copies, host -> shared
// memory (mbarrier completion), shared memory -> device; two word buffers so that the loads of one word overlap the stores
// of the previous one
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) k_fetch_tma(const char* __restrict__ src, char* __restrict__ dst, const unsigned* __restrict__ bits,
                                                  int n_words) {
    extern __shared__ __align__(128) char buf[];                  // [2][32][1024]
    __shared__ uint64_t bar[2];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned phase[2] = {0, 0};
    int it = 0;
    for (int wd = blockIdx.x; wd < n_words; wd += gridDim.x) {
        unsigned m = bits[wd];
        if (!m) continue;
        const int b = it & 1;
        ++it;
        char* slot = buf + b * 32 * 1024;
        // the stores that last read this buffer must have finished reading it (at most one older group may be pending)
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        const int n = __popc(m);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(n * 1024) : "memory");
        unsigned mm = m;
        for (int k = 0; mm; ++k) {
            const int c = __ffs(mm) - 1; mm &= mm - 1;
            const char* s = src + ((long long)wd * 32 + c) * 1024;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(slot + k * 1024)),
                         "l"(s), "r"(1024), "r"(smem_u32(&bar[b])) : "memory");
        }
        unsigned ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(phase[b]) : "memory");
        phase[b] ^= 1;
        mm = m;
        for (int k = 0; mm; ++k) {
            const int c = __ffs(mm) - 1; mm &= mm - 1;
            char* d = dst + ((long long)wd * 32 + c) * 1024;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(smem_u32(slot + k * 1024)), "r"(1024) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const int B = 8, H = 200, W = 336;
    const long long cells = (long long)B * H * W;          // level 0 of the config-2 pyramid, 8 images
    const long long bytes = cells * 1024;
    float* h;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(h, 1, bytes);
    float *d, *hd;
    CK(cudaMalloc(&d, bytes));
    CK(cudaHostGetDevicePointer(&hd, h, 0));
    const int n_words = (int)((cells + 31) / 32);
    std::vector<unsigned> dense(n_words, 0xffffffffu), sparse(n_words, 0u);
    srand(7);
    for (int b = 0; b < B; ++b)
        for (int r = 0; r < 482; ++r) {                     // 482 of 512 sampled RoIs sit on level 0
            const int rw = 6 + rand() % 9, rh = 6 + rand() % 9;
            const int x0 = rand() % (W - rw), y0 = rand() % (H - rh);
            for (int y = y0; y < y0 + rh; ++y)
                for (int x = x0; x < x0 + rw; ++x) {
                    const long long c = ((long long)b * H + y) * W + x;
                    sparse[c >> 5] |= 1u << (c & 31);
                }
        }
    long long set = 0;
    for (unsigned v : sparse) set += __builtin_popcount(v);
    printf("cells %lld (%.1f MB), touched %lld (%.1f %%, %.1f MB)\n", cells, bytes / 1e6, set, 100.0 * set / cells, set * 1024 / 1e6);
    unsigned *d_dense, *d_sparse;
    CK(cudaMalloc(&d_dense, n_words * 4)); CK(cudaMalloc(&d_sparse, n_words * 4));
    CK(cudaMemcpy(d_dense, dense.data(), n_words * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sparse, sparse.data(), n_words * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, double moved, auto fn) {
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            CK(cudaEventRecord(e0)); fn(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-44s %8.3f ms  %6.1f GB/s\n", name, best, moved / best / 1e6);
    };
    time("cudaMemcpyAsync H2D, whole buffer", (double)bytes, [&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); });
    for (int ctas : {148, 296, 592, 1184})
        for (int threads : {256, 1024}) {
            char name[128];
            snprintf(name, sizeof name, "zero-copy dense  U=4 %4d CTAs x %4d", ctas, threads);
            time(name, (double)bytes, [&] { k_fetch<4><<<ctas, threads>>>((const float4*)hd, (float4*)d, d_dense, n_words); });
            snprintf(name, sizeof name, "zero-copy sparse U=4 %4d CTAs x %4d", ctas, threads);
            time(name, set * 1024.0, [&] { k_fetch<4><<<ctas, threads>>>((const float4*)hd, (float4*)d, d_sparse, n_words); });
        }
    CK(cudaFuncSetAttribute(k_fetch_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int ctas : {148, 296, 444})
        for (int pass = 0; pass < 2; ++pass) {
            char name[128];
            snprintf(name, sizeof name, "TMA bulk %s %4d CTAs (1 thread, 2 x 32 KB)", pass ? "sparse" : "dense ", ctas);
            time(name, pass ? set * 1024.0 : (double)bytes, [&] { k_fetch_tma<<<ctas, 32, 65536>>>((const char*)hd, (char*)d, pass ? d_sparse : d_dense, n_words); });
        }
    time("zero-copy sparse U=1  592 CTAs x 1024", set * 1024.0, [&] { k_fetch<1><<<592, 1024>>>((const float4*)hd, (float4*)d, d_sparse, n_words); });
    time("zero-copy sparse U=2  592 CTAs x 1024", set * 1024.0, [&] { k_fetch<2><<<592, 1024>>>((const float4*)hd, (float4*)d, d_sparse, n_words); });
    time("zero-copy sparse U=8  296 CTAs x  512", set * 1024.0, [&] { k_fetch<8><<<296, 512>>>((const float4*)hd, (float4*)d, d_sparse, n_words); });
    return 0;
}
