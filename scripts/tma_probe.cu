// Dev probe: cp.async.bulk (global -> shared) throughput vs copy size and ring depth, one
// producer thread + consumer-less ring per CTA (slots are recycled as soon as the copy lands).
// usage: tma_probe  -> table of GB/s
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// each CTA: `n` groups; a group = `per` copies of `sz` bytes (stride `stride` between them) into one slot, one mbarrier.
__global__ void k_probe(const char* src, long long src_bytes, int sz, int per, long long stride, int depth, int n, int lanes) {
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t bars[64];
    const int slot_bytes = sz * per;
    if (threadIdx.x == 0) for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    uint64_t rng = 0x9E3779B97F4A7C15ull * (blockIdx.x + 1);
    for (int i = 0; i < n; ++i) {
        const int slot = i % depth;
        const uint32_t par = (i / depth) & 1;
        if (i >= depth) mbar_wait(smem_u32(&bars[slot]), par ^ 1);   // previous use of the slot has landed
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        long long off = (long long)((rng >> 20) % (unsigned long long)((src_bytes - (long long)per * stride - sz) / 1024)) * 1024;
        if (lane == 0) mbar_expect_tx(smem_u32(&bars[slot]), slot_bytes);
        __syncwarp();
        if (lanes == 1) {
            if (lane == 0) for (int q = 0; q < per; ++q) bulk_g2s(smem_u32(smem) + slot * slot_bytes + q * sz, src + off + q * stride, sz, smem_u32(&bars[slot]));
        } else {
            for (int q = lane; q < per; q += 32) bulk_g2s(smem_u32(smem) + slot * slot_bytes + q * sz, src + off + q * stride, sz, smem_u32(&bars[slot]));
        }
    }
    for (int i = n; i < n + depth && i >= depth; ++i) mbar_wait(smem_u32(&bars[i % depth]), ((i / depth) & 1) ^ 1);
}

int main() {
    const long long bytes = 1ll << 30;
    char* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int sz, per; long long stride; int lanes; };
    const Cfg cfgs[] = {{16384, 1, 0, 1}, {8192, 1, 0, 1}, {4096, 1, 0, 1}, {2048, 1, 0, 1}, {1024, 1, 0, 1}, {512, 1, 0, 1},
                        {512, 8, 1024, 1}, {512, 8, 1024, 32}, {512, 16, 1024, 32}, {1024, 16, 1024, 32}, {1024, 16, 1024, 1}};
    printf("%-28s %8s %8s %10s\n", "copy", "ctas/sm", "inflightKB", "GB/s");
    for (const Cfg& c : cfgs)
        for (int cps = 1; cps <= 2; ++cps)
            for (int kb : {16, 32, 64, 96}) {
                const int slot = c.sz * c.per;
                int depth = kb * 1024 / slot; if (depth < 1) continue; if (depth > 64) depth = 64;
                const long long per_cta = 24ll << 20;                 // 24 MB per CTA
                const int n = (int)(per_cta / cps / slot);
                const int grid = 148 * cps;
                k_probe<<<grid, 64, depth * slot>>>(src, bytes, c.sz, c.per, c.stride, depth, n, c.lanes);
                cudaEventRecord(e0);
                k_probe<<<grid, 64, depth * slot>>>(src, bytes, c.sz, c.per, c.stride, depth, n, c.lanes);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                cudaError_t err = cudaGetLastError();
                char name[64]; snprintf(name, 64, "%dB x%d stride %lld lanes %d", c.sz, c.per, c.stride, c.lanes);
                printf("%-28s %8d %8d %10.0f %s\n", name, cps, depth * slot / 1024, (double)grid * n * slot / ms / 1e6, err ? cudaGetErrorString(err) : "");
            }
    return 0;
}
