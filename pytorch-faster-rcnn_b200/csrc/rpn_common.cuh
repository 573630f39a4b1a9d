// rpn_common.cuh -- device helpers shared by select.cu (multi-kernel K3) and rpn_front.cu (cluster K3).
#pragma once
#include "common.cuh"
#include "pipeline.cuh"

namespace b2d {

__device__ __forceinline__ float load_logit(const float* __restrict__ cls, int n, int i, int mode, int C) {
    if (mode == 0) return cls[i];
    if (mode == 1) return cls[n + i] - cls[i];            // softmax[1] == sigmoid(l1 - l0)
    float m = cls[i];
    for (int c = 1; c < C; ++c) m = fmaxf(m, cls[(long long)c * n + i]);
    if (mode == 2) return m;                              // best class logit (sigmoid heads: sigmoid is monotone)
    // mode 3, softmax heads (lib/heads/anchor_head.py:232-236): max over the foreground classes c >= 1 of softmax_c
    float s = 0.0f, fg = -INFINITY;
    for (int c = 0; c < C; ++c) {
        const float v = cls[(long long)c * n + i];
        s += expf(v - m);
        if (c >= 1) fg = fmaxf(fg, v);
    }
    return expf(fg - m) / s;
}

__device__ __forceinline__ const float* seg_cls(const RpnLaunch& p, int b, int l) {
    const b2d_level& lv = p.pyr.lv[l];
    const long long n = (long long)lv.A * lv.H * lv.W;
    return p.cls[l] + (long long)b * n * p.cls_ch;
}
__device__ __forceinline__ const float* seg_reg(const RpnLaunch& p, int b, int l) {
    const b2d_level& lv = p.pyr.lv[l];
    const long long n = (long long)lv.A * lv.H * lv.W;
    return p.reg[l] + (long long)b * n * 4;
}

// ---------------------------------------------------------------- block bucket sort
// Descending sort of `total` distinct u64 composites (score key << 32 | ~index) held in
// s_list[0..total), one 1024-thread block.  The candidates are split into <= 4096 buckets
// that are linear in the monotone score key, bucket = (key_max - key) >> shift with the
// smallest shift that fits the candidates' key range, so buckets hold a handful of elements
// even in the dense part of the score distribution.  The sort is then: count per bucket,
// prefix over the buckets, scatter, and a rank-by-counting INSIDE each bucket -- ~7 block
// barriers instead of the 66-78 compare-exchange stages of a bitonic network.  Returns false
// (s_list unchanged) when a bucket is too large for the quadratic in-bucket ranking (heavily
// tied / degenerate score maps); the caller then falls back to the bitonic sort.
constexpr int kBucketCap = kSortCap / 2;             // s_list and s_bkt share the sort buffer
constexpr int kMaxBucket = 128;

static __device__ __noinline__ bool bucket_sort_desc(uint64_t* s_list, uint64_t* s_bkt, int total, uint32_t* s_off /*[kHistBins + 1]*/,
                                 uint32_t* s_cur /*[kHistBins]*/, uint32_t* s_wsum /*[66]*/, int dbg = 0) {
    const int t = threadIdx.x;
    if (dbg == 1) return false;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int i = t; i < total; i += blockDim.x) {
        const uint32_t key = (uint32_t)(s_list[i] >> 32);
        kmin = min(kmin, key); kmax = max(kmax, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if ((t & 31) == 0) { s_wsum[t >> 5] = kmin; s_wsum[33 + (t >> 5)] = kmax; }
    for (int i = t; i < kHistBins; i += blockDim.x) s_cur[i] = 0;
    __syncthreads();
    kmin = s_wsum[0]; kmax = s_wsum[33];
    for (int w = 1; w < 32; ++w) { kmin = min(kmin, s_wsum[w]); kmax = max(kmax, s_wsum[33 + w]); }
    int shift = 0;
    while (((kmax - kmin) >> shift) >= (uint32_t)kHistBins) ++shift;
    __syncthreads();                                  // s_wsum is reused below
    if (dbg == 2) return false;
    for (int i = t; i < total; i += blockDim.x) atomicAdd(&s_cur[(kmax - (uint32_t)(s_list[i] >> 32)) >> shift], 1u);
    __syncthreads();
    // exclusive prefix over the 4096 counts (4 per thread, 1024 threads)
    uint32_t c[4], sum = 0, mx = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { c[q] = s_cur[4 * t + q]; sum += c[q]; mx = max(mx, c[q]); }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((t & 31) == 31) s_wsum[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        uint32_t w = s_wsum[t], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (t >= o) wi += v;
        }
        s_wsum[t] = wi - w;                           // exclusive warp offsets
    }
    const bool too_big = __syncthreads_or(mx > (uint32_t)kMaxBucket);
    if (too_big || dbg == 3) return false;
    uint32_t run = s_wsum[t >> 5] + incl - sum;
#pragma unroll
    for (int q = 0; q < 4; ++q) { s_off[4 * t + q] = run; s_cur[4 * t + q] = run; run += c[q]; }
    if (t == blockDim.x - 1) s_off[kHistBins] = run;
    __syncthreads();
    for (int i = t; i < total; i += blockDim.x) {
        const uint64_t v = s_list[i];
        s_bkt[atomicAdd(&s_cur[(kmax - (uint32_t)(v >> 32)) >> shift], 1u)] = v;
    }
    __syncthreads();
    if (dbg == 4) return false;
    for (int i = t; i < total; i += blockDim.x) {
        const uint64_t v = s_bkt[i];
        const int rb = (int)((kmax - (uint32_t)(v >> 32)) >> shift);
        const int o = (int)s_off[rb], e = (int)s_off[rb + 1];
        int r = o;
        for (int q = o; q < e; ++q) r += (s_bkt[q] > v) ? 1 : 0;
        s_list[r] = v;
    }
    __syncthreads();
    return true;
}


}  // namespace b2d
