// targets_common.cuh -- pieces of the RoI-target stage (lib/bbox.py:6-82) shared by assign.cu (k_roi_targets_small, one
// launch per batch) and rpn_back.cu (the same stage as the tail of the proposal cluster kernel).
#pragma once
#include "common.cuh"

namespace b2d {

constexpr int kGtChunk = 512;      // GTs staged in shared memory per pass
constexpr int kBoxesPerThread = 4;

struct AssignArgs {
    const float* boxes; long long box_ld; const int* box_count; long long N;
    int use_pyr;
    const float* img_hw; float border;
    const float* gt; int gt_ld; const int* gt_count;
    float pos_iou, neg_iou, min_pos_iou;
    int prepend_gt;
    long long out_ld;
};

constexpr int kSmallThreads = 1024;
constexpr int kSmallBoxes = 4;

struct SmallSmem {
    Box gt[kGtChunk];
    float ga[kGtChunk];
    uint32_t mx[kGtChunk];
    int cnt[3];
};

// ---- device-RNG sampler (spec: DESIGN.md "Samplers"; oracle/sampler_spec.py) ----
// 32-bit round function of the Feistel permutation (murmur3-style finaliser, keyed): the 4 rounds of a 64-bit
// splitmix per permutation step were half of the RoI-target kernel's time (r2: ~5 of 11 us sampler + encode).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t key, uint32_t v) {
    uint32_t h = v * 0x9E3779B1u + key;
    h ^= h >> 15; h *= 0x85EBCA77u;
    h ^= h >> 13; h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return h;
}

__device__ __forceinline__ uint32_t feistel(uint32_t x, int half_bits, uint64_t seed) {
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
        const uint64_t ks = seed + 0x1000003ull * (uint64_t)(round + 1);
        const uint32_t f = mix32((uint32_t)ks ^ (uint32_t)(ks >> 32), r) & mask;
        const uint32_t nl = r;
        r = l ^ f;
        l = nl;
    }
    return (l << half_bits) | r;
}

constexpr int kFusedMaxN = kSmallThreads * kSmallBoxes + kGtChunk;     // candidates incl. prepended GT
constexpr int kFusedPer = (kFusedMaxN + kSmallThreads - 1) / kSmallThreads;

struct FusedArgs {
    int* chosen; int* n_chosen; int max_num, pos_num;
    unsigned long long seed;
    const unsigned long long* seed_step;      // optional device counter added to seed (advanced by b2d_counter_add)
    const int64_t* gt_label;
    float* tar_box; float* tar_gt; float* tar_param; int64_t* tar_label; int64_t* tar_is_gt;
    float ms[8];
};


// scratch of fused_sample_encode (shared memory)
struct FusedScratch {
    int lab[kFusedMaxN];                 // labels of the candidates (GT rows first) as int32
    unsigned char flag[kFusedMaxN];      // 1 = sampled
    int chosen[kSmallThreads];
    int w[4][32];
    int tot, red[32];
    unsigned long long lo, hi;
};

// Device-RNG sampler + ascending compaction + gather / encode of the sampled rows for image b, one 1024-thread CTA.
// fs.lab holds the labels of the lead (prepended GT) + proposal candidates; npos_boxes / nneg_boxes count the positive /
// negative PROPOSALS (the prepended GT rows are added here); s_gt: the image's GT boxes in shared memory.
__device__ __forceinline__ void fused_sample_encode(const AssignArgs& p, const FusedArgs& f, const Box* s_gt, FusedScratch& fs,
                                                    int npos_boxes, int nneg_boxes, int b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.gt_count[b];
    const int lead = p.prepend_gt ? K : 0;
    const int n_tot = lead + (p.box_count ? p.box_count[b] : (int)p.N);
    const int npos = npos_boxes + lead, nneg = nneg_boxes;
    const uint64_t sd = f.seed + (f.seed_step ? *f.seed_step : 0ull) + 0x632BE59BD9B4E019ull * (uint64_t)(b + 1);
    for (int i = tid; i < n_tot; i += kSmallThreads) fs.flag[i] = 0;
    // ---- positives: all of them, or the pos_num smallest (mix_key(sd, idx) << 32 | idx)
    const int keep_pos = min(npos, f.pos_num);
    uint64_t thr = ~0ull;
    if (npos > f.pos_num) {
        if (tid == 0) { fs.lo = 0ull; fs.hi = ~0ull; }
        __syncthreads();
        for (int it = 0; it < 64; ++it) {
            const unsigned long long lo = fs.lo, hi = fs.hi;
            if (lo >= hi) break;
            const unsigned long long mid = lo + (hi - lo) / 2;
            int c = 0;
            for (int i = tid; i < n_tot; i += kSmallThreads)
                if (fs.lab[i] > 0) c += ((((uint64_t)mix_key(sd, (uint32_t)i) << 32) | (uint32_t)i) <= mid);
            c = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0) fs.red[warp] = c;
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int w = 0; w < kSmallThreads / 32; ++w) tot += fs.red[w];
                if (tot >= f.pos_num) fs.hi = mid; else fs.lo = mid + 1;
            }
            __syncthreads();
        }
        thr = fs.lo;
    }
    __syncthreads();
    for (int i = tid; i < n_tot; i += kSmallThreads)
        if (fs.lab[i] > 0 && ((((uint64_t)mix_key(sd, (uint32_t)i) << 32) | (uint32_t)i) <= thr)) fs.flag[i] = 1;
    // ---- negatives: the first want_neg label-0 indices along the keyed Feistel permutation
    const int want_neg = min(max(f.max_num - keep_pos, 0), nneg);
    if (want_neg > 0) {
        int bits = 2;
        while ((1 << bits) < n_tot) ++bits;
        if (bits & 1) ++bits;
        const int half = bits >> 1, dom = 1 << bits;
        int taken = 0;
        for (int t0 = 0; t0 < dom && taken < want_neg; t0 += 4 * kSmallThreads) {
            uint32_t y[4];
            unsigned m[4];
            bool hit[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q * kSmallThreads + tid;
                y[q] = feistel((uint32_t)t, half, sd ^ 0xA5A5A5A5DEADBEEFull);
                hit[q] = t < dom && (int)y[q] < n_tot && fs.lab[y[q]] == 0;
                m[q] = __ballot_sync(0xffffffffu, hit[q]);
                if (lane == 0) fs.w[q][warp] = __popc(m[q]);
            }
            __syncthreads();
            if (warp == 0) {                                        // exclusive scan of the 128 warp counts (t order)
                int v[4], sum = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) { v[e] = (&fs.w[0][0])[lane * 4 + e]; sum += v[e]; }
                int incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                int run = incl - sum;
#pragma unroll
                for (int e = 0; e < 4; ++e) { (&fs.w[0][0])[lane * 4 + e] = run; run += v[e]; }
                if (lane == 31) fs.tot = incl;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int rank = taken + fs.w[q][warp] + __popc(m[q] & ((1u << lane) - 1u));
                if (hit[q] && rank < want_neg) fs.flag[y[q]] = 1;
            }
            taken += fs.tot;
            __syncthreads();
        }
    }
    __syncthreads();
    // ---- ascending compaction of the flags
    const int total = keep_pos + want_neg;
    {
        int c = 0;
        const int i0 = tid * kFusedPer;
#pragma unroll
        for (int e = 0; e < kFusedPer; ++e) c += (i0 + e < n_tot) ? fs.flag[i0 + e] : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) fs.red[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = fs.red[lane];
            int iw = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, iw, d);
                if (lane >= d) iw += t;
            }
            fs.red[lane] = iw - v;
        }
        __syncthreads();
        int pos = fs.red[warp] + incl - c;
#pragma unroll
        for (int e = 0; e < kFusedPer; ++e)
            if (i0 + e < n_tot && fs.flag[i0 + e]) { if (pos < kSmallThreads) fs.chosen[pos] = i0 + e; ++pos; }
    }
    __syncthreads();
    // ---- gather + encode of the sampled rows (k_encode_targets)
    if (tid == 0) f.n_chosen[b] = total;
    const int t = tid;
    if (t >= f.max_num) return;
    const long long o = (long long)b * 4 * f.max_num;
    const bool live = t < total;
    Box bx{0, 0, 0, 0}, gb{0, 0, 0, 0};
    float prm[4] = {0, 0, 0, 0};
    int64_t lab_out = 0, isgt = 0;
    int ci = -1;
    if (live) {
        const int i = fs.chosen[t];
        ci = i;
        const int lab = fs.lab[i];
        if (i < lead) {
            bx = s_gt[i];
            isgt = 1;
        } else {
            const float* src = p.boxes + (long long)b * 4 * p.box_ld;
            const long long ii = i - lead;
            bx = Box{src[ii], src[p.box_ld + ii], src[2 * p.box_ld + ii], src[3 * p.box_ld + ii]};
        }
        const int j = max(lab - 1, 0);                       // negatives point at GT 0 (lib/anchor.py:45-47)
        gb = s_gt[j];
        const float bw = (bx.x2 - bx.x1) + 1.0f, bh = (bx.y2 - bx.y1) + 1.0f;
        const float gw = (gb.x2 - gb.x1) + 1.0f, gh = (gb.y2 - gb.y1) + 1.0f;
        const float bcx = (bx.x2 + bx.x1) / 2.0f, bcy = (bx.y2 + bx.y1) / 2.0f;
        const float gcx = (gb.x2 + gb.x1) / 2.0f, gcy = (gb.y2 + gb.y1) / 2.0f;
        prm[0] = ((gcx - bcx) / bw - f.ms[0]) / f.ms[4];
        prm[1] = ((gcy - bcy) / bh - f.ms[1]) / f.ms[5];
        prm[2] = (logf(gw / bw) - f.ms[2]) / f.ms[6];
        prm[3] = (logf(gh / bh) - f.ms[3]) / f.ms[7];
        if (f.gt_label) lab_out = (lab > 0) ? f.gt_label[(long long)b * p.gt_ld + j] : 0;
        else lab_out = (lab > 0) ? 1 : 0;
    }
    const int mn = f.max_num;
    f.chosen[(long long)b * mn + t] = ci;
    if (f.tar_box) { f.tar_box[o + t] = bx.x1; f.tar_box[o + mn + t] = bx.y1; f.tar_box[o + 2 * mn + t] = bx.x2; f.tar_box[o + 3 * mn + t] = bx.y2; }
    if (f.tar_gt) { f.tar_gt[o + t] = gb.x1; f.tar_gt[o + mn + t] = gb.y1; f.tar_gt[o + 2 * mn + t] = gb.x2; f.tar_gt[o + 3 * mn + t] = gb.y2; }
    if (f.tar_param) { f.tar_param[o + t] = prm[0]; f.tar_param[o + mn + t] = prm[1]; f.tar_param[o + 2 * mn + t] = prm[2]; f.tar_param[o + 3 * mn + t] = prm[3]; }
    if (f.tar_label) f.tar_label[(long long)b * mn + t] = lab_out;
    if (f.tar_is_gt) f.tar_is_gt[(long long)b * mn + t] = isgt;
}

}  // namespace b2d
