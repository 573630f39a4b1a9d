// rpn_back.cu -- K4 + merge of the fused RPN proposal path as ONE thread-block-cluster kernel, a cluster of 8 CTAs per
// image: score cut -> x-sweep suppression mask -> fixed-point scan -> (conditional second attempt on all boxes) ->
// 5-way merge to the max_num best survivors.  Replaces k_nms_cut, k_nms_sweep, k_nms_scan_fp, the idle fallback pair
// (k_nms_mask_sym_fb + k_nms_scan_fp) and k_merge_rank (nms.cu / select.cu): six launches, each waiting for the slowest
// image of the batch, become one launch in which an image only waits for itself (cluster barriers).
// Reference: RPNHead.predict_single_image, lib/heads/rpn_head.py:103-118 (tvops.nms per level, [:post_nms], cat,
// topk(max_num)); NMS semantics of torchvision.ops.nms (see nms.cu).
//
// Phases of a cluster (image b); "all" = every CTA redundantly, "split" = work divided over the 8 x 32 warps:
//   C  all    score cut: the 16 leading key bits of the M-th best selected box over all levels (2-pass radix select on
//             the <= 5 x 2048 keys in shared memory) -> boxes per level at or above it.  Exact early termination: a box
//             survives or not depending only on higher-scored boxes of its level, and only the max_num best survivors
//             are kept, so the NMS of the M best boxes is final whenever it leaves >= max_num survivors.
//   Z  split  clear the mask words / row bitmaps the attempt can touch
//   B  all    stage the attempt's boxes in shared memory, bucket every level by x1 (counting sort, 256 cells)
//   S  split  sweep: a warp per box, lanes over the boxes whose x1 lies in [x1_i, x1_i + (1 - 0.9 thr) w_i + slack]
//             (see k_nms_sweep in nms.cu for the pruning bound); bits are OR-ed into the L2-resident mask.
//             Dense variant (all pairs, boxes from global memory) when the sweep cannot be used for this image:
//             malformed boxes, x1 values too concentrated for the cells to prune, more boxes than the staging area.
//   N  CTA l  greedy NMS of level l as the fixed point of kept_j = !(exists i < j: kept_i and M_ij) (k_nms_scan_fp)
//   K  all    enough survivors?  otherwise one more attempt on ALL selected boxes (phases Z..N again)
//   M  CTA 0  rank of every survivor among all levels by binary searches over the other levels' key lists; scatter
//             boxes / sigmoid scores / provenance to the outputs.
// Global memory written by one CTA and read by another goes through __threadfence + cluster barrier + ld.cg.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "nms_common.cuh"
#include "pipeline.cuh"
#include "targets_common.cuh"

namespace cg = cooperative_groups;

namespace b2d {

constexpr int kBkThreads = 1024;
constexpr int kBkCl = 8;                 // CTAs per cluster == images' max level count
constexpr int kBkBoxes = 4096;           // boxes of one attempt staged in shared memory
constexpr int kBkCells = 256;            // x1 cells per level
constexpr int kBkRows = 2;               // scan: rows per thread (2048 boxes per level)
constexpr int kBkEntries = 6;            // scan: non-zero words of a row kept in registers

// RoI-target stage as the tail of the kernel (b2d_rpn_proposals_targets); enable == 0: proposals only
struct BackTargets {
    int enable;
    AssignArgs p;
    FusedArgs f;
    int64_t* labels; float* out_iou; int* census; int* pos_list; int pos_cap;
};

// shared memory of the target tail, aliased over BkShared::box (dead after the merge)
struct TgShared {
    Box gt[kGtChunk];
    float ga[kGtChunk];
    uint32_t cm_local[kGtChunk], cm[kGtChunk];
    int cnt[4];
    FusedScratch fs;
};

struct BackArgs {
    BackTargets tg;
    float thr, thr_lo, thr_hi, prune;
    int M;                               // score cut: attempt 1 on the M best boxes (0: all boxes at once)
    int use_sweep;                       // 0: dense mask only
    float* props; float* scores; int* count; int* prov;
};

struct BkShared {
    float4 box[kBkBoxes];
    uint16_t ord[kBkBoxes];
    int start[kMaxLevels][kBkCells + 1];
    int cur[kMaxLevels][kBkCells];
    int n[kMaxLevels], off[kMaxLevels + 1];      // boxes per level in this attempt, their offsets in `box`
    int ncut[kMaxLevels], nsel[kMaxLevels];
    int keep[kMaxLevels];                        // survivors per level (capped by post_nms)
    unsigned long long best;                     // cut: min over candidates of (count << 32 | tau)
    int dense;                                   // this attempt uses the dense mask
    int next;                                    // sweep: next box of this CTA's share (dynamic distribution)
    int wp[kMaxLevels];                          // mask row pitch in words, per level (kernel parameters indexed by a
    long long moff[kMaxLevels], soff[kMaxLevels], aoff[kMaxLevels];   // register cost a constant-bank round trip each)
    uint32_t kbits[2][64];                       // scan: kept bits, ping-pong
    uint32_t abits[kMaxLevels][64];              // kept bits of every level (written by the level's scan CTA)
    int apre[kMaxLevels][65];                    // prefix popcounts of abits
};

__device__ __forceinline__ int bk_cell(float x, float inv) {
    return min(kBkCells - 1, max(0, (int)(x * inv)));    // monotone in x; boxes are clipped to [0, img_w)
}

__device__ __forceinline__ uint64_t ldcg_u64(const uint64_t* p) {
    return (uint64_t)__ldcg(reinterpret_cast<const unsigned long long*>(p));
}

// number of keys in the descending list k[0..n) that are > key (strict) or >= key
__device__ __forceinline__ int count_before(const uint32_t* k, int n, uint32_t key, bool or_equal) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t km = k[mid];
        if (or_equal ? (km >= key) : (km > key)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kBkThreads, 1) k_rpn_back(RpnLaunch p, BackArgs a) {
    pdl_wait();                                         // the selection kernel's lists (no-op unless launched as its programmatic dependent)
    pdl_launch_dependents();                            // the RoI-target kernel behind us may be set up now
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char s_raw[];
    BkShared& s = *reinterpret_cast<BkShared*>(s_raw);
    uint32_t* s_keys = reinterpret_cast<uint32_t*>(s_raw + ((sizeof(BkShared) + 15) & ~(size_t)15));   // all selected keys, level l at sel_off[l]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int crank = (int)cl.block_rank();
    const int b = blockIdx.y;
    const int L = p.L;
    const long long ibase = (long long)b * p.sel_per_img;
    const float4* g_box = p.sel_box + ibase;
    const uint32_t* g_key = p.sel_key + ibase;
    uint64_t* g_mask = p.mask + (long long)b * p.mask_per_img;
    uint32_t* g_nz = p.nz + ibase;
    const int per_img = (int)p.sel_per_img;
    dbg_stamp(p, b, 32 + crank, 0);

    // ------------------------------------------------------------------ C: score cut (every CTA, redundantly)
    if (tid < kMaxLevels) {
        s.nsel[tid] = tid < L ? p.sel_count[b * L + tid] : 0;
        s.wp[tid] = tid < L ? (p.kcap[tid] + 63) >> 6 : 0;
        s.moff[tid] = tid < L ? p.mask_off[tid] : 0;
        s.soff[tid] = tid < L ? p.sel_off[tid] : 0;
        s.aoff[tid] = tid < L ? p.pyr.lv[tid].offset : 0;
    }
    if (tid == 0) s.best = ~0ull;
    // all keys of the image in one flat pass: independent loads, one L2 round trip (slots past a level's count: 0)
    for (int t = tid; t < per_img; t += kBkThreads) s_keys[t] = g_key[t];
    __syncthreads();
    int total_sel = 0;
    for (int l = 0; l < L; ++l) total_sel += s.nsel[l];
    if (a.M > 0 && total_sel > a.M) {
        // Any threshold tau gives an exact result (a coarser one only admits more boxes than M), so the candidates are
        // every 16th key of every (descending) level list plus each list's last key; count(tau) = sum over levels of
        // the keys >= tau (binary searches in shared memory); the cut is the candidate with the smallest count >= M.
        int ncand = 0;
        for (int l = 0; l < L; ++l) ncand += (s.nsel[l] + 15) >> 4;
        for (int c0 = tid; c0 < ncand; c0 += kBkThreads) {
            int l = 0, c = c0;
            while (c >= ((s.nsel[l] + 15) >> 4)) { c -= (s.nsel[l] + 15) >> 4; ++l; }
            const int pos = min(s.nsel[l], (c + 1) << 4) - 1;
            const uint32_t tau = s_keys[s.soff[l] + pos];
            uint32_t cnt = 0;
            for (int q = 0; q < L; ++q) cnt += (uint32_t)count_before(s_keys + s.soff[q], s.nsel[q], tau, true);
            if (cnt >= (uint32_t)a.M) atomicMin(&s.best, ((unsigned long long)cnt << 32) | tau);
        }
        __syncthreads();
        const uint32_t tau = (uint32_t)s.best;
        if (tid < kMaxLevels) s.ncut[tid] = tid < L ? count_before(s_keys + s.soff[tid], s.nsel[tid], tau, true) : 0;
    } else {
        if (tid < kMaxLevels) s.ncut[tid] = s.nsel[tid];
    }
    __syncthreads();
    dbg_stamp(p, b, 32 + crank, 1);

    const float cell_inv = (float)kBkCells / fmaxf(p.img_hw[2 * b + 1], 1.0f);
    const int gw = crank * (kBkThreads / 32) + warp, nw = kBkCl * (kBkThreads / 32);
    // ------------------------------------------------------------------ attempts
    for (int att = 0; att < 2; ++att) {
        if (tid == 0) {
            int run = 0;
            for (int l = 0; l < kMaxLevels; ++l) {
                s.n[l] = l < L ? (att == 0 ? s.ncut[l] : s.nsel[l]) : 0;
                s.off[l] = run; run += s.n[l];
            }
            s.off[kMaxLevels] = run;
            s.dense = (!a.use_sweep || run > kBkBoxes) ? 1 : 0;
            s.next = 0;
        }
        for (int t = tid; t < kMaxLevels * (kBkCells + 1); t += kBkThreads) (&s.start[0][0])[t] = 0;
        __syncthreads();
        const int total = s.off[kMaxLevels];
        // ---- Z: clear what the attempt can touch: the first n_l rows of every level (contiguous), split over the cluster
        for (int l = 0; l < L; ++l) {
            const int n = s.n[l];
            uint64_t* m = g_mask + s.moff[l];
            const int words = n * s.wp[l];
            for (int t = crank * kBkThreads + tid; t < words; t += kBkCl * kBkThreads) m[t] = 0ull;
            for (int t = crank * kBkThreads + tid; t < n; t += kBkCl * kBkThreads) g_nz[s.soff[l] + t] = 0u;
        }
        if (att == 0) dbg_stamp(p, b, 32 + crank, 12);
        // ---- B: stage the boxes and bucket every level by x1 (every CTA: all of them sweep over all levels)
        if (!s.dense) {
            bool ok = true;
            for (int t = tid; t < total; t += kBkThreads) {
                int l = 0;
                for (int q = 1; q < L; ++q) if (t >= s.off[q]) l = q;
                const float4 bx = g_box[s.soff[l] + (t - s.off[l])];
                s.box[t] = bx;
                ok = ok && well_formed(bx);
                atomicAdd(&s.start[l][bk_cell(bx.x, cell_inv) + 1], 1);
            }
            if (!__syncthreads_and(ok)) { if (tid == 0) s.dense = 1; }
            if (warp < L) {                                // warp l: inclusive scan of level l's 256 counts, 8 per lane
                const int l = warp;
                int c[8], sum = 0;
                long long sq = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { c[q] = s.start[l][lane * 8 + q + 1]; sum += c[q]; sq += (long long)c[q] * c[q]; }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                int run = incl - sum;
#pragma unroll
                for (int q = 0; q < 8; ++q) { s.cur[l][lane * 8 + q] = run; run += c[q]; s.start[l][lane * 8 + q + 1] = run; }
                if (lane == 0 && sq > (long long)s.n[l] * 192) s.dense = 1;     // cells no longer prune
            }
            __syncthreads();
            if (!s.dense) {
                for (int t = tid; t < total; t += kBkThreads) {
                    int l = 0;
                    for (int q = 1; q < L; ++q) if (t >= s.off[q]) l = q;
                    s.ord[s.off[l] + atomicAdd(&s.cur[l][bk_cell(s.box[t].x, cell_inv)], 1)] = (uint16_t)(t - s.off[l]);
                }
            }
        }
        if (att == 0) dbg_stamp(p, b, 32 + crank, 13);
        __threadfence();
        dbg_stamp(p, b, 32 + crank, 2 + 4 * att);
        cl.sync();                                         // mask cleared everywhere; s.dense is the same in every CTA
        // ---- S: suppression bits.  CTA c takes the boxes t = c (mod 8); its warps fetch them dynamically (the candidate
        // ranges are heavy-tailed: mean 40, max ~600 on config 2)
        if (!s.dense) {
            for (;;) {
                int t = 0;
                if (lane == 0) t = atomicAdd(&s.next, 1);
                t = __shfl_sync(0xffffffffu, t, 0) * kBkCl + crank;
                if (t >= total) break;
                int l = 0;
                for (int q = 1; q < L; ++q) if (t >= s.off[q]) l = q;
                const int i = t - s.off[l];
                const float4* lb = s.box + s.off[l];
                const uint16_t* lo = s.ord + s.off[l];
                const float4 bi = lb[i];
                const float wi = bi.z - bi.x;
                if (!(wi > 0.0f) || !(bi.w - bi.y > 0.0f)) continue;          // empty box: inter == 0 with everything
                const float ai = area_of(bi);
                const float xhi = bi.x + a.prune * wi + 1.0e-4f * (fabsf(bi.z) + 1.0f);
                const int wp = s.wp[l];
                uint64_t* m = g_mask + s.moff[l];
                uint32_t* nz = g_nz + s.soff[l];
                const int k1 = s.start[l][bk_cell(xhi, cell_inv) + 1];
                for (int k = s.start[l][bk_cell(bi.x, cell_inv)] + lane; k < k1; k += 32) {
                    const int j = lo[k];
                    const float4 bj = lb[j];
                    if (bj.x < bi.x || (bj.x == bi.x && j <= i) || bj.x > xhi) continue;   // every unordered pair once
                    if (!(bj.y < bi.w && bi.y < bj.w)) continue;                           // no y overlap: inter == 0
                    if (!suppresses_wf(bi, ai, bj, area_of(bj), a.thr, a.thr_lo, a.thr_hi)) continue;
                    atomicOr(reinterpret_cast<unsigned long long*>(&m[(long long)i * wp + (j >> 6)]), 1ull << (j & 63));
                    atomicOr(reinterpret_cast<unsigned long long*>(&m[(long long)j * wp + (i >> 6)]), 1ull << (i & 63));
                    atomicOr(&nz[i], 1u << (j >> 6));
                    atomicOr(&nz[j], 1u << (i >> 6));
                }
            }
        } else {
            // all pairs i < j of every level, a warp per row, boxes from global memory (rare: see the header)
            for (int t = gw; t < total; t += nw) {
                int l = 0;
                for (int q = 1; q < L; ++q) if (t >= s.off[q]) l = q;
                const int i = t - s.off[l], n = s.n[l];
                const float4* lb = g_box + s.soff[l];
                const float4 bi = lb[i];
                const float ai = area_of(bi);
                const int wp = s.wp[l];
                uint64_t* m = g_mask + s.moff[l];
                uint32_t* nz = g_nz + s.soff[l];
                for (int j0 = (i + 1) & ~31; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    bool sup = false;
                    if (j > i && j < n) {
                        const float4 bj = lb[j];
                        sup = suppresses(bi, ai, bj, area_of(bj), a.thr, a.thr_lo, a.thr_hi);
                    }
                    const unsigned bits = __ballot_sync(0xffffffffu, sup);
                    if (bits) {
                        if (lane == 0) {
                            atomicOr(reinterpret_cast<unsigned long long*>(&m[(long long)i * wp + (j0 >> 6)]),
                                     (unsigned long long)bits << (j0 & 32));
                            atomicOr(&nz[i], 1u << (j0 >> 6));
                        }
                        if (sup) {
                            atomicOr(reinterpret_cast<unsigned long long*>(&m[(long long)j * wp + (i >> 6)]), 1ull << (i & 63));
                            atomicOr(&nz[j], 1u << (i >> 6));
                        }
                    }
                }
            }
        }
        if (att == 0) dbg_stamp(p, b, 32 + crank, 14);
        __threadfence();
        dbg_stamp(p, b, 32 + crank, 3 + 4 * att);
        cl.sync();
        if (att == 0) dbg_stamp(p, b, 32 + crank, 15);
        // ---- N: fixed-point scan of level `crank`; the kept bits go to every CTA of the cluster
        if (crank < L) {
            const int l = crank;
            const int n = s.n[l];
            const int wp = s.wp[l];
            const uint64_t* m = g_mask + s.moff[l];
            const uint32_t* nz = g_nz + s.soff[l];
            uint64_t eb[kBkRows][kBkEntries];
            uint32_t ew[kBkRows], over[kBkRows];
#pragma unroll
            for (int q = 0; q < kBkRows; ++q) {
                const int row = tid + q * kBkThreads;
                uint32_t mm = 0;
                if (row < n) {
                    const int dw = row >> 6;
                    mm = __ldcg(&nz[row]) & (dw == 31 ? 0xffffffffu : ((2u << dw) - 1u));
                }
                ew[q] = 0;
#pragma unroll
                for (int e = 0; e < kBkEntries; ++e) {
                    uint64_t word = 0ull;
                    if (mm) {
                        const int w = __ffs(mm) - 1;
                        mm &= mm - 1;
                        word = ldcg_u64(&m[(long long)row * wp + w]);
                        if (w == (row >> 6)) word &= (1ull << (row & 63)) - 1ull;
                        ew[q] |= (uint32_t)w << (5 * e);
                    }
                    eb[q][e] = word;
                }
                over[q] = mm;                                // rare: more than kBkEntries candidate words
            }
            if (tid < 64) {
                const int lo = tid * 32;
                s.kbits[0][tid] = n >= lo + 32 ? 0xffffffffu : (n > lo ? ((1u << (n - lo)) - 1u) : 0u);
            }
            __syncthreads();
            int it = 0;
            for (;; ++it) {
                const uint32_t* cur = s.kbits[it & 1];
                uint32_t* nxt = s.kbits[(it & 1) ^ 1];
                bool changed = false;
#pragma unroll
                for (int q = 0; q < kBkRows; ++q) {
                    const int row = tid + q * kBkThreads;
                    bool sup = false;
#pragma unroll
                    for (int e = 0; e < kBkEntries; ++e) {
                        if (eb[q][e]) {
                            const int w = (ew[q] >> (5 * e)) & 31;
                            const uint64_t k = (uint64_t)cur[2 * w] | ((uint64_t)cur[2 * w + 1] << 32);
                            sup |= (k & eb[q][e]) != 0ull;
                        }
                    }
                    uint32_t mm = over[q];
                    while (mm && !sup) {
                        const int w = __ffs(mm) - 1;
                        mm &= mm - 1;
                        uint64_t word = ldcg_u64(&m[(long long)row * wp + w]);
                        if (w == (row >> 6)) word &= (1ull << (row & 63)) - 1ull;
                        const uint64_t k = (uint64_t)cur[2 * w] | ((uint64_t)cur[2 * w + 1] << 32);
                        sup |= (k & word) != 0ull;
                    }
                    const bool nk = (row < n) && !sup;
                    const uint32_t bits = __ballot_sync(0xffffffffu, nk);
                    const uint32_t old = cur[row >> 5];
                    if (lane == 0) nxt[row >> 5] = bits;
                    changed |= (bits != old);
                }
                if (!__syncthreads_or(changed)) break;
            }
            const uint32_t* fin = s.kbits[(it & 1) ^ 1];
            if (tid < 64 * kBkCl) cl.map_shared_rank(&s.abits[l][0], tid >> 6)[tid & 63] = fin[tid & 63];
        }
        dbg_stamp(p, b, 32 + crank, 4 + 4 * att);
        cl.sync();
        // ---- K: prefix popcounts of every level's kept bits; enough survivors?
        if (warp < L) {
            const uint32_t w0 = s.abits[warp][2 * lane], w1 = s.abits[warp][2 * lane + 1];
            const int c = __popc(w0) + __popc(w1);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s.apre[warp][2 * lane] = incl - c;
            s.apre[warp][2 * lane + 1] = incl - c + __popc(w0);
            if (lane == 31) {
                s.apre[warp][64] = incl;
                s.keep[warp] = (p.post_nms > 0 && incl > p.post_nms) ? p.post_nms : incl;
            }
        }
        __syncthreads();
        int kept = 0;
        bool full = true;
        for (int l = 0; l < L; ++l) { kept += s.keep[l]; full = full && (s.n[l] == s.nsel[l]); }
        if (full || (p.max_num > 0 && kept >= p.max_num)) break;             // cluster-uniform
    }
    dbg_stamp(p, b, 32 + crank, 10);

    // ------------------------------------------------------------------ M: merge, split over the cluster's threads
    // rank of a survivor = survivors before it in its own level + per other level the survivors among the boxes that
    // precede it there (binary search over that level's keys, then the prefix popcount of its kept bits); ties go to
    // the earlier concat position, i.e. the lower level.
    int total_kept = 0;
    for (int l = 0; l < L; ++l) total_kept += s.keep[l];
    const bool topk = (p.max_num > 0) && (total_kept > p.max_num);
    const int nout = topk ? p.max_num : total_kept;
    const int total = s.off[kMaxLevels];
    const bool staged = !s.dense;
    float* pb = a.props + (long long)b * 4 * p.out_ld;
    float* ps = a.scores + (long long)b * p.out_ld;
    int* pv = a.prov ? a.prov + (long long)b * p.out_ld : nullptr;
    for (int t = tid * kBkCl + crank; t < total; t += kBkCl * kBkThreads) {     // interleaved over the CTAs
        int l = 0;
        for (int q = 1; q < L; ++q) if (t >= s.off[q]) l = q;
        const int r = t - s.off[l];
        const uint32_t kw = s.abits[l][r >> 5];
        if (!((kw >> (r & 31)) & 1u)) continue;
        const int own = s.apre[l][r >> 5] + __popc(kw & ((1u << (r & 31)) - 1u));
        if (p.post_nms > 0 && own >= p.post_nms) continue;
        const uint32_t key = s_keys[s.soff[l] + r];
        int rank = own;
        for (int q = 0; q < L; ++q) {
            if (q == l) continue;
            const int pq = count_before(s_keys + s.soff[q], s.n[q], key, q < l);
            const int wq = pq >> 5, bq = pq & 31;
            int c = s.apre[q][wq] + (bq ? __popc(s.abits[q][wq] & ((1u << bq) - 1u)) : 0);
            if (p.post_nms > 0 && c > p.post_nms) c = p.post_nms;
            rank += c;
        }
        if (!topk) {                                      // no global cut: concat order
            rank = own;
            for (int q = 0; q < l; ++q) rank += s.keep[q];
        }
        if (rank < nout) {
            const long long o = s.soff[l] + r;
            const float4 bx = staged ? s.box[t] : g_box[o];
            pb[rank] = bx.x; pb[p.out_ld + rank] = bx.y; pb[2 * p.out_ld + rank] = bx.z; pb[3 * p.out_ld + rank] = bx.w;
            const float sc = 1.0f / (1.0f + expf(-key2f(key)));
            ps[rank] = sc;
            if (pv) pv[rank] = (int)(s.aoff[l] + p.sel_idx[ibase + o]);
            if (p.rec) { float* q = p.rec + ((long long)b * p.out_ld + rank) * 5; q[0] = bx.x; q[1] = bx.y; q[2] = bx.z; q[3] = bx.w; q[4] = sc; }
        }
    }
    for (int t = nout + crank * kBkThreads + tid; t < p.out_ld; t += kBkCl * kBkThreads) {   // padding slots
        pb[t] = 0.f; pb[p.out_ld + t] = 0.f; pb[2 * p.out_ld + t] = 0.f; pb[3 * p.out_ld + t] = 0.f;
        ps[t] = 0.f;
        if (pv) pv[t] = -1;
        if (p.rec) { float* q = p.rec + ((long long)b * p.out_ld + t) * 5; q[0] = q[1] = q[2] = q[3] = q[4] = 0.f; }
    }
    if (crank == 0 && tid == 0) a.count[b] = nout;
    if (tid < kMaxLevels && crank == 0 && tid < L) p.keep_count[b * L + tid] = s.keep[tid];
    dbg_stamp(p, b, 32 + crank, 11);
    if (!a.tg.enable) return;

    // ------------------------------------------------------------------ T: bbox_target on the proposals (lib/bbox.py:6-82)
    // MaxIoUAssigner over the image's nout proposals, a proposal per thread of the cluster (lib/region.py:75-107: per-GT
    // column maxima exchanged through distributed shared memory, then labels); GT rows prepended, device-RNG sampler,
    // gather + encode in the first CTA (fused_sample_encode, the tail of k_roi_targets_small).
    static_assert(sizeof(TgShared) <= sizeof(float4) * kBkBoxes, "target tail must fit the staged-box area");
    __threadfence();
    cl.sync();                                            // the proposals are complete in global memory; s.box is dead
    dbg_stamp(p, b, 32 + crank, 5);
    TgShared& t = *reinterpret_cast<TgShared*>(s.box);
    const AssignArgs& ap = a.tg.p;
    const int K = ap.gt_count[b];
    const int lead = ap.prepend_gt ? K : 0;
    const float* g = ap.gt + (long long)b * 4 * ap.gt_ld;
    const float* src = a.props + (long long)b * 4 * p.out_ld;
    const uint32_t kNegInf = f2key(-INFINITY);
    if (tid < 4) t.cnt[tid] = 0;
    for (int j = tid; j < K; j += kBkThreads) {
        const Box q{g[j], g[ap.gt_ld + j], g[2 * ap.gt_ld + j], g[3 * ap.gt_ld + j]};
        t.gt[j] = q; t.ga[j] = area_plus1(q); t.cm_local[j] = kNegInf;
    }
    const int i = crank * kBkThreads + tid;               // nout <= 4096 < 8 * 1024: at most one proposal per thread
    const bool ok = i < nout;
    Box bx{0.f, 0.f, 0.f, 0.f};
    if (ok) bx = Box{__ldcg(src + i), __ldcg(src + p.out_ld + i), __ldcg(src + 2 * p.out_ld + i), __ldcg(src + 3 * p.out_ld + i)};
    const float ba = ok ? area_plus1(bx) : 0.0f;
    __syncthreads();
    for (int j = 0; j < K; ++j) {                         // pass 1: per-GT column max, -0 folded to +0
        uint32_t m = kNegInf;
        if (ok) m = f2key(iou_plus1(bx, ba, t.gt[j], t.ga[j]) + 0.0f);
        m = __reduce_max_sync(0xffffffffu, m);
        if (lane == 0 && m != kNegInf) atomicMax(&t.cm_local[j], m);
    }
    dbg_stamp(p, b, 32 + crank, 6);
    cl.sync();
    for (int j = tid; j < K; j += kBkThreads) {
        uint32_t m = kNegInf;
        for (int r = 0; r < kBkCl; ++r) m = max(m, cl.map_shared_rank(&t.cm_local[0], r)[j]);
        t.cm[j] = m;
    }
    __syncthreads();
    float best = 0.0f, veq = 0.0f;                        // pass 2: labels (lib/region.py:88-107)
    int arg = 0, eq = -1;
    if (ok) {
        for (int j = 0; j < K; ++j) {
            const float cmj = key2f(t.cm[j]);
            const float v = iou_plus1(bx, ba, t.gt[j], t.ga[j]);
            if (j == 0 || v > best) { best = v; arg = j; }                     // first max wins
            if (eq < 0 && cmj >= ap.min_pos_iou && v == cmj) { eq = j; veq = v; }   // lowest GT wins
        }
    }
    int64_t* lab = a.tg.labels + (long long)b * ap.out_ld;
    float* oiou = a.tg.out_iou + (long long)b * ap.out_ld;
    TgShared* t0 = cl.map_shared_rank(&t, 0);             // labels / counts are collected in the first CTA
    int64_t out_l = -1;
    if (ok) {
        int l = -1;
        if (best < ap.neg_iou) l = 0;
        if (best >= ap.pos_iou) l = 1;
        int aa = arg;
        float out_v = best;
        if (eq >= 0) { l = 1; aa = eq; out_v = veq; }
        out_l = (l == 1) ? (int64_t)(aa + 1) : (int64_t)l;
        lab[lead + i] = out_l; oiou[lead + i] = out_v;
        t0->fs.lab[lead + i] = (int)out_l;
    }
    {
        const bool is_pos = ok && out_l > 0, is_neg = ok && out_l == 0;
        const unsigned mp = __ballot_sync(0xffffffffu, is_pos), mn = __ballot_sync(0xffffffffu, is_neg);
        if (lane == 0) {
            if (mp) atomicAdd(&t0->cnt[0], __popc(mp));
            if (mn) atomicAdd(&t0->cnt[1], __popc(mn));
        }
        if (a.tg.pos_list) {
            int* plist = a.tg.pos_list + (long long)b * a.tg.pos_cap;
            const int slot = warp_alloc(is_pos, &t0->cnt[2]);
            if (is_pos && slot < a.tg.pos_cap) plist[slot] = lead + i;
        }
    }
    dbg_stamp(p, b, 32 + crank, 7);
    cl.sync();
    dbg_stamp(p, b, 32 + crank, 8);
    if (crank != 0) return;
    if (ap.prepend_gt) {                                  // prepended GT rows (lib/bbox.py:27-29): labels 1..K, IoU 1
        int* plist = a.tg.pos_list ? a.tg.pos_list + (long long)b * a.tg.pos_cap : nullptr;
        for (int j = tid; j < K; j += kBkThreads) {
            lab[j] = j + 1; oiou[j] = 1.0f;
            t.fs.lab[j] = j + 1;
            if (plist) { const int slot = atomicAdd(&t.cnt[2], 1); if (slot < a.tg.pos_cap) plist[slot] = j; }
        }
    }
    __syncthreads();
    if (tid == 0) {
        a.tg.census[4 * b + 0] = t.cnt[0] + (ap.prepend_gt ? K : 0);
        a.tg.census[4 * b + 1] = t.cnt[1];
        a.tg.census[4 * b + 2] = t.cnt[2];
        a.tg.census[4 * b + 3] = 0;
    }
    fused_sample_encode(ap, a.tg.f, t.gt, t.fs, t.cnt[0], t.cnt[1], b);
    dbg_stamp(p, b, 32 + crank, 9);
}

static size_t back_smem(const RpnLaunch& p) { return ((sizeof(BkShared) + 15) & ~(size_t)15) + (size_t)p.sel_per_img * 4 + 16; }

bool rpn_back_applicable(const RpnLaunch& p) {
    if (!knobs().rpn_back || !p.do_nms || p.raw || p.L > kBkCl) return false;
    int kmax = 0;
    for (int l = 0; l < p.L; ++l) kmax = kmax > p.kcap[l] ? kmax : p.kcap[l];
    return kmax <= kBkThreads * kBkRows && back_smem(p) <= 200 * 1024;
}

bool rpn_back_takes_targets(const RpnLaunch& p, const ::b2d_roi_target_args* tg) {
    return tg && rpn_back_applicable(p) && p.out_ld <= kBkCl * kBkThreads && p.out_ld <= kSmallThreads * kSmallBoxes &&
           tg->gt_ld >= 1 && tg->gt_ld <= kGtChunk && tg->max_num >= 1 && tg->max_num <= kSmallThreads;
}

// 1: launched; 0: not applicable (the caller runs the multi-kernel path); anything else: error code
int rpn_back_launch(const RpnLaunch& p, int cut_m, float* props, float* scores, int* count, int* prov,
                    const ::b2d_roi_target_args* tg, cudaStream_t st) {
    if (!rpn_back_applicable(p)) return 0;
    const size_t smem = back_smem(p);
    BackArgs a;
    memset(&a, 0, sizeof(a));
    nms_thr_bounds(p.nms_thr, a.thr, a.thr_lo, a.thr_hi);
    a.use_sweep = knobs().nms_sweep != 0 && p.nms_thr >= 0.05f && p.nms_thr < 1.0f;
    a.prune = 1.0f - 0.9f * p.nms_thr;
    a.M = cut_m;
    a.props = props; a.scores = scores; a.count = count; a.prov = prov;
    if (tg) {
        if (!rpn_back_takes_targets(p, tg)) return 0;
        BackTargets& t = a.tg;
        t.enable = 1;
        t.p.boxes = props; t.p.box_ld = p.out_ld; t.p.box_count = count; t.p.N = p.out_ld;
        t.p.use_pyr = 0; t.p.img_hw = nullptr; t.p.border = 0.0f;
        t.p.gt = tg->gt; t.p.gt_ld = tg->gt_ld; t.p.gt_count = tg->gt_count;
        t.p.pos_iou = tg->pos_iou; t.p.neg_iou = tg->neg_iou; t.p.min_pos_iou = tg->min_pos_iou;
        t.p.prepend_gt = tg->prepend_gt; t.p.out_ld = tg->out_ld;
        t.f.chosen = tg->chosen; t.f.n_chosen = tg->n_chosen; t.f.max_num = tg->max_num; t.f.pos_num = tg->pos_num;
        t.f.seed = tg->seed; t.f.seed_step = tg->seed_step; t.f.gt_label = tg->gt_label;
        t.f.tar_box = tg->tar_box; t.f.tar_gt = tg->tar_gt; t.f.tar_param = tg->tar_param; t.f.tar_label = tg->tar_label;
        t.f.tar_is_gt = tg->tar_is_gt;
        for (int i = 0; i < 4; ++i) { t.f.ms[i] = tg->means[i]; t.f.ms[4 + i] = tg->stds[i]; }
        t.labels = tg->labels; t.out_iou = tg->max_iou; t.census = tg->census; t.pos_list = tg->pos_list; t.pos_cap = tg->pos_cap;
    }
    if (set_dyn_smem(k_rpn_back, smem, "k_rpn_back") != 0) return 0;      // (error text set; the caller falls back to the multi-kernel chain)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(kBkCl, (unsigned)p.B, 1);
    cfg.blockDim = dim3(kBkThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kBkCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = knobs().pdl ? 2 : 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_rpn_back, p, a);
    if (e != cudaSuccess) {
        set_error(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    const int rc = check_launch("rpn_proposals/k_rpn_back");
    return rc == B2D_OK ? 1 : rc;
}

}  // namespace b2d
