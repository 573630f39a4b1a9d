// nms.cu -- K4: bitmask NMS with torchvision.ops.nms *CPU* semantics
// (stable descending score order, areas without +1, suppress iff (double)iou > thr;
// reference call sites lib/heads/rpn_head.py:103, lib/region.py:207, lib/utils.py:220).
//
//   k_nms_sort       (generic entry only) per-segment bitonic sort of (score, index)
//   k_nms_mask_sym   n <= 2048: 64x64 tiles of the upper triangle; column boxes live in
//                    registers, row boxes are staged in shared memory and broadcast.  A tile
//                    yields BOTH orientations of the (bitwise symmetric) relation: the row
//                    words by __ballot_sync and the column words by lane-local accumulation,
//                    so afterwards row j of the matrix holds every box that overlaps box j.
//                    Only NON-ZERO words are stored, and a per-row bitmap `nz` says which.
//   k_nms_scan_fp    n <= 2048: one block per segment.  Greedy NMS is the unique fixed point of
//                    kept_j = !(exists i < j: kept_i and M_ij); a thread per row iterates that
//                    map from kept = all (Jacobi), which converges after (longest suppression
//                    chain + 1) sweeps of a few hundred cycles each -- instead of n dependent
//                    steps.  A row keeps its (few) non-zero words in registers.
//   k_nms_mask_dense / k_nms_scan_seq   16384 >= n > 2048: dense upper-triangle words and the
//                    sequential 64-box-chunk scan (latency-bound; generic entry only).
//
// The fp32 threshold passed in is the largest float <= the caller's double threshold, which
// makes `iou > thr_f` identical to torchvision's CPU comparison `(double)iou > thr` for every
// float iou.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "pipeline.cuh"
#include "nms_common.cuh"

namespace b2d {

struct NmsSegs {
    int L;                              // levels per image (1 for the generic entry)
    int lv0, lvn;                       // levels covered by this launch
    long long box_per_img, box_off[kMaxLevels];
    long long mask_per_img, mask_off[kMaxLevels];
    int wp[kMaxLevels];                 // mask row pitch in words
    const float4* boxes;                // score-sorted boxes
    const int* counts;                  // int[S]
    uint64_t* mask;
    uint32_t* nz;                       // sym path: per row, bit w set iff word w of the row is non-zero (laid out like boxes)
    float thr, thr_lo, thr_hi;          // thr_lo/hi: decisive bounds that avoid the divide (see suppresses)
    // fallback pass of the score-cut scheme (rpn_nms_launch): skip every image whose first pass already produced
    // `need` survivors (chk = per-segment survivor counts of the first pass, L per image)
    const int* chk; int need;
    int* force_fb;                      // per image: set by the sweep kernel of the first pass when it gives up
};

// true if the first (score-cut) pass of image b produced enough survivors: nothing left to do for this image
__device__ __forceinline__ bool cut_pass_sufficient(const NmsSegs& s, int b) {
    if (!s.chk) return false;
    if (s.force_fb && s.force_fb[b]) return false;
    int tot = 0;
    for (int l = 0; l < s.L; ++l) tot += s.chk[b * s.L + l];
    return tot >= s.need;
}

__device__ __forceinline__ void tile_of(int t, int wmax, int& rb, int& cb) {
    rb = 0;
    while (t >= wmax - rb) { t -= wmax - rb; ++rb; }
    cb = rb + t;
}

// ---- symmetric, sparse-output mask (n <= 2048) ------------------------------------------------
struct MaskSmem {
    float4 row[64];
    float ra[64];
    uint32_t cw[2][64];
};

// one 64 x 64 tile pair (rb <= cb) of segment (seg, b, l); all 64 threads of the CTA
__device__ __forceinline__ void mask_tile(const NmsSegs& s, MaskSmem& sm, int seg, int b, int l, int rb, int cb) {
    float4* s_row = sm.row;
    float* s_ra = sm.ra;
    uint32_t (*s_cw)[64] = sm.cw;
    const int n = s.counts[seg];
    const int r0 = rb * 64, c0 = cb * 64;
    if (r0 >= n || c0 >= n) return;
    const long long base = (long long)b * s.box_per_img + s.box_off[l];
    const float4* boxes = s.boxes + base;
    uint32_t* nz = s.nz + base;
    uint64_t* mask = s.mask + (long long)b * s.mask_per_img + s.mask_off[l];
    const int wp = s.wp[l];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float4 pad = pad_box();
    {
        const float4 rbx = (r0 + tid < n) ? boxes[r0 + tid] : pad;
        s_row[tid] = rbx; s_ra[tid] = area_of(rbx);
    }
    const int j0 = c0 + lane, j1 = c0 + 32 + lane;
    const float4 cb0 = j0 < n ? boxes[j0] : pad, cb1 = j1 < n ? boxes[j1] : pad;
    const float a0 = area_of(cb0), a1 = area_of(cb1);
    // all 128 boxes of the tile pair well-formed (the normal case) -> cheaper decision
    const bool wf = __syncthreads_and(well_formed(s_row[tid]) && well_formed(cb0) && well_formed(cb1));
    uint64_t myword = 0;
    uint32_t cw0 = 0, cw1 = 0;
    const float4* rowp = s_row + w * 32;
    const float* rap = s_ra + w * 32;
    if (wf) {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
            const float4 rbx = rowp[rr];
            const float ra = rap[rr];
            const bool p0 = suppresses_wf(rbx, ra, cb0, a0, s.thr, s.thr_lo, s.thr_hi);
            const bool p1 = suppresses_wf(rbx, ra, cb1, a1, s.thr, s.thr_lo, s.thr_hi);
            const unsigned lo = __ballot_sync(0xffffffffu, p0), hi = __ballot_sync(0xffffffffu, p1);
            if (lane == rr) myword = ((uint64_t)hi << 32) | lo;
            cw0 |= p0 ? (1u << rr) : 0u;
            cw1 |= p1 ? (1u << rr) : 0u;
        }
    } else {
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            const float4 rbx = rowp[rr];
            const float ra = rap[rr];
            const bool p0 = suppresses(rbx, ra, cb0, a0, s.thr, s.thr_lo, s.thr_hi);
            const bool p1 = suppresses(rbx, ra, cb1, a1, s.thr, s.thr_lo, s.thr_hi);
            const unsigned lo = __ballot_sync(0xffffffffu, p0), hi = __ballot_sync(0xffffffffu, p1);
            if (lane == rr) myword = ((uint64_t)hi << 32) | lo;
            cw0 |= p0 ? (1u << rr) : 0u;
            cw1 |= p1 ? (1u << rr) : 0u;
        }
    }
    if (rb == cb) myword &= ~(1ull << (w * 32 + lane));      // a box does not suppress itself
    const int row = r0 + w * 32 + lane;
    if (row < n && myword) {
        mask[(long long)row * wp + cb] = myword;
        atomicOr(&nz[row], 1u << cb);
    }
    if (rb != cb) {                                          // transposed words: rows of block cb, word rb
        s_cw[w][lane] = cw0; s_cw[w][32 + lane] = cw1;
        __syncthreads();
        const uint64_t cw = (uint64_t)s_cw[0][tid] | ((uint64_t)s_cw[1][tid] << 32);
        const int j = c0 + tid;
        if (j < n && cw) {
            mask[(long long)j * wp + rb] = cw;
            atomicOr(&nz[j], 1u << rb);
        }
    }
}

__global__ void __launch_bounds__(64) k_nms_mask_sym(NmsSegs s, int wmax) {
    __shared__ MaskSmem sm;
    int seg, b, l;
    seg_of(s.lv0, s.lvn, s.L, blockIdx.y, seg, b, l);
    int rb, cb;
    tile_of(blockIdx.x, wmax, rb, cb);
    mask_tile(s, sm, seg, b, l, rb, cb);
}

// Fallback pass of the score-cut scheme: a small persistent grid walks (segment, tile) work items and skips the
// images whose first pass was sufficient -- when every image was (the normal case) the launch costs ~2 us instead
// of the ~10 us that 21 000 immediately-exiting CTAs of the tile grid would.
__global__ void __launch_bounds__(64) k_nms_mask_sym_fb(NmsSegs s, int wmax, int S) {
    __shared__ MaskSmem sm;
    __shared__ int s_todo[64];                          // per image: 1 if the first pass fell short
    __shared__ int s_any;
    const int nimg = S / s.lvn;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < nimg; b += blockDim.x) {
        const int todo = cut_pass_sufficient(s, b) ? 0 : 1;
        if (b < 64) s_todo[b] = todo;
        if (todo) s_any = 1;
    }
    __syncthreads();
    if (!s_any) return;                                 // the normal case: nothing to do for any image
    const int ntiles = wmax * (wmax + 1) / 2;
    for (int sl = 0; sl < S; ++sl) {
        int seg, b, l;
        seg_of(s.lv0, s.lvn, s.L, sl, seg, b, l);
        if (b < 64 ? !s_todo[b] : cut_pass_sufficient(s, b)) continue;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            int rb, cb;
            tile_of(t, wmax, rb, cb);
            mask_tile(s, sm, seg, b, l, rb, cb);
            __syncthreads();                                                 // sm is reused by the next tile
        }
    }
}

// ---- sweep mask (first pass of the score-cut scheme) -----------------------------------------------------------
// iou > thr forces the x-extents to overlap: with i the box whose x1 is smaller, inter > thr * u >= thr * area_i
// gives x1_j < x2_i - thr * w_i.  A CTA buckets the boxes of its segment by x1 (256 uniform cells, counting sort in
// shared memory) and a box only meets the boxes whose x1 lies in [x1_i, x1_i + (1 - 0.9 thr) w_i + slack] -- a few
// dozen candidates instead of 2000 on config 2.  The 0.9 and the absolute slack are far larger than any fp32
// rounding of the quantities involved, so no pair that `suppresses` accepts is skipped; the decision itself is the
// same function on the same operands (it is symmetric in the two boxes).  kSweepSlices CTAs share a segment: each
// repeats the (cheap) bucketing and sweeps the boxes i = slice, slice + kSweepSlices, ..., a warp per box with the
// lanes striding its candidate range.  Bits are OR-ed into the mask, which the launcher zeroes at the start of the step.
// Gives up -- force_fb[image] = 1, the full dense pass then runs for the image -- on malformed boxes or when the
// x1 distribution is so concentrated that the cells stop pruning.
constexpr int kSweepThreads = 256, kSweepCells = 256, kSweepCap = 2048, kSweepSlices = 16;
__global__ void __launch_bounds__(1024) k_nms_sweep(NmsSegs s, float prune) {
    const int nthr = blockDim.x, nslice = gridDim.y;      // launch shape: kSweepThreads x kSweepSlices (dev knobs below)
    __shared__ float4 s_box[kSweepCap];
    __shared__ uint16_t s_ord[kSweepCap];
    __shared__ int s_start[kSweepCells + 1], s_cur[kSweepCells];
    __shared__ float s_mn[32], s_mx[32];
    __shared__ int s_ok, s_heavy;
    int seg, b, l;
    seg_of(s.lv0, s.lvn, s.L, blockIdx.x, seg, b, l);
    const int n = s.counts[seg];
    if (n <= 0) return;
    const long long base = (long long)b * s.box_per_img + s.box_off[l];
    const float4* boxes = s.boxes + base;
    uint32_t* nz = s.nz + base;
    uint64_t* mask = s.mask + (long long)b * s.mask_per_img + s.mask_off[l];
    const int wp = s.wp[l];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_ok = 1; s_heavy = 0; }
    for (int c = tid; c <= kSweepCells; c += nthr) s_start[c] = 0;
    float mn = INFINITY, mx = -INFINITY;
    bool ok = true;
    for (int i = tid; i < n; i += nthr) {
        const float4 bx = boxes[i];
        s_box[i] = bx;
        ok = ok && well_formed(bx);
        mn = fminf(mn, bx.x); mx = fmaxf(mx, bx.x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (!ok) s_ok = 0;
    mn = s_mn[0]; mx = s_mx[0];
    for (int w = 1; w < nthr / 32; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
    const float inv = (mx > mn) ? (float)kSweepCells / (mx - mn) : 0.0f;
    auto cell = [&](float x) { return min(kSweepCells - 1, max(0, (int)((x - mn) * inv))); };   // monotone in x
    __syncthreads();
    if (!s_ok) {                                                         // malformed box: the dense pass decides
        if (tid == 0 && blockIdx.y == 0) s.force_fb[b] = 1;
        return;
    }
    for (int i = tid; i < n; i += nthr) atomicAdd(&s_start[cell(s_box[i].x) + 1], 1);
    __syncthreads();
    if (warp == 0) {                                                     // inclusive scan of the 256 counts, 8 per lane
        int c[8], sum = 0;
        long long sq = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { c[q] = s_start[lane * 8 + q + 1]; sum += c[q]; sq += (long long)c[q] * c[q]; }
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
        for (int q = 0; q < 8; ++q) { s_cur[lane * 8 + q] = run; run += c[q]; s_start[lane * 8 + q + 1] = run; }
        if (lane == 0 && sq > (long long)n * 192) s_heavy = 1;           // cells no longer prune: ~n^2 / 10 pair tests
    }
    __syncthreads();
    if (s_heavy) {
        if (tid == 0 && blockIdx.y == 0) s.force_fb[b] = 1;
        return;
    }
    for (int i = tid; i < n; i += nthr) s_ord[atomicAdd(&s_cur[cell(s_box[i].x)], 1)] = (uint16_t)i;
    __syncthreads();
    // a warp per box, lanes stride its candidate range (the ranges are heavy-tailed: mean 40, max ~600 on config 2)
    for (int i = blockIdx.y + nslice * warp; i < n; i += nslice * (nthr / 32)) {
        const float4 bi = s_box[i];
        const float wi = bi.z - bi.x;
        if (!(wi > 0.0f) || !(bi.w - bi.y > 0.0f)) continue;              // empty box: inter == 0 with everything
        const float ai = area_of(bi);
        const float xhi = bi.x + prune * wi + 1.0e-4f * (fabsf(bi.z) + 1.0f);
        const int k1 = s_start[cell(xhi) + 1];
        for (int k = s_start[cell(bi.x)] + lane; k < k1; k += 32) {
            const int j = s_ord[k];
            const float4 bj = s_box[j];
            if (bj.x < bi.x || (bj.x == bi.x && j <= i) || bj.x > xhi) continue;   // every unordered pair once
            if (!(bj.y < bi.w && bi.y < bj.w)) continue;                           // no y overlap: inter == 0
            if (!suppresses_wf(bi, ai, bj, area_of(bj), s.thr, s.thr_lo, s.thr_hi)) continue;
            atomicOr(reinterpret_cast<unsigned long long*>(&mask[(long long)i * wp + (j >> 6)]), 1ull << (j & 63));
            atomicOr(reinterpret_cast<unsigned long long*>(&mask[(long long)j * wp + (i >> 6)]), 1ull << (i & 63));
            atomicOr(&nz[i], 1u << (j >> 6));
            atomicOr(&nz[j], 1u << (i >> 6));
        }
    }
}

// The tile grid of k_nms_mask_sym is sized by the level capacity (528 tiles per segment at 2000 boxes) while the cut
// prefixes need a fraction of them, and a CTA that exits at once still costs ~2.6 ns of CTA dispatch.  Here a fixed
// grid walks the (segment, tile) items that exist: every CTA builds the per-segment tile prefix from the device-side
// counts (S <= kItemSegs) and takes items blockIdx.x, blockIdx.x + gridDim.x, ...
constexpr int kItemSegs = 1024;
// (A 256-thread variant of the tile -- 8 warps x 8 rows, column words assembled from per-warp bytes -- and grids of
// 8..32 CTAs per SM were measured: 31-38 us for the ~5000 tiles of config 2 either way; the pass is bound by the
// pair arithmetic itself, ~25 instructions per pair at IPC ~1.5.)
constexpr int kItemCtasPerSm = 16;
__global__ void __launch_bounds__(64, kItemCtasPerSm) k_nms_mask_sym_items(NmsSegs s, int S) {
    __shared__ MaskSmem sm;
    __shared__ int s_first[kItemSegs + 1];              // first item of segment slot sl (exclusive prefix of the tile counts)
    for (int sl = threadIdx.x; sl < S; sl += blockDim.x) {
        int seg, b, l;
        seg_of(s.lv0, s.lvn, s.L, sl, seg, b, l);
        const int w = (s.counts[seg] + 63) >> 6;
        s_first[sl + 1] = w * (w + 1) / 2;
    }
    __syncthreads();
    if (threadIdx.x < 32) {                             // inclusive scan, one warp, 32 slots per round
        int carry = 0;
        for (int base = 0; base < S; base += 32) {
            const int i = base + (int)threadIdx.x;
            int v = i < S ? s_first[i + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if ((int)threadIdx.x >= o) v += t;
            }
            if (i < S) s_first[i + 1] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (threadIdx.x == 0) s_first[0] = 0;
    }
    __syncthreads();
    const int total = s_first[S];
    for (int it = blockIdx.x; it < total; it += gridDim.x) {
        int lo = 0, hi = S - 1;                         // last slot with s_first[slot] <= it
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_first[mid] <= it) lo = mid; else hi = mid - 1;
        }
        int seg, b, l;
        seg_of(s.lv0, s.lvn, s.L, lo, seg, b, l);
        const int w = (s.counts[seg] + 63) >> 6;
        int rb, cb;
        tile_of(it - s_first[lo], w, rb, cb);
        mask_tile(s, sm, seg, b, l, rb, cb);
        __syncthreads();                                // sm is reused by the next item
    }
}

struct ScanOut {
    // RPN pipeline: survivors are compacted (score order) from the sel_* arrays into kept_*
    const uint32_t* src_key; const int* src_idx; float4* dst_box; uint32_t* dst_key; int* dst_idx;
    // generic entry: int64 original indices
    int64_t* keep64; const int* sorted_idx; long long keep_ld;
    int* keep_count;
    int* keep_count_aux;                // optional second copy of keep_count (first pass of the score-cut scheme)
};

__device__ __forceinline__ void emit_kept(const NmsSegs& s, const ScanOut& o, int seg, long long base, int r, int pos) {
    if (o.keep64) {
        o.keep64[(long long)seg * o.keep_ld + pos] = (int64_t)o.sorted_idx[(long long)seg * o.keep_ld + r];
    } else {
        o.dst_box[base + pos] = s.boxes[base + r];
        o.dst_key[base + pos] = o.src_key[base + r];
        o.dst_idx[base + pos] = o.src_idx[base + r];
    }
}

// ---- fixed-point scan (n <= 2048) ---------------------------------------------------------------
constexpr int kFpThreads = 1024;
constexpr int kFpRows = 2;                           // rows per thread: 2048 boxes
constexpr int kFpEntries = 6;                        // non-zero words of a row kept in registers

__global__ void __launch_bounds__(kFpThreads) k_nms_scan_fp(NmsSegs s, int max_keep, ScanOut o) {
    __shared__ uint32_t s_k[2][64];                  // kept bits, ping-pong (2048 bits each)
    __shared__ int s_prefix[65];
    int seg, b, l;
    seg_of(s.lv0, s.lvn, s.L, blockIdx.x, seg, b, l);
    if (cut_pass_sufficient(s, b)) return;               // fallback pass: this image is already complete
    const int n = s.counts[seg];
    const long long base = (long long)b * s.box_per_img + s.box_off[l];
    const uint64_t* mask = s.mask + (long long)b * s.mask_per_img + s.mask_off[l];
    const uint32_t* nz = s.nz + base;
    const int wp = s.wp[l];
    const int tid = threadIdx.x, lane = tid & 31;
    // ---- my rows: the words (earlier boxes only) that can suppress them
    uint64_t eb[kFpRows][kFpEntries];
    uint32_t ew[kFpRows], over[kFpRows];
#pragma unroll
    for (int q = 0; q < kFpRows; ++q) {
        const int row = tid + q * kFpThreads;
        uint32_t m = 0;
        if (row < n) {
            const int dw = row >> 6;
            m = nz[row] & (dw == 31 ? 0xffffffffu : ((2u << dw) - 1u));
        }
        ew[q] = 0;
#pragma unroll
        for (int e = 0; e < kFpEntries; ++e) {
            uint64_t word = 0ull;
            if (m) {
                const int w = __ffs(m) - 1;
                m &= m - 1;
                word = mask[(long long)row * wp + w];
                if (w == (row >> 6)) word &= (1ull << (row & 63)) - 1ull;
                ew[q] |= (uint32_t)w << (5 * e);
            }
            eb[q][e] = word;
        }
        over[q] = m;                                 // rare: more than kFpEntries candidate words
    }
    if (tid < 64) {
        const int lo = tid * 32;
        s_k[0][tid] = n >= lo + 32 ? 0xffffffffu : (n > lo ? ((1u << (n - lo)) - 1u) : 0u);
    }
    __syncthreads();
    int it = 0;
    for (;; ++it) {
        const uint32_t* cur = s_k[it & 1];
        uint32_t* nxt = s_k[(it & 1) ^ 1];
        bool changed = false;
#pragma unroll
        for (int q = 0; q < kFpRows; ++q) {
            const int row = tid + q * kFpThreads;
            bool sup = false;
#pragma unroll
            for (int e = 0; e < kFpEntries; ++e) {
                if (eb[q][e]) {
                    const int w = (ew[q] >> (5 * e)) & 31;
                    const uint64_t k = (uint64_t)cur[2 * w] | ((uint64_t)cur[2 * w + 1] << 32);
                    sup |= (k & eb[q][e]) != 0ull;
                }
            }
            uint32_t m = over[q];
            while (m && !sup) {
                const int w = __ffs(m) - 1;
                m &= m - 1;
                uint64_t word = mask[(long long)row * wp + w];
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
    const uint32_t* fin = s_k[(it & 1) ^ 1];
    if (tid == 0) {
        int run = 0;
        for (int w = 0; w < 64; ++w) { s_prefix[w] = run; run += __popc(fin[w]); }
        s_prefix[64] = run;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kFpRows; ++q) {
        const int row = tid + q * kFpThreads;
        const uint32_t kw = fin[row >> 5];
        if (row < n && ((kw >> lane) & 1u)) {
            const int pos = s_prefix[row >> 5] + __popc(kw & ((1u << lane) - 1u));
            if (max_keep <= 0 || pos < max_keep) emit_kept(s, o, seg, base, row, pos);
        }
    }
    if (tid == 0) {
        const int kc = (max_keep > 0 && s_prefix[64] > max_keep) ? max_keep : s_prefix[64];
        o.keep_count[seg] = kc;
        if (o.keep_count_aux) o.keep_count_aux[seg] = kc;
    }
}

// ---- dense upper-triangle mask + sequential scan (2048 < n <= 16384; generic entry) ------------
// Off-diagonal tiles hold bit j of row i iff box i suppresses box j (j > i always); DIAGONAL
// tiles hold the full symmetric relation, so (word_j & lower_bits(j)) is the set of earlier
// boxes of the same chunk that suppress j.
__global__ void __launch_bounds__(64) k_nms_mask_dense(NmsSegs s, int wmax) {
    __shared__ float4 s_row[64];
    int seg, b, l;
    seg_of(s.lv0, s.lvn, s.L, blockIdx.y, seg, b, l);
    const int n = s.counts[seg];
    int rb, cb;
    tile_of(blockIdx.x, wmax, rb, cb);
    const int r0 = rb * 64, c0 = cb * 64;
    if (r0 >= n || c0 >= n) return;
    const float4* boxes = s.boxes + (long long)b * s.box_per_img + s.box_off[l];
    uint64_t* mask = s.mask + (long long)b * s.mask_per_img + s.mask_off[l];
    const int wp = s.wp[l];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float4 pad = pad_box();
    s_row[threadIdx.x] = (r0 + (int)threadIdx.x < n) ? boxes[r0 + threadIdx.x] : pad;
    const int j0 = c0 + lane, j1 = c0 + 32 + lane;
    const float4 cb0 = j0 < n ? boxes[j0] : pad, cb1 = j1 < n ? boxes[j1] : pad;
    const float a0 = area_of(cb0), a1 = area_of(cb1);
    __syncthreads();
    uint64_t myword = 0;
    const int rows = min(32, n - (r0 + w * 32));
    for (int rr = 0; rr < rows; ++rr) {
        const int i = r0 + w * 32 + rr;
        const float4 rbx = s_row[w * 32 + rr];
        const float ra = area_of(rbx);
        const bool p0 = (j0 < n) && (j0 != i) && suppresses(rbx, ra, cb0, a0, s.thr, s.thr_lo, s.thr_hi);
        const bool p1 = (j1 < n) && (j1 != i) && suppresses(rbx, ra, cb1, a1, s.thr, s.thr_lo, s.thr_hi);
        const unsigned lo = __ballot_sync(0xffffffffu, p0), hi = __ballot_sync(0xffffffffu, p1);
        if (lane == rr) myword = ((uint64_t)hi << 32) | lo;
    }
    const int row = r0 + w * 32 + lane;
    if (lane < rows) mask[(long long)row * wp + cb] = myword;
}

// One block (256 threads) per segment.  Per 64-box chunk c: warp 0 finds the kept boxes of the
// chunk as the fixed point of kept_j = alive_j & !(any kept i < j suppresses j) from the
// diagonal word (two ballots per round), then every thread ORs the rows of the kept boxes into
// its words w > c of the removed-set.
constexpr int kScanThreads = 256;

__global__ void __launch_bounds__(kScanThreads) k_nms_scan_seq(NmsSegs s, int max_keep, ScanOut o) {
    __shared__ uint64_t s_removed[kSortCap / 64];
    __shared__ uint64_t s_keepw[kSortCap / 64];
    __shared__ int s_prefix[kSortCap / 64];
    __shared__ uint64_t s_keep;
    int seg, b, l;
    seg_of(s.lv0, s.lvn, s.L, blockIdx.x, seg, b, l);
    const int n = s.counts[seg];
    const int W = (n + 63) / 64;
    const uint64_t* mask = s.mask + (long long)b * s.mask_per_img + s.mask_off[l];
    const int wp = s.wp[l];
    const int lane = threadIdx.x & 31;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        uint64_t r = 0;
        if (w == W - 1 && (n & 63)) r = ~0ull << (n & 63);   // rows past n are "removed"
        s_removed[w] = r; s_keepw[w] = 0;
    }
    uint64_t nd0 = 0, nd1 = 0;                       // diagonal words of the next chunk (warp 0)
    auto load_diag = [&](int c) {
        const int row0 = c * 64 + lane, row1 = row0 + 32;
        nd0 = (c < W && row0 < n) ? mask[(long long)row0 * wp + c] : 0ull;
        nd1 = (c < W && row1 < n) ? mask[(long long)row1 * wp + c] : 0ull;
    };
    if (threadIdx.x < 32) load_diag(0);
    __syncthreads();
    int kept_total = 0;
    for (int c = 0; c < W; ++c) {
        if (threadIdx.x < 32) {
            const uint64_t low0 = nd0 & ((1ull << lane) - 1ull);
            const uint64_t low1 = nd1 & ((1ull << (lane + 32)) - 1ull);
            load_diag(c + 1);                          // in flight during the fixed-point rounds
            const uint64_t alive = ~s_removed[c];
            const bool al0 = (alive >> lane) & 1ull, al1 = (alive >> (lane + 32)) & 1ull;
            uint64_t keep = alive;
            for (int round = 0; round < 65; ++round) {
                const bool k0 = al0 && ((low0 & keep) == 0ull), k1 = al1 && ((low1 & keep) == 0ull);
                const uint64_t nk = (uint64_t)__ballot_sync(0xffffffffu, k0) |
                                    ((uint64_t)__ballot_sync(0xffffffffu, k1) << 32);
                if (nk == keep) break;
                keep = nk;
            }
            if (max_keep > 0 && kept_total + __popcll(keep) > max_keep) {
                int room = max_keep - kept_total;        // keep only the first `room` set bits
                uint64_t trimmed = 0, rest = keep;
                while (room-- > 0) { const uint64_t low = rest & (0ull - rest); trimmed |= low; rest ^= low; }
                keep = trimmed;
            }
            if (lane == 0) { s_keep = keep; s_keepw[c] = keep; }
        }
        __syncthreads();
        const uint64_t keep = s_keep;
        kept_total += __popcll(keep);
        const bool done = (max_keep > 0 && kept_total >= max_keep);
        if (!done) {
            for (int w = c + 1 + threadIdx.x; w < W; w += blockDim.x) {
                uint64_t acc = s_removed[w], rest = keep;
                while (rest) {                       // batches of 4 independent loads
                    uint64_t v[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (rest) {
                            const int bit = __ffsll((long long)rest) - 1;
                            rest &= rest - 1;
                            v[q] = mask[(long long)(c * 64 + bit) * wp + w];
                        }
                    }
                    acc |= (v[0] | v[1]) | (v[2] | v[3]);
                }
                s_removed[w] = acc;
            }
        }
        __syncthreads();
        if (done) break;
    }
    // ordered write-out: position = kept boxes before this one (prefix over words + popc)
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < W; ++w) { s_prefix[w] = run; run += __popcll(s_keepw[w]); }
    }
    __syncthreads();
    const long long base = (long long)b * s.box_per_img + s.box_off[l];
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int w = r >> 6, bit = r & 63;
        const uint64_t kw = s_keepw[w];
        if (!((kw >> bit) & 1ull)) continue;
        emit_kept(s, o, seg, base, r, s_prefix[w] + __popcll(kw & ((1ull << bit) - 1ull)));
    }
    if (threadIdx.x == 0) o.keep_count[seg] = kept_total;
}

static void set_thr(NmsSegs& s, float thr) { nms_thr_bounds(thr, s.thr, s.thr_lo, s.thr_hi); }

// mask + scan of S segments of at most n_max boxes; s.nz must be zeroed by the caller (sym path)
static void launch_mask_scan(const NmsSegs& s, int S, int n_max, int max_keep, const ScanOut& o, cudaStream_t st,
                             bool fallback = false, bool items = false) {
    const int wmax = (n_max + 63) / 64 > 0 ? (n_max + 63) / 64 : 1;
    dim3 grid(wmax * (wmax + 1) / 2, S);
    if (n_max <= kFpThreads * kFpRows) {
        if (fallback) k_nms_mask_sym_fb<<<148 * 4, 64, 0, st>>>(s, wmax, S);
        else if (items && S <= kItemSegs) {
            // sweep: needs a pruning bound (0 < thr < 1), a zeroed mask and a place to report "gave up"
            const bool sweep = s.force_fb != nullptr;
            if (sweep) {
                int nt = knobs().sweep_t ? knobs().sweep_t : kSweepThreads, ng = knobs().sweep_g ? knobs().sweep_g : kSweepSlices;
                if (nt < 32 || nt > 1024 || (nt & 31)) nt = kSweepThreads;       // dev knobs: ignore unusable values
                if (ng < 1 || ng > 64) ng = kSweepSlices;
                k_nms_sweep<<<dim3(S, ng), nt, 0, st>>>(s, 1.0f - 0.9f * s.thr);
            }
            else k_nms_mask_sym_items<<<148 * kItemCtasPerSm, 64, 0, st>>>(s, S);
        }
        else k_nms_mask_sym<<<grid, 64, 0, st>>>(s, wmax);
        k_nms_scan_fp<<<S, kFpThreads, 0, st>>>(s, max_keep, o);
    } else {
        k_nms_mask_dense<<<grid, 64, 0, st>>>(s, wmax);
        k_nms_scan_seq<<<S, kScanThreads, 0, st>>>(s, max_keep, o);
    }
}

// generic entry: sort (score desc, index asc) and gather boxes into score order
__global__ void __launch_bounds__(kSelThreads) k_nms_sort(float4* __restrict__ sorted_box, int* __restrict__ sorted_idx,
                                                          const float4* __restrict__ boxes,
                                                          const float* __restrict__ scores, long long n_ld,
                                                          const int* __restrict__ counts, long long n, int presorted) {
    extern __shared__ __align__(16) uint64_t s_buf[];
    const int s = blockIdx.x;
    const int m = counts ? counts[s] : (int)n;
    if (!presorted) {
        int p2 = 1;
        while (p2 < m) p2 <<= 1;
        for (int i = threadIdx.x; i < p2; i += blockDim.x)
            s_buf[i] = i < m ? make_comp(f2key(scores[(long long)s * n_ld + i]), (uint32_t)i) : 0ull;
        __syncthreads();
        bitonic_sort_desc(s_buf, p2);
    }
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const int i = presorted ? r : (int)comp_idx(s_buf[r]);
        sorted_box[(long long)s * n_ld + r] = boxes[(long long)s * n_ld + i];
        sorted_idx[(long long)s * n_ld + r] = i;
    }
}

__global__ void k_fill_i32(int* p, int v, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) p[i] = v;
}

// true if pass 1 of the score-cut scheme will use the sweep kernel (the launcher then zeroes the mask for the step)
bool rpn_nms_sweep_active(const RpnLaunch& p) {
    int n_max = 1;
    for (int l = 0; l < p.L; ++l) n_max = max(n_max, p.kcap[l]);
    return knobs().nms_sweep != 0 && p.nms_thr >= 0.05f && p.nms_thr < 1.0f && n_max <= kSweepCap &&
           p.B * p.L <= kItemSegs;
}

int rpn_nms_launch(const RpnLaunch& p, cudaStream_t st) {
    NmsSegs s;
    memset(&s, 0, sizeof(s));
    s.L = p.L; s.lv0 = p.lv0; s.lvn = p.lvn; s.box_per_img = p.sel_per_img; s.mask_per_img = p.mask_per_img;
    int wmax = 1;
    for (int l = 0; l < p.L; ++l) {
        s.box_off[l] = p.sel_off[l]; s.mask_off[l] = p.mask_off[l];
        s.wp[l] = (p.kcap[l] + 63) / 64;
        wmax = max(wmax, s.wp[l]);
    }
    s.boxes = p.sel_box; s.counts = p.sel_count; s.mask = p.mask; s.nz = p.nz;   // nz zeroed with the workspace head
    if (p.nms_phase == 1) { s.counts = p.n_cut; s.force_fb = rpn_nms_sweep_active(p) ? p.force_fb : nullptr; }
    if (p.nms_phase == 2) { s.chk = p.keep1; s.need = p.max_num; s.force_fb = p.force_fb; }
    set_thr(s, p.nms_thr);
    const int S = p.B * p.lvn;
    int n_max = 1;
    for (int l = p.lv0; l < p.lv0 + p.lvn; ++l) n_max = max(n_max, p.kcap[l]);
    (void)wmax;
    ScanOut o;
    memset(&o, 0, sizeof(o));
    o.src_key = p.sel_key; o.src_idx = p.sel_idx; o.dst_box = p.kept_box; o.dst_key = p.kept_key; o.dst_idx = p.kept_idx;
    o.keep_count = p.keep_count;
    if (p.nms_phase == 1) o.keep_count_aux = p.keep1;
    launch_mask_scan(s, S, n_max, p.post_nms, o, st, p.nms_phase == 2, p.nms_phase == 1);
    return check_launch("rpn_nms");
}

// ---- score cut (exact early termination of the per-level greedy NMS under a global top-k) ---------------------
// RPNHead.predict_single_image keeps, after the per-level NMS, only the max_num best survivors of all levels
// (lib/heads/rpn_head.py:112-118).  Whether a box survives depends only on HIGHER-scored boxes of its level, so the
// NMS of the M globally best selected boxes (all levels, M >= max_num) already yields the final result whenever it
// leaves >= max_num survivors: every box outside the M has a lower score than all of them.  k_nms_cut finds the key
// of the M-th best box of an image to 16 bits (radix select over the <= 5 x 2048 keys in shared memory) and the
// number of boxes per level at or above it (ties included); pass 1 runs mask + scan on those prefixes only, pass 2
// (the full NMS) runs only for images whose pass 1 fell short.
__global__ void __launch_bounds__(1024) k_nms_cut(RpnLaunch p, int M, int zero_ctas) {
    extern __shared__ uint32_t s_keys[];                 // sel_per_img entries, level l at sel_off[l]
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_prefix, s_need;
    __shared__ int s_cnt[kMaxLevels];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b >= p.B) {
        // CTAs past the images clear the suppression mask for the sweep kernel of pass 1 (it ORs bits into it) while
        // the B selecting CTAs are in their latency-bound radix passes: 13 MB at config 2, off the critical path.
        // (A cudaMemsetAsync on a level chain took 18 us next to the RPN-target kernels and delayed k_select.)
        const long long words = (long long)p.B * p.mask_per_img;
        uint4* m16 = reinterpret_cast<uint4*>(p.mask);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (long long i = (long long)(b - p.B) * 1024 + tid; i < words / 2; i += (long long)zero_ctas * 1024) m16[i] = z;
        if ((words & 1) && b == p.B && tid == 0) p.mask[words - 1] = 0ull;
        return;
    }
    int total = 0;
    {   // all levels' keys requested before the first use (<= 2 per level and thread: kcap <= 2048)
        uint32_t v[kMaxLevels][2];
        int nl[kMaxLevels];
#pragma unroll
        for (int l = 0; l < kMaxLevels; ++l) {
            nl[l] = l < p.L ? p.sel_count[b * p.L + l] : 0;
            total += nl[l];
        }
#pragma unroll
        for (int l = 0; l < kMaxLevels; ++l) {
            const uint32_t* src = p.sel_key + (long long)b * p.sel_per_img + p.sel_off[l < p.L ? l : 0];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = tid + r * 1024;
                v[l][r] = i < nl[l] ? src[i] : 0u;
            }
        }
#pragma unroll
        for (int l = 0; l < kMaxLevels; ++l)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = tid + r * 1024;
                if (i < nl[l]) s_keys[p.sel_off[l] + i] = v[l][r];
            }
    }
    if (total <= M) {                                    // uniform over the block
        if (tid < p.L) p.n_cut[b * p.L + tid] = p.sel_count[b * p.L + tid];
        return;
    }
    if (tid == 0) { s_prefix = 0u; s_need = (uint32_t)M; }
    __syncthreads();
    // Two radix passes (the 16 leading key bits) are enough: any threshold gives an exact result, a coarser one
    // only admits a few more boxes than M (all keys whose leading bits tie with the M-th best).
    for (int shift = 24; shift >= 16; shift -= 8) {
        for (int t = tid; t < 256; t += blockDim.x) s_hist[t] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
        for (int l = 0; l < p.L; ++l) {
            const int n = p.sel_count[b * p.L + l];
            for (int i = tid; i < n; i += blockDim.x) {
                const uint32_t k = s_keys[p.sel_off[l] + i];
                // warp-aggregated: the keys of a score-sorted list share their leading bytes, a plain shared-memory
                // atomic per key serialises on one or two bins (25 us for 8819 keys, timeline r1)
                const bool in = (k & himask) == prefix;
                const uint32_t bin = in ? ((k >> shift) & 255u) : 256u;
                const unsigned peers = __match_any_sync(__activemask(), bin);
                if (in && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        if (tid < 32) {                                  // digit = largest d with sum_{bin >= d} hist >= need (warp suffix scan)
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = s_hist[tid * 8 + q]; sum += loc[q]; }
            uint32_t v = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(0xffffffffu, v, o);
                if (tid + o < 32) v += t;
            }
            const uint32_t above = v - sum, need = s_need;
            if (above < need && above + sum >= need) {
                uint32_t run = above;
#pragma unroll
                for (int q = 7; q >= 0; --q) {
                    if (run + loc[q] >= need) {
                        s_need = need - run;
                        s_prefix = prefix | ((uint32_t)(tid * 8 + q) << shift);
                        break;
                    }
                    run += loc[q];
                }
            }
        }
        __syncthreads();
    }
    const uint32_t tau = s_prefix;                       // leading 16 bits of the key of the M-th best box, rest 0
    if (tid < kMaxLevels) s_cnt[tid] = 0;
    __syncthreads();
    for (int l = 0; l < p.L; ++l) {
        const int n = p.sel_count[b * p.L + l];
        int c = 0;
        for (int i = tid; i < n; i += blockDim.x) c += s_keys[p.sel_off[l] + i] >= tau;
        c = __reduce_add_sync(0xffffffffu, c);
        if ((tid & 31) == 0 && c) atomicAdd(&s_cnt[l], c);
    }
    __syncthreads();
    if (tid < p.L) p.n_cut[b * p.L + tid] = s_cnt[tid];
}

int rpn_nms_cut_launch(const RpnLaunch& p, int M, cudaStream_t st) {
    const size_t smem = (size_t)p.sel_per_img * 4;
    if (smem > 48 * 1024) B2D_SMEM(k_nms_cut, smem, "k_nms_cut");   // per device
    const int zero_ctas = rpn_nms_sweep_active(p) ? 148 : 0;
    k_nms_cut<<<p.B + zero_ctas, 1024, smem, st>>>(p, M, zero_ctas);
    return check_launch("rpn_nms_cut");
}

}  // namespace b2d

using namespace b2d;

extern "C" {

size_t b2d_nms_workspace_bytes(long long n_max, int S) {
    if (n_max < 0 || S < 1) return 0;
    const size_t n = (size_t)(n_max > 0 ? n_max : 1);
    const size_t wp = (n + 63) / 64;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return al((size_t)S * n * 16) + al((size_t)S * n * 4) + al((size_t)S * n * wp * 8) + al((size_t)S * 4) +
           al((size_t)S * n * 4);
}

int b2d_nms(int64_t* keep, int* keep_count, const float* boxes, const float* scores, long long n_ld,
            const int* counts, long long n, int S, float thr_f, int max_keep, int presorted, void* workspace,
            size_t ws_bytes, void* stream) {
    B2D_REQUIRE(keep && keep_count && boxes && scores && S >= 1 && n >= 0 && n_ld >= n, "nms: bad args");
    B2D_REQUIRE(n <= kSortCap, "nms: at most 16384 boxes per segment");
    B2D_REQUIRE(workspace && ws_bytes >= b2d_nms_workspace_bytes(n_ld, S), "nms: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { cudaMemsetAsync(keep_count, 0, sizeof(int) * S, st); return B2D_OK; }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const int wp = (int)((n_ld + 63) / 64);
    char* base = (char*)workspace;
    float4* sorted_box = (float4*)base; base += al((size_t)S * n_ld * 16);
    int* sorted_idx = (int*)base; base += al((size_t)S * n_ld * 4);
    uint64_t* mask = (uint64_t*)base; base += al((size_t)S * n_ld * wp * 8);
    int* cnt = (int*)base; base += al((size_t)S * 4);
    uint32_t* nz = (uint32_t*)base;
    B2D_SMEM(k_nms_sort, kSortCap * 8, "k_nms_sort");   // per device
    if (counts) cudaMemcpyAsync(cnt, counts, sizeof(int) * S, cudaMemcpyDeviceToDevice, st);
    else k_fill_i32<<<cdiv(S, 256), 256, 0, st>>>(cnt, (int)n, S);
    k_nms_sort<<<S, kSelThreads, kSortCap * 8, st>>>(sorted_box, sorted_idx, (const float4*)boxes, scores, n_ld, cnt, n,
                                                     presorted);
    NmsSegs s;
    memset(&s, 0, sizeof(s));
    s.L = 1; s.lv0 = 0; s.lvn = 1; s.box_per_img = n_ld; s.mask_per_img = (long long)n_ld * wp; s.wp[0] = wp;
    s.boxes = sorted_box; s.counts = cnt; s.mask = mask; s.nz = nz;
    set_thr(s, thr_f);
    if (n <= kFpThreads * kFpRows) cudaMemsetAsync(nz, 0, (size_t)S * n_ld * 4, st);
    ScanOut o;
    memset(&o, 0, sizeof(o));
    o.keep64 = keep; o.sorted_idx = sorted_idx; o.keep_ld = n_ld; o.keep_count = keep_count;
    launch_mask_scan(s, S, (int)n, max_keep, o, st);
    return check_launch("nms");
}

}  // extern "C"
