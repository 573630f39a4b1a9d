// nms.cu -- K4: bitmask NMS with torchvision.ops.nms *CPU* semantics
// (stable descending score order, areas without +1, suppress iff (double)iou > thr;
// reference call sites lib/heads/rpn_head.py:103, lib/region.py:207, lib/utils.py:220).
//
//   k_nms_sort  (generic entry only) per-segment bitonic sort of (score, index)
//   k_nms_mask  64x64 tiles of the upper triangle: column boxes live in registers,
//               row boxes are staged in shared memory and broadcast, the 64-bit
//               suppression word of a row is assembled with two __ballot_sync
//   k_nms_scan  one block per segment: 64-box chunks; the intra-chunk dependency is
//               resolved by warp 0 from the diagonal words (register/shuffle only),
//               then every thread ORs the kept rows into its own word of the
//               removed-set.  Latency-bound by design.
//
// The fp32 threshold passed in is the largest float <= the caller's double
// threshold, which makes `iou > thr_f` identical to torchvision's CPU comparison
// `(double)iou > thr` for every float iou.
#include <cstring>

#include "common.cuh"
#include "pipeline.cuh"

namespace b2d {

struct NmsSegs {
    int L;                              // levels per image (1 for the generic entry)
    long long box_per_img, box_off[kMaxLevels];
    long long mask_per_img, mask_off[kMaxLevels];
    int wp[kMaxLevels];                 // mask row pitch in words
    const float4* boxes;                // score-sorted boxes
    const int* counts;                  // int[S]
    uint64_t* mask;
    float thr, thr_lo, thr_hi;          // thr_lo/hi: decisive bounds that avoid the divide (see suppresses)
};

// (double)iou > thr with iou = inter / ((aa + ab) - inter), evaluated exactly like the
// reference.  The IEEE divide is only executed inside the narrow band |iou/thr - 1| <
// 2^-18 where the cheap products cannot decide; outside it the sign of
// inter - thr*u already determines fl(inter/u) > thr (margin >> half an ulp).
__device__ __forceinline__ bool suppresses(const float4& a, float aa, const float4& b, float ab, float thr,
                                           float thr_lo, float thr_hi) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
    const float inter = w * h;
    if (inter == 0.0f && thr >= 0.0f) return false;      // 0/u is 0 or NaN: never > thr
    const float u = (aa + ab) - inter;
    if (thr > 0.0f && u > 0.0f && u < 3.0e38f) {
        if (inter > thr_hi * u) return true;
        if (inter < thr_lo * u) return false;
    }
    return inter / u > thr;
}

// Tile (rb, cb), cb >= rb, of the 64x64-blocked suppression matrix.  Off-diagonal tiles
// hold bit j of row i iff box i suppresses box j (j > i always).  DIAGONAL tiles hold the
// full symmetric relation (bit j set iff i != j and IoU(i,j) > thr): the scan needs, for a
// box j, the set of *earlier* boxes that suppress it, which by symmetry of the IoU
// arithmetic is (word_j & lower_bits(j)).
__global__ void __launch_bounds__(64) k_nms_mask(NmsSegs s, int wmax) {
    __shared__ float4 s_row[64];
    const int seg = blockIdx.y;
    const int l = seg % s.L, b = seg / s.L;
    const int n = s.counts[seg];
    int t = blockIdx.x, rb = 0;
    while (t >= wmax - rb) { t -= wmax - rb; ++rb; }
    const int cb = rb + t;
    const int r0 = rb * 64, c0 = cb * 64;
    if (r0 >= n || c0 >= n) return;
    const float4* boxes = s.boxes + (long long)b * s.box_per_img + s.box_off[l];
    uint64_t* mask = s.mask + (long long)b * s.mask_per_img + s.mask_off[l];
    const int wp = s.wp[l];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    s_row[threadIdx.x] = (r0 + (int)threadIdx.x < n) ? boxes[r0 + threadIdx.x] : zero;
    const int j0 = c0 + lane, j1 = c0 + 32 + lane;
    const float4 cb0 = j0 < n ? boxes[j0] : zero, cb1 = j1 < n ? boxes[j1] : zero;
    const float a0 = (cb0.z - cb0.x) * (cb0.w - cb0.y), a1 = (cb1.z - cb1.x) * (cb1.w - cb1.y);
    __syncthreads();
    uint64_t myword = 0;
    const int rows = min(32, n - (r0 + w * 32));
    for (int rr = 0; rr < rows; ++rr) {
        const int i = r0 + w * 32 + rr;
        const float4 rbx = s_row[w * 32 + rr];
        const float ra = (rbx.z - rbx.x) * (rbx.w - rbx.y);
        const bool p0 = (j0 < n) && (j0 != i) && suppresses(rbx, ra, cb0, a0, s.thr, s.thr_lo, s.thr_hi);
        const bool p1 = (j1 < n) && (j1 != i) && suppresses(rbx, ra, cb1, a1, s.thr, s.thr_lo, s.thr_hi);
        const unsigned lo = __ballot_sync(0xffffffffu, p0), hi = __ballot_sync(0xffffffffu, p1);
        if (lane == rr) myword = ((uint64_t)hi << 32) | lo;
    }
    const int row = r0 + w * 32 + lane;
    if (lane < rows) mask[(long long)row * wp + cb] = myword;
}

// One block (256 threads) per segment.  Per 64-box chunk c:
//   resolve  warp 0 finds the kept boxes of the chunk as the unique fixed point of
//            kept_j = alive_j & !(any kept i < j suppresses j), iterated with two ballots
//            per round (bit j is final after all lower bits are; typically 2-5 rounds)
//   update   thread w ORs the rows of the kept boxes into its word w > c of the removed-set
// STAGE (W <= 32, the RPN case): the 64 x W words of chunk c+1 are prefetched into
// registers by all 256 threads while chunk c is processed, then parked in shared memory, so
// no global-memory latency sits on the serial chain.  Latency-bound by design.
constexpr int kScanThreads = 256;

struct ScanOut {
    // RPN pipeline: survivors are compacted (score order) from the sel_* arrays into kept_*
    const uint32_t* src_key; const int* src_idx; float4* dst_box; uint32_t* dst_key; int* dst_idx;
    // generic entry: int64 original indices
    int64_t* keep64; const int* sorted_idx; long long keep_ld;
    int* keep_count;
};

template <bool STAGE>
__global__ void __launch_bounds__(kScanThreads) k_nms_scan(NmsSegs s, int max_keep, ScanOut o) {
    __shared__ uint64_t s_removed[kSortCap / 64];
    __shared__ uint64_t s_keepw[kSortCap / 64];
    __shared__ uint64_t s_rows[STAGE ? 2 * 64 * 32 : 1];
    __shared__ int s_prefix[kSortCap / 64];
    __shared__ uint64_t s_keep;
    const int seg = blockIdx.x;
    const int l = seg % s.L, b = seg / s.L;
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
    // staging geometry: thread t owns row (t >> 2) of the chunk and words [(t & 3) * 8, +8)
    const int srow = threadIdx.x >> 2, sw0 = (threadIdx.x & 3) * 8;
    uint64_t pre[8];
    auto prefetch = [&](int c) {
        const int row = c * 64 + srow;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int w = sw0 + q;
            pre[q] = (row < n && w > c && w < W) ? mask[(long long)row * wp + w] : 0ull;
        }
    };
    auto park = [&](int c) {
#pragma unroll
        for (int q = 0; q < 8; ++q) s_rows[((c & 1) * 64 + srow) * 32 + sw0 + q] = pre[q];
    };
    uint64_t nd0 = 0, nd1 = 0;                       // diagonal words of the next chunk (warp 0)
    auto load_diag = [&](int c) {
        const int row0 = c * 64 + lane, row1 = row0 + 32;
        nd0 = (c < W && row0 < n) ? mask[(long long)row0 * wp + c] : 0ull;
        nd1 = (c < W && row1 < n) ? mask[(long long)row1 * wp + c] : 0ull;
    };
    if (STAGE) { prefetch(0); park(0); }
    if (threadIdx.x < 32) load_diag(0);
    __syncthreads();
    int kept_total = 0;
    for (int c = 0; c < W; ++c) {
        if (STAGE && c + 1 < W) prefetch(c + 1);
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
            if (STAGE) {
                // warp g ORs the kept rows 8g..8g+7 of the chunk, lane = word: conflict-free LDS
                const int w = lane, g = threadIdx.x >> 5;
                if (w > c && w < W) {
                    uint64_t acc = 0ull;
                    const uint64_t* rows = s_rows + ((c & 1) * 64 + g * 8) * 32 + w;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if ((keep >> (g * 8 + q)) & 1ull) acc |= rows[q * 32];
                    if (acc) atomicOr(reinterpret_cast<unsigned long long*>(&s_removed[w]), (unsigned long long)acc);
                }
            } else {
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
            if (STAGE && c + 1 < W) park(c + 1);
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
        const int pos = s_prefix[w] + __popcll(kw & ((1ull << bit) - 1ull));
        if (o.keep64) {
            o.keep64[(long long)seg * o.keep_ld + pos] = (int64_t)o.sorted_idx[(long long)seg * o.keep_ld + r];
        } else {
            o.dst_box[base + pos] = s.boxes[base + r];
            o.dst_key[base + pos] = o.src_key[base + r];
            o.dst_idx[base + pos] = o.src_idx[base + r];
        }
    }
    if (threadIdx.x == 0) o.keep_count[seg] = kept_total;
}

static void set_thr(NmsSegs& s, float thr) {
    s.thr = thr;
    s.thr_hi = thr * (1.0f + 1.0f / 262144.0f);
    s.thr_lo = thr * (1.0f - 1.0f / 262144.0f);
}

static void launch_scan(const NmsSegs& s, int S, int wmax, int max_keep, const ScanOut& o, cudaStream_t st) {
    if (wmax <= 32) k_nms_scan<true><<<S, kScanThreads, 0, st>>>(s, max_keep, o);
    else k_nms_scan<false><<<S, kScanThreads, 0, st>>>(s, max_keep, o);
}

// generic entry: sort (score desc, index asc) and gather boxes into score order
__global__ void __launch_bounds__(kSelThreads) k_nms_sort(float4* __restrict__ sorted_box, int* __restrict__ sorted_idx,
                                                          const float4* __restrict__ boxes,
                                                          const float* __restrict__ scores, long long n_ld,
                                                          const int* __restrict__ counts, long long n, int presorted) {
    extern __shared__ uint64_t s_buf[];
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

int rpn_nms_launch(const RpnLaunch& p, cudaStream_t st) {
    NmsSegs s;
    memset(&s, 0, sizeof(s));
    s.L = p.L; s.box_per_img = p.sel_per_img; s.mask_per_img = p.mask_per_img;
    int wmax = 1;
    for (int l = 0; l < p.L; ++l) {
        s.box_off[l] = p.sel_off[l]; s.mask_off[l] = p.mask_off[l];
        s.wp[l] = (p.kcap[l] + 63) / 64;
        wmax = max(wmax, s.wp[l]);
    }
    s.boxes = p.sel_box; s.counts = p.sel_count; s.mask = p.mask;
    set_thr(s, p.nms_thr);
    const int S = p.B * p.L;
    dim3 grid(wmax * (wmax + 1) / 2, S);
    k_nms_mask<<<grid, 64, 0, st>>>(s, wmax);
    ScanOut o;
    memset(&o, 0, sizeof(o));
    o.src_key = p.sel_key; o.src_idx = p.sel_idx; o.dst_box = p.kept_box; o.dst_key = p.kept_key; o.dst_idx = p.kept_idx;
    o.keep_count = p.keep_count;
    launch_scan(s, S, wmax, p.post_nms, o, st);
    return check_launch("rpn_nms");
}

}  // namespace b2d

using namespace b2d;

extern "C" {

size_t b2d_nms_workspace_bytes(long long n_max, int S) {
    if (n_max < 0 || S < 1) return 0;
    const size_t n = (size_t)(n_max > 0 ? n_max : 1);
    const size_t wp = (n + 63) / 64;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return al((size_t)S * n * 16) + al((size_t)S * n * 4) + al((size_t)S * n * wp * 8) + al((size_t)S * 4);
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
    int* cnt = (int*)base;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_nms_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortCap * 8);
        attr_set = true;
    }
    if (counts) cudaMemcpyAsync(cnt, counts, sizeof(int) * S, cudaMemcpyDeviceToDevice, st);
    else k_fill_i32<<<cdiv(S, 256), 256, 0, st>>>(cnt, (int)n, S);
    k_nms_sort<<<S, kSelThreads, kSortCap * 8, st>>>(sorted_box, sorted_idx, (const float4*)boxes, scores, n_ld, cnt, n,
                                                     presorted);
    NmsSegs s;
    memset(&s, 0, sizeof(s));
    s.L = 1; s.box_per_img = n_ld; s.mask_per_img = (long long)n_ld * wp; s.wp[0] = wp;
    s.boxes = sorted_box; s.counts = cnt; s.mask = mask;
    set_thr(s, thr_f);
    const int wmax = (int)((n + 63) / 64);
    dim3 grid(wmax * (wmax + 1) / 2, S);
    k_nms_mask<<<grid, 64, 0, st>>>(s, wmax);
    ScanOut o;
    memset(&o, 0, sizeof(o));
    o.keep64 = keep; o.sorted_idx = sorted_idx; o.keep_ld = n_ld; o.keep_count = keep_count;
    launch_scan(s, S, wmax, max_keep, o, st);
    return check_launch("nms");
}

}  // extern "C"
