// roi_align_tma.cu -- K5, TMA-ring variant: FPN level-mapped RoIAlign forward (2x2 samples per
// bin, NHWC features) as a persistent producer/consumer kernel.
//
// Why: the L1-path kernel (roi_align.cu, k_roi_align_win) is bound by the L1-miss latency x
// concurrency product of the LSU path, not by DRAM (profiles/, round 1): every byte in flight
// costs registers.  Here two producer warps stream the feature rows a RoI touches into a
// shared-memory ring with bulk async copies (cp.async.bulk -> UBLKCP, completion counted on
// mbarriers), so the bytes in flight per SM are bounded by shared memory, not by registers,
// and the 14 consumer warps only ever wait on shared memory.
//
//   layout NHWC: the cells [x0, x1] of feature row y are ONE contiguous run of (x1-x0+1)*C*es
//   bytes.  A row is cut into chunks of kSlotBytes (8 cells of 256 fp32 channels); every chunk
//   is one bulk copy into ring slot (g mod kNS), g = running chunk number of this CTA.
//   full[slot]  : armed by the producer with the chunk's byte count, completed by the copy.
//   empty[slot] : one arrival per consumer warp when no later bin row reads the row.
//   filled[slot]: use number of the last fill (guards the 1-bit phase parity: a producer lane
//                 polls `empty` for use n only after fill n-1 has been issued).
//   Only the DISTINCT rows {lo, hi} of the 2*PH sample rows are fetched, in ascending order; a
//   bin row needs at most 4 of them at a time (<= 16 slots for cpr <= 4 chunks per row, i.e.
//   RoIs up to 32 cells wide), the other slots are prefetch depth; wider RoIs take the generic
//   in-kernel path.
//   Issue rate (scripts/tma_probe.cu, B200): one thread issues a bulk copy every ~430 ns when
//   each copy has its own wait + expect_tx, ~55 ns when the lanes of a warp do the mbarrier
//   work in SIMT and only the UBLKCPs serialise -> one lane per chunk, 16 chunks per round,
//   rounds alternate between the two producer warps.
//   Producer warp 0 is also the planner: it writes the per-bin tap table (ring offsets +
//   bilinear weights) of the NEXT RoI while the consumers work on the current one
//   (double-buffered, tab_full / tab_empty mbarriers).
//
// Arithmetic is that of k_roi_align_win (torchvision's order, no FMA): bit-identical results.
// The [C][PH*PW] result tile is staged in shared memory in its HBM layout and leaves as ONE
// bulk store (cp.async.bulk.global.shared::cta).
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "roi_common.cuh"

namespace b2d {
namespace {

constexpr int kSlotBytes = 8192;              // 8 cells x 256 fp32 channels
constexpr int kNS = 20;                       // ring slots (160 KB)
constexpr int kConsWarps = 14;
constexpr int kConsThreads = kConsWarps * 32;
constexpr int kProdWarps = 2;
constexpr int kThreads = kConsThreads + 32 * kProdWarps;
constexpr int kRound = 16;                    // chunks a producer warp polls at once (< kNS: distinct slots)
constexpr int kMaxCpr = 4;                    // chunks per row: 4 live rows x 4 <= kNS
constexpr int kMaxP = 8;                      // PH, PW <= 8

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); }

// What every role derives, identically, from the RoI coordinates.
struct Plan {
    int valid, fits, img, lvl, H, W;
    RoiGeom g;
    int x0, ncols, cpr;
};

__device__ __forceinline__ Plan make_plan(const RoiArgs& a, long long r, int cpc) {
    Plan p;
    const b2d_roi_cfg& c = a.cfg;
    float x1, y1, x2, y2;
    p.valid = roi_fetch(a, r, p.img, x1, y1, x2, y2);
    p.fits = 0; p.lvl = 0; p.H = p.W = 1; p.x0 = 0; p.ncols = 1; p.cpr = 1;
    if (!p.valid) return p;
    if (a.levels) p.lvl = a.levels[r];
    else p.lvl = c.num_levels > 1 ? roi_level(x1, y1, x2, y2, c.finest_scale, c.num_levels) : 0;
    p.H = c.H[p.lvl]; p.W = c.W[p.lvl];
    p.g = roi_geom(x1, y1, x2, y2, c.spatial_scale[p.lvl], c.PH, c.PW, 2, c.aligned);
    const AxisTap xa = axis_tap(p.g.sx, p.g.bw, 0, 0, 2, p.W), xb = axis_tap(p.g.sx, p.g.bw, c.PW - 1, 1, 2, p.W);
    p.x0 = xa.lo;
    p.ncols = xb.hi - xa.lo + 1;
    p.cpr = (p.ncols + cpc - 1) / cpc;
    // monotone sample positions (bin sizes >= 0, no NaN) and a row that fits the ring
    p.fits = (p.g.bw >= 0.0f) && (p.g.bh >= 0.0f) && p.ncols >= 1 && p.cpr <= kMaxCpr;
    return p;
}

struct __align__(16) TBin {
    uint32_t off[16];      // ring byte offsets of the window cells [row r][col c], r, c < 4
    float w[16];           // w1..w4 of the samples (iy, ix) = (0,0), (0,1), (1,0), (1,1)
    int pat, _p[3];
};

struct Hdr {                // per item, written by the planner, double-buffered
    int valid, fits, cpr, nrows;
    int gbase, _pad[3];    // running chunk number of the RoI's first chunk, mod 2 * kNS
    int kmax[kMaxP];       // highest row index bin row ph reads
    int krel[kMaxP];       // rows below this index are dead once bin row ph is done
};

__device__ __forceinline__ float tofloat(float v) { return v; }
__device__ __forceinline__ float tofloat(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename FT>
__device__ __forceinline__ void lds4(const char* p, float (&f)[4]) {
    if constexpr (sizeof(FT) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
    } else {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        f[0] = fa.x; f[1] = fa.y; f[2] = fb.x; f[3] = fb.y;
    }
}

// one bin, 4 channels per lane, window pattern (PY, PX) as in k_roi_align_win
template <typename FT, int PY, int PX>
__device__ __forceinline__ void bin_eval_s(const char* ring_ch, const TBin* t, float (&acc)[4]) {
    float v[4][4][4];
#pragma unroll
    for (int r = 0; r < PY + 2; ++r) {
        const uint4 o = *reinterpret_cast<const uint4*>(&t->off[r * 4]);
        const uint32_t ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int c = 0; c < PX + 2; ++c) lds4<FT>(ring_ch + ov[c], v[r][c]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = 0.0f;
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
            const int r0 = iy ? PY : 0, c0 = ix ? PX : 0;
            const float4 w = *reinterpret_cast<const float4*>(&t->w[(iy * 2 + ix) * 4]);
#pragma unroll
            for (int e = 0; e < 4; ++e)      // torchvision's order: ((w1 v1 + w2 v2) + w3 v3) + w4 v4, summed over samples
                acc[e] += ((w.x * v[r0][c0][e] + w.y * v[r0][c0 + 1][e]) + w.z * v[r0 + 1][c0][e]) +
                          w.w * v[r0 + 1][c0 + 1][e];
        }
    }
}


// Planner: bin table + header of one item, by the 32 lanes of the producer warp.
__device__ __forceinline__ void plan_item(const RoiArgs& a, const Plan& p, int gbase, int cpc, int cellb, TBin* tab,
                                          Hdr* hdr, int lane) {
    const b2d_roi_cfg& c = a.cfg;
    const int bins = c.PH * c.PW;
    if (lane == 0) { hdr->valid = p.valid; hdr->fits = p.fits; hdr->cpr = p.cpr; hdr->gbase = gbase; hdr->nrows = 0; }
    if (!(p.valid && p.fits)) return;
    // distinct feature rows in ascending order: lane s owns sample row s, a warp scan gives every
    // row its index k in that list (= the order the rows are fetched in)
    int lo = 0, hi = 0;
    if (lane < 2 * c.PH) {
        const AxisTap t = axis_tap(p.g.sy, p.g.bh, lane >> 1, lane & 1, 2, p.H);
        lo = t.lo; hi = t.hi;
    }
    int prev_hi = __shfl_up_sync(0xffffffffu, hi, 1);
    if (lane == 0) prev_hi = -1;
    const int new_lo = (lane < 2 * c.PH) && lo > prev_hi;
    const int last2 = max(lo, prev_hi);
    const int new_hi = (lane < 2 * c.PH) && hi > last2;
    int incl = new_lo + new_hi;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int before = incl - (new_lo + new_hi);
    const int k_lo = new_lo ? before : before - 1 - (prev_hi - lo);
    const int n2 = before + new_lo;
    const int k_hi = new_hi ? n2 : n2 - 1 - (last2 - hi);
    const int nrows = __shfl_sync(0xffffffffu, incl, 2 * c.PH - 1);
    for (int b0 = 0; b0 < bins; b0 += 32) {
        const int bin = min(b0 + lane, bins - 1), ph = bin / c.PW, pw = bin - ph * c.PW;
        const int klo0 = __shfl_sync(0xffffffffu, k_lo, 2 * ph), khi0 = __shfl_sync(0xffffffffu, k_hi, 2 * ph);
        const int klo1 = __shfl_sync(0xffffffffu, k_lo, 2 * ph + 1), khi1 = __shfl_sync(0xffffffffu, k_hi, 2 * ph + 1);
        if (b0 + lane >= bins) continue;
        const AxisTap ty0 = axis_tap(p.g.sy, p.g.bh, ph, 0, 2, p.H), ty1 = axis_tap(p.g.sy, p.g.bh, ph, 1, 2, p.H);
        const AxisTap tx0 = axis_tap(p.g.sx, p.g.bw, pw, 0, 2, p.W), tx1 = axis_tap(p.g.sx, p.g.bw, pw, 1, 2, p.W);
        const int dy = ty1.lo - ty0.lo, dx = tx1.lo - tx0.lo;
        const int py = (dy >= 0 && dy < 2) ? dy : 2, px = (dx >= 0 && dx < 2) ? dx : 2;
        // window rows / columns exactly as k_roi_align_win: p < 2 -> lo0 + {0..p+1} (clamped), p = 2 -> {lo0, hi0, lo1, hi1}
        int kr[4], jc[4];
        if (py < 2) { kr[0] = klo0; kr[1] = khi0; kr[2] = khi1; kr[3] = khi1; }
        else { kr[0] = klo0; kr[1] = khi0; kr[2] = klo1; kr[3] = khi1; }
        if (px < 2) { for (int k = 0; k < 4; ++k) jc[k] = min(tx0.lo + k, p.W - 1) - p.x0; }
        else { jc[0] = tx0.lo - p.x0; jc[1] = tx0.hi - p.x0; jc[2] = tx1.lo - p.x0; jc[3] = tx1.hi - p.x0; }
        TBin t;
        for (int rr = 0; rr < 4; ++rr)
            for (int cc = 0; cc < 4; ++cc) {
                const int j = max(0, min(jc[cc], p.ncols - 1));            // (unused window columns stay in range)
                const int g = (gbase + kr[rr] * p.cpr + j / cpc) % kNS;
                t.off[rr * 4 + cc] = (uint32_t)(g * kSlotBytes + (j % cpc) * cellb);
            }
        const AxisTap* tys[2] = {&ty0, &ty1};
        const AxisTap* txs[2] = {&tx0, &tx1};
        for (int iy = 0; iy < 2; ++iy)
            for (int ix = 0; ix < 2; ++ix) {
                const AxisTap& ty = *tys[iy];
                const AxisTap& tx = *txs[ix];
                float* w = &t.w[(iy * 2 + ix) * 4];
                if (ty.valid && tx.valid) { w[0] = ty.h * tx.h; w[1] = ty.h * tx.l; w[2] = ty.l * tx.h; w[3] = ty.l * tx.l; }
                else { w[0] = w[1] = w[2] = w[3] = 0.0f; }
            }
        t.pat = py * 3 + px; t._p[0] = t._p[1] = t._p[2] = 0;
        tab[bin] = t;
        if (pw == 0) {
            hdr->kmax[ph] = khi1;
            if (ph > 0) hdr->krel[ph - 1] = klo0;
            if (ph == c.PH - 1) { hdr->krel[ph] = nrows; hdr->nrows = nrows; }
        }
    }
}

template <typename FT>
__global__ void __launch_bounds__(kThreads, 1) k_roi_align_tma(RoiArgs a, float* __restrict__ out) {
    extern __shared__ __align__(128) char smem[];
    const b2d_roi_cfg& c = a.cfg;
    const int C = c.C, bins = c.PH * c.PW;
    constexpr int es = (int)sizeof(FT);
    const int cellb = C * es;                          // bytes per cell
    const int cpc = kSlotBytes / cellb;                // cells per chunk
    char* ring = smem;
    float* s_tile = reinterpret_cast<float*>(smem + kNS * kSlotBytes);
    TBin* s_tab = reinterpret_cast<TBin*>(reinterpret_cast<char*>(s_tile) + (size_t)C * bins * 4);     // [2][bins]
    Hdr* s_hdr = reinterpret_cast<Hdr*>(s_tab + 2 * bins);                                             // [2]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_hdr + 2);        // full[kNS], empty[kNS], tab_full[2], tab_empty[2]
    volatile int* s_filled = reinterpret_cast<volatile int*>(s_bar + 2 * kNS + 4);                     // [kNS]
    int* s_rows = const_cast<int*>(s_filled) + kNS;                                                    // [kProdWarps][32]
    const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + kNS);
    const uint32_t tab_full = smem_u32(s_bar + 2 * kNS), tab_empty = smem_u32(s_bar + 2 * kNS + 2);
    const uint32_t ring_u32 = smem_u32(ring);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kNS; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, kConsWarps); s_filled[s] = -1; }
        for (int s = 0; s < 2; ++s) { mbar_init(tab_full + 8 * s, 1); mbar_init(tab_empty + 8 * s, kConsWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;

    if (threadIdx.x >= kConsThreads) {
        // ------------------------------------------------------------------ producer warps (warp 0 also plans)
        const int pwarp = (threadIdx.x - kConsThreads) >> 5;
        int* rows = s_rows + pwarp * 32;
        long long g = 0;                               // running chunk number of the CTA
        int it = 0;
        for (long long r = blockIdx.x; r < a.R; r += gridDim.x, ++it) {
            const Plan p = make_plan(a, r, cpc);
            if (pwarp == 0) {
                const int tb = it & 1;
                mbar_wait(tab_empty + 8 * tb, ((it >> 1) & 1) ^ 1);       // consumers are done with RoI it - 2
                plan_item(a, p, (int)(g % (2 * kNS)), cpc, cellb, s_tab + tb * bins, s_hdr + tb, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(tab_full + 8 * tb);
            }
            if (!p.valid || !p.fits) continue;
            // the distinct rows, ascending (same enumeration as plan_item)
            int lo = 0, hi = 0;
            if (lane < 2 * c.PH) {
                const AxisTap t = axis_tap(p.g.sy, p.g.bh, lane >> 1, lane & 1, 2, p.H);
                lo = t.lo; hi = t.hi;
            }
            int prev_hi = __shfl_up_sync(0xffffffffu, hi, 1);
            if (lane == 0) prev_hi = -1;
            const int new_lo = (lane < 2 * c.PH) && lo > prev_hi;
            const int new_hi = (lane < 2 * c.PH) && hi > max(lo, prev_hi);
            int incl = new_lo + new_hi;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int before = incl - (new_lo + new_hi);
            if (new_lo) rows[before] = lo;
            if (new_hi) rows[before + new_lo] = hi;
            const int nrows = __shfl_sync(0xffffffffu, incl, 2 * c.PH - 1);
            __syncwarp();
            const int nchunks = nrows * p.cpr;
            const char* fimg = reinterpret_cast<const char*>(a.feat[p.lvl]) + (long long)p.img * p.H * p.W * cellb;
            for (int base = pwarp * kRound; base < nchunks; base += kRound * kProdWarps) {
                const int j = base + lane;
                bool pending = lane < kRound && j < nchunks;
                const int k = pending ? j / p.cpr : 0, q = pending ? j - k * p.cpr : 0;
                const long long gj = g + j;
                const int slot = (int)(gj % kNS), use = (int)(gj / kNS);
                const int cells = min(cpc, p.ncols - q * cpc);
                const uint32_t bytes = (uint32_t)(cells * cellb);
                const char* src = fimg + ((long long)rows[k] * p.W + p.x0 + q * cpc) * cellb;
                while (__any_sync(0xffffffffu, pending)) {
                    if (pending && s_filled[slot] == use - 1 && mbar_try(bar_empty + 8 * slot, (uint32_t)(use & 1) ^ 1)) {
                        mbar_expect_tx(bar_full + 8 * slot, bytes);
                        bulk_g2s(ring_u32 + slot * kSlotBytes, src, bytes, bar_full + 8 * slot);
                        s_filled[slot] = use;
                        pending = false;
                    }
                }
            }
            __syncwarp();
            g += nchunks;
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int tid = threadIdx.x, warp = tid >> 5;
    const int passes = C / 128;
    const int units = c.PW * passes;                   // (pw, 128-channel pass) per bin row
    bool store_pending = false;
    int it = 0;
    for (long long r = blockIdx.x; r < a.R; r += gridDim.x, ++it) {
        const int tb = it & 1;
        const Hdr* hdr = s_hdr + tb;
        const TBin* tab = s_tab + tb * bins;
        mbar_wait(tab_full + 8 * tb, (uint32_t)((it >> 1) & 1));
        if (!hdr->valid) {                                          // uniform over the CTA
            __syncwarp();
            if (lane == 0) mbar_arrive(tab_empty + 8 * tb);
            continue;
        }
        const bool fits = hdr->fits != 0;
        float* o = out + r * (long long)C * bins;
        if (fits) {
            const int cpr = hdr->cpr, gbase = hdr->gbase;
            int k_wait = 0, k_rel = 0;
            for (int ph = 0; ph < c.PH; ++ph) {
                {   // wait for the chunks of the new rows, one lane per chunk
                    const int need = hdr->kmax[ph] + 1;
                    for (int j = k_wait * cpr + lane; j < need * cpr; j += 32) {
                        const int g = (gbase + j) % (2 * kNS);
                        mbar_wait(bar_full + 8 * (g % kNS), (uint32_t)(g / kNS));
                    }
                    k_wait = max(k_wait, need);
                    __syncwarp();
                }
                if (ph == 0 && store_pending) {                     // the previous tile has left shared memory
                    if (tid == 0) bulk_wait_read0();
                    cons_sync();
                }
                for (int u = warp; u < units; u += kConsWarps) {
                    const int pw = u / passes, cp = u - pw * passes;
                    const int bin = ph * c.PW + pw;
                    const TBin* t = tab + bin;
                    const char* rc = ring + (cp * 128 + lane * 4) * es;
                    float acc[4];
                    switch (t->pat) {                          // uniform over the warp
                        case 0: bin_eval_s<FT, 0, 0>(rc, t, acc); break;
                        case 1: bin_eval_s<FT, 0, 1>(rc, t, acc); break;
                        case 2: bin_eval_s<FT, 0, 2>(rc, t, acc); break;
                        case 3: bin_eval_s<FT, 1, 0>(rc, t, acc); break;
                        case 4: bin_eval_s<FT, 1, 1>(rc, t, acc); break;
                        case 5: bin_eval_s<FT, 1, 2>(rc, t, acc); break;
                        case 6: bin_eval_s<FT, 2, 0>(rc, t, acc); break;
                        case 7: bin_eval_s<FT, 2, 1>(rc, t, acc); break;
                        default: bin_eval_s<FT, 2, 2>(rc, t, acc); break;
                    }
                    float* st = s_tile + (cp * 128 + lane * 4) * bins + bin;    // x / 4 == x * 0.25 exactly
#pragma unroll
                    for (int e = 0; e < 4; ++e) st[e * bins] = acc[e] * 0.25f;
                }
                {   // release the rows no later bin row reads
                    const int rel = hdr->krel[ph];
                    __syncwarp();
                    for (int j = k_rel * cpr + lane; j < rel * cpr; j += 32)
                        mbar_arrive(bar_empty + 8 * ((gbase + j) % kNS));
                    k_rel = max(k_rel, rel);
                }
            }
        } else {
            // generic path (RoI wider than the ring holds): one thread per (bin, channel), taps from global
            if (store_pending) {
                if (tid == 0) bulk_wait_read0();
                cons_sync();
            }
            const Plan p = make_plan(a, r, cpc);
            const FT* feat = reinterpret_cast<const FT*>(a.feat[p.lvl]) + (long long)p.img * p.H * p.W * C;
            for (int t = tid; t < (a.pf_dist == -7 ? 0 : C * bins); t += kConsThreads) {      // (-7: dev knob, timing without this path)
                const int ch = t % C, bin = t / C, ph = bin / c.PW, pw = bin - ph * c.PW;
                float acc = 0.0f;
                for (int iy = 0; iy < 2; ++iy) {
                    const AxisTap ty = axis_tap(p.g.sy, p.g.bh, ph, iy, 2, p.H);
                    for (int ix = 0; ix < 2; ++ix) {
                        const AxisTap tx = axis_tap(p.g.sx, p.g.bw, pw, ix, 2, p.W);
                        float w1 = 0.0f, w2 = 0.0f, w3 = 0.0f, w4 = 0.0f;
                        if (ty.valid && tx.valid) { w1 = ty.h * tx.h; w2 = ty.h * tx.l; w3 = ty.l * tx.h; w4 = ty.l * tx.l; }
                        const FT* f = feat + ch;
                        const float v1 = tofloat(f[((long long)ty.lo * p.W + tx.lo) * C]), v2 = tofloat(f[((long long)ty.lo * p.W + tx.hi) * C]);
                        const float v3 = tofloat(f[((long long)ty.hi * p.W + tx.lo) * C]), v4 = tofloat(f[((long long)ty.hi * p.W + tx.hi) * C]);
                        acc += ((w1 * v1 + w2 * v2) + w3 * v3) + w4 * v4;
                    }
                }
                s_tile[ch * bins + bin] = acc * 0.25f;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(tab_empty + 8 * tb);            // table / header of this RoI are free
        fence_async_smem();                                        // generic-proxy tile writes -> async-proxy read
        cons_sync();
        if (tid == 0) {
            bulk_s2g(o, smem_u32(s_tile), (uint32_t)(C * bins * 4));
            bulk_commit();
        }
        store_pending = true;
    }
    if (tid == 0 && store_pending) bulk_wait0();
}

}  // namespace

int roi_align_tma_try(const RoiArgs& a, float* out, cudaStream_t st) {
    const b2d_roi_cfg& c = a.cfg;
    const int bins = c.PH * c.PW;
    const int es = c.layout == 2 ? 2 : 4;
    if (c.layout < 1 || c.sampling_ratio != 2 || c.PH > kMaxP || c.PW > kMaxP) return 1;
    if (c.C % 128 != 0 || c.C * es > 1024 || kSlotBytes % (c.C * es) != 0) return 1;
    if (((long long)c.C * bins * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return 1;
    for (int l = 0; l < c.num_levels; ++l)
        if (reinterpret_cast<uintptr_t>(a.feat[l]) & 15) return 1;
    const size_t smem = (size_t)kNS * kSlotBytes + (size_t)c.C * bins * 4 + 2 * (size_t)bins * sizeof(TBin) + 2 * sizeof(Hdr) +
                        (2 * kNS + 4) * sizeof(uint64_t) + (kNS + 32 * kProdWarps) * sizeof(int);
    if (smem > 227 * 1024) return 1;
    int sms = 0, dev = 0;                                  // per device, every call (a process may drive several GPUs)
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    B2D_SMEM(k_roi_align_tma<float>, 227 * 1024, "k_roi_align_tma");
    B2D_SMEM(k_roi_align_tma<__nv_bfloat16>, 227 * 1024, "k_roi_align_tma");
    if (sms <= 0) sms = 1;
    const unsigned grid = (unsigned)(a.R < sms ? a.R : sms);
    RoiArgs b = a;
    b.pf_dist = knobs().roi_tma_dev;
    if (c.layout == 1) k_roi_align_tma<float><<<grid, kThreads, smem, st>>>(b, out);
    else k_roi_align_tma<__nv_bfloat16><<<grid, kThreads, smem, st>>>(b, out);
    return check_launch("roi_align_fwd(tma)");
}

}  // namespace b2d
