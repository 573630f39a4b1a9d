// rpn_front.cu -- K3 as ONE thread-block-cluster kernel for all (image, level) segments of a batch:
//   radix histogram -> threshold -> candidate compaction -> exact sort -> anchors in registers -> delta decode ->
//   clip (-> min-size filter), i.e. what k_hist + k_compact + k_select (select.cu) do in three launches with two
//   global-memory round trips of the candidates in between.
// Reference: RPNHead.predict_single_image, the per-level part (lib/heads/rpn_head.py:81-101); also the per-level top-k
// of AnchorHead.predict_single_image (lib/heads/anchor_head.py:224-248) through score_mode 2.
//
// Layout of the work.  A level of an image is owned by a GROUP of 1, 2, 4 or 8 consecutive 1024-thread CTAs of one
// cluster (host-side plan: 25 scores per thread at most); large levels run in clusters of 8, the small ones in a
// second launch with clusters of 2 on a side stream (16 clusters of 8 do not fit a B200 at once, measured).  The
// level's scores are read from HBM exactly once and stay in registers as monotone keys.  Per group:
//   1. 12-bit histogram of the keys in shared memory; the group's CTAs reduce disjoint bin slices of each other's
//      histograms through distributed shared memory, exchange the slice totals, and the CTA whose slice holds the
//      k-th largest key finds the threshold bin and broadcasts it           (3 cluster barriers)
//   2. (degenerate score maps only: more than 8192 candidates would remain) up to three narrowing passes over the next
//      8 + 8 + 4 key bits; if even the full key ties too often the lowest indices win, as everywhere in this library
//   3. candidates are staged locally, their counts exchanged, and every CTA pushes its candidates into the group
//      owner's sort buffer with remote shared-memory stores                 (2 cluster barriers)
//   4. the owner sorts (bucket sort, bitonic fallback), decodes the best k and writes sel_box / sel_key / sel_idx.
// Selection order: score descending, ties by the lowest index (torch.topk leaves ties unspecified).
#include <cooperative_groups.h>

#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "pipeline.cuh"
#include "rpn_common.cuh"

namespace cg = cooperative_groups;

namespace b2d {

constexpr int kFrThreads = 1024;
constexpr int kFrPer = 25;             // scores per thread (registers)
constexpr int kFrCl = 8;               // largest cluster (portable maximum)
constexpr int kFrClSmall = 2;          // cluster size of the launch for the small levels
constexpr int kFrCap = kBucketCap;     // candidates a group owner can sort (8192)
constexpr int kFrMaxSlots = 8;         // clusters per image

struct FrontPlan {
    int slots;                                         // clusters per image
    int cs;                                            // CTAs per cluster of this launch
    int dbg_base;                                      // first debug-stamp row of this launch
    signed char level[kFrMaxSlots * kFrCl];            // level of (slot * cs + cluster rank); -1: idle CTA
    signed char g0[kFrMaxSlots * kFrCl];               // first cluster rank of its group
    signed char gn[kFrMaxSlots * kFrCl];               // CTAs in its group (power of two)
};

struct FrShared {
    uint32_t h[kHistBins + 4];          // local histogram of the current pass; bucket-sort offsets afterwards
    uint32_t red[kHistBins];            // my slice of the group-reduced histogram; bucket-sort cursors afterwards
    uint32_t wsum[80];
    uint32_t tot[kFrCl];                // slice totals of my group (each written by its owner)
    uint32_t res[4];                    // digit, keys above the digit's bin (this pass), keys in the digit's bin
    uint32_t cntA[kFrCl], cntB[kFrCl];  // staged candidates per group member
    uint32_t flag[kFrCl];               // (rank 0 of the cluster) "my group needs narrowing passes"
    uint32_t fnd[2];
    int nA, nB, base;
    int warp[kFrThreads / 32];
};

// sum over the block; every thread gets the result.  Two barriers.
__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* s_w /*>= 33*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                   // s_w may still be read from a previous use
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
    return t;
}

// largest t in [0, nb) with sum_{bin >= t} red[bin] >= need (need >= 1, total >= need); also the number of keys in
// the bins above t and in bin t.  `red` may live in another CTA of the cluster.  nb <= 4096, blockDim.x == 1024.
// Result through s.fnd[0..2].
__device__ void find_in_slice(FrShared& s, const uint32_t* red, int nb, uint32_t need) {
    constexpr int per = kHistBins / kFrThreads;        // 4 consecutive bins per thread
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t loc[per], sum = 0;
#pragma unroll
    for (int q = 0; q < per; ++q) { const int bin = tid * per + q; loc[q] = bin < nb ? red[bin] : 0u; }
#pragma unroll
    for (int q = 0; q < per; ++q) sum += loc[q];
    uint32_t v = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += t;
    }
    __syncthreads();
    if (lane == 0) s.wsum[warp] = v;
    if (tid == 0) { s.fnd[0] = 0u; s.fnd[1] = 0u; s.fnd[2] = 0u; }
    __syncthreads();
    uint32_t above = v - sum;
    for (int w = warp + 1; w < kFrThreads / 32; ++w) above += s.wsum[w];
    if (above < need && above + sum >= need) {
        uint32_t run = above;
        for (int q = per - 1; q >= 0; --q) {
            if (run + loc[q] >= need) { s.fnd[0] = (uint32_t)(tid * per + q); s.fnd[1] = run; s.fnd[2] = loc[q]; break; }
            run += loc[q];
        }
    }
    __syncthreads();
}

// One radix pass of a group: the local histograms s.h[0..nb) of its CTAs are complete (a cluster barrier lies behind
// us).  Returns the digit of the need-th largest participating key, the number of participants in higher bins and in
// the digit's bin.  ONE cluster barrier: every CTA reduces its slice of the bins over the group and publishes the slice
// total; afterwards every CTA locates the slice that holds the digit and searches that slice (remote reads of <= 2 KB)
// itself -- no broadcast round.  Every CTA of the cluster calls this the same number of times (CTAs with work == false
// only take part in the barrier).
__device__ void group_select(cg::cluster_group& cl, FrShared& s, bool work, int g0, int gn, int my, int nb, uint32_t need,
                             uint32_t& digit, uint32_t& above, uint32_t& at) {
    const int tid = threadIdx.x;
    const int sl = work ? nb / gn : 0;                 // bins per slice (nb, gn powers of two, nb >= gn)
    if (work) {
        uint32_t part = 0;
        for (int t = tid; t < sl; t += kFrThreads) {
            uint32_t v[kFrCl];
#pragma unroll
            for (int j = 0; j < kFrCl; ++j) v[j] = j < gn ? cl.map_shared_rank(s.h, g0 + j)[my * sl + t] : 0u;   // all in flight
            const uint32_t sum = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            s.red[t] = sum;
            part += sum;
        }
        part = block_sum_u32(part, s.wsum);
        if (tid < gn) cl.map_shared_rank(s.tot, g0 + tid)[my] = part;
    }
    cl.sync();
    if (work) {
        int js = 0;
        uint32_t ab = 0, run = 0;
        for (int j = gn - 1; j >= 0; --j) {
            if (run + s.tot[j] >= need) { js = j; ab = run; break; }
            run += s.tot[j];
        }
        find_in_slice(s, cl.map_shared_rank(s.red, g0 + js), sl, need - ab);
        digit = (uint32_t)(js * sl) + s.fnd[0]; above = ab + s.fnd[1]; at = s.fnd[2];
    }
}

__global__ void __launch_bounds__(kFrThreads, 1) k_rpn_front(RpnLaunch p, FrontPlan fp) {
    pdl_launch_dependents();                            // k_rpn_back may be set up while this kernel runs (it waits for our completion)
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ __align__(16) uint64_t s_buf[];   // [0, cap): the owner's candidate list; [cap, 2 cap): local staging
    __shared__ FrShared s;
    const int tid = threadIdx.x;
    const int crank = (int)cl.block_rank();
    const int cs = fp.cs;
    const int slot = blockIdx.x / cs, b = blockIdx.y;
    const int pe = slot * cs + crank;
    const int dbg_cta = fp.dbg_base + blockIdx.x;
    const int l = fp.level[pe];
    const bool active = l >= 0;
    const int g0 = active ? fp.g0[pe] : crank, gn = active ? fp.gn[pe] : 1, my = crank - g0;
    const int lq = active ? l : 0;
    const b2d_level& lv = p.pyr.lv[lq];
    const int n = active ? p.n[lq] : 0, k = active ? p.kcap[lq] : 0;
    const int seg = b * p.L + lq;
    const float* cls = seg_cls(p, b, lq);
    const bool sel = active && k < n;                    // the level needs a selection
    const bool identity = active && (k >= n) && !p.do_nms;   // AnchorHead path without top-k: keep index order
    // my part of the level: [start, end)
    int chunk = (n + gn - 1) / gn;
    chunk = (chunk + 3) & ~3;
    const int start = min(n, my * chunk), end = min(n, start + chunk);

    dbg_stamp(p, b, dbg_cta, 9);                           // kernel entry
    uint32_t key[kFrPer];
    if (p.score_mode == 0 && !identity) {
        // sigmoid RPN (every FPN config of the reference): all 25 loads of a thread in flight at once.  Through the generic
        // load_logit() below the score-mode branches keep the compiler from batching them: 25 dependent L2 round trips,
        // 8 us from kernel entry to the first histogram pass (globaltimer stamps) against 1.5 us here.
        float v[kFrPer];
#pragma unroll
        for (int q = 0; q < kFrPer; ++q) {
            const int i = start + q * kFrThreads + tid;
            v[q] = i < end ? __ldg(cls + i) : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < kFrPer; ++q) key[q] = (start + q * kFrThreads + tid < end) ? f2key(v[q]) : 0u;
    } else {
#pragma unroll
        for (int q = 0; q < kFrPer; ++q) {
            const int i = start + q * kFrThreads + tid;
            key[q] = i < end ? (identity ? 0xffffffffu : f2key(load_logit(cls, n, i, p.score_mode, p.cls_ch))) : 0u;
        }
    }
    dbg_stamp(p, b, dbg_cta, 0);
    for (int t = tid; t < kHistBins; t += kFrThreads) s.h[t] = 0u;
    if (tid == 0) { s.nA = 0; s.nB = 0; }
    __syncthreads();
    if (sel) {
#pragma unroll
        for (int q = 0; q < kFrPer; ++q)
            if (start + q * kFrThreads + tid < end) atomicAdd(&s.h[key[q] >> (32 - kHistBits)], 1u);
    }
    dbg_stamp(p, b, dbg_cta, 1);
    cl.sync();
    // ---- pass 1: the 12 leading key bits
    uint32_t prefix = 0u, above = 0u, inbin = (uint32_t)n;
    int pbits = 0;
    {
        uint32_t d, ab, at;
        group_select(cl, s, sel, g0, gn, my, kHistBins, (uint32_t)k, d, ab, at);
        if (sel) { prefix = d; pbits = kHistBits; above = ab; inbin = at; }
    }
    dbg_stamp(p, b, dbg_cta, 2);
    // ---- narrowing passes, taken by the whole cluster if any of its groups still has too many candidates
    const bool over = sel && above + inbin > (uint32_t)kFrCap;
    if (tid == 0) cl.map_shared_rank(s.flag, 0)[crank] = over ? 1u : 0u;
    cl.sync();
    if (tid < 32) {                                        // one warp reads the eight flags of rank 0 (remote), the rest of the CTA locally
        const uint32_t f = tid < cs ? cl.map_shared_rank(s.flag, 0)[tid] : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, f != 0u);
        if (tid == 0) s.fnd[0] = m;
    }
    __syncthreads();
    const bool any_over = s.fnd[0] != 0u;
    dbg_stamp(p, b, dbg_cta, 3);
    if (any_over) {
        const int dbits[3] = {8, 8, 4};
        for (int ps = 0; ps < 3; ++ps) {
            const int d = dbits[ps];
            const bool work = sel && above + inbin > (uint32_t)kFrCap;       // group-uniform
            for (int t = tid; t < (1 << d); t += kFrThreads) s.h[t] = 0u;
            __syncthreads();
            if (work) {
#pragma unroll
                for (int q = 0; q < kFrPer; ++q) {
                    const bool in = (start + q * kFrThreads + tid < end) && (key[q] >> (32 - pbits)) == prefix;
                    if (in) atomicAdd(&s.h[(key[q] >> (32 - pbits - d)) & ((1u << d) - 1u)], 1u);
                }
            }
            cl.sync();
            uint32_t dg, ab, at;
            group_select(cl, s, work, g0, gn, my, 1 << d, (uint32_t)k - above, dg, ab, at);
            if (work) { prefix = (prefix << d) | dg; pbits += d; above += ab; inbin = at; }
        }
    }
    // all 32 bits resolved and still too many: the remaining candidates are exact ties of the k-th key
    const bool tie = sel && above + inbin > (uint32_t)kFrCap;
    // ---- candidates: A = keys above the threshold prefix (all selected), B = keys with the threshold prefix
    uint64_t* stage = s_buf + kFrCap;
    // two passes, no atomics: per-thread counts -> block exclusive scan -> placement (a shared counter costs one
    // serialised returning atomic per candidate: 5 us for the ~270 candidates of a level-0 CTA, measured)
    // key ranges of the two lists: B = [blo, bhi], A = (bhi, 2^32).  nv = this thread's valid slots (a prefix of q).
    const uint32_t blo = pbits ? (pbits < 32 ? prefix << (32 - pbits) : prefix) : 0u;
    const uint32_t bhi = pbits ? (pbits < 32 ? (blo | ((1u << (32 - pbits)) - 1u)) : prefix) : 0xffffffffu;
    const int span = end - start - tid;
    const int nv = active ? (span <= 0 ? 0 : min(kFrPer, (span + kFrThreads - 1) / kFrThreads)) : 0;
    uint32_t cnt = 0;                                     // A in the low half, B in the high half
#pragma unroll
    for (int q = 0; q < kFrPer; ++q) {
        const bool v = q < nv;
        cnt += ((v && key[q] > bhi) ? 1u : 0u) + ((v && key[q] >= blo && key[q] <= bhi) ? 0x10000u : 0u);
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) s.wsum[tid >> 5] = incl;
    __syncthreads();
    uint32_t before = incl - cnt, all = 0;
    for (int w = 0; w < kFrThreads / 32; ++w) {
        const uint32_t c = s.wsum[w];
        if (w < (tid >> 5)) before += c;
        all += c;
    }
    const int nA = (int)(all & 0xffffu), nB = (int)(all >> 16);
    const float* reg_pf = seg_reg(p, b, lq);
    if (cnt != 0u && !(tie && (cnt & 0xffffu) == 0u)) {   // one thread in four holds a candidate at all
        uint32_t pa = before & 0xffffu, pb = before >> 16;
#pragma unroll
        for (int q = 0; q < kFrPer; ++q) {
            if (q < nv && key[q] >= blo) {
                const int idx = start + q * kFrThreads + tid;
                const uint64_t c = make_comp(key[q], (uint32_t)idx);
                if (key[q] > bhi) stage[pa++] = c;
                else if (!tie) stage[kFrCap - 1 - (pb++)] = c;
                // the candidate's four deltas are read by the decode below, after the sort: ask L2 for them now (the step's
                // inputs are cold in L2 -- the previous RoIAlign streamed 600 MB through it: -2 us, scripts/bench_rpn_cold.py)
#pragma unroll
                for (int e = 0; e < 4; ++e) asm volatile("prefetch.global.L2 [%0];" ::"l"(reg_pf + (long long)e * n + idx));
            }
        }
    }
    __syncthreads();
    dbg_stamp(p, b, dbg_cta, 4);
    if (active && tid < gn) {
        cl.map_shared_rank(s.cntA, g0 + tid)[my] = (uint32_t)nA;
        cl.map_shared_rank(s.cntB, g0 + tid)[my] = (uint32_t)nB;
    }
    cl.sync();
    int total = 0;
    if (active) {
        uint32_t offA = 0, offB = 0, totA = 0;
        for (int j = 0; j < gn; ++j) { if (j < my) { offA += s.cntA[j]; offB += s.cntB[j]; } totA += s.cntA[j]; }
        uint64_t* ob = cl.map_shared_rank(s_buf, g0);
        for (int t = tid; t < nA; t += kFrThreads) ob[offA + t] = stage[t];
        if (!tie) {
            for (int t = tid; t < nB; t += kFrThreads) ob[totA + offB + t] = stage[kFrCap - 1 - t];
            total = (int)(above + inbin);
        } else {
            // the first (k - above) ties in index order: CTAs of the group hold ascending index ranges, inside a CTA
            // the order is (q, thread)
            const uint32_t need = (uint32_t)k - above;
            uint32_t run = offB;
#pragma unroll
            for (int q = 0; q < kFrPer; ++q) {                       // (unrolled: key[] stays in registers)
                if (run >= need) break;                              // `run` is block-uniform
                const int i = start + q * kFrThreads + tid;
                const bool f = i < end && key[q] == prefix;
                const unsigned m = __ballot_sync(0xffffffffu, f);
                __syncthreads();
                if ((tid & 31) == 0) s.warp[tid >> 5] = __popc(m);
                __syncthreads();
                uint32_t before = run, all = 0;
                for (int w = 0; w < kFrThreads / 32; ++w) {
                    const uint32_t c = (uint32_t)s.warp[w];
                    if (w < (tid >> 5)) before += c;
                    all += c;
                }
                const uint32_t pos = before + (uint32_t)__popc(m & ((1u << (tid & 31)) - 1u));
                if (f && pos < need) ob[totA + pos] = make_comp(key[q], (uint32_t)i);
                run += all;
            }
            total = k;
        }
    }
    dbg_stamp(p, b, dbg_cta, 5);
    cl.sync();
    dbg_stamp(p, b, dbg_cta, 6);
    const bool owner = active && my == 0;
    const bool split = !(p.min_size > 0.0f);              // no size filter: the group shares the decode (below)
    if (!split && !owner) return;                         // size filter: only the group owner goes on (its own shared memory only)

    // ---- exact order of the candidates (group owner)
    int kk = 0;
    if (owner) {
        if (!bucket_sort_desc(s_buf, s_buf + kFrCap, total, s.h, s.red, s.wsum, p.dbg)) {
            int p2 = 1;
            while (p2 < total) p2 <<= 1;
            for (int i = total + tid; i < p2; i += kFrThreads) s_buf[i] = 0ull;
            __syncthreads();
            bitonic_sort_desc(s_buf, p2);
        }
        kk = min(k, total);
        if (split && tid < gn) cl.map_shared_rank(s.res, g0 + tid)[0] = (uint32_t)kk;
    }
    dbg_stamp(p, b, dbg_cta, 7);

    // ---- decode + clip (+ min-size filter, order preserving: k_select's tail)
    const float* reg = seg_reg(p, b, lq);
    const float img_h = p.img_hw[2 * b], img_w = p.img_hw[2 * b + 1];
    const long long so = (long long)b * p.sel_per_img + p.sel_off[lq];
    float4* sel_box = p.sel_box + so;
    uint32_t* sel_key = p.sel_key + so;
    int* sel_idx = p.sel_idx + so;
    if (split) {
        // Every selected box is kept at its rank, so the CTAs of the group decode interleaved ranks of the owner's
        // sorted list (read through distributed shared memory): one DRAM round trip for the scattered delta loads
        // instead of two per thread of a single CTA.  Two boxes per thread and round, their eight delta loads requested
        // before the first use (as a plain loop the loads of the second box stay behind the stores of the first).
        cl.sync();
        if (active) {
            kk = (int)s.res[0];
            const uint64_t* ob = cl.map_shared_rank(s_buf, g0);
            const int stride = gn * kFrThreads;
            for (int r0 = my * kFrThreads + tid; r0 < kk; r0 += 2 * stride) {
                const int ra = r0, rb = r0 + stride;
                const bool vb = rb < kk;
                const uint64_t ca = ob[ra], cb = vb ? ob[rb] : 0ull;
                const uint32_t ia = comp_idx(ca), ib = vb ? comp_idx(cb) : 0u;
                float da[4], db[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) da[e] = __ldg(reg + (long long)e * n + ia);
#pragma unroll
                for (int e = 0; e < 4; ++e) db[e] = vb ? __ldg(reg + (long long)e * n + ib) : 0.0f;
                const uint32_t ka = identity ? f2key(load_logit(cls, n, (int)ia, p.score_mode, p.cls_ch)) : comp_key(ca);
                const uint32_t kb = (vb && identity) ? f2key(load_logit(cls, n, (int)ib, p.score_mode, p.cls_ch)) : comp_key(cb);
                {
                    const Box o = decode_box(anchor_flat(lv, (int)ia), da[0], da[1], da[2], da[3], p.ms, true, img_h, img_w);
                    sel_box[ra] = make_float4(o.x1, o.y1, o.x2, o.y2);
                    sel_key[ra] = ka;
                    sel_idx[ra] = (int)ia;
                }
                if (vb) {
                    const Box o = decode_box(anchor_flat(lv, (int)ib), db[0], db[1], db[2], db[3], p.ms, true, img_h, img_w);
                    sel_box[rb] = make_float4(o.x1, o.y1, o.x2, o.y2);
                    sel_key[rb] = kb;
                    sel_idx[rb] = (int)ib;
                }
            }
            if (owner && tid == 0) p.sel_count[seg] = kk;
        }
        dbg_stamp(p, b, dbg_cta, 8);
        cl.sync();                                        // the owners' lists stay valid until every reader is done
        return;
    }
    if (tid == 0) s.base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < kk; r0 += kFrThreads) {
        const int r = r0 + tid;
        bool keep = false;
        Box o{0, 0, 0, 0};
        uint32_t ky = 0, idx = 0;
        if (r < kk) {
            const uint64_t c = s_buf[r];
            idx = comp_idx(c);
            ky = identity ? f2key(load_logit(cls, n, (int)idx, p.score_mode, p.cls_ch)) : comp_key(c);
            const Box a = anchor_flat(lv, (int)idx);
            o = decode_box(a, reg[idx], reg[n + idx], reg[2 * n + idx], reg[3 * n + idx], p.ms, true, img_h, img_w);
            keep = ((o.x2 - o.x1) + 1.0f >= p.min_size) && ((o.y2 - o.y1) + 1.0f >= p.min_size);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if ((tid & 31) == 0) s.warp[tid >> 5] = __popc(m);
        __syncthreads();
        int before = s.base;
        for (int w = 0; w < (tid >> 5); ++w) before += s.warp[w];
        if (keep) {
            const int pos = before + __popc(m & ((1u << (tid & 31)) - 1u));
            sel_box[pos] = make_float4(o.x1, o.y1, o.x2, o.y2);
            sel_key[pos] = ky;
            sel_idx[pos] = (int)idx;
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < kFrThreads / 32; ++w) t += s.warp[w];
            s.base += t;
        }
        __syncthreads();
    }
    if (tid == 0) p.sel_count[seg] = s.base;
}

// Host-side plan.  Levels that need more than two CTAs (at 25 scores per thread) go to the launch with clusters of 8,
// the others to the launch with clusters of 2; inside a launch, groups are packed first-fit into clusters and spare
// CTAs of a cluster are given to the group with the most scores per CTA.  false: the multi-kernel path must run
// (a level too large for one cluster, a top-k beyond the owner's sort buffer, or a plain top-k call).
static bool front_plan(const RpnLaunch& p, FrontPlan& big, FrontPlan& small) {
    for (FrontPlan* fp : {&big, &small}) {
        memset(fp, 0, sizeof(*fp));
        memset(fp->level, -1, sizeof(fp->level));
    }
    big.cs = kFrCl; small.cs = kFrClSmall;
    if (p.raw) return false;
    int gmin[kMaxLevels];
    for (int l = 0; l < p.L; ++l) {
        if (p.kcap[l] > kFrCap || p.n[l] < 1) return false;
        const long long per_cta = (long long)kFrThreads * kFrPer;
        int g = 1;
        while ((long long)g * per_cta < (long long)p.n[l] + 4 * g) g <<= 1;      // (+4g: chunks are rounded up to 4)
        if (g > kFrCl) return false;
        gmin[l] = g;
    }
    for (FrontPlan* fp : {&big, &small}) {
        const int cs = fp->cs;
        int gn[kMaxLevels], slot_of[kMaxLevels], used[kFrMaxSlots * kFrCl] = {0}, slots = 0;
        const int max_slots = kFrMaxSlots * kFrCl / cs;
        for (int l = 0; l < p.L; ++l) {
            slot_of[l] = -1;
            const bool is_big = gmin[l] > kFrClSmall;
            if (is_big != (fp == &big)) continue;
            gn[l] = gmin[l];
            int sidx = -1;
            for (int q = 0; q < slots; ++q) if (used[q] + gn[l] <= cs) { sidx = q; break; }
            if (sidx < 0) { if (slots == max_slots) return false; sidx = slots++; }
            slot_of[l] = sidx; used[sidx] += gn[l];
        }
        for (int q = 0; q < slots; ++q) {
            for (;;) {
                int best = -1;
                double load = 0.0;
                for (int l = 0; l < p.L; ++l) {
                    if (slot_of[l] != q || used[q] + gn[l] > cs) continue;
                    const double ld = (double)p.n[l] / gn[l];
                    if (ld > load && ld > 4096.0) { load = ld; best = l; }
                }
                if (best < 0) break;
                used[q] += gn[best]; gn[best] *= 2;
            }
        }
        int next[kFrMaxSlots * kFrCl] = {0};
        for (int l = 0; l < p.L; ++l) {
            const int q = slot_of[l];
            if (q < 0) continue;
            for (int r = 0; r < gn[l]; ++r) {
                const int pe = q * cs + next[q] + r;
                fp->level[pe] = (signed char)l; fp->g0[pe] = (signed char)next[q]; fp->gn[pe] = (signed char)gn[l];
            }
            next[q] += gn[l];
        }
        fp->slots = slots;
    }
    small.dbg_base = big.slots * big.cs;
    return true;
}

static int front_launch_one(const RpnLaunch& p, const FrontPlan& fp, cudaStream_t st) {
    const size_t smem = (size_t)2 * kFrCap * sizeof(uint64_t);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(fp.slots * fp.cs), (unsigned)p.B, 1);
    cfg.blockDim = dim3(kFrThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)fp.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_rpn_front, p, fp);
    if (e != cudaSuccess) {
        set_error(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return check_launch("rpn_proposals/k_rpn_front");
}

int rpn_front_launch_count(const RpnLaunch& p) {
    FrontPlan big, small;
    if (!front_plan(p, big, small)) return 0;
    return (big.slots > 0) + (small.slots > 0);
}

// 1: launched; 0: not applicable (the caller runs the multi-kernel path); anything else: error code.
// `side`, `fork`, `join`: a library-owned stream and two events for the second launch (may be null: both launches
// then go to `st` one after the other).
int rpn_front_launch(const RpnLaunch& p, cudaStream_t st, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join) {
    FrontPlan big, small;
    if (!front_plan(p, big, small)) return 0;
    const size_t smem = (size_t)2 * kFrCap * sizeof(uint64_t);
    // per device, once: opt in to the shared-memory size and check that a cluster of 8 can be co-scheduled
    static int ok_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 0;
    if (set_dyn_smem(k_rpn_front, smem, "k_rpn_front") != 0) return 0;    // (error text set; the caller falls back to the multi-kernel chain)
    if (ok_dev[dev] == 0) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(kFrCl, 1, 1); cfg.blockDim = dim3(kFrThreads, 1, 1); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kFrCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int ncl = 0;
        const cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k_rpn_front, &cfg);
        ok_dev[dev] = (e == cudaSuccess && ncl >= 1) ? 1 : -1;
        if (e != cudaSuccess) cudaGetLastError();
        if (knobs().dbg == 10) fprintf(stderr, "[b2d] k_rpn_front: max active clusters of %d = %d\n", kFrCl, ncl);
    }
    if (ok_dev[dev] < 0) return 0;
    const bool two = big.slots > 0 && small.slots > 0 && side && fork && join;
    if (two) {
        cudaEventRecord(fork, st);
        cudaStreamWaitEvent(side, fork, 0);
        if (int rc = front_launch_one(p, small, side)) return rc;
        cudaEventRecord(join, side);
    }
    if (big.slots > 0) { if (int rc = front_launch_one(p, big, st)) return rc; }
    if (two) cudaStreamWaitEvent(st, join, 0);
    else if (small.slots > 0) { if (int rc = front_launch_one(p, small, st)) return rc; }
    return 1;
}

}  // namespace b2d
