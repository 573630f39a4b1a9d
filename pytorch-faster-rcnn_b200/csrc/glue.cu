// glue.cu -- the remaining "index glue" of the path as kernels (round 1 ran these as eager PyTorch):
//   b2d_multiclass_nms       utils.multiclass_nms (lib/utils.py:224-269): candidate test per (box, class) or per box
//                            ('strict'), ORDERED compaction (box-major, class-minor: the reference's row-major boolean
//                            mask order, which breaks score ties in the NMS sort), bbox.max() of the candidate set,
//                            class-offset boxes box + fp32(label * max) (lib/utils.py:217-219), K4, gather.
//   b2d_batched_nms_boxes    utils.batched_nms (lib/utils.py:211-221): max of all coordinates + offset boxes.
//   b2d_scale_rois           ScalableRoICrop.scale_bbox (lib/region.py:220-225), same operation order.
//   b2d_iou_bin_ids / b2d_sample_iou_balanced   IoUBalancedNegSampler (lib/region.py:128-172): class of every element
//                            (positive / IoU bin / none) and the device-RNG selection (same keyed order as k_sample).
//   b2d_sampled_ce_fwd/bwd   CrossEntropyLoss on sampled rows (lib/losses.py:129-156: F.cross_entropy or
//                            F.binary_cross_entropy_with_logits, 'mean' over the rows, with avg_factor) read in place.
#include <cstring>

#include "common.cuh"

namespace b2d {

constexpr int kMcThreads = 1024;
constexpr int kMcMaxC = 128;                  // classes handled by one warp: 4 per lane

struct McArgs {
    const float* bbox; int box_classes;       // [n][4] (box_classes == 1) or [n][4 * C] viewed (n, 4, C)
    const float* score; long long n; int C;   // [n][C]
    unsigned long long chan[2];               // bit c: class c is in nms_channel
    float min_score;
    const float* factor;                      // [n] score_factor or NULL
    int strict, cap;
};

__device__ __forceinline__ bool mc_chan(const McArgs& p, int c) { return (p.chan[c >> 6] >> (c & 63)) & 1ull; }

// One CTA, a warp per box.  cand_* in candidate order; nms_box = cand_box + label * max(all candidate coordinates).
__global__ void __launch_bounds__(kMcThreads) k_mc_candidates(McArgs p, float4* __restrict__ cand_box,
                                                              float4* __restrict__ nms_box, float* __restrict__ cand_score,
                                                              int* __restrict__ cand_label, int* __restrict__ cand_count,
                                                              int* __restrict__ overflow) {
    __shared__ int s_cnt[kMcThreads / 32], s_off[kMcThreads / 32], s_base;
    __shared__ uint32_t s_maxkey;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.C;
    const long long n = p.n;
    if (tid == 0) { s_base = 0; s_maxkey = f2key(-INFINITY); }
    __syncthreads();
    float vmax = -INFINITY;
    for (long long i0 = 0; i0 < n; i0 += kMcThreads / 32) {
        const long long i = i0 + warp;
        float sc[4];
        bool cand[4];
        if (i < n) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = r * 32 + lane;
                sc[r] = c < C ? p.score[i * C + c] : -INFINITY;
            }
            if (!p.strict) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = r * 32 + lane;
                    cand[r] = c < C && mc_chan(p, c) && sc[r] >= p.min_score;      // tested on the RAW score (lib/utils.py:248)
                }
            } else {
                // score.max(1): first maximum over ALL channels; candidate iff that channel is an NMS channel
                float best = -INFINITY;
                int bc = 0x7fffffff;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = r * 32 + lane;
                    if (c < C && (sc[r] > best || bc == 0x7fffffff)) { best = sc[r]; bc = c; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                    if (oc != 0x7fffffff && (bc == 0x7fffffff || ob > best || (ob == best && oc < bc))) { best = ob; bc = oc; }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = r * 32 + lane;
                    cand[r] = c == bc && c < C && mc_chan(p, c) && best >= p.min_score;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) { cand[r] = false; sc[r] = 0.0f; }
        }
        unsigned bal[4];
        int ncand = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) { bal[r] = __ballot_sync(0xffffffffu, cand[r]); ncand += __popc(bal[r]); }
        if (lane == 0) s_cnt[warp] = ncand;
        __syncthreads();
        if (warp == 0) {                                   // exclusive scan of the 32 per-box counts
            const int v = s_cnt[lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_off[lane] = s_base + incl - v;
            if (lane == 31) s_cnt[0] = incl;               // round total (s_cnt is dead now)
        }
        __syncthreads();
        if (ncand > 0) {
            const float f = p.factor ? p.factor[i] : 1.0f;
            int before = s_off[warp];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (cand[r]) {
                    const int c = r * 32 + lane;
                    const int pos = before + __popc(bal[r] & ((1u << lane) - 1u));
                    float4 bx;
                    if (p.box_classes == 1) {
                        const float* q = p.bbox + i * 4;
                        bx = make_float4(q[0], q[1], q[2], q[3]);
                    } else {
                        const float* q = p.bbox + i * 4 * C + c;
                        bx = make_float4(q[0], q[C], q[2 * C], q[3 * C]);
                    }
                    if (pos < p.cap) {
                        cand_box[pos] = bx;
                        cand_score[pos] = p.factor ? sc[r] * f : sc[r];
                        cand_label[pos] = c;
                    }
                    vmax = fmaxf(vmax, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
                }
                before += __popc(bal[r]);
            }
        }
        __syncthreads();
        if (tid == 0) s_base += s_cnt[0];
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && vmax > -INFINITY) atomicMax(&s_maxkey, f2key(vmax));
    __syncthreads();
    const int total = s_base;
    const int m = min(total, p.cap);
    const float mx = key2f(s_maxkey);
    for (int t = tid; t < m; t += kMcThreads) {
        const float4 v = cand_box[t];
        const float off = (float)cand_label[t] * mx;           // (label * max_range).to(bbox)
        nms_box[t] = make_float4(v.x + off, v.y + off, v.z + off, v.w + off);
    }
    if (tid == 0) {
        cand_count[0] = m;
        if (total > p.cap) atomicOr(overflow, 1);
    }
}

// survivors (score order) -> row-major [max_keep][4] boxes, scores, labels
__global__ void __launch_bounds__(256) k_mc_gather(float* __restrict__ out_box, float* __restrict__ out_score,
                                                   int64_t* __restrict__ out_label, const int64_t* __restrict__ keep,
                                                   const int* __restrict__ keep_count, const float4* __restrict__ cand_box,
                                                   const float* __restrict__ cand_score, const int* __restrict__ cand_label,
                                                   int max_keep) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= max_keep) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.0f;
    int64_t l = 0;
    if (t < keep_count[0]) {
        const long long i = keep[t];
        v = cand_box[i]; s = cand_score[i]; l = cand_label[i];
    }
    reinterpret_cast<float4*>(out_box)[t] = v;
    out_score[t] = s;
    out_label[t] = l;
}

// batched_nms: max over all coordinates, then out = bbox + fp32(label * max)
__global__ void __launch_bounds__(1024) k_offset_boxes(float4* __restrict__ out, const float4* __restrict__ bbox,
                                                       const int64_t* __restrict__ label, long long n) {
    __shared__ uint32_t s_maxkey;
    if (threadIdx.x == 0) s_maxkey = f2key(-INFINITY);
    __syncthreads();
    float vmax = -INFINITY;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float4 v = bbox[i];
        vmax = fmaxf(vmax, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0 && vmax > -INFINITY) atomicMax(&s_maxkey, f2key(vmax));
    __syncthreads();
    const float mx = key2f(s_maxkey);
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float4 v = bbox[i];
        const float off = (float)label[i] * mx;
        out[i] = make_float4(v.x + off, v.y + off, v.z + off, v.w + off);
    }
}

// lib/region.py:220-225, [4][ld] layout: centre (x2 + x1) / 2, half size (x2 - x1 + 1) / 2 * scale
__global__ void __launch_bounds__(256) k_scale_rois(float* __restrict__ out, const float* __restrict__ rois, long long ld,
                                                    long long n, float scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x1 = rois[i], y1 = rois[ld + i], x2 = rois[2 * ld + i], y2 = rois[3 * ld + i];
    const float cx = (x2 + x1) / 2.0f, cy = (y2 + y1) / 2.0f;
    const float hw = (((x2 - x1) + 1.0f) / 2.0f) * scale, hh = (((y2 - y1) + 1.0f) / 2.0f) * scale;
    out[i] = cx - hw; out[ld + i] = cy - hh; out[2 * ld + i] = cx + hw; out[3 * ld + i] = cy + hh;
}

// ---- IoU-balanced negative sampler -----------------------------------------------------------------------------
// class of element i: 0 = positive (label > 0); 1 + j = negative whose IoU lies in bin j, bins ordered from the
// HIGHEST IoU range down as the reference walks them (lib/region.py:152-160); -1 = not a candidate.
constexpr int kIbMaxBins = 15;
struct IbArgs {
    const int64_t* labels; const float* iou; long long n;
    int num_bins;
    float lo[kIbMaxBins], hi[kIbMaxBins];     // fp32 bounds of bin j (highest first): lo <= iou < hi
};

__device__ __forceinline__ int ib_class(const IbArgs& p, long long i) {
    const int64_t lab = p.labels[i];
    if (lab > 0) return 0;
    if (lab != 0) return -1;
    const float v = p.iou[i];
    for (int j = 0; j < p.num_bins; ++j)
        if (v >= p.lo[j] && v < p.hi[j]) return 1 + j;
    return -1;
}

__global__ void __launch_bounds__(256) k_iou_bin_ids(int* __restrict__ ids, IbArgs p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.n) ids[i] = ib_class(p, i);
}

// One CTA.  Per class c a quota: positives pos_num; bin j < last: int(num_neg / num_bins); last bin: what is left of
// num_neg (num_neg = max_num - kept positives).  A class over its quota keeps the quota smallest
// (mix_key(sd + c, i) << 32 | i); out = labels at the kept places, -1 elsewhere.
__global__ void __launch_bounds__(1024) k_sample_iou_balanced(int64_t* __restrict__ out, IbArgs p, int max_num, int pos_num,
                                                              unsigned long long seed) {
    __shared__ int s_cnt[kIbMaxBins + 1], s_quota[kIbMaxBins + 1];
    __shared__ unsigned long long s_lo[kIbMaxBins + 1], s_hi[kIbMaxBins + 1];
    __shared__ int s_le[kIbMaxBins + 1];
    __shared__ int s_active;
    const int tid = threadIdx.x;
    const int nc = p.num_bins + 1;
    if (tid <= kIbMaxBins) s_cnt[tid] = 0;
    __syncthreads();
    for (long long i = tid; i < p.n; i += blockDim.x) {
        const int c = ib_class(p, i);
        if (c >= 0) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    if (tid == 0) {
        const int kept_pos = min(s_cnt[0], pos_num);
        s_quota[0] = pos_num;
        const int num_neg = max_num - kept_pos;
        const int per_bin = num_neg / p.num_bins;            // int(num_neg / num_bins), num_neg >= 0
        int chosen = 0;
        for (int j = 0; j < p.num_bins; ++j) {
            const int allowed = j < p.num_bins - 1 ? per_bin : num_neg - chosen;
            s_quota[1 + j] = allowed;
            chosen += min(s_cnt[1 + j], allowed);
        }
        int act = 0;
        for (int c = 0; c < nc; ++c) {
            s_lo[c] = 0ull; s_hi[c] = ~0ull;
            if (s_cnt[c] > s_quota[c] && s_quota[c] > 0) act = 1;
        }
        s_active = act;
    }
    __syncthreads();
    // bisection on the 64-bit composite, all over-quota classes at once: smallest thr_c with #{comp <= thr_c} >= quota_c
    if (s_active) {
        for (int it = 0; it < 64; ++it) {
            if (tid < nc) s_le[tid] = 0;
            __syncthreads();
            for (long long i = tid; i < p.n; i += blockDim.x) {
                const int c = ib_class(p, i);
                if (c < 0 || s_cnt[c] <= s_quota[c] || s_quota[c] <= 0) continue;
                const unsigned long long mid = s_lo[c] + (s_hi[c] - s_lo[c]) / 2;
                const unsigned long long comp = ((unsigned long long)mix_key(seed + (unsigned long long)c, (uint64_t)i) << 32) | (uint32_t)i;
                if (comp <= mid) atomicAdd(&s_le[c], 1);
            }
            __syncthreads();
            if (tid < nc && s_cnt[tid] > s_quota[tid] && s_quota[tid] > 0 && s_lo[tid] < s_hi[tid]) {
                const unsigned long long mid = s_lo[tid] + (s_hi[tid] - s_lo[tid]) / 2;
                if (s_le[tid] >= s_quota[tid]) s_hi[tid] = mid; else s_lo[tid] = mid + 1;
            }
            __syncthreads();
        }
    }
    for (long long i = tid; i < p.n; i += blockDim.x) {
        const int c = ib_class(p, i);
        bool keep = false;
        if (c >= 0 && s_quota[c] > 0) {
            if (s_cnt[c] <= s_quota[c]) keep = true;
            else {
                const unsigned long long comp = ((unsigned long long)mix_key(seed + (unsigned long long)c, (uint64_t)i) << 32) | (uint32_t)i;
                keep = comp <= s_lo[c];
            }
        }
        out[i] = keep ? p.labels[i] : (int64_t)-1;
    }
}

// ---- cross entropy on sampled rows ------------------------------------------------------------------------------
// logits [C][ld] (class-major columns: the reference's tar_cls_out, lib/anchor.py:55) or [rows][C] (row_major: the
// RCNN head's cls_out), target int64[rows]; rows with target < 0 are ignored.  sigmoid: C == 1 binary cross entropy
// with logits (C > 1: against the one-hot of label - 1); else softmax cross entropy.  out[0] = sum of the row losses, out[1] = rows counted.
struct CeArgs {
    const float* logits; long long ld; int C; int row_major; int sigmoid;
    const int64_t* target; long long rows;
};

__device__ __forceinline__ float ce_logit(const CeArgs& p, long long r, int c) {
    return p.row_major ? p.logits[r * p.ld + c] : p.logits[(long long)c * p.ld + r];
}

__global__ void __launch_bounds__(256) k_ce_fwd(CeArgs p, float* __restrict__ out) {
    __shared__ float s_sum[8];
    __shared__ int s_n[8];
    float acc = 0.0f;
    int cnt = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < p.rows; r += (long long)gridDim.x * blockDim.x) {
        const int64_t t = p.target[r];
        if (t < 0) continue;
        float loss;
        if (p.sigmoid) {
            // max(x, 0) - x * t + log1p(exp(-|x|))  (F.binary_cross_entropy_with_logits); C > 1: one-hot of label - 1
            loss = 0.0f;
            for (int c = 0; c < p.C; ++c) {
                const float x = ce_logit(p, r, c), tf = p.C == 1 ? (float)t : (t == c + 1 ? 1.0f : 0.0f);
                loss += fmaxf(x, 0.0f) - x * tf + log1pf(expf(-fabsf(x)));
            }
        } else {
            float m = -INFINITY;
            for (int c = 0; c < p.C; ++c) m = fmaxf(m, ce_logit(p, r, c));
            float s = 0.0f;
            for (int c = 0; c < p.C; ++c) s += expf(ce_logit(p, r, c) - m);
            loss = (m + logf(s)) - ce_logit(p, r, (int)t);
        }
        acc += loss;
        ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = acc; s_n[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.0f; int n = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_sum[w]; n += s_n[w]; }
        out[2 * blockIdx.x] = a; out[2 * blockIdx.x + 1] = (float)n;
    }
}

// grad[r][c] (same layout as logits) = scale * (p_c - [c == t])   (softmax) / scale * (sigmoid(x) - t); 0 for ignored rows
__global__ void __launch_bounds__(256) k_ce_bwd(CeArgs p, float* __restrict__ grad, const float* __restrict__ scale) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.rows) return;
    const int64_t t = p.target[r];
    const float g = scale[0];
    auto put = [&](int c, float v) {
        if (p.row_major) grad[r * p.ld + c] = v; else grad[(long long)c * p.ld + r] = v;
    };
    if (t < 0) { for (int c = 0; c < p.C; ++c) put(c, 0.0f); return; }
    if (p.sigmoid) {
        for (int c = 0; c < p.C; ++c) {
            const float x = ce_logit(p, r, c), tf = p.C == 1 ? (float)t : (t == c + 1 ? 1.0f : 0.0f);
            put(c, g * (1.0f / (1.0f + expf(-x)) - tf));
        }
        return;
    }
    float m = -INFINITY;
    for (int c = 0; c < p.C; ++c) m = fmaxf(m, ce_logit(p, r, c));
    float s = 0.0f;
    for (int c = 0; c < p.C; ++c) s += expf(ce_logit(p, r, c) - m);
    for (int c = 0; c < p.C; ++c) put(c, g * (expf(ce_logit(p, r, c) - m) / s - (c == (int)t ? 1.0f : 0.0f)));
}

// ---- GA-RPN call sites (lib/heads/guided_head.py:621-669): explicit per-location anchors + a location mask ----
struct GaLevels {
    const float* cls[kMaxLevels];          // [n_l] logits
    const unsigned char* mask[kMaxLevels]; // [n_l] torch.bool
    const float* anchor[kMaxLevels];       // [4][n_l]
    const float* reg[kMaxLevels];          // [4][n_l]
    int n[kMaxLevels];
    int L;
};

// out[l][i] = logit of location i of level l if it is inside the level and unmasked, else -inf (one launch for all levels)
__global__ void __launch_bounds__(256) k_ga_pack_scores(GaLevels g, float* __restrict__ out, long long ld) {
    const int l = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (long long)gridDim.x * blockDim.x)
        out[(long long)l * ld + i] = (i < g.n[l] && g.mask[l][i]) ? g.cls[l][i] : -INFINITY;
}

// per selected location (l, j) of the segmented top-k: gather its anchor and deltas, decode + clamp (utils.param2bbox with
// img_size), sigmoid score; rows whose location is masked out (-inf logit, they sort last) or missing (-1) are invalid.
// A box below min_size keeps its row but gets score -inf: the NMS behind this sorts it to the end of its level.
__global__ void __launch_bounds__(256) k_ga_decode(GaLevels g, const int* __restrict__ idx, const float* __restrict__ packed,
                                                   long long ld, int k, float4 ms_lo, float4 ms_hi, float img_h, float img_w,
                                                   float min_size, float* __restrict__ box /* [L][k][4] */,
                                                   float* __restrict__ score /* [L][k] */, int* __restrict__ nvalid /* [L] */) {
    const int l = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const int i = idx[(long long)l * k + j];
    const float logit = i >= 0 ? packed[(long long)l * ld + i] : -INFINITY;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = -INFINITY;
    if (i >= 0 && logit > -INFINITY) {
        const long long n = g.n[l];
        const Box a{g.anchor[l][i], g.anchor[l][n + i], g.anchor[l][2 * n + i], g.anchor[l][3 * n + i]};
        const float ms[8] = {ms_lo.x, ms_lo.y, ms_lo.z, ms_lo.w, ms_hi.x, ms_hi.y, ms_hi.z, ms_hi.w};
        const Box d = decode_box(a, g.reg[l][i], g.reg[l][n + i], g.reg[l][2 * n + i], g.reg[l][3 * n + i], ms, true, img_h, img_w);
        o = make_float4(d.x1, d.y1, d.x2, d.y2);
        const bool small = min_size > 0.0f && !(((d.x2 - d.x1) + 1.0f >= min_size) && ((d.y2 - d.y1) + 1.0f >= min_size));
        sc = small ? -INFINITY : 1.0f / (1.0f + expf(-logit));
        atomicAdd(&nvalid[l], 1);
    }
    reinterpret_cast<float4*>(box)[(long long)l * k + j] = o;
    score[(long long)l * k + j] = sc;
}

}  // namespace b2d

using namespace b2d;

extern "C" {

static size_t glue_al(size_t v) { return (v + 255) & ~(size_t)255; }

size_t b2d_multiclass_nms_workspace_bytes(int cap) {
    if (cap < 1) return 0;
    const size_t n = (size_t)cap;
    return glue_al(n * 16) * 2 + glue_al(n * 4) * 2 + glue_al(n * 8) + glue_al(4) + b2d_nms_workspace_bytes(cap, 1);
}

int b2d_multiclass_nms(float* out_box, float* out_score, int64_t* out_label, int* out_count, const float* bbox,
                       int box_classes, const float* score, long long n, int C, const unsigned long long* chan_mask_host,
                       float min_score, const float* score_factor, int strict, float nms_thr_f, int max_keep, int cap,
                       int* overflow, void* workspace, size_t ws_bytes, void* stream) {
    B2D_REQUIRE(out_box && out_score && out_label && out_count && bbox && score && chan_mask_host && overflow,
                "multiclass_nms: null pointer");
    B2D_REQUIRE(n >= 0 && C >= 1 && C <= kMcMaxC && (box_classes == 1 || box_classes == C),
                "multiclass_nms: need 1 <= C <= 128 and per-class boxes for all C classes or one box per row");
    B2D_REQUIRE(cap >= 1 && cap <= 16384 && max_keep >= 1 && max_keep <= cap, "multiclass_nms: need max_keep <= cap <= 16384");
    B2D_REQUIRE(workspace && ws_bytes >= b2d_multiclass_nms_workspace_bytes(cap), "multiclass_nms: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    const size_t nb = (size_t)cap;
    float4* cand_box = (float4*)w; w += glue_al(nb * 16);
    float4* nms_box = (float4*)w; w += glue_al(nb * 16);
    float* cand_score = (float*)w; w += glue_al(nb * 4);
    int* cand_label = (int*)w; w += glue_al(nb * 4);
    int64_t* keep = (int64_t*)w; w += glue_al(nb * 8);
    int* cand_count = (int*)w; w += glue_al(4);
    McArgs p;
    memset(&p, 0, sizeof(p));
    p.bbox = bbox; p.box_classes = box_classes; p.score = score; p.n = n; p.C = C;
    p.chan[0] = chan_mask_host[0]; p.chan[1] = chan_mask_host[1];
    p.min_score = min_score; p.factor = score_factor; p.strict = strict; p.cap = cap;
    cudaMemsetAsync(overflow, 0, sizeof(int), st);
    k_mc_candidates<<<1, kMcThreads, 0, st>>>(p, cand_box, nms_box, cand_score, cand_label, cand_count, overflow);
    if (int rc = check_launch("multiclass_nms/candidates")) return rc;
    const int rc = b2d_nms(keep, out_count, (const float*)nms_box, cand_score, cap, cand_count, cap, 1, nms_thr_f, max_keep, 0,
                           w, ws_bytes - (size_t)(w - (char*)workspace), st);
    if (rc != B2D_OK) return rc;
    k_mc_gather<<<cdiv(max_keep, 256), 256, 0, st>>>(out_box, out_score, out_label, keep, out_count, cand_box, cand_score,
                                                     cand_label, max_keep);
    return check_launch("multiclass_nms");
}

// The candidate stage alone, into caller-provided arrays of `cap` entries (any size): for candidate sets beyond the
// 16384 boxes of b2d_nms the caller runs the NMS of its choice on nms_box / cand_score and gathers itself.
int b2d_multiclass_candidates(float* cand_box, float* nms_box, float* cand_score, int* cand_label, int* cand_count,
                              int* overflow, const float* bbox, int box_classes, const float* score, long long n, int C,
                              const unsigned long long* chan_mask_host, float min_score, const float* score_factor,
                              int strict, int cap, void* stream) {
    B2D_REQUIRE(cand_box && nms_box && cand_score && cand_label && cand_count && overflow && bbox && score && chan_mask_host,
                "multiclass_candidates: null pointer");
    B2D_REQUIRE(n >= 0 && C >= 1 && C <= kMcMaxC && (box_classes == 1 || box_classes == C) && cap >= 1,
                "multiclass_candidates: need 1 <= C <= 128 and per-class boxes for all C classes or one box per row");
    cudaStream_t st = (cudaStream_t)stream;
    McArgs p;
    memset(&p, 0, sizeof(p));
    p.bbox = bbox; p.box_classes = box_classes; p.score = score; p.n = n; p.C = C;
    p.chan[0] = chan_mask_host[0]; p.chan[1] = chan_mask_host[1];
    p.min_score = min_score; p.factor = score_factor; p.strict = strict; p.cap = cap;
    cudaMemsetAsync(overflow, 0, sizeof(int), st);
    k_mc_candidates<<<1, kMcThreads, 0, st>>>(p, (float4*)cand_box, (float4*)nms_box, cand_score, cand_label, cand_count, overflow);
    return check_launch("multiclass_candidates");
}

int b2d_batched_nms_boxes(float* out, const float* bbox, const int64_t* label, long long n, void* stream) {
    B2D_REQUIRE(out && bbox && label && n >= 0, "batched_nms_boxes: bad args");
    if (n == 0) return B2D_OK;
    k_offset_boxes<<<1, 1024, 0, (cudaStream_t)stream>>>((float4*)out, (const float4*)bbox, label, n);
    return check_launch("batched_nms_boxes");
}

int b2d_scale_rois(float* out, const float* rois, long long ld, long long n, float scale, void* stream) {
    B2D_REQUIRE(out && rois && ld >= n && n >= 0, "scale_rois: bad args");
    if (n == 0) return B2D_OK;
    k_scale_rois<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(out, rois, ld, n, scale);
    return check_launch("scale_rois");
}

static bool ib_fill(IbArgs& p, const int64_t* labels, const float* iou, long long n, int num_bins, const float* lo, const float* hi) {
    if (!labels || !iou || n < 0 || num_bins < 1 || num_bins > kIbMaxBins || !lo || !hi) return false;
    memset(&p, 0, sizeof(p));
    p.labels = labels; p.iou = iou; p.n = n; p.num_bins = num_bins;
    for (int j = 0; j < num_bins; ++j) { p.lo[j] = lo[j]; p.hi[j] = hi[j]; }
    return true;
}

int b2d_iou_bin_ids(int* ids, const int64_t* labels, const float* iou, long long n, int num_bins, const float* bin_lo_host,
                    const float* bin_hi_host, void* stream) {
    IbArgs p;
    B2D_REQUIRE(ids && ib_fill(p, labels, iou, n, num_bins, bin_lo_host, bin_hi_host), "iou_bin_ids: bad args (1..15 bins)");
    if (n == 0) return B2D_OK;
    k_iou_bin_ids<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(ids, p);
    return check_launch("iou_bin_ids");
}

int b2d_sample_iou_balanced(int64_t* out_labels, const int64_t* labels, const float* iou, long long n, int max_num,
                            int pos_num, int num_bins, const float* bin_lo_host, const float* bin_hi_host,
                            unsigned long long seed, void* stream) {
    IbArgs p;
    B2D_REQUIRE(out_labels && ib_fill(p, labels, iou, n, num_bins, bin_lo_host, bin_hi_host) && max_num >= pos_num && pos_num >= 0 &&
                n < (1ll << 32), "sample_iou_balanced: bad args (1..15 bins, pos_num <= max_num)");
    if (n == 0) return B2D_OK;
    k_sample_iou_balanced<<<1, 1024, 0, (cudaStream_t)stream>>>(out_labels, p, max_num, pos_num, seed);
    return check_launch("sample_iou_balanced");
}

static bool ce_fill(CeArgs& p, const float* logits, long long ld, int C, int row_major, int sigmoid, const int64_t* target,
                    long long rows) {
    if (!logits || !target || rows < 0 || C < 1 || ld < (row_major ? C : rows)) return false;
    memset(&p, 0, sizeof(p));
    p.logits = logits; p.ld = ld; p.C = C; p.row_major = row_major; p.sigmoid = sigmoid; p.target = target; p.rows = rows;
    return true;
}

constexpr int kCeBlocks = 64;
size_t b2d_sampled_ce_workspace_bytes(void) { return (size_t)kCeBlocks * 2 * sizeof(float); }

int b2d_sampled_ce_fwd(float* partial /* [64][2]: per-block (loss sum, rows) */, const float* logits, long long ld, int C,
                       int row_major, int sigmoid, const int64_t* target, long long rows, void* stream) {
    CeArgs p;
    B2D_REQUIRE(partial && ce_fill(p, logits, ld, C, row_major, sigmoid, target, rows), "sampled_ce_fwd: bad args");
    k_ce_fwd<<<kCeBlocks, 256, 0, (cudaStream_t)stream>>>(p, partial);
    return check_launch("sampled_ce_fwd");
}

int b2d_sampled_ce_bwd(float* grad, const float* scale, const float* logits, long long ld, int C, int row_major, int sigmoid,
                       const int64_t* target, long long rows, void* stream) {
    CeArgs p;
    B2D_REQUIRE(grad && scale && ce_fill(p, logits, ld, C, row_major, sigmoid, target, rows), "sampled_ce_bwd: bad args");
    if (rows == 0) return B2D_OK;
    k_ce_bwd<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(p, grad, scale);
    return check_launch("sampled_ce_bwd");
}

static bool ga_fill(GaLevels& g, const void* const* cls, const void* const* mask, const void* const* anchor, const void* const* reg,
                    const int* n_host, int L) {
    if (!cls || !mask || !n_host || L < 1 || L > kMaxLevels) return false;
    memset(&g, 0, sizeof(g));
    g.L = L;
    for (int l = 0; l < L; ++l) {
        g.cls[l] = (const float*)cls[l]; g.mask[l] = (const unsigned char*)mask[l];
        g.anchor[l] = anchor ? (const float*)anchor[l] : nullptr; g.reg[l] = reg ? (const float*)reg[l] : nullptr;
        g.n[l] = n_host[l];
        if (!g.cls[l] || !g.mask[l] || g.n[l] < 0) return false;
    }
    return true;
}

int b2d_ga_pack_scores(float* out, long long ld, const void* const* cls_ptrs_host, const void* const* mask_ptrs_host,
                       const int* n_host, int L, void* stream) {
    GaLevels g;
    B2D_REQUIRE(out && ld >= 1 && ga_fill(g, cls_ptrs_host, mask_ptrs_host, nullptr, nullptr, n_host, L), "ga_pack_scores: bad args");
    k_ga_pack_scores<<<dim3(cdiv(ld, 256), L), 256, 0, (cudaStream_t)stream>>>(g, out, ld);
    return check_launch("ga_pack_scores");
}

int b2d_ga_decode(float* box, float* score, int* nvalid, const int* idx, const float* packed, long long ld, int k,
                  const void* const* cls_ptrs_host, const void* const* mask_ptrs_host, const void* const* anchor_ptrs_host,
                  const void* const* reg_ptrs_host, const int* n_host, int L, const float* means_host, const float* stds_host,
                  float img_h, float img_w, float min_size, void* stream) {
    GaLevels g;
    B2D_REQUIRE(box && score && nvalid && idx && packed && k >= 1 && anchor_ptrs_host && reg_ptrs_host && means_host && stds_host &&
                ga_fill(g, cls_ptrs_host, mask_ptrs_host, anchor_ptrs_host, reg_ptrs_host, n_host, L), "ga_decode: bad args");
    for (int l = 0; l < L; ++l) B2D_REQUIRE(g.anchor[l] && g.reg[l], "ga_decode: null level pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(nvalid, 0, sizeof(int) * L, st) != cudaSuccess) return check_launch("ga_decode(memset)");
    k_ga_decode<<<dim3(cdiv(k, 256), L), 256, 0, st>>>(g, idx, packed, ld, k,
                                                      make_float4(means_host[0], means_host[1], means_host[2], means_host[3]),
                                                      make_float4(stds_host[0], stds_host[1], stds_host[2], stds_host[3]),
                                                      img_h, img_w, min_size, box, score, nvalid);
    return check_launch("ga_decode");
}


}  // extern "C"
