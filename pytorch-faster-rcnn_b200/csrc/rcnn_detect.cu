// rcnn_detect.cu -- SURVEY 8(f-3): the RCNN test-time step after RoIAlign / the head convs,
// BBoxHead.predict_bboxes_single_image (lib/heads/bbox_head.py:122-146):
//     score = softmax(cls_out, dim=1)                               [n, C]
//     preds = batched_param2bbox(props, reg_out.t(), means, stds, img_size)   (lib/utils.py:96-106;
//             reg channel = coord * C + class)
//     multiclass_nms(preds.t(), score, range(1, C), nms_iou, min_score, max_per_img, mode)
//             (lib/utils.py:224-269 -> batched_nms :211-221 -> torchvision nms)
// The reference runs ~35 small torch ops and two boolean-mask compactions for this.  Here:
//   k_rcnn_candidates  one CTA per image, a warp per proposal: softmax, candidate test
//                      ('official': every class >= 1 with score >= min_score; 'strict': the arg-max
//                      class only), per-class decode + clamp of the candidates only, ORDERED
//                      compaction (proposal-major, class-minor: the reference's row-major boolean
//                      mask order, which is what breaks score ties in the NMS sort), the maximum
//                      coordinate of the candidate set, and the class-offset boxes
//                      box + fp32(label * max) (lib/utils.py:217-219) for the NMS.
//   b2d_nms (nms.cu)   on the offset boxes, then
//   k_rcnn_gather      the first max_per_img survivors -> un-shifted boxes, scores, labels.
#include <cstring>

#include "common.cuh"

namespace b2d {

constexpr int kDetThreads = 1024;
constexpr int kDetMaxC = 128;                 // classes (incl. background) handled by one warp: 4 per lane

struct DetArgs {
    const float* props; long long ld; const int* counts; long long n;     // [B][4][ld], count per image
    const float* cls; const float* reg; int C, reg_c;                     // cls [B][ld][C]; reg [B][ld][4 * reg_c]
    float ms[8]; int clamp; const float* img_hw;
    float min_score; int strict;
    int cap;                                                              // candidate slots per image
};

__global__ void __launch_bounds__(kDetThreads) k_rcnn_candidates(DetArgs p, float4* __restrict__ cand_box,
                                                                 float4* __restrict__ nms_box,
                                                                 float* __restrict__ cand_score,
                                                                 int* __restrict__ cand_label,
                                                                 int* __restrict__ cand_count,
                                                                 int* __restrict__ overflow) {
    __shared__ int s_cnt[kDetThreads / 32], s_off[kDetThreads / 32], s_base;
    __shared__ uint32_t s_maxkey;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.counts ? p.counts[b] : (int)p.n;
    const int C = p.C;
    const float* pr = p.props + (long long)b * 4 * p.ld;
    const float* cls = p.cls + (long long)b * p.ld * C;
    const float* reg = p.reg + (long long)b * p.ld * 4 * p.reg_c;
    float4* cbox = cand_box + (long long)b * p.cap;
    float4* nbox = nms_box + (long long)b * p.cap;
    float* cscore = cand_score + (long long)b * p.cap;
    int* clabel = cand_label + (long long)b * p.cap;
    const float img_h = p.clamp ? p.img_hw[2 * b] : 0.0f, img_w = p.clamp ? p.img_hw[2 * b + 1] : 0.0f;
    if (tid == 0) { s_base = 0; s_maxkey = f2key(-INFINITY); }
    __syncthreads();
    float vmax = -INFINITY;
    for (int i0 = 0; i0 < n; i0 += kDetThreads / 32) {
        const int i = i0 + warp;
        // ---- softmax of proposal i over the C classes (4 classes per lane)
        float x[4], e[4];
        bool cand[4];
        int ncand = 0;
        if (i < n) {
            float m = -INFINITY;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = r * 32 + lane;
                x[r] = c < C ? cls[(long long)i * C + c] : -INFINITY;
                m = fmaxf(m, x[r]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < 4; ++r) { e[r] = (r * 32 + lane) < C ? expf(x[r] - m) : 0.0f; s += e[r]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#pragma unroll
            for (int r = 0; r < 4; ++r) e[r] = e[r] / s;
            if (!p.strict) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = r * 32 + lane;
                    cand[r] = c >= 1 && c < C && e[r] >= p.min_score;
                }
            } else {
                // arg-max class (first maximum, like torch.max), candidate iff it is a foreground class
                float best = -1.0f;
                int bc = 0x7fffffff;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = r * 32 + lane;
                    if (c < C && (e[r] > best)) { best = e[r]; bc = c; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                    if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = r * 32 + lane;
                    cand[r] = c == bc && c >= 1 && best >= p.min_score;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) { cand[r] = false; e[r] = 0.0f; }
        }
        unsigned bal[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { bal[r] = __ballot_sync(0xffffffffu, cand[r]); ncand += __popc(bal[r]); }
        if (lane == 0) s_cnt[warp] = ncand;
        __syncthreads();
        if (warp == 0) {                                   // exclusive scan of the 32 per-proposal counts
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
            const Box base{pr[i], pr[p.ld + i], pr[2 * p.ld + i], pr[3 * p.ld + i]};
            int before = s_off[warp];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (cand[r]) {
                    const int c = r * 32 + lane;
                    const int pos = before + __popc(bal[r] & ((1u << lane) - 1u));
                    const int rc = p.reg_c > 1 ? c : 0;                     // class-agnostic regression: one box per proposal
                    const float* q = reg + (long long)i * 4 * p.reg_c + rc;
                    const Box o = decode_box(base, q[0], q[p.reg_c], q[2 * p.reg_c], q[3 * p.reg_c], p.ms, p.clamp != 0,
                                             img_h, img_w);
                    if (pos < p.cap) {
                        cbox[pos] = make_float4(o.x1, o.y1, o.x2, o.y2);
                        cscore[pos] = e[r];
                        clabel[pos] = c;
                    }
                    vmax = fmaxf(vmax, fmaxf(fmaxf(o.x1, o.y1), fmaxf(o.x2, o.y2)));
                }
                before += __popc(bal[r]);
            }
        }
        __syncthreads();
        if (tid == 0) s_base += s_cnt[0];
        __syncthreads();
    }
    // ---- bbox.max() over the candidate set, then the class-offset boxes (lib/utils.py:217-219)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && vmax > -INFINITY) atomicMax(&s_maxkey, f2key(vmax));
    __syncthreads();
    const int total = s_base;
    const int m = min(total, p.cap);
    const float mx = key2f(s_maxkey);
    for (int t = tid; t < m; t += kDetThreads) {
        const float4 v = cbox[t];
        const float off = (float)clabel[t] * mx;               // (label * max_range).to(bbox)
        nbox[t] = make_float4(v.x + off, v.y + off, v.z + off, v.w + off);
    }
    if (tid == 0) {
        cand_count[b] = m;
        if (total > p.cap) atomicOr(overflow, 1);
    }
}

// first max_keep survivors (score order) -> [B][4][max_keep] boxes, scores, labels
__global__ void __launch_bounds__(256) k_rcnn_gather(float* __restrict__ out_box, float* __restrict__ out_score,
                                                     int64_t* __restrict__ out_label, const int64_t* __restrict__ keep,
                                                     const int* __restrict__ keep_count,
                                                     const float4* __restrict__ cand_box,
                                                     const float* __restrict__ cand_score,
                                                     const int* __restrict__ cand_label, int cap, int max_keep) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= max_keep) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.0f;
    int64_t l = 0;
    if (t < keep_count[b]) {
        const long long i = (long long)b * cap + keep[(long long)b * cap + t];
        v = cand_box[i]; s = cand_score[i]; l = cand_label[i];
    }
    float* o = out_box + (long long)b * 4 * max_keep;
    o[t] = v.x; o[max_keep + t] = v.y; o[2 * max_keep + t] = v.z; o[3 * max_keep + t] = v.w;
    out_score[(long long)b * max_keep + t] = s;
    out_label[(long long)b * max_keep + t] = l;
}

}  // namespace b2d

using namespace b2d;

extern "C" {

static size_t det_al(size_t v) { return (v + 255) & ~(size_t)255; }

size_t b2d_rcnn_detect_workspace_bytes(int cap, int B) {
    if (cap < 1 || B < 1) return 0;
    const size_t n = (size_t)B * cap;
    return det_al(n * 16) * 2 + det_al(n * 4) * 2 + det_al(n * 8) + det_al((size_t)B * 4) + det_al(4) +
           b2d_nms_workspace_bytes(cap, B);
}

int b2d_rcnn_detect(float* out_box, float* out_score, int64_t* out_label, int* out_count, const float* props,
                    long long ld, const int* counts, long long n, const float* cls_out, const float* reg_out, int C,
                    int reg_classes, const float* means_host, const float* stds_host, const float* img_hw,
                    float min_score, float nms_thr_f, int max_per_img, int strict, int cap, int B, int* overflow,
                    void* workspace, size_t ws_bytes, void* stream) {
    B2D_REQUIRE(out_box && out_score && out_label && out_count && props && cls_out && reg_out && overflow,
                "rcnn_detect: null pointer");
    B2D_REQUIRE(B >= 1 && n >= 0 && ld >= n && C >= 2 && C <= kDetMaxC && (reg_classes == C || reg_classes == 1),
                "rcnn_detect: need 2 <= C <= 128 and reg_classes in {1, C}");
    B2D_REQUIRE(cap >= 1 && cap <= 16384 && max_per_img >= 1 && max_per_img <= cap, "rcnn_detect: need max_per_img <= cap <= 16384");
    B2D_REQUIRE(workspace && ws_bytes >= b2d_rcnn_detect_workspace_bytes(cap, B), "rcnn_detect: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    const size_t nb = (size_t)B * cap;
    float4* cand_box = (float4*)w; w += det_al(nb * 16);
    float4* nms_box = (float4*)w; w += det_al(nb * 16);
    float* cand_score = (float*)w; w += det_al(nb * 4);
    int* cand_label = (int*)w; w += det_al(nb * 4);
    int64_t* keep = (int64_t*)w; w += det_al(nb * 8);
    int* cand_count = (int*)w; w += det_al((size_t)B * 4);
    w += det_al(4);
    DetArgs p;
    memset(&p, 0, sizeof(p));
    p.props = props; p.ld = ld; p.counts = counts; p.n = n;
    p.cls = cls_out; p.reg = reg_out; p.C = C; p.reg_c = reg_classes;
    for (int i = 0; i < 4; ++i) { p.ms[i] = means_host ? means_host[i] : 0.0f; p.ms[4 + i] = stds_host ? stds_host[i] : 1.0f; }
    p.clamp = img_hw ? 1 : 0; p.img_hw = img_hw;
    p.min_score = min_score; p.strict = strict; p.cap = cap;
    cudaMemsetAsync(overflow, 0, sizeof(int), st);
    k_rcnn_candidates<<<B, kDetThreads, 0, st>>>(p, cand_box, nms_box, cand_score, cand_label, cand_count, overflow);
    if (int rc = check_launch("rcnn_detect/candidates")) return rc;
    const int rc = b2d_nms(keep, out_count, (const float*)nms_box, cand_score, cap, cand_count, cap, B, nms_thr_f, max_per_img,
                           0, w, ws_bytes - (size_t)(w - (char*)workspace), st);
    if (rc != B2D_OK) return rc;
    dim3 g(cdiv(max_per_img, 256), B);
    k_rcnn_gather<<<g, 256, 0, st>>>(out_box, out_score, out_label, keep, out_count, cand_box, cand_score, cand_label, cap,
                                     max_per_img);
    return check_launch("rcnn_detect");
}

}  // extern "C"
