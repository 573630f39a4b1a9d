// anchor_loss.cu -- SURVEY 8(f-2): the loss reductions right after anchor_target, fused with the target
// gather.  AnchorHead.calc_loss (lib/heads/anchor_head.py:113-139) for a head without sampler (RetinaNet):
//     cls_loss = sigmoid_focal_loss(tar_cls_out.t(), tar_labels) / #pos        (lib/losses.py:33-61)
//     reg_loss = smooth_l1_loss_v2(tar_reg_out[:, pos], tar_param[:, pos], beta) / #pos   (:77-83)
// where tar_* are the gathers of anchor_target (lib/anchor.py:49-76) over all non-ignored anchors --
// 201 600 anchors x 20 classes per image at config 4, gathered into [C, s] copies by the reference.
// Here the head maps are read in place ([B, A*C, H, W] viewed (C, A*H*W), channel = class * A + a; regression
// channel = coord * A + a), the class target and the encoded deltas are rebuilt per anchor from the assignment
// labels (gt index + 1 / 0 / -1 of b2d_assign_max_iou) and the anchors are generated in registers:
//   k_anchor_loss_fwd   per-block partial sums of the focal terms, the smooth-L1 terms and #pos (deterministic
//                       two-stage reduction, final sum in double)
//   k_anchor_loss_bwd   d(loss)/d(cls map), d(loss)/d(reg map) written in the maps' own layout (recomputed from
//                       the inputs -- nothing is stored between forward and backward), scaled by device scalars
// HBM roofline: forward reads B * total * (C + 4 pos) * 4 B + 8 B labels; backward reads the same and writes the maps.
#include <cstring>

#include "common.cuh"

namespace b2d {

struct LossArgs {
    b2d_pyramid pyr;
    const float* cls[kMaxLevels]; const float* reg[kMaxLevels];
    float* dcls[kMaxLevels]; float* dreg[kMaxLevels];
    const int64_t* labels; long long label_ld;            // [B][label_ld], -1 / 0 / gt index + 1
    const float* gt; int gt_ld; const int64_t* gt_label;
    int C; float alpha, gamma, beta, ms[8];
    long long total;
};

constexpr int kLossThreads = 256;

struct AnchorRef { int l, li, n_l; };

__device__ __forceinline__ AnchorRef locate(const b2d_pyramid& pyr, long long i) {
    AnchorRef r;
    r.l = 0;
    for (int q = 1; q < pyr.num_levels; ++q) if (i >= pyr.lv[q].offset) r.l = q;
    const b2d_level& lv = pyr.lv[r.l];
    r.li = (int)(i - lv.offset);
    r.n_l = lv.A * lv.H * lv.W;
    return r;
}

// F.binary_cross_entropy_with_logits(x, y) for y in {0, 1}: (1 - y) x + m + log(exp(-m) + exp(-x - m)), m = max(-x, 0)
__device__ __forceinline__ float bce_logits(float x, bool y) {
    const float m = fmaxf(-x, 0.0f);
    return (y ? 0.0f : x) + m + logf(expf(-m) + expf(-x - m));
}
__device__ __forceinline__ float pow_gamma(float v, float gamma) { return gamma == 2.0f ? v * v : powf(v, gamma); }

// encoded + normalised deltas of (anchor, gt) (lib/utils.py:47-70), as k_encode_targets
__device__ __forceinline__ void encode4(const Box& bx, const Box& gb, const float* ms, float (&prm)[4]) {
    const float bw = (bx.x2 - bx.x1) + 1.0f, bh = (bx.y2 - bx.y1) + 1.0f;
    const float gw = (gb.x2 - gb.x1) + 1.0f, gh = (gb.y2 - gb.y1) + 1.0f;
    const float bcx = (bx.x2 + bx.x1) / 2.0f, bcy = (bx.y2 + bx.y1) / 2.0f;
    const float gcx = (gb.x2 + gb.x1) / 2.0f, gcy = (gb.y2 + gb.y1) / 2.0f;
    prm[0] = ((gcx - bcx) / bw - ms[0]) / ms[4];
    prm[1] = ((gcy - bcy) / bh - ms[1]) / ms[5];
    prm[2] = (logf(gw / bw) - ms[2]) / ms[6];
    prm[3] = (logf(gh / bh) - ms[3]) / ms[7];
}

template <bool BWD>
__global__ void __launch_bounds__(kLossThreads) k_anchor_loss(LossArgs p, float* __restrict__ partial /*[B][blocks][3]*/,
                                                              const float* __restrict__ scale /*[2] (BWD)*/) {
    __shared__ float s_red[3][kLossThreads / 32];
    const int b = blockIdx.y;
    const long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x;
    float acc_c = 0.0f, acc_r = 0.0f, npos = 0.0f;
    const float sc = BWD ? scale[0] : 0.0f, sr = BWD ? scale[1] : 0.0f;
    if (i < p.total) {
        const AnchorRef r = locate(p.pyr, i);
        const int64_t lab = p.labels[(long long)b * p.label_ld + i];
        const float* cls = p.cls[r.l] + (long long)b * p.C * r.n_l + r.li;
        float* dcls = BWD ? p.dcls[r.l] + (long long)b * p.C * r.n_l + r.li : nullptr;
        const int j = lab > 0 ? (int)(lab - 1) : 0;
        const int t = lab > 0 ? (int)p.gt_label[(long long)b * p.gt_ld + j] : 0;       // class 1..C, 0 = background
        if (lab >= 0) {
            for (int c = 0; c < p.C; ++c) {
                const float x = cls[(long long)c * r.n_l];
                const bool y = (c + 1) == t;
                const float pr = 1.0f / (1.0f + expf(-x));
                const float bce = bce_logits(x, y);
                if (!BWD) {
                    const float pt = y ? pr : 1.0f - pr;
                    const float fw = (y ? p.alpha : 1.0f - p.alpha) * pow_gamma(1.0f - pt, p.gamma);
                    acc_c += bce * fw;
                } else {
                    // d/dx [bce * w(p)]: y = 1: alpha (1-p)^g (-(1-p) - g p bce);  y = 0: (1-alpha) p^g (p + g (1-p) bce)
                    const float g = y ? p.alpha * pow_gamma(1.0f - pr, p.gamma) * (-(1.0f - pr) - p.gamma * pr * bce)
                                      : (1.0f - p.alpha) * pow_gamma(pr, p.gamma) * (pr + p.gamma * (1.0f - pr) * bce);
                    dcls[(long long)c * r.n_l] = sc * g;
                }
            }
        } else if (BWD) {
            for (int c = 0; c < p.C; ++c) dcls[(long long)c * r.n_l] = 0.0f;
        }
        const float* reg = p.reg[r.l] + (long long)b * 4 * r.n_l + r.li;
        float* dreg = BWD ? p.dreg[r.l] + (long long)b * 4 * r.n_l + r.li : nullptr;
        if (lab > 0) {
            const Box bx = anchor_flat(p.pyr.lv[r.l], r.li);
            const float* g = p.gt + (long long)b * 4 * p.gt_ld;
            const Box gb{g[j], g[p.gt_ld + j], g[2 * p.gt_ld + j], g[3 * p.gt_ld + j]};
            float prm[4];
            encode4(bx, gb, p.ms, prm);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d = reg[(long long)k * r.n_l] - prm[k];
                const float ad = fabsf(d);
                if (!BWD) acc_r += ad < p.beta ? (ad * ad) / (2.0f * p.beta) : ad - 0.5f * p.beta;
                else dreg[(long long)k * r.n_l] = sr * (ad < p.beta ? d / p.beta : (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)));
            }
            npos = 1.0f;
        } else if (BWD) {
#pragma unroll
            for (int k = 0; k < 4; ++k) dreg[(long long)k * r.n_l] = 0.0f;
        }
    }
    if (BWD) return;
    // block reduction in a fixed order -> deterministic partials
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
        acc_r += __shfl_xor_sync(0xffffffffu, acc_r, o);
        npos += __shfl_xor_sync(0xffffffffu, npos, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = acc_c; s_red[1][warp] = acc_r; s_red[2][warp] = npos; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float v = 0.0f;
        for (int w = 0; w < kLossThreads / 32; ++w) v += s_red[threadIdx.x][w];
        partial[((long long)b * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) k_loss_reduce(float* __restrict__ out /*[3]*/, const float* __restrict__ partial,
                                                     long long n) {
    __shared__ double s[3][256];
    double a[3] = {0.0, 0.0, 0.0};
    for (long long i = threadIdx.x; i < n; i += 256)
        for (int k = 0; k < 3; ++k) a[k] += (double)partial[i * 3 + k];
    for (int k = 0; k < 3; ++k) s[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int k = 0; k < 3; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 3) out[threadIdx.x] = (float)s[threadIdx.x][0];
}

}  // namespace b2d

using namespace b2d;

extern "C" {

size_t b2d_anchor_loss_workspace_bytes(const b2d_pyramid* pyr_host, int B) {
    if (!pyr_host || B < 1) return 0;
    return (size_t)B * cdiv(pyr_host->total, kLossThreads) * 3 * sizeof(float);
}

static int loss_args(LossArgs& p, const void* const* cls_ptrs_host, const void* const* reg_ptrs_host,
                     void* const* dcls_ptrs_host, void* const* dreg_ptrs_host, const b2d_pyramid* pyr_host,
                     const int64_t* labels, long long label_ld, const float* gt, int gt_ld, const int64_t* gt_label,
                     int cls_channels, float alpha, float gamma, float beta, const float* means_host,
                     const float* stds_host) {
    B2D_REQUIRE(cls_ptrs_host && reg_ptrs_host && pyr_host && labels && gt && gt_label, "anchor_loss: null pointer");
    B2D_REQUIRE(pyr_host->num_levels >= 1 && pyr_host->num_levels <= B2D_MAX_LEVELS && cls_channels >= 1 && gt_ld >= 1 &&
                label_ld >= pyr_host->total && beta > 0.0f, "anchor_loss: bad sizes");
    memset(&p, 0, sizeof(p));
    p.pyr = *pyr_host;
    for (int l = 0; l < pyr_host->num_levels; ++l) {
        p.cls[l] = (const float*)cls_ptrs_host[l]; p.reg[l] = (const float*)reg_ptrs_host[l];
        p.dcls[l] = dcls_ptrs_host ? (float*)dcls_ptrs_host[l] : nullptr;
        p.dreg[l] = dreg_ptrs_host ? (float*)dreg_ptrs_host[l] : nullptr;
    }
    p.labels = labels; p.label_ld = label_ld; p.gt = gt; p.gt_ld = gt_ld; p.gt_label = gt_label;
    p.C = cls_channels; p.alpha = alpha; p.gamma = gamma; p.beta = beta; p.total = pyr_host->total;
    for (int i = 0; i < 4; ++i) { p.ms[i] = means_host ? means_host[i] : 0.0f; p.ms[4 + i] = stds_host ? stds_host[i] : 1.0f; }
    return B2D_OK;
}

int b2d_anchor_loss_fwd(float* out3, const void* const* cls_ptrs_host, const void* const* reg_ptrs_host,
                        const b2d_pyramid* pyr_host, const int64_t* labels, long long label_ld, const float* gt, int gt_ld,
                        const int64_t* gt_label, int cls_channels, float alpha, float gamma, float beta,
                        const float* means_host, const float* stds_host, int B, void* workspace, size_t ws_bytes,
                        void* stream) {
    LossArgs p;
    if (int rc = loss_args(p, cls_ptrs_host, reg_ptrs_host, nullptr, nullptr, pyr_host, labels, label_ld, gt, gt_ld, gt_label,
                           cls_channels, alpha, gamma, beta, means_host, stds_host)) return rc;
    B2D_REQUIRE(out3 && B >= 1 && workspace && ws_bytes >= b2d_anchor_loss_workspace_bytes(pyr_host, B),
                "anchor_loss_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = cdiv(p.total, kLossThreads);
    dim3 g(blocks, B);
    k_anchor_loss<false><<<g, kLossThreads, 0, st>>>(p, (float*)workspace, nullptr);
    k_loss_reduce<<<1, 256, 0, st>>>(out3, (const float*)workspace, (long long)B * blocks);
    return check_launch("anchor_loss_fwd");
}

int b2d_anchor_loss_bwd(void* const* dcls_ptrs_host, void* const* dreg_ptrs_host, const float* scale2,
                        const void* const* cls_ptrs_host, const void* const* reg_ptrs_host, const b2d_pyramid* pyr_host,
                        const int64_t* labels, long long label_ld, const float* gt, int gt_ld, const int64_t* gt_label,
                        int cls_channels, float alpha, float gamma, float beta, const float* means_host,
                        const float* stds_host, int B, void* stream) {
    LossArgs p;
    if (int rc = loss_args(p, cls_ptrs_host, reg_ptrs_host, dcls_ptrs_host, dreg_ptrs_host, pyr_host, labels, label_ld, gt,
                           gt_ld, gt_label, cls_channels, alpha, gamma, beta, means_host, stds_host)) return rc;
    B2D_REQUIRE(dcls_ptrs_host && dreg_ptrs_host && scale2 && B >= 1, "anchor_loss_bwd: null pointer");
    dim3 g(cdiv(p.total, kLossThreads), B);
    k_anchor_loss<true><<<g, kLossThreads, 0, (cudaStream_t)stream>>>(p, nullptr, scale2);
    return check_launch("anchor_loss_bwd");
}

}  // extern "C"
