/*
 * det_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C, scalar fp32 restatement of the detection post-backbone hot path of
 * pengfeidip/pytorch-faster-rcnn.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (pytorch-faster-rcnn_b200/) never imports anything under oracle/.
 *
 * Parity status: PINNED.  Every function below is checked in
 * tests/test_oracle_golden.py against golden vectors produced by executing the
 * unmodified reference (torch 2.11.0 CPU + torchvision 0.26.0 CPU) with
 * tests/golden/make_golden.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 * (no FMA contraction: every fp32 operation rounds individually, in the order
 * the reference's torch expression evaluates them).
 *
 * Layout convention (reference: lib/utils.py:52, lib/region.py:64-65):
 * boxes are column-major [4, n]: row 0 = x1, 1 = y1, 2 = x2, 3 = y2.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* a1: AnchorCreator.__call__  (lib/anchor.py:107-129)                       */
/* out is [4, A, H, W]; ws/hs are the fp32 anchor sizes of lib/anchor.py:94-99 */
ORC_API void orc_anchor_grid(float *out, const float *ws, const float *hs, int A, int H, int W,
                             float stride, int center_lt) {
    const size_t plane = (size_t)A * H * W;
    for (int a = 0; a < A; ++a) {
        const float hw = ws[a] / 2.0f, hh = hs[a] / 2.0f;
        for (int y = 0; y < H; ++y) {
            float cy = (float)y * stride;          /* linspace(0, s*g, g+1)[:-1]: exact for integer strides */
            if (!center_lt) cy = cy + stride / 2.0f;
            for (int x = 0; x < W; ++x) {
                float cx = (float)x * stride;
                if (!center_lt) cx = cx + stride / 2.0f;
                const size_t i = ((size_t)a * H + y) * W + x;
                out[0 * plane + i] = cx - hw;
                out[1 * plane + i] = cy - hh;
                out[2 * plane + i] = cx + hw;
                out[3 * plane + i] = cy + hh;
            }
        }
    }
}

/* a2: inside_anchor_mask (lib/region.py:19-29) AND inside_grid_mask (:10-16)
 * for one level; mask[a*H*W + y*W + x].  border < 0 => image test is all-True. */
ORC_API void orc_valid_mask(uint8_t *mask, const float *anchors, int A, int H, int W, int img_h,
                            int img_w, float border, int in_h, int in_w) {
    const size_t plane = (size_t)A * H * W;
    for (int a = 0; a < A; ++a)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t i = ((size_t)a * H + y) * W + x;
                int ok = (y < in_h) && (x < in_w);
                if (border >= 0.0f) {
                    ok = ok && anchors[i] >= -border && anchors[plane + i] >= -border &&
                         anchors[2 * plane + i] < (float)img_w + border &&
                         anchors[3 * plane + i] < (float)img_h + border;
                }
                mask[i] = (uint8_t)ok;
            }
}

/* ------------------------------------------------------------------------ */
/* a3: calc_iou (lib/utils.py:151-172): +1 areas, strict tl<br mask, one IEEE divide */
static inline float iou_plus1(float ax1, float ay1, float ax2, float ay2, float bx1, float by1,
                              float bx2, float by2) {
    const float tlx = ax1 > bx1 ? ax1 : bx1, tly = ay1 > by1 ? ay1 : by1;
    const float brx = ax2 < bx2 ? ax2 : bx2, bry = ay2 < by2 ? ay2 : by2;
    const float dx = (brx - tlx) + 1.0f, dy = (bry - tly) + 1.0f;
    float ai = dx * dy;
    ai = ai * ((tlx < brx && tly < bry) ? 1.0f : 0.0f);
    const float aa = ((ax2 - ax1) + 1.0f) * ((ay2 - ay1) + 1.0f);
    const float ab = ((bx2 - bx1) + 1.0f) * ((by2 - by1) + 1.0f);
    return ai / ((aa + ab) - ai);
}

ORC_API void orc_calc_iou(float *out, const float *a, int64_t N, const float *b, int64_t K) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = 0; j < K; ++j)
            out[i * K + j] = iou_plus1(a[i], a[N + i], a[2 * N + i], a[3 * N + i], b[j], b[K + j],
                                       b[2 * K + j], b[3 * K + j]);
}

/* elem_iou (lib/utils.py:174-182): un-paired, no +1 */
ORC_API void orc_elem_iou(float *out, const float *a, const float *b, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        const float tlx = fmaxf(a[i], b[i]), tly = fmaxf(a[N + i], b[N + i]);
        const float brx = fminf(a[2 * N + i], b[2 * N + i]), bry = fminf(a[3 * N + i], b[3 * N + i]);
        float ai = (brx - tlx) * (bry - tly);
        ai = ai * ((tlx < brx && tly < bry) ? 1.0f : 0.0f);
        const float aa = (a[2 * N + i] - a[i]) * (a[3 * N + i] - a[N + i]);
        const float ab = (b[2 * N + i] - b[i]) * (b[3 * N + i] - b[N + i]);
        out[i] = ai / ((aa + ab) - ai);
    }
}

/* a4: MaxIoUAssigner.__call__ (lib/region.py:75-107).
 * labels: -1 ignore, 0 negative, j+1 positive for GT j.  K must be >= 1 (the
 * reference raises on K == 0).  Thresholds are fp32 (torch casts the Python
 * scalar to the tensor dtype before comparing). */
ORC_API int orc_assign_max_iou(int64_t *labels, float *out_iou, const float *boxes, int64_t N,
                               const float *gt, int64_t K, float pos_iou, float neg_iou,
                               float min_pos_iou) {
    if (K < 1) return 1;
    float *colmax = (float *)malloc(sizeof(float) * (size_t)K);
    for (int64_t j = 0; j < K; ++j) colmax[j] = -INFINITY;
    /* column max (torch.max(iou_tab, dim=0), lib/region.py:86) */
#pragma omp parallel
    {
        float *loc = (float *)malloc(sizeof(float) * (size_t)K);
        for (int64_t j = 0; j < K; ++j) loc[j] = -INFINITY;
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < N; ++i)
            for (int64_t j = 0; j < K; ++j) {
                const float v = iou_plus1(boxes[i], boxes[N + i], boxes[2 * N + i], boxes[3 * N + i],
                                          gt[j], gt[K + j], gt[2 * K + j], gt[3 * K + j]);
                if (v > loc[j]) loc[j] = v;
            }
#pragma omp critical
        for (int64_t j = 0; j < K; ++j)
            if (loc[j] > colmax[j]) colmax[j] = loc[j];
        free(loc);
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        float best = 0.0f;
        int64_t arg = 0, eq = -1;
        for (int64_t j = 0; j < K; ++j) {
            const float v = iou_plus1(boxes[i], boxes[N + i], boxes[2 * N + i], boxes[3 * N + i], gt[j],
                                      gt[K + j], gt[2 * K + j], gt[3 * K + j]);
            if (j == 0 || v > best) { best = v; arg = j; }         /* first max wins (:88) */
            if (eq < 0 && v == colmax[j] && colmax[j] >= min_pos_iou) eq = j; /* :95-97 */
        }
        int64_t lab = -1;
        if (best < neg_iou) lab = 0;                                 /* :90 */
        if (best >= pos_iou) lab = 1;                                /* :92 */
        if (eq >= 0) { arg = eq; lab = 1; }                          /* :101-104 */
        labels[i] = (lab == 1) ? arg + 1 : lab;                      /* :105-106 */
        out_iou[i] = iou_plus1(boxes[i], boxes[N + i], boxes[2 * N + i], boxes[3 * N + i], gt[arg],
                               gt[K + arg], gt[2 * K + arg], gt[3 * K + arg]); /* :102 */
    }
    free(colmax);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* a7: bbox2param (lib/utils.py:47-70) followed by the optional second
 * normalisation of lib/anchor.py:70-73 / lib/bbox.py:74-77 (norm2 != 0). */
ORC_API void orc_bbox2param(float *out, const float *base, const float *bbox, int64_t n,
                            const float *means, const float *stds, int norm2,
                            const float *means2, const float *stds2) {
    for (int64_t i = 0; i < n; ++i) {
        const float bw = (base[2 * n + i] - base[i]) + 1.0f, bh = (base[3 * n + i] - base[n + i]) + 1.0f;
        const float gw = (bbox[2 * n + i] - bbox[i]) + 1.0f, gh = (bbox[3 * n + i] - bbox[n + i]) + 1.0f;
        const float bcx = (base[2 * n + i] + base[i]) / 2.0f, bcy = (base[3 * n + i] + base[n + i]) / 2.0f;
        const float gcx = (bbox[2 * n + i] + bbox[i]) / 2.0f, gcy = (bbox[3 * n + i] + bbox[n + i]) / 2.0f;
        float t[4];
        t[0] = (gcx - bcx) / bw;
        t[1] = (gcy - bcy) / bh;
        t[2] = logf(gw / bw);
        t[3] = logf(gh / bh);
        for (int c = 0; c < 4; ++c) {
            float v = (t[c] - means[c]) / stds[c];
            if (norm2) v = (v - means2[c]) / stds2[c];
            out[c * n + i] = v;
        }
    }
}

/* a8: param2bbox / _param2bbox_ / clamp_bbox (lib/utils.py:83-92,134-144,109-120)
 * clamp_h/clamp_w < 0 => no clamp. */
ORC_API void orc_param2bbox(float *out, const float *base, const float *param, int64_t n,
                            const float *means, const float *stds, float clamp_h, float clamp_w) {
    for (int64_t i = 0; i < n; ++i) {
        const float bw = (base[2 * n + i] - base[i]) + 1.0f, bh = (base[3 * n + i] - base[n + i]) + 1.0f;
        const float bcx = (base[2 * n + i] + base[i]) / 2.0f, bcy = (base[3 * n + i] + base[n + i]) / 2.0f;
        const float tx = param[i] * stds[0] + means[0], ty = param[n + i] * stds[1] + means[1];
        const float tw = param[2 * n + i] * stds[2] + means[2], th = param[3 * n + i] * stds[3] + means[3];
        const float cx = tx * bw + bcx, cy = ty * bh + bcy;
        const float w = expf(tw) * bw, h = expf(th) * bh;
        float x1 = cx - w / 2.0f, y1 = cy - h / 2.0f, x2 = cx + w / 2.0f, y2 = cy + h / 2.0f;
        if (clamp_h >= 0.0f) {
            const float mx = clamp_w - 1.0f, my = clamp_h - 1.0f;
            x1 = fminf(fmaxf(x1, 0.0f), mx); x2 = fminf(fmaxf(x2, 0.0f), mx);
            y1 = fminf(fmaxf(y1, 0.0f), my); y2 = fminf(fmaxf(y2, 0.0f), my);
        }
        out[i] = x1; out[n + i] = y1; out[2 * n + i] = x2; out[3 * n + i] = y2;
    }
}

/* ------------------------------------------------------------------------ */
/* top-k, descending, ties -> lowest index first (torch.topk leaves tie order
 * unspecified; SURVEY 7 "Tie-breaking").  idx_out has k entries. */
typedef struct { float v; int64_t i; } orc_pair;
static int cmp_desc(const void *pa, const void *pb) {
    const orc_pair *a = (const orc_pair *)pa, *b = (const orc_pair *)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->i > b->i) - (a->i < b->i);
}
ORC_API void orc_topk_desc(int64_t *idx_out, const float *v, int64_t n, int64_t k) {
    orc_pair *p = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) { p[i].v = v[i]; p[i].i = i; }
    qsort(p, (size_t)n, sizeof(orc_pair), cmp_desc);
    for (int64_t i = 0; i < k && i < n; ++i) idx_out[i] = p[i].i;
    free(p);
}

/* ------------------------------------------------------------------------ */
/* a11: torchvision.ops.nms CPU semantics (third-party, torchvision 0.26.0:
 * stable descending score order, IoU without +1, suppress iff (double)iou > thr).
 * boxes are ROW-major [n,4] here because that is the torchvision boundary
 * (lib/heads/rpn_head.py:103 passes pred_bbox.t()).  Returns kept count. */
ORC_API int64_t orc_nms(int64_t *keep, const float *boxes, const float *scores, int64_t n, double thr) {
    if (n <= 0) return 0;
    orc_pair *p = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) { p[i].v = scores[i]; p[i].i = i; }
    qsort(p, (size_t)n, sizeof(orc_pair), cmp_desc); /* (v desc, i asc) == stable descending */
    uint8_t *sup = (uint8_t *)calloc((size_t)n, 1);
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i)
        area[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    int64_t nk = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        const int64_t i = p[_i].i;
        if (sup[i]) continue;
        keep[nk++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float ia = area[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            const int64_t j = p[_j].i;
            if (sup[j]) continue;
            const float xx1 = fmaxf(ix1, boxes[4 * j]), yy1 = fmaxf(iy1, boxes[4 * j + 1]);
            const float xx2 = fminf(ix2, boxes[4 * j + 2]), yy2 = fminf(iy2, boxes[4 * j + 3]);
            const float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
            const float inter = w * h;
            const float ovr = inter / ((ia + area[j]) - inter);
            if ((double)ovr > thr) sup[j] = 1;
        }
    }
    free(p); free(sup); free(area);
    return nk;
}

/* ------------------------------------------------------------------------ */
/* a15: map_rois_to_levels (lib/region.py:256-264) */
ORC_API void orc_level_map(int64_t *lvl, const float *rois, int64_t K, float finest_scale, int num_lvls) {
    for (int64_t i = 0; i < K; ++i) {
        const float s = sqrtf(((rois[2 * K + i] - rois[i]) + 1.0f) * ((rois[3 * K + i] - rois[K + i]) + 1.0f));
        float t = floorf(log2f(s / finest_scale + 1e-6f));
        if (t < 0.0f) t = 0.0f;
        if (t > (float)(num_lvls - 1)) t = (float)(num_lvls - 1);
        lvl[i] = (int64_t)t;
    }
}

/* a14: torchvision.ops.roi_align forward, CPU semantics (third-party,
 * torchvision 0.26.0 roi_align_kernel).  feat is NCHW [C,H,W] of
 * ONE image (the reference always passes feat[None], lib/region.py:271-276);
 * rois [4,K] column-major; out [K,C,PH,PW]. */
typedef struct { int p1, p2, p3, p4; float w1, w2, w3, w4; } orc_tap;

static void roi_taps(orc_tap *t, const float *rois, int64_t K, int64_t k, float scale, int H, int W,
                     int PH, int PW, int sr, int aligned, int *gh_out, int *gw_out) {
    const float off = aligned ? 0.5f : 0.0f;
    const float sw = rois[k] * scale - off, sh = rois[K + k] * scale - off;
    const float ew = rois[2 * K + k] * scale - off, eh = rois[3 * K + k] * scale - off;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    const float bh = rh / (float)PH, bw = rw / (float)PW;
    const int gh = sr > 0 ? sr : (int)ceilf(rh / (float)PH);
    const int gw = sr > 0 ? sr : (int)ceilf(rw / (float)PW);
    *gh_out = gh; *gw_out = gw;
    int n = 0;
    for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw)
            for (int iy = 0; iy < gh; ++iy) {
                const float yy = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
                for (int ix = 0; ix < gw; ++ix) {
                    const float xx = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                    float x = xx, y = yy;
                    orc_tap *q = &t[n++];
                    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
                        memset(q, 0, sizeof(*q));
                        continue;
                    }
                    if (y <= 0.0f) y = 0.0f;
                    if (x <= 0.0f) x = 0.0f;
                    int yl = (int)y, xl = (int)x, yh, xh;
                    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                    const float ly = y - (float)yl, lx = x - (float)xl;
                    const float hy = 1.0f - ly, hx = 1.0f - lx;
                    q->p1 = yl * W + xl; q->p2 = yl * W + xh; q->p3 = yh * W + xl; q->p4 = yh * W + xh;
                    q->w1 = hy * hx; q->w2 = hy * lx; q->w3 = ly * hx; q->w4 = ly * lx;
                }
            }
}

ORC_API void orc_roi_align_fwd(float *out, const float *feat, int C, int H, int W, const float *rois,
                               int64_t K, float scale, int PH, int PW, int sr, int aligned) {
#pragma omp parallel
    {
        orc_tap *t = NULL;
        size_t cap = 0;
#pragma omp for schedule(dynamic, 4)
        for (int64_t k = 0; k < K; ++k) {
            /* worst-case adaptive grid is bounded by the feature size */
            size_t need = (size_t)PH * PW * (size_t)(sr > 0 ? sr * sr : (H + 2) * (W + 2));
            if (need > cap) { free(t); t = (orc_tap *)malloc(sizeof(orc_tap) * need); cap = need; }
            int gh, gw;
            roi_taps(t, rois, K, k, scale, H, W, PH, PW, sr, aligned, &gh, &gw);
            const int cnt_i = gh * gw > 1 ? gh * gw : 1;
            const float cnt = (float)cnt_i;
            for (int c = 0; c < C; ++c) {
                const float *f = feat + (size_t)c * H * W;
                float *o = out + ((size_t)k * C + c) * PH * PW;
                int n = 0;
                for (int b = 0; b < PH * PW; ++b) {
                    float acc = 0.0f;
                    for (int s = 0; s < gh * gw; ++s) {
                        const orc_tap *q = &t[n++];
                        acc += q->w1 * f[q->p1] + q->w2 * f[q->p2] + q->w3 * f[q->p3] + q->w4 * f[q->p4];
                    }
                    o[b] = acc / cnt;
                }
            }
        }
        free(t);
    }
}

/* roi_align backward (autograd of the above): dfeat [C,H,W] += ...; serial
 * accumulation in (k, c, bin, sample, tap) order like torchvision's CPU kernel. */
ORC_API void orc_roi_align_bwd(float *dfeat, const float *dout, int C, int H, int W, const float *rois,
                               int64_t K, float scale, int PH, int PW, int sr, int aligned) {
    size_t need = (size_t)PH * PW * (size_t)(sr > 0 ? sr * sr : (H + 2) * (W + 2));
    orc_tap *t = (orc_tap *)malloc(sizeof(orc_tap) * need);
    for (int64_t k = 0; k < K; ++k) {
        int gh, gw;
        roi_taps(t, rois, K, k, scale, H, W, PH, PW, sr, aligned, &gh, &gw);
        const float cnt = (float)(gh * gw > 1 ? gh * gw : 1);
#pragma omp parallel for schedule(static)
        for (int c = 0; c < C; ++c) {
            float *g = dfeat + (size_t)c * H * W;
            const float *o = dout + ((size_t)k * C + c) * PH * PW;
            int n = 0;
            for (int b = 0; b < PH * PW; ++b) {
                const float go = o[b];
                for (int s = 0; s < gh * gw; ++s) {
                    const orc_tap *q = &t[n++];
                    if (q->w1 == 0.0f && q->w2 == 0.0f && q->w3 == 0.0f && q->w4 == 0.0f) continue;
                    g[q->p1] += go * q->w1 / cnt; g[q->p2] += go * q->w2 / cnt;
                    g[q->p3] += go * q->w3 / cnt; g[q->p4] += go * q->w4 / cnt;
                }
            }
        }
    }
    free(t);
}

/* K7: torchvision.ops.roi_pool forward, CPU semantics (third-party).
 * argmax [K,C,PH,PW] int32 index into the H*W plane (-1 for empty bins). */
ORC_API void orc_roi_pool_fwd(float *out, int32_t *argmax, const float *feat, int C, int H, int W,
                              const float *rois, int64_t K, float scale, int PH, int PW) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t k = 0; k < K; ++k) {
        const int sw = (int)roundf(rois[k] * scale), sh = (int)roundf(rois[K + k] * scale);
        const int ew = (int)roundf(rois[2 * K + k] * scale), eh = (int)roundf(rois[3 * K + k] * scale);
        const int rw = (ew - sw + 1) > 1 ? (ew - sw + 1) : 1, rh = (eh - sh + 1) > 1 ? (eh - sh + 1) : 1;
        const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw) {
                int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
                int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
                hs = hs + sh; he = he + sh; ws = ws + sw; we = we + sw;
                hs = hs < 0 ? 0 : (hs > H ? H : hs); he = he < 0 ? 0 : (he > H ? H : he);
                ws = ws < 0 ? 0 : (ws > W ? W : ws); we = we < 0 ? 0 : (we > W ? W : we);
                const int empty = (he <= hs) || (we <= ws);
                for (int c = 0; c < C; ++c) {
                    const float *f = feat + (size_t)c * H * W;
                    float mv = empty ? 0.0f : -FLT_MAX;
                    int mi = -1;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w)
                            if (f[h * W + w] > mv) { mv = f[h * W + w]; mi = h * W + w; }
                    const size_t o = (((size_t)k * C + c) * PH + ph) * PW + pw;
                    out[o] = mv;
                    argmax[o] = mi;
                }
            }
    }
}

/* ------------------------------------------------------------------------ */
/* a9: one level of RPNHead.predict_single_image (lib/heads/rpn_head.py:81-106):
 * score -> top pre_nms -> decode+clamp -> min-size filter -> NMS -> [:post_nms].
 * logits[n] is the 1-channel (sigmoid) objectness, deltas [4,n], anchors [4,n].
 * Selection/order is taken on the logit (sigmoid is monotone),
 * ties -> lowest index.  Outputs: boxes_out [4,cap] column-major with stride
 * cap = min(pre_nms, n), scores_out, src_idx (index into the level's anchors). */
ORC_API int64_t orc_rpn_level(float *boxes_out, float *scores_out, int64_t *src_idx, const float *logits,
                              const float *deltas, const float *anchors, int64_t n, int64_t pre_nms,
                              int64_t post_nms, double nms_thr, float min_size, const float *means,
                              const float *stds, float img_h, float img_w) {
    int64_t k = (pre_nms > 0 && pre_nms < n) ? pre_nms : n;
    int64_t *top = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    if (k < n) orc_topk_desc(top, logits, n, k);
    else for (int64_t i = 0; i < n; ++i) top[i] = i;        /* no top-k => original order */
    float *a = (float *)malloc(sizeof(float) * 4 * (size_t)(k > 0 ? k : 1));
    float *d = (float *)malloc(sizeof(float) * 4 * (size_t)(k > 0 ? k : 1));
    float *b = (float *)malloc(sizeof(float) * 4 * (size_t)(k > 0 ? k : 1));
    for (int64_t i = 0; i < k; ++i)
        for (int c = 0; c < 4; ++c) { a[c * k + i] = anchors[c * n + top[i]]; d[c * k + i] = deltas[c * n + top[i]]; }
    orc_param2bbox(b, a, d, k, means, stds, img_h, img_w);
    /* min-size filter (rpn_head.py:97-101), order preserving */
    float *rb = (float *)malloc(sizeof(float) * 4 * (size_t)(k > 0 ? k : 1));
    float *rs = (float *)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
    int64_t *ri = (int64_t *)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < k; ++i) {
        if (min_size > 0.0f) {
            const float w = (b[2 * k + i] - b[i]) + 1.0f, h = (b[3 * k + i] - b[k + i]) + 1.0f;
            if (!(w >= min_size && h >= min_size)) continue;
        }
        rb[4 * m] = b[i]; rb[4 * m + 1] = b[k + i]; rb[4 * m + 2] = b[2 * k + i]; rb[4 * m + 3] = b[3 * k + i];
        rs[m] = 1.0f / (1.0f + expf(-logits[top[i]]));
        ri[m] = top[i];
        ++m;
    }
    /* NMS runs on scores; to keep the logit order authoritative, feed a strictly
     * order-preserving surrogate: rank-based scores (descending position). */
    float *surr = (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
    if (k < n) for (int64_t i = 0; i < m; ++i) surr[i] = (float)(m - i);
    else for (int64_t i = 0; i < m; ++i) surr[i] = logits[ri[i]];
    int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1));
    int64_t nk = orc_nms(keep, rb, surr, m, nms_thr);
    if (post_nms > 0 && post_nms < nk) nk = post_nms;
    const int64_t cap = k;
    for (int64_t i = 0; i < nk; ++i) {
        const int64_t s = keep[i];
        for (int c = 0; c < 4; ++c) boxes_out[c * cap + i] = rb[4 * s + c];
        scores_out[i] = rs[s];
        src_idx[i] = ri[s];
    }
    free(top); free(a); free(d); free(b); free(rb); free(rs); free(ri); free(surr); free(keep);
    return nk;
}
