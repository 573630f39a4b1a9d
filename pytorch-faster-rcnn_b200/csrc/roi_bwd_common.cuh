// roi_bwd_common.cuh -- per-RoI geometry record and the (image, level) bucket lists shared by the two NHWC RoIAlign
// backward forms (roi_align_bwd_tile.cu: tile gather; roi_align_bwd_patch.cu: per-RoI patches + ordered merge).
#pragma once
#include "roi_common.cuh"

namespace b2d {

constexpr int kMaxS = 16;              // samples per axis (PH * 2, PW * 2 <= 16)
constexpr int kMaxBinsT = 64;

struct __align__(16) BwdMeta {
    int img, lvl, y0, y1;
    int x0, x1; float sx, sy;
    float bw, bh; int _p0, _p1;
};

struct MetaArgs {
    b2d_roi_cfg cfg;
    const float* rois; long long roi_ld; const int* roi_img; const int* levels; long long R;
};

__device__ __forceinline__ BwdMeta bwd_meta_of(const MetaArgs& a, long long r) {
    const b2d_roi_cfg& c = a.cfg;
    const float x1 = a.rois[r], y1 = a.rois[a.roi_ld + r], x2 = a.rois[2 * a.roi_ld + r], y2 = a.rois[3 * a.roi_ld + r];
    BwdMeta m;
    m.img = a.roi_img ? a.roi_img[r] : 0;
    m.lvl = a.levels ? a.levels[r] : (c.num_levels > 1 ? roi_level(x1, y1, x2, y2, c.finest_scale, c.num_levels) : 0);
    const int H = c.H[m.lvl], W = c.W[m.lvl];
    const RoiGeom g = roi_geom(x1, y1, x2, y2, c.spatial_scale[m.lvl], c.PH, c.PW, 2, c.aligned);
    m.sx = g.sx; m.sy = g.sy; m.bw = g.bw; m.bh = g.bh;
    // cell bounding box of all taps (sample positions need not be monotone for malformed RoIs: take min / max)
    int y0 = H, y1c = -1, x0 = W, x1c = -1;
    for (int s = 0; s < 2 * c.PH; ++s) {
        const AxisTap t = axis_tap(g.sy, g.bh, s >> 1, s & 1, 2, H);
        if (t.valid) { y0 = min(y0, t.lo); y1c = max(y1c, t.hi); }
    }
    for (int s = 0; s < 2 * c.PW; ++s) {
        const AxisTap t = axis_tap(g.sx, g.bw, s >> 1, s & 1, 2, W);
        if (t.valid) { x0 = min(x0, t.lo); x1c = max(x1c, t.hi); }
    }
    m.y0 = y0; m.y1 = y1c; m.x0 = x0; m.x1 = x1c; m._p0 = m._p1 = 0;
    return m;
}

static __global__ void __launch_bounds__(256) k_bwd_meta(MetaArgs a, BwdMeta* __restrict__ meta) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    meta[r] = bwd_meta_of(a, r);
}

// bucket[(img * L + lvl) * R + k] = k-th RoI (ascending) of that feature map; bcount[img * L + lvl]
static __global__ void __launch_bounds__(256) k_bwd_bucket(const BwdMeta* __restrict__ meta, long long R, int L,
                                                    int* __restrict__ bucket, int* __restrict__ bcount) {
    __shared__ int s_warp[8], s_base;
    const int img = blockIdx.x / L, lvl = blockIdx.x - img * L;
    int* out = bucket + (long long)blockIdx.x * R;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (long long r0 = 0; r0 < R; r0 += 256) {
        const long long r = r0 + threadIdx.x;
        const bool hit = r < R && meta[r].img == img && meta[r].lvl == lvl;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s_warp[w];
        if (hit) out[before + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = (int)r;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) bcount[blockIdx.x] = s_base;
}

}  // namespace b2d
