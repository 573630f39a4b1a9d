// roi_align_tband.cu -- K5, tensor-map TMA variant (round 2): FPN level-mapped RoIAlign forward, 2x2 samples per bin,
// NHWC fp32 features, one CTA per (RoI, 128 channels).
//
// The L1-path kernel (roi_align.cu, k_roi_align_win) keeps every byte in flight in registers and waits ~2500 cycles per
// bin with 24 warps per SM.  Here the feature rows a BIN ROW needs arrive by TMA:
//   * one CUtensorMap per (level, box width) over the level's [B][H][W][C] tensor, box = {128 channels, WB cells, 1 row};
//     the <= 4 distinct feature rows of a bin row ("band") are <= 4 cp.async.bulk.tensor.4d loads (UTMALDG) into one
//     band buffer of a shared-memory ring, completion counted on a `full` mbarrier; a dedicated producer warp runs
//     kBands - 1 bands ahead of the consumers and reuses a buffer when all consumer warps have arrived on its `empty`
//     mbarrier;
//   * consumer warp w evaluates bin (ph, w) of band ph from shared memory with the arithmetic of k_roi_align_win
//     (torchvision's order, packed adds, no FMA): bit-identical results;
//   * the [128][bins] result tile leaves as one bulk store.
// RoIs wider than 16 cells (or with PW > 8 / PH > 8) are evaluated by the same CTA straight from global memory (the
// window path), so the kernel covers every RoI in one launch.
#include <cuda.h>

#include <cstring>

#include "roi_common.cuh"

namespace b2d {
namespace {

constexpr int kTbCT = 128;                    // channels per CTA
constexpr int kTbWarps = 8;                   // consumer warps
constexpr int kTbThreads = (kTbWarps + 1) * 32;   // + 1 producer warp
constexpr int kCellBytes = kTbCT * 4;         // one cell of the channel tile in shared memory
constexpr int kRingBytes = 64 * 1024;         // 2 bands of 4 rows x 16 cells, or 4 bands of 4 rows x 8 cells
constexpr int kMaxBinsTb = 64;

struct __align__(16) TbBin {
    int ry[4];          // global byte offsets of the window rows (slow path)
    int cx[4];          // global byte offsets of the window columns (slow path)
    int sx[4];          // shared-memory byte offsets of the window columns inside a band row
    int pat, _a, _b, _c;
    float w[16];        // w1..w4 of the samples (iy, ix) = (0,0), (0,1), (1,0), (1,1)
};

struct TbMaps { CUtensorMap m[kMaxLevels][2]; };      // [level][0: 8-cell box, 1: 16-cell box]

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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// The bin sum of k_roi_align_win (bin_eval_x2): window cells from shared memory (SM) or global memory.
template <bool SM, int PY, int PX>
__device__ __forceinline__ float4 tb_eval(uint32_t sbase, uint32_t srow, const char* gb, const TbBin* t) {
    float4 v[4][4];
    if constexpr (SM) {
        const int4 sx = *reinterpret_cast<const int4*>(t->sx);
        const int sxv[4] = {sx.x, sx.y, sx.z, sx.w};
#pragma unroll
        for (int r = 0; r < PY + 2; ++r)
#pragma unroll
            for (int c = 0; c < PX + 2; ++c) v[r][c] = lds128(sbase + r * srow + (uint32_t)sxv[c]);
    } else {
        const int4 ry = *reinterpret_cast<const int4*>(t->ry);
        const int4 cx = *reinterpret_cast<const int4*>(t->cx);
        const int ryv[4] = {ry.x, ry.y, ry.z, ry.w}, cxv[4] = {cx.x, cx.y, cx.z, cx.w};
#pragma unroll
        for (int r = 0; r < PY + 2; ++r) {
            const char* rp = gb + (unsigned)ryv[r];
#pragma unroll
            for (int c = 0; c < PX + 2; ++c) v[r][c] = __ldg(reinterpret_cast<const float4*>(rp + (unsigned)cxv[c]));
        }
    }
    unsigned long long a01 = 0ull, a23 = 0ull;
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
            const int r0 = iy ? PY : 0, c0 = ix ? PX : 0;
            const float4 w = *reinterpret_cast<const float4*>(&t->w[(iy * 2 + ix) * 4]);
            const float4 &v1 = v[r0][c0], &v2 = v[r0][c0 + 1], &v3 = v[r0 + 1][c0], &v4 = v[r0 + 1][c0 + 1];
            a01 = add2(a01, add2(add2(add2(pack2(w.x * v1.x, w.x * v1.y), pack2(w.y * v2.x, w.y * v2.y)),
                                      pack2(w.z * v3.x, w.z * v3.y)), pack2(w.w * v4.x, w.w * v4.y)));
            a23 = add2(a23, add2(add2(add2(pack2(w.x * v1.z, w.x * v1.w), pack2(w.y * v2.z, w.y * v2.w)),
                                      pack2(w.z * v3.z, w.z * v3.w)), pack2(w.w * v4.z, w.w * v4.w)));
        }
    }
    float4 acc;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a01));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(a23));
    return acc;
}

template <bool SM>
__device__ __forceinline__ float4 tb_bin(int pat, uint32_t sbase, uint32_t srow, const char* gb, const TbBin* t) {
    const int py = pat / 3, px = pat - py * 3;
    if (py == 0) {
        if (px == 0) return tb_eval<SM, 0, 0>(sbase, srow, gb, t);
        if (px == 1) return tb_eval<SM, 0, 1>(sbase, srow, gb, t);
        return tb_eval<SM, 0, 2>(sbase, srow, gb, t);
    }
    if (py == 1) {
        if (px == 0) return tb_eval<SM, 1, 0>(sbase, srow, gb, t);
        if (px == 1) return tb_eval<SM, 1, 1>(sbase, srow, gb, t);
        return tb_eval<SM, 1, 2>(sbase, srow, gb, t);
    }
    if (px == 0) return tb_eval<SM, 2, 0>(sbase, srow, gb, t);
    if (px == 1) return tb_eval<SM, 2, 1>(sbase, srow, gb, t);
    return tb_eval<SM, 2, 2>(sbase, srow, gb, t);
}

__global__ void __launch_bounds__(kTbThreads, 2) k_roi_align_tband(RoiArgs a, const __grid_constant__ TbMaps maps,
                                                                   float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    // [ring kRingBytes][tile CT * bins floats][TbBin bins][band rows: PH x 4 ints][mbarriers]
    const b2d_roi_cfg& c = a.cfg;
    const int bins = c.PH * c.PW;
    unsigned char* s_ring = s_raw;
    float* s_tile = reinterpret_cast<float*>(s_raw + kRingBytes);
    TbBin* s_tab = reinterpret_cast<TbBin*>(s_tile + kTbCT * bins);
    int* s_rows = reinterpret_cast<int*>(s_tab + bins);                  // [PH][4] feature rows of the band, [PH] counts behind
    int* s_nrows = s_rows + 8 * 4;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_nrows + 8);          // full[4], empty[4]   (8-byte aligned: offsets are multiples of 16)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles = c.C / kTbCT, step = a.sel_m > 1 ? a.sel_m : 1;
    const long long n_units = ((a.R + step - 1) / step) * tiles;
    // one unit = (RoI, channel tile); an ordinary launch has one CTA per unit, the heterogeneous launch one CTA per SM
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const long long r = (unit / tiles) * step;
    const int tile = (int)(unit % tiles);
    float x1, y1, x2, y2;
    int img;
    if (!roi_fetch(a, r, img, x1, y1, x2, y2)) continue;
    int lvl;
    if (a.levels) lvl = a.levels[r];
    else lvl = c.num_levels > 1 ? roi_level(x1, y1, x2, y2, c.finest_scale, c.num_levels) : 0;
    const int H = c.H[lvl], W = c.W[lvl], C = c.C;
    const RoiGeom g = roi_geom(x1, y1, x2, y2, c.spatial_scale[lvl], c.PH, c.PW, 2, c.aligned);
    // column extent of all taps -> box width class
    const AxisTap xa = axis_tap(g.sx, g.bw, 0, 0, 2, W), xb = axis_tap(g.sx, g.bw, c.PW - 1, 1, 2, W);
    int xlo = min(xa.lo, xb.lo), xhi = max(xa.hi, xb.hi);
    bool fast = g.bw >= 0.0f && g.bh >= 0.0f && c.PW <= kTbWarps && c.PH <= 8;    // monotone taps
    if (fast) {                                                          // every sample column inside [xlo, xhi]?  (monotone: yes)
        fast = (xhi - xlo + 1) <= 16;
    }
    const int wbc = (xhi - xlo + 1) <= 8 ? 0 : 1;
    const int WB = wbc ? 16 : 8;
    const uint32_t srow = (uint32_t)WB * kCellBytes;                     // bytes of one band row
    const uint32_t band_bytes = 4u * srow;
    const int nband = kRingBytes / (int)band_bytes;                      // 4 (WB 8) or 2 (WB 16)
    if (tid < bins) {
        const int bin = tid, ph = bin / c.PW, pw = bin - ph * c.PW;
        const AxisTap ty0 = axis_tap(g.sy, g.bh, ph, 0, 2, H), ty1 = axis_tap(g.sy, g.bh, ph, 1, 2, H);
        const AxisTap tx0 = axis_tap(g.sx, g.bw, pw, 0, 2, W), tx1 = axis_tap(g.sx, g.bw, pw, 1, 2, W);
        TbBin t;
        const int dy = ty1.lo - ty0.lo, dx = tx1.lo - tx0.lo;
        const int py = (dy >= 0 && dy < 2) ? dy : 2, px = (dx >= 0 && dx < 2) ? dx : 2;
        int yy[4], xx[4];
        if (py < 2) { for (int k = 0; k < 4; ++k) yy[k] = min(ty0.lo + k, H - 1); }
        else { yy[0] = ty0.lo; yy[1] = ty0.hi; yy[2] = ty1.lo; yy[3] = ty1.hi; }
        if (px < 2) { for (int k = 0; k < 4; ++k) xx[k] = min(tx0.lo + k, W - 1); }
        else { xx[0] = tx0.lo; xx[1] = tx0.hi; xx[2] = tx1.lo; xx[3] = tx1.hi; }
        for (int k = 0; k < 4; ++k) {
            t.ry[k] = yy[k] * W * C * 4; t.cx[k] = xx[k] * C * 4;
            t.sx[k] = max(min(xx[k] - xlo, WB - 1), 0) * kCellBytes;     // (clamped: unused columns of small patterns)
        }
        const AxisTap* tys[2] = {&ty0, &ty1};
        const AxisTap* txs[2] = {&tx0, &tx1};
        for (int iy = 0; iy < 2; ++iy)
            for (int ix = 0; ix < 2; ++ix) {
                const AxisTap& ty = *tys[iy];
                const AxisTap& tx = *txs[ix];
                const bool ok = ty.valid && tx.valid;
                t.w[(iy * 2 + ix) * 4 + 0] = ok ? ty.h * tx.h : 0.0f; t.w[(iy * 2 + ix) * 4 + 1] = ok ? ty.h * tx.l : 0.0f;
                t.w[(iy * 2 + ix) * 4 + 2] = ok ? ty.l * tx.h : 0.0f; t.w[(iy * 2 + ix) * 4 + 3] = ok ? ty.l * tx.l : 0.0f;
            }
        t.pat = py * 3 + px; t._a = t._b = t._c = 0;
        s_tab[bin] = t;
        if (pw == 0) {                                                   // the band's feature rows (shared by its bins)
            for (int k = 0; k < 4; ++k) s_rows[ph * 4 + k] = yy[k];
            s_nrows[ph] = py + 2;
        }
    }
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) { mbar_init(smem_u32(s_bar + k), 1); mbar_init(smem_u32(s_bar + 4 + k), kTbWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const float* feat = reinterpret_cast<const float*>(a.feat[lvl]) + (long long)img * H * W * C;
    float* o = out + r * (long long)C * bins;
    const int c0 = tile * kTbCT;
    const uint32_t ring0 = smem_u32(s_ring);
    if (fast) {
        if (warp == kTbWarps) {                                          // ---- producer warp
            if (lane == 0) {
                const CUtensorMap* map = &maps.m[lvl][wbc];
                for (int ph = 0; ph < c.PH; ++ph) {
                    const int buf = ph % nband, use = ph / nband;
                    if (use > 0) mbar_wait(smem_u32(s_bar + 4 + buf), (uint32_t)((use - 1) & 1));
                    const int nr = s_nrows[ph];
                    const uint32_t full = smem_u32(s_bar + buf);
                    mbar_expect_tx(full, (uint32_t)nr * srow);
                    for (int k = 0; k < nr; ++k)
                        tma_load_4d(ring0 + buf * band_bytes + k * srow, map, c0, xlo, s_rows[ph * 4 + k], img, full);
                }
            }
        } else {                                                         // ---- consumer warps
            const int rot = (lane >> 3) & 3;
            const bool r1 = rot & 1, r2 = rot & 2;
            const unsigned st_base = smem_u32(s_tile + (lane * 4) * bins);
            for (int ph = 0; ph < c.PH; ++ph) {
                const int buf = ph % nband, use = ph / nband;
                // every consumer warp waits for the band, also one without a bin in it: a warp that ran ahead could arrive
                // twice in one phase of `empty` and release the buffer while another warp still reads it
                mbar_wait(smem_u32(s_bar + buf), (uint32_t)(use & 1));
                if (warp < c.PW) {
                    const int bin = ph * c.PW + warp;
                    const TbBin* t = s_tab + bin;
                    const float4 acc = tb_bin<true>(t->pat, ring0 + buf * band_bytes + lane * 16, srow, nullptr, t);
                    const float f0 = acc.x * 0.25f, f1 = acc.y * 0.25f, f2 = acc.z * 0.25f, f3 = acc.w * 0.25f;
                    const float g0 = r1 ? f1 : f0, g1 = r1 ? f2 : f1, g2 = r1 ? f3 : f2, g3 = r1 ? f0 : f3;
                    const unsigned sb = st_base + bin * 4;
                    const unsigned b4 = (unsigned)bins * 4u;
                    sts_f32(sb + ((0 + rot) & 3) * b4, r2 ? g2 : g0);        // component (s + rot) % 4 at step s: all 32 banks
                    sts_f32(sb + ((1 + rot) & 3) * b4, r2 ? g3 : g1);
                    sts_f32(sb + ((2 + rot) & 3) * b4, r2 ? g0 : g2);
                    sts_f32(sb + ((3 + rot) & 3) * b4, r2 ? g1 : g3);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(s_bar + 4 + buf));
            }
        }
    } else if (warp < kTbWarps) {                                        // ---- wide / malformed RoI: window path from global memory
        const char* gb = reinterpret_cast<const char*>(feat + c0 + lane * 4);
        for (int bin = warp; bin < bins; bin += kTbWarps) {
            const TbBin* t = s_tab + bin;
            const float4 acc = tb_bin<false>(t->pat, 0u, 0u, gb, t);
            float* st = s_tile + (lane * 4) * bins + bin;
            st[0] = acc.x * 0.25f; st[bins] = acc.y * 0.25f; st[2 * bins] = acc.z * 0.25f; st[3 * bins] = acc.w * 0.25f;
        }
    }
    // ---- the [CT][bins] tile is contiguous in shared memory and in HBM: one bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        float* dstp = o + (long long)c0 * bins;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstp), "r"(smem_u32(s_tile)),
                     "r"((unsigned)(kTbCT * bins * 4)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        for (int k = 0; k < 8; ++k) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(s_bar + k)) : "memory");
    }
    __syncthreads();                                     // tile, tables and barriers are reused by the next unit
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}

}  // namespace

// returns B2D_OK if it handled the launch, 1 if the configuration is not eligible
bool roi_hetero_streams(cudaStream_t* side, cudaEvent_t* fork, cudaEvent_t* join) {
    struct Set { cudaStream_t s; cudaEvent_t f, j; bool made, ok; };
    static Set table[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    Set& t = table[dev];
    if (!t.made) {
        t.made = true;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        t.ok = cudaStreamCreateWithPriority(&t.s, cudaStreamNonBlocking, hi) == cudaSuccess &&
               cudaEventCreateWithFlags(&t.f, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&t.j, cudaEventDisableTiming) == cudaSuccess;
        if (!t.ok) cudaGetLastError();
    }
    if (!t.ok) return false;
    *side = t.s; *fork = t.f; *join = t.j;
    return true;
}

int roi_align_tband_try(const RoiArgs& a, float* out, int B, cudaStream_t st, int persistent_ctas) {
    const b2d_roi_cfg& c = a.cfg;
    const int bins = c.PH * c.PW;
    if (c.layout != 1 || c.sampling_ratio != 2 || bins > kMaxBinsTb || c.PH > 8 || (c.C % kTbCT) != 0 || B < 1) return 1;
    if ((((long long)kTbCT * bins * 4) % 16) != 0) return 1;
    EncodeFn enc = encode_fn();
    if (!enc) return 1;
    TbMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int l = 0; l < c.num_levels; ++l) {
        if (reinterpret_cast<uintptr_t>(a.feat[l]) & 15) return 1;
        for (int k = 0; k < 2; ++k) {
            const cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W[l], (cuuint64_t)c.H[l], (cuuint64_t)B};
            const cuuint64_t strides[3] = {(cuuint64_t)c.C * 4, (cuuint64_t)c.W[l] * c.C * 4, (cuuint64_t)c.H[l] * c.W[l] * c.C * 4};
            const cuuint32_t box[4] = {(cuuint32_t)kTbCT, (cuuint32_t)(k ? 16 : 8), 1, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            const CUresult rc = enc(&maps.m[l][k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(a.feat[l]), dims, strides, box,
                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) return 1;
        }
    }
    const size_t smem = (size_t)kRingBytes + (size_t)kTbCT * bins * 4 + (size_t)bins * sizeof(TbBin) + (8 * 4 + 8) * 4 + 8 * 8 + 64;
    B2D_SMEM(k_roi_align_tband, smem, "k_roi_align_tband");
    const int step = a.sel_m > 1 ? a.sel_m : 1;
    const long long n_units = ((a.R + step - 1) / step) * (c.C / kTbCT);
    const unsigned grid = (unsigned)(persistent_ctas > 0 && persistent_ctas < n_units ? persistent_ctas : n_units);
    k_roi_align_tband<<<grid, kTbThreads, smem, st>>>(a, maps, out);
    return check_launch("roi_align_fwd(tband)");
}

}  // namespace b2d
