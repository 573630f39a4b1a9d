// roi_fetch.cu -- sparse host -> device feature transfer for the RoI extractor (round 2).
//
// BasicRoIExtractor (lib/region.py:299-375) reads, of a whole FPN pyramid, only the cells under the
// bilinear taps of the sampled RoIs: on BASELINE config 2 that is 59 % of the cells (512 RoIs per
// image, 482 of them ~10 x 10 cells on the finest level).  When the pyramid sits in PINNED HOST
// memory in channels-last layout (one cell = C contiguous values = 1 KB at C = 256) the cells are
// worth fetching one by one: SM-issued 16-byte loads from mapped host memory reach 51 GB/s on the
// PCIe 5 x16 link against 55 GB/s for cudaMemcpyAsync of everything (scripts/zc_probe.cu,
// profiles/r2m_zc_probe.txt), so the step ships 0.59 of the bytes at 0.93 of the speed.
//
//   k_roi_mark   one warp per RoI slot: level map + the cells under the RoI's bilinear taps -- (rows
//                under a tap) x (columns under a tap), from the same roi_level / roi_geom / axis_tap as
//                the RoIAlign kernels, so exactly the cells they read -- OR-ed into a bitmap with one
//                bit per (level, image, y, x).
//   k_fetch_cells one warp per bitmap word (32 cells): every marked cell is copied from the mapped
//                host tensor to the same offset of the device tensor (NHWC), several cells in flight
//                per warp; the number of cells moved is counted for the byte accounting of bench.py.
//
// In NCHW (the reference FPN's layout) a cell is C scattered 4-byte values and the touched runs of a
// row are ~10 cells = 40 bytes per channel plane: at the 32-byte granularity of a PCIe read 75 % of
// the pyramid would have to move in requests too small for the link -- that layout keeps the full copy.
#include "roi_common.cuh"

namespace b2d {

namespace {

struct FetchArgs {
    const char* src[kMaxLevels];
    char* dst[kMaxLevels];
    long long word0[kMaxLevels + 1];      // first bitmap word of each level (levels are padded to whole words)
    long long cells[kMaxLevels];          // B * H * W
    int num_levels;
    int chunks;                           // 16-byte chunks per cell (C * element size / 16)
};

__global__ void __launch_bounds__(256) k_roi_mark(RoiArgs a, unsigned* __restrict__ bits, FetchArgs f) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.R) return;
    const b2d_roi_cfg& c = a.cfg;
    float x1, y1, x2, y2;
    int img;
    if (!roi_fetch(a, r, img, x1, y1, x2, y2)) return;
    int lvl;
    if (a.levels) lvl = a.levels[r];
    else lvl = c.num_levels > 1 ? roi_level(x1, y1, x2, y2, c.finest_scale, c.num_levels) : 0;
    const int H = c.H[lvl], W = c.W[lvl];
    const RoiGeom g = roi_geom(x1, y1, x2, y2, c.spatial_scale[lvl], c.PH, c.PW, c.sampling_ratio, c.aligned);
    // rows: the extent of the taps (no monotonicity assumed: aligned=True may give negative bin sizes; invalid samples still
    // have clamped cell indices, which the kernels read with weight 0); a lane takes the rows ylo + lane, + 32, ...
    int ylo = H, yhi = -1;
    for (int p = 0; p < c.PH; ++p)
        for (int i = 0; i < g.gy; ++i) {
            const AxisTap t = axis_tap(g.sy, g.bh, p, i, g.gy, H);
            ylo = min(ylo, t.lo); yhi = max(yhi, t.hi);
        }
    unsigned* lb = bits + f.word0[lvl];
    for (int y = ylo + lane; y <= yhi; y += 32) {
        // exactly the tap cells: (rows under a tap) x (columns under a tap) -- RoIs with bins wider than two cells have
        // rows and columns no sample touches
        bool tap_row = false;
        for (int p = 0; p < c.PH && !tap_row; ++p)
            for (int i = 0; i < g.gy; ++i) {
                const AxisTap t = axis_tap(g.sy, g.bh, p, i, g.gy, H);
                tap_row = tap_row || t.lo == y || t.hi == y;
            }
        if (!tap_row) continue;
        const long long row0 = ((long long)img * H + y) * W;
        long long cur = -1;                                  // pending bitmap word and its bits
        unsigned m = 0;
        for (int p = 0; p < c.PW; ++p)
            for (int i = 0; i < g.gx; ++i) {
                const AxisTap t = axis_tap(g.sx, g.bw, p, i, g.gx, W);
                for (int k = 0; k < 2; ++k) {
                    const long long cell = row0 + (k ? t.hi : t.lo);
                    if ((cell >> 5) != cur) {
                        if (m) atomicOr(lb + cur, m);
                        cur = cell >> 5; m = 0;
                    }
                    m |= 1u << (cell & 31);
                }
            }
        if (m) atomicOr(lb + cur, m);
    }
}

__device__ __forceinline__ int4 ld_host(const int4* p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// U cells in flight per warp; a cell is `chunks` 16-byte pieces, lane-strided (1 KB cell: 2 per lane)
template <int U>
__global__ void __launch_bounds__(256) k_fetch_cells(FetchArgs f, const unsigned* __restrict__ bits,
                                                     unsigned long long* __restrict__ moved) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_words = f.word0[f.num_levels];
    unsigned long long mine = 0;
    for (long long wd = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wd < n_words; wd += warps) {
        unsigned m = bits[wd];
        if (!m) continue;
        int lvl = 0;
        while (lvl + 1 < f.num_levels && wd >= f.word0[lvl + 1]) ++lvl;
        const long long cell0 = (wd - f.word0[lvl]) * 32;
        const int4* src = reinterpret_cast<const int4*>(f.src[lvl]);
        int4* dst = reinterpret_cast<int4*>(f.dst[lvl]);
        mine += __popc(m);
        while (m) {
            long long off[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (m) { const int k = __ffs(m) - 1; m &= m - 1; off[u] = (cell0 + k) * f.chunks; }
                else off[u] = -1;
            }
            for (int q = lane; q < f.chunks; q += 64) {
                int4 va[U], vb[U];
                const bool two = q + 32 < f.chunks;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (off[u] >= 0) {
                        va[u] = ld_host(src + off[u] + q);
                        if (two) vb[u] = ld_host(src + off[u] + q + 32);
                    }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (off[u] >= 0) {
                        dst[off[u] + q] = va[u];
                        if (two) dst[off[u] + q + 32] = vb[u];
                    }
            }
        }
    }
    if (moved && lane == 0 && mine) atomicAdd(moved, mine);
}

int fill_fetch_args(FetchArgs& f, const b2d_roi_cfg& c, int B, int elem) {
    memset(&f, 0, sizeof(f));
    f.num_levels = c.num_levels;
    f.chunks = c.C * elem / 16;
    long long w = 0;
    for (int l = 0; l < c.num_levels; ++l) {
        f.word0[l] = w;
        f.cells[l] = (long long)B * c.H[l] * c.W[l];
        w += (f.cells[l] + 31) / 32;
    }
    f.word0[c.num_levels] = w;
    return 0;
}

}  // namespace
}  // namespace b2d

using namespace b2d;

extern "C" {

size_t b2d_roi_cell_bitmap_bytes(int B, const b2d_roi_cfg* cfg_host) {
    if (!cfg_host || B < 1 || cfg_host->num_levels < 1 || cfg_host->num_levels > B2D_MAX_LEVELS) return 0;
    FetchArgs f;
    fill_fetch_args(f, *cfg_host, B, 4);
    return (size_t)f.word0[cfg_host->num_levels] * 4;
}

int b2d_roi_mark_cells(void* bitmap, const float* rois, long long ld, const int* counts, int B,
                       const b2d_roi_cfg* cfg_host, void* stream) {
    B2D_REQUIRE(bitmap && rois && counts && cfg_host && ld >= 1 && B >= 1, "roi_mark_cells: bad args");
    const b2d_roi_cfg& c = *cfg_host;
    B2D_REQUIRE(c.num_levels >= 1 && c.num_levels <= B2D_MAX_LEVELS && c.PH >= 1 && c.PW >= 1, "roi_mark_cells: bad cfg");
    B2D_REQUIRE(c.sampling_ratio > 0, "roi_mark_cells: needs a fixed sampling_ratio");
    cudaStream_t st = (cudaStream_t)stream;
    FetchArgs f;
    fill_fetch_args(f, c, B, 4);
    if (cudaMemsetAsync(bitmap, 0, (size_t)f.word0[c.num_levels] * 4, st) != cudaSuccess) return check_launch("roi_mark_cells(memset)");
    RoiArgs a;
    memset(&a, 0, sizeof(a));
    a.cfg = c;
    a.rois = rois; a.roi_ld = ld; a.R = (long long)B * ld; a.batched_ld = ld; a.counts = counts;
    k_roi_mark<<<cdiv(a.R * 32, 256), 256, 0, st>>>(a, reinterpret_cast<unsigned*>(bitmap), f);
    return check_launch("roi_mark_cells");
}

int b2d_fetch_marked_cells(void* const* dst_ptrs_host, const void* const* src_ptrs_host, const void* bitmap, int B,
                           const b2d_roi_cfg* cfg_host, unsigned long long* moved_cells, void* stream) {
    B2D_REQUIRE(dst_ptrs_host && src_ptrs_host && bitmap && cfg_host && B >= 1, "fetch_marked_cells: bad args");
    const b2d_roi_cfg& c = *cfg_host;
    B2D_REQUIRE(c.num_levels >= 1 && c.num_levels <= B2D_MAX_LEVELS, "fetch_marked_cells: bad cfg");
    B2D_REQUIRE(c.layout == 1 || c.layout == 2, "fetch_marked_cells: channels-last features only (layout 1 or 2)");
    const int elem = c.layout == 2 ? 2 : 4;
    B2D_REQUIRE(((long long)c.C * elem) % 16 == 0, "fetch_marked_cells: a cell must be a multiple of 16 bytes");
    FetchArgs f;
    fill_fetch_args(f, c, B, elem);
    for (int l = 0; l < c.num_levels; ++l) {
        B2D_REQUIRE(dst_ptrs_host[l] && src_ptrs_host[l], "fetch_marked_cells: null level pointer");
        B2D_REQUIRE(((uintptr_t)dst_ptrs_host[l] % 16) == 0 && ((uintptr_t)src_ptrs_host[l] % 16) == 0,
                    "fetch_marked_cells: level tensors must be 16-byte aligned");
        f.src[l] = reinterpret_cast<const char*>(src_ptrs_host[l]);
        f.dst[l] = reinterpret_cast<char*>(dst_ptrs_host[l]);
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    k_fetch_cells<4><<<sms * 4, 256, 0, (cudaStream_t)stream>>>(f, reinterpret_cast<const unsigned*>(bitmap), moved_cells);
    return check_launch("fetch_marked_cells");
}

}  // extern "C"
