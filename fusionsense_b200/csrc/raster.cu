// raster.cu — tile rasterisation / alpha compositing, forward and backward, load-balanced by list segments.
//
// Replaces gsplat 1.0.0's rasterize_to_pixels_{fwd,bwd}_kernel and the legacy rasterize_forward /
// rasterize_backward_kernel (Appendix A.5/A.6), reached from
// /root/reference/dn_splatter/dn_model.py:570-591 (RGB + expected depth, D = 4) and :644-653
// (normals, D = 3, white background).
//
// Why segments: real scenes give a few tiles depth-sorted lists 50x longer than the median (object
// silhouettes: on the 300k-Gaussian bench scene the median tile holds 108 entries, the 99th percentile 5171,
// the longest 7589), and one-CTA-per-tile leaves 148 SMs waiting for the handful that own them.  Every tile's
// list is cut into segments of SEG entries and the unit of work is one (tile, segment) CTA.  Compositing is an
// associative scan over (C, T) pairs — (C1, T1) o (C2, T2) = (C1 + T1 C2, T1 T2) — so:
//
//   forward A   raster_seg_kernel: every segment composites its entries from T = 1 with the reference's own
//               per-pixel rules (alpha test, stop when T (1 - alpha) <= 1e-4).  A pixel that stops locally would
//               also stop when started from any T_in <= 1, so its state is flagged "saturated" (negative T).
//               Tiles with a single segment (97 % of them) are finished here.
//   forward B   raster_fold_kernel: one CTA per multi-segment tile folds the segment states in order,
//               C += T C_loc, T *= T_loc — pure arithmetic.  Where the stop rule fires inside a segment
//               (T T_loc <= 1e-4 or the segment saturated locally) the pixel is marked for stage C.
//   forward C   raster_stop_kernel: the marked pixels re-walk that one segment exactly from their incoming
//               state, so the sequential per-pixel semantics (stop position, last id) are kept.  Every pixel
//               stops at most once, so all re-walks are independent and run in parallel.
//               No CTA ever waits for another one (round-1 ncu of a chained look-back version: half of all
//               warp samples sat in the flag spin loop; of a fold-with-inline-re-walk version: 310 us on the
//               critical path of the longest tile).
//   backward    needs no chain at all: the forward leaves, per (segment, pixel), the transmittance after the
//               segment and the colour accumulated through it, so every segment replays independently.
//
// Inside a CTA (tile_size x tile_size threads, one pixel each, a warp owns an 8 x 4 pixel footprint) warps
// walk only the entries whose reach mask has their bit (ballot + find-first-set); forward evaluates four
// entries together for ILP; backward reduces the 8 + D per-Gaussian partials with a transposing butterfly
// (16 shuffles) that leaves each total in its own lane, so one warp-wide red.global.add updates all of them.
// The per-pixel loop is FP32/MUFU bound, not HBM bound.
#include "common.cuh"

namespace {

// 1 / x to 1 ulp (MUFU.RCP): the reference kernels are built with fast-math and divide the same way
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr float ALPHA_MAX = 0.999f;
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float T_MIN = 1e-4f;
constexpr int MAX_BLOCK = 256;  // tile_size <= 16
// list entries per work unit; shorter for wide colour vectors so the staged segment stays under 48 KB of smem
__host__ __device__ constexpr int seg_len(int D) { return D <= 8 ? 512 : (D <= 16 ? 256 : 128); }

struct SegHeader {
    int total_segs;
    int pad[3];
};

// chain_T / chain_last / prefix_C hold the segment-LOCAL state after raster_seg_kernel and the chained state
// (what the backward reads) once raster_combine_kernel has folded the tile.
struct Workspace {
    SegHeader* hdr;
    int32_t* seg_start;  // [n_tiles + 1]
    int32_t* seg_tile;   // [max_segs]
    float* chain_T;      // [max_segs, 256]  transmittance after the segment; negative = pixel finished
    int32_t* chain_last; // [max_segs, 256]  last contributing list position so far (-1: none in a local state)
    float* prefix_C;     // [max_segs, 256, D] colour accumulated through the segment
    int32_t* done_k;     // [n_tiles, 8]  per warp footprint: the first segment after which all its pixels are finished
    int32_t* seg_order;  // [max_segs]  launch order of the segments: longest first (see seg_table_kernel)
    int32_t* tile_order; // [n_tiles]   launch order of the tiles' first segments: longest first
};

inline int64_t max_segments(int64_t n_isects, int64_t n_tiles, int D) { return n_isects / seg_len(D) + n_tiles; }

inline size_t ws_bytes(int64_t n_isects, int64_t n_tiles, int D) {
    int64_t ms = max_segments(n_isects, n_tiles, D);
    size_t b = 256;                                          // header
    b += fsb_align_up((size_t)(n_tiles + 1) * 4, 256);       // seg_start
    b += fsb_align_up((size_t)ms * 4, 256);                  // seg_tile
    b += fsb_align_up((size_t)ms * MAX_BLOCK * 4, 256) * 2;  // chain_T, chain_last
    b += fsb_align_up((size_t)ms * MAX_BLOCK * D * 4, 256);  // prefix_C
    b += fsb_align_up((size_t)n_tiles * 8 * 4, 256);         // done_k
    b += fsb_align_up((size_t)ms * 4, 256);                  // seg_order
    b += fsb_align_up((size_t)n_tiles * 4, 256);             // tile_order
    return b;
}

inline Workspace carve_ws(void* base, int64_t n_isects, int64_t n_tiles, int D) {
    int64_t ms = max_segments(n_isects, n_tiles, D);
    char* p = (char*)base;
    Workspace w;
    w.hdr = (SegHeader*)p; p += 256;
    w.seg_start = (int32_t*)p; p += fsb_align_up((size_t)(n_tiles + 1) * 4, 256);
    w.seg_tile = (int32_t*)p; p += fsb_align_up((size_t)ms * 4, 256);
    w.chain_T = (float*)p; p += fsb_align_up((size_t)ms * MAX_BLOCK * 4, 256);
    w.chain_last = (int32_t*)p; p += fsb_align_up((size_t)ms * MAX_BLOCK * 4, 256);
    w.prefix_C = (float*)p; p += fsb_align_up((size_t)ms * MAX_BLOCK * D * 4, 256);
    w.done_k = (int32_t*)p; p += fsb_align_up((size_t)n_tiles * 8 * 4, 256);
    w.seg_order = (int32_t*)p; p += fsb_align_up((size_t)ms * 4, 256);
    w.tile_order = (int32_t*)p;
    return w;
}

constexpr int LPT_BINS = 16;

// ---- segment table: one block scans ceil(len / SEG) over the tiles -------------------------------------------
__global__ void __launch_bounds__(1024)
seg_table_kernel(int n_tiles, int64_t n_isects, const int64_t* __restrict__ n_dev, int SEG,
                 const int32_t* __restrict__ tile_offsets, int32_t* __restrict__ seg_start,
                 int32_t* __restrict__ seg_tile, SegHeader* __restrict__ hdr, int32_t* __restrict__ seg_order,
                 int32_t* __restrict__ tile_order) {
    n_isects = fsb_eff_n(n_isects, n_dev);
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    __shared__ int s_bins[2][LPT_BINS + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + tid;
        int nseg = 0;
        if (t < n_tiles) {
            const int32_t b = tile_offsets[t];
            const int32_t e = (t == n_tiles - 1) ? (int32_t)n_isects : tile_offsets[t + 1];
            nseg = max(1, (e - b + SEG - 1) / SEG);
        }
        int inc = nseg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_warp[w];
        const int carry = s_carry;
        const int excl = carry + wbase + inc - nseg;
        if (t < n_tiles) {
            seg_start[t] = excl;
            for (int k = 0; k < nseg; ++k) seg_tile[excl + k] = t;
        }
        __syncthreads();
        if (tid == 1023) s_carry = carry + wbase + inc;
        __syncthreads();
    }
    if (tid == 0) {
        seg_start[n_tiles] = s_carry;
        hdr->total_segs = s_carry;
    }
    // Launch order, longest work unit first (LPT): CTA durations follow the segment length (1 .. SEG entries), and
    // with grid order = tile order a few long segments picked up late leave most SMs idle in the kernel's tail
    // (ncu: sm__cycles_active avg / max = 0.65 on the 640x480 bench scene).  Counting sort into LPT_BINS length
    // classes; the order inside a class is arbitrary (results do not depend on which CTA runs which unit).
    if (tid < 2 * (LPT_BINS + 1)) (&s_bins[0][0])[tid] = 0;
    __syncthreads();
    const int total = s_carry;
    auto unit_len = [&](int t, int k) {
        const int32_t b = tile_offsets[t];
        const int32_t e = (t == n_tiles - 1) ? (int32_t)n_isects : tile_offsets[t + 1];
        return min(SEG, max(0, e - b - k * SEG));
    };
    auto bin_of = [&](int n) { return (int)(((int64_t)(SEG - n) * LPT_BINS) / (SEG + 1)); };
    for (int pass = 0; pass < 2; ++pass) {  // 0: count, 1: scatter
        for (int base = 0; base < n_tiles; base += 1024) {
            const int t = base + tid;
            const int bin = (t < n_tiles) ? bin_of(unit_len(t, 0)) : LPT_BINS;
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            const int leader = __ffs(peers) - 1;
            int at = 0;
            if (lane == leader) at = atomicAdd(&s_bins[0][bin], __popc(peers));
            at = __shfl_sync(0xffffffffu, at, leader) + __popc(peers & ((1u << lane) - 1u));
            if (pass == 1 && t < n_tiles) tile_order[at] = t;
        }
        for (int base = 0; base < total; base += 1024) {
            const int sg = base + tid;
            int bin = LPT_BINS;
            if (sg < total) {
                const int t = seg_tile[sg];
                bin = bin_of(unit_len(t, sg - seg_start[t]));
            }
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            const int leader = __ffs(peers) - 1;
            int at = 0;
            if (lane == leader) at = atomicAdd(&s_bins[1][bin], __popc(peers));
            at = __shfl_sync(0xffffffffu, at, leader) + __popc(peers & ((1u << lane) - 1u));
            if (pass == 1 && sg < total) seg_order[at] = sg;
        }
        __syncthreads();
        if (pass == 0 && tid < 2) {  // counts -> exclusive start positions
            int run = 0;
            for (int b = 0; b < LPT_BINS; ++b) {
                const int c = s_bins[tid][b];
                s_bins[tid][b] = run;
                run += c;
            }
            s_bins[tid][LPT_BINS] = 0;
        }
        __syncthreads();
    }
}

struct TileGeom {
    int cam, tile_x, tile_y;
    int block_size, tr, lane, warp, n_warps, warps_x;
    int i, j;
    float px, py;
    bool inside;
};

// Pixel footprint of a warp: 8 x 4 pixels (lane = 8 * row + column), the warps of a 16 x 16 tile laid out 2 across
// and 4 down.  A near-square footprint is hit by fewer splats than a 16 x 2 strip of rows (a splat of diameter d
// touches ~(1 + d/8)(1 + d/4) footprints instead of (1 + d/16)(1 + d/2): 20 % fewer for d = 4..8 px), and every
// footprint a splat misses is a list entry that warp never evaluates.
constexpr int FOOT_W = 8, FOOT_H = 4;

__device__ __forceinline__ TileGeom tile_geom(int64_t tile_lin, int tile_w, int tile_h, int tile_size, int width,
                                              int height) {
    TileGeom g;
    const int n_tiles = tile_w * tile_h;
    g.cam = (int)(tile_lin / n_tiles);
    const int tile_id = (int)(tile_lin - (int64_t)g.cam * n_tiles);
    g.tile_y = tile_id / tile_w;
    g.tile_x = tile_id - g.tile_y * tile_w;
    g.block_size = blockDim.x * blockDim.y;
    g.tr = threadIdx.y * blockDim.x + threadIdx.x;
    g.lane = g.tr & 31;
    g.warp = g.tr >> 5;
    g.n_warps = g.block_size >> 5;
    g.warps_x = tile_size / FOOT_W;
    const int wy = g.warp / g.warps_x, wx = g.warp - wy * g.warps_x;
    g.i = g.tile_y * tile_size + wy * FOOT_H + (g.lane >> 3);
    g.j = g.tile_x * tile_size + wx * FOOT_W + (g.lane & 7);
    g.px = (float)g.j + 0.5f;
    g.py = (float)g.i + 0.5f;
    g.inside = (g.i < height && g.j < width);
    return g;
}

// Which warp footprints of this tile can the Gaussian reach with alpha >= 1/255 ?  alpha >= 1/255 means
// q(d) = a dx^2 + 2 b dx dy + c dy^2 <= 2 ln(255 opacity); the test is the exact minimum of q over the footprint's
// rectangle of pixel centres (0 if the mean lies inside, else the smallest of the four edge minima, each a clamped
// 1-D parabola), compared against an inflated threshold.  Projected surfels are thin rotated ellipses whose
// bounding box is mostly empty, so this removes far more (footprint, entry) visits than a box test.  Anything
// doubtful (non-PD conic, NaN) keeps all bits.
__device__ __forceinline__ uint32_t strip_mask(float gx, float gy, float opac, float a, float b, float c,
                                               float tile_px0, float tile_py0, int tile_size, int warps_x,
                                               int n_warps) {
    const uint32_t all = (1u << n_warps) - 1u;
    const float tau = __logf(255.f * opac);
    if (tau + 2e-3f < 0.f) return 0u;  // opacity below 1/255: can never pass the alpha test
    const float det = a * c - b * b;
    if (!(det > 0.f) || !(a > 0.f) || !(tau < 1e30f)) return all;
    const float thr = 2.f * tau * 1.0002f + 1e-2f;
    const float nb_c = -b / c, nb_a = -b / a;
    uint32_t m = 0u;
    for (int w = 0; w < n_warps; ++w) {
        const int wy = w / warps_x, wx = w - wy * warps_x;
        // rectangle of this footprint's pixel centres, relative to the mean
        const float x0 = tile_px0 + (float)(wx * FOOT_W) + 0.5f - gx, x1 = x0 + (float)(FOOT_W - 1);
        const float y0 = tile_py0 + (float)(wy * FOOT_H) + 0.5f - gy, y1 = y0 + (float)(FOOT_H - 1);
        float q = 0.f;
        if (!(x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f)) {
            float t, v;
            t = fminf(fmaxf(nb_c * x0, y0), y1); q = a * x0 * x0 + 2.f * b * x0 * t + c * t * t;
            t = fminf(fmaxf(nb_c * x1, y0), y1); v = a * x1 * x1 + 2.f * b * x1 * t + c * t * t; q = fminf(q, v);
            t = fminf(fmaxf(nb_a * y0, x0), x1); v = a * t * t + 2.f * b * t * y0 + c * y0 * y0; q = fminf(q, v);
            t = fminf(fmaxf(nb_a * y1, x0), x1); v = a * t * t + 2.f * b * t * y1 + c * y1 * y1; q = fminf(q, v);
        }
        if (!(q > thr)) m |= 1u << w;
    }
    return m;
}

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// A staged list segment.  The per-entry constants are stored in the form the inner loops consume:
//   alpha = min(0.999, 2^(p + L)),  p = a' dx^2 + b' dx dy + c' dy^2  (= -log2e * sigma),  L = log2(opacity)
// so one evaluation is 2 FADD + 6 FMUL/FFMA + 1 FADD + MUFU.EX2 + FMNMX, and "sigma < 0" is "p > 0".
// Slot SEG is a dummy entry with alpha = 0 that pads the per-warp work lists to a multiple of four.
template <int D>
struct Stage {
    static constexpr int SEG = seg_len(D);
    static constexpr int DP = (D == 3) ? 4 : D;  // colour row stride (one 128-bit load for D = 3 and D = 4)
    int32_t id[SEG];
    float4 geo[SEG + 1];  // x, y, log2(opacity), strip mask (as int bits)
    float4 con[SEG + 1];  // -0.5 log2e a, -log2e b, -0.5 log2e c, 1 / opacity
    float col[(SEG + 1) * DP];
    alignas(8) uint16_t wlist[MAX_BLOCK / 32][SEG + 4];  // per warp: the entries whose strip mask has its bit
};

template <int D>
__device__ __forceinline__ void stage_entry(Stage<D>& s, int slot, int32_t g, const float2* __restrict__ means2d,
                                            const float* __restrict__ conics, const float* __restrict__ colors,
                                            const float* __restrict__ opacities, const TileGeom& tg, int tile_size) {
    constexpr int DP = Stage<D>::DP;
    s.id[slot] = g;
    const float2 xy = means2d[g];
    const float o = opacities[g];
    const float a = conics[3 * (size_t)g], b = conics[3 * (size_t)g + 1], c = conics[3 * (size_t)g + 2];
    const uint32_t m = strip_mask(xy.x, xy.y, o, a, b, c, (float)(tg.tile_x * tile_size),
                                  (float)(tg.tile_y * tile_size), tile_size, tg.warps_x, tg.n_warps);
    s.geo[slot] = make_float4(xy.x, xy.y, __log2f(o), __int_as_float((int)m));
    s.con[slot] = make_float4(-0.5f * LOG2E * a, -LOG2E * b, -0.5f * LOG2E * c, fast_rcp(o));
    const float* cp = colors + (size_t)g * D;
#pragma unroll
    for (int k = 0; k < D; ++k) s.col[slot * DP + k] = cp[k];
}

// Stage entries [0, n) of the segment that starts at list position seg_b (all threads), then build the per-warp
// work lists.  Returns the number of entries in this warp's list (it is padded to a multiple of 4 with SEG).
template <int D>
__device__ __forceinline__ int stage_segment(Stage<D>& s, int n, int32_t seg_b, const int32_t* __restrict__ flatten_ids,
                                             const float2* __restrict__ means2d, const float* __restrict__ conics,
                                             const float* __restrict__ colors, const float* __restrict__ opacities,
                                             const TileGeom& tg, int tile_size) {
    constexpr int SEG = Stage<D>::SEG;
    for (int e = tg.tr; e < n; e += tg.block_size)
        stage_entry<D>(s, e, flatten_ids[seg_b + e], means2d, conics, colors, opacities, tg, tile_size);
    if (tg.tr == 0) {
        s.geo[SEG] = make_float4(0.f, 0.f, -1000.f, 0.f);  // 2^-1000 = 0: fails the alpha test
        s.con[SEG] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    uint16_t* wl = s.wlist[tg.warp];
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int tt = k0 + tg.lane;
        const uint32_t m = (tt < n) ? (uint32_t)__float_as_int(s.geo[tt].w) : 0u;
        const bool bit = (m >> tg.warp) & 1u;
        const uint32_t bits = __ballot_sync(0xffffffffu, bit);
        if (bit) wl[base + __popc(bits & ((1u << tg.lane) - 1u))] = (uint16_t)tt;
        base += __popc(bits);
    }
    if (tg.lane < 4) wl[base + tg.lane] = (uint16_t)SEG;
    __syncwarp();
    return base;
}

// alpha of staged entry t at this thread's pixel; `p` receives -log2e * sigma, `au` the unclamped opacity * vis
__device__ __forceinline__ float eval_alpha(const float4& geo, const float4& con, float px, float py, float& dx,
                                            float& dy, float& p, float& au) {
    dx = geo.x - px;
    dy = geo.y - py;
    p = fmaf(con.y * dx, dy, fmaf(con.z * dy, dy, con.x * dx * dx));
    au = fast_ex2(p + geo.z);
    return fminf(ALPHA_MAX, au);
}

// Front-to-back walk of this warp's work list with the reference's per-pixel rules:
// skip when sigma < 0 or alpha < 1/255, stop (entry NOT blended, pixel `done`) when T (1 - alpha) <= 1e-4.
// `done` lanes are frozen.  `last` = list position of the last blended entry.
template <int D>
__device__ __forceinline__ void walk(const Stage<D>& s, int cnt, int seg_b, const TileGeom& tg, float& T,
                                     float (&acc)[D], int32_t& last, bool& done) {
    constexpr int DP = Stage<D>::DP;
    if (__all_sync(0xffffffffu, done)) return;
    const uint16_t* wl = s.wlist[tg.warp];
    for (int i = 0; i < cnt; i += 4) {
        const uint2 pk = *reinterpret_cast<const uint2*>(wl + i);
        const int t[4] = {(int)(pk.x & 0xffffu), (int)(pk.x >> 16), (int)(pk.y & 0xffffu), (int)(pk.y >> 16)};
        float alpha[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float dx, dy, p, au;
            alpha[u] = eval_alpha(s.geo[t[u]], s.con[t[u]], tg.px, tg.py, dx, dy, p, au);
            ok[u] = (p <= 0.f) && (alpha[u] >= ALPHA_MIN);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (ok[u] && !done) {
                const float next_T = fmaf(-T, alpha[u], T);
                if (next_T <= T_MIN) {
                    done = true;
                } else {
                    const float w = alpha[u] * T;
                    const float* cp = s.col + t[u] * DP;
                    if constexpr (D == 3 || D == 4) {
                        const float4 c4 = *reinterpret_cast<const float4*>(cp);
                        acc[0] = fmaf(c4.x, w, acc[0]);
                        acc[1] = fmaf(c4.y, w, acc[1]);
                        acc[2] = fmaf(c4.z, w, acc[2]);
                        if constexpr (D == 4) acc[3] = fmaf(c4.w, w, acc[3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < D; ++k) acc[k] = fmaf(cp[k], w, acc[k]);
                    }
                    last = seg_b + t[u];
                    T = next_T;
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
    }
}

struct FwdOut {
    const float* backgrounds;
    int ed_normalize;
    float* out_colors;
    float* out_alphas;
    int32_t* last_ids;
};

template <int D>
__device__ __forceinline__ void write_pixel(const FwdOut& o, const TileGeom& tg, int64_t pix, bool masked, float T,
                                            const float (&acc)[D], int32_t last) {
    const float alpha_out = masked ? 0.f : 1.f - T;
    o.out_alphas[pix] = alpha_out;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        float v = o.backgrounds ? acc[c] + (1.f - alpha_out) * o.backgrounds[tg.cam * D + c] : acc[c];
        if (o.ed_normalize && c == D - 1) v = v / fmaxf(alpha_out, 1e-10f);
        o.out_colors[pix * D + c] = v;
    }
    o.last_ids[pix] = last;
}

struct SegGeom {
    int64_t tile_lin;
    int k, nseg, n;
    int32_t seg_b;
    bool masked;
};

template <int D>
__device__ __forceinline__ SegGeom seg_geom(const Workspace& ws, int seg, int64_t tile_lin, int64_t n_cam_tiles,
                                            int64_t n_isects, const int32_t* __restrict__ tile_offsets,
                                            const uint8_t* __restrict__ masks) {
    constexpr int SEG = seg_len(D);
    SegGeom g;
    g.tile_lin = tile_lin;
    const int s0 = ws.seg_start[tile_lin];
    g.k = seg - s0;
    g.nseg = ws.seg_start[tile_lin + 1] - s0;
    g.masked = (masks != nullptr && !masks[tile_lin]);
    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end = (tile_lin == n_cam_tiles - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    g.seg_b = range_start + g.k * SEG;
    g.n = g.masked ? 0 : max(0, min(range_end, g.seg_b + SEG) - g.seg_b);
    return g;
}

// forward A: every (tile, segment) composites its own entries from T = 1.
// PHASE 0 runs the first segment of every tile (grid = tiles), PHASE 1 all later segments (grid = segments, the
// first ones exit).  A warp strip whose 32 pixels are all finished after segment k (exactly for k = 0, by the
// local-saturation argument for k > 0) records k in done_k; a later segment skips that strip: nothing it could
// composite is ever used.  The kernel boundary makes every first-segment result visible to phase 1, which is
// where most of the saving is: opaque tiles stop inside their first segment or two.
template <int D, int PHASE>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_seg_kernel(int C, int N, int64_t n_isects, const int64_t* __restrict__ n_dev, const float2* __restrict__ means2d,
                  const float* __restrict__ conics, const float* __restrict__ colors,
                  const float* __restrict__ opacities, const uint8_t* __restrict__ masks, int width, int height,
                  int tile_size, int tile_w, int tile_h, const int32_t* __restrict__ tile_offsets,
                  const int32_t* __restrict__ flatten_ids, Workspace ws, FwdOut o) {
    __shared__ Stage<D> s;
    n_isects = fsb_eff_n(n_isects, n_dev);
    int seg;
    int64_t tile_lin;
    if (PHASE == 0) {
        tile_lin = ws.tile_order[blockIdx.x];
        seg = ws.seg_start[tile_lin];
    } else {
        if ((int)blockIdx.x >= ws.hdr->total_segs) return;
        seg = ws.seg_order[blockIdx.x];
        tile_lin = ws.seg_tile[seg];
        if (seg == ws.seg_start[tile_lin]) return;
    }
    const SegGeom sg = seg_geom<D>(ws, seg, tile_lin, (int64_t)C * tile_w * tile_h, n_isects, tile_offsets, masks);
    const TileGeom tg = tile_geom(sg.tile_lin, tile_w, tile_h, tile_size, width, height);
    int32_t* done_k = ws.done_k + tile_lin * 8 + tg.warp;
    bool skip = false;
    if (PHASE == 1) {
        skip = (*done_k < sg.k);  // written by phase 0 (visible) or, opportunistically, by an earlier segment
        // a skipped strip leaves no state behind, but its slots must not keep a stale STOP_MARK from an earlier
        // call that used the same workspace memory (raster_stop_kernel scans chain_last of every later segment)
        if (skip) ws.chain_last[(size_t)seg * MAX_BLOCK + tg.tr] = -1;
        if (__syncthreads_and(skip)) return;
    }

    const int cnt = stage_segment<D>(s, sg.n, sg.seg_b, flatten_ids, means2d, conics, colors, opacities, tg, tile_size);
    if (skip) return;

    float T = 1.f;
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
    int32_t last = (sg.k == 0) ? 0 : -1;
    bool done = !tg.inside;
    walk<D>(s, cnt, sg.seg_b, tg, T, acc, last, done);
    if (sg.nseg > 1 && __all_sync(0xffffffffu, done) && tg.lane == 0) atomicMin(done_k, sg.k);

    const size_t cidx = (size_t)seg * MAX_BLOCK + tg.tr;
    ws.chain_T[cidx] = (done && tg.inside) ? -T : T;
    ws.chain_last[cidx] = last;
#pragma unroll
    for (int c = 0; c < D; ++c) ws.prefix_C[cidx * D + c] = acc[c];
    if (sg.nseg == 1 && tg.inside) {
        const int64_t pix = ((int64_t)tg.cam * height + tg.i) * width + tg.j;
        write_pixel<D>(o, tg, pix, sg.masked, T, acc, last);
    }
}

constexpr int32_t STOP_MARK = INT32_MIN;  // chain_last value: "this pixel's stop rule fires inside this segment"

// forward B: fold the segment states of every multi-segment tile in list order (pure arithmetic, no list walk).
// Overwrites the local states with the chained ones; a pixel whose stop rule fires inside segment k gets
// chain_last[k] = STOP_MARK and is finished by raster_stop_kernel.
template <int D>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_fold_kernel(int C, int width, int height, int tile_size, int tile_w, int tile_h,
                   const uint8_t* __restrict__ masks, Workspace ws, FwdOut o) {
    const int64_t tile_lin = blockIdx.x;
    const int seg0 = ws.seg_start[tile_lin];
    const int nseg = ws.seg_start[tile_lin + 1] - seg0;
    if (nseg <= 1) return;
    const TileGeom tg = tile_geom(tile_lin, tile_w, tile_h, tile_size, width, height);
    size_t cidx = (size_t)seg0 * MAX_BLOCK + tg.tr;
    const float t0 = ws.chain_T[cidx];
    float T = fabsf(t0);
    bool done = (t0 < 0.f) || !tg.inside;
    bool pending = false;  // stop segment found, result still to be produced by raster_stop_kernel
    int32_t last = ws.chain_last[cidx];
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = ws.prefix_C[cidx * D + c];
    for (int k = 1; k < nseg; ++k) {
        if (__syncthreads_count(done) == tg.block_size) break;
        if (done) continue;
        cidx = (size_t)(seg0 + k) * MAX_BLOCK + tg.tr;
        const float tl = ws.chain_T[cidx];
        const float Tl = fabsf(tl);
        if (!(tl < 0.f) && T * Tl > T_MIN) {  // the stop rule cannot have fired inside this segment
#pragma unroll
            for (int c = 0; c < D; ++c) {
                acc[c] += T * ws.prefix_C[cidx * D + c];
                ws.prefix_C[cidx * D + c] = acc[c];
            }
            T *= Tl;
            ws.chain_T[cidx] = T;
            const int32_t ll = ws.chain_last[cidx];
            if (ll >= 0) last = ll;
            ws.chain_last[cidx] = last;
        } else {
            ws.chain_last[cidx] = STOP_MARK;
            done = true;
            pending = true;
        }
    }
    if (pending) return;
    // the backward takes the tile total from the last segment's slot
    const size_t lidx = (size_t)(seg0 + nseg - 1) * MAX_BLOCK + tg.tr;
#pragma unroll
    for (int c = 0; c < D; ++c) ws.prefix_C[lidx * D + c] = acc[c];
    if (tg.inside) {
        const bool masked = (masks != nullptr && !masks[tile_lin]);
        const int64_t pix = ((int64_t)tg.cam * height + tg.i) * width + tg.j;
        write_pixel<D>(o, tg, pix, masked, T, acc, last);
    }
}

// forward C: the pixels whose stop rule fires inside this segment re-walk it exactly from their incoming
// (chained) state, which keeps the sequential semantics: stop position, last id, nothing blended after it.
template <int D>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_stop_kernel(int C, int N, int64_t n_isects, const int64_t* __restrict__ n_dev, const float2* __restrict__ means2d,
                   const float* __restrict__ conics, const float* __restrict__ colors,
                   const float* __restrict__ opacities, const uint8_t* __restrict__ masks, int width, int height,
                   int tile_size, int tile_w, int tile_h, const int32_t* __restrict__ tile_offsets,
                   const int32_t* __restrict__ flatten_ids, Workspace ws, FwdOut o) {
    __shared__ Stage<D> s;
    n_isects = fsb_eff_n(n_isects, n_dev);
    if ((int)blockIdx.x >= ws.hdr->total_segs) return;
    const int seg = ws.seg_order[blockIdx.x];
    const int64_t tile_lin = ws.seg_tile[seg];
    if (seg == ws.seg_start[tile_lin]) return;  // first segments are exact already
    const int tr = threadIdx.y * blockDim.x + threadIdx.x;
    const size_t cidx = (size_t)seg * MAX_BLOCK + tr;
    const bool redo = (ws.chain_last[cidx] == STOP_MARK);
    if (!__syncthreads_or(redo)) return;
    const SegGeom sg = seg_geom<D>(ws, seg, tile_lin, (int64_t)C * tile_w * tile_h, n_isects, tile_offsets, masks);
    const TileGeom tg = tile_geom(tile_lin, tile_w, tile_h, tile_size, width, height);
    const int cnt = stage_segment<D>(s, sg.n, sg.seg_b, flatten_ids, means2d, conics, colors, opacities, tg, tile_size);
    float T = 1.f;
    float acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
    int32_t last = 0;
    const size_t pidx = cidx - MAX_BLOCK;  // chained state after the previous segment
    if (redo) {
        T = fabsf(ws.chain_T[pidx]);
        last = ws.chain_last[pidx];
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = ws.prefix_C[pidx * D + c];
    }
    bool frozen = !redo;
    walk<D>(s, cnt, sg.seg_b, tg, T, acc, last, frozen);
    if (!redo) return;
    ws.chain_T[cidx] = -T;
    ws.chain_last[cidx] = last;
    const size_t lidx = (size_t)(ws.seg_start[tile_lin + 1] - 1) * MAX_BLOCK + tr;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        ws.prefix_C[cidx * D + c] = acc[c];
        ws.prefix_C[lidx * D + c] = acc[c];
    }
    const int64_t pix = ((int64_t)tg.cam * height + tg.i) * width + tg.j;
    write_pixel<D>(o, tg, pix, sg.masked, T, acc, last);
}

// Reduce-scatter of N per-lane values over the warp: a butterfly that halves the live values at every stage
// (lanes with the stage bit set keep the upper half, the others the lower half), so lane L ends up with the
// warp total of value index slot_of_lane<N>(L).  ceil(N/2) + ceil(N/4) + ... shuffles in all (13 for N = 12,
// 9 for N = 7) against 5 N for plain butterflies.  v[] is clobbered; the lane's total is returned.
template <int N, int XOR>
struct Bfly {
    template <int NV>
    static __device__ __forceinline__ float run(float (&v)[NV], int lane) {
        constexpr int H = (N + 1) / 2;
        const bool hi = lane & XOR;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float lo_v = v[i];
            const float hi_v = (H + i < N) ? v[H + i] : 0.f;
            const float send = hi ? lo_v : hi_v;
            const float keep = hi ? hi_v : lo_v;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, XOR);
        }
        return Bfly<H, XOR / 2>::run(v, lane);
    }
};
template <int N>
struct Bfly<N, 0> {
    template <int NV>
    static __device__ __forceinline__ float run(float (&v)[NV], int) { return v[0]; }
};
template <int N>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[N], int lane) {
    return Bfly<N, 16>::run(v, lane);
}
// value index held by lane L after warp_transpose_sum<N>, or -1 if the lane holds none
template <int N>
__device__ __forceinline__ int slot_of_lane(int lane) {
    int base = 0, n = N, nt = N;  // n: values this lane's group really holds; nt: the stage width (uniform)
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) {
        const int h = (nt + 1) / 2;
        if (lane & x) { base += h; n -= h; } else { n = min(n, h); }
        nt = h;
    }
    return n >= 1 ? base : -1;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Slots of the packed gradient vector: [0, D) colours, then conic a b c, opacity, xy, |xy|.
// XYMODE: 0 = no gradient for the 2-D means (the normals pass detaches them), 1 = xy, 2 = xy and |xy| (absgrad).
template <int D, int XYMODE>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_bwd_kernel(int C, int N, int64_t n_isects, const int64_t* __restrict__ n_dev, const float2* __restrict__ means2d,
                  const float* __restrict__ conics, const float* __restrict__ colors,
                  const float* __restrict__ opacities, const float* __restrict__ backgrounds,
                  const uint8_t* __restrict__ masks, int width, int height, int tile_size, int tile_w, int tile_h,
                  const int32_t* __restrict__ tile_offsets, const int32_t* __restrict__ flatten_ids,
                  int ed_normalize, Workspace ws, const float* __restrict__ render_colors,
                  const float* __restrict__ render_alphas, const int32_t* __restrict__ last_ids,
                  const float* __restrict__ v_render_colors, const float* __restrict__ v_render_alphas,
                  float* __restrict__ v_means2d_abs, float* __restrict__ v_means2d, float* __restrict__ v_conics,
                  float* __restrict__ v_colors, float* __restrict__ v_opacities) {
    constexpr int SEG = seg_len(D);
    __shared__ Stage<D> s;
    __shared__ int32_t s_wmax[MAX_BLOCK / 32];
    n_isects = fsb_eff_n(n_isects, n_dev);
    if ((int)blockIdx.x >= ws.hdr->total_segs) return;
    const int seg = ws.seg_order[blockIdx.x];
    const int64_t tile_lin = ws.seg_tile[seg];
    if (masks != nullptr && !masks[tile_lin]) return;
    const int k = seg - ws.seg_start[tile_lin];
    const int seg_last = ws.seg_start[tile_lin + 1] - 1;
    const TileGeom tg = tile_geom(tile_lin, tile_w, tile_h, tile_size, width, height);
    const int64_t pix = tg.inside ? ((int64_t)tg.cam * height + tg.i) * width + tg.j : 0;

    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end =
        (tile_lin == (int64_t)C * tile_w * tile_h - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    const int32_t seg_b = range_start + k * SEG;
    const int n = max(0, min(range_end, seg_b + SEG) - seg_b);
    if (n <= 0) return;

    int32_t bin_final = -1;
    if (tg.inside) bin_final = last_ids[pix];
    int32_t warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));
    if (tg.lane == 0) s_wmax[tg.warp] = warp_bin_final;
    __syncthreads();
    int32_t cta_bin_final = -1;
    for (int w = 0; w < tg.n_warps; ++w) cta_bin_final = max(cta_bin_final, s_wmax[w]);
    if (cta_bin_final < seg_b) return;  // no pixel of the tile reaches into this segment

    float T_final = 1.f, v_ra = 0.f, T = 1.f;
    float v_rc[D];
    float buffer[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { v_rc[c] = 0.f; buffer[c] = 0.f; }
    if (tg.inside) {
        const float alpha_out = render_alphas[pix];
        T_final = 1.f - alpha_out;
        v_ra = v_render_alphas[pix];
#pragma unroll
        for (int c = 0; c < D; ++c) v_rc[c] = v_render_colors[pix * D + c];
        if (ed_normalize) {
            const float den = fmaxf(alpha_out, 1e-10f);
            const float v_ed = v_rc[D - 1];
            v_rc[D - 1] = v_ed / den;
            if (alpha_out >= 1e-10f) v_ra += -v_ed * render_colors[pix * D + D - 1] / den;
        }
        // state at the END of this segment, left behind by the forward pass
        const size_t cidx = (size_t)seg * MAX_BLOCK + tg.tr;
        const size_t lidx = (size_t)seg_last * MAX_BLOCK + tg.tr;
        T = fabsf(ws.chain_T[cidx]);
#pragma unroll
        for (int c = 0; c < D; ++c) buffer[c] = ws.prefix_C[lidx * D + c] - ws.prefix_C[cidx * D + c];
    }
    float bg_dot = 0.f;
    if (backgrounds) {
#pragma unroll
        for (int c = 0; c < D; ++c) bg_dot += backgrounds[tg.cam * D + c] * v_rc[c];
    }
    constexpr bool want_xy = (XYMODE >= 1);
    constexpr bool want_abs = (XYMODE >= 2);
    constexpr bool kTranspose = (D <= 8);
    constexpr int NG = 4 + 2 * XYMODE;             // conic a b c, opacity (, xy (, |xy|))
    constexpr int NV = kTranspose ? D + NG : NG;   // values reduced by the transposing butterfly
    constexpr int B = kTranspose ? D : 0;          // first geometric slot
    const int slot = slot_of_lane<NV>(tg.lane);
    float* slot_base = nullptr;
    int slot_stride = 0;
    if (slot >= 0) {
        if (slot < B) { slot_base = v_colors + slot; slot_stride = D; }
        else if (slot < B + 3) { slot_base = v_conics + (slot - B); slot_stride = 3; }
        else if (slot == B + 3) { slot_base = v_opacities; slot_stride = 1; }
        else if (slot < B + 6) { slot_base = v_means2d + (slot - B - 4); slot_stride = 2; }
        else { slot_base = v_means2d_abs + (slot - B - 6); slot_stride = 2; }
    }

    // entries behind every pixel's last id are not even staged
    const int n_used = min(n, cta_bin_final - seg_b + 1);
    const int cnt = stage_segment<D>(s, n_used, seg_b, flatten_ids, means2d, conics, colors, opacities, tg, tile_size);
    constexpr int DP = Stage<D>::DP;
    const uint16_t* wl = s.wlist[tg.warp];
    const bool opac_slot = (slot == B + 3);

    const int t_hi = warp_bin_final - seg_b;  // last entry this warp can need
    for (int i = cnt - 1; i >= 0; --i) {
        const int t = wl[i];
        if (t > t_hi) continue;  // warp-uniform
        const float4 con = s.con[t];
        float dx, dy, p, au;
        const float alpha = eval_alpha(s.geo[t], con, tg.px, tg.py, dx, dy, p, au);
        const bool valid = tg.inside && (seg_b + t <= bin_final) && (p <= 0.f) && (alpha >= ALPHA_MIN);
        if (!__any_sync(0xffffffffu, valid)) continue;

        float v[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) v[c] = 0.f;
        float v_colD[D > 8 ? D : 1];
        if (valid) {
            const float ra = fast_rcp(1.f - alpha);
            T *= ra;
            const float fac = alpha * T;
            float v_alpha = 0.f;
            const float* cp = s.col + t * DP;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float ck = cp[c];
                if constexpr (kTranspose) v[c] = fac * v_rc[c];
                else v_colD[c] = fac * v_rc[c];
                v_alpha += (ck * T - buffer[c] * ra) * v_rc[c];
                buffer[c] += ck * fac;
            }
            v_alpha += T_final * ra * v_ra;
            if (backgrounds) v_alpha += -T_final * ra * bg_dot;
            if (au <= ALPHA_MAX) {
                // au = opacity * vis.  v_sigma = -au v_alpha; the opacity gradient vis v_alpha = (au / opacity) v_alpha
                // is reduced as au v_alpha and divided by the opacity once, by the lane that owns the total.
                const float nvs = au * v_alpha;
                const float hs = -0.5f * nvs;
                v[B + 0] = hs * dx * dx;
                v[B + 1] = -nvs * dx * dy;
                v[B + 2] = hs * dy * dy;
                v[B + 3] = nvs;
                if constexpr (want_xy) {
                    // conic a = -2 a' / log2e etc.:  v_sigma (a dx + b dy) = (nvs / log2e) (2 a' dx + b' dy)
                    const float vk = nvs * (1.f / LOG2E);
                    const float gx = vk * fmaf(2.f * con.x, dx, con.y * dy);
                    const float gy = vk * fmaf(2.f * con.z, dy, con.y * dx);
                    v[B + 4] = gx;
                    v[B + 5] = gy;
                    if constexpr (want_abs) {
                        v[B + 6] = fabsf(gx);
                        v[B + 7] = fabsf(gy);
                    }
                }
            }
        } else if constexpr (!kTranspose) {
#pragma unroll
            for (int c = 0; c < D; ++c) v_colD[c] = 0.f;
        }
        const int32_t g = s.id[t];
        if constexpr (!kTranspose) {
            // wide colour vectors: colours by plain butterflies, the geometric values transposed
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float tot = warp_sum(v_colD[c]);
                if (tg.lane == 0) atomicAdd(v_colors + (size_t)g * D + c, tot);
            }
        }
        float total = warp_transpose_sum(v, tg.lane);
        if (opac_slot) total *= con.w;
        if (slot_base != nullptr && total != 0.f) atomicAdd(slot_base + (size_t)g * slot_stride, total);
    }
}

template <int D>
int launch_fwd(int C, int N, int64_t n_isects, const int64_t* n_dev, const float* means2d, const float* conics, const float* colors,
               const float* opacities, const float* backgrounds, const uint8_t* masks, int width, int height,
               int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids,
               int ed_normalize, void* workspace, float* out_colors, float* out_alphas, int32_t* last_ids,
               cudaStream_t st) {
    const int64_t n_tiles = (int64_t)C * tile_w * tile_h;
    Workspace ws = carve_ws(workspace, n_isects, n_tiles, D);
    seg_table_kernel<<<1, 1024, 0, st>>>((int)n_tiles, n_isects, n_dev, seg_len(D), tile_offsets, ws.seg_start, ws.seg_tile, ws.hdr,
                                         ws.seg_order, ws.tile_order);
    FSB_LAUNCH_CHECK();
    dim3 block(tile_size, tile_size);
    FwdOut o{backgrounds, ed_normalize, out_colors, out_alphas, last_ids};
    unsigned grid = (unsigned)max_segments(n_isects, n_tiles, D);
    const bool multi = n_dev != nullptr || n_isects > seg_len(D);  // otherwise no tile can hold more than one segment
    if (multi) FSB_CUDA(cudaMemsetAsync(ws.done_k, 0x7f, (size_t)n_tiles * 8 * 4, st));
    raster_seg_kernel<D, 0><<<(unsigned)n_tiles, block, 0, st>>>(C, N, n_isects, n_dev, (const float2*)means2d, conics, colors,
                                                                 opacities, masks, width, height, tile_size, tile_w,
                                                                 tile_h, tile_offsets, flatten_ids, ws, o);
    FSB_LAUNCH_CHECK();
    if (multi) {
        raster_seg_kernel<D, 1><<<grid, block, 0, st>>>(C, N, n_isects, n_dev, (const float2*)means2d, conics, colors,
                                                        opacities, masks, width, height, tile_size, tile_w, tile_h,
                                                        tile_offsets, flatten_ids, ws, o);
        FSB_LAUNCH_CHECK();
        raster_fold_kernel<D><<<(unsigned)n_tiles, block, 0, st>>>(C, width, height, tile_size, tile_w, tile_h, masks,
                                                                   ws, o);
        FSB_LAUNCH_CHECK();
        raster_stop_kernel<D><<<grid, block, 0, st>>>(C, N, n_isects, n_dev, (const float2*)means2d, conics, colors,
                                                      opacities, masks, width, height, tile_size, tile_w, tile_h,
                                                      tile_offsets, flatten_ids, ws, o);
        FSB_LAUNCH_CHECK();
    }
    return 0;
}

template <int D>
int launch_bwd(int C, int N, int64_t n_isects, const int64_t* n_dev, const float* means2d, const float* conics, const float* colors,
               const float* opacities, const float* backgrounds, const uint8_t* masks, int width, int height,
               int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids,
               int ed_normalize, void* workspace, const float* render_colors, const float* render_alphas,
               const int32_t* last_ids, const float* v_render_colors, const float* v_render_alphas,
               float* v_means2d_abs, float* v_means2d, float* v_conics, float* v_colors, float* v_opacities,
               cudaStream_t st) {
    const int64_t n_tiles = (int64_t)C * tile_w * tile_h;
    Workspace ws = carve_ws(workspace, n_isects, n_tiles, D);
    dim3 block(tile_size, tile_size);
    unsigned grid = (unsigned)max_segments(n_isects, n_tiles, D);
#define FSB_BWD_LAUNCH(MODE)                                                                                       \
    raster_bwd_kernel<D, MODE><<<grid, block, 0, st>>>(                                                             \
        C, N, n_isects, n_dev, (const float2*)means2d, conics, colors, opacities, backgrounds, masks, width, height, \
        tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize, ws, render_colors, render_alphas,       \
        last_ids, v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities)
    if (v_means2d_abs) FSB_BWD_LAUNCH(2);
    else if (v_means2d) FSB_BWD_LAUNCH(1);
    else FSB_BWD_LAUNCH(0);
#undef FSB_BWD_LAUNCH
    FSB_LAUNCH_CHECK();
    return 0;
}

// Measurement aid (bench.py roofline, not on the training path): per tile, the number of (pixel, entry) pairs the
// forward BLENDED (entry at or before the pixel's last id that passes the sigma / alpha tests: exactly the pairs the
// backward differentiates) and the number a list walk has to visit (every entry up to each pixel's last id).
// counts[0] += blended, counts[1] += visited.  One CTA per tile, entries read straight from global memory.
__global__ void __launch_bounds__(MAX_BLOCK)
raster_pair_count_kernel(int C, int64_t n_isects, const int64_t* __restrict__ n_dev, const float2* __restrict__ means2d,
                         const float* __restrict__ conics, const float* __restrict__ opacities, int width, int height,
                         int tile_size, int tile_w, int tile_h, const int32_t* __restrict__ tile_offsets,
                         const int32_t* __restrict__ flatten_ids, const int32_t* __restrict__ last_ids,
                         unsigned long long* __restrict__ counts) {
    __shared__ int32_t s_wmax[MAX_BLOCK / 32];
    __shared__ unsigned long long s_sum[2];
    n_isects = fsb_eff_n(n_isects, n_dev);
    const int64_t tile_lin = blockIdx.x;
    const TileGeom tg = tile_geom(tile_lin, tile_w, tile_h, tile_size, width, height);
    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end =
        (tile_lin == (int64_t)C * tile_w * tile_h - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    int32_t last = -1;
    if (tg.inside && range_end > range_start) last = last_ids[((int64_t)tg.cam * height + tg.i) * width + tg.j];
    int32_t wmax = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (tg.lane == 0) s_wmax[tg.warp] = wmax;
    if (tg.tr < 2) s_sum[tg.tr] = 0ull;
    __syncthreads();
    int32_t cmax = -1;
    for (int w = 0; w < tg.n_warps; ++w) cmax = max(cmax, s_wmax[w]);
    unsigned blended = 0u, visited = 0u;
    for (int32_t e = range_start; e <= min(cmax, range_end - 1); ++e) {
        const int32_t g = flatten_ids[e];
        const float2 xy = means2d[g];
        const float o = opacities[g];
        const float a = conics[3 * (size_t)g], b = conics[3 * (size_t)g + 1], c = conics[3 * (size_t)g + 2];
        const float4 geo = make_float4(xy.x, xy.y, __log2f(o), 0.f);
        const float4 con = make_float4(-0.5f * LOG2E * a, -LOG2E * b, -0.5f * LOG2E * c, 0.f);
        float dx, dy, p, au;
        const float alpha = eval_alpha(geo, con, tg.px, tg.py, dx, dy, p, au);
        if (e <= last) {
            ++visited;
            if (p <= 0.f && alpha >= ALPHA_MIN) ++blended;
        }
    }
    atomicAdd(&s_sum[0], (unsigned long long)blended);
    atomicAdd(&s_sum[1], (unsigned long long)visited);
    __syncthreads();
    if (tg.tr < 2 && s_sum[tg.tr]) atomicAdd(counts + tg.tr, s_sum[tg.tr]);
}

bool bad_geometry(int C, int tile_size, int64_t n_isects) {
    // whole 8 x 4 warp footprints: tile_size 8 or 16 (the reference uses 16, dn_model.py:547)
    return C <= 0 || tile_size < 8 || tile_size > 16 || tile_size % 8 != 0 ||
           n_isects < 0 || n_isects > 0x7fffffffLL;
}

}  // namespace

#define FSB_DISPATCH_D(D_, CALL)            \
    switch (D_) {                           \
        case 1: { constexpr int DD = 1; return CALL; }   \
        case 2: { constexpr int DD = 2; return CALL; }   \
        case 3: { constexpr int DD = 3; return CALL; }   \
        case 4: { constexpr int DD = 4; return CALL; }   \
        case 5: { constexpr int DD = 5; return CALL; }   \
        case 8: { constexpr int DD = 8; return CALL; }   \
        case 16: { constexpr int DD = 16; return CALL; } \
        case 32: { constexpr int DD = 32; return CALL; } \
        default: return FSB_E_ARG;          \
    }

// channel counts the kernels are instantiated for; callers pad up to the next one
FSB_API int fsb_raster_supported_channels(int D) {
    const int s[] = {1, 2, 3, 4, 5, 8, 16, 32};
    for (int i = 0; i < 8; ++i)
        if (s[i] >= D) return s[i];
    return -1;
}

// bytes of the per-call workspace: segment table + the per-(segment, pixel) chain state that the forward
// leaves for the backward (n_tiles = C * tile_w * tile_h)
FSB_API size_t fsb_raster_workspace(int64_t n_isects, int64_t n_tiles, int D) {
    if (n_isects < 0 || n_tiles <= 0 || D <= 0) return 0;
    return ws_bytes(n_isects, n_tiles, D);
}

// n_isects_dev (nullable): static-capacity mode, n_isects is then the capacity of flatten_ids (common.cuh).
FSB_API int fsb_raster_fwd(int C, int N, int D, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* backgrounds,
                           const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                           const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_normalize,
                           void* workspace, size_t workspace_bytes, float* out_colors, float* out_alphas,
                           int32_t* last_ids, void* stream) {
    if (bad_geometry(C, tile_size, n_isects)) return FSB_E_ARG;
    if (tile_w <= 0 || tile_h <= 0) return 0;
    if (!workspace || workspace_bytes < fsb_raster_workspace(n_isects, (int64_t)C * tile_w * tile_h, D))
        return FSB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(D, (launch_fwd<DD>(C, N, n_isects, n_isects_dev, means2d, conics, colors, opacities, backgrounds, masks, width,
                                      height, tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize,
                                      workspace, out_colors, out_alphas, last_ids, st)));
}

// Gradient outputs are ACCUMULATED into (atomicAdd); the caller zero-fills them first.
// `workspace` is the buffer the matching fsb_raster_fwd call filled.
// v_means2d / v_means2d_abs may be NULL (no gradient wanted for the 2-D means: the legacy normals pass of
// dn_model.py:638 detaches them); v_means2d_abs requires v_means2d.
FSB_API int fsb_raster_bwd(int C, int N, int D, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* backgrounds,
                           const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                           const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_normalize,
                           void* workspace, size_t workspace_bytes, const float* render_colors,
                           const float* render_alphas, const int32_t* last_ids, const float* v_render_colors,
                           const float* v_render_alphas, float* v_means2d_abs, float* v_means2d, float* v_conics,
                           float* v_colors, float* v_opacities, void* stream) {
    if (bad_geometry(C, tile_size, n_isects)) return FSB_E_ARG;
    if (ed_normalize && !render_colors) return FSB_E_ARG;
    if (v_means2d_abs && !v_means2d) return FSB_E_ARG;
    if (tile_w <= 0 || tile_h <= 0 || n_isects == 0) return 0;
    if (!workspace || workspace_bytes < fsb_raster_workspace(n_isects, (int64_t)C * tile_w * tile_h, D))
        return FSB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(D, (launch_bwd<DD>(C, N, n_isects, n_isects_dev, means2d, conics, colors, opacities, backgrounds, masks, width,
                                      height, tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize,
                                      workspace, render_colors, render_alphas, last_ids, v_render_colors,
                                      v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities,
                                      st)));
}

// Measurement aid: counts[2] u64 (device, zero-filled by the caller) += {blended, visited} (pixel, entry) pairs of
// a finished forward pass (see raster_pair_count_kernel).  bench.py turns them into the FP32 roofline of R1 / R2.
FSB_API int fsb_raster_pair_count(int C, int N, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d,
                                  const float* conics, const float* opacities, int width, int height, int tile_size,
                                  int tile_w, int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids,
                                  const int32_t* last_ids, uint64_t* counts, void* stream) {
    (void)N;
    if (bad_geometry(C, tile_size, n_isects) || !counts) return FSB_E_ARG;
    if (tile_w <= 0 || tile_h <= 0 || n_isects == 0) return 0;
    dim3 block(tile_size, tile_size);
    raster_pair_count_kernel<<<(unsigned)((int64_t)C * tile_w * tile_h), block, 0, (cudaStream_t)stream>>>(
        C, n_isects, n_isects_dev, (const float2*)means2d, conics, opacities, width, height, tile_size, tile_w, tile_h,
        tile_offsets, flatten_ids, last_ids, (unsigned long long*)counts);
    FSB_LAUNCH_CHECK();
    return 0;
}
