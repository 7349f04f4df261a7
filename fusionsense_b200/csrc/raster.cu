// raster.cu — tile rasterisation / alpha compositing, forward and backward.
//
// Replaces gsplat 1.0.0's rasterize_to_pixels_{fwd,bwd}_kernel and the legacy rasterize_forward /
// rasterize_backward_kernel (Appendix A.5/A.6), reached from
// /root/reference/dn_splatter/dn_model.py:570-591 (RGB + expected depth, D = 4) and :644-653
// (normals, D = 3, white background).
//
// Round-2 design (profiles/r02b_*: the round-1 kernels were instruction-issue bound at 76-83 % issue-active, and
// the legacy normals pass re-binned, re-sorted and re-composited the whole scene because a handful of Gaussians
// get one more tile from the 0.1.x bounding-box rule):
//
//   packed records   After the sort, raster_pack_kernel writes every list entry ONCE, in sorted order, in the form
//                    the inner loops consume: geo = (x, y, log2 opacity, reach mask), con = (-0.5 log2e a,
//                    -log2e b, -0.5 log2e c, Gaussian id | legacy flag), col = colour row.  A tile's range is then
//                    three contiguous arrays: the compositing kernels stage 256-entry chunks with
//                    cp.async.bulk (1-D bulk copies completing on an mbarrier), double-buffered, instead of the
//                    round-1 two-hop gather (flatten_ids -> five scattered loads) and the per-entry reach test
//                    that three kernels recomputed.
//   reach mask       32 bits, one per 4 x 2 pixel block of the 16 x 16 tile: can alpha >= 1/255 be reached on a
//                    pixel centre of that block?  Computed from the exact x-interval of the ellipse on every pixel
//                    row (one sqrt per row band).  A warp owns an 8 x 4 footprint (four blocks) and walks only the
//                    entries that reach it (work lists built by ballot from the masks).
//   units            A tile whose list holds <= 8 chunks (2048 entries; all but the silhouette tiles of an object
//                    scene, and every tile of a uniform scene) is ONE unit: one CTA walks its chunks front to
//                    back, carries (T, colour) in registers, stops loading when every pixel is finished and writes
//                    the pixels — no per-segment state, no fold, no re-walk.  Longer tiles keep the round-1
//                    segment-parallel scheme (units of 512 entries composited from T = 1 in parallel, folded in
//                    order by one CTA per tile, pixels whose stop rule fires inside a unit re-walk that unit) so
//                    that the 50x-longer-than-median lists of an object silhouette do not serialise.
//   two colour sets  DN-Splatter composites the same Gaussians twice per iteration: RGB + depth through
//                    rasterization() and per-Gaussian normals through the legacy rasterize_gaussians().  Both have
//                    the same alphas, so one walk serves both: channel group A = [0, DA) and group B = [DA, D)
//                    with their own colour arrays, backgrounds and outputs; the 2-D mean gradient takes only group
//                    A's dL/dalpha (dn_model.py:638 detaches the means of the normals pass).  The legacy 0.1.x
//                    bounding box can add a tile to a Gaussian; such list entries carry FSB_LEGACY_FLAG in their id
//                    and belong to group B only.  They are rare (an exact-integer tile edge), so a tile that holds
//                    one is composited twice (A without the flagged entries, B with them) by the same code; every
//                    other tile is composited once.
//   backward         every unit replays independently from the state the forward left (transmittance after the
//                    unit, colour accumulated behind it), back to front; the 8 + D per-Gaussian partials are
//                    reduced by a transposing butterfly that leaves each total in its own lane, so one warp-wide
//                    red.global.add updates all of them.
// The per-pixel loops are FP32 / MUFU / issue bound, not HBM bound.
#include "common.cuh"
#include <stdlib.h>

namespace {

// 1 / x to 1 ulp (MUFU.RCP): the reference kernels are built with fast-math and divide the same way
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

constexpr float ALPHA_MAX = 0.999f;
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float T_MIN = 1e-4f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int MAX_BLOCK = 256;  // tile_size <= 16
constexpr int32_t LEGACY_FLAG = (int32_t)0x80000000;  // FSB_LEGACY_FLAG of include/fsb200.h
constexpr int32_t STOP_MARK = INT32_MIN + 1;  // chain_last value: "this pixel's stop rule fires inside this unit"

// list entries per staged chunk; shorter for wide colour vectors so both buffers stay under 48 KB of smem
__host__ __device__ constexpr int chunk_len(int D) { return D <= 8 ? 256 : 128; }
__host__ __device__ constexpr int color_stride(int D) { return (D + 3) / 4 * 4; }  // floats per packed colour row
constexpr int MAX_LIGHT_CHUNKS = 8;  // a tile with at most `light_chunks` (<= this) chunks is one sequential unit
constexpr int HEAVY_CHUNKS = 2;      // chunks per unit of a longer tile

// Sequential tiles are the efficient form (no per-unit state, no fold, early exit), segment-parallel ones the
// low-latency form.  With many tiles per SM (1080p and up) the grid is deep enough for sequential units of up to
// 2048 entries; a 640 x 480 frame has 1200 tiles for 148 x 4 CTA slots, its kernel time IS the longest unit, so there
// only lists of at most 512 entries stay sequential (r02c: cfg2 forward 0.29 ms with 8, [see profiles/] with 2).
inline int light_chunks_for(int64_t n_tiles) { return n_tiles >= 4 * FSB_NUM_SMS * 4 ? MAX_LIGHT_CHUNKS : 2; }

// ---- async bulk copy (TMA 1-D) + mbarrier ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}

// ---- workspace ------------------------------------------------------------------------------------------------
struct Header {
    int total_units;  // all units (every tile has at least one)
    int n_later;      // units k >= 1 of the segment-parallel tiles
    int n_heavy;      // segment-parallel tiles
    int pad;
};

struct Workspace {
    Header* hdr;
    int32_t* unit_start;   // [n_tiles + 1]
    int32_t* unit_tile;    // [max_units]
    int32_t* unit_order;   // [max_units]  all units, longest first (backward, pack)
    int32_t* tile_order;   // [n_tiles]    tiles by the length of their first unit, longest first (forward A)
    int32_t* later_units;  // [max_units]  units k >= 1 of segment-parallel tiles, longest first (forward B / C)
    int32_t* heavy_tiles;  // [n_tiles]    segment-parallel tiles (fold)
    uint8_t* tile_flag;    // [n_tiles]    the tile's list holds a LEGACY_FLAG entry (split mode only)
    int32_t* done_k;       // [n_tiles, 8] per warp footprint: first unit after which all its pixels are finished
    float* chain_T;        // [max_units, 256]  transmittance after the unit; negative = pixel finished
    int32_t* chain_last;   // [max_units, 256]  last contributing list position so far (-1: none in a local state)
    float* prefix_C;       // [max_units, 256, D]  colour accumulated through the unit (segment-parallel tiles)
    float4* geo;           // [n_isects]  x, y, log2(opacity), reach mask
    float4* con;           // [n_isects]  -0.5 log2e a, -log2e b, -0.5 log2e c, id bits
    float* col;            // [n_isects, DP]
    float4* rec;           // [C * N, 2 + DP / 4]  per Gaussian: (x, y, opacity, 0), (a, b, c, 0), colours (prepack)
};

// a flagged tile takes two unit slots (group A state, group B state), hence 2 * n_tiles
inline int64_t max_units(int64_t n_isects, int64_t n_tiles, int D) {
    return n_isects / (HEAVY_CHUNKS * chunk_len(D)) + 2 * n_tiles;
}

template <bool CARVE>
inline size_t ws_layout(void* base, int64_t n_isects, int64_t n_tiles, int64_t n_gauss, int D, Workspace* w) {
    const int64_t mu = max_units(n_isects, n_tiles, D);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = CARVE ? (char*)base + off : nullptr;
        off += fsb_align_up(bytes, 256);
        return p;
    };
    char* p;
    p = take(256); if (CARVE) w->hdr = (Header*)p;
    p = take((size_t)(n_tiles + 1) * 4); if (CARVE) w->unit_start = (int32_t*)p;
    p = take((size_t)mu * 4); if (CARVE) w->unit_tile = (int32_t*)p;
    p = take((size_t)mu * 4); if (CARVE) w->unit_order = (int32_t*)p;
    p = take((size_t)n_tiles * 4); if (CARVE) w->tile_order = (int32_t*)p;
    p = take((size_t)mu * 4); if (CARVE) w->later_units = (int32_t*)p;
    p = take((size_t)n_tiles * 4); if (CARVE) w->heavy_tiles = (int32_t*)p;
    p = take((size_t)n_tiles); if (CARVE) w->tile_flag = (uint8_t*)p;
    p = take((size_t)n_tiles * 8 * 4); if (CARVE) w->done_k = (int32_t*)p;
    p = take((size_t)mu * MAX_BLOCK * 4); if (CARVE) w->chain_T = (float*)p;
    p = take((size_t)mu * MAX_BLOCK * 4); if (CARVE) w->chain_last = (int32_t*)p;
    p = take((size_t)mu * MAX_BLOCK * D * 4); if (CARVE) w->prefix_C = (float*)p;
    p = take((size_t)(n_isects + 1) * 16); if (CARVE) w->geo = (float4*)p;
    p = take((size_t)(n_isects + 1) * 16); if (CARVE) w->con = (float4*)p;
    p = take((size_t)(n_isects + 1) * color_stride(D) * 4); if (CARVE) w->col = (float*)p;
    p = take((size_t)n_gauss * (32 + color_stride(D) * 4)); if (CARVE) w->rec = (float4*)p;
    return off;
}

// ---- geometry -------------------------------------------------------------------------------------------------
// A warp owns an 8 x 4 pixel footprint made of four 4 x 2 blocks; lane = 8 q + s, q = 2 qx + qy the block, s = 4 sy +
// sx the pixel inside it.  The warps of a 16 x 16 tile lie 2 across and 4 down.  Reach-mask bit = 4 warp + q.
struct TileGeom {
    int cam, tile_x, tile_y;
    int block_size, tr, lane, warp, n_warps, warps_x;
    int i, j;
    float px, py;
    bool inside;
};

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
    g.warps_x = tile_size >> 3;
    const int wy = g.warp / g.warps_x, wx = g.warp - wy * g.warps_x;
    const int q = g.lane >> 3, s = g.lane & 7;
    g.i = g.tile_y * tile_size + wy * 4 + (q & 1) * 2 + (s >> 2);
    g.j = g.tile_x * tile_size + wx * 8 + (q >> 1) * 4 + (s & 3);
    g.px = (float)g.j + 0.5f;
    g.py = (float)g.i + 0.5f;
    g.inside = (g.i < height && g.j < width);
    return g;
}

// Reach mask of a Gaussian on a tile (origin tx0, ty0 in pixels): bits 4 w .. 4 w + 3 are set when alpha >= 1/255 can
// hold on a pixel centre of warp w's 8 x 4 footprint.  alpha >= 1/255 means q(d) = a dx^2 + 2 b dx dy + c dy^2 <= thr =
// 2 ln(255 opacity): an ellipse with |dy| <= Y = sqrt(thr a / det), whose right / left boundary at height dy is
// dx = -b dy / a +- sqrt(thr a - det dy^2) / a — concave / convex in dy with the extreme points at dy = -+ s,
// s = b sqrt(thr / (det c)).  For a band of four pixel rows [d0, d0 + 3] (clipped to [-Y, Y]) the x-extent of the
// ellipse over the band is therefore [L(clamp(s)), R(clamp(-s))]: two square roots per band, no loop over rows; the
// continuous band is a superset of its four discrete rows.  The column blocks (8 pixel centres each) that meet the
// extent get their bits.  The threshold is inflated (1e-2 absolute, 2e-4 relative: alpha may be 0.5 % below 1/255),
// Y by 1e-3 and the extent by 2e-3 px, far above the fp32 error of this arithmetic (rsqrt.approx is good to 2 ulp);
// anything doubtful (non-PD conic, NaN, opacity ~ inf) keeps all bits.  Uniform control flow: every lane of the
// packing warp runs the same four (two for 8 x 8 tiles) band iterations.
__device__ __forceinline__ uint32_t reach_mask(float gx, float gy, float opac, float a, float b, float c, float tx0,
                                               float ty0, int tile_size) {
    const int n_warps = (tile_size * tile_size) >> 5;
    const uint32_t all = n_warps >= 8 ? 0xffffffffu : ((1u << (4 * n_warps)) - 1u);
    const float tau = __logf(255.f * opac);
    if (tau + 2e-3f < 0.f) return 0u;  // opacity below 1/255: can never pass the alpha test
    const float det = a * c - b * b;
    if (!(det > 0.f) || !(a > 0.f) || !(c > 0.f) || !(tau < 1e30f)) return all;
    const float thr = 2.f * tau * 1.0002f + 1e-2f;
    const float inv_a = __frcp_rn(a);
    const float ta = thr * a;
    const float nb = -b * inv_a;
    const float Y = ta * rsqrtf(ta * det) * 1.001f + 1e-3f;  // sqrt(thr a / det), inflated
    const float sx = b * (thr * rsqrtf(thr * det * c));      // b sqrt(thr / (det c))
    const float y_rel = ty0 + 0.5f - gy;                     // dy of pixel row 0
    const float x_rel = tx0 + 0.5f - gx;                     // dx of pixel column 0
    const int warps_x = tile_size >> 3, n_bands = tile_size >> 2;
    uint32_t m = 0u;
#pragma unroll
    for (int band = 0; band < 4; ++band) {
        if (band < n_bands) {
            const float d0 = y_rel + (float)(4 * band);
            const float e0 = fmaxf(d0, -Y), e1 = fminf(d0 + 3.f, Y);
            const float tR = fminf(fmaxf(-sx, e0), e1), tL = fminf(fmaxf(sx, e0), e1);
            const float dR = fmaxf(fmaf(-det * tR, tR, ta), 0.f), dL = fmaxf(fmaf(-det * tL, tL, ta), 0.f);
            const float hi = fmaf(nb, tR, dR * rsqrtf(fmaxf(dR, 1e-30f)) * inv_a);
            const float lo = fmaf(nb, tL, -(dL * rsqrtf(fmaxf(dL, 1e-30f)) * inv_a));
            const float pad = 2e-3f + 2e-6f * fmaxf(fabsf(lo), fabsf(hi));
            // pixel centres of column block k: dx in [x_rel + 8 k, x_rel + 8 k + 7]
            const float fl = (lo - pad - x_rel - 7.f) * 0.125f, fh = (hi + pad - x_rel) * 0.125f;
            int k_lo = max(0, __float2int_ru(fmaxf(fl, -1.f)));
            int k_hi = min(warps_x - 1, __float2int_rd(fminf(fh, 64.f)));
            if (!(e1 >= e0)) k_hi = -1;  // the band misses the ellipse's rows
            k_lo = min(k_lo, 4);
            k_hi = max(k_hi, -1);
            // footprint (wy = band, wx = k) owns bits 4 (band warps_x + k) .. + 3
            const uint32_t cols = ((1u << (4 * (k_hi + 1))) - 1u) & ~((1u << (4 * k_lo)) - 1u);
            m |= cols << (4 * band * warps_x);
        }
    }
    return m & all;
}

// ---- unit table: one block scans the per-tile unit counts and orders the work ------------------------------------
constexpr int LPT_BINS = 16;

struct TileRange {
    int32_t b, e;
};
__device__ __forceinline__ TileRange tile_range(const int32_t* __restrict__ tile_offsets, int64_t t, int64_t n_tiles,
                                                int64_t n_isects) {
    TileRange r;
    r.b = tile_offsets[t];
    r.e = (t == n_tiles - 1) ? (int32_t)n_isects : tile_offsets[t + 1];
    if (r.e > (int32_t)n_isects) r.e = (int32_t)n_isects;  // static-capacity overflow: the step is void anyway
    if (r.b > r.e) r.b = r.e;
    return r;
}

struct UnitShape {
    int n_chunks, n_units;
    bool light;
};
__device__ __forceinline__ UnitShape unit_shape(int len, int CH, bool flagged, int light_chunks) {
    UnitShape s;
    s.n_chunks = (len + CH - 1) / CH;
    s.light = flagged || s.n_chunks <= light_chunks;
    s.n_units = s.light ? (flagged ? 2 : 1) : (s.n_chunks + HEAVY_CHUNKS - 1) / HEAVY_CHUNKS;
    return s;
}
// list entries unit k of a tile covers (a flagged tile's two units both cover the whole list)
__device__ __forceinline__ int unit_len(int len, int CH, bool flagged, int light_chunks, int k) {
    const UnitShape s = unit_shape(len, CH, flagged, light_chunks);
    if (s.light) return len;
    return min(HEAVY_CHUNKS * CH, max(0, len - k * HEAVY_CHUNKS * CH));
}

__global__ void __launch_bounds__(1024)
unit_table_kernel(int n_tiles, int64_t n_isects, const int64_t* __restrict__ n_dev, int CH, int light_chunks,
                  const int32_t* __restrict__ tile_offsets, const uint8_t* __restrict__ masks, bool split,
                  Workspace ws) {
    n_isects = fsb_eff_n(n_isects, n_dev);
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    __shared__ int s_bins[3][LPT_BINS + 1];
    __shared__ int s_heavy;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_carry = 0; s_heavy = 0; }
    __syncthreads();
    auto tile_len = [&](int t) {
        if (masks != nullptr && !masks[t]) return 0;
        const TileRange r = tile_range(tile_offsets, t, n_tiles, n_isects);
        return (int)(r.e - r.b);
    };
    auto flagged = [&](int t) { return split && ws.tile_flag[t] != 0; };
    // every thread takes TPT consecutive tiles: its loads are issued together, and the whole table needs one block scan
    // (a 1024-tile stripe per iteration was 8 dependent rounds of global loads + three barriers each at 1080p: this
    // single-CTA kernel sits on the forward's critical path, r02k: 44 us)
    const int TPT = (n_tiles + 1023) / 1024;
    const int t0 = tid * TPT, t1 = min(n_tiles, t0 + TPT);
    int nu_sum = 0;
    for (int t = t0; t < t1; ++t) nu_sum += unit_shape(tile_len(t), CH, flagged(t), light_chunks).n_units;
    int inc = nu_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int excl = inc - nu_sum;
    for (int w = 0; w < warp; ++w) excl += s_warp[w];
    if (tid == 1023) s_carry = excl + nu_sum;
    for (int t = t0; t < t1; ++t) {
        const UnitShape sh = unit_shape(tile_len(t), CH, flagged(t), light_chunks);
        ws.unit_start[t] = excl;
        for (int k = 0; k < sh.n_units; ++k) ws.unit_tile[excl + k] = t;
        if (!sh.light) ws.heavy_tiles[atomicAdd(&s_heavy, 1)] = t;
        excl += sh.n_units;
    }
    __syncthreads();
    const int total = s_carry;
    if (tid == 0) {
        ws.unit_start[n_tiles] = total;
        ws.hdr->total_units = total;
        ws.hdr->n_heavy = s_heavy;
    }
    // Launch orders, longest work unit first (LPT): CTA durations follow the unit length, and with grid order = tile
    // order a few long units picked up late leave most SMs idle in the kernel's tail.  Counting sort into LPT_BINS
    // length classes; the order inside a class is arbitrary (results do not depend on which CTA runs which unit).
    // A deep grid of sequential tiles only (a uniform 1080p scene: 8160 units for ~600 CTA slots, nothing segment-
    // parallel) gains nothing from the ordering, and this single-CTA kernel sits on the forward's critical path
    // (r02d: 43 us with the sort, at cfg4): identity orders there.
    if (s_heavy == 0 && n_tiles >= 4 * FSB_NUM_SMS * 4) {
        for (int t = tid; t < n_tiles; t += 1024) ws.tile_order[t] = t;
        for (int u = tid; u < total; u += 1024) ws.unit_order[u] = u;
        if (tid == 0) ws.hdr->n_later = 0;
        return;
    }
    if (tid < 3 * (LPT_BINS + 1)) (&s_bins[0][0])[tid] = 0;
    __syncthreads();
    const int max_len = MAX_LIGHT_CHUNKS * CH;
    auto bin_of = [&](int n) { return (int)(((int64_t)(max_len - min(n, max_len)) * LPT_BINS) / (max_len + 1)); };
    for (int pass = 0; pass < 2; ++pass) {  // 0: count, 1: scatter
        for (int base = 0; base < n_tiles; base += 1024) {  // list 0: tiles by first unit
            const int t = base + tid;
            const int bin = (t < n_tiles) ? bin_of(unit_len(tile_len(t), CH, flagged(t), light_chunks, 0)) : LPT_BINS;
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            const int leader = __ffs(peers) - 1;
            int at = 0;
            if (lane == leader) at = atomicAdd(&s_bins[0][bin], __popc(peers));
            at = __shfl_sync(0xffffffffu, at, leader) + __popc(peers & ((1u << lane) - 1u));
            if (pass == 1 && t < n_tiles) ws.tile_order[at] = t;
        }
        for (int base = 0; base < total; base += 1024) {  // list 1: all units; list 2: later units of heavy tiles
            const int u = base + tid;
            int bin = LPT_BINS, bin2 = LPT_BINS;
            if (u < total) {
                const int t = ws.unit_tile[u];
                const int k = u - ws.unit_start[t];
                const int len = tile_len(t);
                const bool fl = flagged(t);
                bin = bin_of(unit_len(len, CH, fl, light_chunks, k));
                if (k >= 1 && !unit_shape(len, CH, fl, light_chunks).light) bin2 = bin;
            }
            unsigned peers = __match_any_sync(0xffffffffu, bin);
            int leader = __ffs(peers) - 1;
            int at = 0;
            if (lane == leader) at = atomicAdd(&s_bins[1][bin], __popc(peers));
            at = __shfl_sync(0xffffffffu, at, leader) + __popc(peers & ((1u << lane) - 1u));
            if (pass == 1 && u < total) ws.unit_order[at] = u;
            peers = __match_any_sync(0xffffffffu, bin2);
            leader = __ffs(peers) - 1;
            at = 0;
            if (lane == leader) at = atomicAdd(&s_bins[2][bin2], __popc(peers));
            at = __shfl_sync(0xffffffffu, at, leader) + __popc(peers & ((1u << lane) - 1u));
            if (pass == 1 && bin2 < LPT_BINS) ws.later_units[at] = u;
        }
        __syncthreads();
        if (pass == 0 && tid < 3) {  // counts -> exclusive start positions
            int run = 0;
            for (int b = 0; b < LPT_BINS; ++b) {
                const int c = s_bins[tid][b];
                s_bins[tid][b] = run;
                run += c;
            }
            s_bins[tid][LPT_BINS] = 0;
            if (tid == 2) ws.hdr->n_later = run;
        }
        __syncthreads();
    }
}

// split mode: mark the tiles whose list holds a LEGACY_FLAG entry (rare: binary search only for those)
__global__ void __launch_bounds__(256)
tile_flag_kernel(int64_t n_isects, const int64_t* __restrict__ n_dev, const int32_t* __restrict__ flatten_ids,
                 const int32_t* __restrict__ tile_offsets, int64_t n_tiles, uint8_t* __restrict__ tile_flag) {
    n_isects = fsb_eff_n(n_isects, n_dev);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_isects; e += stride) {
        if (flatten_ids[e] >= 0) continue;
        int64_t lo = 0, hi = n_tiles - 1;  // last tile whose offset is <= e
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if ((int64_t)tile_offsets[mid] <= e) lo = mid; else hi = mid - 1;
        }
        tile_flag[lo] = 1;
    }
}

// ---- chunk range of a unit ------------------------------------------------------------------------------------------
struct UnitGeom {
    int64_t tile_lin;
    int k, n_units;
    int32_t rb, re;  // list range of the tile
    int c0, c1;      // chunk range [c0, c1) of this unit, chunks counted from rb
    bool light, flagged, masked;
};

__device__ __forceinline__ UnitGeom unit_geom(const Workspace& ws, int u, int64_t tile_lin, int64_t n_cam_tiles,
                                              int64_t n_isects, const int32_t* __restrict__ tile_offsets,
                                              const uint8_t* __restrict__ masks, bool split, int CH,
                                              int light_chunks) {
    UnitGeom g;
    g.tile_lin = tile_lin;
    const int s0 = ws.unit_start[tile_lin];
    g.k = u - s0;
    g.n_units = ws.unit_start[tile_lin + 1] - s0;
    g.masked = (masks != nullptr && !masks[tile_lin]);
    g.flagged = split && ws.tile_flag[tile_lin] != 0;
    const TileRange r = tile_range(tile_offsets, tile_lin, n_cam_tiles, n_isects);
    g.rb = r.b;
    g.re = g.masked ? r.b : r.e;
    const UnitShape s = unit_shape(g.re - g.rb, CH, g.flagged, light_chunks);
    g.light = s.light;
    if (s.light) { g.c0 = 0; g.c1 = s.n_chunks; }
    else { g.c0 = g.k * HEAVY_CHUNKS; g.c1 = min(g.c0 + HEAVY_CHUNKS, s.n_chunks); }
    return g;
}

// ---- pack: sorted list -> records the compositing kernels bulk-load -----------------------------------------------
struct PackIn {
    const float2* means2d;
    const float* conics;
    const float* opacities;
    const float* colors_a;
    const float* colors_b;
    const int32_t* flatten_ids;
};

// Per (camera, Gaussian): the fields the pack gathers, side by side in one 32 + 4 DP byte record.  The pack runs once
// per LIST ENTRY (7 per visible Gaussian at 1080p) and was bound by L1 request slots — ten scattered requests over
// five arrays per entry (r02i / r02k ncu: lg_throttle + long_scoreboard, 236 us for 82 M instructions); with the
// record it issues 2 + DP / 4 16-byte requests into two or three adjacent sectors.
template <int D, int DA>
__global__ void __launch_bounds__(256)
raster_prepack_kernel(int64_t n_gauss, PackIn in, float4* __restrict__ rec) {
    constexpr int DP = color_stride(D), DB = D - DA, RS = 2 + DP / 4;
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= n_gauss) return;
    const float2 xy = in.means2d[g];
    float4* r = rec + g * RS;
    r[0] = make_float4(xy.x, xy.y, in.opacities[g], 0.f);
    r[1] = make_float4(in.conics[3 * g], in.conics[3 * g + 1], in.conics[3 * g + 2], 0.f);
    float row[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) row[k] = 0.f;
#pragma unroll
    for (int k = 0; k < DA; ++k) row[k] = in.colors_a[g * DA + k];
    if constexpr (DB > 0) {
#pragma unroll
        for (int k = 0; k < DB; ++k) row[DA + k] = in.colors_b[g * DB + k];
    }
#pragma unroll
    for (int k = 0; k < DP / 4; ++k) r[2 + k] = make_float4(row[4 * k], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]);
}

// REC: gather from the prepack records (lists with several entries per Gaussian), else from the five arrays.
template <int D, int DA, bool REC>
__global__ void __launch_bounds__(256)
raster_pack_kernel(int C, int64_t n_isects, const int64_t* __restrict__ n_dev, PackIn in,
                   const uint8_t* __restrict__ masks, int tile_size, int tile_w, int tile_h, int light_chunks,
                   const int32_t* __restrict__ tile_offsets, Workspace ws) {
    constexpr int CH = chunk_len(D), DP = color_stride(D), DB = D - DA, RS = 2 + DP / 4;
    n_isects = fsb_eff_n(n_isects, n_dev);
    if ((int)blockIdx.x >= ws.hdr->total_units) return;
    const int u = ws.unit_order[blockIdx.x];
    const int64_t tile_lin = ws.unit_tile[u];
    const UnitGeom ug = unit_geom(ws, u, tile_lin, (int64_t)C * tile_w * tile_h, n_isects, tile_offsets, masks, DB > 0, CH,
                                  light_chunks);
    if (ug.flagged && ug.k == 1) return;  // the second unit of a flagged tile shares the first one's records
    const int n_tiles = tile_w * tile_h;
    const int tile_id = (int)(tile_lin % n_tiles);
    const float tx0 = (float)((tile_id % tile_w) * tile_size), ty0 = (float)((tile_id / tile_w) * tile_size);
    const int32_t eb = ug.rb + ug.c0 * CH, ee = min(ug.re, ug.rb + ug.c1 * CH);
    for (int32_t e = eb + (int32_t)threadIdx.x; e < ee; e += 256) {
        const int32_t raw = in.flatten_ids[e];
        const int32_t g = raw & ~LEGACY_FLAG;
        float4* dst = reinterpret_cast<float4*>(ws.col + (size_t)e * DP);
        float x, y, o, a, b, c;
        if constexpr (REC) {
            const float4* r = ws.rec + (size_t)g * RS;
            const float4 r0 = r[0], r1 = r[1];
            x = r0.x; y = r0.y; o = r0.z; a = r1.x; b = r1.y; c = r1.z;
#pragma unroll
            for (int k = 0; k < DP / 4; ++k) dst[k] = r[2 + k];
        } else {
            const float2 xy = in.means2d[g];
            x = xy.x; y = xy.y; o = in.opacities[g];
            a = in.conics[3 * (size_t)g]; b = in.conics[3 * (size_t)g + 1]; c = in.conics[3 * (size_t)g + 2];
            float row[DP];
#pragma unroll
            for (int k = 0; k < DP; ++k) row[k] = 0.f;
            const float* ca = in.colors_a + (size_t)g * DA;
            if constexpr (DA == 4) {
                const float4 c4 = *reinterpret_cast<const float4*>(ca);
                row[0] = c4.x; row[1] = c4.y; row[2] = c4.z; row[3] = c4.w;
            } else {
#pragma unroll
                for (int k = 0; k < DA; ++k) row[k] = ca[k];
            }
            if constexpr (DB > 0) {
                const float* cb = in.colors_b + (size_t)g * DB;
#pragma unroll
                for (int k = 0; k < DB; ++k) row[DA + k] = cb[k];
            }
#pragma unroll
            for (int k = 0; k < DP / 4; ++k) dst[k] = make_float4(row[4 * k], row[4 * k + 1], row[4 * k + 2], row[4 * k + 3]);
        }
        const uint32_t m = reach_mask(x, y, o, a, b, c, tx0, ty0, tile_size);
        ws.geo[e] = make_float4(x, y, __log2f(o), __int_as_float((int)m));
        ws.con[e] = make_float4(-0.5f * LOG2E * a, -LOG2E * b, -0.5f * LOG2E * c, __int_as_float(raw));
    }
}

// ---- staged chunk ----------------------------------------------------------------------------------------------------
// Two buffers of CH records (+ one dummy record with alpha = 0 that pads the work lists), the per-warp work lists and
// the two mbarriers.  alpha = min(0.999, 2^(p + L)),  p = a' dx^2 + b' dx dy + c' dy^2 (= -log2e sigma), L = log2(opacity):
// one evaluation is 2 FADD + 6 FMUL/FFMA + 1 FADD + MUFU.EX2 + FMNMX, and "sigma < 0" is "p > 0".
template <int D>
struct Stage {
    static constexpr int CH = chunk_len(D);
    static constexpr int DP = color_stride(D);
    float4 geo[2][CH + 1];
    float4 con[2][CH + 1];
    float col[2][(CH + 1) * DP];
    alignas(8) uint16_t wlist[MAX_BLOCK / 32][CH + 4];
    alignas(8) uint64_t bar[2];
};

template <int D>
struct Loader {
    Stage<D>& s;
    const float4* g_geo;
    const float4* g_con;
    const float* g_col;
    uint32_t parity;  // bit b: phase parity the next wait on buffer b expects
    __device__ Loader(Stage<D>& s_, const Workspace& ws_)
        : s(s_), g_geo(ws_.geo), g_con(ws_.con), g_col(ws_.col), parity(0u) {}

    // once per CTA, before the first unit
    __device__ __forceinline__ void init(int tr) {
        constexpr int CH = Stage<D>::CH, DP = Stage<D>::DP;
        if (tr == 0) {
            mbar_init(&s.bar[0], 1);
            mbar_init(&s.bar[1], 1);
            mbar_fence_init();
        }
        if (tr < 2) {
            s.geo[tr][CH] = make_float4(0.f, 0.f, -1000.f, 0.f);  // 2^-1000 = 0: fails the alpha test
            s.con[tr][CH] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < DP; ++k) s.col[tr][CH * DP + k] = 0.f;
        }
        __syncthreads();
    }
    // one thread: start the copy of list entries [e0, e0 + n) into buffer b
    __device__ __forceinline__ void issue(int b, int32_t e0, int n) {
        constexpr int DP = Stage<D>::DP;
        if (n <= 0) return;
        mbar_expect_tx(&s.bar[b], (uint32_t)n * (32u + 4u * DP));
        bulk_g2s(&s.geo[b][0], g_geo + e0, (uint32_t)n * 16u, &s.bar[b]);
        bulk_g2s(&s.con[b][0], g_con + e0, (uint32_t)n * 16u, &s.bar[b]);
        bulk_g2s(&s.col[b][0], g_col + (size_t)e0 * DP, (uint32_t)n * 4u * DP, &s.bar[b]);
    }
    // all threads: wait until buffer b holds the copy issued for it (n as given to issue)
    __device__ __forceinline__ void wait(int b, int n) {
        if (n <= 0) return;
        mbar_wait(&s.bar[b], (parity >> b) & 1u);
        parity ^= 1u << b;
    }
};

// Work list of this warp over entries [0, n) of buffer b: the entries whose reach mask touches the warp's footprint
// (skip_flagged: and that are not LEGACY_FLAG entries).  Padded to a multiple of four with the dummy slot.
template <int D>
__device__ __forceinline__ int build_list_bits(Stage<D>& s, int b, int n, int warp, int lane, uint32_t wbits,
                                               bool skip_flagged) {
    constexpr int CH = Stage<D>::CH;
    uint16_t* wl = s.wlist[warp];
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int tt = k0 + lane;
        bool bit = false;
        if (tt < n) {
            bit = ((uint32_t)__float_as_int(s.geo[b][tt].w) & wbits) != 0u;
            if (skip_flagged && __float_as_int(s.con[b][tt].w) < 0) bit = false;
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, bit);
        if (bit) wl[base + __popc(bits & ((1u << lane) - 1u))] = (uint16_t)tt;
        base += __popc(bits);
    }
    if (lane < 4) wl[base + lane] = (uint16_t)CH;
    __syncwarp();
    return base;
}
template <int D>
__device__ __forceinline__ int build_list(Stage<D>& s, int b, int n, const TileGeom& tg, bool skip_flagged) {
    return build_list_bits<D>(s, b, n, tg.warp, tg.lane, 0xfu << (4 * tg.warp), skip_flagged);
}

// alpha of a staged entry at this thread's pixel; `p` receives -log2e * sigma, `au` the unclamped opacity * vis
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
__device__ __forceinline__ void walk(const Stage<D>& s, int b, int cnt, int32_t chunk_b, const TileGeom& tg, float& T,
                                     float (&acc)[D], int32_t& last, bool& done) {
    constexpr int DP = Stage<D>::DP;
    if (__all_sync(0xffffffffu, done)) return;
    const uint16_t* wl = s.wlist[tg.warp];
    const float4* geo = s.geo[b];
    const float4* con = s.con[b];
    const float* col = s.col[b];
    for (int i = 0; i < cnt; i += 4) {
        const uint2 pk = *reinterpret_cast<const uint2*>(wl + i);
        const int t[4] = {(int)(pk.x & 0xffffu), (int)(pk.x >> 16), (int)(pk.y & 0xffffu), (int)(pk.y >> 16)};
        float alpha[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float dx, dy, p, au;
            alpha[u] = eval_alpha(geo[t[u]], con[t[u]], tg.px, tg.py, dx, dy, p, au);
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
                    const float4* cp = reinterpret_cast<const float4*>(col + t[u] * DP);
#pragma unroll
                    for (int k4 = 0; k4 < DP / 4; ++k4) {
                        const float4 c4 = cp[k4];
                        if (4 * k4 + 0 < D) acc[4 * k4 + 0] = fmaf(c4.x, w, acc[4 * k4 + 0]);
                        if (4 * k4 + 1 < D) acc[(4 * k4 + 1) % D] = fmaf(c4.y, w, acc[(4 * k4 + 1) % D]);
                        if (4 * k4 + 2 < D) acc[(4 * k4 + 2) % D] = fmaf(c4.z, w, acc[(4 * k4 + 2) % D]);
                        if (4 * k4 + 3 < D) acc[(4 * k4 + 3) % D] = fmaf(c4.w, w, acc[(4 * k4 + 3) % D]);
                    }
                    last = chunk_b + t[u];
                    T = next_T;
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
    }
}

struct RasterArgs {
    int C, N;
    int64_t n_isects;
    const int64_t* n_dev;
    const uint8_t* masks;
    int width, height, tile_size, tile_w, tile_h;
    const int32_t* tile_offsets;
    const float* backgrounds_a;  // [C, DA] nullable
    const float* backgrounds_b;  // [C, DB] nullable
    int ed_channel;              // channel divided by max(alpha, 1e-10) on output, -1: none
    int light_chunks;            // see light_chunks_for()
};

struct FwdOut {
    float* out_a;        // [C, H, W, DA]
    float* out_b;        // [C, H, W, DB]
    float* out_alphas;   // [C, H, W]
    int32_t* last_ids;   // [C, H, W]
};

// which = 0: both groups (+ alpha, last id), 1: group A (+ alpha, last id), 2: group B only
template <int D, int DA>
__device__ __forceinline__ void write_pixel(const RasterArgs& a, const FwdOut& o, const TileGeom& tg, int64_t pix,
                                            bool masked, float T, const float (&acc)[D], int32_t last, int which) {
    constexpr int DB = D - DA;
    const float alpha_out = masked ? 0.f : 1.f - T;
    if (which != 2) {
        o.out_alphas[pix] = alpha_out;
        o.last_ids[pix] = last;
#pragma unroll
        for (int c = 0; c < DA; ++c) {
            float v = a.backgrounds_a ? acc[c] + (1.f - alpha_out) * a.backgrounds_a[tg.cam * DA + c] : acc[c];
            if (c == a.ed_channel) v = v / fmaxf(alpha_out, 1e-10f);
            o.out_a[pix * DA + c] = v;
        }
    }
    if constexpr (DB > 0) {
        if (which != 1) {
#pragma unroll
            for (int c = 0; c < DB; ++c) {
                const float v = a.backgrounds_b ? acc[DA + c] + (1.f - alpha_out) * a.backgrounds_b[tg.cam * DB + c]
                                                : acc[DA + c];
                o.out_b[pix * DB + c] = v;
            }
        }
    }
}

// Walk chunks [c0, c1) of a tile's list front to back from the state (T, acc, last, done); double-buffered bulk
// loads.  Returns with every issued copy consumed.  All threads of the CTA call it with the same arguments.
template <int D>
__device__ __forceinline__ void walk_chunks(Stage<D>& s, Loader<D>& ld, const UnitGeom& ug, const TileGeom& tg,
                                            bool skip_flagged, float& T, float (&acc)[D], int32_t& last, bool& done) {
    constexpr int CH = Stage<D>::CH;
    auto chunk_n = [&](int c) { return min(CH, (int)(ug.re - ug.rb) - c * CH); };
    if (ug.c0 >= ug.c1) return;
    if (tg.tr == 0) ld.issue(0, ug.rb + ug.c0 * CH, chunk_n(ug.c0));
    for (int c = ug.c0; c < ug.c1; ++c) {
        const int b = (c - ug.c0) & 1;
        const bool more = c + 1 < ug.c1;
        if (more && tg.tr == 0) ld.issue(b ^ 1, ug.rb + (c + 1) * CH, chunk_n(c + 1));
        const int n = chunk_n(c);
        ld.wait(b, n);
        const int cnt = build_list<D>(s, b, n, tg, skip_flagged);
        walk<D>(s, b, cnt, ug.rb + c * CH, tg, T, acc, last, done);
        // the barrier also orders this chunk's reads before the next copy into the same buffer
        const bool all_done = __syncthreads_and(done);
        if (all_done) {
            if (more) ld.wait(b ^ 1, chunk_n(c + 1));  // never leave a copy in flight
            break;
        }
    }
}

// forward A (PHASE 0, grid = tiles): sequential tiles are finished here; the first unit of every segment-parallel tile
// composites from T = 1.  forward B (PHASE 1, grid-stride over the later units of segment-parallel tiles): same, but a
// warp footprint whose 32 pixels were all finished by an earlier unit (recorded in done_k by phase 0, or
// opportunistically by an earlier unit of this phase) is skipped: nothing it could composite is ever used.
template <int D, int DA, int PHASE>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_fwd_kernel(RasterArgs a, Workspace ws, FwdOut o) {
    constexpr int CH = chunk_len(D);
    constexpr bool SPLIT = (DA < D);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage<D>& s = *reinterpret_cast<Stage<D>*>(smem_raw);
    const int64_t n_isects = fsb_eff_n(a.n_isects, a.n_dev);
    const int64_t n_cam_tiles = (int64_t)a.C * a.tile_w * a.tile_h;
    Loader<D> ld(s, ws);
    const int tr0 = threadIdx.y * blockDim.x + threadIdx.x;
    ld.init(tr0);
    const int n_work = (PHASE == 0) ? (int)gridDim.x : ws.hdr->n_later;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        int u;
        int64_t tile_lin;
        if (PHASE == 0) {
            tile_lin = ws.tile_order[w];
            u = ws.unit_start[tile_lin];
        } else {
            u = ws.later_units[w];
            tile_lin = ws.unit_tile[u];
        }
        const UnitGeom ug = unit_geom(ws, u, tile_lin, n_cam_tiles, n_isects, a.tile_offsets, a.masks, SPLIT, CH,
                                      a.light_chunks);
        const TileGeom tg = tile_geom(tile_lin, a.tile_w, a.tile_h, a.tile_size, a.width, a.height);
        const int64_t pix = tg.inside ? ((int64_t)tg.cam * a.height + tg.i) * a.width + tg.j : 0;
        int32_t* done_k = ws.done_k + tile_lin * 8 + tg.warp;
        bool skip = false;
        if (PHASE == 1) {
            skip = (*done_k < ug.k);
            // a skipped footprint leaves no state behind, but its slots must not keep a stale STOP_MARK from an
            // earlier call that used the same workspace memory (raster_stop_kernel scans chain_last of later units)
            if (skip) ws.chain_last[(size_t)u * MAX_BLOCK + tg.tr] = -1;
            if (__syncthreads_and(skip)) continue;
        }
        const int n_pass = ug.flagged ? 2 : 1;
        for (int pass = 0; pass < n_pass; ++pass) {
            float T = 1.f;
            float acc[D];
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = 0.f;
            int32_t last = (ug.k == 0) ? 0 : -1;
            bool done = !tg.inside || skip;
            // flagged tile: pass 0 = group A without the legacy-only entries, pass 1 = group B with them
            walk_chunks<D>(s, ld, ug, tg, ug.flagged && pass == 0, T, acc, last, done);
            if (skip) continue;
            const size_t cidx = (size_t)(u + pass) * MAX_BLOCK + tg.tr;
            if (ug.light) {
                ws.chain_T[cidx] = T;
                if (ug.flagged && pass == 1) ws.chain_last[cidx] = last;  // the backward's last id of group B
                if (tg.inside) write_pixel<D, DA>(a, o, tg, pix, ug.masked, T, acc, last, ug.flagged ? pass + 1 : 0);
            } else {
                if (__all_sync(0xffffffffu, done) && tg.lane == 0) atomicMin(done_k, ug.k);
                ws.chain_T[cidx] = (done && tg.inside) ? -T : T;
                ws.chain_last[cidx] = last;
#pragma unroll
                for (int c = 0; c < D; ++c) ws.prefix_C[cidx * D + c] = acc[c];
            }
        }
    }
}

// forward B': fold the unit states of every segment-parallel tile in list order (pure arithmetic, no list walk).
// Overwrites the local states with the chained ones; a pixel whose stop rule fires inside unit k gets
// chain_last[k] = STOP_MARK and is finished by raster_stop_kernel.
template <int D, int DA>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_fold_kernel(RasterArgs a, Workspace ws, FwdOut o) {
    const int n_heavy = ws.hdr->n_heavy;
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int64_t tile_lin = ws.heavy_tiles[h];
        const int u0 = ws.unit_start[tile_lin];
        const int n_units = ws.unit_start[tile_lin + 1] - u0;
        const TileGeom tg = tile_geom(tile_lin, a.tile_w, a.tile_h, a.tile_size, a.width, a.height);
        size_t cidx = (size_t)u0 * MAX_BLOCK + tg.tr;
        const float t0 = ws.chain_T[cidx];
        float T = fabsf(t0);
        bool done = (t0 < 0.f) || !tg.inside;
        bool pending = false;  // stop unit found, result still to be produced by raster_stop_kernel
        int32_t last = ws.chain_last[cidx];
        float acc[D];
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = ws.prefix_C[cidx * D + c];
        for (int k = 1; k < n_units; ++k) {
            if (__syncthreads_count(done) == tg.block_size) break;
            if (done) continue;
            cidx = (size_t)(u0 + k) * MAX_BLOCK + tg.tr;
            const float tl = ws.chain_T[cidx];
            const float Tl = fabsf(tl);
            // the stop rule cannot have fired inside this unit (the margin keeps the decision on the safe side of the
            // rounding difference between T * Tl and the sequential fmaf chain: a doubtful pixel is re-walked, and
            // a re-walk that does not stop simply carries on, see raster_stop_kernel)
            if (!(tl < 0.f) && T * Tl > T_MIN * 1.001f) {
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
        __syncthreads();  // the early-exit vote above is taken by all threads of the block
        if (pending) continue;
        // the backward takes the tile total from the last unit's slot
        const size_t lidx = (size_t)(u0 + n_units - 1) * MAX_BLOCK + tg.tr;
#pragma unroll
        for (int c = 0; c < D; ++c) ws.prefix_C[lidx * D + c] = acc[c];
        if (tg.inside) {
            const bool masked = (a.masks != nullptr && !a.masks[tile_lin]);
            const int64_t pix = ((int64_t)tg.cam * a.height + tg.i) * a.width + tg.j;
            write_pixel<D, DA>(a, o, tg, pix, masked, T, acc, last, 0);
        }
    }
}

// forward C: the pixels marked in unit k re-walk the list from unit k on, exactly, from their incoming (chained)
// state: stop position and last id are those of the sequential rule.  A marked pixel that does not stop inside unit
// k (the fold's test is conservative) carries on through the following units, writing the chained state as it goes.
template <int D, int DA>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_stop_kernel(RasterArgs a, Workspace ws, FwdOut o) {
    constexpr int CH = chunk_len(D);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage<D>& s = *reinterpret_cast<Stage<D>*>(smem_raw);
    const int64_t n_isects = fsb_eff_n(a.n_isects, a.n_dev);
    const int64_t n_cam_tiles = (int64_t)a.C * a.tile_w * a.tile_h;
    Loader<D> ld(s, ws);
    const int tr0 = threadIdx.y * blockDim.x + threadIdx.x;
    ld.init(tr0);
    const int n_work = ws.hdr->n_later;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int u = ws.later_units[w];
        const int64_t tile_lin = ws.unit_tile[u];
        const size_t cidx = (size_t)u * MAX_BLOCK + tr0;
        const bool redo = (ws.chain_last[cidx] == STOP_MARK);
        if (!__syncthreads_or(redo)) continue;
        UnitGeom ug = unit_geom(ws, u, tile_lin, n_cam_tiles, n_isects, a.tile_offsets, a.masks, DA < D, CH,
                                a.light_chunks);
        const TileGeom tg = tile_geom(tile_lin, a.tile_w, a.tile_h, a.tile_size, a.width, a.height);
        float T = 1.f;
        float acc[D];
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = 0.f;
        int32_t last = 0;
        const size_t pidx = cidx - MAX_BLOCK;  // chained state after the previous unit
        if (redo) {
            T = fabsf(ws.chain_T[pidx]);
            last = ws.chain_last[pidx];
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = ws.prefix_C[pidx * D + c];
        }
        bool frozen = !redo;
        const int u_last = ws.unit_start[tile_lin + 1] - 1;
        for (int uu = u; uu <= u_last; ++uu) {
            walk_chunks<D>(s, ld, ug, tg, false, T, acc, last, frozen);
            const size_t sidx = (size_t)uu * MAX_BLOCK + tg.tr;
            if (redo) {
                // finished here, or still running at the end of this unit: either way this is the pixel's chained state
                ws.chain_T[sidx] = frozen ? -T : T;
                ws.chain_last[sidx] = last;
#pragma unroll
                for (int c = 0; c < D; ++c) ws.prefix_C[sidx * D + c] = acc[c];
            }
            if (__syncthreads_and(frozen) || uu == u_last) break;
            // (rare) some marked pixel ran through the unit without stopping: continue with the next unit
            ug.k += 1;
            ug.c0 += HEAVY_CHUNKS;
            ug.c1 = min(ug.c0 + HEAVY_CHUNKS, (int)((ug.re - ug.rb + CH - 1) / CH));
        }
        if (!redo) continue;
        const size_t lidx = (size_t)u_last * MAX_BLOCK + tg.tr;
#pragma unroll
        for (int c = 0; c < D; ++c) ws.prefix_C[lidx * D + c] = acc[c];
        const int64_t pix = ((int64_t)tg.cam * a.height + tg.i) * a.width + tg.j;
        write_pixel<D, DA>(a, o, tg, pix, ug.masked, T, acc, last, 0);
    }
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

// The same reduce-scatter through shared memory: every lane stores its N values (row k = value k, 32 columns), then
// lanes 2k and 2k + 1 each add 16 columns of row k and exchange the halves, so BOTH hold the warp total of value k
// (lanes >= 2 N hold 0).  N stores + four 16-byte loads + 17 adds + one shuffle per lane, against 16 shuffles + 30
// selects + 18 adds of the butterfly at N = 15 — the butterfly was 43 % of the two-pixel backward's instructions
// (profiles/r02j_misc_cfg4_ncu_summary.txt).  Columns are XOR-swizzled in groups of four by (k & 3) so that the eight
// lanes of a 16-byte load phase hit eight different bank groups; `red` = this warp's N x 32 floats.
template <int N>
__device__ __forceinline__ float warp_smem_sum(const float (&v)[N], int lane, float* __restrict__ red) {
    static_assert(N <= 16, "two lanes per value");
#pragma unroll
    for (int k = 0; k < N; ++k) red[k * 32 + (lane ^ ((k & 3) << 2))] = v[k];
    __syncwarp();
    const int k = lane >> 1;
    float tot = 0.f;
    if (k < N) {
        const int at = k * 32 + ((lane & 1) << 4) + ((k & 3) << 2);
        const float4 q0 = *reinterpret_cast<const float4*>(red + at);
        const float4 q1 = *reinterpret_cast<const float4*>(red + (at ^ 4));
        const float4 q2 = *reinterpret_cast<const float4*>(red + (at ^ 8));
        const float4 q3 = *reinterpret_cast<const float4*>(red + (at ^ 12));
        tot = (((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w))) +
              (((q2.x + q2.y) + (q2.z + q2.w)) + ((q3.x + q3.y) + (q3.z + q3.w)));
    }
    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    __syncwarp();  // the next entry's stores must not overtake these loads
    return tot;
}

// dynamic shared memory of the backward kernels: the Stage, then n_warps x NV x 32 floats of reduction scratch
template <int D>
__host__ __device__ constexpr size_t bwd_red_offset() { return (sizeof(Stage<D>) + 127) / 128 * 128; }
__host__ __device__ constexpr size_t bwd_red_bytes(int nv, int n_warps) { return (size_t)n_warps * nv * 32 * sizeof(float); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct BwdIn {
    const float* render_a;        // [C, H, W, DA] (only the ED channel is read)
    const float* render_alphas;   // [C, H, W]
    const int32_t* last_ids;      // [C, H, W]
    const float* v_render_a;      // [C, H, W, DA]
    const float* v_render_b;      // [C, H, W, DB]
    const float* v_render_alphas; // [C, H, W]
};
struct BwdOut {
    float* v_means2d_abs;  // nullable
    float* v_means2d;      // nullable
    float* v_conics;
    float* v_colors_a;
    float* v_colors_b;
    float* v_opacities;
};

// Slots of the packed gradient vector: [0, D) colours, then conic a b c, opacity, xy, |xy|.
// XYMODE: 0 = no gradient for the 2-D means, 1 = xy, 2 = xy and |xy| (absgrad).  The xy gradient takes dL/dalpha of
// channel group A only (group B = the legacy normals pass, whose 2-D means are detached, dn_model.py:638).
template <int D, int DA, int XYMODE>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_bwd_kernel(RasterArgs a, Workspace ws, BwdIn in, BwdOut out) {
    constexpr int CH = chunk_len(D), DP = color_stride(D), DB = D - DA;
    constexpr bool SPLIT = (DB > 0);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage<D>& s = *reinterpret_cast<Stage<D>*>(smem_raw);
    __shared__ int32_t s_wmax[MAX_BLOCK / 32];
    const int64_t n_isects = fsb_eff_n(a.n_isects, a.n_dev);
    if ((int)blockIdx.x >= ws.hdr->total_units) return;
    const int u = ws.unit_order[blockIdx.x];
    const int64_t tile_lin = ws.unit_tile[u];
    if (a.masks != nullptr && !a.masks[tile_lin]) return;
    const UnitGeom ug = unit_geom(ws, u, tile_lin, (int64_t)a.C * a.tile_w * a.tile_h, n_isects, a.tile_offsets,
                                  a.masks, SPLIT, CH, a.light_chunks);
    if (ug.c0 >= ug.c1) return;
    const TileGeom tg = tile_geom(tile_lin, a.tile_w, a.tile_h, a.tile_size, a.width, a.height);
    const int64_t pix = tg.inside ? ((int64_t)tg.cam * a.height + tg.i) * a.width + tg.j : 0;
    // flagged tile: unit 0 differentiates group A over the unflagged entries, unit 1 group B over all entries
    const int which = ug.flagged ? ug.k + 1 : 0;
    const size_t cidx = (size_t)u * MAX_BLOCK + tg.tr;
    const int u_last = ws.unit_start[tile_lin + 1] - 1;

    const int32_t unit_b = ug.rb + ug.c0 * CH;
    int32_t bin_final = -1;
    if (tg.inside) bin_final = (which == 2) ? ws.chain_last[cidx] : in.last_ids[pix];
    int32_t warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));
    if (tg.lane == 0) s_wmax[tg.warp] = warp_bin_final;
    __syncthreads();
    int32_t cta_bin_final = -1;
    for (int w = 0; w < tg.n_warps; ++w) cta_bin_final = max(cta_bin_final, s_wmax[w]);
    if (cta_bin_final < unit_b) return;  // no pixel of the tile reaches into this unit

    float T_final = 1.f, v_ra = 0.f, T = 1.f;
    // The colour accumulated BEHIND the current entry enters dL/dalpha only through its dot product with the pixel's
    // cotangent, so one scalar per colour set replaces the D-vector gsplat carries:
    //   dL/dalpha = T (c . v) - (1 / (1 - alpha)) S,   S <- S + alpha T (c . v)        (c . v = sum_k colour_k v_rc_k)
    float v_rc[D];
    float S_a = 0.f, S_b = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) v_rc[c] = 0.f;
    if (tg.inside) {
        const float alpha_out = in.render_alphas[pix];
        // state at the END of this unit, left behind by the forward pass
        T = fabsf(ws.chain_T[cidx]);
        T_final = ug.light ? T : 1.f - alpha_out;
        if (which != 2) {
            v_ra = in.v_render_alphas[pix];
#pragma unroll
            for (int c = 0; c < DA; ++c) v_rc[c] = in.v_render_a[pix * DA + c];
            if (a.ed_channel >= 0) {
                const float den = fmaxf(alpha_out, 1e-10f);
#pragma unroll
                for (int c = 0; c < DA; ++c) {
                    if (c == a.ed_channel) {
                        const float v_ed = v_rc[c];
                        v_rc[c] = v_ed / den;
                        if (alpha_out >= 1e-10f) v_ra += -v_ed * in.render_a[pix * DA + c] / den;
                    }
                }
            }
        }
        if constexpr (SPLIT) {
            if (which != 1) {
#pragma unroll
                for (int c = 0; c < DB; ++c) v_rc[DA + c] = in.v_render_b[pix * DB + c];
            }
        }
        if (!ug.light) {
            const size_t lidx = (size_t)u_last * MAX_BLOCK + tg.tr;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float behind = ws.prefix_C[lidx * D + c] - ws.prefix_C[cidx * D + c];
                if (c < DA) S_a = fmaf(behind, v_rc[c], S_a); else S_b = fmaf(behind, v_rc[c], S_b);
            }
        }
    }
    // out = acc + T_final bg  and  alpha_out = 1 - T_final:  dL/dT_final = bg . v_out - v_alpha_out, folded into K
    float bg_dot_a = 0.f, bg_dot_b = 0.f;
    if (a.backgrounds_a) {
#pragma unroll
        for (int c = 0; c < DA; ++c) bg_dot_a += a.backgrounds_a[tg.cam * DA + c] * v_rc[c];
    }
    if constexpr (SPLIT) {
        if (a.backgrounds_b) {
#pragma unroll
            for (int c = 0; c < DB; ++c) bg_dot_b += a.backgrounds_b[tg.cam * DB + c] * v_rc[DA + c];
        }
    }
    const float K_a = T_final * (v_ra - bg_dot_a);
    const float K_b = -T_final * bg_dot_b;

    constexpr bool want_xy = (XYMODE >= 1);
    constexpr bool want_abs = (XYMODE >= 2);
    constexpr bool kTranspose = (D <= 8);
    constexpr int NG = 4 + 2 * XYMODE;             // conic a b c, opacity (, xy (, |xy|))
    constexpr int NV = kTranspose ? D + NG : NG;   // values reduced by the transposing butterfly
    constexpr int B = kTranspose ? D : 0;          // first geometric slot
    // (the shared-memory reduce-scatter of the two-pixel kernel loses here: 0.216 against 0.198 ms at cfg2, r02l — half
    // the values per visit are live and eight warps' scratch costs occupancy)
    const int slot = slot_of_lane<NV>(tg.lane);
    float* slot_base = nullptr;
    int slot_stride = 0;
    if (slot >= 0) {
        if (slot < B) {
            if (slot < DA) { slot_base = out.v_colors_a + slot; slot_stride = DA; }
            else { slot_base = out.v_colors_b + (slot - DA); slot_stride = DB; }
        }
        else if (slot < B + 3) { slot_base = out.v_conics + (slot - B); slot_stride = 3; }
        else if (slot == B + 3) { slot_base = out.v_opacities; slot_stride = 1; }
        else if (slot < B + 6) { slot_base = out.v_means2d + (slot - B - 4); slot_stride = 2; }
        else { slot_base = out.v_means2d_abs + (slot - B - 6); slot_stride = 2; }
    }
    const bool opac_slot = (slot == B + 3);
    const bool skip_flagged = (which == 1);

    Loader<D> ld(s, ws);
    ld.init(tg.tr);
    // entries behind every pixel's last id are not even loaded
    const int c_hi = min(ug.c1 - 1, (int)((cta_bin_final - ug.rb) / CH));
    auto chunk_n = [&](int c) { return min(min(CH, (int)(ug.re - ug.rb) - c * CH), (int)(cta_bin_final - (ug.rb + c * CH) + 1)); };
    if (tg.tr == 0) ld.issue(0, ug.rb + c_hi * CH, chunk_n(c_hi));
    for (int c = c_hi; c >= ug.c0; --c) {
        const int b = (c_hi - c) & 1;
        const bool more = c - 1 >= ug.c0;
        if (more && tg.tr == 0) ld.issue(b ^ 1, ug.rb + (c - 1) * CH, chunk_n(c - 1));
        const int n = chunk_n(c);
        ld.wait(b, n);
        const int cnt = build_list<D>(s, b, n, tg, skip_flagged);
        const int32_t chunk_b = ug.rb + c * CH;
        const uint16_t* wl = s.wlist[tg.warp];
        const float4* sgeo = s.geo[b];
        const float4* scon = s.con[b];
        const float* scol = s.col[b];
        const int t_hi = warp_bin_final - chunk_b;  // last entry of this chunk the warp can need
        for (int i = cnt - 1; i >= 0; --i) {
            const int t = wl[i];
            if (t > t_hi) continue;  // warp-uniform
            const float4 con = scon[t];
            float dx, dy, p, au;
            const float alpha = eval_alpha(sgeo[t], con, tg.px, tg.py, dx, dy, p, au);
            const bool valid = tg.inside && (chunk_b + t <= bin_final) && (p <= 0.f) && (alpha >= ALPHA_MIN);
            if (!__any_sync(0xffffffffu, valid)) continue;

            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = 0.f;
            float v_colD[D > 8 ? D : 1];
            if (valid) {
                const float ra = fast_rcp(1.f - alpha);
                T *= ra;
                const float fac = alpha * T;
                const float* cp = scol + t * DP;
                float cv_a = 0.f, cv_b = 0.f;  // colour . cotangent per colour set
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float ck = cp[k];
                    if constexpr (kTranspose) v[k] = fac * v_rc[k];
                    else v_colD[k] = fac * v_rc[k];
                    if (k < DA) cv_a = fmaf(ck, v_rc[k], cv_a); else cv_b = fmaf(ck, v_rc[k], cv_b);
                }
                // dL/dalpha through group A / group B
                const float va_a = fmaf(T, cv_a, (K_a - S_a) * ra);
                const float va_b = SPLIT ? fmaf(T, cv_b, (K_b - S_b) * ra) : 0.f;
                S_a = fmaf(fac, cv_a, S_a);
                if constexpr (SPLIT) S_b = fmaf(fac, cv_b, S_b);
                if (au <= ALPHA_MAX) {
                    // au = opacity * vis.  v_sigma = -au v_alpha; the opacity gradient vis v_alpha = (au / opacity)
                    // v_alpha is reduced as au v_alpha and divided by the opacity once, by the lane that owns the total.
                    const float nvs_a = au * va_a;
                    const float nvs = SPLIT ? au * (va_a + va_b) : nvs_a;
                    const float hs = -0.5f * nvs;
                    v[B + 0] = hs * dx * dx;
                    v[B + 1] = -nvs * dx * dy;
                    v[B + 2] = hs * dy * dy;
                    v[B + 3] = nvs;
                    if constexpr (want_xy) {
                        // conic a = -2 a' / log2e etc.:  v_sigma (a dx + b dy) = (nvs / log2e) (2 a' dx + b' dy)
                        const float vk = nvs_a * (1.f / LOG2E);
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
                for (int k = 0; k < D; ++k) v_colD[k] = 0.f;
            }
            const int32_t g = __float_as_int(con.w) & ~LEGACY_FLAG;
            if constexpr (!kTranspose) {
                // wide colour vectors: colours by plain butterflies, the geometric values transposed
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float tot = warp_sum(v_colD[k]);
                    if (tg.lane == 0) atomicAdd(out.v_colors_a + (size_t)g * D + k, tot);
                }
            }
            float total = warp_transpose_sum(v, tg.lane);
            if (opac_slot) total *= fast_ex2(-sgeo[t].z);  // 1 / opacity
            if (slot_base != nullptr && total != 0.f) atomicAdd(slot_base + (size_t)g * slot_stride, total);
        }
        __syncthreads();  // this chunk's reads are over before the next copy into the same buffer is issued
    }
}

// ---- backward, two pixels per lane ---------------------------------------------------------------------------------------
// Same arithmetic as raster_bwd_kernel; a warp owns an 8 x 8 pixel block (lane = 8 row + column over the upper four rows,
// its second pixel four rows below), four warps per 16 x 16 tile.  The per-(warp, entry) costs that do not depend on the
// number of pixels — list walk, record loads, the transposing butterfly and its atomics: about half of the 1-pixel
// kernel's instructions at 19 live lanes of 32 (profiles/r02d_ncu_raster_cfg4.txt) — are paid once for 64 pixels; each
// lane adds its two pixels' partials before the butterfly.
struct Px2 {
    float T, K_a, K_b, py;
    int32_t bin_final;
    bool inside;
};

// thread index the FORWARD gave the pixel (ti, tj) of a tile: the per-unit state is stored in that order
__device__ __forceinline__ int fwd_thread_of(int ti, int tj, int warps_x) {
    const int wy = ti >> 2, wx = tj >> 3;
    const int q = (((tj & 7) >> 2) << 1) | ((ti & 3) >> 1);
    const int s_ = ((ti & 1) << 2) | (tj & 3);
    return ((wy * warps_x + wx) << 5) | (q << 3) | s_;
}

// SMEM_RED: warp_smem_sum instead of the butterfly; its scratch follows the Stage in dynamic shared memory.
template <int D, int DA, int XYMODE, bool SMEM_RED>
__global__ void __launch_bounds__(MAX_BLOCK / 2)
raster_bwd2_kernel(RasterArgs a, Workspace ws, BwdIn in, BwdOut out) {
    constexpr int CH = chunk_len(D), DP = color_stride(D), DB = D - DA;
    constexpr bool SPLIT = (DB > 0);
    static_assert(D <= 8, "two-pixel backward: reduce-scatter forms only");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage<D>& s = *reinterpret_cast<Stage<D>*>(smem_raw);
    float* const red_all = reinterpret_cast<float*>(smem_raw + bwd_red_offset<D>());
    __shared__ int32_t s_wmax[MAX_BLOCK / 64];
    const int64_t n_isects = fsb_eff_n(a.n_isects, a.n_dev);
    if ((int)blockIdx.x >= ws.hdr->total_units) return;
    const int u = ws.unit_order[blockIdx.x];
    const int64_t tile_lin = ws.unit_tile[u];
    if (a.masks != nullptr && !a.masks[tile_lin]) return;
    const UnitGeom ug = unit_geom(ws, u, tile_lin, (int64_t)a.C * a.tile_w * a.tile_h, n_isects, a.tile_offsets,
                                  a.masks, SPLIT, CH, a.light_chunks);
    if (ug.c0 >= ug.c1) return;
    // geometry: warp w = 8 x 8 block (bx, by) of the tile
    const int tr = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tr & 31, warp = tr >> 5;
    const int n_warps = (blockDim.x * blockDim.y) >> 5;
    const int blocks_x = a.tile_size >> 3, warps_x = blocks_x;  // forward warps across = 8-pixel columns across
    const int by = warp / blocks_x, bx = warp - by * blocks_x;
    const int n_tiles = a.tile_w * a.tile_h;
    const int cam = (int)(tile_lin / n_tiles);
    const int tile_id = (int)(tile_lin - (int64_t)cam * n_tiles);
    const int tile_y = tile_id / a.tile_w, tile_x = tile_id - tile_y * a.tile_w;
    const int tj = bx * 8 + (lane & 7);
    const int ti0 = by * 8 + (lane >> 3);
    const float px = (float)(tile_x * a.tile_size + tj) + 0.5f;
    // forward warps covered by this block: (wy = 2 by, wx = bx) and (wy = 2 by + 1, wx = bx)
    const uint32_t wbits = (0xfu << (4 * ((2 * by) * warps_x + bx))) | (0xfu << (4 * ((2 * by + 1) * warps_x + bx)));
    const int which = ug.flagged ? ug.k + 1 : 0;
    const int u_last = ws.unit_start[tile_lin + 1] - 1;
    const int32_t unit_b = ug.rb + ug.c0 * CH;

    Px2 P[2];
    float v_rc[2][D], S_a[2], S_b[2];  // S: colour behind the current entry . cotangent (see raster_bwd_kernel)
    size_t cidx[2];
    int64_t pix[2];
    int32_t warp_bin_final = -1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ti = ti0 + 4 * h;
        const int gi = tile_y * a.tile_size + ti, gj = tile_x * a.tile_size + tj;
        P[h].inside = (gi < a.height && gj < a.width);
        P[h].py = (float)gi + 0.5f;
        pix[h] = P[h].inside ? ((int64_t)cam * a.height + gi) * a.width + gj : 0;
        cidx[h] = (size_t)u * MAX_BLOCK + fwd_thread_of(ti, tj, warps_x);
        P[h].bin_final = -1;
        if (P[h].inside) P[h].bin_final = (which == 2) ? ws.chain_last[cidx[h]] : in.last_ids[pix[h]];
        warp_bin_final = max(warp_bin_final, P[h].bin_final);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));
    if (lane == 0) s_wmax[warp] = warp_bin_final;
    __syncthreads();
    int32_t cta_bin_final = -1;
    for (int w = 0; w < n_warps; ++w) cta_bin_final = max(cta_bin_final, s_wmax[w]);
    if (cta_bin_final < unit_b) return;  // no pixel of the tile reaches into this unit

#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float T_final = 1.f, v_ra = 0.f;
        P[h].T = 1.f;
#pragma unroll
        for (int c = 0; c < D; ++c) v_rc[h][c] = 0.f;
        S_a[h] = S_b[h] = 0.f;
        if (P[h].inside) {
            const float alpha_out = in.render_alphas[pix[h]];
            P[h].T = fabsf(ws.chain_T[cidx[h]]);  // state at the END of this unit, left behind by the forward pass
            T_final = ug.light ? P[h].T : 1.f - alpha_out;
            if (which != 2) {
                v_ra = in.v_render_alphas[pix[h]];
#pragma unroll
                for (int c = 0; c < DA; ++c) v_rc[h][c] = in.v_render_a[pix[h] * DA + c];
                if (a.ed_channel >= 0) {
                    const float den = fmaxf(alpha_out, 1e-10f);
#pragma unroll
                    for (int c = 0; c < DA; ++c) {
                        if (c == a.ed_channel) {
                            const float v_ed = v_rc[h][c];
                            v_rc[h][c] = v_ed / den;
                            if (alpha_out >= 1e-10f) v_ra += -v_ed * in.render_a[pix[h] * DA + c] / den;
                        }
                    }
                }
            }
            if constexpr (SPLIT) {
                if (which != 1) {
#pragma unroll
                    for (int c = 0; c < DB; ++c) v_rc[h][DA + c] = in.v_render_b[pix[h] * DB + c];
                }
            }
            if (!ug.light) {
                const size_t lidx = (size_t)u_last * MAX_BLOCK + (cidx[h] - (size_t)u * MAX_BLOCK);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const float behind = ws.prefix_C[lidx * D + c] - ws.prefix_C[cidx[h] * D + c];
                    if (c < DA) S_a[h] = fmaf(behind, v_rc[h][c], S_a[h]); else S_b[h] = fmaf(behind, v_rc[h][c], S_b[h]);
                }
            }
        }
        float bg_dot_a = 0.f, bg_dot_b = 0.f;
        if (a.backgrounds_a) {
#pragma unroll
            for (int c = 0; c < DA; ++c) bg_dot_a += a.backgrounds_a[cam * DA + c] * v_rc[h][c];
        }
        if constexpr (SPLIT) {
            if (a.backgrounds_b) {
#pragma unroll
                for (int c = 0; c < DB; ++c) bg_dot_b += a.backgrounds_b[cam * DB + c] * v_rc[h][DA + c];
            }
        }
        P[h].K_a = T_final * (v_ra - bg_dot_a);
        P[h].K_b = -T_final * bg_dot_b;
    }

    constexpr bool want_xy = (XYMODE >= 1);
    constexpr bool want_abs = (XYMODE >= 2);
    constexpr int NG = 4 + 2 * XYMODE;
    constexpr int NV = D + NG;
    constexpr int B = D;
    // value index whose warp total this lane adds to global memory (-1: none)
    const int slot = SMEM_RED ? (((lane & 1) == 0 && (lane >> 1) < NV) ? (lane >> 1) : -1) : slot_of_lane<NV>(lane);
    float* const red = red_all + warp * (NV * 32);
    float* slot_base = nullptr;
    int slot_stride = 0;
    if (slot >= 0) {
        if (slot < B) {
            if (slot < DA) { slot_base = out.v_colors_a + slot; slot_stride = DA; }
            else { slot_base = out.v_colors_b + (slot - DA); slot_stride = DB; }
        }
        else if (slot < B + 3) { slot_base = out.v_conics + (slot - B); slot_stride = 3; }
        else if (slot == B + 3) { slot_base = out.v_opacities; slot_stride = 1; }
        else if (slot < B + 6) { slot_base = out.v_means2d + (slot - B - 4); slot_stride = 2; }
        else { slot_base = out.v_means2d_abs + (slot - B - 6); slot_stride = 2; }
    }
    const bool opac_slot = (slot == B + 3);
    const bool skip_flagged = (which == 1);

    Loader<D> ld(s, ws);
    ld.init(tr);
    const int c_hi = min(ug.c1 - 1, (int)((cta_bin_final - ug.rb) / CH));
    auto chunk_n = [&](int c) { return min(min(CH, (int)(ug.re - ug.rb) - c * CH), (int)(cta_bin_final - (ug.rb + c * CH) + 1)); };
    if (tr == 0) ld.issue(0, ug.rb + c_hi * CH, chunk_n(c_hi));
    for (int c = c_hi; c >= ug.c0; --c) {
        const int b = (c_hi - c) & 1;
        const bool more = c - 1 >= ug.c0;
        if (more && tr == 0) ld.issue(b ^ 1, ug.rb + (c - 1) * CH, chunk_n(c - 1));
        const int n = chunk_n(c);
        ld.wait(b, n);
        const int cnt = build_list_bits<D>(s, b, n, warp, lane, wbits, skip_flagged);
        const int32_t chunk_b = ug.rb + c * CH;
        const uint16_t* wl = s.wlist[warp];
        const float4* sgeo = s.geo[b];
        const float4* scon = s.con[b];
        const float* scol = s.col[b];
        const int t_hi = warp_bin_final - chunk_b;
        for (int i = cnt - 1; i >= 0; --i) {
            const int t = wl[i];
            if (t > t_hi) continue;  // warp-uniform
            const float4 con = scon[t];
            const float4 geo = sgeo[t];
            float dx, dy[2], p[2], au[2], alpha[2];
            bool valid[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                alpha[h] = eval_alpha(geo, con, px, P[h].py, dx, dy[h], p[h], au[h]);
                valid[h] = P[h].inside && (chunk_b + t <= P[h].bin_final) && (p[h] <= 0.f) && (alpha[h] >= ALPHA_MIN);
            }
            if (!__any_sync(0xffffffffu, valid[0] || valid[1])) continue;
            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = 0.f;
            float col[DP];
            {
                const float4* cp = reinterpret_cast<const float4*>(scol + t * DP);
#pragma unroll
                for (int k4 = 0; k4 < DP / 4; ++k4) {
                    const float4 c4 = cp[k4];
                    col[4 * k4] = c4.x; col[4 * k4 + 1] = c4.y; col[4 * k4 + 2] = c4.z; col[4 * k4 + 3] = c4.w;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (valid[h]) {
                    const float ra = fast_rcp(1.f - alpha[h]);
                    P[h].T *= ra;
                    const float T = P[h].T;
                    const float fac = alpha[h] * T;
                    float cv_a = 0.f, cv_b = 0.f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        v[k] = fmaf(fac, v_rc[h][k], v[k]);
                        if (k < DA) cv_a = fmaf(col[k], v_rc[h][k], cv_a); else cv_b = fmaf(col[k], v_rc[h][k], cv_b);
                    }
                    const float va_a = fmaf(T, cv_a, (P[h].K_a - S_a[h]) * ra);
                    const float va_b = SPLIT ? fmaf(T, cv_b, (P[h].K_b - S_b[h]) * ra) : 0.f;
                    S_a[h] = fmaf(fac, cv_a, S_a[h]);
                    if constexpr (SPLIT) S_b[h] = fmaf(fac, cv_b, S_b[h]);
                    if (au[h] <= ALPHA_MAX) {
                        const float nvs_a = au[h] * va_a;
                        const float nvs = SPLIT ? au[h] * (va_a + va_b) : nvs_a;
                        const float hs = -0.5f * nvs;
                        v[B + 0] = fmaf(hs * dx, dx, v[B + 0]);
                        v[B + 1] = fmaf(-nvs * dx, dy[h], v[B + 1]);
                        v[B + 2] = fmaf(hs * dy[h], dy[h], v[B + 2]);
                        v[B + 3] += nvs;
                        if constexpr (want_xy) {
                            const float vk = nvs_a * (1.f / LOG2E);
                            const float gx = vk * fmaf(2.f * con.x, dx, con.y * dy[h]);
                            const float gy = vk * fmaf(2.f * con.z, dy[h], con.y * dx);
                            v[B + 4] += gx;
                            v[B + 5] += gy;
                            if constexpr (want_abs) {
                                v[B + 6] += fabsf(gx);
                                v[B + 7] += fabsf(gy);
                            }
                        }
                    }
                }
            }
            const int32_t g = __float_as_int(con.w) & ~LEGACY_FLAG;
            float total = SMEM_RED ? warp_smem_sum(v, lane, red) : warp_transpose_sum(v, lane);
            if (opac_slot) total *= fast_ex2(-geo.z);  // 1 / opacity
            if (slot_base != nullptr && total != 0.f) atomicAdd(slot_base + (size_t)g * slot_stride, total);
        }
        __syncthreads();
    }
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

struct FwdCall {
    RasterArgs a;
    PackIn pack;
    void* workspace;
    FwdOut o;
};

template <int D, int DA>
int launch_fwd(const FwdCall& f, cudaStream_t st) {
    const RasterArgs& a = f.a;
    const int64_t n_tiles = (int64_t)a.C * a.tile_w * a.tile_h;
    Workspace ws;
    ws_layout<true>(f.workspace, a.n_isects, n_tiles, (int64_t)a.C * a.N, D, &ws);
    constexpr bool SPLIT = (DA < D);
    if (SPLIT) {
        FSB_CUDA(cudaMemsetAsync(ws.tile_flag, 0, (size_t)n_tiles, st));
        if (a.n_isects > 0) {
            int blocks = fsb_div_up(a.n_isects, 256 * 4);
            if (blocks > FSB_NUM_SMS * 16) blocks = FSB_NUM_SMS * 16;
            tile_flag_kernel<<<blocks, 256, 0, st>>>(a.n_isects, a.n_dev, f.pack.flatten_ids, a.tile_offsets, n_tiles,
                                                     ws.tile_flag);
            FSB_LAUNCH_CHECK();
        }
    }
    unit_table_kernel<<<1, 1024, 0, st>>>((int)n_tiles, a.n_isects, a.n_dev, chunk_len(D), a.light_chunks,
                                          a.tile_offsets, a.masks, SPLIT, ws);
    FSB_LAUNCH_CHECK();
    const unsigned grid_units = (unsigned)max_units(a.n_isects, n_tiles, D);
    if (a.n_isects > 0) {
        // records pay when a Gaussian has several list entries (7 at 1080p / 1M: pack 236 -> 164 + 23 us); with about
        // one entry per Gaussian (cfg2's object scene) the extra pass only costs.  In static-capacity mode n_isects is
        // the capacity, ~1.4 x the count.
        const int64_t n_gauss = (int64_t)a.C * a.N;
        if (a.n_isects >= 3 * n_gauss) {
            raster_prepack_kernel<D, DA><<<fsb_div_up(n_gauss, 256), 256, 0, st>>>(n_gauss, f.pack, ws.rec);
            FSB_LAUNCH_CHECK();
            raster_pack_kernel<D, DA, true><<<grid_units, 256, 0, st>>>(a.C, a.n_isects, a.n_dev, f.pack, a.masks,
                                                                       a.tile_size, a.tile_w, a.tile_h, a.light_chunks,
                                                                       a.tile_offsets, ws);
        } else {
            raster_pack_kernel<D, DA, false><<<grid_units, 256, 0, st>>>(a.C, a.n_isects, a.n_dev, f.pack, a.masks,
                                                                        a.tile_size, a.tile_w, a.tile_h, a.light_chunks,
                                                                        a.tile_offsets, ws);
        }
        FSB_LAUNCH_CHECK();
    }
    dim3 block(a.tile_size, a.tile_size);
    const size_t smem = sizeof(Stage<D>);
    // a tile can only be segment-parallel when its list exceeds LIGHT_CHUNKS chunks
    const bool multi = a.n_dev != nullptr || a.n_isects > (int64_t)a.light_chunks * chunk_len(D);
    if (multi) FSB_CUDA(cudaMemsetAsync(ws.done_k, 0x7f, (size_t)n_tiles * 8 * 4, st));
    {
        auto k = raster_fwd_kernel<D, DA, 0>;
        int e = set_smem(k, smem); if (e) return e;
        k<<<(unsigned)n_tiles, block, smem, st>>>(a, ws, f.o);
        FSB_LAUNCH_CHECK();
    }
    if (multi) {
        const unsigned grid = FSB_NUM_SMS * 4;
        auto k1 = raster_fwd_kernel<D, DA, 1>;
        int e = set_smem(k1, smem); if (e) return e;
        k1<<<grid, block, smem, st>>>(a, ws, f.o);
        FSB_LAUNCH_CHECK();
        raster_fold_kernel<D, DA><<<grid, block, 0, st>>>(a, ws, f.o);
        FSB_LAUNCH_CHECK();
        auto k2 = raster_stop_kernel<D, DA>;
        e = set_smem(k2, smem); if (e) return e;
        k2<<<grid, block, smem, st>>>(a, ws, f.o);
        FSB_LAUNCH_CHECK();
    }
    return 0;
}

struct BwdCall {
    RasterArgs a;
    void* workspace;
    BwdIn in;
    BwdOut out;
};

template <int D, int DA>
int launch_bwd(const BwdCall& f, cudaStream_t st) {
    const RasterArgs& a = f.a;
    const int64_t n_tiles = (int64_t)a.C * a.tile_w * a.tile_h;
    Workspace ws;
    ws_layout<true>(f.workspace, a.n_isects, n_tiles, (int64_t)a.C * a.N, D, &ws);
    dim3 block(a.tile_size, a.tile_size);
    const unsigned grid = (unsigned)max_units(a.n_isects, n_tiles, D);
    const size_t smem = sizeof(Stage<D>);
    // Two pixels per lane (8 x 8 block per warp) for the narrow colour vectors on frames with many tiles: r02h, cfg4
    // (1080p, blob-like splats, 19 of 32 lanes live): 1.06 ms against 1.30 ms.  On the 640 x 480 object scene (thin
    // surfels, 1200 tiles for 592 CTA slots) the finer 8 x 4 footprints and twice the warps win: 0.19 ms against
    // 0.23 ms.  FSB_RASTER_BWD_PX=1 / 2 forces a kernel (A/B runs).
    static const int forced_px = [] { const char* e = getenv("FSB_RASTER_BWD_PX"); return e ? atoi(e) : 0; }();
    const bool one_px = forced_px == 1 || (forced_px != 2 && n_tiles < 4 * FSB_NUM_SMS * 4);
    // FSB_RASTER_BWD_REDUCE=shfl: the butterfly instead of the shared-memory reduce-scatter (A/B runs)
    static const bool smem_red = [] { const char* e = getenv("FSB_RASTER_BWD_REDUCE"); return !(e && e[0] == 's' && e[1] == 'h'); }();
    dim3 block2(a.tile_size, a.tile_size / 2);
#define FSB_BWD_LAUNCH(MODE)                                                   \
    do {                                                                       \
        if constexpr (D <= 8) {                                                \
            if (!one_px && smem_red) {                                         \
                auto k2 = raster_bwd2_kernel<D, DA, MODE, true>;               \
                const size_t sm2 = bwd_red_offset<D>() + bwd_red_bytes(D + 4 + 2 * MODE, MAX_BLOCK / 64); \
                int e2 = set_smem(k2, sm2); if (e2) return e2;                 \
                k2<<<grid, block2, sm2, st>>>(a, ws, f.in, f.out);             \
                break;                                                         \
            }                                                                  \
            if (!one_px) {                                                     \
                auto k2 = raster_bwd2_kernel<D, DA, MODE, false>;              \
                int e2 = set_smem(k2, smem); if (e2) return e2;                \
                k2<<<grid, block2, smem, st>>>(a, ws, f.in, f.out);            \
                break;                                                         \
            }                                                                  \
        }                                                                      \
        auto k = raster_bwd_kernel<D, DA, MODE>;                               \
        int e = set_smem(k, smem); if (e) return e;                            \
        k<<<grid, block, smem, st>>>(a, ws, f.in, f.out);                      \
    } while (0)
    if (f.out.v_means2d_abs) FSB_BWD_LAUNCH(2);
    else if (f.out.v_means2d) FSB_BWD_LAUNCH(1);
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
        const int32_t raw = flatten_ids[e];
        if (raw < 0) continue;  // legacy-only entry: not part of the first colour set's list
        const int32_t g = raw;
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

// (D total, DA) pairs the kernels are instantiated for: single colour set D in {1,2,3,4,5,8,16,32}, and the
// DN-Splatter pair RGB + expected depth (4) | normals (3)
#define FSB_DISPATCH_D(DA_, DB_, CALL)                                          \
    if ((DB_) == 0) {                                                           \
        switch (DA_) {                                                          \
            case 1: { constexpr int DD = 1, DDA = 1; return CALL; }             \
            case 2: { constexpr int DD = 2, DDA = 2; return CALL; }             \
            case 3: { constexpr int DD = 3, DDA = 3; return CALL; }             \
            case 4: { constexpr int DD = 4, DDA = 4; return CALL; }             \
            case 5: { constexpr int DD = 5, DDA = 5; return CALL; }             \
            case 8: { constexpr int DD = 8, DDA = 8; return CALL; }             \
            case 16: { constexpr int DD = 16, DDA = 16; return CALL; }          \
            case 32: { constexpr int DD = 32, DDA = 32; return CALL; }          \
            default: return FSB_E_ARG;                                          \
        }                                                                       \
    } else if ((DA_) == 4 && (DB_) == 3) {                                      \
        constexpr int DD = 7, DDA = 4;                                          \
        return CALL;                                                            \
    } else {                                                                    \
        return FSB_E_ARG;                                                       \
    }

// channel counts the kernels are instantiated for; callers pad up to the next one
FSB_API int fsb_raster_supported_channels(int D) {
    const int s[] = {1, 2, 3, 4, 5, 8, 16, 32};
    for (int i = 0; i < 8; ++i)
        if (s[i] >= D) return s[i];
    return -1;
}

// bytes of the per-call workspace: unit table, the per-(unit, pixel) state that the forward leaves for the backward and
// the packed list records (n_tiles = C * tile_w * tile_h).  DB = 0: one colour set.
FSB_API size_t fsb_raster_dn_workspace(int64_t n_isects, int64_t n_tiles, int64_t n_gauss, int DA, int DB) {
    if (n_isects < 0 || n_tiles <= 0 || DA <= 0 || DB < 0) return 0;
    return ws_layout<false>(nullptr, n_isects, n_tiles, n_gauss, DA + DB, nullptr);
}
FSB_API size_t fsb_raster_workspace(int64_t n_isects, int64_t n_tiles, int64_t n_gauss, int D) {
    return fsb_raster_dn_workspace(n_isects, n_tiles, n_gauss, D, 0);
}

// Forward with two colour sets composited by one walk (see the file header).  DB = 0 (colors_b, backgrounds_b, out_b
// NULL) is the plain rasterizer.  flatten_ids may carry FSB_LEGACY_FLAG (bit 31) when DB > 0.
// n_isects_dev (nullable): static-capacity mode, n_isects is then the capacity of flatten_ids (common.cuh).
FSB_API int fsb_raster_dn_fwd(int C, int N, int DA, int DB, int64_t n_isects, const int64_t* n_isects_dev,
                              const float* means2d, const float* conics, const float* colors_a, const float* colors_b,
                              const float* opacities, const float* backgrounds_a, const float* backgrounds_b,
                              const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                              const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_channel, void* workspace,
                              size_t workspace_bytes, float* out_a, float* out_b, float* out_alphas, int32_t* last_ids,
                              void* stream) {
    if (bad_geometry(C, tile_size, n_isects)) return FSB_E_ARG;
    if (DB < 0 || (DB > 0 && (!colors_b || !out_b)) || ed_channel >= DA) return FSB_E_ARG;
    if (tile_w <= 0 || tile_h <= 0) return 0;
    if (!workspace || workspace_bytes < fsb_raster_dn_workspace(n_isects, (int64_t)C * tile_w * tile_h, (int64_t)C * N, DA, DB))
        return FSB_E_ARG;
    FwdCall f;
    f.a = RasterArgs{C, N, n_isects, n_isects_dev, masks, width, height, tile_size, tile_w, tile_h, tile_offsets,
                     backgrounds_a, backgrounds_b, ed_channel, light_chunks_for((int64_t)C * tile_w * tile_h)};
    f.pack = PackIn{(const float2*)means2d, conics, opacities, colors_a, colors_b, flatten_ids};
    f.workspace = workspace;
    f.o = FwdOut{out_a, out_b, out_alphas, last_ids};
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(DA, DB, (launch_fwd<DD, DDA>(f, st)));
}

// Gradient outputs are ACCUMULATED into (atomicAdd); the caller zero-fills them first.
// `workspace` is the buffer the matching forward call filled.
// v_means2d / v_means2d_abs may be NULL (no gradient wanted for the 2-D means); v_means2d_abs requires v_means2d.
FSB_API int fsb_raster_dn_bwd(int C, int N, int DA, int DB, int64_t n_isects, const int64_t* n_isects_dev,
                              const float* backgrounds_a, const float* backgrounds_b, const uint8_t* masks, int width,
                              int height, int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets,
                              int ed_channel, void* workspace, size_t workspace_bytes, const float* render_a,
                              const float* render_alphas, const int32_t* last_ids, const float* v_render_a,
                              const float* v_render_b, const float* v_render_alphas, float* v_means2d_abs,
                              float* v_means2d, float* v_conics, float* v_colors_a, float* v_colors_b,
                              float* v_opacities, void* stream) {
    if (bad_geometry(C, tile_size, n_isects)) return FSB_E_ARG;
    if (ed_channel >= 0 && !render_a) return FSB_E_ARG;
    if (v_means2d_abs && !v_means2d) return FSB_E_ARG;
    if (DB < 0 || (DB > 0 && (!v_render_b || !v_colors_b)) || ed_channel >= DA) return FSB_E_ARG;
    if (tile_w <= 0 || tile_h <= 0 || n_isects == 0) return 0;
    if (!workspace || workspace_bytes < fsb_raster_dn_workspace(n_isects, (int64_t)C * tile_w * tile_h, (int64_t)C * N, DA, DB))
        return FSB_E_ARG;
    BwdCall f;
    f.a = RasterArgs{C, N, n_isects, n_isects_dev, masks, width, height, tile_size, tile_w, tile_h, tile_offsets,
                     backgrounds_a, backgrounds_b, ed_channel, light_chunks_for((int64_t)C * tile_w * tile_h)};
    f.workspace = workspace;
    f.in = BwdIn{render_a, render_alphas, last_ids, v_render_a, v_render_b, v_render_alphas};
    f.out = BwdOut{v_means2d_abs, v_means2d, v_conics, v_colors_a, v_colors_b, v_opacities};
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(DA, DB, (launch_bwd<DD, DDA>(f, st)));
}

// One colour set (gsplat rasterize_to_pixels): the DB = 0 case of the calls above.
FSB_API int fsb_raster_fwd(int C, int N, int D, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* backgrounds,
                           const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                           const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_normalize,
                           void* workspace, size_t workspace_bytes, float* out_colors, float* out_alphas,
                           int32_t* last_ids, void* stream) {
    return fsb_raster_dn_fwd(C, N, D, 0, n_isects, n_isects_dev, means2d, conics, colors, nullptr, opacities,
                             backgrounds, nullptr, masks, width, height, tile_size, tile_w, tile_h, tile_offsets,
                             flatten_ids, ed_normalize ? D - 1 : -1, workspace, workspace_bytes, out_colors, nullptr,
                             out_alphas, last_ids, stream);
}

FSB_API int fsb_raster_bwd(int C, int N, int D, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* backgrounds,
                           const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                           const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_normalize,
                           void* workspace, size_t workspace_bytes, const float* render_colors,
                           const float* render_alphas, const int32_t* last_ids, const float* v_render_colors,
                           const float* v_render_alphas, float* v_means2d_abs, float* v_means2d, float* v_conics,
                           float* v_colors, float* v_opacities, void* stream) {
    (void)means2d; (void)conics; (void)colors; (void)opacities; (void)flatten_ids;  // packed by the forward
    if (ed_normalize && !render_colors) return FSB_E_ARG;
    return fsb_raster_dn_bwd(C, N, D, 0, n_isects, n_isects_dev, backgrounds, nullptr, masks, width, height, tile_size,
                             tile_w, tile_h, tile_offsets, ed_normalize ? D - 1 : -1, workspace, workspace_bytes,
                             render_colors, render_alphas, last_ids, v_render_colors, nullptr, v_render_alphas,
                             v_means2d_abs, v_means2d, v_conics, v_colors, nullptr, v_opacities, stream);
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
