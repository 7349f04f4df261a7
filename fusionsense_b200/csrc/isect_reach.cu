// isect_reach.cu — tile intersection with an exact reach test: count / emit only the (Gaussian, tile) pairs whose
// Gaussian can pass the alpha test (alpha >= 1/255) on at least one pixel centre of the tile.
//
// On by default in the DN-Splatter step (DNSplatterStepConfig.prune_lists).  The reference bins by the 3-sigma bounding box
// (gsplat isect_tiles, reached from /root/reference/dn_splatter/dn_model.py:570-591 and :644-653); projected
// surfels are thin rotated ellipses whose box is mostly empty: 35 % of the binned pairs of the 300 k bench scene and
// 49 % of cfg1's never reach a pixel (tools/raster_stats.py, profiles/r01_raster_stats_*.json, which also checks
// this very test — same float32 arithmetic — for false prunes: none).  Such a pair composites nothing in either
// direction, so lists without them give the same images and gradients while the sort, the range build and every
// staging round of the compositing kernels handle that many fewer entries.  fsb_isect_count / fsb_isect_emit keep
// producing the reference's full lists; these entry points are for callers that do not expose the lists.
#include "common.cuh"

namespace {

// Same test as csrc/raster.cu::strip_mask on one rectangle of pixel centres [x0, x1] x [y0, y1] (relative to
// nothing: absolute pixel coordinates): min over the rectangle of q(d) = a dx^2 + 2 b dx dy + c dy^2 against
// 2 ln(255 opacity), inflated; anything doubtful (non-PD conic, NaN) answers yes.
__device__ __forceinline__ bool rect_reach(float gx, float gy, float opac, float a, float b, float c, float px0,
                                           float px1, float py0, float py1) {
    const float tau = __logf(255.f * opac);
    if (tau + 2e-3f < 0.f) return false;  // opacity below 1/255: can never pass the alpha test
    const float det = a * c - b * b;
    if (!(det > 0.f) || !(a > 0.f) || !(tau < 1e30f)) return true;
    const float thr = 2.f * tau * 1.0002f + 1e-2f;
    const float x0 = px0 - gx, x1 = px1 - gx, y0 = py0 - gy, y1 = py1 - gy;
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return true;  // the mean lies inside
    const float nb_c = -b / c, nb_a = -b / a;
    float t, v, q;
    t = fminf(fmaxf(nb_c * x0, y0), y1); q = a * x0 * x0 + 2.f * b * x0 * t + c * t * t;
    t = fminf(fmaxf(nb_c * x1, y0), y1); v = a * x1 * x1 + 2.f * b * x1 * t + c * t * t; q = fminf(q, v);
    t = fminf(fmaxf(nb_a * y0, x0), x1); v = a * t * t + 2.f * b * t * y0 + c * y0 * y0; q = fminf(q, v);
    t = fminf(fmaxf(nb_a * y1, x0), x1); v = a * t * t + 2.f * b * t * y1 + c * y1 * y1; q = fminf(q, v);
    return !(q > thr);
}

struct Box {
    int ax, ay, bx, by;
};

// tile bounding box of a projected Gaussian: gsplat 1.0 rule (floor / ceil) or the 0.1.x rule of
// rasterize_gaussians ((int) truncation, +1 on the max side) — the same arithmetic as fsb_isect_emit
__device__ __forceinline__ Box tile_box(float mx, float my, int r, int tile_size, int tile_w, int tile_h,
                                        int legacy_bbox) {
    const float ts = (float)tile_size;
    const float tr = (float)r / ts;
    const float tx = mx / ts, ty = my / ts;
    Box b;
    if (legacy_bbox) {
        b.ax = (int)(tx - tr); b.ay = (int)(ty - tr);
        b.bx = (int)(tx + tr + 1.f); b.by = (int)(ty + tr + 1.f);
    } else {
        b.ax = (int)floorf(tx - tr); b.ay = (int)floorf(ty - tr);
        b.bx = (int)ceilf(tx + tr); b.by = (int)ceilf(ty + tr);
    }
    b.ax = min(max(0, b.ax), tile_w); b.ay = min(max(0, b.ay), tile_h);
    b.bx = min(max(0, b.bx), tile_w); b.by = min(max(0, b.by), tile_h);
    return b;
}

// COUNT_ONLY: counts[idx] = number of reached tiles; else emit keys / values for them.
// legacy_bbox: 0 = gsplat 1.0 box, 1 = 0.1.x box, 2 = UNION list for the two-colour-set compositing pass
// (csrc/raster.cu): the 0.1.x box (a superset of the 1.0 box), with FSB_LEGACY_FLAG set in the value of the tiles
// that only the 0.1.x rule yields.
// A Gaussian whose box holds more than WIDE_TILES tiles is handled by its whole warp (a wall close to the camera
// covers hundreds of tiles; one thread testing them serially was the long pole of the kernel on object scenes).
// The order of one Gaussian's entries among themselves is free: their keys differ (different tiles), and entries with
// equal keys (same tile, same depth, different Gaussians) keep the order of the offsets = Gaussian index, which is what
// the stable sort preserves.
constexpr int WIDE_TILES = 24;  // <= 64: a narrow box's reach results fit the 64-bit hit mask
constexpr int32_t LEGACY_FLAG = (int32_t)0x80000000;

struct GaussTile {
    float mx, my, a, b, c, o;
    Box box, inner;
    int64_t cam_part, depth_part, cur;
    int32_t idx;
};

// one list entry: key = (camera, tile) | depth bits with the value beside it, or (packed) key = (camera, tile) | value
__device__ __forceinline__ void put_entry(int64_t* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids, int64_t at,
                                          int64_t high, int64_t depth_part, int32_t val, int packed) {
    if (packed) {
        isect_ids[at] = high | (int64_t)(uint32_t)val;
    } else {
        isect_ids[at] = high | depth_part;
        flatten_ids[at] = val;
    }
}

template <bool COUNT_ONLY>
__device__ __forceinline__ int reach_one_tile(const GaussTile& g, int i, int j, int tile_size, int tile_w, bool flag_outer,
                                              int64_t limit, int64_t cur, int64_t* __restrict__ isect_ids,
                                              int32_t* __restrict__ flatten_ids, int packed) {
    // pixel centres of the tile; the part of an edge tile beyond the image only makes the test more generous
    const float px0 = (float)(j * tile_size) + 0.5f, py0 = (float)(i * tile_size) + 0.5f;
    if (!rect_reach(g.mx, g.my, g.o, g.a, g.b, g.c, px0, px0 + (float)(tile_size - 1), py0, py0 + (float)(tile_size - 1)))
        return 0;
    if (!COUNT_ONLY && cur < limit) {
        const bool outer = flag_outer && !(i >= g.inner.ay && i < g.inner.by && j >= g.inner.ax && j < g.inner.bx);
        put_entry(isect_ids, flatten_ids, cur, g.cam_part | ((int64_t)(i * tile_w + j) << 32), g.depth_part,
                  outer ? (g.idx | LEGACY_FLAG) : g.idx, packed);
    }
    return 1;
}

template <bool COUNT_ONLY>
__global__ void __launch_bounds__(256)
isect_reach_kernel(int C, int N, const float* __restrict__ means2d, const int32_t* __restrict__ radii,
                   const float* __restrict__ depths, const float* __restrict__ conics,
                   const float* __restrict__ opacities, const int64_t* __restrict__ offsets, int tile_size, int tile_w,
                   int tile_h, int tile_bits, int legacy_bbox, const int64_t* __restrict__ n_dev, int64_t capacity,
                   int32_t* __restrict__ overflow_flag, int32_t* __restrict__ counts, int64_t* __restrict__ isect_ids,
                   int32_t* __restrict__ flatten_ids, unsigned long long* __restrict__ hit_masks,
                   const int32_t* __restrict__ perm, int packed, unsigned long long* __restrict__ depth_keys,
                   int32_t* __restrict__ depth_vals) {
    // Two-level binning (ops.isect_tiles_depth_first):
    //   count pass: depth_keys / depth_vals (nullable) receive (depth bits, index) of every (camera, Gaussian) for the
    //               depth sort; a Gaussian without entries gets the largest key.
    //   emit pass : thread `gid` handles Gaussian perm[gid] (perm = the depth order) and `offsets` is indexed by gid;
    //               packed != 0: the low word of the key is the flatten id (flag included) instead of the depth bits and
    //               flatten_ids is not written — a stable sort on the (camera, tile) bits then finishes the order.
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in_range = gid < (int64_t)C * N;
    const int64_t idx = (in_range && perm != nullptr) ? (int64_t)perm[gid] : gid;
    if (!COUNT_ONLY) {
        if (idx == 0 && n_dev && overflow_flag && *n_dev > capacity) *overflow_flag = 1;
        if (n_dev && *n_dev == 0) return;
    }
    const int r = in_range ? radii[idx] : 0;
    GaussTile g;
    g.box = Box{0, 0, 0, 0};
    if (r > 0) {
        const float2 m = reinterpret_cast<const float2*>(means2d)[idx];
        g.mx = m.x; g.my = m.y;
        g.a = conics[3 * idx + 0]; g.b = conics[3 * idx + 1]; g.c = conics[3 * idx + 2];
        g.o = opacities[idx];
        g.box = tile_box(m.x, m.y, r, tile_size, tile_w, tile_h, legacy_bbox != 0);
        g.inner = (legacy_bbox == 2) ? tile_box(m.x, m.y, r, tile_size, tile_w, tile_h, 0) : g.box;
        g.idx = (int32_t)idx;
        g.cam_part = g.depth_part = g.cur = 0;
        if (!COUNT_ONLY) {
            g.cam_part = (idx / N) << (32 + tile_bits);
            g.depth_part = packed ? 0 : (int64_t)(uint32_t)__float_as_int(depths[idx]);
            g.cur = offsets[gid];
        }
    }
    const int64_t limit = n_dev ? capacity : INT64_MAX;
    const bool flag_outer = (legacy_bbox == 2);
    const int bw = g.box.bx - g.box.ax, bh = g.box.by - g.box.ay;
    const int n_box = (r > 0) ? bw * bh : 0;
    const bool wide = n_box > WIDE_TILES;
    int n = 0;
    if (r > 0 && !wide) {
        // hit_masks (optional, boxes of at most WIDE_TILES <= 64 tiles): the count pass leaves bit k = "tile k of the box,
        // row-major, is reached"; the emit pass reads it instead of repeating the reach tests
        int64_t cur = g.cur;
        if (!COUNT_ONLY && hit_masks != nullptr) {
            unsigned long long hm = hit_masks[idx];
            while (hm) {
                const int k = __ffsll((long long)hm) - 1;
                hm &= hm - 1;
                const int i = g.box.ay + k / bw, j = g.box.ax + k % bw;
                if (cur < limit) {
                    const bool outer = flag_outer && !(i >= g.inner.ay && i < g.inner.by && j >= g.inner.ax && j < g.inner.bx);
                    put_entry(isect_ids, flatten_ids, cur, g.cam_part | ((int64_t)(i * tile_w + j) << 32), g.depth_part,
                              outer ? (g.idx | LEGACY_FLAG) : g.idx, packed);
                }
                ++cur;
            }
        } else {
            unsigned long long hm = 0ull;
            int k = 0;
            for (int i = g.box.ay; i < g.box.by; ++i)
                for (int j = g.box.ax; j < g.box.bx; ++j, ++k) {
                    const int hit = reach_one_tile<COUNT_ONLY>(g, i, j, tile_size, tile_w, flag_outer, limit, cur,
                                                               isect_ids, flatten_ids, packed);
                    cur += hit;
                    n += hit;
                    hm |= (unsigned long long)hit << k;
                }
            if (COUNT_ONLY && hit_masks != nullptr) hit_masks[idx] = hm;
        }
    }
    // wide boxes: the warp takes them one at a time, 32 tiles per round
    unsigned todo = __ballot_sync(0xffffffffu, wide);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        GaussTile w;
        w.mx = __shfl_sync(0xffffffffu, g.mx, src); w.my = __shfl_sync(0xffffffffu, g.my, src);
        w.a = __shfl_sync(0xffffffffu, g.a, src); w.b = __shfl_sync(0xffffffffu, g.b, src);
        w.c = __shfl_sync(0xffffffffu, g.c, src); w.o = __shfl_sync(0xffffffffu, g.o, src);
        w.box.ax = __shfl_sync(0xffffffffu, g.box.ax, src); w.box.ay = __shfl_sync(0xffffffffu, g.box.ay, src);
        w.box.bx = __shfl_sync(0xffffffffu, g.box.bx, src); w.box.by = __shfl_sync(0xffffffffu, g.box.by, src);
        w.inner.ax = __shfl_sync(0xffffffffu, g.inner.ax, src); w.inner.ay = __shfl_sync(0xffffffffu, g.inner.ay, src);
        w.inner.bx = __shfl_sync(0xffffffffu, g.inner.bx, src); w.inner.by = __shfl_sync(0xffffffffu, g.inner.by, src);
        w.idx = __shfl_sync(0xffffffffu, g.idx, src);
        w.cam_part = __shfl_sync(0xffffffffu, g.cam_part, src);
        w.depth_part = __shfl_sync(0xffffffffu, g.depth_part, src);
        int64_t cur = __shfl_sync(0xffffffffu, g.cur, src);
        const int wbw = w.box.bx - w.box.ax;
        const int total = wbw * (w.box.by - w.box.ay);
        int found = 0;
        for (int base = 0; base < total; base += 32) {
            const int k = base + lane;
            int hit = 0;
            int i = 0, j = 0;
            bool reach = false;
            if (k < total) {
                i = w.box.ay + k / wbw;
                j = w.box.ax + k % wbw;
                const float px0 = (float)(j * tile_size) + 0.5f, py0 = (float)(i * tile_size) + 0.5f;
                reach = rect_reach(w.mx, w.my, w.o, w.a, w.b, w.c, px0, px0 + (float)(tile_size - 1), py0,
                                   py0 + (float)(tile_size - 1));
            }
            const unsigned hits = __ballot_sync(0xffffffffu, reach);
            if (!COUNT_ONLY && reach) {
                const int64_t at = cur + __popc(hits & ((1u << lane) - 1u));
                if (at < limit) {
                    const bool outer = flag_outer && !(i >= w.inner.ay && i < w.inner.by && j >= w.inner.ax && j < w.inner.bx);
                    put_entry(isect_ids, flatten_ids, at, w.cam_part | ((int64_t)(i * tile_w + j) << 32), w.depth_part,
                              outer ? (w.idx | LEGACY_FLAG) : w.idx, packed);
                }
            }
            (void)hit;
            cur += __popc(hits);
            found += __popc(hits);
        }
        if (lane == src) n = found;
    }
    if (COUNT_ONLY && in_range) {
        counts[idx] = n;
        if (depth_keys != nullptr) {
            depth_keys[idx] = n > 0 ? (unsigned long long)(uint32_t)__float_as_int(depths[idx]) : 0xffffffffull;
            depth_vals[idx] = (int32_t)idx;
        }
    }
}

}  // namespace

// counts[C*N] = tiles of the bounding box (legacy_bbox: 0.1.x rule) that the Gaussian can reach with alpha >= 1/255.
// conics[C*N,3], opacities[C*N] as handed to the compositing kernels.
// hit_masks (nullable, uint64[C*N]): scratch the count pass fills and the matching emit pass reads (same arguments).
// depths + depth_keys (uint64[C*N]) + depth_vals (int32[C*N]), all three or none: input of the depth sort of the
// two-level binning — key = the depth's float bits (0xffffffff for a Gaussian without entries), value = its index.
FSB_API int fsb_isect_count_reach(int C, int N, const float* means2d, const int32_t* radii, const float* conics,
                                  const float* opacities, int tile_size, int tile_w, int tile_h, int legacy_bbox,
                                  int32_t* counts, uint64_t* hit_masks, const float* depths, uint64_t* depth_keys,
                                  int32_t* depth_vals, void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0 || !conics || !opacities || !counts) return FSB_E_ARG;
    if ((depth_keys != nullptr) != (depth_vals != nullptr) || (depth_keys != nullptr && depths == nullptr)) return FSB_E_ARG;
    if ((int64_t)C * N > 0x7fffffffLL) return FSB_E_ARG;
    if (N == 0) return 0;
    const int64_t total = (int64_t)C * N;
    isect_reach_kernel<true><<<fsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        C, N, means2d, radii, depths, conics, opacities, nullptr, tile_size, tile_w, tile_h, 0, legacy_bbox, nullptr, 0,
        nullptr, counts, nullptr, nullptr, (unsigned long long*)hit_masks, nullptr, 0,
        (unsigned long long*)depth_keys, depth_vals);
    FSB_LAUNCH_CHECK();
    return 0;
}

// fsb_isect_emit restricted to the reached tiles; `offsets` = exclusive scan of fsb_isect_count_reach's counts.
// Static-capacity mode (n_dev, capacity, overflow_flag) as in fsb_isect_emit.
// perm (nullable, int32[C*N]): entry i of `offsets` belongs to Gaussian perm[i] (fsb_isect_scan_perm).
// packed != 0: isect_ids[k] = (camera, tile) << 32 | flatten id (flag included), flatten_ids may be null and is not
// written; feed the result to fsb_radix_sort_keys(begin_bit = 32).
FSB_API int fsb_isect_emit_reach(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                                 const float* conics, const float* opacities, const int64_t* offsets, int tile_size,
                                 int tile_w, int tile_h, int tile_bits, int legacy_bbox, const int64_t* n_dev,
                                 int64_t capacity, int32_t* overflow_flag, int64_t* isect_ids, int32_t* flatten_ids,
                                 const uint64_t* hit_masks, const int32_t* perm, int packed, void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0 || tile_bits < 0 || tile_bits > 30) return FSB_E_ARG;
    if (!conics || !opacities || !offsets || !isect_ids || (!flatten_ids && !packed)) return FSB_E_ARG;
    if (!packed && !depths) return FSB_E_ARG;
    if (n_dev && capacity < 0) return FSB_E_ARG;
    if ((int64_t)C * N > 0x7fffffffLL) return FSB_E_ARG;
    if (N == 0) return 0;
    const int64_t total = (int64_t)C * N;
    isect_reach_kernel<false><<<fsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        C, N, means2d, radii, depths, conics, opacities, offsets, tile_size, tile_w, tile_h, tile_bits, legacy_bbox, n_dev,
        capacity, overflow_flag, nullptr, isect_ids, flatten_ids, (unsigned long long*)hit_masks, perm, packed, nullptr,
        nullptr);
    FSB_LAUNCH_CHECK();
    return 0;
}
