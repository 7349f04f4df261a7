// isect_reach.cu — tile intersection with an exact reach test: count / emit only the (Gaussian, tile) pairs whose
// Gaussian can pass the alpha test (alpha >= 1/255) on at least one pixel centre of the tile.
//
// EXPERIMENTAL (round 1: compiled and declared, NOT yet exercised on a GPU; off by default, see
// DNSplatterStepConfig.prune_lists).  The reference bins by the 3-sigma bounding box
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

// COUNT_ONLY: counts[idx] = number of reached tiles; else emit keys / values for them in row-major tile order
// (the order fsb_isect_emit uses, so ties in the sort resolve identically).
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(256)
isect_reach_kernel(int C, int N, const float* __restrict__ means2d, const int32_t* __restrict__ radii,
                   const float* __restrict__ depths, const float* __restrict__ conics,
                   const float* __restrict__ opacities, const int64_t* __restrict__ offsets, int tile_size, int tile_w,
                   int tile_h, int tile_bits, int legacy_bbox, const int64_t* __restrict__ n_dev, int64_t capacity,
                   int32_t* __restrict__ overflow_flag, int32_t* __restrict__ counts, int64_t* __restrict__ isect_ids,
                   int32_t* __restrict__ flatten_ids) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)C * N) return;
    if (!COUNT_ONLY) {
        if (idx == 0 && n_dev && overflow_flag && *n_dev > capacity) *overflow_flag = 1;
        if (n_dev && *n_dev == 0) return;
    }
    const int r = radii[idx];
    if (r <= 0) {
        if (COUNT_ONLY) counts[idx] = 0;
        return;
    }
    const float2 m = reinterpret_cast<const float2*>(means2d)[idx];
    const float a = conics[3 * idx + 0], b = conics[3 * idx + 1], c = conics[3 * idx + 2];
    const float o = opacities[idx];
    const Box bx = tile_box(m.x, m.y, r, tile_size, tile_w, tile_h, legacy_bbox);
    const int64_t limit = n_dev ? capacity : INT64_MAX;
    int64_t cam_part = 0, depth_part = 0, cur = 0;
    if (!COUNT_ONLY) {
        cam_part = (idx / N) << (32 + tile_bits);
        depth_part = (int64_t)(uint32_t)__float_as_int(depths[idx]);
        cur = offsets[idx];
    }
    int n = 0;
    for (int i = bx.ay; i < bx.by; ++i)
        for (int j = bx.ax; j < bx.bx; ++j) {
            // pixel centres of the tile; the part of an edge tile beyond the image only makes the test more generous
            const float px0 = (float)(j * tile_size) + 0.5f, py0 = (float)(i * tile_size) + 0.5f;
            if (!rect_reach(m.x, m.y, o, a, b, c, px0, px0 + (float)(tile_size - 1), py0,
                            py0 + (float)(tile_size - 1)))
                continue;
            if (COUNT_ONLY) {
                ++n;
            } else {
                if (cur < limit) {
                    isect_ids[cur] = cam_part | ((int64_t)(i * tile_w + j) << 32) | depth_part;
                    flatten_ids[cur] = (int32_t)idx;
                }
                ++cur;
            }
        }
    if (COUNT_ONLY) counts[idx] = n;
}

}  // namespace

// counts[C*N] = tiles of the bounding box (legacy_bbox: 0.1.x rule) that the Gaussian can reach with alpha >= 1/255.
// conics[C*N,3], opacities[C*N] as handed to the compositing kernels.
FSB_API int fsb_isect_count_reach(int C, int N, const float* means2d, const int32_t* radii, const float* conics,
                                  const float* opacities, int tile_size, int tile_w, int tile_h, int legacy_bbox,
                                  int32_t* counts, void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0 || !conics || !opacities || !counts) return FSB_E_ARG;
    if (N == 0) return 0;
    const int64_t total = (int64_t)C * N;
    isect_reach_kernel<true><<<fsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        C, N, means2d, radii, nullptr, conics, opacities, nullptr, tile_size, tile_w, tile_h, 0, legacy_bbox, nullptr, 0,
        nullptr, counts, nullptr, nullptr);
    FSB_LAUNCH_CHECK();
    return 0;
}

// fsb_isect_emit restricted to the reached tiles; `offsets` = exclusive scan of fsb_isect_count_reach's counts.
// Static-capacity mode (n_dev, capacity, overflow_flag) as in fsb_isect_emit.
FSB_API int fsb_isect_emit_reach(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                                 const float* conics, const float* opacities, const int64_t* offsets, int tile_size,
                                 int tile_w, int tile_h, int tile_bits, int legacy_bbox, const int64_t* n_dev,
                                 int64_t capacity, int32_t* overflow_flag, int64_t* isect_ids, int32_t* flatten_ids,
                                 void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0 || tile_bits < 0 || tile_bits > 30) return FSB_E_ARG;
    if (!conics || !opacities || !offsets || !isect_ids || !flatten_ids) return FSB_E_ARG;
    if (n_dev && capacity < 0) return FSB_E_ARG;
    if ((int64_t)C * N > 0x7fffffffLL) return FSB_E_ARG;
    if (N == 0) return 0;
    const int64_t total = (int64_t)C * N;
    isect_reach_kernel<false><<<fsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        C, N, means2d, radii, depths, conics, opacities, offsets, tile_size, tile_w, tile_h, tile_bits, legacy_bbox, n_dev,
        capacity, overflow_flag, nullptr, isect_ids, flatten_ids);
    FSB_LAUNCH_CHECK();
    return 0;
}
