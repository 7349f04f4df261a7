// hull_prune.cu — visual-hull pruning mask without the N x V distance matrix.
//
// Replaces the every-100-steps block of /root/reference/dn_splatter/dn_model.py:1254-1264:
//   center = hull.mean(0) ; close = |means - center| <= 0.2 s
//   d_min  = torch.cdist(means[close], hull).min(-1)          (materialises an N_close x V fp32 matrix: 36 GB at
//                                                              300k Gaussians x 30k hull points)
//   mask[close] = (d_min > 0.005 s) & (d_min <= 0.02 s)
// One thread per Gaussian; hull points stream through shared memory in tiles; a thread stops as soon as it has
// seen a hull point within the lower threshold (its mask bit is then 0 whatever comes later).
// Distances are direct differences in fp32 (torch.cdist switches to the |a|^2 + |b|^2 - 2ab matmul form above 25
// rows, which is LESS accurate; the parity test checks against an fp64 brute force and allows flips only where
// d_min is within 1e-6 relative of a threshold).
// FP32-pipe bound: N_close * V * 8 flop; HBM traffic is N * 12 + V * 12 + N bytes.
#include "common.cuh"

namespace {

constexpr int HP_THREADS = 256;
constexpr int HP_TILE = 1024;  // hull points per shared-memory tile (12 KB)

__global__ void __launch_bounds__(HP_THREADS)
hull_min_dist_kernel(int N, const float* __restrict__ pts, int V, const float* __restrict__ hull,
                     const float* __restrict__ center, float r_close, float stop_below,
                     float* __restrict__ min_dist) {
    __shared__ float hx[HP_TILE], hy[HP_TILE], hz[HP_TILE];
    const int n = blockIdx.x * HP_THREADS + threadIdx.x;
    float px = 0.f, py = 0.f, pz = 0.f;
    bool active = false;
    if (n < N) {
        px = pts[3 * (size_t)n]; py = pts[3 * (size_t)n + 1]; pz = pts[3 * (size_t)n + 2];
        active = true;
        if (center) {
            const float dx = px - center[0], dy = py - center[1], dz = pz - center[2];
            active = sqrtf(dx * dx + dy * dy + dz * dz) <= r_close;
        }
    }
    const float stop2 = stop_below > 0.f ? stop_below * stop_below : -1.f;
    float best = INFINITY;  // squared
    for (int base = 0; base < V; base += HP_TILE) {
        // whole block finished (everyone inactive or already below the stop threshold)?
        if (__syncthreads_count(active && !(best <= stop2)) == 0) break;
        const int cnt = min(HP_TILE, V - base);
        for (int e = threadIdx.x; e < cnt; e += HP_THREADS) {
            const float* h = hull + 3 * (size_t)(base + e);
            hx[e] = h[0]; hy[e] = h[1]; hz[e] = h[2];
        }
        __syncthreads();
        if (active && !(best <= stop2)) {
            float b0 = best, b1 = best, b2 = best, b3 = best;
            int e = 0;
            for (; e + 4 <= cnt; e += 4) {
                float dx, dy, dz;
                dx = px - hx[e]; dy = py - hy[e]; dz = pz - hz[e];
                b0 = fminf(b0, dx * dx + dy * dy + dz * dz);
                dx = px - hx[e + 1]; dy = py - hy[e + 1]; dz = pz - hz[e + 1];
                b1 = fminf(b1, dx * dx + dy * dy + dz * dz);
                dx = px - hx[e + 2]; dy = py - hy[e + 2]; dz = pz - hz[e + 2];
                b2 = fminf(b2, dx * dx + dy * dy + dz * dz);
                dx = px - hx[e + 3]; dy = py - hy[e + 3]; dz = pz - hz[e + 3];
                b3 = fminf(b3, dx * dx + dy * dy + dz * dz);
            }
            for (; e < cnt; ++e) {
                const float dx = px - hx[e], dy = py - hy[e], dz = pz - hz[e];
                b0 = fminf(b0, dx * dx + dy * dy + dz * dz);
            }
            best = fminf(fminf(b0, b1), fminf(b2, b3));
        }
    }
    if (n < N) min_dist[n] = active ? sqrtf(best) : INFINITY;
}

__global__ void __launch_bounds__(HP_THREADS)
hull_mask_kernel(int N, const float* __restrict__ min_dist, float lo, float hi, const uint8_t* __restrict__ protect,
                 uint8_t* __restrict__ mask) {
    const int n = blockIdx.x * HP_THREADS + threadIdx.x;
    if (n >= N) return;
    const float d = min_dist[n];
    bool m = (d > lo) && (d <= hi);
    if (protect && protect[n]) m = false;
    mask[n] = m ? 1 : 0;
}

}  // namespace

// min_dist[n] = min_v |pts[n] - hull[v]| for the points within r_close of `center` (device [3], nullable = all
// points), +inf for the others.  stop_below > 0: a point may stop searching once a hull point is closer than
// stop_below (its min_dist is then only an upper bound <= stop_below).  V == 0 gives +inf everywhere.
FSB_API int fsb_hull_min_dist(int N, const float* pts, int V, const float* hull, const float* center, float r_close,
                              float stop_below, float* min_dist, void* stream) {
    if (N < 0 || V < 0) return FSB_E_ARG;
    if (N == 0) return 0;
    hull_min_dist_kernel<<<fsb_div_up(N, HP_THREADS), HP_THREADS, 0, (cudaStream_t)stream>>>(
        N, pts, V, hull, center, r_close, stop_below, min_dist);
    FSB_LAUNCH_CHECK();
    return 0;
}

// mask[n] = (lo < min_dist[n] <= hi) and not protect[n]   (protect: u8, nullable — dn_model.py:1268-1269 add_mask)
FSB_API int fsb_hull_prune_mask(int N, const float* min_dist, float lo, float hi, const uint8_t* protect,
                                uint8_t* mask, void* stream) {
    if (N < 0) return FSB_E_ARG;
    if (N == 0) return 0;
    hull_mask_kernel<<<fsb_div_up(N, HP_THREADS), HP_THREADS, 0, (cudaStream_t)stream>>>(N, min_dist, lo, hi, protect,
                                                                                      mask);
    FSB_LAUNCH_CHECK();
    return 0;
}
