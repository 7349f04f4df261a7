// raster.cu — tile rasterisation / alpha compositing, forward and backward.
//
// Replaces gsplat 1.0.0's rasterize_to_pixels_{fwd,bwd}_kernel and the legacy rasterize_forward /
// rasterize_backward_kernel (Appendix A.5/A.6), reached from
// /root/reference/dn_splatter/dn_model.py:570-591 (RGB + expected depth, D = 4) and :644-653
// (normals, D = 3, white background).
//
// One CTA per tile (tile_size x tile_size threads, one pixel each; a warp owns a strip of 32/tile_size rows).
// The tile's depth-sorted list is staged through shared memory in batches of blockDim entries.  While staging,
// each thread computes for its Gaussian the exact bounding box of the ellipse {alpha >= 1/255} and from it
// the set of warp strips the Gaussian can touch (an 8-bit mask).  Every warp then walks only the entries
// whose bit is set (ballot + find-first-set): entries it skips are entries all of its lanes would have
// rejected with the alpha < 1/255 test, so results are unchanged.  The per-pixel loop is FP32/MUFU bound.
//
// Forward: front to back, four list entries evaluated together (independent sigma/exp chains) before the
// sequential blend; warp and CTA early termination.
// Backward: back to front replay from last_ids; the 8 + D per-Gaussian partial gradients of a warp are
// reduced with a transposing butterfly (16 shuffles instead of 5 * (8 + D)) that leaves each total in a
// different lane, so one warp-wide red.global.add updates all of them.
#include "common.cuh"

namespace {

constexpr float ALPHA_MAX = 0.999f;
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float T_MIN = 1e-4f;
constexpr int MAX_BLOCK = 256;  // tile_size <= 16

struct TileGeom {
    int cam, tile_x, tile_y;
    int block_size, tr, lane, warp, n_warps, rows_per_warp;
    int i, j;
    float px, py;
    bool inside;
};

__device__ __forceinline__ TileGeom tile_geom(int tile_w, int tile_h, int tile_size, int width, int height) {
    TileGeom g;
    const int n_tiles = tile_w * tile_h;
    const int64_t tile_lin = blockIdx.x;
    g.cam = (int)(tile_lin / n_tiles);
    const int tile_id = (int)(tile_lin - (int64_t)g.cam * n_tiles);
    g.tile_y = tile_id / tile_w;
    g.tile_x = tile_id - g.tile_y * tile_w;
    g.block_size = blockDim.x * blockDim.y;
    g.tr = threadIdx.y * blockDim.x + threadIdx.x;
    g.lane = g.tr & 31;
    g.warp = g.tr >> 5;
    g.n_warps = g.block_size >> 5;
    g.rows_per_warp = 32 / tile_size;
    g.i = g.tile_y * tile_size + threadIdx.y;
    g.j = g.tile_x * tile_size + threadIdx.x;
    g.px = (float)g.j + 0.5f;
    g.py = (float)g.i + 0.5f;
    g.inside = (g.i < height && g.j < width);
    return g;
}

// Which warp strips of this tile can the Gaussian reach with alpha >= 1/255 ?  Conservative by construction:
// the ellipse {sigma <= ln(255 * opacity)} is inflated, and anything doubtful (non-PD conic, NaN) keeps all bits.
__device__ __forceinline__ uint32_t strip_mask(float gx, float gy, float opac, float a, float b, float c,
                                               float tile_px0, float tile_py0, int tile_size, int rows_per_warp,
                                               int n_warps) {
    const uint32_t all = (1u << n_warps) - 1u;
    const float tau = __logf(255.f * opac) + 2e-3f;
    if (tau < 0.f) return 0u;  // opacity below 1/255: can never pass the alpha test
    const float det = a * c - b * b;
    if (!(det > 0.f) || !(tau < 1e30f)) return all;
    const float k = 2.f * tau / det;
    const float ex = sqrtf(k * c) * 1.0001f + 1e-3f;
    const float ey = sqrtf(k * a) * 1.0001f + 1e-3f;
    if (gx + ex < tile_px0 + 0.5f || gx - ex > tile_px0 + (float)tile_size - 0.5f) return 0u;
    uint32_t m = 0u;
    for (int w = 0; w < n_warps; ++w) {
        const float y_lo = tile_py0 + (float)(w * rows_per_warp) + 0.5f;
        const float y_hi = y_lo + (float)(rows_per_warp - 1);
        if (!(gy + ey < y_lo) && !(gy - ey > y_hi)) m |= 1u << w;
    }
    return m;
}

template <int D>
struct Stage {
    int32_t id[MAX_BLOCK];
    float4 xyo[MAX_BLOCK];  // x, y, opacity, strip mask (as int bits)
    float4 con[MAX_BLOCK];  // conic a, b, c
    float col[MAX_BLOCK * D];
};

template <int D>
__device__ __forceinline__ void stage_entry(Stage<D>& s, int slot, int32_t g, const float2* __restrict__ means2d,
                                            const float* __restrict__ conics, const float* __restrict__ colors,
                                            const float* __restrict__ opacities, const TileGeom& tg, int tile_size) {
    s.id[slot] = g;
    const float2 xy = means2d[g];
    const float o = opacities[g];
    const float a = conics[3 * (size_t)g], b = conics[3 * (size_t)g + 1], c = conics[3 * (size_t)g + 2];
    const uint32_t m = strip_mask(xy.x, xy.y, o, a, b, c, (float)(tg.tile_x * tile_size),
                                  (float)(tg.tile_y * tile_size), tile_size, tg.rows_per_warp, tg.n_warps);
    s.xyo[slot] = make_float4(xy.x, xy.y, o, __int_as_float((int)m));
    s.con[slot] = make_float4(a, b, c, 0.f);
    const float* cp = colors + (size_t)g * D;
#pragma unroll
    for (int k = 0; k < D; ++k) s.col[slot * D + k] = cp[k];
}

template <int D>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_fwd_kernel(int C, int N, int64_t n_isects, const float2* __restrict__ means2d,
                  const float* __restrict__ conics, const float* __restrict__ colors,
                  const float* __restrict__ opacities, const float* __restrict__ backgrounds,
                  const uint8_t* __restrict__ masks, int width, int height, int tile_size, int tile_w, int tile_h,
                  const int32_t* __restrict__ tile_offsets, const int32_t* __restrict__ flatten_ids,
                  int ed_normalize, float* __restrict__ out_colors, float* __restrict__ out_alphas,
                  int32_t* __restrict__ last_ids) {
    __shared__ Stage<D> s;
    const TileGeom tg = tile_geom(tile_w, tile_h, tile_size, width, height);
    const int64_t tile_lin = blockIdx.x;
    const int64_t pix = ((int64_t)tg.cam * height + tg.i) * width + tg.j;

    float acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.f;

    if (masks != nullptr && !masks[tile_lin]) {
        if (tg.inside) {
#pragma unroll
            for (int k = 0; k < D; ++k) out_colors[pix * D + k] = backgrounds ? backgrounds[tg.cam * D + k] : 0.f;
            out_alphas[pix] = 0.f;
            last_ids[pix] = 0;
        }
        return;
    }

    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end =
        (tile_lin == (int64_t)C * tile_w * tile_h - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    const int num_batches = (range_end - range_start + tg.block_size - 1) / tg.block_size;

    bool done = !tg.inside;
    float T = 1.f;
    int32_t cur_idx = 0;

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= tg.block_size) break;
        const int32_t batch_start = range_start + tg.block_size * b;
        const int32_t idx = batch_start + tg.tr;
        if (idx < range_end) stage_entry<D>(s, tg.tr, flatten_ids[idx], means2d, conics, colors, opacities, tg, tile_size);
        __syncthreads();
        const int batch_size = min(tg.block_size, range_end - batch_start);
        bool warp_done = __all_sync(0xffffffffu, done);
        for (int k0 = 0; k0 < batch_size && !warp_done; k0 += 32) {
            const int tt = k0 + tg.lane;
            const uint32_t m = (tt < batch_size) ? (uint32_t)__float_as_int(s.xyo[tt].w) : 0u;
            uint32_t bits = __ballot_sync(0xffffffffu, (m >> tg.warp) & 1u);
            while (bits) {
                // up to four entries of this warp's strip, evaluated together
                int t[4];
                float alpha[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (bits) {
                        t[u] = k0 + __ffs(bits) - 1;
                        bits &= bits - 1;
                    } else {
                        t[u] = -1;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ok[u] = false;
                    alpha[u] = 0.f;
                    if (t[u] >= 0) {
                        const float4 xyo = s.xyo[t[u]];
                        const float4 con = s.con[t[u]];
                        const float dx = xyo.x - tg.px, dy = xyo.y - tg.py;
                        const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
                        alpha[u] = fminf(ALPHA_MAX, xyo.z * __expf(-sigma));
                        ok[u] = !(sigma < 0.f || alpha[u] < ALPHA_MIN);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (ok[u] && !done) {
                        const float next_T = T * (1.f - alpha[u]);
                        if (next_T <= T_MIN) {
                            done = true;
                        } else {
                            const float w = alpha[u] * T;
#pragma unroll
                            for (int k = 0; k < D; ++k) acc[k] += s.col[t[u] * D + k] * w;
                            cur_idx = batch_start + t[u];
                            T = next_T;
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done)) {
                    warp_done = true;
                    break;
                }
            }
        }
    }

    if (tg.inside) {
        const float alpha_out = 1.f - T;
        out_alphas[pix] = alpha_out;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float v = backgrounds ? acc[k] + T * backgrounds[tg.cam * D + k] : acc[k];
            if (ed_normalize && k == D - 1) v = v / fmaxf(alpha_out, 1e-10f);
            out_colors[pix * D + k] = v;
        }
        last_ids[pix] = cur_idx;
    }
}

// Sum NV (<= 16) per-lane values over the warp so that lane L ends up with the total of value index
// slot_of_lane(L); 16 shuffles in all.  v[] is clobbered; the total is returned.
template <int NV>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[16], int lane) {
    static_assert(NV <= 16, "at most 16 values");
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = hi ? v[i] : v[i + 8];
            const float keep = hi ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = hi ? v[i] : v[i + 4];
            const float keep = hi ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = hi ? v[i] : v[i + 2];
            const float keep = hi ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// value index held by lane L after warp_transpose_sum
__device__ __forceinline__ int slot_of_lane(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// Generic (D > 8) fallback reduction: plain butterflies.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Slots of the packed gradient vector: [0, D) colours, then conic a b c, opacity, xy, |xy|.
template <int D>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_bwd_kernel(int C, int N, int64_t n_isects, const float2* __restrict__ means2d,
                  const float* __restrict__ conics, const float* __restrict__ colors,
                  const float* __restrict__ opacities, const float* __restrict__ backgrounds,
                  const uint8_t* __restrict__ masks, int width, int height, int tile_size, int tile_w, int tile_h,
                  const int32_t* __restrict__ tile_offsets, const int32_t* __restrict__ flatten_ids,
                  int ed_normalize, const float* __restrict__ render_colors,
                  const float* __restrict__ render_alphas, const int32_t* __restrict__ last_ids,
                  const float* __restrict__ v_render_colors, const float* __restrict__ v_render_alphas,
                  float* __restrict__ v_means2d_abs, float* __restrict__ v_means2d, float* __restrict__ v_conics,
                  float* __restrict__ v_colors, float* __restrict__ v_opacities) {
    __shared__ Stage<D> s;
    const TileGeom tg = tile_geom(tile_w, tile_h, tile_size, width, height);
    const int64_t tile_lin = blockIdx.x;
    if (masks != nullptr && !masks[tile_lin]) return;
    const int64_t pix = tg.inside ? ((int64_t)tg.cam * height + tg.i) * width + tg.j : 0;

    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end =
        (tile_lin == (int64_t)C * tile_w * tile_h - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    const int num_batches = (range_end - range_start + tg.block_size - 1) / tg.block_size;
    if (num_batches <= 0) return;

    float T_final = 1.f, v_ra = 0.f;
    float v_rc[D];
    float buffer[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { v_rc[k] = 0.f; buffer[k] = 0.f; }
    int32_t bin_final = -1;
    if (tg.inside) {
        const float alpha_out = render_alphas[pix];
        T_final = 1.f - alpha_out;
        v_ra = v_render_alphas[pix];
#pragma unroll
        for (int k = 0; k < D; ++k) v_rc[k] = v_render_colors[pix * D + k];
        if (ed_normalize) {
            // out[D-1] = acc / max(alpha, 1e-10)
            const float den = fmaxf(alpha_out, 1e-10f);
            const float v_ed = v_rc[D - 1];
            v_rc[D - 1] = v_ed / den;
            if (alpha_out >= 1e-10f) v_ra += -v_ed * render_colors[pix * D + D - 1] / den;
        }
        bin_final = last_ids[pix];
    }
    float T = T_final;
    float bg_dot = 0.f;
    if (backgrounds) {
#pragma unroll
        for (int k = 0; k < D; ++k) bg_dot += backgrounds[tg.cam * D + k] * v_rc[k];
    }
    int32_t warp_bin_final = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));
    // deepest list position any pixel of the tile reached: batches entirely behind it are never staged
    __shared__ int32_t s_wmax[MAX_BLOCK / 32];
    if (tg.lane == 0) s_wmax[tg.warp] = warp_bin_final;
    __syncthreads();
    int32_t cta_bin_final = -1;
    for (int w = 0; w < tg.n_warps; ++w) cta_bin_final = max(cta_bin_final, s_wmax[w]);
    if (cta_bin_final < range_start) return;
    const int b_first = (range_end - 1 - cta_bin_final) / tg.block_size;
    const bool want_xy = (v_means2d != nullptr);
    const bool want_abs = (v_means2d_abs != nullptr);

    // per-lane atomic target for the transposed reduction (D <= 8: 8 + D <= 16 slots)
    constexpr bool kTranspose = (D <= 8);
    const int slot = slot_of_lane(tg.lane);
    float* slot_base = nullptr;
    int slot_stride = 0;
    if (kTranspose && !(tg.lane & 1)) {
        if (slot < D) { slot_base = v_colors + slot; slot_stride = D; }
        else if (slot < D + 3) { slot_base = v_conics + (slot - D); slot_stride = 3; }
        else if (slot == D + 3) { slot_base = v_opacities; slot_stride = 1; }
        else if (slot < D + 6) { slot_base = want_xy ? v_means2d + (slot - D - 4) : nullptr; slot_stride = 2; }
        else if (slot < D + 8) { slot_base = want_abs ? v_means2d_abs + (slot - D - 6) : nullptr; slot_stride = 2; }
    }

    for (int b = b_first; b < num_batches; ++b) {
        __syncthreads();
        const int32_t batch_end = range_end - 1 - tg.block_size * b;
        const int batch_size = min(tg.block_size, batch_end + 1 - range_start);
        const int32_t idx = batch_end - tg.tr;
        if (idx >= range_start) stage_entry<D>(s, tg.tr, flatten_ids[idx], means2d, conics, colors, opacities, tg, tile_size);
        __syncthreads();
        const int t_first = max(0, batch_end - warp_bin_final);
        for (int k0 = (t_first & ~31); k0 < batch_size; k0 += 32) {
            const int tt = k0 + tg.lane;
            const uint32_t m = (tt < batch_size && tt >= t_first) ? (uint32_t)__float_as_int(s.xyo[tt].w) : 0u;
            uint32_t bits = __ballot_sync(0xffffffffu, (m >> tg.warp) & 1u);
            while (bits) {
                const int t = k0 + __ffs(bits) - 1;
                bits &= bits - 1;
                bool valid = tg.inside && (batch_end - t <= bin_final);
                float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
                float4 con = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    const float4 xyo = s.xyo[t];
                    con = s.con[t];
                    opac = xyo.z;
                    dx = xyo.x - tg.px; dy = xyo.y - tg.py;
                    const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
                    vis = __expf(-sigma);
                    alpha = fminf(ALPHA_MAX, opac * vis);
                    if (sigma < 0.f || alpha < ALPHA_MIN) valid = false;
                }
                if (!__any_sync(0xffffffffu, valid)) continue;

                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = 0.f;
                float v_colD[D > 8 ? D : 1];
                if (valid) {
                    const float ra = 1.f / (1.f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    float v_alpha = 0.f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const float ck = s.col[t * D + k];
                        if constexpr (kTranspose) v[k] = fac * v_rc[k];
                        else v_colD[k] = fac * v_rc[k];
                        v_alpha += (ck * T - buffer[k] * ra) * v_rc[k];
                        buffer[k] += ck * fac;
                    }
                    v_alpha += T_final * ra * v_ra;
                    if (backgrounds) v_alpha += -T_final * ra * bg_dot;
                    if (opac * vis <= ALPHA_MAX) {
                        const float v_sigma = -opac * vis * v_alpha;
                        constexpr int B = kTranspose ? D : 0;
                        v[B + 0] = 0.5f * v_sigma * dx * dx;
                        v[B + 1] = v_sigma * dx * dy;
                        v[B + 2] = 0.5f * v_sigma * dy * dy;
                        v[B + 3] = vis * v_alpha;
                        if (want_xy) {
                            const float gx = v_sigma * (con.x * dx + con.y * dy);
                            const float gy = v_sigma * (con.y * dx + con.z * dy);
                            v[B + 4] = gx;
                            v[B + 5] = gy;
                            if (want_abs) {
                                v[B + 6] = fabsf(gx);
                                v[B + 7] = fabsf(gy);
                            }
                        }
                    }
                } else if constexpr (!kTranspose) {
#pragma unroll
                    for (int k = 0; k < D; ++k) v_colD[k] = 0.f;
                }
                const int32_t g = s.id[t];
                if constexpr (kTranspose) {
                    const float total = warp_transpose_sum<8 + D>(v, tg.lane);
                    if (slot_base != nullptr && total != 0.f) atomicAdd(slot_base + (size_t)g * slot_stride, total);
                } else {
                    // wide colour vectors: colours by plain butterflies, the 8 geometric values transposed
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const float tot = warp_sum(v_colD[k]);
                        if (tg.lane == 0) atomicAdd(v_colors + (size_t)g * D + k, tot);
                    }
                    const float total = warp_transpose_sum<8>(v, tg.lane);
                    if (!(tg.lane & 1) && total != 0.f) {
                        float* p = nullptr;
                        if (slot < 3) p = v_conics + 3 * (size_t)g + slot;
                        else if (slot == 3) p = v_opacities + g;
                        else if (slot < 6) p = want_xy ? v_means2d + 2 * (size_t)g + (slot - 4) : nullptr;
                        else if (slot < 8) p = want_abs ? v_means2d_abs + 2 * (size_t)g + (slot - 6) : nullptr;
                        if (p) atomicAdd(p, total);
                    }
                }
            }
        }
    }
}

template <int D>
int launch_fwd(int C, int N, int64_t n_isects, const float* means2d, const float* conics, const float* colors,
               const float* opacities, const float* backgrounds, const uint8_t* masks, int width, int height,
               int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids,
               int ed_normalize, float* out_colors, float* out_alphas, int32_t* last_ids, cudaStream_t st) {
    dim3 block(tile_size, tile_size);
    unsigned grid = (unsigned)((int64_t)C * tile_w * tile_h);
    raster_fwd_kernel<D><<<grid, block, 0, st>>>(C, N, n_isects, (const float2*)means2d, conics, colors, opacities,
                                                 backgrounds, masks, width, height, tile_size, tile_w, tile_h,
                                                 tile_offsets, flatten_ids, ed_normalize, out_colors, out_alphas,
                                                 last_ids);
    FSB_LAUNCH_CHECK();
    return 0;
}

template <int D>
int launch_bwd(int C, int N, int64_t n_isects, const float* means2d, const float* conics, const float* colors,
               const float* opacities, const float* backgrounds, const uint8_t* masks, int width, int height,
               int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids,
               int ed_normalize, const float* render_colors, const float* render_alphas, const int32_t* last_ids,
               const float* v_render_colors, const float* v_render_alphas, float* v_means2d_abs, float* v_means2d,
               float* v_conics, float* v_colors, float* v_opacities, cudaStream_t st) {
    dim3 block(tile_size, tile_size);
    unsigned grid = (unsigned)((int64_t)C * tile_w * tile_h);
    raster_bwd_kernel<D><<<grid, block, 0, st>>>(
        C, N, n_isects, (const float2*)means2d, conics, colors, opacities, backgrounds, masks, width, height,
        tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize, render_colors, render_alphas, last_ids,
        v_render_colors, v_render_alphas, v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities);
    FSB_LAUNCH_CHECK();
    return 0;
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

FSB_API int fsb_raster_fwd(int C, int N, int D, int64_t n_isects, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* backgrounds,
                           const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                           const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_normalize,
                           float* out_colors, float* out_alphas, int32_t* last_ids, void* stream) {
    if (C <= 0 || tile_size < 2 || tile_size > 16 || (tile_size * tile_size) % 32 != 0 || 32 % tile_size != 0 ||
        n_isects < 0 || n_isects > 0x7fffffffLL)
        return FSB_E_ARG;  // whole warps made of whole rows: tile_size 8 or 16 (the reference uses 16, dn_model.py:547)
    if (tile_w <= 0 || tile_h <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(D, (launch_fwd<DD>(C, N, n_isects, means2d, conics, colors, opacities, backgrounds, masks, width,
                                      height, tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize,
                                      out_colors, out_alphas, last_ids, st)));
}

// Gradient outputs are ACCUMULATED into (atomicAdd); the caller zero-fills them first.
// v_means2d / v_means2d_abs may be NULL (no gradient wanted for the 2-D means: the legacy normals pass of
// dn_model.py:638 detaches them); v_means2d_abs requires v_means2d.
FSB_API int fsb_raster_bwd(int C, int N, int D, int64_t n_isects, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* backgrounds,
                           const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                           const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_normalize,
                           const float* render_colors, const float* render_alphas, const int32_t* last_ids,
                           const float* v_render_colors, const float* v_render_alphas, float* v_means2d_abs,
                           float* v_means2d, float* v_conics, float* v_colors, float* v_opacities, void* stream) {
    if (C <= 0 || tile_size < 2 || tile_size > 16 || (tile_size * tile_size) % 32 != 0 || 32 % tile_size != 0 ||
        n_isects < 0 || n_isects > 0x7fffffffLL)
        return FSB_E_ARG;
    if (ed_normalize && !render_colors) return FSB_E_ARG;
    if (v_means2d_abs && !v_means2d) return FSB_E_ARG;
    if (tile_w <= 0 || tile_h <= 0 || n_isects == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(D, (launch_bwd<DD>(C, N, n_isects, means2d, conics, colors, opacities, backgrounds, masks, width,
                                      height, tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize,
                                      render_colors, render_alphas, last_ids, v_render_colors, v_render_alphas,
                                      v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities, st)));
}
