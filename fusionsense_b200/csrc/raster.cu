// raster.cu — tile rasterisation / alpha compositing, forward and backward.
//
// Replaces gsplat 1.0.0's rasterize_to_pixels_{fwd,bwd}_kernel and the legacy rasterize_forward /
// rasterize_backward_kernel (Appendix A.5/A.6), reached from
// /root/reference/dn_splatter/dn_model.py:570-591 (RGB + expected depth, D = 4) and :644-653
// (normals, D = 3, white background).
//
// One CTA per tile (tile_size x tile_size threads, one pixel each).  The tile's depth-sorted list is
// staged through shared memory in batches of blockDim threads; the per-pixel loop is FP32/MUFU bound.
// Forward: front-to-back, warp/CTA early termination.  Backward: back-to-front replay from last_ids,
// warp-shuffle reduction of the per-Gaussian partial gradients, one atomic per warp and value.
#include "common.cuh"

namespace {

constexpr float ALPHA_MAX = 0.999f;
constexpr float ALPHA_MIN = 1.f / 255.f;
constexpr float T_MIN = 1e-4f;
constexpr int MAX_BLOCK = 256;  // tile_size <= 16

template <int D>
__global__ void __launch_bounds__(MAX_BLOCK)
raster_fwd_kernel(int C, int N, int64_t n_isects, const float2* __restrict__ means2d,
                  const float* __restrict__ conics, const float* __restrict__ colors,
                  const float* __restrict__ opacities, const float* __restrict__ backgrounds,
                  const uint8_t* __restrict__ masks, int width, int height, int tile_size, int tile_w, int tile_h,
                  const int32_t* __restrict__ tile_offsets, const int32_t* __restrict__ flatten_ids,
                  int ed_normalize, float* __restrict__ out_colors, float* __restrict__ out_alphas,
                  int32_t* __restrict__ last_ids) {
    __shared__ int32_t s_id[MAX_BLOCK];
    __shared__ float4 s_xyo[MAX_BLOCK];
    __shared__ float4 s_con[MAX_BLOCK];
    __shared__ float s_col[MAX_BLOCK * D];

    const int n_tiles = tile_w * tile_h;
    const int64_t tile_lin = blockIdx.x;
    const int cam = (int)(tile_lin / n_tiles);
    const int tile_id = (int)(tile_lin - (int64_t)cam * n_tiles);
    const int tile_y = tile_id / tile_w, tile_x = tile_id - tile_y * tile_w;
    const int block_size = blockDim.x * blockDim.y;
    const int tr = threadIdx.y * blockDim.x + threadIdx.x;
    const int i = tile_y * tile_size + threadIdx.y;
    const int j = tile_x * tile_size + threadIdx.x;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < height && j < width);
    const int64_t pix = ((int64_t)cam * height + i) * width + j;

    float acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.f;

    if (masks != nullptr && !masks[tile_lin]) {
        if (inside) {
#pragma unroll
            for (int k = 0; k < D; ++k) out_colors[pix * D + k] = backgrounds ? backgrounds[cam * D + k] : 0.f;
            out_alphas[pix] = 0.f;
            last_ids[pix] = 0;
        }
        return;
    }

    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end =
        (tile_lin == (int64_t)C * n_tiles - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    const int num_batches = (range_end - range_start + block_size - 1) / block_size;

    bool done = !inside;
    float T = 1.f;
    int32_t cur_idx = 0;

    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= block_size) break;
        const int32_t batch_start = range_start + block_size * b;
        const int32_t idx = batch_start + tr;
        if (idx < range_end) {
            int32_t g = flatten_ids[idx];  // index into the flattened [C*N] arrays
            s_id[tr] = g;
            float2 xy = means2d[g];
            s_xyo[tr] = make_float4(xy.x, xy.y, opacities[g], 0.f);
            s_con[tr] = make_float4(conics[3 * (size_t)g], conics[3 * (size_t)g + 1], conics[3 * (size_t)g + 2], 0.f);
            const float* cp = colors + (size_t)g * D;
#pragma unroll
            for (int k = 0; k < D; ++k) s_col[tr * D + k] = cp[k];
        }
        __syncthreads();
        const int batch_size = min(block_size, range_end - batch_start);
        for (int t = 0; t < batch_size && !done; ++t) {
            const float4 xyo = s_xyo[t];
            const float4 con = s_con[t];
            const float dx = xyo.x - px, dy = xyo.y - py;
            const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
            const float alpha = fminf(ALPHA_MAX, xyo.z * __expf(-sigma));
            if (sigma < 0.f || alpha < ALPHA_MIN) continue;
            const float next_T = T * (1.f - alpha);
            if (next_T <= T_MIN) {
                done = true;
                break;
            }
            const float w = alpha * T;
#pragma unroll
            for (int k = 0; k < D; ++k) acc[k] += s_col[t * D + k] * w;
            cur_idx = batch_start + t;
            T = next_T;
        }
    }

    if (inside) {
        const float alpha_out = 1.f - T;
        out_alphas[pix] = alpha_out;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float v = backgrounds ? acc[k] + T * backgrounds[cam * D + k] : acc[k];
            if (ed_normalize && k == D - 1) v = v / fmaxf(alpha_out, 1e-10f);
            out_colors[pix * D + k] = v;
        }
        last_ids[pix] = cur_idx;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

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
                  float2* __restrict__ v_means2d_abs, float2* __restrict__ v_means2d, float* __restrict__ v_conics,
                  float* __restrict__ v_colors, float* __restrict__ v_opacities) {
    __shared__ int32_t s_id[MAX_BLOCK];
    __shared__ float4 s_xyo[MAX_BLOCK];
    __shared__ float4 s_con[MAX_BLOCK];
    __shared__ float s_col[MAX_BLOCK * D];

    const int n_tiles = tile_w * tile_h;
    const int64_t tile_lin = blockIdx.x;
    if (masks != nullptr && !masks[tile_lin]) return;
    const int cam = (int)(tile_lin / n_tiles);
    const int tile_id = (int)(tile_lin - (int64_t)cam * n_tiles);
    const int tile_y = tile_id / tile_w, tile_x = tile_id - tile_y * tile_w;
    const int block_size = blockDim.x * blockDim.y;
    const int tr = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tr & 31;
    const int i = tile_y * tile_size + threadIdx.y;
    const int j = tile_x * tile_size + threadIdx.x;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < height && j < width);
    const int64_t pix = inside ? ((int64_t)cam * height + i) * width + j : 0;

    const int32_t range_start = tile_offsets[tile_lin];
    const int32_t range_end =
        (tile_lin == (int64_t)C * n_tiles - 1) ? (int32_t)n_isects : tile_offsets[tile_lin + 1];
    const int num_batches = (range_end - range_start + block_size - 1) / block_size;
    if (num_batches <= 0) return;

    float T_final = 1.f, v_ra = 0.f;
    float v_rc[D];
    float buffer[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { v_rc[k] = 0.f; buffer[k] = 0.f; }
    int32_t bin_final = 0;
    if (inside) {
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
        for (int k = 0; k < D; ++k) bg_dot += backgrounds[cam * D + k] * v_rc[k];
    }
    int32_t warp_bin_final = inside ? bin_final : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, o));

    for (int b = 0; b < num_batches; ++b) {
        __syncthreads();
        const int32_t batch_end = range_end - 1 - block_size * b;
        const int batch_size = min(block_size, batch_end + 1 - range_start);
        const int32_t idx = batch_end - tr;
        if (idx >= range_start) {
            int32_t g = flatten_ids[idx];
            s_id[tr] = g;
            float2 xy = means2d[g];
            s_xyo[tr] = make_float4(xy.x, xy.y, opacities[g], 0.f);
            s_con[tr] = make_float4(conics[3 * (size_t)g], conics[3 * (size_t)g + 1], conics[3 * (size_t)g + 2], 0.f);
            const float* cp = colors + (size_t)g * D;
#pragma unroll
            for (int k = 0; k < D; ++k) s_col[tr * D + k] = cp[k];
        }
        __syncthreads();
        for (int t = max(0, batch_end - warp_bin_final); t < batch_size; ++t) {
            bool valid = inside && (batch_end - t <= bin_final);
            float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
            float4 con = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4 xyo = s_xyo[t];
                con = s_con[t];
                opac = xyo.z;
                dx = xyo.x - px; dy = xyo.y - py;
                const float sigma = 0.5f * (con.x * dx * dx + con.z * dy * dy) + con.y * dx * dy;
                vis = __expf(-sigma);
                alpha = fminf(ALPHA_MAX, opac * vis);
                if (sigma < 0.f || alpha < ALPHA_MIN) valid = false;
            }
            if (!__any_sync(0xffffffffu, valid)) continue;

            float v_col[D];
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = 0.f;
            float v_ca = 0.f, v_cb = 0.f, v_cc = 0.f, v_x = 0.f, v_y = 0.f, v_xa = 0.f, v_ya = 0.f, v_o = 0.f;
            if (valid) {
                const float ra = 1.f / (1.f - alpha);
                T *= ra;
                const float fac = alpha * T;
                float v_alpha = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float ck = s_col[t * D + k];
                    v_col[k] = fac * v_rc[k];
                    v_alpha += (ck * T - buffer[k] * ra) * v_rc[k];
                    buffer[k] += ck * fac;
                }
                v_alpha += T_final * ra * v_ra;
                if (backgrounds) v_alpha += -T_final * ra * bg_dot;
                if (opac * vis <= ALPHA_MAX) {
                    const float v_sigma = -opac * vis * v_alpha;
                    v_ca = 0.5f * v_sigma * dx * dx;
                    v_cb = v_sigma * dx * dy;
                    v_cc = 0.5f * v_sigma * dy * dy;
                    v_x = v_sigma * (con.x * dx + con.y * dy);
                    v_y = v_sigma * (con.y * dx + con.z * dy);
                    v_xa = fabsf(v_x);
                    v_ya = fabsf(v_y);
                    v_o = vis * v_alpha;
                }
            }
#pragma unroll
            for (int k = 0; k < D; ++k) v_col[k] = warp_sum(v_col[k]);
            v_ca = warp_sum(v_ca); v_cb = warp_sum(v_cb); v_cc = warp_sum(v_cc);
            v_x = warp_sum(v_x); v_y = warp_sum(v_y);
            v_o = warp_sum(v_o);
            if (v_means2d_abs) { v_xa = warp_sum(v_xa); v_ya = warp_sum(v_ya); }
            if (lane == 0) {
                const int32_t g = s_id[t];
                float* vc = v_colors + (size_t)g * D;
#pragma unroll
                for (int k = 0; k < D; ++k) atomicAdd(vc + k, v_col[k]);
                atomicAdd(v_conics + 3 * (size_t)g + 0, v_ca);
                atomicAdd(v_conics + 3 * (size_t)g + 1, v_cb);
                atomicAdd(v_conics + 3 * (size_t)g + 2, v_cc);
                atomicAdd(&v_means2d[g].x, v_x);
                atomicAdd(&v_means2d[g].y, v_y);
                if (v_means2d_abs) {
                    atomicAdd(&v_means2d_abs[g].x, v_xa);
                    atomicAdd(&v_means2d_abs[g].y, v_ya);
                }
                atomicAdd(v_opacities + g, v_o);
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
        v_render_colors, v_render_alphas, (float2*)v_means2d_abs, (float2*)v_means2d, v_conics, v_colors,
        v_opacities);
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
    if (C <= 0 || tile_size < 2 || tile_size > 16 || (tile_size * tile_size) % 32 != 0 || n_isects < 0 ||
        n_isects > 0x7fffffffLL)
        return FSB_E_ARG;  // whole warps only: tile_size 8 or 16 (the reference uses 16, dn_model.py:547)
    if (tile_w <= 0 || tile_h <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(D, (launch_fwd<DD>(C, N, n_isects, means2d, conics, colors, opacities, backgrounds, masks, width,
                                      height, tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize,
                                      out_colors, out_alphas, last_ids, st)));
}

// Gradient outputs are ACCUMULATED into (atomicAdd); the caller zero-fills them first.
FSB_API int fsb_raster_bwd(int C, int N, int D, int64_t n_isects, const float* means2d, const float* conics,
                           const float* colors, const float* opacities, const float* backgrounds,
                           const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                           const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_normalize,
                           const float* render_colors, const float* render_alphas, const int32_t* last_ids,
                           const float* v_render_colors, const float* v_render_alphas, float* v_means2d_abs,
                           float* v_means2d, float* v_conics, float* v_colors, float* v_opacities, void* stream) {
    if (C <= 0 || tile_size < 2 || tile_size > 16 || (tile_size * tile_size) % 32 != 0 || n_isects < 0 ||
        n_isects > 0x7fffffffLL)
        return FSB_E_ARG;  // whole warps only: tile_size 8 or 16 (the reference uses 16, dn_model.py:547)
    if (ed_normalize && !render_colors) return FSB_E_ARG;
    if (tile_w <= 0 || tile_h <= 0 || n_isects == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_DISPATCH_D(D, (launch_bwd<DD>(C, N, n_isects, means2d, conics, colors, opacities, backgrounds, masks, width,
                                      height, tile_size, tile_w, tile_h, tile_offsets, flatten_ids, ed_normalize,
                                      render_colors, render_alphas, last_ids, v_render_colors, v_render_alphas,
                                      v_means2d_abs, v_means2d, v_conics, v_colors, v_opacities, st)));
}
