// compose.cu — the image-space glue of DNSplatterModel.get_outputs and the flatness regulariser, fused.
//
// Replaces the elementwise / reduction launches (and their autograd mirrors) that torch issues for
//   /root/reference/dn_splatter/dn_model.py:602-604   rgb = clamp(render[..., :3] + (1 - alpha) * background, 0, 1)
//   /root/reference/dn_splatter/dn_model.py:609-613   depth = where(alpha > 0, render[..., 3:4], render[..., 3:4].detach().max())
//   /root/reference/dn_splatter/dn_model.py:655-656   normal = (n / ||n|| + 1) / 2
//   /root/reference/dn_splatter/dn_model.py:817-819   two_d_gaussians: mean_i min_k exp(scales[i, k])
// about 35 launches per iteration forward + backward at 640x480 become 7.  One thread per pixel / Gaussian;
// HBM-trivial (a 640x480 frame is 6 MB), the point is launch count on a launch/latency-bound step.
// Operation order follows torch's (separate mul and add, true division) so results agree to the last bit where
// torch's own kernels are exactly rounded.
#include "common.cuh"

namespace {

constexpr int C_THREADS = 256;

// float max through integer atomics, valid for any mix of signs; the cell starts at 0xffffffff, which loses
// against every float both as a signed int (-1 < bits of any v >= 0) and as an unsigned int (max > bits of any v < 0)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
    else atomicMin((unsigned*)addr, __float_as_uint(v));
}

__global__ void __launch_bounds__(C_THREADS)
compose_rgbd_kernel(int64_t P, const float4* __restrict__ render, const float* __restrict__ alpha,
                    const float* __restrict__ bg, float* __restrict__ rgb, float* __restrict__ depth_max) {
    __shared__ float s_max[C_THREADS / 32];
    const int64_t p = (int64_t)blockIdx.x * C_THREADS + threadIdx.x;
    float d = -INFINITY;
    if (p < P) {
        const float4 r = render[p];
        const float om = __fsub_rn(1.f, alpha[p]);
        const float x0 = __fadd_rn(r.x, __fmul_rn(om, bg[0]));
        const float x1 = __fadd_rn(r.y, __fmul_rn(om, bg[1]));
        const float x2 = __fadd_rn(r.z, __fmul_rn(om, bg[2]));
        rgb[3 * p + 0] = fminf(fmaxf(x0, 0.f), 1.f);
        rgb[3 * p + 1] = fminf(fmaxf(x1, 0.f), 1.f);
        rgb[3 * p + 2] = fminf(fmaxf(x2, 0.f), 1.f);
        d = r.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_max[0];
#pragma unroll
        for (int w = 1; w < C_THREADS / 32; ++w) m = fmaxf(m, s_max[w]);
        if (m > -INFINITY) atomic_max_float(depth_max, m);
    }
}

__global__ void __launch_bounds__(C_THREADS)
compose_depth_kernel(int64_t P, const float4* __restrict__ render, const float* __restrict__ alpha,
                     const float* __restrict__ depth_max, float* __restrict__ depth) {
    const int64_t p = (int64_t)blockIdx.x * C_THREADS + threadIdx.x;
    if (p >= P) return;
    depth[p] = (alpha[p] > 0.f) ? render[p].w : *depth_max;
}

__global__ void __launch_bounds__(C_THREADS)
compose_rgbd_bwd_kernel(int64_t P, const float4* __restrict__ render, const float* __restrict__ alpha,
                        const float* __restrict__ bg, const float* __restrict__ v_rgb,
                        const float* __restrict__ v_depth, float4* __restrict__ v_render,
                        float* __restrict__ v_alpha) {
    const int64_t p = (int64_t)blockIdx.x * C_THREADS + threadIdx.x;
    if (p >= P) return;
    const float4 r = render[p];
    const float a = alpha[p];
    const float om = __fsub_rn(1.f, a);
    float g[3] = {0.f, 0.f, 0.f};
    float va = 0.f;
    if (v_rgb) {
        const float x[3] = {__fadd_rn(r.x, __fmul_rn(om, bg[0])), __fadd_rn(r.y, __fmul_rn(om, bg[1])),
                            __fadd_rn(r.z, __fmul_rn(om, bg[2]))};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // clamp passes the gradient where min <= x <= max (torch clamp_backward)
            g[c] = (x[c] >= 0.f && x[c] <= 1.f) ? v_rgb[3 * p + c] : 0.f;
            va -= g[c] * bg[c];
        }
    }
    const float gd = (v_depth && a > 0.f) ? v_depth[p] : 0.f;
    v_render[p] = make_float4(g[0], g[1], g[2], gd);
    v_alpha[p] = va;
}

__global__ void __launch_bounds__(C_THREADS)
normal_map_kernel(int64_t P, const float* __restrict__ n_raw, float* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * C_THREADS + threadIdx.x;
    if (p >= P) return;
    const float x = n_raw[3 * p + 0], y = n_raw[3 * p + 1], z = n_raw[3 * p + 2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    // no epsilon, like the reference: a zero vector gives NaN there as well
    out[3 * p + 0] = __fmul_rn(__fadd_rn(__fdiv_rn(x, nrm), 1.f), 0.5f);
    out[3 * p + 1] = __fmul_rn(__fadd_rn(__fdiv_rn(y, nrm), 1.f), 0.5f);
    out[3 * p + 2] = __fmul_rn(__fadd_rn(__fdiv_rn(z, nrm), 1.f), 0.5f);
}

__global__ void __launch_bounds__(C_THREADS)
normal_map_bwd_kernel(int64_t P, const float* __restrict__ n_raw, const float* __restrict__ v_out,
                      float* __restrict__ v_raw) {
    const int64_t p = (int64_t)blockIdx.x * C_THREADS + threadIdx.x;
    if (p >= P) return;
    const float x = n_raw[3 * p + 0], y = n_raw[3 * p + 1], z = n_raw[3 * p + 2];
    const float nrm = sqrtf(x * x + y * y + z * z);
    const float inv = 1.f / nrm;
    const float ux = x * inv, uy = y * inv, uz = z * inv;
    const float gx = 0.5f * v_out[3 * p + 0], gy = 0.5f * v_out[3 * p + 1], gz = 0.5f * v_out[3 * p + 2];
    const float dot = ux * gx + uy * gy + uz * gz;
    v_raw[3 * p + 0] = (gx - ux * dot) * inv;
    v_raw[3 * p + 1] = (gy - uy * dot) * inv;
    v_raw[3 * p + 2] = (gz - uz * dot) * inv;
}

// mean over Gaussians of exp(min_k log_scale_k)  (= min_k exp(log_scale_k): exp is monotone)
__global__ void __launch_bounds__(C_THREADS)
flatness_fwd_kernel(int N, const float* __restrict__ log_scales, double* __restrict__ sum, unsigned* __restrict__ ticket,
                    float* __restrict__ out) {
    __shared__ float red[C_THREADS / 32];
    const int n = blockIdx.x * C_THREADS + threadIdx.x;
    float v = 0.f;
    if (n < N) {
        const float s = fminf(fminf(log_scales[3 * (size_t)n], log_scales[3 * (size_t)n + 1]),
                              log_scales[3 * (size_t)n + 2]);
        v = expf(s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = 0.f;
#pragma unroll
        for (int w = 0; w < C_THREADS / 32; ++w) b += red[w];
        atomicAdd(sum, (double)b);
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence();
            *out = (float)(*(volatile double*)sum / (double)N);
        }
    }
}

// d/d log_scale: the mean's 1/N reaches the arg-min axis only (first one on ties), times exp(min)
__global__ void __launch_bounds__(C_THREADS)
flatness_bwd_kernel(int N, const float* __restrict__ log_scales, const float* __restrict__ v_loss,
                    float* __restrict__ v_log_scales) {
    const int n = blockIdx.x * C_THREADS + threadIdx.x;
    if (n >= N) return;
    const float s0 = log_scales[3 * (size_t)n], s1 = log_scales[3 * (size_t)n + 1], s2 = log_scales[3 * (size_t)n + 2];
    int k = 0;
    float s = s0;
    if (s1 < s) { s = s1; k = 1; }
    if (s2 < s) { s = s2; k = 2; }
    const float g = (*v_loss) * expf(s) / (float)N;
    v_log_scales[3 * (size_t)n + 0] = (k == 0) ? g : 0.f;
    v_log_scales[3 * (size_t)n + 1] = (k == 1) ? g : 0.f;
    v_log_scales[3 * (size_t)n + 2] = (k == 2) ? g : 0.f;
}

// main_loss = w_ssim * (1 - ssim) + reg + w_flat * flat, evaluated in the order torch evaluates
// dn_model.py:683-690 / :925 (each product and sum rounded on its own), and its three scalar cotangents.
__global__ void loss_combine_fwd_kernel(const float* __restrict__ ssim, const float* __restrict__ reg,
                                        const float* __restrict__ flat, float w_ssim, float w_flat,
                                        float* __restrict__ out) {
    float v = 0.f;
    if (ssim) v = __fmul_rn(w_ssim, __fsub_rn(1.f, *ssim));
    if (reg) v = __fadd_rn(v, *reg);
    if (flat) v = __fadd_rn(v, __fmul_rn(w_flat, *flat));
    *out = v;
}

__global__ void loss_combine_bwd_kernel(const float* __restrict__ v_out, float w_ssim, float w_flat,
                                        float* __restrict__ v_ssim, float* __restrict__ v_reg,
                                        float* __restrict__ v_flat) {
    const float v = *v_out;
    if (v_ssim) *v_ssim = -(w_ssim * v);
    if (v_reg) *v_reg = v;
    if (v_flat) *v_flat = w_flat * v;
}

}  // namespace

// render[P,4] (RGB premultiplied + expected depth), alpha[P], background[3] (device) ->
// rgb[P,3] = clamp(render.rgb + (1 - alpha) background, 0, 1);  depth[P] = alpha > 0 ? render.w : max_p render.w.
// scratch: 4 bytes (the running maximum).
FSB_API int fsb_compose_rgbd_fwd(int64_t P, const float* render, const float* alpha, const float* background,
                                 float* rgb, float* depth, float* scratch, void* stream) {
    if (P < 0 || !render || !alpha || !background || !rgb || !depth || !scratch) return FSB_E_ARG;
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // 0xffffffff: below every float in both orders atomic_max_float uses (signed -1, unsigned max)
    FSB_CUDA(cudaMemsetAsync(scratch, 0xff, 4, st));
    const int grid = fsb_div_up(P, C_THREADS);
    compose_rgbd_kernel<<<grid, C_THREADS, 0, st>>>(P, (const float4*)render, alpha, background, rgb, scratch);
    FSB_LAUNCH_CHECK();
    compose_depth_kernel<<<grid, C_THREADS, 0, st>>>(P, (const float4*)render, alpha, scratch, depth);
    FSB_LAUNCH_CHECK();
    return 0;
}

// v_rgb[P,3], v_depth[P] (either may be NULL = zero) -> v_render[P,4], v_alpha[P] (overwritten).
FSB_API int fsb_compose_rgbd_bwd(int64_t P, const float* render, const float* alpha, const float* background,
                                 const float* v_rgb, const float* v_depth, float* v_render, float* v_alpha,
                                 void* stream) {
    if (P < 0 || !render || !alpha || !background || !v_render || !v_alpha) return FSB_E_ARG;
    if (P == 0) return 0;
    compose_rgbd_bwd_kernel<<<fsb_div_up(P, C_THREADS), C_THREADS, 0, (cudaStream_t)stream>>>(
        P, (const float4*)render, alpha, background, v_rgb, v_depth, (float4*)v_render, v_alpha);
    FSB_LAUNCH_CHECK();
    return 0;
}

// normals_raw[P,3] -> out[P,3] = (n / ||n|| + 1) / 2
FSB_API int fsb_normal_map_fwd(int64_t P, const float* normals_raw, float* out, void* stream) {
    if (P < 0 || !normals_raw || !out) return FSB_E_ARG;
    if (P == 0) return 0;
    normal_map_kernel<<<fsb_div_up(P, C_THREADS), C_THREADS, 0, (cudaStream_t)stream>>>(P, normals_raw, out);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_normal_map_bwd(int64_t P, const float* normals_raw, const float* v_out, float* v_normals_raw,
                               void* stream) {
    if (P < 0 || !normals_raw || !v_out || !v_normals_raw) return FSB_E_ARG;
    if (P == 0) return 0;
    normal_map_bwd_kernel<<<fsb_div_up(P, C_THREADS), C_THREADS, 0, (cudaStream_t)stream>>>(P, normals_raw, v_out,
                                                                                           v_normals_raw);
    FSB_LAUNCH_CHECK();
    return 0;
}

// out = mean_i min_k exp(log_scales[i,k]).  workspace: 16 bytes.
FSB_API int fsb_flatness_fwd(int N, const float* log_scales, void* workspace, float* out, void* stream) {
    if (N <= 0 || !log_scales || !workspace || !out) return FSB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_CUDA(cudaMemsetAsync(workspace, 0, 16, st));
    flatness_fwd_kernel<<<fsb_div_up(N, C_THREADS), C_THREADS, 0, st>>>(N, log_scales, (double*)workspace,
                                                                         (unsigned*)((char*)workspace + 8), out);
    FSB_LAUNCH_CHECK();
    return 0;
}

// v_loss: DEVICE scalar.  v_log_scales[N,3] overwritten.
FSB_API int fsb_flatness_bwd(int N, const float* log_scales, const float* v_loss, float* v_log_scales, void* stream) {
    if (N <= 0 || !log_scales || !v_loss || !v_log_scales) return FSB_E_ARG;
    flatness_bwd_kernel<<<fsb_div_up(N, C_THREADS), C_THREADS, 0, (cudaStream_t)stream>>>(N, log_scales, v_loss,
                                                                                         v_log_scales);
    FSB_LAUNCH_CHECK();
    return 0;
}

// out = w_ssim * (1 - *ssim) + *reg + w_flat * *flat  (device scalars; a NULL term is left out).
// Replaces the scalar torch launches that assemble main_loss (dn_splatter/dn_model.py:683-690, :925).
FSB_API int fsb_loss_combine_fwd(const float* ssim, const float* reg, const float* flat, float w_ssim, float w_flat,
                                 float* out, void* stream) {
    if (!out) return FSB_E_ARG;
    loss_combine_fwd_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ssim, reg, flat, w_ssim, w_flat, out);
    FSB_LAUNCH_CHECK();
    return 0;
}

// v_out: DEVICE scalar; v_ssim / v_reg / v_flat: device scalars, nullable, overwritten.
FSB_API int fsb_loss_combine_bwd(const float* v_out, float w_ssim, float w_flat, float* v_ssim, float* v_reg,
                                 float* v_flat, void* stream) {
    if (!v_out) return FSB_E_ARG;
    loss_combine_bwd_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(v_out, w_ssim, w_flat, v_ssim, v_reg, v_flat);
    FSB_LAUNCH_CHECK();
    return 0;
}

// ---- 8-bit targets -> float32 in [0, 1] on the device --------------------------------------------------------------
// The step's RGB and normal targets cross PCIe as the 8-bit images they are on disk (20.7 MB instead of 58 MB per 1080p
// view; 8 ranks feeding float32 targets are bound by the host) and become float32 here, with the bits the
// reference produces:
//   recip = 1: x * (1.0f / 255.0f) — what `image.float() / 255.0` evaluates to ON THE DEVICE (splatfacto get_gt_img,
//              torch's CUDA division by a Python scalar multiplies by the fp32 reciprocal);
//   recip = 0: IEEE x / 255.0f — numpy / torch-CPU division (dn_dataset.py:205 `normal_map.astype("float32") / 255.0`).
namespace {
template <bool RECIP>
__device__ __forceinline__ float unit_of(uint32_t b) {
    return RECIP ? (float)b * (1.0f / 255.0f) : __fdiv_rn((float)b, 255.0f);
}

template <bool RECIP>
__global__ void __launch_bounds__(256)
u8_to_unit_float_kernel(int64_t n, const uint8_t* __restrict__ src, float* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n; i += stride) {
        if (i + 16 <= n && ((((uintptr_t)(src + i)) & 15) == 0) && ((((uintptr_t)(dst + i)) & 15) == 0)) {
            const uint4 q = *reinterpret_cast<const uint4*>(src + i);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float4 f;
                f.x = unit_of<RECIP>(w[k] & 0xffu);
                f.y = unit_of<RECIP>((w[k] >> 8) & 0xffu);
                f.z = unit_of<RECIP>((w[k] >> 16) & 0xffu);
                f.w = unit_of<RECIP>(w[k] >> 24);
                *reinterpret_cast<float4*>(dst + i + 4 * k) = f;
            }
        } else {
            for (int64_t j = i; j < n && j < i + 16; ++j) dst[j] = unit_of<RECIP>(src[j]);
        }
    }
}
}  // namespace

// dst[i] = float(src[i]) * (1.0f / 255.0f)  (recip != 0)  or  float(src[i]) / 255.0f  (recip == 0), i in [0, n)
FSB_API int fsb_u8_to_unit_float(int64_t n, const uint8_t* src, float* dst, int recip, void* stream) {
    if (n < 0 || (n > 0 && (!src || !dst))) return FSB_E_ARG;
    if (n == 0) return 0;
    int blocks = fsb_div_up(n, 256 * 16);
    if (blocks > FSB_NUM_SMS * 8) blocks = FSB_NUM_SMS * 8;
    if (recip)
        u8_to_unit_float_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, src, dst);
    else
        u8_to_unit_float_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, src, dst);
    FSB_LAUNCH_CHECK();
    return 0;
}
