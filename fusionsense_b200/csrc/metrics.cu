// metrics.cu — the per-step image metrics of DNSplatterModel.get_metrics_dict in one launch, results left on the device.
//
// Replaces, for /root/reference/dn_splatter/dn_model.py:962-1000 (get_metrics_dict, called every training iteration
// by nerfstudio's Trainer), the torchmetrics PSNR + torch MSELoss on the RGB image and dn_splatter/metrics.py:111-150
// (DepthMetrics: abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 over the pixels with gt > tolerance) — ~45 elementwise /
// masked-select / reduction launches and eleven float() host syncs in the reference.  SSIM is fsb_ssim_fwd.
// One thread per pixel, ten partial sums block-reduced in fp32 and accumulated in fp64; the last CTA to finish turns
// them into the metric vector.  HBM-trivial; the point is launch count and the absence of host syncs.
#include "common.cuh"

namespace {

constexpr int M_THREADS = 256;
enum { M_SE = 0, M_CNT, M_ABSREL, M_SQREL, M_SQ, M_LOG, M_LOGCNT, M_A1, M_A2, M_A3, M_COUNT };

__global__ void __launch_bounds__(M_THREADS)
image_metrics_kernel(int64_t P, const float* __restrict__ pred_rgb, const float* __restrict__ gt_rgb,
                     const float* __restrict__ pred_depth, const float* __restrict__ gt_depth, float depth_tol,
                     double* __restrict__ sums, unsigned* __restrict__ ticket, float* __restrict__ out) {
    __shared__ float red[M_COUNT][M_THREADS / 32];
    __shared__ bool s_last;
    const int64_t p = (int64_t)blockIdx.x * M_THREADS + threadIdx.x;
    float s[M_COUNT];
#pragma unroll
    for (int k = 0; k < M_COUNT; ++k) s[k] = 0.f;
    if (p < P) {
        if (pred_rgb != nullptr) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float d = pred_rgb[3 * p + c] - gt_rgb[3 * p + c];
                s[M_SE] += d * d;
            }
        }
        if (pred_depth != nullptr) {
            const float g = gt_depth[p], d = pred_depth[p];
            if (g > depth_tol) {
                s[M_CNT] = 1.f;
                const float diff = g - d;
                s[M_ABSREL] = fabsf(diff) / g;
                s[M_SQREL] = diff * diff / g;
                s[M_SQ] = diff * diff;
                const float l = fabsf(logf(g) - logf(d));  // torch.sqrt(x ** 2) of the reference, then nanmean
                if (l == l) { s[M_LOG] = l; s[M_LOGCNT] = 1.f; }
                const float th = fmaxf(g / d, d / g);
                s[M_A1] = th < 1.25f ? 1.f : 0.f;
                s[M_A2] = th < 1.25f * 1.25f ? 1.f : 0.f;
                s[M_A3] = th < 1.25f * 1.25f * 1.25f ? 1.f : 0.f;
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < M_COUNT; ++k) {
        float v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < M_COUNT) {
        float v = 0.f;
        for (int w = 0; w < M_THREADS / 32; ++w) v += red[threadIdx.x][w];
        if (v != 0.f) atomicAdd(sums + threadIdx.x, (double)v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    volatile double* S = sums;
    const double n_rgb = 3.0 * (double)P, cnt = S[M_CNT];
    const double mse = S[M_SE] / n_rgb;
    out[0] = (float)mse;                                            // rgb_mse
    out[1] = (float)(10.0 * log10(1.0 / mse));                      // rgb_psnr (data_range 1)
    out[2] = (float)(S[M_ABSREL] / cnt);                            // depth_abs_rel
    out[3] = (float)(S[M_SQREL] / cnt);                             // depth_sq_rel
    out[4] = (float)sqrt(S[M_SQ] / cnt);                            // depth_rmse
    out[5] = (float)(S[M_LOG] / S[M_LOGCNT]);                       // depth_rmse_log
    out[6] = (float)(S[M_A1] / cnt);
    out[7] = (float)(S[M_A2] / cnt);
    out[8] = (float)(S[M_A3] / cnt);
    out[9] = (float)cnt;
    for (int k = 0; k < M_COUNT; ++k) sums[k] = 0.0;  // ready for the next launch (and the next graph replay)
    *ticket = 0u;
}

}  // namespace

FSB_API size_t fsb_image_metrics_workspace(void) { return 16 * sizeof(double); }

// out[10] fp32 (device) = { rgb_mse, rgb_psnr, depth_abs_rel, depth_sq_rel, depth_rmse, depth_rmse_log, a1, a2, a3,
// number of valid depth pixels }.  pred_rgb / gt_rgb [H,W,3] and pred_depth / gt_depth [H,W] may each pair be NULL
// (its metrics are then NaN / inf).  workspace: fsb_image_metrics_workspace() bytes, zero-filled ONCE by the caller
// (the kernel leaves it zeroed).
FSB_API int fsb_image_metrics(int H, int W, const float* pred_rgb, const float* gt_rgb, const float* pred_depth,
                              const float* gt_depth, float depth_tol, void* workspace, float* out, void* stream) {
    if (H <= 0 || W <= 0 || !workspace || !out) return FSB_E_ARG;
    if ((pred_rgb == nullptr) != (gt_rgb == nullptr) || (pred_depth == nullptr) != (gt_depth == nullptr)) return FSB_E_ARG;
    const int64_t P = (int64_t)H * W;
    double* sums = (double*)workspace;
    unsigned* ticket = (unsigned*)(sums + 12);
    image_metrics_kernel<<<fsb_div_up(P, M_THREADS), M_THREADS, 0, (cudaStream_t)stream>>>(
        P, pred_rgb, gt_rgb, pred_depth, gt_depth, depth_tol, sums, ticket, out);
    FSB_LAUNCH_CHECK();
    return 0;
}
