// losses.cu — the DN-Splatter depth / normal regulariser (and the RGB L1 term) in one forward and one
// backward kernel.
//
// Replaces the ~40 elementwise / reduction launches behind
//   /root/reference/dn_splatter/dn_model.py:722-736  sensor-depth EdgeAwareLogL1   (losses.py:177-214)
//   /root/reference/dn_splatter/dn_model.py:753-756  TV(depth)                      (losses.py:269-285)
//   /root/reference/dn_splatter/dn_model.py:806      |gt_normal - pred_normal|.mean()
//   /root/reference/dn_splatter/dn_model.py:814-815  TV(pred_normal)
//   splatfacto get_loss_dict                          |gt_rgb - pred_rgb|.mean()     
// loss = l_sensor * EALogL1 + l_smooth * TV(d) + l_nl1 * L1(n) + l_ntv * TV(n) + l_rgb * L1(rgb)
// One thread per pixel; 10 partial sums are block-reduced in fp32 and accumulated in fp64; the last CTA to
// finish turns them into the scalar loss.  The backward kernel is the analytic gradient (abs'(0) = 0 like torch).
// HBM-trivial (a 640x480 frame is 3.7 MB of inputs): the point is launch count, not bandwidth.
#include "common.cuh"

namespace {

constexpr int L_THREADS = 256;
enum { S_EAX = 0, S_CNTX, S_EAY, S_CNTY, S_TVX, S_TVY, S_NL1, S_NTVX, S_NTVY, S_RGB, S_COUNT };

struct LossArgs {
    int H, W;
    const float* depth;        // [H,W]   predicted depth
    const float* sensor;       // [H,W]   sensor depth (ground truth)
    const float* edge_rgb;     // [H,W,3] image whose gradients weight the depth term (clamped at rgb_clamp_min)
    const float* pred_normal;  // [H,W,3]
    const float* gt_normal;    // [H,W,3]
    const float* pred_rgb;     // [H,W,3] nullable
    const float* gt_rgb;       // [H,W,3] nullable
    float depth_tol, rgb_clamp_min;
    float l_sensor, l_smooth, l_nl1, l_ntv, l_rgb;
};

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__device__ __forceinline__ float edge_weight(const float* __restrict__ rgb, int64_t a, int64_t b, float cmin) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) s += fabsf(fmaxf(rgb[3 * a + c], cmin) - fmaxf(rgb[3 * b + c], cmin));
    return expf(-(s / 3.f));
}

__global__ void __launch_bounds__(L_THREADS)
dn_loss_fwd_kernel(LossArgs a, double* __restrict__ sums, unsigned* __restrict__ ticket, float* __restrict__ loss) {
    __shared__ float red[S_COUNT][L_THREADS / 32];
    const int64_t P = (int64_t)a.H * a.W;
    float s[S_COUNT];
#pragma unroll
    for (int k = 0; k < S_COUNT; ++k) s[k] = 0.f;
    // grid-stride: a CTA per 256 pixels made 8100 CTAs x 10 fp64 atomics on the same ten addresses at 1080p
    for (int64_t p = (int64_t)blockIdx.x * L_THREADS + threadIdx.x; p < P; p += (int64_t)gridDim.x * L_THREADS) {
        const int i = (int)(p / a.W), j = (int)(p - (int64_t)i * a.W);
        const bool hx = j < a.W - 1, hy = i < a.H - 1;
        const float d = a.depth[p];
        if (a.l_sensor != 0.f) {
            const float g = a.sensor[p];
            if (g > a.depth_tol) {
                const float ll = logf(1.f + fabsf(d - g));
                if (hx) { s[S_EAX] += edge_weight(a.edge_rgb, p, p + 1, a.rgb_clamp_min) * ll; s[S_CNTX] += 1.f; }
                if (hy) { s[S_EAY] += edge_weight(a.edge_rgb, p, p + a.W, a.rgb_clamp_min) * ll; s[S_CNTY] += 1.f; }
            }
        }
        if (a.l_smooth != 0.f) {
            if (hx) s[S_TVX] += fabsf(d - a.depth[p + 1]);
            if (hy) s[S_TVY] += fabsf(d - a.depth[p + a.W]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float n = a.pred_normal ? a.pred_normal[3 * p + c] : 0.f;
            if (a.l_nl1 != 0.f) s[S_NL1] += fabsf(a.gt_normal[3 * p + c] - n);
            if (a.l_ntv != 0.f) {
                if (hx) s[S_NTVX] += fabsf(n - a.pred_normal[3 * (p + 1) + c]);
                if (hy) s[S_NTVY] += fabsf(n - a.pred_normal[3 * (p + a.W) + c]);
            }
            if (a.l_rgb != 0.f) s[S_RGB] += fabsf(a.gt_rgb[3 * p + c] - a.pred_rgb[3 * p + c]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < S_COUNT; ++k) {
        float v = s[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < S_COUNT) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < L_THREADS / 32; ++w) v += red[threadIdx.x][w];
        if (v != 0.f) atomicAdd(sums + threadIdx.x, (double)v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence();
            volatile double* S = sums;
            const double H = a.H, W = a.W;
            double L = 0.0;
            // masked means: an empty selection is 0/0 = NaN in the reference as well
            if (a.l_sensor != 0.f) L += a.l_sensor * (S[S_EAX] / S[S_CNTX] + S[S_EAY] / S[S_CNTY]);
            if (a.l_smooth != 0.f) L += a.l_smooth * (S[S_TVX] / (H * (W - 1)) + S[S_TVY] / ((H - 1) * W));
            if (a.l_nl1 != 0.f) L += a.l_nl1 * (S[S_NL1] / (3.0 * H * W));
            if (a.l_ntv != 0.f) L += a.l_ntv * (S[S_NTVX] / (3.0 * H * (W - 1)) + S[S_NTVY] / (3.0 * (H - 1) * W));
            if (a.l_rgb != 0.f) L += a.l_rgb * (S[S_RGB] / (3.0 * H * W));
            *loss = (float)L;
        }
    }
}

__global__ void __launch_bounds__(L_THREADS)
dn_loss_bwd_kernel(LossArgs a, const double* __restrict__ sums, const float* __restrict__ v_loss,
                   float* __restrict__ v_depth, float* __restrict__ v_normal, float* __restrict__ v_rgb) {
    const int64_t P = (int64_t)a.H * a.W;
    const int64_t p = (int64_t)blockIdx.x * L_THREADS + threadIdx.x;
    if (p >= P) return;
    const float vl = *v_loss;
    const int i = (int)(p / a.W), j = (int)(p - (int64_t)i * a.W);
    const bool hx = j < a.W - 1, hy = i < a.H - 1, lx = j > 0, ly = i > 0;
    const float H = (float)a.H, W = (float)a.W;
    if (v_depth) {
        const float d = a.depth[p];
        float g = 0.f;
        if (a.l_sensor != 0.f) {
            const float gt = a.sensor[p];
            if (gt > a.depth_tol) {
                const float diff = d - gt;
                const float dl = sgn(diff) / (1.f + fabsf(diff));
                float w = 0.f;
                if (hx) w += edge_weight(a.edge_rgb, p, p + 1, a.rgb_clamp_min) / (float)sums[S_CNTX];
                if (hy) w += edge_weight(a.edge_rgb, p, p + a.W, a.rgb_clamp_min) / (float)sums[S_CNTY];
                g += a.l_sensor * w * dl;
            }
        }
        if (a.l_smooth != 0.f) {
            float tx = 0.f, ty = 0.f;
            if (hx) tx += sgn(d - a.depth[p + 1]);
            if (lx) tx -= sgn(a.depth[p - 1] - d);
            if (hy) ty += sgn(d - a.depth[p + a.W]);
            if (ly) ty -= sgn(a.depth[p - a.W] - d);
            g += a.l_smooth * (tx / (H * (W - 1.f)) + ty / ((H - 1.f) * W));
        }
        v_depth[p] = g * vl;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (v_normal) {
            const float n = a.pred_normal[3 * p + c];
            float g = 0.f;
            if (a.l_nl1 != 0.f) g += -a.l_nl1 * sgn(a.gt_normal[3 * p + c] - n) / (3.f * H * W);
            if (a.l_ntv != 0.f) {
                float tx = 0.f, ty = 0.f;
                if (hx) tx += sgn(n - a.pred_normal[3 * (p + 1) + c]);
                if (lx) tx -= sgn(a.pred_normal[3 * (p - 1) + c] - n);
                if (hy) ty += sgn(n - a.pred_normal[3 * (p + a.W) + c]);
                if (ly) ty -= sgn(a.pred_normal[3 * (p - a.W) + c] - n);
                g += a.l_ntv * (tx / (3.f * H * (W - 1.f)) + ty / (3.f * (H - 1.f) * W));
            }
            v_normal[3 * p + c] = g * vl;
        }
        if (v_rgb) {
            float g = 0.f;
            if (a.l_rgb != 0.f) g = -a.l_rgb * sgn(a.gt_rgb[3 * p + c] - a.pred_rgb[3 * p + c]) / (3.f * H * W);
            v_rgb[3 * p + c] = g * vl;
        }
    }
}

int fill_args(LossArgs& a, int H, int W, const float* depth, const float* sensor, const float* edge_rgb,
              const float* pred_normal, const float* gt_normal, const float* pred_rgb, const float* gt_rgb,
              float depth_tol, float rgb_clamp_min, float l_sensor, float l_smooth, float l_nl1, float l_ntv,
              float l_rgb) {
    if (H < 2 || W < 2) return FSB_E_ARG;
    if ((l_sensor != 0.f || l_smooth != 0.f) && !depth) return FSB_E_ARG;
    if (l_sensor != 0.f && (!sensor || !edge_rgb)) return FSB_E_ARG;
    if ((l_nl1 != 0.f || l_ntv != 0.f) && !pred_normal) return FSB_E_ARG;
    if (l_nl1 != 0.f && !gt_normal) return FSB_E_ARG;
    if (l_rgb != 0.f && (!pred_rgb || !gt_rgb)) return FSB_E_ARG;
    a.H = H; a.W = W; a.depth = depth; a.sensor = sensor; a.edge_rgb = edge_rgb; a.pred_normal = pred_normal;
    a.gt_normal = gt_normal; a.pred_rgb = pred_rgb; a.gt_rgb = gt_rgb; a.depth_tol = depth_tol;
    a.rgb_clamp_min = rgb_clamp_min; a.l_sensor = l_sensor; a.l_smooth = l_smooth; a.l_nl1 = l_nl1; a.l_ntv = l_ntv;
    a.l_rgb = l_rgb;
    return 0;
}

}  // namespace

// bytes of the reduction workspace (10 fp64 sums + ticket); fwd fills it, bwd reads the mask counts from it
FSB_API size_t fsb_dn_loss_workspace(void) { return 16 * sizeof(double); }

FSB_API int fsb_dn_loss_fwd(int H, int W, const float* depth, const float* sensor, const float* edge_rgb,
                            const float* pred_normal, const float* gt_normal, const float* pred_rgb,
                            const float* gt_rgb, float depth_tol, float rgb_clamp_min, float l_sensor,
                            float l_smooth, float l_nl1, float l_ntv, float l_rgb, void* workspace, float* loss_out,
                            void* stream) {
    LossArgs a;
    int rc = fill_args(a, H, W, depth, sensor, edge_rgb, pred_normal, gt_normal, pred_rgb, gt_rgb, depth_tol,
                       rgb_clamp_min, l_sensor, l_smooth, l_nl1, l_ntv, l_rgb);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_CUDA(cudaMemsetAsync(workspace, 0, fsb_dn_loss_workspace(), st));
    double* sums = (double*)workspace;
    unsigned* ticket = (unsigned*)(sums + 12);
    int blocks = fsb_div_up((int64_t)H * W, L_THREADS);
    if (blocks > FSB_NUM_SMS * 8) blocks = FSB_NUM_SMS * 8;
    dn_loss_fwd_kernel<<<blocks, L_THREADS, 0, st>>>(a, sums, ticket, loss_out);
    FSB_LAUNCH_CHECK();
    return 0;
}

// v_loss: DEVICE scalar (dL/dloss).  v_depth[H,W], v_normal[H,W,3], v_rgb[H,W,3]: nullable, overwritten.
FSB_API int fsb_dn_loss_bwd(int H, int W, const float* depth, const float* sensor, const float* edge_rgb,
                            const float* pred_normal, const float* gt_normal, const float* pred_rgb,
                            const float* gt_rgb, float depth_tol, float rgb_clamp_min, float l_sensor,
                            float l_smooth, float l_nl1, float l_ntv, float l_rgb, const void* workspace,
                            const float* v_loss, float* v_depth, float* v_normal, float* v_rgb, void* stream) {
    LossArgs a;
    int rc = fill_args(a, H, W, depth, sensor, edge_rgb, pred_normal, gt_normal, pred_rgb, gt_rgb, depth_tol,
                       rgb_clamp_min, l_sensor, l_smooth, l_nl1, l_ntv, l_rgb);
    if (rc) return rc;
    if (!v_loss) return FSB_E_ARG;
    dn_loss_bwd_kernel<<<fsb_div_up((int64_t)H * W, L_THREADS), L_THREADS, 0, (cudaStream_t)stream>>>(
        a, (const double*)workspace, v_loss, v_depth, v_normal, v_rgb);
    FSB_LAUNCH_CHECK();
    return 0;
}
