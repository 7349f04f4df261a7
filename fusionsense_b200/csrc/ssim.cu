// ssim.cu — structural similarity (gaussian 11x11, sigma 1.5) forward + analytic backward in two launches.
//
// Replaces the SSIM half of the base splatfacto loss, `1 - self.ssim(gt, pred)` with
// self.ssim = StructuralSimilarityIndexMeasure(data_range=1.0, kernel_size=11)
// (/root/reference/dn_splatter/dn_model.py:244, consumed through super().get_loss_dict at :683), which in torch
// is 2 reflection pads + a 15-channel depthwise conv + ~20 elementwise launches forward and as many backward
// (ncu round 1: 0.74 ms of a 3.4 ms step for the two depthwise-conv kernels alone).
//
// torchmetrics pads both images by 5 px (reflect), convolves, and then CROPS the 5 px border of the SSIM map
// before averaging, so every window that survives lies fully inside the image: the result is the mean, over the
// (H-10) x (W-10) interior and the channels, of the valid-convolution SSIM map.  The padding never matters.
//
//   mu_x = w*x, mu_y = w*y, e_xx = w*x^2, e_yy = w*y^2, e_xy = w*xy          (w = separable gaussian)
//   S = (2 mu_x mu_y + c1)(2 s_xy + c2) / ((mu_x^2 + mu_y^2 + c1)(s_x + s_y + c2)),  s_x = e_xx - mu_x^2, ...
//
// forward   one CTA per 32x32 block of interior pixels; per channel a 42x42 tile of x and y is staged in shared
//           memory, 5 moments are convolved horizontally then vertically (register-blocked, see below); S is block-reduced (fp32) and
//           accumulated in fp64, the last CTA writes the mean.  The three partials dS/dmu_x, dS/de_xx, dS/de_xy
//           are stored per interior pixel for the backward.
// backward  dL/dx(p) = v * [ (w * dS/dmu_x)(p) + 2 x(p) (w * dS/de_xx)(p) + y(p) (w * dS/de_xy)(p) ] / count,
//           the same separable convolution applied to the stored maps (zero outside the interior).
// Images are [H, W, C] row-major (the layout the rasteriser writes), C <= 4.  HBM-trivial; launch-count bound.
#include "common.cuh"

namespace {

// Round 2: register-blocked separable filter.  One CTA = 32 x 32 outputs, 256 threads; the horizontal pass gives every
// thread eight adjacent outputs of one staged row (18 loads per image instead of 8 x 11), the vertical pass four
// vertically adjacent outputs of one column (14 loads per moment instead of 4 x 11): ~30 shared-memory loads per output
// and channel against ~90 for the one-output-per-thread form (r02b: 0.145 ms forward at 1080p, LSU bound).
constexpr int TS = 32;           // output tile edge
constexpr int KS = 11;           // gaussian taps
constexpr int TL = TS + KS - 1;  // staged tile edge (42)
constexpr int TLP = TL + 1;      // padded row length (43: conflict-free for the 8-wide horizontal items)
constexpr int HP = TS + 1;       // padded row length of the horizontally filtered rows
constexpr int S_THREADS = 256;
constexpr int HC = 8;            // horizontal outputs per work item
constexpr int VR = 4;            // vertical outputs per thread

struct SsimArgs {
    int H, W, C;
    float c1, c2;
    float w[KS];
};

// horizontal then vertical 11-tap pass over NQ staged quantities; out[v][q] = filtered quantity q of this thread's
// v-th output pixel (tile row VR * (tid / 32) + v, tile column tid % 32).  `src(r, i, v)` reads the NQ staged
// quantities at tile row r, tile column i.
template <int NQ, typename Src>
__device__ __forceinline__ void separable(const SsimArgs& a, Src src, float (*hbuf)[TL][HP], float (&out)[VR][NQ]) {
    for (int e = threadIdx.x; e < TL * (TS / HC); e += S_THREADS) {
        const int r = e / (TS / HC), j0 = (e - r * (TS / HC)) * HC;
        float acc[HC][NQ];
#pragma unroll
        for (int o = 0; o < HC; ++o)
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[o][q] = 0.f;
#pragma unroll
        for (int i = 0; i < HC + KS - 1; ++i) {
            float v[NQ];
            src(r, j0 + i, v);
#pragma unroll
            for (int o = 0; o < HC; ++o) {
                const int t = i - o;  // tap index of input i for output o (compile-time after unrolling)
                if (t >= 0 && t < KS) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) acc[o][q] = fmaf(a.w[t], v[q], acc[o][q]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < HC; ++o)
#pragma unroll
            for (int q = 0; q < NQ; ++q) hbuf[q][r][j0 + o] = acc[o][q];
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, y0 = (threadIdx.x >> 5) * VR;
#pragma unroll
    for (int v = 0; v < VR; ++v)
#pragma unroll
        for (int q = 0; q < NQ; ++q) out[v][q] = 0.f;
#pragma unroll
    for (int i = 0; i < VR + KS - 1; ++i) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float h = hbuf[q][y0 + i][tx];
#pragma unroll
            for (int v = 0; v < VR; ++v) {
                const int t = i - v;
                if (t >= 0 && t < KS) out[v][q] = fmaf(a.w[t], h, out[v][q]);
            }
        }
    }
}

__global__ void __launch_bounds__(S_THREADS)
ssim_fwd_kernel(SsimArgs a, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ d_mu,
                float* __restrict__ d_xx, float* __restrict__ d_xy, double* __restrict__ sum,
                unsigned* __restrict__ ticket, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char ssim_smem[];
    float (*sx)[TLP] = reinterpret_cast<float (*)[TLP]>(ssim_smem);
    float (*sy)[TLP] = reinterpret_cast<float (*)[TLP]>(ssim_smem + sizeof(float) * TL * TLP);
    float (*hbuf)[TL][HP] = reinterpret_cast<float (*)[TL][HP]>(ssim_smem + 2 * sizeof(float) * TL * TLP);
    __shared__ float red[S_THREADS / 32];
    const int Hi = a.H - (KS - 1), Wi = a.W - (KS - 1);
    const int oy0 = blockIdx.y * TS, ox0 = blockIdx.x * TS;  // interior coordinates == image coordinates of the window start
    const int tx = threadIdx.x & 31, y0 = (threadIdx.x >> 5) * VR;
    const int qx = ox0 + tx;
    float local = 0.f;
    for (int c = 0; c < a.C; ++c) {
        __syncthreads();  // previous channel's hbuf / tiles are no longer read
        for (int e = threadIdx.x; e < TL * TL; e += S_THREADS) {
            const int r = e / TL, i = e - r * TL;
            const int iy = oy0 + r, ix = ox0 + i;
            float vx = 0.f, vy = 0.f;
            if (iy < a.H && ix < a.W) {
                const size_t o = ((size_t)iy * a.W + ix) * a.C + c;
                vx = x[o];
                vy = y[o];
            }
            sx[r][i] = vx;
            sy[r][i] = vy;
        }
        __syncthreads();
        float m[VR][5];
        separable<5>(a, [&](int r, int i, float (&v)[5]) {
            const float p = sx[r][i], t = sy[r][i];
            v[0] = p; v[1] = t; v[2] = p * p; v[3] = t * t; v[4] = p * t;
        }, hbuf, m);
#pragma unroll
        for (int v = 0; v < VR; ++v) {
            const int qy = oy0 + y0 + v;
            if (qy < Hi && qx < Wi) {
                const float mu_x = m[v][0], mu_y = m[v][1];
                const float s_x = m[v][2] - mu_x * mu_x, s_y = m[v][3] - mu_y * mu_y, s_xy = m[v][4] - mu_x * mu_y;
                const float num1 = 2.f * mu_x * mu_y + a.c1, num2 = 2.f * s_xy + a.c2;
                const float den1 = mu_x * mu_x + mu_y * mu_y + a.c1, den2 = s_x + s_y + a.c2;
                const float inv = 1.f / (den1 * den2);
                const float S = num1 * num2 * inv;
                local += S;
                if (d_mu) {
                    const float dS_dnum1 = num2 * inv, dS_dnum2 = num1 * inv;
                    const float dS_dden1 = -S / den1, dS_dden2 = -S / den2;
                    const size_t o = ((size_t)qy * Wi + qx) * a.C + c;
                    d_mu[o] = 2.f * (mu_y * (dS_dnum1 - dS_dnum2) + mu_x * (dS_dden1 - dS_dden2));
                    d_xx[o] = dS_dden2;
                    d_xy[o] = 2.f * dS_dnum2;
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0) red[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < S_THREADS / 32; ++w) v += red[w];
        atomicAdd(sum, (double)v);
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x * gridDim.y - 1) {
            __threadfence();
            *out = (float)(*((volatile double*)sum) / ((double)Hi * (double)Wi * (double)a.C));
        }
    }
}

__global__ void __launch_bounds__(S_THREADS)
ssim_bwd_kernel(SsimArgs a, const float* __restrict__ x, const float* __restrict__ y,
                const float* __restrict__ d_mu, const float* __restrict__ d_xx, const float* __restrict__ d_xy,
                const float* __restrict__ v_out, float* __restrict__ v_x) {
    extern __shared__ __align__(16) unsigned char ssim_smem[];
    float (*s0)[TLP] = reinterpret_cast<float (*)[TLP]>(ssim_smem);
    float (*s1)[TLP] = reinterpret_cast<float (*)[TLP]>(ssim_smem + sizeof(float) * TL * TLP);
    float (*s2)[TLP] = reinterpret_cast<float (*)[TLP]>(ssim_smem + 2 * sizeof(float) * TL * TLP);
    float (*hbuf)[TL][HP] = reinterpret_cast<float (*)[TL][HP]>(ssim_smem + 3 * sizeof(float) * TL * TLP);
    const int Hi = a.H - (KS - 1), Wi = a.W - (KS - 1);
    const int py0 = blockIdx.y * TS, px0 = blockIdx.x * TS;
    const int tx = threadIdx.x & 31, y0 = (threadIdx.x >> 5) * VR;
    const int px = px0 + tx;
    const float scale = *v_out / ((float)Hi * (float)Wi * (float)a.C);
    for (int c = 0; c < a.C; ++c) {
        __syncthreads();
        for (int e = threadIdx.x; e < TL * TL; e += S_THREADS) {
            const int r = e / TL, i = e - r * TL;
            const int qy = py0 - (KS - 1) + r, qx = px0 - (KS - 1) + i;  // interior coordinates
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (qy >= 0 && qy < Hi && qx >= 0 && qx < Wi) {
                const size_t o = ((size_t)qy * Wi + qx) * a.C + c;
                v0 = d_mu[o]; v1 = d_xx[o]; v2 = d_xy[o];
            }
            s0[r][i] = v0; s1[r][i] = v1; s2[r][i] = v2;
        }
        __syncthreads();
        float g[VR][3];
        separable<3>(a, [&](int r, int i, float (&v)[3]) { v[0] = s0[r][i]; v[1] = s1[r][i]; v[2] = s2[r][i]; },
                     hbuf, g);
#pragma unroll
        for (int v = 0; v < VR; ++v) {
            const int py = py0 + y0 + v;
            if (py < a.H && px < a.W) {
                const size_t o = ((size_t)py * a.W + px) * a.C + c;
                v_x[o] = scale * (g[v][0] + 2.f * x[o] * g[v][1] + y[o] * g[v][2]);
            }
        }
    }
}

constexpr size_t FWD_SMEM = sizeof(float) * (2 * TL * TLP + 5 * TL * HP);
constexpr size_t BWD_SMEM = sizeof(float) * (3 * TL * TLP + 3 * TL * HP);

int fill(SsimArgs& a, int H, int W, int C, float data_range, float k1, float k2, const float* taps) {
    if (H < KS || W < KS || C < 1 || C > 4 || !taps) return FSB_E_ARG;
    a.H = H; a.W = W; a.C = C;
    a.c1 = (k1 * data_range) * (k1 * data_range);
    a.c2 = (k2 * data_range) * (k2 * data_range);
    for (int t = 0; t < KS; ++t) a.w[t] = taps[t];
    return 0;
}

}  // namespace

FSB_API int fsb_ssim_taps(void) { return KS; }

// bytes of the reduction workspace (fp64 sum + ticket)
FSB_API size_t fsb_ssim_workspace(void) { return 16; }

// x, y: [H,W,C] fp32 images; taps: HOST array of the 11 normalised 1-D gaussian weights (the caller computes them
// the way torchmetrics does).  ssim_out: device scalar = mean SSIM over the (H-10)x(W-10) interior and channels.
// d_mu / d_xx / d_xy: [(H-10),(W-10),C] partial derivatives w.r.t. x's window moments, nullable (all or none).
FSB_API int fsb_ssim_fwd(int H, int W, int C, const float* x, const float* y, float data_range, float k1, float k2,
                         const float* taps, void* workspace, float* ssim_out, float* d_mu, float* d_xx, float* d_xy,
                         void* stream) {
    SsimArgs a;
    int rc = fill(a, H, W, C, data_range, k1, k2, taps);
    if (rc) return rc;
    if (!x || !y || !workspace || !ssim_out) return FSB_E_ARG;
    if ((d_mu || d_xx || d_xy) && !(d_mu && d_xx && d_xy)) return FSB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_CUDA(cudaMemsetAsync(workspace, 0, fsb_ssim_workspace(), st));
    dim3 grid(fsb_div_up(W - (KS - 1), TS), fsb_div_up(H - (KS - 1), TS));
    ssim_fwd_kernel<<<grid, S_THREADS, FWD_SMEM, st>>>(a, x, y, d_mu, d_xx, d_xy, (double*)workspace,
                                                (unsigned*)((char*)workspace + 8), ssim_out);
    FSB_LAUNCH_CHECK();
    return 0;
}

// v_out: DEVICE scalar dL/dssim.  v_x[H,W,C]: overwritten with dL/dx.
FSB_API int fsb_ssim_bwd(int H, int W, int C, const float* x, const float* y, float data_range, float k1, float k2,
                         const float* taps, const float* d_mu, const float* d_xx, const float* d_xy,
                         const float* v_out, float* v_x, void* stream) {
    SsimArgs a;
    int rc = fill(a, H, W, C, data_range, k1, k2, taps);
    if (rc) return rc;
    if (!x || !y || !d_mu || !d_xx || !d_xy || !v_out || !v_x) return FSB_E_ARG;
    dim3 grid(fsb_div_up(W, TS), fsb_div_up(H, TS));
    ssim_bwd_kernel<<<grid, S_THREADS, BWD_SMEM, (cudaStream_t)stream>>>(a, x, y, d_mu, d_xx, d_xy, v_out, v_x);
    FSB_LAUNCH_CHECK();
    return 0;
}
