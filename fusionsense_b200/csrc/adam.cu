// adam.cu — multi-tensor Adam: every Gaussian parameter group updated by ONE launch.
//
// Replaces the eight torch.optim.Adam.step() calls nerfstudio issues per iteration for the optimizers
// declared at /root/reference/dn_splatter/dn_config.py:36-75 (lr per group, eps = 1e-15, betas (0.9, 0.999),
// no weight decay, no amsgrad).  Arithmetic order follows torch.optim.Adam's single-tensor path:
//   m += (g - m) * (1 - b1) ; v = v * b2 + (1 - b2) * g * g ;
//   denom = sqrt(v) / sqrt(1 - b2^t) + eps ; p -= (lr / (1 - b1^t)) * m / denom
// HBM-bound: 28 B per element (read p, g, m, v; write p, m, v).
#include "common.cuh"

#define FSB_ADAM_MAX_TENSORS 8

struct FsbAdamArgs {
    float* p[FSB_ADAM_MAX_TENSORS];
    const float* g[FSB_ADAM_MAX_TENSORS];
    float* m[FSB_ADAM_MAX_TENSORS];
    float* v[FSB_ADAM_MAX_TENSORS];
    long long n[FSB_ADAM_MAX_TENSORS];
    int block_start[FSB_ADAM_MAX_TENSORS + 1];  // first CTA of each tensor
    float step_size[FSB_ADAM_MAX_TENSORS];      // lr / (1 - b1^t)
    float inv_bc2_sqrt[FSB_ADAM_MAX_TENSORS];   // 1 / sqrt(1 - b2^t)  (applied as a division, see below)
    float bc2_sqrt[FSB_ADAM_MAX_TENSORS];
};

namespace {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC = 4;
constexpr int ADAM_PER_BLOCK = ADAM_THREADS * ADAM_VEC * 4;  // 4096 elements per CTA

// omb1 / omb2 = 1 - beta computed in double on the host and rounded once, like torch's Python floats
// (1.f - 0.999f is 4.7e-5 off 0.001f, which shows up in exp_avg_sq after a few steps).
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float omb1, float b2, float omb2,
                                         float eps, float step_size, float bc2_sqrt) {
    m = m + (g - m) * omb1;
    v = v * b2 + omb2 * g * g;
    float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_multi_kernel(FsbAdamArgs a, int n_tensors, float omb1, float b2, float omb2, float eps,
                  const float* __restrict__ hyper_dev, const int32_t* __restrict__ skip_flag) {
    if (skip_flag != nullptr && *skip_flag != 0) return;
    int t = 0;
#pragma unroll
    for (int i = 1; i < FSB_ADAM_MAX_TENSORS; ++i)
        if (i < n_tensors && (int)blockIdx.x >= a.block_start[i]) t = i;
    const long long n = a.n[t];
    const long long base = (long long)(blockIdx.x - a.block_start[t]) * ADAM_PER_BLOCK;
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    const float ss = hyper_dev ? hyper_dev[t] : a.step_size[t];
    const float bc2s = hyper_dev ? hyper_dev[FSB_ADAM_MAX_TENSORS + t] : a.bc2_sqrt[t];
    const bool aligned = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        long long i = base + ((long long)it * ADAM_THREADS + threadIdx.x) * ADAM_VEC;
        if (i >= n) break;
        if (aligned && i + ADAM_VEC <= n) {
            float4 P = *reinterpret_cast<float4*>(p + i);
            float4 G = *reinterpret_cast<const float4*>(g + i);
            float4 M = *reinterpret_cast<float4*>(m + i);
            float4 V = *reinterpret_cast<float4*>(v + i);
            adam_one(P.x, G.x, M.x, V.x, omb1, b2, omb2, eps, ss, bc2s);
            adam_one(P.y, G.y, M.y, V.y, omb1, b2, omb2, eps, ss, bc2s);
            adam_one(P.z, G.z, M.z, V.z, omb1, b2, omb2, eps, ss, bc2s);
            adam_one(P.w, G.w, M.w, V.w, omb1, b2, omb2, eps, ss, bc2s);
            *reinterpret_cast<float4*>(p + i) = P;
            *reinterpret_cast<float4*>(m + i) = M;
            *reinterpret_cast<float4*>(v + i) = V;
        } else {
            for (int k = 0; k < ADAM_VEC && i + k < n; ++k) {
                float P = p[i + k], M = m[i + k], V = v[i + k];
                adam_one(P, g[i + k], M, V, omb1, b2, omb2, eps, ss, bc2s);
                p[i + k] = P; m[i + k] = M; v[i + k] = V;
            }
        }
    }
}

}  // namespace

FSB_API int fsb_adam_max_tensors(void) { return FSB_ADAM_MAX_TENSORS; }

// All array arguments are HOST arrays of length n_tensors; the pointers inside p/g/m/v are device pointers.
// step[i] is the 1-based step count of tensor i AFTER this update (torch's state["step"] post-increment).
FSB_API int fsb_adam_multi(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                           const int64_t* n, const double* lr, const int64_t* step, double beta1, double beta2,
                           double eps, void* stream) {
    if (n_tensors <= 0 || n_tensors > FSB_ADAM_MAX_TENSORS) return FSB_E_ARG;
    FsbAdamArgs a;
    int blocks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (n[i] < 0 || step[i] < 1) return FSB_E_ARG;
        a.p[i] = p[i]; a.g[i] = g[i]; a.m[i] = m[i]; a.v[i] = v[i]; a.n[i] = n[i];
        a.block_start[i] = blocks;
        blocks += (int)((n[i] + ADAM_PER_BLOCK - 1) / ADAM_PER_BLOCK);
        // bias corrections in double like Python's floats in torch.optim.Adam, then rounded once
        double bc1 = 1.0 - pow(beta1, (double)step[i]);
        double bc2 = 1.0 - pow(beta2, (double)step[i]);
        a.step_size[i] = (float)(lr[i] / bc1);
        a.bc2_sqrt[i] = (float)sqrt(bc2);
        a.inv_bc2_sqrt[i] = (float)(1.0 / sqrt(bc2));
    }
    a.block_start[n_tensors] = blocks;
    for (int i = n_tensors; i < FSB_ADAM_MAX_TENSORS; ++i) {
        a.p[i] = nullptr; a.g[i] = nullptr; a.m[i] = nullptr; a.v[i] = nullptr; a.n[i] = 0;
        a.block_start[i + 1] = blocks;
        a.step_size[i] = 0.f; a.bc2_sqrt[i] = 1.f; a.inv_bc2_sqrt[i] = 1.f;
    }
    if (blocks == 0) return 0;
    adam_multi_kernel<<<blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(
        a, n_tensors, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, nullptr, nullptr);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_adam_multi_dev(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                               const int64_t* n, const float* hyper_dev, const int32_t* skip_flag, double beta1,
                               double beta2, double eps, void* stream) {
    if (n_tensors <= 0 || n_tensors > FSB_ADAM_MAX_TENSORS || !hyper_dev) return FSB_E_ARG;
    FsbAdamArgs a;
    int blocks = 0;
    for (int i = 0; i < FSB_ADAM_MAX_TENSORS; ++i) {
        const bool live = i < n_tensors;
        if (live && n[i] < 0) return FSB_E_ARG;
        a.p[i] = live ? p[i] : nullptr; a.g[i] = live ? g[i] : nullptr;
        a.m[i] = live ? m[i] : nullptr; a.v[i] = live ? v[i] : nullptr;
        a.n[i] = live ? n[i] : 0;
        a.block_start[i] = blocks;
        if (live) blocks += (int)((n[i] + ADAM_PER_BLOCK - 1) / ADAM_PER_BLOCK);
        a.step_size[i] = 0.f; a.bc2_sqrt[i] = 1.f; a.inv_bc2_sqrt[i] = 1.f;
    }
    a.block_start[FSB_ADAM_MAX_TENSORS] = blocks;
    if (blocks == 0) return 0;
    adam_multi_kernel<<<blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(
        a, n_tensors, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, hyper_dev, skip_flag);
    FSB_LAUNCH_CHECK();
    return 0;
}
