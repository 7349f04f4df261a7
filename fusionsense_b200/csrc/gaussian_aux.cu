// gaussian_aux.cu — per-Gaussian side computations of the DN-Splatter step that are not part of gsplat:
//   * per-Gaussian normals (smallest-scale axis of the rotation, flipped towards the camera, rotated into the
//     camera frame) and their backward          -> /root/reference/dn_splatter/dn_model.py:617-636
//   * per-step densification statistics          -> nerfstudio splatfacto after_train, consumed by
//                                                   /root/reference/dn_splatter/dn_model.py:326-451
// Both are one-thread-per-Gaussian streaming kernels (HBM-bound, tens of bytes per Gaussian) that replace
// ~10 and ~40 small torch launches respectively.
#include "common.cuh"
#include "fs_math.cuh"

namespace {

__device__ __forceinline__ int argmin3(float a, float b, float c) {
    int k = 0;
    float m = a;
    if (b < m) { m = b; k = 1; }
    if (c < m) { k = 2; }
    return k;
}

__global__ void __launch_bounds__(256)
normals_fwd_kernel(int N, const float* __restrict__ quats, const float* __restrict__ scales,
                   const float* __restrict__ means, const float* __restrict__ c2w /*[3,4] or [4,4] row-major, stride 4*/,
                   float* __restrict__ normals_world, float* __restrict__ normals_cam) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float4 q = reinterpret_cast<const float4*>(quats)[n];
    fs::Mat3 R = fs::quat_to_rotmat(q.x, q.y, q.z, q.w, nullptr);
    int k = argmin3(scales[3 * (size_t)n], scales[3 * (size_t)n + 1], scales[3 * (size_t)n + 2]);
    float nx = R.m[0][k], ny = R.m[1][k], nz = R.m[2][k];
    float inv = 1.f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);  // F.normalize(eps=1e-12)
    nx *= inv; ny *= inv; nz *= inv;
    // viewdir = cam_pos - mean (the sign of the dot product does not depend on its normalisation)
    float vx = __ldg(c2w + 3) - means[3 * (size_t)n], vy = __ldg(c2w + 7) - means[3 * (size_t)n + 1],
          vz = __ldg(c2w + 11) - means[3 * (size_t)n + 2];
    float vn = sqrtf(vx * vx + vy * vy + vz * vz);
    float dot = (nx * vx + ny * vy + nz * vz) / vn;
    if (dot < 0.f) { nx = -nx; ny = -ny; nz = -nz; }
    if (normals_world) {
        normals_world[3 * (size_t)n] = nx; normals_world[3 * (size_t)n + 1] = ny; normals_world[3 * (size_t)n + 2] = nz;
    }
    // normals @ c2w[:3,:3]
#pragma unroll
    for (int j = 0; j < 3; ++j)
        normals_cam[3 * (size_t)n + j] = nx * __ldg(c2w + j) + ny * __ldg(c2w + 4 + j) + nz * __ldg(c2w + 8 + j);
}

__global__ void __launch_bounds__(256)
normals_bwd_kernel(int N, const float* __restrict__ quats, const float* __restrict__ scales,
                   const float* __restrict__ means, const float* __restrict__ c2w,
                   const float* __restrict__ v_normals_cam, float* __restrict__ v_quats) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float4 q = reinterpret_cast<const float4*>(quats)[n];
    fs::Mat3 R = fs::quat_to_rotmat(q.x, q.y, q.z, q.w, nullptr);
    int k = argmin3(scales[3 * (size_t)n], scales[3 * (size_t)n + 1], scales[3 * (size_t)n + 2]);
    float cx = R.m[0][k], cy = R.m[1][k], cz = R.m[2][k];
    float len = fmaxf(sqrtf(cx * cx + cy * cy + cz * cz), 1e-12f);
    float nx = cx / len, ny = cy / len, nz = cz / len;
    float vx = __ldg(c2w + 3) - means[3 * (size_t)n], vy = __ldg(c2w + 7) - means[3 * (size_t)n + 1],
          vz = __ldg(c2w + 11) - means[3 * (size_t)n + 2];
    float vn = sqrtf(vx * vx + vy * vy + vz * vz);
    float sign = ((nx * vx + ny * vy + nz * vz) / vn < 0.f) ? -1.f : 1.f;
    float g0 = v_normals_cam[3 * (size_t)n], g1 = v_normals_cam[3 * (size_t)n + 1], g2 = v_normals_cam[3 * (size_t)n + 2];
    // out = n_flipped @ Rc  ->  v_n_flipped[i] = sum_j Rc[i][j] * v_out[j]
    float wx = sign * (__ldg(c2w + 0) * g0 + __ldg(c2w + 1) * g1 + __ldg(c2w + 2) * g2);
    float wy = sign * (__ldg(c2w + 4) * g0 + __ldg(c2w + 5) * g1 + __ldg(c2w + 6) * g2);
    float wz = sign * (__ldg(c2w + 8) * g0 + __ldg(c2w + 9) * g1 + __ldg(c2w + 10) * g2);
    // through n = col / |col|
    float d = nx * wx + ny * wy + nz * wz;
    fs::Mat3 vR;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vR.m[i][j] = 0.f;
    vR.m[0][k] = (wx - d * nx) / len;
    vR.m[1][k] = (wy - d * ny) / len;
    vR.m[2][k] = (wz - d * nz) / len;
    float vq[4];
    fs::quat_to_rotmat_vjp(q.x, q.y, q.z, q.w, vR, vq);
    reinterpret_cast<float4*>(v_quats)[n] = make_float4(vq[0], vq[1], vq[2], vq[3]);
}

__global__ void __launch_bounds__(256)
densify_stats_kernel(int N, const int32_t* __restrict__ radii, const float2* __restrict__ grads2d, float max_dim,
                     float* __restrict__ xys_grad_norm, float* __restrict__ vis_counts,
                     float* __restrict__ max_2Dsize, const int32_t* __restrict__ skip_flag) {
    if (skip_flag != nullptr && *skip_flag != 0) return;
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int r = radii[n];
    if (r <= 0) return;
    float2 g = grads2d[n];
    vis_counts[n] += 1.f;
    xys_grad_norm[n] += sqrtf(g.x * g.x + g.y * g.y);
    max_2Dsize[n] = fmaxf(max_2Dsize[n], (float)r / max_dim);
}

}  // namespace

FSB_API int fsb_gaussian_normals_fwd(int N, const float* quats, const float* scales, const float* means,
                                     const float* c2w, float* normals_world, float* normals_cam, void* stream) {
    if (N < 0) return FSB_E_ARG;
    if (N == 0) return 0;
    normals_fwd_kernel<<<fsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(N, quats, scales, means, c2w,
                                                                           normals_world, normals_cam);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_gaussian_normals_bwd(int N, const float* quats, const float* scales, const float* means,
                                     const float* c2w, const float* v_normals_cam, float* v_quats, void* stream) {
    if (N < 0) return FSB_E_ARG;
    if (N == 0) return 0;
    normals_bwd_kernel<<<fsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(N, quats, scales, means, c2w,
                                                                           v_normals_cam, v_quats);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_densify_stats(int N, const int32_t* radii, const float* grads2d, float max_dim, float* xys_grad_norm,
                              float* vis_counts, float* max_2Dsize, const int32_t* skip_flag, void* stream) {
    if (N < 0 || !(max_dim > 0.f)) return FSB_E_ARG;
    if (N == 0) return 0;
    densify_stats_kernel<<<fsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        N, radii, (const float2*)grads2d, max_dim, xys_grad_norm, vis_counts, max_2Dsize, skip_flag);
    FSB_LAUNCH_CHECK();
    return 0;
}
