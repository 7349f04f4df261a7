// pseudo_normals.cu — surface normals from a depth image.
//
// Replaces normal_from_depth_image (/root/reference/dn_splatter/utils/normal_utils.py:23-46), i.e.
// get_means3d_backproj (/root/reference/dn_splatter/utils/camera_utils.py:92-144) followed by pcd_to_normal
// (normal_utils.py:7-20), reached from /root/reference/dn_splatter/dn_model.py:779-789 when
// normal_supervision == "depth":  back-project every pixel centre, cross product of the central differences
// (right - left) x (top - bottom), normalise, one pixel of zero padding.
//
// One thread per pixel; the four neighbours are back-projected on the fly (4 depth loads that hit L1/L2, no
// point-cloud round trip through HBM).  Algorithmic bytes: P * (4 + 12).
#include "common.cuh"

namespace {

struct BackprojArgs {
    float fx, fy, cx, cy;
    float m[9];  // inverse of c2w[:3,:3], row-major: world = cam_point (row vector) @ m + t
    float t[3];
};

__device__ __forceinline__ void backproject(const float* __restrict__ depth, const float* __restrict__ xyz, int W,
                                            int y, int x, const BackprojArgs& a, float (&p)[3]) {
    const size_t idx = (size_t)y * W + x;
    if (xyz != nullptr) {
        p[0] = xyz[3 * idx]; p[1] = xyz[3 * idx + 1]; p[2] = xyz[3 * idx + 2];
        return;
    }
    const float d = depth[idx];
    // same operation order as camera_utils.py:125-127: ((u - cx) * d) / fx, IEEE division, no contraction
    const float X = __fdiv_rn(__fmul_rn(__fsub_rn((float)x + 0.5f, a.cx), d), a.fx);
    const float Y = __fdiv_rn(__fmul_rn(__fsub_rn((float)y + 0.5f, a.cy), d), a.fy);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        p[k] = __fadd_rn(fmaf(d, a.m[6 + k], fmaf(Y, a.m[3 + k], __fmul_rn(X, a.m[k]))), a.t[k]);
}

__global__ void __launch_bounds__(256)
normal_from_depth_kernel(int H, int W, const float* __restrict__ depth, const float* __restrict__ xyz,
                         BackprojArgs a, float* __restrict__ normals) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    float n[3] = {0.f, 0.f, 0.f};
    if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
        float r[3], l[3], t[3], b[3];
        backproject(depth, xyz, W, y, x + 1, a, r);
        backproject(depth, xyz, W, y, x - 1, a, l);
        backproject(depth, xyz, W, y - 1, x, a, t);
        backproject(depth, xyz, W, y + 1, x, a, b);
        float u[3], v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { u[k] = __fsub_rn(r[k], l[k]); v[k] = __fsub_rn(t[k], b[k]); }
        n[0] = __fsub_rn(__fmul_rn(u[1], v[2]), __fmul_rn(u[2], v[1]));
        n[1] = __fsub_rn(__fmul_rn(u[2], v[0]), __fmul_rn(u[0], v[2]));
        n[2] = __fsub_rn(__fmul_rn(u[0], v[1]), __fmul_rn(u[1], v[0]));
        // torch.nn.functional.normalize(p=2, eps=1e-12): v / max(|v|, eps)
        const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])), __fmul_rn(n[2], n[2])));
        const float den = fmaxf(len, 1e-12f);
#pragma unroll
        for (int k = 0; k < 3; ++k) n[k] = __fdiv_rn(n[k], den);
    }
    float* o = normals + ((size_t)y * W + x) * 3;
    o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
}

}  // namespace

FSB_API int fsb_normal_from_depth(int H, int W, const float* depth, const float* xyz, float fx, float fy, float cx,
                                  float cy, const float* rot_inv, const float* trans, float* normals, void* stream) {
    if (H < 0 || W < 0 || (!depth && !xyz) || !normals) return FSB_E_ARG;
    if (depth && (!(fx != 0.f) || !(fy != 0.f))) return FSB_E_ARG;
    if (H == 0 || W == 0) return 0;
    BackprojArgs a;
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    for (int i = 0; i < 9; ++i) a.m[i] = rot_inv ? rot_inv[i] : ((i % 4 == 0) ? 1.f : 0.f);
    for (int i = 0; i < 3; ++i) a.t[i] = trans ? trans[i] : 0.f;
    dim3 block(32, 8);
    dim3 grid((unsigned)fsb_div_up(W, 32), (unsigned)fsb_div_up(H, 8));
    normal_from_depth_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(H, W, xyz ? nullptr : depth, xyz, a, normals);
    FSB_LAUNCH_CHECK();
    return 0;
}
