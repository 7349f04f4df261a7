// fs_math.cuh — per-Gaussian math shared by the projection / SH kernels.
//
// Everything here is a plain inline function marked FS_HD so the same source
// compiles (a) into the sm_100a kernels and (b) into the host-only math probe
// that tests/ uses to check formulas against fp64 autograd without a GPU.
// The product path never calls the host build.
//
// Semantics follow gsplat==1.0.0 as restated in SURVEY.md Appendix A.2/A.3
// (call sites: /root/reference/dn_splatter/dn_model.py:570-591).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FS_HD __host__ __device__ __forceinline__
#else
#define FS_HD inline
#endif

namespace fs {

// ---- rounding-exact helpers ------------------------------------------------
// The camera-space mean feeds the 32-bit depth half of the sort key, which must
// be bit-identical to the oracle's.  Each product and sum is rounded separately
// (no FMA contraction) in a fixed left-to-right order.
FS_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
FS_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
FS_HD float dot3_rn(float a0, float a1, float a2, float x, float y, float z, float t) {
    return add_rn(add_rn(add_rn(mul_rn(a0, x), mul_rn(a1, y)), mul_rn(a2, z)), t);
}

FS_HD float inv_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

struct Mat3 {
    float m[3][3];
};

FS_HD Mat3 mat3_mul(const Mat3& A, const Mat3& B) {
    Mat3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
FS_HD Mat3 mat3_mul_bt(const Mat3& A, const Mat3& B) {  // A * B^T
    Mat3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C.m[i][j] = A.m[i][0] * B.m[j][0] + A.m[i][1] * B.m[j][1] + A.m[i][2] * B.m[j][2];
    return C;
}
FS_HD Mat3 mat3_mul_at(const Mat3& A, const Mat3& B) {  // A^T * B
    Mat3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C.m[i][j] = A.m[0][i] * B.m[0][j] + A.m[1][i] * B.m[1][j] + A.m[2][i] * B.m[2][j];
    return C;
}

// rotation matrix of the NORMALISED quaternion (w,x,y,z); returns 1/|q| too.
FS_HD Mat3 quat_to_rotmat(float qw, float qx, float qy, float qz, float* inv_norm_out) {
    float inv = inv_sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    float w = qw * inv, x = qx * inv, y = qy * inv, z = qz * inv;
    Mat3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z);
    R.m[0][1] = 2.f * (x * y - w * z);
    R.m[0][2] = 2.f * (x * z + w * y);
    R.m[1][0] = 2.f * (x * y + w * z);
    R.m[1][1] = 1.f - 2.f * (x * x + z * z);
    R.m[1][2] = 2.f * (y * z - w * x);
    R.m[2][0] = 2.f * (x * z - w * y);
    R.m[2][1] = 2.f * (y * z + w * x);
    R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    if (inv_norm_out) *inv_norm_out = inv;
    return R;
}

// gradient of a loss w.r.t. the UN-normalised quaternion given dL/dR.
FS_HD void quat_to_rotmat_vjp(float qw, float qx, float qy, float qz, const Mat3& vR, float vq[4]) {
    float inv = inv_sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    float w = qw * inv, x = qx * inv, y = qy * inv, z = qz * inv;
    float gw = 2.f * (x * (vR.m[2][1] - vR.m[1][2]) + y * (vR.m[0][2] - vR.m[2][0]) + z * (vR.m[1][0] - vR.m[0][1]));
    float gx = 2.f * (-2.f * x * (vR.m[1][1] + vR.m[2][2]) + y * (vR.m[1][0] + vR.m[0][1]) +
                      z * (vR.m[2][0] + vR.m[0][2]) + w * (vR.m[2][1] - vR.m[1][2]));
    float gy = 2.f * (x * (vR.m[1][0] + vR.m[0][1]) - 2.f * y * (vR.m[0][0] + vR.m[2][2]) +
                      z * (vR.m[2][1] + vR.m[1][2]) + w * (vR.m[0][2] - vR.m[2][0]));
    float gz = 2.f * (x * (vR.m[2][0] + vR.m[0][2]) + y * (vR.m[2][1] + vR.m[1][2]) -
                      2.f * z * (vR.m[0][0] + vR.m[1][1]) + w * (vR.m[1][0] - vR.m[0][1]));
    // through q_hat = q / |q|
    float d = gw * w + gx * x + gy * y + gz * z;
    vq[0] = (gw - d * w) * inv;
    vq[1] = (gx - d * x) * inv;
    vq[2] = (gy - d * y) * inv;
    vq[3] = (gz - d * z) * inv;
}

struct Camera {
    float V[12];  // rows 0..2 of the 4x4 world->camera matrix, row-major
    float fx, fy, cx, cy;
};

struct ProjFwd {
    int radius;  // 0 = culled
    float mx, my, depth;
    float ca, cb, cc;  // conic (inverse of the blurred 2D covariance): xx, xy, yy
    float comp;
};

// intermediate state the backward pass recomputes
struct ProjState {
    float mcx, mcy, mcz;
    Mat3 Rq;       // rotation of the normalised quaternion
    Mat3 Sigma;    // world covariance
    Mat3 SigmaC;   // camera-space covariance
    float J[2][3];
    float c2xx, c2xy, c2yy;  // un-blurred 2D covariance
    float det0, det1;
    bool in_x, in_y;  // x/z, y/z inside the 1.3*tan(fov/2) clip
};

FS_HD bool project_core(const Camera& cam, float px, float py, float pz, float qw, float qx, float qy, float qz,
                        float sx, float sy, float sz, int width, int height, float eps2d, float near_plane,
                        float far_plane, ProjState& st) {
    st.mcx = dot3_rn(cam.V[0], cam.V[1], cam.V[2], px, py, pz, cam.V[3]);
    st.mcy = dot3_rn(cam.V[4], cam.V[5], cam.V[6], px, py, pz, cam.V[7]);
    st.mcz = dot3_rn(cam.V[8], cam.V[9], cam.V[10], px, py, pz, cam.V[11]);
    if (st.mcz < near_plane || st.mcz > far_plane) return false;

    st.Rq = quat_to_rotmat(qw, qx, qy, qz, nullptr);
    Mat3 M;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        M.m[i][0] = st.Rq.m[i][0] * sx;
        M.m[i][1] = st.Rq.m[i][1] * sy;
        M.m[i][2] = st.Rq.m[i][2] * sz;
    }
    st.Sigma = mat3_mul_bt(M, M);
    Mat3 Rv;
    Rv.m[0][0] = cam.V[0]; Rv.m[0][1] = cam.V[1]; Rv.m[0][2] = cam.V[2];
    Rv.m[1][0] = cam.V[4]; Rv.m[1][1] = cam.V[5]; Rv.m[1][2] = cam.V[6];
    Rv.m[2][0] = cam.V[8]; Rv.m[2][1] = cam.V[9]; Rv.m[2][2] = cam.V[10];
    st.SigmaC = mat3_mul_bt(mat3_mul(Rv, st.Sigma), Rv);

    float x = st.mcx, y = st.mcy, z = st.mcz;
    float tan_fovx = 0.5f * (float)width / cam.fx;
    float tan_fovy = 0.5f * (float)height / cam.fy;
    float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
    float rz = 1.f / z;
    float rz2 = rz * rz;
    float xz = x * rz, yz = y * rz;
    st.in_x = (xz <= lim_x) && (xz >= -lim_x);
    st.in_y = (yz <= lim_y) && (yz >= -lim_y);
    float tx = z * fminf(lim_x, fmaxf(-lim_x, xz));
    float ty = z * fminf(lim_y, fmaxf(-lim_y, yz));
    st.J[0][0] = cam.fx * rz; st.J[0][1] = 0.f;         st.J[0][2] = -cam.fx * tx * rz2;
    st.J[1][0] = 0.f;         st.J[1][1] = cam.fy * rz; st.J[1][2] = -cam.fy * ty * rz2;

    // Sigma2 = J SigmaC J^T
    float a0[3], a1[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        a0[j] = st.J[0][0] * st.SigmaC.m[0][j] + st.J[0][2] * st.SigmaC.m[2][j];
        a1[j] = st.J[1][1] * st.SigmaC.m[1][j] + st.J[1][2] * st.SigmaC.m[2][j];
    }
    st.c2xx = a0[0] * st.J[0][0] + a0[2] * st.J[0][2];
    st.c2xy = a0[1] * st.J[1][1] + a0[2] * st.J[1][2];
    st.c2yy = a1[1] * st.J[1][1] + a1[2] * st.J[1][2];
    st.det0 = st.c2xx * st.c2yy - st.c2xy * st.c2xy;
    float bxx = st.c2xx + eps2d, byy = st.c2yy + eps2d;
    st.det1 = bxx * byy - st.c2xy * st.c2xy;
    return true;
}

FS_HD ProjFwd project_fwd(const Camera& cam, float px, float py, float pz, float qw, float qx, float qy, float qz,
                          float sx, float sy, float sz, int width, int height, float eps2d, float near_plane,
                          float far_plane, float radius_clip) {
    ProjFwd o;
    o.radius = 0; o.mx = o.my = o.depth = 0.f; o.ca = o.cb = o.cc = 0.f; o.comp = 0.f;
    ProjState st;
    if (!project_core(cam, px, py, pz, qw, qx, qy, qz, sx, sy, sz, width, height, eps2d, near_plane, far_plane, st))
        return o;
    float det = st.det1;
    if (det <= 0.f) return o;
    float bxx = st.c2xx + eps2d, byy = st.c2yy + eps2d;
    float rdet = 1.f / det;
    float b = 0.5f * (bxx + byy);
    float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
    float radius = ceilf(3.f * sqrtf(v1));
    if (radius <= radius_clip) return o;
    float rz = 1.f / st.mcz;
    float mx = cam.fx * st.mcx * rz + cam.cx;
    float my = cam.fy * st.mcy * rz + cam.cy;
    if (mx + radius <= 0.f || mx - radius >= (float)width || my + radius <= 0.f || my - radius >= (float)height)
        return o;
    o.radius = (int)radius;
    o.mx = mx; o.my = my; o.depth = st.mcz;
    o.ca = byy * rdet; o.cb = -st.c2xy * rdet; o.cc = bxx * rdet;
    o.comp = sqrtf(fmaxf(0.f, st.det0 * rdet));
    return o;
}

// Backward of project_fwd for one visible (camera, Gaussian) pair.
// In: v_mean2d (2), v_depth, v_conic (a,b,c), v_comp.  Out (overwritten): v_mean[3], v_quat[4], v_scale[3];
// if v_Rv / v_tv are non-null, dL/d(viewmat rotation, translation) are written too.
FS_HD void project_bwd(const Camera& cam, float px, float py, float pz, float qw, float qx, float qy, float qz,
                       float sx, float sy, float sz, int width, int height, float eps2d, float vmx, float vmy,
                       float vdepth, float vca, float vcb, float vcc, float vcomp, float v_mean[3], float v_quat[4],
                       float v_scale[3], float* v_Rv /*9 or null*/, float* v_tv /*3 or null*/) {
    ProjState st;
    project_core(cam, px, py, pz, qw, qx, qy, qz, sx, sy, sz, width, height, eps2d, -INFINITY, INFINITY, st);
    float bxx = st.c2xx + eps2d, byy = st.c2yy + eps2d;
    float rdet = 1.f / st.det1;
    float a = byy * rdet, b = -st.c2xy * rdet, c = bxx * rdet;  // conic
    // dL/dSigma2' (full 2x2) = -X * [[va, vb/2],[vb/2, vc]] * X with X = conic matrix
    float h = 0.5f * vcb;
    float t00 = a * vca + b * h, t01 = a * h + b * vcc;
    float t10 = b * vca + c * h, t11 = b * h + c * vcc;
    float g00 = -(t00 * a + t01 * b);
    float g01 = -(t00 * b + t01 * c);
    float g10 = -(t10 * a + t11 * b);
    float g11 = -(t10 * b + t11 * c);
    if (vcomp != 0.f) {
        // comp = sqrt(max(0, det0/det1)); det0 = xx*yy - xy^2 ; det1 = (xx+e)(yy+e) - xy^2
        float ratio = st.det0 * rdet;
        if (ratio > 0.f) {
            float comp = sqrtf(ratio);
            float k = 0.5f * vcomp / comp;  // dL/dratio
            float d_xx = k * (st.c2yy * rdet - st.det0 * byy * rdet * rdet);
            float d_yy = k * (st.c2xx * rdet - st.det0 * bxx * rdet * rdet);
            float d_xy = k * (-2.f * st.c2xy * rdet + 2.f * st.det0 * st.c2xy * rdet * rdet);
            g00 += d_xx; g11 += d_yy; g01 += 0.5f * d_xy; g10 += 0.5f * d_xy;
        }
    }
    // V_SigmaC = J^T G J   (3x3)
    float GJ[2][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        GJ[0][j] = g00 * st.J[0][j] + g01 * st.J[1][j];
        GJ[1][j] = g10 * st.J[0][j] + g11 * st.J[1][j];
    }
    Mat3 VSc;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) VSc.m[i][j] = st.J[0][i] * GJ[0][j] + st.J[1][i] * GJ[1][j];
    // V_J = G J SigmaC^T + G^T J SigmaC   (2x3)
    float GtJ[2][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        GtJ[0][j] = g00 * st.J[0][j] + g10 * st.J[1][j];
        GtJ[1][j] = g01 * st.J[0][j] + g11 * st.J[1][j];
    }
    float VJ[2][3];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += GJ[i][k] * st.SigmaC.m[j][k] + GtJ[i][k] * st.SigmaC.m[k][j];
            VJ[i][j] = s;
        }
    float x = st.mcx, y = st.mcy, z = st.mcz;
    float rz = 1.f / z, rz2 = rz * rz, rz3 = rz2 * rz;
    float tan_fovx = 0.5f * (float)width / cam.fx;
    float tan_fovy = 0.5f * (float)height / cam.fy;
    float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
    float tx = z * fminf(lim_x, fmaxf(-lim_x, x * rz));
    float ty = z * fminf(lim_y, fmaxf(-lim_y, y * rz));
    float vx = cam.fx * rz * vmx;
    float vy = cam.fy * rz * vmy;
    float vz = -(cam.fx * x * vmx + cam.fy * y * vmy) * rz2 + vdepth;
    vz += -cam.fx * rz2 * VJ[0][0] - cam.fy * rz2 * VJ[1][1] + 2.f * cam.fx * tx * rz3 * VJ[0][2] +
          2.f * cam.fy * ty * rz3 * VJ[1][2];
    if (st.in_x) vx += -cam.fx * rz2 * VJ[0][2];
    else vz += -cam.fx * rz3 * VJ[0][2] * tx;
    if (st.in_y) vy += -cam.fy * rz2 * VJ[1][2];
    else vz += -cam.fy * rz3 * VJ[1][2] * ty;

    Mat3 Rv;
    Rv.m[0][0] = cam.V[0]; Rv.m[0][1] = cam.V[1]; Rv.m[0][2] = cam.V[2];
    Rv.m[1][0] = cam.V[4]; Rv.m[1][1] = cam.V[5]; Rv.m[1][2] = cam.V[6];
    Rv.m[2][0] = cam.V[8]; Rv.m[2][1] = cam.V[9]; Rv.m[2][2] = cam.V[10];
    // mean_c = Rv * p + t
    v_mean[0] = Rv.m[0][0] * vx + Rv.m[1][0] * vy + Rv.m[2][0] * vz;
    v_mean[1] = Rv.m[0][1] * vx + Rv.m[1][1] * vy + Rv.m[2][1] * vz;
    v_mean[2] = Rv.m[0][2] * vx + Rv.m[1][2] * vy + Rv.m[2][2] * vz;
    // SigmaC = Rv Sigma Rv^T  ->  V_Sigma = Rv^T V_SigmaC Rv
    Mat3 VS = mat3_mul(mat3_mul_at(Rv, VSc), Rv);
    if (v_Rv) {
        // dL/dRv = v_mean_c p^T + V_SigmaC Rv Sigma^T + V_SigmaC^T Rv Sigma
        Mat3 RS = mat3_mul(Rv, st.Sigma);  // Sigma symmetric
        float vmc[3] = {vx, vy, vz};
        float p[3] = {px, py, pz};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float s = vmc[i] * p[j];
#pragma unroll
                for (int k = 0; k < 3; ++k) s += (VSc.m[i][k] + VSc.m[k][i]) * RS.m[k][j];
                v_Rv[i * 3 + j] = s;
            }
        v_tv[0] = vx; v_tv[1] = vy; v_tv[2] = vz;
    }
    // Sigma = M M^T -> V_M = (V_Sigma + V_Sigma^T) M, M = Rq diag(s)
    Mat3 VM;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
            float sc[3] = {sx, sy, sz};
#pragma unroll
            for (int k = 0; k < 3; ++k) s += (VS.m[i][k] + VS.m[k][i]) * st.Rq.m[k][j] * sc[j];
            VM.m[i][j] = s;
        }
    Mat3 VRq;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        VRq.m[i][0] = VM.m[i][0] * sx;
        VRq.m[i][1] = VM.m[i][1] * sy;
        VRq.m[i][2] = VM.m[i][2] * sz;
    }
    v_scale[0] = st.Rq.m[0][0] * VM.m[0][0] + st.Rq.m[1][0] * VM.m[1][0] + st.Rq.m[2][0] * VM.m[2][0];
    v_scale[1] = st.Rq.m[0][1] * VM.m[0][1] + st.Rq.m[1][1] * VM.m[1][1] + st.Rq.m[2][1] * VM.m[2][1];
    v_scale[2] = st.Rq.m[0][2] * VM.m[0][2] + st.Rq.m[1][2] * VM.m[1][2] + st.Rq.m[2][2] * VM.m[2][2];
    quat_to_rotmat_vjp(qw, qx, qy, qz, VRq, v_quat);
}

// ---- spherical harmonics (Sloan fast form) -------------------
// coeffs: pointer to this Gaussian's [K,3] block.  degree <= 3 (the reference never goes higher: dn_model.py:562-565).
FS_HD void sh_basis(int degree, float x, float y, float z, float b[16]) {
    b[0] = 0.2820947917738781f;
    if (degree < 1) return;
    b[1] = -0.48860251190292f * y;
    b[2] = 0.48860251190292f * z;
    b[3] = -0.48860251190292f * x;
    if (degree < 2) return;
    float z2 = z * z;
    float fTmp0B = -1.092548430592079f * z;
    float fC1 = x * x - y * y;
    float fS1 = 2.f * x * y;
    b[4] = 0.5462742152960395f * fS1;
    b[5] = fTmp0B * y;
    b[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
    b[7] = fTmp0B * x;
    b[8] = 0.5462742152960395f * fC1;
    if (degree < 3) return;
    float fTmp0C = -2.285228997322329f * z2 + 0.4570457994644658f;
    float fTmp1B = 1.445305721320277f * z;
    float fC2 = x * fC1 - y * fS1;
    float fS2 = x * fS1 + y * fC1;
    b[9] = -0.5900435899266435f * fS2;
    b[10] = fTmp1B * fS1;
    b[11] = fTmp0C * y;
    b[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
    b[13] = fTmp0C * x;
    b[14] = fTmp1B * fC1;
    b[15] = -0.5900435899266435f * fC2;
}

// d basis / d (x,y,z) of the UNIT direction, degree <= 3.
FS_HD void sh_basis_grad(int degree, float x, float y, float z, float dx[16], float dy[16], float dz[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dx[i] = dy[i] = dz[i] = 0.f;
    if (degree < 1) return;
    dy[1] = -0.48860251190292f;
    dz[2] = 0.48860251190292f;
    dx[3] = -0.48860251190292f;
    if (degree < 2) return;
    const float c2 = 0.5462742152960395f, cT = 1.092548430592079f;
    float fC1 = x * x - y * y, fS1 = 2.f * x * y;
    dx[4] = c2 * 2.f * y; dy[4] = c2 * 2.f * x;
    dy[5] = -cT * z;      dz[5] = -cT * y;
    dz[6] = 2.f * 0.9461746957575601f * z;
    dx[7] = -cT * z;      dz[7] = -cT * x;
    dx[8] = c2 * 2.f * x; dy[8] = -c2 * 2.f * y;
    if (degree < 3) return;
    const float c3 = 0.5900435899266435f, cU = 2.285228997322329f, cV = 1.445305721320277f;
    float z2 = z * z;
    float U = -cU * z2 + 0.4570457994644658f;
    float V = cV * z;
    dx[9] = -c3 * 3.f * fS1;  dy[9] = -c3 * 3.f * fC1;
    dx[10] = V * 2.f * y;     dy[10] = V * 2.f * x;   dz[10] = cV * fS1;
    dy[11] = U;               dz[11] = -2.f * cU * z * y;
    dz[12] = 3.f * 1.865881662950577f * z2 - 1.119528997770346f;
    dx[13] = U;               dz[13] = -2.f * cU * z * x;
    dx[14] = V * 2.f * x;     dy[14] = -V * 2.f * y;  dz[14] = cV * fC1;
    dx[15] = -c3 * 3.f * fC1; dy[15] = c3 * 3.f * fS1;
}

}  // namespace fs
