// seed_points.cu — seed point cloud of Module 1: depth back-projection and voxel down-sampling.
//
// Replaces utils/generate_pcd.py:15-48 `get_pointcloud` (a dozen torch launches and a [3,3] @ [3,P] matmul per view,
// then two boolean-mask gathers) and the per-view `voxel_down_sample(voxel_size=0.02)` that open3d runs on the CPU
// (utils/generate_pcd.py:98-101), i.e. the cloud `init_pcd_generate` writes to merged_pcd.ply (scripts/train.py:95).
//
//   back-projection: one thread per pixel; a pixel with lo < depth < hi becomes the row
//       (R ((u - cx) / fx d, (v - cy) / fy d, d) + T, r, g, b), rows in pixel order (what the mask gather produces);
//       the rank of every kept pixel comes from the library's int32 -> int64 scan (fsb_isect_scan);
//   voxel down-sample: open3d's rule — voxel = floor((p - (min_bound - voxel / 2)) / voxel), one output point per
//       occupied voxel = the mean of its points and colours in fp64 — as key build, the library's radix sort (stable,
//       so a voxel's points keep their input order and the fp64 sums add up in the order open3d adds them), and one
//       thread per voxel run.  Output order: ascending voxel key (open3d's is its hash map's iteration order).
#include "common.cuh"

namespace {

struct BackprojArgs {
    float R[9];  // camera-to-world rotation, row-major
    float T[3];
    float fx, fy, cx, cy, lo, hi;
};

__global__ void __launch_bounds__(256)
backproject_flags_kernel(int64_t P, const float* __restrict__ depth, float lo, float hi, int32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float d = depth[i];
    flags[i] = (d > lo && d < hi) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
backproject_emit_kernel(int H, int W, const float* __restrict__ depth, const float* __restrict__ color_chw,
                        BackprojArgs a, const int64_t* __restrict__ offsets, float* __restrict__ out) {
    const int64_t P = (int64_t)H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float d = depth[i];
    if (!(d > a.lo && d < a.hi)) return;
    const int v = (int)(i / W), u = (int)(i - (int64_t)v * W);
    // generate_pcd.py:20-28: xx = (x - CX) / FX ; pts_cam = (xx * d, yy * d, d)
    const float xx = __fdiv_rn((float)u - a.cx, a.fx), yy = __fdiv_rn((float)v - a.cy, a.fy);
    const float xc = xx * d, yc = yy * d, zc = d;
    float* o = out + 6 * offsets[i];
    // generate_pcd.py:32-35: (R @ pts_cam.T) + T
    o[0] = fmaf(a.R[2], zc, fmaf(a.R[1], yc, a.R[0] * xc)) + a.T[0];
    o[1] = fmaf(a.R[5], zc, fmaf(a.R[4], yc, a.R[3] * xc)) + a.T[1];
    o[2] = fmaf(a.R[8], zc, fmaf(a.R[7], yc, a.R[6] * xc)) + a.T[2];
    o[3] = color_chw[i]; o[4] = color_chw[P + i]; o[5] = color_chw[2 * P + i];
}

// ---- voxel down-sample ----------------------------------------------------------------------------------------
constexpr int VB = 21;  // bits per axis of the voxel key

__global__ void __launch_bounds__(256)
voxel_min_partial_kernel(int64_t N, const float* __restrict__ pts, int stride, double* __restrict__ partial) {
    double mn[3] = {1e300, 1e300, 1e300};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) mn[a] = fmin(mn[a], (double)pts[i * stride + a]);
    }
    __shared__ double sh[3][256];
#pragma unroll
    for (int a = 0; a < 3; ++a) sh[a][threadIdx.x] = mn[a];
    __syncthreads();
    if (threadIdx.x < 3) {
        double r = sh[threadIdx.x][0];
        for (int t = 1; t < 256; ++t) r = fmin(r, sh[threadIdx.x][t]);
        partial[(size_t)blockIdx.x * 3 + threadIdx.x] = r;
    }
}

__global__ void voxel_min_final_kernel(int n_partials, const double* __restrict__ partial, double* __restrict__ out) {
    const int a = threadIdx.x;
    if (a >= 3) return;
    double r = partial[a];
    for (int b = 1; b < n_partials; ++b) r = fmin(r, partial[(size_t)b * 3 + a]);
    out[a] = r;
}

__global__ void __launch_bounds__(256)
voxel_keys_kernel(int64_t N, const float* __restrict__ pts, int stride, const double* __restrict__ min_bound,
                  double voxel, uint64_t* __restrict__ keys, int32_t* __restrict__ vals, int32_t* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint64_t key = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // open3d VoxelDownSample: ref_coord = (point - voxel_min_bound) / voxel_size ; voxel_index = floor(ref_coord)
        const double ref = ((double)pts[i * stride + a] - (min_bound[a] - voxel * 0.5)) / voxel;
        double f = floor(ref);
        if (!(f >= 0.0 && f < (double)(1 << VB))) { atomicExch(overflow, 1); f = 0.0; }
        key |= (uint64_t)f << (VB * (2 - a));  // x in the high bits: ascending (x, y, z) voxel order
    }
    keys[i] = key;
    vals[i] = (int32_t)i;
}

// heads[i] = 1 where sorted position i starts a voxel run
__global__ void __launch_bounds__(256)
voxel_heads_kernel(int64_t N, const uint64_t* __restrict__ keys, int32_t* __restrict__ heads) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    heads[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per run head: fp64 mean of the run's rows, in input order
__global__ void __launch_bounds__(128)
voxel_mean_kernel(int64_t N, const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                  const int32_t* __restrict__ heads, const int64_t* __restrict__ offsets, const float* __restrict__ pts,
                  int stride, int width, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !heads[i]) return;
    const uint64_t key = keys[i];
    double acc[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] = 0.0;
    int64_t j = i;
    for (; j < N && keys[j] == key; ++j) {
        const float* p = pts + (size_t)vals[j] * stride;
        for (int c = 0; c < width; ++c) acc[c] += (double)p[c];
    }
    const double cnt = (double)(j - i);
    double* o = out + offsets[i] * width;
    for (int c = 0; c < width; ++c) o[c] = acc[c] / cnt;
}

}  // namespace

// flags[i] = lo < depth[i] < hi (i32); scan them with fsb_isect_scan, then
FSB_API int fsb_backproject_flags(int64_t P, const float* depth, float lo, float hi, int32_t* flags, void* stream) {
    if (P < 0 || (P > 0 && (!depth || !flags))) return FSB_E_ARG;
    if (P == 0) return 0;
    backproject_flags_kernel<<<fsb_div_up(P, 256), 256, 0, (cudaStream_t)stream>>>(P, depth, lo, hi, flags);
    FSB_LAUNCH_CHECK();
    return 0;
}

// out[offsets[i]] = (R p_cam + T, colour) for every kept pixel i.  c2w_rot (9 floats, row-major) and c2w_trans (3) are
// HOST pointers; color_chw is the [3,H,W] image ToTensor() makes.
FSB_API int fsb_backproject_emit(int H, int W, const float* depth, const float* color_chw, const float* c2w_rot,
                                 const float* c2w_trans, float fx, float fy, float cx, float cy, float lo, float hi,
                                 const int64_t* offsets, float* out, void* stream) {
    if (H <= 0 || W <= 0 || !depth || !color_chw || !c2w_rot || !c2w_trans || !offsets || !out) return FSB_E_ARG;
    BackprojArgs a;
    for (int i = 0; i < 9; ++i) a.R[i] = c2w_rot[i];
    for (int i = 0; i < 3; ++i) a.T[i] = c2w_trans[i];
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.lo = lo; a.hi = hi;
    backproject_emit_kernel<<<fsb_div_up((int64_t)H * W, 256), 256, 0, (cudaStream_t)stream>>>(H, W, depth, color_chw, a,
                                                                                            offsets, out);
    FSB_LAUNCH_CHECK();
    return 0;
}

// Voxel down-sample of rows pts[N, stride] (xyz first, `width` <= 9 averaged columns):
//   fsb_voxel_keys : min_bound[3] f64 (device, written here: the per-axis minimum), keys / vals for fsb_radix_sort_pairs
//                    (63 bits); *overflow (device i32, not zeroed here) is set when a voxel index leaves [0, 2^21)
//   fsb_voxel_heads: heads[i] i32 = sorted position i starts a voxel; scan with fsb_isect_scan -> offsets, total
//   fsb_voxel_mean : out[n_voxels, width] f64 = mean of every voxel's rows
FSB_API size_t fsb_voxel_workspace(void) { return (size_t)FSB_NUM_SMS * 4 * 3 * sizeof(double); }

FSB_API int fsb_voxel_keys(int64_t N, const float* pts, int stride, double voxel, double* min_bound, uint64_t* keys,
                           int32_t* vals, int32_t* overflow, void* workspace, size_t workspace_bytes, void* stream) {
    if (N <= 0 || N > 0x7fffffff || stride < 3 || !(voxel > 0.0) || !pts || !min_bound || !keys || !vals || !overflow ||
        !workspace || workspace_bytes < fsb_voxel_workspace())
        return FSB_E_ARG;
    int blocks = fsb_div_up(N, 256);
    if (blocks > FSB_NUM_SMS * 4) blocks = FSB_NUM_SMS * 4;
    voxel_min_partial_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(N, pts, stride, (double*)workspace);
    FSB_LAUNCH_CHECK();
    voxel_min_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(blocks, (const double*)workspace, min_bound);
    FSB_LAUNCH_CHECK();
    voxel_keys_kernel<<<fsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(N, pts, stride, min_bound, voxel, keys, vals,
                                                                          overflow);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_voxel_heads(int64_t N, const uint64_t* sorted_keys, int32_t* heads, void* stream) {
    if (N <= 0 || !sorted_keys || !heads) return FSB_E_ARG;
    voxel_heads_kernel<<<fsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(N, sorted_keys, heads);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_voxel_mean(int64_t N, const uint64_t* sorted_keys, const int32_t* sorted_vals, const int32_t* heads,
                           const int64_t* offsets, const float* pts, int stride, int width, double* out, void* stream) {
    if (N <= 0 || width < 3 || width > 9 || width > stride || !sorted_keys || !sorted_vals || !heads || !offsets || !pts ||
        !out)
        return FSB_E_ARG;
    voxel_mean_kernel<<<fsb_div_up(N, 128), 128, 0, (cudaStream_t)stream>>>(N, sorted_keys, sorted_vals, heads, offsets,
                                                                          pts, stride, width, out);
    FSB_LAUNCH_CHECK();
    return 0;
}
