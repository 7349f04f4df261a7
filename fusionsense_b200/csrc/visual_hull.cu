// visual_hull.cu — voxel carving from silhouette masks (votes, threshold, ordered compaction).
//
// Replaces the projection / vote loop and the occupied-voxel extraction of the reference's
// /root/reference/utils/VisualHull.py:149-191 (with InitializeVoxels :15-57 folded in: voxel coordinates
// are looked up from the three axis tables instead of materialising the [V,4] array).
//
// fp64 arithmetic in the reference's order so occupancy is bit exact:
//   p = M[r,0]*x (+fma) M[r,1]*y (+fma) M[r,2]*z (+fma) M[r,3]        (np.matmul, k = 4 FMA chain, :158)
//   q = floor(p / p_z + 1e-6) -> int32 with x86 cvttsd2si semantics (NaN / out of range -> INT_MIN)  (:159)
//   negatives -> 0 per element (:160); row >= H or col >= W -> (0,0) (:163-166); votes += mask[row, col]/255 (:170)
// One thread per voxel, voxel order = z (outer, as given), x, y (inner) like the reference's loops (:51-55).
// Round 2: (i) the two IEEE divisions per (voxel, view) — 18 per voxel, the whole kernel at 17 % of the FP64 pipe —
// became one correctly-rounded reciprocal and two multiplications, with the exact divisions kept as a fallback for
// the ~1e-7 of projections whose quotient lands within 1e-7 of an integer (only there can the floor differ);
// (ii) when the masks are binary (0 / 255, what Grounded-SAM-2 writes: grounded_sam2_hf_model_imgs_MaskExtract.py:97,154)
// a vote is an exact small integer, so it is stored as one byte per voxel instead of a float64 (1 GiB at 512^3) and
// the count / compaction passes read one byte: SURVEY.md §8d's V * 1 bytes.  Generic masks keep the float64 sums.
#include "common.cuh"

namespace {

constexpr int VH_THREADS = 256;
constexpr int VH_MAX_VIEWS = 40;  // 40 * 96 B of matrices stay under the 4 KB kernel-parameter limit

struct VhMats {
    double m[VH_MAX_VIEWS][12];
};

__device__ __forceinline__ int cvt_i32_x86(double v) {
    // numpy's float64 -> int32 cast on x86 (cvttsd2si): anything unrepresentable becomes INT_MIN
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return (int)v;
}

// floor(RN(RN(p / d) + 1e-6)) as the reference computes it, from the reciprocal r = RN(1 / d): q' = RN(p r) is within
// a few ulp of p / d, so the floors agree unless RN(q' + 1e-6) lies within 1e-7 of an integer (or is huge / not finite);
// exactly those cases redo the IEEE division.
__device__ __forceinline__ int pixel_index(double p, double d, double r) {
    const double x = __dadd_rn(__dmul_rn(p, r), 1e-6);
    const double f = floor(x);
    const double frac = x - f;
    if (frac > 1e-7 && frac < 1.0 - 1e-7 && fabs(x) < 1.0e6) return (int)f;  // |q' - p/d| < 4 ulp(1e6) = 5e-10
    return cvt_i32_x86(floor(__dadd_rn(__ddiv_rn(p, d), 1e-6)));
}

__device__ __forceinline__ double voxel_vote(const VhMats& mats, int n_views, int H, int W,
                                             const uint8_t* __restrict__ masks, const double* __restrict__ lut,
                                             double x, double y, double z) {
    double vote = 0.0;
    for (int i = 0; i < n_views; ++i) {
        const double* m = mats.m[i];
        double p0 = fma(m[3], 1.0, fma(m[2], z, fma(m[1], y, m[0] * x)));
        double p1 = fma(m[7], 1.0, fma(m[6], z, fma(m[5], y, m[4] * x)));
        double p2 = fma(m[11], 1.0, fma(m[10], z, fma(m[9], y, m[8] * x)));
        const double r = __drcp_rn(p2);
        int u = pixel_index(p0, p2, r);
        int v = pixel_index(p1, p2, r);
        if (u < 0) u = 0;
        if (v < 0) v = 0;
        if (v >= H) { u = 0; v = 0; }
        if (u >= W) { u = 0; v = 0; }
        vote += lut[masks[((size_t)i * H + v) * W + u]];
    }
    return vote;
}

template <typename VT>
__global__ void __launch_bounds__(VH_THREADS)
vh_votes_kernel(VhMats mats, int n_views, int H, int W, const uint8_t* __restrict__ masks,
                const double* __restrict__ lut, const double* __restrict__ xs, int nx,
                const double* __restrict__ ys, int ny, const double* __restrict__ zs, int nz,
                VT* __restrict__ votes, unsigned long long* __restrict__ max_bits) {
    __shared__ double s_lut[256];
    __shared__ unsigned long long s_max[VH_THREADS / 32];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const int64_t V = (int64_t)nz * nx * ny;
    int64_t l = (int64_t)blockIdx.x * VH_THREADS + threadIdx.x;
    double vote = 0.0;
    if (l < V) {
        int iy = (int)(l % ny);
        int64_t r = l / ny;
        int ix = (int)(r % nx);
        int iz = (int)(r / nx);
        vote = voxel_vote(mats, n_views, H, W, masks, s_lut, xs[ix], ys[iy], zs[iz]);
        votes[l] = (VT)vote;  // uint8: binary masks, the vote is an exact integer <= n_views
    }
    // votes are >= 0, so the IEEE bit pattern is monotone: max over bits == max over values
    unsigned long long b = (unsigned long long)__double_as_longlong(vote);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = t > b ? t : b;
    }
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < VH_THREADS / 32; ++i) b = s_max[i] > b ? s_max[i] : b;
        if (b) atomicMax(max_bits, b);
    }
}

// per-block number of voxels with votes > iso (block = 1024 consecutive voxels)
constexpr int VH_CBLOCK = 1024;

template <typename VT>
__global__ void __launch_bounds__(VH_THREADS)
vh_count_kernel(int64_t V, const VT* __restrict__ votes, double iso, int32_t* __restrict__ block_counts) {
    __shared__ int s_cnt[VH_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * VH_CBLOCK;
    int c = 0;
#pragma unroll
    for (int k = 0; k < VH_CBLOCK / VH_THREADS; ++k) {
        int64_t l = base + k * VH_THREADS + threadIdx.x;
        c += (l < V && (double)votes[l] > iso) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < VH_THREADS / 32; ++i) t += s_cnt[i];
        block_counts[blockIdx.x] = t;
    }
}

// order-preserving compaction: points[block_offsets[b] + rank] = (x, y, z) of each occupied voxel
template <typename VT>
__global__ void __launch_bounds__(VH_THREADS)
vh_compact_kernel(int64_t V, const VT* __restrict__ votes, double iso, const int64_t* __restrict__ block_offsets,
                  const double* __restrict__ xs, int nx, const double* __restrict__ ys, int ny,
                  const double* __restrict__ zs, double* __restrict__ points, int64_t* __restrict__ indices) {
    __shared__ int s_warp[VH_THREADS / 32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t base = (int64_t)blockIdx.x * VH_CBLOCK;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int k = 0; k < VH_CBLOCK / VH_THREADS; ++k) {
        int64_t l = base + k * VH_THREADS + threadIdx.x;
        bool occ = (l < V) && (double)votes[l] > iso;
        unsigned ballot = __ballot_sync(0xffffffffu, occ);
        if (lane == 0) s_warp[warp] = __popc(ballot);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (occ) {
            int64_t dst = block_offsets[blockIdx.x] + before + __popc(ballot & ((1u << lane) - 1u));
            int iy = (int)(l % ny);
            int64_t r = l / ny;
            int ix = (int)(r % nx);
            int iz = (int)(r / nx);
            points[3 * dst + 0] = xs[ix];
            points[3 * dst + 1] = ys[iy];
            points[3 * dst + 2] = zs[iz];
            if (indices) indices[dst] = l;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < VH_THREADS / 32; ++w) t += s_warp[w];
            s_base += t;
        }
        __syncthreads();
    }
}

}  // namespace

FSB_API int fsb_vh_max_views(void) { return VH_MAX_VIEWS; }
FSB_API int fsb_vh_count_block(void) { return VH_CBLOCK; }

// mats_host: HOST pointer to [n_views,12] doubles (K @ [R|t], row-major 3x4); passed by value to the kernel.
// max_bits: device u64, zero-initialised by the caller; receives the bit pattern of max(votes).
// votes_u8 != 0: `votes` is uint8[V] (caller guarantees binary 0 / 255 masks and n_views <= 255), else float64[V].
FSB_API int fsb_vh_votes(int n_views, int H, int W, const uint8_t* masks, const double* mats_host,
                         const double* lut, const double* xs, int nx, const double* ys, int ny, const double* zs,
                         int nz, void* votes, int votes_u8, uint64_t* max_bits, void* stream) {
    if (n_views <= 0 || n_views > VH_MAX_VIEWS || H <= 0 || W <= 0 || nx <= 0 || ny <= 0 || nz < 0) return FSB_E_ARG;
    if (nz == 0) return 0;
    VhMats mats;
    for (int i = 0; i < n_views; ++i)
        for (int k = 0; k < 12; ++k) mats.m[i][k] = mats_host[i * 12 + k];
    int64_t V = (int64_t)nz * nx * ny;
    if (votes_u8)
        vh_votes_kernel<uint8_t><<<fsb_div_up(V, VH_THREADS), VH_THREADS, 0, (cudaStream_t)stream>>>(
            mats, n_views, H, W, masks, lut, xs, nx, ys, ny, zs, nz, (uint8_t*)votes, (unsigned long long*)max_bits);
    else
        vh_votes_kernel<double><<<fsb_div_up(V, VH_THREADS), VH_THREADS, 0, (cudaStream_t)stream>>>(
            mats, n_views, H, W, masks, lut, xs, nx, ys, ny, zs, nz, (double*)votes, (unsigned long long*)max_bits);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_vh_count(int64_t V, const void* votes, int votes_u8, double iso, int32_t* block_counts, void* stream) {
    if (V < 0) return FSB_E_ARG;
    if (V == 0) return 0;
    if (votes_u8)
        vh_count_kernel<uint8_t><<<fsb_div_up(V, VH_CBLOCK), VH_THREADS, 0, (cudaStream_t)stream>>>(
            V, (const uint8_t*)votes, iso, block_counts);
    else
        vh_count_kernel<double><<<fsb_div_up(V, VH_CBLOCK), VH_THREADS, 0, (cudaStream_t)stream>>>(
            V, (const double*)votes, iso, block_counts);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_vh_compact(int64_t V, const void* votes, int votes_u8, double iso, const int64_t* block_offsets,
                           const double* xs, int nx, const double* ys, int ny, const double* zs, double* points,
                           int64_t* indices, void* stream) {
    if (V < 0) return FSB_E_ARG;
    if (V == 0) return 0;
    if (votes_u8)
        vh_compact_kernel<uint8_t><<<fsb_div_up(V, VH_CBLOCK), VH_THREADS, 0, (cudaStream_t)stream>>>(
            V, (const uint8_t*)votes, iso, block_offsets, xs, nx, ys, ny, zs, points, indices);
    else
        vh_compact_kernel<double><<<fsb_div_up(V, VH_CBLOCK), VH_THREADS, 0, (cudaStream_t)stream>>>(
            V, (const double*)votes, iso, block_offsets, xs, nx, ys, ny, zs, points, indices);
    FSB_LAUNCH_CHECK();
    return 0;
}
