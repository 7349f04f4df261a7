// isect.cu — tile-intersection bookkeeping: exclusive scan of tiles-per-Gaussian, key/value emission,
// and per-(camera, tile) range build from the sorted keys.
//
// Replaces gsplat 1.0.0's isect_tiles (second pass) + torch.cumsum + isect_offset_encode, and the
// legacy map_gaussian_to_intersects / get_tile_bin_edges (Appendix A.4/A.6),
// reached from /root/reference/dn_splatter/dn_model.py:570-591 and :644-653.
// Integer work, bit-exact against oracle/gsplat_ref.py.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 counts per block

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one int64 per thread (256 threads); returns the exclusive prefix and the total
__device__ __forceinline__ int64_t block_excl_scan(int64_t v, int64_t* total, int64_t* smem /*>=9*/) {
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    int64_t inc = warp_incl_scan(v);
    __syncthreads();
    if (l == 31) smem[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
#pragma unroll
        for (int i = 0; i < SCAN_THREADS / 32; ++i) {
            int64_t t = smem[i];
            smem[i] = run;
            run += t;
        }
        smem[8] = run;
    }
    __syncthreads();
    *total = smem[8];
    return smem[w] + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums_kernel(int64_t M, const int32_t* __restrict__ counts, const int32_t* __restrict__ perm,
                       int64_t* __restrict__ block_sums) {
    __shared__ int64_t sm[9];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < M) s += counts[perm ? perm[base + i] : base + i];
    int64_t total;
    block_excl_scan(s, &total, sm);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// one block scans the block sums in place (exclusive) and writes the grand total
__global__ void __launch_bounds__(SCAN_THREADS)
scan_spine_kernel(int64_t n_blocks, int64_t* __restrict__ block_sums, int64_t* __restrict__ total_out) {
    __shared__ int64_t sm[9];
    int64_t carry = 0;
    for (int64_t base = 0; base < n_blocks; base += SCAN_THREADS) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < n_blocks ? block_sums[i] : 0;
        int64_t total;
        int64_t ex = block_excl_scan(v, &total, sm);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(int64_t M, const int32_t* __restrict__ counts, const int32_t* __restrict__ perm,
                  const int64_t* __restrict__ block_sums, int64_t* __restrict__ offsets) {
    __shared__ int64_t sm[9];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < M) ? counts[perm ? perm[base + i] : base + i] : 0;
        s += v[i];
    }
    int64_t total;
    int64_t run = block_excl_scan(s, &total, sm) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < M) offsets[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(256)
isect_emit_kernel(int C, int N, const float* __restrict__ means2d, const int32_t* __restrict__ radii,
                  const float* __restrict__ depths, const int64_t* __restrict__ offsets, int tile_size, int tile_w,
                  int tile_h, int tile_bits, int legacy_bbox, const int64_t* __restrict__ n_dev, int64_t capacity,
                  int32_t* __restrict__ overflow_flag, int64_t* __restrict__ isect_ids,
                  int32_t* __restrict__ flatten_ids) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)C * N) return;
    // static-capacity mode: entries past the end of the buffers are dropped and the step is flagged as invalid
    const int64_t limit = n_dev ? capacity : INT64_MAX;
    if (idx == 0 && n_dev && overflow_flag && *n_dev > capacity) *overflow_flag = 1;
    if (n_dev && *n_dev == 0) return;  // nothing to emit (also the "lists are shared" gate of fsb_isect_share_gate)
    int r = radii[idx];
    if (r <= 0) return;
    float2 m = reinterpret_cast<const float2*>(means2d)[idx];
    float ts = (float)tile_size;
    float tr = (float)r / ts;
    float tx = m.x / ts, ty = m.y / ts;
    int ax, ay, bx, by;
    if (legacy_bbox) {
        ax = (int)(tx - tr); ay = (int)(ty - tr);
        bx = (int)(tx + tr + 1.f); by = (int)(ty + tr + 1.f);
    } else {
        ax = (int)floorf(tx - tr); ay = (int)floorf(ty - tr);
        bx = (int)ceilf(tx + tr); by = (int)ceilf(ty + tr);
    }
    ax = min(max(0, ax), tile_w); ay = min(max(0, ay), tile_h);
    bx = min(max(0, bx), tile_w); by = min(max(0, by), tile_h);
    int64_t c = idx / N;
    int64_t cam_part = c << (32 + tile_bits);
    int64_t depth_part = (int64_t)(uint32_t)__float_as_int(depths[idx]);
    // gsplat sign-extends the int32 view of the depth; depths are > near_plane > 0 so both agree
    int64_t cur = offsets[idx];
    int32_t val = (int32_t)idx;
    for (int i = ay; i < by; ++i)
        for (int j = ax; j < bx; ++j) {
            int64_t tile = (int64_t)i * tile_w + j;
            if (cur < limit) {
                isect_ids[cur] = cam_part | (tile << 32) | depth_part;
                flatten_ids[cur] = val;
            }
            ++cur;
        }
}

// isect_offsets[c, ty, tx] = first sorted position whose (camera, tile) id is >= this one
__global__ void __launch_bounds__(256)
isect_offsets_kernel(int64_t n_isects, const int64_t* __restrict__ n_dev, const int64_t* __restrict__ sorted_ids, int C,
                     int n_tiles, int tile_bits, int32_t* __restrict__ offsets) {
    n_isects = fsb_eff_n(n_isects, n_dev);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total_tiles = (int64_t)C * n_tiles;
    if (n_isects == 0) {
        if (i < total_tiles) offsets[i] = 0;
        return;
    }
    if (i >= n_isects) return;
    int64_t tile_mask = ((int64_t)1 << tile_bits) - 1;
    int64_t k = sorted_ids[i] >> 32;
    int64_t cur = (k >> tile_bits) * n_tiles + (k & tile_mask);
    if (i == 0) {
        for (int64_t t = 0; t <= cur; ++t) offsets[t] = 0;
    }
    if (i == n_isects - 1) {
        for (int64_t t = cur + 1; t < total_tiles; ++t) offsets[t] = (int32_t)n_isects;
    }
    if (i > 0) {
        int64_t kp = sorted_ids[i - 1] >> 32;
        int64_t prev = (kp >> tile_bits) * n_tiles + (kp & tile_mask);
        for (int64_t t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
    }
}

// ---- static-capacity mode: device-side decision to share the first pass's sorted lists with the legacy pass ----
// The 0.1.x bbox of a Gaussian is a superset of its 1.0 bbox (truncation vs floor agree after the clamp at 0;
// (int)(x + 1) >= ceil(x)), so equal intersection totals <=> identical tile sets <=> identical keys and sort order.
__global__ void isect_share_gate_kernel(const int64_t* __restrict__ n_first, const int64_t* __restrict__ n_legacy,
                                        int64_t* __restrict__ gate) {
    if (threadIdx.x == 0 && blockIdx.x == 0) gate[0] = (*n_first == *n_legacy) ? 0 : *n_legacy;
}

__global__ void __launch_bounds__(256)
isect_share_copy_kernel(const int64_t* __restrict__ gate, const int64_t* __restrict__ n_list, int64_t capacity,
                        const int32_t* __restrict__ src_flat, const int32_t* __restrict__ src_offsets,
                        int64_t n_offsets, int32_t* __restrict__ dst_flat, int32_t* __restrict__ dst_offsets) {
    if (*gate != 0) return;  // the legacy pass built its own lists
    int64_t n = *n_list;
    if (n > capacity) n = capacity;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst_flat[i] = src_flat[i];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_offsets; i += stride)
        dst_offsets[i] = src_offsets[i];
}

}  // namespace

// Static-capacity mode (no host read), legacy normals pass right after a rasterization() on the same Gaussians:
// gate[0] = 0 when the legacy bbox rule yields the same number of intersections as the 1.0 rule (*n_first ==
// *n_legacy: the sorted lists are then identical and can be shared), else *n_legacy.  Passed as `n_dev` to
// fsb_isect_emit / fsb_radix_sort_pairs / fsb_isect_offsets it turns them into no-ops in the shared case.
// Replaces the host-side decision the eager path takes in gsplat/cuda_legacy/_wrapper.py (one D2H read there).
FSB_API int fsb_isect_share_gate(const int64_t* n_first, const int64_t* n_legacy, int64_t* gate, void* stream) {
    if (!n_first || !n_legacy || !gate) return FSB_E_ARG;
    isect_share_gate_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(n_first, n_legacy, gate);
    FSB_LAUNCH_CHECK();
    return 0;
}

// Shared case (gate[0] == 0): dst_flat[0 .. min(*n_list, capacity)) = src_flat, dst_offsets[0 .. n_offsets) =
// src_offsets; otherwise nothing is touched.
FSB_API int fsb_isect_share_copy(const int64_t* gate, const int64_t* n_list, int64_t capacity, const int32_t* src_flat,
                                 const int32_t* src_offsets, int64_t n_offsets, int32_t* dst_flat,
                                 int32_t* dst_offsets, void* stream) {
    if (!gate || !n_list || capacity < 0 || n_offsets < 0) return FSB_E_ARG;
    int64_t work = capacity > n_offsets ? capacity : n_offsets;
    if (work == 0) return 0;
    int blocks = fsb_div_up(work, 256 * 4);
    if (blocks > FSB_NUM_SMS * 8) blocks = FSB_NUM_SMS * 8;
    isect_share_copy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(gate, n_list, capacity, src_flat, src_offsets,
                                                                      n_offsets, dst_flat, dst_offsets);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API size_t fsb_isect_scan_workspace(int64_t M) {
    int64_t n_blocks = (M + SCAN_TILE - 1) / SCAN_TILE;
    if (n_blocks < 1) n_blocks = 1;
    return (size_t)n_blocks * sizeof(int64_t);
}

// offsets[i] = sum(counts[perm[0..i)]) as int64 (perm nullable = identity); *total_dev = sum of all counts
static int scan_launch(int64_t M, const int32_t* counts, const int32_t* perm, int64_t* offsets, int64_t* total_dev,
                       void* workspace, size_t workspace_bytes, void* stream) {
    if (M < 0) return FSB_E_ARG;
    if (workspace_bytes < fsb_isect_scan_workspace(M)) return FSB_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        FSB_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st));
        return 0;
    }
    int64_t n_blocks = (M + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* sums = (int64_t*)workspace;
    scan_block_sums_kernel<<<(unsigned)n_blocks, SCAN_THREADS, 0, st>>>(M, counts, perm, sums);
    FSB_LAUNCH_CHECK();
    scan_spine_kernel<<<1, SCAN_THREADS, 0, st>>>(n_blocks, sums, total_dev);
    FSB_LAUNCH_CHECK();
    scan_apply_kernel<<<(unsigned)n_blocks, SCAN_THREADS, 0, st>>>(M, counts, perm, sums, offsets);
    FSB_LAUNCH_CHECK();
    return 0;
}

// offsets[i] = sum(counts[0..i)) as int64 ; *total_dev = sum of all counts (device scalar)
FSB_API int fsb_isect_scan(int64_t M, const int32_t* counts, int64_t* offsets, int64_t* total_dev, void* workspace,
                           size_t workspace_bytes, void* stream) {
    return scan_launch(M, counts, nullptr, offsets, total_dev, workspace, workspace_bytes, stream);
}

// The same scan taken in the order `perm` (int32[M], a permutation of 0 .. M-1): offsets[i] = sum(counts[perm[0..i)]).
// Two-level binning: perm = the Gaussians in depth order, offsets[i] = where the entries of the i-th nearest start.
FSB_API int fsb_isect_scan_perm(int64_t M, const int32_t* counts, const int32_t* perm, int64_t* offsets,
                                int64_t* total_dev, void* workspace, size_t workspace_bytes, void* stream) {
    if (!perm) return FSB_E_ARG;
    return scan_launch(M, counts, perm, offsets, total_dev, workspace, workspace_bytes, stream);
}

FSB_API int fsb_isect_emit(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                           const int64_t* offsets, int tile_size, int tile_w, int tile_h, int tile_bits,
                           int legacy_bbox, const int64_t* n_dev, int64_t capacity, int32_t* overflow_flag,
                           int64_t* isect_ids, int32_t* flatten_ids, void* stream) {
    if (C <= 0 || N < 0 || tile_size <= 0 || tile_bits < 0 || tile_bits > 30) return FSB_E_ARG;
    if (n_dev && capacity < 0) return FSB_E_ARG;
    if ((int64_t)C * N > 0x7fffffffLL) return FSB_E_ARG;  // flatten_ids are int32 (same limit as gsplat)
    if (N == 0) return 0;
    int64_t total = (int64_t)C * N;
    isect_emit_kernel<<<fsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        C, N, means2d, radii, depths, offsets, tile_size, tile_w, tile_h, tile_bits, legacy_bbox, n_dev, capacity,
        overflow_flag, isect_ids, flatten_ids);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_isect_offsets(int64_t n_isects, const int64_t* n_dev, const int64_t* sorted_ids, int C, int n_tiles,
                              int tile_bits, int32_t* offsets, void* stream) {
    if (n_isects < 0 || n_isects > 0x7fffffffLL || C <= 0 || n_tiles <= 0) return FSB_E_ARG;
    int64_t work = n_isects > 0 ? n_isects : (int64_t)C * n_tiles;
    if (n_dev && work < (int64_t)C * n_tiles) work = (int64_t)C * n_tiles;  // the true count may be zero
    isect_offsets_kernel<<<fsb_div_up(work, 256), 256, 0, (cudaStream_t)stream>>>(n_isects, n_dev, sorted_ids, C,
                                                                                 n_tiles, tile_bits, offsets);
    FSB_LAUNCH_CHECK();
    return 0;
}
