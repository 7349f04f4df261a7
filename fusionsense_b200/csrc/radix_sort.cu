// radix_sort.cu — stable LSD radix sort of (uint64 key, int32 value) pairs, 8-bit digits,
// one read + one write of the pairs per pass ("onesweep": chained-scan with decoupled look-back).
//
// Replaces the cub::DeviceRadixSort::SortPairs call inside gsplat 1.0.0's isect_tiles and the
// torch.sort of the legacy bin_and_sort_gaussians, reached from
// /root/reference/dn_splatter/dn_model.py:570-591 and :644-653.  Only key bits [0, end_bit) are
// sorted (32 depth bits + tile bits + camera bits).  Stability is part of the contract: ties keep
// emission order, which is what makes the sorted order bit-exact against the oracle.
//
// HBM-bound: algorithmic traffic = n * 12 B * 2 per pass + one n * 8 B histogram read.
#include "common.cuh"

namespace {

constexpr int RADIX = 256;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int KPT = 16;                         // keys per thread
constexpr int SORT_TILE = SORT_THREADS * KPT;   // 4096 pairs per CTA
constexpr int MAX_PASSES = 8;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_INC = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VAL_MASK = ~FLAG_MASK;

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(int64_t n, const int64_t* __restrict__ n_dev, const uint64_t* __restrict__ keys, int passes,
                  uint32_t* __restrict__ hist) {
    n = fsb_eff_n(n, n_dev);
    __shared__ uint32_t sh[MAX_PASSES][RADIX];
    for (int i = threadIdx.x; i < MAX_PASSES * RADIX; i += SORT_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    int lane = threadIdx.x & 31;
    int64_t stride = (int64_t)gridDim.x * SORT_THREADS;
    // every lane of a warp runs the same trip count so the full-mask match below is legal
    int64_t start = (int64_t)blockIdx.x * SORT_THREADS + (threadIdx.x & ~31);
    for (int64_t base = start; base < n; base += stride) {
        int64_t i = base + lane;
        bool valid = i < n;
        uint64_t key = valid ? keys[i] : 0;
        for (int p = 0; p < passes; ++p) {
            uint32_t d = valid ? (uint32_t)((key >> (8 * p)) & 0xff) : 0xffffffffu;
            uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&sh[p][d], __popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += SORT_THREADS) {
        uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
onesweep_pass_kernel(int64_t n, const int64_t* __restrict__ n_dev, const uint64_t* __restrict__ keys_in,
                     const int32_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                     const uint32_t* __restrict__ pass_hist, volatile uint32_t* status, uint32_t* tile_counter,
                     int shift) {
    __shared__ uint32_t warp_hist[SORT_WARPS][RADIX];
    __shared__ uint32_t digit_base[RADIX];
    __shared__ uint32_t scan_tmp[SORT_WARPS];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    n = fsb_eff_n(n, n_dev);
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    // static-capacity mode launches tiles for the capacity; a tile past the true count holds nothing and no later
    // tile exists that would look back at it
    if ((int64_t)tile * SORT_TILE >= n) return;
    const int64_t tile_base = (int64_t)tile * SORT_TILE + (int64_t)warp * (32 * KPT);

    uint64_t key[KPT];
    int32_t val[KPT];
    uint32_t rank[KPT];
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        int64_t i = tile_base + k * 32 + lane;
        bool valid = i < n;
        key[k] = valid ? keys_in[i] : ~0ull;
        val[k] = valid ? vals_in[i] : 0;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        int64_t i = tile_base + k * 32 + lane;
        bool valid = i < n;
        uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
        uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
        int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = warp_hist[warp][d];
            warp_hist[warp][d] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[k] = pre + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: exclusive prefix over the CTA's warps, chained scan over earlier tiles
    {
        const int d = tid;
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = total;
            total += c;
        }
        volatile uint32_t* my = status + (size_t)tile * RADIX + d;
        *my = (tile == 0 ? FLAG_INC : FLAG_AGG) | total;
        uint32_t excl = 0;
        if (tile > 0) {
            int64_t t = (int64_t)tile - 1;
            while (true) {
                uint32_t v = status[(size_t)t * RADIX + d];
                uint32_t f = v & FLAG_MASK;
                if (f == 0) continue;  // predecessor has not published yet
                excl += v & VAL_MASK;
                if (f == FLAG_INC) break;
                --t;
            }
            *my = FLAG_INC | (excl + total);
        }
        // exclusive scan of the pass histogram over digits -> global base of digit d
        uint32_t h = pass_hist[d];
        uint32_t inc = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t2 = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t2;
        }
        if (lane == 31) scan_tmp[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            if (w < warp) wbase += scan_tmp[w];
        digit_base[d] = wbase + inc - h + excl;
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        int64_t i = tile_base + k * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
            uint32_t pos = digit_base[d] + warp_hist[warp][d] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = val[k];
        }
    }
}

}  // namespace

static inline int64_t sort_num_tiles(int64_t n) { return n > 0 ? (n + SORT_TILE - 1) / SORT_TILE : 1; }
static inline int sort_num_passes(int end_bit) { return (end_bit + 7) / 8; }

FSB_API size_t fsb_radix_sort_workspace(int64_t n, int end_bit) {
    int passes = sort_num_passes(end_bit);
    size_t bytes = (size_t)passes * RADIX * sizeof(uint32_t);                  // histograms
    bytes += fsb_align_up((size_t)passes * sizeof(uint32_t), 256);             // tile counters
    bytes += (size_t)passes * (size_t)sort_num_tiles(n) * RADIX * sizeof(uint32_t);  // look-back status
    return fsb_align_up(bytes, 256);
}

// Sorts pairs by key bits [0, end_bit).  Buffers A (input, clobbered) and B ping-pong;
// *result_in_b tells the caller which one holds the sorted pairs (it depends only on end_bit).
// n_dev (nullable): device pointer to the true count; n is then the capacity (see common.cuh).
FSB_API int fsb_radix_sort_pairs(int64_t n, const int64_t* n_dev, int end_bit, uint64_t* keys_a, int32_t* vals_a,
                                 uint64_t* keys_b, int32_t* vals_b, void* workspace, size_t workspace_bytes,
                                 int* result_in_b, void* stream) {
    if (n < 0 || n >= (1ll << 30) || end_bit < 1 || end_bit > 64) return FSB_E_ARG;
    int passes = sort_num_passes(end_bit);
    if (passes > MAX_PASSES) return FSB_E_ARG;
    if (workspace_bytes < fsb_radix_sort_workspace(n, end_bit)) return FSB_E_ARG;
    if (result_in_b) *result_in_b = passes & 1;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t tiles = sort_num_tiles(n);
    uint32_t* hist = (uint32_t*)workspace;
    uint32_t* counters = hist + (size_t)passes * RADIX;
    uint32_t* status = (uint32_t*)((char*)counters + fsb_align_up((size_t)passes * sizeof(uint32_t), 256));
    FSB_CUDA(cudaMemsetAsync(workspace, 0, fsb_radix_sort_workspace(n, end_bit), st));
    int hist_blocks = (int)(tiles < FSB_NUM_SMS * 8 ? tiles : FSB_NUM_SMS * 8);
    radix_hist_kernel<<<hist_blocks, SORT_THREADS, 0, st>>>(n, n_dev, keys_a, passes, hist);
    FSB_LAUNCH_CHECK();
    uint64_t* kin = keys_a; int32_t* vin = vals_a;
    uint64_t* kout = keys_b; int32_t* vout = vals_b;
    for (int p = 0; p < passes; ++p) {
        onesweep_pass_kernel<<<(unsigned)tiles, SORT_THREADS, 0, st>>>(
            n, n_dev, kin, vin, kout, vout, hist + (size_t)p * RADIX, status + (size_t)p * tiles * RADIX, counters + p,
            8 * p);
        FSB_LAUNCH_CHECK();
        uint64_t* tk = kin; kin = kout; kout = tk;
        int32_t* tv = vin; vin = vout; vout = tv;
    }
    return 0;
}
