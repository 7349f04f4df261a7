// radix_sort.cu — stable LSD radix sort of (uint64 key, int32 value) pairs, 8- or 9-bit digits,
// one read + one write of the pairs per pass ("onesweep": chained-scan with decoupled look-back).
// The digit width is the one that needs fewer passes for the caller's end_bit: the intersection keys of a
// single-camera 640x480 .. 1080p frame have 43 .. 45 significant bits, which is 6 passes of 8 bits but 5 of 9.
//
// Replaces the cub::DeviceRadixSort::SortPairs call inside gsplat 1.0.0's isect_tiles and the
// torch.sort of the legacy bin_and_sort_gaussians, reached from
// /root/reference/dn_splatter/dn_model.py:570-591 and :644-653.  Only key bits [0, end_bit) are
// sorted (32 depth bits + tile bits + camera bits).  Stability is part of the contract: ties keep
// emission order, which is what makes the sorted order bit-exact against the oracle.
//
// HBM-bound: algorithmic traffic = n * 12 B * 2 per pass + one n * 8 B histogram read.  Pairs are reordered by digit
// in shared memory before they are written, so every digit's run of a 4096-pair tile leaves as whole sectors.
//
// fsb_radix_sort_keys is the keys-only form on a bit window [begin_bit, end_bit): the second level of the two-level
// binning (ops.isect_tiles_depth_first): the (camera, tile) id sits in the high word, the Gaussian id rides in the low
// word, and because the entries were emitted in depth order a stable sort on the 11 .. 13 tile bits alone (2 passes,
// 8 B per entry) gives the order the reference gets from sorting the full 64-bit (tile | depth) keys (5 passes, 12 B).
#include "common.cuh"

namespace {

constexpr int MAX_RADIX = 512;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int KPT = 16;                         // keys per thread
constexpr int SORT_TILE = SORT_THREADS * KPT;   // 4096 pairs per CTA
constexpr int MAX_PASSES = 8;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_INC = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VAL_MASK = ~FLAG_MASK;

// digit of a pass: key bits [shift, shift + width)
__device__ __forceinline__ uint32_t digit_of(uint64_t key, int shift, int width) {
    return (uint32_t)((key >> shift) & ((1u << width) - 1u));
}

template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(int64_t n, const int64_t* __restrict__ n_dev, const uint64_t* __restrict__ keys, int passes,
                  int begin_bit, int digit_w, int end_bit, uint32_t* __restrict__ hist) {
    constexpr int RADIX = 1 << BITS;
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
            const int shift = begin_bit + digit_w * p;
            uint32_t d = valid ? digit_of(key, shift, min(digit_w, end_bit - shift)) : 0xffffffffu;
            if (shift + digit_w <= 23) {
                // digit inside the mantissa of the depth: values are spread, same-address conflicts are rare, and a
                // plain shared-memory atomic is cheaper than MATCH.ANY (the unit the ranking loop also leans on)
                if (valid) atomicAdd(&sh[p][d], 1u);
            } else {
                // exponent / tile / camera bits: neighbouring keys share them, so combine equal digits first
                uint32_t peers = __match_any_sync(0xffffffffu, d);
                if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&sh[p][d], __popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += SORT_THREADS) {
        uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

// PAIRS: (key, value) pairs; else keys only, and `low32_out` (nullable, last pass) also receives the low word of every
// key at its sorted position.
template <int BITS, bool PAIRS>
__global__ void __launch_bounds__(SORT_THREADS, PAIRS ? 2 : 3)
onesweep_pass_kernel(int64_t n, const int64_t* __restrict__ n_dev, const uint64_t* __restrict__ keys_in,
                     const int32_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, int32_t* __restrict__ vals_out, int32_t* __restrict__ low32_out,
                     const uint32_t* __restrict__ pass_hist, volatile uint32_t* status, uint32_t* tile_counter,
                     int shift, int width) {
    constexpr int RADIX = 1 << BITS;
    constexpr int DPT = RADIX / SORT_THREADS;  // digits owned by a thread: DPT t .. DPT t + DPT - 1
    static_assert(DPT >= 1 && RADIX <= MAX_RADIX, "digit width");
    extern __shared__ __align__(16) unsigned char sort_smem[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(sort_smem);                        // [SORT_TILE]
    int32_t* s_vals = reinterpret_cast<int32_t*>(sort_smem + SORT_TILE * 8);          // [SORT_TILE], PAIRS only
    uint32_t (*warp_hist)[RADIX] =
        reinterpret_cast<uint32_t (*)[RADIX]>(sort_smem + SORT_TILE * (PAIRS ? 12 : 8));  // [SORT_WARPS]
    __shared__ uint32_t digit_base[RADIX];
    __shared__ uint32_t cta_total[RADIX];
    __shared__ uint32_t local_start[RADIX];
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
    int32_t val[PAIRS ? KPT : 1];
    uint32_t rank[KPT];
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        int64_t i = tile_base + k * 32 + lane;
        bool valid = i < n;
        key[k] = valid ? keys_in[i] : ~0ull;
        if (PAIRS) val[k] = valid ? vals_in[i] : 0;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        int64_t i = tile_base + k * 32 + lane;
        bool valid = i < n;
        uint32_t d = digit_of(key[k], shift, width);
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

    // a thread owns DPT adjacent digits: exclusive prefix over the CTA's warps, chained scan over earlier tiles
    {
        uint32_t total[DPT], excl[DPT];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = tid * DPT + j;
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) {
                uint32_t c = warp_hist[w][d];
                warp_hist[w][d] = tot;
                tot += c;
            }
            total[j] = tot;
            cta_total[d] = tot;
            status[(size_t)tile * RADIX + d] = (tile == 0 ? FLAG_INC : FLAG_AGG) | tot;
        }
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = tid * DPT + j;
            uint32_t ex = 0;
            if (tile > 0) {
                // Look-back in windows of LB predecessors: the LB status words are requested together and then
                // consumed nearest-first up to the first one that is not published yet (re-read from there) or
                // the first inclusive prefix.  At these sizes every tile is resident at once, so a one-word-at-a-
                // time walk is a serial chain of ~tiles / 2 dependent L2 round trips per pass (r01h: 14 us per
                // pass for 104 tiles); the window divides that chain by LB.
                constexpr int LB = 8;
                int64_t t = (int64_t)tile - 1;
                bool found = false;
                while (!found) {
                    uint32_t v[LB];
#pragma unroll
                    for (int u = 0; u < LB; ++u) {
                        const int64_t tt = t - u;
                        v[u] = FLAG_INC + 0u;  // before tile 0: an inclusive prefix of 0
                        if (tt >= 0) v[u] = status[(size_t)tt * RADIX + d];
                    }
                    int adv = 0;
                    bool stop = false;
#pragma unroll
                    for (int u = 0; u < LB; ++u) {
                        if (stop) continue;
                        const uint32_t f = v[u] & FLAG_MASK;
                        if (f == 0) {
                            stop = true;  // not published yet: everything behind it is read again
                        } else {
                            ex += v[u] & VAL_MASK;
                            adv = u + 1;
                            if (f == FLAG_INC) { found = true; stop = true; }
                        }
                    }
                    t -= adv;
                }
                status[(size_t)tile * RADIX + d] = FLAG_INC | (ex + total[j]);
            }
            excl[j] = ex;
        }
        // exclusive scan of the pass histogram over digits -> global base of each digit
        uint32_t h[DPT];
        uint32_t hsum = 0;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            h[j] = pass_hist[tid * DPT + j];
            hsum += h[j];
        }
        uint32_t inc = hsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t2 = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t2;
        }
        if (lane == 31) scan_tmp[warp] = inc;
        __syncthreads();
        uint32_t run = inc - hsum;  // digits of earlier threads of this warp
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            if (w < warp) run += scan_tmp[w];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            digit_base[tid * DPT + j] = run + excl[j];
            run += h[j];
        }
    }
    __syncthreads();

    // Reorder inside the CTA before the scatter: every pair goes to its CTA-local sorted slot in shared memory, then
    // consecutive threads write consecutive slots, so each digit's run leaves as one contiguous piece.  A direct
    // scatter writes 8 + 4 bytes per pair into 32-byte sectors all over the output; the L2 then has to fetch every
    // partially written sector from DRAM first (r02b ncu: 151 MB read per pass for 75 MB of pairs).
    // local_start[d] = exclusive prefix of the CTA's digit totals; computed by the owner threads from cta_total.
    {
        uint32_t tot[DPT];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            tot[j] = cta_total[tid * DPT + j];
            sum += tot[j];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t2 = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t2;
        }
        if (lane == 31) scan_tmp[warp] = inc;
        __syncthreads();
        uint32_t run = inc - sum;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            if (w < warp) run += scan_tmp[w];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            local_start[tid * DPT + j] = run;
            run += tot[j];
        }
    }
    __syncthreads();
    const int64_t cta_first = (int64_t)tile * SORT_TILE;
    const int cta_n = (int)min((int64_t)SORT_TILE, n - cta_first);
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        int64_t i = tile_base + k * 32 + lane;
        if (i < n) {
            uint32_t d = digit_of(key[k], shift, width);
            uint32_t slot = local_start[d] + warp_hist[warp][d] + rank[k];
            s_keys[slot] = key[k];
            if (PAIRS) s_vals[slot] = val[k];
        }
    }
    __syncthreads();
    for (int slot = tid; slot < cta_n; slot += SORT_THREADS) {
        const uint64_t kk = s_keys[slot];
        const uint32_t d = digit_of(kk, shift, width);
        const uint32_t pos = digit_base[d] + ((uint32_t)slot - local_start[d]);
        keys_out[pos] = kk;
        if (PAIRS) vals_out[pos] = s_vals[slot];
        else if (low32_out != nullptr) low32_out[pos] = (int32_t)(uint32_t)kk;
    }
}

}  // namespace

static inline int64_t sort_num_tiles(int64_t n) { return n > 0 ? (n + SORT_TILE - 1) / SORT_TILE : 1; }

// Pass plan of a window of `nbits` key bits: 9-bit digits when they save a pass (41..45, 49..54, 57..63 bits), else
// 8-bit; the window is then split into equal digits (45 bits -> 5 x 9, 32 -> 4 x 8, 13 -> 2 x 7).
struct SortPlan {
    int radix_bits;  // kernel instantiation (table sizes)
    int passes;
    int digit_w;     // bits per pass (the last pass takes what is left)
};
static inline SortPlan sort_plan(int nbits) {
    SortPlan pl;
    pl.radix_bits = (nbits + 8) / 9 < (nbits + 7) / 8 ? 9 : 8;
    pl.passes = (nbits + pl.radix_bits - 1) / pl.radix_bits;
    pl.digit_w = (nbits + pl.passes - 1) / pl.passes;
    return pl;
}

static size_t sort_workspace_bytes(int64_t n, int nbits) {
    const SortPlan pl = sort_plan(nbits);
    const size_t radix = (size_t)1 << pl.radix_bits;
    size_t bytes = (size_t)pl.passes * radix * sizeof(uint32_t);                        // histograms
    bytes += fsb_align_up((size_t)pl.passes * sizeof(uint32_t), 256);                   // tile counters
    bytes += (size_t)pl.passes * (size_t)sort_num_tiles(n) * radix * sizeof(uint32_t);  // look-back status
    return fsb_align_up(bytes, 256);
}

FSB_API size_t fsb_radix_sort_workspace(int64_t n, int end_bit) { return sort_workspace_bytes(n, end_bit); }
FSB_API size_t fsb_radix_sort_keys_workspace(int64_t n, int begin_bit, int end_bit) {
    return sort_workspace_bytes(n, end_bit - begin_bit);
}

template <int BITS, bool PAIRS>
static int sort_launch(int64_t n, const int64_t* n_dev, int begin_bit, int end_bit, SortPlan pl, uint64_t* keys_a,
                       int32_t* vals_a, uint64_t* keys_b, int32_t* vals_b, int32_t* low32_out, void* workspace,
                       cudaStream_t st) {
    constexpr int RADIX = 1 << BITS;
    const int passes = pl.passes;
    int64_t tiles = sort_num_tiles(n);
    uint32_t* hist = (uint32_t*)workspace;
    uint32_t* counters = hist + (size_t)passes * RADIX;
    uint32_t* status = (uint32_t*)((char*)counters + fsb_align_up((size_t)passes * sizeof(uint32_t), 256));
    int hist_blocks = (int)(tiles < FSB_NUM_SMS * 8 ? tiles : FSB_NUM_SMS * 8);
    radix_hist_kernel<BITS><<<hist_blocks, SORT_THREADS, 0, st>>>(n, n_dev, keys_a, passes, begin_bit, pl.digit_w, end_bit,
                                                                  hist);
    FSB_LAUNCH_CHECK();
    uint64_t* kin = keys_a; int32_t* vin = vals_a;
    uint64_t* kout = keys_b; int32_t* vout = vals_b;
    const size_t smem = (size_t)SORT_TILE * (PAIRS ? 12 : 8) + (size_t)SORT_WARPS * RADIX * 4;
    static bool smem_opted_in = false;  // once per instantiation, outside any stream capture (the first call is eager)
    if (!smem_opted_in) {
        FSB_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<BITS, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        smem_opted_in = true;
    }
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + pl.digit_w * p;
        const int width = pl.digit_w < end_bit - shift ? pl.digit_w : end_bit - shift;
        onesweep_pass_kernel<BITS, PAIRS><<<(unsigned)tiles, SORT_THREADS, smem, st>>>(
            n, n_dev, kin, vin, kout, vout, p == passes - 1 ? low32_out : nullptr, hist + (size_t)p * RADIX,
            status + (size_t)p * tiles * RADIX, counters + p, shift, width);
        FSB_LAUNCH_CHECK();
        uint64_t* tk = kin; kin = kout; kout = tk;
        int32_t* tv = vin; vin = vout; vout = tv;
    }
    return 0;
}

// Sorts pairs by key bits [0, end_bit).  Buffers A (input, clobbered) and B ping-pong;
// *result_in_b tells the caller which one holds the sorted pairs (it depends only on end_bit).
// n_dev (nullable): device pointer to the true count; n is then the capacity (see common.cuh).
FSB_API int fsb_radix_sort_pairs(int64_t n, const int64_t* n_dev, int end_bit, uint64_t* keys_a, int32_t* vals_a,
                                 uint64_t* keys_b, int32_t* vals_b, void* workspace, size_t workspace_bytes,
                                 int* result_in_b, void* stream) {
    if (n < 0 || n >= (1ll << 30) || end_bit < 1 || end_bit > 64) return FSB_E_ARG;
    const SortPlan pl = sort_plan(end_bit);
    if (pl.passes > MAX_PASSES) return FSB_E_ARG;
    if (workspace_bytes < sort_workspace_bytes(n, end_bit)) return FSB_E_ARG;
    if (result_in_b) *result_in_b = pl.passes & 1;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_CUDA(cudaMemsetAsync(workspace, 0, sort_workspace_bytes(n, end_bit), st));
    if (pl.radix_bits == 9)
        return sort_launch<9, true>(n, n_dev, 0, end_bit, pl, keys_a, vals_a, keys_b, vals_b, nullptr, workspace, st);
    return sort_launch<8, true>(n, n_dev, 0, end_bit, pl, keys_a, vals_a, keys_b, vals_b, nullptr, workspace, st);
}

// Stable sort of 64-bit keys on bits [begin_bit, end_bit) only (the other bits travel with the key).  Buffers A (input,
// clobbered) and B ping-pong, *result_in_b as above.  low32_out (nullable, int32[n]): the low word of every key at its
// sorted position, written by the last pass.  n_dev as above.
FSB_API int fsb_radix_sort_keys(int64_t n, const int64_t* n_dev, int begin_bit, int end_bit, uint64_t* keys_a,
                                uint64_t* keys_b, int32_t* low32_out, void* workspace, size_t workspace_bytes,
                                int* result_in_b, void* stream) {
    if (n < 0 || n >= (1ll << 30) || begin_bit < 0 || end_bit <= begin_bit || end_bit > 64) return FSB_E_ARG;
    const int nbits = end_bit - begin_bit;
    const SortPlan pl = sort_plan(nbits);
    if (pl.passes > MAX_PASSES) return FSB_E_ARG;
    if (workspace_bytes < sort_workspace_bytes(n, nbits)) return FSB_E_ARG;
    if (result_in_b) *result_in_b = pl.passes & 1;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    FSB_CUDA(cudaMemsetAsync(workspace, 0, sort_workspace_bytes(n, nbits), st));
    if (pl.radix_bits == 9)
        return sort_launch<9, false>(n, n_dev, begin_bit, end_bit, pl, keys_a, nullptr, keys_b, nullptr, low32_out,
                                     workspace, st);
    return sort_launch<8, false>(n, n_dev, begin_bit, end_bit, pl, keys_a, nullptr, keys_b, nullptr, low32_out,
                                 workspace, st);
}
