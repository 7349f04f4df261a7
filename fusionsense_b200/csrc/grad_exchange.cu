// grad_exchange.cu — the multi-GPU exchange step of the DN-Splatter iteration as kernels of our own over NVLink
// peer memory (no NCCL call on the step's path): pack -> barrier -> reduce-scatter -> barrier -> Adam reading the
// reduced gradient straight out of its owners' memory (the all-gather is folded into the optimizer's gradient load).
//
// Replaces what DDP would do for /root/reference/dn_splatter/dn_pipeline.py:161-167 (one all-reduce of every Gaussian
// parameter gradient per iteration; DDP itself cannot follow densification).  Every rank holds a full
// replica; rank r owns slice r = [r S, (r + 1) S) of the flat 59-floats-per-Gaussian gradient:
//
//   fsb_xchg_pack            the step's gradient tensors -> this rank's flat buffer G_r (symmetric memory: mapped into
//                            every peer), tensor offsets rounded to 4 floats so no float4 straddles a tensor or slice
//   fsb_xchg_barrier (A)     "my G is complete"; carries the static-capacity overflow flag (any rank -> every rank)
//   fsb_xchg_reduce_scatter  R_r[i] = sum over w of G_w[r S + i]: plain peer loads over NVLink, or — when the buffers
//                            have an NVSwitch multicast mapping — one multimem.ld_reduce per 16 bytes (the switch adds)
//   fsb_xchg_barrier (B)     "my R is complete" (also: every rank is done reading G, it may be overwritten)
//   fsb_adam_multi_xchg      csrc/adam.cu with the gradient of flat element F read from R_{F / S} in peer memory
//
// Each element is reduced exactly once, by its owner, in rank order: all replicas apply bit-identical updates.
// A barrier is one tiny kernel: thread w stores this barrier's epoch into peer w's signal pad (release, system scope)
// and spins on its own pad until peer w's epoch arrives (acquire).  Epochs only grow, so pads are never reset and a
// replayed CUDA graph needs no host involvement.
#include "common.cuh"

#define FSB_XCHG_MAX_WORLD 8
#define FSB_XCHG_MAX_TENSORS 8

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct PadPtrs {
    uint32_t* p[FSB_XCHG_MAX_WORLD];
};

// pads[w] -> rank w's pad: uint32 [n_slots][FSB_XCHG_MAX_WORLD]; entry [slot][src] is written by rank src only
__global__ void xchg_barrier_kernel(int world, int rank, PadPtrs pads, int slot, uint32_t* __restrict__ epoch_dev,
                                    int32_t* __restrict__ flag) {
    __shared__ uint32_t s_epoch;
    __shared__ int s_any;
    if (threadIdx.x == 0) {
        s_epoch = epoch_dev[slot] + 1u;
        epoch_dev[slot] = s_epoch;
        s_any = 0;
    }
    __syncthreads();
    const uint32_t e = s_epoch;
    const int w = threadIdx.x;
    if (w < world) {
        const uint32_t mine = (e << 1) | ((flag != nullptr && *flag != 0) ? 1u : 0u);
        __threadfence_system();
        st_release_sys(pads.p[w] + slot * FSB_XCHG_MAX_WORLD + rank, mine);
        const uint32_t* src = pads.p[rank] + slot * FSB_XCHG_MAX_WORLD + w;
        uint32_t got;
        do {
            got = ld_acquire_sys(src);
        } while ((got >> 1) < e);
        if (got & 1u) atomicOr(&s_any, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && flag != nullptr && s_any) *flag = 1;
}

struct PackArgs {
    const float* src[FSB_XCHG_MAX_TENSORS];
    long long n[FSB_XCHG_MAX_TENSORS];
    long long off[FSB_XCHG_MAX_TENSORS + 1];  // flat offsets (multiples of 4); off[n_tensors] = end of the payload
};

// dst[off[t] + j] = src[t][j]; the gaps up to the next offset and up to `total` are zero-filled
__global__ void __launch_bounds__(256)
xchg_pack_kernel(PackArgs a, int n_tensors, float* __restrict__ dst, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += stride) {
        int t = 0;
#pragma unroll
        for (int k = 1; k < FSB_XCHG_MAX_TENSORS; ++k)
            if (k < n_tensors && i >= a.off[k]) t = k;
        const long long j = i - a.off[t];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* s = a.src[t];
        if (i < a.off[n_tensors] && j < a.n[t]) {
            if (j + 4 <= a.n[t] && (((uintptr_t)(s + j)) & 15) == 0) {
                v = *reinterpret_cast<const float4*>(s + j);
            } else {
                v.x = s[j];
                if (j + 1 < a.n[t]) v.y = s[j + 1];
                if (j + 2 < a.n[t]) v.z = s[j + 2];
                if (j + 3 < a.n[t]) v.w = s[j + 3];
            }
        }
        *reinterpret_cast<float4*>(dst + i) = v;
    }
}

struct PeerPtrs {
    const float* p[FSB_XCHG_MAX_WORLD];
};

__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// out[i] = sum_w g.p[w][rank S + i], i in [0, S), S a multiple of 4; rank order, so the sum is the same bit pattern
// whoever computes it
__global__ void __launch_bounds__(256)
xchg_reduce_scatter_kernel(int world, int rank, PeerPtrs g, long long S, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    const long long base = (long long)rank * S;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < S; i += stride) {
        float4 acc = ld_peer(g.p[0] + base + i);
        for (int w = 1; w < world; ++w) {
            const float4 v = ld_peer(g.p[w] + base + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(out + i) = acc;
    }
}

// the same through the NVSwitch multicast mapping of the gradient buffers: the switch fetches the 16 bytes from every
// GPU of the group and adds them (NVLS), one response per request
__global__ void __launch_bounds__(256)
xchg_reduce_scatter_mc_kernel(int rank, const float* __restrict__ g_mc, long long S, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    const long long base = (long long)rank * S;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < S; i += stride) {
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(g_mc + base + i)
                     : "memory");
        *reinterpret_cast<float4*>(out + i) = v;
    }
}

// ---- in-place all-reduce: the owner reduces its slice and writes the sum back into every replica ------------------
// r02m, 8 GPUs: with the all-gather folded into Adam's gradient load (above) the exchange took 1.6 ms against NCCL's
// 0.68 ms — an HBM-bound kernel whose every seventh load crosses NVLink and comes back after microseconds waits for
// those loads.  Stores are posted, and the switch can replicate them: one multimem.ld_reduce (NVSwitch adds the
// replicas' 16 bytes) and one multimem.st (NVSwitch writes the sum into all of them) per 16 bytes of the owner's
// slice, the one-kernel NVLS all-reduce.  Adam then reads a local buffer.
struct PeerPtrsRW {
    float* p[FSB_XCHG_MAX_WORLD];
};

constexpr int AR_THREADS = 512;
constexpr int AR_UNROLL = 4;

__global__ void __launch_bounds__(AR_THREADS)
xchg_allreduce_mc_kernel(int rank, float* __restrict__ g_mc, long long S) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    float* base = g_mc + (long long)rank * S;
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < S; i0 += stride * AR_UNROLL) {
        float4 v[AR_UNROLL];
#pragma unroll
        for (int u = 0; u < AR_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < S)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "l"(base + i)
                             : "memory");
        }
#pragma unroll
        for (int u = 0; u < AR_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < S)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + i),
                             "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w)
                             : "memory");
        }
    }
}

// no multicast mapping: peer loads in rank order, then one store per replica
__global__ void __launch_bounds__(AR_THREADS)
xchg_allreduce_peer_kernel(int world, int rank, PeerPtrsRW g, long long S) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    const long long base = (long long)rank * S;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < S; i += stride) {
        float4 acc = ld_peer(g.p[0] + base + i);
        for (int w = 1; w < world; ++w) {
            const float4 v = ld_peer(g.p[w] + base + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        for (int w = 0; w < world; ++w) *reinterpret_cast<float4*>(g.p[w] + base + i) = acc;
    }
}

// ---- Adam with the gradient gathered from the owners' reduced slices ---------------------------------------------
struct AdamXArgs {
    float* p[FSB_XCHG_MAX_TENSORS];
    float* m[FSB_XCHG_MAX_TENSORS];
    float* v[FSB_XCHG_MAX_TENSORS];
    long long n[FSB_XCHG_MAX_TENSORS];
    long long off[FSB_XCHG_MAX_TENSORS];  // flat offset of the tensor (multiple of 4)
    int block_start[FSB_XCHG_MAX_TENSORS + 1];
    const float* r[FSB_XCHG_MAX_WORLD];   // reduced slices, one per rank (peer memory)
    long long S;
};

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC = 4;
constexpr int ADAM_PER_BLOCK = ADAM_THREADS * ADAM_VEC * 4;

// torch.optim.Adam's single-tensor arithmetic order (see csrc/adam.cu)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float omb1, float b2, float omb2,
                                         float eps, float step_size, float bc2_sqrt) {
    m = m + (g - m) * omb1;
    v = v * b2 + omb2 * g * g;
    float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_multi_xchg_kernel(AdamXArgs a, int n_tensors, float omb1, float b2, float omb2, float eps,
                       const float* __restrict__ hyper_dev, int hyper_stride, const int32_t* __restrict__ skip_flag) {
    if (skip_flag != nullptr && *skip_flag != 0) return;
    int t = 0;
#pragma unroll
    for (int i = 1; i < FSB_XCHG_MAX_TENSORS; ++i)
        if (i < n_tensors && (int)blockIdx.x >= a.block_start[i]) t = i;
    const long long n = a.n[t];
    const long long base = (long long)(blockIdx.x - a.block_start[t]) * ADAM_PER_BLOCK;
    float* __restrict__ p = a.p[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    const float ss = hyper_dev[t];
    const float bc2s = hyper_dev[hyper_stride + t];
    const bool aligned = ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const long long i = base + ((long long)it * ADAM_THREADS + threadIdx.x) * ADAM_VEC;
        if (i >= n) break;
        const long long F = a.off[t] + i;  // multiple of 4: the float4 lies inside one slice
        const int w = (int)(F / a.S);
        const float* gp = a.r[w] + (F - (long long)w * a.S);
        const float4 G = ld_peer(gp);  // the padding behind a tensor's end is zero-filled by the pack
        if (aligned && i + ADAM_VEC <= n) {
            float4 P = *reinterpret_cast<float4*>(p + i);
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
            const float g4[4] = {G.x, G.y, G.z, G.w};
            for (int k = 0; k < ADAM_VEC && i + k < n; ++k) {
                float P = p[i + k], M = m[i + k], V = v[i + k];
                adam_one(P, g4[k], M, V, omb1, b2, omb2, eps, ss, bc2s);
                p[i + k] = P; m[i + k] = M; v[i + k] = V;
            }
        }
    }
}

}  // namespace

FSB_API int fsb_xchg_max_world(void) { return FSB_XCHG_MAX_WORLD; }

// Cross-GPU barrier number `slot` (0 .. n_slots - 1) of this step.  pads: HOST array of `world` device pointers, pads[w]
// = rank w's signal pad (uint32 [n_slots][fsb_xchg_max_world()], zero-initialised once, mapped into this process);
// epoch_dev: this rank's uint32 [n_slots] counters (zero-initialised once).  flag (nullable, int32[1]): on return it is
// non-zero on every rank if it was non-zero on any rank when the barrier was entered.
FSB_API int fsb_xchg_barrier(int world, int rank, uint32_t* const* pads, int slot, uint32_t* epoch_dev, int32_t* flag,
                             void* stream) {
    if (world < 1 || world > FSB_XCHG_MAX_WORLD || rank < 0 || rank >= world || !pads || slot < 0 || !epoch_dev)
        return FSB_E_ARG;
    PadPtrs pp;
    for (int w = 0; w < FSB_XCHG_MAX_WORLD; ++w) pp.p[w] = w < world ? pads[w] : nullptr;
    xchg_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(world, rank, pp, slot, epoch_dev, flag);
    FSB_LAUNCH_CHECK();
    return 0;
}

// Flat gradient buffer from the step's gradient tensors.  src / n / off: HOST arrays of n_tensors (+ 1 for off) entries;
// off[t] = flat offset of tensor t (multiple of 4, ascending, off[n_tensors] = end); dst[0 .. total) is written
// (total a multiple of 4: payload, gaps and tail zero-filled).
FSB_API int fsb_xchg_pack(int n_tensors, const float* const* src, const int64_t* n, const int64_t* off, float* dst,
                          int64_t total, void* stream) {
    if (n_tensors < 1 || n_tensors > FSB_XCHG_MAX_TENSORS || !dst || total < 0 || (total & 3)) return FSB_E_ARG;
    PackArgs a;
    for (int t = 0; t < FSB_XCHG_MAX_TENSORS; ++t) {
        const bool live = t < n_tensors;
        a.src[t] = live ? src[t] : nullptr;
        a.n[t] = live ? n[t] : 0;
        a.off[t] = live ? off[t] : off[n_tensors];
        if (live && ((off[t] & 3) || n[t] < 0 || off[t] + n[t] > off[t + 1])) return FSB_E_ARG;
    }
    a.off[FSB_XCHG_MAX_TENSORS] = off[n_tensors];
    for (int t = n_tensors; t <= FSB_XCHG_MAX_TENSORS; ++t) a.off[t] = off[n_tensors];
    if (off[n_tensors] > total) return FSB_E_ARG;
    if (total == 0) return 0;
    int blocks = fsb_div_up(total / 4, 256 * 4);
    if (blocks > FSB_NUM_SMS * 8) blocks = FSB_NUM_SMS * 8;
    xchg_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, n_tensors, dst, total);
    FSB_LAUNCH_CHECK();
    return 0;
}

// out[0 .. S) = sum over ranks w of grads[w][rank * S ...], S a multiple of 4.  grads: HOST array of `world` device
// pointers to the ranks' flat gradient buffers (peer mappings).  grads_mc (nullable): NVSwitch multicast mapping of
// the same buffers; when given, the reduction is done by the switch (multimem.ld_reduce).
FSB_API int fsb_xchg_reduce_scatter(int world, int rank, const float* const* grads, const float* grads_mc, int64_t S,
                                    float* out, void* stream) {
    if (world < 1 || world > FSB_XCHG_MAX_WORLD || rank < 0 || rank >= world || !grads || !out || S < 0 || (S & 3))
        return FSB_E_ARG;
    if (S == 0) return 0;
    int blocks = fsb_div_up(S / 4, 256 * 2);
    if (blocks > FSB_NUM_SMS * 8) blocks = FSB_NUM_SMS * 8;
    if (grads_mc != nullptr) {
        xchg_reduce_scatter_mc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rank, grads_mc, S, out);
    } else {
        PeerPtrs g;
        for (int w = 0; w < FSB_XCHG_MAX_WORLD; ++w) g.p[w] = w < world ? grads[w] : nullptr;
        xchg_reduce_scatter_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(world, rank, g, S, out);
    }
    FSB_LAUNCH_CHECK();
    return 0;
}

// In place: afterwards EVERY rank's buffer holds sum over ranks of the buffers, in elements [rank S, (rank + 1) S) as far
// as this call is concerned — all ranks call it between two fsb_xchg_barrier's, each for its own slice.
// grads: HOST array of `world` device pointers (peer mappings, written to); grads_mc (nullable): the multicast mapping.
FSB_API int fsb_xchg_allreduce(int world, int rank, float* const* grads, float* grads_mc, int64_t S, void* stream) {
    if (world < 1 || world > FSB_XCHG_MAX_WORLD || rank < 0 || rank >= world || !grads || S < 0 || (S & 3))
        return FSB_E_ARG;
    if (S == 0) return 0;
    // two 512-thread CTAs per SM keep ~10 MB on the wire (the links need about 3) and leave half of every SM to the
    // Adam launch of the previous chunk that runs beside this kernel (dist.PeerGradExchange.exchange_and_adam)
    static const int per_sm = [] { const char* e = getenv("FSB_XCHG_AR_CTAS_PER_SM"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > 4 ? 4 : v); }();
    int blocks = fsb_div_up(S / 4, AR_THREADS * (grads_mc != nullptr ? AR_UNROLL : 1));
    if (blocks > FSB_NUM_SMS * per_sm) blocks = FSB_NUM_SMS * per_sm;
    if (grads_mc != nullptr) {
        xchg_allreduce_mc_kernel<<<blocks, AR_THREADS, 0, (cudaStream_t)stream>>>(rank, grads_mc, S);
    } else {
        PeerPtrsRW g;
        for (int w = 0; w < FSB_XCHG_MAX_WORLD; ++w) g.p[w] = w < world ? grads[w] : nullptr;
        xchg_allreduce_peer_kernel<<<blocks, AR_THREADS, 0, (cudaStream_t)stream>>>(world, rank, g, S);
    }
    FSB_LAUNCH_CHECK();
    return 0;
}

// fsb_adam_multi_dev (csrc/adam.cu) with the gradient of tensor t's element j read from the reduced slices:
// flat index F = off[t] + j lives in reduced[F / S][F % S].  reduced: HOST array of `world` device pointers (peer
// mappings of every rank's fsb_xchg_reduce_scatter output).  hyper_dev as in fsb_adam_multi_dev.
FSB_API int fsb_adam_multi_xchg(int n_tensors, float* const* p, float* const* m, float* const* v, const int64_t* n,
                                const int64_t* off, int world, const float* const* reduced, int64_t S,
                                const float* hyper_dev, int hyper_stride, const int32_t* skip_flag, double beta1,
                                double beta2, double eps, void* stream) {
    if (n_tensors <= 0 || n_tensors > FSB_XCHG_MAX_TENSORS || !hyper_dev || world < 1 || world > FSB_XCHG_MAX_WORLD ||
        !reduced || S <= 0 || (S & 3))
        return FSB_E_ARG;
    AdamXArgs a;
    int blocks = 0;
    for (int i = 0; i < FSB_XCHG_MAX_TENSORS; ++i) {
        const bool live = i < n_tensors;
        if (live && (n[i] < 0 || (off[i] & 3) || off[i] + n[i] > S * world)) return FSB_E_ARG;
        a.p[i] = live ? p[i] : nullptr;
        a.m[i] = live ? m[i] : nullptr;
        a.v[i] = live ? v[i] : nullptr;
        a.n[i] = live ? n[i] : 0;
        a.off[i] = live ? off[i] : 0;
        a.block_start[i] = blocks;
        if (live) blocks += (int)((n[i] + ADAM_PER_BLOCK - 1) / ADAM_PER_BLOCK);
    }
    a.block_start[FSB_XCHG_MAX_TENSORS] = blocks;
    for (int w = 0; w < FSB_XCHG_MAX_WORLD; ++w) a.r[w] = w < world ? reduced[w] : nullptr;
    a.S = S;
    if (blocks == 0) return 0;
    adam_multi_xchg_kernel<<<blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(
        a, n_tensors, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, hyper_dev, hyper_stride,
        skip_flag);
    FSB_LAUNCH_CHECK();
    return 0;
}
