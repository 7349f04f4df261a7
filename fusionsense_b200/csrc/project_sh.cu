// project_sh.cu — fused EWA projection + SH colour + tile count (forward) and its backward.
//
// Replaces gsplat 1.0.0's fully_fused_projection_{fwd,bwd}, compute_sh_{fwd,bwd} and the first
// pass of isect_tiles (SURVEY.md §2b stages P1, P2, I1-count, P3), reached from
// /root/reference/dn_splatter/dn_model.py:570-591.
//
// HBM-bound streaming kernels: one thread per (camera, Gaussian) forward, one thread per Gaussian
// looping over cameras backward (so the cross-camera sum is a register accumulation, no atomics).
#include "common.cuh"
#include <stdlib.h>
#include "fs_math.cuh"

namespace {

__device__ __forceinline__ fs::Camera load_camera(const float* __restrict__ viewmats, const float* __restrict__ Ks,
                                                  int c) {
    fs::Camera cam;
    const float* V = viewmats + (size_t)c * 16;
#pragma unroll
    for (int i = 0; i < 12; ++i) cam.V[i] = __ldg(V + i);
    const float* K = Ks + (size_t)c * 9;
    cam.fx = __ldg(K + 0);
    cam.cx = __ldg(K + 2);
    cam.fy = __ldg(K + 4);
    cam.cy = __ldg(K + 5);
    return cam;
}

// camera centre = -R^-1 t of the (affine) view matrix; what rendering.rasterization takes from
// torch.linalg.inv(viewmats)[:, :3, 3].  Used when the caller passes campos = NULL (13 tiny LU launches saved).
__device__ __forceinline__ void camera_centre(const fs::Camera& cam, const float* __restrict__ campos, int c,
                                              float* o) {
    if (campos) {
        o[0] = __ldg(campos + 3 * c + 0); o[1] = __ldg(campos + 3 * c + 1); o[2] = __ldg(campos + 3 * c + 2);
        return;
    }
    const float* V = cam.V;  // row-major 3x4
    const float r00 = V[0], r01 = V[1], r02 = V[2], t0 = V[3];
    const float r10 = V[4], r11 = V[5], r12 = V[6], t1 = V[7];
    const float r20 = V[8], r21 = V[9], r22 = V[10], t2 = V[11];
    const float c00 = r11 * r22 - r12 * r21, c01 = r12 * r20 - r10 * r22, c02 = r10 * r21 - r11 * r20;
    const float det = r00 * c00 + r01 * c01 + r02 * c02;
    const float id = 1.f / det;
    // inverse = adjugate / det; row i of the inverse dotted with t
    const float i00 = c00, i01 = r02 * r21 - r01 * r22, i02 = r01 * r12 - r02 * r11;
    const float i10 = c01, i11 = r00 * r22 - r02 * r20, i12 = r02 * r10 - r00 * r12;
    const float i20 = c02, i21 = r01 * r20 - r00 * r21, i22 = r00 * r11 - r01 * r10;
    o[0] = -(i00 * t0 + i01 * t1 + i02 * t2) * id;
    o[1] = -(i10 * t0 + i11 * t1 + i12 * t2) * id;
    o[2] = -(i20 * t0 + i21 * t1 + i22 * t2) * id;
}

__device__ __forceinline__ int tile_count(float mx, float my, int radius, int tile_size, int tile_w, int tile_h,
                                          int legacy_bbox, int* x0, int* y0, int* x1, int* y1) {
    float ts = (float)tile_size;
    float tr = (float)radius / ts;
    float tx = mx / ts, ty = my / ts;
    int ax, ay, bx, by;
    if (legacy_bbox) {
        // gsplat 0.1.x map_gaussian_to_intersects: (int) truncation, +1 on the max side
        ax = (int)(tx - tr); ay = (int)(ty - tr);
        bx = (int)(tx + tr + 1.f); by = (int)(ty + tr + 1.f);
    } else {
        ax = (int)floorf(tx - tr); ay = (int)floorf(ty - tr);
        bx = (int)ceilf(tx + tr); by = (int)ceilf(ty + tr);
    }
    ax = min(max(0, ax), tile_w); ay = min(max(0, ay), tile_h);
    bx = min(max(0, bx), tile_w); by = min(max(0, by), tile_h);
    *x0 = ax; *y0 = ay; *x1 = bx; *y1 = by;
    return (by - ay) * (bx - ax);
}

// SH coefficient j (0 .. 3K-1) of Gaussian n.  One tensor coeffs[N,K,3], or — when `rest` is given — the layout the
// model stores (dn_model.py:294-304): coeffs = features_dc[N,3] (band 0), rest = features_rest[N,K-1,3]; reading the
// two tensors in place saves the torch.cat that builds [N,K,3] every step (and the split of its gradient).
__device__ __forceinline__ size_t sh_index(bool split, int n, int L, int j, bool* in_rest) {
    if (!split) { *in_rest = false; return (size_t)n * L + j; }
    if (j < 3) { *in_rest = false; return (size_t)n * 3 + j; }
    *in_rest = true;
    return (size_t)n * (L - 3) + (j - 3);
}
__device__ __forceinline__ float sh_load(const float* __restrict__ coeffs, const float* __restrict__ rest, int n,
                                         int L, int j) {
    bool r;
    const size_t i = sh_index(rest != nullptr, n, L, j, &r);
    return r ? rest[i] : coeffs[i];
}

// act_flags: the parameters arrive as the model stores them and the activation is applied here
constexpr int ACT_VEC4_ROWS = 2;   // backward: the warp's run of features_rest rows moves as 16-byte requests
constexpr int ACT_FLAT_ROWS = 4;   // backward: that run keeps its global layout in shared memory (no row/column arithmetic)
constexpr int ACT_EXP_SCALES = 1;  // scales are log-scales: s = exp(raw)  (dn_model.py:573)

__global__ void __launch_bounds__(256)
project_sh_fwd_kernel(int C, int N, const float* __restrict__ means, const float* __restrict__ quats,
                      const float* __restrict__ scales, const float* __restrict__ viewmats,
                      const float* __restrict__ Ks, int width, int height, float eps2d, float near_plane,
                      float far_plane, float radius_clip, int tile_size, int tile_w, int tile_h, int sh_degree,
                      int K, const float* __restrict__ coeffs, const float* __restrict__ coeffs_rest, int act_flags,
                      const float* __restrict__ campos, int color_stride,
                      int depth_channel, int32_t* __restrict__ radii, float* __restrict__ means2d,
                      float* __restrict__ depths, float* __restrict__ conics, float* __restrict__ comps,
                      float* __restrict__ colors, int32_t* __restrict__ tiles_per_gauss,
                      unsigned long long* __restrict__ legacy_extra) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)C * N) return;
    int c = (int)(idx / N);
    int n = (int)(idx - (int64_t)c * N);
    fs::Camera cam = load_camera(viewmats, Ks, c);
    float px = means[3 * (size_t)n + 0], py = means[3 * (size_t)n + 1], pz = means[3 * (size_t)n + 2];
    float4 q = reinterpret_cast<const float4*>(quats)[n];
    float sx = scales[3 * (size_t)n + 0], sy = scales[3 * (size_t)n + 1], sz = scales[3 * (size_t)n + 2];
    if (act_flags & ACT_EXP_SCALES) { sx = expf(sx); sy = expf(sy); sz = expf(sz); }
    fs::ProjFwd o = fs::project_fwd(cam, px, py, pz, q.x, q.y, q.z, q.w, sx, sy, sz, width, height, eps2d,
                                    near_plane, far_plane, radius_clip);
    radii[idx] = o.radius;
    reinterpret_cast<float2*>(means2d)[idx] = make_float2(o.mx, o.my);
    depths[idx] = o.depth;
    conics[3 * idx + 0] = o.ca;
    conics[3 * idx + 1] = o.cb;
    conics[3 * idx + 2] = o.cc;
    if (comps) comps[idx] = o.comp;
    int cnt = 0;
    if (o.radius > 0) {
        int x0, y0, x1, y1;
        cnt = tile_count(o.mx, o.my, o.radius, tile_size, tile_w, tile_h, 0, &x0, &y0, &x1, &y1);
        if (legacy_extra) {
            // tiles the gsplat 0.1.x bbox rule would add (it differs only when (x + r) / tile_size rounds to an
            // integer: ~1 Gaussian per frame); zero means the legacy normals pass can share these sorted lists
            const int lc = tile_count(o.mx, o.my, o.radius, tile_size, tile_w, tile_h, 1, &x0, &y0, &x1, &y1);
            if (lc != cnt) atomicAdd(legacy_extra, (unsigned long long)(lc > cnt ? lc - cnt : cnt - lc));
        }
    }
    tiles_per_gauss[idx] = cnt;
    if (colors) {
        float* out = colors + (size_t)idx * color_stride;
        float r = 0.f, g = 0.f, b = 0.f;
        if (o.radius > 0 && sh_degree >= 0) {
            float cc[3];
            camera_centre(cam, campos, c, cc);
            float dx = px - cc[0], dy = py - cc[1], dz = pz - cc[2];
            float inorm = fs::inv_sqrt(dx * dx + dy * dy + dz * dz);
            float basis[16];
            fs::sh_basis(sh_degree, dx * inorm, dy * inorm, dz * inorm, basis);
            int nb = (sh_degree + 1) * (sh_degree + 1);
            if (coeffs_rest) {
                const float* dc = coeffs + (size_t)n * 3;
                const float* cf = coeffs_rest + (size_t)n * (K - 1) * 3 - 3;  // band k >= 1 at cf[3 k + c]
                r = basis[0] * dc[0]; g = basis[0] * dc[1]; b = basis[0] * dc[2];
#pragma unroll 4
                for (int k = 1; k < nb; ++k) {
                    r += basis[k] * cf[3 * k + 0];
                    g += basis[k] * cf[3 * k + 1];
                    b += basis[k] * cf[3 * k + 2];
                }
            } else {
                const float* cf = coeffs + (size_t)n * K * 3;
#pragma unroll 4
                for (int k = 0; k < nb; ++k) {
                    r += basis[k] * cf[3 * k + 0];
                    g += basis[k] * cf[3 * k + 1];
                    b += basis[k] * cf[3 * k + 2];
                }
            }
            // host-side clamp_min(colors + 0.5, 0) of rendering.rasterization, fused here
            r = fmaxf(r + 0.5f, 0.f);
            g = fmaxf(g + 0.5f, 0.f);
            b = fmaxf(b + 0.5f, 0.f);
        } else if (sh_degree >= 0) {
            // masked-out entries are 0 from spherical_harmonics, then +0.5 by the host clamp
            r = g = b = 0.5f;
        }
        if (sh_degree >= 0) {
            out[0] = r; out[1] = g; out[2] = b;
        }
        if (depth_channel >= 0) out[depth_channel] = o.depth;
    }
}

constexpr int PB_THREADS = 128;           // backward block size
constexpr int PB_WARPS = PB_THREADS / 32;
constexpr int SH_ROW_MAX = 48;            // K * 3 floats of SH coefficients per Gaussian staged through smem (K <= 16)

// block-wide sum of `v` (blockDim.x == PB_THREADS), result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) smem[w] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < PB_WARPS; ++i) r += smem[i];
    }
    return r;
}

// One thread per Gaussian, looping over cameras.  The SH rows (192 B per Gaussian at K = 16) are the bulk of the
// traffic; a thread-per-row access pattern touches 32 sectors per request, so each warp moves its 32 rows
// through a shared-memory tile with row-contiguous (coalesced) global loads and stores.
// MINB = minimum resident CTAs per SM the register allocation targets: 4 -> 115 registers (round 1), 6 -> 80 registers
// with 48 bytes of spills, 24 instead of 16 warps per SM for a kernel that is latency / HBM bound (20 % of HBM in r01).
template <int MINB>
__global__ void __launch_bounds__(PB_THREADS, MINB)
project_sh_bwd_kernel(int C, int N, const float* __restrict__ means, const float* __restrict__ quats,
                      const float* __restrict__ scales, const float* __restrict__ viewmats,
                      const float* __restrict__ Ks, int width, int height, float eps2d, int sh_degree, int K,
                      const float* __restrict__ coeffs, const float* __restrict__ coeffs_rest, int act_flags,
                      const float* __restrict__ campos, int color_stride,
                      int depth_channel, const int32_t* __restrict__ radii, const float* __restrict__ v_means2d,
                      const float* __restrict__ v_depths, const float* __restrict__ v_conics,
                      const float* __restrict__ v_comps, const float* __restrict__ v_colors,
                      float* __restrict__ v_means, float* __restrict__ v_quats, float* __restrict__ v_scales,
                      float* __restrict__ v_coeffs, float* __restrict__ v_coeffs_rest, float* __restrict__ v_viewmats,
                      float* __restrict__ v_campos) {
    __shared__ float red[PB_WARPS];
    __shared__ __align__(16) float tiles[PB_WARPS][32][SH_ROW_MAX + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*tile)[SH_ROW_MAX + 1] = tiles[warp];
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = n - lane;  // first Gaussian of this warp
    bool live = n < N;
    const int L = K * 3;
    const bool do_sh = (sh_degree >= 0 && v_colors != nullptr && v_coeffs != nullptr);
    const bool staged = do_sh && L <= SH_ROW_MAX;  // rows go through the tile
    float px = 0, py = 0, pz = 0, sx = 1, sy = 1, sz = 1;
    float4 q = make_float4(1, 0, 0, 0);
    if (live) {
        px = means[3 * (size_t)n + 0]; py = means[3 * (size_t)n + 1]; pz = means[3 * (size_t)n + 2];
        q = reinterpret_cast<const float4*>(quats)[n];
        sx = scales[3 * (size_t)n + 0]; sy = scales[3 * (size_t)n + 1]; sz = scales[3 * (size_t)n + 2];
        if (act_flags & ACT_EXP_SCALES) { sx = expf(sx); sy = expf(sy); sz = expf(sz); }
    }
    const bool split = (coeffs_rest != nullptr);  // the host entry point guarantees `staged` in this case
    // Flat rows (split layout): the warp's 32 rows of features_rest are ONE contiguous run of 32 (L - 3) floats in
    // global memory; shared memory keeps exactly that layout, so moving it in and out is a plain 16-byte copy with no
    // (row, column) bookkeeping — that bookkeeping was ~40 % of the kernel's instructions (r02o source page) — and
    // lane l's row starts at float l (L - 3): conflict-free whenever L - 3 is odd (45 at K = 16).  Band 0
    // (features_dc) stays in registers.
    const int LR = L - 3;
    const bool flat = split && staged && (act_flags & ACT_FLAT_ROWS);
    float* const flat_buf = &tiles[warp][0][0];  // 32 * 49 floats >= 32 * LR, 16-byte aligned
    float* const frow = flat_buf + lane * LR;
    float c0[3] = {0, 0, 0}, g0[3] = {0, 0, 0};  // band-0 coefficients / gradient (flat mode)
    float am[3] = {0, 0, 0}, aq[4] = {0, 0, 0, 0}, as[3] = {0, 0, 0};
    int nb = sh_degree >= 0 ? (sh_degree + 1) * (sh_degree + 1) : 0;
    bool any_sh = false;
    // gradient rows accumulate in the tile (staged) across cameras; with one camera the tile first carries the
    // coefficients themselves (coalesced load), with several they are read straight from global memory
    const bool coeffs_in_tile = staged && C == 1;
    if (staged && !coeffs_in_tile) {
        if (flat) { for (int k = 0; k < LR; ++k) frow[k] = 0.f; }
        else { for (int k = 0; k < L; ++k) tile[lane][k] = 0.f; }
    }
    for (int c = 0; c < C; ++c) {
        size_t idx = (size_t)c * N + n;
        bool vis = live && radii[idx] > 0;
        if (coeffs_in_tile) {
            const unsigned vmask = __ballot_sync(0xffffffffu, vis);
            if (flat) {
                if (vis) {
                    c0[0] = coeffs[3 * (size_t)n + 0]; c0[1] = coeffs[3 * (size_t)n + 1]; c0[2] = coeffs[3 * (size_t)n + 2];
                }
                const int total = min(32, N - n0) * LR;
                const float* base = coeffs_rest + (size_t)n0 * LR;
                int e_done = 0;
                if ((((uintptr_t)base) & 15) == 0) {
                    constexpr int UB4 = 4;  // independent 16-byte requests per lane in flight
                    const int total4 = total >> 2;
                    const float4* base4 = reinterpret_cast<const float4*>(base);
                    float4* flat4 = reinterpret_cast<float4*>(flat_buf);
                    const float inv_lr = 1.0f / (float)LR;
                    for (int e0 = lane; e0 < total4; e0 += 32 * UB4) {
                        float4 v[UB4];
#pragma unroll
                        for (int u = 0; u < UB4; ++u) {
                            const int e = e0 + 32 * u;
                            // rows the four floats belong to ((a + 0.5) / LR is >= 0.5 / LR away from an integer: the
                            // float quotient floors exactly); rows of culled Gaussians are not fetched
                            const int r_a = (int)(((float)(4 * e) + 0.5f) * inv_lr);
                            const int r_b = min((int)(((float)(4 * e + 3) + 0.5f) * inv_lr), 31);
                            const bool ok = e < total4 && (((vmask >> r_a) | (vmask >> r_b)) & 1u);
                            v[u] = ok ? base4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < UB4; ++u) {
                            const int e = e0 + 32 * u;
                            if (e < total4) flat4[e] = v[u];
                        }
                    }
                    e_done = total4 << 2;
                }
                for (int e = e_done + lane; e < total; e += 32) flat_buf[e] = base[e];
            } else if (split) {
                // The warp's 32 rows of features_rest are one contiguous run of 32 (L - 3) floats: copy it flat,
                // lane-contiguous (every request is a full 128-B line), and in batches of independent loads — a
                // row-at-a-time loop waits for each row's data before it asks for the next (first split version:
                // 108 us against 60 us for the same bytes).  Rows of culled Gaussians are skipped.
                if (vis) {
                    tile[lane][0] = coeffs[3 * (size_t)n + 0];
                    tile[lane][1] = coeffs[3 * (size_t)n + 1];
                    tile[lane][2] = coeffs[3 * (size_t)n + 2];
                }
                const int LR = L - 3;
                const int total = min(32, N - n0) * LR;
                const float* base = coeffs_rest + (size_t)n0 * LR;
                int e_done = 0;  // floats [0, e_done) of the run are in the tile
                if ((act_flags & ACT_VEC4_ROWS) && ((((uintptr_t)base) & 15) == 0)) {
                    // 16-byte requests, four per lane in flight: at 16 warps per SM the 4-byte version keeps 16 KB per
                    // SM on the wire, about a third of what the HBM latency needs (r02l: 268 us for 0.5 GB)
                    constexpr int UB4 = 4;
                    const int total4 = total >> 2;
                    const float4* base4 = reinterpret_cast<const float4*>(base);
                    int r = 0, col = 4 * lane;  // (row, column) of the first float of this lane's current float4
                    while (col >= LR) { col -= LR; ++r; }
                    for (int e0 = lane; e0 < total4; e0 += 32 * UB4) {
                        float4 v[UB4];
                        int rr[UB4], cc[UB4];
#pragma unroll
                        for (int u = 0; u < UB4; ++u) {
                            const int e = e0 + 32 * u;
                            rr[u] = r; cc[u] = col;
                            const int r_last = (col + 3 >= LR) ? r + 1 : r;
                            const bool ok = e < total4 && (((vmask >> r) | (vmask >> min(r_last, 31))) & 1u);
                            v[u] = ok ? base4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
                            if (!ok) rr[u] = -1;
                            col += 128;
                            while (col >= LR) { col -= LR; ++r; }
                        }
#pragma unroll
                        for (int u = 0; u < UB4; ++u) {
                            if (rr[u] < 0) continue;
                            const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                int rk = rr[u], ck = cc[u] + k;
                                if (ck >= LR) { ck -= LR; ++rk; }
                                if (rk < 32) tile[rk][3 + ck] = f[k];
                            }
                        }
                    }
                    e_done = total4 << 2;
                }
                constexpr int UB = 8;
                int r = 0, col = e_done + lane;          // element e = e_done + lane + 32 it  ->  (row r, column col) of the run
                while (col >= LR) { col -= LR; ++r; }
                for (int e0 = e_done + lane; e0 < total; e0 += 32 * UB) {
                    float v[UB];
                    int rr[UB], cc[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int e = e0 + 32 * u;
                        rr[u] = r; cc[u] = col;
                        const bool ok = e < total && ((vmask >> r) & 1u);
                        v[u] = ok ? base[e] : 0.f;
                        if (!ok) rr[u] = -1;
                        col += 32;
                        while (col >= LR) { col -= LR; ++r; }
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u)
                        if (rr[u] >= 0) tile[rr[u]][3 + cc[u]] = v[u];
                }
            } else {
                for (int r = 0; r < 32; ++r) {
                    if (!((vmask >> r) & 1u)) continue;
                    const float* row = coeffs + (size_t)(n0 + r) * L;
                    if (lane < L) tile[r][lane] = row[lane];
                    if (lane + 32 < L) tile[r][lane + 32] = row[lane + 32];
                }
            }
            __syncwarp();
        }
        float vRv[9], vtv[3], vcp[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 9; ++i) vRv[i] = 0.f;
        vtv[0] = vtv[1] = vtv[2] = 0.f;
        if (vis) {
            fs::Camera cam = load_camera(viewmats, Ks, c);
            float2 vm = reinterpret_cast<const float2*>(v_means2d)[idx];
            float vd = v_depths ? v_depths[idx] : 0.f;
            float vca = v_conics[3 * idx + 0], vcb = v_conics[3 * idx + 1], vcc = v_conics[3 * idx + 2];
            float vcomp = v_comps ? v_comps[idx] : 0.f;
            const float* vc = v_colors ? v_colors + idx * color_stride : nullptr;
            if (vc && depth_channel >= 0) vd += vc[depth_channel];
            float gm[3], gq[4], gs[3];
            fs::project_bwd(cam, px, py, pz, q.x, q.y, q.z, q.w, sx, sy, sz, width, height, eps2d, vm.x, vm.y, vd,
                            vca, vcb, vcc, vcomp, gm, gq, gs, v_viewmats ? vRv : nullptr, v_viewmats ? vtv : nullptr);
            if (vc && sh_degree >= 0) {
                float cc[3];
                camera_centre(cam, campos, c, cc);
                float dx = px - cc[0], dy = py - cc[1], dz = pz - cc[2];
                float inorm = fs::inv_sqrt(dx * dx + dy * dy + dz * dz);
                float ux = dx * inorm, uy = dy * inorm, uz = dz * inorm;
                float basis[16];
                fs::sh_basis(sh_degree, ux, uy, uz, basis);
                // coefficient (band k >= 1, channel ch) = crow[3 k + ch]; band 0 = cb[ch]
                const float* crow;
                float cb[3];
                if (coeffs_in_tile) {
                    if (flat) { crow = frow - 3; cb[0] = c0[0]; cb[1] = c0[1]; cb[2] = c0[2]; }
                    else { crow = &tile[lane][0]; cb[0] = crow[0]; cb[1] = crow[1]; cb[2] = crow[2]; }
                } else if (split) {
                    crow = coeffs_rest + (size_t)n * LR - 3;
                    cb[0] = coeffs[3 * (size_t)n + 0]; cb[1] = coeffs[3 * (size_t)n + 1]; cb[2] = coeffs[3 * (size_t)n + 2];
                } else {
                    crow = coeffs + (size_t)n * L;
                    cb[0] = crow[0]; cb[1] = crow[1]; cb[2] = crow[2];
                }
                auto cf = [&](int j) -> float { return crow[j]; };  // j >= 3
                float r = basis[0] * cb[0], g = basis[0] * cb[1], b = basis[0] * cb[2];
                for (int k = 1; k < nb; ++k) {
                    r += basis[k] * cf(3 * k + 0);
                    g += basis[k] * cf(3 * k + 1);
                    b += basis[k] * cf(3 * k + 2);
                }
                // clamp_min(x + 0.5, 0): gradient passes where x + 0.5 >= 0
                float vr = (r + 0.5f >= 0.f) ? vc[0] : 0.f;
                float vg = (g + 0.5f >= 0.f) ? vc[1] : 0.f;
                float vb = (b + 0.5f >= 0.f) ? vc[2] : 0.f;
                float gx = 0.f, gy = 0.f, gz = 0.f;
                if (sh_degree >= 1) {
                    float bx[16], by[16], bz[16];
                    fs::sh_basis_grad(sh_degree, ux, uy, uz, bx, by, bz);
                    for (int k = 1; k < nb; ++k) {
                        float w = cf(3 * k + 0) * vr + cf(3 * k + 1) * vg + cf(3 * k + 2) * vb;
                        gx += bx[k] * w; gy += by[k] * w; gz += bz[k] * w;
                    }
                }
                // the coefficients of this row are not needed any more: the tile row now takes the gradient
                if (do_sh) {
                    if (staged) {
                        float* vrow = flat ? frow - 3 : &tile[lane][0];  // band k >= 1 at vrow[3 k + ch]
                        if (coeffs_in_tile) {
                            if (flat) { g0[0] = basis[0] * vr; g0[1] = basis[0] * vg; g0[2] = basis[0] * vb; }
                            else { vrow[0] = basis[0] * vr; vrow[1] = basis[0] * vg; vrow[2] = basis[0] * vb; }
                            for (int k = 1; k < nb; ++k) {
                                vrow[3 * k + 0] = basis[k] * vr;
                                vrow[3 * k + 1] = basis[k] * vg;
                                vrow[3 * k + 2] = basis[k] * vb;
                            }
                            for (int k = nb * 3; k < L; ++k) vrow[k] = 0.f;
                        } else {
                            if (flat) { g0[0] += basis[0] * vr; g0[1] += basis[0] * vg; g0[2] += basis[0] * vb; }
                            else { vrow[0] += basis[0] * vr; vrow[1] += basis[0] * vg; vrow[2] += basis[0] * vb; }
                            for (int k = 1; k < nb; ++k) {
                                vrow[3 * k + 0] += basis[k] * vr;
                                vrow[3 * k + 1] += basis[k] * vg;
                                vrow[3 * k + 2] += basis[k] * vb;
                            }
                        }
                    } else {
                        float* vcf = v_coeffs + (size_t)n * L;
                        for (int k = 0; k < nb; ++k) {
                            float o0 = any_sh ? vcf[3 * k + 0] : 0.f, o1 = any_sh ? vcf[3 * k + 1] : 0.f,
                                  o2 = any_sh ? vcf[3 * k + 2] : 0.f;
                            vcf[3 * k + 0] = o0 + basis[k] * vr;
                            vcf[3 * k + 1] = o1 + basis[k] * vg;
                            vcf[3 * k + 2] = o2 + basis[k] * vb;
                        }
                    }
                }
                any_sh = true;
                if (sh_degree >= 1) {
                    // through u = d / |d|
                    float dot = gx * ux + gy * uy + gz * uz;
                    float vdx = (gx - dot * ux) * inorm, vdy = (gy - dot * uy) * inorm, vdz = (gz - dot * uz) * inorm;
                    gm[0] += vdx; gm[1] += vdy; gm[2] += vdz;
                    vcp[0] = -vdx; vcp[1] = -vdy; vcp[2] = -vdz;
                }
            }
            am[0] += gm[0]; am[1] += gm[1]; am[2] += gm[2];
            aq[0] += gq[0]; aq[1] += gq[1]; aq[2] += gq[2]; aq[3] += gq[3];
            as[0] += gs[0]; as[1] += gs[1]; as[2] += gs[2];
        }
        if (v_viewmats) {  // uniform branch: every thread of the block takes part in the reductions
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                float s = block_sum(vRv[i], red);
                if (threadIdx.x == 0 && s != 0.f) atomicAdd(v_viewmats + (size_t)c * 16 + (i / 3) * 4 + (i % 3), s);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float s = block_sum(vtv[i], red);
                if (threadIdx.x == 0 && s != 0.f) atomicAdd(v_viewmats + (size_t)c * 16 + i * 4 + 3, s);
            }
        }
        if (v_campos) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float s = block_sum(vcp[i], red);
                if (threadIdx.x == 0 && s != 0.f) atomicAdd(v_campos + (size_t)c * 3 + i, s);
            }
        }
    }
    if (v_coeffs) {
        if (staged) {
            // rows of Gaussians that were never visible (or hold stale coefficients) are zero
            if (!any_sh) {
                if (flat) { for (int k = 0; k < LR; ++k) frow[k] = 0.f; }
                else { for (int k = 0; k < L; ++k) tile[lane][k] = 0.f; }
            }
            __syncwarp();
            const int rows = min(32, N - n0);
            if (flat) {
                if (live) {
                    v_coeffs[3 * (size_t)n + 0] = any_sh ? g0[0] : 0.f;
                    v_coeffs[3 * (size_t)n + 1] = any_sh ? g0[1] : 0.f;
                    v_coeffs[3 * (size_t)n + 2] = any_sh ? g0[2] : 0.f;
                }
                const int total = rows * LR;
                float* base = v_coeffs_rest + (size_t)n0 * LR;
                int e_done = 0;
                if ((((uintptr_t)base) & 15) == 0) {
                    const int total4 = total >> 2;
                    float4* base4 = reinterpret_cast<float4*>(base);
                    const float4* flat4 = reinterpret_cast<const float4*>(flat_buf);
#pragma unroll 4
                    for (int e = lane; e < total4; e += 32) base4[e] = flat4[e];
                    e_done = total4 << 2;
                }
                for (int e = e_done + lane; e < total; e += 32) base[e] = flat_buf[e];
            } else if (split) {
                if (live) {
                    v_coeffs[3 * (size_t)n + 0] = tile[lane][0];
                    v_coeffs[3 * (size_t)n + 1] = tile[lane][1];
                    v_coeffs[3 * (size_t)n + 2] = tile[lane][2];
                }
                const int LR = L - 3;
                const int total = rows * LR;
                float* base = v_coeffs_rest + (size_t)n0 * LR;
                int e_done = 0;
                if ((act_flags & ACT_VEC4_ROWS) && ((((uintptr_t)base) & 15) == 0)) {
                    const int total4 = total >> 2;
                    float4* base4 = reinterpret_cast<float4*>(base);
                    int r = 0, col = 4 * lane;
                    while (col >= LR) { col -= LR; ++r; }
#pragma unroll 2
                    for (int e = lane; e < total4; e += 32) {
                        float f[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            int rk = r, ck = col + k;
                            if (ck >= LR) { ck -= LR; ++rk; }
                            f[k] = tile[min(rk, 31)][3 + ck];
                        }
                        base4[e] = make_float4(f[0], f[1], f[2], f[3]);
                        col += 128;
                        while (col >= LR) { col -= LR; ++r; }
                    }
                    e_done = total4 << 2;
                }
                int r = 0, col = e_done + lane;
                while (col >= LR) { col -= LR; ++r; }
#pragma unroll 4
                for (int e = e_done + lane; e < total; e += 32) {
                    base[e] = tile[r][3 + col];
                    col += 32;
                    while (col >= LR) { col -= LR; ++r; }
                }
            } else {
                for (int r = 0; r < rows; ++r) {
                    float* row = v_coeffs + (size_t)(n0 + r) * L;
                    if (lane < L) row[lane] = tile[r][lane];
                    if (lane + 32 < L) row[lane + 32] = tile[r][lane + 32];
                }
            }
        } else if (live) {
            float* vcf = v_coeffs + (size_t)n * L;
            int start = any_sh ? nb : 0;  // bases above the active degree (or everything, if never visible) get zero
            for (int k = start * 3; k < L; ++k) vcf[k] = 0.f;
        }
    }
    if (!live) return;
    v_means[3 * (size_t)n + 0] = am[0]; v_means[3 * (size_t)n + 1] = am[1]; v_means[3 * (size_t)n + 2] = am[2];
    reinterpret_cast<float4*>(v_quats)[n] = make_float4(aq[0], aq[1], aq[2], aq[3]);
    if (act_flags & ACT_EXP_SCALES) { as[0] *= sx; as[1] *= sy; as[2] *= sz; }  // d exp(raw) / d raw = exp(raw)
    v_scales[3 * (size_t)n + 0] = as[0]; v_scales[3 * (size_t)n + 1] = as[1]; v_scales[3 * (size_t)n + 2] = as[2];
}

// stand-alone tile count for callers that bring their own xys/radii (legacy rasterize_gaussians)
__global__ void __launch_bounds__(256)
isect_count_kernel(int64_t M, const float* __restrict__ means2d, const int32_t* __restrict__ radii, int tile_size,
                   int tile_w, int tile_h, int legacy_bbox, int32_t* __restrict__ tiles_per_gauss) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    int r = radii[idx];
    int cnt = 0;
    if (r > 0) {
        float2 m = reinterpret_cast<const float2*>(means2d)[idx];
        int x0, y0, x1, y1;
        cnt = tile_count(m.x, m.y, r, tile_size, tile_w, tile_h, legacy_bbox, &x0, &y0, &x1, &y1);
    }
    tiles_per_gauss[idx] = cnt;
}

}  // namespace

FSB_API int fsb_project_sh_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                               const float* viewmats, const float* Ks, int width, int height, float eps2d,
                               float near_plane, float far_plane, float radius_clip, int tile_size, int tile_w,
                               int tile_h, int sh_degree, int K, const float* coeffs, const float* campos,
                               int color_stride, int depth_channel, int32_t* radii, float* means2d, float* depths,
                               float* conics, float* comps, float* colors, int32_t* tiles_per_gauss,
                               int64_t* legacy_extra, void* stream) {
    if (C <= 0 || N < 0 || sh_degree > 3 || tile_size <= 0) return FSB_E_ARG;
    if (sh_degree >= 0 && (!coeffs || !colors || (sh_degree + 1) * (sh_degree + 1) > K)) return FSB_E_ARG;
    if (colors && (color_stride < 3 || depth_channel >= color_stride)) return FSB_E_ARG;
    if (N == 0) return 0;
    int64_t total = (int64_t)C * N;
    project_sh_fwd_kernel<<<fsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        C, N, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip, tile_size,
        tile_w, tile_h, sh_degree, K, coeffs, nullptr, 0, campos, color_stride, depth_channel, radii, means2d, depths,
        conics, comps, colors, tiles_per_gauss, (unsigned long long*)legacy_extra);
    FSB_LAUNCH_CHECK();
    return 0;
}

// The same stage fed with the model's parameters as stored: log-scales (exp applied here when exp_scales != 0; the
// quaternion is normalised inside either way) and the SH coefficients as two tensors, features_dc[N,3] and
// features_rest[N,K-1,3].  Saves the activation launches and the [N,K,3] concatenation of dn_model.py:566-574.
FSB_API int fsb_project_params_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                                   int exp_scales, const float* viewmats, const float* Ks, int width, int height,
                                   float eps2d, float near_plane, float far_plane, float radius_clip, int tile_size,
                                   int tile_w, int tile_h, int sh_degree, int K, const float* features_dc,
                                   const float* features_rest, int color_stride, int depth_channel, int32_t* radii,
                                   float* means2d, float* depths, float* conics, float* colors,
                                   int32_t* tiles_per_gauss, int64_t* legacy_extra, void* stream) {
    if (C <= 0 || N < 0 || sh_degree < 0 || sh_degree > 3 || tile_size <= 0) return FSB_E_ARG;
    if (!features_dc || !colors || K < 1 || K > 16 || (sh_degree + 1) * (sh_degree + 1) > K) return FSB_E_ARG;
    if (K > 1 && !features_rest) return FSB_E_ARG;
    if (color_stride < 3 || depth_channel >= color_stride) return FSB_E_ARG;
    if (N == 0) return 0;
    int64_t total = (int64_t)C * N;
    project_sh_fwd_kernel<<<fsb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        C, N, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip, tile_size,
        tile_w, tile_h, sh_degree, K, features_dc, K > 1 ? features_rest : nullptr, exp_scales ? ACT_EXP_SCALES : 0,
        nullptr, color_stride, depth_channel, radii, means2d, depths, conics, nullptr, colors, tiles_per_gauss,
        (unsigned long long*)legacy_extra);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_project_sh_bwd(int C, int N, const float* means, const float* quats, const float* scales,
                               const float* viewmats, const float* Ks, int width, int height, float eps2d,
                               int sh_degree, int K, const float* coeffs, const float* campos, int color_stride,
                               int depth_channel, const int32_t* radii, const float* v_means2d, const float* v_depths,
                               const float* v_conics, const float* v_comps, const float* v_colors, float* v_means,
                               float* v_quats, float* v_scales, float* v_coeffs, float* v_viewmats, float* v_campos,
                               void* stream) {
    if (C <= 0 || N < 0 || sh_degree > 3) return FSB_E_ARG;
    if (sh_degree >= 0 && v_colors && (!coeffs || !v_coeffs)) return FSB_E_ARG;
    if (v_campos && !campos) return FSB_E_ARG;
    if (N == 0) return 0;
    project_sh_bwd_kernel<4><<<fsb_div_up(N, PB_THREADS), PB_THREADS, 0, (cudaStream_t)stream>>>(
        C, N, means, quats, scales, viewmats, Ks, width, height, eps2d, sh_degree, K, coeffs, nullptr, 0, campos,
        color_stride, depth_channel, radii, v_means2d, v_depths, v_conics, v_comps, v_colors, v_means, v_quats, v_scales,
        v_coeffs, nullptr, v_viewmats, v_campos);
    FSB_LAUNCH_CHECK();
    return 0;
}

// Backward of fsb_project_params_fwd: gradients w.r.t. the stored parameters (v_scales is d/d log-scale when
// exp_scales != 0; v_quats is w.r.t. the un-normalised quaternion), v_features_dc[N,3] and v_features_rest[N,K-1,3]
// written in place of a [N,K,3] gradient that torch would have to split.  All outputs are overwritten.
FSB_API int fsb_project_params_bwd(int C, int N, const float* means, const float* quats, const float* scales,
                                   int exp_scales, const float* viewmats, const float* Ks, int width, int height,
                                   float eps2d, int sh_degree, int K, const float* features_dc,
                                   const float* features_rest, int color_stride, int depth_channel,
                                   const int32_t* radii, const float* v_means2d, const float* v_depths,
                                   const float* v_conics, const float* v_colors, float* v_means, float* v_quats,
                                   float* v_scales, float* v_features_dc, float* v_features_rest, void* stream) {
    if (C <= 0 || N < 0 || sh_degree < 0 || sh_degree > 3) return FSB_E_ARG;
    if (!features_dc || !v_features_dc || !v_colors || K < 1 || K > 16) return FSB_E_ARG;
    if (K > 1 && (!features_rest || !v_features_rest)) return FSB_E_ARG;
    if (N == 0) return 0;
    // FSB_PROJ_BWD_MINB=6 selects the 80-register build (24 instead of 16 warps per SM); measured SLOWER on B200
    // (r02i, cfg4: 0.304 ms against 0.276 ms: the spills and the longer dependent chains cost more than the extra
    // warps hide), so the 115-register build stays the default
    static const int minb = [] { const char* e = getenv("FSB_PROJ_BWD_MINB"); return e ? atoi(e) : 4; }();
    // FSB_PROJ_BWD_VEC4=0: 4-byte requests for the SH rows (A/B runs)
    static const int vec4 = [] {
        const char* e = getenv("FSB_PROJ_BWD_VEC4");
        const char* f = getenv("FSB_PROJ_BWD_FLAT");  // =0: the row-tile staging of r02l (A/B runs)
        return ((e && e[0] == '0') ? 0 : ACT_VEC4_ROWS) | ((f && f[0] == '0') ? 0 : ACT_FLAT_ROWS);
    }();
#define FSB_PBWD(MINB)                                                                                              \
    project_sh_bwd_kernel<MINB><<<fsb_div_up(N, PB_THREADS), PB_THREADS, 0, (cudaStream_t)stream>>>(                \
        C, N, means, quats, scales, viewmats, Ks, width, height, eps2d, sh_degree, K, features_dc,                  \
        K > 1 ? features_rest : nullptr, (exp_scales ? ACT_EXP_SCALES : 0) | vec4, nullptr, color_stride, depth_channel, \
        radii, v_means2d, v_depths, v_conics, nullptr, v_colors, v_means, v_quats, v_scales, v_features_dc,         \
        K > 1 ? v_features_rest : nullptr, nullptr, nullptr)
    if (minb <= 4) FSB_PBWD(4); else FSB_PBWD(6);
#undef FSB_PBWD
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_isect_count(int64_t M, const float* means2d, const int32_t* radii, int tile_size, int tile_w,
                            int tile_h, int legacy_bbox, int32_t* tiles_per_gauss, void* stream) {
    if (M < 0 || tile_size <= 0) return FSB_E_ARG;
    if (M == 0) return 0;
    isect_count_kernel<<<fsb_div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(M, means2d, radii, tile_size, tile_w,
                                                                           tile_h, legacy_bbox, tiles_per_gauss);
    FSB_LAUNCH_CHECK();
    return 0;
}
