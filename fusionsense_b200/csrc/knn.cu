// knn.cu — exact k nearest neighbours of 3-D points on a quantile grid, and the local-density field built on it.
//
// Replaces, on the mesh-export / SDF path of the reference:
//   * dn_splatter/utils/knn.py:29-43  knn_sk(x, y, k): sklearn NearestNeighbors(k + 1).fit(x).kneighbors(y)[:, 1:]
//     on the CPU (every call copies both clouds to the host and the indices back);
//   * dn_splatter/dn_model.py:1575-1635  get_density: ~15 torch launches over [S, K, 3, 3] temporaries.
//
// Structure: a g x g x g grid whose cell boundaries along each axis are the i/g QUANTILES of that coordinate (from a
// strided sample sorted by the library's radix sort), so a dense object inside a sparse room — the shape of a
// FusionSense scene — gets fine cells where the points are and coarse ones elsewhere; border cells are open-ended.
// Points are sorted by cell, and each query grows a box of cells around its own cell, always on the face nearest to
// it, until the k-th best distance is provably smaller than the distance to the nearest face.  Queries that would need
// more than `max_steps` growth steps (outliers in empty space) go to a list that a brute-force kernel finishes, so
// the result is exact for every query.  Distances are fp64 of the fp32 coordinates — the differences and
// squares are exact, like the KD-tree sklearn picks for 3-D data — and ties are ordered by index.
#include "common.cuh"

namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_MAX_G = 320;  // cells per axis (320^3 = 32.8 M cells)

struct KnnGrid {
    int g;                 // cells per axis
    const float* edges;    // [3][g + 1]: cell c of axis a covers [edges[a][c], edges[a][c + 1]); [0] and [g] are the
                           // sample's extremes, never used as boundaries (the border cells are open-ended)
};

__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// monotone map float -> uint32 (NaN / inf never get here)
__device__ __forceinline__ uint32_t ordered_bits(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// cell of coordinate p: the number of interior edges e[1 .. g-1] that are <= p (pure comparisons: a point and a query
// with the same coordinate always land in the same cell, and p in cell c  <=>  e[c] <= p < e[c + 1] away from the borders)
__device__ __forceinline__ int cell_of(const float* __restrict__ e, int g, float p) {
    int lo = 0, hi = g - 1;  // answer in [0, g - 1]
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (e[mid] <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- per-axis quantile edges --------------------------------------------------------------------------------
// keys[a * S + i] = (a << 32) | ordered bits of coordinate a of sample i (point i * stride); non-finite points sort
// to the end of their axis segment
__global__ void __launch_bounds__(256)
knn_axis_keys_kernel(int64_t S, int64_t stride, const float* __restrict__ pts, uint64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const float* p = pts + 3 * (i * stride);
    const bool fin = finite3(p[0], p[1], p[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a)
        keys[(int64_t)a * S + i] = ((uint64_t)a << 32) | (fin ? ordered_bits(p[a]) : 0xffffffffu);
}

__global__ void __launch_bounds__(256)
knn_edges_kernel(int64_t S, const uint64_t* __restrict__ sorted_keys, int g, float* __restrict__ edges) {
    const int a = blockIdx.x;
    const uint64_t* k = sorted_keys + (int64_t)a * S;
    // number of finite samples of this axis: first position whose low word is the non-finite marker
    int64_t lo = 0, hi = S;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((uint32_t)k[mid] == 0xffffffffu) hi = mid; else lo = mid + 1;
    }
    const int64_t n_fin = lo;
    for (int c = threadIdx.x; c <= g; c += blockDim.x) {
        float e = 0.f;
        if (n_fin > 0) {
            int64_t r = ((int64_t)c * n_fin) / g;  // c <= 320, n_fin < 2^31
            if (r >= n_fin) r = n_fin - 1;
            e = from_ordered_bits((uint32_t)k[r]);
        }
        edges[a * (g + 1) + c] = e;
    }
}

// ---- binning ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_cells_kernel(int64_t N, const float* __restrict__ pts, KnnGrid gr, uint64_t* __restrict__ keys,
                 int32_t* __restrict__ vals, int32_t* __restrict__ n_nonfinite) {
    extern __shared__ float s_edges[];
    for (int i = threadIdx.x; i < 3 * (gr.g + 1); i += blockDim.x) s_edges[i] = gr.edges[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const int g = gr.g;
    uint64_t key = (uint64_t)g * g * g;  // the cell after the grid: never visited
    if (finite3(x, y, z)) {
        const int cx = cell_of(s_edges, g, x);
        const int cy = cell_of(s_edges + (g + 1), g, y);
        const int cz = cell_of(s_edges + 2 * (g + 1), g, z);
        key = ((uint64_t)cz * g + cy) * g + cx;
    } else if (n_nonfinite) {
        atomicAdd(n_nonfinite, 1);
    }
    keys[i] = key;
    vals[i] = (int32_t)i;
}

// cell_start[c] = first sorted position whose key is >= c, c in [0, n_cells]
__global__ void __launch_bounds__(256)
knn_cell_start_kernel(int64_t N, const uint64_t* __restrict__ keys, int64_t n_cells, int32_t* __restrict__ cell_start) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < (uint64_t)c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = (int32_t)lo;
}

// sorted_pts[i] = point vals[i] with its index in .w
__global__ void __launch_bounds__(256)
knn_gather_kernel(int64_t N, const int32_t* __restrict__ vals, const float* __restrict__ pts,
                  float4* __restrict__ sorted_pts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int32_t v = vals[i];
    sorted_pts[i] = make_float4(pts[3 * (size_t)v], pts[3 * (size_t)v + 1], pts[3 * (size_t)v + 2], __int_as_float(v));
}

// ---- query --------------------------------------------------------------------------------------------------
// sorted candidate list of the K best (distance^2, index), lexicographic order
template <int KCAP>
struct Best {
    double d[KCAP];
    int32_t id[KCAP];
    int n;
    __device__ __forceinline__ void offer(double dd, int32_t ii, int K) {
        if (n == K) {
            if (dd > d[K - 1] || (dd == d[K - 1] && ii > id[K - 1])) return;
        }
        int j = n < K ? n : K - 1;
        while (j > 0 && (d[j - 1] > dd || (d[j - 1] == dd && id[j - 1] > ii))) {
            d[j] = d[j - 1]; id[j] = id[j - 1];
            --j;
        }
        d[j] = dd; id[j] = ii;
        if (n < K) ++n;
    }
};

template <int KCAP>
__global__ void __launch_bounds__(KNN_THREADS)
knn_query_kernel(int64_t Ny, const float* __restrict__ y, const int32_t* __restrict__ order, KnnGrid gr,
                 const int32_t* __restrict__ cell_start, const float4* __restrict__ sorted_pts, int K, int drop_first,
                 int max_steps, int64_t* __restrict__ out_idx, double* __restrict__ out_dist,
                 int32_t* __restrict__ unresolved, int32_t* __restrict__ n_unresolved) {
    extern __shared__ float s_edges[];
    const int g = gr.g;
    for (int i = threadIdx.x; i < 3 * (g + 1); i += blockDim.x) s_edges[i] = gr.edges[i];
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Ny) return;
    const int64_t qi = order ? order[t] : t;
    const float qxf = y[3 * qi], qyf = y[3 * qi + 1], qzf = y[3 * qi + 2];
    const int KO = K - drop_first;
    if (!finite3(qxf, qyf, qzf)) {  // sklearn refuses NaN input; the caller nan_to_num's (dn_model.py:180)
        for (int k = 0; k < KO; ++k) {
            out_idx[qi * KO + k] = -1;
            if (out_dist) out_dist[qi * KO + k] = NAN;
        }
        return;
    }
    const float* ex = s_edges;
    const float* ey = s_edges + (g + 1);
    const float* ez = s_edges + 2 * (g + 1);
    const double qx = qxf, qy = qyf, qz = qzf;
    const int cx = cell_of(ex, g, qxf), cy = cell_of(ey, g, qyf), cz = cell_of(ez, g, qzf);
    Best<KCAP> best;
    best.n = 0;
    auto scan = [&](int b, int e) {
        for (int i = b; i < e; ++i) {
            const float4 p = sorted_pts[i];
            const double dx = qx - (double)p.x, dy = qy - (double)p.y, dz = qz - (double)p.z;
            // every product and sum rounded on its own (no FMA): the bits numpy and sklearn's KD-tree compute
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            best.offer(d2, __float_as_int(p.w), K);
        }
    };
    // The visited set is a box of cells [lo, hi] per axis, grown one slab at a time on the face that is NEAREST to the
    // query (cells are as wide as the local quantiles make them — a few millimetres across a dense object, decimetres in
    // the room around it — so growing all six faces in step would crawl along the thin axes).  Every unvisited point lies
    // beyond one of the six faces; border cells are open-ended, so a face on the border has nothing behind it.
    int lo[3] = {cx, cy, cz}, hi[3] = {cx, cy, cz};
    {
        const int64_t c = ((int64_t)cz * g + cy) * g + cx;
        scan(cell_start[c], cell_start[c + 1]);
    }
    const double q3[3] = {qx, qy, qz};
    const float* e3[3] = {ex, ey, ez};
    bool done = false;
    for (int step = 0; step <= max_steps; ++step) {
        double reach = INFINITY;
        int face = -1;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (lo[a] >= 1) {
                const double d = q3[a] - (double)e3[a][lo[a]];
                if (d < reach) { reach = d; face = 2 * a; }
            }
            if (hi[a] + 1 <= g - 1) {
                const double d = (double)e3[a][hi[a] + 1] - q3[a];
                if (d < reach) { reach = d; face = 2 * a + 1; }
            }
        }
        if (face < 0) { done = true; break; }  // the whole grid has been visited
        // strict: an unvisited point at exactly the k-th distance could still win the tie on its index
        if (best.n == K && best.d[K - 1] < reach * reach) { done = true; break; }
        if (step == max_steps) break;
        const int a = face >> 1;
        const int idx = (face & 1) ? ++hi[a] : --lo[a];
        if (a == 0) {          // a y-z sheet of single cells
            for (int zz = lo[2]; zz <= hi[2]; ++zz)
                for (int yy = lo[1]; yy <= hi[1]; ++yy) {
                    const int64_t c = ((int64_t)zz * g + yy) * g + idx;
                    scan(cell_start[c], cell_start[c + 1]);
                }
        } else if (a == 1) {   // one run of cells (consecutive in the sorted order too) per z
            for (int zz = lo[2]; zz <= hi[2]; ++zz) {
                const int64_t row = ((int64_t)zz * g + idx) * g;
                scan(cell_start[row + lo[0]], cell_start[row + hi[0] + 1]);
            }
        } else {               // one run per y
            for (int yy = lo[1]; yy <= hi[1]; ++yy) {
                const int64_t row = ((int64_t)idx * g + yy) * g;
                scan(cell_start[row + lo[0]], cell_start[row + hi[0] + 1]);
            }
        }
    }
    if (!done) {
        const int slot = atomicAdd(n_unresolved, 1);
        unresolved[slot] = (int32_t)qi;
        return;
    }
    for (int k = 0; k < KO; ++k) {
        const int s = k + drop_first;
        out_idx[qi * KO + k] = s < best.n ? (int64_t)best.id[s] : -1;
        if (out_dist) out_dist[qi * KO + k] = s < best.n ? sqrt(best.d[s]) : INFINITY;
    }
}

// Queries the grid walk gave up on: one CTA per query, K rounds of "smallest (distance, index) above the previous pick"
// over all points.  O(K N) per query, for the handful of outliers only.
__global__ void __launch_bounds__(256)
knn_brute_kernel(int64_t Nx, const float* __restrict__ x, const float* __restrict__ y,
                 const int32_t* __restrict__ unresolved, const int32_t* __restrict__ n_unresolved, int K, int drop_first,
                 int64_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    __shared__ double sd[256];
    __shared__ int32_t si[256];
    const int KO = K - drop_first;
    const int total = *n_unresolved;
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        const int64_t qi = unresolved[u];
        const double qx = y[3 * qi], qy = y[3 * qi + 1], qz = y[3 * qi + 2];
        double pd = -1.0;
        int32_t pi = -1;
        for (int k = 0; k < K; ++k) {
            double bd = INFINITY;
            int32_t bi = 0x7fffffff;
            for (int64_t i = threadIdx.x; i < Nx; i += blockDim.x) {
                const float px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
                if (!finite3(px, py, pz)) continue;
                const double dx = qx - (double)px, dy = qy - (double)py, dz = qz - (double)pz;
                const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                const int32_t ii = (int32_t)i;
                const bool after_prev = d > pd || (d == pd && ii > pi);
                if (after_prev && (d < bd || (d == bd && ii < bi))) { bd = d; bi = ii; }
            }
            sd[threadIdx.x] = bd; si[threadIdx.x] = bi;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if ((int)threadIdx.x < o) {
                    const double od = sd[threadIdx.x + o];
                    const int32_t oi = si[threadIdx.x + o];
                    if (od < sd[threadIdx.x] || (od == sd[threadIdx.x] && oi < si[threadIdx.x])) {
                        sd[threadIdx.x] = od; si[threadIdx.x] = oi;
                    }
                }
                __syncthreads();
            }
            pd = sd[0]; pi = si[0];
            const bool found = pi != 0x7fffffff;
            if (threadIdx.x == 0 && k >= drop_first) {
                out_idx[qi * KO + (k - drop_first)] = found ? (int64_t)pi : -1;
                if (out_dist) out_dist[qi * KO + (k - drop_first)] = found ? sqrt(pd) : INFINITY;
            }
            __syncthreads();
            if (!found) { pd = INFINITY; pi = 0x7fffffff; }
        }
    }
}

// ---- density ------------------------------------------------------------------------------------------------
// dn_model.py:1596-1634 with scale_rot_to_inv_cov3d (:2141-2150, return_sqrt=True) and gsplat's quat_to_rotmat
// (normalised wxyz) folded in: one thread per sample, K gathered Gaussians.
__global__ void __launch_bounds__(256)
gaussian_density_kernel(int64_t S, const float* __restrict__ samples, int K, const int64_t* __restrict__ knn,
                        const float* __restrict__ means, const float* __restrict__ log_scales,
                        const float* __restrict__ quats, const float* __restrict__ opacity_logits, float clamp_min,
                        float* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float sx = samples[3 * s], sy = samples[3 * s + 1], sz = samples[3 * s + 2];
    float dens = 0.f;
    for (int k = 0; k < K; ++k) {
        const int64_t g = knn[s * K + k];
        if (g < 0) continue;  // a neighbour slot the search left empty (non-finite query)
        const float dx = sx - means[3 * g], dy = sy - means[3 * g + 1], dz = sz - means[3 * g + 2];
        const float i0 = 1.0f / fmaxf(expf(log_scales[3 * g]), 1e-3f);
        const float i1 = 1.0f / fmaxf(expf(log_scales[3 * g + 1]), 1e-3f);
        const float i2 = 1.0f / fmaxf(expf(log_scales[3 * g + 2]), 1e-3f);
        float w = quats[4 * g], x = quats[4 * g + 1], y = quats[4 * g + 2], z = quats[4 * g + 3];
        const float qn = fmaxf(sqrtf(w * w + x * x + y * y + z * z), 1e-12f);  // F.normalize's eps
        w /= qn; x /= qn; y /= qn; z /= qn;
        const float r00 = 1.f - 2.f * (y * y + z * z), r01 = 2.f * (x * y - w * z), r02 = 2.f * (x * z + w * y);
        const float r10 = 2.f * (x * y + w * z), r11 = 1.f - 2.f * (x * x + z * z), r12 = 2.f * (y * z - w * x);
        const float r20 = 2.f * (x * z - w * y), r21 = 2.f * (y * z + w * x), r22 = 1.f - 2.f * (x * x + y * y);
        // (M^T d)_j = inv_scale_j * sum_i R[i][j] d_i
        const float m0 = (r00 * i0) * dx + (r10 * i0) * dy + (r20 * i0) * dz;
        const float m1 = (r01 * i1) * dx + (r11 * i1) * dy + (r21 * i1) * dz;
        const float m2 = (r02 * i2) * dx + (r12 * i2) * dy + (r22 * i2) * dz;
        const float maha = fminf(fmaxf(m0 * m0 + m1 * m1 + m2 * m2, 0.f), 1e8f);
        const float op = 1.f / (1.f + expf(-opacity_logits[g]));
        dens += op * expf(-0.5f * maha);
    }
    if (dens >= 1.0f) dens = dens / (dens + 1e-5f);
    out[s] = fmaxf(dens, clamp_min);
}


// ---- level-set search along camera rays -----------------------------------------------------------------------
// dn_model.py:1766-1870 (compute_level_surface_points) for one back-projected point per thread: the first neighbour's
// standard deviation along the ray, 21 samples at linspace(-3, 3, 21) * std around the point, the density of the K
// neighbours at every sample (the inlined get_density of :1806-1840: normalised above 1, NOT clamped below), and per
// surface level the first sample above it with the linear interpolation of :1858-1880.  Every neighbour's M = R diag(1/s)
// is built once and used for the 21 samples; nothing of the reference's [P*21, K, 3, 3] temporaries exists.
constexpr int LS_SAMPLES = 21;
constexpr int LS_MAX_LEVELS = 4;

struct LevelArgs {
    float cam[3];
    float lin[LS_SAMPLES];        // torch.linspace(-3, 3, 21), from the host
    float levels[LS_MAX_LEVELS];
    int n_levels;
};

__device__ __forceinline__ void quat_rot(float w, float x, float y, float z, float (&r)[9]) {
    const float qn = fmaxf(sqrtf(w * w + x * x + y * y + z * z), 1e-12f);  // F.normalize inside quat_to_rotmat
    w /= qn; x /= qn; y /= qn; z /= qn;
    r[0] = 1.f - 2.f * (y * y + z * z); r[1] = 2.f * (x * y - w * z); r[2] = 2.f * (x * z + w * y);
    r[3] = 2.f * (x * y + w * z); r[4] = 1.f - 2.f * (x * x + z * z); r[5] = 2.f * (y * z - w * x);
    r[6] = 2.f * (x * z - w * y); r[7] = 2.f * (y * z + w * x); r[8] = 1.f - 2.f * (x * x + y * y);
}

__global__ void __launch_bounds__(128)
level_crossings_kernel(int64_t P, const float* __restrict__ points, int K, const int64_t* __restrict__ knn,
                       const float* __restrict__ means, const float* __restrict__ log_scales,
                       const float* __restrict__ quats, const float* __restrict__ opacity_logits, LevelArgs a,
                       float* __restrict__ t_out, uint8_t* __restrict__ valid_out, float* __restrict__ dens_out,
                       float* __restrict__ std_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float px = points[3 * i], py = points[3 * i + 1], pz = points[3 * i + 2];
    // camera_to_samples = F.normalize(points - cam)  (:1790-1792)
    float dx = px - a.cam[0], dy = py - a.cam[1], dz = pz - a.cam[2];
    {
        const float n = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
        dx /= n; dy /= n; dz /= n;
    }
    // std of the FIRST neighbour along its own view direction (:1766-1776): |exp(s) * (R^T v)|, v = (cam - mu) / |cam - mu|,
    // R from the normalised quaternion (quat_to_rotmat(invert_quaternion(q / |q|)) = R^T)
    float std = 0.f;
    if (knn[i * K] >= 0) {
        const int64_t g = knn[i * K];
        float w = quats[4 * g], x = quats[4 * g + 1], y = quats[4 * g + 2], z = quats[4 * g + 3];
        const float qn = sqrtf(w * w + x * x + y * y + z * z);
        w /= qn; x /= qn; y /= qn; z /= qn;
        float r[9];
        quat_rot(w, x, y, z, r);
        float vx = a.cam[0] - means[3 * g], vy = a.cam[1] - means[3 * g + 1], vz = a.cam[2] - means[3 * g + 2];
        const float vn = sqrtf(vx * vx + vy * vy + vz * vz);
        vx /= vn; vy /= vn; vz /= vn;
        const float e0 = expf(log_scales[3 * g]) * (r[0] * vx + r[3] * vy + r[6] * vz);
        const float e1 = expf(log_scales[3 * g + 1]) * (r[1] * vx + r[4] * vy + r[7] * vz);
        const float e2 = expf(log_scales[3 * g + 2]) * (r[2] * vx + r[5] * vy + r[8] * vz);
        std = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    }
    if (std_out) std_out[i] = std;
    float dens[LS_SAMPLES];
#pragma unroll
    for (int s = 0; s < LS_SAMPLES; ++s) dens[s] = 0.f;
    for (int k = 0; k < K; ++k) {
        const int64_t g = knn[i * K + k];
        if (g < 0) continue;
        const float mx = means[3 * g], my = means[3 * g + 1], mz = means[3 * g + 2];
        const float i0 = 1.0f / fmaxf(expf(log_scales[3 * g]), 1e-3f);
        const float i1 = 1.0f / fmaxf(expf(log_scales[3 * g + 1]), 1e-3f);
        const float i2 = 1.0f / fmaxf(expf(log_scales[3 * g + 2]), 1e-3f);
        float r[9];
        quat_rot(quats[4 * g], quats[4 * g + 1], quats[4 * g + 2], quats[4 * g + 3], r);
        const float op = 1.f / (1.f + expf(-opacity_logits[g]));
        const float m00 = r[0] * i0, m10 = r[3] * i0, m20 = r[6] * i0;
        const float m01 = r[1] * i1, m11 = r[4] * i1, m21 = r[7] * i1;
        const float m02 = r[2] * i2, m12 = r[5] * i2, m22 = r[8] * i2;
#pragma unroll
        for (int s = 0; s < LS_SAMPLES; ++s) {
            const float t = __fmul_rn(a.lin[s], std);
            // samples = points + points_range * camera_to_samples: product and sum rounded separately, like torch
            const float sx = __fadd_rn(px, __fmul_rn(t, dx)) - mx;
            const float sy = __fadd_rn(py, __fmul_rn(t, dy)) - my;
            const float sz = __fadd_rn(pz, __fmul_rn(t, dz)) - mz;
            const float a0 = m00 * sx + m10 * sy + m20 * sz;
            const float a1 = m01 * sx + m11 * sy + m21 * sz;
            const float a2 = m02 * sx + m12 * sy + m22 * sz;
            const float maha = fminf(fmaxf(a0 * a0 + a1 * a1 + a2 * a2, 0.f), 1e8f);
            dens[s] += op * expf(-0.5f * maha);
        }
    }
#pragma unroll
    for (int s = 0; s < LS_SAMPLES; ++s) {
        if (dens[s] >= 1.0f) dens[s] = dens[s] / (dens[s] + 1e-5f);
        if (dens_out) dens_out[i * LS_SAMPLES + s] = dens[s];
    }
    for (int l = 0; l < a.n_levels; ++l) {
        const float L = a.levels[l];
        // first sample above the level (:1848-1850); a ray is empty when its first sample is not under the level or no
        // sample is above it
        int first = 0;
#pragma unroll
        for (int s = LS_SAMPLES - 1; s >= 1; --s)
            if (dens[s] - L > 0.f) first = s;
        const bool first_under = (dens[0] - L < 0.f);
        const bool ok = first_under && first > 0;
        float t = 0.f;
        if (ok) {
            float v1 = 0.f, v0 = 0.f;
#pragma unroll
            for (int s = 1; s < LS_SAMPLES; ++s)
                if (s == first) { v1 = dens[s]; v0 = dens[s - 1]; }
            const float t1 = __fmul_rn(a.lin[first], std), t0 = __fmul_rn(a.lin[first - 1], std);
            // (L - v0) / (v1 - v0) * (t1 - t0) + t0   (:1873-1875)
            t = __fadd_rn(__fmul_rn(__fdiv_rn(L - v0, v1 - v0), t1 - t0), t0);
        }
        t_out[(int64_t)l * P + i] = t;
        valid_out[(int64_t)l * P + i] = ok ? 1 : 0;
    }
}

}  // namespace

static int knn_g_ok(int g) { return g >= 1 && g <= KNN_MAX_G; }

// Per-axis quantile edges from a strided sample of the cloud (S = number of samples, point i * stride each):
//   fsb_knn_axis_keys: keys[3 S] u64 = (axis << 32) | order-preserving bits of the coordinate; sort them on bits [0, 34)
//   fsb_knn_edges    : edges[3][g + 1] f32 = the i/g quantiles of each axis (i = 0 .. g) from the sorted keys
FSB_API int fsb_knn_axis_keys(int64_t S, int64_t stride, const float* pts, uint64_t* keys, void* stream) {
    if (S <= 0 || stride < 1 || !pts || !keys) return FSB_E_ARG;
    knn_axis_keys_kernel<<<fsb_div_up(S, 256), 256, 0, (cudaStream_t)stream>>>(S, stride, pts, keys);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_knn_edges(int64_t S, const uint64_t* sorted_keys, int g, float* edges, void* stream) {
    if (S <= 0 || !knn_g_ok(g) || !sorted_keys || !edges) return FSB_E_ARG;
    knn_edges_kernel<<<3, 256, 0, (cudaStream_t)stream>>>(S, sorted_keys, g, edges);
    FSB_LAUNCH_CHECK();
    return 0;
}

// keys[i] = linear cell ((cz g + cy) g + cx) of point i in the g^3 grid of `edges` (non-finite points -> g^3 and counted
// into n_nonfinite, nullable i32, not zeroed here), vals[i] = i; sort the pairs with fsb_radix_sort_pairs
FSB_API int fsb_knn_cells(int64_t N, const float* pts, int g, const float* edges, uint64_t* keys, int32_t* vals,
                          int32_t* n_nonfinite, void* stream) {
    if (N < 0 || N > 0x7fffffff || !knn_g_ok(g) || !edges || (N > 0 && (!pts || !keys || !vals))) return FSB_E_ARG;
    if (N == 0) return 0;
    KnnGrid gr{g, edges};
    knn_cells_kernel<<<fsb_div_up(N, 256), 256, 3 * (g + 1) * sizeof(float), (cudaStream_t)stream>>>(N, pts, gr, keys,
                                                                                                    vals, n_nonfinite);
    FSB_LAUNCH_CHECK();
    return 0;
}

// from the pairs sorted by key: cell_start[g^3 + 1] and the points in sorted order (xyz + index bits)
FSB_API int fsb_knn_build(int64_t N, const uint64_t* sorted_keys, const int32_t* sorted_vals, const float* pts, int g,
                          int32_t* cell_start, float* sorted_pts, void* stream) {
    if (N <= 0 || N > 0x7fffffff || !knn_g_ok(g) || !sorted_keys || !sorted_vals || !pts || !cell_start || !sorted_pts)
        return FSB_E_ARG;
    const int64_t n_cells = (int64_t)g * g * g;
    knn_cell_start_kernel<<<fsb_div_up(n_cells + 1, 256), 256, 0, (cudaStream_t)stream>>>(N, sorted_keys, n_cells,
                                                                                        cell_start);
    FSB_LAUNCH_CHECK();
    knn_gather_kernel<<<fsb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(N, sorted_vals, pts, (float4*)sorted_pts);
    FSB_LAUNCH_CHECK();
    return 0;
}

// The K nearest points of every query, nearest first, ties by index; the first `drop_first` of them are not written
// (knn_sk drops the nearest one: the query itself when y is x).  out_idx [Ny, K - drop_first] i64, out_dist (nullable)
// f64 distances.  order (nullable): the sequence in which queries are processed (sorted by cell for locality).
// unresolved i32[Ny] + n_unresolved i32[1] (zeroed here): queries left to fsb_knn_brute.
FSB_API int fsb_knn_query(int64_t Ny, const float* y, const int32_t* order, int g, const float* edges,
                          const int32_t* cell_start, const float* sorted_pts, int K, int drop_first, int max_steps,
                          int64_t* out_idx, double* out_dist, int32_t* unresolved, int32_t* n_unresolved,
                          void* stream) {
    if (Ny < 0 || Ny > 0x7fffffff || !knn_g_ok(g) || K < 1 || K > 33 || drop_first < 0 || drop_first >= K ||
        max_steps < 0 || !n_unresolved)
        return FSB_E_ARG;
    FSB_CUDA(cudaMemsetAsync(n_unresolved, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (Ny == 0) return 0;
    if (!y || !edges || !cell_start || !sorted_pts || !out_idx || !unresolved) return FSB_E_ARG;
    KnnGrid gr{g, edges};
    const int blocks = fsb_div_up(Ny, KNN_THREADS);
    const size_t sh = 3 * (g + 1) * sizeof(float);
    if (K <= 17)
        knn_query_kernel<17><<<blocks, KNN_THREADS, sh, (cudaStream_t)stream>>>(
            Ny, y, order, gr, cell_start, (const float4*)sorted_pts, K, drop_first, max_steps, out_idx, out_dist,
            unresolved, n_unresolved);
    else
        knn_query_kernel<33><<<blocks, KNN_THREADS, sh, (cudaStream_t)stream>>>(
            Ny, y, order, gr, cell_start, (const float4*)sorted_pts, K, drop_first, max_steps, out_idx, out_dist,
            unresolved, n_unresolved);
    FSB_LAUNCH_CHECK();
    return 0;
}

FSB_API int fsb_knn_brute(int64_t Nx, const float* x, const float* y, const int32_t* unresolved,
                          const int32_t* n_unresolved, int K, int drop_first, int64_t* out_idx, double* out_dist,
                          void* stream) {
    if (Nx <= 0 || Nx > 0x7fffffff || !x || !y || !unresolved || !n_unresolved || K < 1 || drop_first < 0 ||
        drop_first >= K || !out_idx)
        return FSB_E_ARG;
    knn_brute_kernel<<<FSB_NUM_SMS * 2, 256, 0, (cudaStream_t)stream>>>(Nx, x, y, unresolved, n_unresolved, K,
                                                                       drop_first, out_idx, out_dist);
    FSB_LAUNCH_CHECK();
    return 0;
}

// out[s] = max(normalised sum_k sigmoid(opacity[g]) exp(-0.5 |M_g^T (sample - mean_g)|^2), clamp_min), g = knn[s, k]
// (get_density clamps at 1e-4; the copy of it inlined in compute_level_surface_points does not: clamp_min = 0)
FSB_API int fsb_gaussian_density(int64_t S, const float* samples, int K, const int64_t* knn, const float* means,
                                 const float* log_scales, const float* quats, const float* opacity_logits,
                                 float clamp_min, float* out, void* stream) {
    if (S < 0 || K < 1) return FSB_E_ARG;
    if (S == 0) return 0;
    if (!samples || !knn || !means || !log_scales || !quats || !opacity_logits || !out) return FSB_E_ARG;
    gaussian_density_kernel<<<fsb_div_up(S, 256), 256, 0, (cudaStream_t)stream>>>(S, samples, K, knn, means, log_scales,
                                                                                 quats, opacity_logits, clamp_min, out);
    FSB_LAUNCH_CHECK();
    return 0;
}

// Level-set search along the camera rays of P back-projected points (dn_model.py:1766-1880): per point 21 samples at
// lin[s] * std (lin = torch.linspace(-3, 3, 21) from the host, std = the first neighbour's extent along its view
// direction), densities from the K neighbours knn[P, K], and for each of n_levels <= 4 surface levels the first crossing:
// t_out[l, p] (ray parameter of the intersection, 0 where none) and valid_out[l, p] (u8).  dens_out (nullable) [P, 21],
// std_out (nullable) [P].  cam, lin, levels are HOST pointers.
FSB_API int fsb_level_crossings(int64_t P, const float* points, const float* cam, int K, const int64_t* knn,
                                const float* means, const float* log_scales, const float* quats,
                                const float* opacity_logits, const float* lin, int n_levels, const float* levels,
                                float* t_out, uint8_t* valid_out, float* dens_out, float* std_out, void* stream) {
    if (P < 0 || K < 1 || n_levels < 1 || n_levels > LS_MAX_LEVELS || !cam || !lin || !levels) return FSB_E_ARG;
    if (P == 0) return 0;
    if (!points || !knn || !means || !log_scales || !quats || !opacity_logits || !t_out || !valid_out) return FSB_E_ARG;
    LevelArgs a;
    for (int i = 0; i < 3; ++i) a.cam[i] = cam[i];
    for (int i = 0; i < LS_SAMPLES; ++i) a.lin[i] = lin[i];
    for (int i = 0; i < LS_MAX_LEVELS; ++i) a.levels[i] = i < n_levels ? levels[i] : 0.f;
    a.n_levels = n_levels;
    level_crossings_kernel<<<fsb_div_up(P, 128), 128, 0, (cudaStream_t)stream>>>(
        P, points, K, knn, means, log_scales, quats, opacity_logits, a, t_out, valid_out, dens_out, std_out);
    FSB_LAUNCH_CHECK();
    return 0;
}
