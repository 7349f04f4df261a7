// refine.cu — densify / prune bookkeeping of the DN-Splatter model as a handful of streaming kernels.
//
// Replaces what /root/reference/dn_splatter/dn_model.py:326-451 (refinement_after) does every refine_every steps
// through nerfstudio 1.1.3's split_gaussians / dup_gaussians / cull_gaussians / dup_in_optim / remove_from_optim
// (SURVEY.md A.7): ~100 torch launches with boolean-mask indexing, torch.cat and repeat() over all 7 parameter
// tensors and both Adam moments of each.
//
//   classify   per Gaussian: split / duplicate / nothing (dn_model.py:347-376)                 1 launch
//   index      order-preserving lists of the split and duplicated Gaussians (torch.where)      scan + 1 launch
//   keep       for every candidate row of the concatenated model
//              [N originals | samps x n_split children (sample-major) | n_dup copies]
//              its parent row and whether cull_gaussians would keep it                         1 launch
//   gather     out[o] = src[parent] for kept rows (zeros for the Adam moments of new rows),
//              order-preserving like tensor[~culls]                                            1 launch per tensor
//   split fix  children: mean = mu + R(q / |q|) (exp(s) * eps), scale = log(exp(s) / 1.6)      1 launch
//
// The normal samples eps come from the caller (torch.randn with the generator the reference would use), so a
// seeded run reproduces the reference's children exactly.  HBM-bound: ~59 * 4 * 3 * 2 bytes per surviving Gaussian.
#include "common.cuh"
#include "fs_math.cuh"

namespace {

constexpr int RF_THREADS = 256;

__global__ void __launch_bounds__(RF_THREADS)
refine_classify_kernel(int N, const float* __restrict__ xys_grad_norm, const float* __restrict__ vis_counts,
                       const float* __restrict__ max_2Dsize, const float* __restrict__ scales, float half_max_dim,
                       float densify_grad_thresh, float densify_size_thresh, float split_screen_size,
                       float size_fac, const uint8_t* __restrict__ add_mask, uint8_t* __restrict__ action,
                       int32_t* __restrict__ split_flag, int32_t* __restrict__ dup_flag) {
    const int n = blockIdx.x * RF_THREADS + threadIdx.x;
    if (n >= N) return;
    // avg_grad_norm = (xys_grad_norm / vis_counts) * 0.5 * max(H, W)
    const float avg = (xys_grad_norm[n] / vis_counts[n]) * 0.5f * half_max_dim;
    const bool high = avg > densify_grad_thresh;
    const float e0 = expf(scales[3 * (size_t)n]), e1 = expf(scales[3 * (size_t)n + 1]),
                e2 = expf(scales[3 * (size_t)n + 2]);
    const float smax = fmaxf(fmaxf(e0, e1), e2);
    bool split = smax > densify_size_thresh;
    if (split_screen_size > 0.f && max_2Dsize) split = split || (max_2Dsize[n] > split_screen_size);
    split = split && high;
    if (add_mask && add_mask[n]) split = false;
    // `dups` is evaluated AFTER split_gaussians has shrunk the split parents' scales in place (dn_model.py:369-375
    // over nerfstudio's split_gaussians): a split parent whose shrunk scale fits the threshold is duplicated too
    float dmax = smax;
    if (split) dmax = fmaxf(fmaxf(expf(logf(e0 / size_fac)), expf(logf(e1 / size_fac))), expf(logf(e2 / size_fac)));
    bool dup = (dmax <= densify_size_thresh) && high;
    if (add_mask && add_mask[n]) dup = false;
    action[n] = (split ? 1 : 0) | (dup ? 2 : 0);
    split_flag[n] = split ? 1 : 0;
    dup_flag[n] = dup ? 1 : 0;
}

__global__ void __launch_bounds__(RF_THREADS)
refine_index_kernel(int N, const uint8_t* __restrict__ action, const int64_t* __restrict__ split_rank,
                    const int64_t* __restrict__ dup_rank, int32_t* __restrict__ split_idcs,
                    int32_t* __restrict__ dup_idcs) {
    const int n = blockIdx.x * RF_THREADS + threadIdx.x;
    if (n >= N) return;
    const uint8_t a = action[n];
    if (a & 1) split_idcs[split_rank[n]] = n;
    if (a & 2) dup_idcs[dup_rank[n]] = n;
}

// candidate row j of the concatenated model -> parent row, keep flag
__global__ void __launch_bounds__(RF_THREADS)
refine_keep_kernel(int64_t M, int N, int n_split, int n_dup, int samps, const uint8_t* __restrict__ action,
                   const int32_t* __restrict__ split_idcs, const int32_t* __restrict__ dup_idcs,
                   const float* __restrict__ opacities, const float* __restrict__ scales,
                   const float* __restrict__ max_2Dsize, float cull_alpha_thresh, float cull_scale_thresh,
                   float cull_screen_size, float size_fac, const uint8_t* __restrict__ extra_cull,
                   int32_t* __restrict__ parent, int32_t* __restrict__ keep) {
    const int64_t j = (int64_t)blockIdx.x * RF_THREADS + threadIdx.x;
    if (j >= M) return;
    int p;
    bool child = false, is_new = false;
    if (j < N) {
        p = (int)j;
    } else if (j < (int64_t)N + (int64_t)samps * n_split) {
        p = split_idcs[(j - N) % n_split];  // sample-major: .repeat(samps, 1)
        child = true;
        is_new = true;
    } else {
        p = dup_idcs[j - N - (int64_t)samps * n_split];
        is_new = true;
    }
    // cull_gaussians(extra): sigmoid(opacity) < thresh | extra ; then optionally the too-big tests
    const float o = opacities[p];
    bool cull = (1.f / (1.f + expf(-o))) < cull_alpha_thresh;
    if (!is_new) {
        if (action && (action[p] & 1)) cull = true;  // a split Gaussian is replaced by its children
        if (extra_cull && extra_cull[p]) cull = true;
    }
    if (cull_scale_thresh > 0.f) {
        float sx = expf(scales[3 * (size_t)p]), sy = expf(scales[3 * (size_t)p + 1]), sz = expf(scales[3 * (size_t)p + 2]);
        if (child || (is_new && action && (action[p] & 1))) {
            // children (and duplicates of already-shrunk split parents) carry log(exp(s) / size_fac)
            sx = expf(logf(sx / size_fac)); sy = expf(logf(sy / size_fac)); sz = expf(logf(sz / size_fac));
        }
        bool toobig = fmaxf(fmaxf(sx, sy), sz) > cull_scale_thresh;
        if (cull_screen_size > 0.f && max_2Dsize && !is_new) toobig = toobig || (max_2Dsize[p] > cull_screen_size);
        cull = cull || toobig;
    }
    parent[j] = p;
    keep[j] = cull ? 0 : 1;
}

// out[offset[j], :] = new_row_zero && j >= N ? 0 : src[parent[j], :]   for kept j
__global__ void __launch_bounds__(RF_THREADS)
refine_gather_kernel(int64_t M, int N, int width, const float* __restrict__ src, const int32_t* __restrict__ parent,
                     const int32_t* __restrict__ keep, const int64_t* __restrict__ offsets, int zero_new,
                     float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * RF_THREADS + threadIdx.x;
    if (e >= M * width) return;
    const int64_t j = e / width;
    const int c = (int)(e - j * width);
    if (!keep[j]) return;
    const float v = (zero_new && j >= N) ? 0.f : src[(size_t)parent[j] * width + c];
    out[(size_t)offsets[j] * width + c] = v;
}

__global__ void __launch_bounds__(RF_THREADS)
refine_split_fixup_kernel(int64_t n_new, int64_t n_children, int N, const uint8_t* __restrict__ action,
                          const float* __restrict__ means, const float* __restrict__ scales,
                          const float* __restrict__ quats, const float* __restrict__ samples,
                          const int32_t* __restrict__ parent, const int32_t* __restrict__ keep,
                          const int64_t* __restrict__ offsets, float size_fac, float* __restrict__ out_means,
                          float* __restrict__ out_scales) {
    const int64_t i = (int64_t)blockIdx.x * RF_THREADS + threadIdx.x;
    if (i >= n_new) return;
    const int64_t j = (int64_t)N + i;
    if (!keep[j]) return;
    const int p = parent[j];
    if (i >= n_children) {
        // a duplicate: only the copies of split parents differ from a plain gather (their scales were shrunk)
        if (!(action[p] & 1)) return;
        const size_t o = (size_t)offsets[j];
#pragma unroll
        for (int r = 0; r < 3; ++r) out_scales[3 * o + r] = logf(expf(scales[3 * (size_t)p + r]) / size_fac);
        return;
    }
    const float es0 = expf(scales[3 * (size_t)p]), es1 = expf(scales[3 * (size_t)p + 1]),
                es2 = expf(scales[3 * (size_t)p + 2]);
    const float s0 = es0 * samples[3 * i], s1 = es1 * samples[3 * i + 1], s2 = es2 * samples[3 * i + 2];
    const float4 q = reinterpret_cast<const float4*>(quats)[p];
    const fs::Mat3 R = fs::quat_to_rotmat(q.x, q.y, q.z, q.w, nullptr);
    const size_t o = (size_t)offsets[j];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        out_means[3 * o + r] = (R.m[r][0] * s0 + R.m[r][1] * s1 + R.m[r][2] * s2) + means[3 * (size_t)p + r];
    out_scales[3 * o + 0] = logf(es0 / size_fac);
    out_scales[3 * o + 1] = logf(es1 / size_fac);
    out_scales[3 * o + 2] = logf(es2 / size_fac);
}

}  // namespace

// action[N] u8: bit 0 = split, bit 1 = duplicate; split_flag / dup_flag [N] i32: the same as 0/1 counts for
// fsb_isect_scan.  scales are the LOG scales the model stores; max_2Dsize / add_mask nullable;
// split_screen_size <= 0 disables the screen-size test (step >= stop_screen_size_at).
FSB_API int fsb_refine_classify(int N, const float* xys_grad_norm, const float* vis_counts, const float* max_2Dsize,
                                const float* scales, float max_dim, float densify_grad_thresh,
                                float densify_size_thresh, float split_screen_size, float size_fac,
                                const uint8_t* add_mask, uint8_t* action, int32_t* split_flag, int32_t* dup_flag,
                                void* stream) {
    if (N < 0) return FSB_E_ARG;
    if (N == 0) return 0;
    refine_classify_kernel<<<fsb_div_up(N, RF_THREADS), RF_THREADS, 0, (cudaStream_t)stream>>>(
        N, xys_grad_norm, vis_counts, max_2Dsize, scales, max_dim, densify_grad_thresh, densify_size_thresh,
        split_screen_size, size_fac, add_mask, action, split_flag, dup_flag);
    FSB_LAUNCH_CHECK();
    return 0;
}

// split_idcs[n_split], dup_idcs[n_dup] i32 = torch.where(mask)[0]; ranks = exclusive scans of the flags
FSB_API int fsb_refine_index(int N, const uint8_t* action, const int64_t* split_rank, const int64_t* dup_rank,
                             int32_t* split_idcs, int32_t* dup_idcs, void* stream) {
    if (N < 0) return FSB_E_ARG;
    if (N == 0) return 0;
    refine_index_kernel<<<fsb_div_up(N, RF_THREADS), RF_THREADS, 0, (cudaStream_t)stream>>>(
        N, action, split_rank, dup_rank, split_idcs, dup_idcs);
    FSB_LAUNCH_CHECK();
    return 0;
}

// M = N + samps * n_split + n_dup candidate rows -> parent[M] i32, keep[M] i32 (0/1, scan it with fsb_isect_scan).
// action nullable (no densification: plain cull_gaussians); extra_cull u8 [N] nullable (hull / touch masks);
// cull_scale_thresh <= 0 disables the too-big tests (step <= refine_every * reset_alpha_every),
// cull_screen_size <= 0 the screen-size one (step >= stop_screen_size_at).
FSB_API int fsb_refine_keep(int64_t M, int N, int n_split, int n_dup, int samps, const uint8_t* action,
                            const int32_t* split_idcs, const int32_t* dup_idcs, const float* opacities,
                            const float* scales, const float* max_2Dsize, float cull_alpha_thresh,
                            float cull_scale_thresh, float cull_screen_size, float size_fac,
                            const uint8_t* extra_cull, int32_t* parent, int32_t* keep, void* stream) {
    if (M < 0 || N < 0 || n_split < 0 || n_dup < 0 || samps < 0) return FSB_E_ARG;
    if (M != (int64_t)N + (int64_t)samps * n_split + n_dup) return FSB_E_ARG;
    if (M == 0) return 0;
    refine_keep_kernel<<<fsb_div_up(M, RF_THREADS), RF_THREADS, 0, (cudaStream_t)stream>>>(
        M, N, n_split, n_dup, samps, action, split_idcs, dup_idcs, opacities, scales, max_2Dsize, cull_alpha_thresh,
        cull_scale_thresh, cull_screen_size, size_fac, extra_cull, parent, keep);
    FSB_LAUNCH_CHECK();
    return 0;
}

// order-preserving compaction of one [N, width] tensor into out[n_kept, width]; zero_new = 1 writes zeros for
// rows j >= N (Adam moments of new Gaussians, dup_in_optim)
FSB_API int fsb_refine_gather(int64_t M, int N, int width, const float* src, const int32_t* parent,
                              const int32_t* keep, const int64_t* offsets, int zero_new, float* out, void* stream) {
    if (M < 0 || width <= 0) return FSB_E_ARG;
    if (M == 0) return 0;
    refine_gather_kernel<<<fsb_div_up(M * width, RF_THREADS), RF_THREADS, 0, (cudaStream_t)stream>>>(
        M, N, width, src, parent, keep, offsets, zero_new, out);
    FSB_LAUNCH_CHECK();
    return 0;
}

// new rows (candidates N .. N + n_new; the first n_children of them are split children): children get the sampled
// position and the shrunk scale, duplicates of split parents the shrunk scale.  samples[n_children, 3] ~ N(0, I)
// from the caller.
FSB_API int fsb_refine_split_fixup(int64_t n_new, int64_t n_children, int N, const uint8_t* action,
                                   const float* means, const float* scales, const float* quats,
                                   const float* samples, const int32_t* parent, const int32_t* keep,
                                   const int64_t* offsets, float size_fac, float* out_means, float* out_scales,
                                   void* stream) {
    if (n_new < 0 || n_children < 0 || n_children > n_new) return FSB_E_ARG;
    if (n_new == 0) return 0;
    refine_split_fixup_kernel<<<fsb_div_up(n_new, RF_THREADS), RF_THREADS, 0, (cudaStream_t)stream>>>(
        n_new, n_children, N, action, means, scales, quats, samples, parent, keep, offsets, size_fac, out_means,
        out_scales);
    FSB_LAUNCH_CHECK();
    return 0;
}
