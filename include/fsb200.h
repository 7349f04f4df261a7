/* fsb200.h — C ABI of libfsb200.so, the B200 (sm_100a) hot path of FusionSense / DN-Splatter.
 *
 * Each entry point replaces one stage that the reference reaches through gsplat==1.0.0 /
 * nerfstudio==1.1.3 / torch (not vendored in the reference tree), or one
 * first-party function of the reference.  The "replaces" notes cite the reference call site
 * (paths relative to the reference root) that ends up in that stage.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked "host";
 *  - fp32 row-major contiguous tensors; shapes in brackets; "nullable" pointers may be NULL;
 *  - the caller owns every buffer including workspaces; nothing here allocates or frees;
 *  - stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant, no global state;
 *  - return value: 0 on success, otherwise a cudaError_t value or FSB_E_ARG (10001) for a refused
 *    argument; nothing throws across the ABI.
 */
#ifndef FSB200_H
#define FSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSB_E_ARG 10001

/* ------------------------------------------------------------------------------------------------
 * P1+P2+I1(count): fused EWA projection, SH->RGB and tiles-per-Gaussian.
 * replaces gsplat fully_fused_projection + spherical_harmonics + clamp_min(+0.5) + isect_tiles pass 1,
 * called from dn_splatter/dn_model.py:570-591.
 *   means[N,3] quats[N,4](wxyz) scales[N,3] viewmats[C,4,4] Ks[C,3,3]
 *   sh_degree -1: no colour evaluation; else coeffs[N,K,3] and colors[C,N,color_stride] must be given:
 *   channels 0..2 = max(SH+0.5,0).  campos[C,3] = camera centres (torch.linalg.inv(viewmats)[:, :3, 3]); nullable:
 *   NULL means -R^-1 t is evaluated inside the kernel.
 *   depth_channel >= 0: depths are also written into colors[..., depth_channel] ("RGB+ED").
 *   outputs: radii[C,N] i32 (0 = culled), means2d[C,N,2], depths[C,N], conics[C,N,3],
 *   comps[C,N] (nullable), tiles_per_gauss[C,N] i32.  Culled entries are written as zeros.
 *   legacy_extra (nullable, i64, caller zero-fills): ACCUMULATES the number of tiles by which the gsplat 0.1.x
 *   bbox rule (fsb_isect_count legacy_bbox=1) differs from tiles_per_gauss; 0 means the legacy normals pass
 *   (dn_model.py:644-653) bins to exactly the same sorted lists. */
int fsb_project_sh_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                       const float* viewmats, const float* Ks, int width, int height, float eps2d,
                       float near_plane, float far_plane, float radius_clip, int tile_size, int tile_w,
                       int tile_h, int sh_degree, int K, const float* coeffs, const float* campos,
                       int color_stride, int depth_channel, int32_t* radii, float* means2d, float* depths,
                       float* conics, float* comps, float* colors, int32_t* tiles_per_gauss,
                       int64_t* legacy_extra, void* stream);

/* P3: backward of the above.  replaces gsplat fully_fused_projection_bwd + compute_sh_bwd
 * (loss.backward() of dn_splatter/dn_model.py:570-591).
 *   in : v_means2d[C,N,2] v_depths[C,N](nullable) v_conics[C,N,3] v_comps[C,N](nullable)
 *        v_colors[C,N,color_stride](nullable; channel depth_channel is added to v_depths)
 *   out: v_means[N,3] v_quats[N,4] v_scales[N,3] v_coeffs[N,K,3](nullable) — overwritten, summed over cameras;
 *        v_viewmats[C,4,4], v_campos[C,3] (nullable) — ACCUMULATED (atomicAdd), caller zero-fills. */
int fsb_project_sh_bwd(int C, int N, const float* means, const float* quats, const float* scales,
                       const float* viewmats, const float* Ks, int width, int height, float eps2d,
                       int sh_degree, int K, const float* coeffs, const float* campos, int color_stride,
                       int depth_channel, const int32_t* radii, const float* v_means2d, const float* v_depths,
                       const float* v_conics, const float* v_comps, const float* v_colors, float* v_means,
                       float* v_quats, float* v_scales, float* v_coeffs, float* v_viewmats, float* v_campos,
                       void* stream);

/* P1+P2+I1(count) and P3 fed with the model's parameters AS STORED (dn_splatter/dn_model.py:294-304), fusing the
 * activations and the concatenation that dn_model.py:566-574 issues as separate torch launches before the call:
 *   scales: log-scales when exp_scales != 0 (exp applied inside; v_scales is then d/d log-scale);
 *   quats: any norm (normalised inside either way; v_quats is w.r.t. the stored quaternion);
 *   SH coefficients as two tensors, features_dc[N,3] (band 0) and features_rest[N,K-1,3] (NULL iff K == 1), read
 *   and differentiated in place instead of through torch.cat -> [N,K,3] and the split of its gradient.
 * K <= 16, 0 <= sh_degree <= 3, camera centres are evaluated inside, no compensation output, no view-matrix
 * gradient (camera optimiser off: dn_model.py:128-130).  Outputs as fsb_project_sh_fwd / _bwd, all overwritten. */
int fsb_project_params_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                           int exp_scales, const float* viewmats, const float* Ks, int width, int height,
                           float eps2d, float near_plane, float far_plane, float radius_clip, int tile_size,
                           int tile_w, int tile_h, int sh_degree, int K, const float* features_dc,
                           const float* features_rest, int color_stride, int depth_channel, int32_t* radii,
                           float* means2d, float* depths, float* conics, float* colors,
                           int32_t* tiles_per_gauss, int64_t* legacy_extra, void* stream);
int fsb_project_params_bwd(int C, int N, const float* means, const float* quats, const float* scales,
                           int exp_scales, const float* viewmats, const float* Ks, int width, int height,
                           float eps2d, int sh_degree, int K, const float* features_dc,
                           const float* features_rest, int color_stride, int depth_channel,
                           const int32_t* radii, const float* v_means2d, const float* v_depths,
                           const float* v_conics, const float* v_colors, float* v_means, float* v_quats,
                           float* v_scales, float* v_features_dc, float* v_features_rest, void* stream);

/* I1 (count only) for callers that bring their own 2-D means / radii.
 * legacy_bbox = 1 selects the gsplat 0.1.x rule ((int) truncation, +1 on the max side) used by
 * gsplat.rasterize_gaussians (dn_splatter/dn_model.py:644-653); 0 = floor/ceil rule of gsplat 1.0. */
int fsb_isect_count(int64_t M, const float* means2d, const int32_t* radii, int tile_size, int tile_w,
                    int tile_h, int legacy_bbox, int32_t* tiles_per_gauss, void* stream);

/* I1 with an exact reach test: count / emit only the tiles of the bounding box on which the Gaussian can pass the
 * alpha test (alpha >= 1/255) at some pixel centre — the pairs the compositing kernels would stage and then skip
 * (35-49 % of the bounding-box pairs; validated on B200 in round 2, tests/test_gpu_prune_lists.py).
 * The reference's lists (fsb_isect_count / fsb_isect_emit: gsplat isect_tiles, dn_splatter/dn_model.py:570-591) stay
 * the bit-exact contract; these serve callers that do not expose the lists (the fused DN-Splatter step).
 *   conics[C*N,3], opacities[C*N]: what the compositing kernels receive.
 *   legacy_bbox: 0 / 1 as in fsb_isect_count; 2 = UNION list for fsb_raster_dn_*: the 0.1.x box (a superset of the 1.0
 *   box) with FSB_LEGACY_FLAG set in flatten_ids for the tiles that only the 0.1.x rule of rasterize_gaussians yields
 *   (dn_model.py:644-653), so ONE sorted list serves rasterization() and the legacy normals pass.
 *   emit: offsets = exclusive scan of the reach counts; static-capacity arguments as in fsb_isect_emit.
 *   hit_masks (nullable, u64[C*N]): scratch; the count pass records the reached tiles of every narrow box there and the
 *   emit pass called with the same arguments reads them instead of repeating the tests. */
#define FSB_LEGACY_FLAG 0x80000000u
int fsb_isect_count_reach(int C, int N, const float* means2d, const int32_t* radii, const float* conics,
                          const float* opacities, int tile_size, int tile_w, int tile_h, int legacy_bbox,
                          int32_t* counts, uint64_t* hit_masks, const float* depths, uint64_t* depth_keys,
                          int32_t* depth_vals, void* stream);
int fsb_isect_emit_reach(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                         const float* conics, const float* opacities, const int64_t* offsets, int tile_size,
                         int tile_w, int tile_h, int tile_bits, int legacy_bbox, const int64_t* n_dev,
                         int64_t capacity, int32_t* overflow_flag, int64_t* isect_ids, int32_t* flatten_ids,
                         const uint64_t* hit_masks, const int32_t* perm, int packed, void* stream);
/* Two-level binning (the default of the pruned path): the reference sorts 64-bit (camera | tile | depth bits) keys; the
 * same order comes from (1) a sort of the C*N Gaussians by depth bits (fsb_isect_count_reach's depth_keys / depth_vals
 * through fsb_radix_sort_pairs, end_bit 32), (2) emission in that order (fsb_isect_scan_perm + fsb_isect_emit_reach
 * with perm, packed = 1) and (3) a stable sort of the entries on the (camera, tile) bits only
 * (fsb_radix_sort_keys, begin_bit 32): 2 passes over 8-byte entries instead of 5 over 12-byte pairs at 1080p. */

/* exclusive int64 prefix sum of counts[M] (replaces torch.cumsum inside gsplat isect_tiles);
 * total_dev receives the grand total (= n_isects), a device scalar the caller copies back. */
size_t fsb_isect_scan_workspace(int64_t M);
int fsb_isect_scan_perm(int64_t M, const int32_t* counts, const int32_t* perm, int64_t* offsets, int64_t* total_dev,
                        void* workspace, size_t workspace_bytes, void* stream);
int fsb_isect_scan(int64_t M, const int32_t* counts, int64_t* offsets, int64_t* total_dev, void* workspace,
                   size_t workspace_bytes, void* stream);

/* STATIC-CAPACITY MODE (every function below that takes an `n_dev` / `n_isects_dev` pointer).
 * gsplat reads the intersection count back to the host between the count and the emit stage (one stream sync per
 * render).  When the pointer is non-NULL the host passes the CAPACITY of the list buffers as the count and a
 * device pointer to the true count (the total written by fsb_isect_scan); kernels work on min(true, capacity),
 * grids are sized for the capacity, and the whole render -> backward -> Adam step becomes a fixed launch sequence
 * that a CUDA graph can capture.  fsb_isect_emit raises *overflow_flag when the true count exceeds the capacity;
 * fsb_adam_multi_dev / fsb_densify_stats then leave the state untouched, so an overflowed step is a no-op that
 * the host re-runs with a larger capacity. */

/* I1 (emit): isect_ids[n_isects] = (cam << (32+tile_bits)) | (tile << 32) | float_bits(depth),
 * flatten_ids[n_isects] = cam*N + n.  replaces gsplat isect_tiles pass 2 / map_gaussian_to_intersects.
 * n_dev nullable; with it, `capacity` entries are the most that is written and overflow_flag (nullable, i32,
 * never cleared here) is set when *n_dev > capacity. */
int fsb_isect_emit(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                   const int64_t* offsets, int tile_size, int tile_w, int tile_h, int tile_bits,
                   int legacy_bbox, const int64_t* n_dev, int64_t capacity, int32_t* overflow_flag,
                   int64_t* isect_ids, int32_t* flatten_ids, void* stream);

/* I2: stable LSD radix sort of (u64 key, i32 value) pairs on key bits [0, end_bit).
 * replaces cub::DeviceRadixSort::SortPairs in gsplat isect_tiles and torch.sort in the legacy path.
 * Buffers A (input, clobbered) and B ping-pong; *result_in_b (host int) = 1 if B holds the result. */
size_t fsb_radix_sort_workspace(int64_t n, int end_bit);
int fsb_radix_sort_pairs(int64_t n, const int64_t* n_dev, int end_bit, uint64_t* keys_a, int32_t* vals_a,
                         uint64_t* keys_b, int32_t* vals_b, void* workspace, size_t workspace_bytes,
                         int* result_in_b, void* stream);
/* keys only, on key bits [begin_bit, end_bit); low32_out (nullable, i32[n]) = low word of each key at its sorted
 * position, written by the last pass. */
size_t fsb_radix_sort_keys_workspace(int64_t n, int begin_bit, int end_bit);
int fsb_radix_sort_keys(int64_t n, const int64_t* n_dev, int begin_bit, int end_bit, uint64_t* keys_a,
                        uint64_t* keys_b, int32_t* low32_out, void* workspace, size_t workspace_bytes,
                        int* result_in_b, void* stream);

/* I3: offsets[C, n_tiles] i32 = first sorted position of each (camera, tile).
 * replaces gsplat isect_offset_encode / legacy get_tile_bin_edges. */
int fsb_isect_offsets(int64_t n_isects, const int64_t* n_dev, const int64_t* sorted_ids, int C, int n_tiles,
                      int tile_bits, int32_t* offsets, void* stream);

/* Static-capacity mode, legacy normals pass (dn_model.py:644-653) right after rasterization() (:570-591) on the same
 * Gaussians: decide ON THE DEVICE whether the 0.1.x binning equals the 1.0 binning (equal intersection totals; the
 * legacy bbox is a superset per Gaussian) and, if so, reuse the first pass's sorted lists instead of emitting and
 * sorting again.  gate[0] = 0 (shared) or *n_legacy; pass `gate` as n_dev to emit / sort / offsets (no-ops when 0),
 * then fsb_isect_share_copy fills the legacy buffers from the first pass's.  The eager path takes the same decision
 * on the host (one D2H read). */
int fsb_isect_share_gate(const int64_t* n_first, const int64_t* n_legacy, int64_t* gate, void* stream);
int fsb_isect_share_copy(const int64_t* gate, const int64_t* n_list, int64_t capacity, const int32_t* src_flat,
                         const int32_t* src_offsets, int64_t n_offsets, int32_t* dst_flat, int32_t* dst_offsets,
                         void* stream);

/* R1: tile compositing forward.  replaces gsplat rasterize_to_pixels fwd / legacy rasterize_forward.
 *   means2d[C*N,2] conics[C*N,3] colors[C*N,D] opacities[C*N] ; backgrounds[C,D] nullable ;
 *   masks[C*tiles] u8 nullable ; ed_normalize: divide channel D-1 by max(alpha,1e-10) ("ED" modes).
 *   out_colors[C,H,W,D] out_alphas[C,H,W] last_ids[C,H,W] i32.
 * The forward first packs the sorted list into contiguous per-entry records inside `workspace`
 * (fsb_raster_workspace() bytes, n_tiles = C*tile_w*tile_h) that the compositing kernels stage with bulk async
 * copies; it also leaves the per-unit state there that the backward consumes (csrc/raster.cu).
 * D must be one of fsb_raster_supported_channels(); tile_size 8 or 16. */
int fsb_raster_supported_channels(int D);
size_t fsb_raster_workspace(int64_t n_isects, int64_t n_tiles, int64_t n_gauss, int D);  /* n_gauss = C*N */
int fsb_raster_fwd(int C, int N, int D, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d, const float* conics,
                   const float* colors, const float* opacities, const float* backgrounds, const uint8_t* masks,
                   int width, int height, int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets,
                   const int32_t* flatten_ids, int ed_normalize, void* workspace, size_t workspace_bytes,
                   float* out_colors, float* out_alphas, int32_t* last_ids, void* stream);

/* R2: tile compositing backward.  Gradient outputs are ACCUMULATED; the caller zero-fills them.
 *   workspace: the buffer filled by the matching fsb_raster_fwd call.
 *   v_means2d_abs nullable (absgrad=True in dn_splatter/dn_model.py:587); v_means2d nullable (2-D means detached,
 *   dn_model.py:638). */
int fsb_raster_bwd(int C, int N, int D, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d, const float* conics,
                   const float* colors, const float* opacities, const float* backgrounds, const uint8_t* masks,
                   int width, int height, int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets,
                   const int32_t* flatten_ids, int ed_normalize, void* workspace, size_t workspace_bytes,
                   const float* render_colors, const float* render_alphas, const int32_t* last_ids,
                   const float* v_render_colors, const float* v_render_alphas, float* v_means2d_abs,
                   float* v_means2d, float* v_conics, float* v_colors, float* v_opacities, void* stream);

/* R1 + R2 for the two compositing passes of one DN-Splatter iteration in ONE walk: colour set A (DA channels:
 * rasterization() RGB + expected depth, dn_splatter/dn_model.py:570-591) and colour set B (DB channels: the
 * per-Gaussian normals of the legacy gsplat.rasterize_gaussians pass, dn_model.py:644-653) share means2d / conics /
 * opacities, hence every alpha.  (DA, DB) must be (D, 0) with D from fsb_raster_supported_channels() — then these are
 * fsb_raster_fwd / fsb_raster_bwd — or (4, 3).
 *   flatten_ids may carry FSB_LEGACY_FLAG (fsb_isect_emit_reach legacy_bbox=2): such an entry exists only in the
 *   0.1.x list, i.e. it is composited into set B only; set A, out_alphas and last_ids follow the unflagged entries.
 *   ed_channel: channel of set A divided by max(alpha, 1e-10) on output (-1: none).
 *   backward: the 2-D mean gradients take set A's dL/dalpha only (dn_model.py:638 detaches the means of the legacy
 *   pass); conics / opacities receive both.  Gradient outputs are ACCUMULATED; the caller zero-fills them. */
size_t fsb_raster_dn_workspace(int64_t n_isects, int64_t n_tiles, int64_t n_gauss, int DA, int DB);
int fsb_raster_dn_fwd(int C, int N, int DA, int DB, int64_t n_isects, const int64_t* n_isects_dev,
                      const float* means2d, const float* conics, const float* colors_a, const float* colors_b,
                      const float* opacities, const float* backgrounds_a, const float* backgrounds_b,
                      const uint8_t* masks, int width, int height, int tile_size, int tile_w, int tile_h,
                      const int32_t* tile_offsets, const int32_t* flatten_ids, int ed_channel, void* workspace,
                      size_t workspace_bytes, float* out_a, float* out_b, float* out_alphas, int32_t* last_ids,
                      void* stream);
int fsb_raster_dn_bwd(int C, int N, int DA, int DB, int64_t n_isects, const int64_t* n_isects_dev,
                      const float* backgrounds_a, const float* backgrounds_b, const uint8_t* masks, int width,
                      int height, int tile_size, int tile_w, int tile_h, const int32_t* tile_offsets,
                      int ed_channel, void* workspace, size_t workspace_bytes, const float* render_a,
                      const float* render_alphas, const int32_t* last_ids, const float* v_render_a,
                      const float* v_render_b, const float* v_render_alphas, float* v_means2d_abs,
                      float* v_means2d, float* v_conics, float* v_colors_a, float* v_colors_b,
                      float* v_opacities, void* stream);

/* Measurement aid for bench.py's FP32 roofline (SURVEY.md §8d "Q = pair count"): counts[2] u64 (device, zeroed by
 * the caller) += { (pixel, entry) pairs the finished forward blended, pairs a per-pixel list walk visits }.
 * Nothing comparable exists in the reference; not on the training path. */
int fsb_raster_pair_count(int C, int N, int64_t n_isects, const int64_t* n_isects_dev, const float* means2d,
                          const float* conics, const float* opacities, int width, int height, int tile_size,
                          int tile_w, int tile_h, const int32_t* tile_offsets, const int32_t* flatten_ids,
                          const int32_t* last_ids, uint64_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-step image metrics.  replaces, in dn_splatter/dn_model.py:962-1000 (get_metrics_dict, run
 * every training iteration), torchmetrics PSNR + MSELoss and dn_splatter/metrics.py:111-150 DepthMetrics: one launch,
 * results stay on the device (the reference does eleven float() host syncs per step here).
 *   out[10] fp32 = { rgb_mse, rgb_psnr, depth_abs_rel, depth_sq_rel, depth_rmse, depth_rmse_log, a1, a2, a3, n_valid } */
size_t fsb_image_metrics_workspace(void);
int fsb_image_metrics(int H, int W, const float* pred_rgb, const float* gt_rgb, const float* pred_depth,
                      const float* gt_depth, float depth_tol, void* workspace, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU gradient exchange over NVLink peer memory (csrc/grad_exchange.cu): what DDP's all-reduce would do for
 * dn_splatter/dn_pipeline.py:161-167, as kernels of this library that a CUDA graph can capture — no NCCL call on the
 * step's path.  One process per GPU; rank r owns slice r = [r S, (r + 1) S) of the flat gradient.  "peer mapping" =
 * a device pointer into another rank's buffer, valid in this process (torch symmetric memory / CUDA IPC).
 *   pack            the step's gradient tensors -> this rank's flat buffer (offsets multiples of 4, gaps zeroed)
 *   barrier         all ranks arrived; ORs an int32 flag over the ranks (the static-capacity overflow flag)
 *   reduce_scatter  out[i] = sum_w grads[w][rank S + i]  (peer loads, or NVSwitch multimem.ld_reduce via grads_mc)
 *   adam_multi_xchg fsb_adam_multi_dev whose gradient load gathers from the ranks' reduced slices (peer loads)
 * Every pointer-array argument is a HOST array; signal pads and epoch counters are zero-initialised once. */
int fsb_xchg_max_world(void);
int fsb_xchg_barrier(int world, int rank, uint32_t* const* pads, int slot, uint32_t* epoch_dev, int32_t* flag,
                     void* stream);
int fsb_xchg_pack(int n_tensors, const float* const* src, const int64_t* n, const int64_t* off, float* dst,
                  int64_t total, void* stream);
int fsb_xchg_reduce_scatter(int world, int rank, const float* const* grads, const float* grads_mc, int64_t S,
                            float* out, void* stream);
/* in-place all-reduce of the flat gradient buffers: this rank reduces slice [rank S, (rank+1) S) and writes the sum
 * into every rank's buffer (NVSwitch multimem.ld_reduce + multimem.st when grads_mc is given, else peer loads and
 * stores).  Between two fsb_xchg_barrier's; afterwards the optimizer reads its own buffer. */
int fsb_xchg_allreduce(int world, int rank, float* const* grads, float* grads_mc, int64_t S, void* stream);
int fsb_adam_multi_xchg(int n_tensors, float* const* p, float* const* m, float* const* v, const int64_t* n,
                        const int64_t* off, int world, const float* const* reduced, int64_t S,
                        const float* hyper_dev, int hyper_stride, const int32_t* skip_flag, double beta1,
                        double beta2, double eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Visual hull (voxel carving).  replaces utils/VisualHull.py:149-191 (projection/vote loop, threshold
 * mask, occupied-voxel extraction) with InitializeVoxels (:15-57) folded into axis-table lookups.
 * Voxel l = (iz*nx + ix)*ny + iy, zs given in the reference's loop order (descending).
 *   masks[n_views,H,W] u8 (device), mats_host[n_views,12] f64 (HOST: K @ [R|t] rows), lut[256] f64 (device,
 *   value/255 as numpy computes it), xs/ys/zs f64 axis tables (device).
 *   votes[nz*nx*ny]: f64, or u8 when votes_u8 != 0 (binary 0 / 255 masks: a vote is an exact integer <= n_views, one
 *   byte per voxel instead of eight) ; max_bits: device u64, zero-filled by the caller, gets bits of max(votes) as f64. */
int fsb_vh_max_views(void);
int fsb_vh_count_block(void);
int fsb_vh_votes(int n_views, int H, int W, const uint8_t* masks, const double* mats_host, const double* lut,
                 const double* xs, int nx, const double* ys, int ny, const double* zs, int nz, void* votes,
                 int votes_u8, uint64_t* max_bits, void* stream);
/* block_counts[ceil(V / fsb_vh_count_block())] i32 = voxels with votes > iso per block of consecutive voxels */
int fsb_vh_count(int64_t V, const void* votes, int votes_u8, double iso, int32_t* block_counts, void* stream);
/* order-preserving compaction: points[n_occ,3] f64 (x,y,z) and indices[n_occ] i64 (nullable) in voxel order;
 * block_offsets = exclusive scan of block_counts (fsb_isect_scan). */
int fsb_vh_compact(int64_t V, const void* votes, int votes_u8, double iso, const int64_t* block_offsets,
                   const double* xs, int nx, const double* ys, int ny, const double* zs, double* points,
                   int64_t* indices, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a15: densify / prune bookkeeping.  replaces dn_splatter/dn_model.py:326-451 (refinement_after) and the
 * nerfstudio 1.1.3 helpers it drives (split_gaussians, dup_gaussians, cull_gaussians, dup_in_optim,
 * remove_from_optim).  The concatenated candidate model has M = N + samps*n_split + n_dup rows:
 * the N originals, the split children sample-major (.repeat(samps, 1)), the duplicates.
 *   classify : action[N] u8 (bit0 split, bit1 dup) + 0/1 i32 flags for fsb_isect_scan.  scales = LOG scales.
 *              max_dim = max(H, W); split_screen_size <= 0 disables that test; max_2Dsize / add_mask nullable.
 *   index    : split_idcs[n_split], dup_idcs[n_dup] i32 from the exclusive scans of the flags.
 *   keep     : parent[M] i32, keep[M] i32 = !cull_gaussians(...) evaluated on the candidate row; action nullable
 *              (cull only); extra_cull[N] u8 nullable; cull_scale_thresh <= 0 / cull_screen_size <= 0 disable
 *              the too-big tests.
 *   gather   : out[offsets[j], :] = src[parent[j], :] for kept j (offsets = exclusive scan of keep);
 *              zero_new = 1 writes zeros for j >= N (Adam moments of new rows).
 *   split_fixup: over the n_new = samps*n_split + n_dup new rows: kept children get mean = mu + R(q/|q|)(exp(s)*eps)
 *              and scale = log(exp(s)/size_fac); kept duplicates of split parents the shrunk scale (the reference
 *              shrinks split parents in place before it evaluates `dups`).  samples[samps*n_split, 3] ~ N(0, I)
 *              supplied by the caller. */
int fsb_refine_classify(int N, const float* xys_grad_norm, const float* vis_counts, const float* max_2Dsize,
                        const float* scales, float max_dim, float densify_grad_thresh, float densify_size_thresh,
                        float split_screen_size, float size_fac, const uint8_t* add_mask, uint8_t* action,
                        int32_t* split_flag, int32_t* dup_flag, void* stream);
int fsb_refine_index(int N, const uint8_t* action, const int64_t* split_rank, const int64_t* dup_rank,
                     int32_t* split_idcs, int32_t* dup_idcs, void* stream);
int fsb_refine_keep(int64_t M, int N, int n_split, int n_dup, int samps, const uint8_t* action,
                    const int32_t* split_idcs, const int32_t* dup_idcs, const float* opacities,
                    const float* scales, const float* max_2Dsize, float cull_alpha_thresh, float cull_scale_thresh,
                    float cull_screen_size, float size_fac, const uint8_t* extra_cull, int32_t* parent,
                    int32_t* keep, void* stream);
int fsb_refine_gather(int64_t M, int N, int width, const float* src, const int32_t* parent, const int32_t* keep,
                      const int64_t* offsets, int zero_new, float* out, void* stream);
int fsb_refine_split_fixup(int64_t n_new, int64_t n_children, int N, const uint8_t* action, const float* means,
                           const float* scales, const float* quats, const float* samples, const int32_t* parent,
                           const int32_t* keep, const int64_t* offsets, float size_fac, float* out_means,
                           float* out_scales, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a16: visual-hull pruning without the distance matrix.  replaces torch.cdist(means[close], hull).min(-1) and
 * the threshold mask of dn_splatter/dn_model.py:1254-1269 (every refine_every steps).
 *   fsb_hull_min_dist: pts[N,3], hull[V,3], center[3] (device, nullable: no radius filter), r_close;
 *     min_dist[N] = distance to the nearest hull point for points within r_close of center, +inf otherwise.
 *     stop_below > 0 lets a point stop once a hull point is closer than that (result then an upper bound).
 *   fsb_hull_prune_mask: mask[N] u8 = (lo < min_dist <= hi) && !protect[n]  (protect u8, nullable). */
int fsb_hull_min_dist(int N, const float* pts, int V, const float* hull, const float* center, float r_close,
                      float stop_below, float* min_dist, void* stream);
int fsb_hull_prune_mask(int N, const float* min_dist, float lo, float hi, const uint8_t* protect, uint8_t* mask,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * SSIM of the base splatfacto loss: `1 - self.ssim(gt, pred)` with torchmetrics
 * StructuralSimilarityIndexMeasure(data_range=1.0, kernel_size=11, sigma=1.5) — dn_splatter/dn_model.py:244,
 * reached through super().get_loss_dict at dn_model.py:683.  Mean over the (H-10)x(W-10) interior (torchmetrics
 * crops the reflect-padded border) and the channels.  x, y: [H,W,C] fp32, C <= 4; taps: HOST array of
 * fsb_ssim_taps() normalised gaussian weights; workspace: fsb_ssim_workspace() bytes.
 * fwd: ssim_out = device scalar; d_mu/d_xx/d_xy [(H-10),(W-10),C] (nullable, all or none) = partials w.r.t. the
 * window moments of x, consumed by bwd.  bwd: v_out = DEVICE scalar dL/dssim, v_x[H,W,C] overwritten. */
int fsb_ssim_taps(void);
size_t fsb_ssim_workspace(void);
int fsb_ssim_fwd(int H, int W, int C, const float* x, const float* y, float data_range, float k1, float k2,
                 const float* taps, void* workspace, float* ssim_out, float* d_mu, float* d_xx, float* d_xy,
                 void* stream);
int fsb_ssim_bwd(int H, int W, int C, const float* x, const float* y, float data_range, float k1, float k2,
                 const float* taps, const float* d_mu, const float* d_xx, const float* d_xy, const float* v_out,
                 float* v_x, void* stream);

/* ------------------------------------------------------------------------------------------------
 * O1: multi-tensor Adam, one launch for all Gaussian parameter groups.  replaces the torch.optim.Adam.step()
 * calls for the optimizers of dn_splatter/dn_config.py:36-75 (eps 1e-15, betas (0.9, 0.999), no decay).
 * Every array argument is a HOST array of length n_tensors (<= fsb_adam_max_tensors()); p/g/m/v hold device
 * pointers to fp32 tensors of n[i] elements; step[i] = 1-based step count of tensor i for this update. */
int fsb_adam_max_tensors(void);
int fsb_adam_multi(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                   const int64_t* n, const double* lr, const int64_t* step, double beta1, double beta2, double eps,
                   void* stream);
/* Same update with the per-tensor scalars read from DEVICE memory, for a CUDA-graph-captured step whose launch
 * arguments are frozen: hyper_dev[2 * fsb_adam_max_tensors()] fp32 = step_size[i] = lr_i / (1 - beta1^t_i) at
 * [i] and sqrt(1 - beta2^t_i) at [fsb_adam_max_tensors() + i], refreshed by the host before every replay
 * (fsb_upload_small).  skip_flag (nullable, device i32): non-zero -> nothing is updated (overflowed step). */
int fsb_adam_multi_dev(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                       const int64_t* n, const float* hyper_dev, const int32_t* skip_flag, double beta1,
                       double beta2, double eps, void* stream);

/* Stream-ordered upload of up to fsb_upload_small_max() bytes from HOST memory to device memory: the bytes
 * travel as kernel arguments, so the host buffer may be reused as soon as the call returns and nothing
 * synchronises (per-replay scalars of a captured step: camera index, Adam step sizes). */
int fsb_upload_small_max(void);
int fsb_upload_small(void* dst_dev, const void* src_host, int bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DN-Splatter regulariser, fused.  replaces dn_splatter/dn_model.py:722-736 (EdgeAwareLogL1 on sensor depth,
 * losses.py:177-214), :753-756 (TV on depth, losses.py:269-285), :806 (normal L1), :814-815 (normal TV) and the
 * RGB L1 term of the base splatfacto loss:
 *   loss = l_sensor*EALogL1(depth, sensor | sensor > depth_tol, edges of max(edge_rgb, rgb_clamp_min))
 *        + l_smooth*TV(depth) + l_nl1*mean|gt_normal - pred_normal| + l_ntv*TV(pred_normal)
 *        + l_rgb*mean|gt_rgb - pred_rgb|
 * Images are [H,W] / [H,W,3] fp32; a term with weight 0 may have NULL inputs.  loss_out: device scalar.
 * workspace: fsb_dn_loss_workspace() bytes, filled by fwd and read by bwd (mask counts).
 * bwd: v_loss is a DEVICE scalar; v_depth / v_normal / v_rgb are nullable and overwritten. */
size_t fsb_dn_loss_workspace(void);
int fsb_dn_loss_fwd(int H, int W, const float* depth, const float* sensor, const float* edge_rgb,
                    const float* pred_normal, const float* gt_normal, const float* pred_rgb, const float* gt_rgb,
                    float depth_tol, float rgb_clamp_min, float l_sensor, float l_smooth, float l_nl1, float l_ntv,
                    float l_rgb, void* workspace, float* loss_out, void* stream);
int fsb_dn_loss_bwd(int H, int W, const float* depth, const float* sensor, const float* edge_rgb,
                    const float* pred_normal, const float* gt_normal, const float* pred_rgb, const float* gt_rgb,
                    float depth_tol, float rgb_clamp_min, float l_sensor, float l_smooth, float l_nl1, float l_ntv,
                    float l_rgb, const void* workspace, const float* v_loss, float* v_depth, float* v_normal,
                    float* v_rgb, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Image-space glue of DNSplatterModel.get_outputs, fused (P = H*W pixels of ONE camera).
 * fsb_compose_rgbd_*: replaces dn_splatter/dn_model.py:602-604 and :609-613
 *   rgb[P,3] = clamp(render[P,0:3] + (1 - alpha[P]) * background[3], 0, 1)
 *   depth[P] = alpha > 0 ? render[P,3] : max_p render[p,3]     (the maximum is detached, as in the reference)
 *   render[P,4] must be 16-byte aligned; background is a DEVICE pointer; scratch: 4 bytes.
 *   bwd: v_rgb[P,3] / v_depth[P] (nullable = zero) -> v_render[P,4], v_alpha[P] (overwritten).
 * fsb_normal_map_*: replaces dn_model.py:655-656, out[P,3] = (n / ||n|| + 1) / 2 (no epsilon, as there). */
int fsb_compose_rgbd_fwd(int64_t P, const float* render, const float* alpha, const float* background, float* rgb,
                         float* depth, float* scratch, void* stream);
int fsb_compose_rgbd_bwd(int64_t P, const float* render, const float* alpha, const float* background,
                         const float* v_rgb, const float* v_depth, float* v_render, float* v_alpha, void* stream);
int fsb_normal_map_fwd(int64_t P, const float* normals_raw, float* out, void* stream);
int fsb_normal_map_bwd(int64_t P, const float* normals_raw, const float* v_out, float* v_normals_raw, void* stream);
/* 8-bit image targets -> float32 on the device, with the reference's bits: recip = 1: src * (1.0f / 255.0f) — what
 * splatfacto get_gt_img's `image.float() / 255.0` evaluates to on the device (torch's CUDA division by a Python scalar
 * multiplies by the fp32 reciprocal); recip = 0: IEEE src / 255.0f — numpy's division in dn_dataset.py:205. */
int fsb_u8_to_unit_float(int64_t n, const uint8_t* src, float* dst, int recip, void* stream);

/* Flatness regulariser ("two_d_gaussians").  replaces dn_splatter/dn_model.py:817-819:
 *   out = mean_i min_k exp(log_scales[i,k])     (device scalar; workspace: 16 bytes)
 * bwd: v_loss is a DEVICE scalar; v_log_scales[N,3] overwritten (non-zero on the arg-min axis only). */
int fsb_flatness_fwd(int N, const float* log_scales, void* workspace, float* out, void* stream);
int fsb_flatness_bwd(int N, const float* log_scales, const float* v_loss, float* v_log_scales, void* stream);

/* main_loss assembly.  replaces the scalar torch launches of dn_splatter/dn_model.py:683-690 and :925:
 *   out = w_ssim * (1 - *ssim) + *reg + w_flat * *flat     (device scalars; a NULL term is left out)
 * bwd: v_out is a DEVICE scalar; v_ssim = -w_ssim v, v_reg = v, v_flat = w_flat v (nullable, overwritten). */
int fsb_loss_combine_fwd(const float* ssim, const float* reg, const float* flat, float w_ssim, float w_flat,
                         float* out, void* stream);
int fsb_loss_combine_bwd(const float* v_out, float w_ssim, float w_flat, float* v_ssim, float* v_reg, float* v_flat,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-Gaussian normals.  replaces dn_splatter/dn_model.py:617-636: one_hot(argmin(scales)) -> R(q) column ->
 * normalize -> flip towards the camera -> rotate into the camera frame (normals @ c2w[:3,:3]).
 *   quats[N,4] wxyz (any norm), scales[N,3] (log or linear: only the argmin matters), means[N,3],
 *   c2w: 3 rows of 4 floats (camera-to-world, row-major).  normals_world[N,3] nullable, normals_cam[N,3].
 * bwd: v_normals_cam[N,3] -> v_quats[N,4] (overwritten); nothing else is differentiable in the reference. */
int fsb_gaussian_normals_fwd(int N, const float* quats, const float* scales, const float* means, const float* c2w,
                             float* normals_world, float* normals_cam, void* stream);
int fsb_gaussian_normals_bwd(int N, const float* quats, const float* scales, const float* means, const float* c2w,
                             const float* v_normals_cam, float* v_quats, void* stream);

/* Per-step densification statistics.  replaces nerfstudio splatfacto after_train, whose outputs
 * dn_splatter/dn_model.py:326-451 (refinement_after) consumes.  For radii[n] > 0:
 *   vis_counts[n] += 1 ; xys_grad_norm[n] += |grads2d[n]|_2 ; max_2Dsize[n] = max(., radii[n] / max_dim)
 * skip_flag (nullable, device i32): non-zero -> no-op (static-capacity mode, overflowed step). */
int fsb_densify_stats(int N, const int32_t* radii, const float* grads2d, float max_dim, float* xys_grad_norm,
                      float* vis_counts, float* max_2Dsize, const int32_t* skip_flag, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a13: pseudo ground-truth normals from a depth image.  replaces normal_from_depth_image
 * (dn_splatter/utils/normal_utils.py:23-46 = get_means3d_backproj, utils/camera_utils.py:92-144, then
 * pcd_to_normal, normal_utils.py:7-20), called from dn_splatter/dn_model.py:779-789 when
 * normal_supervision == "depth".
 *   depth[H,W]: back-projected at pixel centres (u + 0.5 - cx) * d / fx; rot_inv: HOST 3x3 row-major
 *   inverse of c2w[:3,:3] (nullable = identity), trans: HOST c2w[:3,3] (nullable = 0): world = p @ rot_inv + trans.
 *   xyz[H,W,3] (nullable): when given, the points are taken from it instead (pcd_to_normal alone).
 *   normals[H,W,3] = normalize((right - left) x (top - bottom)), the one-pixel border is zero. */
int fsb_normal_from_depth(int H, int W, const float* depth, const float* xyz, float fx, float fy, float cx,
                          float cy, const float* rot_inv, const float* trans, float* normals, void* stream);

/* ------------------------------------------------------------------------------------------------
 * f2: exact k nearest neighbours of 3-D points and the local density field of the mesh-export /
 * SDF path.  replaces knn_sk (dn_splatter/utils/knn.py:29-43: sklearn NearestNeighbors(k + 1) on the CPU, first
 * neighbour dropped; call sites dn_splatter/dn_model.py:183-189, :306-310, :1562-1572) and get_density
 * (dn_model.py:1575-1635 with scale_rot_to_inv_cov3d, :2141-2150).
 *   fsb_knn_axis_keys : keys[3 S] u64 = (axis << 32) | order-preserving bits of coordinate `axis` of the S sampled
 *                    points (point i * stride); sort them with fsb_radix_sort_keys on bits [0, 34)
 *   fsb_knn_edges  : edges[3][g + 1] f32 = the i/g quantiles of every axis: a g^3 grid that is fine where the points are
 *   fsb_knn_cells  : keys[i] = linear cell ((cz g + cy) g + cx) of point i (border cells open-ended, non-finite points
 *                    -> g^3 and counted into n_nonfinite, nullable), vals[i] = i; sort with fsb_radix_sort_pairs
 *   fsb_knn_build  : cell_start[g^3 + 1] i32 and sorted_pts[N,4] (xyz + index bits) from the sorted pairs
 *   fsb_knn_query  : the K <= 33 nearest points of every y, nearest first, ties by index (distances in fp64 of the fp32
 *                    coordinates); the first drop_first are not written.  out_idx [Ny, K - drop_first] i64, out_dist
 *                    (nullable) f64.  order (nullable) = processing order.  Queries needing more than max_steps growth
 *                    steps of their cell box are appended to unresolved[Ny] / n_unresolved[1] and finished by
 *   fsb_knn_brute  : K selection rounds over all points, one CTA per unresolved query.
 *   fsb_gaussian_density : out[s] = max(norm(sum_k sigmoid(opacity[g]) exp(-0.5 |M_g^T (x_s - mu_g)|^2)), clamp_min = 1e-4),
 *                    g = knn[s,k], M = R(quat) diag(1 / clamp_min(exp(log_scales), 1e-3)), norm(d) = d / (d + 1e-5) if d >= 1 */
int fsb_knn_axis_keys(int64_t S, int64_t stride, const float* pts, uint64_t* keys, void* stream);
int fsb_knn_edges(int64_t S, const uint64_t* sorted_keys, int g, float* edges, void* stream);
int fsb_knn_cells(int64_t N, const float* pts, int g, const float* edges, uint64_t* keys, int32_t* vals,
                  int32_t* n_nonfinite, void* stream);
int fsb_knn_build(int64_t N, const uint64_t* sorted_keys, const int32_t* sorted_vals, const float* pts, int g,
                  int32_t* cell_start, float* sorted_pts, void* stream);
int fsb_knn_query(int64_t Ny, const float* y, const int32_t* order, int g, const float* edges,
                  const int32_t* cell_start, const float* sorted_pts, int K, int drop_first, int max_steps,
                  int64_t* out_idx, double* out_dist, int32_t* unresolved, int32_t* n_unresolved, void* stream);
int fsb_knn_brute(int64_t Nx, const float* x, const float* y, const int32_t* unresolved, const int32_t* n_unresolved,
                  int K, int drop_first, int64_t* out_idx, double* out_dist, void* stream);
int fsb_gaussian_density(int64_t S, const float* samples, int K, const int64_t* knn, const float* means,
                         const float* log_scales, const float* quats, const float* opacity_logits, float clamp_min,
                         float* out, void* stream);
/* Level-set search along the camera rays of P back-projected points.  replaces dn_splatter/dn_model.py:1766-1880 of
 * compute_level_surface_points: per point the first neighbour's standard deviation along its view direction, 21 samples
 * at lin[s] * std (lin = torch.linspace(-3, 3, 21), HOST) around the point, the density of the K neighbours knn[P,K] at
 * every sample (the inlined get_density, normalised above 1, not clamped) and, for each of n_levels <= 4 surface levels
 * (HOST), the first sample above the level with the linear interpolation of :1858-1880.  cam = camera position (HOST, 3).
 *   t_out[l,p] f32 = ray parameter of the crossing (0 where none), valid_out[l,p] u8; dens_out (nullable) [P,21],
 *   std_out (nullable) [P].  The reference materialises [P*21, K, 3, 3] temporaries in passes of 2M samples. */
int fsb_level_crossings(int64_t P, const float* points, const float* cam, int K, const int64_t* knn, const float* means,
                        const float* log_scales, const float* quats, const float* opacity_logits, const float* lin,
                        int n_levels, const float* levels, float* t_out, uint8_t* valid_out, float* dens_out,
                        float* std_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * f3: the seed point cloud of Module 1.  replaces get_pointcloud (utils/generate_pcd.py:15-48:
 * meshgrid, (x - CX) / FX, [3,3] @ [3,P] matmul, colour permute, two boolean-mask gathers per view) and open3d's
 * voxel_down_sample(0.02) on the CPU (utils/generate_pcd.py:98-101), i.e. what init_pcd_generate (scripts/train.py:95)
 * writes to merged_pcd.ply.
 *   fsb_backproject_flags : flags[i] i32 = lo < depth[i] < hi; scan them with fsb_isect_scan -> offsets, count
 *   fsb_backproject_emit  : out[offsets[i], 0:6] = (R p_cam + T, r, g, b) for every kept pixel, in pixel order;
 *                           c2w_rot (9, row-major) / c2w_trans (3) are HOST pointers, color_chw the [3,H,W] image
 *   fsb_voxel_keys        : open3d's voxel = floor((p - (min_bound - voxel / 2)) / voxel) in fp64 as a 63-bit key (21 bits
 *                           per axis, x high); min_bound[3] f64 (device) is computed here; *overflow (device i32, not
 *                           zeroed here) is set when an index leaves [0, 2^21); sort with fsb_radix_sort_pairs(63)
 *   fsb_voxel_heads       : heads[i] i32 = sorted position i starts a voxel; scan with fsb_isect_scan
 *   fsb_voxel_mean        : out[n_voxels, width] f64 = mean of each voxel's rows pts[., 0:width] (fp64 sums in input
 *                           order, like open3d's AccumulatedPoint), voxels in ascending key order */
int fsb_backproject_flags(int64_t P, const float* depth, float lo, float hi, int32_t* flags, void* stream);
int fsb_backproject_emit(int H, int W, const float* depth, const float* color_chw, const float* c2w_rot,
                         const float* c2w_trans, float fx, float fy, float cx, float cy, float lo, float hi,
                         const int64_t* offsets, float* out, void* stream);
size_t fsb_voxel_workspace(void);
int fsb_voxel_keys(int64_t N, const float* pts, int stride, double voxel, double* min_bound, uint64_t* keys,
                   int32_t* vals, int32_t* overflow, void* workspace, size_t workspace_bytes, void* stream);
int fsb_voxel_heads(int64_t N, const uint64_t* sorted_keys, int32_t* heads, void* stream);
int fsb_voxel_mean(int64_t N, const uint64_t* sorted_keys, const int32_t* sorted_vals, const int32_t* heads,
                   const int64_t* offsets, const float* pts, int stride, int width, double* out, void* stream);

/* library bookkeeping: kernels launched by libfsb200 since load (monotone), ABI revision */
uint64_t fsb_launch_count(void);
int fsb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_H */
