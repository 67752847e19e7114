/*
 * den_b200.h — C ABI of the B200-native Deblur e-NeRF renderer kernels.
 *
 * This is the drop-in boundary under the reference's Python operator API
 * ("C ABI under B1/B2").  Every entry point is `extern "C"`,
 * takes plain device pointers + sizes + a `cudaStream_t` (passed as `void*`),
 * never allocates, never synchronises, never touches the default stream unless
 * it is the stream handed in, and returns 0 or a negative `den_status`; the
 * message of the last failure on the calling thread is `den_last_error()`.
 * Buffers (outputs, workspaces, the sample arena) are owned by the caller — the
 * torch wrapper in `deblur-e-nerf_b200/` (loaded with ctypes, no torch types in
 * any signature).
 *
 * Each declaration cites the reference interface it replaces.  The reference
 * (wengflow/deblur-e-nerf) has no native code; its kernels live in two
 * un-vendored third-party packages (nerfacc==0.3.1, tiny-cuda-nn) reached
 * through the Python call sites cited below (paths relative to the reference
 * root, `deblur_e_nerf/` prefix omitted).
 *
 * Conventions: all floating-point data is fp32 unless the name says f64;
 * "samples" are ray-major, front-to-back; `offsets` is an (R+1) int32 exclusive
 * prefix of per-ray sample counts (samples of ray r are
 * [offsets[r], offsets[r+1])); matrices are row-major.
 */
#ifndef DEN_B200_H_
#define DEN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEN_ABI_VERSION 2
#define DEN_MAX_LEVELS 32

typedef enum den_status {
    DEN_OK = 0,
    DEN_ERR_INVALID_ARGUMENT = -1,
    DEN_ERR_CUDA = -2,
    DEN_ERR_UNSUPPORTED = -3,
    DEN_ERR_WORKSPACE = -4
} den_status;

/* nerfacc.ContractionType (models/deblur_e_nerf.py:271-275) */
typedef enum den_contraction {
    DEN_CONTRACT_AABB = 0,
    DEN_CONTRACT_TANH = 1,
    DEN_CONTRACT_SPHERE = 2
} den_contraction;

/* Library identity / error reporting. */
int den_version(void);
const char* den_last_error(void);
/* sm count of the current device (grid sizing), or a negative den_status */
int den_device_sm_count(void);

/* ------------------------------------------------------------------------- *
 * Multiresolution hash grid — replaces tinycudann.Encoding(HashGrid, Linear,
 * fp32) constructed at external/ngp.py:166-170 and called at ngp.py:240.
 * Level table per tcnn grid.h (grid_scale / grid_resolution / offset table),
 * evaluated once on the host (oracle/tcnn_ref.py: grid_level_table).
 * F (features per level) is 2.
 * ------------------------------------------------------------------------- */
typedef struct den_hashgrid_desc {
    int32_t n_levels;
    int32_t n_features;                 /* must be 2 */
    int32_t n_agg_levels;               /* bwd: levels [0, n_agg) use warp-aggregated atomics */
    int32_t reserved;
    float scale[DEN_MAX_LEVELS];        /* exp2(l*log2 s)*base - 1, fp32 */
    uint32_t resolution[DEN_MAX_LEVELS];/* ceil(scale) + 1 */
    uint32_t size[DEN_MAX_LEVELS];      /* entries in the level's table */
    uint32_t offset[DEN_MAX_LEVELS];    /* entry offset of the level */
} den_hashgrid_desc;

/* Device-side sample counts.  Every per-sample entry point takes `n_samples` (the number of rows
 * the host KNOWS the buffers hold — their capacity) and `n_samples_dev`: NULL, or a device int32 with
 * the true count written earlier in the stream (the march total, the survivor total of the
 * visibility filter); the kernel then works on min(n_samples, *n_samples_dev) rows.  With it a
 * training step has no host read-back (upstream reads both totals back: SURVEY.md App. C rows 1, 3,
 * 4) and can be captured in a CUDA graph. */
/* x (M,3) in unit-cube coordinates -> out (M, L*2).  tcnn kernel_grid. */
int den_hashgrid_fwd(const den_hashgrid_desc* desc, const float* x, const float* table,
                     float* out, int64_t n_samples, const int32_t* n_samples_dev, void* stream);
/* dtable (n_entries*2, pre-zeroed or accumulating) += scatter of dout (M, L*2).
 * tcnn kernel_grid_backward.  If dx != NULL also writes dL/dx (M,3)
 * (kernel_grid_backward_input; needs `table`). */
int den_hashgrid_bwd(const den_hashgrid_desc* desc, const float* x, const float* dout,
                     const float* table, float* dtable, float* dx, int64_t n_samples,
                     const int32_t* n_samples_dev, void* stream);

/* ------------------------------------------------------------------------- *
 * Ray marching — replaces nerfacc.ray_marching (external/utils.py:106-119)
 * and nerfacc.ray_aabb_intersect reached through it (models/nerf.py:248-251).
 * ------------------------------------------------------------------------- */
typedef struct den_march_params {
    float roi[6];                       /* grid.roi_aabb */
    int32_t res[3];                     /* grid.binary.shape */
    int32_t contraction;                /* den_contraction */
    float step_size;                    /* render_step_size */
    float cone_angle;
} den_march_params;

int den_ray_aabb_intersect(const float* rays_o, const float* rays_d, const float* aabb6_host,
                           float* t_min, float* t_max, int64_t n_rays, void* stream);
/* t_min = max(t_min, near); t_max = min(t_max, far); t_min += jitter*step (jitter may be NULL) */
int den_clamp_jitter(float* t_min, float* t_max, const float* jitter, int has_near, float near_plane,
                     int has_far, float far_plane, float step_size, int64_t n_rays, void* stream);
/* pass 1: num_steps[r] = samples ray r emits */
int den_march_count(const den_march_params* p, const float* rays_o, const float* rays_d,
                    const float* t_min, const float* t_max, const uint8_t* binary,
                    int32_t* num_steps, int64_t n_rays, void* stream);
/* out[0..n) = exclusive prefix of in, out[n] = total.  workspace >= den_scan_workspace_bytes(n) */
size_t den_scan_workspace_bytes(int64_t n);
int den_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* workspace,
                           size_t workspace_bytes, void* stream);
/* pass 2: writes ray_indices/t_starts/t_ends at offsets[r].. (clamped to capacity) */
int den_march_write(const den_march_params* p, const float* rays_o, const float* rays_d,
                    const float* t_min, const float* t_max, const uint8_t* binary,
                    const int32_t* offsets, int32_t* ray_indices, float* t_starts, float* t_ends,
                    int64_t n_rays, int64_t capacity, void* stream);

/* Single-pass alternative to count + write (the marching loop runs once): den_march_bound gives a
 * per-ray upper bound of the samples ray r can emit (<= max_per_ray); after an exclusive scan of the
 * bounds (seg_offsets, n_rays+1) den_march_single marches every ray once, writing t_starts/t_ends
 * into the ray's own segment of the arena and its true count into num_steps; after a scan of the
 * counts (offsets) den_march_pack copies the segments to their packed place and fills ray_indices.
 * A ray whose count exceeds its segment keeps counting but stops writing (the caller compares
 * num_steps with the bounds and falls back to den_march_write). */
int den_march_bound(const den_march_params* p, const float* t_min, const float* t_max,
                    int32_t max_per_ray, int32_t* bound, int64_t n_rays, void* stream);
int den_march_single(const den_march_params* p, const float* rays_o, const float* rays_d,
                     const float* t_min, const float* t_max, const uint8_t* binary,
                     const int32_t* seg_offsets, int32_t* num_steps, float* arena_t0,
                     float* arena_t1, int64_t n_rays, void* stream);
int den_march_pack(const int32_t* seg_offsets, const int32_t* offsets, const float* arena_t0,
                   const float* arena_t1, int64_t n_rays, int32_t* ray_indices, float* t_starts,
                   float* t_ends, void* stream);

/* ------------------------------------------------------------------------- *
 * Occupancy-grid update — replaces nerfacc.OccupancyGrid._update (grid.py)
 * as called through every_n_step at models/nerf.py:200-204, and the cone-aware
 * occ_eval_fn of NeRF.update_occ_grid (models/nerf.py:171-198).  The random draws (cell
 * indices, jitter, camera ids) are made by the caller in upstream order and passed in.
 * ------------------------------------------------------------------------- */
typedef struct den_occgrid_desc {
    float roi[6];                       /* grid.roi_aabb */
    int32_t res[3];                     /* grid.resolution */
    int32_t contraction;                /* den_contraction */
} den_occgrid_desc;

/* point i: cell = indices ? indices[i] : i (x-slowest flat index); unit = (coords + jitter[i]) /
 * res; world[i] = contract_inv(unit) (AABB: unit_to_roi; sphere: v = 4(unit - 0.5), n > 1:
 * v /= max(2n - n^2, 1e-10)); keep[i] (may be NULL) = sphere ? ||unit - 0.5|| < 0.5 : 1 */
int den_occgrid_cell_points(const den_occgrid_desc* g, const int64_t* indices, const float* jitter,
                            int64_t n, float* world, uint8_t* keep, void* stream);
/* occ[i] = sigma[i] * step_i; cone_angle > 0: t = ||camera_pos[camera_ids[i]] - world[i]||,
 * step_i = max(t * cone_angle, step_size), 0 outside (near, far) when has_planes; else step_size */
int den_occgrid_occ(const float* sigma, const float* world, const int64_t* camera_ids,
                    const float* camera_pos, float cone_angle, float step_size, int has_planes,
                    float near_plane, float far_plane, int64_t n, float* occ, void* stream);
/* occs[cell] = max(occs[cell] * ema_decay, occ) over the n points (duplicates: the largest
 * candidate wins — one of the outcomes of upstream's unordered indexed assignment; points with
 * keep[i] == 0 are skipped), then binary = occs > min(mean(occs), occ_thre) with the mean
 * accumulated in fp64 in a fixed order; mean_out (device, may be NULL) receives mean(occs).
 * workspace: den_occgrid_workspace_bytes(n_cells), initialised ONCE by den_occgrid_workspace_init
 * (the update leaves it ready for the next call). */
size_t den_occgrid_workspace_bytes(int64_t n_cells);
int den_occgrid_workspace_init(void* workspace, int64_t n_cells, void* stream);
int den_occgrid_ema_update(const int64_t* indices, const uint8_t* keep, const float* occ, int64_t n,
                           float ema_decay, float occ_thre, float* occs, int64_t n_cells,
                           uint8_t* binary, float* mean_out, void* workspace, void* stream);

/* Timestamps + pixels -> rays: LinearTrajectory.forward (models/trajectories.py:30-90: searchsorted,
 * f64 weight, LERP, shortest-path SLERP via utils/tensor_ops.py:118-184, quaternion -> R) fused with
 * NeRF.pixel_params_to_ray (models/nerf.py:206-228).  timestamps (n_rays) f64 ns; ray i looks through
 * pixel i % n_pixels of pixels (n_pixels, 2); pose_ts (C) int64 ascending, pose_pos (C,3), pose_quat
 * (C,4) xyzw; kinv9_host: row-major K^-1 on the HOST. */
int den_rays_from_trajectory(const double* timestamps, const float* pixels, int64_t n_pixels,
                             const int64_t* pose_ts, const float* pose_pos, const float* pose_quat,
                             int32_t n_poses, const float* kinv9_host, float* rays_o, float* rays_d,
                             int64_t n_rays, void* stream);
/* Reverse mode of the above with respect to the timestamps (the refractory-period path,
 * models/deblur_e_nerf.py:465-469: timestamps = event time - tau; the autograd of trajectories.py:30-90 +
 * utils/tensor_ops.py:118-184 + nerf.py:206-228, ~80 launches): d_timestamps[i] (f64, per ns) =
 * (d_rays_o[i] . (p_right - p_left) + d_rays_d[i] . ((R_left r) x rays_d[i])) / interval width, r = the
 * rotation vector of q_left^-1 q_right.  rays_d = the forward output; d_rays_o / d_rays_d may be NULL
 * (no gradient through that output). */
int den_rays_from_trajectory_bwd(const double* timestamps, const int64_t* pose_ts, const float* pose_pos,
                                 const float* pose_quat, int32_t n_poses, const float* rays_d,
                                 const float* d_rays_o, const float* d_rays_d, double* d_timestamps,
                                 int64_t n_rays, void* stream);

/* Visibility filter + compaction — replaces nerfacc.render_visibility and the
 * three boolean-mask compactions inside nerfacc.ray_marching. */
/* alpha = 1 - exp(-sigma * (t1 - t0)) */
int den_alpha_from_sigma(const float* sigmas, const float* t_starts, const float* t_ends,
                         float* alphas, int64_t n_samples, const int32_t* n_samples_dev, void* stream);
/* mask[i] = (T_i >= eps) && (alpha_thre <= 0 || alpha_i >= alpha_thre), T sequential fp32 per ray;
 * vis_count[r] = number of visible samples of ray r */
int den_visibility(const float* alphas, const int32_t* offsets, int64_t n_rays, float early_stop_eps,
                   float alpha_thre, uint8_t* mask, int32_t* vis_count, void* stream);
int den_compact_samples(const uint8_t* mask, const int32_t* offsets_in, const int32_t* offsets_out,
                        const int32_t* ray_indices_in, const float* t_starts_in, const float* t_ends_in,
                        int32_t* ray_indices_out, float* t_starts_out, float* t_ends_out,
                        int64_t n_rays, void* stream);
/* The same compaction, also carrying the pre-pass outputs of the survivors along: sigmas (M) and
 * rgbs (M, channels) are copied (8 B per sample), and src_rows[k] = index of survivor k in the
 * un-compacted arrays, which lets den_mlp_bwd read the 128-byte encodings of the pre-pass IN PLACE
 * instead of from a compacted copy.  Any of sig / rgb / src_rows may be NULL. */
int den_compact_samples_ex(const uint8_t* mask, const int32_t* offsets_in, const int32_t* offsets_out,
                           const int32_t* ray_indices_in, const float* t_starts_in,
                           const float* t_ends_in, int32_t* ray_indices_out, float* t_starts_out,
                           float* t_ends_out, int64_t n_rays, const float* sigmas_in,
                           const float* rgbs_in, int32_t channels, float* sigmas_out, float* rgbs_out,
                           int32_t* src_rows, void* stream);
/* offsets[i] = min(offsets[i], capacity) for the n_rays + 1 entries; *overflow = 1 (device int32,
 * may be NULL) when the total exceeded the capacity.  Run after the scan of a march whose sample
 * buffers were sized WITHOUT reading the total back: every per-ray consumer then stays inside the
 * buffers, and the step that lost samples is flagged instead of synchronised on. */
int den_clamp_offsets(int32_t* offsets, int64_t n_rays_plus_1, int32_t capacity, int32_t* overflow,
                      void* stream);

/* ------------------------------------------------------------------------- *
 * Transmittance / weights / accumulation — replaces
 * nerfacc.render_weight_from_density, render_weight_from_alpha and
 * accumulate_along_rays (external/vol_rendering.py:89-122).
 * ------------------------------------------------------------------------- */
int den_weight_from_density_fwd(const float* sigmas, const float* t_starts, const float* t_ends,
                                const int32_t* offsets, int64_t n_rays, float* weights, void* stream);
int den_weight_from_density_bwd(const float* sigmas, const float* t_starts, const float* t_ends,
                                const int32_t* offsets, int64_t n_rays, const float* dweights,
                                float* dsigmas, void* stream);
int den_weight_from_alpha_fwd(const float* alphas, const int32_t* offsets, int64_t n_rays,
                              float* weights, void* stream);
int den_weight_from_alpha_bwd(const float* alphas, const int32_t* offsets, int64_t n_rays,
                              const float* dweights, float* dalphas, void* stream);
/* out (R,D) = per-ray sum of w * v (v NULL -> D = 1, v = 1; w NULL -> 1); deterministic */
int den_accumulate_fwd(const float* weights, const float* values, const int32_t* offsets,
                       int64_t n_rays, int32_t dim, float* out, void* stream);
int den_accumulate_bwd(const float* weights, const float* values, const int32_t* ray_indices,
                       const float* dout, int64_t n_samples, int32_t dim, float* dweights,
                       float* dvalues, void* stream);

/* Fused front-to-back compositing — replaces external/vol_rendering.py:16-128
 * (`rendering`: weights + three accumulations + background blend) in one pass.
 * colour (R,C) = sum w c + bkgd (1 - opacity); opacity (R); depth (R) = sum w (t0+t1)/2. */
int den_composite_fwd(const float* sigmas, const float* rgbs, const float* t_starts,
                      const float* t_ends, const int32_t* offsets, int64_t n_rays, int32_t channels,
                      const float* bkgd, float* colour, float* opacity, float* depth, void* stream);
/* d_bkgd (C) is accumulated with atomics (pre-zeroed by the caller); may be NULL.
 * colour / opacity / depth are den_composite_fwd's outputs: with all three the backward is a single
 * front-to-back sweep (the suffix sums come from total - prefix); colour and depth may be NULL, which
 * selects the two-pass form (total optical depth, then a back-to-front sweep). */
int den_composite_bwd(const float* sigmas, const float* rgbs, const float* t_starts,
                      const float* t_ends, const int32_t* offsets, int64_t n_rays, int32_t channels,
                      const float* bkgd, const float* colour, const float* opacity,
                      const float* depth, const float* d_colour, const float* d_opacity,
                      const float* d_depth, float* d_sigmas, float* d_rgbs, float* d_bkgd,
                      void* stream);

/* ------------------------------------------------------------------------- *
 * Fused radiance field — replaces NGPradianceField.query_density / forward
 * (external/ngp.py:230-280) together with the tcnn.Encoding call (ngp.py:240),
 * MLP.forward (external/mlp.py:99-113), SHEncoder.forward
 * (external/sh_encoder.py:28-77) and the position computation of the
 * sigma_fn / rgb_sigma_fn closures (external/utils.py:68-96).
 * Architecture: the one every shipped config uses (base 32->64->1+15, SH degree
 * 4, head 31->64->64->C); anything else returns DEN_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------- */
typedef struct den_field_desc {
    den_hashgrid_desc grid;
    float aabb[6];                      /* radiance_field.aabb */
    int32_t contraction;                /* den_contraction */
    int32_t channels;                   /* radiance_dim: 1 (mono) or 3 (Bayer) */
    int32_t hidden_act;                 /* 0 relu, 1 softplus(beta=100)  (models/nerf.py:17-20) */
    int32_t density_act;                /* 0 shifted_trunc_exp, 1 softplus, 2 shifted_softplus */
    int32_t radiance_act;               /* 0 softplus, 1 sigmoid */
    int32_t width;                      /* n_neurons (64) */
    int32_t geo_feat_dim;               /* 15 */
    int32_t sh_degree;                  /* 4 */
    int32_t n_hidden_base;              /* 1 */
    int32_t n_hidden_head;              /* 2 */
} den_field_desc;

/* device pointers; weights in nn.Linear layout (out, in) row-major — the reference's
 * state-dict tensors are passed as they are (SURVEY.md §5 checkpoint keys) */
typedef struct den_field_params {
    const float* table;                 /* mlp_base.0.params */
    const float* wb1; const float* bb1; /* mlp_base.1.hidden_layers.0 : (64, L*2), (64) */
    const float* wb2; const float* bb2; /* mlp_base.1.output_layer    : (16, 64), (16) */
    const float* w1;  const float* b1;  /* mlp_head.hidden_layers.0   : (64, 31), (64) */
    const float* w2;  const float* b2;  /* mlp_head.hidden_layers.1   : (64, 64), (64) */
    const float* w3;  const float* b3;  /* mlp_head.output_layer      : (C, 64), (C) */
} den_field_params;

/* Samples are given by the march output: pos = o[ray] + d[ray] * (t0 + t1) / 2, dir = d[ray].
 * sigmas (M); rgbs (M,C) or NULL for a density-only evaluation (the visibility pre-pass).
 * If n_samples_dev != NULL the kernel reads the sample count from device memory
 * (min(*n_samples_dev, n_samples)) so the host never has to synchronise on the march. */
int den_field_fwd(const den_field_desc* f, const den_field_params* p, const float* rays_o,
                  const float* rays_d, const int32_t* ray_indices, const float* t_starts,
                  const float* t_ends, int64_t n_samples, const int32_t* n_samples_dev,
                  float* sigmas, float* rgbs, void* stream);
/* density at explicit world positions (n,3): the occupancy-grid update
 * (models/nerf.py:171-198 occ_eval_fn) */
int den_field_density_at(const den_field_desc* f, const den_field_params* p, const float* positions,
                         int64_t n, float* sigmas, void* stream);

/* ------------------------------------------------------------------------- *
 * Tensor-core MLP (tcgen05.mma, TMEM accumulators) on pre-encoded samples — replaces
 * MLP.forward (external/mlp.py:99-113) for mlp_base[1] / mlp_head, SHEncoder.forward and the
 * activations, as reached from NGPradianceField.query_density/_query_rgb
 * (external/ngp.py:239-267).  `enc` (M, L*2) is the output of den_hashgrid_fwd on the
 * positions written by den_contract_samples.
 * ------------------------------------------------------------------------- */
/* unit-cube positions (M,3) of marched samples: contraction of o + d (t0+t1)/2
 * (external/utils.py:68-96 + external/ngp.py:231-237) */
int den_contract_samples(const den_field_desc* f, const float* rays_o, const float* rays_d,
                         const int32_t* ray_indices, const float* t_starts, const float* t_ends,
                         int64_t n_samples, const int32_t* n_samples_dev, float* unit_pos, void* stream);
/* sigmas (M); rgbs (M,C) or NULL (density only) */
int den_mlp_fwd(const den_field_desc* f, const den_field_params* p, const float* enc,
                const float* rays_o, const float* rays_d, const int32_t* ray_indices,
                const float* t_starts, const float* t_ends, int64_t n_samples,
                const int32_t* n_samples_dev, float* sigmas, float* rgbs, void* stream);

/* fp32 gradient accumulators of the MLP parameters (same shapes as den_field_params; the
 * kernel ADDS into them with atomics, so the caller zeroes them or passes .grad buffers) */
typedef struct den_field_grads {
    float* wb1; float* bb1; float* wb2; float* bb2;
    float* w1;  float* b1;  float* w2;  float* b2;  float* w3;  float* b3;
} den_field_grads;

/* Backward of den_mlp_fwd with forward recompute: d_enc (M, L*2) is written, the weight /
 * bias gradients are accumulated into `g`.  d_sigmas (M), d_rgbs (M,C).  d_dirs (M,3) or NULL:
 * dL/d(view direction) through the SH encoding.  enc_rows (M) int32 or NULL: sample i reads row
 * enc_rows[i] of `enc` (den_compact_samples_ex's src_rows: the pre-pass encodings in place). */
int den_mlp_bwd(const den_field_desc* f, const den_field_params* p, const den_field_grads* g,
                const float* enc, const float* rays_o, const float* rays_d,
                const int32_t* ray_indices, const float* t_starts, const float* t_ends,
                const float* d_sigmas, const float* d_rgbs, int64_t n_samples,
                const int32_t* n_samples_dev, const int32_t* enc_rows, float* d_enc, float* d_dirs,
                void* stream);
/* Reverse mode of den_contract_samples w.r.t. the rays (the refractory-period gradient path,
 * models/trajectories.py -> models/nerf.py:206-228 -> external/utils.py:83-96): per sample
 * d_pos = J^T d_unit and d_pos_t = d_pos * (t0+t1)/2; den_accumulate_fwd sums them per ray
 * into dL/d rays_o and dL/d rays_d. */
int den_contract_samples_bwd(const den_field_desc* f, const float* rays_o, const float* rays_d,
                             const int32_t* ray_indices, const float* t_starts, const float* t_ends,
                             const float* d_unit, int64_t n_samples, const int32_t* n_samples_dev,
                             float* d_pos, float* d_pos_t, void* stream);

/* ------------------------------------------------------------------------- *
 * Pixel-bandwidth low-pass filter — replaces PixelBandwidth.intensity_sample_to_weight,
 * linearize_sys, discretized_sys_to_weight and the normalised weighted sum of
 * weighted_it_sample_to_output_log_it (models/pixel_bandwidth.py:181-228,260-296,369-415)
 * with control.foh_cont2discrete (utils/control.py:29-123).
 * intensity (S,N) fp32; sample_dt_ns (S-1,N) fp32 [ns]; coef: 5 fp64 on the DEVICE
 * (alpha0, alpha1, beta, omega_sf, omega_diff: a = alpha0 + alpha1 I, b = beta I);
 * out (N, n_channels) fp32: n_channels == 2 -> (source-follower output, diff-amp output) for
 * the reset call, 1 -> diff-amp output.  fp64 internally.
 * ------------------------------------------------------------------------- */
int den_lpf_fwd(const float* intensity, const float* sample_dt_ns, const double* coef, int32_t S,
                int64_t N, int32_t n_channels, float* out, void* stream);
/* d_intensity (S,N) written; d_coef (5, fp64, device) accumulated with atomics (may be NULL) */
int den_lpf_bwd(const float* intensity, const float* sample_dt_ns, const double* coef, int32_t S,
                int64_t N, int32_t n_channels, const float* d_out, float* d_intensity,
                double* d_coef, void* stream);

/* ------------------------------------------------------------------------- *
 * Pixel-bandwidth filter fused with the event loss, every render request of a training step in
 * ONE launch — replaces, per step, the filter half of the 2P calls of PixelBandwidth.forward
 * (models/pixel_bandwidth.py:369-448, including the differencing-amplifier reset carried from the
 * first request to the others, :419-446) and Loss.compute (loss_metric/loss.py:34-96).
 * Requests k = 0 .. K-1 come in P = K/2 pairs (2p, 2p+1) = (start, end) of a supervision interval
 * (models/deblur_e_nerf.py:472-521: pair 0 = `diff`, pair 1 = `subdiff`); with has_reset request 0 is
 * the reset call: delta = diff-amp output - source-follower output, its own value is the
 * source-follower output, and request k >= 1 yields out_k - delta exp(-omega_diff 1e-9 reset_dt_k).
 *   pred_p = final[2p+1] - final[2p];   err_p = E_kind(pred_p inv_k[p], target[p]);
 *   terms[p] = mean of err_p over the events with valid[p] != 0   (counts[p] of them).
 * intensity (K,S,N) fp32; sample_dt_ns (K,S-1,N) fp32; coef 5 fp64 (den_lpf_fwd); reset_dt_ns (K,N)
 * fp64 = output_ts[k] - output_ts[0] (row 0 unused; NULL without reset); target (P,N) fp32 or NULL;
 * inv_k (P) fp32 on the device; valid (P,N) uint8.  S <= 32, K even <= 8.  fp64 inside; the means are
 * summed in event order by the last CTA to finish (deterministic).  log_intensity (K,N), optional:
 * the K filter outputs after the reset.  workspace: den_lpf_loss_workspace_bytes(P, N), zeroed ONCE.
 * ------------------------------------------------------------------------- */
typedef struct den_lpf_loss_desc {
    int32_t it_sample_size;             /* S */
    int32_t n_requests;                 /* K */
    int32_t has_reset;
    int32_t error_kind[4];              /* per pair: 0 l1, 1 mse, 2 huber (delta 1), 3 mape */
    int32_t has_target[4];              /* per pair: 0 -> target is zero (the TV term) */
} den_lpf_loss_desc;

size_t den_lpf_loss_workspace_bytes(int32_t n_pairs, int64_t N);
int den_lpf_loss_fwd(const den_lpf_loss_desc* d, const float* intensity, const float* sample_dt_ns,
                     const double* coef, const double* reset_dt_ns, const float* target,
                     const float* inv_k, const uint8_t* valid, int64_t N, float* terms, int32_t* counts,
                     float* log_intensity, void* workspace, void* stream);
/* d_terms (P) fp32 on the device; d_intensity (K,S,N) written; d_coef (5 fp64), d_inv_k (P fp64)
 * accumulated with atomics (pre-zeroed, may be NULL); d_reset_dt_ns (K,N) fp64 and d_target (P,N)
 * fp32 written (may be NULL; rows of pairs without a target are left untouched). */
int den_lpf_loss_bwd(const den_lpf_loss_desc* d, const float* intensity, const float* sample_dt_ns,
                     const double* coef, const double* reset_dt_ns, const float* target,
                     const float* inv_k, const uint8_t* valid, int64_t N, const int32_t* counts,
                     const float* d_terms, float* d_intensity, double* d_coef, double* d_reset_dt_ns,
                     float* d_target, double* d_inv_k, void* stream);

/* ------------------------------------------------------------------------- *
 * Optimiser — replaces torch.optim.Adam as set up by DeblurENeRF.configure_optimizers
 * (models/deblur_e_nerf.py:1055-1112) for the fp32 parameters: one step t (1-based) of
 *   g = grad_scale * grad + weight_decay * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
 * over a HOST array of tensor descriptors (device pointers inside), each with its own lr / decay.
 * ------------------------------------------------------------------------- */
typedef struct den_adam_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t n;
    float lr;
    float weight_decay;
} den_adam_tensor;

/* grad_scale: 1 for a single process; 1 / world_size when the gradients were SUM all-reduced
 * (the mean of Lightning's DDP, scripts/run.py:84-89, folded into the update). */
/* step_dev: NULL, or a device int64 holding the step number t (then `step` is ignored): a step
 * captured in a CUDA graph keeps its step counter on the device and increments it inside the graph. */
/* skip_flag: NULL, or a device int32; when it is non-zero the launch changes nothing (the overflow
 * flag of den_clamp_offsets: a step that lost samples applies no update). */
int den_adam_step(const den_adam_tensor* tensors_host, int32_t n_tensors, double beta1, double beta2,
                  double eps, int64_t step, const int64_t* step_dev, double grad_scale,
                  const int32_t* skip_flag, void* stream);

/* ------------------------------------------------------------------------- *
 * Evaluation post-processing on the device — replaces the host part of
 * DeblurENeRF.evaluation_epoch_end (models/deblur_e_nerf.py:705-969): images (B, C, HW) fp32;
 * every entry point is one pass over the images with fp64 accumulation (atomics into pre-zeroed
 * outputs).  params (C, 5) fp64 on the device: a, b (the log-space affine correction of :789-797),
 * s, gamma, o (OffsetGammaCorrection, models/offset_gamma_correction.py:37-40); gain (B) fp64: the
 * mean-normalised gain-exposure products (:707-712); log_gain their logs.
 *   affine moments (C, 5): n, sum x, sum y, sum xx, sum xy, x = log pred, y = log target - log_gain_b
 *   lm moments (C, 10): J^T J (ss, sg, so, gg, go, oo), J^T r (s, g, o), sum r^2 for
 *       f = gain_b (s x^gamma - o), x = exp(a log pred + b), r = f - target
 *       (external/optimizer.py:86-92 builds exactly these normal equations from a B*HW x 3 Jacobian)
 *   apply: out = f (fp32) and image_sums (B, 2): sum |out - target|, sum (out - target)^2
 *       (the L1 and PSNR terms of loss_metric/metric.py:57-72)
 * ------------------------------------------------------------------------- */
int den_eval_affine_moments(const float* pred, const float* target, const double* log_gain, int32_t B,
                            int32_t C, int64_t HW, double* moments, void* stream);
int den_eval_lm_moments(const float* pred, const float* target, const double* gain, const double* params,
                        int32_t B, int32_t C, int64_t HW, double* moments, void* stream);
int den_eval_apply(const float* pred, const float* target, const double* gain, const double* params,
                   int32_t B, int32_t C, int64_t HW, float* out, double* image_sums, void* stream);
/* SSIM term of Metric.compute (loss_metric/metric.py:74-81: torchmetrics 0.6.2 functional.ssim with
 * data_range = max_target_val): images (B, C, H, W) fp32; Gaussian window kernel_size x kernel_size (odd,
 * <= 15; upstream default 11) with `sigma` (1.5); c1 = (k1 data_range)^2, c2 = (k2 data_range)^2
 * (k1 0.01, k2 0.03).  image_sums (B) fp64, pre-zeroed: the sum of the SSIM index over the channels and
 * over the (H - k + 1) x (W - k + 1) pixels whose window lies inside the image (upstream crops the
 * reflect-padded border before its mean); the mean divides by B C (H - k + 1)(W - k + 1). */
int den_eval_ssim(const float* pred, const float* target, int32_t B, int32_t C, int32_t H, int32_t W,
                  int32_t kernel_size, double sigma, double c1, double c2, double* image_sums, void* stream);

/* ------------------------------------------------------------------------- *
 * Raw event stream -> queued events (the data format in front of the hot path) — replaces the per-event
 * Python loops of Event.queue_raw_events (data/datasets.py:186-276) and
 * Event.extract_max_refractory_period (:131-183).  Both need, per raw event i, the PREVIOUS raw event at the
 * same pixel: a stable sort of the stream indices by pixel id y * width + x puts it next to i.
 *   den_radix_sort_pairs_u32: stable LSD radix sort of n (u32 key, u32 value) pairs on the low `key_bits`
 *       bits (8 per pass); the result lands in (keys_out, vals_out), (keys_tmp, vals_tmp) is scratch of the
 *       same size; the inputs are not modified.  workspace >= den_radix_sort_workspace_bytes(n).
 *   den_queue_raw_events: position_xy (n, 2) int32 (x, y), timestamp (n) int64 in stream order ->
 *       an event is KEPT iff an earlier event exists at its pixel and the latest one has a different
 *           timestamp (the sliding-window test of :246-253);
 *       start_ts (n) int64: that earlier event's timestamp (0 where not kept); end_ts is `timestamp` itself,
 *           num_pos / num_neg are polarity / 1 - polarity (:255-267);
 *       min_interval (1) int64, PRE-SET by the caller to INT64_MAX: atomically lowered to the smallest
 *           non-zero timestamp[i] - timestamp[prev(i)] — the maximum refractory period of :131-183
 *           (still INT64_MAX: no pixel saw two distinct timestamps; upstream keeps +inf);
 *       kept_offsets (n + 1) int32: exclusive prefix sum of the keep flags — the row of event i among the
 *           kept events (event i is kept iff kept_offsets[i + 1] > kept_offsets[i]); kept_offsets[n] = their
 *           number M;
 *       out_of_range (1) int32, pre-zeroed: set to 1 if a position lies outside width x height (upstream
 *           raises IndexError; the caller checks the flag).
 *       workspace >= den_queue_events_workspace_bytes(n).  n < 2^31.
 *   den_compact_queued_events: the kept events in stream order in upstream's layout (:232-238,270-274):
 *       out_position (M, 2) int64, out_start_ts / out_end_ts / out_num_pos / out_num_neg (M) int64;
 *       polarity (n) u8 (0 / 1).
 * ------------------------------------------------------------------------- */
size_t den_radix_sort_workspace_bytes(int64_t n);
int den_radix_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                             uint32_t* vals_out, uint32_t* keys_tmp, uint32_t* vals_tmp, int64_t n,
                             int32_t key_bits, void* workspace, size_t workspace_bytes, void* stream);
size_t den_queue_events_workspace_bytes(int64_t n);
int den_queue_raw_events(const int32_t* position_xy, const int64_t* timestamp, int64_t n, int32_t width,
                         int32_t height, void* workspace, size_t workspace_bytes, int64_t* start_ts,
                         int32_t* kept_offsets, int64_t* min_interval, int32_t* out_of_range, void* stream);
int den_compact_queued_events(const int32_t* position_xy, const int64_t* timestamp, const uint8_t* polarity,
                              const int64_t* start_ts, const int32_t* kept_offsets, int64_t n,
                              int64_t* out_position, int64_t* out_start_ts, int64_t* out_end_ts,
                              int64_t* out_num_pos, int64_t* out_num_neg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEN_B200_H_ */
