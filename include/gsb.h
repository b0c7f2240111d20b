/* gsb.h — C ABI of libgsb.so, the B200-native (sm_100a) replacement for the reference's
 * kernel boundary on the render → loss → backward → Adam path.
 *
 * Reference boundary being replaced (paths relative to tatsuya-ogawa/GaussianSplattingMlx):
 *   Trainer/SlangKernelSpecLoader.swift:35-48   loadKernel(named:) -> MLXFastKernel   (12 kernels)
 *   Trainer/GaussianRenderer.swift:124-147,187-226,282-299,388-482,510-573,605-701   (kernel call sites)
 *   Trainer/GaussianTrainer.swift:555-625,719,1066-1086                              (SSIM, valueAndGrad, Adam)
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  Unless a parameter is named host_*, every
 *     pointer is a DEVICE pointer owned by the caller, row-major, f32/u32, byte-identical to the
 *     MLX arrays of the reference (layouts quoted per function).
 *   - every function returns 0 (GSB_OK) or a negative gsb_status; gsb_last_error(ctx) gives text.
 *   - work is enqueued on the context's CUDA stream (gsb_set_stream); functions do not synchronise
 *     unless documented.  One context = one caller thread at a time, like the reference's
 *     renderer/trainer objects (closure-captured saved state, GaussianRenderer.swift:119-122).
 *   - there is NO CPU fallback: without a CUDA device gsb_create fails with GSB_ERR_CUDA.
 */
#ifndef GSB_H_
#define GSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_ABI_VERSION 2

#if defined(__GNUC__)
#define GSB_API __attribute__((visibility("default")))
#else
#define GSB_API
#endif

typedef enum gsb_status {
    GSB_OK = 0,
    GSB_ERR_INVALID = -1,      /* bad argument (null pointer, size mismatch, N > max_gaussians ...) */
    GSB_ERR_CUDA = -2,         /* CUDA runtime error; text in gsb_last_error */
    GSB_ERR_UNSUPPORTED = -3,  /* configuration outside what the kernels implement */
    GSB_ERR_STATE = -4,        /* backward called without a matching forward, etc. */
    GSB_ERR_CAPACITY = -5      /* intersection list outgrew its buffer; retried internally when possible */
} gsb_status;

/* Mirrors GaussianRenderer.init(active_sh_degree:W:H:TILE_SIZE:whiteBackground:)
 * (Trainer/GaussianRenderer.swift:703-710) plus trainer constants (GaussianTrainer.swift:277-300). */
typedef struct gsb_config {
    int32_t width, height;        /* W, H */
    int32_t tile_w, tile_h;       /* TILE_SIZE.w/.h (any size >= 1; 16x16 is the fast path) */
    int32_t sh_degree;            /* active_sh_degree, 0..4 */
    int32_t sh_coeffs;            /* K = coefficient stride of shs[N,K,3]; >= (sh_degree+1)^2, <= 25 */
    int32_t white_background;     /* whiteBackground */
    int32_t max_gaussians;        /* capacity; grows on demand when 0 */
    int32_t device;               /* CUDA device ordinal */
    int32_t flags;                /* GSB_FLAG_* */
    float lambda_dssim;           /* 0.2  (GaussianTrainer.swift:277) */
    float adam_beta1, adam_beta2; /* 0.9, 0.999 (GaussianTrainer.swift:941-945) */
    float adam_eps;               /* 1e-15 */
} gsb_config;

#define GSB_FLAG_SORT_CUB 1       /* use the CUB baseline instead of the hand-written onesweep (checking only) */
#define GSB_FLAG_NO_OVERLAP 2     /* trainer: run every view's kernels back to back on one stream (per-kernel timing) */
#define GSB_FLAG_ASYNC_LOSS 4     /* gsb_trainer_accumulate / gsb_train_step: host_loss is PINNED host memory; the loss is copied
                                   * into it asynchronously on the work stream and the call does not synchronise */
#define GSB_FLAG_NVTX 16          /* wrap every stage in an NVTX range named "<reference profiler section>/<stage>", e.g.
                                   * "bwd.globalTileComposite/raster_bwd" (sections of GaussianTrainer.swift:122-241) */
#define GSB_FLAG_NO_SEGMENTS 8    /* raster backward: one work item per 16x16 block instead of checkpointed 256-Gaussian segments
                                   * (checking only: same gradients up to rounding, worse load balance) */

/* Camera block = the 7 camera arrays of TrainStepInputIndex (GaussianTrainer.swift:254-272),
 * produced exactly as Trainer/CameraUtil.swift:5-102 does (row-vector convention, proj = P^T). */
typedef struct gsb_camera {
    float view[16];       /* worldViewTransform, row-major */
    float proj[16];       /* projectionMatrix,  row-major */
    float cam_center[3];  /* cameraCenter */
    float fov_x, fov_y;   /* FoVx, FoVy (radians) */
    float focal_x, focal_y;
} gsb_camera;

typedef struct gsb_ctx gsb_ctx;

/* ---- lifetime ------------------------------------------------------------------------------ */
GSB_API int gsb_abi_version(void);
GSB_API void gsb_default_config(gsb_config* cfg);
GSB_API int gsb_create(const gsb_config* cfg, gsb_ctx** out);
GSB_API void gsb_destroy(gsb_ctx* ctx);
GSB_API const char* gsb_last_error(const gsb_ctx* ctx);       /* ctx may be NULL: last creation error */
GSB_API int gsb_set_stream(gsb_ctx* ctx, void* cuda_stream);  /* cudaStream_t; NULL = the legacy default stream.
                                                              * Before the first call a context uses a private non-blocking stream. */
GSB_API int gsb_synchronize(gsb_ctx* ctx);
GSB_API int gsb_set_flags(gsb_ctx* ctx, int32_t flags);      /* replaces gsb_config.flags (GSB_FLAG_*); synchronises */

/* ---- activations: GaussianRenderer.get_*_from (GaussianRenderer.swift:936-963) -------------- */
/* in : f_dc[N,1,3] f_rest[N,K-1,3] scales_log[N,3] rot_raw[N,4] opacity_logit[N,1]
 * out: shs[N,K,3] scales[N,3] rotations[N,4] opacity[N,1]         (means3d = xyz, identity) */
GSB_API int gsb_activate_fwd(gsb_ctx*, int32_t N, const float* f_dc, const float* f_rest, const float* scales_log,
                     const float* rot_raw, const float* opacity_logit, float* shs, float* scales,
                     float* rotations, float* opacity);
GSB_API int gsb_activate_bwd(gsb_ctx*, int32_t N, const float* scales_log, const float* rot_raw,
                     const float* opacity_logit, const float* g_shs, const float* g_scales,
                     const float* g_rotations, const float* g_opacity, float* g_f_dc, float* g_f_rest,
                     float* g_scales_log, float* g_rot_raw, float* g_opacity_logit);

/* ---- K1 / K2: gaussian_projection_screen_fused_{forward,backward} ---------------------------
 * (slang/gaussian_projection_kernels.slang:36-173,205-398; call sites GaussianRenderer.swift:510-573,605-701)
 * in : scales[N,3] rotations[N,4] means3d[N,3] shs[N,K,3] (ACTIVATED tensors) + camera
 * out: means2d[N,2] depths[N] color[N,3] cov2d[N,2,2] conic[N,2,2] radii[N] rectMin[N,2] rectMax[N,2] */
GSB_API int gsb_project_fwd(gsb_ctx*, int32_t N, const float* scales, const float* rotations, const float* means3d,
                    const float* shs, const gsb_camera* host_cam, float* means2d, float* depths, float* color,
                    float* cov2d, float* conic, float* radii, float* rect_min, float* rect_max);
/* cot*: cotangents of depths[N], means2d[N,2], cov2d[N,4], color[N,3], conic[N,4]
 * out : g_scales[N,3] g_rotations[N,4] g_means3d[N,3] g_shs[N,K,3] g_cam_center_point[N,3] */
GSB_API int gsb_project_bwd(gsb_ctx*, int32_t N, const float* scales, const float* rotations, const float* means3d,
                    const float* shs, const gsb_camera* host_cam, const float* cot_depths,
                    const float* cot_means2d, const float* cot_cov2d, const float* cot_color,
                    const float* cot_conic, float* g_scales, float* g_rotations, float* g_means3d, float* g_shs,
                    float* g_cam_center_point);

/* ---- K3..K8: buildGlobalTileSliceInfo (GaussianRenderer.swift:333-490) ----------------------
 * count_tiles_per_gaussian → cumsum → generate_keys → radix sort on (tile id, depth bits) → tile
 * ranges / counts.  Output is CSR (ranges into the sorted list); the reference's dense
 * [numTiles,maxTilePairs] padding (K8) holds exactly sorted_gauss_idx[range.start + slot].
 * in : rect_min[N,2] rect_max[N,2] radii[N] depths[N]
 * out: tiles_touched u32[N], *host_M = number of (Gaussian,tile) pairs (synchronises),
 *      tile_ranges u32[numTiles,2], tile_counts u32[numTiles]; the sorted lists stay inside the
 *      context and are read back with gsb_bin_read. */
GSB_API int gsb_bin(gsb_ctx*, int32_t N, const float* rect_min, const float* rect_max, const float* radii,
            const float* depths, uint32_t* tiles_touched, uint32_t* tile_ranges, uint32_t* tile_counts,
            uint32_t* host_M);
/* Copies the last gsb_bin / render-forward lists into caller DEVICE buffers of M elements each (any may
 * be NULL): unsorted keys (emission order) and sorted keys/values.  For the bit-exact parity tests. */
GSB_API int gsb_bin_read(gsb_ctx*, uint32_t* keys_high, uint32_t* keys_low, uint32_t* gauss_idx,
                 uint32_t* sorted_keys_high, uint32_t* sorted_keys_low, uint32_t* sorted_gauss_idx);
/* Stand-alone stable sort of M (key_high[tile_bits], key_low[32]) pairs with 32-bit payloads —
 * the replacement of radix_sort_tile_keys_fused_forward (slang/gaussian_tile_global_kernels.slang:143-305,
 * GaussianRenderer.swift:271-301).  use_cub != 0 runs cub::DeviceRadixSort (checked baseline). */
GSB_API int gsb_sort_tile_keys(gsb_ctx*, uint32_t M, uint32_t tile_bits, const uint32_t* keys_high,
                       const uint32_t* keys_low, const uint32_t* values, uint32_t* sorted_high,
                       uint32_t* sorted_low, uint32_t* sorted_values, int32_t use_cub);

/* ---- K9 / K10: gaussian_tile_global_{forward,backward} --------------------------------------
 * (slang/gaussian_tile_global_kernels.slang:523-614,648-881; GaussianRenderer.swift:124-147,187-226)
 * Uses the tile lists of the preceding gsb_bin on this context.
 * in : packed[N,11] = mean2d(2) conic(4) color(3) opacity depth   (GaussianRenderer.swift:45-51)
 * out: color[P,3] depth[P] alpha[P] last_contrib u32[P],  P = W*H */
GSB_API int gsb_raster_fwd(gsb_ctx*, int32_t N, const float* packed, float* out_color, float* out_depth,
                   float* out_alpha, uint32_t* out_last_contrib);
/* out: grad_packed[N,11] (overwritten, not accumulated) */
GSB_API int gsb_raster_bwd(gsb_ctx*, int32_t N, const float* packed, const float* cot_color, const float* cot_depth,
                   const float* cot_alpha, const float* out_color, const float* out_depth,
                   const float* out_alpha, const uint32_t* last_contrib, float* grad_packed);

/* ---- K11 / K12: ssim_{forward,backward} (slang/ssim_kernels.slang:94-155,181-266;
 *      GaussianTrainer.swift:555-625).  HWC f32 images, 11x11 window of LossUtil.swift:47-54
 *      (sigma 1.5, centre 5.5), zero padding.  Any of the five saved maps may be NULL. */
GSB_API int gsb_ssim_fwd(gsb_ctx*, int32_t H, int32_t W, int32_t C, const float* img1, const float* img2,
                 float* ssim_map, float* mu1, float* mu2, float* sigma1_sq, float* sigma2_sq, float* sigma12);
GSB_API int gsb_ssim_bwd(gsb_ctx*, int32_t H, int32_t W, int32_t C, const float* grad_out, const float* img1,
                 const float* img2, float* grad_img1);

/* ---- fused renderer path: GaussianRenderer.forwardWithCameraParams (GaussianRenderer.swift:823-880)
 * Takes the RAW model tensors (GaussianModel.swift:33-55) and fuses activations + K1 + K3..K9.
 * in : xyz[N,3] f_dc[N,1,3] f_rest[N,K-1,3] scales_log[N,3] rot_raw[N,4] opacity_logit[N,1]
 * out: render[H,W,3] depth[H,W,1] alpha[H,W,1] visibility u8[N] radii[N]  (any may be NULL)
 * State needed by gsb_render_backward is saved inside the context (one forward in flight). */
GSB_API int gsb_render_forward(gsb_ctx*, int32_t N, const float* xyz, const float* f_dc, const float* f_rest,
                       const float* scales_log, const float* rot_raw, const float* opacity_logit,
                       const gsb_camera* host_cam, float* render, float* depth, float* alpha,
                       uint8_t* visibility, float* radii);
/* cot_depth / cot_alpha may be NULL (= zeros).  g_* are the gradients wrt the RAW tensors; when
 * accumulate != 0 they are added to (reduce-add), otherwise overwritten. */
GSB_API int gsb_render_backward(gsb_ctx*, const float* cot_render, const float* cot_depth, const float* cot_alpha,
                        float* g_xyz, float* g_f_dc, float* g_f_rest, float* g_scales_log, float* g_rot_raw,
                        float* g_opacity_logit, int32_t accumulate);

/* ---- loss: GaussianTrainer lossFn (GaussianTrainer.swift:688-716) ---------------------------
 * total = (1-lambda)*mean|render-target| + lambda*(1-mean(ssim_map)); writes d total/d render
 * (scaled by grad_scale, e.g. 1/B for a batch of B views) into cot_render[H,W,3] and adds
 * grad_scale*total into *loss_accum (device scalar). */
GSB_API int gsb_loss_fwd_bwd(gsb_ctx*, const float* render, const float* target_rgb, float grad_scale,
                     float* cot_render, float* loss_accum);
/* The same with the depth-supervision term of lossFn (GaussianTrainer.swift:693-699,710-714; the trainer enables it
 * whenever the dataset carries depth, :949):
 *   total += lambda_depth * sum(|depth - target_depth| * mask) / max(sum(mask), 1e-6)
 * depth[H,W,1] = the renderer's depth output, target_depth[H,W] f32, depth_mask[H,W] one byte per pixel (the MLX bool
 * array `alpha > 0.5`, GaussianTrainer.swift:492).  Also writes cot_depth[H,W,1] = d total / d depth (scaled by grad_scale),
 * to be handed to gsb_render_backward. */
GSB_API int gsb_loss_fwd_bwd_depth(gsb_ctx*, const float* render, const float* depth, const float* target_rgb,
                           const uint8_t* depth_mask, const float* target_depth, float lambda_depth, float grad_scale,
                           float* cot_render, float* cot_depth, float* loss_accum);

/* ---- Adam (MLXOptimizers.Adam.applySingle x6, GaussianTrainer.swift:1066-1079) + D1
 *      accum_grad_norm (GaussianTrainer.swift:321-339), one launch.  No bias correction.
 * params/grads/m/v: 6 device pointers each in GaussModel.getParams() order
 * (xyz, f_dc, f_rest, scales, rotation, opacity); counts[6] = number of floats of each tensor;
 * host_lrs[6] from GaussModel.getLearningRates; grad_norm_accum[N] may be NULL. */
GSB_API int gsb_adam_step(gsb_ctx*, int32_t N, float* const* host_params, const float* const* host_grads,
                  float* const* host_m, float* const* host_v, const int64_t* host_counts,
                  const float* host_lrs, float* grad_norm_accum);

/* ---- trainer: one GaussianTrainer.startTrain iteration (GaussianTrainer.swift:958-1086) over a
 *      batch of B views: L = (1/B) sum_v L_v.  Parameters, Adam state and gradient buffers live in
 *      the context (gsb_trainer_*).  host_targets[b] may be pinned HOST memory when
 *      targets_on_host != 0 (copied H2D on a side stream, overlapped with the previous view). */
GSB_API int gsb_trainer_init(gsb_ctx*, int32_t N, const float* host_xyz, const float* host_f_dc,
                     const float* host_f_rest, const float* host_scales_log, const float* host_rot_raw,
                     const float* host_opacity_logit);
GSB_API int gsb_trainer_param_ptrs(gsb_ctx*, float** host_params6, float** host_grads6, float** host_m6,
                           float** host_v6, float** grad_norm_accum);
/* The six gradient tensors live in ONE contiguous device block of *floats f32 (128-byte aligned
 * segments, padding kept at zero) so that a data-parallel host can all-reduce them with a single
 * collective between gsb_trainer_accumulate and gsb_trainer_apply. */
GSB_API int gsb_trainer_grad_block(gsb_ctx*, float** grad_block, int64_t* floats);
/* Forward+loss+backward for B views into the context's gradient buffers (no optimiser step).
 * host_loss (may be NULL) receives the mean loss and makes the call synchronise (unless GSB_FLAG_ASYNC_LOSS). */
GSB_API int gsb_trainer_accumulate(gsb_ctx*, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                           int32_t targets_on_host, int32_t zero_grads, float grad_scale, float* host_loss);
/* gsb_trainer_accumulate with depth supervision (GaussianTrainer.swift:949: effectiveLambdaDepth = lambda_depth when the
 * dataset has depth): host_target_depths[b] -> f32[H,W], host_depth_masks[b] -> one byte per pixel; both live where the
 * RGB targets live (device, or pinned host memory when targets_on_host != 0). */
GSB_API int gsb_trainer_accumulate_depth(gsb_ctx*, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                                 const float* const* host_target_depths, const uint8_t* const* host_depth_masks,
                                 float lambda_depth, int32_t targets_on_host, int32_t zero_grads, float grad_scale,
                                 float* host_loss);
/* Adam + D1 on the context's buffers; learning rates from (iteration, total_iterations)
 * (GaussianModel.swift:56-65).  reset_state != 0 re-zeroes m/v first (GaussianTrainer.swift:1104-1109). */
GSB_API int gsb_trainer_apply(gsb_ctx*, int32_t iteration, int32_t total_iterations, int32_t reset_state);
/* Data-parallel step fused with its collective, over NVLink peer memory (one process per GPU on one node, <= 8):
 *   every replica calls gsb_trainer_peers_export (GSB_PEER_BLOB_BYTES of host memory: CUDA IPC handles of its trainer
 *   slab + layout), the host exchanges the blobs (any transport) and hands all of them, in rank order, to
 *   gsb_trainer_peers_import.  Per step, INSTEAD of "all-reduce the gradient block; gsb_trainer_apply":
 *       barrier (stream-ordered: every replica's gsb_trainer_accumulate is complete)
 *       gsb_trainer_apply_peers   -- rank r sums slice r of ALL replicas' gradients (peer loads), applies Adam + D1 with
 *                                    its local m / v, stores the new parameters into EVERY replica (peer stores)
 *       barrier (every replica's parameters are written)
 * The mapping survives gsb_trainer_init with the same N; it is dropped by a different N and by gsb_destroy.
 * gsb_trainer_apply_peers fails with GSB_ERR_STATE without it.  Densification swaps (and may free) the slab the other
 * replicas have mapped: EVERY replica must call gsb_trainer_peers_close and pass a host barrier before
 * gsb_trainer_densify (which otherwise fails with GSB_ERR_STATE), then export / import the new slabs; the same close +
 * barrier is required before any replica's gsb_destroy. */
/* gsb_trainer_step_peers: the WHOLE data-parallel step with device-side synchronisation - no host barrier, no NCCL call.
 *   = gsb_trainer_accumulate(B views, zero_grads, grad_scale) + the exchange, where the replicas meet only in flag words
 *   kept in peer-mapped memory: the projection backward of the step's last view runs in Gaussian chunks, each chunk is
 *   announced to every replica, and the exchange kernel of chunk k (reduce the owned slice over all replicas + Adam + D1 +
 *   store the parameters into every replica; NVLS multimem variant when symmetric buffers are attached) starts as soon as
 *   all replicas have announced chunk k - so it overlaps the projection backward of chunk k+1 - and announces its stores
 *   back.  The next batch on any replica begins by waiting for those announcements.  Every replica must call it the same
 *   number of times (B may be 0 on a replica without views).  Waits are bounded (2 s): gsb_trainer_peers_check reports a
 *   replica that never arrived.  May be mixed with the barrier-bracketed gsb_trainer_apply_peers form. */
#define GSB_PEER_BLOB_BYTES 256
GSB_API int gsb_trainer_peers_export(gsb_ctx*, void* host_blob, int64_t blob_bytes);
GSB_API int gsb_trainer_peers_import(gsb_ctx*, int32_t world, int32_t rank, const void* host_blobs, int64_t blob_bytes);
GSB_API int gsb_trainer_apply_peers(gsb_ctx*, int32_t iteration, int32_t total_iterations, int32_t reset_state);
GSB_API int gsb_trainer_step_peers(gsb_ctx*, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                           int32_t targets_on_host, float grad_scale, int32_t iteration, int32_t total_iterations,
                           int32_t reset_state, float* host_loss);
/* tuning of the step protocol: Gaussian chunks per step (1..8, 0 = keep; default 2) and CTAs of the peer / multicast
 * exchange kernels (0 = default: 2 per SM beside the projection backward, 1 per SM for the NVLS kernel) */
GSB_API int gsb_trainer_peers_tune(gsb_ctx*, int32_t chunks, int32_t peer_blocks, int32_t multicast_blocks);
GSB_API int gsb_trainer_peers_check(gsb_ctx*);   /* GSB_ERR_STATE when a bounded wait of the step protocol ran out; synchronises */
/* Unmaps the other replicas' slabs.  Call it on every replica (and synchronise the replicas) BEFORE any of them destroys
 * its context: memory exported through CUDA IPC must not be freed while another process still has it open. */
GSB_API int gsb_trainer_peers_close(gsb_ctx*);
/* The same step through the NVSwitch (NVLS).  The host allocates, per replica, two SYMMETRIC buffers of at least
 * gsb_trainer_grad_block's float count (same size on every replica, bound to one multicast range - e.g.
 * torch.distributed._symmetric_memory) and attaches them: the parameters move into params_local, gradients are
 * accumulated into grads_local; params_mc / grads_mc are the multicast addresses of those buffers.
 * gsb_trainer_apply_multicast (between the same two barriers as above) then reads slice r of the gradient ALREADY SUMMED
 * by the switch (multimem.ld_reduce), applies Adam + D1 and multicasts the new parameters (multimem.st): 1/world of
 * the two blocks crosses this GPU's links instead of (world-1)/world.  gsb_trainer_init and densification detach. */
GSB_API int gsb_trainer_attach_symmetric(gsb_ctx*, int32_t world, int32_t rank, float* params_local, float* grads_local,
                                 float* params_mc, float* grads_mc, int64_t floats);
GSB_API int gsb_trainer_apply_multicast(gsb_ctx*, int32_t iteration, int32_t total_iterations, int32_t reset_state);
/* accumulate (zero_grads=1, scale 1/B) followed by apply. */
GSB_API int gsb_train_step(gsb_ctx*, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                   int32_t targets_on_host, int32_t iteration, int32_t total_iterations, float* host_loss);

/* ---- densification: split_and_prune (Trainer/GaussianTrainer.swift:766-908) ------------------
 * D2 classify_gaussians (:344-392): actions i32[N] = 0 keep / 1 split / 2 clone / 3 prune and the per-Gaussian
 * output count 1 / 2 / 2 / 0.  denom = denomGradAccumulation (iterations accumulated into grad_accum). */
GSB_API int gsb_densify_classify(gsb_ctx*, int32_t N, const float* grad_accum, float denom, const float* scales_log,
                         const float* opacity_logit, float grad_threshold, float max_scale, float min_opacity,
                         int32_t allow_densify, int32_t* actions, int32_t* output_counts);
/* cumsum -> offsets (:813-817) + D3 build_densify_output_map (:397-427).  offsets i32[N] may be NULL.
 * gather_indices / noise_mode have `capacity` slots (2N always suffices); *host_total = total output count
 * (synchronises).  noise_mode: 0 none, 1 split first, 2 split second, 3 clone copy. */
GSB_API int gsb_densify_map(gsb_ctx*, int32_t N, const int32_t* actions, const int32_t* output_counts, int32_t* offsets,
                    int32_t capacity, int32_t* gather_indices, int32_t* noise_mode, int32_t* host_total);
/* Phases 4-5 (:866-897): gather the six raw tensors into N_out slots; split children get scale / 1.6 and
 * +-0.1*mean(exp(scale))*noise, clone copies 0.01*noise.  base_noise f32[N_out,3] stands for
 * MLXRandom.normal([totalOutput,3]); NULL draws it from a counter-based generator keyed by (seed, slot). */
GSB_API int gsb_densify_apply(gsb_ctx*, int32_t N_out, const int32_t* gather_indices, const int32_t* noise_mode,
                      const float* base_noise, uint64_t seed, const float* xyz, const float* f_dc, const float* f_rest,
                      const float* scales_log, const float* rot_raw, const float* opacity_logit, float* o_xyz, float* o_f_dc,
                      float* o_f_rest, float* o_scales_log, float* o_rot_raw, float* o_opacity_logit);
/* split_and_prune on the trainer's own tensors: classify with the accumulated gradient norms, rebuild parameters
 * (new Gaussian count), re-create the Adam state and reset the accumulation (GaussianTrainer.swift:1098-1109).
 * max_gaussians: densification is allowed while N < max_gaussians (:785).  base_noise: device f32[>= 2N,3] or NULL.
 * host_counts5 (may be NULL) = {keep, split, clone, prune, total}.  Synchronises.  GSB_ERR_STATE while peer mappings
 * (gsb_trainer_peers_import) are open. */
GSB_API int gsb_trainer_densify(gsb_ctx*, float grad_threshold, float max_scale, float min_opacity, int32_t max_gaussians,
                        uint64_t seed, const float* base_noise, int32_t* host_counts5);
/* current Gaussian count of the trainer and the number of iterations accumulated since the last reset */
GSB_API int gsb_trainer_count(gsb_ctx*, int32_t* host_N, int32_t* host_accum_steps);

/* ---- introspection for bench.py / profiles ------------------------------------------------- */
typedef struct gsb_stats {
    uint64_t kernel_launches;   /* CUDA kernels launched by this library since the last reset */
    uint64_t pairs_last_view;   /* M of the last view */
    uint64_t pairs_total;       /* sum of M since reset */
    uint64_t views;             /* views rendered since reset */
    uint64_t pair_capacity;     /* current capacity of the intersection buffers */
    double   stage_ms[16];      /* accumulated CUDA-event time per stage when timing is enabled */
    uint64_t stage_calls[16];
    uint64_t sb_pairs_last_view; /* (Gaussian, superblock) pairs of the last view: the only list that is radix-sorted */
} gsb_stats;
enum { GSB_STAGE_PROJECT_FWD = 0, GSB_STAGE_SCAN, GSB_STAGE_KEYGEN, GSB_STAGE_SORT, GSB_STAGE_TILE_LISTS,
       GSB_STAGE_RASTER_FWD, GSB_STAGE_LOSS, GSB_STAGE_RASTER_BWD, GSB_STAGE_PROJECT_BWD, GSB_STAGE_ADAM,
       GSB_STAGE_H2D, GSB_STAGE_DEPTH_SORT, GSB_STAGE_COUNT };
/* Sum of lastContrib over the image of the last gsb_render_forward = number of (pixel, Gaussian)
 * blend evaluations E of that view (the unit of the raster rooflines).  Synchronises. */
GSB_API int gsb_last_contrib_sum(gsb_ctx*, uint64_t* host_out);
/* geometry of the two-level tile lists: superblock size in tiles, superblock count, onesweep passes of the level-1 sort */
GSB_API int gsb_tile_list_info(gsb_ctx*, int32_t* sb_w, int32_t* sb_h, int32_t* num_superblocks, int32_t* sort_passes);
GSB_API int gsb_stats_reset(gsb_ctx*);
GSB_API int gsb_stats_get(gsb_ctx*, gsb_stats* host_out);
GSB_API int gsb_enable_stage_timing(gsb_ctx*, int32_t on);   /* brackets every stage with CUDA-event pairs; no syncs until stats are read */
GSB_API const char* gsb_stage_name(int32_t stage);
/* the section of the reference's IntervalProfiler report (GaussianTrainer.swift:122-241) the stage belongs to:
 * "train.forward", "train.loss.ssim", "bwd.globalTileComposite", "bwd.projectionScreenFused",
 * "train.optimizer.applySingle", "train.makeTrainStepInputs" */
GSB_API const char* gsb_stage_section(int32_t stage);
/* Counter bumped whenever the context's tile lists are rebuilt (gsb_bin, gsb_render_forward, the trainer).  A caller
 * that differentiates gsb_raster_fwd later (autograd node, the reference's CustomFunction closure,
 * GaussianRenderer.swift:119-122,150-184) stamps it at forward time and must find it unchanged at backward time:
 * gsb_raster_bwd differentiates with the binning the context holds NOW. */
GSB_API int gsb_bin_generation(gsb_ctx*, uint64_t* host_out);

#ifdef __cplusplus
}
#endif
#endif /* GSB_H_ */
