// kernels.h — internal launcher declarations (one per .cu), all enqueue on the given stream.
#pragma once
#include "common.cuh"

namespace gsb {

// superblock = SBW x SBH tiles, walked by L2_WARPS warps (tilelists.cu); SBW * SBH <= 32 (one ballot mask)
#ifndef GSB_SBW
#define GSB_SBW 4
#endif
#ifndef GSB_SBH
#define GSB_SBH 2
#endif
#ifndef GSB_L2_WARPS
#define GSB_L2_WARPS 8
#endif
// a superblock's slice is walked by L2_SPLIT CTAs of L2_WARPS warps = L2_SLICES warp slices: the kernel lasts as long as the
// heaviest superblock (3x the mean at C3), so the heavy ones must spread over several CTAs
#ifndef GSB_L2_SPLIT
#define GSB_L2_SPLIT 4
#endif
constexpr int SBW = GSB_SBW, SBH = GSB_SBH, L2_WARPS = GSB_L2_WARPS, SB_TILES = SBW * SBH;
constexpr int L2_SPLIT = GSB_L2_SPLIT, L2_SLICES = L2_WARPS * L2_SPLIT;
static_assert(SB_TILES <= 32 && L2_WARPS * 32 <= 1024, "superblock geometry");

// ---- project.cu --------------------------------------------------------------------------------
cudaError_t launch_activate_fwd(cudaStream_t st, int N, int K, const float* f_dc, const float* f_rest,
                                const float* scales_log, const float* rot_raw, const float* op_logit, float* shs,
                                float* scales, float* rotations, float* opacity);
cudaError_t launch_activate_bwd(cudaStream_t st, int N, int K, const float* scales_log, const float* rot_raw,
                                const float* op_logit, const float* g_shs, const float* g_scales, const float* g_rot,
                                const float* g_op, float* g_f_dc, float* g_f_rest, float* g_scales_log, float* g_rot_raw,
                                float* g_op_logit);
cudaError_t launch_project_fwd_api(cudaStream_t st, int N, const ViewParams& vp, const float* scales,
                                   const float* rotations, const float* means3d, const float* shs, float* means2d,
                                   float* depths, float* color, float* cov2d, float* conic, float* radii, float* rectMin,
                                   float* rectMax);
cudaError_t launch_project_bwd_api(cudaStream_t st, int N, const ViewParams& vp, const float* scales,
                                   const float* rotations, const float* means3d, const float* shs,
                                   const float* cotDepths, const float* cotMeans2d, const float* cotCov2d,
                                   const float* cotColor, const float* cotConic, float* gScales, float* gRot,
                                   float* gMeans3d, float* gShs, float* gCam);
cudaError_t launch_project_fused_fwd(cudaStream_t st, int N, const ViewParams& vp, const float* xyz, const float* f_dc,
                                     const float* f_rest, const float* scales_log, const float* rot_raw,
                                     const float* op_logit, float* rec, uint2* tile_rects, uint32_t* touched,
                                     uint32_t* depth_keys, float* radii_out, uint8_t* vis_out, int part = 3);   // 1 geometry, 2 colour, 3 both
size_t project_fused_smem_bytes(int K);
cudaError_t launch_project_fused_bwd(cudaStream_t st, int N, const ViewParams& vp, const float* xyz, const float* f_dc,
                                     const float* f_rest, const float* scales_log, const float* rot_raw,
                                     const float* op_logit, const float* grad_rec, float* g_xyz, float* g_f_dc,
                                     float* g_f_rest, float* g_scales, float* g_rot, float* g_op, int accumulate);

// ---- binning.cu --------------------------------------------------------------------------------
// Tile rectangles + counts + depth sort keys from reference-layout rect/radii/depths (K3).
cudaError_t launch_count_tiles(cudaStream_t st, int N, const ViewParams& vp, const float* rectMin, const float* rectMax,
                               const float* radii, const float* depths, uint2* tile_rects, uint32_t* touched,
                               uint32_t* depth_keys);
// Exclusive scan of in[perm[i]] (perm = perm_sel && *perm_sel ? perm1 : perm0; NULL perm0 = identity)
// → offsets[N]; *total (device) = M.  scan_ws: >= scan_ws_bytes(N).
size_t scan_ws_bytes(int N);
cudaError_t launch_exclusive_scan(cudaStream_t st, int N, const uint32_t* in, const uint32_t* perm0, const uint32_t* perm1,
                                  const uint32_t* perm_sel, uint32_t* offsets, uint32_t* total, void* scan_ws);
// The level-1 scan: offsets[i] = exclusive prefix, in DEPTH order (perm), of the superblocks the tile rect of Gaussian
// perm[i] touches; *total_sb_pairs = their sum; *total_pairs = sum of tiles-touched = M.
cudaError_t launch_sb_scan(cudaStream_t st, int N, const uint2* tile_rects, const uint32_t* touched, const uint32_t* perm0,
                           const uint32_t* perm1, const uint32_t* perm_sel, uint32_t* offsets, uint32_t* total_sb_pairs,
                           uint32_t* total_pairs, void* scan_ws);
// K4 in depth order: cell-id keys + Gaussian indices, one thread per output pair (coalesced).  A cell is cw x ch
// tiles (1 x 1 = the reference's tiles, SBW x SBH = the superblocks of tilelists.cu); cellGridW = cells per row.
cudaError_t launch_generate_keys(cudaStream_t st, int N, int cellGridW, int cw, int ch, const uint2* tile_rects,
                                 const uint32_t* offsets, const uint32_t* perm0, const uint32_t* perm1,
                                 const uint32_t* perm_sel, uint32_t* keys, uint32_t* vals, uint32_t capacity,
                                 const uint32_t* total, uint32_t* overflow_flag);
// K4 in the reference's emission order with 64-bit (tile << 32 | depth bits) keys (parity API only).
cudaError_t launch_generate_keys_ref(cudaStream_t st, int N, const ViewParams& vp, const uint2* tile_rects,
                                     const uint32_t* offsets, const float* depth_ptr, int depth_stride, uint64_t* keys,
                                     uint32_t* vals, uint32_t capacity);

// Onesweep LSD radix sort of (key, u32 value) pairs over key bits [0, end_bit).
struct SortPlan {
    uint32_t capacity = 0;      // max number of pairs
    uint32_t max_tiles = 0;
    uint32_t passes = 0;
    uint32_t end_bit = 0;
    size_t ws_bytes = 0;        // workspace: histograms, look-back state, counters, control block
};
SortPlan sort_plan(uint32_t capacity, uint32_t end_bit);
uint32_t sort_tile_items();   // pairs per onesweep CTA
// keys/vals double buffers; count read from device *d_count (clamped to capacity).  On return the
// sorted data sits in buffer index *d_result_buf (device u32 inside the workspace control block,
// pointer returned via result_buf_ptr).  sort32: iota != 0 synthesises the payload (= element index).
cudaError_t launch_onesweep_sort(cudaStream_t st, const SortPlan& plan, uint64_t* keys0, uint64_t* keys1, uint32_t* vals0,
                                 uint32_t* vals1, const uint32_t* d_count, void* ws, const uint32_t** result_buf_ptr,
                                 int* launches);
cudaError_t launch_onesweep_sort32(cudaStream_t st, const SortPlan& plan, uint32_t* keys0, uint32_t* keys1, uint32_t* vals0,
                                   uint32_t* vals1, int iota, const uint32_t* d_count, void* ws,
                                   const uint32_t** result_buf_ptr, int* launches);
// CUB baselines (checked against, never the product path unless GSB_FLAG_SORT_CUB): sort count pairs
// (host-known) from keys0/vals0 into keys1/vals1.
cudaError_t cub_sort_pairs(cudaStream_t st, uint64_t* keys0, uint64_t* keys1, uint32_t* vals0, uint32_t* vals1,
                           uint32_t count, uint32_t end_bit, void* tmp, size_t tmp_bytes, size_t* tmp_needed);
cudaError_t cub_sort_pairs32(cudaStream_t st, uint32_t* keys0, uint32_t* keys1, uint32_t* vals0, uint32_t* vals1,
                             uint32_t count, uint32_t end_bit, void* tmp, size_t tmp_bytes, size_t* tmp_needed);
cudaError_t launch_iota(cudaStream_t st, uint32_t n, uint32_t* v);

// K6 on a sorted key list: ranges[key] = (first, last + 1), (0, 0) for absent keys
cudaError_t launch_key_ranges(cudaStream_t st, const uint32_t* keys0, const uint32_t* keys1, const uint32_t* d_result_buf,
                              const uint32_t* d_count, uint32_t capacity, uint32_t* ranges, int numKeys);
cudaError_t launch_tile_counts(cudaStream_t st, int numTiles, const uint32_t* tile_ranges, uint32_t* tile_counts);
// packed[N,11] (reference layout) → rec[N,12]
cudaError_t launch_packed_to_rec(cudaStream_t st, int N, const float* packed, float* rec);
cudaError_t launch_rec_to_packed(cudaStream_t st, int N, const float* rec, float* packed);
// split / merge helpers for the parity API
cudaError_t launch_split_keys(cudaStream_t st, uint32_t M, const uint64_t* keys, uint32_t* hi, uint32_t* lo);
cudaError_t launch_merge_keys(cudaStream_t st, uint32_t M, const uint32_t* hi, const uint32_t* lo, uint32_t hi_mask,
                              uint64_t* keys);

// ---- tilelists.cu ------------------------------------------------------------------------------
cudaError_t launch_l2_count(cudaStream_t st, int numSB, int sbGridW, int gridW, int gridH, const uint32_t* sb_ranges, const uint32_t* vals0,
                            const uint32_t* vals1, const uint32_t* d_result_buf, const uint2* tile_rects, uint32_t* slice_counts,
                            uint32_t* tile_counts);
cudaError_t launch_l2_fill(cudaStream_t st, int numSB, int sbGridW, const uint32_t* sb_ranges, const uint32_t* vals0,
                           const uint32_t* vals1, const uint32_t* d_result_buf, const uint2* tile_rects, const uint32_t* slice_base,
                           uint32_t* list, uint32_t capacity);
constexpr int TO_BUCKETS = 256;   // length buckets of the heavy-first tile order
// CSR ranges + slice bases, and (order_ws != NULL: 2 * TO_BUCKETS words of scratch) the heavy-first tile order
cudaError_t launch_tile_bases(cudaStream_t st, int gridW, int gridH, int sbGridW, const uint32_t* tile_counts, const uint32_t* tile_starts,
                              const uint32_t* slice_counts, uint32_t* slice_base, uint32_t* tile_ranges, uint32_t* order_ws, uint32_t* order);
cudaError_t launch_expand_sorted_keys(cudaStream_t st, uint32_t M, int numTiles, const uint32_t* tile_starts, const uint32_t* list,
                                      const float* depth_ptr, int depth_stride, uint32_t* hi, uint32_t* lo);

// ---- raster.cu ---------------------------------------------------------------------------------
// Blend-state checkpoints written by the forward every 256 Gaussians of a 16x16 block that is still running, so that
// the backward can differentiate the segments of a long list as independent, bounded work items.  A pool of slots
// handed out by a device counter; a null `state` disables checkpointing (forward-only rendering) and an empty or
// exhausted pool only makes the backward fall back to longer items.
struct RasterCkpt {
    float4* state = nullptr;      // [capacity][256] colour summed over one segment + transmittance at its end, 16x16 pixels
    float* depth = nullptr;       // [capacity][256] depth summed over one segment (written only when depth is requested)
    uint2* header = nullptr;      // [capacity] (block work id, checkpoint index | 0xffffffff = final partial segment sums)
    uint32_t* table = nullptr;    // [blocks][17] slot of checkpoint c (c < written), [16] = slot of the final segment sums
    uint32_t* count = nullptr;    // slots requested by the last forward (may exceed capacity)
    uint32_t* written = nullptr;  // [blocks] checkpoints each block managed to write
    uint32_t capacity = 0;
};
int raster_block_count(const ViewParams& vp);
// rec = record table [N,12]; (vals0|vals1 selected by *d_result_buf) = Gaussian indices in (tile, depth) order
cudaError_t launch_raster_fwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges,
                              const uint32_t* tile_order, const float* rec, const uint32_t* vals0, const uint32_t* vals1,
                              const uint32_t* d_result_buf,
                              float* out_color, float* out_depth, float* out_alpha, uint32_t* out_last, uint32_t* work_counter,
                              const RasterCkpt& ck);
cudaError_t launch_raster_bwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges,
                              const uint32_t* tile_order, const float* rec, const uint32_t* vals0, const uint32_t* vals1,
                              const uint32_t* d_result_buf,
                              const float* cot_color, const float* cot_depth, const float* cot_alpha,
                              const float* out_color, const float* out_depth, const float* out_alpha,
                              const uint32_t* last_contrib, float* grad_rec, uint32_t* work_counter, const RasterCkpt& ck);

cudaError_t launch_sum_u32(cudaStream_t st, size_t n, const uint32_t* v, unsigned long long* out);

// ---- loss.cu -----------------------------------------------------------------------------------
// separable SSIM (11x11, sigma 1.5, centre 5.5) forward with optional saved maps
cudaError_t launch_ssim_fwd(cudaStream_t st, int H, int W, int C, const float* img1, const float* img2, float* ssim_map,
                            float* mu1, float* mu2, float* s1, float* s2, float* s12);
// fused loss forward: writes the three "dSSIM/d(window statistic) x upstream" maps and accumulates
// sum|d| and sum(ssim) into partial[2] (double).
cudaError_t launch_loss_fwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            float upstream, float* mapA, float* mapB, float* mapC, double* partial);
// fused loss backward: cot_render = l1_scale*sign(render-target) + conv^T(maps); finalises the loss scalar.
cudaError_t launch_loss_bwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            const float* mapA, const float* mapB, const float* mapC, float l1_scale, float* cot_render);
cudaError_t launch_loss_finalize(cudaStream_t st, const double* partial, double inv_count, float lambda, float scale,
                                 float* loss_accum);
// masked-L1 depth supervision: cot_depth[P] = lambda * scale * mask * sign(depth - target) / max(sum mask, 1e-6) and
// loss_accum += lambda * scale * masked mean; partial2 = two doubles of scratch
cudaError_t launch_depth_loss(cudaStream_t st, size_t P, const float* depth, const float* target_depth, const uint8_t* mask,
                              float lambda_depth, float scale, float* cot_depth, double* partial2, float* loss_accum);
// generic ssim backward for the parity API (upstream map given)
cudaError_t launch_ssim_bwd_api(cudaStream_t st, int H, int W, int C, const float* grad_out, const float* img1,
                                const float* img2, float* mapA, float* mapB, float* mapC, float* grad_img1);

// ---- adam.cu -----------------------------------------------------------------------------------
struct AdamTensors {
    float* p[6];
    const float* g[6];
    float* m[6];
    float* v[6];
    long long count[6];
    float lr[6];
};
// Peer-memory data parallelism (adam.cu k_adam_peers): replica r's gradient / parameter copies and D1 accumulator, and the
// slice of Gaussians [g0, g1) this rank owns.  AdamTensors passed with it holds the LOCAL p / m / v base pointers of the
// six tensors and count[k] = floats of the owned slice of tensor k.
constexpr int GSB_MAX_PEERS = 8;
constexpr int GSB_MAX_CHUNKS = 8;
// Flags of the device-side step protocol (gsb_trainer_step_peers), one block per replica, written by the OTHER replicas
// through peer memory.  Every word holds the id of the last step for which the event happened (monotone, compared >=).
struct PeerSync {
    uint32_t grads_ready[GSB_MAX_CHUNKS][GSB_MAX_PEERS];    // [c][r]: replica r's gradients of Gaussian chunk c are complete
    uint32_t params_ready[GSB_MAX_CHUNKS][GSB_MAX_PEERS];   // [c][r]: owner r has written its part of chunk c into THIS replica
    uint32_t done[GSB_MAX_CHUNKS];                          // local: finished CTAs of the exchange kernel of chunk c
    uint32_t error;                                         // local: a bounded wait ran out (the step's results are invalid)
    uint32_t pad[7];
};
struct PeerStepSync {                      // what one exchange kernel launch waits for and announces
    const uint32_t* wait_flags = nullptr;  // local grads_ready[c][0..world): all must reach `step` before any gradient is read
    uint32_t* done = nullptr;              // local CTA completion counter (self-resetting)
    uint32_t* announce[GSB_MAX_PEERS] = {};   // params_ready[c][rank] of every replica: set to `step` by the last CTA
    uint32_t* error = nullptr;
    uint32_t step = 0;
};
struct AdamPeers {
    const float* grads[GSB_MAX_PEERS];   // base of the gradient copy (six tensors, padded) of every replica
    float* params[GSB_MAX_PEERS];        // base of the parameter copy of every replica
    float* accum[GSB_MAX_PEERS];         // D1 accumulators (accum[0] == nullptr: skip D1)
    long long tensor_off[6];             // float offset of tensor k inside one copy
    long long first[6];                  // first float of the owned slice inside tensor k (multiple of 4)
    long long g0, g1;                    // owned Gaussians
    int world, rank;
    int accum_all;                       // accum[r] is valid for EVERY replica (peer-mapped), not only for accum[rank]
};
// sync == nullptr: the caller brackets the launch with two barriers (gsb_trainer_apply_peers); otherwise the kernel
// itself waits for the replicas' gradients and announces its parameter stores (gsb_trainer_step_peers)
cudaError_t launch_adam_peers(cudaStream_t st, const AdamTensors& t, const AdamPeers& pr, float beta1, float beta2, float eps,
                              float gscale, const PeerStepSync* sync, int blocks, int* launches);
// NVLS variant: mc_grads / mc_params = multicast addresses of the replicas' gradient / parameter blocks (symmetric memory);
// pr supplies the owned slice, the tensor offsets and the D1 accumulators (accum[r] of every replica when the slabs are
// peer-mapped: D1 is then computed for the owned slice only and stored to every replica; otherwise accum[rank] alone and
// every replica computes all of D1 from the switch-reduced gradient).
cudaError_t launch_adam_multicast(cudaStream_t st, const AdamTensors& t, const AdamPeers& pr, const float* mc_grads, float* mc_params,
                                  float beta1, float beta2, float eps, float gscale, int N, const PeerStepSync* sync, int blocks, int* launches);
// system-scope release of `value` into up to GSB_MAX_PEERS flag words (after everything queued before it on the stream)
struct PeerFlagList {
    uint32_t* dst[GSB_MAX_PEERS] = {};
    int n = 0;
};
cudaError_t launch_peer_signal(cudaStream_t st, const PeerFlagList& flags, uint32_t value);
// bounded wait until the rows x cols words flags[row * row_stride + col] have all reached `value` (ids modulo 2^32)
cudaError_t launch_peer_wait(cudaStream_t st, const uint32_t* flags, int rows, int cols, int row_stride, uint32_t value, uint32_t* error);
// gscale multiplies every gradient before use (1.0 on the training path); launches (may be NULL) is
// incremented by the number of kernels launched.
cudaError_t launch_adam(cudaStream_t st, const AdamTensors& t, float beta1, float beta2, float eps, float gscale, int N,
                        float* grad_norm_accum, const uint32_t* skip_flag, int* launches);

// ---- densify.cu --------------------------------------------------------------------------------
cudaError_t launch_densify_classify(cudaStream_t st, int N, const float* grad_accum, float denom, const float* scales_log,
                                    const float* opacity_logit, float grad_threshold, float max_scale, float min_opacity,
                                    int allow_densify, int* actions, int* counts, uint32_t* stats4);
cudaError_t launch_densify_map(cudaStream_t st, int N, const int* actions, const uint32_t* offsets, uint32_t capacity, int* gather,
                               int* noise_mode);
cudaError_t launch_densify_apply(cudaStream_t st, int Nout, int K, const int* gather, const int* noise_mode, const float* base_noise,
                                 uint64_t seed, const float* const* in6, float* const* out6);

}  // namespace gsb
