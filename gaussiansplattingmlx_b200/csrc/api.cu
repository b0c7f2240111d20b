// api.cu — the C ABI of libgsb.so (include/gsb.h): context, buffer management, stage sequencing.
//
// Sequencing mirrors the reference's host glue:
//   GaussianRenderer.forwardWithCameraParams → render → buildGlobalTileSliceInfo
//                                       (Trainer/GaussianRenderer.swift:823-880, 769-821, 333-490)
//   VJPs                                (Trainer/GaussianRenderer.swift:187-226, 605-701)
//   lossFn / Adam loop                  (Trainer/GaussianTrainer.swift:634-716, 1060-1086)
// but with no host synchronisation on the data path: the pair count M stays on the device, kernels
// are launched for the buffer capacity, and M is read back asynchronously only to detect overflow
// (the view is then redone with larger buffers).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>

#include <nvtx3/nvToolsExt.h>   // header-only: ranges are no-ops unless a profiler injects itself

#include "kernels.h"

namespace gsb {

static thread_local std::string g_create_error;

// view working sets: with 3 the front of view b+2 can start as soon as the raster backward of view b-1 is done,
// so the front stream never waits for the view that is being rasterised
#ifndef GSB_VIEW_SETS_N
#define GSB_VIEW_SETS_N 3
#endif
constexpr int GSB_VIEW_SETS = GSB_VIEW_SETS_N;
constexpr int GSB_FRONT_AHEAD = GSB_VIEW_SETS - 1;

struct Ctx {
    gsb_config cfg{};
    cudaStream_t stream = nullptr;       // work stream (caller's, or own_stream)
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of targets
    std::string err;
    int gridW = 0, gridH = 0, numTiles = 0, tileBits = 1, P = 0;
    int sbGridW = 0, sbGridH = 0, numSB = 0, sbBits = 1;   // superblocks of SBW x SBH tiles (tilelists.cu)

    // Per-view working set.  There are two: while the rasteriser / loss / backward kernels of view v run on the
    // work stream, the projection + binning of view v+1 run on the (higher-priority) front stream into the
    // other set (trainer path only; the single-view API always uses set `cur`).
    struct ViewBufs {
        // per Gaussian (capacity capN)
        float* rec = nullptr;          // [N,12]
        uint2* tile_rects = nullptr;   // [N]
        uint32_t* touched = nullptr;   // [N]
        uint32_t* offsets = nullptr;   // [N] exclusive scan of touched, in DEPTH order
        uint32_t* dkeys[2] = {nullptr, nullptr};   // [N] asuint(depth) sort keys (ping-pong)
        uint32_t* dvals[2] = {nullptr, nullptr};   // [N] Gaussian indices in depth order (ping-pong)
        void* dsort_ws = nullptr;
        const uint32_t* d_dresult_buf = nullptr;   // which dvals buffer holds the depth order
        uint32_t* d_nvalue = nullptr;              // device copy of N (count of the depth sort)
        void* scan_ws = nullptr;
        const float* depth_src = nullptr;          // depth of Gaussian g at depth_src[g * depth_src_stride]
        int depth_src_stride = 1;
        // level 1: (superblock id, Gaussian) pairs (capacity capL1), ping-pong for the sort
        uint32_t* keys[2] = {nullptr, nullptr};
        uint32_t* vals[2] = {nullptr, nullptr};
        void* sort_ws = nullptr;
        const uint32_t* d_result_buf = nullptr;  // device flag: which ping-pong buffer holds the sorted level-1 list
        // level 2 / per tile
        uint32_t* list = nullptr;         // [capM] Gaussian indices in (tile, depth, index) order = the tile lists
        uint32_t* sb_ranges = nullptr;    // [numSB,2] slice of every superblock in the sorted level-1 list
        uint32_t* slice_counts = nullptr; // [numSB, L2_SLICES warp slices, 8 tiles]
        uint32_t* slice_base = nullptr;   // [numSB, L2_SLICES warp slices, 8 tiles]
        uint32_t* tile_starts = nullptr;  // [numTiles+1] monotone CSR offsets
        uint32_t* tile_counts = nullptr;  // [numTiles]
        void* tile_scan_ws = nullptr;
        uint32_t* tile_ranges = nullptr;  // [numTiles,2] (reference convention: (0,0) for an empty tile)
        uint32_t* tile_order = nullptr;   // [numTiles] tile ids, longest list first
        uint32_t* order_ws = nullptr;     // [2 * TO_BUCKETS] bucket histogram + cursors of the heavy-first order
        // control words: [0] = M of the current view, [1] = overflow flag, [2] scratch, [3] raster work counter,
        //                [4] = level-1 pair count, [5] = list total after the tile scan
        uint32_t* d_ctl = nullptr;
        uint32_t* h_ctl = nullptr;        // pinned mirror
        cudaEvent_t ev_ctl = nullptr;     // fires when h_ctl[0] holds M
        cudaEvent_t ev_front = nullptr;   // projection + binning of this set finished (front stream)
        cudaEvent_t ev_back = nullptr;    // the work stream no longer reads this set
        uint32_t last_M = 0;
        uint32_t last_L1 = 0;
    } vb[GSB_VIEW_SETS];
    int cur = 0;                       // set used by the single-view API / holding the saved forward
    cudaStream_t front_stream = nullptr;                  // = front_streams[0]
    cudaStream_t front_streams[GSB_VIEW_SETS] = {};       // one per view set: the fronts of different views overlap each other
    cudaEvent_t ev_fork = nullptr;     // work stream -> front stream dependency at the start of a batch

    int capN = 0;
    SortPlan dplan;
    float* grad_rec = nullptr;         // [N,12] raster backward -> projection backward
    float* grad_rec2 = nullptr;        // second buffer (trainer: projection backward of view b overlaps the raster of b+1)
    cudaStream_t tail_stream = nullptr;
    cudaStream_t xchg_stream = nullptr;   // exchange kernels of gsb_trainer_step_peers (ordered against the rest by flags and events)
    cudaEvent_t ev_rb[2] = {nullptr, nullptr};   // raster backward into gradient-record buffer i finished (work stream)
    cudaEvent_t ev_pb[2] = {nullptr, nullptr};   // projection backward out of buffer i finished (tail stream)
    uint32_t* offsets_ref = nullptr;   // [N] scan in index order (parity API: reference emission order)
    float* act_tmp = nullptr;          // scratch for the reference-layout parity API ([N,12] floats)

    uint32_t capM = 0;                 // capacity of the tile lists (pairs)
    uint32_t capL1 = 0;                // capacity of the level-1 (superblock) pair buffers
    SortPlan plan;
    void* cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    uint64_t* dbg_keys = nullptr;  // unsorted copies for gsb_bin_read
    uint32_t* dbg_vals = nullptr;
    uint32_t dbg_cap = 0;
    uint32_t* d_zero = nullptr;              // constant 0 (CUB path result buffer selector = 1 → d_one)
    uint32_t* d_one = nullptr;

    // per pixel
    float* out_color = nullptr;       // saved forward outputs
    float* out_depth = nullptr;
    float* out_alpha = nullptr;
    uint32_t* out_last = nullptr;
    gsb::RasterCkpt ck;               // forward blend-state checkpoints for the segmented raster backward (raster.cu)
    bool ck_has_depth = false;        // the last fused forward also checkpointed the depth prefix
    bool ck_valid = false;            // the last fused forward wrote checkpoints at all (GSB_FLAG_NO_SEGMENTS clear)
    float* mapA = nullptr;            // SSIM backward maps, [P*3] each
    float* mapB = nullptr;
    float* mapC = nullptr;
    float* cot_render = nullptr;      // [P*3]
    float* cot_depth = nullptr;       // [P] depth-supervision cotangent (allocated on first use)
    double* partial = nullptr;        // [4]: L1 / SSIM sums, depth-loss sums
    float* loss_accum = nullptr;      // device scalar
    float* h_loss = nullptr;          // pinned

    // saved forward (one in flight, like the reference's closure-captured state)
    struct Saved {
        bool valid = false;
        int N = 0;
        const float *xyz = nullptr, *f_dc = nullptr, *f_rest = nullptr, *scales_log = nullptr, *rot_raw = nullptr,
                    *op_logit = nullptr;
        ViewParams vp{};
    } saved;
    uint64_t bin_gen = 0;             // bumped whenever the tile lists of the single-view API are rebuilt (gsb_bin_generation)
    bool bin_valid = false;           // gsb_bin ran (parity API)
    ViewParams bin_vp{};

    // trainer state
    int tN = 0;
    int t_accum_steps = 0;           // iterations accumulated into t_accum since the last reset (denomGradAccumulation)
    float* t_block = nullptr;         // params | grads | m | v, each 6 tensors, 128-byte aligned segments (= t_slab[t_cur])
    // Two capacity-sized slabs (block + accumulator): densification gathers from the current one into the other, so a
    // clone/split/prune every 100 iterations costs no cudaMalloc/cudaFree unless the capacity itself has to grow.
    float* t_slab[2] = {nullptr, nullptr};
    float* t_accum_slab[2] = {nullptr, nullptr};
    int t_cur = 0;
    int t_cap = 0;                    // Gaussians each slab can hold
    float* t_p[6]{}; float* t_g[6]{}; float* t_m[6]{}; float* t_v[6]{};
    long long t_count[6]{};
    size_t t_floats = 0;              // padded floats of one copy of the six tensors
    float* t_accum = nullptr;         // D1 accumulator [N]
    // peer-memory data parallelism (gsb_trainer_peers_*): the other replicas' slab / accumulator mapped through CUDA IPC
    int peer_world = 0, peer_rank = 0;
    float* peer_block[gsb::GSB_MAX_PEERS] = {};    // base of replica r's trainer slab (own pointer at [peer_rank])
    float* peer_accum[gsb::GSB_MAX_PEERS] = {};
    bool peer_opened[gsb::GSB_MAX_PEERS] = {};
    // device-side step protocol (gsb_trainer_step_peers): this replica's flag block (written by the others), the others'
    // blocks, the id of the last step and the events "projection backward of Gaussian chunk c is done"
    gsb::PeerSync* t_sync = nullptr;
    gsb::PeerSync* peer_sync[gsb::GSB_MAX_PEERS] = {};
    uint32_t peer_step_id = 0;
    uint32_t peer_wait_step = 0;        // != 0: the next batch must first wait for this step's "parameters written" flags
    int peer_wait_chunks = 0;           // flag rows of that step: chunks, or 2 x chunks (geometry rows, then SH rows) when phased
    bool peer_wait_phased = false;
    int peer_phased = 1;                // exchange the geometry tensors first, the SH tensors behind the next step's binning
    cudaEvent_t ev_sh = nullptr;        // the SH rows of the previous step have been announced (recorded on the tail stream)
    int peer_chunks = 2;
    int peer_blocks = 0, mc_blocks = 0;   // CTAs of the exchange kernels (0 = default), gsb_trainer_peers_tune
    cudaEvent_t ev_chunk[gsb::GSB_MAX_CHUNKS] = {};
    uint32_t* h_peer_error = nullptr;   // pinned mirror of t_sync->error
    // symmetric-memory / NVLS variant (gsb_trainer_attach_symmetric): parameters and gradients live in caller-owned
    // symmetric buffers, mc_* are the multicast addresses of the same buffers on all replicas
    bool sym = false;
    int sym_world = 0, sym_rank = 0;
    float* sym_params = nullptr; float* sym_grads = nullptr; float* mc_params = nullptr; float* mc_grads = nullptr;
    float* t_target[2] = {nullptr, nullptr};
    float* t_dtarget[2] = {nullptr, nullptr};     // depth targets / masks of host-resident depth supervision (first use)
    uint8_t* t_dmask[2] = {nullptr, nullptr};
    cudaEvent_t t_target_ready[2] = {nullptr, nullptr};
    cudaEvent_t t_target_free[2] = {nullptr, nullptr};

    // stats.  Stage timing records CUDA-event pairs on the work stream WITHOUT synchronising; the
    // pairs are resolved (one stream sync) when the statistics are read.
    gsb_stats stats{};
    bool timing = false;
    struct StageEv { cudaEvent_t a = nullptr, b = nullptr; int stage = 0; };
    std::vector<StageEv> ev_pool;
    size_t ev_used = 0;
};

static void sync_all_streams(Ctx* c);
static void resolve_stage_events(Ctx* c)
{
    if (c->ev_used == 0) return;
    sync_all_streams(c);
    for (size_t i = 0; i < c->ev_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_pool[i].a, c->ev_pool[i].b) == cudaSuccess) {
            c->stats.stage_ms[c->ev_pool[i].stage] += ms;
            c->stats.stage_calls[c->ev_pool[i].stage] += 1;
        }
    }
    c->ev_used = 0;
}

void set_error(Ctx* ctx, const std::string& msg)
{
    if (ctx) ctx->err = msg; else g_create_error = msg;
}

static const char* kStageNames[GSB_STAGE_COUNT] = {"project_fwd", "scan", "keygen", "sort", "tile_lists", "raster_fwd",
                                                   "loss", "raster_bwd", "project_bwd", "adam", "h2d", "depth_sort"};
// The section of the reference's IntervalProfiler report (GaussianTrainer.swift:122-241; names at :652-716,1060-1086 and
// GaussianRenderer.swift:187-226,605-701) each stage belongs to, so that a trace of a Swift shim over this library lines
// up with the reference's own "[Profile]" report.  The L1 term ("train.loss.l1") is fused into the SSIM kernels.
static const char* kStageSections[GSB_STAGE_COUNT] = {"train.forward", "train.forward", "train.forward", "train.forward", "train.forward",
                                                      "train.forward", "train.loss.ssim", "bwd.globalTileComposite",
                                                      "bwd.projectionScreenFused", "train.optimizer.applySingle",
                                                      "train.makeTrainStepInputs", "train.forward"};
static const char* kStageNvtx[GSB_STAGE_COUNT] = {
    "train.forward/project_fwd", "train.forward/scan", "train.forward/keygen", "train.forward/sort", "train.forward/tile_lists",
    "train.forward/raster_fwd", "train.loss.ssim/loss", "bwd.globalTileComposite/raster_bwd", "bwd.projectionScreenFused/project_bwd",
    "train.optimizer.applySingle/adam", "train.makeTrainStepInputs/h2d", "train.forward/depth_sort"};

struct StageTimer {
    Ctx* c; int stage; bool on; cudaStream_t st; bool nvtx;
    StageTimer(Ctx* ctx, int s, cudaStream_t stream = nullptr)
        : c(ctx), stage(s), on(ctx->timing), st(stream ? stream : ctx->stream), nvtx((ctx->cfg.flags & GSB_FLAG_NVTX) != 0)
    {
        if (nvtx) nvtxRangePushA(kStageNvtx[stage]);   // host range around the enqueue; profilers project it onto the kernels
        if (!on) return;
        if (c->ev_used >= 16384) resolve_stage_events(c);
        if (c->ev_used == c->ev_pool.size()) {
            Ctx::StageEv e;
            if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) { on = false; return; }
            c->ev_pool.push_back(e);
        }
        c->ev_pool[c->ev_used].stage = stage;
        cudaEventRecord(c->ev_pool[c->ev_used].a, st);
    }
    ~StageTimer()
    {
        if (on) {
            cudaEventRecord(c->ev_pool[c->ev_used].b, st);
            c->ev_used += 1;
        }
        if (nvtx) nvtxRangePop();
    }
};

#define GSB_REQUIRE(ctx, cond, msg)                 \
    do {                                            \
        if (!(cond)) {                              \
            gsb::set_error(ctx, msg);               \
            return GSB_ERR_INVALID;                 \
        }                                           \
    } while (0)

template <class T>
static cudaError_t dev_alloc(T** p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}
template <class T>
static void dev_free(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

static int tile_bit_count(int numTiles)  // GaussianRenderer.swift:246-255
{
    if (numTiles <= 1) return 1;
    int v = numTiles - 1, bits = 0;
    while (v > 0) { ++bits; v >>= 1; }
    return bits;
}

static ViewParams make_view(const Ctx* c, const gsb_camera* cam)
{
    ViewParams vp{};
    memcpy(vp.V, cam->view, sizeof(vp.V));
    memcpy(vp.P, cam->proj, sizeof(vp.P));
    memcpy(vp.cam, cam->cam_center, sizeof(vp.cam));
    vp.fovX = cam->fov_x; vp.fovY = cam->fov_y; vp.focalX = cam->focal_x; vp.focalY = cam->focal_y;
    vp.tanHalfX = tanf(cam->fov_x * 0.5f);   // gaussian_projection_screen_shared.slang:200-201, host libm
    vp.tanHalfY = tanf(cam->fov_y * 0.5f);
    vp.imageW = (float)c->cfg.width; vp.imageH = (float)c->cfg.height;
    vp.W = c->cfg.width; vp.H = c->cfg.height; vp.tileW = c->cfg.tile_w; vp.tileH = c->cfg.tile_h;
    vp.gridW = c->gridW; vp.gridH = c->gridH;
    vp.degree = c->cfg.sh_degree; vp.K = c->cfg.sh_coeffs;
    vp.coeffCount = std::min((c->cfg.sh_degree + 1) * (c->cfg.sh_degree + 1), 25);
    vp.whiteBg = c->cfg.white_background;
    return vp;
}
static ViewParams make_view_nocam(const Ctx* c)
{
    gsb_camera cam{};
    return make_view(c, &cam);
}

static void sync_all_streams(Ctx* c)
{
    cudaStreamSynchronize(c->stream);
    if (c->xchg_stream) cudaStreamSynchronize(c->xchg_stream);
    for (cudaStream_t fs : c->front_streams) if (fs) cudaStreamSynchronize(fs);
    if (c->tail_stream) cudaStreamSynchronize(c->tail_stream);
}

static int ensure_gaussians(Ctx* c, int N)
{
    if (N <= c->capN) return GSB_OK;
    if ((uint32_t)N > gsb::REC_IDX_MASK) {   // the two top bits of a record's index word are flags (common.cuh)
        set_error(c, "more than 2^30 Gaussians are not supported");
        return GSB_ERR_INVALID;
    }
    if (c->cfg.max_gaussians > 0 && N > c->cfg.max_gaussians) {
        set_error(c, "N exceeds gsb_config.max_gaussians");
        return GSB_ERR_INVALID;
    }
    sync_all_streams(c);
    // max_gaussians == 0: grow on demand, geometrically (densification raises N every 100 iterations)
    const int cap = c->cfg.max_gaussians > 0 ? c->cfg.max_gaussians : std::max(std::max(N, 1024), c->capN + c->capN / 2);
    dev_free(c->grad_rec); dev_free(c->grad_rec2); dev_free(c->act_tmp); dev_free(c->offsets_ref);
    GSB_CUDA_CHECK(c, dev_alloc(&c->grad_rec, (size_t)cap * REC_FLOATS));
    GSB_CUDA_CHECK(c, dev_alloc(&c->grad_rec2, (size_t)cap * REC_FLOATS));
    GSB_CUDA_CHECK(c, dev_alloc(&c->act_tmp, (size_t)cap * REC_FLOATS));
    GSB_CUDA_CHECK(c, dev_alloc(&c->offsets_ref, (size_t)cap));
    c->dplan = sort_plan((uint32_t)cap, 32u);
    for (Ctx::ViewBufs& v : c->vb) {
        dev_free(v.rec); dev_free(v.tile_rects); dev_free(v.touched); dev_free(v.offsets);
        for (int i = 0; i < 2; ++i) { dev_free(v.dkeys[i]); dev_free(v.dvals[i]); }
        if (v.dsort_ws) { cudaFree(v.dsort_ws); v.dsort_ws = nullptr; }
        if (v.scan_ws) { cudaFree(v.scan_ws); v.scan_ws = nullptr; }
        GSB_CUDA_CHECK(c, dev_alloc(&v.rec, (size_t)cap * REC_FLOATS));
        GSB_CUDA_CHECK(c, dev_alloc(&v.tile_rects, (size_t)cap));
        GSB_CUDA_CHECK(c, dev_alloc(&v.touched, (size_t)cap));
        GSB_CUDA_CHECK(c, dev_alloc(&v.offsets, (size_t)cap));
        GSB_CUDA_CHECK(c, cudaMalloc(&v.scan_ws, scan_ws_bytes(cap)));
        for (int i = 0; i < 2; ++i) {
            GSB_CUDA_CHECK(c, dev_alloc(&v.dkeys[i], (size_t)cap));
            GSB_CUDA_CHECK(c, dev_alloc(&v.dvals[i], (size_t)cap));
        }
        GSB_CUDA_CHECK(c, cudaMalloc(&v.dsort_ws, c->dplan.ws_bytes));
    }
    c->capN = cap;
    return GSB_OK;
}

static int ensure_pairs(Ctx* c, uint64_t M, uint64_t L1)
{
    if (M <= c->capM && L1 <= c->capL1) return GSB_OK;
    if (M > 0xfffffff0ull / 3ull) {  // u32 offsets (the reference's cumsum is u32 too)
        set_error(c, "intersection list exceeds the 32-bit index range");
        return GSB_ERR_CAPACITY;
    }
    sync_all_streams(c);
    if (c->cub_tmp) { cudaFree(c->cub_tmp); c->cub_tmp = nullptr; c->cub_tmp_bytes = 0; }
    auto grow = [](uint64_t need) { uint64_t cap = std::max<uint64_t>(need + need / 4, 1u << 16); return (cap + 4095) & ~4095ull; };
    if (M > c->capM) {
        const uint64_t cap = grow(M);
        for (Ctx::ViewBufs& v : c->vb) {
            dev_free(v.list);
            GSB_CUDA_CHECK(c, dev_alloc(&v.list, (size_t)cap));
        }
        c->capM = (uint32_t)cap;
        c->stats.pair_capacity = cap;
    }
    if (L1 > c->capL1) {
        const uint64_t cap = grow(L1);
        c->plan = sort_plan((uint32_t)cap, (uint32_t)c->sbBits);
        for (Ctx::ViewBufs& v : c->vb) {
            for (int i = 0; i < 2; ++i) { dev_free(v.keys[i]); dev_free(v.vals[i]); }
            if (v.sort_ws) { cudaFree(v.sort_ws); v.sort_ws = nullptr; }
            for (int i = 0; i < 2; ++i) {
                GSB_CUDA_CHECK(c, dev_alloc(&v.keys[i], (size_t)cap));
                GSB_CUDA_CHECK(c, dev_alloc(&v.vals[i], (size_t)cap));
            }
            GSB_CUDA_CHECK(c, cudaMalloc(&v.sort_ws, c->plan.ws_bytes));
        }
        c->capL1 = (uint32_t)cap;
    }
    return GSB_OK;
}

static int cub_sort32(Ctx* c, cudaStream_t st, uint32_t* k0, uint32_t* k1, uint32_t* v0, uint32_t* v1, uint32_t count,
                      uint32_t end_bit)
{
    size_t need = 0;
    GSB_CUDA_CHECK(c, cub_sort_pairs32(st, k0, k1, v0, v1, count, end_bit, nullptr, 0, &need));
    if (need > c->cub_tmp_bytes) {
        sync_all_streams(c);
        if (c->cub_tmp) cudaFree(c->cub_tmp);
        c->cub_tmp = nullptr; c->cub_tmp_bytes = 0;
        GSB_CUDA_CHECK(c, cudaMalloc(&c->cub_tmp, need));
        c->cub_tmp_bytes = need;
    }
    if (count > 0) GSB_CUDA_CHECK(c, cub_sort_pairs32(st, k0, k1, v0, v1, count, end_bit, c->cub_tmp, c->cub_tmp_bytes, nullptr));
    return GSB_OK;
}

// depth sort → scan → keys → tile sort → ranges, enqueued on `st` into view set `v`.
// v.tile_rects / v.touched / v.dkeys[0] must be filled; depth of Gaussian g is depth_ptr[g * depth_stride].
// Does not wait for the GPU (except the very first call, which sizes the pair buffers, and the CUB baseline):
// the pair count M is copied to pinned memory behind v.ev_ctl and checked later by finish_binning.
static int enqueue_binning(Ctx* c, Ctx::ViewBufs& v, cudaStream_t st, int N, const ViewParams& vp, const float* depth_ptr,
                           int depth_stride)
{
    int launches = 0;
    const bool use_cub = (c->cfg.flags & GSB_FLAG_SORT_CUB) != 0;
    v.depth_src = depth_ptr;
    v.depth_src_stride = depth_stride;
    {
        StageTimer t(c, GSB_STAGE_DEPTH_SORT, st);
        // 1. Gaussians into depth order (stable: ties keep index order)
        const uint32_t n32 = (uint32_t)N;
        GSB_CUDA_CHECK(c, cudaMemsetAsync(v.d_ctl, 0, 2 * sizeof(uint32_t), st));
        GSB_CUDA_CHECK(c, cudaMemcpyAsync(v.d_nvalue, &n32, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (use_cub) {
            GSB_CUDA_CHECK(c, launch_iota(st, n32, v.dvals[0]));
            int rc = cub_sort32(c, st, v.dkeys[0], v.dkeys[1], v.dvals[0], v.dvals[1], n32, 32u);
            if (rc != GSB_OK) return rc;
            v.d_dresult_buf = N > 0 ? c->d_one : c->d_zero;
            launches += 2;
        } else {
            SortPlan dp = c->dplan;
            dp.capacity = n32;
            dp.max_tiles = (uint32_t)cdiv(N > 0 ? N : 1, sort_tile_items());
            GSB_CUDA_CHECK(c, launch_onesweep_sort32(st, dp, v.dkeys[0], v.dkeys[1], v.dvals[0], v.dvals[1], 1, v.d_nvalue,
                                                     v.dsort_ws, &v.d_dresult_buf, &launches));
        }
    }
    {
        StageTimer t(c, GSB_STAGE_SCAN, st);
        // 2. superblocks touched per Gaussian in depth order (+ M = sum of tiles touched), offsets of the level-1 pairs
        GSB_CUDA_CHECK(c, launch_sb_scan(st, N, v.tile_rects, v.touched, v.dvals[0], v.dvals[1], v.d_dresult_buf, v.offsets, &v.d_ctl[4],
                                         &v.d_ctl[0], v.scan_ws));
        launches += 1;
    }
    GSB_CUDA_CHECK(c, cudaMemcpyAsync(v.h_ctl, v.d_ctl, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GSB_CUDA_CHECK(c, cudaEventRecord(v.ev_ctl, st));
    if (c->capM == 0 || c->capL1 == 0) {  // first use: size the pair buffers from the actual counts
        GSB_CUDA_CHECK(c, cudaEventSynchronize(v.ev_ctl));
        int rc = ensure_pairs(c, std::max<uint64_t>(v.h_ctl[0], 1), std::max<uint64_t>(v.h_ctl[4], 1));
        if (rc != GSB_OK) return rc;
    }
    {
        StageTimer t(c, GSB_STAGE_KEYGEN, st);
        // 3. level-1 pairs (superblock id, Gaussian) in depth order
        GSB_CUDA_CHECK(c, launch_generate_keys(st, N, c->sbGridW, SBW, SBH, v.tile_rects, v.offsets, v.dvals[0], v.dvals[1],
                                               v.d_dresult_buf, v.keys[0], v.vals[0], c->capL1, &v.d_ctl[4], &v.d_ctl[1]));
        ++launches;
    }
    if (use_cub) GSB_CUDA_CHECK(c, cudaStreamSynchronize(st));   // checked baseline: needs the host-known count
    {
        StageTimer t(c, GSB_STAGE_SORT, st);
        // 4. stable sort on the superblock id alone (the pairs are already in (depth, index) order)
        if (use_cub) {
            const uint32_t L1 = std::min(v.h_ctl[4], c->capL1);
            int rc = cub_sort32(c, st, v.keys[0], v.keys[1], v.vals[0], v.vals[1], L1, (uint32_t)c->sbBits);
            if (rc != GSB_OK) return rc;
            v.d_result_buf = L1 > 0 ? c->d_one : c->d_zero;
            launches += 1;
        } else {
            GSB_CUDA_CHECK(c, launch_onesweep_sort32(st, c->plan, v.keys[0], v.keys[1], v.vals[0], v.vals[1], 0, &v.d_ctl[4],
                                                     v.sort_ws, &v.d_result_buf, &launches));
        }
    }
    {
        StageTimer t(c, GSB_STAGE_TILE_LISTS, st);
        // 5. superblock slices -> per-tile counts -> CSR ranges + heavy-first order -> the tile lists (tilelists.cu)
        GSB_CUDA_CHECK(c, launch_key_ranges(st, v.keys[0], v.keys[1], v.d_result_buf, &v.d_ctl[4], c->capL1, v.sb_ranges, c->numSB));
        GSB_CUDA_CHECK(c, launch_l2_count(st, c->numSB, c->sbGridW, c->gridW, c->gridH, v.sb_ranges, v.vals[0], v.vals[1], v.d_result_buf,
                                          v.tile_rects, v.slice_counts, v.tile_counts));
        GSB_CUDA_CHECK(c, launch_exclusive_scan(st, c->numTiles, v.tile_counts, nullptr, nullptr, nullptr, v.tile_starts, &v.d_ctl[5],
                                                v.tile_scan_ws));
        GSB_CUDA_CHECK(c, launch_tile_bases(st, c->gridW, c->gridH, c->sbGridW, v.tile_counts, v.tile_starts, v.slice_counts, v.slice_base,
                                            v.tile_ranges, v.order_ws, v.tile_order));
        launches += 2;
        GSB_CUDA_CHECK(c, launch_l2_fill(st, c->numSB, c->sbGridW, v.sb_ranges, v.vals[0], v.vals[1], v.d_result_buf, v.tile_rects,
                                         v.slice_base, v.list, c->capM));
        launches += 4;
    }
    c->stats.kernel_launches += launches;
    return GSB_OK;
}

// Waits for the pair count of the binning enqueued on set v (the event fired right after the scan, long
// before the queue drains).  Returns GSB_OK, or GSB_ERR_CAPACITY when the lists did not fit: the pair
// buffers have then been regrown and the caller must redo projection + binning of that view.
static int finish_binning(Ctx* c, Ctx::ViewBufs& v)
{
    GSB_CUDA_CHECK(c, cudaEventSynchronize(v.ev_ctl));
    const uint32_t M = v.h_ctl[0], L1 = v.h_ctl[4];
    if (M <= c->capM && L1 <= c->capL1) {
        v.last_M = M;
        v.last_L1 = L1;
        return GSB_OK;
    }
    int rc = ensure_pairs(c, M, L1);
    if (rc != GSB_OK) return rc;
    set_error(c, "intersection buffers regrown");
    return GSB_ERR_CAPACITY;
}

// blocking variant on the work stream for the single-view API
static int run_binning(Ctx* c, Ctx::ViewBufs& v, int N, const ViewParams& vp, const float* depth_ptr, int depth_stride,
                       bool keep_unsorted)
{
    for (int attempt = 0;; ++attempt) {
        int rc = enqueue_binning(c, v, c->stream, N, vp, depth_ptr, depth_stride);
        if (rc != GSB_OK) return rc;
        rc = finish_binning(c, v);
        if (rc == GSB_OK) break;
        if (rc != GSB_ERR_CAPACITY || attempt == 3) {
            if (rc == GSB_ERR_CAPACITY) set_error(c, "intersection buffers kept overflowing");
            return rc;
        }
    }
    if (keep_unsorted) {
        // the reference's UNSORTED lists (emission order = Gaussian index order, 64-bit keys): parity API only
        if (c->dbg_cap < c->capM) {
            sync_all_streams(c);
            dev_free(c->dbg_keys); dev_free(c->dbg_vals);
            GSB_CUDA_CHECK(c, dev_alloc(&c->dbg_keys, (size_t)c->capM));
            GSB_CUDA_CHECK(c, dev_alloc(&c->dbg_vals, (size_t)c->capM));
            c->dbg_cap = c->capM;
        }
        GSB_CUDA_CHECK(c, launch_exclusive_scan(c->stream, N, v.touched, nullptr, nullptr, nullptr, c->offsets_ref, &v.d_ctl[2],
                                                v.scan_ws));
        GSB_CUDA_CHECK(c, launch_generate_keys_ref(c->stream, N, vp, v.tile_rects, c->offsets_ref, depth_ptr, depth_stride,
                                                   c->dbg_keys, c->dbg_vals, c->capM));
        c->stats.kernel_launches += 2;
    }
    return GSB_OK;
}

static int check_ctx(Ctx* c)
{
    if (!c) return GSB_ERR_INVALID;
    cudaError_t e = cudaSetDevice(c->cfg.device);
    if (e != cudaSuccess) {
        set_error(c, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        return GSB_ERR_CUDA;
    }
    return GSB_OK;
}

// unmap the other replicas' trainer slabs (peer-memory data parallelism, gsb_trainer_peers_import)
static void peers_close(Ctx* c)
{
    for (int r = 0; r < gsb::GSB_MAX_PEERS; ++r) {
        if (c->peer_opened[r]) {
            if (c->peer_block[r]) cudaIpcCloseMemHandle(c->peer_block[r]);
            if (c->peer_accum[r]) cudaIpcCloseMemHandle(c->peer_accum[r]);
            if (c->peer_sync[r]) cudaIpcCloseMemHandle(c->peer_sync[r]);
        }
        c->peer_block[r] = nullptr; c->peer_accum[r] = nullptr; c->peer_sync[r] = nullptr; c->peer_opened[r] = false;
    }
    c->peer_world = 0;
}


// After a gsb_trainer_step_peers the other replicas may still be storing parameters into this replica (and reading its
// gradients): anything that reads the trainer state from the stream, or overwrites the gradients, first waits for their
// "parameters written" flags of that step.
static cudaError_t drain_peer_step(Ctx* c)
{
    if (!c->peer_wait_step || !c->t_sync || c->peer_world < 1) return cudaSuccess;
    cudaError_t e = launch_peer_wait(c->stream, &c->t_sync->params_ready[0][0], c->peer_wait_chunks, c->peer_world, GSB_MAX_PEERS,
                                     c->peer_wait_step, &c->t_sync->error);
    c->stats.kernel_launches += 1;
    c->peer_wait_step = 0;
    return e;
}

static void destroy_ctx(Ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (cudaStream_t fs : c->front_streams) if (fs) cudaStreamSynchronize(fs);
    if (c->tail_stream) cudaStreamSynchronize(c->tail_stream);
    if (c->xchg_stream) cudaStreamSynchronize(c->xchg_stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (Ctx::ViewBufs& v : c->vb) {
        dev_free(v.rec); dev_free(v.tile_rects); dev_free(v.touched); dev_free(v.offsets); dev_free(v.d_nvalue);
        for (int i = 0; i < 2; ++i) { dev_free(v.dkeys[i]); dev_free(v.dvals[i]); dev_free(v.keys[i]); dev_free(v.vals[i]); }
        if (v.dsort_ws) cudaFree(v.dsort_ws);
        if (v.scan_ws) cudaFree(v.scan_ws);
        if (v.sort_ws) cudaFree(v.sort_ws);
        dev_free(v.tile_ranges); dev_free(v.tile_order); dev_free(v.order_ws); dev_free(v.d_ctl); dev_free(v.list);
        dev_free(v.sb_ranges); dev_free(v.slice_counts); dev_free(v.slice_base); dev_free(v.tile_starts); dev_free(v.tile_counts);
        if (v.tile_scan_ws) cudaFree(v.tile_scan_ws);
        if (v.h_ctl) cudaFreeHost(v.h_ctl);
        cudaEvent_t* evs[] = {&v.ev_ctl, &v.ev_front, &v.ev_back};
        for (cudaEvent_t* e : evs) if (*e) cudaEventDestroy(*e);
    }
    dev_free(c->grad_rec); dev_free(c->grad_rec2); dev_free(c->act_tmp); dev_free(c->offsets_ref);
    for (int i = 0; i < 2; ++i) { dev_free(c->t_target[i]); dev_free(c->t_dtarget[i]); dev_free(c->t_dmask[i]); }
    dev_free(c->cot_depth);
    if (c->cub_tmp) cudaFree(c->cub_tmp);
    dev_free(c->dbg_keys); dev_free(c->dbg_vals);
    dev_free(c->out_color); dev_free(c->out_depth); dev_free(c->out_alpha); dev_free(c->out_last);
    if (c->ck.state) cudaFree(c->ck.state);
    if (c->ck.header) cudaFree(c->ck.header);
    dev_free(c->ck.depth); dev_free(c->ck.count); dev_free(c->ck.written); dev_free(c->ck.table);
    dev_free(c->mapA); dev_free(c->mapB); dev_free(c->mapC); dev_free(c->cot_render); dev_free(c->partial);
    dev_free(c->loss_accum); dev_free(c->d_zero);
    peers_close(c);
    if (c->t_sync) cudaFree(c->t_sync);
    if (c->h_peer_error) cudaFreeHost(c->h_peer_error);
    for (cudaEvent_t& e : c->ev_chunk) if (e) cudaEventDestroy(e);
    if (c->ev_sh) cudaEventDestroy(c->ev_sh);
    for (int i = 0; i < 2; ++i) { dev_free(c->t_slab[i]); dev_free(c->t_accum_slab[i]); }
    if (c->h_loss) cudaFreeHost(c->h_loss);
    for (auto& e : c->ev_pool) { if (e.a) cudaEventDestroy(e.a); if (e.b) cudaEventDestroy(e.b); }
    cudaEvent_t* evs[] = {&c->t_target_ready[0], &c->t_target_ready[1], &c->t_target_free[0], &c->t_target_free[1], &c->ev_fork,
                          &c->ev_rb[0], &c->ev_rb[1], &c->ev_pb[0], &c->ev_pb[1]};
    for (cudaEvent_t* e : evs) if (*e) cudaEventDestroy(*e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (cudaStream_t fs : c->front_streams) if (fs) cudaStreamDestroy(fs);
    if (c->tail_stream) cudaStreamDestroy(c->tail_stream);
    if (c->xchg_stream) cudaStreamDestroy(c->xchg_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

}  // namespace gsb

using gsb::Ctx;
struct gsb_ctx : gsb::Ctx {};
static inline Ctx* C(gsb_ctx* p) { return static_cast<Ctx*>(p); }

#define CTX_PROLOGUE(ctxp)                        \
    Ctx* c = C(ctxp);                             \
    {                                             \
        int _rc = gsb::check_ctx(c);              \
        if (_rc != GSB_OK) return _rc;            \
    }

extern "C" {

int gsb_abi_version(void) { return GSB_ABI_VERSION; }

void gsb_default_config(gsb_config* cfg)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->width = 64; cfg->height = 64;
    cfg->tile_w = 16; cfg->tile_h = 16;
    cfg->sh_degree = 3; cfg->sh_coeffs = 16;
    cfg->white_background = 0;
    cfg->max_gaussians = 0;
    cfg->device = 0;
    cfg->flags = 0;
    cfg->lambda_dssim = 0.2f;
    cfg->adam_beta1 = 0.9f; cfg->adam_beta2 = 0.999f; cfg->adam_eps = 1e-15f;
}

const char* gsb_last_error(const gsb_ctx* ctx)
{
    if (!ctx) return gsb::g_create_error.c_str();
    return static_cast<const Ctx*>(ctx)->err.c_str();
}

int gsb_create(const gsb_config* cfg, gsb_ctx** out)
{
    using namespace gsb;
    if (!cfg || !out) { set_error(nullptr, "gsb_create: null argument"); return GSB_ERR_INVALID; }
    *out = nullptr;
    if (cfg->width < 1 || cfg->height < 1 || cfg->tile_w < 1 || cfg->tile_h < 1 || cfg->sh_degree < 0 || cfg->sh_degree > 4 ||
        cfg->sh_coeffs < (cfg->sh_degree + 1) * (cfg->sh_degree + 1) || cfg->sh_coeffs > 25 || cfg->sh_coeffs < 1) {
        set_error(nullptr, "gsb_create: invalid configuration (size, tile, sh_degree 0..4, (deg+1)^2 <= sh_coeffs <= 25)");
        return GSB_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(nullptr, std::string("gsb_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
        return GSB_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error(nullptr, "gsb_create: bad device ordinal"); return GSB_ERR_INVALID; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { set_error(nullptr, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); return GSB_ERR_CUDA; }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10) {
        set_error(nullptr, "gsb_create: this library contains sm_100a code only (Blackwell B200 required)");
        return GSB_ERR_UNSUPPORTED;
    }
    gsb_ctx* h = new (std::nothrow) gsb_ctx();
    if (!h) { set_error(nullptr, "out of host memory"); return GSB_ERR_INVALID; }
    Ctx* c = h;
    c->cfg = *cfg;
    c->gridW = (cfg->width + cfg->tile_w - 1) / cfg->tile_w;
    c->gridH = (cfg->height + cfg->tile_h - 1) / cfg->tile_h;
    if (c->gridW > 65535 || c->gridH > 65535) { delete h; set_error(nullptr, "tile grid too large"); return GSB_ERR_UNSUPPORTED; }
    c->numTiles = c->gridW * c->gridH;
    c->tileBits = tile_bit_count(c->numTiles);
    c->sbGridW = (c->gridW + gsb::SBW - 1) / gsb::SBW;
    c->sbGridH = (c->gridH + gsb::SBH - 1) / gsb::SBH;
    c->numSB = c->sbGridW * c->sbGridH;
    c->sbBits = tile_bit_count(c->numSB);
    c->P = cfg->width * cfg->height;
#define CREATE_CHECK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            set_error(nullptr, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
            destroy_ctx(c);                                                                 \
            return GSB_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)
    int prio_lo = 0, prio_hi = 0;
    CREATE_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // hi = numerically lowest = most urgent
    CREATE_CHECK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    // The tail stream (projection backward, which the next Adam waits for) and the exchange stream are urgent; the front
    // streams run at the work stream's own priority: urgent fronts take the SMs from the rasteriser that is running to finish
    // lists that are not needed yet (measured on B200, profiles/r2/r2zj_*: 13.59 -> 13.44 ms per step, end to end 14.11 ->
    // 13.70).  Tuning: GSB_FRONT_PRIO / GSB_TAIL_PRIO = levels above the work stream's priority (0 ... 5).
    auto level = [&](const char* name, int dflt) {
        const char* e = getenv(name);
        const int v = (e && *e) ? atoi(e) : dflt;
        return std::max(prio_hi, prio_lo - std::max(v, 0));
    };
    const int prio_front = level("GSB_FRONT_PRIO", 0);
    const int prio_tail = level("GSB_TAIL_PRIO", 5);
    for (cudaStream_t& fs : c->front_streams) CREATE_CHECK(cudaStreamCreateWithPriority(&fs, cudaStreamNonBlocking, prio_front));
    c->front_stream = c->front_streams[0];
    CREATE_CHECK(cudaStreamCreateWithPriority(&c->tail_stream, cudaStreamNonBlocking, prio_tail));
    CREATE_CHECK(cudaStreamCreateWithPriority(&c->xchg_stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 2; ++i) {
        CREATE_CHECK(cudaEventCreateWithFlags(&c->ev_rb[i], cudaEventDisableTiming));
        CREATE_CHECK(cudaEventCreateWithFlags(&c->ev_pb[i], cudaEventDisableTiming));
    }
    CREATE_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CREATE_CHECK(cudaEventCreateWithFlags(&c->t_target_ready[i], cudaEventDisableTiming));
        CREATE_CHECK(cudaEventCreateWithFlags(&c->t_target_free[i], cudaEventDisableTiming));
    }
    const size_t P = (size_t)c->P;
    CREATE_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (Ctx::ViewBufs& v : c->vb) {
        CREATE_CHECK(cudaEventCreateWithFlags(&v.ev_ctl, cudaEventDisableTiming));
        CREATE_CHECK(cudaEventCreateWithFlags(&v.ev_front, cudaEventDisableTiming));
        CREATE_CHECK(cudaEventCreateWithFlags(&v.ev_back, cudaEventDisableTiming));
        CREATE_CHECK(dev_alloc(&v.tile_ranges, (size_t)c->numTiles * 2));
        CREATE_CHECK(dev_alloc(&v.tile_order, (size_t)c->numTiles));
        CREATE_CHECK(dev_alloc(&v.order_ws, (size_t)2 * gsb::TO_BUCKETS));
        CREATE_CHECK(dev_alloc(&v.tile_starts, (size_t)c->numTiles + 1));
        CREATE_CHECK(dev_alloc(&v.tile_counts, (size_t)c->numTiles));
        CREATE_CHECK(cudaMalloc(&v.tile_scan_ws, scan_ws_bytes(c->numTiles)));
        CREATE_CHECK(dev_alloc(&v.sb_ranges, (size_t)c->numSB * 2));
        CREATE_CHECK(dev_alloc(&v.slice_counts, (size_t)c->numSB * gsb::L2_SLICES * gsb::SB_TILES));
        CREATE_CHECK(dev_alloc(&v.slice_base, (size_t)c->numSB * gsb::L2_SLICES * gsb::SB_TILES));
        CREATE_CHECK(dev_alloc(&v.d_ctl, 8));
        CREATE_CHECK(dev_alloc(&v.d_nvalue, 4));
        CREATE_CHECK(cudaMallocHost(reinterpret_cast<void**>(&v.h_ctl), 8 * sizeof(uint32_t)));
    }
    CREATE_CHECK(dev_alloc(&c->out_color, P * 3));
    CREATE_CHECK(dev_alloc(&c->out_depth, P));
    CREATE_CHECK(dev_alloc(&c->out_alpha, P));
    CREATE_CHECK(dev_alloc(&c->out_last, P));
    {
        // checkpoint pool: ~1.1 slots per block at C3 (lists are cut short by early termination); 4 per block leaves
        // room for translucent scenes, and an exhausted pool only costs load balance (5 KB per slot)
        const int blocks = gsb::raster_block_count(make_view_nocam(c));
        c->ck.capacity = (uint32_t)std::max(4096, 4 * blocks);
        if (const char* e = getenv("GSB_CKPT_SLOTS")) c->ck.capacity = (uint32_t)std::max(1, atoi(e));   // tests: force exhaustion
        CREATE_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->ck.state), (size_t)c->ck.capacity * 256 * sizeof(float4)));
        CREATE_CHECK(dev_alloc(&c->ck.depth, (size_t)c->ck.capacity * 256));
        CREATE_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->ck.header), (size_t)c->ck.capacity * sizeof(uint2)));
        CREATE_CHECK(dev_alloc(&c->ck.count, 4));
        CREATE_CHECK(dev_alloc(&c->ck.written, (size_t)std::max(blocks, 1)));
        CREATE_CHECK(dev_alloc(&c->ck.table, (size_t)std::max(blocks, 1) * 17));
        CREATE_CHECK(cudaMemset(c->ck.count, 0, 4 * sizeof(uint32_t)));
        CREATE_CHECK(cudaMemset(c->ck.written, 0, (size_t)std::max(blocks, 1) * sizeof(uint32_t)));
    }
    CREATE_CHECK(dev_alloc(&c->mapA, P * 3));
    CREATE_CHECK(dev_alloc(&c->mapB, P * 3));
    CREATE_CHECK(dev_alloc(&c->mapC, P * 3));
    CREATE_CHECK(dev_alloc(&c->cot_render, P * 3));
    CREATE_CHECK(dev_alloc(&c->partial, 4));
    CREATE_CHECK(dev_alloc(&c->loss_accum, 4));
    CREATE_CHECK(dev_alloc(&c->d_zero, 4));
    c->d_one = c->d_zero + 1;
    const uint32_t zo[4] = {0u, 1u, 0u, 0u};
    CREATE_CHECK(cudaMemcpy(c->d_zero, zo, sizeof(zo), cudaMemcpyHostToDevice));
    CREATE_CHECK(cudaMemset(c->loss_accum, 0, 4 * sizeof(float)));
    CREATE_CHECK(cudaMallocHost(reinterpret_cast<void**>(&c->h_loss), 4 * sizeof(float)));
#undef CREATE_CHECK
    if (cfg->max_gaussians > 0) {
        int rc = ensure_gaussians(c, cfg->max_gaussians);
        if (rc != GSB_OK) { g_create_error = c->err; destroy_ctx(c); return rc; }
    }
    *out = h;
    return GSB_OK;
}

void gsb_destroy(gsb_ctx* ctx) { gsb::destroy_ctx(C(ctx)); }

int gsb_set_stream(gsb_ctx* ctx, void* cuda_stream)
{
    CTX_PROLOGUE(ctx);
    gsb::sync_all_streams(c);
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : cudaStreamLegacy;
    return GSB_OK;
}

int gsb_set_flags(gsb_ctx* ctx, int32_t flags)
{
    CTX_PROLOGUE(ctx);
    gsb::sync_all_streams(c);
    c->cfg.flags = flags;
    return GSB_OK;
}

int gsb_synchronize(gsb_ctx* ctx)
{
    CTX_PROLOGUE(ctx);
    GSB_CUDA_CHECK(c, gsb::drain_peer_step(c));
    gsb::sync_all_streams(c);
    GSB_CUDA_CHECK(c, cudaGetLastError());
    return GSB_OK;
}

// ---- activations ---------------------------------------------------------------------------------
int gsb_activate_fwd(gsb_ctx* ctx, int32_t N, const float* f_dc, const float* f_rest, const float* scales_log,
                     const float* rot_raw, const float* opacity_logit, float* shs, float* scales, float* rotations,
                     float* opacity)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && f_dc && scales_log && rot_raw && opacity_logit && shs && scales && rotations && opacity &&
                       (f_rest || c->cfg.sh_coeffs == 1),
                "gsb_activate_fwd: null argument");
    GSB_CUDA_CHECK(c, gsb::launch_activate_fwd(c->stream, N, c->cfg.sh_coeffs, f_dc, f_rest, scales_log, rot_raw, opacity_logit,
                                               shs, scales, rotations, opacity));
    c->stats.kernel_launches += N > 0;
    return GSB_OK;
}

int gsb_activate_bwd(gsb_ctx* ctx, int32_t N, const float* scales_log, const float* rot_raw, const float* opacity_logit,
                     const float* g_shs, const float* g_scales, const float* g_rotations, const float* g_opacity,
                     float* g_f_dc, float* g_f_rest, float* g_scales_log, float* g_rot_raw, float* g_opacity_logit)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && scales_log && rot_raw && opacity_logit && g_shs && g_scales && g_rotations && g_opacity &&
                       g_f_dc && (g_f_rest || c->cfg.sh_coeffs == 1) && g_scales_log && g_rot_raw && g_opacity_logit,
                "gsb_activate_bwd: null argument");
    GSB_CUDA_CHECK(c, gsb::launch_activate_bwd(c->stream, N, c->cfg.sh_coeffs, scales_log, rot_raw, opacity_logit, g_shs,
                                               g_scales, g_rotations, g_opacity, g_f_dc, g_f_rest, g_scales_log, g_rot_raw,
                                               g_opacity_logit));
    c->stats.kernel_launches += N > 0;
    return GSB_OK;
}

// ---- K1 / K2 -------------------------------------------------------------------------------------
int gsb_project_fwd(gsb_ctx* ctx, int32_t N, const float* scales, const float* rotations, const float* means3d,
                    const float* shs, const gsb_camera* host_cam, float* means2d, float* depths, float* color, float* cov2d,
                    float* conic, float* radii, float* rect_min, float* rect_max)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && scales && rotations && means3d && shs && host_cam && means2d && depths && color && cov2d &&
                       conic && radii && rect_min && rect_max,
                "gsb_project_fwd: null argument");
    const gsb::ViewParams vp = gsb::make_view(c, host_cam);
    gsb::StageTimer t(c, GSB_STAGE_PROJECT_FWD);
    GSB_CUDA_CHECK(c, gsb::launch_project_fwd_api(c->stream, N, vp, scales, rotations, means3d, shs, means2d, depths, color,
                                                  cov2d, conic, radii, rect_min, rect_max));
    c->stats.kernel_launches += N > 0;
    return GSB_OK;
}

int gsb_project_bwd(gsb_ctx* ctx, int32_t N, const float* scales, const float* rotations, const float* means3d,
                    const float* shs, const gsb_camera* host_cam, const float* cot_depths, const float* cot_means2d,
                    const float* cot_cov2d, const float* cot_color, const float* cot_conic, float* g_scales,
                    float* g_rotations, float* g_means3d, float* g_shs, float* g_cam_center_point)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && scales && rotations && means3d && shs && host_cam && cot_depths && cot_means2d && cot_cov2d &&
                       cot_color && cot_conic && g_scales && g_rotations && g_means3d && g_shs && g_cam_center_point,
                "gsb_project_bwd: null argument");
    const gsb::ViewParams vp = gsb::make_view(c, host_cam);
    gsb::StageTimer t(c, GSB_STAGE_PROJECT_BWD);
    GSB_CUDA_CHECK(c, gsb::launch_project_bwd_api(c->stream, N, vp, scales, rotations, means3d, shs, cot_depths, cot_means2d,
                                                  cot_cov2d, cot_color, cot_conic, g_scales, g_rotations, g_means3d, g_shs,
                                                  g_cam_center_point));
    c->stats.kernel_launches += N > 0;
    return GSB_OK;
}

// ---- K3..K8 --------------------------------------------------------------------------------------
int gsb_bin(gsb_ctx* ctx, int32_t N, const float* rect_min, const float* rect_max, const float* radii, const float* depths,
            uint32_t* tiles_touched, uint32_t* tile_ranges, uint32_t* tile_counts, uint32_t* host_M)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && rect_min && rect_max && radii && depths, "gsb_bin: null argument");
    int rc = gsb::ensure_gaussians(c, N);
    if (rc != GSB_OK) return rc;
    const gsb::ViewParams vp = gsb::make_view_nocam(c);
    Ctx::ViewBufs& v = c->vb[c->cur];
    c->saved.valid = false;
    c->bin_gen += 1;
    // keep a private copy of the depths: gsb_bin_read rebuilds the reference's sortedKeysLow from them
    if (N > 0) GSB_CUDA_CHECK(c, cudaMemcpyAsync(c->act_tmp, depths, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    GSB_CUDA_CHECK(c, gsb::launch_count_tiles(c->stream, N, vp, rect_min, rect_max, radii, c->act_tmp, v.tile_rects, v.touched,
                                              v.dkeys[0]));
    c->stats.kernel_launches += N > 0;
    rc = gsb::run_binning(c, v, N, vp, c->act_tmp, 1, true);
    if (rc != GSB_OK) return rc;
    c->bin_valid = true;
    c->bin_vp = vp;
    if (tiles_touched && N > 0)
        GSB_CUDA_CHECK(c, cudaMemcpyAsync(tiles_touched, v.touched, (size_t)N * 4, cudaMemcpyDeviceToDevice, c->stream));
    if (tile_ranges)
        GSB_CUDA_CHECK(c, cudaMemcpyAsync(tile_ranges, v.tile_ranges, (size_t)c->numTiles * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (tile_counts) {
        GSB_CUDA_CHECK(c, gsb::launch_tile_counts(c->stream, c->numTiles, v.tile_ranges, tile_counts));
        c->stats.kernel_launches += 1;
    }
    if (host_M) *host_M = v.last_M;
    c->stats.pairs_last_view = v.last_M;
    return GSB_OK;
}

int gsb_bin_read(gsb_ctx* ctx, uint32_t* keys_high, uint32_t* keys_low, uint32_t* gauss_idx, uint32_t* sorted_keys_high,
                 uint32_t* sorted_keys_low, uint32_t* sorted_gauss_idx)
{
    CTX_PROLOGUE(ctx);
    if (!c->bin_valid && !c->saved.valid) { gsb::set_error(c, "gsb_bin_read: no binning result on this context"); return GSB_ERR_STATE; }
    Ctx::ViewBufs& v = c->vb[c->cur];
    const uint32_t M = v.last_M;
    if (M == 0) return GSB_OK;
    if (keys_high || keys_low || gauss_idx) {
        if (!c->dbg_keys) { gsb::set_error(c, "gsb_bin_read: unsorted keys are only kept by gsb_bin"); return GSB_ERR_STATE; }
        GSB_CUDA_CHECK(c, gsb::launch_split_keys(c->stream, M, c->dbg_keys, keys_high, keys_low));
        if (gauss_idx) GSB_CUDA_CHECK(c, cudaMemcpyAsync(gauss_idx, c->dbg_vals, (size_t)M * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (sorted_keys_high || sorted_keys_low || sorted_gauss_idx) {
        // the lists are stored without keys: tile id = position in the CSR offsets, depth bits = the listed Gaussian's depth
        GSB_CUDA_CHECK(c, gsb::launch_expand_sorted_keys(c->stream, M, c->numTiles, v.tile_starts, v.list, v.depth_src, v.depth_src_stride,
                                                         sorted_keys_high, sorted_keys_low));
        if (sorted_gauss_idx)
            GSB_CUDA_CHECK(c, cudaMemcpyAsync(sorted_gauss_idx, v.list, (size_t)M * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    return GSB_OK;
}

int gsb_sort_tile_keys(gsb_ctx* ctx, uint32_t M, uint32_t tile_bits, const uint32_t* keys_high, const uint32_t* keys_low,
                       const uint32_t* values, uint32_t* sorted_high, uint32_t* sorted_low, uint32_t* sorted_values,
                       int32_t use_cub)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, tile_bits >= 1 && tile_bits <= 32, "gsb_sort_tile_keys: tile_bits must be 1..32");
    if (M == 0) return GSB_OK;
    GSB_REQUIRE(c, keys_high && keys_low && values && sorted_high && sorted_low && sorted_values, "gsb_sort_tile_keys: null argument");
    // stand-alone buffers (does not disturb the context's tile lists)
    uint64_t* k[2] = {nullptr, nullptr};
    uint32_t* v[2] = {nullptr, nullptr};
    void* ws = nullptr;
    void* tmp = nullptr;
    uint32_t* d_count = nullptr;
    int rc = GSB_OK;
    cudaError_t e = cudaSuccess;
    const uint32_t end_bit = 32u + tile_bits;
    const uint32_t hi_mask = tile_bits >= 32 ? 0xffffffffu : ((1u << tile_bits) - 1u);
    auto fail = [&](cudaError_t err, const char* what) {
        gsb::set_error(c, std::string(what) + ": " + cudaGetErrorString(err));
        rc = GSB_ERR_CUDA;
    };
    do {
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = gsb::dev_alloc(&k[i], (size_t)M);
            if (e == cudaSuccess) e = gsb::dev_alloc(&v[i], (size_t)M);
        }
        if (e == cudaSuccess) e = gsb::dev_alloc(&d_count, 1);
        if (e != cudaSuccess) { fail(e, "cudaMalloc"); break; }
        e = cudaMemcpyAsync(d_count, &M, 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = gsb::launch_merge_keys(c->stream, M, keys_high, keys_low, hi_mask, k[0]);
        if (e == cudaSuccess) e = cudaMemcpyAsync(v[0], values, (size_t)M * 4, cudaMemcpyDeviceToDevice, c->stream);
        if (e != cudaSuccess) { fail(e, "merge keys"); break; }
        uint32_t buf = 1;
        if (use_cub) {
            size_t need = 0;
            e = gsb::cub_sort_pairs(c->stream, k[0], k[1], v[0], v[1], M, end_bit, nullptr, 0, &need);
            if (e == cudaSuccess) e = cudaMalloc(&tmp, need ? need : 1);
            if (e == cudaSuccess) e = gsb::cub_sort_pairs(c->stream, k[0], k[1], v[0], v[1], M, end_bit, tmp, need, nullptr);
            if (e != cudaSuccess) { fail(e, "cub sort"); break; }
        } else {
            gsb::SortPlan plan = gsb::sort_plan(M, end_bit);
            e = cudaMalloc(&ws, plan.ws_bytes);
            const uint32_t* d_res = nullptr;
            int launches = 0;
            if (e == cudaSuccess) e = gsb::launch_onesweep_sort(c->stream, plan, k[0], k[1], v[0], v[1], d_count, ws, &d_res, &launches);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&c->vb[0].h_ctl[2], d_res, 4, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) { fail(e, "onesweep sort"); break; }
            buf = c->vb[0].h_ctl[2];
            c->stats.kernel_launches += launches;
        }
        e = gsb::launch_split_keys(c->stream, M, k[buf], sorted_high, sorted_low);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sorted_values, v[buf], (size_t)M * 4, cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { fail(e, "split keys"); break; }
    } while (0);
    if (rc != GSB_OK) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 2; ++i) { if (k[i]) cudaFree(k[i]); if (v[i]) cudaFree(v[i]); }
    if (ws) cudaFree(ws);
    if (tmp) cudaFree(tmp);
    if (d_count) cudaFree(d_count);
    return rc;
}

// ---- K9 / K10 ------------------------------------------------------------------------------------
static int restage_packed(Ctx* c, int32_t N, const float* packed)
{
    if (!c->bin_valid && !c->saved.valid) { gsb::set_error(c, "raster: call gsb_bin (or gsb_render_forward) first"); return GSB_ERR_STATE; }
    if (N > c->capN) { gsb::set_error(c, "raster: N larger than the binned scene"); return GSB_ERR_INVALID; }
    GSB_CUDA_CHECK(c, gsb::launch_packed_to_rec(c->stream, N, packed, c->vb[c->cur].rec));
    c->stats.kernel_launches += 1;
    return GSB_OK;
}

int gsb_raster_fwd(gsb_ctx* ctx, int32_t N, const float* packed, float* out_color, float* out_depth, float* out_alpha,
                   uint32_t* out_last_contrib)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && packed && out_color && out_depth && out_alpha && out_last_contrib, "gsb_raster_fwd: null argument");
    if (c->saved.valid) c->bin_vp = c->saved.vp;
    int rc = restage_packed(c, N, packed);
    if (rc != GSB_OK) return rc;
    Ctx::ViewBufs& v = c->vb[c->cur];
    gsb::StageTimer t(c, GSB_STAGE_RASTER_FWD);
    // kernel-level mirror of K9/K10: the caller owns the saved outputs, so no checkpoints are kept for them
    GSB_CUDA_CHECK(c, gsb::launch_raster_fwd(c->stream, c->bin_vp, v.tile_ranges, v.tile_order, v.rec, v.list, v.list, c->d_zero, out_color, out_depth, out_alpha,
                                             out_last_contrib, &v.d_ctl[3], gsb::RasterCkpt{}));
    c->stats.kernel_launches += 1;
    return GSB_OK;
}

int gsb_raster_bwd(gsb_ctx* ctx, int32_t N, const float* packed, const float* cot_color, const float* cot_depth,
                   const float* cot_alpha, const float* out_color, const float* out_depth, const float* out_alpha,
                   const uint32_t* last_contrib, float* grad_packed)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && packed && cot_color && out_color && out_depth && out_alpha && last_contrib && grad_packed,
                "gsb_raster_bwd: null argument");
    if (c->saved.valid) c->bin_vp = c->saved.vp;
    int rc = restage_packed(c, N, packed);
    if (rc != GSB_OK) return rc;
    Ctx::ViewBufs& v = c->vb[c->cur];
    GSB_CUDA_CHECK(c, cudaMemsetAsync(c->grad_rec, 0, (size_t)N * gsb::REC_FLOATS * 4, c->stream));
    {
        gsb::StageTimer t(c, GSB_STAGE_RASTER_BWD);
        GSB_CUDA_CHECK(c, gsb::launch_raster_bwd(c->stream, c->bin_vp, v.tile_ranges, v.tile_order, v.rec, v.list, v.list, c->d_zero, cot_color, cot_depth, cot_alpha,
                                                 out_color, out_depth, out_alpha, last_contrib, c->grad_rec, &v.d_ctl[3], gsb::RasterCkpt{}));
    }
    GSB_CUDA_CHECK(c, gsb::launch_rec_to_packed(c->stream, N, c->grad_rec, grad_packed));
    c->stats.kernel_launches += 2;
    return GSB_OK;
}

// ---- K11 / K12 -----------------------------------------------------------------------------------
int gsb_ssim_fwd(gsb_ctx* ctx, int32_t H, int32_t W, int32_t Cn, const float* img1, const float* img2, float* ssim_map,
                 float* mu1, float* mu2, float* sigma1_sq, float* sigma2_sq, float* sigma12)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, H >= 0 && W >= 0 && Cn >= 0 && img1 && img2 && ssim_map, "gsb_ssim_fwd: null argument");
    GSB_CUDA_CHECK(c, gsb::launch_ssim_fwd(c->stream, H, W, Cn, img1, img2, ssim_map, mu1, mu2, sigma1_sq, sigma2_sq, sigma12));
    c->stats.kernel_launches += 1;
    return GSB_OK;
}

int gsb_ssim_bwd(gsb_ctx* ctx, int32_t H, int32_t W, int32_t Cn, const float* grad_out, const float* img1, const float* img2,
                 float* grad_img1)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, H >= 0 && W >= 0 && Cn >= 0 && grad_out && img1 && img2 && grad_img1, "gsb_ssim_bwd: null argument");
    const size_t n = (size_t)H * W * Cn;
    float *a = c->mapA, *b = c->mapB, *m = c->mapC;
    float* tmp = nullptr;
    if (n > (size_t)c->P * 3) {  // image larger than the context's: temporary maps
        GSB_CUDA_CHECK(c, gsb::dev_alloc(&tmp, n * 3));
        a = tmp; b = tmp + n; m = tmp + 2 * n;
    }
    cudaError_t e = gsb::launch_ssim_bwd_api(c->stream, H, W, Cn, grad_out, img1, img2, a, b, m, grad_img1);
    c->stats.kernel_launches += 3;
    if (tmp) { cudaStreamSynchronize(c->stream); cudaFree(tmp); }
    GSB_CUDA_CHECK(c, e);
    return GSB_OK;
}

// ---- fused renderer ------------------------------------------------------------------------------
struct RawParams {
    const float *xyz, *f_dc, *f_rest, *scales_log, *rot_raw, *op_logit;
};

// projection (+ fused activations) and binning of one view into set v, enqueued on st
// color_after != nullptr: the projection runs without the SH colour (geometry only), and the colour kernel follows the
// binning once that event has fired (data-parallel step: the SH parameters of the previous exchange arrive last)
static int enqueue_front(Ctx* c, Ctx::ViewBufs& v, cudaStream_t st, int32_t N, const RawParams& p, const gsb::ViewParams& vp,
                         float* radii, uint8_t* visibility, cudaEvent_t color_after = nullptr)
{
    {
        gsb::StageTimer t(c, GSB_STAGE_PROJECT_FWD, st);
        GSB_CUDA_CHECK(c, gsb::launch_project_fused_fwd(st, N, vp, p.xyz, p.f_dc, p.f_rest, p.scales_log, p.rot_raw, p.op_logit, v.rec,
                                                        v.tile_rects, v.touched, v.dkeys[0], radii, visibility, color_after ? 1 : 3));
        c->stats.kernel_launches += N > 0;
    }
    int rc = gsb::enqueue_binning(c, v, st, N, vp, v.rec + 10, gsb::REC_FLOATS);
    if (rc != GSB_OK || !color_after) return rc;
    GSB_CUDA_CHECK(c, cudaStreamWaitEvent(st, color_after, 0));
    {
        gsb::StageTimer t(c, GSB_STAGE_PROJECT_FWD, st);
        GSB_CUDA_CHECK(c, gsb::launch_project_fused_fwd(st, N, vp, p.xyz, p.f_dc, p.f_rest, p.scales_log, p.rot_raw, p.op_logit, v.rec,
                                                        v.tile_rects, v.touched, v.dkeys[0], nullptr, nullptr, 2));
        c->stats.kernel_launches += N > 0;
    }
    return GSB_OK;
}

// K9 on the work stream from set v; marks the forward of this view as the saved one
static int enqueue_raster_fwd(Ctx* c, Ctx::ViewBufs& v, int32_t N, const RawParams& p, const gsb::ViewParams& vp, bool want_depth)
{
    {
        gsb::StageTimer t(c, GSB_STAGE_RASTER_FWD);
        GSB_CUDA_CHECK(c, gsb::launch_raster_fwd(c->stream, vp, v.tile_ranges, v.tile_order, v.rec, v.list, v.list, c->d_zero,
                                                 c->out_color, want_depth ? c->out_depth : nullptr, c->out_alpha, c->out_last, &v.d_ctl[3],
                                                 (c->cfg.flags & GSB_FLAG_NO_SEGMENTS) ? gsb::RasterCkpt{} : c->ck));
        c->ck_has_depth = want_depth;
        c->ck_valid = !(c->cfg.flags & GSB_FLAG_NO_SEGMENTS);
        c->stats.kernel_launches += 1;
    }
    c->cur = (int)(&v - &c->vb[0]);
    c->saved.valid = true;
    c->saved.N = N;
    c->saved.xyz = p.xyz; c->saved.f_dc = p.f_dc; c->saved.f_rest = p.f_rest; c->saved.scales_log = p.scales_log;
    c->saved.rot_raw = p.rot_raw; c->saved.op_logit = p.op_logit;
    c->saved.vp = vp;
    c->bin_vp = vp;
    c->stats.pairs_last_view = v.last_M;
    c->stats.sb_pairs_last_view = v.last_L1;
    c->stats.pairs_total += v.last_M;
    c->stats.views += 1;
    return GSB_OK;
}

static int render_forward_impl(Ctx* c, int32_t N, const RawParams& p, const gsb::ViewParams& vp, float* radii, uint8_t* visibility,
                               bool want_depth)
{
    int rc = gsb::ensure_gaussians(c, N);
    if (rc != GSB_OK) return rc;
    c->saved.valid = false;
    c->bin_valid = false;
    c->bin_gen += 1;
    Ctx::ViewBufs& v = c->vb[c->cur];
    for (int attempt = 0;; ++attempt) {
        rc = enqueue_front(c, v, c->stream, N, p, vp, radii, visibility);
        if (rc != GSB_OK) return rc;
        rc = gsb::finish_binning(c, v);
        if (rc == GSB_OK) break;
        if (rc != GSB_ERR_CAPACITY || attempt == 3) return rc;
    }
    return enqueue_raster_fwd(c, v, N, p, vp, want_depth);
}

// K10 on the work stream, K2 (+ activation VJPs) either right behind it (gbuf < 0: single-view API, serial
// trainer) or on the tail stream out of gradient-record buffer gbuf (pipelined trainer).
// Peer step protocol: the projection backward of the step's LAST view runs as one launch per Gaussian chunk; after
// chunk k the replica raises grads_ready[k][rank] in every replica and records ev_chunk[k].
struct ChunkSignal {
    uint32_t step = 0;
    int chunks = 0;
    long long begin[gsb::GSB_MAX_CHUNKS + 1] = {};
};
static int signal_chunk(Ctx* c, cudaStream_t st, const ChunkSignal& sig, int k)
{
    gsb::PeerFlagList fl;
    fl.n = c->peer_world;
    for (int r = 0; r < c->peer_world; ++r) fl.dst[r] = &c->peer_sync[r]->grads_ready[k][c->peer_rank];
    GSB_CUDA_CHECK(c, gsb::launch_peer_signal(st, fl, sig.step));
    GSB_CUDA_CHECK(c, cudaEventRecord(c->ev_chunk[k], st));
    c->stats.kernel_launches += 1;
    return GSB_OK;
}

static int render_backward_impl(Ctx* c, const float* cot_render, const float* cot_depth, const float* cot_alpha, float* g_xyz,
                                float* g_f_dc, float* g_f_rest, float* g_scales_log, float* g_rot_raw, float* g_opacity_logit,
                                int accumulate, int gbuf = -1, const ChunkSignal* sig = nullptr)
{
    if (!c->saved.valid) { gsb::set_error(c, "gsb_render_backward: no forward saved on this context"); return GSB_ERR_STATE; }
    const int N = c->saved.N;
    const gsb::ViewParams& vp = c->saved.vp;
    Ctx::ViewBufs& v = c->vb[c->cur];
    const bool split = gbuf >= 0;
    float* grec = (split && gbuf == 1) ? c->grad_rec2 : c->grad_rec;
    if (split) GSB_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_pb[gbuf], 0));   // previous reader of this buffer
    GSB_CUDA_CHECK(c, cudaMemsetAsync(grec, 0, (size_t)N * gsb::REC_FLOATS * 4, c->stream));
    {
        gsb::StageTimer t(c, GSB_STAGE_RASTER_BWD);
        GSB_CUDA_CHECK(c, gsb::launch_raster_bwd(c->stream, vp, v.tile_ranges, v.tile_order, v.rec, v.list, v.list, c->d_zero, cot_render, cot_depth, cot_alpha,
                                                 c->out_color, c->out_depth, c->out_alpha, c->out_last, grec, &v.d_ctl[3],
                                                 (!c->ck_valid || (cot_depth && !c->ck_has_depth)) ? gsb::RasterCkpt{} : c->ck));
    }
    cudaStream_t pst = c->stream;
    if (split) {
        GSB_CUDA_CHECK(c, cudaEventRecord(c->ev_rb[gbuf], c->stream));
        GSB_CUDA_CHECK(c, cudaStreamWaitEvent(c->tail_stream, c->ev_rb[gbuf], 0));
        pst = c->tail_stream;
    }
    if (!sig) {
        gsb::StageTimer t(c, GSB_STAGE_PROJECT_BWD, pst);
        GSB_CUDA_CHECK(c, gsb::launch_project_fused_bwd(pst, N, vp, c->saved.xyz, c->saved.f_dc, c->saved.f_rest,
                                                        c->saved.scales_log, c->saved.rot_raw, c->saved.op_logit, grec,
                                                        g_xyz, g_f_dc, g_f_rest, g_scales_log, g_rot_raw, g_opacity_logit,
                                                        accumulate));
        c->stats.kernel_launches += N > 0;
    } else {
        const long long rest = (long long)(c->cfg.sh_coeffs - 1) * 3;
        for (int k = 0; k < sig->chunks; ++k) {
            const long long n0 = sig->begin[k], n1 = sig->begin[k + 1];
            if (n1 > n0) {
                gsb::StageTimer t(c, GSB_STAGE_PROJECT_BWD, pst);
                GSB_CUDA_CHECK(c, gsb::launch_project_fused_bwd(
                    pst, (int)(n1 - n0), vp, c->saved.xyz + n0 * 3, c->saved.f_dc + n0 * 3, c->saved.f_rest ? c->saved.f_rest + n0 * rest : nullptr,
                    c->saved.scales_log + n0 * 3, c->saved.rot_raw + n0 * 4, c->saved.op_logit + n0, grec + n0 * gsb::REC_FLOATS,
                    g_xyz + n0 * 3, g_f_dc + n0 * 3, g_f_rest ? g_f_rest + n0 * rest : nullptr, g_scales_log + n0 * 3, g_rot_raw + n0 * 4,
                    g_opacity_logit + n0, accumulate));
                c->stats.kernel_launches += 1;
            }
            int rc = signal_chunk(c, pst, *sig, k);
            if (rc != GSB_OK) return rc;
        }
    }
    if (split) GSB_CUDA_CHECK(c, cudaEventRecord(c->ev_pb[gbuf], c->tail_stream));
    c->stats.kernel_launches += 1;
    return GSB_OK;
}

int gsb_render_forward(gsb_ctx* ctx, int32_t N, const float* xyz, const float* f_dc, const float* f_rest,
                       const float* scales_log, const float* rot_raw, const float* opacity_logit, const gsb_camera* host_cam,
                       float* render, float* depth, float* alpha, uint8_t* visibility, float* radii)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && xyz && f_dc && (f_rest || c->cfg.sh_coeffs == 1) && scales_log && rot_raw && opacity_logit && host_cam,
                "gsb_render_forward: null argument");
    const gsb::ViewParams vp = gsb::make_view(c, host_cam);
    const RawParams rp{xyz, f_dc, f_rest, scales_log, rot_raw, opacity_logit};
    int rc = render_forward_impl(c, N, rp, vp, radii, visibility, depth != nullptr);
    if (rc != GSB_OK) return rc;
    const size_t P = (size_t)c->P;
    if (render) GSB_CUDA_CHECK(c, cudaMemcpyAsync(render, c->out_color, P * 12, cudaMemcpyDeviceToDevice, c->stream));
    if (depth) GSB_CUDA_CHECK(c, cudaMemcpyAsync(depth, c->out_depth, P * 4, cudaMemcpyDeviceToDevice, c->stream));
    if (alpha) GSB_CUDA_CHECK(c, cudaMemcpyAsync(alpha, c->out_alpha, P * 4, cudaMemcpyDeviceToDevice, c->stream));
    return GSB_OK;
}

int gsb_render_backward(gsb_ctx* ctx, const float* cot_render, const float* cot_depth, const float* cot_alpha, float* g_xyz,
                        float* g_f_dc, float* g_f_rest, float* g_scales_log, float* g_rot_raw, float* g_opacity_logit,
                        int32_t accumulate)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, cot_render && g_xyz && g_f_dc && (g_f_rest || c->cfg.sh_coeffs == 1) && g_scales_log && g_rot_raw && g_opacity_logit,
                "gsb_render_backward: null argument");
    return render_backward_impl(c, cot_render, cot_depth, cot_alpha, g_xyz, g_f_dc, g_f_rest, g_scales_log, g_rot_raw,
                                g_opacity_logit, accumulate);
}

// ---- loss ----------------------------------------------------------------------------------------
// Optional depth supervision of one view (GaussianTrainer.swift:693-699): rendered depth, target, mask, weight.
struct DepthTerm {
    const float* depth = nullptr;
    const float* target = nullptr;
    const uint8_t* mask = nullptr;
    float lambda = 0.0f;
    float* cot = nullptr;
};

static int loss_impl(Ctx* c, const float* render, const float* target, float grad_scale, float* cot_render, float* loss_accum,
                     const DepthTerm* dt = nullptr)
{
    gsb::StageTimer t(c, GSB_STAGE_LOSS);
    const int H = c->cfg.height, W = c->cfg.width;
    const double n = (double)H * W * 3.0;
    const float lambda = c->cfg.lambda_dssim;
    // d total / d ssim_map = -lambda/n ; d total / d render (L1) = (1-lambda)/n * sign
    const float upstream = (float)(-(double)lambda / n) * grad_scale;
    const float l1_scale = (float)((1.0 - (double)lambda) / n) * grad_scale;
    GSB_CUDA_CHECK(c, gsb::launch_loss_fwd(c->stream, H, W, 3, render, target, upstream, c->mapA, c->mapB, c->mapC, c->partial));
    if (loss_accum) GSB_CUDA_CHECK(c, gsb::launch_loss_finalize(c->stream, c->partial, 1.0 / n, lambda, grad_scale, loss_accum));
    GSB_CUDA_CHECK(c, gsb::launch_loss_bwd(c->stream, H, W, 3, render, target, c->mapA, c->mapB, c->mapC, l1_scale, cot_render));
    c->stats.kernel_launches += 2 + (loss_accum != nullptr);
    if (dt) {
        GSB_CUDA_CHECK(c, gsb::launch_depth_loss(c->stream, (size_t)c->P, dt->depth, dt->target, dt->mask, dt->lambda, grad_scale, dt->cot,
                                                 c->partial + 2, loss_accum));
        c->stats.kernel_launches += 2;
    }
    return GSB_OK;
}

int gsb_loss_fwd_bwd(gsb_ctx* ctx, const float* render, const float* target_rgb, float grad_scale, float* cot_render,
                     float* loss_accum)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, render && target_rgb && cot_render, "gsb_loss_fwd_bwd: null argument");
    return loss_impl(c, render, target_rgb, grad_scale, cot_render, loss_accum);
}

int gsb_loss_fwd_bwd_depth(gsb_ctx* ctx, const float* render, const float* depth, const float* target_rgb,
                           const uint8_t* depth_mask, const float* target_depth, float lambda_depth, float grad_scale,
                           float* cot_render, float* cot_depth, float* loss_accum)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, render && depth && target_rgb && depth_mask && target_depth && cot_render && cot_depth,
                "gsb_loss_fwd_bwd_depth: null argument");
    DepthTerm dt;
    dt.depth = depth; dt.target = target_depth; dt.mask = depth_mask; dt.lambda = lambda_depth; dt.cot = cot_depth;
    return loss_impl(c, render, target_rgb, grad_scale, cot_render, loss_accum, &dt);
}

// ---- Adam ----------------------------------------------------------------------------------------
int gsb_adam_step(gsb_ctx* ctx, int32_t N, float* const* host_params, const float* const* host_grads, float* const* host_m,
                  float* const* host_v, const int64_t* host_counts, const float* host_lrs, float* grad_norm_accum)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && host_params && host_grads && host_m && host_v && host_counts && host_lrs, "gsb_adam_step: null argument");
    gsb::AdamTensors t{};
    for (int k = 0; k < 6; ++k) {
        GSB_REQUIRE(c, host_counts[k] >= 0 && (host_counts[k] == 0 || (host_params[k] && host_grads[k] && host_m[k] && host_v[k])),
                    "gsb_adam_step: null tensor");
        t.p[k] = host_params[k]; t.g[k] = host_grads[k]; t.m[k] = host_m[k]; t.v[k] = host_v[k];
        t.count[k] = host_counts[k]; t.lr[k] = host_lrs[k];
    }
    GSB_REQUIRE(c, !grad_norm_accum || host_counts[0] >= (int64_t)N * 3, "gsb_adam_step: xyz tensor smaller than N*3");
    gsb::StageTimer tm(c, GSB_STAGE_ADAM);
    int launches = 0;
    GSB_CUDA_CHECK(c, gsb::launch_adam(c->stream, t, c->cfg.adam_beta1, c->cfg.adam_beta2, c->cfg.adam_eps, 1.0f, N,
                                       grad_norm_accum, nullptr, &launches));
    c->stats.kernel_launches += launches;
    return GSB_OK;
}

// ---- trainer -------------------------------------------------------------------------------------
// Layout of one trainer block: params | grads | m | v, each the six tensors in 128-byte aligned segments.
struct TrainerLayout {
    long long cnt[6];
    size_t off[7];
    size_t floats;
};
static TrainerLayout trainer_layout(int N, int K)
{
    TrainerLayout L;
    const long long cnt[6] = {(long long)N * 3, (long long)N * 3, (long long)N * (K - 1) * 3, (long long)N * 3, (long long)N * 4, (long long)N};
    L.off[0] = 0;
    for (int k = 0; k < 6; ++k) {
        L.cnt[k] = cnt[k];
        L.off[k + 1] = L.off[k] + (((size_t)cnt[k] + 31) & ~(size_t)31);
    }
    L.floats = L.off[6];
    return L;
}
// makes slab `which` (params already in place, grads/m/v/accum zeroed) the trainer state for N Gaussians
static void trainer_adopt(Ctx* c, int N, const TrainerLayout& L, int which)
{
    float* block = c->t_slab[which];
    c->t_cur = which;
    c->t_block = block;
    c->t_accum = c->t_accum_slab[which];
    c->t_floats = L.floats;
    for (int k = 0; k < 6; ++k) {
        c->t_count[k] = L.cnt[k];
        c->t_p[k] = block + L.off[k];
        c->t_g[k] = block + L.floats + L.off[k];
        c->t_m[k] = block + 2 * L.floats + L.off[k];
        c->t_v[k] = block + 3 * L.floats + L.off[k];
    }
    c->tN = N;
    c->t_accum_steps = 0;
    c->sym = false;   // parameters / gradients are back in the slab: a symmetric attachment has to be renewed
}
static cudaError_t trainer_alloc_slab(Ctx* c, int which, int cap)
{
    const TrainerLayout L = trainer_layout(cap, c->cfg.sh_coeffs);
    gsb::dev_free(c->t_slab[which]); gsb::dev_free(c->t_accum_slab[which]);
    cudaError_t e = gsb::dev_alloc(&c->t_slab[which], L.floats * 4);
    if (e == cudaSuccess) e = gsb::dev_alloc(&c->t_accum_slab[which], (size_t)cap);
    return e;
}

int gsb_trainer_init(gsb_ctx* ctx, int32_t N, const float* host_xyz, const float* host_f_dc, const float* host_f_rest,
                     const float* host_scales_log, const float* host_rot_raw, const float* host_opacity_logit)
{
    CTX_PROLOGUE(ctx);
    const int K = c->cfg.sh_coeffs;
    GSB_REQUIRE(c, N > 0 && host_xyz && host_f_dc && (host_f_rest || K == 1) && host_scales_log && host_rot_raw && host_opacity_logit,
                "gsb_trainer_init: null argument");
    int rc = gsb::ensure_gaussians(c, N);
    if (rc != GSB_OK) return rc;
    gsb::sync_all_streams(c);
    if (N != c->tN) gsb::peers_close(c);   // a different layout invalidates what the other replicas mapped
    c->tN = 0;
    c->t_block = nullptr; c->t_accum = nullptr;
    if (!c->t_slab[0] || N > c->t_cap) {   // re-initialisation within the capacity keeps the slabs (and their addresses)
        gsb::peers_close(c);                 // ... and the peers' mappings of them
        for (int i = 0; i < 2; ++i) { gsb::dev_free(c->t_slab[i]); gsb::dev_free(c->t_accum_slab[i]); }
        c->t_cap = N;
        GSB_CUDA_CHECK(c, trainer_alloc_slab(c, 0, c->t_cap));
    }
    const TrainerLayout L = trainer_layout(N, K);
    trainer_adopt(c, N, L, 0);
    GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_block, 0, c->t_floats * 4 * sizeof(float), c->stream));
    GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_accum, 0, (size_t)N * sizeof(float), c->stream));
    const float* src[6] = {host_xyz, host_f_dc, host_f_rest, host_scales_log, host_rot_raw, host_opacity_logit};
    for (int k = 0; k < 6; ++k)
        if (L.cnt[k] > 0)
            GSB_CUDA_CHECK(c, cudaMemcpyAsync(c->t_p[k], src[k], (size_t)L.cnt[k] * 4, cudaMemcpyDefault, c->stream));
    for (int i = 0; i < 2; ++i)
        if (!c->t_target[i]) GSB_CUDA_CHECK(c, gsb::dev_alloc(&c->t_target[i], (size_t)c->P * 3));
    GSB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return GSB_OK;
}

int gsb_trainer_param_ptrs(gsb_ctx* ctx, float** host_params6, float** host_grads6, float** host_m6, float** host_v6,
                           float** grad_norm_accum)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_CUDA_CHECK(c, gsb::drain_peer_step(c));   // reads of the returned buffers queued behind this call see the finished step
    for (int k = 0; k < 6; ++k) {
        if (host_params6) host_params6[k] = c->t_p[k];
        if (host_grads6) host_grads6[k] = c->t_g[k];
        if (host_m6) host_m6[k] = c->t_m[k];
        if (host_v6) host_v6[k] = c->t_v[k];
    }
    if (grad_norm_accum) *grad_norm_accum = c->t_accum;
    return GSB_OK;
}

int gsb_trainer_grad_block(gsb_ctx* ctx, float** grad_block, int64_t* floats)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    if (grad_block) *grad_block = c->t_g[0];   // base of the gradient copy (slab, or the attached symmetric buffer)
    if (floats) *floats = (int64_t)c->t_floats;
    return GSB_OK;
}

// One batch of views into the context's gradient buffers.  host_depths / host_masks (both or neither) switch on the
// depth-supervision term with weight lambda_depth (GaussianTrainer.swift:693-714,949).
static int trainer_accumulate_impl(Ctx* c, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                                   const float* const* host_depths, const uint8_t* const* host_masks, float lambda_depth,
                                   int32_t targets_on_host, int32_t zero_grads, float grad_scale, float* host_loss,
                                   const ChunkSignal* sig = nullptr)
{
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, B >= 0 && (B == 0 || (host_cams && host_targets)), "gsb_trainer_accumulate: null argument");
    GSB_REQUIRE(c, (host_depths == nullptr) == (host_masks == nullptr), "gsb_trainer_accumulate_depth: depth targets and masks go together");
    const bool with_depth = host_depths != nullptr && B > 0;
    const int N = c->tN;
    const size_t img_bytes = (size_t)c->P * 3 * sizeof(float);
    if (with_depth) {
        if (!c->cot_depth) GSB_CUDA_CHECK(c, gsb::dev_alloc(&c->cot_depth, (size_t)c->P));
        for (int i = 0; i < 2 && targets_on_host; ++i) {
            if (!c->t_dtarget[i]) GSB_CUDA_CHECK(c, gsb::dev_alloc(&c->t_dtarget[i], (size_t)c->P));
            if (!c->t_dmask[i]) GSB_CUDA_CHECK(c, gsb::dev_alloc(&c->t_dmask[i], (size_t)c->P));
        }
    }
    // the previous step may have ended without a barrier (gsb_trainer_step_peers): before this replica reads its parameters
    // or overwrites its gradients, every owner must have announced "my stores into your replica are done" - which also
    // means it no longer reads this replica's gradients
    cudaEvent_t sh_event = nullptr;
    if (sig && c->peer_wait_step && c->peer_wait_phased && c->t_sync && B > 0) {
        // Phased exchange of the previous step: its geometry rows gate this batch (work stream, the fronts fork from it),
        // its SH rows only gate the colour kernels (waited for on the tail stream, idle at this point) - the binning of
        // the first view runs while the SH parameters are still crossing the links.  The SH rows also tell that every
        // owner is done with this replica's gradients: the first projection backward that overwrites them comes after a
        // rasteriser, i.e. after a colour kernel.
        const int rows = c->peer_wait_chunks / 2;
        GSB_CUDA_CHECK(c, gsb::launch_peer_wait(c->stream, &c->t_sync->params_ready[0][0], rows, c->peer_world, gsb::GSB_MAX_PEERS,
                                                c->peer_wait_step, &c->t_sync->error));
        GSB_CUDA_CHECK(c, gsb::launch_peer_wait(c->tail_stream, &c->t_sync->params_ready[rows][0], rows, c->peer_world, gsb::GSB_MAX_PEERS,
                                                c->peer_wait_step, &c->t_sync->error));
        GSB_CUDA_CHECK(c, cudaEventRecord(c->ev_sh, c->tail_stream));
        c->stats.kernel_launches += 2;
        c->peer_wait_step = 0;
        sh_event = c->ev_sh;
    }
    GSB_CUDA_CHECK(c, gsb::drain_peer_step(c));
    if (host_loss) GSB_CUDA_CHECK(c, cudaMemsetAsync(c->loss_accum, 0, sizeof(float), c->stream));
    if (zero_grads && B == 0) GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_g[0], 0, c->t_floats * 4, c->stream));
    if (sig && B == 0)   // a replica without views this step: its (zero) gradients are complete right away
        for (int k = 0; k < sig->chunks; ++k) {
            int rc = signal_chunk(c, c->stream, *sig, k);
            if (rc != GSB_OK) return rc;
        }
    // prefetch of view 0's target
    auto prefetch = [&](int b) -> int {
        const int s = b & 1;
        GSB_CUDA_CHECK(c, cudaStreamWaitEvent(c->copy_stream, c->t_target_free[s], 0));
        GSB_CUDA_CHECK(c, cudaMemcpyAsync(c->t_target[s], host_targets[b], img_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        if (with_depth) {
            GSB_REQUIRE(c, host_depths[b] && host_masks[b], "gsb_trainer_accumulate_depth: null depth target / mask");
            GSB_CUDA_CHECK(c, cudaMemcpyAsync(c->t_dtarget[s], host_depths[b], (size_t)c->P * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
            GSB_CUDA_CHECK(c, cudaMemcpyAsync(c->t_dmask[s], host_masks[b], (size_t)c->P, cudaMemcpyHostToDevice, c->copy_stream));
        }
        GSB_CUDA_CHECK(c, cudaEventRecord(c->t_target_ready[s], c->copy_stream));
        c->stats.stage_calls[GSB_STAGE_H2D] += 1;
        return GSB_OK;
    };
    if (targets_on_host && B > 0) {
        int rc = prefetch(0);
        if (rc != GSB_OK) return rc;
    }
    // View pipeline: projection + binning ("front") of views b+1.. run on high-priority front streams (one per view
    // working set) while the FP32-bound rasteriser / loss / backward kernels ("back") of view b occupy the work
    // stream.  The front is HBM/latency-bound integer work, so it fills issue slots the back leaves idle.  Measured
    // on B200: more sets / more fronts in flight, or finishing every front before the first rasteriser, change the
    // step by < 1 % - the front's ~0.4 ms of GPU time per view is real work, not exposed latency.
    const RawParams rp{c->t_p[0], c->t_p[1], c->t_p[2], c->t_p[3], c->t_p[4], c->t_p[5]};
    const bool overlap = B >= 2 && !(c->cfg.flags & (GSB_FLAG_SORT_CUB | GSB_FLAG_NO_OVERLAP));
    c->saved.valid = false;
    c->bin_valid = false;
    c->bin_gen += 1;
    if (overlap) {   // the front stream must see everything already queued on the work stream (Adam of the last step)
        GSB_CUDA_CHECK(c, cudaEventRecord(c->ev_fork, c->stream));
        for (cudaStream_t fs : c->front_streams) GSB_CUDA_CHECK(c, cudaStreamWaitEvent(fs, c->ev_fork, 0));
    }
    int front_issued = -1;
    auto issue_front = [&](int b) -> int {
        Ctx::ViewBufs& v = c->vb[b % gsb::GSB_VIEW_SETS];
        const gsb::ViewParams vp = gsb::make_view(c, &host_cams[b]);
        cudaStream_t st = overlap ? c->front_streams[b % gsb::GSB_VIEW_SETS] : c->stream;
        if (overlap) GSB_CUDA_CHECK(c, cudaStreamWaitEvent(st, v.ev_back, 0));   // set free again
        // phased exchange: only the step's FIRST view splits its projection around the binning (its front is the one on the
        // critical path while the SH parameters arrive); the later views' fronts run ahead of their rasterisers anyway
        // and simply start once the SH rows have been announced
        if (sh_event && b > 0) GSB_CUDA_CHECK(c, cudaStreamWaitEvent(st, sh_event, 0));
        int rc = enqueue_front(c, v, st, N, rp, vp, nullptr, nullptr, b == 0 ? sh_event : nullptr);
        if (rc != GSB_OK) return rc;
        if (overlap) GSB_CUDA_CHECK(c, cudaEventRecord(v.ev_front, st));
        front_issued = b;
        return GSB_OK;
    };
    // measurement aid (GSB_DBG_FRONT_ALL=1, needs a build with GSB_VIEW_SETS_N >= B): every front of the batch is issued
    // first, each on its own stream, the work stream waits for all of them and the span is printed (synchronises)
    static const bool dbg_front_all = getenv("GSB_DBG_FRONT_ALL") != nullptr;
    if (dbg_front_all && overlap && B <= gsb::GSB_VIEW_SETS) {
        static cudaEvent_t ea = nullptr, eb = nullptr;
        if (!ea) { cudaEventCreate(&ea); cudaEventCreate(&eb); }
        cudaEventRecord(ea, c->stream);
        for (int b = 0; b < B; ++b) {
            int rc = issue_front(b);
            if (rc != GSB_OK) return rc;
        }
        for (int b = 0; b < B; ++b) GSB_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->vb[b % gsb::GSB_VIEW_SETS].ev_front, 0));
        cudaEventRecord(eb, c->stream);
        cudaEventSynchronize(eb);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ea, eb);
        fprintf(stderr, "[GSB_DBG_FRONT_ALL] %d fronts on %d streams, alone on the GPU: %.3f ms\n", B, gsb::GSB_VIEW_SETS, ms);
    }
    for (int b = 0; b < B; ++b) {
        GSB_REQUIRE(c, host_targets[b] != nullptr, "gsb_trainer_accumulate: null target");
        if (targets_on_host && b + 1 < B) {
            int rc = prefetch(b + 1);
            if (rc != GSB_OK) return rc;
        }
        Ctx::ViewBufs& v = c->vb[b % gsb::GSB_VIEW_SETS];
        for (int attempt = 0;; ++attempt) {
            int rc = GSB_OK;
            if (front_issued < b) rc = issue_front(b);
            for (int a = 1; a <= gsb::GSB_FRONT_AHEAD && rc == GSB_OK && overlap; ++a)
                if (b + a < B && front_issued < b + a) rc = issue_front(b + a);
            if (rc != GSB_OK) return rc;
            rc = gsb::finish_binning(c, v);
            if (rc == GSB_OK) break;
            if (rc != GSB_ERR_CAPACITY || attempt == 3) return rc;
            front_issued = b - 1;   // the regrow synchronised and reallocated every set: redo what was in flight
            if (overlap) {
                GSB_CUDA_CHECK(c, cudaEventRecord(c->ev_fork, c->stream));
                for (cudaStream_t fs : c->front_streams) GSB_CUDA_CHECK(c, cudaStreamWaitEvent(fs, c->ev_fork, 0));
            }
        }
        const gsb::ViewParams vp = gsb::make_view(c, &host_cams[b]);
        if (overlap) GSB_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, v.ev_front, 0));
        int rc = enqueue_raster_fwd(c, v, N, rp, vp, with_depth);
        if (rc != GSB_OK) return rc;
        const float* target = host_targets[b];
        DepthTerm dt;
        if (with_depth) {
            GSB_REQUIRE(c, host_depths[b] && host_masks[b], "gsb_trainer_accumulate_depth: null depth target / mask");
            dt.depth = c->out_depth; dt.target = host_depths[b]; dt.mask = host_masks[b]; dt.lambda = lambda_depth; dt.cot = c->cot_depth;
        }
        if (targets_on_host) {
            GSB_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->t_target_ready[b & 1], 0));
            target = c->t_target[b & 1];
            dt.target = c->t_dtarget[b & 1]; dt.mask = c->t_dmask[b & 1];
        }
        rc = loss_impl(c, c->out_color, target, grad_scale, c->cot_render, host_loss ? c->loss_accum : nullptr, with_depth ? &dt : nullptr);
        if (rc != GSB_OK) return rc;
        if (targets_on_host) GSB_CUDA_CHECK(c, cudaEventRecord(c->t_target_free[b & 1], c->stream));
        const int accumulate = (b > 0 || !zero_grads) ? 1 : 0;
        // peer step protocol: the last view's projection backward always goes to the tail stream, chunk by chunk, so that
        // the exchange of chunk k (work stream) overlaps the projection backward of chunk k + 1
        const bool last_sig = sig && b == B - 1;
        const bool split = overlap || (last_sig && !(c->cfg.flags & (GSB_FLAG_SORT_CUB | GSB_FLAG_NO_OVERLAP)));
        rc = render_backward_impl(c, c->cot_render, with_depth ? c->cot_depth : nullptr, nullptr, c->t_g[0], c->t_g[1], c->t_g[2], c->t_g[3], c->t_g[4],
                                  c->t_g[5], accumulate, split ? (b & 1) : -1, last_sig ? sig : nullptr);
        if (rc != GSB_OK) return rc;
        if (overlap) GSB_CUDA_CHECK(c, cudaEventRecord(v.ev_back, c->stream));   // raster backward done: the set is free
    }
    if (overlap && B > 0 && !sig)   // the gradients are complete when the last projection backward has run (tail stream is in order)
        GSB_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_pb[(B - 1) & 1], 0));   // (peer step protocol: per-chunk events instead)
    if (host_loss) {
        if (c->cfg.flags & GSB_FLAG_ASYNC_LOSS) {   // pinned destination, the caller synchronises when it wants the value
            GSB_CUDA_CHECK(c, cudaMemcpyAsync(host_loss, c->loss_accum, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        } else {
            GSB_CUDA_CHECK(c, cudaMemcpyAsync(c->h_loss, c->loss_accum, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            GSB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
            *host_loss = c->h_loss[0];
        }
    }
    return GSB_OK;
}

int gsb_trainer_accumulate(gsb_ctx* ctx, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                           int32_t targets_on_host, int32_t zero_grads, float grad_scale, float* host_loss)
{
    CTX_PROLOGUE(ctx);
    return trainer_accumulate_impl(c, B, host_cams, host_targets, nullptr, nullptr, 0.0f, targets_on_host, zero_grads, grad_scale, host_loss);
}

int gsb_trainer_accumulate_depth(gsb_ctx* ctx, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                                 const float* const* host_target_depths, const uint8_t* const* host_depth_masks, float lambda_depth,
                                 int32_t targets_on_host, int32_t zero_grads, float grad_scale, float* host_loss)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, B == 0 || (host_target_depths && host_depth_masks), "gsb_trainer_accumulate_depth: null argument");
    return trainer_accumulate_impl(c, B, host_cams, host_targets, host_target_depths, host_depth_masks, lambda_depth, targets_on_host,
                                   zero_grads, grad_scale, host_loss);
}

static void learning_rates(int iteration, int total, float* lrs)
{
    // GaussianModel.swift:56-65 (f32 arithmetic)
    const float frac = 1.0f - (float)iteration / (float)total;
    lrs[0] = 0.00016f * (frac > 0.01f ? frac : 0.01f);
    lrs[1] = 0.0025f;
    lrs[2] = 0.0025f / 20.0f;
    lrs[3] = 0.005f;
    lrs[4] = 0.001f;
    lrs[5] = 0.025f;
}

int gsb_trainer_apply(gsb_ctx* ctx, int32_t iteration, int32_t total_iterations, int32_t reset_state)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, total_iterations > 0, "gsb_trainer_apply: total_iterations must be positive");
    GSB_CUDA_CHECK(c, gsb::drain_peer_step(c));
    if (reset_state) GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_block + 2 * c->t_floats, 0, 2 * c->t_floats * sizeof(float), c->stream));
    gsb::AdamTensors t{};
    learning_rates(iteration, total_iterations, t.lr);
    for (int k = 0; k < 6; ++k) { t.p[k] = c->t_p[k]; t.g[k] = c->t_g[k]; t.m[k] = c->t_m[k]; t.v[k] = c->t_v[k]; t.count[k] = c->t_count[k]; }
    gsb::StageTimer tm(c, GSB_STAGE_ADAM);
    int launches = 0;
    GSB_CUDA_CHECK(c, gsb::launch_adam(c->stream, t, c->cfg.adam_beta1, c->cfg.adam_beta2, c->cfg.adam_eps, 1.0f, c->tN,
                                       c->t_accum, nullptr, &launches));
    c->stats.kernel_launches += launches;
    c->t_accum_steps += 1;   // addGradientAccumulation (GaussianTrainer.swift:724-742)
    return GSB_OK;
}

// ---- peer-memory data parallelism ----------------------------------------------------------------
struct PeerExport {                 // what one replica publishes: 3 IPC handles + the layout the others must share
    cudaIpcMemHandle_t block, accum, sync;
    int64_t n, floats, cap;
};
static_assert(sizeof(PeerExport) <= GSB_PEER_BLOB_BYTES, "GSB_PEER_BLOB_BYTES too small");
int gsb_trainer_peers_export(gsb_ctx* ctx, void* host_blob, int64_t blob_bytes)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, host_blob && blob_bytes >= (int64_t)sizeof(PeerExport), "gsb_trainer_peers_export: blob too small");
    PeerExport e{};
    if (!c->t_sync) {   // allocated once per context: survives densification (the slabs do not)
        GSB_CUDA_CHECK(c, cudaMalloc(reinterpret_cast<void**>(&c->t_sync), sizeof(gsb::PeerSync)));
        GSB_CUDA_CHECK(c, cudaMemset(c->t_sync, 0, sizeof(gsb::PeerSync)));
        GSB_CUDA_CHECK(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_peer_error), sizeof(uint32_t)));
        *c->h_peer_error = 0;
        for (cudaEvent_t& ev : c->ev_chunk) GSB_CUDA_CHECK(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        GSB_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->ev_sh, cudaEventDisableTiming));
        if (const char* env = getenv("GSB_PEER_CHUNKS")) c->peer_chunks = std::max(1, std::min(gsb::GSB_MAX_CHUNKS, atoi(env)));
        if (const char* env = getenv("GSB_PEER_PHASED")) c->peer_phased = atoi(env);
    }
    GSB_CUDA_CHECK(c, cudaIpcGetMemHandle(&e.block, c->t_block));
    GSB_CUDA_CHECK(c, cudaIpcGetMemHandle(&e.accum, c->t_accum));
    GSB_CUDA_CHECK(c, cudaIpcGetMemHandle(&e.sync, c->t_sync));
    e.n = c->tN; e.floats = (int64_t)c->t_floats; e.cap = c->t_cap;
    memset(host_blob, 0, (size_t)blob_bytes);
    memcpy(host_blob, &e, sizeof(e));
    return GSB_OK;
}

int gsb_trainer_peers_import(gsb_ctx* ctx, int32_t world, int32_t rank, const void* host_blobs, int64_t blob_bytes)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, world >= 1 && world <= gsb::GSB_MAX_PEERS && rank >= 0 && rank < world && host_blobs &&
                       blob_bytes >= (int64_t)sizeof(PeerExport),
                "gsb_trainer_peers_import: bad arguments (at most 8 replicas)");
    gsb::sync_all_streams(c);
    gsb::peers_close(c);
    for (int r = 0; r < world; ++r) {
        PeerExport e;
        memcpy(&e, static_cast<const char*>(host_blobs) + (size_t)r * (size_t)blob_bytes, sizeof(e));
        if (e.n != c->tN || e.floats != (int64_t)c->t_floats) {
            gsb::peers_close(c);
            gsb::set_error(c, "gsb_trainer_peers_import: replica layouts differ");
            return GSB_ERR_INVALID;
        }
        if (r == rank) {
            if (!c->t_sync) { gsb::peers_close(c); gsb::set_error(c, "gsb_trainer_peers_import: call gsb_trainer_peers_export first"); return GSB_ERR_STATE; }
            c->peer_block[r] = c->t_block; c->peer_accum[r] = c->t_accum; c->peer_sync[r] = c->t_sync;
            continue;
        }
        void* pb = nullptr; void* pa = nullptr; void* ps = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&pb, e.block, cudaIpcMemLazyEnablePeerAccess);
        if (err == cudaSuccess) err = cudaIpcOpenMemHandle(&pa, e.accum, cudaIpcMemLazyEnablePeerAccess);
        if (err == cudaSuccess) err = cudaIpcOpenMemHandle(&ps, e.sync, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            if (pb) cudaIpcCloseMemHandle(pb);
            if (pa) cudaIpcCloseMemHandle(pa);
            gsb::peers_close(c);
            gsb::set_error(c, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(err));
            cudaGetLastError();
            return GSB_ERR_CUDA;
        }
        c->peer_block[r] = static_cast<float*>(pb); c->peer_accum[r] = static_cast<float*>(pa);
        c->peer_sync[r] = static_cast<gsb::PeerSync*>(ps); c->peer_opened[r] = true;
    }
    c->peer_world = world; c->peer_rank = rank;
    // A new generation of the step protocol: nobody writes into this block any more (every replica closed its mappings
    // and passed a barrier before exporting again) and nobody starts a step before all replicas have imported.
    GSB_CUDA_CHECK(c, cudaMemset(c->t_sync, 0, sizeof(gsb::PeerSync)));
    c->peer_step_id = 0;
    c->peer_wait_step = 0;
    return GSB_OK;
}

int gsb_trainer_peers_close(gsb_ctx* ctx)
{
    CTX_PROLOGUE(ctx);
    GSB_CUDA_CHECK(c, gsb::drain_peer_step(c));
    gsb::sync_all_streams(c);
    gsb::peers_close(c);
    return GSB_OK;
}

// Owned slice of the Gaussians [n0, n1) for replica R of W: equal parts on multiples of 4 Gaussians (every tensor slice then
// starts 16-byte aligned).  Fills the per-launch descriptors of the exchange kernels (adam.cu).
// tensor_mask: bit k = tensor k of (xyz, f_dc, f_rest, scales, rotation, opacity) takes part; D1 goes with xyz (bit 0)
constexpr int TENSORS_ALL = 0x3f, TENSORS_GEOMETRY = 0x39, TENSORS_SH = 0x06;
static void fill_exchange(Ctx* c, int W, int R, long long n0, long long n1, int iteration, int total_iterations, const float* param_base,
                          gsb::AdamTensors& t, gsb::AdamPeers& pr, int tensor_mask = TENSORS_ALL)
{
    const long long n = n1 - n0;
    const long long per = ((n + W - 1) / W + 3) & ~3LL;
    const long long g0 = n0 + std::min<long long>((long long)R * per, n), g1 = n0 + std::min<long long>((long long)R * per + per, n);
    learning_rates(iteration, total_iterations, t.lr);
    const int K = c->cfg.sh_coeffs;
    const long long row[6] = {3, 3, (long long)(K - 1) * 3, 3, 4, 1};
    for (int k = 0; k < 6; ++k) {
        t.p[k] = c->t_p[k]; t.g[k] = c->t_g[k]; t.m[k] = c->t_m[k]; t.v[k] = c->t_v[k];
        t.count[k] = ((tensor_mask >> k) & 1) ? (g1 - g0) * row[k] : 0;
        pr.first[k] = g0 * row[k];
        pr.tensor_off[k] = (long long)(c->t_p[k] - param_base);
    }
    pr.g0 = g0; pr.g1 = g1; pr.world = W; pr.rank = R;
    pr.accum_all = 0;
    const bool d1 = (tensor_mask & 1) != 0;
    if (!d1) pr.g1 = pr.g0;   // no D1 in this launch
    if (c->peer_world == W && c->peer_accum[R] == c->t_accum) {   // the replicas' slabs are peer-mapped
        for (int r = 0; r < W; ++r) {
            pr.params[r] = c->peer_block[r];
            pr.grads[r] = c->peer_block[r] + c->t_floats;
            pr.accum[r] = d1 ? c->peer_accum[r] : nullptr;
        }
        pr.accum_all = 1;
    } else {
        pr.accum[R] = d1 ? c->t_accum : nullptr;
    }
}

int gsb_trainer_apply_peers(gsb_ctx* ctx, int32_t iteration, int32_t total_iterations, int32_t reset_state)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, total_iterations > 0, "gsb_trainer_apply_peers: total_iterations must be positive");
    if (c->peer_world < 1 || c->peer_block[c->peer_rank] != c->t_block || c->sym) {
        gsb::set_error(c, "gsb_trainer_apply_peers: peers not imported for the current trainer state");
        return GSB_ERR_STATE;
    }
    if (reset_state) GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_block + 2 * c->t_floats, 0, 2 * c->t_floats * sizeof(float), c->stream));
    gsb::AdamTensors t{};
    gsb::AdamPeers pr{};
    fill_exchange(c, c->peer_world, c->peer_rank, 0, c->tN, iteration, total_iterations, c->t_block, t, pr);
    gsb::StageTimer tm(c, GSB_STAGE_ADAM);
    int launches = 0;
    GSB_CUDA_CHECK(c, gsb::launch_adam_peers(c->stream, t, pr, c->cfg.adam_beta1, c->cfg.adam_beta2, c->cfg.adam_eps, 1.0f, nullptr, c->peer_blocks, &launches));
    c->stats.kernel_launches += launches;
    c->t_accum_steps += 1;
    return GSB_OK;
}

// ---- the whole data-parallel step with device-side synchronisation ---------------------------------
// wait (previous step's parameter stores have landed, every owner is done with my gradients)
//   -> views (the last view's projection backward chunk by chunk, each chunk announced to every replica)
//   -> per chunk: exchange kernel (waits for every replica's announcement, reduces + Adam + stores, announces back)
// No host barrier and no NCCL call on the step: the replicas only meet in the flags.
int gsb_trainer_step_peers(gsb_ctx* ctx, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                           int32_t targets_on_host, float grad_scale, int32_t iteration, int32_t total_iterations, int32_t reset_state,
                           float* host_loss)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, total_iterations > 0 && B >= 0, "gsb_trainer_step_peers: bad arguments");
    const int W = c->peer_world, R = c->peer_rank;
    if (W < 1 || c->peer_block[R] != c->t_block || !c->t_sync) {
        gsb::set_error(c, "gsb_trainer_step_peers: peers not imported for the current trainer state");
        return GSB_ERR_STATE;
    }
    const uint32_t step = ++c->peer_step_id;
    ChunkSignal sig;
    sig.step = step;
    sig.chunks = std::max(1, std::min(c->peer_chunks, gsb::GSB_MAX_CHUNKS));
    {   // chunk boundaries on multiples of 128 Gaussians (one projection CTA, 16-byte aligned tensor rows)
        const long long per = (((long long)c->tN + sig.chunks - 1) / sig.chunks + 127) & ~127LL;
        for (int k = 0; k <= sig.chunks; ++k) sig.begin[k] = std::min<long long>((long long)k * per, c->tN);
    }
    int rc = trainer_accumulate_impl(c, B, host_cams, host_targets, nullptr, nullptr, 0.0f, targets_on_host, 1, grad_scale, host_loss, &sig);
    if (rc != GSB_OK) return rc;
    // The exchange kernels run on their own stream: what they wait for (every replica's gradients of a chunk) and what
    // waits for them (the next batch, through the "parameters written" flags - this replica announces to itself as well)
    // is expressed in flags, so the work stream is free to start the next step's first front as soon as the geometry rows
    // are in, while this replica's own SH exchange kernel is still running.
    cudaStream_t xs = c->xchg_stream;
    if (reset_state) GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_block + 2 * c->t_floats, 0, 2 * c->t_floats * sizeof(float), xs));
    const float* param_base = c->sym ? c->sym_params : c->t_block;
    // Phased: the geometry tensors (19 % of the bytes) of every chunk first, then the SH tensors - the next step's
    // projection + binning only need the former (enqueue_front), so most of the exchange hides behind them.
    const bool phased = c->peer_phased != 0 && 2 * sig.chunks <= gsb::GSB_MAX_CHUNKS;
    for (int phase = 0; phase < (phased ? 2 : 1); ++phase) {
        for (int k = 0; k < sig.chunks; ++k) {
            GSB_CUDA_CHECK(c, cudaStreamWaitEvent(xs, c->ev_chunk[k], 0));   // my own chunk k is done (tail / work stream)
            gsb::AdamTensors t{};
            gsb::AdamPeers pr{};
            fill_exchange(c, W, R, sig.begin[k], sig.begin[k + 1], iteration, total_iterations, param_base, t, pr,
                          phased ? (phase == 0 ? TENSORS_GEOMETRY : TENSORS_SH) : TENSORS_ALL);
            const int row = phase * sig.chunks + k;
            gsb::PeerStepSync sy;
            sy.wait_flags = &c->t_sync->grads_ready[k][0];
            sy.done = &c->t_sync->done[row];
            for (int r = 0; r < W; ++r) sy.announce[r] = &c->peer_sync[r]->params_ready[row][R];
            sy.error = &c->t_sync->error;
            sy.step = step;
            gsb::StageTimer tm(c, GSB_STAGE_ADAM, xs);
            int launches = 0;
            if (c->sym)
                GSB_CUDA_CHECK(c, gsb::launch_adam_multicast(xs, t, pr, c->mc_grads, c->mc_params, c->cfg.adam_beta1, c->cfg.adam_beta2,
                                                             c->cfg.adam_eps, 1.0f, c->tN, &sy, c->mc_blocks, &launches));
            else
                GSB_CUDA_CHECK(c, gsb::launch_adam_peers(xs, t, pr, c->cfg.adam_beta1, c->cfg.adam_beta2, c->cfg.adam_eps, 1.0f, &sy,
                                                         c->peer_blocks, &launches));
            c->stats.kernel_launches += launches;
        }
    }
    c->t_accum_steps += 1;
    c->peer_wait_step = step;        // the next batch (fused or not) starts by waiting for every replica's stores of this step
    c->peer_wait_chunks = phased ? 2 * sig.chunks : sig.chunks;
    c->peer_wait_phased = phased;
    return GSB_OK;
}

int gsb_trainer_peers_tune(gsb_ctx* ctx, int32_t chunks, int32_t peer_blocks, int32_t multicast_blocks)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, chunks >= 0 && chunks <= gsb::GSB_MAX_CHUNKS && peer_blocks >= 0 && multicast_blocks >= 0, "gsb_trainer_peers_tune: bad arguments");
    GSB_CUDA_CHECK(c, gsb::drain_peer_step(c));
    gsb::sync_all_streams(c);
    if (chunks > 0) c->peer_chunks = chunks;
    if (const char* e = getenv("GSB_PEER_PHASED")) c->peer_phased = atoi(e);   // 0: one exchange kernel per chunk for all six tensors
    c->peer_blocks = peer_blocks;
    c->mc_blocks = multicast_blocks;
    return GSB_OK;
}

// A bounded wait of the step protocol ran out on this replica (a replica died or the replicas disagree about the step
// count): the results since then are invalid.  Reads one word; synchronises the context.
int gsb_trainer_peers_check(gsb_ctx* ctx)
{
    CTX_PROLOGUE(ctx);
    if (!c->t_sync) return GSB_OK;
    GSB_CUDA_CHECK(c, gsb::drain_peer_step(c));
    gsb::sync_all_streams(c);
    GSB_CUDA_CHECK(c, cudaMemcpy(c->h_peer_error, &c->t_sync->error, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (*c->h_peer_error) {
        gsb::set_error(c, "gsb_trainer_step_peers: a wait for another replica timed out (2 s); the step results are invalid");
        return GSB_ERR_STATE;
    }
    return GSB_OK;
}

int gsb_trainer_attach_symmetric(gsb_ctx* ctx, int32_t world, int32_t rank, float* params_local, float* grads_local,
                                 float* params_mc, float* grads_mc, int64_t floats)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, world >= 1 && world <= gsb::GSB_MAX_PEERS && rank >= 0 && rank < world && params_local && grads_local &&
                       params_mc && grads_mc && floats >= (int64_t)c->t_floats,
                "gsb_trainer_attach_symmetric: bad arguments (buffers must hold gsb_trainer_grad_block's float count)");
    GSB_REQUIRE(c, ((reinterpret_cast<uintptr_t>(params_local) | reinterpret_cast<uintptr_t>(grads_local) |
                     reinterpret_cast<uintptr_t>(params_mc) | reinterpret_cast<uintptr_t>(grads_mc)) & 127u) == 0,
                "gsb_trainer_attach_symmetric: buffers must be 128-byte aligned");
    gsb::sync_all_streams(c);
    const float* cur_params = c->t_p[0];   // off[0] == 0: base of the current parameter copy
    if (cur_params != params_local)
        GSB_CUDA_CHECK(c, cudaMemcpyAsync(params_local, cur_params, c->t_floats * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    GSB_CUDA_CHECK(c, cudaMemsetAsync(grads_local, 0, c->t_floats * sizeof(float), c->stream));
    for (int k = 0; k < 6; ++k) {
        const ptrdiff_t off = c->t_m[k] - (c->t_block + 2 * c->t_floats);   // segment offset inside one copy
        c->t_p[k] = params_local + off;
        c->t_g[k] = grads_local + off;
    }
    c->sym = true; c->sym_world = world; c->sym_rank = rank;
    c->sym_params = params_local; c->sym_grads = grads_local; c->mc_params = params_mc; c->mc_grads = grads_mc;
    GSB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return GSB_OK;
}

int gsb_trainer_apply_multicast(gsb_ctx* ctx, int32_t iteration, int32_t total_iterations, int32_t reset_state)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    GSB_REQUIRE(c, total_iterations > 0, "gsb_trainer_apply_multicast: total_iterations must be positive");
    if (!c->sym) { gsb::set_error(c, "gsb_trainer_apply_multicast: no symmetric buffers attached to the current trainer state"); return GSB_ERR_STATE; }
    if (reset_state) GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_block + 2 * c->t_floats, 0, 2 * c->t_floats * sizeof(float), c->stream));
    gsb::AdamTensors t{};
    gsb::AdamPeers pr{};
    fill_exchange(c, c->sym_world, c->sym_rank, 0, c->tN, iteration, total_iterations, c->sym_params, t, pr);
    gsb::StageTimer tm(c, GSB_STAGE_ADAM);
    int launches = 0;
    GSB_CUDA_CHECK(c, gsb::launch_adam_multicast(c->stream, t, pr, c->mc_grads, c->mc_params, c->cfg.adam_beta1, c->cfg.adam_beta2,
                                                 c->cfg.adam_eps, 1.0f, c->tN, nullptr, c->mc_blocks, &launches));
    c->stats.kernel_launches += launches;
    c->t_accum_steps += 1;
    return GSB_OK;
}

int gsb_train_step(gsb_ctx* ctx, int32_t B, const gsb_camera* host_cams, const float* const* host_targets,
                   int32_t targets_on_host, int32_t iteration, int32_t total_iterations, float* host_loss)
{
    GSB_REQUIRE(C(ctx), B > 0, "gsb_train_step: B must be positive");
    int rc = gsb_trainer_accumulate(ctx, B, host_cams, host_targets, targets_on_host, 1, 1.0f / (float)B, host_loss);
    if (rc != GSB_OK) return rc;
    return gsb_trainer_apply(ctx, iteration, total_iterations, 0);
}

// ---- densification -------------------------------------------------------------------------------
int gsb_densify_classify(gsb_ctx* ctx, int32_t N, const float* grad_accum, float denom, const float* scales_log,
                         const float* opacity_logit, float grad_threshold, float max_scale, float min_opacity,
                         int32_t allow_densify, int32_t* actions, int32_t* output_counts)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && grad_accum && scales_log && opacity_logit && actions && output_counts, "gsb_densify_classify: null argument");
    GSB_CUDA_CHECK(c, gsb::launch_densify_classify(c->stream, N, grad_accum, denom, scales_log, opacity_logit, grad_threshold, max_scale,
                                                   min_opacity, allow_densify, actions, output_counts, nullptr));
    c->stats.kernel_launches += N > 0;
    return GSB_OK;
}

int gsb_densify_map(gsb_ctx* ctx, int32_t N, const int32_t* actions, const int32_t* output_counts, int32_t* offsets, int32_t capacity,
                    int32_t* gather_indices, int32_t* noise_mode, int32_t* host_total)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, N >= 0 && capacity >= 0 && actions && output_counts && gather_indices && noise_mode, "gsb_densify_map: null argument");
    uint32_t* off = reinterpret_cast<uint32_t*>(offsets);
    uint32_t* tmp = nullptr;
    void* ws = nullptr;
    if (!off) { GSB_CUDA_CHECK(c, gsb::dev_alloc(&tmp, (size_t)N)); off = tmp; }
    cudaError_t e = cudaMalloc(&ws, gsb::scan_ws_bytes(N));
    Ctx::ViewBufs& v = c->vb[c->cur];
    if (e == cudaSuccess) e = gsb::launch_exclusive_scan(c->stream, N, reinterpret_cast<const uint32_t*>(output_counts), nullptr, nullptr, nullptr,
                                                         off, &v.d_ctl[2], ws);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&v.h_ctl[2], &v.d_ctl[2], 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = gsb::launch_densify_map(c->stream, N, actions, off, (uint32_t)capacity, gather_indices, noise_mode);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (tmp) cudaFree(tmp);
    if (ws) cudaFree(ws);
    GSB_CUDA_CHECK(c, e);
    if (host_total) *host_total = (int32_t)v.h_ctl[2];
    c->stats.kernel_launches += 2;
    if ((int64_t)v.h_ctl[2] > capacity) { gsb::set_error(c, "gsb_densify_map: output map larger than capacity"); return GSB_ERR_CAPACITY; }
    return GSB_OK;
}

int gsb_densify_apply(gsb_ctx* ctx, int32_t N_out, const int32_t* gather_indices, const int32_t* noise_mode, const float* base_noise,
                      uint64_t seed, const float* xyz, const float* f_dc, const float* f_rest, const float* scales_log,
                      const float* rot_raw, const float* opacity_logit, float* o_xyz, float* o_f_dc, float* o_f_rest, float* o_scales_log,
                      float* o_rot_raw, float* o_opacity_logit)
{
    CTX_PROLOGUE(ctx);
    const int K = c->cfg.sh_coeffs;
    GSB_REQUIRE(c, N_out >= 0 && gather_indices && noise_mode && xyz && f_dc && (f_rest || K == 1) && scales_log && rot_raw &&
                       opacity_logit && o_xyz && o_f_dc && (o_f_rest || K == 1) && o_scales_log && o_rot_raw && o_opacity_logit,
                "gsb_densify_apply: null argument");
    const float* in6[6] = {xyz, f_dc, f_rest, scales_log, rot_raw, opacity_logit};
    float* out6[6] = {o_xyz, o_f_dc, o_f_rest, o_scales_log, o_rot_raw, o_opacity_logit};
    GSB_CUDA_CHECK(c, gsb::launch_densify_apply(c->stream, N_out, K, gather_indices, noise_mode, base_noise, seed, in6, out6));
    c->stats.kernel_launches += N_out > 0;
    return GSB_OK;
}

int gsb_trainer_count(gsb_ctx* ctx, int32_t* host_N, int32_t* host_accum_steps)
{
    CTX_PROLOGUE(ctx);
    if (host_N) *host_N = c->tN;
    if (host_accum_steps) *host_accum_steps = c->t_accum_steps;
    return GSB_OK;
}

int gsb_trainer_densify(gsb_ctx* ctx, float grad_threshold, float max_scale, float min_opacity, int32_t max_gaussians, uint64_t seed,
                        const float* base_noise, int32_t* host_counts5)
{
    CTX_PROLOGUE(ctx);
    if (c->tN == 0) { gsb::set_error(c, "trainer not initialised"); return GSB_ERR_STATE; }
    const int N = c->tN, K = c->cfg.sh_coeffs;
    if (c->peer_world > 0) {
        // densification swaps - and, when the capacity grows, frees - the slab the other replicas have mapped through CUDA
        // IPC (freeing exported memory that an importer still has open is undefined behaviour)
        gsb::set_error(c, "gsb_trainer_densify: peer mappings are open - every replica must call gsb_trainer_peers_close and pass a "
                          "barrier first, and exchange the new slabs (export / import) afterwards");
        return GSB_ERR_STATE;
    }
    gsb::sync_all_streams(c);
    // scratch: actions | counts | offsets in the [capN,12] float parity buffer (capN >= N)
    int* actions = reinterpret_cast<int*>(c->act_tmp);
    int* counts = actions + N;
    uint32_t* offsets = reinterpret_cast<uint32_t*>(counts + N);
    Ctx::ViewBufs& v = c->vb[0];
    uint32_t* d_stats = reinterpret_cast<uint32_t*>(c->loss_accum);   // 4 words of scratch
    const int allow = N < max_gaussians ? 1 : 0;   // GaussianTrainer.swift:785
    GSB_CUDA_CHECK(c, gsb::launch_densify_classify(c->stream, N, c->t_accum, (float)c->t_accum_steps, c->t_p[3], c->t_p[5], grad_threshold,
                                                   max_scale, min_opacity, allow, actions, counts, d_stats));
    GSB_CUDA_CHECK(c, gsb::launch_exclusive_scan(c->stream, N, reinterpret_cast<const uint32_t*>(counts), nullptr, nullptr, nullptr, offsets,
                                                 &v.d_ctl[2], v.scan_ws));
    uint32_t h[5] = {0, 0, 0, 0, 0};
    GSB_CUDA_CHECK(c, cudaMemcpyAsync(h, d_stats, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    GSB_CUDA_CHECK(c, cudaMemcpyAsync(&h[4], &v.d_ctl[2], sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    GSB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    GSB_CUDA_CHECK(c, cudaMemsetAsync(c->loss_accum, 0, 4 * sizeof(float), c->stream));
    c->stats.kernel_launches += 2;
    if (host_counts5) for (int i = 0; i < 5; ++i) host_counts5[i] = (int32_t)h[i];
    const uint32_t total = h[4];
    const bool unchanged = total == 0 || (h[1] == 0 && h[2] == 0 && h[3] == 0);   // :820-838
    if (unchanged) {   // resetGradientAccumulation on every exit path
        GSB_CUDA_CHECK(c, cudaMemsetAsync(c->t_accum, 0, (size_t)N * sizeof(float), c->stream));
        c->t_accum_steps = 0;
        return GSB_OK;
    }
    const int Nout = (int)total;
    if (c->cfg.max_gaussians > 0 && Nout > c->cfg.max_gaussians) {
        gsb::set_error(c, "gsb_trainer_densify: output count exceeds gsb_config.max_gaussians (use 0 = grow on demand)");
        return GSB_ERR_CAPACITY;
    }
    // destination slab: the other one; (re)allocated only when it does not exist yet or the capacity has to grow
    const int dst = c->t_cur ^ 1;
    const TrainerLayout L = trainer_layout(Nout, K);
    bool grew = false;
    if (Nout > c->t_cap) {
        c->t_cap = std::max(Nout, c->t_cap + c->t_cap / 2);
        grew = true;
    }
    if (grew || !c->t_slab[dst]) GSB_CUDA_CHECK(c, trainer_alloc_slab(c, dst, c->t_cap));
    float* block = c->t_slab[dst];
    float* accum = c->t_accum_slab[dst];
    // gather map in the destination's (still unused) Adam-state region: 2 * Nout ints << 2 * L.floats floats
    int* gather = reinterpret_cast<int*>(block + 2 * L.floats);
    int* mode = gather + Nout;
    GSB_CUDA_CHECK(c, cudaMemsetAsync(block + L.floats, 0, L.floats * sizeof(float), c->stream));   // grads
    GSB_CUDA_CHECK(c, cudaMemsetAsync(accum, 0, (size_t)Nout * sizeof(float), c->stream));
    GSB_CUDA_CHECK(c, gsb::launch_densify_map(c->stream, N, actions, offsets, (uint32_t)Nout, gather, mode));
    {
        const float* in6[6] = {c->t_p[0], c->t_p[1], c->t_p[2], c->t_p[3], c->t_p[4], c->t_p[5]};
        float* out6[6];
        for (int k = 0; k < 6; ++k) out6[k] = block + L.off[k];
        GSB_CUDA_CHECK(c, gsb::launch_densify_apply(c->stream, Nout, K, gather, mode, base_noise, seed, in6, out6));
    }
    GSB_CUDA_CHECK(c, cudaMemsetAsync(block + 2 * L.floats, 0, 2 * L.floats * sizeof(float), c->stream));   // m, v (after the map is consumed)
    c->stats.kernel_launches += 2;
    // commit (GaussianTrainer.swift:899-907) + fresh optimiser state (:1104-1109)
    const int old = c->t_cur;
    trainer_adopt(c, Nout, L, dst);
    if (grew) {   // the old slab is too small to ever be a destination again
        GSB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
        gsb::dev_free(c->t_slab[old]); gsb::dev_free(c->t_accum_slab[old]);
    }
    c->saved.valid = false;
    c->bin_valid = false;
    return gsb::ensure_gaussians(c, Nout);
}

// ---- stats ---------------------------------------------------------------------------------------
int gsb_last_contrib_sum(gsb_ctx* ctx, uint64_t* host_out)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, host_out, "gsb_last_contrib_sum: null argument");
    if (!c->saved.valid) { gsb::set_error(c, "gsb_last_contrib_sum: no forward saved on this context"); return GSB_ERR_STATE; }
    unsigned long long* d = reinterpret_cast<unsigned long long*>(c->partial);   // 16-byte scratch
    GSB_CUDA_CHECK(c, gsb::launch_sum_u32(c->stream, (size_t)c->P, c->out_last, d));
    unsigned long long h = 0;
    GSB_CUDA_CHECK(c, cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    GSB_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    *host_out = h;
    c->stats.kernel_launches += 1;
    return GSB_OK;
}

int gsb_tile_list_info(gsb_ctx* ctx, int32_t* sb_w, int32_t* sb_h, int32_t* num_superblocks, int32_t* sort_passes)
{
    CTX_PROLOGUE(ctx);
    if (sb_w) *sb_w = gsb::SBW;
    if (sb_h) *sb_h = gsb::SBH;
    if (num_superblocks) *num_superblocks = c->numSB;
    if (sort_passes) *sort_passes = (c->sbBits + 7) / 8;
    return GSB_OK;
}

int gsb_stats_reset(gsb_ctx* ctx)
{
    CTX_PROLOGUE(ctx);
    gsb::resolve_stage_events(c);
    const uint64_t cap = c->stats.pair_capacity;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.pair_capacity = cap;
    return GSB_OK;
}
int gsb_stats_get(gsb_ctx* ctx, gsb_stats* host_out)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, host_out, "gsb_stats_get: null argument");
    gsb::resolve_stage_events(c);
    *host_out = c->stats;
    return GSB_OK;
}
int gsb_enable_stage_timing(gsb_ctx* ctx, int32_t on)
{
    CTX_PROLOGUE(ctx);
    gsb::resolve_stage_events(c);
    c->timing = on != 0;
    return GSB_OK;
}
const char* gsb_stage_name(int32_t stage)
{
    return (stage >= 0 && stage < GSB_STAGE_COUNT) ? gsb::kStageNames[stage] : "";
}
const char* gsb_stage_section(int32_t stage)
{
    return (stage >= 0 && stage < GSB_STAGE_COUNT) ? gsb::kStageSections[stage] : "";
}

int gsb_bin_generation(gsb_ctx* ctx, uint64_t* host_out)
{
    CTX_PROLOGUE(ctx);
    GSB_REQUIRE(c, host_out, "gsb_bin_generation: null argument");
    *host_out = c->bin_gen;
    return GSB_OK;
}

}  // extern "C"
