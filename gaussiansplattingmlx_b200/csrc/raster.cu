// raster.cu — tile alpha-blend forward (K9) and back-to-front backward (K10) for sm_100a.
//
// Reference: slang/gaussian_tile_global_kernels.slang:437-614 (forward), :501-521 + :648-881
// (backward), call sites Trainer/GaussianRenderer.swift:124-147,187-226.
//
// Both kernels are bound by their packed FP32 arithmetic (tools/microbench/issue_model.cu: an FFMA2 with two or three
// vector operands holds the FMA pipe for 2.8-3.8 cycles per warp and nothing issues in its shadow; ncu: issue slots
// ~62 % busy, insensitive to occupancy), so the design goal is the least arithmetic per (pixel, Gaussian) evaluation:
//   * a tile's list is (tile_ranges, Gaussian indices in depth order) from tilelists.cu; the 48-byte records are
//     gathered from the L2-resident record table straight into shared memory by 16-byte async copies (LDGSTS)
//     completing on mbarriers, double-buffered;
//   * records carry the conic/opacity pre-folded into log2 units (common.cuh), so
//     alpha = min(0.99, ex2(A dx^2 + B dx dy + C dy^2 + lo));
//   * a thread owns a COLUMN of pixels (forward 4, backward 8): the record is read once per column and the exponent
//     is a quadratic in the compile-time row offset, 2 FFMA per pixel;
//   * forward: 64 threads per 16x16 block; a terminated pixel keeps running with T = 0 (adds exact zeros), the
//     termination bookkeeping is a per-pixel replay in the epilogue: 15 instructions per evaluation;
//   * backward: ONE WARP per 16x16 block.  The per-Gaussian gradient sums are first accumulated over the thread's 8
//     pixels in registers (geometry as three moments of h), parked in shared memory and summed across lanes by
//     column reads every 3 Gaussians; one vector red per (block, Gaussian, quad) then goes to L2.  Records flagged
//     "cannot reach the alpha clamp" take a path without the clamp logic; the sparse tail of an item (<= 32 of the 256
//     pixels still active) runs one pixel per lane;
//   * both kernels are persistent: CTAs pull blocks heavy-first from a device counter.
// Work per evaluation: forward 27 flop + 1 ex2; backward ~80 flop + 1 ex2 + 1 rcp.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "kernels.h"

namespace gsb {

constexpr int RT = 64;           // forward: threads per CTA (16x16 pixels, 4 per thread)
constexpr int RB_FWD = 128;      // records per forward batch (6 KB)
#ifndef GSB_FWD_CTAS
#define GSB_FWD_CTAS 12
#endif
#ifndef GSB_RB_BWD
#define GSB_RB_BWD 64
#endif
#ifndef GSB_BWD_WARPS
#define GSB_BWD_WARPS 16
#endif
#ifndef GSB_BWD_COMPACT
#define GSB_BWD_COMPACT 1    // the sparse tail of a backward item one pixel per lane (0: every Gaussian on the dense 8-rows-per-lane body)
#endif
constexpr int RB_BWD = GSB_RB_BWD;   // records per backward batch
constexpr int BWD_WARPS = GSB_BWD_WARPS;   // resident one-warp CTAs per SM the backward is compiled for (register budget)
constexpr int BPPT = 8;          // backward: pixels per thread (rows)
constexpr int BGRP = 3;          // backward: Gaussians per shared-memory reduction round (3 x 10 sums <= 32 lanes)
constexpr int CK = 256;          // Gaussians between transmittance/colour checkpoints (forward) = backward segment length
constexpr int CK_MAX = 16;       // checkpoints per block (a longer list ends in one long backward item)
constexpr int BPAD = 36;         // padded column length (floats): 144-byte stride keeps the LDS.128 of 8 lanes on distinct banks
// Gather staging of the 48-byte records: three 16-byte cp.async (LDGSTS) per record completing on the batch's mbarrier.
// One TMA bulk copy per record (cp.async.bulk, 48 bytes, UBLKCP) was measured and removed: a list is a GATHER of
// scattered records, so TMA can only move 48 bytes per instruction and its per-copy issue cost loses to LDGSTS - forward
// 0.452 vs 0.430 ms, backward 0.793 vs 0.764 ms on the same box (profiles/r2/r2h_forward_variants_and_tma_gather.txt).
// TMA is used where the bytes are contiguous: the projection kernels stage 128 Gaussians per CTA with bulk copies.

// (tile, 16x16 sub-block) of this CTA; tiles larger than 16x16 are covered by several CTAs
struct BlockMap {
    int tile, x0, y0;   // pixel origin of the 16x16 block
    int xmax, ymax;     // exclusive pixel bounds of the tile clipped to the image
};
__device__ __forceinline__ BlockMap map_block(const ViewParams& vp, const uint32_t* __restrict__ tile_order, uint32_t work)
{
    const int subX = (vp.tileW + 15) >> 4, subY = (vp.tileH + 15) >> 4;
    const int per = subX * subY;
    BlockMap m;
    const int slot = (int)work / per;
    const int sb = (int)work - slot * per;
    m.tile = (int)tile_order[slot];   // heavy tiles first (binning.cu k_tile_order)
    const int tileX = m.tile % vp.gridW, tileY = m.tile / vp.gridW;
    m.x0 = tileX * vp.tileW + (sb % subX) * 16;
    m.y0 = tileY * vp.tileH + (sb / subX) * 16;
    m.xmax = min((tileX + 1) * vp.tileW, vp.W);
    m.ymax = min((tileY + 1) * vp.tileH, vp.H);
    return m;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// A thread owns a column of FPPT = 4 vertically adjacent pixels; a warp covers 16 x 8 pixels and the
// 64-thread CTA one 16 x 16 block.  Per Gaussian a thread reads the record ONCE (4 LDS instead of 16)
// and evaluates the exponent as a quadratic in the compile-time row offset r:
//     e(r) = E0 + r*E1 + r^2*C          (log2 units, log2(opacity) folded into E0)
// so a row costs 2 FFMA + ex2 + min + the blend FMAs.  Termination bookkeeping is kept out of the inner
// loop: a pixel whose transmittance drops below 1e-4 just gets T = 0 (all later Gaussians then add
// exact zeros); lastContrib and the transmittance AT termination are recovered once per pixel in the
// epilogue by replaying the (at most FCHUNK) Gaussians of the chunk in which the pixel died, starting
// from the transmittance saved at that chunk's start, with bit-identical arithmetic (fwd_exponents).
constexpr int FPPT = 4;          // forward: pixels (rows) per thread
#ifndef GSB_FCHUNK
#define GSB_FCHUNK 8
#endif
constexpr int FCHUNK = GSB_FCHUNK;   // Gaussians between termination cuts / warp votes
static_assert(RB_FWD % FCHUNK == 0, "batch must be a whole number of chunks");
static_assert(CK % RB_FWD == 0 && CK % RB_BWD == 0, "checkpoints sit on batch boundaries");
static_assert(RB_BWD == 32 || RB_BWD == 64, "a lane stages one or two records per backward batch");

struct FwdExp {
    float E0, E1, C;
};
// shared by the blend loop and the termination replay: identical instruction sequence => identical bits
__device__ __forceinline__ FwdExp fwd_exponents(const float4& a, float C, float lo, float pxf, float pyf)
{
    const float dx = pxf - a.x, dyb = pyf - a.y;
    const float Cd = __fmul_rn(C, dyb);
    FwdExp e;
    e.E0 = fmaf(dx, fmaf(a.z, dx, __fmul_rn(a.w, dyb)), fmaf(Cd, dyb, lo));
    e.E1 = fmaf(a.w, dx, __fadd_rn(Cd, Cd));
    e.C = C;
    return e;
}
template <int R>
__device__ __forceinline__ float fwd_alpha(const FwdExp& e)
{
    // same two roundings per row as the packed f32x2 evaluation in the blend loop (row 0 included: fma(0, x, E0) = E0)
    const float rf = (float)R;
    const float p = fmaf(rf, fmaf(rf, e.C, e.E1), e.E0);
    return fminf(ex2_approx(p), 0.99f);
}
__device__ __forceinline__ float fwd_alpha_rt(const FwdExp& e, int r)
{
    switch (r) {
        case 0: return fwd_alpha<0>(e);
        case 1: return fwd_alpha<1>(e);
        case 2: return fwd_alpha<2>(e);
        default: return fwd_alpha<3>(e);
    }
}

// Persistent: the grid is (SM count x resident CTAs per SM); a CTA pulls 16x16 blocks from a device counter in
// heavy-first order until none is left (dynamic balancing of the tail; 8160 block launches become 1776).
// The trainer runs the projection/binning kernels of the NEXT view concurrently on a second stream (api.cu);
// measured on B200, capping the residency below the register limit to leave them room does not pay
// (tools/sweep_res.sh: 12/16 CTAs per SM beats 9/12), so the defaults fill the SM.
template <bool DEPTH>
__global__ void __launch_bounds__(RT, GSB_FWD_CTAS) k_raster_fwd(const __grid_constant__ ViewParams vp,
                                                       const uint32_t* __restrict__ tile_ranges,
                                                       const uint32_t* __restrict__ tile_order,
                                                       const float4* __restrict__ rec, const uint32_t* __restrict__ vals0,
                                                       const uint32_t* __restrict__ vals1,
                                                       const uint32_t* __restrict__ d_result_buf, float* __restrict__ out_color,
                                                       float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                       uint32_t* __restrict__ out_last, uint32_t* __restrict__ work_counter,
                                                       uint32_t total_work, const RasterCkpt ck)
{
    __shared__ __align__(128) float4 s_rec[2][RB_FWD * 3];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_work, s_slot, s_slots[CK_MAX + 1];
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&s_bar[0], RT);
        mbar_init(&s_bar[1], RT);
        mbar_fence_init();
    }
    const uint32_t* __restrict__ vals = (*d_result_buf) ? vals1 : vals0;
    const uint32_t rec_base = smem_u32(&s_rec[0][0]);
    uint32_t seq = 0;   // batches staged so far by this CTA: stage = seq & 1, mbarrier parity = (seq >> 1) & 1

    for (;;) {
        __syncthreads();   // previous block fully consumed (also orders the mbarrier init before first use)
        if (tid == 0) s_work = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t work = s_work;
        if (work >= total_work) break;
        const BlockMap bm = map_block(vp, tile_order, work);
        const int pxi = bm.x0 + (tid & 15), py0 = bm.y0 + (tid >> 4) * FPPT;
        const uint32_t start = tile_ranges[bm.tile * 2], end = tile_ranges[bm.tile * 2 + 1];
        const uint32_t count = end > start ? end - start : 0u;
        const int nb = (int)((count + RB_FWD - 1) / RB_FWD);

        // Gather staging: the tile's list is (tile_ranges, sorted Gaussian indices); every thread pulls RB_FWD / RT
        // 48-byte records of the next batch from the (L2-resident) record table straight into shared memory with
        // 16-byte async copies that complete on the batch's mbarrier.  Indices are prefetched two batches ahead.
        // Slots past the end of the list (up to the next chunk boundary) are filled with a null record
        // (log2(opacity) = -inf => alpha = +0), so the blend loop always runs whole chunks.
        constexpr int SLOTS = RB_FWD / RT;
        uint32_t idx_next[SLOTS];
        auto load_idx = [&](int b) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const uint32_t j = (uint32_t)b * RB_FWD + s * RT + tid;
                idx_next[s] = (b < nb && j < count) ? vals[start + j] : 0xffffffffu;
            }
        };
        auto issue = [&](int b) {
            const uint32_t sq = seq + (uint32_t)b;
            uint64_t* bar = &s_bar[sq & 1];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                float4* dst = &s_rec[sq & 1][(s * RT + tid) * 3];
                if (idx_next[s] != 0xffffffffu) {
                    const float4* src = rec + (size_t)idx_next[s] * 3;
                    cp_async16(dst, src);
                    cp_async16(dst + 1, src + 1);
                    cp_async16(dst + 2, src + 2);
                } else {
                    dst[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    dst[1] = make_float4(0.f, __int_as_float(0xff800000), 0.f, 0.f);
                    dst[2] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            cp_async_mbar_arrive_noinc(bar);   // fires when this thread's copies have landed
        };
        load_idx(0);
        if (nb > 0) issue(0);
        load_idx(1);
        __syncthreads();   // null-record fills of batch 0 are plain shared stores

        const float pxf = (float)pxi, pyf = (float)py0;
        // Rows (2k, 2k+1) of the thread's column form one f32x2 lane pair.  The colour/depth accumulators hold the
        // NEGATED sums and the blend works with -alpha, so that T - T*alpha is a single packed FMA with no operand
        // negation (f32x2 instructions have none); every half is bit-identical to the scalar fmaf sequence.
        constexpr int FPAIRS = FPPT / 2;
        f32x2 T2[FPAIRS], ncx[FPAIRS], ncy[FPAIRS], ncz[FPAIRS], ndep[FPAIRS];
        float Ts[FPPT];
        uint32_t ci[FPPT];
        bool active[FPPT];
#pragma unroll
        for (int r = 0; r < FPPT; ++r) {
            active[r] = pxi < bm.xmax && py0 + r < bm.ymax;
            Ts[r] = 0.f;
            ci[r] = 0u;
        }
#pragma unroll
        for (int k = 0; k < FPAIRS; ++k) {
            // T == 0  <=>  this pixel is finished (terminated or outside)
            T2[k] = f2_make(active[2 * k] ? 1.0f : 0.0f, active[2 * k + 1] ? 1.0f : 0.0f);
            ncx[k] = ncy[k] = ncz[k] = ndep[k] = f2_bc(0.0f);
        }
        bool warp_done = false;
        bool ck_ok = ck.state != nullptr;
        uint32_t ck_written = 0;

        // one Gaussian for the thread's FPPT pixels: slang/gaussian_tile_global_kernels.slang:437-499
        auto blend = [&](uint32_t addr) {
            const float4 a = lds128(addr), q = lds128(addr + 16), c = lds128(addr + 32);
            const FwdExp e = fwd_exponents(a, q.x, q.y, pxf, pyf);
            const f32x2 E0b = f2_bc(e.E0), E1b = f2_bc(e.E1), Cb = f2_bc(e.C);
            const f32x2 colR = f2_bc(q.z), colG = f2_bc(q.w), colB = f2_bc(c.x), colD = f2_bc(c.z);
#pragma unroll
            for (int k = 0; k < FPAIRS; ++k) {
                const f32x2 rf = f2_make((float)(2 * k), (float)(2 * k + 1));
                const f32x2 p = f2_fma(rf, f2_fma(rf, Cb, E1b), E0b);
                const f32x2 na = f2_make(fmaxf(-ex2_approx(f2_lo(p)), -0.99f), fmaxf(-ex2_approx(f2_hi(p)), -0.99f));
                const f32x2 nc = f2_mul(T2[k], na);   // -T * alpha
                ncx[k] = f2_fma(nc, colR, ncx[k]);
                ncy[k] = f2_fma(nc, colG, ncy[k]);
                ncz[k] = f2_fma(nc, colB, ncz[k]);
                if (DEPTH) ndep[k] = f2_fma(nc, colD, ndep[k]);
                // The T < 1e-4 cut (:599-603) is NOT applied per Gaussian: it costs a compare + select per pixel and
                // Gaussian on an issue-bound loop.  It is applied once per chunk of FCHUNK Gaussians (below); a pixel that
                // crossed the threshold inside the chunk has blended up to FCHUNK - 1 Gaussians too many, with weights
                // below 1e-4 - the epilogue replays that one chunk per pixel, finds the exact terminating Gaussian and
                // takes the surplus terms out again (of the pixel and of the checkpoint sums the backward reads).
                T2[k] = f2_fma(T2[k], na, T2[k]);
            }
        };
        auto Trow = [&](int r) { return (r & 1) ? f2_hi(T2[r >> 1]) : f2_lo(T2[r >> 1]); };

        // this thread's FPPT pixels of checkpoint slot `slot`: (sum r, sum g, sum b, T) [+ depth plane]
        auto store_sums = [&](uint32_t slot) {
#pragma unroll
            for (int r = 0; r < FPPT; ++r) {
                const int k = r >> 1;
                const bool hi = (r & 1) != 0;
                const size_t o = (size_t)slot * 256 + (size_t)(((tid >> 4) * FPPT + r) * 16 + (tid & 15));
                ck.state[o] = make_float4(-(hi ? f2_hi(ncx[k]) : f2_lo(ncx[k])), -(hi ? f2_hi(ncy[k]) : f2_lo(ncy[k])),
                                          -(hi ? f2_hi(ncz[k]) : f2_lo(ncz[k])), Trow(r));
                if (DEPTH) ck.depth[o] = -(hi ? f2_hi(ndep[k]) : f2_lo(ndep[k]));
            }
        };

        int b = 0;
        for (; b < nb; ++b) {
            if (b + 1 < nb) issue(b + 1);
            load_idx(b + 2);
            const uint32_t sq = seq + (uint32_t)b;
            mbar_wait(&s_bar[sq & 1], (sq >> 1) & 1u);
            if (!warp_done) {
                const int n = (int)min((uint32_t)RB_FWD, count - (uint32_t)b * RB_FWD);
                uint32_t addr = rec_base + (sq & 1u) * (RB_FWD * 48u);
                uint32_t chunk = (uint32_t)b * (RB_FWD / FCHUNK);
                for (int j = 0; j < n; j += FCHUNK, addr += FCHUNK * 48u, ++chunk) {
                    // remember where (and with which transmittance) each live pixel entered this chunk
#pragma unroll
                    for (int r = 0; r < FPPT; ++r) {
                        const float t = Trow(r);
                        const bool live = t != 0.0f;
                        Ts[r] = live ? t : Ts[r];
                        ci[r] = live ? chunk : ci[r];
                    }
#pragma unroll
                    for (int g = 0; g < FCHUNK; ++g) blend(addr + g * 48u);
                    // the termination cut, once per chunk: a finished pixel continues with T = 0 (adds exact zeros)
                    float tmax = 0.0f;
#pragma unroll
                    for (int k = 0; k < FPAIRS; ++k) {
                        float t0 = f2_lo(T2[k]), t1 = f2_hi(T2[k]);
                        t0 = t0 < 1e-4f ? 0.0f : t0;
                        t1 = t1 < 1e-4f ? 0.0f : t1;
                        T2[k] = f2_make(t0, t1);
                        tmax = fmaxf(tmax, fmaxf(t0, t1));
                    }
                    if (__all_sync(0xffffffffu, tmax == 0.0f)) { warp_done = true; break; }
                }
            }
            // releases the stage buffer for the copy issued two batches later, and votes on early exit
            if (__syncthreads_and(warp_done)) {
                if (b + 1 < nb) {   // drain the in-flight copy
                    mbar_wait(&s_bar[(sq + 1) & 1], ((sq + 1) >> 1) & 1u);
                    ++b;
                }
                ++b;
                break;
            }
            // The block goes on past a multiple of CK Gaussians: checkpoint the blend state of its 256 pixels for the
            // segmented backward (RasterCkpt) - the colour/depth summed SINCE THE PREVIOUS CHECKPOINT and the
            // transmittance - and restart the sums, so a late segment's small terms keep their relative precision
            // (the backward needs sum_{j >= e} w_j c_j; as C_final - C_prefix(e) it would lose it to cancellation).
            // A terminated pixel stores T = 0 and is never read back.
            const uint32_t done = (uint32_t)(b + 1) * RB_FWD;
            if (ck_ok && done % CK == 0u && done < count) {
                if (tid == 0) {
                    // the first checkpoint also reserves the slot of the block's final (partial) segment sums
                    const uint32_t want = ck_written == 0u ? 2u : 1u;
                    uint32_t slot = 0xffffffffu;
                    if (ck_written < (uint32_t)CK_MAX) {
                        slot = atomicAdd(ck.count, want);
                        if (slot + want > ck.capacity) {   // pool exhausted; every slot below min(count, capacity) keeps a valid header
                            if (slot < ck.capacity) ck.header[slot] = make_uint2(work, 0xffffffffu);
                            slot = 0xffffffffu;
                        }
                    }
                    if (slot != 0xffffffffu) {
                        if (want == 2u) {
                            ck.header[slot] = make_uint2(work, 0xffffffffu);   // not a work item
                            s_slots[CK_MAX] = slot++;
                        }
                        ck.header[slot] = make_uint2(work, ck_written);
                        s_slots[ck_written] = slot;
                    }
                    s_slot = slot;
                }
                __syncthreads();
                const uint32_t slot = s_slot;
                if (slot != 0xffffffffu) {
                    store_sums(slot);
#pragma unroll
                    for (int k = 0; k < FPAIRS; ++k) ncx[k] = ncy[k] = ncz[k] = ndep[k] = f2_bc(0.0f);
                    ++ck_written;
                } else {
                    ck_ok = false;   // pool exhausted: the backward handles the rest of this block as one item
                }
            }
        }
        seq += (uint32_t)b;   // batches actually staged (and waited for) in this block
        // Segment sums -> totals.  The block's final (partial) segment sums go to the slot reserved with the first
        // checkpoint; the per-block slot table lets a backward item find the sums of all later segments.
        float tot[FPPT][4];
#pragma unroll
        for (int r = 0; r < FPPT; ++r) tot[r][0] = tot[r][1] = tot[r][2] = tot[r][3] = 0.0f;
        if (ck_written) {   // CTA-uniform
            if (tid <= CK_MAX) ck.table[(size_t)work * (CK_MAX + 1) + tid] = s_slots[tid < (int)ck_written || tid == CK_MAX ? tid : CK_MAX];
            for (uint32_t c = 0; c < ck_written; ++c) {   // this thread re-reads its own stores
                const uint32_t slot = s_slots[c];
#pragma unroll
                for (int r = 0; r < FPPT; ++r) {
                    const size_t o = (size_t)slot * 256 + (size_t)(((tid >> 4) * FPPT + r) * 16 + (tid & 15));
                    const float4 v = ck.state[o];
                    tot[r][0] += v.x; tot[r][1] += v.y; tot[r][2] += v.z;
                    if (DEPTH) tot[r][3] += ck.depth[o];
                }
            }
        }
        if (tid == 0 && ck.written) ck.written[work] = ck_written;
        // epilogue: exact lastContrib / transmittance at termination, then the outputs
#pragma unroll
        for (int r = 0; r < FPPT; ++r) {
            if (!active[r]) continue;
            float Tend = Trow(r);
            uint32_t nContrib = count;
            auto row = [&](const f32x2* v) { return -((r & 1) ? f2_hi(v[r >> 1]) : f2_lo(v[r >> 1])); };
            float cur[4] = {row(ncx), row(ncy), row(ncz), DEPTH ? row(ndep) : 0.0f};   // sums since the last checkpoint
            const size_t po = (size_t)(((tid >> 4) * FPPT + r) * 16 + (tid & 15));      // pixel inside a checkpoint slot
            if (Tend == 0.0f) {
                // The pixel finished inside chunk ci[r], which it entered with transmittance Ts[r]: replay that chunk with
                // the blend loop's own bits, find the terminating Gaussian (included, :599-603) and collect what the loop
                // blended beyond it (the loop only cuts at chunk ends)
                float t = Ts[r];
                uint32_t i = ci[r] * FCHUNK;
                float sur[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                bool found = false;
                for (int g = 0; g < FCHUNK && i < count; ++g, ++i) {
                    const float4* src = rec + (size_t)vals[start + i] * 3;
                    const float4 a = __ldg(src), q = __ldg(src + 1);
                    const FwdExp e = fwd_exponents(a, q.x, q.y, pxf, pyf);
                    const float alpha = fwd_alpha_rt(e, r);
                    if (found) {
                        const float4 c = __ldg(src + 2);
                        const float w = t * alpha;
                        sur[0] = fmaf(w, q.z, sur[0]); sur[1] = fmaf(w, q.w, sur[1]); sur[2] = fmaf(w, c.x, sur[2]);
                        if (DEPTH) sur[3] = fmaf(w, c.z, sur[3]);
                    }
                    const float Tn = fmaf(-t, alpha, t);
                    if (!found && Tn < 1e-4f) {
                        nContrib = i + 1u;
                        Tend = Tn;
                        found = true;
                    }
                    t = Tn;
                }
                const uint32_t seg = ci[r] * FCHUNK / (uint32_t)CK;   // the segment (between checkpoints) the surplus went into
                if (seg < ck_written) {
                    const size_t o = (size_t)s_slots[seg] * 256 + po;
                    float4 v = ck.state[o];
                    v.x -= sur[0]; v.y -= sur[1]; v.z -= sur[2];
                    ck.state[o] = v;
                    if (DEPTH) ck.depth[o] -= sur[3];
#pragma unroll
                    for (int q = 0; q < 4; ++q) tot[r][q] -= sur[q];
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) cur[q] -= sur[q];
                }
            }
            const size_t p = (size_t)(py0 + r) * vp.W + pxi;
            const float bg = vp.whiteBg ? Tend : 0.0f;
            if (ck_written) {
                // final (partial) segment sums + the exact final transmittance (out_alpha = 1 - T loses its low bits)
                const size_t o = (size_t)s_slots[CK_MAX] * 256 + po;
                ck.state[o] = make_float4(cur[0], cur[1], cur[2], Tend);
                if (DEPTH) ck.depth[o] = cur[3];
            }
            out_color[p * 3 + 0] = (tot[r][0] + cur[0]) + bg;
            out_color[p * 3 + 1] = (tot[r][1] + cur[1]) + bg;
            out_color[p * 3 + 2] = (tot[r][2] + cur[2]) + bg;
            if (DEPTH) out_depth[p] = tot[r][3] + cur[3];
            out_alpha[p] = 1.0f - Tend;
            out_last[p] = nContrib;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <bool DEPTH>
__global__ void __launch_bounds__(32, BWD_WARPS) k_raster_bwd(const __grid_constant__ ViewParams vp,
                                                       const uint32_t* __restrict__ tile_ranges,
                                                       const uint32_t* __restrict__ tile_order,
                                                       const float4* __restrict__ rec, const uint32_t* __restrict__ vals0,
                                                       const uint32_t* __restrict__ vals1,
                                                       const uint32_t* __restrict__ d_result_buf,
                                                       const float* __restrict__ cot_color,
                                                       const float* __restrict__ cot_depth, const float* __restrict__ cot_alpha,
                                                       const float* __restrict__ out_color, const float* __restrict__ out_depth,
                                                       const float* __restrict__ out_alpha,
                                                       const uint32_t* __restrict__ last_contrib, float* __restrict__ grad_rec,
                                                       uint32_t* __restrict__ work_counter, uint32_t num_blocks, const RasterCkpt ck)
{
    __shared__ __align__(128) float4 s_rec[2][RB_BWD * 3];
    __shared__ __align__(16) float s_out[RB_BWD][12];   // per-Gaussian sums of this block for one batch
    __shared__ __align__(16) float s_part[BGRP][10][BPAD];   // per-lane partial sums of up to BGRP Gaussians
    __shared__ __align__(8) uint64_t s_bar[2];
    const int lane = threadIdx.x;
    // Cross-lane reduction through shared memory: each lane parks its 10 partial sums per Gaussian (10 conflict-free
    // STS); after BGRP Gaussians lane L sums column L of the BGRP x 10 parked columns (8 LDS.128 + 31 FADD) and writes
    // the total to s_out.  ~24 instructions per Gaussian instead of the 66 of a 13-shuffle select/butterfly.
    const int red_g = lane / 10, red_c = lane - red_g * 10;   // lanes 30, 31 idle
    int gcnt = 0;
    uint32_t jpack = 0u;   // batch slots j (< 256) of the parked Gaussians, 8 bits each
    auto reduce_group = [&](int n) {
        __syncwarp();
        if (red_g < n && (DEPTH || red_c != 3)) {   // without depth column 3 is never parked; s_out[.][3] keeps its zero
            const float4* col = reinterpret_cast<const float4*>(&s_part[red_g][red_c][0]);
            float4 a = col[0];
#pragma unroll
            for (int m = 1; m < 8; ++m) {
                const float4 v = col[m];
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            s_out[(jpack >> (8 * red_g)) & 0xffu][red_c] = (a.x + a.y) + (a.z + a.w);
        }
        __syncwarp();
    };
    if (lane == 0) {
        mbar_init(&s_bar[0], 32);
        mbar_init(&s_bar[1], 32);
        mbar_fence_init();
    }
    __syncwarp();
    const uint32_t* __restrict__ vals = (*d_result_buf) ? vals1 : vals0;
    uint32_t seq = 0;   // batches staged so far by this CTA: stage = seq & 1, mbarrier parity = (seq >> 1) & 1
    // Persistent: one warp pulls work items from a device counter.  A work item is a SEGMENT of a 16x16 block's list:
    // the forward saved the per-pixel blend state every CK Gaussians (RasterCkpt), so the Gaussians
    // [c CK, (c+1) CK) of a block can be differentiated on their own, starting from the checkpointed state at (c+1) CK:
    // the items are bounded by CK Gaussians and the warps of the grid finish together (a whole-block item is up to
    // ~7x the mean, longer than the ideal makespan at 16 warps per SM).  Items: first one per checkpoint slot (all
    // full segments), then one per block for the remainder past the block's last checkpoint, heavy tiles first.
    const uint32_t num_slots = ck.count ? min(*ck.count, ck.capacity) : 0u;
    const uint32_t total_work = num_slots + num_blocks;
    for (;;) {
    __syncwarp();       // every lane is done with the previous block's shared memory
    uint32_t work = 0;
    if (lane == 0) work = atomicAdd(work_counter, 1u);
    work = __shfl_sync(0xffffffffu, work, 0);
    if (work >= total_work) break;
    const bool interior = work < num_slots;
    uint32_t block_id, seg_begin;
    uint32_t ck_index = 0, ck_written = 0;
    if (interior) {
        const uint2 h = ck.header[work];   // (block work id, checkpoint index)
        if (h.y == 0xffffffffu) continue;  // the slot holding a block's final segment sums is not a work item
        block_id = h.x;
        ck_index = h.y;
        ck_written = ck.written[block_id];
        seg_begin = h.y * (uint32_t)CK;
    } else {
        block_id = work - num_slots;
        seg_begin = ck.written ? ck.written[block_id] * (uint32_t)CK : 0u;
    }
    const BlockMap bm = map_block(vp, tile_order, block_id);
    const uint32_t start = tile_ranges[bm.tile * 2], end = tile_ranges[bm.tile * 2 + 1];
    const uint32_t count = end > start ? end - start : 0u;
    const uint32_t seg_limit = interior ? seg_begin + (uint32_t)CK : count;   // Gaussians >= seg_limit belong to later items

    // thread = column (lane & 15) x 8 consecutive rows starting at (lane >> 4) * 8
    const int pxi = bm.x0 + (lane & 15);
    const int py0 = bm.y0 + (lane >> 4) * BPPT;
    // rows (2k, 2k+1) of the thread's column form one f32x2 lane pair (see the forward)
    constexpr int BPAIRS = BPPT / 2;
    f32x2 sT2[BPAIRS], kX2[BPAIRS], kY2[BPAIRS], kZ2[BPAIRS], kD2[BPAIRS], kT2[BPAIRS];
    uint32_t nC[BPPT];
    uint32_t nmax = 0, nmin = 0xffffffffu;
    {
        float sT[BPPT], kX[BPPT], kY[BPPT], kZ[BPPT], kD[BPPT], kT[BPPT];
#pragma unroll
        for (int p = 0; p < BPPT; ++p) {
            const int pyi = py0 + p;
            sT[p] = 0.f; kX[p] = 0.f; kY[p] = 0.f; kZ[p] = 0.f; kD[p] = 0.f; kT[p] = 0.f;
            nC[p] = 0;
            if (pxi < bm.xmax && pyi < bm.ymax) {
                // slang/gaussian_tile_global_kernels.slang:696-723
                const size_t pix = (size_t)pyi * vp.W + pxi;
                kX[p] = cot_color[pix * 3];
                kY[p] = cot_color[pix * 3 + 1];
                kZ[p] = cot_color[pix * 3 + 2];
                kD[p] = cot_depth ? cot_depth[pix] : 0.0f;
                const float cotA = cot_alpha ? cot_alpha[pix] : 0.0f;
                const float Tfin = 1.0f - out_alpha[pix];
                sT[p] = Tfin;
                kT[p] = -cotA + (vp.whiteBg ? (kX[p] + kY[p] + kZ[p]) : 0.0f);
                const uint32_t nc = min(last_contrib[pix], count);
                nC[p] = min(nc, seg_limit);
                if (nc > seg_limit) {
                    // The pixel is still live at the end of this segment: start from the forward's checkpoint there.
                    // The back-to-front recursion kT <- (1 - alpha) kT + alpha (cot . colour) has the closed form
                    //   kT_e = (cot . sum_{j >= e} w_j c_j + T_final kT_init) / T(e),
                    // and the sum is the forward's segment sums of all later segments (+ the final partial one)
                    const int pi = (p + (lane >> 4) * BPPT) * 16 + (lane & 15);
                    const uint32_t* tab = ck.table + (size_t)block_id * (CK_MAX + 1);
                    float acc = 0.0f, Tend = Tfin;
                    for (uint32_t t = ck_index + 1; t <= ck_written; ++t) {
                        const size_t o = (size_t)tab[t < ck_written ? t : (uint32_t)CK_MAX] * 256 + pi;
                        const float4 s4 = ck.state[o];
                        acc = fmaf(kX[p], s4.x, fmaf(kY[p], s4.y, fmaf(kZ[p], s4.z, acc)));
                        if (DEPTH) acc = fmaf(kD[p], ck.depth[o], acc);
                        Tend = s4.w;   // the last slot read holds the exact final transmittance
                    }
                    const float Te = ck.state[(size_t)work * 256 + pi].w;   // T after seg_limit Gaussians
                    kT[p] = fmaf(kT[p], Tend, acc) / Te;
                    // The reference starts its undo chain from T = 1 - out_alpha (tile_global.slang:696-723), whose
                    // rounding (up to 1e-3 relative for a saturated pixel) scales every transmittance of the chain:
                    // apply the same factor to the checkpoint so that segmented and whole-list gradients agree
                    sT[p] = Te * (Tfin / Tend);
                }
            }
            nmax = max(nmax, nC[p]);
            nmin = min(nmin, nC[p]);
        }
#pragma unroll
        for (int k = 0; k < BPAIRS; ++k) {
            sT2[k] = f2_make(sT[2 * k], sT[2 * k + 1]);
            kX2[k] = f2_make(kX[2 * k], kX[2 * k + 1]);
            kY2[k] = f2_make(kY[2 * k], kY[2 * k + 1]);
            kZ2[k] = f2_make(kZ[2 * k], kZ[2 * k + 1]);
            kD2[k] = f2_make(kD[2 * k], kD[2 * k + 1]);
            kT2[k] = f2_make(kT[2 * k], kT[2 * k + 1]);
        }
    }
    // only Gaussians below the block-wide max nContrib (capped at the segment end) can contribute
    const uint32_t used_abs = __reduce_max_sync(0xffffffffu, nmax);
    const uint32_t nmin_block = __reduce_min_sync(0xffffffffu, nmin);
    const uint32_t used = used_abs > seg_begin ? used_abs - seg_begin : 0u;   // Gaussians of this item
    const int nb = (int)((used + RB_BWD - 1) / RB_BWD);
    if (nb == 0) continue;
    const uint32_t lstart = start + seg_begin;   // list position of the item's first Gaussian

#if GSB_BWD_COMPACT
    // ---- Sparse tail.  Back to front an item begins where FEWEST pixels are still active: at C3 11 % of all
    // (block, Gaussian) iterations have at most 32 of the 256 pixels active (6 % at most 16; oracle work model, DESIGN.md
    // section 7), yet the dense body below costs the same whatever the number of active pixels.  Those iterations run
    // ONE PIXEL PER LANE instead: i_split = the smallest Gaussian index from which on at most 32 pixels are active
    // (binary search over the block's 256 nContrib values, warp-wide counts); the <= 32 pixels with nContrib > i_split
    // hand transmittance and kT to one lane each through shared memory (the constant part of a pixel's state - cotangents,
    // nContrib - is re-read from where the dense prologue took it), the Gaussians
    // [i_split, end) are differentiated by the scalar body gaussian_c (a fifth of the dense body's arithmetic, same
    // parked sums -> same cross-lane reduction and flush), then transmittance and kT go back to the owning lanes and the
    // dense bodies take over at i_split.
#ifndef GSB_BWD_CMIN
#define GSB_BWD_CMIN 12
#endif
    constexpr uint32_t CMIN = GSB_BWD_CMIN;   // fewer compact Gaussians than this do not pay for the search and the two hand-overs
    uint32_t i_split = used_abs, cnC = 0u;
    bool compact_on = false;
    float cpx = 0.f, cpy = 0.f, csT = 0.f, ckT = 0.f, ckX = 0.f, ckY = 0.f, ckZ = 0.f, ckD = 0.f;
    // Hand-over area (s_part is idle at both hand-overs): transmittance and kT of all 256 pixels as [column][row] with a
    // row stride of 18 floats - a dense lane moves its four row pairs as 64-bit words (its packed registers are never
    // taken apart: conditional half updates made the compiler keep every pair split across the dense loops, +45 MOV per
    // Gaussian), a compact lane reads / writes the one float of its pixel - and the pixel of each compact lane.
    constexpr int CSTRIDE = 18;
    float* const s_cst = &s_part[0][0][0];
    float* const s_ckt = s_cst + 16 * CSTRIDE;
    uint32_t* const s_cpid = reinterpret_cast<uint32_t*>(s_ckt + 16 * CSTRIDE);
    static_assert(2 * 16 * CSTRIDE + 32 <= BGRP * 10 * BPAD, "hand-over area must fit into s_part");
    const uint32_t cpair_addr = smem_u32(s_cst) + (uint32_t)(((lane & 15) * CSTRIDE + (lane >> 4) * BPPT) * 4);   // this lane's pair 0
    uint32_t cidx = 0u;   // compact lane: its pixel inside the hand-over arrays
    {
        auto count_gt = [&](uint32_t t) {
            uint32_t cnt = 0;
#pragma unroll
            for (int p = 0; p < BPPT; ++p) cnt += nC[p] > t ? 1u : 0u;
            return __reduce_add_sync(0xffffffffu, cnt);
        };
        const uint32_t lo_bound = max(nmin_block, seg_begin);   // below it every pixel of the block is active
        if (used_abs > lo_bound + CMIN && count_gt(used_abs - CMIN) <= 32u) {
            uint32_t lo = lo_bound, hi = used_abs - CMIN;        // invariant: count_gt(hi) <= 32
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (count_gt(mid) <= 32u) hi = mid;
                else lo = mid + 1u;
            }
            i_split = hi;
            uint32_t sel = 0u;
#pragma unroll
            for (int p = 0; p < BPPT; ++p) sel |= (nC[p] > i_split ? 1u : 0u) << p;
            const uint32_t mine = (uint32_t)__popc(sel);
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t ctotal = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t r = incl - mine;
#pragma unroll
            for (int p = 0; p < BPPT; ++p)
                if ((sel >> p) & 1u) s_cpid[r++] = (uint32_t)(((lane >> 4) * BPPT + p) * 16 + (lane & 15));
#pragma unroll
            for (int k = 0; k < BPAIRS; ++k) {
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(cpair_addr + 8u * k), "l"(sT2[k].v) : "memory");
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(cpair_addr + 8u * k + 16u * CSTRIDE * 4u), "l"(kT2[k].v) : "memory");
            }
            __syncwarp();
            if ((uint32_t)lane < ctotal) {
                // the constant part of the pixel state comes from where the dense prologue took it (L2)
                const uint32_t pid = s_cpid[lane];
                const int col = (int)(pid & 15u), row = (int)(pid >> 4);
                cidx = (uint32_t)(col * CSTRIDE + row);
                cpx = (float)(bm.x0 + col);
                cpy = (float)(bm.y0 + row);
                const size_t pix = (size_t)(bm.y0 + row) * vp.W + (bm.x0 + col);
                ckX = cot_color[pix * 3]; ckY = cot_color[pix * 3 + 1]; ckZ = cot_color[pix * 3 + 2];
                ckD = cot_depth ? cot_depth[pix] : 0.0f;
                cnC = min(min(last_contrib[pix], count), seg_limit);
                csT = s_cst[cidx];
                ckT = s_ckt[cidx];
            }
            __syncwarp();   // s_part is about to park sums again
            compact_on = true;
        }
    }
#endif

    // batches are visited last -> first; sequence number s = nb-1-b selects stage / parity
    // gather staging as in the forward: each lane pulls two 48-byte records per batch
    uint32_t ia = 0xffffffffu, ib = 0xffffffffu;   // indices of this lane's two slots of the next batch to issue
    auto load_idx = [&](int b) {
        const uint32_t j0 = (uint32_t)b * RB_BWD + lane, j1 = j0 + 32;
        ia = (b >= 0 && j0 < used) ? vals[lstart + j0] : 0xffffffffu;
        ib = (RB_BWD > 32 && b >= 0 && j1 < used) ? vals[lstart + j1] : 0xffffffffu;
    };
    auto issue = [&](int b) {
        const uint32_t s = seq + (uint32_t)(nb - 1 - b);
        uint64_t* bar = &s_bar[s & 1];
        if (ia != 0xffffffffu) {
            float4* dst = &s_rec[s & 1][lane * 3];
            const float4* src = rec + (size_t)ia * 3;
            cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
        }
        if (ib != 0xffffffffu) {
            float4* dst = &s_rec[s & 1][(lane + 32) * 3];
            const float4* src = rec + (size_t)ib * 3;
            cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
        }
        cp_async_mbar_arrive_noinc(bar);   // fires when this lane's copies have landed (immediately if it has none)
    };
    load_idx(nb - 1);
    issue(nb - 1);
    load_idx(nb - 2);

    const float pxf = (float)pxi, pyf = (float)py0;
    const uint32_t rec_base = smem_u32(&s_rec[0][0]);

    // per batch: begin (refill the other stage, clear the batch sums, wait for this batch's records) - the Gaussians,
    // back to front - end (cross-lane sums of the last group, flush).  The state of the open batch:
    int n = 0;                     // Gaussians in it
    uint32_t stage_addr = 0u;      // its records in shared memory
    uint32_t bstart = 0u;          // absolute index of its first Gaussian
    auto begin_batch = [&](int b) {
        const uint32_t s = seq + (uint32_t)(nb - 1 - b);
        __syncwarp();                               // every lane is done with the stage being refilled
        if (b > 0) issue(b - 1);
        load_idx(b - 2);
        {
            float4* z = reinterpret_cast<float4*>(&s_out[0][0]);
            for (int i = lane; i < RB_BWD * 3; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&s_bar[s & 1], (s >> 1) & 1u);
        __syncwarp();
        n = (int)min((uint32_t)RB_BWD, used - (uint32_t)b * RB_BWD);
        stage_addr = rec_base + (s & 1u) * (RB_BWD * 48u);
        bstart = seg_begin + (uint32_t)b * RB_BWD;
    };
    auto end_batch = [&]() {
        if (gcnt) {
            reduce_group(gcnt);
            gcnt = 0;
            jpack = 0u;
        }
        __syncwarp();
        // flush: finish the per-Gaussian chain rule on the block sums and send one 16-byte vector
        // reduction per quad to L2 (records that no pixel of this block reached are skipped)
        for (int j = lane; j < n; j += 32) {
            const float4 s0 = *reinterpret_cast<const float4*>(&s_out[j][0]);   // Cr Cg Cb Cd
            const float4 s1 = *reinterpret_cast<const float4*>(&s_out[j][4]);   // S0 Sx Sy Sxx
            const float4 s2 = *reinterpret_cast<const float4*>(&s_out[j][8]);   // Sxy Syy - -
            const bool any = s0.x != 0.f || s0.y != 0.f || s0.z != 0.f || s0.w != 0.f || s1.x != 0.f || s1.y != 0.f ||
                             s1.z != 0.f || s1.w != 0.f || s2.x != 0.f || s2.y != 0.f;
            if (!any) continue;
            const uint32_t addr = stage_addr + (uint32_t)j * 48u;
            const float4 a = lds128(addr), q = lds128(addr + 16), c = lds128(addr + 32);
            // h carries the opacity factor (alpha-weighted): the opacity gradient is H0 / opacity.  opacity == 0
            // (sigmoid underflow) gives 0; the activation VJP multiplies it by opacity (1 - opacity) = 0 anyway
            const float g_op = c.y > 0.0f ? s1.x / c.y : 0.0f;
            const float kc = -0.5f;                      // d(natural exponent)/d conic = -0.5 * (dx^2, dx dy, dy^2)
            const float km = -(1.0f / LOG2E_F);          // d(natural exponent)/d mean = -(2a dx + b dy, ...), (a,b,c) = (A,B,C)/log2 e
            const float g_mx = km * fmaf(a.z + a.z, s1.y, a.w * s1.z);
            const float g_my = km * fmaf(q.x + q.x, s1.z, a.w * s1.y);
            const float g_c00 = kc * s1.w, g_c01 = kc * s2.x, g_c11 = kc * s2.y;
            float* dst = grad_rec + (size_t)(__float_as_uint(c.w) & REC_IDX_MASK) * REC_FLOATS;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g_mx), "f"(g_my), "f"(g_c00), "f"(g_c01) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(g_c01), "f"(g_c11), "f"(s0.x), "f"(s0.y) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 8), "f"(s0.z), "f"(g_op), "f"(s0.w), "f"(0.0f) : "memory");
        }
    };
    int b_dense = nb - 1;          // first (highest) batch of the dense phase
    int j_open = -2;               // >= -1: batch b_dense is already open, the dense bodies continue below Gaussian j_open + 1
#if GSB_BWD_COMPACT
    if (compact_on) {
        // One Gaussian for this lane's ONE pixel: the scalar restatement of `gaussian` (masked), parking the same ten sums
        auto gaussian_c = [&](int jj, const float4& a, const float4& q, const float4& c, auto clamp_c) {
            constexpr bool CLAMP = decltype(clamp_c)::value;
            const uint32_t i = bstart + (uint32_t)jj;
            const float dx = cpx - a.x, dy = cpy - a.y;
            const float e = fmaf(dx, fmaf(a.z, dx, a.w * dy), fmaf(q.x * dy, dy, q.y));
            float al = ex2_approx(e);
            al = i < cnC ? al : 0.0f;
            const float raw = al;
            if (CLAMP) al = fminf(al, 0.99f);
            const float prevT = csT * rcp_approx(1.0f - al);
            const float contrib = prevT * al;
            csT = prevT;
            float dotc = fmaf(ckZ, c.x, fmaf(ckY, q.w, ckX * q.z));
            if (DEPTH) dotc = fmaf(ckD, c.z, dotc);
            const float d = dotc - ckT;
            ckT = fmaf(al, d, ckT);
            float hh = contrib * d;
            if (CLAMP) hh = raw > 0.99f ? 0.0f : hh;
            const float Sx = hh * dx, Sy = hh * dy;
            float* dst = &s_part[gcnt][0][lane];
            dst[0 * BPAD] = contrib * ckX; dst[1 * BPAD] = contrib * ckY; dst[2 * BPAD] = contrib * ckZ;
            if (DEPTH) dst[3 * BPAD] = contrib * ckD;
            dst[4 * BPAD] = hh;
            dst[5 * BPAD] = Sx; dst[6 * BPAD] = Sy; dst[7 * BPAD] = dx * Sx; dst[8 * BPAD] = dx * Sy; dst[9 * BPAD] = dy * Sy;
            jpack |= (uint32_t)jj << (8 * gcnt);
            if (++gcnt == BGRP) {
                reduce_group(BGRP);
                gcnt = 0;
                jpack = 0u;
            }
        };
        int b = nb - 1, j = 0;
        for (;; --b) {
            begin_batch(b);
            const int jc = (int)min((uint32_t)n, max(i_split, bstart) - bstart);
            for (j = n - 1; j >= jc; --j) {
                const uint32_t addr = stage_addr + (uint32_t)j * 48u;
                const float4 a = lds128(addr), q = lds128(addr + 16), c = lds128(addr + 32);
                if (__float_as_uint(c.w) & REC_MAYCLAMP) gaussian_c(j, a, q, c, std::true_type{});   // warp-uniform
                else gaussian_c(j, a, q, c, std::false_type{});
            }
            if (i_split >= bstart) break;   // the compact phase ends in this batch (i_split >= seg_begin: at the latest in batch 0)
            end_batch();
        }
        // transmittance and kT return to the owning lanes; the dense bodies continue in the open batch below Gaussian j + 1
        if (gcnt) {
            reduce_group(gcnt);
            gcnt = 0;
            jpack = 0u;
        }
        __syncwarp();
        b_dense = b;
        j_open = j;
    }
    // The dense loops always start from packed state that came back through the hand-over area - whether or not a compact
    // phase ran - so that they see ONE definition of it (a second, conditional definition made the compiler keep the pairs
    // split: 12 MOV per Gaussian).  The parked sums of the compact phase have overwritten the area: every lane lays its
    // pairs out again, then the compact lanes put the new values of their pixels on top.
#pragma unroll
    for (int k = 0; k < BPAIRS; ++k) {
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(cpair_addr + 8u * k), "l"(sT2[k].v) : "memory");
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(cpair_addr + 8u * k + 16u * CSTRIDE * 4u), "l"(kT2[k].v) : "memory");
    }
    __syncwarp();
    if (compact_on && cnC) {   // a lane without a pixel has cnC == 0 (a selected pixel has nContrib > i_split >= 0)
        s_cst[cidx] = csT;
        s_ckt[cidx] = ckT;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < BPAIRS; ++k) {
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(sT2[k].v) : "r"(cpair_addr + 8u * k) : "memory");
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(kT2[k].v) : "r"(cpair_addr + 8u * k + 16u * CSTRIDE * 4u) : "memory");
    }
    __syncwarp();
#endif
    for (int b = b_dense; b >= 0; --b) {
        int j;
        if (j_open >= -1) {
            j = j_open;
            j_open = -2;
        } else {
            begin_batch(b);
            j = n - 1;
        }
        // One Gaussian for the thread's 8 pixels.  MASKED = false when every pixel of the block is known to be active
        // (i < block-wide min nContrib); CLAMP = false when the record says alpha cannot reach 0.99 (REC_MAYCLAMP clear).
        // The four variants are whole loop bodies (no join inside an iteration), so the pixel state is updated in
        // place, and no variant branches, so the compiler interleaves the 4 dependent pair chains.  An inactive pixel
        // (i >= nContrib) of the masked variant sees alpha = 0: contrib = 0, kT unchanged, h = 0, transmittance unchanged.
        auto gaussian = [&](int j, const float4& a, const float4& q, const float4& c, auto masked_c, auto clamp_c) {
            constexpr bool MASKED = decltype(masked_c)::value, CLAMP = decltype(clamp_c)::value;
            const uint32_t i = bstart + (uint32_t)j;
            const float dx = pxf - a.x, dyb = pyf - a.y;
            // The thread's 8 pixels share dx and have dy = dyb + p, so the exponent is a quadratic in the
            // compile-time row offset p:  e(p) = E0 + p*E1 + p^2*C  (log2 units).  log2(opacity) is folded in, so
            // ex2 yields alpha itself and h = dL/d(ln of the opacity-free alpha) = contrib * d; the opacity gradient
            // is H0 / opacity (flush).
            const float E0 = fmaf(dx, fmaf(a.z, dx, a.w * dyb), fmaf(q.x * dyb, dyb, q.y));
            const float E1 = fmaf(a.w, dx, 2.0f * q.x * dyb);
            // Rows (2k, 2k + 1) as one pair: e(2k + b) = [E0 + b E1 + b^2 C] + 2k [E1 + 2b C] + 4k^2 C = P0 + 2k P1 + 4k^2 C, so
            // the multipliers of a pair are the SAME in both halves (broadcast immediates): pair 0 is free and the others
            // take two FFMA2 - the row offsets (2k, 2k + 1) as a packed constant cost two UMOV per pair on top of them
            const f32x2 Cqb = f2_bc(q.x);
            const f32x2 P0 = f2_make(E0, (E0 + E1) + q.x), P1 = f2_make(E1, fmaf(2.0f, q.x, E1));
            const f32x2 colR = f2_bc(q.z), colG = f2_bc(q.w), colB = f2_bc(c.x), colD = f2_bc(c.z);
            f32x2 Cr2 = f2_bc(0.f), Cg2 = f2_bc(0.f), Cb2 = f2_bc(0.f), Cd2 = f2_bc(0.f);
            f32x2 h[BPAIRS];
#pragma unroll
            for (int k = 0; k < BPAIRS; ++k) {
                const f32x2 e = k == 0 ? P0 : f2_fma(P1, f2_bc((float)(2 * k)), f2_fma(Cqb, f2_bc((float)(4 * k * k)), P0));    // :437-483
                float a0 = ex2_approx(f2_lo(e)), a1 = ex2_approx(f2_hi(e));
                const bool act0 = !MASKED || (i < nC[2 * k]), act1 = !MASKED || (i < nC[2 * k + 1]);
                if (MASKED) {
                    a0 = act0 ? a0 : 0.0f;
                    a1 = act1 ? a1 : 0.0f;
                }
                const float raw0 = a0, raw1 = a1;
                if (CLAMP) {
                    a0 = fminf(a0, 0.99f);
                    a1 = fminf(a1, 0.99f);
                }
                const f32x2 al = f2_make(a0, a1);
                // undoTileGlobalPixelState (:501-521): only the transmittance matters for the gradients;
                // max(1 - alpha, 1e-6) == 1 - alpha because alpha <= 0.99
                const f32x2 om = f2_fma(al, f2_bc(-1.0f), f2_bc(1.0f));
                const f32x2 prevT = f2_mul(sT2[k], f2_make(rcp_approx(f2_lo(om)), rcp_approx(f2_hi(om))));
                const f32x2 contrib = f2_mul(prevT, al);
                // an inactive pixel has alpha = 0: 1 - alpha = 1, rcp.approx(1) = 1 exactly, prevT = sT - no select needed
                sT2[k] = prevT;
                // VJP of updateTileGlobalPixelState (:485-499)
                f32x2 dotc = f2_fma(kZ2[k], colB, f2_fma(kY2[k], colG, f2_mul(kX2[k], colR)));
                if (DEPTH) dotc = f2_fma(kD2[k], colD, dotc);
                const f32x2 d = f2_sub(dotc, kT2[k]);
                kT2[k] = f2_fma(al, d, kT2[k]);
                Cr2 = f2_fma(contrib, kX2[k], Cr2);
                Cg2 = f2_fma(contrib, kY2[k], Cg2);
                Cb2 = f2_fma(contrib, kZ2[k], Cb2);
                if (DEPTH) Cd2 = f2_fma(contrib, kD2[k], Cd2);
                // VJP of evaluateTileGlobalSample: everything geometric is a moment of h; the alpha clamp branch has
                // zero gradient
                f32x2 hh = f2_mul(contrib, d);
                if (CLAMP) hh = f2_make(raw0 > 0.99f ? 0.0f : f2_lo(hh), raw1 > 0.99f ? 0.0f : f2_hi(hh));
                h[k] = hh;
            }
            // per-thread sums over the 8 pixels: colour/depth terms and the moments of h in the row offset
            // p = 2k + (0 | 1):  H0 = sum h,  H1 = sum p h = 2 sum k (he + ho) + sum ho,
            //                    H2 = sum p^2 h = 4 sum k^2 (he + ho) + 4 sum k ho + sum ho
            static_assert(BPAIRS == 4, "moment recombination below is written for 4 pairs");
            const f32x2 S = f2_add(f2_add(h[0], h[1]), f2_add(h[2], h[3]));
            const f32x2 M1 = f2_fma(h[3], f2_bc(3.0f), f2_fma(h[2], f2_bc(2.0f), h[1]));
            const f32x2 M2 = f2_fma(h[3], f2_bc(9.0f), f2_fma(h[2], f2_bc(4.0f), h[1]));
            const float So = f2_hi(S);
            const float H0 = f2_lo(S) + So;
            const float H1 = fmaf(2.0f, f2_lo(M1) + f2_hi(M1), So);
            const float H2 = fmaf(4.0f, (f2_lo(M2) + f2_hi(M2)) + f2_hi(M1), So);
            // moments in (dx, dy) of this thread's pixels: sum h dy = dyb H0 + H1, sum h dy^2 = dyb^2 H0 + 2 dyb H1 + H2
            const float Sy = fmaf(dyb, H0, H1);
            const float Syy = fmaf(dyb, fmaf(dyb, H0, H1 + H1), H2);
            const float Sx = dx * H0;
            // park this lane's 10 partial sums; every BGRP Gaussians the warp sums the parked columns (reduce_group)
            float* dst = &s_part[gcnt][0][lane];
            dst[0 * BPAD] = f2_lo(Cr2) + f2_hi(Cr2); dst[1 * BPAD] = f2_lo(Cg2) + f2_hi(Cg2); dst[2 * BPAD] = f2_lo(Cb2) + f2_hi(Cb2);
            if (DEPTH) dst[3 * BPAD] = f2_lo(Cd2) + f2_hi(Cd2);
            dst[4 * BPAD] = H0;
            dst[5 * BPAD] = Sx; dst[6 * BPAD] = Sy; dst[7 * BPAD] = dx * Sx; dst[8 * BPAD] = dx * Sy; dst[9 * BPAD] = Syy;
            jpack |= (uint32_t)j << (8 * gcnt);
            if (++gcnt == BGRP) {
                reduce_group(BGRP);
                gcnt = 0;
                jpack = 0u;
            }
        };
        // back to front: the Gaussians at or beyond the block-wide min nContrib come first (masked), the rest have
        // every pixel of the block active
        const int jsplit = (int)min((uint32_t)n, max(nmin_block, bstart) - bstart);
        for (; j >= jsplit; --j) {
            const uint32_t addr = stage_addr + (uint32_t)j * 48u;
            const float4 a = lds128(addr), q = lds128(addr + 16), c = lds128(addr + 32);
            if (__float_as_uint(c.w) & REC_MAYCLAMP) gaussian(j, a, q, c, std::true_type{}, std::true_type{});   // warp-uniform
            else gaussian(j, a, q, c, std::true_type{}, std::false_type{});
        }
        for (; j >= 0; --j) {
            const uint32_t addr = stage_addr + (uint32_t)j * 48u;
            const float4 a = lds128(addr), q = lds128(addr + 16), c = lds128(addr + 32);
            if (__float_as_uint(c.w) & REC_MAYCLAMP) gaussian(j, a, q, c, std::false_type{}, std::true_type{});
            else gaussian(j, a, q, c, std::false_type{}, std::false_type{});
        }
        end_batch();
    }
    seq += (uint32_t)nb;
    }   // persistent loop
}

// sum of lastContrib over the image = number of (pixel, Gaussian) blend evaluations (bench statistics)
__global__ void __launch_bounds__(256) k_sum_u32(size_t n, const uint32_t* __restrict__ v, unsigned long long* __restrict__ out)
{
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += v[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

cudaError_t launch_sum_u32(cudaStream_t st, size_t n, const uint32_t* v, unsigned long long* out)
{
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    if (n > 0) k_sum_u32<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(n, v, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static int raster_blocks(const ViewParams& vp)
{
    const int subX = (vp.tileW + 15) / 16, subY = (vp.tileH + 15) / 16;
    return vp.gridW * vp.gridH * subX * subY;
}
int raster_block_count(const ViewParams& vp) { return raster_blocks(vp); }

// resident CTAs per SM of the persistent rasterisers (GSB_FWD_RES / GSB_BWD_RES override, for tuning)
static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
static int sm_count()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

cudaError_t launch_raster_fwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges,
                              const uint32_t* tile_order, const float* rec, const uint32_t* vals0, const uint32_t* vals1,
                              const uint32_t* d_result_buf,
                              float* out_color, float* out_depth, float* out_alpha, uint32_t* out_last, uint32_t* work_counter,
                              const RasterCkpt& ck)
{
    const int blocks = raster_blocks(vp);
    if (blocks > 0) {
        static int res = 0;
        if (!res) {
            cudaFuncSetAttribute(k_raster_fwd<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_raster_fwd<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            res = std::max(1, env_int("GSB_FWD_RES", GSB_FWD_CTAS));   // resident 64-thread CTAs per SM the forward is compiled for
        }
        cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        if (ck.state) {
            e = cudaMemsetAsync(ck.count, 0, sizeof(uint32_t), st);
            if (e != cudaSuccess) return e;
        }
        const int grid = std::min(blocks, sm_count() * res);
        if (out_depth)
            k_raster_fwd<true><<<grid, RT, 0, st>>>(vp, tile_ranges, tile_order, reinterpret_cast<const float4*>(rec), vals0, vals1,
                                                    d_result_buf, out_color, out_depth, out_alpha, out_last, work_counter, (uint32_t)blocks, ck);
        else
            k_raster_fwd<false><<<grid, RT, 0, st>>>(vp, tile_ranges, tile_order, reinterpret_cast<const float4*>(rec), vals0, vals1,
                                                     d_result_buf, out_color, out_depth, out_alpha, out_last, work_counter, (uint32_t)blocks, ck);
    }
    return cudaGetLastError();
}

cudaError_t launch_raster_bwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges,
                              const uint32_t* tile_order, const float* rec, const uint32_t* vals0, const uint32_t* vals1,
                              const uint32_t* d_result_buf,
                              const float* cot_color, const float* cot_depth, const float* cot_alpha,
                              const float* out_color, const float* out_depth, const float* out_alpha,
                              const uint32_t* last_contrib, float* grad_rec, uint32_t* work_counter, const RasterCkpt& ck)
{
    // out_color / out_depth are only read where a work item starts from a forward checkpoint (see kernel)
    const int blocks = raster_blocks(vp);
    if (blocks > 0) {
        static int res = 0;
        if (!res) {
            cudaFuncSetAttribute(k_raster_bwd<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_raster_bwd<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            res = std::max(1, env_int("GSB_BWD_RES", BWD_WARPS));
        }
        cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        const int grid = sm_count() * res;   // the item count (checkpoint slots + blocks) is only known on the device
        if (cot_depth)
            k_raster_bwd<true><<<grid, 32, 0, st>>>(vp, tile_ranges, tile_order, reinterpret_cast<const float4*>(rec), vals0, vals1,
                                                    d_result_buf, cot_color, cot_depth, cot_alpha, out_color, out_depth, out_alpha,
                                                    last_contrib, grad_rec, work_counter, (uint32_t)blocks, ck);
        else
            k_raster_bwd<false><<<grid, 32, 0, st>>>(vp, tile_ranges, tile_order, reinterpret_cast<const float4*>(rec), vals0, vals1,
                                                     d_result_buf, cot_color, cot_depth, cot_alpha, out_color, out_depth, out_alpha,
                                                     last_contrib, grad_rec, work_counter, (uint32_t)blocks, ck);
    }
    return cudaGetLastError();
}

}  // namespace gsb
