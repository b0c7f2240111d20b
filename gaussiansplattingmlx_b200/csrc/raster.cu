// raster.cu — tile alpha-blend forward (K9) and back-to-front backward (K10) for sm_100a.
//
// Reference: slang/gaussian_tile_global_kernels.slang:437-614 (forward), :501-521 + :648-881
// (backward), call sites Trainer/GaussianRenderer.swift:124-147,187-226.
//
// Both kernels are FP32-issue bound (ncu: issue slots ~90 % busy), so the design goal is the fewest
// instructions per (pixel, Gaussian) evaluation:
//   * the tile's depth-ordered 48-byte records are contiguous in the `staged` stream (binning.cu);
//     batches are pulled into shared memory by TMA 1-D bulk copies (cp.async.bulk -> SASS UBLKCP)
//     signalled through mbarriers, double-buffered against the blend loop;
//   * records carry the conic/opacity pre-folded into log2 units (common.cuh), so
//     alpha = min(0.99, ex2(A dx^2 + B dx dy + C dy^2 + lo)): 5 FMA-pipe ops + 1 MUFU;
//   * forward: one pixel per thread, 256 threads = one 16x16 block; a terminated pixel keeps running
//     with T = 0 (adds exact zeros) so the loop body has no per-lane predicate, and the warp votes
//     only every 4 Gaussians;
//   * backward: ONE WARP per 16x16 block, 8 pixels per thread.  The 11 per-Gaussian gradient sums are
//     first accumulated over the thread's 8 pixels in registers (the accumulation is the FMA that
//     produces the term), so the 13-shuffle warp butterfly is paid once per 256 evaluations instead
//     of once per 32, and there is no cross-warp reduction at all.  One vector red per (block,
//     Gaussian, quad) then goes to L2.
// Work per evaluation: forward 27 flop + 1 ex2; backward ~80 flop + 1 ex2 + 1 rcp.
#include <algorithm>

#include "kernels.h"

namespace gsb {

constexpr int RT = 256;          // forward: threads per CTA = 16x16 pixels
constexpr int RB_FWD = 128;      // records per forward batch (6 KB)
constexpr int RB_BWD = 64;       // records per backward batch
constexpr int BPPT = 8;          // backward: pixels per thread (rows)
// Gather staging engine for the 48-byte records: 1 = one TMA bulk copy per record (cp.async.bulk, UBLKCP),
// 0 = three 16-byte cp.async (LDGSTS) per record.  Both complete on the batch's mbarrier.
#ifndef GSB_GATHER_TMA
#define GSB_GATHER_TMA 0
#endif

// (tile, 16x16 sub-block) of this CTA; tiles larger than 16x16 are covered by several CTAs
struct BlockMap {
    int tile, x0, y0;   // pixel origin of the 16x16 block
    int xmax, ymax;     // exclusive pixel bounds of the tile clipped to the image
};
__device__ __forceinline__ BlockMap map_block(const ViewParams& vp, const uint32_t* __restrict__ tile_order)
{
    const int subX = (vp.tileW + 15) >> 4, subY = (vp.tileH + 15) >> 4;
    const int per = subX * subY;
    BlockMap m;
    const int slot = blockIdx.x / per;
    const int sb = blockIdx.x - slot * per;
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
__global__ void __launch_bounds__(RT) k_raster_fwd(const __grid_constant__ ViewParams vp,
                                                   const uint32_t* __restrict__ tile_ranges,
                                                   const uint32_t* __restrict__ tile_order,
                                                   const float4* __restrict__ rec, const uint32_t* __restrict__ vals0,
                                                   const uint32_t* __restrict__ vals1,
                                                   const uint32_t* __restrict__ d_result_buf, float* __restrict__ out_color,
                                                   float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                   uint32_t* __restrict__ out_last)
{
    __shared__ __align__(128) float4 s_rec[2][RB_FWD * 3];
    __shared__ __align__(8) uint64_t s_bar[2];
    const BlockMap bm = map_block(vp, tile_order);
    const int pxi = bm.x0 + (threadIdx.x & 15), pyi = bm.y0 + (threadIdx.x >> 4);
    const bool active = pxi < bm.xmax && pyi < bm.ymax;
    const uint32_t start = tile_ranges[bm.tile * 2], end = tile_ranges[bm.tile * 2 + 1];
    const uint32_t count = end > start ? end - start : 0u;
    const int nb = (int)((count + RB_FWD - 1) / RB_FWD);

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], GSB_GATHER_TMA ? 1 : RB_FWD);
        mbar_init(&s_bar[1], GSB_GATHER_TMA ? 1 : RB_FWD);
        mbar_fence_init();
    }
    __syncthreads();
    // Gather staging: the tile's list is (tile_ranges, sorted Gaussian indices); threads 0..RB_FWD-1 each pull
    // ONE 48-byte record of the next batch from the (L2-resident) record table straight into shared memory with
    // a TMA bulk copy; all copies of a batch complete on one mbarrier.  Indices are prefetched two batches ahead.
    const uint32_t* __restrict__ vals = (*d_result_buf) ? vals1 : vals0;
    auto load_idx = [&](int b) -> uint32_t {
        const uint32_t j = (uint32_t)b * RB_FWD + threadIdx.x;
        return (threadIdx.x < RB_FWD && b < nb && j < count) ? vals[start + j] : 0xffffffffu;
    };
    auto issue = [&](int b, uint32_t idx) {
        uint64_t* bar = &s_bar[b & 1];
#if GSB_GATHER_TMA
        if (threadIdx.x == 0) mbar_expect_tx(bar, min((uint32_t)RB_FWD, count - (uint32_t)b * RB_FWD) * 48u);
        if (idx != 0xffffffffu) bulk_g2s(&s_rec[b & 1][threadIdx.x * 3], rec + (size_t)idx * 3, 48u, bar);
#else
        if (threadIdx.x < RB_FWD) {
            if (idx != 0xffffffffu) {
                float4* dst = &s_rec[b & 1][threadIdx.x * 3];
                const float4* src = rec + (size_t)idx * 3;
                cp_async16(dst, src);
                cp_async16(dst + 1, src + 1);
                cp_async16(dst + 2, src + 2);
                cp_async_mbar_arrive_noinc(bar);
            } else {
                mbar_arrive(bar);
            }
        }
#endif
    };
    uint32_t idx_next = load_idx(0);
    if (nb > 0) issue(0, idx_next);
    idx_next = load_idx(1);

    const float px = (float)pxi, py = (float)pyi;
    float cx = 0.f, cy = 0.f, cz = 0.f, dep = 0.f;
    float T = active ? 1.0f : 0.0f;   // T == 0  <=>  this pixel is finished (terminated or outside)
    float Tfin = -1.0f;               // transmittance at termination (valid when >= 0)
    uint32_t nContrib = count;
    const uint32_t rec_base = smem_u32(&s_rec[0][0]);

    // one Gaussian: slang/gaussian_tile_global_kernels.slang:437-499 + the termination test :599-603
    auto blend = [&](uint32_t addr, uint32_t index) {
        const float4 a = lds128(addr), q = lds128(addr + 16), c = lds128(addr + 32);
        const float dx = px - a.x, dy = py - a.y;
        const float t = fmaf(a.w, dy, a.z * dx);
        const float u = q.x * dy;
        const float p = fmaf(u, dy, fmaf(dx, t, q.y));
        const float alpha = fminf(ex2_approx(p), 0.99f);
        const float contrib = T * alpha;
        cx = fmaf(contrib, q.z, cx);
        cy = fmaf(contrib, q.w, cy);
        cz = fmaf(contrib, c.x, cz);
        dep = fmaf(contrib, c.z, dep);
        float Tn = fmaf(-T, alpha, T);
        if (Tn < 1e-4f && T != 0.0f) {   // terminating Gaussian is included (reference quirk)
            nContrib = index + 1u;
            Tfin = Tn;
            Tn = 0.0f;
        }
        T = Tn;
    };

    for (int b = 0; b < nb; ++b) {
        if (b + 1 < nb) issue(b + 1, idx_next);
        idx_next = load_idx(b + 2);
        mbar_wait(&s_bar[b & 1], (uint32_t)(b >> 1) & 1u);
        const int n = (int)min((uint32_t)RB_FWD, count - (uint32_t)b * RB_FWD);
        if (!__all_sync(0xffffffffu, T == 0.0f)) {
            uint32_t addr = rec_base + (uint32_t)(b & 1) * (RB_FWD * 48u);
            const uint32_t i0 = (uint32_t)b * RB_FWD;
            int j = 0;
            for (; j + 4 <= n; j += 4, addr += 4 * 48u) {
                blend(addr, i0 + j);
                blend(addr + 48u, i0 + j + 1);
                blend(addr + 96u, i0 + j + 2);
                blend(addr + 144u, i0 + j + 3);
                if (__all_sync(0xffffffffu, T == 0.0f)) { j = n; break; }
            }
            for (; j < n; ++j, addr += 48u) blend(addr, i0 + j);
        }
        // releases the stage buffer for the copy issued two batches later, and votes on early exit
        if (__syncthreads_and(T == 0.0f)) {
            if (b + 1 < nb) mbar_wait(&s_bar[(b + 1) & 1], (uint32_t)((b + 1) >> 1) & 1u);  // drain the in-flight copy
            break;
        }
    }
    if (active) {
        const float Tend = Tfin >= 0.0f ? Tfin : T;
        const size_t p = (size_t)pyi * vp.W + pxi;
        const float bg = vp.whiteBg ? Tend : 0.0f;
        out_color[p * 3 + 0] = cx + bg;
        out_color[p * 3 + 1] = cy + bg;
        out_color[p * 3 + 2] = cz + bg;
        out_depth[p] = dep;
        out_alpha[p] = 1.0f - Tend;
        out_last[p] = nContrib;
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Sums 12 per-lane values across the warp with a halving butterfly (13 shuffles instead of 60):
// afterwards lane L holds the warp total of component comp(L) = 6*b4 + 3*b3 + (b2 ? 2 : b1)
// (invalid when b2 && b1); lanes differing only in bit 0 hold duplicates.
__device__ __forceinline__ float warp_reduce12(float (&v)[12], int lane)
{
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float send = u16 ? v[k] : v[k + 6];
        const float keep = u16 ? v[k + 6] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float send = u8 ? v[k] : v[k + 3];
        const float keep = u8 ? v[k + 3] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float send0 = u4 ? v[0] : v[2];
        const float keep0 = u4 ? v[2] : v[0];
        const float send1 = u4 ? v[1] : 0.0f;
        const float keep1 = u4 ? 0.0f : v[1];
        v[0] = keep0 + __shfl_xor_sync(0xffffffffu, send0, 4);
        v[1] = keep1 + __shfl_xor_sync(0xffffffffu, send1, 4);
    }
    {
        const float send = u2 ? v[0] : v[1];
        const float keep = u2 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

template <bool DEPTH>
__global__ void __launch_bounds__(32, 12) k_raster_bwd(const __grid_constant__ ViewParams vp,
                                                       const uint32_t* __restrict__ tile_ranges,
                                                       const uint32_t* __restrict__ tile_order,
                                                       const float4* __restrict__ rec, const uint32_t* __restrict__ vals0,
                                                       const uint32_t* __restrict__ vals1,
                                                       const uint32_t* __restrict__ d_result_buf,
                                                       const float* __restrict__ cot_color,
                                                       const float* __restrict__ cot_depth, const float* __restrict__ cot_alpha,
                                                       const float* __restrict__ out_alpha,
                                                       const uint32_t* __restrict__ last_contrib, float* __restrict__ grad_rec)
{
    __shared__ __align__(128) float4 s_rec[2][RB_BWD * 3];
    __shared__ __align__(16) float s_out[RB_BWD][12];   // per-Gaussian sums of this block for one batch
    __shared__ __align__(8) uint64_t s_bar[2];
    const BlockMap bm = map_block(vp, tile_order);
    const int lane = threadIdx.x;
    const uint32_t start = tile_ranges[bm.tile * 2], end = tile_ranges[bm.tile * 2 + 1];
    const uint32_t count = end > start ? end - start : 0u;

    // thread = column (lane & 15) x 8 consecutive rows starting at (lane >> 4) * 8
    const int pxi = bm.x0 + (lane & 15);
    const int py0 = bm.y0 + (lane >> 4) * BPPT;
    float sT[BPPT], kX[BPPT], kY[BPPT], kZ[BPPT], kD[BPPT], kT[BPPT];
    uint32_t nC[BPPT];
    uint32_t nmax = 0, nmin = 0xffffffffu;
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
            sT[p] = 1.0f - out_alpha[pix];
            kT[p] = -cotA + (vp.whiteBg ? (kX[p] + kY[p] + kZ[p]) : 0.0f);
            nC[p] = min(last_contrib[pix], count);
        }
        nmax = max(nmax, nC[p]);
        nmin = min(nmin, nC[p]);
    }
    // only Gaussians below the block-wide max nContrib can contribute
    const uint32_t used = __reduce_max_sync(0xffffffffu, nmax);
    const int nb = (int)((used + RB_BWD - 1) / RB_BWD);
    if (nb == 0) return;
    if (lane == 0) {
        mbar_init(&s_bar[0], GSB_GATHER_TMA ? 1 : 32);
        mbar_init(&s_bar[1], GSB_GATHER_TMA ? 1 : 32);
        mbar_fence_init();
    }
    __syncwarp();

    // batches are visited last -> first; sequence number s = nb-1-b selects stage / parity
    // gather staging as in the forward: each lane pulls two 48-byte records per batch with TMA bulk copies
    const uint32_t* __restrict__ vals = (*d_result_buf) ? vals1 : vals0;
    uint32_t ia = 0xffffffffu, ib = 0xffffffffu;   // indices of this lane's two slots of the next batch to issue
    auto load_idx = [&](int b) {
        const uint32_t j0 = (uint32_t)b * RB_BWD + lane, j1 = j0 + 32;
        ia = (b >= 0 && j0 < used) ? vals[start + j0] : 0xffffffffu;
        ib = (b >= 0 && j1 < used) ? vals[start + j1] : 0xffffffffu;
    };
    auto issue = [&](int b) {
        const int s = nb - 1 - b;
        uint64_t* bar = &s_bar[s & 1];
#if GSB_GATHER_TMA
        if (lane == 0) mbar_expect_tx(bar, min((uint32_t)RB_BWD, used - (uint32_t)b * RB_BWD) * 48u);
        if (ia != 0xffffffffu) bulk_g2s(&s_rec[s & 1][lane * 3], rec + (size_t)ia * 3, 48u, bar);
        if (ib != 0xffffffffu) bulk_g2s(&s_rec[s & 1][(lane + 32) * 3], rec + (size_t)ib * 3, 48u, bar);
#else
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
#endif
    };
    load_idx(nb - 1);
    issue(nb - 1);
    load_idx(nb - 2);

    const float pxf = (float)pxi, pyf = (float)py0;
    const uint32_t rec_base = smem_u32(&s_rec[0][0]);
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    const int comp = 6 * b4 + 3 * b3 + (b2 ? 2 : b1);
    const bool writer = !(lane & 1) && !(b2 && b1);

    for (int b = nb - 1; b >= 0; --b) {
        const int s = nb - 1 - b;
        __syncwarp();                               // every lane is done with the stage being refilled
        if (b > 0) issue(b - 1);
        load_idx(b - 2);
        {
            float4* z = reinterpret_cast<float4*>(&s_out[0][0]);
            for (int i = lane; i < RB_BWD * 3; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&s_bar[s & 1], (uint32_t)(s >> 1) & 1u);
        __syncwarp();
        const int n = (int)min((uint32_t)RB_BWD, used - (uint32_t)b * RB_BWD);
        const uint32_t stage_addr = rec_base + (uint32_t)(s & 1) * (RB_BWD * 48u);
        for (int j = n - 1; j >= 0; --j) {
            const uint32_t i = (uint32_t)(b * RB_BWD + j);
            if (!__any_sync(0xffffffffu, i < nmax)) continue;
            const uint32_t addr = stage_addr + (uint32_t)j * 48u;
            const float4 a = lds128(addr), q = lds128(addr + 16), c = lds128(addr + 32);
            const float dx = pxf - a.x, dyb = pyf - a.y;
            // The thread's 8 pixels share dx and have dy = dyb + p, so the exponent is a quadratic in the
            // compile-time row offset p:  e(p) = E0 + p*E1 + p^2*C  (log2 units, opacity not included)
            const float E0 = fmaf(dx, fmaf(a.z, dx, a.w * dyb), q.x * dyb * dyb);
            const float E1 = fmaf(a.w, dx, 2.0f * q.x * dyb);
            // per-thread partial sums over the 8 pixels: colour/depth terms and the moments of
            // h = dL/d(opacity-free alpha) in the row offset: H0 = sum h, H1 = sum p h, H2 = sum p^2 h
            float Cr = 0.f, Cg = 0.f, Cb = 0.f, Cd = 0.f, H0 = 0.f, H1 = 0.f, H2 = 0.f;
            // one pixel of this thread.  MASKED = false when every pixel of the block is known to be active.
            // Neither variant branches, so the compiler interleaves the 8 dependent chains; in the masked
            // variant an inactive pixel (i >= nContrib) computes and discards (selects keep its state).
            auto pixel = [&](int p, bool masked) {
                const bool act = !masked || (i < nC[p]);
                const float pf = (float)p;
                const float ex = ex2_approx(fmaf(pf, fmaf(pf, q.x, E1), E0));    // :437-483
                const float raw = ex * c.y;
                const bool keep = act && !(raw > 0.99f);   // the alpha clamp branch has zero gradient
                const float alpha = fminf(raw, 0.99f);
                // undoTileGlobalPixelState (:501-521): only the transmittance matters for the gradients;
                // max(1 - alpha, 1e-6) == 1 - alpha because alpha <= 0.99
                const float prevT = sT[p] * rcp_approx(1.0f - alpha);
                const float contrib = act ? prevT * alpha : 0.0f;
                sT[p] = act ? prevT : sT[p];
                // VJP of updateTileGlobalPixelState (:485-499)
                float dotc = fmaf(kZ[p], c.x, fmaf(kY[p], q.w, kX[p] * q.z));
                if (DEPTH) dotc = fmaf(kD[p], c.z, dotc);
                const float d = dotc - kT[p];
                const float g_alpha = prevT * d;
                kT[p] = act ? fmaf(alpha, d, kT[p]) : kT[p];
                Cr = fmaf(contrib, kX[p], Cr);
                Cg = fmaf(contrib, kY[p], Cg);
                Cb = fmaf(contrib, kZ[p], Cb);
                if (DEPTH) Cd = fmaf(contrib, kD[p], Cd);
                // VJP of evaluateTileGlobalSample: everything geometric is a moment of h
                const float h = keep ? g_alpha * ex : 0.0f;
                H0 += h;
                if (p > 0) {
                    H1 = fmaf(h, pf, H1);
                    H2 = fmaf(h, pf * pf, H2);
                }
            };
            if (__all_sync(0xffffffffu, i < nmin)) {
#pragma unroll
                for (int p = 0; p < BPPT; ++p) pixel(p, false);
            } else {
#pragma unroll
                for (int p = 0; p < BPPT; ++p) pixel(p, true);
            }
            // moments in (dx, dy) of this thread's pixels: sum h dy = dyb H0 + H1, sum h dy^2 = dyb^2 H0 + 2 dyb H1 + H2
            const float Sy = fmaf(dyb, H0, H1);
            const float Syy = fmaf(dyb, fmaf(dyb, H0, H1 + H1), H2);
            const float Sx = dx * H0;
            float g[12] = {Cr, Cg, Cb, Cd, H0, Sx, Sy, dx * Sx, dx * Sy, Syy, 0.0f, 0.0f};
            const float tot = warp_reduce12(g, lane);
            if (writer) s_out[j][comp] = tot;
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
            const float op = c.y;
            const float kc = -0.5f * op;                 // d(natural exponent)/d conic = -0.5 * (dx^2, dx dy, dy^2)
            const float km = -op * (1.0f / LOG2E_F);     // d(natural exponent)/d mean = -(2a dx + b dy, ...), (a,b,c) = (A,B,C)/log2 e
            const float g_mx = km * fmaf(a.z + a.z, s1.y, a.w * s1.z);
            const float g_my = km * fmaf(q.x + q.x, s1.z, a.w * s1.y);
            const float g_c00 = kc * s1.w, g_c01 = kc * s2.x, g_c11 = kc * s2.y;
            float* dst = grad_rec + (size_t)__float_as_uint(c.w) * REC_FLOATS;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g_mx), "f"(g_my), "f"(g_c00), "f"(g_c01) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(g_c01), "f"(g_c11), "f"(s0.x), "f"(s0.y) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 8), "f"(s0.z), "f"(s1.x), "f"(s0.w), "f"(0.0f) : "memory");
        }
    }
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

cudaError_t launch_raster_fwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges,
                              const uint32_t* tile_order, const float* rec, const uint32_t* vals0, const uint32_t* vals1,
                              const uint32_t* d_result_buf,
                              float* out_color, float* out_depth, float* out_alpha, uint32_t* out_last)
{
    const int blocks = raster_blocks(vp);
    if (blocks > 0)
        k_raster_fwd<<<blocks, RT, 0, st>>>(vp, tile_ranges, tile_order, reinterpret_cast<const float4*>(rec), vals0, vals1,
                                            d_result_buf, out_color, out_depth, out_alpha, out_last);
    return cudaGetLastError();
}

cudaError_t launch_raster_bwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges,
                              const uint32_t* tile_order, const float* rec, const uint32_t* vals0, const uint32_t* vals1,
                              const uint32_t* d_result_buf,
                              const float* cot_color, const float* cot_depth, const float* cot_alpha,
                              const float* out_color, const float* out_depth, const float* out_alpha,
                              const uint32_t* last_contrib, float* grad_rec)
{
    (void)out_color; (void)out_depth;   // the colour/depth state is not needed by the gradients (see kernel)
    const int blocks = raster_blocks(vp);
    if (blocks > 0) {
        static bool carveout_set = false;
        if (!carveout_set) {
            cudaFuncSetAttribute(k_raster_bwd<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_raster_bwd<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carveout_set = true;
        }
        if (cot_depth)
            k_raster_bwd<true><<<blocks, 32, 0, st>>>(vp, tile_ranges, tile_order, reinterpret_cast<const float4*>(rec), vals0, vals1,
                                                      d_result_buf, cot_color, cot_depth, cot_alpha, out_alpha, last_contrib, grad_rec);
        else
            k_raster_bwd<false><<<blocks, 32, 0, st>>>(vp, tile_ranges, tile_order, reinterpret_cast<const float4*>(rec), vals0, vals1,
                                                       d_result_buf, cot_color, cot_depth, cot_alpha, out_alpha, last_contrib, grad_rec);
    }
    return cudaGetLastError();
}

}  // namespace gsb
