// raster.cu — tile alpha-blend forward (K9) and back-to-front backward (K10) for sm_100a.
//
// Reference: slang/gaussian_tile_global_kernels.slang:437-614 (forward), :501-521 + :648-881
// (backward), call sites Trainer/GaussianRenderer.swift:124-147,187-226.
//
// One CTA = one 16x16 pixel block of one tile (a tile larger than 16x16 is covered by several
// CTAs that share the tile's list).  The tile's depth-ordered 48-byte records are contiguous in the
// `staged` stream (binning.cu), so batches are pulled into shared memory with TMA 1-D bulk copies
// (cp.async.bulk → SASS UBLKCP) signalled through mbarriers, double-buffered against the blend loop.
// FP32-pipe bound: 27 flop + 1 ex2 per (pixel, Gaussian) forward, ≈80 flop + 1 ex2 + 1 div backward.
#include <algorithm>

#include "kernels.h"

namespace gsb {

constexpr int RT = 256;          // threads per CTA = 16x16 pixels
constexpr int RB_FWD = 128;      // records per forward batch (6 KB)
constexpr int RB_BWD = 64;       // records per backward batch
constexpr int RWARPS = RT / 32;

struct PixelMap {
    int tile, px, py;
    bool active;
};

__device__ __forceinline__ PixelMap map_pixel(const ViewParams& vp)
{
    const int subX = (vp.tileW + 15) >> 4, subY = (vp.tileH + 15) >> 4;
    const int per = subX * subY;
    PixelMap m;
    m.tile = blockIdx.x / per;
    const int sb = blockIdx.x - m.tile * per;
    const int tileX = m.tile % vp.gridW, tileY = m.tile / vp.gridW;
    const int lx = (sb % subX) * 16 + (threadIdx.x & 15);
    const int ly = (sb / subX) * 16 + (threadIdx.x >> 4);
    m.px = tileX * vp.tileW + lx;
    m.py = tileY * vp.tileH + ly;
    m.active = lx < vp.tileW && ly < vp.tileH && m.px < vp.W && m.py < vp.H;
    return m;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RT) k_raster_fwd(const __grid_constant__ ViewParams vp,
                                                   const uint32_t* __restrict__ tile_ranges,
                                                   const float4* __restrict__ staged, float* __restrict__ out_color,
                                                   float* __restrict__ out_depth, float* __restrict__ out_alpha,
                                                   uint32_t* __restrict__ out_last)
{
    __shared__ __align__(128) float4 s_rec[2][RB_FWD * 3];
    __shared__ __align__(8) uint64_t s_bar[2];
    const PixelMap pm = map_pixel(vp);
    const uint32_t start = tile_ranges[pm.tile * 2], end = tile_ranges[pm.tile * 2 + 1];
    const uint32_t count = end > start ? end - start : 0u;
    const int nb = (int)((count + RB_FWD - 1) / RB_FWD);

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int b) {
        const uint32_t n = min((uint32_t)RB_FWD, count - (uint32_t)b * RB_FWD);
        uint64_t* bar = &s_bar[b & 1];
        mbar_expect_tx(bar, n * 48u);
        bulk_g2s(&s_rec[b & 1][0], staged + ((size_t)start + (size_t)b * RB_FWD) * 3, n * 48u, bar);
    };
    if (threadIdx.x == 0 && nb > 0) issue(0);

    const float px = (float)pm.px, py = (float)pm.py;
    float cx = 0.f, cy = 0.f, cz = 0.f, dep = 0.f, T = 1.0f;
    uint32_t nContrib = count;
    bool done = !pm.active;

    for (int b = 0; b < nb; ++b) {
        if (threadIdx.x == 0 && b + 1 < nb) issue(b + 1);
        mbar_wait(&s_bar[b & 1], (uint32_t)(b >> 1) & 1u);
        const int n = (int)min((uint32_t)RB_FWD, count - (uint32_t)b * RB_FWD);
        const float4* r = &s_rec[b & 1][0];
        if (!__all_sync(0xffffffffu, done)) {
            for (int j = 0; j < n; ++j) {
                if (!done) {
                    const float4 a = r[j * 3], q = r[j * 3 + 1], c = r[j * 3 + 2];
                    const float dx = px - a.x, dy = py - a.y;
                    const float dxdy = dx * dy;
                    const float expo = -0.5f * (dx * dx * a.z + dy * dy * q.y + dxdy * a.w + dxdy * q.x);
                    const float raw = __expf(expo) * c.y;
                    const float alpha = raw > 0.99f ? 0.99f : raw;
                    const float contrib = T * alpha;
                    cx += contrib * q.z;
                    cy += contrib * q.w;
                    cz += contrib * c.x;
                    dep += contrib * c.z;
                    T *= (1.0f - alpha);
                    if (T < 1e-4f) {
                        nContrib = (uint32_t)(b * RB_FWD + j + 1);
                        done = true;
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        // releases the stage buffer for the copy issued two batches later, and votes on early exit
        if (__syncthreads_and(done)) {
            if (b + 1 < nb) mbar_wait(&s_bar[(b + 1) & 1], (uint32_t)((b + 1) >> 1) & 1u);  // drain the in-flight copy
            break;
        }
    }
    if (pm.active) {
        const size_t p = (size_t)pm.py * vp.W + pm.px;
        const float bg = vp.whiteBg ? T : 0.0f;
        out_color[p * 3 + 0] = cx + bg;
        out_color[p * 3 + 1] = cy + bg;
        out_color[p * 3 + 2] = cz + bg;
        out_depth[p] = dep;
        out_alpha[p] = 1.0f - T;
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

__global__ void __launch_bounds__(RT) k_raster_bwd(const __grid_constant__ ViewParams vp,
                                                   const uint32_t* __restrict__ tile_ranges,
                                                   const float4* __restrict__ staged, const float* __restrict__ cot_color,
                                                   const float* __restrict__ cot_depth, const float* __restrict__ cot_alpha,
                                                   const float* __restrict__ out_color, const float* __restrict__ out_depth,
                                                   const float* __restrict__ out_alpha, const uint32_t* __restrict__ last_contrib,
                                                   float* __restrict__ grad_rec)
{
    __shared__ __align__(128) float4 s_rec[2][RB_BWD * 3];
    __shared__ __align__(16) float s_acc[RWARPS][RB_BWD][12];   // per-warp partial sums, no atomics
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_max[RWARPS];
    const PixelMap pm = map_pixel(vp);
    const uint32_t start = tile_ranges[pm.tile * 2], end = tile_ranges[pm.tile * 2 + 1];
    const uint32_t count = end > start ? end - start : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    float sX = 0.f, sY = 0.f, sZ = 0.f, sD = 0.f, sT = 0.f;
    float kX = 0.f, kY = 0.f, kZ = 0.f, kD = 0.f, kT = 0.f;
    uint32_t nContrib = 0;
    const float px = (float)pm.px, py = (float)pm.py;
    if (pm.active) {
        // slang/gaussian_tile_global_kernels.slang:696-723
        const size_t p = (size_t)pm.py * vp.W + pm.px;
        kX = cot_color[p * 3];
        kY = cot_color[p * 3 + 1];
        kZ = cot_color[p * 3 + 2];
        kD = cot_depth ? cot_depth[p] : 0.0f;
        const float cotA = cot_alpha ? cot_alpha[p] : 0.0f;
        const float trans = 1.0f - out_alpha[p];
        const float bg = vp.whiteBg ? trans : 0.0f;
        sX = out_color[p * 3] - bg;
        sY = out_color[p * 3 + 1] - bg;
        sZ = out_color[p * 3 + 2] - bg;
        sD = out_depth[p];
        sT = trans;
        kT = -cotA + (vp.whiteBg ? (kX + kY + kZ) : 0.0f);
        nContrib = min(last_contrib[p], count);
    }
    // only Gaussians below the block-wide max nContrib can contribute
    uint32_t wmax = __reduce_max_sync(0xffffffffu, nContrib);
    if (lane == 0) s_max[warp] = wmax;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t used = 0;
#pragma unroll
    for (int w = 0; w < RWARPS; ++w) used = max(used, s_max[w]);
    const int nb = (int)((used + RB_BWD - 1) / RB_BWD);
    if (nb == 0) return;

    // batches are visited last → first; sequence number s = nb-1-b selects stage / parity
    auto issue = [&](int b) {
        const int s = nb - 1 - b;
        const uint32_t n = min((uint32_t)RB_BWD, used - (uint32_t)b * RB_BWD);
        uint64_t* bar = &s_bar[s & 1];
        mbar_expect_tx(bar, n * 48u);
        bulk_g2s(&s_rec[s & 1][0], staged + ((size_t)start + (size_t)b * RB_BWD) * 3, n * 48u, bar);
    };
    if (threadIdx.x == 0) issue(nb - 1);

    for (int b = nb - 1; b >= 0; --b) {
        const int s = nb - 1 - b;
        if (threadIdx.x == 0 && b > 0) issue(b - 1);
        // zero this warp's partial sums (own slice only → no barrier needed before use)
        {
            float4* z = reinterpret_cast<float4*>(&s_acc[warp][0][0]);
            for (int i = lane; i < RB_BWD * 3; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        mbar_wait(&s_bar[s & 1], (uint32_t)(s >> 1) & 1u);
        const int n = (int)min((uint32_t)RB_BWD, used - (uint32_t)b * RB_BWD);
        const float4* r = &s_rec[s & 1][0];
        for (int j = n - 1; j >= 0; --j) {
            const uint32_t i = (uint32_t)(b * RB_BWD + j);
            const bool act = i < nContrib;
            if (!__any_sync(0xffffffffu, act)) continue;
            float g[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) g[k] = 0.0f;
            if (act) {
                const float4 a = r[j * 3], q = r[j * 3 + 1], c = r[j * 3 + 2];
                const float dx = px - a.x, dy = py - a.y;
                const float dxdy = dx * dy;
                const float expo = -0.5f * (dx * dx * a.z + dy * dy * q.y + dxdy * a.w + dxdy * q.x);
                const float ex = __expf(expo);
                const float raw = ex * c.y;
                const bool clamped = raw > 0.99f;
                const float alpha = clamped ? 0.99f : raw;
                // undoTileGlobalPixelState (:501-521)
                float denom = 1.0f - alpha;
                if (denom < 1e-6f) denom = 1e-6f;
                const float prevT = sT / denom;
                const float contrib = prevT * alpha;
                sX -= contrib * q.z;
                sY -= contrib * q.w;
                sZ -= contrib * c.x;
                sD -= contrib * c.z;
                sT = prevT;
                // VJP of updateTileGlobalPixelState (:485-499)
                const float dotc = kX * q.z + kY * q.w + kZ * c.x + kD * c.z;
                const float g_alpha = prevT * (dotc - kT);
                kT = alpha * dotc + (1.0f - alpha) * kT;
                g[6] = contrib * kX;
                g[7] = contrib * kY;
                g[8] = contrib * kZ;
                g[10] = contrib * kD;
                // VJP of evaluateTileGlobalSample / tileGlobalAlphaFromGaussian (:437-483)
                const float g_raw = clamped ? 0.0f : g_alpha;
                g[9] = g_raw * ex;
                const float gq = -0.5f * (g_raw * raw);
                g[2] = gq * dx * dx;
                g[5] = gq * dy * dy;
                g[3] = gq * dxdy;
                g[4] = g[3];
                const float csum = a.w + q.x;
                g[0] = -gq * (2.0f * dx * a.z + dy * csum);
                g[1] = -gq * (2.0f * dy * q.y + dx * csum);
            }
            const float tot = warp_reduce12(g, lane);
            const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
            const int comp = 6 * b4 + 3 * b3 + (b2 ? 2 : b1);
            if (!(lane & 1) && !(b2 && b1)) s_acc[warp][j][comp] = tot;
        }
        __syncthreads();
        // flush: sum the 8 warp slices, one 16-byte vector reduction per (record, quad)
        for (int it = threadIdx.x; it < n * 3; it += RT) {
            const int j = it / 3, qd = it - j * 3;
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < RWARPS; ++w) {
                const float4 v = *reinterpret_cast<const float4*>(&s_acc[w][j][qd * 4]);
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
            if (sum.x != 0.f || sum.y != 0.f || sum.z != 0.f || sum.w != 0.f) {
                const uint32_t gi = __float_as_uint(r[j * 3 + 2].w);
                float* dst = grad_rec + (size_t)gi * REC_FLOATS + qd * 4;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(sum.x), "f"(sum.y), "f"(sum.z),
                             "f"(sum.w)
                             : "memory");
            }
        }
        __syncthreads();
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

cudaError_t launch_raster_fwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges, const float* staged,
                              float* out_color, float* out_depth, float* out_alpha, uint32_t* out_last)
{
    const int blocks = raster_blocks(vp);
    if (blocks > 0)
        k_raster_fwd<<<blocks, RT, 0, st>>>(vp, tile_ranges, reinterpret_cast<const float4*>(staged), out_color, out_depth,
                                            out_alpha, out_last);
    return cudaGetLastError();
}

cudaError_t launch_raster_bwd(cudaStream_t st, const ViewParams& vp, const uint32_t* tile_ranges, const float* staged,
                              const float* cot_color, const float* cot_depth, const float* cot_alpha,
                              const float* out_color, const float* out_depth, const float* out_alpha,
                              const uint32_t* last_contrib, float* grad_rec)
{
    const int blocks = raster_blocks(vp);
    if (blocks > 0)
        k_raster_bwd<<<blocks, RT, 0, st>>>(vp, tile_ranges, reinterpret_cast<const float4*>(staged), cot_color, cot_depth,
                                            cot_alpha, out_color, out_depth, out_alpha, last_contrib, grad_rec);
    return cudaGetLastError();
}

}  // namespace gsb
