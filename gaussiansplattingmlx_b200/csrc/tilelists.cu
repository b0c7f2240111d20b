// tilelists.cu — per-tile Gaussian lists by a two-level stable partition (sm_100a).
//
// Reference semantics (slang/gaussian_tile_global_kernels.slang:73-404, Trainer/GaussianRenderer.swift:333-490):
// tile t's list = the Gaussians whose clamped rect covers t, ascending asuint(depth), ties by Gaussian index.
// The Gaussians arrive already in that order (depth pre-sort, binning.cu), so a tile's list is a FILTER of the
// depth-ordered sequence.  Sorting M = 12 M (Gaussian, tile) pairs by tile id moves 36 B/pair through HBM; this
// file builds the same lists while only ever sorting the ~3.5x fewer (Gaussian, superblock) pairs:
//
//   level 1   superblock = SBW x SBH = 4 x 2 tiles.  Emit one (superblock id, Gaussian) pair per superblock a rect
//             touches, in depth order (k_generate_keys of binning.cu with a superblock grid), stable radix sort on
//             the superblock id (the onesweep of binning.cu): every superblock now owns a depth-ordered slice.
//   level 2   a CTA per superblock, its slice cut into 8 contiguous warp slices.  COUNT: each warp walks its slice
//             32 entries at a time, rebuilds the 8-bit "which of my 8 tiles does this rect cover" mask from the
//             packed tile rect and counts per tile with ballots.  The per-tile totals are scanned
//             (binning.cu), k_tile_bases turns them into CSR ranges and (superblock, warp slice, tile) bases, and
//             k_tile_order2 gives the heavy-first launch order.  FILL: the same walk again; lane ranks from the ballots give every (Gaussian, tile) pair its
//             final slot — stable by construction, no atomics, writes coalesced per tile.
//
// HBM traffic per view at C3: ~3.4 M pairs x (8 B keygen + 36 B sort) + 2 x (3.4 M x 12 B) walks + 48 MB of list
// writes = ~280 MB instead of ~580 MB, and the result is bit-identical (tests/test_gpu_parity.py).
#include "kernels.h"

namespace gsb {

// packed tile rect (binning.cu / project.cu): x = x0 | y0 << 16, y = x1 | y1 << 16, exclusive upper bounds, 0 = empty
__device__ __forceinline__ uint32_t sb_tile_mask(uint2 r, int tx0, int ty0)
{
    const int x0 = (int)(r.x & 0xffff) - tx0, y0 = (int)(r.x >> 16) - ty0;
    const int x1 = (int)(r.y & 0xffff) - tx0, y1 = (int)(r.y >> 16) - ty0;
    const int lx = max(x0, 0), hx = min(x1, SBW), ly = max(y0, 0), hy = min(y1, SBH);
    if (hx <= lx || hy <= ly) return 0u;
    const uint32_t xm = ((1u << hx) - 1u) & ~((1u << lx) - 1u);          // SBW-bit row mask
    uint32_t m = 0u;
#pragma unroll
    for (int y = 0; y < SBH; ++y)
        if (y >= ly && y < hy) m |= xm << (y * SBW);
    return m;
}

// ------------------------------------------------------------------------------------------------
// level 2: count / fill.  One CTA (8 warps) per superblock, warp w owns slice w of the superblock's depth-ordered
// list.  counts / bases are indexed [superblock][warp][tile-in-superblock].
// ------------------------------------------------------------------------------------------------

template <bool FILL>
__global__ void __launch_bounds__(L2_WARPS * 32) k_l2_walk(int sbGridW, const uint32_t* __restrict__ sb_ranges,
                                                           const uint32_t* __restrict__ vals0, const uint32_t* __restrict__ vals1,
                                                           const uint32_t* __restrict__ d_result_buf,
                                                           const uint2* __restrict__ tile_rects, uint32_t* __restrict__ slice_counts,
                                                           const uint32_t* __restrict__ slice_base, uint32_t* __restrict__ list,
                                                           uint32_t capacity, uint32_t* __restrict__ tile_counts, int gridW, int gridH)
{
    __shared__ uint32_t s_cnt[L2_WARPS][SB_TILES];
    const int s = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t begin = sb_ranges[s * 2], end = sb_ranges[s * 2 + 1];
    const uint32_t n = end > begin ? end - begin : 0u;
    const int slice = blockIdx.y * L2_WARPS + warp;      // blockIdx.y < L2_SPLIT
    const uint32_t per = (n + L2_SLICES - 1) / L2_SLICES;
    const uint32_t w0 = min(begin + slice * per, end), w1 = min(w0 + per, end);
    const uint32_t* __restrict__ vals = (d_result_buf && *d_result_buf) ? vals1 : vals0;
    const int tx0 = (s % sbGridW) * SBW, ty0 = (s / sbGridW) * SBH;
    const uint32_t lt = (1u << lane) - 1u;
    const size_t slot = ((size_t)s * L2_SLICES + slice) * SB_TILES;
    uint32_t run[SB_TILES];
#pragma unroll
    for (int t = 0; t < SB_TILES; ++t) run[t] = FILL ? slice_base[slot + t] : 0u;
    // software pipeline over the dependent gather (index -> rect): the index of step k+2 and the rect of step k+1 are
    // in flight while step k is balloted, so no load is consumed in the iteration that issues it
    auto ld_idx = [&](uint32_t j) { return j < w1 ? vals[j] : 0xffffffffu; };
    auto ld_rect = [&](uint32_t g) { return g != 0xffffffffu ? tile_rects[g] : make_uint2(0u, 0u); };
    uint32_t g1 = ld_idx(w0 + lane), g2 = ld_idx(w0 + 32 + lane);
    uint2 r1 = ld_rect(g1);
    for (uint32_t j = w0; j < w1; j += 32) {
        const uint32_t g = g1;
        const uint2 r = r1;
        g1 = g2;
        g2 = ld_idx(j + 64 + lane);
        r1 = ld_rect(g1);
        const uint32_t mask = g != 0xffffffffu ? sb_tile_mask(r, tx0, ty0) : 0u;
#pragma unroll
        for (int t = 0; t < SB_TILES; ++t) {
            const uint32_t b = __ballot_sync(0xffffffffu, (mask >> t) & 1u);
            if (FILL) {
                if ((mask >> t) & 1u) {
                    const uint32_t dst = run[t] + __popc(b & lt);
                    if (dst < capacity) list[dst] = g;
                }
            }
            run[t] += __popc(b);
        }
    }
    if (!FILL) {
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < SB_TILES; ++t) {
                slice_counts[slot + t] = run[t];
                s_cnt[warp][t] = run[t];
            }
        }
        __syncthreads();
        if (threadIdx.x < SB_TILES) {   // tile totals of this superblock (each tile belongs to exactly one superblock)
            const int t = threadIdx.x;
            const int tx = tx0 + t % SBW, ty = ty0 + t / SBW;
            if (tx < gridW && ty < gridH) {
                uint32_t c = 0;
#pragma unroll
                for (int w = 0; w < L2_WARPS; ++w) c += s_cnt[w][t];
                if (c) atomicAdd(&tile_counts[ty * gridW + tx], c);   // L2_SPLIT CTAs per superblock: zeroed by the launcher
            }
        }
    }
}

cudaError_t launch_l2_count(cudaStream_t st, int numSB, int sbGridW, int gridW, int gridH, const uint32_t* sb_ranges, const uint32_t* vals0,
                            const uint32_t* vals1, const uint32_t* d_result_buf, const uint2* tile_rects, uint32_t* slice_counts,
                            uint32_t* tile_counts)
{
    if (numSB > 0) {
        cudaError_t e = cudaMemsetAsync(tile_counts, 0, (size_t)gridW * gridH * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        k_l2_walk<false><<<dim3(numSB, L2_SPLIT), L2_WARPS * 32, 0, st>>>(sbGridW, sb_ranges, vals0, vals1, d_result_buf, tile_rects, slice_counts,
                                                                          nullptr, nullptr, 0u, tile_counts, gridW, gridH);
    }
    return cudaGetLastError();
}
cudaError_t launch_l2_fill(cudaStream_t st, int numSB, int sbGridW, const uint32_t* sb_ranges, const uint32_t* vals0,
                           const uint32_t* vals1, const uint32_t* d_result_buf, const uint2* tile_rects, const uint32_t* slice_base,
                           uint32_t* list, uint32_t capacity)
{
    if (numSB > 0)
        k_l2_walk<true><<<dim3(numSB, L2_SPLIT), L2_WARPS * 32, 0, st>>>(sbGridW, sb_ranges, vals0, vals1, d_result_buf, tile_rects, nullptr, slice_base,
                                                         list, capacity, nullptr, 0, 0);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// tile counts (written by the count walk) -> exclusive scan (binning.cu, decoupled look-back) -> this kernel:
// tile ranges in the reference convention ((0,0) for an empty tile) and the base of every (superblock, warp
// slice, tile) run.  One thread per tile.
// ------------------------------------------------------------------------------------------------
// Heavy-first launch order of the rasterisers: tiles are dropped into TO_BUCKETS log-spaced length buckets (8 per octave,
// longest first).  k_tile_bases counts the tiles per bucket, k_tile_order2 turns the counts into bucket bases and hands out the
// slots with one atomic per tile - the order inside a bucket is arbitrary (it only schedules work; results do not depend on
// it).  Replaces a single-CTA counting sort (18 us on the critical path of every view) by a 32-CTA kernel.
__device__ __forceinline__ uint32_t tile_order_bucket(uint32_t c)
{
    if (c == 0u) return TO_BUCKETS - 1;
    const int msb = 31 - __clz(c);
    const uint32_t sub = msb >= 3 ? ((c >> (msb - 3)) & 7u) : ((c << (3 - msb)) & 7u);
    const uint32_t rank = (uint32_t)msb * 8u + sub;          // monotone in c, < 256
    return (TO_BUCKETS - 2) - min(rank, (uint32_t)(TO_BUCKETS - 2));
}

__global__ void __launch_bounds__(256) k_tile_bases(int gridW, int gridH, int sbGridW, const uint32_t* __restrict__ tile_counts,
                                                    const uint32_t* __restrict__ tile_starts, const uint32_t* __restrict__ slice_counts,
                                                    uint32_t* __restrict__ slice_base, uint32_t* __restrict__ tile_ranges,
                                                    uint32_t* __restrict__ bucket_hist)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= gridW * gridH) return;
    if (bucket_hist) atomicAdd(&bucket_hist[tile_order_bucket(tile_counts[t])], 1u);
    const int tx = t % gridW, ty = t / gridW;
    const int s = (ty / SBH) * sbGridW + tx / SBW;
    const size_t sl = (size_t)s * L2_SLICES * SB_TILES + (size_t)((ty % SBH) * SBW + tx % SBW);
    const uint32_t start = tile_starts[t], c = tile_counts[t];
    uint32_t b = start;
#pragma unroll 8
    for (int w = 0; w < L2_SLICES; ++w) {
        slice_base[sl + (size_t)w * SB_TILES] = b;
        b += slice_counts[sl + (size_t)w * SB_TILES];
    }
    tile_ranges[t * 2 + 0] = c ? start : 0u;       // compute_tile_ranges leaves (0,0) for tiles nobody touches
    tile_ranges[t * 2 + 1] = c ? start + c : 0u;
}

__global__ void __launch_bounds__(256) k_tile_order2(int numTiles, const uint32_t* __restrict__ tile_counts, const uint32_t* __restrict__ bucket_hist,
                                                     uint32_t* __restrict__ bucket_cursor, uint32_t* __restrict__ order)
{
    __shared__ uint32_t s_base[TO_BUCKETS];
    __shared__ uint32_t s_warp[8];
    static_assert(TO_BUCKETS == 256, "one bucket per thread");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c = bucket_hist[tid];
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t pre = 0;
    for (int w = 0; w < warp; ++w) pre += s_warp[w];
    s_base[tid] = pre + inc - c;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + tid;
    if (t < numTiles) {
        const uint32_t b = tile_order_bucket(tile_counts[t]);
        order[s_base[b] + atomicAdd(&bucket_cursor[b], 1u)] = (uint32_t)t;
    }
}

// order_ws: 2 * TO_BUCKETS words (bucket histogram, bucket cursors), zeroed here
cudaError_t launch_tile_bases(cudaStream_t st, int gridW, int gridH, int sbGridW, const uint32_t* tile_counts, const uint32_t* tile_starts,
                              const uint32_t* slice_counts, uint32_t* slice_base, uint32_t* tile_ranges, uint32_t* order_ws, uint32_t* order)
{
    const int n = gridW * gridH;
    if (n <= 0) return cudaSuccess;
    if (order_ws) {
        cudaError_t e = cudaMemsetAsync(order_ws, 0, 2 * TO_BUCKETS * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
    }
    k_tile_bases<<<cdiv(n, 256), 256, 0, st>>>(gridW, gridH, sbGridW, tile_counts, tile_starts, slice_counts, slice_base, tile_ranges, order_ws);
    if (order_ws) k_tile_order2<<<cdiv(n, 256), 256, 0, st>>>(n, tile_counts, order_ws, order_ws + TO_BUCKETS, order);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// parity API: the reference's sortedKeysHigh / sortedKeysLow for list position j
// (tile id by binary search in the monotone tile starts, depth bits of the listed Gaussian)
// ------------------------------------------------------------------------------------------------
__global__ void k_expand_sorted_keys(uint32_t M, int numTiles, const uint32_t* __restrict__ tile_starts, const uint32_t* __restrict__ list,
                                     const float* __restrict__ depth_ptr, int depth_stride, uint32_t* __restrict__ hi,
                                     uint32_t* __restrict__ lo)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    if (hi) {
        int a = 0, b = numTiles;       // invariant: tile_starts[a] <= j < tile_starts[b]
        while (b - a > 1) {
            const int m = (a + b) >> 1;
            if (tile_starts[m] <= j) a = m; else b = m;
        }
        hi[j] = (uint32_t)a;
    }
    if (lo) lo[j] = __float_as_uint(depth_ptr[(size_t)list[j] * depth_stride]);
}

cudaError_t launch_expand_sorted_keys(cudaStream_t st, uint32_t M, int numTiles, const uint32_t* tile_starts, const uint32_t* list,
                                      const float* depth_ptr, int depth_stride, uint32_t* hi, uint32_t* lo)
{
    if (M > 0) k_expand_sorted_keys<<<cdiv(M, 256), 256, 0, st>>>(M, numTiles, tile_starts, list, depth_ptr, depth_stride, hi, lo);
    return cudaGetLastError();
}

}  // namespace gsb
