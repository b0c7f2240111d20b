// binning.cu — tile-intersection lists (K3..K8 of the reference) for sm_100a.
//
// Reference: slang/gaussian_tile_global_kernels.slang:17-404 and the driver
// Trainer/GaussianRenderer.swift:333-490 (two host syncs, ONE 128-thread threadgroup doing a 4-bit LSD
// sort of (tile, depth) keys over 12 passes, dense [numTiles,maxTilePairs] padding).
//
// The reference order is: ascending tile id, then ascending asuint(depth), ties in emission order
// (Gaussian index ascending).  It is produced here WITHOUT ever sorting 64-bit keys over the M pairs:
//
//   1. sort the N Gaussians by asuint(depth) (stable, 32-bit keys, payload = Gaussian index);
//   2. exclusive scan, in that order, of the SUPERBLOCKS (4 x 2 tiles) each rect touches (single pass, decoupled
//      look-back);
//   3. emit the (superblock id, Gaussian) pairs in that order — the pair list is then already sorted by
//      (depth, index) — with one thread per OUTPUT pair (binary search in shared memory), fully coalesced;
//   4. one STABLE radix sort of those ~3.4 M pairs on the superblock id alone (10 bits at 1080p: 2 onesweep passes
//      over 8-byte pairs; the reference sorts 12 M twelve-byte triples in 12 passes);
//   5. tilelists.cu turns every superblock's depth-ordered slice into its 8 tile lists (ballot-ranked count / fill).
//
// Nothing synchronises with the host: every kernel reads its element count from device memory and is launched for
// the buffer capacity.  All of this is HBM/latency-bound integer work; the figure of merit is bytes moved per pair.
// The hand-written onesweep (8-bit digits, chained-scan look-back with relaxed status words, stable) is also exposed
// stand-alone on 64-bit keys (gsb_sort_tile_keys = the reference's K5), with cub::DeviceRadixSort as the checked
// baseline.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

#include "kernels.h"

namespace gsb {

// ------------------------------------------------------------------------------------------------
// K3 on reference-layout inputs (parity API): tile rect, count and the depth sort key
// ------------------------------------------------------------------------------------------------
__global__ void k_count_tiles(int N, const __grid_constant__ ViewParams vp, const float* __restrict__ rectMin,
                              const float* __restrict__ rectMax, const float* __restrict__ radii,
                              const float* __restrict__ depths, uint2* __restrict__ tile_rects,
                              uint32_t* __restrict__ touched, uint32_t* __restrict__ depth_keys)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    uint32_t cnt = 0;
    if (radii[i] > 0.0f) {
        // slang/gaussian_tile_global_kernels.slang:39-57
        int tMinX = (int)floorf(rectMin[i * 2 + 0] / (float)vp.tileW);
        int tMinY = (int)floorf(rectMin[i * 2 + 1] / (float)vp.tileH);
        int tMaxX = (int)floorf(rectMax[i * 2 + 0] / (float)vp.tileW) + 1;
        int tMaxY = (int)floorf(rectMax[i * 2 + 1] / (float)vp.tileH) + 1;
        x0 = max(0, min(tMinX, vp.gridW));
        y0 = max(0, min(tMinY, vp.gridH));
        x1 = max(0, min(tMaxX, vp.gridW));
        y1 = max(0, min(tMaxY, vp.gridH));
        cnt = (uint32_t)((x1 - x0) * (y1 - y0));
    }
    if (cnt == 0) { x0 = y0 = x1 = y1 = 0; }
    tile_rects[i] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
    touched[i] = cnt;
    // Gaussians without tiles emit no pairs, so their place in the depth order is irrelevant: they keep a key in the
    // range of the others (|depth| clamped to the near plane) instead of 0xffffffff, which would make the top digit
    // place non-trivial and cost the depth sort a fourth pass
    depth_keys[i] = __float_as_uint(cnt ? depths[i] : fminf(fmaxf(fabsf(depths[i]), 0.2f), 3.0e38f));
}

cudaError_t launch_count_tiles(cudaStream_t st, int N, const ViewParams& vp, const float* rectMin, const float* rectMax,
                               const float* radii, const float* depths, uint2* tile_rects, uint32_t* touched,
                               uint32_t* depth_keys)
{
    if (N > 0) k_count_tiles<<<cdiv(N, 256), 256, 0, st>>>(N, vp, rectMin, rectMax, radii, depths, tile_rects, touched, depth_keys);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// exclusive scan (replaces MLX cumsum + `.item()` host sync, GaussianRenderer.swift:398-409)
// single pass, decoupled look-back; 2048 items per CTA; optional gather through a permutation
// ------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
#ifndef GSB_SC_IPT
#define GSB_SC_IPT 8
#endif
constexpr int SC_IPT = GSB_SC_IPT;
constexpr int SC_TILE = SC_THREADS * SC_IPT;
constexpr uint64_t SC_FLAG_LOCAL = 1ull << 62;
constexpr uint64_t SC_FLAG_INCL = 2ull << 62;
constexpr uint64_t SC_VALUE_MASK = (1ull << 62) - 1;

size_t scan_ws_bytes(int N) { return 16 + (size_t)cdiv(N > 0 ? N : 1, SC_TILE) * sizeof(uint64_t); }

__device__ __forceinline__ uint64_t ld_acquire_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// SBCOUNT: the scanned value of element i is the number of SUPERBLOCKS the tile rect of Gaussian perm[i] touches, computed
// on the fly (rects = packed tile rects), and the tiles-touched of the same Gaussians are summed into *touched_total
// (the pair count M) - what used to be a kernel of its own (k_sb_counts) with a 4 MB round trip.
template <bool SBCOUNT>
__global__ void __launch_bounds__(SC_THREADS) k_exclusive_scan(int N, const uint32_t* __restrict__ in,
                                                               const uint32_t* __restrict__ perm,
                                                               const uint32_t* __restrict__ perm_sel,
                                                               const uint32_t* __restrict__ perm_alt,
                                                               uint32_t* __restrict__ out, uint32_t* __restrict__ total,
                                                               uint32_t* counter, uint64_t* status,
                                                               const uint2* __restrict__ rects, uint32_t* __restrict__ touched_total)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SC_THREADS / 32];
    __shared__ uint32_t s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    if (perm && perm_sel && *perm_sel) perm = perm_alt;   // which ping-pong buffer holds the sorted order
    const uint32_t tile = s_tile;
    const int base = tile * SC_TILE + threadIdx.x * SC_IPT;
    uint32_t v[SC_IPT];
    uint32_t sum = 0;
    if (SBCOUNT) {
        uint32_t g[SC_IPT];
#pragma unroll
        for (int i = 0; i < SC_IPT; ++i) g[i] = (base + i < N) ? perm[base + i] : 0xffffffffu;
        uint32_t tsum = 0;
#pragma unroll
        for (int i = 0; i < SC_IPT; ++i) {
            v[i] = 0u;
            if (g[i] != 0xffffffffu) {
                const uint2 r = rects[g[i]];
                const int x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff, y1 = r.y >> 16;
                if (x1 > x0 && y1 > y0) v[i] = (uint32_t)((((x1 - 1) / SBW) - (x0 / SBW) + 1) * (((y1 - 1) / SBH) - (y0 / SBH) + 1));
                tsum += in[g[i]];   // in = tiles-touched per Gaussian
            }
            sum += v[i];
        }
        tsum = __reduce_add_sync(0xffffffffu, tsum);
        if ((threadIdx.x & 31) == 0 && tsum) atomicAdd(touched_total, tsum);
    } else {
#pragma unroll
        for (int i = 0; i < SC_IPT; ++i) {
            v[i] = (base + i < N) ? (perm ? in[perm[base + i]] : in[base + i]) : 0u;
            sum += v[i];
        }
    }
    // block inclusive scan of per-thread sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SC_THREADS / 32 ? s_warp[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < SC_THREADS / 32) s_warp[lane] = winc - w;  // exclusive warp prefix
        uint32_t block_total = __shfl_sync(0xffffffffu, winc, SC_THREADS / 32 - 1);
        // Decoupled look-back, one warp wide: lane i inspects predecessor t - i, so a chain of LOCAL aggregates is
        // consumed 32 tiles per L2 round trip (a one-thread walk made the scan of 489 tiles a ~50-hop serial chain:
        // all tiles publish LOCAL at the same time and the INCLUSIVE frontier meets the walkers half way).
        uint64_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(&status[0], SC_FLAG_INCL | block_total);
        } else {
            if (lane == 0) st_relaxed_u64(&status[tile], SC_FLAG_LOCAL | block_total);
            int t = (int)tile - 1;
            while (true) {
                const int tt = t - lane;
                const uint64_t s = tt >= 0 ? ld_relaxed_u64(&status[tt]) : SC_FLAG_INCL;   // virtual tile -1: inclusive 0
                const uint32_t flag = (uint32_t)(s >> 62);
                const uint32_t not_ready = __ballot_sync(0xffffffffu, flag == 0);
                const uint32_t incl = __ballot_sync(0xffffffffu, flag == 2);
                // lanes [0, stop) can be consumed: up to and including the nearest INCLUSIVE, or up to the first tile
                // that has not published yet
                const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
                const int first_in = incl ? __ffs(incl) - 1 : 32;
                const bool done = first_in < first_nr;
                const int stop = done ? first_in + 1 : first_nr;
                uint64_t part = lane < stop ? (s & SC_VALUE_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                excl += part;
                if (done) break;
                t -= stop;
            }
            if (lane == 0) st_relaxed_u64(&status[tile], SC_FLAG_INCL | (excl + block_total));
        }
        if (lane == 0) {
            s_prefix = (uint32_t)excl;
            if ((long long)(tile + 1) * SC_TILE >= N) *total = (uint32_t)excl + block_total;
        }
    }
    __syncthreads();
    uint32_t run = s_prefix + s_warp[warp] + (inc - sum);
#pragma unroll
    for (int i = 0; i < SC_IPT; ++i) {
        if (base + i < N) out[base + i] = run;
        run += v[i];
    }
}

cudaError_t launch_exclusive_scan(cudaStream_t st, int N, const uint32_t* in, const uint32_t* perm0, const uint32_t* perm1,
                                  const uint32_t* perm_sel, uint32_t* offsets, uint32_t* total, void* scan_ws)
{
    if (N <= 0) return cudaMemsetAsync(total, 0, sizeof(uint32_t), st);
    cudaError_t e = cudaMemsetAsync(scan_ws, 0, scan_ws_bytes(N), st);
    if (e != cudaSuccess) return e;
    uint32_t* counter = reinterpret_cast<uint32_t*>(scan_ws);
    uint64_t* status = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(scan_ws) + 16);
    k_exclusive_scan<false><<<cdiv(N, SC_TILE), SC_THREADS, 0, st>>>(N, in, perm0, perm_sel, perm1, offsets, total, counter, status, nullptr, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_sb_scan(cudaStream_t st, int N, const uint2* tile_rects, const uint32_t* touched, const uint32_t* perm0,
                           const uint32_t* perm1, const uint32_t* perm_sel, uint32_t* offsets, uint32_t* total_sb_pairs,
                           uint32_t* total_pairs, void* scan_ws)
{
    cudaError_t e = cudaMemsetAsync(total_pairs, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (N <= 0) return cudaMemsetAsync(total_sb_pairs, 0, sizeof(uint32_t), st);
    e = cudaMemsetAsync(scan_ws, 0, scan_ws_bytes(N), st);
    if (e != cudaSuccess) return e;
    uint32_t* counter = reinterpret_cast<uint32_t*>(scan_ws);
    uint64_t* status = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(scan_ws) + 16);
    k_exclusive_scan<true><<<cdiv(N, SC_TILE), SC_THREADS, 0, st>>>(N, touched, perm0, perm_sel, perm1, offsets, total_sb_pairs, counter, status,
                                                                      tile_rects, total_pairs);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K4 generate_keys (slang/gaussian_tile_global_kernels.slang:73-126).
// (a) reference emission order, 64-bit keys: only for the unsorted lists of the parity API;
// (b) depth order, tile-id keys, one thread per OUTPUT pair: the production path.
// ------------------------------------------------------------------------------------------------
__global__ void k_generate_keys_ref(int N, const __grid_constant__ ViewParams vp, const uint2* __restrict__ tile_rects,
                                    const uint32_t* __restrict__ offsets, const float* __restrict__ depth_ptr,
                                    int depth_stride, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                    uint32_t capacity)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint2 r = tile_rects[i];
    int x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff, y1 = r.y >> 16;
    if (x1 <= x0 || y1 <= y0) return;
    uint32_t off = offsets[i];
    uint32_t cnt = (uint32_t)((x1 - x0) * (y1 - y0));
    if (off + cnt > capacity) return;
    uint32_t depthBits = __float_as_uint(depth_ptr[(size_t)i * depth_stride]);
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) {
            uint32_t tile = (uint32_t)(ty * vp.gridW + tx);
            keys[off] = ((uint64_t)tile << 32) | depthBits;
            vals[off] = (uint32_t)i;
            ++off;
        }
}

cudaError_t launch_generate_keys_ref(cudaStream_t st, int N, const ViewParams& vp, const uint2* tile_rects,
                                     const uint32_t* offsets, const float* depth_ptr, int depth_stride, uint64_t* keys,
                                     uint32_t* vals, uint32_t capacity)
{
    if (N > 0)
        k_generate_keys_ref<<<cdiv(N, 256), 256, 0, st>>>(N, vp, tile_rects, offsets, depth_ptr, depth_stride, keys, vals, capacity);
    return cudaGetLastError();
}

constexpr int KG_THREADS = 256;

__global__ void __launch_bounds__(KG_THREADS) k_generate_keys(int N, int gridW, int cw, int ch, const uint2* __restrict__ tile_rects,
                                                              const uint32_t* __restrict__ offsets,
                                                              const uint32_t* __restrict__ perm0,
                                                              const uint32_t* __restrict__ perm1,
                                                              const uint32_t* __restrict__ perm_sel,
                                                              uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                              uint32_t capacity, const uint32_t* __restrict__ total,
                                                              uint32_t* overflow_flag)
{
    __shared__ uint32_t s_off[KG_THREADS + 1];
    __shared__ uint2 s_rect[KG_THREADS];
    __shared__ uint32_t s_g[KG_THREADS];
    const uint32_t* perm = (perm_sel && *perm_sel) ? perm1 : perm0;
    const int i0 = blockIdx.x * KG_THREADS;
    const int i = i0 + threadIdx.x;
    const uint32_t M = *total;
    if (i == 0 && M > capacity) *overflow_flag = 1u;
    if (i < N) {
        const uint32_t g = perm[i];
        s_g[threadIdx.x] = g;
        uint2 r = tile_rects[g];
        if (cw != 1 || ch != 1) {   // tile rect -> rect of cw x ch cells
            const uint32_t x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff, y1 = r.y >> 16;
            if (x1 > x0 && y1 > y0)
                r = make_uint2((x0 / cw) | ((y0 / ch) << 16), ((x1 - 1) / cw + 1) | (((y1 - 1) / ch + 1) << 16));
        }
        s_rect[threadIdx.x] = r;
        s_off[threadIdx.x] = offsets[i];
    } else {
        s_off[threadIdx.x] = M;
        s_rect[threadIdx.x] = make_uint2(0u, 0u);
        s_g[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) s_off[KG_THREADS] = (i0 + KG_THREADS < N) ? offsets[i0 + KG_THREADS] : M;
    __syncthreads();
    const uint32_t begin = s_off[0];
    const uint32_t end = min(s_off[KG_THREADS], capacity);
    for (uint32_t j = begin + threadIdx.x; j < end; j += KG_THREADS) {
        // owner of output slot j = the LAST index k with s_off[k] <= j (indices on a plateau before it own nothing)
        int lo = 0, hi = KG_THREADS;   // invariant: s_off[lo] <= j < s_off[hi]
#pragma unroll
        for (int step = 0; step < 8; ++step) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= j) lo = mid; else hi = mid;
        }
        const uint2 r = s_rect[lo];
        const uint32_t x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff;
        const uint32_t w = x1 - x0;
        const uint32_t rel = j - s_off[lo];
        const uint32_t dy = rel / w;
        const uint32_t dx = rel - dy * w;
        keys[j] = (y0 + dy) * (uint32_t)gridW + (x0 + dx);
        vals[j] = s_g[lo];
    }
}

cudaError_t launch_generate_keys(cudaStream_t st, int N, int cellGridW, int cw, int ch, const uint2* tile_rects,
                                 const uint32_t* offsets, const uint32_t* perm0, const uint32_t* perm1,
                                 const uint32_t* perm_sel, uint32_t* keys, uint32_t* vals, uint32_t capacity,
                                 const uint32_t* total, uint32_t* overflow_flag)
{
    if (N > 0)
        k_generate_keys<<<cdiv(N, KG_THREADS), KG_THREADS, 0, st>>>(N, cellGridW, cw, ch, tile_rects, offsets, perm0, perm1, perm_sel,
                                                                    keys, vals, capacity, total, overflow_flag);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// ONESWEEP radix sort: 8-bit digits, one upfront histogram of every digit place, then one
// chained-scan (decoupled look-back) + scatter kernel per digit place.  Stable.  KeyT = u32 | u64.
// Replaces radix_sort_tile_keys_fused_forward (one 128-thread threadgroup, 4-bit digits,
// slang/gaussian_tile_global_kernels.slang:143-305).
// Digit places whose histogram has a single occupied bin are skipped (device-side decision).
// iota != 0 means "payload = element index", synthesised by the first executed pass.
// ------------------------------------------------------------------------------------------------
constexpr int OS_THREADS = 256;
constexpr int OS_WARPS = OS_THREADS / 32;
#ifndef GSB_OS_HIST_MULT
#define GSB_OS_HIST_MULT 2
#endif
#ifndef GSB_OS_IPT
#define GSB_OS_IPT 16
#endif
constexpr int OS_IPT = GSB_OS_IPT;
constexpr int OS_TILE = OS_THREADS * OS_IPT;  // 4096 pairs per CTA
constexpr int OS_RADIX = 256;
constexpr int OS_MAX_PASSES = 8;
#ifndef GSB_OS_LB
#define GSB_OS_LB 8
#endif
constexpr int OS_LB = GSB_OS_LB;   // look-back: predecessors fetched per round trip
constexpr uint32_t OS_FLAG_LOCAL = 1u << 30;
constexpr uint32_t OS_FLAG_INCL = 2u << 30;
constexpr uint32_t OS_VALUE_MASK = (1u << 30) - 1;

__host__ __device__ inline uint32_t os_digit_mask(int end_bit, int pass)
{
    int bits = end_bit - 8 * pass;
    return bits >= 8 ? 255u : ((1u << (bits > 0 ? bits : 0)) - 1u);
}

struct SortCtl {
    uint32_t tile_counter[OS_MAX_PASSES];
    uint32_t skip[OS_MAX_PASSES];
    uint32_t src_buf[OS_MAX_PASSES];
    uint32_t result_buf;
    uint32_t count;
    uint32_t num_tiles;
    uint32_t first_pass;   // first pass that is not skipped (OS_MAX_PASSES when all are)
    uint32_t pad[4];
};
static_assert(sizeof(SortCtl) == 128, "SortCtl layout");

uint32_t sort_tile_items() { return (uint32_t)OS_TILE; }

SortPlan sort_plan(uint32_t capacity, uint32_t end_bit)
{
    SortPlan p;
    p.capacity = capacity;
    p.max_tiles = (uint32_t)cdiv(capacity > 0 ? capacity : 1, OS_TILE);
    p.end_bit = end_bit;
    p.passes = (end_bit + 7) / 8;
    if (p.passes < 1) p.passes = 1;
    if (p.passes > OS_MAX_PASSES) p.passes = OS_MAX_PASSES;
    p.ws_bytes = sizeof(SortCtl) + (size_t)OS_MAX_PASSES * OS_RADIX * 4 + (size_t)p.passes * p.max_tiles * OS_RADIX * 4;
    return p;
}

template <typename KeyT>
__global__ void __launch_bounds__(256) k_os_histogram(const KeyT* __restrict__ keys, const uint32_t* __restrict__ d_count,
                                                      uint32_t capacity, int passes, int end_bit, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t sh[OS_MAX_PASSES * OS_RADIX];
    for (int i = threadIdx.x; i < passes * OS_RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t count = min(*d_count, capacity);
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += stride) {
        uint32_t idx = base + lane;
        bool valid = idx < count;
        KeyT key = valid ? keys[idx] : (KeyT)0;
        uint32_t m = __ballot_sync(0xffffffffu, valid);
        int leader = __ffs(m) - 1;
        for (int p = 0; p < passes; ++p) {
            uint32_t d = (uint32_t)(key >> (8 * p)) & os_digit_mask(end_bit, p);
            uint32_t d0 = __shfl_sync(0xffffffffu, d, leader);
            bool uni = __all_sync(0xffffffffu, !valid || d == d0);
            if (uni) {
                if (lane == leader) atomicAdd(&sh[p * OS_RADIX + d], (uint32_t)__popc(m));
            } else if (valid) {
                atomicAdd(&sh[p * OS_RADIX + d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * OS_RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one CTA: per digit place, exclusive scan of the 256 bins, skip decision, buffer ping-pong plan
__global__ void __launch_bounds__(OS_RADIX) k_os_scan(SortCtl* ctl, uint32_t* hist, const uint32_t* __restrict__ d_count,
                                                      uint32_t capacity, int passes)
{
    __shared__ uint32_t s_warp[OS_RADIX / 32];
    __shared__ uint32_t s_skip[OS_MAX_PASSES];
    const uint32_t count = min(*d_count, capacity);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = 0; p < passes; ++p) {
        uint32_t c = hist[p * OS_RADIX + threadIdx.x];
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        int trivial = __syncthreads_or(count > 0 && c == count);
        uint32_t wprefix = 0;
        for (int w = 0; w < warp; ++w) wprefix += s_warp[w];
        hist[p * OS_RADIX + threadIdx.x] = wprefix + inc - c;
        if (threadIdx.x == 0) s_skip[p] = (trivial || count <= 1) ? 1u : 0u;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t cur = 0, first = OS_MAX_PASSES;
        for (int p = 0; p < passes; ++p) {
            ctl->skip[p] = s_skip[p];
            ctl->src_buf[p] = cur;
            if (!s_skip[p]) {
                if (first == OS_MAX_PASSES) first = (uint32_t)p;
                cur ^= 1u;
            }
        }
        ctl->first_pass = first;
        ctl->result_buf = cur;
        ctl->count = count;
        ctl->num_tiles = (count + OS_TILE - 1) / OS_TILE;
    }
}

template <typename KeyT>
struct OsSmem {
    KeyT keys[OS_TILE];
    uint32_t vals[OS_TILE];
    uint32_t warp_hist[OS_WARPS][OS_RADIX];
    uint32_t block_excl[OS_RADIX];
    uint32_t digit_base[OS_RADIX];
    uint32_t warp_tot[OS_RADIX / 32];
    uint32_t tile;
};

template <typename KeyT>
__global__ void __launch_bounds__(OS_THREADS, 4) k_os_pass(int pass, uint32_t dmask, SortCtl* ctl, const uint32_t* __restrict__ hist_excl,
                                                        uint32_t* lookback, uint32_t max_tiles, KeyT* keys0, KeyT* keys1,
                                                        uint32_t* vals0, uint32_t* vals1, int iota)
{
    extern __shared__ __align__(16) unsigned char os_raw[];
    OsSmem<KeyT>& S = *reinterpret_cast<OsSmem<KeyT>*>(os_raw);
    if (ctl->skip[pass]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) S.tile = atomicAdd(&ctl->tile_counter[pass], 1u);
    for (int i = tid; i < OS_WARPS * OS_RADIX; i += OS_THREADS) (&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    const uint32_t count = ctl->count;
    if (tile >= ctl->num_tiles) return;
    const uint32_t src = ctl->src_buf[pass];
    const bool synth = iota && (uint32_t)pass == ctl->first_pass;   // payload = element index
    const KeyT* kin = src ? keys1 : keys0;
    const uint32_t* vin = src ? vals1 : vals0;
    KeyT* kout = src ? keys0 : keys1;
    uint32_t* vout = src ? vals0 : vals1;
    const uint32_t tile_base = tile * OS_TILE;
    const uint32_t valid = min((uint32_t)OS_TILE, count - tile_base);
    const int shift = pass * 8;

    // warp-striped load: stability order is (warp, item, lane)
    KeyT key[OS_IPT];
    const uint32_t wbase = warp * (OS_IPT * 32);
#pragma unroll
    for (int i = 0; i < OS_IPT; ++i) {
        uint32_t idx = wbase + i * 32 + lane;
        key[i] = idx < valid ? kin[tile_base + idx] : ~(KeyT)0;
    }
    // payloads are fetched now as well, so their latency hides behind the ranking instead of stalling the scatter
    uint32_t val[OS_IPT];
#pragma unroll
    for (int i = 0; i < OS_IPT; ++i) {
        uint32_t idx = wbase + i * 32 + lane;
        val[i] = idx < valid ? (synth ? tile_base + idx : vin[tile_base + idx]) : 0u;
    }
    // per-warp stable ranking with match.any
    uint16_t rank[OS_IPT];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < OS_IPT; ++i) {
        uint32_t d = (uint32_t)(key[i] >> shift) & dmask;
        uint32_t m = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(m) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = S.warp_hist[warp][d];
            S.warp_hist[warp][d] = old + __popc(m);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = (uint16_t)(old + __popc(m & lt_mask));
        __syncwarp();
    }
    __syncthreads();
    // per digit: prefix over warps, publish, look back, block-exclusive scan
    {
        const int d = tid;  // OS_THREADS == OS_RADIX
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) {
            uint32_t c = S.warp_hist[w][d];
            S.warp_hist[w][d] = run;
            run += c;
        }
        uint32_t total_valid = run;
        if ((uint32_t)d == dmask) total_valid -= (OS_TILE - valid);
        uint32_t* lb = lookback + ((size_t)pass * max_tiles + tile) * OS_RADIX + d;
        uint32_t excl_prev = 0;
        if (tile == 0) {
            st_relaxed_u32(lb, OS_FLAG_INCL | total_valid);
        } else {
            st_relaxed_u32(lb, OS_FLAG_LOCAL | total_valid);
            // Decoupled look-back with OS_LB predecessors in flight per round trip: every tile publishes LOCAL at about
            // the same time, so a one-at-a-time walk is a serial chain of ~tiles/2 L2 latencies (it was most of a
            // 1 M-key pass); the status rows of a tile are 1 KB apart, one coalesced row per predecessor.
            int t = (int)tile - 1;
            const uint32_t* row = lookback + (size_t)pass * max_tiles * OS_RADIX + d;
            for (bool done = false; !done;) {
                uint32_t v[OS_LB];
#pragma unroll
                for (int k = 0; k < OS_LB; ++k)
                    v[k] = t - k >= 0 ? ld_relaxed_u32(row + (size_t)(t - k) * OS_RADIX) : OS_FLAG_INCL;   // virtual tile -1
                int consumed = 0, state = 0;   // 0 consuming, 1 blocked on a tile that has not published, 2 done
#pragma unroll
                for (int k = 0; k < OS_LB; ++k) {
                    const uint32_t f = v[k] >> 30;
                    if (state == 0) {
                        if (f == 0) {
                            state = 1;
                        } else {
                            excl_prev += v[k] & OS_VALUE_MASK;
                            ++consumed;
                            if (f == 2) state = 2;
                        }
                    }
                }
                t -= consumed;
                done = state == 2;
            }
            st_relaxed_u32(lb, OS_FLAG_INCL | (excl_prev + total_valid));
        }
        // block-exclusive scan of `run` over the 256 digits
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) S.warp_tot[warp] = inc;
        __syncthreads();
        uint32_t wprefix = 0;
        for (int w = 0; w < warp; ++w) wprefix += S.warp_tot[w];
        uint32_t bex = wprefix + inc - run;
        S.block_excl[d] = bex;
        S.digit_base[d] = hist_excl[pass * OS_RADIX + d] + excl_prev - bex;
    }
    __syncthreads();
    // scatter into block-sorted order in shared memory
#pragma unroll
    for (int i = 0; i < OS_IPT; ++i) {
        uint32_t d = (uint32_t)(key[i] >> shift) & dmask;
        uint32_t pos = S.block_excl[d] + S.warp_hist[warp][d] + rank[i];
        S.keys[pos] = key[i];
        S.vals[pos] = val[i];
    }
    __syncthreads();
    // coalesced runs out
    for (uint32_t idx = tid; idx < valid; idx += OS_THREADS) {
        KeyT k = S.keys[idx];
        uint32_t d = (uint32_t)(k >> shift) & dmask;
        uint32_t dst = S.digit_base[d] + idx;
        kout[dst] = k;
        vout[dst] = S.vals[idx];
    }
}

// when every pass was skipped and the payload is synthetic, the "sorted" payload is the identity
__global__ void k_os_iota_if_unsorted(const SortCtl* ctl, uint32_t* vals0)
{
    if (ctl->first_pass != OS_MAX_PASSES) return;
    const uint32_t count = ctl->count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) vals0[i] = i;
}

template <typename KeyT>
static cudaError_t onesweep_impl(cudaStream_t st, const SortPlan& plan, KeyT* keys0, KeyT* keys1, uint32_t* vals0, uint32_t* vals1,
                                 int iota, const uint32_t* d_count, void* ws, const uint32_t** result_buf_ptr, int* launches)
{
    SortCtl* ctl = reinterpret_cast<SortCtl*>(ws);
    uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + sizeof(SortCtl));
    uint32_t* lookback = hist + OS_MAX_PASSES * OS_RADIX;
    if (result_buf_ptr) *result_buf_ptr = &ctl->result_buf;
    cudaError_t e = cudaMemsetAsync(ws, 0, plan.ws_bytes, st);
    if (e != cudaSuccess) return e;
    // the histogram is latency-bound (load -> shared atomics per key): a few keys per thread, up to one full wave of CTAs
    int hist_blocks = (int)std::min<uint32_t>(plan.max_tiles * (uint32_t)GSB_OS_HIST_MULT, 148u * 8u);
    k_os_histogram<KeyT><<<hist_blocks, 256, 0, st>>>(keys0, d_count, plan.capacity, (int)plan.passes, (int)plan.end_bit, hist);
    k_os_scan<<<1, OS_RADIX, 0, st>>>(ctl, hist, d_count, plan.capacity, (int)plan.passes);
    static bool attr_set = false;
    if (!attr_set) {
        e = cudaFuncSetAttribute(k_os_pass<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OsSmem<KeyT>));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    for (uint32_t p = 0; p < plan.passes; ++p)
        k_os_pass<KeyT><<<plan.max_tiles, OS_THREADS, sizeof(OsSmem<KeyT>), st>>>((int)p, os_digit_mask((int)plan.end_bit, (int)p), ctl,
                                                                                    hist, lookback, plan.max_tiles, keys0, keys1, vals0,
                                                                                    vals1, iota);
    int n = 2 + (int)plan.passes;
    if (iota) {
        k_os_iota_if_unsorted<<<std::max(1, std::min((int)plan.max_tiles, 148 * 4)), 256, 0, st>>>(ctl, vals0);
        ++n;
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

cudaError_t launch_onesweep_sort(cudaStream_t st, const SortPlan& plan, uint64_t* keys0, uint64_t* keys1, uint32_t* vals0,
                                 uint32_t* vals1, const uint32_t* d_count, void* ws, const uint32_t** result_buf_ptr,
                                 int* launches)
{
    return onesweep_impl<uint64_t>(st, plan, keys0, keys1, vals0, vals1, 0, d_count, ws, result_buf_ptr, launches);
}

cudaError_t launch_onesweep_sort32(cudaStream_t st, const SortPlan& plan, uint32_t* keys0, uint32_t* keys1, uint32_t* vals0,
                                   uint32_t* vals1, int iota, const uint32_t* d_count, void* ws,
                                   const uint32_t** result_buf_ptr, int* launches)
{
    return onesweep_impl<uint32_t>(st, plan, keys0, keys1, vals0, vals1, iota, d_count, ws, result_buf_ptr, launches);
}

cudaError_t cub_sort_pairs(cudaStream_t st, uint64_t* keys0, uint64_t* keys1, uint32_t* vals0, uint32_t* vals1,
                           uint32_t count, uint32_t end_bit, void* tmp, size_t tmp_bytes, size_t* tmp_needed)
{
    size_t need = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, keys0, keys1, vals0, vals1, (int)count, 0, (int)end_bit, st);
    if (e != cudaSuccess) return e;
    if (tmp_needed) *tmp_needed = need;
    if (!tmp || tmp_bytes < need) return cudaSuccess;  // size query only
    return cub::DeviceRadixSort::SortPairs(tmp, need, keys0, keys1, vals0, vals1, (int)count, 0, (int)end_bit, st);
}

cudaError_t cub_sort_pairs32(cudaStream_t st, uint32_t* keys0, uint32_t* keys1, uint32_t* vals0, uint32_t* vals1,
                             uint32_t count, uint32_t end_bit, void* tmp, size_t tmp_bytes, size_t* tmp_needed)
{
    size_t need = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, keys0, keys1, vals0, vals1, (int)count, 0, (int)end_bit, st);
    if (e != cudaSuccess) return e;
    if (tmp_needed) *tmp_needed = need;
    if (!tmp || tmp_bytes < need) return cudaSuccess;
    return cub::DeviceRadixSort::SortPairs(tmp, need, keys0, keys1, vals0, vals1, (int)count, 0, (int)end_bit, st);
}

__global__ void k_iota(uint32_t n, uint32_t* v)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
cudaError_t launch_iota(cudaStream_t st, uint32_t n, uint32_t* v)
{
    if (n > 0) k_iota<<<cdiv(n, 256), 256, 0, st>>>(n, v);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K6 compute_tile_ranges (slang/gaussian_tile_global_kernels.slang:314-344) on a sorted key list: boundary
// detection -> ranges[key] = (first, last + 1).  Used on the sorted superblock ids of the level-1 list.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_key_ranges(const uint32_t* __restrict__ keys0, const uint32_t* __restrict__ keys1,
                                                    const uint32_t* __restrict__ d_result_buf, const uint32_t* __restrict__ d_count,
                                                    uint32_t capacity, uint32_t* __restrict__ ranges)
{
    const uint32_t M = min(*d_count, capacity);
    const uint32_t* keys = (d_result_buf && *d_result_buf) ? keys1 : keys0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const uint32_t cur = keys[j];
        if (j == 0) {
            ranges[cur * 2 + 0] = 0;
        } else {
            const uint32_t prev = keys[j - 1];
            if (cur != prev) {
                ranges[prev * 2 + 1] = j;
                ranges[cur * 2 + 0] = j;
            }
        }
        if (j == M - 1) ranges[cur * 2 + 1] = M;
    }
}

cudaError_t launch_key_ranges(cudaStream_t st, const uint32_t* keys0, const uint32_t* keys1, const uint32_t* d_result_buf,
                              const uint32_t* d_count, uint32_t capacity, uint32_t* ranges, int numKeys)
{
    cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)numKeys * 2 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (capacity == 0) return cudaSuccess;
    const int blocks = (int)std::min<uint64_t>(((uint64_t)capacity + 255) / 256, 148ull * 16ull);
    k_key_ranges<<<blocks, 256, 0, st>>>(keys0, keys1, d_result_buf, d_count, capacity, ranges);
    return cudaGetLastError();
}

__global__ void k_tile_counts(int numTiles, const uint32_t* __restrict__ ranges, uint32_t* __restrict__ counts)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numTiles) return;
    uint32_t s = ranges[t * 2], e = ranges[t * 2 + 1];
    counts[t] = e > s ? e - s : 0u;
}

cudaError_t launch_tile_counts(cudaStream_t st, int numTiles, const uint32_t* tile_ranges, uint32_t* tile_counts)
{
    if (numTiles > 0) k_tile_counts<<<cdiv(numTiles, 256), 256, 0, st>>>(numTiles, tile_ranges, tile_counts);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// layout helpers for the parity API
// ------------------------------------------------------------------------------------------------
__global__ void k_packed_to_rec(int N, const float* __restrict__ packed, float* __restrict__ rec)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float* s = packed + (size_t)p * 11;
    make_raster_record(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], s[9], s[10], (uint32_t)p,
                       reinterpret_cast<float4*>(rec) + (size_t)p * 3);
}
__global__ void k_rec_to_packed(int N, const float* __restrict__ rec, float* __restrict__ packed)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 11) return;
    int p = i / 11, c = i - p * 11;
    packed[i] = rec[(size_t)p * REC_FLOATS + c];
}
cudaError_t launch_packed_to_rec(cudaStream_t st, int N, const float* packed, float* rec)
{
    if (N > 0) k_packed_to_rec<<<cdiv(N, 256), 256, 0, st>>>(N, packed, rec);
    return cudaGetLastError();
}
cudaError_t launch_rec_to_packed(cudaStream_t st, int N, const float* rec, float* packed)
{
    if (N > 0) k_rec_to_packed<<<cdiv((long long)N * 11, 256), 256, 0, st>>>(N, rec, packed);
    return cudaGetLastError();
}

__global__ void k_split_keys(uint32_t M, const uint64_t* __restrict__ keys, uint32_t* __restrict__ hi, uint32_t* __restrict__ lo)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    uint64_t k = keys[i];
    if (hi) hi[i] = (uint32_t)(k >> 32);
    if (lo) lo[i] = (uint32_t)k;
}
__global__ void k_merge_keys(uint32_t M, const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo, uint32_t hi_mask,
                             uint64_t* __restrict__ keys)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    keys[i] = ((uint64_t)(hi[i] & hi_mask) << 32) | lo[i];
}
cudaError_t launch_split_keys(cudaStream_t st, uint32_t M, const uint64_t* keys, uint32_t* hi, uint32_t* lo)
{
    if (M > 0) k_split_keys<<<cdiv(M, 256), 256, 0, st>>>(M, keys, hi, lo);
    return cudaGetLastError();
}
cudaError_t launch_merge_keys(cudaStream_t st, uint32_t M, const uint32_t* hi, const uint32_t* lo, uint32_t hi_mask,
                              uint64_t* keys)
{
    if (M > 0) k_merge_keys<<<cdiv(M, 256), 256, 0, st>>>(M, hi, lo, hi_mask, keys);
    return cudaGetLastError();
}
}  // namespace gsb
