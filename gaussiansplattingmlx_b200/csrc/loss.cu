// loss.cu — separable SSIM forward/backward (K11/K12 of the reference) and the fused L1+SSIM loss.
//
// Reference: slang/ssim_kernels.slang:94-155 (forward, direct 11x11 taps), :181-266 (backward, gather
// over the 121 window centres), window Trainer/LossUtil.swift:47-54 + GaussianTrainer.swift:308-314
// (g[x] = exp(-(x-5.5)^2 / (2*1.5^2)), x = 0..10, normalised: the centre 11/2 = 5.5 makes it
// ASYMMETRIC), loss Trainer/GaussianTrainer.swift:688-716.
//
// The reference's 2-D window is the outer product g (x) g, so both passes separate into an
// 11-tap horizontal and an 11-tap vertical pass (zero padding = taps outside the image skipped).
// The backward is the transposed (flipped-window) convolution of three per-centre maps
//     A = dL/dmu1,  B = dL/dE[x^2] (= dL/dsigma1^2),  C = dL/dE[xy] (= dL/dsigma12)
//     grad1(p) = convT(A)(p) + 2*v1(p)*convT(B)(p) + v2(p)*convT(C)(p)
// which is exactly what the reference's gather (ssim_kernels.slang:214-262) sums.
// FP32-pipe work: ~225 flop (fwd) / ~170 flop (bwd) per pixel-channel; HWC f32 images.
#include <algorithm>

#include "kernels.h"

namespace gsb {

__device__ __forceinline__ f32x2 lds64_pair(uint32_t addr)
{
    f32x2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v.v) : "r"(addr));
    return v;
}

constexpr int SPAD = 5;           // window radius (K/2)
constexpr int SK = 11;
constexpr float SSIM_C1 = 0.0001f, SSIM_C2 = 0.0009f;

struct SsimWindow {
    float g[SK];
};

// f32 evaluation identical to the oracle's gso_ssim_window / LossUtil.swift:47-54
static SsimWindow make_window()
{
    SsimWindow w;
    const float center = (float)SK / 2.0f;
    float sum = 0.f;
    for (int x = 0; x < SK; ++x) {
        w.g[x] = expf(-powf((float)x - center, 2.0f) / (2.0f * powf(1.5f, 2.0f)));
        sum += w.g[x];
    }
    for (int x = 0; x < SK; ++x) w.g[x] = w.g[x] / sum;
    return w;
}

// Per-centre SSIM value and its partial derivatives (ssim_kernels.slang:70-92 and the AD thereof).
struct SsimPoint {
    float ssim, gm1, gs1, gs12;
};
__device__ __forceinline__ SsimPoint ssim_point(float mu1, float mu2, float e11, float e22, float e12)
{
    const float s1 = e11 - mu1 * mu1, s2 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
    const float A = 2.0f * mu1 * mu2 + SSIM_C1, B = 2.0f * s12 + SSIM_C2;
    const float Cc = mu1 * mu1 + mu2 * mu2 + SSIM_C1, D = s1 + s2 + SSIM_C2;
    const float CD = Cc * D;
    const float inv = 1.0f / CD;       // one IEEE division; the products below differ from (x / CD) by <= 1 ulp
    SsimPoint r;
    r.ssim = (A * B) * inv;
    const float gA = B * inv, gB = A * inv;
    const float gCD = -r.ssim * inv;
    const float gC = gCD * D, gD = gCD * Cc;
    r.gs1 = gD;
    r.gs12 = 2.0f * gB;
    r.gm1 = 2.0f * mu2 * gA + 2.0f * mu1 * gC - 2.0f * mu1 * gD - mu2 * r.gs12;
    return r;
}

// ------------------------------------------------------------------------------------------------
// Streaming separable convolution.  The HWC image is treated as H rows of W*C floats; a horizontal tap
// k of the per-channel window sits at element offset (k - 5) * C.  A CTA owns a strip of SW consecutive
// row elements and a chunk of `srows` output rows (chosen by the launcher so that the grid is one full wave) and marches down the rows.  Per input row:
//   * the strip (+ 5*C halo on both sides) is staged in shared memory, one barrier per row (two row
//     buffers alternate; row r+1 is fetched into registers while row r is filtered);
//   * each thread filters its element horizontally (11 taps: the only shared-memory reads, 22 per
//     output instead of the 77 of a shared-memory ring);
//   * the vertical pass lives entirely in REGISTERS: the filtered row is scattered into the 11 pending
//     output-row accumulators of the thread's column (acc[(phase + d) % 11] += g[10 - d] * h), the one that
//     just received its last tap is emitted and recycled.  The row loop is unrolled by 11 so every
//     accumulator index is a compile-time constant.
// MODE 0: parity API (writes ssim + optional saved maps)
// MODE 1: training (writes upstream-scaled A/B/C maps, accumulates sum|d| and sum(ssim))
// ------------------------------------------------------------------------------------------------
constexpr int SW = 128;        // strip width in floats = threads per CTA (1080p: 5760 = 45 strips)
constexpr int SROWS_MIN = 24;  // fewest output rows per CTA (10 halo rows are filtered on top of them)
constexpr int SMAXC = 4;       // channels supported by the halo buffer
constexpr int SHALO = SPAD * SMAXC;

// Stages row r of NIMG images: registers -> shared, barrier, prefetch of the next row into registers.
// Also prefetches, one row ahead, the two image values at the thread's own element of the OUTPUT row r - 5
// (the pointwise terms of the epilogue), so no global load is consumed in the iteration that issues it.
template <int NIMG>
struct RowStager {
    float pre[NIMG][2];
    float ctr_next[2], ctr[2];
    const float* img[NIMG];
    const float* cimg[2];
    int H, RW, e0, halo, nload, t, r0;
    // column validity and element offsets are row-invariant: set once (begin), advanced by one image row per fetch -
    // recomputing r * RW + e with its 64-bit selects cost ~100 of the ~260 instructions of a row
    bool colok[2], cok_col;
    long long off[2], coff;
    __device__ __forceinline__ void begin(int r)
    {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            const int ee = e0 - halo + i;
            colok[u] = i < nload && ee >= 0 && ee < RW;
            off[u] = (long long)r * RW + ee;
        }
        cok_col = e0 + t < RW;
        coff = (long long)(r - SPAD) * RW + e0 + t;
    }
    // rows must be fetched in order: begin(r), fetch(r), fetch(r + 1), ...
    __device__ __forceinline__ void fetch(int r)
    {
        const bool row_ok = r >= 0 && r < H;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool ok = row_ok && colok[u];
#pragma unroll
            for (int m = 0; m < NIMG; ++m) pre[m][u] = ok ? __ldg(img[m] + off[u]) : 0.f;
            off[u] += RW;
        }
        const int ro = r - SPAD;
        const bool cok = ro >= r0 && ro < H && cok_col;
        ctr_next[0] = cok ? __ldg(cimg[0] + coff) : 0.f;
        ctr_next[1] = cok ? __ldg(cimg[1] + coff) : 0.f;
        coff += RW;
    }
    __device__ __forceinline__ void store(float (*row)[SW + 2 * SHALO])
    {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            if (i < nload) {
#pragma unroll
                for (int m = 0; m < NIMG; ++m) row[m][i] = pre[m][u];
            }
        }
        ctr[0] = ctr_next[0];
        ctr[1] = ctr_next[1];
    }
    // Training forward (NIMG == 2): every staged element is stored ONCE as two packed pairs - (a, b) and the products
    // (a^2 + b^2, a b) - so that a horizontal tap is two 64-bit shared loads feeding two FFMA2 (the products were
    // recomputed by each of the 11 threads whose window covers the element: 8 instructions per tap instead of 4)
    __device__ __forceinline__ void store_pairs(float2* ab, float2* pp)
    {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            if (i < nload) {
                const float a = pre[0][u], b = pre[1][u];
                ab[i] = make_float2(a, b);
                pp[i] = make_float2(fmaf(b, b, a * a), a * b);
            }
        }
        ctr[0] = ctr_next[0];
        ctr[1] = ctr_next[1];
    }
    // Backward (NIMG == 3): maps A and B as one packed pair, C on its own
    __device__ __forceinline__ void store_ab_c(float2* ab, float* cc)
    {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            if (i < nload) {
                ab[i] = make_float2(pre[0][u], pre[1][u]);
                cc[i] = pre[2][u];
            }
        }
        ctr[0] = ctr_next[0];
        ctr[1] = ctr_next[1];
    }
};

// CT = compile-time channel count (tap offsets become immediates); CT == 0 uses the runtime C.
template <int MODE, int CT>
__global__ void __launch_bounds__(SW) k_ssim_fwd(int H, int W, int Crt, const float* __restrict__ img1,
                                                 const float* __restrict__ img2, const __grid_constant__ SsimWindow win,
                                                 float upstream, float* __restrict__ o0, float* __restrict__ o1,
                                                 float* __restrict__ o2, float* __restrict__ o3, float* __restrict__ o4,
                                                 float* __restrict__ o5, double* __restrict__ partial, int srows)
{
    __shared__ __align__(16) float s_row[MODE == 0 ? 4 : 1][SW + 2 * SHALO];   // MODE 0: [parity][image]
    __shared__ __align__(16) float2 s_ab[MODE == 1 ? 2 : 1][SW + 2 * SHALO];   // MODE 1: [parity] (a, b)
    __shared__ __align__(16) float2 s_pp[MODE == 1 ? 2 : 1][SW + 2 * SHALO];   //         [parity] (a^2 + b^2, a b)
    __shared__ double s_red[2][SW / 32];
    const int C = CT ? CT : Crt;
    const int RW = W * C;                          // floats per image row
    const int e0 = blockIdx.x * SW;                // first row element of the strip
    const int r0 = blockIdx.y * srows, r1 = min(r0 + srows, H);
    const int t = threadIdx.x;
    const int e = e0 + t;
    float accL1 = 0.f, accS = 0.f;               // <= srows terms each per thread: f32 is ample
    RowStager<2> st;
    st.img[0] = img1; st.img[1] = img2; st.cimg[0] = img1; st.cimg[1] = img2;
    st.H = H; st.RW = RW; st.e0 = e0; st.halo = SPAD * C; st.nload = SW + 2 * SPAD * C; st.t = t; st.r0 = r0;
    // MODE 1 needs sigma1^2 + sigma2^2 only as a SUM, so it filters E[a^2 + b^2] as one map (4 maps, not 5) - and
    // keeps the 4 maps as two packed f32x2 pairs (mu1, mu2), (E[a^2 + b^2], E[ab]): one FFMA2 per pair and tap, each
    // half bit-identical to the scalar fmaf (the kernel is issue-bound)
    constexpr int NQ = MODE == 0 ? 5 : 4;
    constexpr int NP = MODE == 0 ? 1 : 2;   // packed pairs (MODE 1 only)
    float acc[SK][MODE == 0 ? NQ : 1];
    f32x2 accp[SK][NP];
#pragma unroll
    for (int j = 0; j < SK; ++j) {
#pragma unroll
        for (int q = 0; q < (MODE == 0 ? NQ : 1); ++q) acc[j][q] = 0.f;
#pragma unroll
        for (int q = 0; q < NP; ++q) accp[j][q] = f2_bc(0.f);
    }

    const int rbeg = r0 - SPAD, rend = r1 + SPAD;   // input rows [rbeg, rend)
    long long oidx = (long long)(rbeg - SPAD) * RW + e;   // element index of output row r - 5, advanced by one row per iteration
    st.begin(rbeg);
    st.fetch(rbeg);
    for (int rb = rbeg; rb < rend; rb += SK) {
#pragma unroll
        for (int ph = 0; ph < SK; ++ph) {
            const int r = rb + ph;
            if (r < rend) {   // uniform across the CTA
                const int par = (r - rbeg) & 1;
                float (*row)[SW + 2 * SHALO] = reinterpret_cast<float (*)[SW + 2 * SHALO]>(&s_row[MODE == 0 ? par * 2 : 0][0]);
                if (MODE == 0) st.store(row);
                else st.store_pairs(&s_ab[par][0], &s_pp[par][0]);
                __syncthreads();   // the only barrier per row: the two row buffers alternate
                if (r + 1 < rend) st.fetch(r + 1);
                const uint32_t ab_addr = smem_u32(&s_ab[MODE == 1 ? par : 0][t]), pp_addr = smem_u32(&s_pp[MODE == 1 ? par : 0][t]);
                // ---- horizontal 11 taps of input row r
                float h[NQ];
                f32x2 hp[NP];
#pragma unroll
                for (int q = 0; q < NQ; ++q) h[q] = 0.f;
#pragma unroll
                for (int q = 0; q < NP; ++q) hp[q] = f2_bc(0.f);
#pragma unroll
                for (int k = 0; k < SK; ++k) {
                    const float w = win.g[k];
                    if (MODE == 0) {
                        const float a = row[0][t + k * C], b = row[1][t + k * C];
                        h[0] = fmaf(w, a, h[0]);
                        h[1] = fmaf(w, b, h[1]);
                        h[2] = fmaf(w, a * a, h[2]);
                        h[3] = fmaf(w, b * b, h[3]);
                        h[4] = fmaf(w, a * b, h[4]);
                    } else {
                        const f32x2 wb = f2_bc(w);
                        hp[0] = f2_fma(wb, lds64_pair(ab_addr + (uint32_t)(k * C) * 8u), hp[0]);   // (mu1, mu2)
                        hp[1] = f2_fma(wb, lds64_pair(pp_addr + (uint32_t)(k * C) * 8u), hp[1]);   // (E[a^2 + b^2], E[ab])
                    }
                }
                // ---- vertical: input row r feeds output rows r-5 .. r+5; output ro lives in slot (ro - rbeg) % 11
                //      = (ph + d - 5 + 11) % 11 for ro = r + d - 5, with tap k = r - ro + 5 = 10 - d
#pragma unroll
                for (int d = 0; d < SK; ++d) {
                    const int slot = (ph + d + SK - SPAD) % SK;
                    const float w = win.g[SK - 1 - d];
                    if (MODE == 0) {
#pragma unroll
                        for (int q = 0; q < NQ; ++q) acc[slot][q] = fmaf(w, h[q], acc[slot][q]);
                    } else {
#pragma unroll
                        for (int q = 0; q < NP; ++q) accp[slot][q] = f2_fma(f2_bc(w), hp[q], accp[slot][q]);
                    }
                }
                // ---- output row ro = r - 5 just received its last tap (d = 0)
                const int ro = r - SPAD;
                const int oslot = (ph + SK - SPAD) % SK;
                if (ro >= r0 && e < RW) {
                    const float m1 = MODE == 0 ? acc[oslot][0] : f2_lo(accp[oslot][0]);
                    const float m2 = MODE == 0 ? acc[oslot][MODE == 0 ? 1 : 0] : f2_hi(accp[oslot][0]);
                    const float e11 = MODE == 0 ? acc[oslot][MODE == 0 ? 2 : 0] : f2_lo(accp[oslot][NP - 1]);
                    const float e22 = MODE == 0 ? acc[oslot][MODE == 0 ? 3 : 0] : 0.f;
                    const float e12 = MODE == 0 ? acc[oslot][MODE == 0 ? 4 : 0] : f2_hi(accp[oslot][NP - 1]);
                    // (MODE 1: e11 carries E[a^2 + b^2], e22 = 0: ssim_point only uses their sum)
                    const SsimPoint sp = ssim_point(m1, m2, e11, e22, e12);
                    const long long idx = oidx;
                    if (MODE == 0) {
                        o0[idx] = sp.ssim;
                        if (o1) o1[idx] = m1;
                        if (o2) o2[idx] = m2;
                        if (o3) o3[idx] = e11 - m1 * m1;
                        if (o4) o4[idx] = e22 - m2 * m2;
                        if (o5) o5[idx] = e12 - m1 * m2;
                    } else {
                        o0[idx] = upstream * sp.gm1;
                        o1[idx] = upstream * sp.gs1;
                        o2[idx] = upstream * sp.gs12;
                        accS += sp.ssim;
                        accL1 += fabsf(st.ctr[0] - st.ctr[1]);
                    }
                }
#pragma unroll
                for (int q = 0; q < (MODE == 0 ? NQ : 1); ++q) acc[oslot][q] = 0.f;
#pragma unroll
                for (int q = 0; q < NP; ++q) accp[oslot][q] = f2_bc(0.f);
                oidx += RW;
            }
        }
    }
    if (MODE == 1) {
        double dL1 = (double)accL1, dS = (double)accS;
        for (int o = 16; o > 0; o >>= 1) {
            dL1 += __shfl_xor_sync(0xffffffffu, dL1, o);
            dS += __shfl_xor_sync(0xffffffffu, dS, o);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { s_red[0][warp] = dL1; s_red[1][warp] = dS; }
        __syncthreads();
        if (threadIdx.x == 0 && partial) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < SW / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
            atomicAdd(&partial[0], a);
            atomicAdd(&partial[1], b);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: transposed (flipped-window) separable convolution of the three maps, then the pointwise
// combination  grad1 = convT(A) + 2 v1 convT(B) + v2 convT(C)  [+ l1_scale * sign(v1 - v2)].
// Pixel x receives from centre x - (k - 5) with weight g[k]  ->  tap k reads element offset (5 - k) * C.
// Same streaming structure as the forward (horizontal taps from shared memory, vertical taps in registers).
// ------------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(SW) k_ssim_bwd(int H, int W, int Crt, const float* __restrict__ img1,
                                                 const float* __restrict__ img2, const float* __restrict__ mapA,
                                                 const float* __restrict__ mapB, const float* __restrict__ mapC,
                                                 const __grid_constant__ SsimWindow win, float l1_scale,
                                                 float* __restrict__ grad1, int srows)
{
    __shared__ __align__(16) float2 s_ab[2][SW + 2 * SHALO];   // [parity] maps (A, B) as one packed pair
    __shared__ __align__(16) float s_c[2][SW + 2 * SHALO];     // [parity] map C
    const int C = CT ? CT : Crt;
    const int RW = W * C;
    const int e0 = blockIdx.x * SW;
    const int r0 = blockIdx.y * srows, r1 = min(r0 + srows, H);
    const int t = threadIdx.x;
    const int e = e0 + t;
    RowStager<3> st;
    st.img[0] = mapA; st.img[1] = mapB; st.img[2] = mapC; st.cimg[0] = img1; st.cimg[1] = img2;
    st.H = H; st.RW = RW; st.e0 = e0; st.halo = SPAD * C; st.nload = SW + 2 * SPAD * C; st.t = t; st.r0 = r0;
    // maps A and B travel as one packed f32x2 pair, C as a scalar
    f32x2 accp[SK];
    float acc2[SK];
#pragma unroll
    for (int j = 0; j < SK; ++j) { accp[j] = f2_bc(0.f); acc2[j] = 0.f; }

    const int rbeg = r0 - SPAD, rend = r1 + SPAD;
    long long oidx = (long long)(rbeg - SPAD) * RW + e;
    st.begin(rbeg);
    st.fetch(rbeg);
    for (int rb = rbeg; rb < rend; rb += SK) {
#pragma unroll
        for (int ph = 0; ph < SK; ++ph) {
            const int r = rb + ph;
            if (r < rend) {
                const int par = (r - rbeg) & 1;
                st.store_ab_c(&s_ab[par][0], &s_c[par][0]);
                __syncthreads();
                if (r + 1 < rend) st.fetch(r + 1);
                f32x2 hp = f2_bc(0.f);
                float h2 = 0.f;
                const uint32_t ab_addr = smem_u32(&s_ab[par][t]);
#pragma unroll
                for (int k = 0; k < SK; ++k) {
                    const float w = win.g[k];
                    const int off = (2 * SPAD - k) * C;
                    hp = f2_fma(f2_bc(w), lds64_pair(ab_addr + (uint32_t)off * 8u), hp);
                    h2 = fmaf(w, s_c[par][t + off], h2);
                }
                // transposed vertical pass: output row ro receives centre row rc = ro + 5 - k with weight g[k];
                // centre row r feeds ro = r + d - 5 with k = r - ro + 5 ... flipped: k = d
#pragma unroll
                for (int d = 0; d < SK; ++d) {
                    const int slot = (ph + d + SK - SPAD) % SK;
                    const float w = win.g[d];
                    accp[slot] = f2_fma(f2_bc(w), hp, accp[slot]);
                    acc2[slot] = fmaf(w, h2, acc2[slot]);
                }
                const int ro = r - SPAD;
                const int oslot = (ph + SK - SPAD) % SK;
                if (ro >= r0 && e < RW) {
                    const long long idx = oidx;
                    const float v1 = st.ctr[0], v2 = st.ctr[1];
                    float g = f2_lo(accp[oslot]) + 2.0f * v1 * f2_hi(accp[oslot]) + v2 * acc2[oslot];
                    if (l1_scale != 0.0f) {
                        const float dd = v1 - v2;
                        g += dd > 0.0f ? l1_scale : (dd < 0.0f ? -l1_scale : 0.0f);
                    }
                    grad1[idx] = g;
                }
                accp[oslot] = f2_bc(0.f); acc2[oslot] = 0.f;
                oidx += RW;
            }
        }
    }
}

// pointwise scaling of the three maps by a per-centre upstream gradient (parity API only)
__global__ void k_scale_maps(size_t n, const float* __restrict__ up, float* __restrict__ a, float* __restrict__ b,
                             float* __restrict__ c)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float u = up[i];
    a[i] *= u; b[i] *= u; c[i] *= u;
}

__global__ void k_loss_finalize(const double* __restrict__ partial, double inv_count, float lambda, float scale,
                                float* __restrict__ loss_accum)
{
    // total = (1-lambda)*mean|d| + lambda*(1 - mean(ssim))   (GaussianTrainer.swift:710-714)
    const double l1 = partial[0] * inv_count;
    const double ssim_loss = 1.0 - partial[1] * inv_count;
    const double total = (1.0 - (double)lambda) * l1 + (double)lambda * ssim_loss;
    loss_accum[0] += (float)(total * (double)scale);
}

// ------------------------------------------------------------------------------------------------
// Depth supervision (Trainer/GaussianTrainer.swift:693-699,710-714; enabled when the dataset carries depth, :949):
//     depthLoss = sum(|depth - target| * mask) / max(sum(mask), 1e-6),   total += lambda_depth * depthLoss
// mask = target alpha > 0.5 as bytes (MLX bool array).  Two small HBM-bound passes: the reduction (sum of the masked
// differences and of the mask, doubles), then the cotangent  lambda * scale * mask * sign(depth - target) / weight
// (MLX abs has derivative 0 at 0) and the contribution to the loss scalar.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_depth_loss_reduce(size_t P, const float* __restrict__ depth, const float* __restrict__ target,
                                                           const uint8_t* __restrict__ mask, double* __restrict__ partial)
{
    double a = 0.0, w = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mask[i] ? 1.0f : 0.0f;
        a += (double)(fabsf(depth[i] - target[i]) * m);
        w += (double)m;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        w += __shfl_xor_sync(0xffffffffu, w, o);
    }
    __shared__ double s_a[8], s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_a[warp] = a; s_w[warp] = w; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tw = 0.0;
        for (int i = 0; i < 8; ++i) { ta += s_a[i]; tw += s_w[i]; }
        atomicAdd(&partial[0], ta);
        atomicAdd(&partial[1], tw);
    }
}

__global__ void __launch_bounds__(256) k_depth_loss_bwd(size_t P, const float* __restrict__ depth, const float* __restrict__ target,
                                                        const uint8_t* __restrict__ mask, const double* __restrict__ partial,
                                                        float lambda_depth, float scale, float* __restrict__ cot_depth,
                                                        float* __restrict__ loss_accum)
{
    const float weight = fmaxf((float)partial[1], 1e-6f);   // safeDepthWeight
    const float g = lambda_depth * scale / weight;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
        const float d = depth[i] - target[i];
        cot_depth[i] = mask[i] ? (d > 0.0f ? g : (d < 0.0f ? -g : 0.0f)) : 0.0f;
    }
    if (loss_accum && blockIdx.x == 0 && threadIdx.x == 0)
        loss_accum[0] += (float)((double)lambda_depth * (partial[0] / (double)weight) * (double)scale);
}

cudaError_t launch_depth_loss(cudaStream_t st, size_t P, const float* depth, const float* target_depth, const uint8_t* mask,
                              float lambda_depth, float scale, float* cot_depth, double* partial2, float* loss_accum)
{
    cudaError_t e = cudaMemsetAsync(partial2, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess || P == 0) return e;
    const int blocks = (int)std::min<size_t>((P + 255) / 256, 148 * 8);
    k_depth_loss_reduce<<<blocks, 256, 0, st>>>(P, depth, target_depth, mask, partial2);
    k_depth_loss_bwd<<<blocks, 256, 0, st>>>(P, depth, target_depth, mask, partial2, lambda_depth, scale, cot_depth, loss_accum);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static const SsimWindow& window()
{
    static const SsimWindow w = make_window();
    return w;
}

// Output rows per CTA: every CTA walks its rows serially, so the kernel lasts as long as one CTA; the image is cut into
// as many row chunks as fit in ONE wave of resident CTAs (occupancy x SM count / strips), but not below SROWS_MIN rows
// (each chunk filters 10 halo rows on top of its own).
template <typename K>
static int rows_per_cta(K kernel, int H, int strips)
{
    int dev = 0, sms = 148, per_sm = 4;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SW, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    const int chunks = std::max(1, (sms * per_sm) / std::max(strips, 1));
    return std::max(SROWS_MIN, cdiv(H, chunks));
}

cudaError_t launch_ssim_fwd(cudaStream_t st, int H, int W, int C, const float* img1, const float* img2, float* ssim_map,
                            float* mu1, float* mu2, float* s1, float* s2, float* s12)
{
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    const int strips = cdiv((long long)W * C, SW);
    const int srows = C == 3 ? rows_per_cta(k_ssim_fwd<0, 3>, H, strips) : rows_per_cta(k_ssim_fwd<0, 0>, H, strips);
    dim3 grid(strips, cdiv(H, srows));
    if (C == 3) k_ssim_fwd<0, 3><<<grid, SW, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, ssim_map, mu1, mu2, s1, s2, s12, nullptr, srows);
    else k_ssim_fwd<0, 0><<<grid, SW, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, ssim_map, mu1, mu2, s1, s2, s12, nullptr, srows);
    return cudaGetLastError();
}

cudaError_t launch_loss_fwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            float upstream, float* mapA, float* mapB, float* mapC, double* partial)
{
    cudaError_t e = cudaMemsetAsync(partial, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    const int strips = cdiv((long long)W * C, SW);
    const int srows = C == 3 ? rows_per_cta(k_ssim_fwd<1, 3>, H, strips) : rows_per_cta(k_ssim_fwd<1, 0>, H, strips);
    dim3 grid(strips, cdiv(H, srows));
    if (C == 3) k_ssim_fwd<1, 3><<<grid, SW, 0, st>>>(H, W, C, render, target, window(), upstream, mapA, mapB, mapC, nullptr, nullptr,
                                                      nullptr, partial, srows);
    else k_ssim_fwd<1, 0><<<grid, SW, 0, st>>>(H, W, C, render, target, window(), upstream, mapA, mapB, mapC, nullptr, nullptr,
                                               nullptr, partial, srows);
    return cudaGetLastError();
}

cudaError_t launch_loss_bwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            const float* mapA, const float* mapB, const float* mapC, float l1_scale, float* cot_render)
{
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    const int strips = cdiv((long long)W * C, SW);
    const int srows = C == 3 ? rows_per_cta(k_ssim_bwd<3>, H, strips) : rows_per_cta(k_ssim_bwd<0>, H, strips);
    dim3 grid(strips, cdiv(H, srows));
    if (C == 3) k_ssim_bwd<3><<<grid, SW, 0, st>>>(H, W, C, render, target, mapA, mapB, mapC, window(), l1_scale, cot_render, srows);
    else k_ssim_bwd<0><<<grid, SW, 0, st>>>(H, W, C, render, target, mapA, mapB, mapC, window(), l1_scale, cot_render, srows);
    return cudaGetLastError();
}

cudaError_t launch_loss_finalize(cudaStream_t st, const double* partial, double inv_count, float lambda, float scale,
                                 float* loss_accum)
{
    k_loss_finalize<<<1, 1, 0, st>>>(partial, inv_count, lambda, scale, loss_accum);
    return cudaGetLastError();
}

cudaError_t launch_ssim_bwd_api(cudaStream_t st, int H, int W, int C, const float* grad_out, const float* img1,
                                const float* img2, float* mapA, float* mapB, float* mapC, float* grad_img1)
{
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    const int strips = cdiv((long long)W * C, SW);
    const int srows = C == 3 ? rows_per_cta(k_ssim_fwd<1, 3>, H, strips) : rows_per_cta(k_ssim_fwd<1, 0>, H, strips);
    dim3 grid(strips, cdiv(H, srows));
    // maps with unit upstream, then scaled by the caller's per-centre gradient
    if (C == 3) k_ssim_fwd<1, 3><<<grid, SW, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, mapA, mapB, mapC, nullptr, nullptr, nullptr, nullptr, srows);
    else k_ssim_fwd<1, 0><<<grid, SW, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, mapA, mapB, mapC, nullptr, nullptr, nullptr, nullptr, srows);
    const size_t n = (size_t)H * W * C;
    k_scale_maps<<<cdiv((long long)n, 256), 256, 0, st>>>(n, grad_out, mapA, mapB, mapC);
    const int srows_b = C == 3 ? rows_per_cta(k_ssim_bwd<3>, H, strips) : rows_per_cta(k_ssim_bwd<0>, H, strips);
    dim3 grid_b(strips, cdiv(H, srows_b));
    if (C == 3) k_ssim_bwd<3><<<grid_b, SW, 0, st>>>(H, W, C, img1, img2, mapA, mapB, mapC, window(), 0.0f, grad_img1, srows_b);
    else k_ssim_bwd<0><<<grid_b, SW, 0, st>>>(H, W, C, img1, img2, mapA, mapB, mapC, window(), 0.0f, grad_img1, srows_b);
    return cudaGetLastError();
}

}  // namespace gsb
