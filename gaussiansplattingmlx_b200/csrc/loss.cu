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
#include "kernels.h"

namespace gsb {

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
    SsimPoint r;
    r.ssim = (A * B) / CD;
    const float gA = B / CD, gB = A / CD;
    const float gCD = -(A * B) / (CD * CD);
    const float gC = gCD * D, gD = gCD * Cc;
    r.gs1 = gD;
    r.gs12 = 2.0f * gB;
    r.gm1 = 2.0f * mu2 * gA + 2.0f * mu1 * gC - 2.0f * mu1 * gD - mu2 * r.gs12;
    return r;
}

// ------------------------------------------------------------------------------------------------
// Streaming separable convolution.  The HWC image is treated as H rows of W*C floats; a horizontal tap
// k of the per-channel window sits at element offset (k - 5) * C.  A CTA owns a strip of SW consecutive
// row elements and a chunk of SROWS output rows and marches down the rows: each input row is loaded once
// into shared memory (strip + 5*C halo on both sides), filtered horizontally (11 taps) into a ring of
// the last 11 filtered rows, and one output row is produced from the ring (11 vertical taps).  Compared
// with 2-D tiles the halo overhead is 10 rows per SROWS-row chunk and 10*C floats per SW-float strip.
// MODE 0: parity API (writes ssim + optional saved maps)
// MODE 1: training (writes upstream-scaled A/B/C maps, accumulates sum|d| and sum(ssim))
// ------------------------------------------------------------------------------------------------
constexpr int SW = 192;        // strip width in floats = threads per CTA (1080p: 5760 = 30 strips)
constexpr int SROWS = 60;      // output rows per CTA
constexpr int SMAXC = 4;       // channels supported by the halo buffer
constexpr int SHALO = SPAD * SMAXC;

template <int MODE>
__global__ void __launch_bounds__(SW) k_ssim_fwd(int H, int W, int C, const float* __restrict__ img1,
                                                 const float* __restrict__ img2, const __grid_constant__ SsimWindow win,
                                                 float upstream, float* __restrict__ o0, float* __restrict__ o1,
                                                 float* __restrict__ o2, float* __restrict__ o3, float* __restrict__ o4,
                                                 float* __restrict__ o5, double* __restrict__ partial)
{
    __shared__ float s_row[4][SW + 2 * SHALO];   // [parity][image]
    __shared__ float s_ring[SK][5][SW];
    __shared__ double s_red[2][SW / 32];
    const int RW = W * C;                          // floats per image row
    const int e0 = blockIdx.x * SW;                // first row element of the strip
    const int r0 = blockIdx.y * SROWS, r1 = min(r0 + SROWS, H);
    const int t = threadIdx.x;
    const int e = e0 + t;
    const int halo = SPAD * C;
    double accL1 = 0.0, accS = 0.0;
    // row r+1 is fetched into registers while row r is filtered (each thread owns elements t and t + SW of the
    // haloed strip); zero outside the image
    const int nload = SW + 2 * halo;
    float pa[2], pb[2];
    auto fetch = [&](int r) {
        const bool row_ok = r >= 0 && r < H;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            const int ee = e0 - halo + i;
            pa[u] = 0.f; pb[u] = 0.f;
            if (row_ok && i < nload && ee >= 0 && ee < RW) {
                const size_t si = (size_t)r * RW + ee;
                pa[u] = img1[si];
                pb[u] = img2[si];
            }
        }
    };
    fetch(r0 - SPAD);
    for (int r = r0 - SPAD; r < r1 + SPAD; ++r) {
        float (*row)[SW + 2 * SHALO] = reinterpret_cast<float (*)[SW + 2 * SHALO]>(&s_row[(r & 1) * 2][0]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            if (i < nload) { row[0][i] = pa[u]; row[1][i] = pb[u]; }
        }
        __syncthreads();   // the only barrier per row: the two row buffers alternate
        if (r + 1 < r1 + SPAD) fetch(r + 1);
        // ---- horizontal 11 taps -> ring slot of row r
        {
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                const float a = row[0][t + k * C], b = row[1][t + k * C];
                m1 = fmaf(w, a, m1);
                m2 = fmaf(w, b, m2);
                e11 = fmaf(w, a * a, e11);
                e22 = fmaf(w, b * b, e22);
                e12 = fmaf(w, a * b, e12);
            }
            const int slot = (r + SK) % SK;
            s_ring[slot][0][t] = m1; s_ring[slot][1][t] = m2; s_ring[slot][2][t] = e11; s_ring[slot][3][t] = e22;
            s_ring[slot][4][t] = e12;
        }
        // ---- vertical 11 taps -> output row ro = r - 5 (each thread reads only its own ring column)
        const int ro = r - SPAD;
        if (ro >= r0 && e < RW) {
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
            const int base = (ro - SPAD + 2 * SK) % SK;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                int slot = base + k;
                if (slot >= SK) slot -= SK;
                m1 = fmaf(w, s_ring[slot][0][t], m1);
                m2 = fmaf(w, s_ring[slot][1][t], m2);
                e11 = fmaf(w, s_ring[slot][2][t], e11);
                e22 = fmaf(w, s_ring[slot][3][t], e22);
                e12 = fmaf(w, s_ring[slot][4][t], e12);
            }
            const SsimPoint sp = ssim_point(m1, m2, e11, e22, e12);
            const size_t idx = (size_t)ro * RW + e;
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
                accS += (double)sp.ssim;
                accL1 += (double)fabsf(img1[idx] - img2[idx]);
            }
        }
    }
    if (MODE == 1) {
        for (int o = 16; o > 0; o >>= 1) {
            accL1 += __shfl_xor_sync(0xffffffffu, accL1, o);
            accS += __shfl_xor_sync(0xffffffffu, accS, o);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { s_red[0][warp] = accL1; s_red[1][warp] = accS; }
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
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SW) k_ssim_bwd(int H, int W, int C, const float* __restrict__ img1,
                                                 const float* __restrict__ img2, const float* __restrict__ mapA,
                                                 const float* __restrict__ mapB, const float* __restrict__ mapC,
                                                 const __grid_constant__ SsimWindow win, float l1_scale,
                                                 float* __restrict__ grad1)
{
    __shared__ float s_row[6][SW + 2 * SHALO];   // [parity][map]
    __shared__ float s_ring[SK][3][SW];
    const int RW = W * C;
    const int e0 = blockIdx.x * SW;
    const int r0 = blockIdx.y * SROWS, r1 = min(r0 + SROWS, H);
    const int t = threadIdx.x;
    const int e = e0 + t;
    const int halo = SPAD * C;
    const int nload = SW + 2 * halo;
    float pa[2], pb[2], pc[2];
    auto fetch = [&](int r) {
        const bool row_ok = r >= 0 && r < H;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            const int ee = e0 - halo + i;
            pa[u] = 0.f; pb[u] = 0.f; pc[u] = 0.f;
            if (row_ok && i < nload && ee >= 0 && ee < RW) {
                const size_t si = (size_t)r * RW + ee;
                pa[u] = mapA[si];
                pb[u] = mapB[si];
                pc[u] = mapC[si];
            }
        }
    };
    fetch(r0 - SPAD);
    for (int r = r0 - SPAD; r < r1 + SPAD; ++r) {
        float (*row)[SW + 2 * SHALO] = reinterpret_cast<float (*)[SW + 2 * SHALO]>(&s_row[(r & 1) * 3][0]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * SW;
            if (i < nload) { row[0][i] = pa[u]; row[1][i] = pb[u]; row[2][i] = pc[u]; }
        }
        __syncthreads();
        if (r + 1 < r1 + SPAD) fetch(r + 1);
        {
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                const int off = t + (2 * SPAD - k) * C;
                a = fmaf(w, row[0][off], a);
                b = fmaf(w, row[1][off], b);
                c = fmaf(w, row[2][off], c);
            }
            const int slot = (r + SK) % SK;
            s_ring[slot][0][t] = a; s_ring[slot][1][t] = b; s_ring[slot][2][t] = c;
        }
        const int ro = r - SPAD;
        if (ro >= r0 && e < RW) {
            float a = 0.f, b = 0.f, c = 0.f;
            const int base = (ro + SPAD + SK) % SK;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                int slot = base - k;
                if (slot < 0) slot += SK;
                a = fmaf(w, s_ring[slot][0][t], a);
                b = fmaf(w, s_ring[slot][1][t], b);
                c = fmaf(w, s_ring[slot][2][t], c);
            }
            const size_t idx = (size_t)ro * RW + e;
            const float v1 = img1[idx], v2 = img2[idx];
            float g = a + 2.0f * v1 * b + v2 * c;
            if (l1_scale != 0.0f) {
                const float d = v1 - v2;
                g += d > 0.0f ? l1_scale : (d < 0.0f ? -l1_scale : 0.0f);
            }
            grad1[idx] = g;
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
// launchers
// ------------------------------------------------------------------------------------------------
static const SsimWindow& window()
{
    static const SsimWindow w = make_window();
    return w;
}

cudaError_t launch_ssim_fwd(cudaStream_t st, int H, int W, int C, const float* img1, const float* img2, float* ssim_map,
                            float* mu1, float* mu2, float* s1, float* s2, float* s12)
{
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    dim3 grid(cdiv((long long)W * C, SW), cdiv(H, SROWS));
    k_ssim_fwd<0><<<grid, SW, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, ssim_map, mu1, mu2, s1, s2, s12, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_loss_fwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            float upstream, float* mapA, float* mapB, float* mapC, double* partial)
{
    cudaError_t e = cudaMemsetAsync(partial, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    dim3 grid(cdiv((long long)W * C, SW), cdiv(H, SROWS));
    k_ssim_fwd<1><<<grid, SW, 0, st>>>(H, W, C, render, target, window(), upstream, mapA, mapB, mapC, nullptr, nullptr,
                                        nullptr, partial);
    return cudaGetLastError();
}

cudaError_t launch_loss_bwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            const float* mapA, const float* mapB, const float* mapC, float l1_scale, float* cot_render)
{
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    dim3 grid(cdiv((long long)W * C, SW), cdiv(H, SROWS));
    k_ssim_bwd<<<grid, SW, 0, st>>>(H, W, C, render, target, mapA, mapB, mapC, window(), l1_scale, cot_render);
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
    dim3 grid(cdiv((long long)W * C, SW), cdiv(H, SROWS));
    // maps with unit upstream, then scaled by the caller's per-centre gradient
    k_ssim_fwd<1><<<grid, SW, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, mapA, mapB, mapC, nullptr, nullptr, nullptr,
                                        nullptr);
    const size_t n = (size_t)H * W * C;
    k_scale_maps<<<cdiv((long long)n, 256), 256, 0, st>>>(n, grad_out, mapA, mapB, mapC);
    k_ssim_bwd<<<grid, SW, 0, st>>>(H, W, C, img1, img2, mapA, mapB, mapC, window(), 0.0f, grad_img1);
    return cudaGetLastError();
}

}  // namespace gsb
