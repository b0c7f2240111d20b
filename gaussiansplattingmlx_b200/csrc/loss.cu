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

constexpr int ST = 16;            // tile edge in pixels
constexpr int SPAD = 5;           // window radius (K/2)
constexpr int SK = 11;
constexpr int SH = ST + 2 * SPAD; // halo edge (26)
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
// forward statistics for one 16x16xC tile: horizontal pass into s_h, vertical pass per thread.
// MODE 0: parity API (writes ssim + optional saved maps)
// MODE 1: training (writes upstream-scaled A/B/C maps, accumulates sum|d| and sum(ssim))
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_ssim_fwd(int H, int W, int C, const float* __restrict__ img1,
                                                  const float* __restrict__ img2, const __grid_constant__ SsimWindow win,
                                                  float upstream, float* __restrict__ o0, float* __restrict__ o1,
                                                  float* __restrict__ o2, float* __restrict__ o3, float* __restrict__ o4,
                                                  float* __restrict__ o5, double* __restrict__ partial)
{
    __shared__ float s_a[SH][SH + 1];
    __shared__ float s_b[SH][SH + 1];
    __shared__ float s_h[5][SH][ST + 1];
    __shared__ double s_red[2][8];
    const int tx0 = blockIdx.x * ST, ty0 = blockIdx.y * ST;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    double accL1 = 0.0, accS = 0.0;
    for (int c = 0; c < C; ++c) {
        __syncthreads();
        for (int i = threadIdx.x; i < SH * SH; i += 256) {
            const int hy = i / SH, hx = i - hy * SH;
            const int y = ty0 + hy - SPAD, x = tx0 + hx - SPAD;
            float a = 0.f, b = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const size_t si = ((size_t)y * W + x) * C + c;
                a = img1[si];
                b = img2[si];
            }
            s_a[hy][hx] = a;
            s_b[hy][hx] = b;
        }
        __syncthreads();
        // horizontal: SH rows x ST columns x 5 statistics
        for (int i = threadIdx.x; i < SH * ST; i += 256) {
            const int hy = i / ST, ox = i - hy * ST;
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                const float a = s_a[hy][ox + k], b = s_b[hy][ox + k];
                m1 += w * a;
                m2 += w * b;
                e11 += w * (a * a);
                e22 += w * (b * b);
                e12 += w * (a * b);
            }
            s_h[0][hy][ox] = m1; s_h[1][hy][ox] = m2; s_h[2][hy][ox] = e11; s_h[3][hy][ox] = e22; s_h[4][hy][ox] = e12;
        }
        __syncthreads();
        const int x = tx0 + lx, y = ty0 + ly;
        if (x < W && y < H) {
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                m1 += w * s_h[0][ly + k][lx];
                m2 += w * s_h[1][ly + k][lx];
                e11 += w * s_h[2][ly + k][lx];
                e22 += w * s_h[3][ly + k][lx];
                e12 += w * s_h[4][ly + k][lx];
            }
            const SsimPoint sp = ssim_point(m1, m2, e11, e22, e12);
            const size_t idx = ((size_t)y * W + x) * C + c;
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
                accL1 += (double)fabsf(s_a[ly + SPAD][lx + SPAD] - s_b[ly + SPAD][lx + SPAD]);
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
            for (int w = 0; w < 8; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
            atomicAdd(&partial[0], a);
            atomicAdd(&partial[1], b);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: transposed separable convolution of the three maps, then the pointwise combination.
// PRE = true : maps already hold upstream * (gm1, gs1, gs12)            (training path)
// PRE = false: maps are computed here from (grad_out, img1, img2)        (parity API) — done by the
//              launcher through k_ssim_fwd<1>-style statistics, see k_ssim_maps below.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ssim_bwd(int H, int W, int C, const float* __restrict__ img1,
                                                  const float* __restrict__ img2, const float* __restrict__ mapA,
                                                  const float* __restrict__ mapB, const float* __restrict__ mapC,
                                                  const __grid_constant__ SsimWindow win, float l1_scale,
                                                  float* __restrict__ grad1)
{
    __shared__ float s_m[3][SH][SH + 1];
    __shared__ float s_h[3][SH][ST + 1];
    const int tx0 = blockIdx.x * ST, ty0 = blockIdx.y * ST;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    for (int c = 0; c < C; ++c) {
        __syncthreads();
        for (int i = threadIdx.x; i < SH * SH; i += 256) {
            const int hy = i / SH, hx = i - hy * SH;
            const int y = ty0 + hy - SPAD, x = tx0 + hx - SPAD;
            float a = 0.f, b = 0.f, cc = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const size_t si = ((size_t)y * W + x) * C + c;
                a = mapA[si];
                b = mapB[si];
                cc = mapC[si];
            }
            s_m[0][hy][hx] = a; s_m[1][hy][hx] = b; s_m[2][hy][hx] = cc;
        }
        __syncthreads();
        // pixel x receives from centre cx = x - k + pad with weight g[k]: halo column (ox + 2*pad - k)
        for (int i = threadIdx.x; i < SH * ST; i += 256) {
            const int hy = i / ST, ox = i - hy * ST;
            float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                const int hx = ox + 2 * SPAD - k;
                a += w * s_m[0][hy][hx];
                b += w * s_m[1][hy][hx];
                cc += w * s_m[2][hy][hx];
            }
            s_h[0][hy][ox] = a; s_h[1][hy][ox] = b; s_h[2][hy][ox] = cc;
        }
        __syncthreads();
        const int x = tx0 + lx, y = ty0 + ly;
        if (x < W && y < H) {
            float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float w = win.g[k];
                const int hy = ly + 2 * SPAD - k;
                a += w * s_h[0][hy][lx];
                b += w * s_h[1][hy][lx];
                cc += w * s_h[2][hy][lx];
            }
            const size_t idx = ((size_t)y * W + x) * C + c;
            const float v1 = img1[idx], v2 = img2[idx];
            float g = a + 2.0f * v1 * b + v2 * cc;
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
    dim3 grid(cdiv(W, ST), cdiv(H, ST));
    k_ssim_fwd<0><<<grid, 256, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, ssim_map, mu1, mu2, s1, s2, s12, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_loss_fwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            float upstream, float* mapA, float* mapB, float* mapC, double* partial)
{
    cudaError_t e = cudaMemsetAsync(partial, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    dim3 grid(cdiv(W, ST), cdiv(H, ST));
    k_ssim_fwd<1><<<grid, 256, 0, st>>>(H, W, C, render, target, window(), upstream, mapA, mapB, mapC, nullptr, nullptr,
                                        nullptr, partial);
    return cudaGetLastError();
}

cudaError_t launch_loss_bwd(cudaStream_t st, int H, int W, int C, const float* render, const float* target,
                            const float* mapA, const float* mapB, const float* mapC, float l1_scale, float* cot_render)
{
    if (H <= 0 || W <= 0 || C <= 0) return cudaSuccess;
    dim3 grid(cdiv(W, ST), cdiv(H, ST));
    k_ssim_bwd<<<grid, 256, 0, st>>>(H, W, C, render, target, mapA, mapB, mapC, window(), l1_scale, cot_render);
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
    dim3 grid(cdiv(W, ST), cdiv(H, ST));
    // maps with unit upstream, then scaled by the caller's per-centre gradient
    k_ssim_fwd<1><<<grid, 256, 0, st>>>(H, W, C, img1, img2, window(), 1.0f, mapA, mapB, mapC, nullptr, nullptr, nullptr,
                                        nullptr);
    const size_t n = (size_t)H * W * C;
    k_scale_maps<<<cdiv((long long)n, 256), 256, 0, st>>>(n, grad_out, mapA, mapB, mapC);
    k_ssim_bwd<<<grid, 256, 0, st>>>(H, W, C, img1, img2, mapA, mapB, mapC, window(), 0.0f, grad_img1);
    return cudaGetLastError();
}

}  // namespace gsb
