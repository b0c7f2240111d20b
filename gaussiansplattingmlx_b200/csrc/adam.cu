// adam.cu — fused Adam over the six parameter tensors + D1 accum_grad_norm, one launch.
//
// Reference: MLXOptimizers.Adam.applySingle called once per tensor with a per-tensor learning
// rate (Trainer/GaussianTrainer.swift:941-948,1066-1079; mlx-swift 0.30.6, un-vendored — published
// update rule, NO bias correction):
//     m = b1*m + (1-b1)*g ;  v = b2*v + (1-b2)*g*g ;  p = p - lr * m / (sqrt(v) + eps)
// and accum_grad_norm (GaussianTrainer.swift:321-339): accum[i] += ||grad_xyz_i||_2.
// HBM-bound: 28 B per parameter float (read p,g,m,v; write p,m,v).  Compiled with --fmad=false so
// the update rounds exactly like the CPU oracle (bit-exact Adam parity).
#include "kernels.h"

namespace gsb {

constexpr int AD_THREADS = 256;

struct AdamSeg {
    long long vec_begin[7];  // prefix of float4-chunk counts per tensor (tensor t owns chunks [vec_begin[t], vec_begin[t+1]))
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps)
{
    const float mi = b1 * m + (1.0f - b1) * g;
    const float vi = b2 * v + (1.0f - b2) * (g * g);
    m = mi;
    v = vi;
    p = p - lr * mi / (sqrtf(vi) + eps);
}

__global__ void __launch_bounds__(AD_THREADS) k_adam(const __grid_constant__ AdamTensors t, const __grid_constant__ AdamSeg seg,
                                                     float b1, float b2, float eps, float gscale, int N,
                                                     float* __restrict__ grad_norm_accum,
                                                     const uint32_t* __restrict__ skip_flag)
{
    if (skip_flag && *skip_flag) return;
    const long long total = seg.vec_begin[6];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < 6; ++j) k += (i >= seg.vec_begin[j]) ? 1 : 0;
        const long long c = i - seg.vec_begin[k];
        const long long base = c * 4;
        const long long cnt = t.count[k];
        const float lr = t.lr[k];
        float* p = t.p[k] + base;
        const float* g = t.g[k] + base;
        float* m = t.m[k] + base;
        float* v = t.v[k] + base;
        if (base + 4 <= cnt) {
            float4 P = *reinterpret_cast<float4*>(p);
            float4 G = *reinterpret_cast<const float4*>(g);
            float4 M = *reinterpret_cast<float4*>(m);
            float4 V = *reinterpret_cast<float4*>(v);
            G.x = G.x * gscale; G.y = G.y * gscale; G.z = G.z * gscale; G.w = G.w * gscale;
            adam1(P.x, G.x, M.x, V.x, lr, b1, b2, eps);
            adam1(P.y, G.y, M.y, V.y, lr, b1, b2, eps);
            adam1(P.z, G.z, M.z, V.z, lr, b1, b2, eps);
            adam1(P.w, G.w, M.w, V.w, lr, b1, b2, eps);
            *reinterpret_cast<float4*>(p) = P;
            *reinterpret_cast<float4*>(m) = M;
            *reinterpret_cast<float4*>(v) = V;
        } else {
            for (long long e = 0; base + e < cnt; ++e) {
                float pp = p[e], mm = m[e], vv = v[e];
                adam1(pp, g[e] * gscale, mm, vv, lr, b1, b2, eps);
                p[e] = pp; m[e] = mm; v[e] = vv;
            }
        }
    }
    // D1: one thread per Gaussian (gradient re-read hits L2)
    if (grad_norm_accum) {
        const float* gx = t.g[0];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
            const float a = gx[i * 3] * gscale, b = gx[i * 3 + 1] * gscale, c = gx[i * 3 + 2] * gscale;
            grad_norm_accum[i] = grad_norm_accum[i] + sqrtf(a * a + b * b + c * c);
        }
    }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// scalar fallback for unaligned caller pointers (parity API only)
__global__ void k_adam_scalar(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                              long long n, float lr, float b1, float b2, float eps, float gscale)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pp = p[i], mm = m[i], vv = v[i];
    adam1(pp, g[i] * gscale, mm, vv, lr, b1, b2, eps);
    p[i] = pp; m[i] = mm; v[i] = vv;
}
__global__ void k_accum_grad_norm(int N, const float* __restrict__ gx, float gscale, float* __restrict__ accum)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float a = gx[i * 3] * gscale, b = gx[i * 3 + 1] * gscale, c = gx[i * 3 + 2] * gscale;
    accum[i] = accum[i] + sqrtf(a * a + b * b + c * c);
}

cudaError_t launch_adam(cudaStream_t st, const AdamTensors& t, float beta1, float beta2, float eps, float gscale, int N,
                        float* grad_norm_accum, const uint32_t* skip_flag, int* launches)
{
    bool aligned = true;
    for (int k = 0; k < 6; ++k) aligned = aligned && al16(t.p[k]) && al16(t.g[k]) && al16(t.m[k]) && al16(t.v[k]);
    if (!aligned) {
        for (int k = 0; k < 6; ++k)
            if (t.count[k] > 0) {
                k_adam_scalar<<<cdiv(t.count[k], 256), 256, 0, st>>>(t.p[k], t.g[k], t.m[k], t.v[k], t.count[k], t.lr[k], beta1,
                                                                     beta2, eps, gscale);
                if (launches) ++*launches;
            }
        if (grad_norm_accum && N > 0) {
            k_accum_grad_norm<<<cdiv(N, 256), 256, 0, st>>>(N, t.g[0], gscale, grad_norm_accum);
            if (launches) ++*launches;
        }
        return cudaGetLastError();
    }
    AdamSeg seg;
    seg.vec_begin[0] = 0;
    for (int k = 0; k < 6; ++k) seg.vec_begin[k + 1] = seg.vec_begin[k] + (t.count[k] + 3) / 4;
    const long long total = seg.vec_begin[6];
    if (total == 0) return cudaSuccess;
    long long blocks = (total + AD_THREADS - 1) / AD_THREADS;
    const long long cap = 148LL * 8 * 4;   // grid-stride: 32 resident-CTA waves' worth at most
    if (blocks > cap) blocks = cap;
    k_adam<<<(int)blocks, AD_THREADS, 0, st>>>(t, seg, beta1, beta2, eps, gscale, N, grad_norm_accum, skip_flag);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace gsb
