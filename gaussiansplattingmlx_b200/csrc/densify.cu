// densify.cu — clone / split / prune for sm_100a.  Compiled with --fmad=false so that the
// classification thresholds and the split/clone arithmetic round like the CPU oracle.
//
// Reference: split_and_prune (Trainer/GaussianTrainer.swift:766-908) and its inline Metal kernels
//   D2 classify_gaussians          :344-392   keep 0 / split 1 / clone 2 / prune 3 + output count 1/2/2/0
//   D3 build_densify_output_map    :397-427   gather index + noise mode per output slot
// plus the MLX ops around them: cumsum -> offsets (:813-817, `.item()` host sync), gathers of the six tensors
// (:866-871), scale reduction and position noise (:874-897).  All HBM-bound streaming work; the only data
// dependent size is the output count, which is read back once per densification (every 100 iterations).
#include "kernels.h"

namespace gsb {

// ------------------------------------------------------------------------------------------------
// D2 + action statistics
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_densify_classify(int N, const float* __restrict__ grad_accum, float denom,
                                                          const float* __restrict__ scales_log,
                                                          const float* __restrict__ opacity_logit, float grad_threshold,
                                                          float max_scale_thresh, float min_opacity_thresh, int allow_densify,
                                                          int* __restrict__ actions, int* __restrict__ output_counts,
                                                          uint32_t* __restrict__ stats4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int action = -1;
    if (i < N) {
        // GaussianTrainer.swift:358-388
        const float g = grad_accum[i];
        const float avg = denom > 0.0f ? g / denom : 0.0f;
        const float s0 = expf(scales_log[i * 3]), s1 = expf(scales_log[i * 3 + 1]), s2 = expf(scales_log[i * 3 + 2]);
        const float mx = fmaxf(fmaxf(s0, s1), s2);
        const float op = 1.0f / (1.0f + expf(-opacity_logit[i]));
        int cnt;
        if (op < min_opacity_thresh) { action = 3; cnt = 0; }
        else if (allow_densify && avg > grad_threshold) {
            if (mx > max_scale_thresh) { action = 1; cnt = 2; } else { action = 2; cnt = 2; }
        } else { action = 0; cnt = 1; }
        actions[i] = action;
        output_counts[i] = cnt;
    }
    if (stats4) {   // keep / split / clone / prune totals, warp-aggregated
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const uint32_t m = __ballot_sync(0xffffffffu, action == a);
            if ((threadIdx.x & 31) == 0 && m) atomicAdd(&stats4[a], (uint32_t)__popc(m));
        }
    }
}

cudaError_t launch_densify_classify(cudaStream_t st, int N, const float* grad_accum, float denom, const float* scales_log,
                                    const float* opacity_logit, float grad_threshold, float max_scale, float min_opacity,
                                    int allow_densify, int* actions, int* counts, uint32_t* stats4)
{
    if (stats4) {
        cudaError_t e = cudaMemsetAsync(stats4, 0, 4 * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
    }
    if (N > 0)
        k_densify_classify<<<cdiv(N, 256), 256, 0, st>>>(N, grad_accum, denom, scales_log, opacity_logit, grad_threshold, max_scale,
                                                         min_opacity, allow_densify, actions, counts, stats4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// D3
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_densify_map(int N, const int* __restrict__ actions, const uint32_t* __restrict__ offsets,
                                                     uint32_t capacity, int* __restrict__ gather, int* __restrict__ noise_mode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int a = actions[i];
    const uint32_t o = offsets[i];
    if (a == 0) {
        if (o < capacity) { gather[o] = i; noise_mode[o] = 0; }
    } else if (a == 1) {
        if (o + 1 < capacity) { gather[o] = i; noise_mode[o] = 1; gather[o + 1] = i; noise_mode[o + 1] = 2; }
    } else if (a == 2) {
        if (o + 1 < capacity) { gather[o] = i; noise_mode[o] = 0; gather[o + 1] = i; noise_mode[o + 1] = 3; }
    }
}

cudaError_t launch_densify_map(cudaStream_t st, int N, const int* actions, const uint32_t* offsets, uint32_t capacity, int* gather,
                               int* noise_mode)
{
    if (N > 0) k_densify_map<<<cdiv(N, 256), 256, 0, st>>>(N, actions, offsets, capacity, gather, noise_mode);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// counter-based normal noise (stands for MLXRandom.normal, GaussianTrainer.swift:881): Philox4x32-10 keyed by the
// caller's seed, counter = output slot.  Every data-parallel replica draws the same numbers for the same slot.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ void normal3(uint64_t seed, uint32_t slot, float (&n)[3])
{
    uint32_t c[4] = {slot, 0u, 0x67736231u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u0 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f, u1 = (float)c[1] * 2.3283064365386963e-10f;
    const float u2 = ((float)c[2] + 0.5f) * 2.3283064365386963e-10f, u3 = (float)c[3] * 2.3283064365386963e-10f;
    const float r0 = sqrtf(-2.0f * logf(fminf(u0, 1.0f))), r1 = sqrtf(-2.0f * logf(fminf(u2, 1.0f)));
    float s, co;
    sincospif(2.0f * u1, &s, &co);
    n[0] = r0 * co; n[1] = r0 * s;
    sincospif(2.0f * u3, &s, &co);
    n[2] = r1 * co;
}

// ------------------------------------------------------------------------------------------------
// Phases 4-5: gather + per-slot modification.  A CTA owns DG consecutive output slots: the small tensors are
// handled by one thread per slot, the f_rest rows ((K-1)*3 floats each) are copied by the whole CTA so that both
// the gathered reads (one contiguous row per slot) and the writes (one contiguous block per CTA) are coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int DG = 64;

__global__ void __launch_bounds__(256) k_densify_apply(int Nout, int K, const int* __restrict__ gather,
                                                       const int* __restrict__ noise_mode, const float* __restrict__ base_noise,
                                                       uint64_t seed, const float* __restrict__ xyz, const float* __restrict__ f_dc,
                                                       const float* __restrict__ f_rest, const float* __restrict__ scales_log,
                                                       const float* __restrict__ rot, const float* __restrict__ opacity,
                                                       float* __restrict__ o_xyz, float* __restrict__ o_f_dc, float* __restrict__ o_f_rest,
                                                       float* __restrict__ o_scales_log, float* __restrict__ o_rot,
                                                       float* __restrict__ o_opacity)
{
    __shared__ int s_src[DG];
    const int j0 = blockIdx.x * DG;
    const int nj = min(DG, Nout - j0);
    if ((int)threadIdx.x < nj) s_src[threadIdx.x] = gather[j0 + threadIdx.x];
    __syncthreads();
    if ((int)threadIdx.x < nj) {
        const int j = j0 + threadIdx.x;
        const int s = s_src[threadIdx.x];
        const int mode = noise_mode[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) o_f_dc[(size_t)j * 3 + c] = f_dc[(size_t)s * 3 + c];
        const float4 q = *reinterpret_cast<const float4*>(rot + (size_t)s * 4);
        *reinterpret_cast<float4*>(o_rot + (size_t)j * 4) = q;
        o_opacity[j] = opacity[s];
        // GaussianTrainer.swift:874-897
        const float red = -0.47000362924573563f;   // Float(-log(1.6))
        const float isSplit = (mode == 1 || mode == 2) ? 1.0f : 0.0f;
        const float l0 = scales_log[(size_t)s * 3], l1 = scales_log[(size_t)s * 3 + 1], l2 = scales_log[(size_t)s * 3 + 2];
        o_scales_log[(size_t)j * 3 + 0] = l0 + isSplit * red;
        o_scales_log[(size_t)j * 3 + 1] = l1 + isSplit * red;
        o_scales_log[(size_t)j * 3 + 2] = l2 + isSplit * red;
        const float mean = ((expf(l0) + expf(l1)) + expf(l2)) / 3.0f;
        const float sign = (mode == 1 ? 1.0f : 0.0f) - (mode == 2 ? 1.0f : 0.0f);
        const float isClone = mode == 3 ? 1.0f : 0.0f;
        float n[3];
        if (base_noise) { n[0] = base_noise[(size_t)j * 3]; n[1] = base_noise[(size_t)j * 3 + 1]; n[2] = base_noise[(size_t)j * 3 + 2]; }
        else normal3(seed, (uint32_t)j, n);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float splitNoise = ((sign * mean) * 0.1f) * n[c];
            const float cloneNoise = (isClone * 0.01f) * n[c];
            o_xyz[(size_t)j * 3 + c] = (xyz[(size_t)s * 3 + c] + splitNoise) + cloneNoise;
        }
    }
    const int R = (K - 1) * 3;
    for (int idx = threadIdx.x; idx < nj * R; idx += blockDim.x) {
        const int jj = idx / R, c = idx - jj * R;
        o_f_rest[(size_t)j0 * R + idx] = f_rest[(size_t)s_src[jj] * R + c];
    }
}

cudaError_t launch_densify_apply(cudaStream_t st, int Nout, int K, const int* gather, const int* noise_mode, const float* base_noise,
                                 uint64_t seed, const float* const* in6, float* const* out6)
{
    if (Nout > 0)
        k_densify_apply<<<cdiv(Nout, DG), 256, 0, st>>>(Nout, K, gather, noise_mode, base_noise, seed, in6[0], in6[1], in6[2], in6[3],
                                                        in6[4], in6[5], out6[0], out6[1], out6[2], out6[3], out6[4], out6[5]);
    return cudaGetLastError();
}

}  // namespace gsb
