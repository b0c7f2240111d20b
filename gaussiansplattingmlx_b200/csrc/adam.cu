// adam.cu — fused Adam over the six parameter tensors + D1 accum_grad_norm, one launch.
//
// Reference: MLXOptimizers.Adam.applySingle called once per tensor with a per-tensor learning
// rate (Trainer/GaussianTrainer.swift:941-948,1066-1079; mlx-swift 0.30.6, un-vendored — published
// update rule, NO bias correction):
//     m = b1*m + (1-b1)*g ;  v = b2*v + (1-b2)*g*g ;  p = p - lr * m / (sqrt(v) + eps)
// and accum_grad_norm (GaussianTrainer.swift:321-339): accum[i] += ||grad_xyz_i||_2.
// HBM-bound: 28 B per parameter float (read p,g,m,v; write p,m,v).  Compiled with --fmad=false so
// the update rounds exactly like the CPU oracle (bit-exact Adam parity).
#include <stdlib.h>

#include <algorithm>

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

// ------------------------------------------------------------------------------------------------
// Device-side synchronisation between the replicas (one process per GPU, flags in peer-mapped memory; PeerSync in
// kernels.h).  A flag holds the id of the last step for which its event happened; ids are compared modulo 2^32.
// Every wait is bounded (2 s of %globaltimer): a replica that never arrives sets the local error word instead of
// hanging the GPU.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void spin_until(const uint32_t* flag, uint32_t value, uint32_t* error)
{
    if ((int32_t)(ld_acquire_sys_u32(flag) - value) >= 0) return;
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys_u32(flag) - value) < 0) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > 2000000000ull) {
            if (error) *error = 1u;
            return;
        }
    }
}
// all threads of the CTA: returns when every replica's flag has reached sy.step
__device__ __forceinline__ void peer_wait_cta(const PeerStepSync& sy, int world)
{
    if (sy.wait_flags) {
        if ((int)threadIdx.x < world) spin_until(sy.wait_flags + threadIdx.x, sy.step, sy.error);
        __syncthreads();
    }
}
// all threads of the CTA, after their last peer store: the last CTA of the grid announces sy.step to every replica
__device__ __forceinline__ void peer_announce_cta(const PeerStepSync& sy, int world)
{
    if (!sy.done) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(sy.done, 1u);
        if (prev == gridDim.x - 1) {
            *sy.done = 0u;
            __threadfence_system();
            for (int r = 0; r < world; ++r) st_release_sys_u32(sy.announce[r], sy.step);
        }
    }
}

__global__ void k_peer_signal(const __grid_constant__ PeerFlagList flags, uint32_t value)
{
    __threadfence_system();
    if ((int)threadIdx.x < flags.n) st_release_sys_u32(flags.dst[threadIdx.x], value);
}
__global__ void k_peer_wait(const uint32_t* flags, int rows, int cols, int row_stride, uint32_t value, uint32_t* error)
{
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) spin_until(flags + (i / cols) * row_stride + (i % cols), value, error);
}
cudaError_t launch_peer_signal(cudaStream_t st, const PeerFlagList& flags, uint32_t value)
{
    if (flags.n > 0) k_peer_signal<<<1, 32, 0, st>>>(flags, value);
    return cudaGetLastError();
}
cudaError_t launch_peer_wait(cudaStream_t st, const uint32_t* flags, int rows, int cols, int row_stride, uint32_t value, uint32_t* error)
{
    if (rows * cols > 0) k_peer_wait<<<1, 64, 0, st>>>(flags, rows, cols, row_stride, value, error);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Data-parallel step fused with its collective: gradient reduction + Adam + parameter broadcast in ONE
// kernel over NVLink peer memory.  Rank r owns the Gaussians [g0, g1) of every tensor: it sums that slice of the gradient
// blocks of ALL ranks (peer loads, fixed rank order, so the sum has one owner and no replica can diverge), applies Adam
// with its local m / v / p, and stores the new parameters into every rank's replica (peer stores).  Per rank and step
// this moves (world - 1) / world of the gradient block in and of the parameter block out over NVLink - what a
// reduce-scatter + all-gather moves - and 1 / world of Adam's HBM traffic, with no staging buffers and no second pass.
// Link traffic per GPU and DIRECTION is twice that (206 MB at 8 GPUs and C3): its own pulls / pushes plus the pulls /
// pushes of the 7 other owners that target it - 412 MB, i.e. 0.72 ms = 570 GB/s per direction.
// Synchronisation: either the caller brackets the launch with two stream-ordered barriers (sync.wait_flags == nullptr,
// gsb_trainer_apply_peers), or the kernel itself waits for the replicas' "gradients complete" flags and its last CTA
// raises "parameters written" in every replica (gsb_trainer_step_peers: one launch per Gaussian chunk, so the exchange
// of chunk c overlaps the projection backward of chunk c + 1).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AD_THREADS, 4) k_adam_peers(const __grid_constant__ AdamTensors t, const __grid_constant__ AdamPeers pr,
                                                           const __grid_constant__ AdamSeg seg, const __grid_constant__ PeerStepSync sy,
                                                           float b1, float b2, float eps, float gscale)
{
    peer_wait_cta(sy, pr.world);
    const long long total = seg.vec_begin[6];
    const long long stride = (long long)gridDim.x * blockDim.x;
    // chunk i of the owned slice -> tensor k, float offset inside the tensor / inside one copy of the six tensors
    auto locate = [&](long long i, int& k, long long& base, long long& off) {
        k = 0;
#pragma unroll
        for (int j = 1; j < 6; ++j) k += (i >= seg.vec_begin[j]) ? 1 : 0;
        base = pr.first[k] + (i - seg.vec_begin[k]) * 4;   // the slice starts on a multiple of 4 floats
        off = pr.tensor_off[k] + base;
    };
    // Software pipeline: the peer loads of chunk i + stride are in flight while chunk i is updated and stored to every
    // replica, so NVLink carries gradients in and parameters out at the same time.
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4 gn[GSB_MAX_PEERS];
    bool vec_n = false;
    auto prefetch = [&](long long ii) {
        int k; long long base, off;
        locate(ii, k, base, off);
        vec_n = base + 4 <= pr.first[k] + t.count[k];
        if (vec_n) {
#pragma unroll
            for (int r = 0; r < GSB_MAX_PEERS; ++r)
                if (r < pr.world) gn[r] = __ldcg(reinterpret_cast<const float4*>(pr.grads[r] + off));
        }
    };
    if (i < total) prefetch(i);
    while (i < total) {
        int k; long long base, off;
        locate(i, k, base, off);
        const long long end = pr.first[k] + t.count[k];   // t.count = floats of the slice
        const float lr = t.lr[k];
        float* m = t.m[k] + base;
        float* v = t.v[k] + base;
        const bool vec = vec_n;
        float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
#pragma unroll
            for (int r = 0; r < GSB_MAX_PEERS; ++r)
                if (r < pr.world) { G.x += gn[r].x; G.y += gn[r].y; G.z += gn[r].z; G.w += gn[r].w; }
        }
        const long long inext = i + stride;
        if (inext < total) prefetch(inext);
        if (vec) {
            float4 P = *reinterpret_cast<const float4*>(t.p[k] + base);
            float4 M = *reinterpret_cast<float4*>(m);
            float4 V = *reinterpret_cast<float4*>(v);
            G.x = G.x * gscale; G.y = G.y * gscale; G.z = G.z * gscale; G.w = G.w * gscale;
            adam1(P.x, G.x, M.x, V.x, lr, b1, b2, eps);
            adam1(P.y, G.y, M.y, V.y, lr, b1, b2, eps);
            adam1(P.z, G.z, M.z, V.z, lr, b1, b2, eps);
            adam1(P.w, G.w, M.w, V.w, lr, b1, b2, eps);
            *reinterpret_cast<float4*>(m) = M;
            *reinterpret_cast<float4*>(v) = V;
            for (int r = 0; r < pr.world; ++r) *reinterpret_cast<float4*>(pr.params[r] + off) = P;
        } else {
            for (long long e = 0; base + e < end; ++e) {
                float g = 0.f;
                for (int r = 0; r < pr.world; ++r) g += __ldcg(pr.grads[r] + off + e);
                float pp = t.p[k][base + e], mm = m[e], vv = v[e];
                adam1(pp, g * gscale, mm, vv, lr, b1, b2, eps);
                m[e] = mm; v[e] = vv;
                for (int r = 0; r < pr.world; ++r) pr.params[r][off + e] = pp;
            }
        }
        i = inext;
    }
    // D1 for the owned Gaussians: the norm of the SUMMED position gradient, written to every replica's accumulator
    if (pr.accum[0]) {
        for (long long i = pr.g0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pr.g1; i += stride) {
            float a = 0.f, b = 0.f, c = 0.f;
            for (int r = 0; r < pr.world; ++r) {
                const float* gx = pr.grads[r] + pr.tensor_off[0] + i * 3;
                a += __ldcg(gx); b += __ldcg(gx + 1); c += __ldcg(gx + 2);
            }
            a *= gscale; b *= gscale; c *= gscale;
            const float acc = pr.accum[pr.rank][i] + sqrtf(a * a + b * b + c * c);
            for (int r = 0; r < pr.world; ++r) pr.accum[r][i] = acc;
        }
    }
    peer_announce_cta(sy, pr.world);
}

cudaError_t launch_adam_peers(cudaStream_t st, const AdamTensors& t, const AdamPeers& pr, float beta1, float beta2, float eps,
                              float gscale, const PeerStepSync* sync, int blocks_override, int* launches)
{
    AdamSeg seg;
    seg.vec_begin[0] = 0;
    for (int k = 0; k < 6; ++k) seg.vec_begin[k + 1] = seg.vec_begin[k] + (t.count[k] + 3) / 4;
    const long long total = seg.vec_begin[6];
    if (!sync && total == 0 && pr.g1 <= pr.g0) return cudaSuccess;   // with sync even an empty slice must announce itself
    long long blocks = (std::max<long long>(total, 1) + AD_THREADS - 1) / AD_THREADS;
    // persistent: every thread pipelines its chunks.  One resident wave of 4 CTAs per SM when the launch is alone on the GPU;
    // 2 per SM under the step protocol, where the projection backward of the next Gaussian chunk must fit beside it (the
    // kernel is NVLink-bound: 296 CTAs keep ~10 MB of peer loads in flight)
    long long cap = sync ? 148LL * 2 : 148LL * 4;
    if (blocks_override > 0) cap = blocks_override;   // gsb_trainer_peers_tune
    if (blocks > cap) blocks = cap;
    PeerStepSync sy{};
    if (sync) sy = *sync;
    k_adam_peers<<<(int)blocks, AD_THREADS, 0, st>>>(t, pr, seg, sy, beta1, beta2, eps, gscale);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// The same step through the NVSwitch (NVLS): the gradient and parameter blocks of all replicas are bound to ONE multicast
// address range (symmetric memory).  multimem.ld_reduce returns the SUM over all replicas of a 16-byte chunk - the switch
// adds, one read crosses this GPU's links - and multimem.st writes a chunk into every replica.  Rank r still owns slice r:
// per step it pulls 1 / world of the gradient block (already reduced) and pushes 1 / world of the parameter block, instead
// of (world - 1) / world each with plain peer loads / stores: ~265 MB per direction and GPU instead of 412 MB at 8 GPUs.
// Four 16-byte reductions are in flight per thread before the first is consumed (the switch round trip is several
// microseconds).  D1 is computed for the owned Gaussians only and stored into every replica's accumulator through peer
// memory (one owner per value: replicas cannot diverge by the summation order of the switch).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mc_ld_reduce_v4(const float* mc)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ float mc_ld_reduce(const float* mc)
{
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_v4(float* mc, const float4& v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st(float* mc, float v)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}

constexpr int MC_UNROLL = 4;

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_adam_multicast(const __grid_constant__ AdamTensors t, const __grid_constant__ AdamPeers pr,
                                                               const __grid_constant__ AdamSeg seg, const __grid_constant__ PeerStepSync sy,
                                                               const float* __restrict__ mc_grads, float* __restrict__ mc_params, float b1,
                                                               float b2, float eps, float gscale, int N)
{
    peer_wait_cta(sy, pr.world);
    const long long total = seg.vec_begin[6];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * MC_UNROLL) {
        int kk[MC_UNROLL];
        long long bb[MC_UNROLL], oo[MC_UNROLL];
        bool vec[MC_UNROLL], live[MC_UNROLL];
        float4 G[MC_UNROLL];
#pragma unroll
        for (int u = 0; u < MC_UNROLL; ++u) {
            const long long i = i0 + (long long)u * stride;
            live[u] = i < total;
            vec[u] = false;
            if (live[u]) {
                int k = 0;
#pragma unroll
                for (int j = 1; j < 6; ++j) k += (i >= seg.vec_begin[j]) ? 1 : 0;
                kk[u] = k;
                bb[u] = pr.first[k] + (i - seg.vec_begin[k]) * 4;
                oo[u] = pr.tensor_off[k] + bb[u];
                vec[u] = bb[u] + 4 <= pr.first[k] + t.count[k];
                if (vec[u]) G[u] = mc_ld_reduce_v4(mc_grads + oo[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < MC_UNROLL; ++u) {
            if (!live[u]) continue;
            const int k = kk[u];
            const long long base = bb[u], off = oo[u];
            const float lr = t.lr[k];
            float* m = t.m[k] + base;
            float* v = t.v[k] + base;
            if (vec[u]) {
                float4 Gu = G[u];
                float4 P = *reinterpret_cast<const float4*>(t.p[k] + base);
                float4 M = *reinterpret_cast<float4*>(m);
                float4 V = *reinterpret_cast<float4*>(v);
                Gu.x = Gu.x * gscale; Gu.y = Gu.y * gscale; Gu.z = Gu.z * gscale; Gu.w = Gu.w * gscale;
                adam1(P.x, Gu.x, M.x, V.x, lr, b1, b2, eps);
                adam1(P.y, Gu.y, M.y, V.y, lr, b1, b2, eps);
                adam1(P.z, Gu.z, M.z, V.z, lr, b1, b2, eps);
                adam1(P.w, Gu.w, M.w, V.w, lr, b1, b2, eps);
                *reinterpret_cast<float4*>(m) = M;
                *reinterpret_cast<float4*>(v) = V;
                mc_st_v4(mc_params + off, P);
            } else {
                const long long end = pr.first[k] + t.count[k];
                for (long long e = 0; base + e < end; ++e) {
                    const float g = mc_ld_reduce(mc_grads + off + e);
                    float pp = t.p[k][base + e], mm = m[e], vv = v[e];
                    adam1(pp, g * gscale, mm, vv, lr, b1, b2, eps);
                    m[e] = mm; v[e] = vv;
                    mc_st(mc_params + off + e, pp);
                }
            }
        }
    }
    const float* gx = mc_grads + pr.tensor_off[0];
    if (pr.accum[pr.rank] && pr.accum_all) {
        // the replicas' accumulators are peer-mapped: D1 of the owned Gaussians, one owner per value
        for (long long i = pr.g0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pr.g1; i += stride) {
            const float a = mc_ld_reduce(gx + i * 3) * gscale, b = mc_ld_reduce(gx + i * 3 + 1) * gscale, c = mc_ld_reduce(gx + i * 3 + 2) * gscale;
            const float acc = pr.accum[pr.rank][i] + sqrtf(a * a + b * b + c * c);
            for (int r = 0; r < pr.world; ++r) pr.accum[r][i] = acc;
        }
    } else if (pr.accum[pr.rank]) {
        // no peer mapping of the accumulators: every replica computes all of D1 from the switch-reduced gradient
        float* accum = pr.accum[pr.rank];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
            const float x = mc_ld_reduce(gx + i * 3) * gscale, y = mc_ld_reduce(gx + i * 3 + 1) * gscale, z = mc_ld_reduce(gx + i * 3 + 2) * gscale;
            accum[i] = accum[i] + sqrtf(x * x + y * y + z * z);
        }
    }
    peer_announce_cta(sy, pr.world);
}

cudaError_t launch_adam_multicast(cudaStream_t st, const AdamTensors& t, const AdamPeers& pr, const float* mc_grads, float* mc_params,
                                  float beta1, float beta2, float eps, float gscale, int N, const PeerStepSync* sync, int blocks_override,
                                  int* launches)
{
    AdamSeg seg;
    seg.vec_begin[0] = 0;
    for (int k = 0; k < 6; ++k) seg.vec_begin[k + 1] = seg.vec_begin[k] + (t.count[k] + 3) / 4;
    const bool own_d1 = pr.accum[pr.rank] && pr.accum_all;
    const long long d1 = pr.accum[pr.rank] ? (own_d1 ? pr.g1 - pr.g0 : N) : 0;
    const long long total = std::max<long long>(seg.vec_begin[6], d1);
    // Few, fat CTAs: the NVSwitch, not the SMs, sets the pace (148 x 256 threads 0.65 ms, 1184 x 256 0.70 ms per step at 8
    // GPUs, profiles/r2/r2e_exchange_8gpu.jsonl), and every SM that hosts a CTA of this kernel has its memory pipeline full
    // of multi-microsecond multimem requests - the projection / binning kernels of the next step that share those SMs slow
    // down 3-4x (r2r).  32 CTAs of 1024 threads keep the same bytes in flight on 32 SMs and leave the other 116 alone.
    static const int env_threads = getenv("GSB_MC_THREADS") ? atoi(getenv("GSB_MC_THREADS")) : 0;
    // (2 replicas: each owns half of Adam's state, the launch is bound by LOCAL HBM traffic and wants every SM:
    // 148 x 256 threads 7.32 ms per strong-scaling step against 7.59 ms with 32 x 1024, profiles/r2/r2q vs r2z)
    const bool wide = pr.world <= 2;
    const int threads = env_threads > 0 ? std::min(1024, (env_threads + 31) & ~31) : (wide ? 256 : 1024);
    long long blocks = (std::max<long long>((total + MC_UNROLL - 1) / MC_UNROLL, 1) + threads - 1) / threads;
    long long cap = wide ? 148 : 32;
    if (blocks_override > 0) cap = blocks_override;   // gsb_trainer_peers_tune
    if (blocks > cap) blocks = cap;
    PeerStepSync sy{};
    if (sync) sy = *sync;
    if (threads <= 256)   // compiled for 256 threads: no register cap, no spills
        k_adam_multicast<256><<<(int)blocks, threads, 0, st>>>(t, pr, seg, sy, mc_grads, mc_params, beta1, beta2, eps, gscale, N);
    else
        k_adam_multicast<1024><<<(int)blocks, threads, 0, st>>>(t, pr, seg, sy, mc_grads, mc_params, beta1, beta2, eps, gscale, N);
    if (launches) ++*launches;
    return cudaGetLastError();
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
