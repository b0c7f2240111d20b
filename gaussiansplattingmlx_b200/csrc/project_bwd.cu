// project_bwd.cu — backward of the EWA projection + SH colour (K2 of the reference) and of the activations, sm_100a.
//
// Reference: slang/gaussian_projection_kernels.slang:205-398 (Slang reverse-AD of K1; derivative rules for max / clamp /
// sqrt restated in project_math.cuh), Swift VJP Trainer/GaussianRenderer.swift:605-701, activation VJPs :936-963.
//
// Its own translation unit so that it is compiled WITH FMA contraction: the forward (project.cu) must round like the
// -ffp-contract=off CPU oracle to keep the tile lists bit-exact, the backward only has to meet 1e-3 relative and was
// instruction-bound (~2 800 instructions per Gaussian without contraction).  gsb_expf uses explicit _rn intrinsics, so the
// activations recomputed here are the forward's bit for bit.
#include "kernels.h"
#include "project_math.cuh"
#include "project_stage.cuh"

namespace gsb {

__device__ __forceinline__ void rot_activation_bwd(const float* q, const float* gq, float* out)
{
    float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    float n = sqrtf(n2);
    float d = n + 1e-8f;
    float dot = gq[0] * q[0] + gq[1] * q[1] + gq[2] * q[2] + gq[3] * q[3];
    float gd = -dot / (d * d);
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = gq[c] / d + (n > 0.0f ? gd * q[c] / n : 0.0f);
}

__global__ void k_activate_bwd(int N, int K, const float* __restrict__ scales_log, const float* __restrict__ rot_raw,
                               const float* __restrict__ op_logit, const float* __restrict__ g_shs,
                               const float* __restrict__ g_scales, const float* __restrict__ g_rot,
                               const float* __restrict__ g_op, float* __restrict__ g_f_dc, float* __restrict__ g_f_rest,
                               float* __restrict__ g_scales_log, float* __restrict__ g_rot_raw, float* __restrict__ g_op_logit)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    for (int c = 0; c < 3; ++c) g_f_dc[(size_t)p * 3 + c] = g_shs[(size_t)p * K * 3 + c];
    for (int j = 0; j < (K - 1) * 3; ++j) g_f_rest[(size_t)p * (K - 1) * 3 + j] = g_shs[(size_t)p * K * 3 + 3 + j];
    for (int c = 0; c < 3; ++c) g_scales_log[p * 3 + c] = g_scales[p * 3 + c] * gsb_expf(scales_log[p * 3 + c]);
    float q[4] = {rot_raw[p * 4], rot_raw[p * 4 + 1], rot_raw[p * 4 + 2], rot_raw[p * 4 + 3]};
    float gq[4] = {g_rot[p * 4], g_rot[p * 4 + 1], g_rot[p * 4 + 2], g_rot[p * 4 + 3]};
    float out[4];
    rot_activation_bwd(q, gq, out);
    for (int c = 0; c < 4; ++c) g_rot_raw[p * 4 + c] = out[c];
    float s = 1.0f / (1.0f + gsb_expf(-op_logit[p]));
    g_op_logit[p] = g_op[p] * s * (1.0f - s);
}

template <int MAXK>
__global__ void __launch_bounds__(128) k_project_bwd_api(
    int N, const float* __restrict__ scales, const float* __restrict__ rotations, const float* __restrict__ means3d,
    const float* __restrict__ shs, const __grid_constant__ ViewParams vp, const float* __restrict__ cotDepths,
    const float* __restrict__ cotMeans2d, const float* __restrict__ cotCov2d, const float* __restrict__ cotColor,
    const float* __restrict__ cotConic, float* __restrict__ gScales, float* __restrict__ gRot,
    float* __restrict__ gMeans3d, float* __restrict__ gShs, float* __restrict__ gCam)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float* shp = shs + (size_t)p * vp.K * 3;
    float* gshp = gShs + (size_t)p * vp.K * 3;
    float cc2[4] = {cotCov2d[p * 4], cotCov2d[p * 4 + 1], cotCov2d[p * 4 + 2], cotCov2d[p * 4 + 3]};
    float ccol[3] = {cotColor[p * 3], cotColor[p * 3 + 1], cotColor[p * 3 + 2]};
    float ccon[4] = {cotConic[p * 4], cotConic[p * 4 + 1], cotConic[p * 4 + 2], cotConic[p * 4 + 3]};
    ProjGrad g;
    project_backward<MAXK>(means3d[p * 3], means3d[p * 3 + 1], means3d[p * 3 + 2], scales[p * 3], scales[p * 3 + 1],
                           scales[p * 3 + 2], rotations[p * 4], rotations[p * 4 + 1], rotations[p * 4 + 2],
                           rotations[p * 4 + 3], [&](int k, int c) { return shp[k * 3 + c]; }, vp, cotDepths[p],
                           cotMeans2d[p * 2], cotMeans2d[p * 2 + 1], cc2, ccol, ccon,
                           [&](int k, int c, float v) { gshp[k * 3 + c] = v; }, g);
    for (int c = 0; c < 3; ++c) { gScales[p * 3 + c] = g.gs[c]; gMeans3d[p * 3 + c] = g.gm[c]; gCam[p * 3 + c] = g.gcam[c]; }
    for (int c = 0; c < 4; ++c) gRot[p * 4 + c] = g.gr[c];
}

template <int MAXK>
__global__ void __launch_bounds__(PB) k_project_fused_bwd(
    int N, const float* __restrict__ xyz, const float* __restrict__ f_dc, const float* __restrict__ f_rest,
    const float* __restrict__ scales_log, const float* __restrict__ rot_raw, const float* __restrict__ op_logit,
    const __grid_constant__ ViewParams vp, int tma_ok, const float* __restrict__ grad_rec, float* __restrict__ g_xyz,
    float* __restrict__ g_f_dc, float* __restrict__ g_f_rest, float* __restrict__ g_scales, float* __restrict__ g_rot,
    float* __restrict__ g_op, int accumulate)
{
    extern __shared__ __align__(128) float sm[];
    const int K = vp.K;
    const FusedSmem L = fused_layout(K);
    const int base = blockIdx.x * PB;
    const int count = min(PB, N - base);
    const bool tma = tma_ok && count == PB;
    const int restF = (K - 1) * 3;
    stage_inputs(sm, L, base, count, K, tma, xyz, f_dc, f_rest, scales_log, rot_raw, op_logit, grad_rec, REC_FLOATS);

    const int t = threadIdx.x;
    if (t < count) {
        float sl0 = sm[L.scales + t * 3], sl1 = sm[L.scales + t * 3 + 1], sl2 = sm[L.scales + t * 3 + 2];
        float s0 = gsb_expf(sl0), s1 = gsb_expf(sl1), s2 = gsb_expf(sl2);
        float4 q = *reinterpret_cast<const float4*>(sm + L.rot + t * 4);
        float qr[4] = {q.x, q.y, q.z, q.w};
        float d = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w) + 1e-8f;
        float rw = q.x / d, rx = q.y / d, ry = q.z / d, rz = q.w / d;
        float sig = 1.0f / (1.0f + gsb_expf(-sm[L.op + t]));
        float* dc = sm + L.fdc + t * 3;
        float* rest = sm + L.frest + t * restF;
        const float4* gr4 = reinterpret_cast<const float4*>(sm + L.rec + t * REC_FLOATS);
        float4 ga = gr4[0], gb = gr4[1], gc = gr4[2];
        float cotConic[4] = {ga.z, ga.w, gb.x, gb.y};
        float cotColor[3] = {gb.z, gb.w, gc.x};
        float cotCov2d[4] = {0.f, 0.f, 0.f, 0.f};
        float m0 = sm[L.xyz + t * 3], m1 = sm[L.xyz + t * 3 + 1], m2 = sm[L.xyz + t * 3 + 2];
        ProjGrad g;
        project_backward<MAXK>(m0, m1, m2, s0, s1, s2, rw, rx, ry, rz,
                               [&](int k, int c) { return k == 0 ? dc[c] : rest[(k - 1) * 3 + c]; }, vp, gc.z, ga.x, ga.y,
                               cotCov2d, cotColor, cotConic,
                               [&](int k, int c, float v) {
                                   if (k == 0) dc[c] = v; else rest[(k - 1) * 3 + c] = v;
                               }, g);
        // SH slots beyond the active coefficient count carry zero gradient
        for (int k = vp.coeffCount; k < K; ++k)
            for (int c = 0; c < 3; ++c) rest[(k - 1) * 3 + c] = 0.0f;
        // activation VJPs, written in place over this thread's own input slots
        sm[L.xyz + t * 3] = g.gm[0]; sm[L.xyz + t * 3 + 1] = g.gm[1]; sm[L.xyz + t * 3 + 2] = g.gm[2];
        sm[L.scales + t * 3] = g.gs[0] * s0; sm[L.scales + t * 3 + 1] = g.gs[1] * s1; sm[L.scales + t * 3 + 2] = g.gs[2] * s2;
        float gq[4];
        rot_activation_bwd(qr, g.gr, gq);
        *reinterpret_cast<float4*>(sm + L.rot + t * 4) = make_float4(gq[0], gq[1], gq[2], gq[3]);
        sm[L.op + t] = gc.y * sig * (1.0f - sig);
    }
    if (tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (accumulate) {
                bulk_s2g_add_f32(g_xyz + (size_t)base * 3, sm + L.xyz, PB * 3 * 4);
                bulk_s2g_add_f32(g_scales + (size_t)base * 3, sm + L.scales, PB * 3 * 4);
                bulk_s2g_add_f32(g_rot + (size_t)base * 4, sm + L.rot, PB * 4 * 4);
                bulk_s2g_add_f32(g_op + (size_t)base, sm + L.op, PB * 4);
                bulk_s2g_add_f32(g_f_dc + (size_t)base * 3, sm + L.fdc, PB * 3 * 4);
                if (restF > 0) bulk_s2g_add_f32(g_f_rest + (size_t)base * restF, sm + L.frest, PB * restF * 4);
            } else {
                bulk_s2g(g_xyz + (size_t)base * 3, sm + L.xyz, PB * 3 * 4);
                bulk_s2g(g_scales + (size_t)base * 3, sm + L.scales, PB * 3 * 4);
                bulk_s2g(g_rot + (size_t)base * 4, sm + L.rot, PB * 4 * 4);
                bulk_s2g(g_op + (size_t)base, sm + L.op, PB * 4);
                bulk_s2g(g_f_dc + (size_t)base * 3, sm + L.fdc, PB * 3 * 4);
                if (restF > 0) bulk_s2g(g_f_rest + (size_t)base * restF, sm + L.frest, PB * restF * 4);
            }
            bulk_commit();
            bulk_wait_all_read();
        }
    } else {
        __syncthreads();
        auto put = [&](float* dst, int off, int n) {
            for (int i = threadIdx.x; i < n; i += PB) dst[i] = accumulate ? dst[i] + sm[off + i] : sm[off + i];
        };
        put(g_xyz + (size_t)base * 3, L.xyz, count * 3);
        put(g_scales + (size_t)base * 3, L.scales, count * 3);
        put(g_rot + (size_t)base * 4, L.rot, count * 4);
        put(g_op + (size_t)base, L.op, count);
        put(g_f_dc + (size_t)base * 3, L.fdc, count * 3);
        put(g_f_rest + (size_t)base * restF, L.frest, count * restF);
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_activate_bwd(cudaStream_t st, int N, int K, const float* scales_log, const float* rot_raw,
                                const float* op_logit, const float* g_shs, const float* g_scales, const float* g_rot,
                                const float* g_op, float* g_f_dc, float* g_f_rest, float* g_scales_log, float* g_rot_raw,
                                float* g_op_logit)
{
    if (N > 0)
        k_activate_bwd<<<cdiv(N, 256), 256, 0, st>>>(N, K, scales_log, rot_raw, op_logit, g_shs, g_scales, g_rot, g_op,
                                                     g_f_dc, g_f_rest, g_scales_log, g_rot_raw, g_op_logit);
    return cudaGetLastError();
}

cudaError_t launch_project_bwd_api(cudaStream_t st, int N, const ViewParams& vp, const float* scales,
                                   const float* rotations, const float* means3d, const float* shs,
                                   const float* cotDepths, const float* cotMeans2d, const float* cotCov2d,
                                   const float* cotColor, const float* cotConic, float* gScales, float* gRot,
                                   float* gMeans3d, float* gShs, float* gCam)
{
    if (N <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(gShs, 0, (size_t)N * vp.K * 3 * sizeof(float), st);  // initValue 0 (GaussianRenderer.swift:674)
    if (e != cudaSuccess) return e;
    if (vp.coeffCount > 16)
        k_project_bwd_api<25><<<cdiv(N, 128), 128, 0, st>>>(N, scales, rotations, means3d, shs, vp, cotDepths, cotMeans2d, cotCov2d, cotColor, cotConic, gScales, gRot, gMeans3d, gShs, gCam);
    else
        k_project_bwd_api<16><<<cdiv(N, 128), 128, 0, st>>>(N, scales, rotations, means3d, shs, vp, cotDepths, cotMeans2d, cotCov2d, cotColor, cotConic, gScales, gRot, gMeans3d, gShs, gCam);
    return cudaGetLastError();
}

cudaError_t launch_project_fused_bwd(cudaStream_t st, int N, const ViewParams& vp, const float* xyz, const float* f_dc,
                                     const float* f_rest, const float* scales_log, const float* rot_raw,
                                     const float* op_logit, const float* grad_rec, float* g_xyz, float* g_f_dc,
                                     float* g_f_rest, float* g_scales, float* g_rot, float* g_op, int accumulate)
{
    if (N <= 0) return cudaSuccess;
    int tma_ok = aligned16(xyz) && aligned16(f_dc) && aligned16(f_rest) && aligned16(scales_log) && aligned16(rot_raw) &&
                 aligned16(op_logit) && aligned16(grad_rec) && aligned16(g_xyz) && aligned16(g_f_dc) && aligned16(g_f_rest) &&
                 aligned16(g_scales) && aligned16(g_rot) && aligned16(g_op);
    size_t smem = fused_smem_bytes(vp.K);
    cudaError_t e;
    if (vp.coeffCount > 16 || vp.K > 16) {
        e = cudaFuncSetAttribute(k_project_fused_bwd<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_project_fused_bwd<25><<<cdiv(N, PB), PB, smem, st>>>(N, xyz, f_dc, f_rest, scales_log, rot_raw, op_logit, vp, tma_ok, grad_rec, g_xyz, g_f_dc, g_f_rest, g_scales, g_rot, g_op, accumulate);
    } else {
        e = cudaFuncSetAttribute(k_project_fused_bwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_project_fused_bwd<16><<<cdiv(N, PB), PB, smem, st>>>(N, xyz, f_dc, f_rest, scales_log, rot_raw, op_logit, vp, tma_ok, grad_rec, g_xyz, g_f_dc, g_f_rest, g_scales, g_rot, g_op, accumulate);
    }
    return cudaGetLastError();
}

}  // namespace gsb
