// project.cu — EWA 3D→2D projection + SH colour (K1/K2 of the reference) for sm_100a.
// Compiled with --fmad=false (see project_math.cuh for why); the backward kernels live in project_bwd.cu, which is not.
//
// Two families of kernels:
//  * k_project_{fwd,bwd}_api   reference-layout I/O on ACTIVATED tensors (parity API gsb_project_fwd/bwd)
//  * k_project_fused_{fwd,bwd} the training path: RAW model tensors in, activations fused, one 48-byte
//    projected record per Gaussian out.  A CTA owns 128 consecutive Gaussians, whose bytes form one
//    contiguous range in each of the six AoS parameter tensors: they are staged into shared memory with
//    TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) and the results leave through bulk stores /
//    bulk reduce-adds (gradient accumulation across views happens in L2, not in the SM).
#include "kernels.h"
#include "project_math.cuh"
#include "project_stage.cuh"

namespace gsb {


// ------------------------------------------------------------------------------------------------
// activations (GaussianRenderer.swift:936-963)
// ------------------------------------------------------------------------------------------------
__global__ void k_activate_fwd(int N, int K, const float* __restrict__ f_dc, const float* __restrict__ f_rest,
                               const float* __restrict__ scales_log, const float* __restrict__ rot_raw,
                               const float* __restrict__ op_logit, float* __restrict__ shs, float* __restrict__ scales,
                               float* __restrict__ rotations, float* __restrict__ opacity)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    for (int c = 0; c < 3; ++c) shs[(size_t)p * K * 3 + c] = f_dc[(size_t)p * 3 + c];
    for (int j = 0; j < (K - 1) * 3; ++j) shs[(size_t)p * K * 3 + 3 + j] = f_rest[(size_t)p * (K - 1) * 3 + j];
    for (int c = 0; c < 3; ++c) scales[p * 3 + c] = gsb_expf(scales_log[p * 3 + c]);
    float q0 = rot_raw[p * 4], q1 = rot_raw[p * 4 + 1], q2 = rot_raw[p * 4 + 2], q3 = rot_raw[p * 4 + 3];
    float d = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3) + 1e-8f;
    rotations[p * 4] = q0 / d; rotations[p * 4 + 1] = q1 / d; rotations[p * 4 + 2] = q2 / d; rotations[p * 4 + 3] = q3 / d;
    opacity[p] = 1.0f / (1.0f + gsb_expf(-op_logit[p]));
}

// ------------------------------------------------------------------------------------------------
// parity API kernels (reference I/O layout)
// ------------------------------------------------------------------------------------------------
template <int MAXK>
__global__ void __launch_bounds__(128) k_project_fwd_api(int N, const float* __restrict__ scales,
                                                         const float* __restrict__ rotations,
                                                         const float* __restrict__ means3d, const float* __restrict__ shs,
                                                         const __grid_constant__ ViewParams vp, float* __restrict__ means2d,
                                                         float* __restrict__ depths, float* __restrict__ color,
                                                         float* __restrict__ cov2d, float* __restrict__ conic,
                                                         float* __restrict__ radii, float* __restrict__ rectMin,
                                                         float* __restrict__ rectMax)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const float* shp = shs + (size_t)p * vp.K * 3;
    ProjOut o;
    project_forward<MAXK>(means3d[p * 3], means3d[p * 3 + 1], means3d[p * 3 + 2], scales[p * 3], scales[p * 3 + 1],
                          scales[p * 3 + 2], rotations[p * 4], rotations[p * 4 + 1], rotations[p * 4 + 2],
                          rotations[p * 4 + 3], [&](int k, int c) { return shp[k * 3 + c]; }, vp, o);
    means2d[p * 2] = o.sx; means2d[p * 2 + 1] = o.sy;
    depths[p] = o.depth;
    for (int c = 0; c < 3; ++c) color[p * 3 + c] = o.color[c];
    for (int c = 0; c < 4; ++c) { cov2d[p * 4 + c] = o.cov2d[c]; conic[p * 4 + c] = o.conic[c]; }
    radii[p] = o.radius;
    rectMin[p * 2] = o.rect[0]; rectMin[p * 2 + 1] = o.rect[1];
    rectMax[p * 2] = o.rect[2]; rectMax[p * 2 + 1] = o.rect[3];
}

// ------------------------------------------------------------------------------------------------
// fused training-path kernels (staging helpers: project_stage.cuh)
// ------------------------------------------------------------------------------------------------
// PART 3: the whole forward.  PART 1: geometry only (positions, scales, rotations, opacities in; record with a zero colour,
// tile rect, tiles-touched, depth key out).  PART 2: colour only (positions + SH coefficients in; the three colour floats of
// the record out).  The data-parallel step (api.cu gsb_trainer_step_peers) runs 1, the binning, then 2: the SH parameters
// - 81 % of the bytes of the previous step's parameter exchange - are only needed by 2.
template <int MAXK, int PART>
__global__ void __launch_bounds__(PB) k_project_fused_fwd(int N, const float* __restrict__ xyz,
                                                          const float* __restrict__ f_dc, const float* __restrict__ f_rest,
                                                          const float* __restrict__ scales_log,
                                                          const float* __restrict__ rot_raw,
                                                          const float* __restrict__ op_logit,
                                                          const __grid_constant__ ViewParams vp, int tma_ok,
                                                          float* __restrict__ rec, uint2* __restrict__ tile_rects,
                                                          uint32_t* __restrict__ touched, uint32_t* __restrict__ depth_keys,
                                                          float* __restrict__ radii_out,
                                                          uint8_t* __restrict__ vis_out)
{
    extern __shared__ __align__(128) float sm[];
    const int K = vp.K;
    const FusedSmem L = fused_layout(K);
    const int base = blockIdx.x * PB;
    const int count = min(PB, N - base);
    const bool tma = tma_ok && count == PB;
    stage_inputs(sm, L, base, count, K, tma, xyz, f_dc, f_rest, scales_log, rot_raw, op_logit, nullptr, 0,
                 STAGE_XYZ | ((PART & 1) ? STAGE_GEO : 0) | ((PART & 2) ? STAGE_SH : 0));

    const int t = threadIdx.x;
    if (PART == 2) {
        if (t < count) {
            const float* dc = sm + L.fdc + t * 3;
            const float* rest = sm + L.frest + t * (K - 1) * 3;
            float col[3];
            project_color<MAXK>(sm[L.xyz + t * 3], sm[L.xyz + t * 3 + 1], sm[L.xyz + t * 3 + 2],
                                [&](int k, int c) { return k == 0 ? dc[c] : rest[(k - 1) * 3 + c]; }, vp, col);
            float* r = rec + (size_t)(base + t) * REC_FLOATS;   // record floats 6, 7, 8 = r, g, b (common.cuh)
            r[6] = col[0]; r[7] = col[1]; r[8] = col[2];
        }
        return;
    }
    if (t < count) {
        const int p = base + t;
        float s0 = gsb_expf(sm[L.scales + t * 3]), s1 = gsb_expf(sm[L.scales + t * 3 + 1]), s2 = gsb_expf(sm[L.scales + t * 3 + 2]);
        float4 q = *reinterpret_cast<const float4*>(sm + L.rot + t * 4);
        float d = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w) + 1e-8f;
        float rw = q.x / d, rx = q.y / d, ry = q.z / d, rz = q.w / d;
        float opac = 1.0f / (1.0f + gsb_expf(-sm[L.op + t]));
        const float* dc = sm + L.fdc + t * 3;
        const float* rest = sm + L.frest + t * (K - 1) * 3;
        ProjOut o;
        o.color[0] = o.color[1] = o.color[2] = 0.0f;
        project_forward<MAXK, PART>(sm[L.xyz + t * 3], sm[L.xyz + t * 3 + 1], sm[L.xyz + t * 3 + 2], s0, s1, s2, rw, rx, ry, rz,
                                    [&](int k, int c) { return k == 0 ? dc[c] : rest[(k - 1) * 3 + c]; }, vp, o);
        float4* r4 = reinterpret_cast<float4*>(sm + L.rec + t * REC_FLOATS);
        make_raster_record(o.sx, o.sy, o.conic[0], o.conic[1], o.conic[2], o.conic[3], o.color[0], o.color[1], o.color[2],
                           opac, o.depth, (uint32_t)p, r4);
        int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
        uint32_t cnt = 0;
        if (o.radius > 0.0f) {
            tile_rect(o.rect, vp, x0, y0, x1, y1);
            cnt = (uint32_t)((x1 - x0) * (y1 - y0));
        }
        if (cnt == 0) { x0 = y0 = x1 = y1 = 0; }
        tile_rects[p] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
        touched[p] = cnt;
        // Gaussians without tiles emit no pairs, so their place in the depth order is irrelevant: they keep a key in the
        // range of the others (|depth| clamped to the near plane) instead of 0xffffffff, which would make the top digit
        // place non-trivial and cost the depth sort a fourth pass
        depth_keys[p] = __float_as_uint(cnt ? o.depth : fminf(fmaxf(fabsf(o.depth), 0.2f), 3.0e38f));
        if (radii_out) radii_out[p] = o.radius;
        if (vis_out) vis_out[p] = o.radius > 0.0f ? 1 : 0;
    }
    // records leave through one bulk store per CTA (48 B per record keeps every size a multiple of 16)
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_s2g(rec + (size_t)base * REC_FLOATS, sm + L.rec, (uint32_t)(count * REC_FLOATS * 4));
        bulk_commit();
        bulk_wait_all_read();
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------

cudaError_t launch_activate_fwd(cudaStream_t st, int N, int K, const float* f_dc, const float* f_rest,
                                const float* scales_log, const float* rot_raw, const float* op_logit, float* shs,
                                float* scales, float* rotations, float* opacity)
{
    if (N > 0) k_activate_fwd<<<cdiv(N, 256), 256, 0, st>>>(N, K, f_dc, f_rest, scales_log, rot_raw, op_logit, shs, scales, rotations, opacity);
    return cudaGetLastError();
}

cudaError_t launch_project_fwd_api(cudaStream_t st, int N, const ViewParams& vp, const float* scales,
                                   const float* rotations, const float* means3d, const float* shs, float* means2d,
                                   float* depths, float* color, float* cov2d, float* conic, float* radii, float* rectMin,
                                   float* rectMax)
{
    if (N <= 0) return cudaSuccess;
    if (vp.coeffCount > 16)
        k_project_fwd_api<25><<<cdiv(N, 128), 128, 0, st>>>(N, scales, rotations, means3d, shs, vp, means2d, depths, color, cov2d, conic, radii, rectMin, rectMax);
    else
        k_project_fwd_api<16><<<cdiv(N, 128), 128, 0, st>>>(N, scales, rotations, means3d, shs, vp, means2d, depths, color, cov2d, conic, radii, rectMin, rectMax);
    return cudaGetLastError();
}

size_t project_fused_smem_bytes(int K) { return (size_t)fused_layout(K).total * sizeof(float); }

template <int MAXK, int PART>
static cudaError_t launch_fused_fwd_t(cudaStream_t st, int N, const ViewParams& vp, const float* xyz, const float* f_dc, const float* f_rest,
                                      const float* scales_log, const float* rot_raw, const float* op_logit, int tma_ok, size_t smem, float* rec,
                                      uint2* tile_rects, uint32_t* touched, uint32_t* depth_keys, float* radii_out, uint8_t* vis_out)
{
    cudaError_t e = cudaFuncSetAttribute(k_project_fused_fwd<MAXK, PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_project_fused_fwd<MAXK, PART><<<cdiv(N, PB), PB, smem, st>>>(N, xyz, f_dc, f_rest, scales_log, rot_raw, op_logit, vp, tma_ok, rec, tile_rects,
                                                                      touched, depth_keys, radii_out, vis_out);
    return cudaGetLastError();
}

cudaError_t launch_project_fused_fwd(cudaStream_t st, int N, const ViewParams& vp, const float* xyz, const float* f_dc,
                                     const float* f_rest, const float* scales_log, const float* rot_raw,
                                     const float* op_logit, float* rec, uint2* tile_rects, uint32_t* touched,
                                     uint32_t* depth_keys, float* radii_out, uint8_t* vis_out, int part)
{
    if (N <= 0) return cudaSuccess;
    int tma_ok = aligned16(xyz) && aligned16(f_dc) && aligned16(f_rest) && aligned16(scales_log) && aligned16(rot_raw) &&
                 aligned16(op_logit) && aligned16(rec);
    const size_t smem = project_fused_smem_bytes(vp.K);
    const bool big = vp.coeffCount > 16 || vp.K > 16;
#define GSB_FWD_CASE(MAXK, PART)                                                                                                             \
    return launch_fused_fwd_t<MAXK, PART>(st, N, vp, xyz, f_dc, f_rest, scales_log, rot_raw, op_logit, tma_ok, smem, rec, tile_rects, touched, \
                                          depth_keys, radii_out, vis_out)
    if (part == 1) { if (big) GSB_FWD_CASE(25, 1); GSB_FWD_CASE(16, 1); }
    if (part == 2) { if (big) GSB_FWD_CASE(25, 2); GSB_FWD_CASE(16, 2); }
    if (big) GSB_FWD_CASE(25, 3);
    GSB_FWD_CASE(16, 3);
#undef GSB_FWD_CASE
}

}  // namespace gsb
