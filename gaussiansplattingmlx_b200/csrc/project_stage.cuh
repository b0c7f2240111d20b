// project_stage.cuh — shared-memory staging of the fused projection kernels (project.cu forward, project_bwd.cu
// backward): a CTA owns PB consecutive Gaussians = one contiguous byte range in each AoS parameter tensor, pulled in
// with TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) on one mbarrier.
#pragma once
#include "kernels.h"

namespace gsb {

constexpr int PB = 128;  // Gaussians per CTA

struct FusedSmem {
    // offsets in floats into the dynamic smem block (all 16-byte aligned)
    int xyz, scales, rot, op, fdc, frest, rec, total;
};
__host__ __device__ inline FusedSmem fused_layout(int K)
{
    FusedSmem L;
    int o = 4;  // first 16 bytes: mbarrier
    L.xyz = o; o += PB * 3;
    L.scales = o; o += PB * 3;
    L.rot = o; o += PB * 4;
    L.op = o; o += PB;
    L.fdc = o; o += PB * 3;
    L.frest = o; o += PB * (K - 1) * 3;
    L.rec = o; o += PB * REC_FLOATS;
    L.total = o;
    return L;
}

// what: bit 0 = positions, bit 1 = scales / rotations / opacities (geometry), bit 2 = SH coefficients (colour)
constexpr int STAGE_XYZ = 1, STAGE_GEO = 2, STAGE_SH = 4, STAGE_ALL = 7;
__device__ __forceinline__ void stage_inputs(float* sm, const FusedSmem& L, int base, int count, int K, bool tma,
                                             const float* xyz, const float* f_dc, const float* f_rest,
                                             const float* scales_log, const float* rot_raw, const float* op_logit,
                                             const float* extra, int extra_floats_per,  // extra → L.rec region
                                             int what = STAGE_ALL)
{
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    const int restF = (K - 1) * 3;
    if (tma) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t floats = 0;
            if (what & STAGE_XYZ) floats += 3;
            if (what & STAGE_GEO) floats += 3 + 4 + 1;
            if (what & STAGE_SH) floats += 3 + restF;
            if (extra) floats += extra_floats_per;
            mbar_expect_tx(bar, (uint32_t)PB * floats * 4u);
            if (what & STAGE_XYZ) bulk_g2s(sm + L.xyz, xyz + (size_t)base * 3, PB * 3 * 4, bar);
            if (what & STAGE_GEO) {
                bulk_g2s(sm + L.scales, scales_log + (size_t)base * 3, PB * 3 * 4, bar);
                bulk_g2s(sm + L.rot, rot_raw + (size_t)base * 4, PB * 4 * 4, bar);
                bulk_g2s(sm + L.op, op_logit + (size_t)base, PB * 4, bar);
            }
            if (what & STAGE_SH) {
                bulk_g2s(sm + L.fdc, f_dc + (size_t)base * 3, PB * 3 * 4, bar);
                if (restF > 0) bulk_g2s(sm + L.frest, f_rest + (size_t)base * restF, PB * restF * 4, bar);
            }
            if (extra) bulk_g2s(sm + L.rec, extra + (size_t)base * extra_floats_per, PB * extra_floats_per * 4, bar);
        }
        mbar_wait(bar, 0);
    } else {
        for (int i = threadIdx.x; i < count * 3; i += PB) {
            if (what & STAGE_XYZ) sm[L.xyz + i] = xyz[(size_t)base * 3 + i];
            if (what & STAGE_GEO) sm[L.scales + i] = scales_log[(size_t)base * 3 + i];
            if (what & STAGE_SH) sm[L.fdc + i] = f_dc[(size_t)base * 3 + i];
        }
        if (what & STAGE_GEO) {
            for (int i = threadIdx.x; i < count * 4; i += PB) sm[L.rot + i] = rot_raw[(size_t)base * 4 + i];
            for (int i = threadIdx.x; i < count; i += PB) sm[L.op + i] = op_logit[(size_t)base + i];
        }
        if (what & STAGE_SH)
            for (int i = threadIdx.x; i < count * restF; i += PB) sm[L.frest + i] = f_rest[(size_t)base * restF + i];
        if (extra)
            for (int i = threadIdx.x; i < count * extra_floats_per; i += PB)
                sm[L.rec + i] = extra[(size_t)base * extra_floats_per + i];
        __syncthreads();
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t fused_smem_bytes(int K) { return (size_t)fused_layout(K).total * sizeof(float); }

}  // namespace gsb
