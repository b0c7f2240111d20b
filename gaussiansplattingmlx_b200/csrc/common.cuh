// common.cuh — shared declarations for libgsb.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/gsb.h"

namespace gsb {

constexpr int REC_FLOATS = 12;  // internal projected record (48 B, three 16-byte quads)
// Raster record = the reference's packed[N,11] (GaussianRenderer.swift:45-51) with the conic and the
// opacity pre-folded for the blend loop (exponent evaluated directly in log2 units):
//   0 meanX 1 meanY 2 A 3 B | 4 C 5 lo 6 r 7 g | 8 b 9 opacity 10 depth 11 idx(bits)
//   A = -0.5*log2(e)*c00,  B = -0.5*log2(e)*(c01 + c10),  C = -0.5*log2(e)*c11,  lo = log2(opacity)
//   alpha = min(0.99, 2^(A dx^2 + B dx dy + C dy^2 + lo))   ( == min(0.99, exp(-0.5 d^T conic d) * opacity) )
// Gradient record (raster backward -> projection backward) keeps the reference's packed order:
//   0 g_meanX 1 g_meanY 2 g_c00 3 g_c01 | 4 g_c10 5 g_c11 6 g_r 7 g_g | 8 g_b 9 g_opacity 10 g_depth 11 unused
constexpr float LOG2E_F = 1.4426950408889634f;

// Per-view constants handed to kernels by value (__grid_constant__).
struct ViewParams {
    float V[16];
    float P[16];
    float cam[3];
    float fovX, fovY, focalX, focalY;
    float tanHalfX, tanHalfY;   // tanf(fov*0.5f) computed on the host with libm
    float imageW, imageH;
    int W, H, tileW, tileH, gridW, gridH;
    int degree, K, coeffCount;
    int whiteBg;
};

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
};

struct Ctx;  // defined in api.cu

#define GSB_CUDA_CHECK(ctx, expr)                                                                   \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            gsb::set_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
            return GSB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

void set_error(Ctx* ctx, const std::string& msg);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Device helpers: mbarrier + bulk async copies (TMA 1-D, SASS UBLKCP) and misc PTX.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed transaction count traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    for (uint32_t it = 0; it < (1u << 26); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy / reduce-add (f32), tracked by the bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g_add_f32(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// Ampere-style 16-byte async copy (SASS LDGSTS) + completion reported to an mbarrier
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to smem must be made visible to the async proxy before a bulk s2g copy
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Look-back status words carry their payload in the same 32/64-bit word as the flag, so polling needs no
// acquire/release ordering: relaxed gpu-scope accesses avoid the MEMBAR + CCTL.IVALL (L1 invalidate) that
// ld.acquire.gpu / st.release.gpu cost on every spin iteration (22 % of the onesweep pass's stall samples).
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// packed-order fields -> raster record quads
// Bits 31/30 of the index word are flags.  Bit 31 = "alpha may reach the 0.99 clamp (or the conic is not negative definite in log2 form)":
// when it is clear, exponent <= 0 everywhere and opacity <= 0.98, so min(0.99, .) is the identity and the backward
// takes a path without the clamp logic.
constexpr uint32_t REC_MAYCLAMP = 0x80000000u;
// Bit 30 is reserved; Gaussian indices stay below 2^30.
constexpr uint32_t REC_IDX_MASK = 0x3fffffffu;
__device__ __forceinline__ void make_raster_record(float mx, float my, float c00, float c01, float c10, float c11, float r,
                                                   float g, float b, float opacity, float depth, uint32_t idx, float4* out)
{
    const float s = -0.5f * LOG2E_F;
    const float A = s * c00, B = s * (c01 + c10), C = s * c11;
    const bool safe = opacity <= 0.98f && A <= 0.0f && C <= 0.0f && B * B <= 4.0f * A * C;   // false for NaNs
    out[0] = make_float4(mx, my, A, B);
    out[1] = make_float4(C, log2f(opacity), r, g);
    out[2] = make_float4(b, opacity, depth, __uint_as_float(idx | (safe ? 0u : REC_MAYCLAMP)));
}
// Activation exponential, pinned by convention (the CPU checker restates the same sequence of
// IEEE-754 f32 operations): the reference's MLX.exp / MLX.sigmoid (Trainer/GaussianRenderer.swift:936-963) cannot be reproduced
// bit for bit off-device, and a 1-ulp difference in exp(scale) moves ceil() in the radius of a few Gaussians, i.e. the
// tile lists.  Cephes-style: k = rint(x log2 e), two-step Cody-Waite reduction, degree-5 polynomial, 2^k applied in two
// multiplies; explicit _rn intrinsics so that no translation unit contracts them into FMAs.  <= 1 ulp.
__device__ __forceinline__ float gsb_expf(float x)
{
    if (x != x) return x;
    if (x > 88.8f) return __int_as_float(0x7f800000);
    if (x < -104.0f) return 0.0f;
    const float kf = rintf(__fmul_rn(x, 1.44269504088896341f));
    float r = __fsub_rn(x, __fmul_rn(kf, 0.693359375f));
    r = __fsub_rn(r, __fmul_rn(kf, -2.12194440e-4f));
    float p = 1.9875691500e-4f;
    p = __fadd_rn(__fmul_rn(p, r), 1.3981999507e-3f);
    p = __fadd_rn(__fmul_rn(p, r), 8.3334519073e-3f);
    p = __fadd_rn(__fmul_rn(p, r), 4.1665795894e-2f);
    p = __fadd_rn(__fmul_rn(p, r), 1.6666665459e-1f);
    p = __fadd_rn(__fmul_rn(p, r), 5.0000001201e-1f);
    float y = __fadd_rn(__fmul_rn(p, __fmul_rn(r, r)), r);
    y = __fadd_rn(y, 1.0f);
    const int k = (int)kf;
    const int k1 = k / 2, k2 = k - k1;
    return __fmul_rn(__fmul_rn(y, __int_as_float((k1 + 127) << 23)), __int_as_float((k2 + 127) << 23));
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Packed FP32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): one issue slot for two IEEE-rn operations, each half bit-identical
// to the scalar fmaf / __fmul_rn / __fadd_rn.  Measured on B200 (tools/microbench/f32x2_rate.cu): same FMA rate as scalar
// FFMA (123 of 128 fma/clk/SM) at half the issue slots when one source is a broadcast scalar; the rasterisers are
// issue-bound, not pipe-bound, so pairs of vertically adjacent pixels are evaluated as one f32x2 lane.
struct f32x2 {
    unsigned long long v;
};
__device__ __forceinline__ f32x2 f2_make(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 f2_bc(float x) { return f2_make(x, x); }
__device__ __forceinline__ float f2_lo(f32x2 a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return lo;
}
__device__ __forceinline__ float f2_hi(f32x2 a)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
    return hi;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
#endif  // __CUDACC__

}  // namespace gsb
