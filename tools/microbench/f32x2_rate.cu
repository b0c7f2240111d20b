// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and mixed with MUFU.EX2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* in, float* out, int iters)
{
    const int i = threadIdx.x;
    const float a = in[i & 31], b = in[(i + 7) & 31];
    float s[16];
    u64 v[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = in[(i + j) & 31];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = pack(s[2 * j], s[2 * j + 1]);
    const u64 a2 = pack(a, a), b2 = pack(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {   // 16 scalar FFMA
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = fmaf(s[j], a, b);
        } else if (MODE == 1) {   // 8 FFMA2 (= 16 fma)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(v[j], a2, b2);
        } else if (MODE == 2) {   // 16 FFMA + 2 MUFU
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = fmaf(s[j], a, b);
            s[0] = ex2(s[0]); s[8] = ex2(s[8]);
        } else if (MODE == 3) {   // 8 FFMA2 + 2 MUFU
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(v[j], a2, b2);
            float lo, hi; unpack(v[0], lo, hi); lo = ex2(lo); v[0] = pack(lo, hi);
            unpack(v[4], lo, hi); lo = ex2(lo); v[4] = pack(lo, hi);
        } else if (MODE == 4) {   // 16 FFMA2 non-broadcast operands
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(v[j], v[(j + 1) & 7], v[(j + 2) & 7]);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += s[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) { float lo, hi; unpack(v[j], lo, hi); acc += lo + hi; }
    out[blockIdx.x * blockDim.x + i] = acc;
}

template <int MODE>
void run(const char* name, const float* in, float* out, int sms, int clock_khz)
{
    const int iters = 1 << 14, grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(in, out, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(in, out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)grid * 256 * iters * 16;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  %6.1f fma/clk/SM (at %d MHz max clock)\n", name, ms, 2 * fma / ms * 1e-9,
           fma / (ms * 1e-3) / sms / (clock_khz * 1e3), clock_khz / 1000);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float h[32]; for (int i = 0; i < 32; ++i) h[i] = 0.5f + 0.01f * i;
    float *in, *out; cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, p.multiProcessorCount * 8 * 256 * 4);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("FFMA x16", in, out, p.multiProcessorCount, khz);
    run<1>("FFMA2 x8 (bcast operands)", in, out, p.multiProcessorCount, khz);
    run<4>("FFMA2 x8 (vector operands)", in, out, p.multiProcessorCount, khz);
    run<2>("FFMA x16 + MUFU x2", in, out, p.multiProcessorCount, khz);
    run<3>("FFMA2 x8 + MUFU x2", in, out, p.multiProcessorCount, khz);
    return 0;
}
