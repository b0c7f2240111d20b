// Microbenchmark: what does an instruction COST on a B200 SM sub-partition when the rasterisers' mixes are issued?
// Every mode is a loop body of independent chains run by 8 warps per sub-partition; the figure printed is
//   cycles per loop iteration per warp  =  elapsed cycles / (iterations x warps per sub-partition),
// i.e. the issue/pipe time the body occupies.  The questions it answers (cost model of k_raster_*):
//   * does a packed FFMA2 occupy ONE issue slot (other pipes run in its shadow) or TWO;
//   * do the operand forms of FFMA2 differ (accumulator + broadcast scalar, two distinct vectors, three distinct vectors);
//   * FMUL2 / FADD2; MUFU.EX2 alone and beside FFMA / FFMA2; FMNMX, IADD3, broadcast LDS.128 beside FFMA2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_model issue_model.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

enum {
    M_FFMA16, M_FFMA2_ACC_BC, M_FFMA2_TFORM, M_FFMA2_3VEC, M_FMUL2, M_FADD2, M_FFMA2_IADD, M_FFMA2_FMNMX, M_FFMA2_MUFU4,
    M_FFMA16_MUFU4, M_FFMA2_LDS4, M_MUFU16, M_FFMA16_FMNMX8, M_FFMA2_8_FFMA8, M_FFMA2_RCP4, M_FWD_BODY, M_COUNT
};

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* in, float* out, int iters)
{
    __shared__ float4 sm[64];
    const int i = threadIdx.x;
    if (i < 64) sm[i] = make_float4(in[i & 31], in[(i + 1) & 31], in[(i + 2) & 31], in[(i + 3) & 31]);
    __syncthreads();
    float s[16], t[8];
    u64 v[8], w[8], x[8];
    int n[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = in[(i + j) & 31];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = pack(s[2 * j], s[2 * j + 1]);
        w[j] = pack(s[(2 * j + 3) & 15], s[(2 * j + 5) & 15]);
        x[j] = pack(s[(2 * j + 7) & 15], s[(2 * j + 9) & 15]);
        t[j] = in[(i + 3 * j) & 31];
        n[j] = i + j;
    }
    const float a = in[i & 31], b = in[(i + 7) & 31];
    for (int it = 0; it < iters; ++it) {
        if (MODE == M_FFMA16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = fmaf(s[j], a, b);
        } else if (MODE == M_FFMA2_ACC_BC) {   // colour form: acc = fma2(weight (vector), colour (broadcast scalar), acc)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(w[j], pack(t[j & 3], t[j & 3]), v[j]);
        } else if (MODE == M_FFMA2_TFORM) {    // transmittance form: T = fma2(T, na, T): two distinct vectors
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(v[j], w[j], v[j]);
        } else if (MODE == M_FFMA2_3VEC) {     // colour-gradient form of the backward: C = fma2(contrib, k, C): three distinct vectors
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(w[j], x[j], v[j]);
        } else if (MODE == M_FMUL2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = mul2(v[j], w[j]);
        } else if (MODE == M_FADD2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = add2(v[j], w[j]);
        } else if (MODE == M_FFMA2_IADD) {     // 8 FFMA2 + 8 integer adds (ALU pipe)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = fma2(w[j], pack(t[j & 3], t[j & 3]), v[j]);
                n[j] = n[j] * 3 + (n[(j + 1) & 7] ^ it);
            }
        } else if (MODE == M_FFMA2_FMNMX) {    // 8 FFMA2 + 8 FMNMX
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = fma2(w[j], pack(t[j & 3], t[j & 3]), v[j]);
                s[j] = fmaxf(-s[j], -0.99f);
            }
        } else if (MODE == M_FFMA2_MUFU4) {    // 8 FFMA2 + 4 MUFU.EX2
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(w[j], pack(t[j & 3], t[j & 3]), v[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] = ex2(-s[j]);
        } else if (MODE == M_FFMA16_MUFU4) {   // 16 FFMA + 4 MUFU.EX2
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = fmaf(s[j], a, b);
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = ex2(-t[j]);
        } else if (MODE == M_FFMA2_LDS4) {     // 8 FFMA2 + 4 broadcast LDS.128 feeding the scalars
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 q = sm[(it + j) & 63];
                v[2 * j] = fma2(w[2 * j], pack(q.x, q.x), v[2 * j]);
                v[2 * j + 1] = fma2(w[2 * j + 1], pack(q.z, q.z), v[2 * j + 1]);
            }
        } else if (MODE == M_MUFU16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = ex2(-s[j]);
        } else if (MODE == M_FFMA16_FMNMX8) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = fmaf(s[j], a, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = fmaxf(-t[j], -0.99f);
        } else if (MODE == M_FFMA2_8_FFMA8) {  // 8 FFMA2 + 8 scalar FFMA
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = fma2(w[j], pack(t[j & 3], t[j & 3]), v[j]);
                s[j] = fmaf(s[j], a, b);
            }
        } else if (MODE == M_FFMA2_RCP4) {     // 8 FFMA2 + 2 EX2 + 2 RCP
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma2(w[j], pack(t[j & 3], t[j & 3]), v[j]);
            s[0] = ex2(-s[0]); s[1] = ex2(-s[1]); s[2] = rcp(s[2]); s[3] = rcp(s[3]);
        } else if (MODE == M_FWD_BODY) {
            // one Gaussian of the forward blend loop for one thread (4 pixels = 2 pairs), block-relative form:
            // 3 LDS.128, 3 FFMA of set-up, per pair 2 FFMA2 + 2 EX2 + 2 FMNMX + FMUL2 + FFMA2 (T) + 3 FFMA2 (colour)
            const float4 A = sm[(3 * it) & 63], Q = sm[(3 * it + 1) & 63], Cc = sm[(3 * it + 2) & 63];
            const float E1 = fmaf(a, A.w, Q.y);
            const float E0 = fmaf(a, fmaf(a, A.z, A.y), A.x);
#pragma unroll
            for (int kq = 0; kq < 2; ++kq) {
                const u64 rf = x[kq];
                const u64 p = fma2(rf, fma2(rf, pack(Q.x, Q.x), pack(E1, E1)), pack(E0, E0));
                float p0, p1;
                unpack(p, p0, p1);
                const u64 na = pack(fmaxf(-ex2(p0), -0.99f), fmaxf(-ex2(p1), -0.99f));
                const u64 nc = mul2(v[kq], na);
                w[kq] = fma2(nc, pack(Q.z, Q.z), w[kq]);
                w[kq + 2] = fma2(nc, pack(Q.w, Q.w), w[kq + 2]);
                w[kq + 4] = fma2(nc, pack(Cc.x, Cc.x), w[kq + 4]);
                v[kq] = fma2(v[kq], na, v[kq]);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += s[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float lo, hi;
        unpack(v[j], lo, hi); acc += lo + hi;
        unpack(w[j], lo, hi); acc += lo + hi;
        unpack(x[j], lo, hi); acc += lo + hi;
        acc += t[j] + (float)n[j];
    }
    out[blockIdx.x * blockDim.x + i] = acc;
}

template <int MODE>
void run(const char* name, const float* in, float* out, int sms, int clock_khz)
{
    const int iters = 1 << 13, per_sm = 4, grid = sms * per_sm;   // 4 x 256 threads = 32 warps per SM = 8 per sub-partition
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(in, out, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(in, out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * clock_khz * 1e3;
    printf("%-44s %8.3f ms  %7.2f cycles per iteration per warp\n", name, ms, cycles / iters / (per_sm * 8 / 4));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float h[32]; for (int i = 0; i < 32; ++i) h[i] = 0.5f + 0.01f * i;
    float *in, *out; cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, p.multiProcessorCount * 4 * 256 * 4);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("%s, %d SMs, %d MHz (max clock; cycles assume it)\n", p.name, p.multiProcessorCount, khz / 1000);
    const int S = p.multiProcessorCount;
    run<M_FFMA16>("16 FFMA", in, out, S, khz);
    run<M_FFMA2_ACC_BC>("8 FFMA2 acc += vec * bcast scalar", in, out, S, khz);
    run<M_FFMA2_TFORM>("8 FFMA2 T = T * na + T (2 vectors)", in, out, S, khz);
    run<M_FFMA2_3VEC>("8 FFMA2 C += a * b (3 vectors)", in, out, S, khz);
    run<M_FMUL2>("8 FMUL2 (2 vectors)", in, out, S, khz);
    run<M_FADD2>("8 FADD2 (2 vectors)", in, out, S, khz);
    run<M_FFMA2_8_FFMA8>("8 FFMA2 + 8 FFMA", in, out, S, khz);
    run<M_FFMA2_IADD>("8 FFMA2 + 8 x (IMAD + LOP3)", in, out, S, khz);
    run<M_FFMA2_FMNMX>("8 FFMA2 + 8 FMNMX", in, out, S, khz);
    run<M_FFMA16_FMNMX8>("16 FFMA + 8 FMNMX", in, out, S, khz);
    run<M_MUFU16>("16 MUFU.EX2", in, out, S, khz);
    run<M_FFMA2_MUFU4>("8 FFMA2 + 4 MUFU.EX2", in, out, S, khz);
    run<M_FFMA16_MUFU4>("16 FFMA + 4 MUFU.EX2", in, out, S, khz);
    run<M_FFMA2_RCP4>("8 FFMA2 + 2 MUFU.EX2 + 2 MUFU.RCP", in, out, S, khz);
    run<M_FFMA2_LDS4>("8 FFMA2 + 4 LDS.128 (broadcast)", in, out, S, khz);
    run<M_FWD_BODY>("forward blend body, 1 Gaussian x 4 pixels", in, out, S, khz);
    return 0;
}
