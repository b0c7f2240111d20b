/* gsb_oracle.c — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (libgsb.so, CUDA) never links, calls or falls back to it.
 *
 * Plain C, f32 arithmetic, compiled with
 *     gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 * so that every +,-,*,/ and sqrt rounds exactly once like the reference's f32 Metal kernels.
 * Each function cites the reference file:line it restates (paths relative to the reference
 * root tatsuya-ogawa/GaussianSplattingMlx).  Parity of this file is PINNED against the
 * reference's own shipped kernels compiled for CPU (oracle/_ref, see oracle/build_ref.py) and
 * against the reference's known-answer unit tests (tests/test_oracle_pinning.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GSO_API __attribute__((visibility("default")))

static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float clampf_(float x, float lo, float hi) { return fminf_(fmaxf_(x, lo), hi); }

GSO_API int gso_abi_version(void) { return 1; }

/* ------------------------------------------------------------------------------------------
 * Activation exponential, PINNED BY CONVENTION.  The reference computes exp / sigmoid with MLX ops
 * (Trainer/GaussianRenderer.swift:936-963: MLX.exp, MLX.sigmoid) whose Metal fast-math bits cannot be
 * reproduced off-device ("parity unpinned").  A 1-ulp difference in exp(scale) flips
 * ceil() in the radius of a handful of Gaussians, so the tile lists of the fused path would only
 * match "almost".  The arithmetic is therefore fixed here and restated op for op in the CUDA library
 * (csrc/common.cuh gsb_expf): Cephes-style range reduction k = rint(x log2 e), r = x - k ln2 (two-step
 * Cody-Waite), degree-5 polynomial, every operation a single IEEE-754 f32 rounding (no FMA), scaling
 * by 2^k in two exact-or-once-rounded multiplies.  <= 1 ulp from the exact value
 * (tests/test_oracle_pinning.py), identical bits on CPU and GPU.
 * ---------------------------------------------------------------------------------------- */
static inline float gso_pow2i(int k)   /* 2^k for -126 <= k <= 127 */
{
    union { unsigned u; float f; } c;
    c.u = (unsigned)(k + 127) << 23;
    return c.f;
}
GSO_API float gso_expf(float x)
{
    if (x != x) return x;
    if (x > 88.8f) return INFINITY;
    if (x < -104.0f) return 0.0f;
    const float kf = rintf(x * 1.44269504088896341f);
    float r = x - kf * 0.693359375f;
    r = r - kf * -2.12194440e-4f;
    float p = 1.9875691500e-4f;
    p = p * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    float y = p * (r * r) + r;
    y = y + 1.0f;
    const int k = (int)kf;
    const int k1 = k / 2, k2 = k - k1;
    return (y * gso_pow2i(k1)) * gso_pow2i(k2);
}
GSO_API void gso_expf_array(int n, const float* x, float* y)
{
    for (int i = 0; i < n; ++i) y[i] = gso_expf(x[i]);
}

/* ------------------------------------------------------------------------------------------
 * Activations.  Trainer/GaussianRenderer.swift:936-963 (get_*_from), called from
 * Trainer/GaussianTrainer.swift:652-666.
 *   means3d = xyz; opacity = sigmoid(o); scales = exp(s); rotations = q/(||q||+1e-8);
 *   shs = concat(f_dc[N,1,3], f_rest[N,K-1,3]) -> [N,K,3]
 * ---------------------------------------------------------------------------------------- */
GSO_API void gso_activate_fwd(int N, int K, const float* f_dc, const float* f_rest, const float* scales_log,
                              const float* rot_raw, const float* opacity_logit, float* shs, float* scales,
                              float* rotations, float* opacity)
{
    #pragma omp parallel for
    for (int p = 0; p < N; ++p) {
        for (int c = 0; c < 3; ++c) shs[(size_t)p * K * 3 + c] = f_dc[(size_t)p * 3 + c];
        for (int j = 0; j < (K - 1) * 3; ++j) shs[(size_t)p * K * 3 + 3 + j] = f_rest[(size_t)p * (K - 1) * 3 + j];
        for (int c = 0; c < 3; ++c) scales[p * 3 + c] = gso_expf(scales_log[p * 3 + c]);
        const float* q = rot_raw + (size_t)p * 4;
        float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        float d = n + 1e-8f;
        for (int c = 0; c < 4; ++c) rotations[p * 4 + c] = q[c] / d;
        opacity[p] = 1.0f / (1.0f + gso_expf(-opacity_logit[p]));
    }
}

/* VJP of the activations (MLX autodiff of the ops above; textbook rules). */
GSO_API void gso_activate_bwd(int N, int K, const float* scales_log, const float* rot_raw, const float* opacity_logit,
                              const float* g_shs, const float* g_scales, const float* g_rotations,
                              const float* g_opacity, float* g_f_dc, float* g_f_rest, float* g_scales_log,
                              float* g_rot_raw, float* g_opacity_logit)
{
    #pragma omp parallel for
    for (int p = 0; p < N; ++p) {
        for (int c = 0; c < 3; ++c) g_f_dc[(size_t)p * 3 + c] = g_shs[(size_t)p * K * 3 + c];
        for (int j = 0; j < (K - 1) * 3; ++j) g_f_rest[(size_t)p * (K - 1) * 3 + j] = g_shs[(size_t)p * K * 3 + 3 + j];
        for (int c = 0; c < 3; ++c) g_scales_log[p * 3 + c] = g_scales[p * 3 + c] * gso_expf(scales_log[p * 3 + c]);
        const float* q = rot_raw + (size_t)p * 4;
        const float* gq = g_rotations + (size_t)p * 4;
        float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
        float n = sqrtf(n2);
        float d = n + 1e-8f;
        /* y_c = q_c / d ; dn/dq_c = q_c / n  (0 when n == 0) */
        float dot = gq[0] * q[0] + gq[1] * q[1] + gq[2] * q[2] + gq[3] * q[3];
        float gd = -dot / (d * d);
        for (int c = 0; c < 4; ++c) g_rot_raw[p * 4 + c] = gq[c] / d + (n > 0.0f ? gd * q[c] / n : 0.0f);
        float s = 1.0f / (1.0f + gso_expf(-opacity_logit[p]));
        g_opacity_logit[p] = g_opacity[p] * s * (1.0f - s);
    }
}

/* ------------------------------------------------------------------------------------------
 * SH basis (un-normalised direction).  slang/gaussian_projection_screen_shared.slang:257-319,
 * constants Trainer/ShUtils.swift:4-32.  basis[k] for k < 25; dbasis = d basis / d(x,y,z).
 * ---------------------------------------------------------------------------------------- */
static void sh_basis(float x, float y, float z, int degree, float* b)
{
    for (int k = 0; k < 25; ++k) b[k] = 0.0f;
    b[0] = 0.28209479177387814f;
    if (degree > 0) {
        b[1] = -0.4886025119029199f * y;
        b[2] = 0.4886025119029199f * z;
        b[3] = -0.4886025119029199f * x;
        if (degree > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = 1.0925484305920792f * xy;
            b[5] = -1.0925484305920792f * yz;
            b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
            b[7] = -1.0925484305920792f * xz;
            b[8] = 0.5462742152960396f * (xx - yy);
            if (degree > 2) {
                b[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
                b[10] = 2.890611442640554f * xy * z;
                b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
                b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
                b[14] = 1.445305721320277f * z * (xx - yy);
                b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
                if (degree > 3) {
                    b[16] = 2.5033429417967046f * xy * (xx - yy);
                    b[17] = -1.7701307697799304f * yz * (3.0f * xx - yy);
                    b[18] = 0.9461746957575601f * xy * (7.0f * zz - 1.0f);
                    b[19] = -0.6690465435572892f * yz * (7.0f * zz - 3.0f);
                    b[20] = 0.10578554691520431f * (zz * (35.0f * zz - 30.0f) + 3.0f);
                    b[21] = -0.6690465435572892f * xz * (7.0f * zz - 3.0f);
                    b[22] = 0.47308734787878004f * (xx - yy) * (7.0f * zz - 1.0f);
                    b[23] = -1.7701307697799304f * xz * (xx - 3.0f * yy);
                    b[24] = 0.6258357354491761f * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
                }
            }
        }
    }
}

/* analytic gradient of the basis above wrt (x,y,z) */
static void sh_basis_grad(float x, float y, float z, int degree, float (*d)[3])
{
    for (int k = 0; k < 25; ++k) d[k][0] = d[k][1] = d[k][2] = 0.0f;
    if (degree < 1) return;
    const float C1 = 0.4886025119029199f;
    d[1][1] = -C1; d[2][2] = C1; d[3][0] = -C1;
    if (degree < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const float C20 = 1.0925484305920792f, C22 = 0.31539156525252005f, C24 = 0.5462742152960396f;
    d[4][0] = C20 * y;  d[4][1] = C20 * x;
    d[5][1] = -C20 * z; d[5][2] = -C20 * y;
    d[6][0] = -2.0f * C22 * x; d[6][1] = -2.0f * C22 * y; d[6][2] = 4.0f * C22 * z;
    d[7][0] = -C20 * z; d[7][2] = -C20 * x;
    d[8][0] = 2.0f * C24 * x; d[8][1] = -2.0f * C24 * y;
    if (degree < 3) return;
    const float C30 = 0.5900435899266435f, C31 = 2.890611442640554f, C32 = 0.4570457994644658f;
    const float C33 = 0.3731763325901154f, C35 = 1.445305721320277f;
    /* b9 = -C30*y*(3xx-yy) */
    d[9][0] = -C30 * 6.0f * xy; d[9][1] = -C30 * (3.0f * xx - 3.0f * yy);
    /* b10 = C31*x*y*z */
    d[10][0] = C31 * yz; d[10][1] = C31 * xz; d[10][2] = C31 * xy;
    /* b11 = -C32*y*(4zz-xx-yy) */
    d[11][0] = C32 * 2.0f * xy; d[11][1] = -C32 * (4.0f * zz - xx - 3.0f * yy); d[11][2] = -C32 * 8.0f * yz;
    /* b12 = C33*z*(2zz-3xx-3yy) */
    d[12][0] = -C33 * 6.0f * xz; d[12][1] = -C33 * 6.0f * yz; d[12][2] = C33 * (6.0f * zz - 3.0f * xx - 3.0f * yy);
    /* b13 = -C32*x*(4zz-xx-yy) */
    d[13][0] = -C32 * (4.0f * zz - 3.0f * xx - yy); d[13][1] = C32 * 2.0f * xy; d[13][2] = -C32 * 8.0f * xz;
    /* b14 = C35*z*(xx-yy) */
    d[14][0] = C35 * 2.0f * xz; d[14][1] = -C35 * 2.0f * yz; d[14][2] = C35 * (xx - yy);
    /* b15 = -C30*x*(xx-3yy) */
    d[15][0] = -C30 * (3.0f * xx - 3.0f * yy); d[15][1] = C30 * 6.0f * xy;
    if (degree < 4) return;
    const float C40 = 2.5033429417967046f, C41 = 1.7701307697799304f, C42 = 0.9461746957575601f;
    const float C43 = 0.6690465435572892f, C44 = 0.10578554691520431f, C46 = 0.47308734787878004f;
    const float C48 = 0.6258357354491761f;
    /* b16 = C40*xy*(xx-yy) = C40*(x^3 y - x y^3) */
    d[16][0] = C40 * (3.0f * xx * y - yy * y); d[16][1] = C40 * (xx * x - 3.0f * x * yy);
    /* b17 = -C41*yz*(3xx-yy) */
    d[17][0] = -C41 * 6.0f * xy * z; d[17][1] = -C41 * z * (3.0f * xx - 3.0f * yy); d[17][2] = -C41 * y * (3.0f * xx - yy);
    /* b18 = C42*xy*(7zz-1) */
    d[18][0] = C42 * y * (7.0f * zz - 1.0f); d[18][1] = C42 * x * (7.0f * zz - 1.0f); d[18][2] = C42 * 14.0f * xy * z;
    /* b19 = -C43*yz*(7zz-3) */
    d[19][1] = -C43 * z * (7.0f * zz - 3.0f); d[19][2] = -C43 * y * (21.0f * zz - 3.0f);
    /* b20 = C44*(35 z^4 - 30 z^2 + 3) */
    d[20][2] = C44 * (140.0f * zz * z - 60.0f * z);
    /* b21 = -C43*xz*(7zz-3) */
    d[21][0] = -C43 * z * (7.0f * zz - 3.0f); d[21][2] = -C43 * x * (21.0f * zz - 3.0f);
    /* b22 = C46*(xx-yy)*(7zz-1) */
    d[22][0] = C46 * 2.0f * x * (7.0f * zz - 1.0f); d[22][1] = -C46 * 2.0f * y * (7.0f * zz - 1.0f);
    d[22][2] = C46 * (xx - yy) * 14.0f * z;
    /* b23 = -C41*xz*(xx-3yy) */
    d[23][0] = -C41 * z * (3.0f * xx - 3.0f * yy); d[23][1] = C41 * 6.0f * xy * z; d[23][2] = -C41 * x * (xx - 3.0f * yy);
    /* b24 = C48*(x^4 - 6 x^2 y^2 + y^4) */
    d[24][0] = C48 * (4.0f * xx * x - 12.0f * x * yy); d[24][1] = C48 * (4.0f * yy * y - 12.0f * xx * y);
}

/* known-answer hook for tests/test_oracle_pinning.py (GaussianSplattingMlxTests/ShUtilsTests.swift) */
GSO_API void gso_sh_basis(float x, float y, float z, int degree, float* basis25) { sh_basis(x, y, z, degree, basis25); }

typedef struct {
    float l[9];      /* L = R*S rows */
    float q[4];      /* normalised quaternion */
    float safeNorm, norm;
    float cov3d[9];
} cov3d_ctx;

/* slang/gaussian_projection_screen_shared.slang:117-168 buildCov3dFromScaleRotation */
static void build_cov3d(float sx, float sy, float sz, float rw, float rx, float ry, float rz, cov3d_ctx* o)
{
    float norm = sqrtf(rw * rw + rx * rx + ry * ry + rz * rz);
    float safeNorm = fmaxf_(norm, 1e-8f);
    float qw = rw / safeNorm, qx = rx / safeNorm, qy = ry / safeNorm, qz = rz / safeNorm;
    float r00 = 1.0f - 2.0f * (qy * qy + qz * qz);
    float r01 = 2.0f * (qx * qy - qw * qz);
    float r02 = 2.0f * (qx * qz + qw * qy);
    float r10 = 2.0f * (qx * qy + qw * qz);
    float r11 = 1.0f - 2.0f * (qx * qx + qz * qz);
    float r12 = 2.0f * (qy * qz - qw * qx);
    float r20 = 2.0f * (qx * qz - qw * qy);
    float r21 = 2.0f * (qy * qz + qw * qx);
    float r22 = 1.0f - 2.0f * (qx * qx + qy * qy);
    float l00 = r00 * sx, l01 = r01 * sy, l02 = r02 * sz;
    float l10 = r10 * sx, l11 = r11 * sy, l12 = r12 * sz;
    float l20 = r20 * sx, l21 = r21 * sy, l22 = r22 * sz;
    float* c = o->cov3d;
    c[0] = l00 * l00 + l01 * l01 + l02 * l02;
    c[1] = l00 * l10 + l01 * l11 + l02 * l12;
    c[2] = l00 * l20 + l01 * l21 + l02 * l22;
    c[3] = l10 * l00 + l11 * l01 + l12 * l02;
    c[4] = l10 * l10 + l11 * l11 + l12 * l12;
    c[5] = l10 * l20 + l11 * l21 + l12 * l22;
    c[6] = l20 * l00 + l21 * l01 + l22 * l02;
    c[7] = l20 * l10 + l21 * l11 + l22 * l12;
    c[8] = l20 * l20 + l21 * l21 + l22 * l22;
    float L[9] = {l00, l01, l02, l10, l11, l12, l20, l21, l22};
    memcpy(o->l, L, sizeof L);
    o->q[0] = qw; o->q[1] = qx; o->q[2] = qy; o->q[3] = qz;
    o->safeNorm = safeNorm; o->norm = norm;
}

/* known-answer hook: GaussianSplattingMlxTests.swift:73-130 (build_rotation / build_scaling_rotation) */
GSO_API void gso_build_scaling_rotation(const float* s3, const float* q4, float* L9, float* cov9)
{
    cov3d_ctx c;
    build_cov3d(s3[0], s3[1], s3[2], q4[0], q4[1], q4[2], q4[3], &c);
    memcpy(L9, c.l, sizeof c.l);
    memcpy(cov9, c.cov3d, sizeof c.cov3d);
}

typedef struct {
    float t0, t1, t2, clipX, clipY, limX, limY, tx, ty;
    float j00, j02, j11, j12;
    float b[6];   /* b00 b01 b02 b10 b11 b12 */
    float t[6];   /* t00 t01 t02 t10 t11 t12 */
    float cov2d[4];
} cov2d_ctx;

/* slang/gaussian_projection_screen_shared.slang:170-243 buildCov2dFromCov3d */
static void build_cov2d(float m0, float m1, float m2, const float* c3, const float* V, float fovX, float fovY,
                        float focalX, float focalY, cov2d_ctx* o)
{
    float a00 = V[0], a01 = V[1], a02 = V[2], a10 = V[4], a11 = V[5], a12 = V[6], a20 = V[8], a21 = V[9], a22 = V[10];
    float t30 = V[12], t31 = V[13], t32 = V[14];
    float t0 = m0 * a00 + m1 * a10 + m2 * a20 + t30;
    float t1 = m0 * a01 + m1 * a11 + m2 * a21 + t31;
    float t2 = m0 * a02 + m1 * a12 + m2 * a22 + t32;
    float tanFovX = tanf(fovX * 0.5f);
    float tanFovY = tanf(fovY * 0.5f);
    float limX = tanFovX * 1.3f, limY = tanFovY * 1.3f;
    float clipX = clampf_(t2, -tanFovX * 1.3f, limX);
    float clipY = clampf_(t2, -tanFovY * 1.3f, limY);
    float tx = t0 / clipX * t2;
    float ty = t1 / clipY * t2;
    float tz = t2;
    float j00 = focalX / tz;
    float j02 = -tx * focalX / (tz * tz);
    float j11 = focalY / tz;
    float j12 = -ty * focalY / (tz * tz);
    float w00 = a00, w01 = a10, w02 = a20, w10 = a01, w11 = a11, w12 = a21, w20 = a02, w21 = a12, w22 = a22;
    float b00 = j00 * w00 + j02 * w20;
    float b01 = j00 * w01 + j02 * w21;
    float b02 = j00 * w02 + j02 * w22;
    float b10 = j11 * w10 + j12 * w20;
    float b11 = j11 * w11 + j12 * w21;
    float b12 = j11 * w12 + j12 * w22;
    float t00 = b00 * c3[0] + b01 * c3[3] + b02 * c3[6];
    float t01 = b00 * c3[1] + b01 * c3[4] + b02 * c3[7];
    float t02 = b00 * c3[2] + b01 * c3[5] + b02 * c3[8];
    float t10 = b10 * c3[0] + b11 * c3[3] + b12 * c3[6];
    float t11 = b10 * c3[1] + b11 * c3[4] + b12 * c3[7];
    float t12 = b10 * c3[2] + b11 * c3[5] + b12 * c3[8];
    o->cov2d[0] = t00 * b00 + t01 * b01 + t02 * b02 + 0.3f;
    o->cov2d[1] = t00 * b10 + t01 * b11 + t02 * b12;
    o->cov2d[2] = t10 * b00 + t11 * b01 + t12 * b02;
    o->cov2d[3] = t10 * b10 + t11 * b11 + t12 * b12 + 0.3f;
    o->t0 = t0; o->t1 = t1; o->t2 = t2; o->clipX = clipX; o->clipY = clipY; o->limX = limX; o->limY = limY;
    o->tx = tx; o->ty = ty; o->j00 = j00; o->j02 = j02; o->j11 = j11; o->j12 = j12;
    o->b[0] = b00; o->b[1] = b01; o->b[2] = b02; o->b[3] = b10; o->b[4] = b11; o->b[5] = b12;
    o->t[0] = t00; o->t[1] = t01; o->t[2] = t02; o->t[3] = t10; o->t[4] = t11; o->t[5] = t12;
}

/* ------------------------------------------------------------------------------------------
 * K1 gaussian_projection_screen_fused_forward.  slang/gaussian_projection_kernels.slang:36-173
 * with the shared math of slang/gaussian_projection_screen_shared.slang:53-115,245-255,375-382.
 * Inputs are ACTIVATED tensors (the reference applies activations as MLX ops beforehand).
 * ---------------------------------------------------------------------------------------- */
GSO_API void gso_project_fwd(int N, int degree, int K, const float* scales, const float* rotations,
                             const float* means3d, const float* shs, const float* camCenter, const float* V,
                             const float* P, float fovX, float fovY, float focalX, float focalY, float imageW,
                             float imageH, float* means2d, float* depths, float* color, float* cov2d, float* conic,
                             float* radii, float* rectMin, float* rectMax)
{
    int coeffCount = (degree + 1) * (degree + 1);
    if (coeffCount > 25) coeffCount = 25;
    #pragma omp parallel for
    for (int p = 0; p < N; ++p) {
        float m0 = means3d[p * 3 + 0], m1 = means3d[p * 3 + 1], m2 = means3d[p * 3 + 2];
        /* evaluateProjectionNdcOutputs :53-107 */
        float pv0 = m0 * V[0] + m1 * V[4] + m2 * V[8] + V[12];
        float pv1 = m0 * V[1] + m1 * V[5] + m2 * V[9] + V[13];
        float pv2 = m0 * V[2] + m1 * V[6] + m2 * V[10] + V[14];
        float pv3 = m0 * V[3] + m1 * V[7] + m2 * V[11] + V[15];
        float pc0 = pv0 * P[0] + pv1 * P[4] + pv2 * P[8] + pv3 * P[12];
        float pc1 = pv0 * P[1] + pv1 * P[5] + pv2 * P[9] + pv3 * P[13];
        float pc3 = pv0 * P[3] + pv1 * P[7] + pv2 * P[11] + pv3 * P[15];
        float wInv = 1.0f / (pc3 + 0.000001f);
        float ndcX = pc0 * wInv, ndcY = pc1 * wInv;
        float visibleMask = (pv2 >= 0.2f) ? 1.0f : 0.0f;
        /* ndcToScreen :109-115 */
        float sx = ((ndcX + 1.0f) * imageW - 1.0f) * 0.5f;
        float sy = ((ndcY + 1.0f) * imageH - 1.0f) * 0.5f;
        /* evaluateShColorFromPoint :257-319 */
        float dx = m0 - camCenter[0], dy = m1 - camCenter[1], dz = m2 - camCenter[2];
        float basis[25];
        sh_basis(dx, dy, dz, degree, basis);
        const float* sh = shs + (size_t)p * K * 3;
        float col[3];
        for (int c = 0; c < 3; ++c) {
            float acc = basis[0] * sh[c];
            for (int k = 1; k < coeffCount; ++k) acc += basis[k] * sh[k * 3 + c];
            acc += 0.5f;
            col[c] = fmaxf_(acc, 0.0f);
        }
        cov3d_ctx c3;
        build_cov3d(scales[p * 3], scales[p * 3 + 1], scales[p * 3 + 2], rotations[p * 4], rotations[p * 4 + 1],
                    rotations[p * 4 + 2], rotations[p * 4 + 3], &c3);
        cov2d_ctx c2;
        build_cov2d(m0, m1, m2, c3.cov3d, V, fovX, fovY, focalX, focalY, &c2);
        /* inverseCov2d :245-255 */
        float det = c2.cov2d[0] * c2.cov2d[3] - c2.cov2d[1] * c2.cov2d[2];
        means2d[p * 2] = sx; means2d[p * 2 + 1] = sy;
        depths[p] = pv2;
        color[p * 3] = col[0]; color[p * 3 + 1] = col[1]; color[p * 3 + 2] = col[2];
        for (int c = 0; c < 4; ++c) cov2d[p * 4 + c] = c2.cov2d[c];
        conic[p * 4 + 0] = c2.cov2d[3] / det;
        conic[p * 4 + 1] = -c2.cov2d[1] / det;
        conic[p * 4 + 2] = -c2.cov2d[2] / det;
        conic[p * 4 + 3] = c2.cov2d[0] / det;
        /* computeRadiusFromCov2d :375-382 */
        float mid = 0.5f * (c2.cov2d[0] + c2.cov2d[3]);
        float delta = fmaxf_(mid * mid - det, 1e-5f);
        float lambdaMax = mid + sqrtf(delta);
        float radius = 3.0f * ceilf(sqrtf(lambdaMax));
        float vr = radius * visibleMask;
        radii[p] = vr;
        /* kernels.slang:158-172 */
        float maxX = imageW - 1.0f, maxY = imageH - 1.0f;
        float minX = sx - vr, minY = sy - vr, maxRX = sx + vr, maxRY = sy + vr;
        if (minX < 0.0f) minX = 0.0f;
        if (minY < 0.0f) minY = 0.0f;
        if (maxRX > maxX) maxRX = maxX;
        if (maxRY > maxY) maxRY = maxY;
        rectMin[p * 2] = minX; rectMin[p * 2 + 1] = minY;
        rectMax[p * 2] = maxRX; rectMax[p * 2 + 1] = maxRY;
    }
}

/* ------------------------------------------------------------------------------------------
 * K2 gaussian_projection_screen_fused_backward.  slang/gaussian_projection_kernels.slang:205-398
 * is Slang reverse-mode AD of K1; this is the hand-derived VJP of the same function with the
 * derivative conventions of the shipped MSL header: max ties 0.5 (_d_max_0), clamp passes the
 * gradient iff lo <= x <= hi (_d_clamp_0), sqrt' = 0.5/sqrt(max(1e-7,x)) (_d_sqrt_0).
 * ---------------------------------------------------------------------------------------- */
GSO_API void gso_project_bwd(int N, int degree, int K, const float* scales, const float* rotations,
                             const float* means3d, const float* shs, const float* camCenter, const float* V,
                             const float* P, float fovX, float fovY, float focalX, float focalY, float imageW,
                             float imageH, const float* cotDepths, const float* cotMeans2d, const float* cotCov2d,
                             const float* cotColor, const float* cotConic, float* gScales, float* gRot,
                             float* gMeans3d, float* gShs, float* gCamCenterPoint)
{
    int coeffCount = (degree + 1) * (degree + 1);
    if (coeffCount > 25) coeffCount = 25;
    #pragma omp parallel for
    for (int p = 0; p < N; ++p) {
        float m0 = means3d[p * 3 + 0], m1 = means3d[p * 3 + 1], m2 = means3d[p * 3 + 2];
        float gm[3] = {0.f, 0.f, 0.f};
        /* ---- colour path ---- */
        float dx = m0 - camCenter[0], dy = m1 - camCenter[1], dz = m2 - camCenter[2];
        float basis[25], dbasis[25][3];
        sh_basis(dx, dy, dz, degree, basis);
        sh_basis_grad(dx, dy, dz, degree, dbasis);
        const float* sh = shs + (size_t)p * K * 3;
        float* gsh = gShs + (size_t)p * K * 3;
        float gdir[3] = {0.f, 0.f, 0.f};
        for (int c = 0; c < 3; ++c) {
            float acc = basis[0] * sh[c];
            for (int k = 1; k < coeffCount; ++k) acc += basis[k] * sh[k * 3 + c];
            acc += 0.5f;
            float g = cotColor[p * 3 + c];
            float gpre = acc > 0.0f ? g : (acc < 0.0f ? 0.0f : 0.5f * g);
            for (int k = 0; k < coeffCount; ++k) {
                gsh[k * 3 + c] = basis[k] * gpre;
                float s = sh[k * 3 + c] * gpre;
                gdir[0] += dbasis[k][0] * s; gdir[1] += dbasis[k][1] * s; gdir[2] += dbasis[k][2] * s;
            }
        }
        gm[0] += gdir[0]; gm[1] += gdir[1]; gm[2] += gdir[2];
        gCamCenterPoint[p * 3] = -gdir[0]; gCamCenterPoint[p * 3 + 1] = -gdir[1]; gCamCenterPoint[p * 3 + 2] = -gdir[2];

        /* ---- recompute cov chain ---- */
        float sxs = scales[p * 3], sys = scales[p * 3 + 1], szs = scales[p * 3 + 2];
        float rw = rotations[p * 4], rx = rotations[p * 4 + 1], ry = rotations[p * 4 + 2], rz = rotations[p * 4 + 3];
        cov3d_ctx c3;
        build_cov3d(sxs, sys, szs, rw, rx, ry, rz, &c3);
        cov2d_ctx c2;
        build_cov2d(m0, m1, m2, c3.cov3d, V, fovX, fovY, focalX, focalY, &c2);
        float c00 = c2.cov2d[0], c01 = c2.cov2d[1], c10 = c2.cov2d[2], c11 = c2.cov2d[3];
        float det = c00 * c11 - c01 * c10;
        /* ---- conic -> cov2d ---- */
        float q0 = cotConic[p * 4], q1 = cotConic[p * 4 + 1], q2 = cotConic[p * 4 + 2], q3 = cotConic[p * 4 + 3];
        float G00 = cotCov2d[p * 4], G01 = cotCov2d[p * 4 + 1], G10 = cotCov2d[p * 4 + 2], G11 = cotCov2d[p * 4 + 3];
        float invdet = 1.0f / det;
        float gdet = (-q0 * c11 + q1 * c01 + q2 * c10 - q3 * c00) * invdet * invdet;
        G11 += q0 * invdet + gdet * c00;
        G01 += -q1 * invdet - gdet * c10;
        G10 += -q2 * invdet - gdet * c01;
        G00 += q3 * invdet + gdet * c11;
        /* ---- cov2d = T b^T (+0.3 I), T = b * Sigma ---- */
        const float* b = c2.b; const float* t = c2.t;
        float gt[6], gb[6];
        for (int k = 0; k < 3; ++k) {
            gt[k]     = G00 * b[k] + G01 * b[3 + k];
            gt[3 + k] = G10 * b[k] + G11 * b[3 + k];
            gb[k]     = G00 * t[k] + G10 * t[3 + k];
            gb[3 + k] = G01 * t[k] + G11 * t[3 + k];
        }
        float gS[9];
        const float* S = c3.cov3d;
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) {
                gS[k * 3 + c] = b[k] * gt[c] + b[3 + k] * gt[3 + c];
            }
        for (int k = 0; k < 3; ++k) {
            gb[k]     += gt[0] * S[k * 3 + 0] + gt[1] * S[k * 3 + 1] + gt[2] * S[k * 3 + 2];
            gb[3 + k] += gt[3] * S[k * 3 + 0] + gt[4] * S[k * 3 + 1] + gt[5] * S[k * 3 + 2];
        }
        /* ---- b = J W ---- */
        float a00 = V[0], a01 = V[1], a02 = V[2], a10 = V[4], a11 = V[5], a12 = V[6], a20 = V[8], a21 = V[9], a22 = V[10];
        float w0[3] = {a00, a10, a20}, w1[3] = {a01, a11, a21}, w2[3] = {a02, a12, a22};
        float gj00 = gb[0] * w0[0] + gb[1] * w0[1] + gb[2] * w0[2];
        float gj02 = gb[0] * w2[0] + gb[1] * w2[1] + gb[2] * w2[2];
        float gj11 = gb[3] * w1[0] + gb[4] * w1[1] + gb[5] * w1[2];
        float gj12 = gb[3] * w2[0] + gb[4] * w2[1] + gb[5] * w2[2];
        float tz = c2.t2, tz2 = tz * tz, tz3 = tz2 * tz;
        float gtz = -focalX / tz2 * gj00 - focalY / tz2 * gj11 + 2.0f * c2.tx * focalX / tz3 * gj02
                    + 2.0f * c2.ty * focalY / tz3 * gj12;
        float gtx = -focalX / tz2 * gj02;
        float gty = -focalY / tz2 * gj12;
        /* tx = t0/clipX*t2 */
        float gt0 = gtx * c2.t2 / c2.clipX;
        float gclipX = -gtx * c2.t0 * c2.t2 / (c2.clipX * c2.clipX);
        float gt2 = gtx * c2.t0 / c2.clipX;
        float gt1 = gty * c2.t2 / c2.clipY;
        float gclipY = -gty * c2.t1 * c2.t2 / (c2.clipY * c2.clipY);
        gt2 += gty * c2.t1 / c2.clipY;
        if (c2.t2 >= -c2.limX && c2.t2 <= c2.limX) gt2 += gclipX;
        if (c2.t2 >= -c2.limY && c2.t2 <= c2.limY) gt2 += gclipY;
        gt2 += gtz;
        gm[0] += gt0 * a00 + gt1 * a01 + gt2 * a02;
        gm[1] += gt0 * a10 + gt1 * a11 + gt2 * a12;
        gm[2] += gt0 * a20 + gt1 * a21 + gt2 * a22;
        /* ---- NDC path (evaluateProjectionNdcOutputs bwd) ---- */
        float pv0 = m0 * V[0] + m1 * V[4] + m2 * V[8] + V[12];
        float pv1 = m0 * V[1] + m1 * V[5] + m2 * V[9] + V[13];
        float pv2 = m0 * V[2] + m1 * V[6] + m2 * V[10] + V[14];
        float pv3 = m0 * V[3] + m1 * V[7] + m2 * V[11] + V[15];
        float pc0 = pv0 * P[0] + pv1 * P[4] + pv2 * P[8] + pv3 * P[12];
        float pc1 = pv0 * P[1] + pv1 * P[5] + pv2 * P[9] + pv3 * P[13];
        float pc3 = pv0 * P[3] + pv1 * P[7] + pv2 * P[11] + pv3 * P[15];
        float wInv = 1.0f / (pc3 + 0.000001f);
        float gndcX = cotMeans2d[p * 2] * 0.5f * imageW;
        float gndcY = cotMeans2d[p * 2 + 1] * 0.5f * imageH;
        float gpc0 = gndcX * wInv, gpc1 = gndcY * wInv;
        float gwInv = gndcX * pc0 + gndcY * pc1;
        float gpc3 = -gwInv * wInv * wInv;
        float gpv[4];
        for (int j = 0; j < 4; ++j) gpv[j] = gpc0 * P[j * 4 + 0] + gpc1 * P[j * 4 + 1] + gpc3 * P[j * 4 + 3];
        gpv[2] += cotDepths[p];
        for (int i = 0; i < 3; ++i)
            gm[i] += gpv[0] * V[i * 4 + 0] + gpv[1] * V[i * 4 + 1] + gpv[2] * V[i * 4 + 2] + gpv[3] * V[i * 4 + 3];
        gMeans3d[p * 3] = gm[0]; gMeans3d[p * 3 + 1] = gm[1]; gMeans3d[p * 3 + 2] = gm[2];
        /* ---- cov3d = L L^T  ->  gL = (gS + gS^T) L ---- */
        const float* L = c3.l;
        float gL[9];
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) {
                float acc = 0.f;
                for (int c = 0; c < 3; ++c) acc += (gS[r * 3 + c] + gS[c * 3 + r]) * L[c * 3 + k];
                gL[r * 3 + k] = acc;
            }
        /* L[r][k] = R[r][k] * s[k] */
        float sv[3] = {sxs, sys, szs};
        float gR[9];
        float gs[3] = {0.f, 0.f, 0.f};
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) {
                float Rrk = (sv[k] != 0.0f) ? L[r * 3 + k] / sv[k] : 0.0f;
                gs[k] += gL[r * 3 + k] * Rrk;
                gR[r * 3 + k] = gL[r * 3 + k] * sv[k];
            }
        /* recompute R exactly for the scale gradient (avoid the division above) */
        {
            float qw = c3.q[0], qx = c3.q[1], qy = c3.q[2], qz = c3.q[3];
            float R[9] = {1.0f - 2.0f * (qy * qy + qz * qz), 2.0f * (qx * qy - qw * qz), 2.0f * (qx * qz + qw * qy),
                          2.0f * (qx * qy + qw * qz), 1.0f - 2.0f * (qx * qx + qz * qz), 2.0f * (qy * qz - qw * qx),
                          2.0f * (qx * qz - qw * qy), 2.0f * (qy * qz + qw * qx), 1.0f - 2.0f * (qx * qx + qy * qy)};
            gs[0] = gs[1] = gs[2] = 0.f;
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k) gs[k] += gL[r * 3 + k] * R[r * 3 + k];
            gScales[p * 3] = gs[0]; gScales[p * 3 + 1] = gs[1]; gScales[p * 3 + 2] = gs[2];
            /* R -> normalised quaternion */
            float gqw = 2.0f * (-qz * gR[1] + qy * gR[2] + qz * gR[3] - qx * gR[5] - qy * gR[6] + qx * gR[7]);
            float gqx = 2.0f * (qy * gR[1] + qz * gR[2] + qy * gR[3] - 2.0f * qx * gR[4] - qw * gR[5] + qz * gR[6]
                                + qw * gR[7] - 2.0f * qx * gR[8]);
            float gqy = 2.0f * (-2.0f * qy * gR[0] + qx * gR[1] + qw * gR[2] + qx * gR[3] + qz * gR[5] - qw * gR[6]
                                + qz * gR[7] - 2.0f * qy * gR[8]);
            float gqz = 2.0f * (-2.0f * qz * gR[0] - qw * gR[1] + qx * gR[2] + qw * gR[3] - 2.0f * qz * gR[4]
                                + qy * gR[5] + qx * gR[6] + qy * gR[7]);
            /* q = r / safeNorm, safeNorm = max(sqrt(n2), 1e-8) */
            float sn = c3.safeNorm;
            float gsn = -(gqw * rw + gqx * rx + gqy * ry + gqz * rz) / (sn * sn);
            float gnorm = c3.norm > 1e-8f ? gsn : (c3.norm < 1e-8f ? 0.0f : 0.5f * gsn);
            float n2 = rw * rw + rx * rx + ry * ry + rz * rz;
            float gn2 = 0.5f / sqrtf(fmaxf_(1e-7f, n2)) * gnorm;
            gRot[p * 4 + 0] = gqw / sn + gn2 * 2.0f * rw;
            gRot[p * 4 + 1] = gqx / sn + gn2 * 2.0f * rx;
            gRot[p * 4 + 2] = gqy / sn + gn2 * 2.0f * ry;
            gRot[p * 4 + 3] = gqz / sn + gn2 * 2.0f * rz;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Tile rectangle of a Gaussian.  Shared by K3/K4:
 * slang/gaussian_tile_global_kernels.slang:39-55 and :96-112.
 * ---------------------------------------------------------------------------------------- */
static inline void tile_rect(const float* rectMin, const float* rectMax, int idx, int tileW, int tileH, int imageW,
                             int imageH, int* x0, int* y0, int* x1, int* y1, int* gridWo)
{
    int tMinX = (int)floorf(rectMin[idx * 2 + 0] / (float)tileW);
    int tMinY = (int)floorf(rectMin[idx * 2 + 1] / (float)tileH);
    int tMaxX = (int)floorf(rectMax[idx * 2 + 0] / (float)tileW) + 1;
    int tMaxY = (int)floorf(rectMax[idx * 2 + 1] / (float)tileH) + 1;
    int gridW = (imageW + tileW - 1) / tileW, gridH = (imageH + tileH - 1) / tileH;
    #define CL(v, hi) ((v) < (hi) ? ((v) > 0 ? (v) : 0) : ((hi) > 0 ? (hi) : 0))
    *x0 = CL(tMinX, gridW); *y0 = CL(tMinY, gridH); *x1 = CL(tMaxX, gridW); *y1 = CL(tMaxY, gridH);
    #undef CL
    *gridWo = gridW;
}

/* K3 count_tiles_per_gaussian.  slang/gaussian_tile_global_kernels.slang:17-58 */
GSO_API void gso_count_tiles(int N, const float* rectMin, const float* rectMax, const float* radii, int tileW, int tileH,
                             int imageW, int imageH, uint32_t* tilesTouched)
{
    #pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        if (radii[i] <= 0.0f) { tilesTouched[i] = 0; continue; }
        int x0, y0, x1, y1, gw;
        tile_rect(rectMin, rectMax, i, tileW, tileH, imageW, imageH, &x0, &y0, &x1, &y1, &gw);
        tilesTouched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
    }
}

/* cumsum + exclusive offsets.  Trainer/GaussianRenderer.swift:398-409.  Returns M. */
GSO_API uint32_t gso_exclusive_scan(int N, const uint32_t* touched, uint32_t* offsets)
{
    uint32_t run = 0;
    for (int i = 0; i < N; ++i) { offsets[i] = run; run += touched[i]; }
    return run;
}

/* K4 generate_keys.  slang/gaussian_tile_global_kernels.slang:73-126 */
GSO_API void gso_generate_keys(int N, const float* depths, const float* rectMin, const float* rectMax,
                               const float* radii, const uint32_t* offsets, int tileW, int tileH, int imageW,
                               int imageH, uint32_t* keysHigh, uint32_t* keysLow, uint32_t* gaussIdx)
{
    #pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        if (radii[i] <= 0.0f) continue;
        uint32_t depthBits;
        memcpy(&depthBits, &depths[i], 4);
        int x0, y0, x1, y1, gw;
        tile_rect(rectMin, rectMax, i, tileW, tileH, imageW, imageH, &x0, &y0, &x1, &y1, &gw);
        uint32_t off = offsets[i];
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx) {
                keysHigh[off] = (uint32_t)(ty * gw + tx);
                keysLow[off] = depthBits;
                gaussIdx[off] = (uint32_t)i;
                ++off;
            }
    }
}

/* K5 radix_sort_tile_keys_fused_forward.  slang/gaussian_tile_global_kernels.slang:143-305.
 * Restated as the same stable LSD radix sort with 4-bit digits: 8 passes over keyLow, then
 * max(1, ceil(tileBits/4)) passes over keyHigh (single-threaded counting sort per pass; the
 * reference's 128-lane chunking only changes who computes which slice, not the permutation). */
GSO_API void gso_radix_sort_tile_keys(uint32_t M, uint32_t tileBits, const uint32_t* keysHigh, const uint32_t* keysLow,
                                      const uint32_t* values, uint32_t* sortedHigh, uint32_t* sortedLow,
                                      uint32_t* sortedValues)
{
    if (M == 0) return;
    uint32_t highPasses = (tileBits + 3u) / 4u;
    if (highPasses < 1u) highPasses = 1u;
    uint32_t total = 8u + highPasses;
    uint32_t* bufH[2]; uint32_t* bufL[2]; uint32_t* bufV[2];
    for (int i = 0; i < 2; ++i) {
        bufH[i] = (uint32_t*)malloc((size_t)M * 4); bufL[i] = (uint32_t*)malloc((size_t)M * 4);
        bufV[i] = (uint32_t*)malloc((size_t)M * 4);
    }
    memcpy(bufH[0], keysHigh, (size_t)M * 4); memcpy(bufL[0], keysLow, (size_t)M * 4);
    memcpy(bufV[0], values, (size_t)M * 4);
    int cur = 0;
    for (uint32_t pass = 0; pass < total; ++pass) {
        uint32_t hist[16] = {0};
        const uint32_t* src = pass < 8u ? bufL[cur] : bufH[cur];
        uint32_t shift = pass < 8u ? pass * 4u : (pass - 8u) * 4u;
        for (uint32_t i = 0; i < M; ++i) hist[(src[i] >> shift) & 15u]++;
        uint32_t base[16], run = 0;
        for (int d = 0; d < 16; ++d) { base[d] = run; run += hist[d]; }
        int nxt = cur ^ 1;
        for (uint32_t i = 0; i < M; ++i) {
            uint32_t d = (src[i] >> shift) & 15u;
            uint32_t dst = base[d]++;
            bufH[nxt][dst] = bufH[cur][i]; bufL[nxt][dst] = bufL[cur][i]; bufV[nxt][dst] = bufV[cur][i];
        }
        cur = nxt;
    }
    memcpy(sortedHigh, bufH[cur], (size_t)M * 4); memcpy(sortedLow, bufL[cur], (size_t)M * 4);
    memcpy(sortedValues, bufV[cur], (size_t)M * 4);
    for (int i = 0; i < 2; ++i) { free(bufH[i]); free(bufL[i]); free(bufV[i]); }
}

/* K6 + K7 compute_tile_ranges / compute_tile_counts_from_ranges.
 * slang/gaussian_tile_global_kernels.slang:314-367.  ranges is zero-initialised by the caller
 * (initValue 0, Trainer/GaussianRenderer.swift:441-450). */
GSO_API void gso_tile_ranges(uint32_t M, uint32_t numTiles, const uint32_t* sortedHigh, uint32_t* ranges,
                             uint32_t* counts)
{
    memset(ranges, 0, (size_t)numTiles * 8);
    for (uint32_t i = 0; i < M; ++i) {
        uint32_t cur = sortedHigh[i];
        if (i == 0) ranges[cur * 2] = 0;
        else {
            uint32_t prev = sortedHigh[i - 1];
            if (cur != prev) { ranges[prev * 2 + 1] = i; ranges[cur * 2] = i; }
        }
        if (i == M - 1) ranges[cur * 2 + 1] = M;
    }
    for (uint32_t t = 0; t < numTiles; ++t) {
        uint32_t s = ranges[t * 2], e = ranges[t * 2 + 1];
        counts[t] = e > s ? e - s : 0;
    }
}

/* ------------------------------------------------------------------------------------------
 * K9 gaussian_tile_global_forward.  slang/gaussian_tile_global_kernels.slang:523-614, sample
 * math :437-499.  Consumes the CSR tile lists (ranges + sorted Gaussian indices) directly; K8's
 * dense [numTiles,maxTilePairs] padding (:377-404) holds exactly sortedIdx[start+slot].
 * packed = [N,11]: mean2d(2) conic(4) color(3) opacity depth (Trainer/GaussianRenderer.swift:45-51).
 * ---------------------------------------------------------------------------------------- */
static inline float sample_alpha(const float* g, float px, float py)
{
    float dx = px - g[0], dy = py - g[1];
    float dxdy = dx * dy;
    float exponent = -0.5f * (dx * dx * g[2] + dy * dy * g[5] + dxdy * g[3] + dxdy * g[4]);
    float raw = expf(exponent) * g[9];
    return raw > 0.99f ? 0.99f : raw;
}

GSO_API void gso_raster_fwd(int imageW, int imageH, int tileW, int tileH, int whiteBg, const float* packed,
                            const uint32_t* sortedIdx, const uint32_t* ranges, float* outColor, float* outDepth,
                            float* outAlpha, uint32_t* lastContrib)
{
    int gridW = (imageW + tileW - 1) / tileW;
    long P = (long)imageW * imageH;
    #pragma omp parallel for schedule(dynamic, 256)
    for (long p = 0; p < P; ++p) {
        int y = (int)(p / imageW), x = (int)(p % imageW);
        int tile = (y / tileH) * gridW + (x / tileW);
        uint32_t s = ranges[tile * 2], e = ranges[tile * 2 + 1];
        uint32_t count = e > s ? e - s : 0;
        float px = (float)x, py = (float)y;
        float cx = 0.f, cy = 0.f, cz = 0.f, dep = 0.f, T = 1.0f;
        uint32_t nContrib = count;
        for (uint32_t i = 0; i < count; ++i) {
            const float* g = packed + (size_t)sortedIdx[s + i] * 11;
            float alpha = sample_alpha(g, px, py);
            float contrib = T * alpha;
            cx = cx + contrib * g[6]; cy = cy + contrib * g[7]; cz = cz + contrib * g[8];
            dep = dep + contrib * g[10];
            T = T * (1.0f - alpha);
            if (T < 1e-4f) { nContrib = i + 1; break; }
        }
        float bg = whiteBg ? T : 0.0f;
        outColor[p * 3] = cx + bg; outColor[p * 3 + 1] = cy + bg; outColor[p * 3 + 2] = cz + bg;
        outDepth[p] = dep;
        outAlpha[p] = 1.0f - T;
        lastContrib[p] = nContrib;
    }
}

/* ------------------------------------------------------------------------------------------
 * K10 gaussian_tile_global_backward.  slang/gaussian_tile_global_kernels.slang:648-881 with
 * undoTileGlobalPixelState :501-521.  Per pixel, reverse traversal i = count-1..0 (skipping
 * i >= nContrib), state reconstructed by division, hand-derived VJP of
 * updateTileGlobalPixelState (:485-499) and evaluateTileGlobalSample (:437-483); the alpha clamp
 * branch has zero gradient.  Per-Gaussian sums in f64 (reference order is nondeterministic).
 * ---------------------------------------------------------------------------------------- */
GSO_API void gso_raster_bwd(int imageW, int imageH, int tileW, int tileH, int whiteBg, const float* packed,
                            const uint32_t* sortedIdx, const uint32_t* ranges, const float* cotColor,
                            const float* cotDepth, const float* cotAlpha, const float* outColor,
                            const float* outDepth, const float* outAlpha, const uint32_t* lastContrib,
                            double* gradPacked64)
{
    int gridW = (imageW + tileW - 1) / tileW;
    long P = (long)imageW * imageH;
    #pragma omp parallel for schedule(dynamic, 64)
    for (long p = 0; p < P; ++p) {
        int y = (int)(p / imageW), x = (int)(p % imageW);
        int tile = (y / tileH) * gridW + (x / tileW);
        uint32_t s = ranges[tile * 2], e = ranges[tile * 2 + 1];
        uint32_t count = e > s ? e - s : 0;
        float px = (float)x, py = (float)y;
        float gcx = cotColor[p * 3], gcy = cotColor[p * 3 + 1], gcz = cotColor[p * 3 + 2];
        float trans = 1.0f - outAlpha[p];
        float bg = whiteBg ? trans : 0.0f;
        float sX = outColor[p * 3] - bg, sY = outColor[p * 3 + 1] - bg, sZ = outColor[p * 3 + 2] - bg;
        float sD = outDepth[p], sT = trans;
        float kX = gcx, kY = gcy, kZ = gcz, kD = cotDepth[p];
        float kT = -cotAlpha[p] + (whiteBg ? (gcx + gcy + gcz) : 0.0f);
        uint32_t nContrib = lastContrib[p];
        for (long ii = (long)count - 1; ii >= 0; --ii) {
            if ((uint32_t)ii >= nContrib) continue;
            uint32_t gi = sortedIdx[s + ii];
            const float* g = packed + (size_t)gi * 11;
            float dx = px - g[0], dy = py - g[1];
            float dxdy = dx * dy;
            float exponent = -0.5f * (dx * dx * g[2] + dy * dy * g[5] + dxdy * g[3] + dxdy * g[4]);
            float ex = expf(exponent);
            float raw = ex * g[9];
            int clamped = raw > 0.99f;
            float alpha = clamped ? 0.99f : raw;
            /* undo :501-521 */
            float denom = 1.0f - alpha;
            if (denom < 1e-6f) denom = 1e-6f;
            float prevT = sT / denom;
            float contrib = prevT * alpha;
            float pX = sX - contrib * g[6], pY = sY - contrib * g[7], pZ = sZ - contrib * g[8];
            float pD = sD - contrib * g[10];
            /* VJP of update: next.c = prev.c + prev.T*alpha*col ; next.T = prev.T*(1-alpha) */
            float dotc = kX * g[6] + kY * g[7] + kZ * g[8] + kD * g[10];
            float g_contrib = dotc;
            float g_alpha = prevT * g_contrib - prevT * kT;
            float g_prevT = alpha * g_contrib + (1.0f - alpha) * kT;
            float g_colX = contrib * kX, g_colY = contrib * kY, g_colZ = contrib * kZ, g_dep = contrib * kD;
            /* cot of prev state: colours pass through */
            kT = g_prevT;
            sX = pX; sY = pY; sZ = pZ; sD = pD; sT = prevT;
            /* VJP of sample */
            float g_raw = clamped ? 0.0f : g_alpha;
            float g_op = g_raw * ex;
            float g_exp = g_raw * g[9] * ex;      /* d/d exponent */
            float gq = -0.5f * g_exp;
            float g_c00 = gq * dx * dx, g_c11 = gq * dy * dy, g_c01 = gq * dxdy, g_c10 = gq * dxdy;
            float g_dx = gq * (2.0f * dx * g[2] + dy * (g[3] + g[4]));
            float g_dy = gq * (2.0f * dy * g[5] + dx * (g[3] + g[4]));
            double vals[11] = {-g_dx, -g_dy, g_c00, g_c01, g_c10, g_c11, g_colX, g_colY, g_colZ, g_op, g_dep};
            double* out = gradPacked64 + (size_t)gi * 11;
            for (int k = 0; k < 11; ++k) {
                #pragma omp atomic
                out[k] += vals[k];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * SSIM window.  Trainer/LossUtil.swift:47-54 + Trainer/GaussianTrainer.swift:308-314:
 * g[x] = exp(-(x - K/2)^2 / (2 sigma^2)) (centre K/2 = 5.5 for K = 11), normalised, outer product.
 * ---------------------------------------------------------------------------------------- */
GSO_API void gso_ssim_window(int K, float sigma, float* g1d, float* window2d)
{
    float center = (float)K / 2.0f;
    float sum = 0.f;
    for (int x = 0; x < K; ++x) {
        g1d[x] = expf(-powf((float)x - center, 2.0f) / (2.0f * powf(sigma, 2.0f)));
        sum += g1d[x];
    }
    for (int x = 0; x < K; ++x) g1d[x] = g1d[x] / sum;
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) window2d[i * K + j] = g1d[i] * g1d[j];
}

/* K11 ssim_forward.  slang/ssim_kernels.slang:94-155 (direct KxK taps, zero padding, HWC). */
GSO_API void gso_ssim_fwd(int H, int W, int C, int K, const float* img1, const float* img2, const float* window,
                          float* ssim, float* mu1o, float* mu2o, float* s1o, float* s2o, float* s12o)
{
    int pad = K / 2;
    long total = (long)H * W * C;
    #pragma omp parallel for schedule(static)
    for (long idx = 0; idx < total; ++idx) {
        int c = (int)(idx % C);
        long tmp = idx / C;
        int w = (int)(tmp % W), h = (int)(tmp / W);
        float mu1 = 0, mu2 = 0, a = 0, b = 0, ab = 0;
        for (int ki = 0; ki < K; ++ki) {
            int sh = h + ki - pad;
            if (sh < 0 || sh >= H) continue;
            for (int kj = 0; kj < K; ++kj) {
                int sw = w + kj - pad;
                if (sw < 0 || sw >= W) continue;
                float wt = window[ki * K + kj];
                size_t si = ((size_t)sh * W + sw) * C + c;
                float v1 = img1[si], v2 = img2[si];
                mu1 = mu1 + wt * v1; mu2 = mu2 + wt * v2;
                a = a + wt * v1 * v1; b = b + wt * v2 * v2; ab = ab + wt * v1 * v2;
            }
        }
        float s1 = a - mu1 * mu1, s2 = b - mu2 * mu2, s12 = ab - mu1 * mu2;
        const float C1 = 0.0001f, C2 = 0.0009f;
        float A = 2.0f * mu1 * mu2 + C1, B = 2.0f * s12 + C2;
        float Cc = mu1 * mu1 + mu2 * mu2 + C1, D = s1 + s2 + C2;
        ssim[idx] = (A * B) / (Cc * D);
        mu1o[idx] = mu1; mu2o[idx] = mu2; s1o[idx] = s1; s2o[idx] = s2; s12o[idx] = s12;
    }
}

/* K12 ssim_backward.  slang/ssim_kernels.slang:181-266: gather over the KxK window centres that
 * contain this pixel; VJP of ssimFromAccumState (:70-92) then of updateSsimAccumState (:37-52). */
GSO_API void gso_ssim_bwd(int H, int W, int C, int K, const float* gradOut, const float* img1, const float* img2,
                          const float* window, const float* mu1a, const float* mu2a, const float* s1a,
                          const float* s2a, const float* s12a, float* grad1, float* grad2)
{
    int pad = K / 2;
    long total = (long)H * W * C;
    const float C1 = 0.0001f, C2 = 0.0009f;
    #pragma omp parallel for schedule(static)
    for (long idx = 0; idx < total; ++idx) {
        int c = (int)(idx % C);
        long tmp = idx / C;
        int w = (int)(tmp % W), h = (int)(tmp / W);
        float v1 = img1[idx], v2 = img2[idx];
        float g1 = 0.f, g2 = 0.f;
        for (int ki = 0; ki < K; ++ki) {
            int cx = h - ki + pad;
            if (cx < 0 || cx >= H) continue;
            for (int kj = 0; kj < K; ++kj) {
                int cy = w - kj + pad;
                if (cy < 0 || cy >= W) continue;
                float wt = window[ki * K + kj];
                size_t ci = ((size_t)cx * W + cy) * C + c;
                float up = gradOut[ci];
                float m1 = mu1a[ci], m2 = mu2a[ci], s1 = s1a[ci], s2 = s2a[ci], s12 = s12a[ci];
                /* state: mu1, mu2, E11 = s1+m1^2, E22 = s2+m2^2, E12 = s12+m1*m2 */
                float A = 2.0f * m1 * m2 + C1, B = 2.0f * s12 + C2;
                float Cc = m1 * m1 + m2 * m2 + C1, D = s1 + s2 + C2;
                float CD = Cc * D;
                float gA = up * B / CD, gB = up * A / CD;
                float gCD = -up * (A * B) / (CD * CD);
                float gC = gCD * D, gD = gCD * Cc;
                float g_s1 = gD, g_s2 = gD, g_s12 = 2.0f * gB;
                float g_m1 = 2.0f * m2 * gA + 2.0f * m1 * gC - 2.0f * m1 * g_s1 - m2 * g_s12;
                float g_m2 = 2.0f * m1 * gA + 2.0f * m2 * gC - 2.0f * m2 * g_s2 - m1 * g_s12;
                /* update: mu1 += wt v1; E11 += wt v1^2; E12 += wt v1 v2 ... */
                g1 += wt * (g_m1 + 2.0f * v1 * g_s1 + v2 * g_s12);
                g2 += wt * (g_m2 + 2.0f * v2 * g_s2 + v1 * g_s12);
            }
        }
        grad1[idx] = g1; grad2[idx] = g2;
    }
}

/* ------------------------------------------------------------------------------------------
 * Adam (no bias correction).  mlx-swift 0.30.6 MLXOptimizers.Adam.applySingle, called from
 * Trainer/GaussianTrainer.swift:941-948,1066-1079 (source not vendored in the reference; published
 * form: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr * m / (sqrt(v) + eps)).
 * ---------------------------------------------------------------------------------------- */
GSO_API void gso_adam(long n, float* p, const float* g, float* m, float* v, float lr, float b1, float b2, float eps)
{
    #pragma omp parallel for
    for (long i = 0; i < n; ++i) {
        float mi = b1 * m[i] + (1.0f - b1) * g[i];
        float vi = b2 * v[i] + (1.0f - b2) * (g[i] * g[i]);
        m[i] = mi; v[i] = vi;
        p[i] = p[i] - lr * mi / (sqrtf(vi) + eps);
    }
}

/* D1 accum_grad_norm.  Trainer/GaussianTrainer.swift:321-339 */
GSO_API void gso_accum_grad_norm(int N, const float* xyzGrad, float* accum)
{
    #pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        float gx = xyzGrad[i * 3], gy = xyzGrad[i * 3 + 1], gz = xyzGrad[i * 3 + 2];
        accum[i] = accum[i] + sqrtf(gx * gx + gy * gy + gz * gz);
    }
}

/* ------------------------------------------------------------------------------------------
 * Densification.  D2 classify_gaussians (Trainer/GaussianTrainer.swift:344-392):
 *   avg_grad = denom > 0 ? accum/denom : 0;  max_scale = max(exp(s0), exp(s1), exp(s2));
 *   op = 1/(1+exp(-raw));  prune (3, count 0) if op < min_opacity; else if allow_densify and
 *   avg_grad > grad_threshold: split (1, count 2) if max_scale > max_scale_thresh else clone (2, count 2);
 *   else keep (0, count 1).
 * ---------------------------------------------------------------------------------------- */
GSO_API void gso_classify_gaussians(int N, const float* grad_accum, float denom, const float* scales_log,
                                    const float* opacity_logit, float grad_threshold, float max_scale_thresh,
                                    float min_opacity_thresh, int allow_densify, int* actions, int* output_counts)
{
    #pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        float g = grad_accum[i];
        float avg = denom > 0.0f ? g / denom : 0.0f;
        float s0 = expf(scales_log[i * 3]), s1 = expf(scales_log[i * 3 + 1]), s2 = expf(scales_log[i * 3 + 2]);
        float mx = fmaxf_(fmaxf_(s0, s1), s2);
        float op = 1.0f / (1.0f + expf(-opacity_logit[i]));
        int action, cnt;
        if (op < min_opacity_thresh) { action = 3; cnt = 0; }
        else if (allow_densify && avg > grad_threshold) {
            if (mx > max_scale_thresh) { action = 1; cnt = 2; } else { action = 2; cnt = 2; }
        } else { action = 0; cnt = 1; }
        actions[i] = action;
        output_counts[i] = cnt;
    }
}

/* D3 build_densify_output_map (Trainer/GaussianTrainer.swift:397-427): offsets = exclusive scan of counts.
 * noise_mode: 0 none (keep / clone original), 1 split first, 2 split second, 3 clone copy. */
GSO_API void gso_build_densify_output_map(int N, const int* actions, const int* offsets, int* gather, int* noise_mode)
{
    #pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        int a = actions[i], o = offsets[i];
        if (a == 0) { gather[o] = i; noise_mode[o] = 0; }
        else if (a == 1) { gather[o] = i; noise_mode[o] = 1; gather[o + 1] = i; noise_mode[o + 1] = 2; }
        else if (a == 2) { gather[o] = i; noise_mode[o] = 0; gather[o + 1] = i; noise_mode[o + 1] = 3; }
    }
}

/* Phases 4-5 of split_and_prune (Trainer/GaussianTrainer.swift:866-897): gather the six tensors; split
 * children get scales + Float(-log(1.6)) and position offset sign * mean(exp(source scales)) * 0.1 * noise
 * (sign +1 first / -1 second), clone copies get 0.01 * noise.  The MLX expression is evaluated left to right in
 * f32: ((sign * mean) * 0.1) * n, (flag * 0.01) * n, (xyz + split) + clone; mean over 3 = sum / 3
 * (MLX mean: parity unpinned).  base_noise[Nout,3] stands for MLXRandom.normal([totalOutput,3]). */
GSO_API void gso_densify_apply(int Nout, int K, const int* gather, const int* noise_mode, const float* base_noise,
                               const float* xyz, const float* f_dc, const float* f_rest, const float* scales_log,
                               const float* rot, const float* opacity, float* o_xyz, float* o_f_dc, float* o_f_rest,
                               float* o_scales_log, float* o_rot, float* o_opacity)
{
    const float red = (float)(-log(1.6));
    const int R = (K - 1) * 3;
    #pragma omp parallel for
    for (int j = 0; j < Nout; ++j) {
        const int s = gather[j], mode = noise_mode[j];
        for (int c = 0; c < 3; ++c) o_f_dc[(size_t)j * 3 + c] = f_dc[(size_t)s * 3 + c];
        for (int c = 0; c < R; ++c) o_f_rest[(size_t)j * R + c] = f_rest[(size_t)s * R + c];
        for (int c = 0; c < 4; ++c) o_rot[(size_t)j * 4 + c] = rot[(size_t)s * 4 + c];
        o_opacity[j] = opacity[s];
        const float isSplit = (mode == 1 || mode == 2) ? 1.0f : 0.0f;
        for (int c = 0; c < 3; ++c) o_scales_log[(size_t)j * 3 + c] = scales_log[(size_t)s * 3 + c] + isSplit * red;
        const float e0 = expf(scales_log[(size_t)s * 3]), e1 = expf(scales_log[(size_t)s * 3 + 1]), e2 = expf(scales_log[(size_t)s * 3 + 2]);
        const float mean = ((e0 + e1) + e2) / 3.0f;
        const float sign = (mode == 1 ? 1.0f : 0.0f) - (mode == 2 ? 1.0f : 0.0f);
        const float isClone = mode == 3 ? 1.0f : 0.0f;
        for (int c = 0; c < 3; ++c) {
            const float n = base_noise[(size_t)j * 3 + c];
            const float splitNoise = ((sign * mean) * 0.1f) * n;
            const float cloneNoise = (isClone * 0.01f) * n;
            o_xyz[(size_t)j * 3 + c] = (xyz[(size_t)s * 3 + c] + splitNoise) + cloneNoise;
        }
    }
}
