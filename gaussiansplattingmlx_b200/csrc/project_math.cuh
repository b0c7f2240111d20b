// project_math.cuh — per-Gaussian EWA projection + SH colour, forward and hand-derived VJP.
//
// Restates slang/gaussian_projection_screen_shared.slang (reference) op-for-op in f32.  The
// translation unit that includes this file is compiled with --fmad=false so that every
// +,-,*,/ and sqrt rounds once, exactly like the -ffp-contract=off CPU oracle: the geometric
// outputs (means2d, depth, radius, rect → tile lists) are then bit-identical to the reference
// kernels compiled for the CPU.  The stage is HBM-bound (≈500 flop vs 284 B per Gaussian), so
// giving up FMA contraction costs nothing measurable.
#pragma once
#include "common.cuh"

namespace gsb {

// SH basis for an UN-NORMALISED direction (reference quirk; shared.slang:257-319,
// constants Trainer/ShUtils.swift:4-32).  Writes coeffCount entries.
template <int MAXK>
__device__ __forceinline__ void sh_basis(float x, float y, float z, int degree, float* b)
{
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
                if (MAXK > 16 && degree > 3) {
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

// d basis_k / d(x,y,z) contracted with a per-k weight vector w_k (= sum_c sh[k][c] * gpre[c]):
// returns gdir = sum_k w_k * grad basis_k.  wfun(k) supplies w_k lazily.
template <int MAXK, class WFun>
__device__ __forceinline__ void sh_basis_grad_dot(float x, float y, float z, int degree, WFun w, float* gdir)
{
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (degree > 0) {
        const float C1 = 0.4886025119029199f;
        gy += -C1 * w(1);
        gz += C1 * w(2);
        gx += -C1 * w(3);
        if (degree > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            const float C20 = 1.0925484305920792f, C22 = 0.31539156525252005f, C24 = 0.5462742152960396f;
            float w4 = w(4), w5 = w(5), w6 = w(6), w7 = w(7), w8 = w(8);
            gx += C20 * y * w4;  gy += C20 * x * w4;
            gy += -C20 * z * w5; gz += -C20 * y * w5;
            gx += -2.0f * C22 * x * w6; gy += -2.0f * C22 * y * w6; gz += 4.0f * C22 * z * w6;
            gx += -C20 * z * w7; gz += -C20 * x * w7;
            gx += 2.0f * C24 * x * w8; gy += -2.0f * C24 * y * w8;
            if (degree > 2) {
                const float C30 = 0.5900435899266435f, C31 = 2.890611442640554f, C32 = 0.4570457994644658f;
                const float C33 = 0.3731763325901154f, C35 = 1.445305721320277f;
                float w9 = w(9), w10 = w(10), w11 = w(11), w12 = w(12), w13 = w(13), w14 = w(14), w15 = w(15);
                gx += -C30 * 6.0f * xy * w9;  gy += -C30 * (3.0f * xx - 3.0f * yy) * w9;
                gx += C31 * yz * w10; gy += C31 * xz * w10; gz += C31 * xy * w10;
                gx += C32 * 2.0f * xy * w11; gy += -C32 * (4.0f * zz - xx - 3.0f * yy) * w11; gz += -C32 * 8.0f * yz * w11;
                gx += -C33 * 6.0f * xz * w12; gy += -C33 * 6.0f * yz * w12; gz += C33 * (6.0f * zz - 3.0f * xx - 3.0f * yy) * w12;
                gx += -C32 * (4.0f * zz - 3.0f * xx - yy) * w13; gy += C32 * 2.0f * xy * w13; gz += -C32 * 8.0f * xz * w13;
                gx += C35 * 2.0f * xz * w14; gy += -C35 * 2.0f * yz * w14; gz += C35 * (xx - yy) * w14;
                gx += -C30 * (3.0f * xx - 3.0f * yy) * w15; gy += C30 * 6.0f * xy * w15;
                if (MAXK > 16 && degree > 3) {
                    const float C40 = 2.5033429417967046f, C41 = 1.7701307697799304f, C42 = 0.9461746957575601f;
                    const float C43 = 0.6690465435572892f, C44 = 0.10578554691520431f, C46 = 0.47308734787878004f;
                    const float C48 = 0.6258357354491761f;
                    float w16 = w(16), w17 = w(17), w18 = w(18), w19 = w(19), w20 = w(20), w21 = w(21), w22 = w(22),
                          w23 = w(23), w24 = w(24);
                    gx += C40 * (3.0f * xx * y - yy * y) * w16; gy += C40 * (xx * x - 3.0f * x * yy) * w16;
                    gx += -C41 * 6.0f * xy * z * w17; gy += -C41 * z * (3.0f * xx - 3.0f * yy) * w17; gz += -C41 * y * (3.0f * xx - yy) * w17;
                    gx += C42 * y * (7.0f * zz - 1.0f) * w18; gy += C42 * x * (7.0f * zz - 1.0f) * w18; gz += C42 * 14.0f * xy * z * w18;
                    gy += -C43 * z * (7.0f * zz - 3.0f) * w19; gz += -C43 * y * (21.0f * zz - 3.0f) * w19;
                    gz += C44 * (140.0f * zz * z - 60.0f * z) * w20;
                    gx += -C43 * z * (7.0f * zz - 3.0f) * w21; gz += -C43 * x * (21.0f * zz - 3.0f) * w21;
                    gx += C46 * 2.0f * x * (7.0f * zz - 1.0f) * w22; gy += -C46 * 2.0f * y * (7.0f * zz - 1.0f) * w22; gz += C46 * (xx - yy) * 14.0f * z * w22;
                    gx += -C41 * z * (3.0f * xx - 3.0f * yy) * w23; gy += C41 * 6.0f * xy * z * w23; gz += -C41 * x * (xx - 3.0f * yy) * w23;
                    gx += C48 * (4.0f * xx * x - 12.0f * x * yy) * w24; gy += C48 * (4.0f * yy * y - 12.0f * xx * y) * w24;
                }
            }
        }
    }
    gdir[0] = gx; gdir[1] = gy; gdir[2] = gz;
}

struct Cov3d {
    float L[9];   // R*S
    float q[4];   // normalised quaternion (w,x,y,z)
    float safeNorm, norm;
    float S[9];   // covariance, both triangles
};

// shared.slang:117-168
__device__ __forceinline__ void build_cov3d(float sx, float sy, float sz, float rw, float rx, float ry, float rz, Cov3d& o)
{
    float norm = sqrtf(rw * rw + rx * rx + ry * ry + rz * rz);
    float safeNorm = fmaxf(norm, 1e-8f);
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
    o.S[0] = l00 * l00 + l01 * l01 + l02 * l02;
    o.S[1] = l00 * l10 + l01 * l11 + l02 * l12;
    o.S[2] = l00 * l20 + l01 * l21 + l02 * l22;
    o.S[3] = l10 * l00 + l11 * l01 + l12 * l02;
    o.S[4] = l10 * l10 + l11 * l11 + l12 * l12;
    o.S[5] = l10 * l20 + l11 * l21 + l12 * l22;
    o.S[6] = l20 * l00 + l21 * l01 + l22 * l02;
    o.S[7] = l20 * l10 + l21 * l11 + l22 * l12;
    o.S[8] = l20 * l20 + l21 * l21 + l22 * l22;
    o.L[0] = l00; o.L[1] = l01; o.L[2] = l02; o.L[3] = l10; o.L[4] = l11; o.L[5] = l12; o.L[6] = l20; o.L[7] = l21; o.L[8] = l22;
    o.q[0] = qw; o.q[1] = qx; o.q[2] = qy; o.q[3] = qz;
    o.safeNorm = safeNorm; o.norm = norm;
}

struct Cov2d {
    float t0, t1, t2, clipX, clipY, limX, limY, tx, ty;
    float b[6], t[6];
    float c[4];  // c00 c01 c10 c11
};

// shared.slang:170-243 (incl. the clamp(t.z) quirk at :202-205)
__device__ __forceinline__ void build_cov2d(float m0, float m1, float m2, const float* S, const ViewParams& vp, Cov2d& o)
{
    const float* V = vp.V;
    float a00 = V[0], a01 = V[1], a02 = V[2], a10 = V[4], a11 = V[5], a12 = V[6], a20 = V[8], a21 = V[9], a22 = V[10];
    float t0 = m0 * a00 + m1 * a10 + m2 * a20 + V[12];
    float t1 = m0 * a01 + m1 * a11 + m2 * a21 + V[13];
    float t2 = m0 * a02 + m1 * a12 + m2 * a22 + V[14];
    float limX = vp.tanHalfX * 1.3f, limY = vp.tanHalfY * 1.3f;
    float clipX = fminf(fmaxf(t2, -vp.tanHalfX * 1.3f), limX);
    float clipY = fminf(fmaxf(t2, -vp.tanHalfY * 1.3f), limY);
    float tx = t0 / clipX * t2;
    float ty = t1 / clipY * t2;
    float tz = t2;
    float j00 = vp.focalX / tz;
    float j02 = -tx * vp.focalX / (tz * tz);
    float j11 = vp.focalY / tz;
    float j12 = -ty * vp.focalY / (tz * tz);
    float b00 = j00 * a00 + j02 * a02;
    float b01 = j00 * a10 + j02 * a12;
    float b02 = j00 * a20 + j02 * a22;
    float b10 = j11 * a01 + j12 * a02;
    float b11 = j11 * a11 + j12 * a12;
    float b12 = j11 * a21 + j12 * a22;
    float t00 = b00 * S[0] + b01 * S[3] + b02 * S[6];
    float t01 = b00 * S[1] + b01 * S[4] + b02 * S[7];
    float t02 = b00 * S[2] + b01 * S[5] + b02 * S[8];
    float t10 = b10 * S[0] + b11 * S[3] + b12 * S[6];
    float t11 = b10 * S[1] + b11 * S[4] + b12 * S[7];
    float t12 = b10 * S[2] + b11 * S[5] + b12 * S[8];
    o.c[0] = t00 * b00 + t01 * b01 + t02 * b02 + 0.3f;
    o.c[1] = t00 * b10 + t01 * b11 + t02 * b12;
    o.c[2] = t10 * b00 + t11 * b01 + t12 * b02;
    o.c[3] = t10 * b10 + t11 * b11 + t12 * b12 + 0.3f;
    o.t0 = t0; o.t1 = t1; o.t2 = t2; o.clipX = clipX; o.clipY = clipY; o.limX = limX; o.limY = limY; o.tx = tx; o.ty = ty;
    o.b[0] = b00; o.b[1] = b01; o.b[2] = b02; o.b[3] = b10; o.b[4] = b11; o.b[5] = b12;
    o.t[0] = t00; o.t[1] = t01; o.t[2] = t02; o.t[3] = t10; o.t[4] = t11; o.t[5] = t12;
}

struct ProjOut {
    float sx, sy, depth;
    float color[3];
    float cov2d[4], conic[4];
    float radius;                // visible radius (0 when culled)
    float rect[4];               // minX minY maxX maxY (clamped)
};

// SH colour of K1 (shared.slang:257-319): un-normalised direction, max(sum + 0.5, 0).  It depends on the position, the
// camera centre and the SH coefficients only, so it can run as its own kernel (project.cu: the data-parallel step evaluates
// it after the binning, when the SH parameters of the previous step's exchange have landed).
template <int MAXK, class ShFun>
__device__ __forceinline__ void project_color(float m0, float m1, float m2, ShFun sh, const ViewParams& vp, float* color)
{
    float dx = m0 - vp.cam[0], dy = m1 - vp.cam[1], dz = m2 - vp.cam[2];
    float basis[MAXK];
    sh_basis<MAXK>(dx, dy, dz, vp.degree, basis);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = basis[0] * sh(0, c);
#pragma unroll
        for (int k = 1; k < MAXK; ++k)
            if (k < vp.coeffCount) acc += basis[k] * sh(k, c);
        acc += 0.5f;
        color[c] = fmaxf(acc, 0.0f);
    }
}

// K1 body (kernels.slang:36-173).  shfun(k, c) returns SH coefficient (k, c).  PART: bit 0 = geometry, bit 1 = colour.
template <int MAXK, int PART = 3, class ShFun>
__device__ __forceinline__ void project_forward(float m0, float m1, float m2, float s0, float s1, float s2, float rw,
                                                float rx, float ry, float rz, ShFun sh, const ViewParams& vp, ProjOut& o)
{
    const float* V = vp.V;
    const float* P = vp.P;
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
    o.sx = ((ndcX + 1.0f) * vp.imageW - 1.0f) * 0.5f;
    o.sy = ((ndcY + 1.0f) * vp.imageH - 1.0f) * 0.5f;
    o.depth = pv2;
    if (PART & 2) project_color<MAXK>(m0, m1, m2, sh, vp, o.color);
    if (!(PART & 1)) return;
    Cov3d c3;
    build_cov3d(s0, s1, s2, rw, rx, ry, rz, c3);
    Cov2d c2;
    build_cov2d(m0, m1, m2, c3.S, vp, c2);
    float det = c2.c[0] * c2.c[3] - c2.c[1] * c2.c[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) o.cov2d[i] = c2.c[i];
    o.conic[0] = c2.c[3] / det;
    o.conic[1] = -c2.c[1] / det;
    o.conic[2] = -c2.c[2] / det;
    o.conic[3] = c2.c[0] / det;
    float mid = 0.5f * (c2.c[0] + c2.c[3]);
    float delta = fmaxf(mid * mid - det, 1e-5f);
    float lambdaMax = mid + sqrtf(delta);
    float radius = 3.0f * ceilf(sqrtf(lambdaMax));
    float vr = radius * visibleMask;
    o.radius = vr;
    float maxX = vp.imageW - 1.0f, maxY = vp.imageH - 1.0f;
    float minX = o.sx - vr, minY = o.sy - vr, maxRX = o.sx + vr, maxRY = o.sy + vr;
    if (minX < 0.0f) minX = 0.0f;
    if (minY < 0.0f) minY = 0.0f;
    if (maxRX > maxX) maxRX = maxX;
    if (maxRY > maxY) maxRY = maxY;
    o.rect[0] = minX; o.rect[1] = minY; o.rect[2] = maxRX; o.rect[3] = maxRY;
}

// K3 tile rectangle (tile_global.slang:39-57): returns (x0,y0,x1,y1) clamped to the grid.
__device__ __forceinline__ void tile_rect(const float* rect, const ViewParams& vp, int& x0, int& y0, int& x1, int& y1)
{
    int tMinX = (int)floorf(rect[0] / (float)vp.tileW);
    int tMinY = (int)floorf(rect[1] / (float)vp.tileH);
    int tMaxX = (int)floorf(rect[2] / (float)vp.tileW) + 1;
    int tMaxY = (int)floorf(rect[3] / (float)vp.tileH) + 1;
    x0 = max(0, min(tMinX, vp.gridW));
    y0 = max(0, min(tMinY, vp.gridH));
    x1 = max(0, min(tMaxX, vp.gridW));
    y1 = max(0, min(tMaxY, vp.gridH));
}

struct ProjGrad {
    float gm[3], gs[3], gr[4], gcam[3];
};

// K2 body: hand-derived VJP of project_forward (reference: Slang reverse-AD, kernels.slang:205-398).
// Conventions of the shipped MSL header: max ties → 0.5, clamp passes gradient iff lo<=x<=hi,
// sqrt' = 0.5/sqrt(max(1e-7,x)).  gshfun(k, c, v) stores d/d sh(k,c).
template <int MAXK, class ShFun, class GShFun>
__device__ __forceinline__ void project_backward(float m0, float m1, float m2, float s0, float s1, float s2, float rw,
                                                 float rx, float ry, float rz, ShFun sh, const ViewParams& vp,
                                                 float cotDepth, float cotMx, float cotMy, const float* cotCov2d,
                                                 const float* cotColor, const float* cotConic, GShFun gsh, ProjGrad& g)
{
    const float* V = vp.V;
    const float* P = vp.P;
    float gm0 = 0.f, gm1 = 0.f, gm2 = 0.f;
    // ---- colour ----
    float dx = m0 - vp.cam[0], dy = m1 - vp.cam[1], dz = m2 - vp.cam[2];
    float basis[MAXK];
    sh_basis<MAXK>(dx, dy, dz, vp.degree, basis);
    float gpre[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = basis[0] * sh(0, c);
#pragma unroll
        for (int k = 1; k < MAXK; ++k)
            if (k < vp.coeffCount) acc += basis[k] * sh(k, c);
        acc += 0.5f;
        float gc = cotColor[c];
        gpre[c] = acc > 0.0f ? gc : (acc < 0.0f ? 0.0f : 0.5f * gc);
    }
    // NOTE: gsh may alias the storage sh reads from (in-place smem), so read everything first.
    float gdir[3];
    sh_basis_grad_dot<MAXK>(dx, dy, dz, vp.degree,
                            [&](int k) { return sh(k, 0) * gpre[0] + sh(k, 1) * gpre[1] + sh(k, 2) * gpre[2]; }, gdir);
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k < vp.coeffCount) {
            gsh(k, 0, basis[k] * gpre[0]);
            gsh(k, 1, basis[k] * gpre[1]);
            gsh(k, 2, basis[k] * gpre[2]);
        }
    gm0 += gdir[0]; gm1 += gdir[1]; gm2 += gdir[2];
    g.gcam[0] = -gdir[0]; g.gcam[1] = -gdir[1]; g.gcam[2] = -gdir[2];
    // ---- covariance chain ----
    Cov3d c3;
    build_cov3d(s0, s1, s2, rw, rx, ry, rz, c3);
    Cov2d c2;
    build_cov2d(m0, m1, m2, c3.S, vp, c2);
    float c00 = c2.c[0], c01 = c2.c[1], c10 = c2.c[2], c11 = c2.c[3];
    float det = c00 * c11 - c01 * c10;
    float invdet = 1.0f / det;
    float q0 = cotConic[0], q1 = cotConic[1], q2 = cotConic[2], q3 = cotConic[3];
    float gdet = (-q0 * c11 + q1 * c01 + q2 * c10 - q3 * c00) * invdet * invdet;
    float G00 = cotCov2d[0] + q3 * invdet + gdet * c11;
    float G01 = cotCov2d[1] - q1 * invdet - gdet * c10;
    float G10 = cotCov2d[2] - q2 * invdet - gdet * c01;
    float G11 = cotCov2d[3] + q0 * invdet + gdet * c00;
    const float* b = c2.b;
    const float* t = c2.t;
    float gt[6], gb[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        gt[k] = G00 * b[k] + G01 * b[3 + k];
        gt[3 + k] = G10 * b[k] + G11 * b[3 + k];
        gb[k] = G00 * t[k] + G10 * t[3 + k];
        gb[3 + k] = G01 * t[k] + G11 * t[3 + k];
    }
    float gS[9];
    const float* S = c3.S;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) gS[k * 3 + c] = b[k] * gt[c] + b[3 + k] * gt[3 + c];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        gb[k] += gt[0] * S[k * 3 + 0] + gt[1] * S[k * 3 + 1] + gt[2] * S[k * 3 + 2];
        gb[3 + k] += gt[3] * S[k * 3 + 0] + gt[4] * S[k * 3 + 1] + gt[5] * S[k * 3 + 2];
    }
    float a00 = V[0], a01 = V[1], a02 = V[2], a10 = V[4], a11 = V[5], a12 = V[6], a20 = V[8], a21 = V[9], a22 = V[10];
    float gj00 = gb[0] * a00 + gb[1] * a10 + gb[2] * a20;
    float gj02 = gb[0] * a02 + gb[1] * a12 + gb[2] * a22;
    float gj11 = gb[3] * a01 + gb[4] * a11 + gb[5] * a21;
    float gj12 = gb[3] * a02 + gb[4] * a12 + gb[5] * a22;
    float tz = c2.t2, tz2 = tz * tz, tz3 = tz2 * tz;
    float gtz = -vp.focalX / tz2 * gj00 - vp.focalY / tz2 * gj11 + 2.0f * c2.tx * vp.focalX / tz3 * gj02 +
                2.0f * c2.ty * vp.focalY / tz3 * gj12;
    float gtx = -vp.focalX / tz2 * gj02;
    float gty = -vp.focalY / tz2 * gj12;
    float gt0 = gtx * c2.t2 / c2.clipX;
    float gclipX = -gtx * c2.t0 * c2.t2 / (c2.clipX * c2.clipX);
    float gt2 = gtx * c2.t0 / c2.clipX;
    float gt1 = gty * c2.t2 / c2.clipY;
    float gclipY = -gty * c2.t1 * c2.t2 / (c2.clipY * c2.clipY);
    gt2 += gty * c2.t1 / c2.clipY;
    if (c2.t2 >= -c2.limX && c2.t2 <= c2.limX) gt2 += gclipX;
    if (c2.t2 >= -c2.limY && c2.t2 <= c2.limY) gt2 += gclipY;
    gt2 += gtz;
    gm0 += gt0 * a00 + gt1 * a01 + gt2 * a02;
    gm1 += gt0 * a10 + gt1 * a11 + gt2 * a12;
    gm2 += gt0 * a20 + gt1 * a21 + gt2 * a22;
    // ---- NDC path ----
    float pv0 = m0 * V[0] + m1 * V[4] + m2 * V[8] + V[12];
    float pv1 = m0 * V[1] + m1 * V[5] + m2 * V[9] + V[13];
    float pv2 = m0 * V[2] + m1 * V[6] + m2 * V[10] + V[14];
    float pv3 = m0 * V[3] + m1 * V[7] + m2 * V[11] + V[15];
    float pc0 = pv0 * P[0] + pv1 * P[4] + pv2 * P[8] + pv3 * P[12];
    float pc1 = pv0 * P[1] + pv1 * P[5] + pv2 * P[9] + pv3 * P[13];
    float pc3 = pv0 * P[3] + pv1 * P[7] + pv2 * P[11] + pv3 * P[15];
    float wInv = 1.0f / (pc3 + 0.000001f);
    float gndcX = cotMx * 0.5f * vp.imageW;
    float gndcY = cotMy * 0.5f * vp.imageH;
    float gpc0 = gndcX * wInv, gpc1 = gndcY * wInv;
    float gwInv = gndcX * pc0 + gndcY * pc1;
    float gpc3 = -gwInv * wInv * wInv;
    float gpv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gpv[j] = gpc0 * P[j * 4 + 0] + gpc1 * P[j * 4 + 1] + gpc3 * P[j * 4 + 3];
    gpv[2] += cotDepth;
    gm0 += gpv[0] * V[0] + gpv[1] * V[1] + gpv[2] * V[2] + gpv[3] * V[3];
    gm1 += gpv[0] * V[4] + gpv[1] * V[5] + gpv[2] * V[6] + gpv[3] * V[7];
    gm2 += gpv[0] * V[8] + gpv[1] * V[9] + gpv[2] * V[10] + gpv[3] * V[11];
    g.gm[0] = gm0; g.gm[1] = gm1; g.gm[2] = gm2;
    // ---- cov3d = L L^T ----
    const float* L = c3.L;
    float gL[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc += (gS[r * 3 + c] + gS[c * 3 + r]) * L[c * 3 + k];
            gL[r * 3 + k] = acc;
        }
    float qw = c3.q[0], qx = c3.q[1], qy = c3.q[2], qz = c3.q[3];
    float R[9] = {1.0f - 2.0f * (qy * qy + qz * qz), 2.0f * (qx * qy - qw * qz), 2.0f * (qx * qz + qw * qy),
                  2.0f * (qx * qy + qw * qz), 1.0f - 2.0f * (qx * qx + qz * qz), 2.0f * (qy * qz - qw * qx),
                  2.0f * (qx * qz - qw * qy), 2.0f * (qy * qz + qw * qx), 1.0f - 2.0f * (qx * qx + qy * qy)};
    float sv[3] = {s0, s1, s2};
    float gR[9];
    g.gs[0] = g.gs[1] = g.gs[2] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            g.gs[k] += gL[r * 3 + k] * R[r * 3 + k];
            gR[r * 3 + k] = gL[r * 3 + k] * sv[k];
        }
    float gqw = 2.0f * (-qz * gR[1] + qy * gR[2] + qz * gR[3] - qx * gR[5] - qy * gR[6] + qx * gR[7]);
    float gqx = 2.0f * (qy * gR[1] + qz * gR[2] + qy * gR[3] - 2.0f * qx * gR[4] - qw * gR[5] + qz * gR[6] + qw * gR[7] -
                        2.0f * qx * gR[8]);
    float gqy = 2.0f * (-2.0f * qy * gR[0] + qx * gR[1] + qw * gR[2] + qx * gR[3] + qz * gR[5] - qw * gR[6] + qz * gR[7] -
                        2.0f * qy * gR[8]);
    float gqz = 2.0f * (-2.0f * qz * gR[0] - qw * gR[1] + qx * gR[2] + qw * gR[3] - 2.0f * qz * gR[4] + qy * gR[5] +
                        qx * gR[6] + qy * gR[7]);
    float sn = c3.safeNorm;
    float gsn = -(gqw * rw + gqx * rx + gqy * ry + gqz * rz) / (sn * sn);
    float gnorm = c3.norm > 1e-8f ? gsn : (c3.norm < 1e-8f ? 0.0f : 0.5f * gsn);
    float n2 = rw * rw + rx * rx + ry * ry + rz * rz;
    float gn2 = 0.5f / sqrtf(fmaxf(1e-7f, n2)) * gnorm;
    g.gr[0] = gqw / sn + gn2 * 2.0f * rw;
    g.gr[1] = gqx / sn + gn2 * 2.0f * rx;
    g.gr[2] = gqy / sn + gn2 * 2.0f * ry;
    g.gr[3] = gqz / sn + gn2 * 2.0f * rz;
}

}  // namespace gsb
