// Device helpers shared by the forward and backward kernels (envmap lookup, SH / RGB sample colour).
#pragma once
#include "egn_device.cuh"

#define FULL 0xffffffffu
__device__ __forceinline__ float hsum4(float4 v) { return (v.x + v.y) + (v.z + v.w); }

// =================================================================================================
// Envmap (models/envmap.py:6-34): equirect (3, 2h, h) bilinear + sigmoid.
// =================================================================================================
struct EnvTap { int x0, y0; float fx, fy; };
__device__ __forceinline__ EnvTap egn_env_tap(float dx, float dy, float dz, int h) {
    const float nrm = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);     // F.normalize eps
    dx /= nrm; dy /= nrm; dz /= nrm;
    const float u = (dz + 1.f) * 0.5f;
    const float v = (atan2f(dy, dx) + 3.14159265358979323846f) / 6.28318530717958647692f;
    const float ix = egn_unnorm(2.f * u - 1.f, h), iy = egn_unnorm(2.f * v - 1.f, 2 * h);
    EnvTap t;
    const float flx = floorf(ix), fly = floorf(iy);
    t.x0 = (int)flx; t.y0 = (int)fly; t.fx = ix - flx; t.fy = iy - fly;
    return t;
}
__device__ __forceinline__ void egn_env_radiance(const float* __restrict__ em, int h, float dx, float dy, float dz, float out[3]) {
    const EnvTap t = egn_env_tap(dx, dy, dz, h);
    const int W = h, H = 2 * h;
    const float w[4] = {(1.f - t.fx) * (1.f - t.fy), t.fx * (1.f - t.fy), (1.f - t.fx) * t.fy, t.fx * t.fy};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int x = t.x0 + (q & 1), y = t.y0 + (q >> 1);
            if (x >= 0 && x < W && y >= 0 && y < H) acc = fmaf(w[q], __ldg(em + ((long long)ch * H + y) * W + x), acc);
        }
        out[ch] = egn_sigmoid(acc);
    }
}


// =================================================================================================
// Per-sample colour for the compositing kernels (tensorBase.py:30-39, models/sh.py:87-116)
// =================================================================================================
#define K4_MAXE 16

__device__ __forceinline__ void egn_sample_color(const EgnKernelCfg& k, const float* __restrict__ feat,
                                                 const float* __restrict__ rgbs, long long m, const float sh[9], float c[3]) {
    if (k.shading == EGN_SHADE_RGB) {              // RGBRender (tensorBase.py:37-39)
        c[0] = feat[m * EGN_FEAT_STRIDE]; c[1] = feat[m * EGN_FEAT_STRIDE + 1]; c[2] = feat[m * EGN_FEAT_STRIDE + 2];
    } else if (k.shading == EGN_SHADE_SH) {        // SHRender (tensorBase.py:30-34)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float a = 0.f;
#pragma unroll
            for (int b = 0; b < 9; ++b) a += sh[b] * feat[m * EGN_FEAT_STRIDE + ch * 9 + b];
            c[ch] = fmaxf(a + 0.5f, 0.f);
        }
    } else {
        c[0] = rgbs[m * 3]; c[1] = rgbs[m * 3 + 1]; c[2] = rgbs[m * 3 + 2];
    }
}
__device__ __forceinline__ void egn_sh_basis(float x, float y, float z, float sh[9]) {   // models/sh.py:87-116
    sh[0] = 0.28209479177387814f;
    sh[1] = -0.4886025119029199f * y; sh[2] = 0.4886025119029199f * z; sh[3] = -0.4886025119029199f * x;
    const float xx = x * x, yy = y * y, zz = z * z;
    sh[4] = 1.0925484305920792f * (x * y); sh[5] = -1.0925484305920792f * (y * z);
    sh[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    sh[7] = -1.0925484305920792f * (x * z); sh[8] = 0.5462742152960396f * (xx - yy);
}

