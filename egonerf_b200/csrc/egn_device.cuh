// Shared host/device definitions for libegn_b200 (sm_100a).  See include/egn.h for the C ABI and
// DESIGN.md for the data layout.  Reference citations are into changwoonchoi/EgoNeRF.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/egn.h"

#define EGN_CS 16            // density channels per plane   (n_lamb_sigma, configs/EgoNeRF/common.txt:29)
#define EGN_CA 48            // appearance channels per plane (n_lamb_sh,    configs/EgoNeRF/common.txt:30)
#define EGN_CF (EGN_CS + EGN_CA)   // interleaved channels per texel: one tap = 256 contiguous bytes
#define EGN_FEAT_STRIDE 28   // app feature row (27 used) padded to a multiple of 16 bytes
#define EGN_MAX_KNOTS 1024
#define EGN_FUSED_MAX_KNOTS 384   // r-ladder entries the fused fine pass keeps in shared memory (N_r <= 381)
#define EGN_HID 128

// matMode [[0,1],[0,2],[1,2]] / vecMode [2,1,0] (EgoNeRF.py:30-33): plane i is indexed x = c[MX[i]] (width),
// y = c[MY[i]] (height); line i by c[VL[i]].
__host__ __device__ __forceinline__ int egn_mx(int i) { return i == 2 ? 1 : 0; }
__host__ __device__ __forceinline__ int egn_my(int i) { return i == 0 ? 1 : 2; }
__host__ __device__ __forceinline__ int egn_vl(int i) { return 2 - i; }

// Offsets (in floats) of every section of the packed render tables.
struct EgnLayout {
    long long pf[2][3];   // fine planes   [H][W][EGN_CF]
    long long lf[2][3];   // fine lines    [L][EGN_CF]
    long long pc[2][3];   // coarse planes [H/2][W/2][EGN_CS]   (AvgPool2d(2,2), EgoNeRF.py:128)
    long long lc[2][3];   // coarse lines  [L/2][EGN_CS]        (AvgPool1d(2,2), EgoNeRF.py:129)
    long long total;
    int G[3];             // N_r, N_theta, N_phi
    int Gc[3];            // floor(G/2)
};

static inline EgnLayout egn_make_layout(const int32_t grid[3]) {
    EgnLayout L;
    for (int a = 0; a < 3; ++a) { L.G[a] = grid[a]; L.Gc[a] = grid[a] / 2; }
    long long off = 0;
    auto take = [&](long long n) { long long o = off; off += (n + 63) / 64 * 64; return o; };
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) {
            L.pf[h][i] = take((long long)L.G[egn_my(i)] * L.G[egn_mx(i)] * EGN_CF);
            L.lf[h][i] = take((long long)L.G[egn_vl(i)] * EGN_CF);
        }
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) {
            L.pc[h][i] = take((long long)L.Gc[egn_my(i)] * L.Gc[egn_mx(i)] * EGN_CS);
            L.lc[h][i] = take((long long)L.Gc[egn_vl(i)] * EGN_CS);
        }
    L.total = off;
    return L;
}

// "Half" tables of the throughput mode (fused fine pass): every fine texel of the 12 plane / line sections gets ONE global
// texel index (sections in the order [h][i] plane, line), and two arrays indexed by it:
//   app  [texel][64] fp16 : the 48 appearance channels + 16 halfs of zero padding = one aligned 128-byte line per tap
//   dens [texel][16] fp32 : the density channels, exact (alpha stays inside the 1e-4 parity bound)
// laid out as [app array][dens array] in one buffer.
#define EGN_APP_TEXEL_HALFS 64
struct EgnLayoutH {
    int texp[2][3];       // first global texel of fine plane [h][i]
    int texl[2][3];       // first global texel of fine line  [h][i]
    long long n_texels;
    long long dens_byte_offset;   // = n_texels * 128
    long long total_bytes;        // = n_texels * (128 + 64)
};
static inline EgnLayoutH egn_make_layout_h(const int32_t grid[3]) {
    EgnLayoutH L;
    long long t = 0;
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) {
            L.texp[h][i] = (int)t; t += (long long)grid[egn_my(i)] * grid[egn_mx(i)];
            L.texl[h][i] = (int)t; t += grid[egn_vl(i)];
        }
    L.n_texels = t;
    L.dens_byte_offset = t * (EGN_APP_TEXEL_HALFS * 2);
    L.total_bytes = t * (EGN_APP_TEXEL_HALFS * 2 + EGN_CS * 4);
    return L;
}

// Everything a render kernel needs, passed by value.
struct EgnKernelCfg {
    EgnLayout lay;
    const float* tables;
    const float* r_knots;     // knots_last + 1 entries (fine pass)
    const float* r_knots_c;   // knots_last_c + 1 entries (coarse pass; the same ladder unless plain_ladders)
    int knots_last, knots_last_c;   // last knot index = upper clamp of the search
    int r_div, r_div_c;       // cells of the fine / coarse-pass ladder (N_r; coarse: N_r/2 with plain ladders)
    int plain_ladders;
    float jitter_ratio, jitter_r0;
    const float* z_coarse;    // n_coarse
    float center[3];
    float ang_near[2], ang_inv[2];
    float density_shift, distance_scale;
    int n_coarse, n_fine, S, use_coarse_sample, resampling;
    int fea2dense, shading, app_dim, view_pe, fea_pe, env_h;
    int mlp_mode;             // EGN_MLP_*
    int march;                // 1: uniform march (TensorBase.sample_ray) instead of the exponential schedule
    float step_size, far_plane, aabb[6];
    const void* tables_bf16;  // optional bf16 copy of the fine tables (same element offsets)
    const void* tables_h;     // half tables of the fused fine pass (EgnLayoutH), or NULL
    int texp[2][3], texl[2][3];   // EgnLayoutH section starts (global texel indices)
    long long dens_byte_offset;
    float* coords;            // optional [samples][4]: {r, polar, azimuth (index space), hemisphere} written by the fused training
                              // forward and read back by the tcgen05 gather backward instead of recomputing acosf / atan2f / the
                              // knot search per sample (bit-identical: same function, same inputs); NULL = recompute
};

#ifdef __CUDACC__
// -------------------------------------------------------------------------------------------------
// device helpers
// -------------------------------------------------------------------------------------------------
#define EGN_PI_4   0.78539816339744830962f   // fp32(pi/4)   : thresholds are compared in fp32 (coordinates.py:483-486)
#define EGN_3PI_4  2.35619449019234492885f   // fp32(3pi/4)

struct YYCoord {       // index-space coordinate of one sample in its active hemisphere
    float c[3];        // normalised [-1,1]: r, polar, azimuth
    int yang;
};

// YinYangSphericalCoords.from_cartesian (coordinates.py:468-498) followed by normalize_coord (:442-466) and
// GenericSphericalCoords.normalize_r, interval_th branch (:112-131,156).  IEEE sqrtf/acosf/atan2f, no FMA
// contraction (the library is compiled with -fmad=false): a 1-ulp change flips the hemisphere of a sample.
__device__ __forceinline__ YYCoord egn_cart_to_yinyang(float px, float py, float pz, const EgnKernelCfg& k,
                                                       const float* __restrict__ knots /*smem or global*/, int last, int n_r) {
    float qx = px - k.center[0], qy = py - k.center[1], qz = pz - k.center[2];
    float r = sqrtf(qx * qx + qy * qy + qz * qz);
    float th = acosf(qz / r);
    if (th != th) th = 0.f;                       // nan_to_num_ (origin)
    float ph = atan2f(qy, qx);
    bool yin = (EGN_PI_4 <= th) && (th <= EGN_3PI_4) && (-EGN_3PI_4 <= ph) && (ph <= EGN_3PI_4);
    YYCoord o;
    if (!yin) {
        th = acosf(qy / r);
        if (th != th) th = 0.f;
        ph = atan2f(qz, -qx);
    }
    o.yang = yin ? 0 : 1;
    o.c[1] = (th - k.ang_near[0]) * k.ang_inv[0] * 2.f - 1.f;
    o.c[2] = (ph - k.ang_near[1]) * k.ang_inv[1] * 2.f - 1.f;
    // searchsorted(knots, r, right=True) clamped to [1, last]; with the plain ladders in + frac == 1 + k + lin of
    // coordinates.py:141-155 (knots[i] = r0 * ratio^(i-1)), r < r0 falls in cell 0: frac = r / r0
    int lo = 0, hi = last + 1;                     // first index with knots[idx] > r
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (knots[mid] > r) hi = mid; else lo = mid + 1;
    }
    int out = min(max(lo, 1), last);
    int in = out - 1;
    float g0 = knots[in], g1 = knots[out];
    float frac = (r - g0) / (g1 - g0);
    o.c[0] = ((float)in + frac) / (float)n_r * 2.f - 1.f;
    return o;
}

__device__ __forceinline__ YYCoord egn_cart_to_yinyang(float px, float py, float pz, const EgnKernelCfg& k,
                                                       const float* __restrict__ knots) {       // fine pass
    return egn_cart_to_yinyang(px, py, pz, k, knots, k.knots_last, k.r_div);
}

// sample index -> ray index.  S is a multiple of 32 (validated: n_coarse, n_fine % 32 == 0), so m / S == (m >> 5) / (S >> 5): a
// 32-bit division (~20 instructions) instead of the ~100-instruction 64-bit one; exact for m < 2^37
__device__ __forceinline__ long long egn_ray_of(long long m, int S) {
    return (long long)((unsigned)(m >> 5) / (unsigned)(S >> 5));
}

// F.grid_sample(align_corners=True) un-normalisation: ((x+1)/2)*(size-1)
__device__ __forceinline__ float egn_unnorm(float x, int size) { return ((x + 1.f) / 2.f) * (float)(size - 1); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4fma(float w, float4 a, float4 acc) {
    return make_float4(fmaf(w, a.x, acc.x), fmaf(w, a.y, acc.y), fmaf(w, a.z, acc.z), fmaf(w, a.w, acc.w));
}
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// softplus(beta=1, threshold=20) (tensorBase.py:415-417, torch semantics) / relu
__device__ __forceinline__ float egn_density_act(float f, float shift, int act) {
    if (act == EGN_ACT_SOFTPLUS) {
        float x = f + shift;
        return x > 20.f ? x : log1pf(expf(x));
    }
    return fmaxf(f, 0.f);
}
__device__ __forceinline__ float egn_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// Philox4x32-10 counter-based generator (train-mode jitter when the caller passes no uniforms).
__device__ __forceinline__ uint4 egn_philox(uint64_t seed, uint64_t ctr_lo, uint32_t ctr_hi) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = ctr_hi, c3 = 0x9E3779B9u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float egn_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0,1)
#endif  // __CUDACC__
