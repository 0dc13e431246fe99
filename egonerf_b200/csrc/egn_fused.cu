// Fused fine pass (throughput mode, EGN_MLP_TC_F16): Yin-Yang coordinates -> 18-tap factor gather -> VM products
// -> basis contraction -> positional encoding -> 3-layer MLP -> sample colour, in ONE persistent warp-specialised kernel.
// Replaces egn_gather_kernel + egn_mlp_*_kernel for one ray chunk (EgoNeRF.py:544-556: from_cartesian / normalize_coord,
// compute_densityfeature, compute_appfeature, renderModule).  One CTA per SM, 800 threads at 72 registers (FU_GATHER_WARPS = 16;
// the round-1 cut, 8 gather warps x 16 rows at 128 registers with the MMA issue inside MLP warp 0, is still selectable):
//
//   warps 8..23  GATHER group   per 128-sample tile, each warp owns 8 rows and runs three phases on them:
//                  1. ADDRESS   one lane per sample: coordinates (every fourth tile for four tiles at once, all 32 lanes busy),
//                               then per-axis texel indices + tap weights -> the 18 global texel indices and 18 fp32 weights
//                               of the sample go to a 144-byte record in shared memory.  Done ONCE per sample instead of
//                               once per lane of the sample (the r01 kernel spent half its gather instructions here).
//                  2. DENSITY   4 lanes per sample, one pass of 8 samples: 16 fp32 channels per tap (exact: alpha stays inside
//                               the parity bound), FFMA interpolation, relu-sum -> sigma feature.
//                  3. APPEARANCE 8 lanes per sample (6 active), two passes of 4 samples: 48 fp16 channels per tap = one aligned
//                               128-byte line, packed half2 interpolation (HFMA2: two channels per instruction, no unpacking),
//                               the 144 products P*L go straight into shared memory as an fp16 row of the tcgen05 A operand V.
//                               V is double buffered: tile i+1 is gathered while tile i runs through the MLP.
//                 Two factor pairs of taps (12 loads per lane) are in flight per warp, 192 per SM.
//   warps 0..7   MLP group      feat2 = V [B_yin | B_yang]^T (N = 64; the epilogue picks the sample's hemisphere) -> PE -> X -> D1 ->
//                               relu (folded into the fp16 conversion) -> H1 -> D2 -> + b2, relu -> H2 -> D3 = H2 W3^T (N = 16) ->
//                               sigmoid [-> compositing].  X, H1 and H2 share one operand buffer; the warps announce each finished
//                               operand on an mbarrier and run on (no group barrier in front of an MMA)
//   warp 24      MMA warp       one elected thread issues every tcgen05.mma in the order layer 1, layer 2, layer 0 of the next
//                               tile, layer 3; completions come back through tcgen05.commit -> mbarrier
//
// All MMA operands are FP16 (fp32 accumulate in TMEM): against bf16 the rounding error of every operand is 8x smaller, which
// brings rgb to ~1e-5 of the exact render (scripts/error_budget.py) at the same bytes and tensor throughput.
// Nothing of size (samples x features) touches HBM in between: per sample the kernel reads 24 B of ray, 4 B of depth and
// its taps, and writes 4 B (sigma feature) + 12 B (colour) [+ 112 B app feature when the backward pass will need it].
#include <cuda_fp16.h>
#include <type_traits>
#include "egn_tc.cuh"
#include "egn_host.h"
#include "egn_shared.cuh"

#ifndef FU_L3_MMA
#define FU_L3_MMA 1                          // layer 3 as a fourth MMA: H2 = relu(D2 + b2) goes back to shared memory as an fp16 operand, D3 = H2 . W3^T (N = 16)
#endif
#ifndef FU_GATHER_WARPS
#define FU_GATHER_WARPS 16                   // 8: 16 rows of every tile per gather warp (512 threads, 128 registers); 16: 8 rows (768 threads, 80 registers)
#endif
#ifndef FU_DEPTH
#define FU_DEPTH 2                           // factor pairs of taps in flight per gather warp in the 16-warp variant (3 spills at 80 registers: 6.03 ms)
#endif
#ifndef FU_MMA_WARP
#define FU_MMA_WARP (FU_GATHER_WARPS == 16 && FU_L3_MMA)     // a 25th warp issues every tcgen05.mma (800 threads x 80 registers)
#endif
#ifndef FU_ALPHA_IN_GATHER
#define FU_ALPHA_IN_GATHER 0                 // 1: the density lanes of the gather group turn the sigma feature into alpha (needs FU_MMA_WARP); measured 4.88 vs 4.82 ms: off
#endif
#define FU_THREADS (256 + 32 * FU_GATHER_WARPS + 32 * FU_MMA_WARP)
#define FU_GROUP 256
#define FU_VK (3 * EGN_CA)                   // 144
#define FU_VCHUNKS (FU_VK / 8)               // 18
// K-chunk stride of the V operand: 2048 B of data + 16 B of padding.  The 6 appearance lanes of a sample store chunks
// kc, kc+1, .. of the SAME row; with a 2048 B stride they would all fall into one 16-byte bank group (6-way conflict), with
// 2064 B consecutive chunks land in consecutive bank groups: conflict-free.
#define FU_VCHUNK (TC_CHUNK + 16)
#define FU_VBYTES (FU_VCHUNKS * FU_VCHUNK)   // 37 152
#define FU_BB_CHUNK 1024                     // basis operand: 64 rows x 16 B per K chunk
#define FU_REC_WORDS 36                      // address record of one sample: 3 x {4 plane texels, 2 line texels, 4 + 2 weights}
#ifndef FU_VBUFS
#define FU_VBUFS 2                           // V operand buffers: 2 = gather of tile i+1 overlaps layer 0 of tile i; 1 = -37 KB shared memory
#endif
#ifndef FU_APP_UNROLL
#define FU_APP_UNROLL 1
#endif
#define FU_STR(x) #x
#define FU_UNROLL(n) _Pragma(FU_STR(unroll n))
#ifndef FU_PAIRED_W
#define FU_PAIRED_W 0                        // measured: 5.50 ms with the paired conversions vs 5.41 ms (profiles/r02_fused.md)
#endif
#ifndef FU_MLP_BACKOFF
#define FU_MLP_BACKOFF 0                     // ns between polls of the MLP group's wait for the next gathered tile (256 / 1024: no effect)
#endif
#define FU_IDESC_128x128 0x08200010u         // kind::f16: D fp32, A/B fp16, both K-major, N = 128, M = 128
#define FU_IDESC_128x64  0x08100010u         // same, N = 64
#define FU_IDESC_128x16  0x08040010u         // same, N = 16
#define FU_W3_CHUNK 128                      // layer-3 operand: 8 rows (3 used) x 16 B per K chunk; the N = 16 tile reads them twice (SBO = 0)

// Layer-3 table {b2, W3[0], W3[1], W3[2]} per hidden unit, as constant-bank operands of the epilogue's FADD / FFMA (FU_L3_CONST):
// the 64 broadcast LDS.128 per thread and tile it replaces were 1 053 of the 3 122 shared-memory wavefronts of a tile, i.e. 13 %
// of the traffic on the L1 data pipe the kernel is bound by (profiles/r02_fused.md).  The symbol is refreshed by a device-to-
// device copy from the operand image before every launch, on the launch's stream; launches on DIFFERENT streams are ordered
// against each other by an event per device (a later copy waits for the earlier kernel), under a host mutex.
#ifndef FU_FAST_SIGMOID
#define FU_FAST_SIGMOID 1                    // sample colour with ex2.approx / rcp.approx instead of expf and an IEEE division
#endif
#ifndef FU_PREFETCH
#define FU_PREFETCH 1                        // epilogue inputs of the previous tile requested at the top of the iteration
#endif
#ifndef FU_RELU_CVT
#define FU_RELU_CVT 1                        // relu of H1 / H2 folded into the fp16 conversion (cvt.rn.relu.f16x2.f32), b2 added with packed FADD2
#endif
#ifndef FU_FFMA2
#define FU_FFMA2 0                           // layer 3 with packed fp32 FFMA2 / FADD2 (two hidden units per instruction): 5.40 vs 5.37 ms with 16 gather warps
#endif
#ifndef FU_L3_CONST
#define FU_L3_CONST 0                        // measured: 5.54 ms with the constant-bank table vs 5.41 ms (LDCU.128 into uniform registers is slower than the LDS it saves)
#endif
#if FU_L3_CONST
#include <mutex>
__constant__ float4 c_fu_l3[EGN_HID];
static std::mutex g_fu_l3_mutex;
static cudaEvent_t g_fu_l3_event[64] = {};
#endif

struct FuLayout {
    static constexpr int W1 = 0;
    static constexpr int W2 = W1 + (TC_K1 / 8) * TC_CHUNK;            // 40 960
    static constexpr int BB = W2 + (EGN_HID / 8) * TC_CHUNK;          // + 32 768
    static constexpr int L3 = BB + FU_VCHUNKS * FU_BB_CHUNK;          // + 18 432; layer 3: per hidden unit {b2, W3[0], W3[1], W3[2]}: 128 x float4
                                                                      // (FU_L3_MMA: W3 as an fp16 K-major operand, 16 chunks x 128 B)
    static constexpr int B2 = L3 + EGN_HID * 16;                      // FU_L3_MMA: b2, 128 floats
    static constexpr int IMAGE = B2 + (FU_L3_MMA ? EGN_HID * 4 : 0);  // W1 .. = the operand image, one bulk copy (94 208 / 94 720 bytes)
    static constexpr int A = IMAGE;
    static constexpr int V = A + (TC_K1 / 8) * TC_CHUNK;              // + 40 960 ; two buffers
    static constexpr int REC = V + FU_VBUFS * FU_VBYTES;                     // address records: 128 samples x 144 B
    static constexpr int YANG = REC + (FU_GATHER_WARPS == 16 ? TC_TM * (FU_REC_WORDS / 4) : 8 * (16 * (FU_REC_WORDS / 4) + 1)) * 16;   // records (8-warp cut: + 1 swizzle slot per warp); then 4 x 128 bytes
    static constexpr int KNOTS = YANG + 4 * TC_TM * (FU_ALPHA_IN_GATHER ? 4 : 1);   // FU_ALPHA_IN_GATHER: 4 x 128 floats {alpha | hemisphere in the sign bit}
    static constexpr int MBAR = KNOTS + ((EGN_FUSED_MAX_KNOTS + 1) * 4 + 15) / 16 * 16;
    static constexpr int TMEM = MBAR + 10 * 8;
    static constexpr int PART = TMEM + 16;                            // layer-3 partial sums of the upper column half: 128 x float4 (FU_L3_MMA: alpha only)
    static constexpr int RED = PART + (FU_ALPHA_IN_GATHER ? 0 : TC_TM * (FU_L3_MMA ? 4 : 16));   // fused compositing: 4 warp products + 4 x 5 warp sums
    static constexpr int TOTAL = RED + 4 * 8 * 4;
};
static_assert(FuLayout::TOTAL <= 227 * 1024, "fused kernel exceeds the shared memory of one SM");

// exact n / d for the small divisors of this kernel (d <= 16, n < 2^27: n * (magic * d - 2^32) < 2^32), one IMAD.HI instead of
// the ~100-instruction 64-bit division every thread used to run several times per tile
struct FuDiv {
    uint32_t magic, d;
    __device__ __forceinline__ void init(uint32_t div) { d = div; magic = (uint32_t)((0x100000000ull + div - 1) / div); }
    __device__ __forceinline__ uint32_t operator()(uint32_t n) const { return d == 1 ? n : __umulhi(n, magic); }
};

// packed fp32 arithmetic of sm_100 (SASS FFMA2 / FADD2): two independent fp32 operations per instruction, each IEEE-rounded
__device__ __forceinline__ float2 fu_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fu_fadd2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fu_fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

// L2 residency: the factor tables (74 MB) fit the 126 MB L2, but every launch also streams z, sigma feature and alpha (3 x 67 MB at
// cfg2) through it.  FU_L2_HINTS: table taps are loaded with an evict_last policy, the streams with evict_first.
#ifndef FU_L2_HINTS
#define FU_L2_HINTS 0                        // 1: streaming stores, 2: + evict_last table loads -- measured: no effect (5.38 / 5.38 / 5.39 ms)
#endif
__device__ __forceinline__ uint64_t fu_policy_keep() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 fu_ldg_keep(const uint4* ptr, uint64_t pol) {
#if FU_L2_HINTS >= 2
    uint4 v;
    asm("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol));
    return v;
#else
    return __ldg(ptr);
#endif
}
__device__ __forceinline__ float4 fu_ldg_keep(const float4* ptr, uint64_t pol) {
    const uint4 v = fu_ldg_keep(reinterpret_cast<const uint4*>(ptr), pol);
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ void fu_st_stream(float* ptr, float v) {
#if FU_L2_HINTS >= 1
    __stcs(ptr, v);
#else
    *ptr = v;
#endif
}

__device__ __forceinline__ __half2 as_h2(uint32_t w) { return *reinterpret_cast<__half2*>(&w); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// ---------------------------------------------------------------------------------------------------------------
// GATHER group, phase 1: address record of the sample with index-space coordinate `cc` (EgoNeRF.py:291-347 / 349-413 tap
// geometry, F.grid_sample bilinear / zeros / align_corners=True).  Out-of-range taps get weight 0 and a clamped in-range
// texel, so every later load is unconditional.  Record layout per factor pair i (12 words):
//   [t00 t01 t10 t11] [l0 l1 w00 w01] [w10 w11 u0 u1]     t*, l*: global texel indices (EgnLayoutH); w*, u*: fp32 weights
// (Keeping taps of the previous row in registers was measured and dropped: 59 % of all taps of the cfg2 workload equal the
// previous sample's (profiles/r02_locality.md), but holding even the 8 angular taps across passes cost more in lost L1-tag
// coalescing and registers than the skipped loads saved -- profiles/r02_fused.md.)
// ---------------------------------------------------------------------------------------------------------------
struct FuRecord { unsigned idx[3][6]; float w[3][6]; };
// slot (in uint4 units) of local row r of a warp's 16 records: 9 uint4 per record, rows 8..15 shifted by one uint4 so that
// the rows a gather pass reads together (4g + p for the appearance quarters, 8p + h for the density groups) and the rows
// a quarter-warp of phase 1 writes together sit in distinct 16-byte bank groups
__device__ __forceinline__ int fu_rec_slot(int r) { return r * (FU_REC_WORDS / 4) + (r >> 3); }
#define FU_REC_WARP_UINT4 (16 * (FU_REC_WORDS / 4) + 1)

__device__ __forceinline__ void fused_address_record(const EgnKernelCfg& k, const YYCoord& cc, FuRecord& R) {
    unsigned j0[3], j1[3];
    float wa0[3], wa1[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int G = k.lay.G[a];
        const float ix = egn_unnorm(cc.c[a], G);
        const float fl = floorf(ix);
        const float fr = ix - fl;
        const int i0 = (int)fminf(fmaxf(fl, -2.f), (float)G + 1.f);
        wa0[a] = ((i0 >= 0) & (i0 < G)) ? 1.f - fr : 0.f;
        wa1[a] = ((i0 + 1 >= 0) & (i0 + 1 < G)) ? fr : 0.f;
        j0[a] = (unsigned)min(max(i0, 0), G - 1);
        j1[a] = (unsigned)min(max(i0 + 1, 0), G - 1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ax = egn_mx(i), ay = egn_my(i), al = egn_vl(i);
        const unsigned W = (unsigned)k.lay.G[ax];
        const unsigned pb = (unsigned)k.texp[cc.yang][i], lb = (unsigned)k.texl[cc.yang][i];
        const unsigned ra = pb + j0[ay] * W, rb = pb + j1[ay] * W;
        R.idx[i][0] = ra + j0[ax]; R.idx[i][1] = ra + j1[ax]; R.idx[i][2] = rb + j0[ax]; R.idx[i][3] = rb + j1[ax];
        R.idx[i][4] = lb + j0[al]; R.idx[i][5] = lb + j1[al];
        R.w[i][0] = wa0[ax] * wa0[ay]; R.w[i][1] = wa1[ax] * wa0[ay]; R.w[i][2] = wa0[ax] * wa1[ay]; R.w[i][3] = wa1[ax] * wa1[ay];
        R.w[i][4] = wa0[al]; R.w[i][5] = wa1[al];
    }
}
__device__ __forceinline__ void fused_store_record(const FuRecord& R, uint4* __restrict__ rec) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        rec[3 * i] = make_uint4(R.idx[i][0], R.idx[i][1], R.idx[i][2], R.idx[i][3]);
        rec[3 * i + 1] = make_uint4(R.idx[i][4], R.idx[i][5], __float_as_uint(R.w[i][0]), __float_as_uint(R.w[i][1]));
        rec[3 * i + 2] = make_uint4(__float_as_uint(R.w[i][2]), __float_as_uint(R.w[i][3]), __float_as_uint(R.w[i][4]),
                                    __float_as_uint(R.w[i][5]));
    }
}

// ---- phases 2 and 3 are software pipelines over (pass, factor pair) units: the six taps of a unit are requested one whole
// pass ahead -- right after the previous pass has consumed the registers they land in -- so two to three units (12-18 loads
// per lane) are in flight while a unit is being computed.  The shared memory of this kernel leaves no L1: every tap is an L2
// round trip, and the gather warps are bound by how many of them they keep in flight.

// density: 4 lanes x 4 fp32 channels per sample; lane group h = lane >> 2 takes row 8p + h in pass p (8 consecutive rows per
// load instruction: identical texels of neighbouring samples are fetched once by the L1 tag stage)
__device__ __forceinline__ void fused_density_issue(const float4* __restrict__ dens, const uint4* __restrict__ rec, int i, unsigned sub,
                                                    float4 (&t)[6], uint64_t pol) {
    const uint4 a = rec[3 * i], b = rec[3 * i + 1];
    t[0] = fu_ldg_keep(dens + (a.x * 4u + sub), pol); t[1] = fu_ldg_keep(dens + (a.y * 4u + sub), pol);
    t[2] = fu_ldg_keep(dens + (a.z * 4u + sub), pol); t[3] = fu_ldg_keep(dens + (a.w * 4u + sub), pol);
    t[4] = fu_ldg_keep(dens + (b.x * 4u + sub), pol); t[5] = fu_ldg_keep(dens + (b.y * 4u + sub), pol);
}
__device__ __forceinline__ float fused_density_unit(const uint4* __restrict__ rec, int i, const float4 (&t)[6]) {
    const uint4 b = rec[3 * i + 1], c = rec[3 * i + 2];
    const float w0 = __uint_as_float(b.z), w1 = __uint_as_float(b.w);
    const float w2 = __uint_as_float(c.x), w3 = __uint_as_float(c.y);
    const float u0 = __uint_as_float(c.z), u1 = __uint_as_float(c.w);
    float4 P = make_float4(w0 * t[0].x, w0 * t[0].y, w0 * t[0].z, w0 * t[0].w);
    P = f4fma(w1, t[1], P); P = f4fma(w2, t[2], P); P = f4fma(w3, t[3], P);
    float4 Lv = make_float4(u0 * t[4].x, u0 * t[4].y, u0 * t[4].z, u0 * t[4].w);
    Lv = f4fma(u1, t[5], Lv);
    float s = fmaf(P.w, Lv.w, fmaf(P.z, Lv.z, fmaf(P.y, Lv.y, P.x * Lv.x)));
    s += __shfl_xor_sync(FULL, s, 1);
    s += __shfl_xor_sync(FULL, s, 2);
    return fmaxf(s, 0.f);                                     // relu per factor pair (EgoNeRF.py:346)
}

// appearance: lanes 0..5 of each quarter-warp x 8 fp16 channels (the 128-byte texel line holds 96 bytes of channels: lanes
// 6, 7 shadow lane 5 -- same addresses, no extra wavefront, no divergent control flow -- and only skip the store); the four
// quarters of pass p take the four consecutive rows 4p + g
__device__ __forceinline__ void fused_app_issue(const uint4* __restrict__ app, const uint4* __restrict__ rec, int i, unsigned q,
                                                uint4 (&t)[6], uint64_t pol) {
    const uint4 a = rec[3 * i], b = rec[3 * i + 1];
    t[0] = fu_ldg_keep(app + (a.x * 8u + q), pol); t[1] = fu_ldg_keep(app + (a.y * 8u + q), pol);
    t[2] = fu_ldg_keep(app + (a.z * 8u + q), pol); t[3] = fu_ldg_keep(app + (a.w * 8u + q), pol);
    t[4] = fu_ldg_keep(app + (b.x * 8u + q), pol); t[5] = fu_ldg_keep(app + (b.y * 8u + q), pol);
}
__device__ __forceinline__ uint4 fused_app_unit(const uint4* __restrict__ rec, int i, const uint4 (&t)[6]) {
    const uint4 b = rec[3 * i + 1], c = rec[3 * i + 2];
    const float f0 = __uint_as_float(b.z), f1 = __uint_as_float(b.w);
    const float f2 = __uint_as_float(c.x), f3 = __uint_as_float(c.y);
    const float g0 = __uint_as_float(c.z), g1 = __uint_as_float(c.w);
#if FU_PAIRED_W
    // three conversions make the six weights (two per register); the multiplies read one half of a register broadcast
    // (SASS operand modifiers .H0_H0 / .H1_H1)
    const __half2 w01 = as_h2(pack_h2(f0, f1)), w23 = as_h2(pack_h2(f2, f3)), u01 = as_h2(pack_h2(g0, g1));
    const __half2 w0 = __low2half2(w01), w1 = __high2half2(w01), w2 = __low2half2(w23), w3 = __high2half2(w23);
    const __half2 u0 = __low2half2(u01), u1 = __high2half2(u01);
#else
    const __half2 w0 = as_h2(pack_h2(f0, f0)), w1 = as_h2(pack_h2(f1, f1)), w2 = as_h2(pack_h2(f2, f2)), w3 = as_h2(pack_h2(f3, f3));
    const __half2 u0 = as_h2(pack_h2(g0, g0)), u1 = as_h2(pack_h2(g1, g1));
#endif
    uint4 o;
#define FU_PL(m) { __half2 P = __hmul2(w0, as_h2(t[0].m)); P = __hfma2(w1, as_h2(t[1].m), P); P = __hfma2(w2, as_h2(t[2].m), P); \
                   P = __hfma2(w3, as_h2(t[3].m), P); const __half2 Lv = __hfma2(u1, as_h2(t[5].m), __hmul2(u0, as_h2(t[4].m))); \
                   o.m = as_u32(__hmul2(P, Lv)); }
    FU_PL(x) FU_PL(y) FU_PL(z) FU_PL(w)
#undef FU_PL
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// Operand image of the MLP: W1 (bias folded in, PE column order), W2, [B_yin | B_yang] as fp16 canonical K-major tcgen05 tiles
// and the layer-3 table, laid out exactly as the fused kernel keeps them in shared memory (FuLayout W1 .. L3).  Built once
// per launch by this small kernel; every CTA then fetches it with ONE bulk copy (cp.async.bulk -> mbarrier, SASS UBLKCP)
// instead of 512 threads converting ~46 000 scattered fp32 weights per CTA.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
egn_fused_image_kernel(int AD, const float* __restrict__ basis0, const float* __restrict__ basis1, const float* __restrict__ w1,
                       const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                       const float* __restrict__ w3, unsigned char* __restrict__ img) {
    using L = FuLayout;
    const int in_dim = 5 * AD + 15;
    const int nthreads = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i < EGN_HID * TC_K1; i += nthreads) {
        const int n = i / TC_K1, kk = i % TC_K1;
        const int src = tc_input_index(kk, AD);
        store_elem_h(img + L::W1, n, kk, src >= 0 ? w1[n * in_dim + src] : (src == -2 ? b1[n] : 0.f));
    }
    for (int i = t0; i < EGN_HID * EGN_HID; i += nthreads) store_elem_h(img + L::W2, i / EGN_HID, i % EGN_HID, w2[i]);
    for (int i = t0; i < 64 * FU_VK; i += nthreads) {               // rows 0..31: basis_mat_yin, 32..63: basis_mat_yang
        const int n = i / FU_VK, kk = i % FU_VK, o = n & 31;
        const float* B = (n >> 5) ? basis1 : basis0;
        store_elem_h(img + L::BB, n, kk, o < AD ? B[o * FU_VK + kk] : 0.f, FU_BB_CHUNK);
    }
#if FU_L3_MMA
    for (int i = t0; i < 8 * EGN_HID; i += nthreads) {                // rows 0..2 = W3, rows 3..7 zero
        const int n = i / EGN_HID, kk = i % EGN_HID;
        store_elem_h(img + L::L3, n, kk, n < 3 ? w3[n * EGN_HID + kk] : 0.f, FU_W3_CHUNK);
    }
    for (int i = t0; i < EGN_HID; i += nthreads) reinterpret_cast<float*>(img + L::B2)[i] = b2[i];
#elif FU_FFMA2 && !FU_L3_CONST
    for (int i = t0; i < EGN_HID / 2; i += nthreads) {               // per PAIR of hidden units (2i, 2i+1): {b2, b2, W3[0], W3[0]} {W3[1], W3[1], W3[2], W3[2]}
        reinterpret_cast<float4*>(img + L::L3)[2 * i] = make_float4(b2[2 * i], b2[2 * i + 1], w3[2 * i], w3[2 * i + 1]);
        reinterpret_cast<float4*>(img + L::L3)[2 * i + 1] = make_float4(w3[EGN_HID + 2 * i], w3[EGN_HID + 2 * i + 1], w3[2 * EGN_HID + 2 * i],
                                                                        w3[2 * EGN_HID + 2 * i + 1]);
    }
#else
    for (int i = t0; i < EGN_HID; i += nthreads)
        reinterpret_cast<float4*>(img + L::L3)[i] = make_float4(b2[i], w3[i], w3[EGN_HID + i], w3[2 * EGN_HID + i]);
#endif
}

// COMP = true (forward-only calls with S a multiple of 128): compositing runs inside the kernel too (EgoNeRF.py:579-598,
// tensorBase.py:22-27).  A CTA then walks whole rays (all tiles of a ray back to back); the density lanes turn the sigma
// feature into alpha (the only per-sample output, 4 B) and write it to the caller's alpha tensor; the layer-3 epilogue reads
// it back, runs the transmittance product scan over the tile's 128 rows (warp shuffles + 4 warp totals through shared
// memory), carries T across the tiles of the ray and accumulates sum(w c), sum(w), sum(w z): sample colours, sigma
// features and weights never leave the SM, and egn_composite_kernel is not launched.
template <bool COMP>
__global__ void __launch_bounds__(FU_THREADS, 1)
egn_fused_fine_kernel(const __grid_constant__ EgnKernelCfg k, const unsigned char* __restrict__ image,
                      const float* __restrict__ b3, const float* __restrict__ rays, long long M,
                      const float* __restrict__ zs, float* __restrict__ fsig, float* __restrict__ feat_out,
                      float* __restrict__ rgbs, const float* __restrict__ emission, EgnOutputs out) {
    using L = FuLayout;
    extern __shared__ __align__(128) unsigned char smem[];
    // warp index through a shuffle from lane 0: the compiler then KNOWS it is warp-uniform, keeps the role branches uniform and
    // the load descriptors in uniform registers (without it every LDG of the kernel was preceded by two R2UR moves)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int AD = k.app_dim;
    unsigned char* w1s = smem + L::W1;
    unsigned char* w2s = smem + L::W2;
    unsigned char* bbs = smem + L::BB;
    unsigned char* as = smem + L::A;
    unsigned char* vs = smem + L::V;
#if FU_ALPHA_IN_GATHER
    float* s_ay = reinterpret_cast<float*>(smem + L::YANG);      // ring of 4 tiles x 128 {alpha, sign bit = hemisphere}
#else
    unsigned char* s_yang = smem + L::YANG;
#endif
    float* s_knots = reinterpret_cast<float*>(smem + L::KNOTS);
#if !FU_ALPHA_IN_GATHER
    float* part = reinterpret_cast<float*>(smem + L::PART);
#endif
#if !FU_L3_CONST && !FU_L3_MMA
    float4* l3s = reinterpret_cast<float4*>(smem + L::L3);
#endif
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::TMEM);
    const uint32_t bar = smem_u32(smem + L::MBAR);
    const uint32_t v_full0 = bar, v_empty0 = bar + 16, feat_full = bar + 32, d1_full = bar + 40, d2_full = bar + 48, img_full = bar + 56, d3_full = bar + 64, a_ready = bar + 72;

    // ---- one-time setup ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(v_full0, FU_GATHER_WARPS); mbar_init(v_full0 + 8, FU_GATHER_WARPS);          // one arrive per gather warp
        mbar_init(v_empty0, 1); mbar_init(v_empty0 + 8, 1);        // tcgen05.commit
        mbar_init(feat_full, 1); mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(img_full, 1); mbar_init(d3_full, 1); mbar_init(a_ready, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the operand image (W1, W2, basis, layer-3 table: FuLayout W1 .. IMAGE) in one bulk copy, completion on img_full
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(img_full), "r"((uint32_t)L::IMAGE) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(smem + L::W1)), "l"(image), "r"((uint32_t)L::IMAGE), "r"(img_full) : "memory");
    }
    for (int i = tid; i <= k.knots_last; i += FU_THREADS) s_knots[i] = k.r_knots[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    // sample indices are 32-bit inside the kernel (the launcher refuses M >= 2^32 - 256); S is a multiple of 32
    const uint32_t M32 = (uint32_t)M, S32 = (uint32_t)k.S;
    const uint32_t tiles = (M32 + TC_TM - 1) / TC_TM;
    bool ok = mbar_wait(img_full, 0);                           // operand image landed (async-proxy write -> visible after the wait)
    // CTA-local iteration -> tile.  Plain: tiles strided over the CTAs.  COMP: rays strided over the CTAs, the tpr tiles of a
    // ray consecutive (S = tpr * 128), so the transmittance can be carried from tile to tile.
    const uint32_t tpr = COMP ? S32 / TC_TM : 1;
    const uint32_t n_local = COMP ? (((M32 / S32) - blockIdx.x + gridDim.x - 1) / gridDim.x) * tpr
                                   : (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    FuDiv div_tpr, div_s;
    div_tpr.init(tpr);
    div_s.init(S32 >> 5);
    auto tile_of = [&](uint32_t i) -> uint32_t {
        if (!COMP) return blockIdx.x + i * gridDim.x;
        const uint32_t q = div_tpr(i);
        return (q * gridDim.x + blockIdx.x) * tpr + (i - q * tpr);
    };
    auto ray_of = [&](uint32_t m) -> uint32_t { return div_s(m >> 5); };       // sample index -> ray

#if FU_GATHER_WARPS == 16
    if (warp >= 8 && warp < 8 + FU_GATHER_WARPS) {
        // =========================== GATHER group, 16 warps x 8 rows ===========================
        // Twice the warps over the same rows: the gather chain of a tile is split over 16 instruction streams (4 per scheduler
        // next to 2 MLP warps) instead of 8.  Coordinates: every fourth iteration for four tiles at once (lanes 8g .. 8g+7 take
        // the warp's rows of tile it + g).  Density: one pass of 8 samples; appearance: two passes of 4 samples.
        const int gwarp = warp - 8, row0 = 8 * gwarp;
        uint4* recs = reinterpret_cast<uint4*>(smem + L::REC) + gwarp * (8 * (FU_REC_WORDS / 4));
        const uint4* app = reinterpret_cast<const uint4*>(k.tables_h);
        const float4* dens = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(k.tables_h) + k.dens_byte_offset);
        const uint64_t pol = fu_policy_keep();
        YYCoord held;
        held.c[0] = held.c[1] = held.c[2] = -3.f; held.yang = 0;
        for (uint32_t it = 0; it < n_local; ++it) {
            const uint32_t tile = tile_of(it);
            const uint32_t b = FU_VBUFS == 2 ? (it & 1) : 0, u = FU_VBUFS == 2 ? (it >> 1) : it;
            if ((it & 3) == 0) {
                const uint32_t g = lane >> 3;
                const uint32_t m = tile_of(it + g) * TC_TM + row0 + (lane & 7);
                held.c[0] = held.c[1] = held.c[2] = -3.f;            // out of range -> every tap gets weight zero
                held.yang = 0;
                if (m < M32 && it + g < n_local) {
                    const float z = zs[m];
                    const float* ry = rays + (size_t)ray_of(m) * 6;
                    held = egn_cart_to_yinyang(ry[0] + ry[3] * z, ry[1] + ry[4] * z, ry[2] + ry[5] * z, k, s_knots);
                }
            }
            YYCoord cc;
            {
                const int src = (lane & 7) + 8 * (it & 3);
                cc.c[0] = __shfl_sync(FULL, held.c[0], src);
                cc.c[1] = __shfl_sync(FULL, held.c[1], src);
                cc.c[2] = __shfl_sync(FULL, held.c[2], src);
                cc.yang = __shfl_sync(FULL, held.yang, src);
            }
            if (lane < 8) {
                FuRecord R;
                fused_address_record(k, cc, R);
                fused_store_record(R, recs + lane * (FU_REC_WORDS / 4));
                if (!COMP && k.coords != nullptr) {                  // training forward: the backward reads the coordinates back
                    const uint32_t mrow = tile * TC_TM + row0 + lane;
                    if (mrow < M32) reinterpret_cast<float4*>(k.coords)[mrow] = make_float4(cc.c[0], cc.c[1], cc.c[2], __int_as_float(cc.yang));
                }
#if !FU_ALPHA_IN_GATHER
                s_yang[(it & 3) * TC_TM + row0 + lane] = (unsigned char)cc.yang;
#endif
            }
            __syncwarp();
            // ---- density (fp32): 8 samples x 4 lanes, three factor pairs ----
#if FU_ALPHA_IN_GATHER
            float ay = 0.f;                                          // {alpha, sign bit = hemisphere} of row lane >> 2
#endif
            {
                const unsigned sub = lane & 3;
                const uint4* rec = recs + (lane >> 2) * (FU_REC_WORDS / 4);
                float4 t[FU_DEPTH][6];
#pragma unroll
                for (int i = 0; i < FU_DEPTH; ++i) fused_density_issue(dens, rec, i, sub, t[i], pol);
#if FU_ALPHA_IN_GATHER
                // the row's depth step, requested behind the taps (tensorBase.py:22-27 with the distances of EgoNeRF.py:541-542,553:
                // z[j+1] - z[j], the last one of a ray repeated)
                const uint32_t m = tile * TC_TM + row0 + (lane >> 2);
                const int yang_row = __shfl_sync(FULL, cc.yang, lane >> 2);
                float dist = 0.f;
                uint32_t ray_a = 0, j_a = 0;
                if (COMP && m < M32) {
                    ray_a = ray_of(m);
                    j_a = m - ray_a * S32;
                    const float z0 = zs[m];
                    dist = (j_a + 1 < S32 ? zs[m + 1] - z0 : z0 - zs[m - 1]) * k.distance_scale;
                }
#endif
                float f = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    f += fused_density_unit(rec, i, t[i % FU_DEPTH]);
                    if (i + FU_DEPTH < 3) fused_density_issue(dens, rec, i + FU_DEPTH, sub, t[i % FU_DEPTH], pol);
                }
#if FU_ALPHA_IN_GATHER
                float a = 0.f;
                if (sub == 0 && m < M32) {
                    if constexpr (COMP) {
                        a = 1.f - expf(-egn_density_act(f, k.density_shift, k.fea2dense) * dist);
                        fu_st_stream(out.alpha + (size_t)ray_a * (S32 + (k.env_h > 0 ? 1 : 0)) + j_a, a);
                    } else {
                        fu_st_stream(fsig + m, f);
                    }
                }
                ay = __uint_as_float(__float_as_uint(a) | ((unsigned)yang_row << 31));
#else
                const uint32_t m = tile * TC_TM + row0 + (lane >> 2);
                if (sub == 0 && m < M32) fu_st_stream(fsig + m, f);
#endif
            }
            // ---- appearance (fp16) into the V operand: 2 passes x 4 samples x 3 factor pairs = 6 units ----
            {
                const unsigned q = min(lane & 7, 5);
                const bool owner = (lane & 7) < 6;
                const uint4* recA = recs + (lane >> 3) * (FU_REC_WORDS / 4);
                const uint4* recB = recA + 4 * (FU_REC_WORDS / 4);
                unsigned char* vrow = vs + b * FU_VBYTES + (row0 + (lane >> 3)) * 16 + q * FU_VCHUNK;
                uint4 t[FU_DEPTH][6];
#pragma unroll
                for (int un = 0; un < FU_DEPTH; ++un) fused_app_issue(app, un / 3 ? recB : recA, un % 3, q, t[un], pol);
                ok &= mbar_wait(v_empty0 + 8 * b, (u & 1) ^ 1);      // layer-0 MMAs of the tile that used this buffer are done
#if FU_ALPHA_IN_GATHER
                // ring slot it & 3: layer 0 of tile it - 2 is complete, so every MLP warp has finished the epilogue of tile it - 4
                if ((lane & 3) == 0) s_ay[(it & 3) * TC_TM + row0 + (lane >> 2)] = ay;
#endif
#pragma unroll
                for (int un = 0; un < 6; ++un) {
                    const uint4 o = fused_app_unit(un / 3 ? recB : recA, un % 3, t[un % FU_DEPTH]);
                    if (owner) *reinterpret_cast<uint4*>(vrow + (un / 3) * (4 * 16) + (un % 3) * (EGN_CA / 8) * FU_VCHUNK) = o;
                    if (un + FU_DEPTH < 6) fused_app_issue(app, (un + FU_DEPTH) / 3 ? recB : recA, (un + FU_DEPTH) % 3, q, t[un % FU_DEPTH], pol);
                }
            }
            fence_async_smem();                                      // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(v_full0 + 8 * b);
        }
#else
    if (warp >= 8) {
        // =========================== GATHER group ===========================
        const int gwarp = warp - 8, row0 = 16 * gwarp;
        uint4* recs = reinterpret_cast<uint4*>(smem + L::REC) + gwarp * FU_REC_WARP_UINT4;
        const uint4* app = reinterpret_cast<const uint4*>(k.tables_h);
        const float4* dens = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(k.tables_h) + k.dens_byte_offset);
        const uint64_t pol = fu_policy_keep();
        YYCoord held;                                                // coordinates computed one tile ahead (lanes 16..31)
        held.c[0] = held.c[1] = held.c[2] = -3.f; held.yang = 0;
        for (uint32_t it = 0; it < n_local; ++it) {
            const uint32_t tile = tile_of(it);
            const uint32_t b = FU_VBUFS == 2 ? (it & 1) : 0, u = FU_VBUFS == 2 ? (it >> 1) : it;
            // ---- phase 1a: coordinates.  Even iterations: lanes 0..15 take this tile's rows, lanes 16..31 the same rows of
            // the CTA's next tile (kept in `held`); odd iterations just fetch them ----
            YYCoord cc;
            if (b == 0) {
                const uint32_t m = (lane < 16 ? tile : tile_of(it + 1)) * TC_TM + row0 + (lane & 15);
                cc.c[0] = cc.c[1] = cc.c[2] = -3.f;                  // out of range -> every tap gets weight zero
                cc.yang = 0;
                if (m < M32 && (lane < 16 || it + 1 < n_local)) {
                    const float z = zs[m];
                    const float* ry = rays + (size_t)ray_of(m) * 6;
                    cc = egn_cart_to_yinyang(ry[0] + ry[3] * z, ry[1] + ry[4] * z, ry[2] + ry[5] * z, k, s_knots);
                }
                held = cc;
            } else {
                cc.c[0] = __shfl_sync(FULL, held.c[0], (lane & 15) + 16);
                cc.c[1] = __shfl_sync(FULL, held.c[1], (lane & 15) + 16);
                cc.c[2] = __shfl_sync(FULL, held.c[2], (lane & 15) + 16);
                cc.yang = __shfl_sync(FULL, held.yang, (lane & 15) + 16);
            }
            // ---- phase 1b: address records of the warp's 16 rows (lane = row) ----
            if (lane < 16) {
                FuRecord R;
                fused_address_record(k, cc, R);
                fused_store_record(R, recs + fu_rec_slot(lane));
                if (!COMP && k.coords != nullptr) {                  // training forward: the backward reads the coordinates back
                    const uint32_t mrow = tile * TC_TM + row0 + lane;
                    if (mrow < M32) reinterpret_cast<float4*>(k.coords)[mrow] = make_float4(cc.c[0], cc.c[1], cc.c[2], __int_as_float(cc.yang));
                }
                s_yang[(it & 3) * TC_TM + row0 + lane] = (unsigned char)cc.yang;
            }
            __syncwarp();
            // ---- phase 2: density (fp32), 2 passes x 8 samples, pipelined per factor pair ----
            {
                const unsigned sub = lane & 3;
                const uint4* rec = recs + fu_rec_slot(lane >> 2);
                float4 t[3][6];
#pragma unroll
                for (int i = 0; i < 3; ++i) fused_density_issue(dens, rec, i, sub, t[i], pol);
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const uint4* nxt = recs + fu_rec_slot(8 * (p + 1) + (lane >> 2));
                    float f = 0.f;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        f += fused_density_unit(rec, i, t[i]);
                        if (p == 0) fused_density_issue(dens, nxt, i, sub, t[i], pol);
                    }
                    const uint32_t m = tile * TC_TM + row0 + 8 * p + (lane >> 2);
                    if (sub == 0 && m < M32) fu_st_stream(fsig + m, f);
                    rec = nxt;
                }
            }
            // ---- phase 3: appearance (fp16) into the V operand, 4 passes x 4 samples, pipelined per factor pair.  The first
            // pass's taps are requested BEFORE waiting for the V buffer ----
            {
                const unsigned q = min(lane & 7, 5);
                const bool owner = (lane & 7) < 6;
                const uint4* rec = recs + fu_rec_slot(lane >> 3);
                unsigned char* vrow = vs + b * FU_VBYTES + (row0 + (lane >> 3)) * 16 + q * FU_VCHUNK;
                uint4 t[3][6];
#pragma unroll
                for (int i = 0; i < 3; ++i) fused_app_issue(app, rec, i, q, t[i], pol);
                ok &= mbar_wait(v_empty0 + 8 * b, (u & 1) ^ 1);      // layer-0 MMAs of the tile that used this buffer are done
FU_UNROLL(FU_APP_UNROLL)
                for (int p = 0; p < 4; ++p) {
                    const uint4* nxt = recs + fu_rec_slot(min(4 * (p + 1), 12) + (lane >> 3));
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const uint4 o = fused_app_unit(rec, i, t[i]);
                        if (owner) *reinterpret_cast<uint4*>(vrow + i * (EGN_CA / 8) * FU_VCHUNK) = o;   // K index i*48 + q*8 .. +7
                        if (p < 3) fused_app_issue(app, nxt, i, q, t[i], pol);
                    }
                    rec = nxt;
                    vrow += 4 * 16;
                }
            }
            fence_async_smem();                                      // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(v_full0 + 8 * b);
        }
#endif
#if FU_MMA_WARP
    } else if (warp == 8 + FU_GATHER_WARPS) {
        // =========================== MMA warp ===========================
        // One elected thread issues every tcgen05.mma of the CTA in the order the tensor pipe should run them: layer 1 and 2 of
        // tile t, layer 0 of tile t+1, layer 3 of tile t.  The MLP warps announce each finished operand (X, H1, H2 -- the same
        // buffer, three phases per tile) on the `a_ready` mbarrier and run on; they meet the results on d1 / d2 / d3 / feat_full.
        // (With the issue inside MLP warp 0 that warp executed ~450 instructions per tile more than the other seven, and every
        // group barrier waited for it: 27 % of the group's stall samples.)
        if (lane == 0) {
            const uint32_t a_s = smem_u32(as), w1_s = smem_u32(w1s), w2_s = smem_u32(w2s), bb_s = smem_u32(bbs), v_s = smem_u32(vs);
            const uint32_t w3_s = smem_u32(smem + L::L3);
            auto issue_layer0 = [&](uint32_t i) {                   // thread 0: feat2 = V . [B_yin | B_yang]^T for CTA-local tile i
                const uint32_t bb = FU_VBUFS == 2 ? (i & 1) : 0;
                ok &= mbar_wait(v_full0 + 8 * bb, (FU_VBUFS == 2 ? (i >> 1) : i) & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < FU_VK / 16; ++ks)
                    tc_mma(tmem + 256, tc_desc(v_s + bb * FU_VBYTES + ks * 2 * FU_VCHUNK, FU_VCHUNK),
                           tc_desc(bb_s + ks * 2 * FU_BB_CHUNK, FU_BB_CHUNK), FU_IDESC_128x64, ks > 0);
                tc_commit(v_empty0 + 8 * bb);
                tc_commit(feat_full);
            };
            uint32_t ph = 0;
            if (n_local > 0) issue_layer0(0);
            for (uint32_t it = 0; it < n_local; ++it) {
                ok &= mbar_wait(a_ready, ph++ & 1);              // X of tile it
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < TC_K1 / 16; ++ks)
                    tc_mma(tmem, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc(w1_s + ks * 2 * TC_CHUNK), FU_IDESC_128x128, ks > 0);
                tc_commit(d1_full);
                ok &= mbar_wait(a_ready, ph++ & 1);              // H1
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < EGN_HID / 16; ++ks)
                    tc_mma(tmem + 128, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc(w2_s + ks * 2 * TC_CHUNK), FU_IDESC_128x128, ks > 0);
                tc_commit(d2_full);
                if (it + 1 < n_local) issue_layer0(it + 1);
                ok &= mbar_wait(a_ready, ph++ & 1);              // H2
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < EGN_HID / 16; ++ks)
                    tc_mma(tmem + 384, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc_sbo(w3_s + ks * 2 * FU_W3_CHUNK, FU_W3_CHUNK, 0), FU_IDESC_128x16, ks > 0);
                tc_commit(d3_full);
            }
        }
        __syncwarp();
#endif
    } else {
        // =========================== MLP group ===========================
        const int row = tid & 127, half = tid >> 7;
        const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t a_s = smem_u32(as), w1_s = smem_u32(w1s), w2_s = smem_u32(w2s), bb_s = smem_u32(bbs), v_s = smem_u32(vs);
        const float bias3[3] = {__ldg(b3), __ldg(b3 + 1), __ldg(b3 + 2)};
        // Software pipeline over tiles.  The tensor pipe works on layer 0 of tile t+1 while this group waits for layer 2 of
        // tile t, and on layer 1 of tile t while the group runs layer 3 of tile t-1 out of TMEM: only the layer-2 wait is
        // exposed.  TMEM: D1 = columns 0..127, D2 = 128..255, feat2 = 256..319.
#if !FU_MMA_WARP
        auto issue_layer0 = [&](uint32_t i) {                   // thread 0: feat2 = V . [B_yin | B_yang]^T for CTA-local tile i
            const uint32_t bb = FU_VBUFS == 2 ? (i & 1) : 0;
            ok &= mbar_wait(v_full0 + 8 * bb, (FU_VBUFS == 2 ? (i >> 1) : i) & 1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < FU_VK / 16; ++ks)
                tc_mma(tmem + 256, tc_desc(v_s + bb * FU_VBYTES + ks * 2 * FU_VCHUNK, FU_VCHUNK),
                       tc_desc(bb_s + ks * 2 * FU_BB_CHUNK, FU_BB_CHUNK), FU_IDESC_128x64, ks > 0);
            tc_commit(v_empty0 + 8 * bb);
            tc_commit(feat_full);
        };
#endif
        // layer 3 + sigmoid of the tile whose D2 sits in TMEM (its layer-2 completion has been waited for); COMP: then the
        // compositing of the tile's 128 samples (they belong to one ray), carried over the tiles of the ray
        float carryT = 1.f, sum_w = 0.f, sum_r = 0.f, sum_g = 0.f, sum_b = 0.f, sum_z = 0.f;     // sums live in thread 0
        float* red = reinterpret_cast<float*>(smem + L::RED);
        const int acols = k.S + (k.env_h > 0 ? 1 : 0);
        // COMP: the upper-half threads turn the row's sigma feature -- written by the gather group, read back through L2 -- into
        // alpha (tensorBase.py:22-27 with the distances of EgoNeRF.py:541-542,553: z[j+1] - z[j], the last one repeated).  Its
        // inputs {sigma feature, z, next z} are requested a whole operand phase before they are used (FU_PREFETCH).
        auto prefetch = [&](uint32_t gm3) -> float3 {
            float3 pre = make_float3(0.f, 0.f, 0.f);
            if constexpr (COMP) {
                if (gm3 < M32) {
#if FU_ALPHA_IN_GATHER
                    if (half == 0) pre.y = zs[gm3];
#else
                    pre.y = zs[gm3];
                    if (half == 1) {
                        const uint32_t ray = ray_of(gm3);
                        pre.x = __ldcg(fsig + gm3);
                        pre.z = (gm3 - ray * S32 + 1 < S32) ? zs[gm3 + 1] : zs[gm3 - 1];
                    }
#endif
                }
            }
            return pre;
        };
        auto layer3 = [&](uint32_t gm3, uint32_t it3, float3 pre) {
            const float fs = pre.x, zrow = pre.y, znext = pre.z;
            uint32_t ray3 = 0;
            int j3 = 0;
            if constexpr (COMP) {
                ray3 = ray_of(gm3);
                j3 = (int)(gm3 - ray3 * S32);
            }
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#if FU_L3_MMA
            if (half == 0) {                                         // D3 = H2 . W3^T of this tile sits in TMEM columns 384..386
                uint32_t r4[4];
                tmem_ld4(tmem_lane + 384, r4);
                p0 = __uint_as_float(r4[0]); p1 = __uint_as_float(r4[1]); p2 = __uint_as_float(r4[2]);
            }
#else
#if FU_L3_CONST
            // the column half is warp-uniform: two copies of the loop with compile-time constant-bank offsets
            auto dot64 = [&](auto HALF) {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    constexpr int col = 64 * decltype(HALF)::value + 0;
                    uint32_t r[32];
                    tmem_ld32(tmem_lane + 128 + col + 32 * cc, r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float4 w = c_fu_l3[col + 32 * cc + j];
                        const float h = fmaxf(__uint_as_float(r[j]) + w.x, 0.f);
                        p0 = fmaf(h, w.y, p0); p1 = fmaf(h, w.z, p1); p2 = fmaf(h, w.w, p2);
                    }
                }
            };
            if (half == 0) dot64(std::integral_constant<int, 0>{}); else dot64(std::integral_constant<int, 1>{});
#elif FU_FFMA2
            {
                float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0;     // even / odd hidden units accumulate separately
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = 64 * half + 32 * cc;
                    uint32_t r[32];
                    tmem_ld32(tmem_lane + 128 + col, r);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 wa = l3s[col + 2 * j], wb = l3s[col + 2 * j + 1];          // two broadcast LDS.128 per unit pair
                        float2 h = fu_fadd2(make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), make_float2(wa.x, wa.y));
                        h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f);
                        q0 = fu_ffma2(h, make_float2(wa.z, wa.w), q0);
                        q1 = fu_ffma2(h, make_float2(wb.x, wb.y), q1);
                        q2 = fu_ffma2(h, make_float2(wb.z, wb.w), q2);
                    }
                }
                p0 = q0.x + q0.y; p1 = q1.x + q1.y; p2 = q2.x + q2.y;
            }
#else
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col = 64 * half + 32 * cc;
                uint32_t r[32];
                tmem_ld32(tmem_lane + 128 + col, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float4 w = l3s[col + j];                     // broadcast LDS.128: {b2, W3[0], W3[1], W3[2]} of unit col + j
                    const float h = fmaxf(__uint_as_float(r[j]) + w.x, 0.f);
                    p0 = fmaf(h, w.y, p0); p1 = fmaf(h, w.z, p1); p2 = fmaf(h, w.w, p2);
                }
            }
#endif
#endif
            tc_fence_before();
#if FU_ALPHA_IN_GATHER
            (void)fs; (void)znext;
            if (half == 0 && gm3 < M32) {
                const float4 q = make_float4(0.f, 0.f, 0.f, fabsf(s_ay[(it3 & 3) * TC_TM + row]));
#else
            if (half == 1) {
                float a = 0.f;
                if constexpr (COMP) {
                    const float dist = ((j3 + 1 < k.S) ? (znext - zrow) : (zrow - znext)) * k.distance_scale;
                    a = 1.f - expf(-egn_density_act(fs, k.density_shift, k.fea2dense) * dist);
                    fu_st_stream(out.alpha + (size_t)ray3 * acols + j3, a);
                }
#if FU_L3_MMA
                part[row] = a;
#else
                *reinterpret_cast<float4*>(part + row * 4) = make_float4(p0, p1, p2, a);
#endif
            }
            named_bar_sync(1, FU_GROUP);
            if (half == 0 && gm3 < M32) {
#if FU_L3_MMA
                const float4 q = make_float4(0.f, 0.f, 0.f, part[row]);
#else
                const float4 q = *reinterpret_cast<const float4*>(part + row * 4);
#endif
#endif
                const float alpha = q.w;
#if FU_FAST_SIGMOID
                const float c0 = __fdividef(1.f, 1.f + __expf(-(p0 + q.x + bias3[0]))), c1 = __fdividef(1.f, 1.f + __expf(-(p1 + q.y + bias3[1]))),
                            c2 = __fdividef(1.f, 1.f + __expf(-(p2 + q.z + bias3[2])));
#else
                const float c0 = egn_sigmoid(p0 + q.x + bias3[0]), c1 = egn_sigmoid(p1 + q.y + bias3[1]), c2 = egn_sigmoid(p2 + q.z + bias3[2]);
#endif
                if constexpr (!COMP) {
                    float* o3 = rgbs + (size_t)gm3 * 3;
                    o3[0] = c0; o3[1] = c1; o3[2] = c2;
                } else {
                    // transmittance: T_j = prod_{i<j} (1 - alpha_i + 1e-10) (tensorBase.py:24-26); rows = consecutive samples
                    const int lane_ = tid & 31, w4 = tid >> 5;         // w4 in 0..3
                    float incl = 1.f - alpha + 1e-10f;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const float o = __shfl_up_sync(FULL, incl, d);
                        if (lane_ >= d) incl *= o;
                    }
                    float excl = __shfl_up_sync(FULL, incl, 1);
                    if (lane_ == 0) excl = 1.f;
                    if (lane_ == 31) red[w4] = incl;
                    named_bar_sync(2, TC_TM);
                    const float q0 = red[0], q1 = red[1], q2 = red[2], q3 = red[3];
                    const float before = w4 == 0 ? 1.f : (w4 == 1 ? q0 : (w4 == 2 ? q0 * q1 : (q0 * q1) * q2));
                    const float wgt = alpha * ((carryT * before) * excl);
                    float v[5] = {wgt, wgt * c0, wgt * c1, wgt * c2, wgt * zrow};
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1)
#pragma unroll
                        for (int e = 0; e < 5; ++e) v[e] += __shfl_xor_sync(FULL, v[e], d);
                    if (lane_ == 0) {
#pragma unroll
                        for (int e = 0; e < 5; ++e) red[4 + w4 * 5 + e] = v[e];
                    }
                    named_bar_sync(2, TC_TM);
                    carryT *= ((q0 * q1) * q2) * q3;
                    const size_t ray = ray3;
                    const bool last = j3 + TC_TM - row >= k.S;                     // this tile is the ray's last one
                    if (tid == 0) {
                        sum_w += (red[4] + red[9]) + (red[14] + red[19]);
                        sum_r += (red[5] + red[10]) + (red[15] + red[20]);
                        sum_g += (red[6] + red[11]) + (red[16] + red[21]);
                        sum_b += (red[7] + red[12]) + (red[17] + red[22]);
                        sum_z += (red[8] + red[13]) + (red[18] + red[23]);
                        if (last) {
                            const float* ry = rays + ray * 6;
                            float cr = sum_r, cg = sum_g, cb = sum_b;
                            if (k.env_h > 0) {                         // EgoNeRF.py:586-590
                                float e[3];
                                egn_env_radiance(emission, k.env_h, ry[3], ry[4], ry[5], e);
                                const float b0 = carryT * e[0], b1 = carryT * e[1], b2v = carryT * e[2];
                                out.env[ray * 3] = e[0]; out.env[ray * 3 + 1] = e[1]; out.env[ray * 3 + 2] = e[2];
                                out.bg[ray * 3] = b0; out.bg[ray * 3 + 1] = b1; out.bg[ray * 3 + 2] = b2v;
                                cr += b0; cg += b1; cb += b2v;
                                out.alpha[ray * acols + k.S] = 1.f;     // EgoNeRF.py:587
                            }
                            out.rgb[ray * 3] = fminf(fmaxf(cr, 0.f), 1.f);
                            out.rgb[ray * 3 + 1] = fminf(fmaxf(cg, 0.f), 1.f);
                            out.rgb[ray * 3 + 2] = fminf(fmaxf(cb, 0.f), 1.f);
                            out.depth[ray] = sum_z + (1.f - sum_w) * ry[5];     // EgoNeRF.py:598: rays_chunk[..., -1] is d_z
                            sum_w = sum_r = sum_g = sum_b = sum_z = 0.f;
                        }
                    }
                    if (last) carryT = 1.f;
                }
            }
        };
#if !FU_MMA_WARP
        if (tid == 0 && n_local > 0) issue_layer0(0);
#endif
        uint32_t it = 0;
        uint32_t gm_prev = M32;
        for (; it < n_local; ++it) {
            const uint32_t tile = tile_of(it);
            const uint32_t gm = tile * TC_TM + row;
            const bool live = gm < M32;
            float3 pre = make_float3(0.f, 0.f, 0.f);
            if (FU_PREFETCH && it > 0) pre = prefetch(gm_prev);
            // ---- layer 0 of this tile was issued one iteration ago ----
            ok &= mbar_wait<FU_MLP_BACKOFF>(feat_full, it & 1);
#if FU_L3_MMA
            if (it > 0) ok &= mbar_wait(d3_full, (it - 1) & 1);      // layer 3 of the previous tile has finished reading the operand buffer
#endif
            tc_fence_after();
            // ---- A. this thread's 16 elements: features of its hemisphere (from TMEM), then view direction / 1 / padding ----
            {
#if FU_ALPHA_IN_GATHER
                const int yang = (int)(__float_as_uint(s_ay[(it & 3) * TC_TM + row]) >> 31);
#else
                const int yang = s_yang[(it & 3) * TC_TM + row];
#endif
                uint32_t r0[16], r1[16];
                tmem_ld16(tmem_lane + 256 + 16 * half, r0);            // yin block
                tmem_ld16(tmem_lane + 256 + 32 + 16 * half, r1);       // yang block
                float el[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) el[j] = __uint_as_float(yang ? r1[j] : r0[j]);
                if (feat_out != nullptr && live) {                     // saved for the backward pass (28 floats / sample)
                    float4* dst = reinterpret_cast<float4*>(feat_out + (size_t)gm * EGN_FEAT_STRIDE + 16 * half);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (half == 0 || q < 3) dst[q] = make_float4(el[4 * q], el[4 * q + 1], el[4 * q + 2], el[4 * q + 3]);
                }
                const float* dir = rays + (size_t)(live ? ray_of(gm) : 0u) * 6 + 3;
                if (AD == 27) {                                      // every shipped config: elements 27..29 = view direction, 30 = 1
                    if (half) { el[11] = __ldg(dir); el[12] = __ldg(dir + 1); el[13] = __ldg(dir + 2); el[14] = 1.f; el[15] = 0.f; }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int e = 16 * half + j;
                        if (e >= AD) el[j] = (e < AD + 3) ? __ldg(dir + (e - AD)) : (e == AD + 3 ? 1.f : 0.f);
                    }
                }
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    float v[40];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float x = el[8 * pass + j];
                        float s1, c1;
                        __sincosf(x, &s1, &c1);
                        v[5 * j] = x; v[5 * j + 1] = s1; v[5 * j + 2] = c1;
                        v[5 * j + 3] = 2.f * s1 * c1; v[5 * j + 4] = 1.f - 2.f * s1 * s1;   // W1's padding columns are zero
                    }
#pragma unroll
                    for (int c = 0; c < 5; ++c) store_chunk_h(as, 10 * half + 5 * pass + c, row, v + 8 * c);
                }
            }
            fence_async_smem();
            tc_fence_before();
#if FU_MMA_WARP
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);                     // X ready: the MMA warp issues layer 1
#else
            named_bar_sync(1, FU_GROUP);
            // ---- B. layer 1 (runs on the tensor pipe during layer 3 of the previous tile) ----
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < TC_K1 / 16; ++ks)
                    tc_mma(tmem, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc(w1_s + ks * 2 * TC_CHUNK), FU_IDESC_128x128, ks > 0);
                tc_commit(d1_full);
            }
#endif
            if (it > 0) layer3(gm_prev, it - 1, FU_PREFETCH ? pre : prefetch(gm_prev));
            ok &= mbar_wait(d1_full, it & 1);
            tc_fence_after();
            // ---- C. H1 = relu(D1) -> operand of layer 2 ----
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col = 64 * half + 32 * cc;
                uint32_t r[32];
                tmem_ld32(tmem_lane + col, r);
                float v[32];
#if FU_RELU_CVT
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk_h_relu(as, (col >> 3) + c, row, v + 8 * c);
#else
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]), 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk_h(as, (col >> 3) + c, row, v + 8 * c);
#endif
            }
            fence_async_smem();
            tc_fence_before();
#if FU_MMA_WARP
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);                     // H1 ready: layer 2, then layer 0 of the next tile
#else
            named_bar_sync(1, FU_GROUP);
            // ---- D. layer 2, then layer 0 of the next tile ----
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < EGN_HID / 16; ++ks)
                    tc_mma(tmem + 128, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc(w2_s + ks * 2 * TC_CHUNK), FU_IDESC_128x128, ks > 0);
                tc_commit(d2_full);
                if (it + 1 < n_local) issue_layer0(it + 1);
            }
#endif
            ok &= mbar_wait(d2_full, it & 1);
            tc_fence_after();
#if FU_L3_MMA
            // ---- E. H2 = relu(D2 + b2) -> operand of layer 3 (same buffer), D3 = H2 . W3^T ----
            {
                const float4* b2s = reinterpret_cast<const float4*>(smem + L::B2);
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = 64 * half + 32 * cc;
                    uint32_t r[32];
                    tmem_ld32(tmem_lane + 128 + col, r);
                    float v[32];
#if FU_RELU_CVT
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 bq = b2s[(col >> 2) + g];
                        const float2 s0 = fu_fadd2(make_float2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1])), make_float2(bq.x, bq.y));
                        const float2 s1 = fu_fadd2(make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])), make_float2(bq.z, bq.w));
                        v[4 * g] = s0.x; v[4 * g + 1] = s0.y; v[4 * g + 2] = s1.x; v[4 * g + 3] = s1.y;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_chunk_h_relu(as, (col >> 3) + c, row, v + 8 * c);
#else
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 bq = b2s[(col >> 2) + g];
                        v[4 * g] = fmaxf(__uint_as_float(r[4 * g]) + bq.x, 0.f); v[4 * g + 1] = fmaxf(__uint_as_float(r[4 * g + 1]) + bq.y, 0.f);
                        v[4 * g + 2] = fmaxf(__uint_as_float(r[4 * g + 2]) + bq.z, 0.f); v[4 * g + 3] = fmaxf(__uint_as_float(r[4 * g + 3]) + bq.w, 0.f);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_chunk_h(as, (col >> 3) + c, row, v + 8 * c);
#endif
                }
                fence_async_smem();
                tc_fence_before();
#if FU_MMA_WARP
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);                 // H2 ready: layer 3
#else
                named_bar_sync(1, FU_GROUP);
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t w3_s = smem_u32(smem + L::L3);
#pragma unroll
                    for (int ks = 0; ks < EGN_HID / 16; ++ks)
                        tc_mma(tmem + 384, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc_sbo(w3_s + ks * 2 * FU_W3_CHUNK, FU_W3_CHUNK, 0), FU_IDESC_128x16, ks > 0);
                    tc_commit(d3_full);
                }
#endif
            }
#endif
            gm_prev = gm;
        }
#if FU_L3_MMA
        if (it > 0) { ok &= mbar_wait(d3_full, (it - 1) & 1); tc_fence_after(); }
#endif
        if (it > 0) layer3(gm_prev, it - 1, prefetch(gm_prev));   // layer 3 of the last tile
    }
    if (!ok) __trap();                                          // a lost mbarrier arrive: fail loudly, never hang
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

int egn_launch_fused_fine(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                          float* fsig, float* feat_out, float* rgbs, const EgnOutputs* composite_out, void* image_buf,
                          cudaStream_t st) {
    const long long M = n * k.S;
    const long long tiles = (M + TC_TM - 1) / TC_TM;
    if (M >= 0xffffff00ll) return (int)cudaErrorInvalidValue;      // 32-bit sample indices inside the kernel: render in smaller chunks
    unsigned char* img = reinterpret_cast<unsigned char*>(image_buf);
    egn_fused_image_kernel<<<48, 256, 0, st>>>(k.app_dim, p->basis[0], p->basis[1], p->mlp_w[0], p->mlp_b[0], p->mlp_w[1], p->mlp_b[1],
                                               p->mlp_w[2], img);
#if FU_L3_CONST
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    std::lock_guard<std::mutex> lock(g_fu_l3_mutex);
    if (g_fu_l3_event[dev] == nullptr) cudaEventCreateWithFlags(&g_fu_l3_event[dev], cudaEventDisableTiming);
    else cudaStreamWaitEvent(st, g_fu_l3_event[dev], 0);          // the previous launch (any stream) has finished reading the symbol
    cudaMemcpyToSymbolAsync(c_fu_l3, img + FuLayout::L3, EGN_HID * sizeof(float4), 0, cudaMemcpyDeviceToDevice, st);
    struct Mark { cudaEvent_t e; cudaStream_t s; ~Mark() { cudaEventRecord(e, s); } } mark{g_fu_l3_event[dev], st};   // after the launch below
#endif
    if (composite_out != nullptr) {           // compositing inside the kernel: CTAs walk whole rays
#ifndef FU_MAX_BLOCKS
#define FU_MAX_BLOCKS 148
#endif
        const int blocks = (int)(n < FU_MAX_BLOCKS ? n : FU_MAX_BLOCKS);
        cudaFuncSetAttribute(egn_fused_fine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuLayout::TOTAL);
#ifdef FU_CARVEOUT
        cudaFuncSetAttribute(egn_fused_fine_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, FU_CARVEOUT);
#endif
        egn_fused_fine_kernel<true><<<blocks, FU_THREADS, FuLayout::TOTAL, st>>>(k, img, p->mlp_b[2], rays, M, z, fsig, nullptr, rgbs,
                                                                                 p->emission, *composite_out);
        return (int)cudaGetLastError();
    }
    const int blocks = (int)(tiles < 148 ? tiles : 148);
    cudaFuncSetAttribute(egn_fused_fine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuLayout::TOTAL);
    egn_fused_fine_kernel<false><<<blocks, FU_THREADS, FuLayout::TOTAL, st>>>(k, img, p->mlp_b[2], rays, M, z, fsig, feat_out, rgbs,
                                                                              nullptr, EgnOutputs{});
    return (int)cudaGetLastError();
}

// bytes of the operand image a caller of egn_launch_fused_fine must provide (16-byte aligned)
long long egn_fused_image_bytes() { return FuLayout::IMAGE; }
