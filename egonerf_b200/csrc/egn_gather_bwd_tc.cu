// Gather backward, throughput mode: same function as egn_gather_bwd_kernel (egn_backward.cu) with its two dense
// contractions moved from shared-memory FFMA loops onto the tensor cores (ncu on the FFMA version: L1/LSU pipe 80 % busy,
// 1.1 G shared-memory wavefronts per 4 M samples for dv = B_h^T dfeat and d(basis) += dfeat^T v).
//
// Per 128-sample tile, one CTA per SM, 512 threads:
//   Ph0  warp w: Yin-Yang coordinates + d(sigma feature) of rows 8w..8w+7 (kept in registers), hemisphere flags -> smem
//   Ph1  thread (row, q): 8 values of d_feat -> fp16 tile S * dF2[128][64] (S: launch-wide power of two, tc_grad_scale()) (column block 32*hemisphere, other block zero)
//        MMA1  dV[128 x 144] = dF2 . [B_yin ; B_yang]        (A K-major K = 64; B = MN-major view of the basis operand)
//   Ph2  dV: TMEM -> fp32 smem tile (rows -> the gather lanes that need them)
//   Ph3  warp w, half-warp per sample: re-gather the 18 taps (branch-free clamped loads), P, L, v = P*L;
//        d(plane) = up*L, d(line) = up*P scattered with 128-bit reductions; v rows -> bf16 tile V[128][144]
//        MMA2  dB[64(+64 zero) x 144] += dF2^T . V            (both operands MN-major views; accumulator stays in TMEM)
//   end  dB: TMEM -> atomics into d(basis_mat_yin / yang)
#include "egn_tc.cuh"
#include "egn_host.h"
#include "egn_shared.cuh"

#define GT_THREADS 512
#define GT_VK (3 * EGN_CA)                 // 144
#define GT_BB_CHUNK 1024
#define GT_DVS 148                         // dv smem row stride (floats)
#define GT_VCHUNK (TC_CHUNK + 64)          // padded K-chunk stride of the V tile (bank-conflict-free row stores, see egn_fused.cu)
#define IDESC_DV TC_IDESC_F16(0x08250490u)               // M128 N144, A K-major, B MN-major
#define IDESC_DB TC_IDESC_F16(0x08258490u)               // M128 N144, A MN-major, B MN-major
// (Staging dF2 / MMA1 of tile t+1 before the scatter phase of tile t -- two dF2 tiles, two dV accumulators -- was built and
// measured at 6.91 vs 6.90 ms per training step: the barriers do not wait for the MMAs.  Removed; commit a1a3ea4 has it.)
#define GT_TM_DV 0                   // dV accumulator: columns 0..143
#define GT_TM_DB 160

struct GtLayout {
    static constexpr int BB = 0;                                          // [64 n][144 k]      18 432
    static constexpr int DF = BB + (GT_VK / 8) * GT_BB_CHUNK;             // [128 m][128]       32 768 (cols 64.. zero)
    static constexpr int V = DF + 16 * TC_CHUNK;                // [128 m][144] bf16  36 864
    static constexpr int DV = V + (GT_VK / 8) * GT_VCHUNK;                 // [128 m][148] fp32  75 776
    static constexpr int KNOTS = DV + TC_TM * GT_DVS * 4;
    static constexpr int YANG = KNOTS + ((EGN_MAX_KNOTS + 1) * 4 + 15) / 16 * 16;
    static constexpr int COORD = YANG + TC_TM;                   // [128] float4: normalised r, polar, azimuth, flags
    static constexpr int DSG = COORD + TC_TM * 16;               // [128] d(sigma feature)
    static constexpr int MBAR = DSG + TC_TM * 4;
    static constexpr int TMEM = MBAR + 16;
    // GT_PREFETCH: inputs of the NEXT tile, fetched by cp.async while this tile is being scattered
    static constexpr int NXT_DF = TMEM + 16;                               // [128 m][28] d_feat rows   14 336
    static constexpr int NXT_CO = NXT_DF + TC_TM * EGN_FEAT_STRIDE * 4;     // [128] saved coordinates     2 048
    static constexpr int NXT_DS = NXT_CO + TC_TM * 16;                      // [128] d(sigma feature)        512
    static constexpr int TOTAL = NXT_DS + TC_TM * 4;
};
static_assert(GtLayout::TOTAL <= 227 * 1024, "gather backward kernel exceeds the shared memory of one SM");

#ifndef GT_PREFETCH
#define GT_PREFETCH 1                // 1: d_feat / saved coordinates / d(sigma feature) of tile t+1 land in shared memory (cp.async) during Ph3 of tile t
#endif
__device__ __forceinline__ void gt_cp16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gt_cp4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gt_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gt_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifndef GT_DIAG
#define GT_DIAG 0                    // timing diagnostics only (WRONG gradients): 1 = no reductions for the (r, angle) planes, 2 = no reductions at all
#endif
__device__ __forceinline__ void gt_red4(float* addr, float4 v) {
#if GT_DIAG < 2
    atomicAdd(reinterpret_cast<float4*>(addr), v);
#endif
}
#ifndef GT_FFMA2
#define GT_FFMA2 0                   // 1: Ph3 arithmetic with the packed fp32 instructions of sm_100 (FFMA2 / FMUL2; same roundings) -- measured 6.915 vs 6.912 ms per step: off
#endif
#if GT_FFMA2
__device__ __forceinline__ unsigned long long gt_pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void gt_unpk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ float4 gt_fma4(float w, float4 a, float4 acc) {
    const unsigned long long ww = gt_pk(w, w);
    unsigned long long lo, hi;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(lo) : "l"(ww), "l"(gt_pk(a.x, a.y)), "l"(gt_pk(acc.x, acc.y)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(hi) : "l"(ww), "l"(gt_pk(a.z, a.w)), "l"(gt_pk(acc.z, acc.w)));
    float4 r;
    gt_unpk(lo, r.x, r.y); gt_unpk(hi, r.z, r.w);
    return r;
}
__device__ __forceinline__ float4 gt_mul4(float4 a, float4 b) {
    unsigned long long lo, hi;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(lo) : "l"(gt_pk(a.x, a.y)), "l"(gt_pk(b.x, b.y)));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hi) : "l"(gt_pk(a.z, a.w)), "l"(gt_pk(b.z, b.w)));
    float4 r;
    gt_unpk(lo, r.x, r.y); gt_unpk(hi, r.z, r.w);
    return r;
}
__device__ __forceinline__ float4 gt_scale(float s, float4 a) { return gt_mul4(make_float4(s, s, s, s), a); }
#else
__device__ __forceinline__ float4 gt_fma4(float w, float4 a, float4 acc) { return f4fma(w, a, acc); }
__device__ __forceinline__ float4 gt_mul4(float4 a, float4 b) { return f4mul(a, b); }
__device__ __forceinline__ float4 gt_scale(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
#endif

// register cache of the angular-tap gradients of one half-warp (see Ph3)
#ifndef GT_LINES_CACHED
#define GT_LINES_CACHED 2            // 2: phi / theta lines; 3: the r line too (10 more registers)
#endif
#ifndef GT_TAP_DEPTH
#define GT_TAP_DEPTH 3               // factor pairs of taps in flight per half-warp (3 = all 18 taps up front; 2 frees no registers in practice)
#endif
#ifndef GT_SHIFT
#define GT_SHIFT 0                   // 1: the (r, angle) planes and the r line keep the contributions of their j1 column / tap pending for one sample: along
#endif                               //    a ray r advances about one texel per sample, so the next sample's j0 column is that very column (4 -> 2 and 2 -> 1
                                     //    reductions).  Correct (gradient tests green) but 25 more registers at the 128 cap: 104 B of spills in the scatter loop and
                                     //    half-warp-divergent flushes -- measured 8.18 vs 6.90 ms per training step: off (profiles/r02_backward.md)
struct GtCache {
    unsigned po[4], lo[GT_LINES_CACHED][2];        // element offsets of the cached taps (0xffffffff = empty)
    float4 pa[4], la[GT_LINES_CACHED][2];
#if GT_SHIFT
    unsigned so[2][2], ro;                          // pending j1 column (rows a, b) of planes 0 / 1; pending j1 tap of the r line
    float4 sa[2][2], ra;
    __device__ __forceinline__ void flush_shift(float* d_tab, int i) {
        if (so[i][0] != 0xffffffffu) {
            if (nz(sa[i][0])) gt_red4(d_tab + so[i][0], sa[i][0]);
            if (nz(sa[i][1])) gt_red4(d_tab + so[i][1], sa[i][1]);
        }
    }
    __device__ __forceinline__ void flush_rline(float* d_tab) {
        if (ro != 0xffffffffu && nz(ra)) gt_red4(d_tab + ro, ra);
    }
#endif
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int t = 0; t < 4; ++t) { po[t] = 0xffffffffu; pa[t] = f4zero(); }
#pragma unroll
        for (int i = 0; i < GT_LINES_CACHED; ++i) { lo[i][0] = lo[i][1] = 0xffffffffu; la[i][0] = la[i][1] = f4zero(); }
#if GT_SHIFT
#pragma unroll
        for (int i = 0; i < 2; ++i) { so[i][0] = so[i][1] = 0xffffffffu; sa[i][0] = sa[i][1] = f4zero(); }
        ro = 0xffffffffu; ra = f4zero();
#endif
    }
    __device__ __forceinline__ static bool nz(const float4& v) { return (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f); }
    __device__ __forceinline__ void flush_plane(float* d_tab) {
        if (po[0] != 0xffffffffu) {
#pragma unroll
            for (int t = 0; t < 4; ++t) { if (nz(pa[t])) gt_red4(d_tab + po[t], pa[t]); pa[t] = f4zero(); }
        }
    }
    __device__ __forceinline__ void flush_line(float* d_tab, int i) {
        if (lo[i][0] != 0xffffffffu) {
            if (nz(la[i][0])) gt_red4(d_tab + lo[i][0], la[i][0]);
            if (nz(la[i][1])) gt_red4(d_tab + lo[i][1], la[i][1]);
            la[i][0] = la[i][1] = f4zero();
        }
    }
};

// one lane's 4 channels of a tap: fp32 tables (16 B) or the bf16 copy the throughput-mode forward gathered from (8 B —
// half the bytes and half the L2 footprint next to the 99 MB gradient table)
template <bool BF16>
__device__ __forceinline__ float4 gt_tap(const EgnKernelCfg& k, unsigned off) {
    if constexpr (BF16) {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(k.tables_bf16) + off));
        return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                           __uint_as_float(q.y & 0xffff0000u));
    } else {
        return ldg4(k.tables + off);
    }
}
#ifndef GT_L2_KEEP
#define GT_L2_KEEP 0                 // 1: table taps loaded with an L2 evict_last policy (the 50 MB bf16 table competes with the 99 MB gradient table for L2)
#endif
__device__ __forceinline__ uint2 gt_raw(const EgnKernelCfg& k, unsigned off) {
#if GT_L2_KEEP
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y)
                 : "l"(reinterpret_cast<const __nv_bfloat16*>(k.tables_bf16) + off), "l"(pol));
    return v;
#else
    return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(k.tables_bf16) + off));
#endif
}
__device__ __forceinline__ float4 gt_widen(const uint2& q) {
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u), __uint_as_float(q.y << 16),
                       __uint_as_float(q.y & 0xffff0000u));
}

template <bool BF16>
__global__ void __launch_bounds__(GT_THREADS, 1)
egn_gather_bwd_tc_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ basis0,
                         const float* __restrict__ basis1, const float* __restrict__ rays, long long M,
                         const float* __restrict__ zs, const float* __restrict__ d_fsig, const float* __restrict__ d_feat,
                         const unsigned* __restrict__ gmax_bits, float* __restrict__ d_tab, float* __restrict__ d_basis0,
                         float* __restrict__ d_basis1) {
    using L = GtLayout;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // shuffle: provably warp-uniform
    const int row = tid & 127, q = tid >> 7;
    const int AD = k.app_dim;
    unsigned char* bbs = smem + L::BB;
    unsigned char* dfs = smem + L::DF;
    unsigned char* vs = smem + L::V;
    float* dvs = reinterpret_cast<float*>(smem + L::DV);
    float* s_knots = reinterpret_cast<float*>(smem + L::KNOTS);
    unsigned char* s_yang = smem + L::YANG;
    float4* s_coord = reinterpret_cast<float4*>(smem + L::COORD);
    float* s_dsg = reinterpret_cast<float*>(smem + L::DSG);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::TMEM);
    const uint32_t bar = smem_u32(smem + L::MBAR);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(bar, 1);
        mbar_init(bar + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 64 * GT_VK; i += GT_THREADS) {
        const int n = i / GT_VK, kk = i % GT_VK, o = n & 31;
        const float* B = (n >> 5) ? basis1 : basis0;
        store_elem_h(bbs, n, kk, o < AD ? B[o * GT_VK + kk] : 0.f, GT_BB_CHUNK);
    }
    for (int i = tid; i < 16 * TC_CHUNK / 16; i += GT_THREADS) reinterpret_cast<uint4*>(dfs)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i <= k.knots_last; i += GT_THREADS) s_knots[i] = k.r_knots[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t bb_s = smem_u32(bbs), df_s = smem_u32(dfs), v_s = smem_u32(vs);
    const int sub = lane & 15;
    float inv_scale;
    const float scale = tc_grad_scale(gmax_bits, inv_scale);

    const long long tiles = (M + TC_TM - 1) / TC_TM;
    uint32_t it = 0;
    bool ok = true;
#if GT_PREFETCH
    float4* nx_df = reinterpret_cast<float4*>(smem + L::NXT_DF);
    float4* nx_co = reinterpret_cast<float4*>(smem + L::NXT_CO);
    float* nx_ds = reinterpret_cast<float*>(smem + L::NXT_DS);
    // the tile's inputs are contiguous in global memory: 128 rows x 7 float4 of d_feat, 128 float4 of coordinates, 128 floats
    auto prefetch = [&](long long t) {
        if (t < tiles) {
            const long long m0 = t * TC_TM;
            const long long live = (M - m0 < TC_TM) ? (M - m0) : TC_TM;
            const float4* src = reinterpret_cast<const float4*>(d_feat + m0 * EGN_FEAT_STRIDE);
            for (int i = tid; i < live * (EGN_FEAT_STRIDE / 4); i += GT_THREADS) gt_cp16(nx_df + i, src + i);
            if (tid < live) {
                if (k.coords != nullptr) gt_cp16(nx_co + tid, reinterpret_cast<const float4*>(k.coords) + m0 + tid);
                gt_cp4(nx_ds + tid, d_fsig + m0 + tid);
            }
        }
        gt_cp_commit();
    };
    prefetch(blockIdx.x);
    gt_cp_wait();                                                // each thread's own copies have landed ...
    __syncthreads();                                             // ... and so have everybody else's
#endif
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        // ---- Ph1a. this thread's 8 values of d_feat ----
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
#if GT_PREFETCH
        {                                                        // staged by cp.async during the previous tile's Ph3
            const long long gm = tile * TC_TM + row;
            if (gm < M) {
                const float4 a = nx_df[row * (EGN_FEAT_STRIDE / 4) + 2 * q];
                v[0] = a.x * scale; v[1] = a.y * scale; v[2] = a.z * scale; v[3] = a.w * scale;
                if (q < 3) { const float4 b = nx_df[row * (EGN_FEAT_STRIDE / 4) + 2 * q + 1]; v[4] = b.x * scale; v[5] = b.y * scale; v[6] = b.z * scale; v[7] = b.w * scale; }
            }
        }
#else
        {
            const long long gm = tile * TC_TM + row;
            if (gm < M) {
                const float4* f4 = reinterpret_cast<const float4*>(d_feat + gm * EGN_FEAT_STRIDE) + 2 * q;
                const float4 a = __ldg(f4);
                v[0] = a.x * scale; v[1] = a.y * scale; v[2] = a.z * scale; v[3] = a.w * scale;
                if (q < 3) { const float4 b = __ldg(f4 + 1); v[4] = b.x * scale; v[5] = b.y * scale; v[6] = b.z * scale; v[7] = b.w * scale; }
            }
        }
#endif
        // ---- Ph0. coordinates + d(sigma feature) of the tile's 128 samples: one sample per lane of warps 0..3 -> smem ----
        if (warp < 4) {
            const int r = 32 * warp + lane;
            const long long mg = tile * TC_TM + r;
            YYCoord cc;
            cc.c[0] = cc.c[1] = cc.c[2] = -3.f;
            cc.yang = 0;
            float dsg = 0.f;
            const bool glive = mg < M;
            if (glive) {
                if (k.coords != nullptr) {                           // saved by the fused forward (egn_fused.cu, phase 1b)
#if GT_PREFETCH
                    const float4 sv = nx_co[r];
#else
                    const float4 sv = __ldg(reinterpret_cast<const float4*>(k.coords) + mg);
#endif
                    cc.c[0] = sv.x; cc.c[1] = sv.y; cc.c[2] = sv.z; cc.yang = __float_as_int(sv.w);
                } else {
                    const long long ray = egn_ray_of(mg, k.S);
                    const float z = zs[mg];
                    const float* ry = rays + ray * 6;
                    cc = egn_cart_to_yinyang(ry[0] + ry[3] * z, ry[1] + ry[4] * z, ry[2] + ry[5] * z, k, s_knots);
                }
#if GT_PREFETCH
                dsg = nx_ds[r];
#else
                dsg = d_fsig[mg];
#endif
            }
            s_coord[r] = make_float4(cc.c[0], cc.c[1], cc.c[2], __int_as_float(cc.yang | (glive ? 2 : 0)));
            s_dsg[r] = dsg;
            s_yang[r] = (unsigned char)cc.yang;
        }
        if (it > 0) ok &= mbar_wait(bar + 8, (it - 1) & 1);     // MMA2 of the previous tile has finished with dF2 and V
        __syncthreads();
        // ---- Ph1b. dF2 tile ----
        {
            const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int yang = s_yang[row];
            store_chunk_h(dfs, 4 * yang + q, row, v);
            store_chunk_h(dfs, 4 * (1 - yang) + q, row, zero);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
#if GT_PREFETCH
        prefetch(tile + gridDim.x);                             // staging consumed by everybody (Ph0, Ph1a): refill it behind Ph2 / Ph3
#endif
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)                      // K = 64 basis outputs (yin | yang)
                tc_mma(tmem + GT_TM_DV, desc_k(df_s + ks * 2 * TC_CHUNK), desc_mn(bb_s + ks * 256, GT_BB_CHUNK), IDESC_DV, ks > 0);
            tc_commit(bar);
        }
        // ---- Ph2. dV rows -> smem ----
        ok &= mbar_wait(bar, par);
        tc_fence_after();
        {
            uint32_t r[32], r4[4];
            tmem_ld32(tmem_lane + GT_TM_DV + 36 * q, r);
            tmem_ld4(tmem_lane + GT_TM_DV + 36 * q + 32, r4);
            float4* dst = reinterpret_cast<float4*>(dvs + row * GT_DVS + 36 * q);
#pragma unroll
            for (int g = 0; g < 8; ++g)
                dst[g] = make_float4(__uint_as_float(r[4 * g]) * inv_scale, __uint_as_float(r[4 * g + 1]) * inv_scale,
                                     __uint_as_float(r[4 * g + 2]) * inv_scale, __uint_as_float(r[4 * g + 3]) * inv_scale);
            dst[8] = make_float4(__uint_as_float(r4[0]) * inv_scale, __uint_as_float(r4[1]) * inv_scale, __uint_as_float(r4[2]) * inv_scale,
                                 __uint_as_float(r4[3]) * inv_scale);
        }
        tc_fence_before();
        __syncthreads();
        // ---- Ph3. re-gather, local gradients, scatter, V rows ----
        // Ray coherence: the 8 samples of this warp are consecutive along one ray, so the angular plane (theta x phi) and
        // the theta / phi lines are hit at the same texels by (almost) all of them.  Their contributions are summed in
        // registers per half-warp and flushed with ONE reduction per tap when the texel changes or the tile ends
        // (8 of the 18 reductions per sample become ~2 per 4 samples).
        GtCache cache;
        cache.reset();
#pragma unroll 1
        for (int itr = 0; itr < 4; ++itr) {
#ifndef GT_CONSECUTIVE
#define GT_CONSECUTIVE 1
#endif
            // each half-warp walks 4 CONSECUTIVE samples of the ray (0..3 / 4..7): neighbours share texels far more often than
            // samples two apart, so the register cache below flushes less
            const int src = GT_CONSECUTIVE ? 4 * (lane >> 4) + itr : 2 * itr + (lane >> 4);              // sample within the warp's 8
            const int srow = 8 * warp + src;
            const float4 sc = s_coord[srow];
            const float c[3] = {sc.x, sc.y, sc.z};
            const int yang = __float_as_int(sc.w) & 1;
            const bool slive = (__float_as_int(sc.w) & 2) != 0;
            const float dsig = s_dsg[srow];
            unsigned j0[3], j1[3];
            float wa0[3], wa1[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int G = k.lay.G[a];
                const float ix = egn_unnorm(c[a], G);
                const float fl = floorf(ix);
                const float fr = ix - fl;
                const int i0 = (int)fminf(fmaxf(fl, -2.f), (float)G + 1.f);
                wa0[a] = ((i0 >= 0) & (i0 < G)) ? 1.f - fr : 0.f;
                wa1[a] = ((i0 + 1 >= 0) & (i0 + 1 < G)) ? fr : 0.f;
                j0[a] = (unsigned)min(max(i0, 0), G - 1);
                j1[a] = (unsigned)min(max(i0 + 1, 0), G - 1);
            }
            // bf16 tables: all 18 taps of the sample are requested before the first one is used (8 B per lane and tap stay
            // packed until then) -- three times the loads in flight of a per-factor-pair loop
            // (GT_TAP_DEPTH 2: the third factor pair is requested as soon as the first one has been widened -- 12 registers
            // less for the pending-column caches of GT_SHIFT)
            uint2 raw[BF16 ? GT_TAP_DEPTH : 1][6];
            auto request = [&](int i, uint2 (&dst)[6]) {
                const int ax = egn_mx(i), ay = egn_my(i), al = egn_vl(i);
                const unsigned W = (unsigned)k.lay.G[ax];
                const unsigned pbase = (unsigned)k.lay.pf[yang][i] + sub * 4, lbase = (unsigned)k.lay.lf[yang][i] + sub * 4;
                const unsigned ra = j0[ay] * W, rb = j1[ay] * W;
                dst[0] = gt_raw(k, pbase + (ra + j0[ax]) * EGN_CF); dst[1] = gt_raw(k, pbase + (ra + j1[ax]) * EGN_CF);
                dst[2] = gt_raw(k, pbase + (rb + j0[ax]) * EGN_CF); dst[3] = gt_raw(k, pbase + (rb + j1[ax]) * EGN_CF);
                dst[4] = gt_raw(k, lbase + j0[al] * EGN_CF); dst[5] = gt_raw(k, lbase + j1[al] * EGN_CF);
            };
            if constexpr (BF16) {
#pragma unroll
                for (int i = 0; i < GT_TAP_DEPTH; ++i) request(i, raw[i]);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int ax = egn_mx(i), ay = egn_my(i), al = egn_vl(i);
                const unsigned W = (unsigned)k.lay.G[ax];
                const unsigned pbase = (unsigned)k.lay.pf[yang][i] + sub * 4, lbase = (unsigned)k.lay.lf[yang][i] + sub * 4;
                const unsigned ra = j0[ay] * W, rb = j1[ay] * W;
                const unsigned o0 = pbase + (ra + j0[ax]) * EGN_CF, o1 = pbase + (ra + j1[ax]) * EGN_CF;
                const unsigned o2 = pbase + (rb + j0[ax]) * EGN_CF, o3 = pbase + (rb + j1[ax]) * EGN_CF;
                const unsigned q0 = lbase + j0[al] * EGN_CF, q1 = lbase + j1[al] * EGN_CF;
                float4 t0, t1, t2, t3, l0, l1;
                if constexpr (BF16) {
                    uint2 (&rw)[6] = raw[i % GT_TAP_DEPTH];
                    t0 = gt_widen(rw[0]); t1 = gt_widen(rw[1]); t2 = gt_widen(rw[2]); t3 = gt_widen(rw[3]);
                    l0 = gt_widen(rw[4]); l1 = gt_widen(rw[5]);
                    if (i + GT_TAP_DEPTH < 3) request(i + GT_TAP_DEPTH, rw);
                } else {
                    t0 = gt_tap<BF16>(k, o0); t1 = gt_tap<BF16>(k, o1); t2 = gt_tap<BF16>(k, o2); t3 = gt_tap<BF16>(k, o3);
                    l0 = gt_tap<BF16>(k, q0); l1 = gt_tap<BF16>(k, q1);
                }
                const float w0 = wa0[ax] * wa0[ay], w1 = wa1[ax] * wa0[ay], w2 = wa0[ax] * wa1[ay], w3 = wa1[ax] * wa1[ay];
                const float u0 = wa0[al], u1 = wa1[al];
                float4 P = f4zero();
                P = gt_fma4(w0, t0, P); P = gt_fma4(w1, t1, P); P = gt_fma4(w2, t2, P); P = gt_fma4(w3, t3, P);
                float4 Lv = f4zero();
                Lv = gt_fma4(u0, l0, Lv); Lv = gt_fma4(u1, l1, Lv);
                const float4 prod = gt_mul4(P, Lv);
                float s = hsum4(prod);                          // density lanes: relu mask of this product (EgoNeRF.py:346)
                s += __shfl_xor_sync(FULL, s, 1);
                s += __shfl_xor_sync(FULL, s, 2);
                float4 up;
                if (sub < EGN_CS / 4) {
                    const float ds = (s > 0.f) ? dsig : 0.f;
                    up = make_float4(ds, ds, ds, ds);
                } else {
                    const int kk = i * EGN_CA + (sub - EGN_CS / 4) * 4;
                    up = *reinterpret_cast<const float4*>(dvs + srow * GT_DVS + kk);
                    *reinterpret_cast<uint2*>(vs + (kk >> 3) * GT_VCHUNK + srow * 16 + (kk & 7) * 2) =
                        make_uint2(pack_h2(prod.x, prod.y), pack_h2(prod.z, prod.w));
                }
                if (slive) {
                    const float4 dP = gt_mul4(up, Lv), dL = gt_mul4(up, P);
                    if (i == 2) {                                   // angular plane: cached
                        if (cache.po[0] != o0 || cache.po[3] != o3) {
                            cache.flush_plane(d_tab);
                            cache.po[0] = o0; cache.po[1] = o1; cache.po[2] = o2; cache.po[3] = o3;
                        }
                        cache.pa[0] = gt_fma4(w0, dP, cache.pa[0]); cache.pa[1] = gt_fma4(w1, dP, cache.pa[1]);
                        cache.pa[2] = gt_fma4(w2, dP, cache.pa[2]); cache.pa[3] = gt_fma4(w3, dP, cache.pa[3]);
                    } else if (GT_DIAG == 0) {
#if GT_SHIFT
                        // planes (r, theta) / (r, phi): column j0 = taps 0, 2; column j1 = taps 1, 3 (r is the fast axis)
                        float4 c0 = gt_scale(w0, dP), c2 = gt_scale(w2, dP);
                        const float4 c1 = gt_scale(w1, dP), c3 = gt_scale(w3, dP);
                        const bool shift = cache.so[i][0] == o0 && cache.so[i][1] == o2;      // pending column is this sample's j0 column
                        const bool same = cache.so[i][0] == o1 && cache.so[i][1] == o3;       // r did not advance: pending column is the j1 column again
                        if (shift) {
                            c0.x += cache.sa[i][0].x; c0.y += cache.sa[i][0].y; c0.z += cache.sa[i][0].z; c0.w += cache.sa[i][0].w;
                            c2.x += cache.sa[i][1].x; c2.y += cache.sa[i][1].y; c2.z += cache.sa[i][1].z; c2.w += cache.sa[i][1].w;
                        } else if (!same) {
                            cache.flush_shift(d_tab, i);
                        }
                        if (GtCache::nz(c0)) gt_red4(d_tab + o0, c0);
                        if (GtCache::nz(c2)) gt_red4(d_tab + o2, c2);
                        if (same) {
                            cache.sa[i][0].x += c1.x; cache.sa[i][0].y += c1.y; cache.sa[i][0].z += c1.z; cache.sa[i][0].w += c1.w;
                            cache.sa[i][1].x += c3.x; cache.sa[i][1].y += c3.y; cache.sa[i][1].z += c3.z; cache.sa[i][1].w += c3.w;
                        } else {
                            cache.so[i][0] = o1; cache.so[i][1] = o3;
                            cache.sa[i][0] = c1; cache.sa[i][1] = c3;
                        }
#else
                        if (w0 != 0.f) gt_red4(d_tab + o0, gt_scale(w0, dP));
                        if (w1 != 0.f) gt_red4(d_tab + o1, gt_scale(w1, dP));
                        if (w2 != 0.f) gt_red4(d_tab + o2, gt_scale(w2, dP));
                        if (w3 != 0.f) gt_red4(d_tab + o3, gt_scale(w3, dP));
#endif
                    }
                    if (i < GT_LINES_CACHED) {                      // phi / theta (/ r) lines: cached
                        if (cache.lo[i][0] != q0 || cache.lo[i][1] != q1) {
                            cache.flush_line(d_tab, i);
                            cache.lo[i][0] = q0; cache.lo[i][1] = q1;
                        }
                        cache.la[i][0] = gt_fma4(u0, dL, cache.la[i][0]);
                        cache.la[i][1] = gt_fma4(u1, dL, cache.la[i][1]);
                    } else {
#if GT_SHIFT
                        float4 d0 = gt_scale(u0, dL);
                        const float4 d1 = gt_scale(u1, dL);
                        const bool shift = cache.ro == q0, same = cache.ro == q1;
                        if (shift) { d0.x += cache.ra.x; d0.y += cache.ra.y; d0.z += cache.ra.z; d0.w += cache.ra.w; }
                        else if (!same) cache.flush_rline(d_tab);
                        if (GtCache::nz(d0)) gt_red4(d_tab + q0, d0);
                        if (same) { cache.ra.x += d1.x; cache.ra.y += d1.y; cache.ra.z += d1.z; cache.ra.w += d1.w; }
                        else { cache.ro = q1; cache.ra = d1; }
#else
                        if (u0 != 0.f) gt_red4(d_tab + q0, gt_scale(u0, dL));
                        if (u1 != 0.f) gt_red4(d_tab + q1, gt_scale(u1, dL));
#endif
                    }
                }
            }
        }
        cache.flush_plane(d_tab);
#pragma unroll
        for (int i = 0; i < GT_LINES_CACHED; ++i) cache.flush_line(d_tab, i);
#if GT_SHIFT
        cache.flush_shift(d_tab, 0);
        cache.flush_shift(d_tab, 1);
        cache.flush_rline(d_tab);
#endif
#if GT_PREFETCH
        gt_cp_wait();                                            // next tile's inputs: own copies landed; the barrier below covers the rest
#endif
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- MMA2: d(basis) += dF2^T V over the tile's 128 samples ----
        if (tid == 0) {
            tc_fence_after();
            const uint32_t first = it > 0 ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < TC_TM / 16; ++ks)
                tc_mma(tmem + GT_TM_DB, desc_mn(df_s + ks * 256), desc_mn(v_s + ks * 256, GT_VCHUNK), IDESC_DB, first | (ks > 0));
            tc_commit(bar + 8);
        }
    }
    if (it > 0) ok &= mbar_wait(bar + 8, (it - 1) & 1);
    if (!ok) __trap();
    tc_fence_after();
    // ---- flush d(basis): accumulator row n = 32 * hemisphere + output, column = input index ----
    if ((warp & 3) < 2) {
        uint32_t r[32], r4[4];
        tmem_ld32(tmem_lane + GT_TM_DB + 36 * q, r);
        tmem_ld4(tmem_lane + GT_TM_DB + 36 * q + 32, r4);
        const int o = row & 31;
        float* dB = (row >> 5) ? d_basis1 : d_basis0;
        if (o < AD && dB != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dB + o * GT_VK + 36 * q + j, __uint_as_float(r[j]) * inv_scale);
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(dB + o * GT_VK + 36 * q + 32 + j, __uint_as_float(r4[j]) * inv_scale);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

int egn_launch_gather_bwd_tc(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                             const float* d_fsig, const float* d_feat, const unsigned* gmax_bits, float* d_tables, const EgnGrads* g,
                             cudaStream_t st) {
    const long long M = n * k.S;
    if (M <= 0) return 0;
    const long long tiles = (M + TC_TM - 1) / TC_TM;
    const int blocks = (int)(tiles < 148 ? tiles : 148);
    if (k.tables_bf16 != nullptr) {
        cudaFuncSetAttribute(egn_gather_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GtLayout::TOTAL);
        egn_gather_bwd_tc_kernel<true><<<blocks, GT_THREADS, GtLayout::TOTAL, st>>>(k, p->basis[0], p->basis[1], rays, M, z, d_fsig,
                                                                                     d_feat, gmax_bits, d_tables, g->basis[0], g->basis[1]);
    } else {
        cudaFuncSetAttribute(egn_gather_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GtLayout::TOTAL);
        egn_gather_bwd_tc_kernel<false><<<blocks, GT_THREADS, GtLayout::TOTAL, st>>>(k, p->basis[0], p->basis[1], rays, M, z, d_fsig,
                                                                                      d_feat, gmax_bits, d_tables, g->basis[0], g->basis[1]);
    }
    return (int)cudaGetLastError();
}
