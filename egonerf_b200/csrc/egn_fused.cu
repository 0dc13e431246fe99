// Fused fine pass (throughput mode, EGN_MLP_TC_BF16): Yin-Yang coordinates -> 18-tap factor gather -> VM products
// -> basis contraction -> positional encoding -> 3-layer MLP -> sample colour, in ONE persistent warp-specialised kernel.
// Replaces egn_gather_kernel + egn_mlp_*_kernel for one ray chunk (EgoNeRF.py:544-556: from_cartesian / normalize_coord,
// compute_densityfeature, compute_appfeature, renderModule).  One CTA per SM, 512 threads:
//
//   warps 8..15  GATHER group   per 128-sample tile: coordinates, taps (fp32 or bf16 tables, 128-bit loads), P*L
//                               products; the 144 appearance products of a sample go straight into shared memory as a
//                               bf16 row of the tcgen05 A operand V (canonical K-major layout), sigma feature -> HBM.
//                               V is double buffered: tile i+1 is gathered while tile i runs through the MLP.
//   warps 0..7   MLP group      feat2 = V [B_yin | B_yang]^T (tcgen05, N = 64; the epilogue picks the sample's
//                               hemisphere) -> PE -> X -> D1 -> relu -> H1 -> D2 -> relu . W3 -> sigmoid   (as egn_mlp_tc.cu)
//   thread 0                    issues every tcgen05.mma; completions come back through tcgen05.commit -> mbarrier
//
// Nothing of size (samples x features) touches HBM in between: per sample the kernel reads 24 B of ray, 4 B of depth and
// its taps, and writes 4 B (sigma feature) + 12 B (colour) [+ 112 B app feature when the backward pass will need it].
#include "egn_tc.cuh"
#include "egn_host.h"
#include "egn_shared.cuh"

#define FU_THREADS 512
#define FU_GROUP 256
#define FU_VK (3 * EGN_CA)                   // 144
#define FU_VCHUNKS (FU_VK / 8)               // 18
// K-chunk stride of the V operand: 2048 B of data + 64 B of padding.  The 6 appearance lanes of a sample store chunks
// kc, kc+1, .. of the SAME row; with a 2048 B stride they all fall into one 16-byte bank group (6-way conflict), with
// 2112 B consecutive chunks alternate between two groups and the 24 lanes of a store spread 3 per group = the minimum.
#define FU_VCHUNK (TC_CHUNK + 64)
#define FU_VBYTES (FU_VCHUNKS * FU_VCHUNK)   // 38 016
#define FU_BB_CHUNK 1024                     // basis operand: 64 rows x 16 B per K chunk

struct FuLayout {
    static constexpr int W1 = 0;
    static constexpr int W2 = W1 + (TC_K1 / 8) * TC_CHUNK;            // 40 960
    static constexpr int BB = W2 + (EGN_HID / 8) * TC_CHUNK;          // + 32 768
    static constexpr int A = BB + FU_VCHUNKS * FU_BB_CHUNK;           // + 18 432
    static constexpr int V = A + (TC_K1 / 8) * TC_CHUNK;              // + 40 960 ; two buffers
    static constexpr int YANG = V + 2 * FU_VBYTES;                    // 4 x 128 bytes
    static constexpr int KNOTS = YANG + 4 * TC_TM;
    static constexpr int MBAR = KNOTS + ((EGN_MAX_KNOTS + 1) * 4 + 15) / 16 * 16;
    static constexpr int TMEM = MBAR + 8 * 8;
    static constexpr int PART = TMEM + 16;                            // layer-3 partial sums of the upper column half: 128 x float4
    static constexpr int TOTAL = PART + TC_TM * 16;
};
static_assert(FuLayout::TOTAL <= 227 * 1024, "fused kernel exceeds the shared memory of one SM");

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// one lane's share of a tap is kept as loaded (4 registers): 8 bf16 channels or 4 fp32 channels
template <bool BF16>
__device__ __forceinline__ float tap_val(const uint4& q, int ch) {
    if constexpr (BF16) {
        const uint32_t w = (ch >> 1) == 0 ? q.x : (ch >> 1) == 1 ? q.y : (ch >> 1) == 2 ? q.z : q.w;
        return (ch & 1) ? bf_hi(w) : bf_lo(w);
    } else {
        return __uint_as_float(ch == 0 ? q.x : ch == 1 ? q.y : ch == 2 ? q.z : q.w);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GATHER group: the warp's 16 samples (rows row0 .. row0+15 of the tile) -> V rows + sigma features + hemisphere flags
// ---------------------------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ void fused_gather_rows(const EgnKernelCfg& k, const void* __restrict__ tab, const float* __restrict__ knots,
                                                  const float* __restrict__ rays, const float* __restrict__ zs, long long M,
                                                  long long m_base, int row0, int lane, unsigned char* vbuf,
                                                  unsigned char* yang_out, float* __restrict__ fsig) {
    constexpr int NCH = BF16 ? 8 : 4;              // channels per lane
    constexpr int LPS = EGN_CF / NCH;              // lanes per sample: 8 / 16
    constexpr int SPI = 32 / LPS;                  // samples per iteration: 4 / 2
    constexpr int DL = EGN_CS / NCH;               // density lanes per sample: 2 / 4
    // ---- coordinates of sample (lane & 15) ----
    const long long m = m_base + row0 + (lane & 15);
    YYCoord cc;
    cc.c[0] = cc.c[1] = cc.c[2] = -3.f;            // out of range -> every tap reads zero
    cc.yang = 0;
    if (m < M) {
        const long long ray = m / k.S;
        const float z = zs[m];
        const float* ry = rays + ray * 6;
        cc = egn_cart_to_yinyang(ry[0] + ry[3] * z, ry[1] + ry[4] * z, ry[2] + ry[5] * z, k, knots);
    }
    if (lane < 16) yang_out[row0 + lane] = (unsigned char)cc.yang;
    const int sub = lane % LPS;
    float myf = 0.f;
#pragma unroll 1
    for (int it = 0; it < 16 / SPI; ++it) {
        const int src = SPI * it + lane / LPS;
        float c[3];
        c[0] = __shfl_sync(FULL, cc.c[0], src);
        c[1] = __shfl_sync(FULL, cc.c[1], src);
        c[2] = __shfl_sync(FULL, cc.c[2], src);
        const int yang = __shfl_sync(FULL, cc.yang, src);
        // per axis: clamped texel indices and tap weights with the zero padding of F.grid_sample folded in — an
        // out-of-range tap gets weight 0 and reads a clamped in-range texel, so every load is unconditional
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
        constexpr unsigned ES = BF16 ? 2u : 4u;                 // bytes per table element
        const char* tabc = reinterpret_cast<const char*>(tab);
        uint4 t[3][4], l[3][2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int ax = egn_mx(i), ay = egn_my(i), al = egn_vl(i);
            const unsigned W = (unsigned)k.lay.G[ax];
            const unsigned pbase = (unsigned)k.lay.pf[yang][i] + sub * NCH, lbase = (unsigned)k.lay.lf[yang][i] + sub * NCH;
            const unsigned ra = j0[ay] * W, rb = j1[ay] * W;
            t[i][0] = __ldg(reinterpret_cast<const uint4*>(tabc + (size_t)(pbase + (ra + j0[ax]) * EGN_CF) * ES));
            t[i][1] = __ldg(reinterpret_cast<const uint4*>(tabc + (size_t)(pbase + (ra + j1[ax]) * EGN_CF) * ES));
            t[i][2] = __ldg(reinterpret_cast<const uint4*>(tabc + (size_t)(pbase + (rb + j0[ax]) * EGN_CF) * ES));
            t[i][3] = __ldg(reinterpret_cast<const uint4*>(tabc + (size_t)(pbase + (rb + j1[ax]) * EGN_CF) * ES));
            l[i][0] = __ldg(reinterpret_cast<const uint4*>(tabc + (size_t)(lbase + j0[al] * EGN_CF) * ES));
            l[i][1] = __ldg(reinterpret_cast<const uint4*>(tabc + (size_t)(lbase + j1[al] * EGN_CF) * ES));
        }
        float f = 0.f;
        const int row = row0 + src;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int ax = egn_mx(i), ay = egn_my(i), al = egn_vl(i);
            const float w0 = wa0[ax] * wa0[ay], w1 = wa1[ax] * wa0[ay], w2 = wa0[ax] * wa1[ay], w3 = wa1[ax] * wa1[ay];
            const float u0 = wa0[al], u1 = wa1[al];
            float prod[NCH];
            float s = 0.f;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                float P = w0 * tap_val<BF16>(t[i][0], ch);
                P = fmaf(w1, tap_val<BF16>(t[i][1], ch), P);
                P = fmaf(w2, tap_val<BF16>(t[i][2], ch), P);
                P = fmaf(w3, tap_val<BF16>(t[i][3], ch), P);
                const float Lv = fmaf(u1, tap_val<BF16>(l[i][1], ch), u0 * tap_val<BF16>(l[i][0], ch));
                prod[ch] = P * Lv;
                s += prod[ch];
            }
            // density: sum over the 16 density channels = DL lanes, then relu (EgoNeRF.py:346)
#pragma unroll
            for (int d = 1; d < DL; d <<= 1) s += __shfl_xor_sync(FULL, s, d);
            f += fmaxf(s, 0.f);
            if (sub >= DL) {                                   // appearance lanes: bf16 row of the A operand
                const int kk = i * EGN_CA + (sub - DL) * NCH;  // first K index of this lane's products
                unsigned char* dst = vbuf + (kk >> 3) * FU_VCHUNK + row * 16 + (kk & 7) * 2;
                if constexpr (BF16) {
                    uint4 q;
                    q.x = pack_hi(prod[0], prod[1]); q.y = pack_hi(prod[2], prod[3]);
                    q.z = pack_hi(prod[4], prod[5]); q.w = pack_hi(prod[6], prod[7]);
                    *reinterpret_cast<uint4*>(dst) = q;
                } else {
                    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_hi(prod[0], prod[1]), pack_hi(prod[2], prod[3]));
                }
            }
        }
        // the density lanes of sample `src` all hold f; hand it to the lane that owns the sample's coordinate slot
#pragma unroll
        for (int q = 0; q < SPI; ++q) {
            const float fq2 = __shfl_sync(FULL, f, q * LPS);
            if ((lane & 15) == SPI * it + q) myf = fq2;
        }
    }
    if (lane < 16 && m < M) fsig[m] = myf;
}

// ---------------------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(FU_THREADS, 1)
egn_fused_fine_kernel(const __grid_constant__ EgnKernelCfg k, const void* __restrict__ tab, const float* __restrict__ basis0,
                      const float* __restrict__ basis1, const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                      const float* __restrict__ b3, const float* __restrict__ rays, long long M,
                      const float* __restrict__ zs, float* __restrict__ fsig, float* __restrict__ feat_out,
                      float* __restrict__ rgbs) {
    using L = FuLayout;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int AD = k.app_dim;
    const int in_dim = 5 * AD + 15;
    unsigned char* w1s = smem + L::W1;
    unsigned char* w2s = smem + L::W2;
    unsigned char* bbs = smem + L::BB;
    unsigned char* as = smem + L::A;
    unsigned char* vs = smem + L::V;
    unsigned char* s_yang = smem + L::YANG;
    float* s_knots = reinterpret_cast<float*>(smem + L::KNOTS);
    float* part = reinterpret_cast<float*>(smem + L::PART);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::TMEM);
    const uint32_t bar = smem_u32(smem + L::MBAR);
    const uint32_t v_full0 = bar, v_empty0 = bar + 16, feat_full = bar + 32, d1_full = bar + 40, d2_full = bar + 48;

    // ---- one-time setup ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(v_full0, 8); mbar_init(v_full0 + 8, 8);          // one arrive per gather warp
        mbar_init(v_empty0, 1); mbar_init(v_empty0 + 8, 1);        // tcgen05.commit
        mbar_init(feat_full, 1); mbar_init(d1_full, 1); mbar_init(d2_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < EGN_HID * TC_K1; i += FU_THREADS) {
        const int n = i / TC_K1, kk = i % TC_K1;
        const int src = tc_input_index(kk, AD);
        store_elem(w1s, nullptr, false, n, kk, src >= 0 ? w1[n * in_dim + src] : (src == -2 ? b1[n] : 0.f));
    }
    for (int i = tid; i < EGN_HID * EGN_HID; i += FU_THREADS) store_elem(w2s, nullptr, false, i / EGN_HID, i % EGN_HID, w2[i]);
    for (int i = tid; i < 64 * FU_VK; i += FU_THREADS) {            // rows 0..31: basis_mat_yin, 32..63: basis_mat_yang
        const int n = i / FU_VK, kk = i % FU_VK, o = n & 31;
        const float* B = (n >> 5) ? basis1 : basis0;
        store_elem(bbs, nullptr, false, n, kk, o < AD ? B[o * FU_VK + kk] : 0.f, FU_BB_CHUNK);
    }
    for (int i = tid; i <= k.knots_last; i += FU_THREADS) s_knots[i] = k.r_knots[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const long long tiles = (M + TC_TM - 1) / TC_TM;
    bool ok = true;

    if (warp >= 8) {
        // =========================== GATHER group ===========================
        const int gwarp = warp - 8;
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1, u = it >> 1;
            ok &= mbar_wait(v_empty0 + 8 * b, (u & 1) ^ 1);          // layer-0 MMAs of the tile that used this buffer are done
            fused_gather_rows<BF16>(k, tab, s_knots, rays, zs, M, tile * TC_TM, 16 * gwarp, lane, vs + b * FU_VBYTES,
                                    s_yang + (it & 3) * TC_TM, fsig);
            fence_async_smem();                                      // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(v_full0 + 8 * b);
        }
    } else {
        // =========================== MLP group ===========================
        const int row = tid & 127, half = tid >> 7;
        const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t a_s = smem_u32(as), w1_s = smem_u32(w1s), w2_s = smem_u32(w2s), bb_s = smem_u32(bbs), v_s = smem_u32(vs);
        const float bias3[3] = {__ldg(b3), __ldg(b3 + 1), __ldg(b3 + 2)};
        // Software pipeline over tiles.  The tensor pipe works on layer 0 of tile t+1 while this group waits for layer 2 of
        // tile t, and on layer 1 of tile t while the group runs layer 3 of tile t-1 out of TMEM: only the layer-2 wait is
        // exposed.  TMEM: D1 = columns 0..127, D2 = 128..255, feat2 = 256..319.
        auto issue_layer0 = [&](uint32_t i) {                   // thread 0: feat2 = V . [B_yin | B_yang]^T for CTA-local tile i
            const uint32_t bb = i & 1;
            ok &= mbar_wait(v_full0 + 8 * bb, (i >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < FU_VK / 16; ++ks)
                tc_mma(tmem + 256, tc_desc(v_s + bb * FU_VBYTES + ks * 2 * FU_VCHUNK, FU_VCHUNK),
                       tc_desc(bb_s + ks * 2 * FU_BB_CHUNK, FU_BB_CHUNK), TC_IDESC_128x64, ks > 0);
            tc_commit(v_empty0 + 8 * bb);
            tc_commit(feat_full);
        };
        // layer 3 + sigmoid of the tile whose D2 sits in TMEM (its layer-2 completion has been waited for)
        auto layer3 = [&](long long gm3) {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col = 64 * half + 32 * cc;
                uint32_t r[32];
                tmem_ld32(tmem_lane + 128 + col, r);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + col) + q);
                    const float4 wa = __ldg(reinterpret_cast<const float4*>(w3 + col) + q);
                    const float4 wb = __ldg(reinterpret_cast<const float4*>(w3 + EGN_HID + col) + q);
                    const float4 wc = __ldg(reinterpret_cast<const float4*>(w3 + 2 * EGN_HID + col) + q);
                    const float h0 = fmaxf(__uint_as_float(r[4 * q]) + bb.x, 0.f), h1 = fmaxf(__uint_as_float(r[4 * q + 1]) + bb.y, 0.f);
                    const float h2 = fmaxf(__uint_as_float(r[4 * q + 2]) + bb.z, 0.f), h3 = fmaxf(__uint_as_float(r[4 * q + 3]) + bb.w, 0.f);
                    p0 = fmaf(h0, wa.x, p0); p0 = fmaf(h1, wa.y, p0); p0 = fmaf(h2, wa.z, p0); p0 = fmaf(h3, wa.w, p0);
                    p1 = fmaf(h0, wb.x, p1); p1 = fmaf(h1, wb.y, p1); p1 = fmaf(h2, wb.z, p1); p1 = fmaf(h3, wb.w, p1);
                    p2 = fmaf(h0, wc.x, p2); p2 = fmaf(h1, wc.y, p2); p2 = fmaf(h2, wc.z, p2); p2 = fmaf(h3, wc.w, p2);
                }
            }
            tc_fence_before();
            if (half == 1) *reinterpret_cast<float4*>(part + row * 4) = make_float4(p0, p1, p2, 0.f);
            named_bar_sync(1, FU_GROUP);
            if (half == 0 && gm3 < M) {
                const float4 q = *reinterpret_cast<const float4*>(part + row * 4);
                rgbs[gm3 * 3 + 0] = egn_sigmoid(p0 + q.x + bias3[0]);
                rgbs[gm3 * 3 + 1] = egn_sigmoid(p1 + q.y + bias3[1]);
                rgbs[gm3 * 3 + 2] = egn_sigmoid(p2 + q.z + bias3[2]);
            }
        };
        if (tid == 0 && (long long)blockIdx.x < tiles) issue_layer0(0);
        uint32_t it = 0;
        long long gm_prev = M;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const long long gm = tile * TC_TM + row;
            const bool live = gm < M;
            // ---- layer 0 of this tile was issued one iteration ago ----
            ok &= mbar_wait(feat_full, it & 1);
            tc_fence_after();
            // ---- A. this thread's 16 elements: features of its hemisphere (from TMEM), then view direction / 1 / padding ----
            {
                const int yang = s_yang[(it & 3) * TC_TM + row];
                uint32_t r0[16], r1[16];
                tmem_ld16(tmem_lane + 256 + 16 * half, r0);            // yin block
                tmem_ld16(tmem_lane + 256 + 32 + 16 * half, r1);       // yang block
                float el[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) el[j] = __uint_as_float(yang ? r1[j] : r0[j]);
                if (feat_out != nullptr && live) {                     // saved for the backward pass (28 floats / sample)
                    float4* dst = reinterpret_cast<float4*>(feat_out + gm * EGN_FEAT_STRIDE + 16 * half);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (half == 0 || q < 3) dst[q] = make_float4(el[4 * q], el[4 * q + 1], el[4 * q + 2], el[4 * q + 3]);
                }
                const float* dir = rays + (live ? gm / k.S : 0) * 6 + 3;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int e = 16 * half + j;
                    if (e >= AD) el[j] = (e < AD + 3) ? __ldg(dir + (e - AD)) : (e == AD + 3 ? 1.f : 0.f);
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
                    for (int c = 0; c < 5; ++c) store_chunk<false>(as, nullptr, 10 * half + 5 * pass + c, row, v + 8 * c);
                }
            }
            fence_async_smem();
            tc_fence_before();
            named_bar_sync(1, FU_GROUP);
            // ---- B. layer 1 (runs on the tensor pipe during layer 3 of the previous tile) ----
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < TC_K1 / 16; ++ks)
                    tc_mma(tmem, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc(w1_s + ks * 2 * TC_CHUNK), TC_IDESC_128x128, ks > 0);
                tc_commit(d1_full);
            }
            if (it > 0) layer3(gm_prev);
            ok &= mbar_wait(d1_full, it & 1);
            tc_fence_after();
            // ---- C. H1 = relu(D1) -> operand of layer 2 ----
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col = 64 * half + 32 * cc;
                uint32_t r[32];
                tmem_ld32(tmem_lane + col, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]), 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk<false>(as, nullptr, (col >> 3) + c, row, v + 8 * c);
            }
            fence_async_smem();
            tc_fence_before();
            named_bar_sync(1, FU_GROUP);
            // ---- D. layer 2, then layer 0 of the next tile ----
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < EGN_HID / 16; ++ks)
                    tc_mma(tmem + 128, tc_desc(a_s + ks * 2 * TC_CHUNK), tc_desc(w2_s + ks * 2 * TC_CHUNK), TC_IDESC_128x128, ks > 0);
                tc_commit(d2_full);
                if (tile + gridDim.x < tiles) issue_layer0(it + 1);
            }
            ok &= mbar_wait(d2_full, it & 1);
            tc_fence_after();
            gm_prev = gm;
        }
        if (it > 0) layer3(gm_prev);                            // layer 3 of the last tile
    }
    if (!ok) __trap();                                          // a lost mbarrier arrive: fail loudly, never hang
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

int egn_launch_fused_fine(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                          float* fsig, float* feat_out, float* rgbs, cudaStream_t st) {
    const long long M = n * k.S;
    const long long tiles = (M + TC_TM - 1) / TC_TM;
    const int blocks = (int)(tiles < 148 ? tiles : 148);
    if (k.tables_bf16 != nullptr) {
        cudaFuncSetAttribute(egn_fused_fine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuLayout::TOTAL);
        egn_fused_fine_kernel<true><<<blocks, FU_THREADS, FuLayout::TOTAL, st>>>(
            k, k.tables_bf16, p->basis[0], p->basis[1], p->mlp_w[0], p->mlp_b[0], p->mlp_w[1], p->mlp_b[1], p->mlp_w[2],
            p->mlp_b[2], rays, M, z, fsig, feat_out, rgbs);
    } else {
        cudaFuncSetAttribute(egn_fused_fine_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuLayout::TOTAL);
        egn_fused_fine_kernel<false><<<blocks, FU_THREADS, FuLayout::TOTAL, st>>>(
            k, k.tables, p->basis[0], p->basis[1], p->mlp_w[0], p->mlp_b[0], p->mlp_w[1], p->mlp_b[1], p->mlp_w[2],
            p->mlp_b[2], rays, M, z, fsig, feat_out, rgbs);
    }
    return (int)cudaGetLastError();
}
