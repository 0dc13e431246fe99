// Backward of the colour-decode MLP on the tensor cores (tcgen05 + TMEM), fp16 operands / fp32 accumulation — the
// throughput-mode counterpart of egn_mlp_bwd.cu (which stays the exact-fp32 reference implementation).
// Activations and weights are fp16 exactly as in the fused forward (same recomputed H1 / H2, same relu masks); the
// gradient operands dO, dZ2, dZ1 are fp16 after multiplication by the launch-wide power of two S of tc_grad_scale()
// (egn_tc.cuh), every fp32 result is multiplied by 1/S.  Against bf16 operands this cuts the rounding error of every
// operand 8x (measured per-tensor error vs the reference's gradients: profiles/r02_parity.md).
//
// Per 128-sample tile (one CTA per SM, 512 threads = 4 threads per row, 32 / 40 columns each):
//   recompute   X -> D1 = X W1^T -> H1 = relu -> D2 = H1 W2^T (+b2 through a constant-1 column) -> H2 = relu
//   dO  = d_rgbs * c (1 - c)                       (sigmoid', c = sample colour saved by the forward)
//   dZ2 = (dO W3) [H2 > 0]                         CUDA cores (K = 3)
//   dW3^T += H2^T dO                               tcgen05, A = H2 tile viewed MN-major, N = 16 (3 used)
//   G1  = dZ2 W2        dW2 += dZ2^T [H1 | 1]      tcgen05; the transposed operands are MN-major VIEWS of the very same
//   dZ1 = G1 [H1 > 0]                              shared-memory tiles (canonical K-major tiles read with LBO/SBO swapped:
//   dX  = dZ1 W1        dW1 += dZ1^T X              LBO = 128 B, SBO = 2048 B) — no transposed copies, no extra smem
//   d_feat_e = dX_e + cos(x) dX_sin - sin(x) dX_cos + 2 (cos(2x) dX_sin2 - sin(2x) dX_cos2)   (PE chain rule)
// dW1 / dW2 / dW3^T live in TMEM for the whole kernel (persistent accumulators, 320 of 512 columns) and are flushed with
// atomics once per CTA; db1 / db2 are the columns of dW1 / dW2 that face the constant-1 inputs.
#include "egn_tc.cuh"
#include "egn_host.h"

#define BT_THREADS 512
#define BT_K2 144                        // hidden width + constant-1 column, padded to a multiple of 16
#define IDESC_F      TC_IDESC_F16(0x08200490u)         // M128 N128, A K-major, B K-major
#define IDESC_B3W    TC_IDESC_F16(0x08048490u)         // M128 N16,  A MN-major, B K-major
#define IDESC_B2X    TC_IDESC_F16(0x08210490u)         // M128 N128, A K-major,  B MN-major
#define IDESC_B2W    TC_IDESC_F16(0x08258490u)         // M128 N144, A MN-major, B MN-major
#define IDESC_B1X    TC_IDESC_F16(0x08290490u)         // M128 N160, A K-major,  B MN-major
#define IDESC_B1W    TC_IDESC_F16(0x08298490u)         // M128 N160, A MN-major, B MN-major

struct BtLayout {
    static constexpr int W1 = 0;                                        // [128 n][160 k]  40 960
    static constexpr int W2 = W1 + (TC_K1 / 8) * TC_CHUNK;              // [128 n][144 k]  36 864
    static constexpr int X = W2 + (BT_K2 / 8) * TC_CHUNK;               // [128 m][160]    40 960
    static constexpr int H1 = X + (TC_K1 / 8) * TC_CHUNK;               // [128 m][144]    36 864
    static constexpr int H2 = H1 + (BT_K2 / 8) * TC_CHUNK;              // [128 m][128]    32 768
    static constexpr int DZ = H2 + (EGN_HID / 8) * TC_CHUNK;            // dZ2, then dZ1   32 768
    static constexpr int DO = DZ + (EGN_HID / 8) * TC_CHUNK;            // [16 ch][128 m]   4 096 (K chunk = 256 B)
    static constexpr int MBAR = DO + 16 * 256;
    static constexpr int TMEM = MBAR + 4 * 8;
    static constexpr int TOTAL = TMEM + 16;
};
static_assert(BtLayout::TOTAL <= 227 * 1024, "MLP backward kernel exceeds the shared memory of one SM");

// TMEM columns
#define BT_WORK 0        // D1 / D2 / G1 (128) / dX (160)
#define BT_DW2 160       // 144
#define BT_DW1 304       // 160
#define BT_DW3 464       // 16

__global__ void __launch_bounds__(BT_THREADS, 1)
egn_mlp_bwd_tc_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                      const float* __restrict__ rays, long long M, const float* __restrict__ feat,
                      const float* __restrict__ rgbs, const float* __restrict__ d_rgbs, const unsigned* __restrict__ gmax_bits,
                      float* __restrict__ d_feat,
                      float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2,
                      float* __restrict__ dW3, float* __restrict__ db3) {
    using L = BtLayout;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // shuffle: provably warp-uniform
    const int row = tid & 127, q = tid >> 7;
    const int AD = k.app_dim;
    const int in_dim = 5 * AD + 15;
    unsigned char* w1s = smem + L::W1; unsigned char* w2s = smem + L::W2;
    unsigned char* xs = smem + L::X;   unsigned char* h1s = smem + L::H1;
    unsigned char* h2s = smem + L::H2; unsigned char* dzs = smem + L::DZ;
    unsigned char* dos = smem + L::DO;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::TMEM);
    const uint32_t bar = smem_u32(smem + L::MBAR);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < 4; ++i) mbar_init(bar + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < EGN_HID * TC_K1; i += BT_THREADS) {
        const int n = i / TC_K1, kk = i % TC_K1;
        const int src = tc_input_index(kk, AD);
        store_elem_h(w1s, n, kk, src >= 0 ? w1[n * in_dim + src] : (src == -2 ? b1[n] : 0.f));
    }
    for (int i = tid; i < EGN_HID * BT_K2; i += BT_THREADS) {
        const int n = i / BT_K2, kk = i % BT_K2;
        store_elem_h(w2s, n, kk, kk < EGN_HID ? w2[n * EGN_HID + kk] : (kk == EGN_HID ? b2[n] : 0.f));
    }
    // constant-1 column of H1 (k = 128) and zero padding (k = 129..143); zero dO tile (channels 3..15 stay zero)
    for (int i = tid; i < TC_TM * 16; i += BT_THREADS) store_elem_h(h1s, i / 16, EGN_HID + i % 16, (i % 16) == 0 ? 1.f : 0.f);
    for (int i = tid; i < 16 * 256 / 4; i += BT_THREADS) reinterpret_cast<uint32_t*>(dos)[i] = 0u;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t w1_s = smem_u32(w1s), w2_s = smem_u32(w2s), x_s = smem_u32(xs), h1_s = smem_u32(h1s), h2_s = smem_u32(h2s),
                   dz_s = smem_u32(dzs), do_s = smem_u32(dos);
    const float4* w3v = reinterpret_cast<const float4*>(w3);
    float inv_scale;
    const float scale = tc_grad_scale(gmax_bits, inv_scale);

    const long long tiles = (M + TC_TM - 1) / TC_TM;
    uint32_t it = 0;
    bool ok = true;
    float acc3[3] = {0.f, 0.f, 0.f};
    // this thread's 8 input elements of a tile (features, then view direction / constant 1 / padding), fetched one tile
    // ahead so that the global-memory latency overlaps the last MMA batch of the previous tile
    auto load_elements = [&](long long t, float (&e8)[8]) {
        const long long g_m = t * TC_TM + row;
        const bool lv = t < tiles && g_m < M;
#pragma unroll
        for (int j = 0; j < 8; ++j) e8[j] = 0.f;
        if (lv) {
            const float4* f4 = reinterpret_cast<const float4*>(feat + g_m * EGN_FEAT_STRIDE) + 2 * q;
            const float4 a = __ldg(f4);
            e8[0] = a.x; e8[1] = a.y; e8[2] = a.z; e8[3] = a.w;
            if (q < 3) { const float4 b = __ldg(f4 + 1); e8[4] = b.x; e8[5] = b.y; e8[6] = b.z; e8[7] = b.w; }
        }
        const float* dir = rays + (lv ? egn_ray_of(g_m, k.S) : 0) * 6 + 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = 8 * q + j;
            if (e >= AD) e8[j] = (e < AD + 3) ? (lv ? __ldg(dir + (e - AD)) : 0.f) : (e == AD + 3 ? 1.f : 0.f);
        }
    };
    float el_next[8];
    load_elements(blockIdx.x, el_next);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const long long gm = tile * TC_TM + row;
        const bool live = gm < M;
        const uint32_t par = it & 1;
        // ---- P1. X rows: 8 elements per thread -> 5 chunks (the elements were loaded one tile ahead) ----
        float el[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) el[j] = el_next[j];
        {
            float v[40];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s1, c1;
                __sincosf(el[j], &s1, &c1);
                v[5 * j] = el[j]; v[5 * j + 1] = s1; v[5 * j + 2] = c1;
                v[5 * j + 3] = 2.f * s1 * c1; v[5 * j + 4] = 1.f - 2.f * s1 * s1;
            }
#pragma unroll
            for (int c = 0; c < 5; ++c) store_chunk_h(xs, 5 * q + c, row, v + 8 * c);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < TC_K1 / 16; ++ks)
                tc_mma(tmem + BT_WORK, desc_k(x_s + ks * 2 * TC_CHUNK), desc_k(w1_s + ks * 2 * TC_CHUNK), IDESC_F, ks > 0);
            tc_commit(bar);
        }
        // ---- P2. H1 = relu(D1) (b1 rides in W1), keep the relu mask ----
        ok &= mbar_wait(bar, par);
        tc_fence_after();
        uint32_t m1 = 0;
        {
            uint32_t r[32];
            tmem_ld32(tmem_lane + BT_WORK + 32 * q, r);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float h = __uint_as_float(r[j]);
                m1 |= (h > 0.f ? 1u : 0u) << j;
                v[j] = fmaxf(h, 0.f);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) store_chunk_h(h1s, 4 * q + c, row, v + 8 * c);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < BT_K2 / 16; ++ks)
                tc_mma(tmem + BT_WORK, desc_k(h1_s + ks * 2 * TC_CHUNK), desc_k(w2_s + ks * 2 * TC_CHUNK), IDESC_F, ks > 0);
            tc_commit(bar + 8);
        }
        // ---- P3. H2 = relu(D2) (b2 rides in W2); dO; dZ2 = (dO W3) [H2 > 0] ----
        float dq[3] = {0.f, 0.f, 0.f};                     // loaded while the layer-2 MMAs run
        if (live) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float c = __ldg(rgbs + gm * 3 + ch);
                dq[ch] = __ldg(d_rgbs + gm * 3 + ch) * c * (1.f - c);
            }
        }
        ok &= mbar_wait(bar + 8, par);
        tc_fence_after();
        {
            if (q == 0) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    acc3[ch] += dq[ch];
                    *reinterpret_cast<__half*>(dos + (row >> 3) * 256 + ch * 16 + (row & 7) * 2) = __float2half_rn(dq[ch] * scale);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) dq[ch] *= scale;
            uint32_t r[32];
            tmem_ld32(tmem_lane + BT_WORK + 32 * q, r);
            float h2[32], dz[32];
#pragma unroll
            for (int g = 0; g < 8; ++g) {                      // warp-uniform addresses
                const float4 wa = __ldg(w3v + 8 * q + g), wb = __ldg(w3v + 32 + 8 * q + g), wc = __ldg(w3v + 64 + 8 * q + g);
                const float wva[4] = {wa.x, wa.y, wa.z, wa.w}, wvb[4] = {wb.x, wb.y, wb.z, wb.w}, wvc[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float h = __uint_as_float(r[4 * g + j]);
                    h2[4 * g + j] = fmaxf(h, 0.f);
                    const float d = fmaf(dq[2], wvc[j], fmaf(dq[1], wvb[j], dq[0] * wva[j]));
                    dz[4 * g + j] = h > 0.f ? d : 0.f;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                store_chunk_h(h2s, 4 * q + c, row, h2 + 8 * c);
                store_chunk_h(dzs, 4 * q + c, row, dz + 8 * c);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t first = it > 0 ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < TC_TM / 16; ++ks) {           // reductions over the 128 samples of the tile
                tc_mma(tmem + BT_DW3, desc_mn(h2_s + ks * 256), desc_k(do_s + ks * 2 * 256, 256), IDESC_B3W, first | (ks > 0));
                tc_mma(tmem + BT_DW2, desc_mn(dz_s + ks * 256), desc_mn(h1_s + ks * 256), IDESC_B2W, first | (ks > 0));
            }
#pragma unroll
            for (int ks = 0; ks < EGN_HID / 16; ++ks)           // G1 = dZ2 . W2: reduction over W2's output index = its rows
                tc_mma(tmem + BT_WORK, desc_k(dz_s + ks * 2 * TC_CHUNK), desc_mn(w2_s + ks * 256), IDESC_B2X, ks > 0);
            tc_commit(bar + 16);
        }
        // ---- P4. dZ1 = G1 [H1 > 0] ----
        ok &= mbar_wait(bar + 16, par);
        tc_fence_after();
        {
            uint32_t r[32];
            tmem_ld32(tmem_lane + BT_WORK + 32 * q, r);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((m1 >> j) & 1u) ? __uint_as_float(r[j]) : 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) store_chunk_h(dzs, 4 * q + c, row, v + 8 * c);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t first = it > 0 ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < TC_TM / 16; ++ks)
                tc_mma(tmem + BT_DW1, desc_mn(dz_s + ks * 256), desc_mn(x_s + ks * 256), IDESC_B1W, first | (ks > 0));
#pragma unroll
            for (int ks = 0; ks < EGN_HID / 16; ++ks)
                tc_mma(tmem + BT_WORK, desc_k(dz_s + ks * 2 * TC_CHUNK), desc_mn(w1_s + ks * 256), IDESC_B1X, ks > 0);
            tc_commit(bar + 24);
        }
        load_elements(tile + gridDim.x, el_next);          // in flight while the last MMA batch runs
        // ---- P5. d_feat through the positional encoding ----
        ok &= mbar_wait(bar + 24, par);
        tc_fence_after();
        {
            uint32_t r[32], r2[8];
            tmem_ld32(tmem_lane + BT_WORK + 40 * q, r);
            tmem_ld8(tmem_lane + BT_WORK + 40 * q + 32, r2);
            float g[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float dv[5];
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const int idx = 5 * j + c;
                    dv[c] = __uint_as_float(idx < 32 ? r[idx % 32] : r2[(idx - 32) & 7]);
                }
                float s1, c1;
                __sincosf(el[j], &s1, &c1);
                const float s2 = 2.f * s1 * c1, c2 = 1.f - 2.f * s1 * s1;
                g[j] = (dv[0] + (c1 * dv[1] - s1 * dv[2]) + 2.f * (c2 * dv[3] - s2 * dv[4])) * inv_scale;
                if (8 * q + j >= AD) g[j] = 0.f;
            }
            if (live) {
                float4* dst = reinterpret_cast<float4*>(d_feat + gm * EGN_FEAT_STRIDE) + 2 * q;
                dst[0] = make_float4(g[0], g[1], g[2], g[3]);
                if (q < 3) dst[1] = make_float4(g[4], g[5], g[6], g[7]);
            }
        }
        tc_fence_before();
    }
    if (!ok) __trap();
    // ---- flush the persistent accumulators ----
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    for (int c = q; c < BT_K2 / 16; c += 4) {                     // dW2[k = row][n]; column 128 is db2
        uint32_t r[16];
        tmem_ld16(tmem_lane + BT_DW2 + 16 * c, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = 16 * c + j;
            if (n < EGN_HID) atomicAdd(dW2 + row * EGN_HID + n, __uint_as_float(r[j]) * inv_scale);
            else if (n == EGN_HID) atomicAdd(db2 + row, __uint_as_float(r[j]) * inv_scale);
        }
    }
    for (int c = q; c < TC_K1 / 16; c += 4) {                     // dW1[k = row][our K order]; the constant-1 column is db1
        uint32_t r[16];
        tmem_ld16(tmem_lane + BT_DW1 + 16 * c, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int src = tc_input_index(16 * c + j, AD);
            if (src >= 0) atomicAdd(dW1 + row * in_dim + src, __uint_as_float(r[j]) * inv_scale);
            else if (src == -2) atomicAdd(db1 + row, __uint_as_float(r[j]) * inv_scale);
        }
    }
    if (q == 0) {                                                  // dW3[ch][n = row]
        uint32_t r[16];
        tmem_ld16(tmem_lane + BT_DW3, r);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) atomicAdd(dW3 + ch * EGN_HID + row, __uint_as_float(r[ch]) * inv_scale);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float s = acc3[ch];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (lane == 0) atomicAdd(db3 + ch, s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

int egn_launch_mlp_bwd_tc(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                          const float* rgbs, const float* d_rgbs, const unsigned* gmax_bits, float* d_feat, const EgnGrads* g,
                          cudaStream_t st) {
    const long long M = n * k.S;
    if (M <= 0) return 0;
    const long long tiles = (M + TC_TM - 1) / TC_TM;
    const int blocks = (int)(tiles < 148 ? tiles : 148);
    cudaFuncSetAttribute(egn_mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BtLayout::TOTAL);
    egn_mlp_bwd_tc_kernel<<<blocks, BT_THREADS, BtLayout::TOTAL, st>>>(
        k, p->mlp_w[0], p->mlp_b[0], p->mlp_w[1], p->mlp_b[1], p->mlp_w[2], rays, M, feat, rgbs, d_rgbs, gmax_bits, d_feat,
        g->mlp_w[0], g->mlp_b[0], g->mlp_w[1], g->mlp_b[1], g->mlp_w[2], g->mlp_b[2]);
    return (int)cudaGetLastError();
}
