// Colour decode on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same function as egn_mlp.cu (MLPRender_Fea, models/tensorBase.py:54-78 with fea_pe = view_pe = 2): per 128-sample tile
//     X[128 x 160] (bf16)  --tcgen05.mma-->  D1 (TMEM, fp32)  --+b1, relu-->  H1[128 x 128] (bf16, shared memory)
//     --tcgen05.mma-->  D2 (TMEM)  --+b2, relu, . W3 (fp32 FFMA), sigmoid-->  rgb
// This kernel is the stand-alone MLP of the parity mode (EgnConfig.mlp_mode = EGN_MLP_TC_SPLIT): every operand is split
// x = hi + lo (two bf16), D += A_lo B_hi + A_hi B_lo + A_hi B_hi with fp32 accumulation: products carry ~2^-17 relative
// error -- fp32-equivalent for the 1e-4 parity bound.  The plain-bf16 mode (EGN_MLP_TC_BF16) runs the same MLP inside the
// fused fine pass (egn_fused.cu); the SPLIT = false instantiation of this template is no longer launched.
// The tile loop is software-pipelined: layer 3 of tile t-1 runs on the CUDA cores while the tensor pipe works on layer 1 of
// tile t (see `layer3` below).
//
// Operand layout in shared memory: canonical K-major, no swizzle (UMMA "interleave"): 8 x 8 core matrices of 128 B,
// element (row, k) at (k / 8) * 2048 + row * 16 + (k % 8) * 2 for 128-row tiles, i.e. LBO (K step) = 2048 B,
// SBO (8-row step) = 128 B.  A thread that owns a row writes 16-byte chunks; consecutive rows are consecutive 16 B
// -> conflict-free st.shared.v4.  The K order of X is OURS to choose (W1's columns are permuted to match when the
// weights are staged): element e contributes [x, sin x, cos x, sin 2x, cos 2x] at k = 5e .. 5e+4 (e < app_dim: feature,
// then the 3 view-direction components), so that a thread produces whole 16-byte chunks from 8 elements.
//
// 512 threads: thread t owns row (t & 127) and column block (t >> 7) — warps w, w+4, w+8, w+12 share TMEM lane quadrant w & 3.
// One elected thread issues the MMAs; completion comes back through tcgen05.commit -> mbarrier.
#include "egn_tc.cuh"
#include "egn_host.h"


// ---- shared-memory carve-up (bytes) --------------------------------------------------------------------------------
template <bool SPLIT>
struct TcLayout {
    static constexpr int W1 = 0;                                             // [hi | lo] 20 chunks each
    static constexpr int W1_BYTES = (TC_K1 / 8) * TC_CHUNK;
    static constexpr int W2 = W1 + W1_BYTES * (SPLIT ? 2 : 1);               // 16 chunks each
    static constexpr int W2_BYTES = (EGN_HID / 8) * TC_CHUNK;
    static constexpr int A = W2 + W2_BYTES * (SPLIT ? 2 : 1);                // X (20 chunks) / H1 (16 chunks)
    static constexpr int A_BYTES = (TC_K1 / 8) * TC_CHUNK;
    static constexpr int MBAR = A + A_BYTES * (SPLIT ? 2 : 1);               // 2 x 8 bytes
    static constexpr int TMEM = MBAR + 16;                                   // 4 bytes
    static constexpr int TOTAL = TMEM + 16;
    // bf16 mode: 114 720 B -> two CTAs per SM (the second CTA's CUDA-core phases hide the first one's MMAs);
    // split mode: 229 408 B -> one CTA per SM
    // layer-3 partial sums of the column blocks 1.., (NQ-1) x 128 x 4 floats (<= 6 KB): aliases K chunks 16..19 (8 KB) of the
    // A buffer, which the layer-2 operand (16 chunks) never touches
    static constexpr int PART = A + 16 * TC_CHUNK;
};

template <bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? 512 : 256, SPLIT ? 1 : 2)
egn_mlp_tc_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                  const float* __restrict__ b3, const float* __restrict__ rays, long long M,
                  const float* __restrict__ feat, float* __restrict__ rgbs) {
    // NQ threads share a row: thread t owns row (t & 127) and column block (t >> 7).  bf16 mode: 2 per row, 256 threads,
    // two CTAs per SM (the second CTA's CUDA-core phases hide the first one's MMAs); split mode (one CTA per SM because
    // of its 229 KB of operands): 4 per row, 512 threads, so that the serial phases between the MMAs are half as long.
    constexpr int NQ = SPLIT ? 4 : 2;
    constexpr int NT = 128 * NQ;
    constexpr int EPT = 32 / NQ;                   // input elements per thread (16 / 8)
    constexpr int CPT = EGN_HID / NQ;              // hidden columns per thread (64 / 32)
    using L = TcLayout<SPLIT>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // shuffle: provably warp-uniform (keeps descriptors in uniform registers)
    const int row = tid & 127, q = tid >> 7;
    const int AD = k.app_dim;
    const int in_dim = AD + 3 + 4 * AD + 12;
    unsigned char* w1hi = smem + L::W1; unsigned char* w1lo = w1hi + L::W1_BYTES;
    unsigned char* w2hi = smem + L::W2; unsigned char* w2lo = w2hi + L::W2_BYTES;
    unsigned char* ahi = smem + L::A;   unsigned char* alo = ahi + L::A_BYTES;
    float* part = reinterpret_cast<float*>(smem + L::PART);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::TMEM);
    const uint32_t bar0 = smem_u32(smem + L::MBAR), bar1 = bar0 + 8;

    // ---- one-time setup: TMEM, barriers, weights ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(SPLIT ? 512 : 256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < EGN_HID * TC_K1; i += NT) {          // W1[n][kk], columns permuted to our K order
        const int n = i / TC_K1, kk = i % TC_K1;
        const int src = tc_input_index(kk, AD);
        store_elem(w1hi, w1lo, SPLIT, n, kk, src >= 0 ? w1[n * in_dim + src] : (src == -2 ? b1[n] : 0.f));
    }
    for (int i = tid; i < EGN_HID * EGN_HID; i += NT) {
        const int n = i / EGN_HID, kk = i % EGN_HID;
        store_elem(w2hi, w2lo, SPLIT, n, kk, w2[i]);
    }
    const float bias3[3] = {__ldg(b3), __ldg(b3 + 1), __ldg(b3 + 2)};
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);      // this warp's lane quadrant
    const uint32_t a_hi = smem_u32(ahi), a_lo = smem_u32(alo);
    const uint32_t w1_hi = smem_u32(w1hi), w1_lo = smem_u32(w1lo), w2_hi = smem_u32(w2hi), w2_lo = smem_u32(w2lo);

    const long long tiles = (M + TC_TM - 1) / TC_TM;
    uint32_t it = 0;
    bool ok = true;
    // this thread's input elements of a tile: app features [EPT*q, EPT*q + EPT), then the view direction, then the
    // constant 1 that carries b1, then padding.  Loaded one tile ahead (during the layer-2 MMA) so that the global-memory
    // latency is off the critical path of the serial phase chain.
    auto load_elements = [&](long long t, float (&el)[EPT]) {
        const long long g_m = t * TC_TM + row;
#pragma unroll
        for (int j = 0; j < EPT; ++j) el[j] = 0.f;
        if (t < tiles && g_m < M) {
            const float4* f4 = reinterpret_cast<const float4*>(feat + g_m * EGN_FEAT_STRIDE) + q * (EPT / 4);
#pragma unroll
            for (int g = 0; g < EPT / 4; ++g) {
                if (q * EPT + 4 * g < EGN_FEAT_STRIDE) {
                    const float4 v = __ldg(f4 + g);
                    el[4 * g] = v.x; el[4 * g + 1] = v.y; el[4 * g + 2] = v.z; el[4 * g + 3] = v.w;
                }
            }
            const float* dir = rays + egn_ray_of(g_m, k.S) * 6 + 3;
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                const int e = EPT * q + j;
                if (e >= AD) el[j] = (e < AD + 3) ? __ldg(dir + (e - AD)) : (e == AD + 3 ? 1.f : 0.f);
            }
        }
    };
    // layer 3 of a finished tile (its D2 sits in TMEM columns 128..255): H2 = relu(D2 + b2); rgb = sigmoid(W3 H2 + b3) in fp32.
    // The partial sums of the NQ column blocks of a row meet in TMEM columns 256.. of the row's lane (the four threads of a
    // row live in warps w, w+4, w+8, w+12 = the same lane quadrant): no shared memory, so this can run while the tensor pipe
    // reads the operand tiles of the NEXT tile's layer 1.
    auto layer3 = [&](long long gm3) {
        float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int cc = 0; cc < CPT / 32; ++cc) {
            const int col = CPT * q + 32 * cc;
            uint32_t r[32];
            tmem_ld32(tmem_lane + 128 + col, r);
#pragma unroll
            for (int g = 0; g < 8; ++g) {                 // warp-uniform addresses: one L1 transaction per load
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + col) + g);
                const float4 wa = __ldg(reinterpret_cast<const float4*>(w3 + col) + g);
                const float4 wb = __ldg(reinterpret_cast<const float4*>(w3 + EGN_HID + col) + g);
                const float4 wc = __ldg(reinterpret_cast<const float4*>(w3 + 2 * EGN_HID + col) + g);
                const float h0 = fmaxf(__uint_as_float(r[4 * g]) + bb.x, 0.f), h1 = fmaxf(__uint_as_float(r[4 * g + 1]) + bb.y, 0.f);
                const float h2 = fmaxf(__uint_as_float(r[4 * g + 2]) + bb.z, 0.f), h3 = fmaxf(__uint_as_float(r[4 * g + 3]) + bb.w, 0.f);
                p0 = fmaf(h0, wa.x, p0); p0 = fmaf(h1, wa.y, p0); p0 = fmaf(h2, wa.z, p0); p0 = fmaf(h3, wa.w, p0);
                p1 = fmaf(h0, wb.x, p1); p1 = fmaf(h1, wb.y, p1); p1 = fmaf(h2, wb.z, p1); p1 = fmaf(h3, wb.w, p1);
                p2 = fmaf(h0, wc.x, p2); p2 = fmaf(h1, wc.y, p2); p2 = fmaf(h2, wc.z, p2); p2 = fmaf(h3, wc.w, p2);
            }
        }
        if (q > 0) tmem_st4(tmem_lane + 256 + 4 * (q - 1), __float_as_uint(p0), __float_as_uint(p1), __float_as_uint(p2), 0u);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (q == 0) {
            uint32_t t[16];
            tmem_ld16(tmem_lane + 256, t);
            if (gm3 < M) {
#pragma unroll
                for (int j = 0; j < NQ - 1; ++j) {
                    p0 += __uint_as_float(t[4 * j]); p1 += __uint_as_float(t[4 * j + 1]); p2 += __uint_as_float(t[4 * j + 2]);
                }
                rgbs[gm3 * 3 + 0] = egn_sigmoid(p0 + bias3[0]);
                rgbs[gm3 * 3 + 1] = egn_sigmoid(p1 + bias3[1]);
                rgbs[gm3 * 3 + 2] = egn_sigmoid(p2 + bias3[2]);
            }
        }
    };
    float el[EPT];
    load_elements(blockIdx.x, el);
    long long gm_prev = M;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const long long gm = tile * TC_TM + row;
        // ---- A. input rows: EPT elements per thread -> chunks of [x, sin x, cos x, sin 2x, cos 2x] ----
        {
#pragma unroll
            for (int pass = 0; pass < EPT / 8; ++pass) {
                float v[40];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = el[8 * pass + j];
                    float s1, c1;
                    if (SPLIT) sincosf(x, &s1, &c1); else __sincosf(x, &s1, &c1);   // bf16 operands keep 8 bits: MUFU is ample
                    v[5 * j] = x; v[5 * j + 1] = s1; v[5 * j + 2] = c1;
                    v[5 * j + 3] = 2.f * s1 * c1; v[5 * j + 4] = 1.f - 2.f * s1 * s1;   // double angle; W1's padding columns are zero
                }
#pragma unroll
                for (int c = 0; c < 5; ++c) store_chunk<SPLIT>(ahi, alo, (EPT / 8) * 5 * q + 5 * pass + c, row, v + 8 * c);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- B. layer 1: D1 (TMEM columns 0..127) = X . W1^T ----
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < TC_K1 / 16; ++ks) {
                const uint32_t o = ks * 2 * TC_CHUNK;
                if (SPLIT) {
                    tc_mma(tmem, tc_desc(a_lo + o), tc_desc(w1_hi + o), TC_IDESC_128x128, ks > 0);
                    tc_mma(tmem, tc_desc(a_hi + o), tc_desc(w1_lo + o), TC_IDESC_128x128, 1);
                    tc_mma(tmem, tc_desc(a_hi + o), tc_desc(w1_hi + o), TC_IDESC_128x128, 1);
                } else {
                    tc_mma(tmem, tc_desc(a_hi + o), tc_desc(w1_hi + o), TC_IDESC_128x128, ks > 0);
                }
            }
            tc_commit(bar0);
        }
        // ---- E(previous tile): its layer 3 runs on the CUDA cores while the tensor pipe works on this tile's layer 1 ----
        if (it > 0) layer3(gm_prev);
        // ---- C. H1 = relu(D1) (b1 is inside D1) -> bf16 operand of layer 2 (overwrites X: its MMAs have completed) ----
        ok &= mbar_wait(bar0, it & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < CPT / 32; ++cc) {
            const int col = CPT * q + 32 * cc;
            uint32_t r[32];
            tmem_ld32(tmem_lane + col, r);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r[j]), 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) store_chunk<SPLIT>(ahi, alo, (col >> 3) + c, row, v + 8 * c);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- D. layer 2: D2 (TMEM columns 128..255) = H1 . W2^T ----
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < EGN_HID / 16; ++ks) {
                const uint32_t o = ks * 2 * TC_CHUNK;
                if (SPLIT) {
                    tc_mma(tmem + 128, tc_desc(a_lo + o), tc_desc(w2_hi + o), TC_IDESC_128x128, ks > 0);
                    tc_mma(tmem + 128, tc_desc(a_hi + o), tc_desc(w2_lo + o), TC_IDESC_128x128, 1);
                    tc_mma(tmem + 128, tc_desc(a_hi + o), tc_desc(w2_hi + o), TC_IDESC_128x128, 1);
                } else {
                    tc_mma(tmem + 128, tc_desc(a_hi + o), tc_desc(w2_hi + o), TC_IDESC_128x128, ks > 0);
                }
            }
            tc_commit(bar1);
        }
        load_elements(tile + gridDim.x, el);              // next tile's inputs: in flight while layer 2 runs
        ok &= mbar_wait(bar1, it & 1);                    // A is free for the next tile's X, D2 is ready for layer 3
        tc_fence_after();
        gm_prev = gm;
    }
    if (it > 0) layer3(gm_prev);                          // layer 3 of the last tile
    if (!ok) __trap();                                      // a lost tcgen05.commit arrive: fail loudly, never hang
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(SPLIT ? 512 : 256) : "memory");
}

int egn_launch_mlp_tc(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                      float* rgbs, int split, cudaStream_t st) {
    const long long M = n * k.S;
    const long long tiles = (M + TC_TM - 1) / TC_TM;
    // Only the 3-term split instantiation is launched: plain bf16 operands (EGN_MLP_TC_BF16) always take the fused fine pass
    // (egn_fused.cu), which contains the same MLP.
    (void)split;
    const int blocks = (int)(tiles < 148 ? tiles : 148);
    cudaFuncSetAttribute(egn_mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcLayout<true>::TOTAL);
    egn_mlp_tc_kernel<true><<<blocks, 512, TcLayout<true>::TOTAL, st>>>(
        k, p->mlp_w[0], p->mlp_b[0], p->mlp_w[1], p->mlp_b[1], p->mlp_w[2], p->mlp_b[2], rays, M, feat, rgbs);
    return (int)cudaGetLastError();
}
