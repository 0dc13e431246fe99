// Backward of the colour-decode MLP (MLPRender_Fea / MLPRender, models/tensorBase.py:54-129), exact fp32 FFMA.
// For one sub-chunk of rays the hidden activations h1, h2 are recomputed by the forward kernel (nothing of size
// M x 128 is kept between forward and backward), then:
//   egn_mlp_bwd_out_kernel     do = d_rgbs * c (1 - c);  dz2 = (do W3) [h2 > 0];  dW3, db3
//   egn_mlp_bwd_hidden_kernel  dz1 = (dz2 W2) [h1 > 0]
//   egn_mlp_bwd_input_kernel   dx = dz1 W1;  d_feat_j = dx_j + sum_f 2^f (cos(feat_j 2^f) dx_sin[j,f] - sin(feat_j 2^f) dx_cos[j,f])
//   egn_mlp_wgrad_kernel       dW2 += dz2^T h1, db2;  dW1 += dz1^T x, db1   (reduction over samples, per-CTA register tiles)
#include "egn_mlp.cuh"
#include "egn_host.h"

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
egn_mlp_bwd_out_kernel(const float* __restrict__ w3, long long M, const float* __restrict__ rgbs,
                       const float* __restrict__ d_rgbs, const float* __restrict__ h2, float* __restrict__ dz2,
                       float* __restrict__ dW3, float* __restrict__ db3) {
    __shared__ float s_acc[3][EGN_HID];
    __shared__ float s_b[3];
    for (int i = threadIdx.x; i < 3 * EGN_HID; i += blockDim.x) (&s_acc[0][0])[i] = 0.f;
    if (threadIdx.x < 3) s_b[threadIdx.x] = 0.f;
    __syncthreads();
    const int n4 = threadIdx.x & 31, rr = threadIdx.x >> 5;
    const float4 wa = *reinterpret_cast<const float4*>(w3 + 0 * EGN_HID + n4 * 4);
    const float4 wb = *reinterpret_cast<const float4*>(w3 + 1 * EGN_HID + n4 * 4);
    const float4 wc = *reinterpret_cast<const float4*>(w3 + 2 * EGN_HID + n4 * 4);
    float4 acc[3] = {f4zero(), f4zero(), f4zero()};
    float accb[3] = {0.f, 0.f, 0.f};
    for (long long m = (long long)blockIdx.x * 8 + rr; m < M; m += (long long)gridDim.x * 8) {
        float dq[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float c = rgbs[m * 3 + ch];
            dq[ch] = d_rgbs[m * 3 + ch] * c * (1.f - c);          // sigmoid'
            accb[ch] += dq[ch];
        }
        const float4 h = *reinterpret_cast<const float4*>(h2 + m * EGN_HID + n4 * 4);
        float4 d;
        d.x = h.x > 0.f ? fmaf(dq[2], wc.x, fmaf(dq[1], wb.x, dq[0] * wa.x)) : 0.f;
        d.y = h.y > 0.f ? fmaf(dq[2], wc.y, fmaf(dq[1], wb.y, dq[0] * wa.y)) : 0.f;
        d.z = h.z > 0.f ? fmaf(dq[2], wc.z, fmaf(dq[1], wb.z, dq[0] * wa.z)) : 0.f;
        d.w = h.w > 0.f ? fmaf(dq[2], wc.w, fmaf(dq[1], wb.w, dq[0] * wa.w)) : 0.f;
        *reinterpret_cast<float4*>(dz2 + m * EGN_HID + n4 * 4) = d;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) acc[ch] = f4fma(dq[ch], h, acc[ch]);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        atomicAdd(&s_acc[ch][n4 * 4 + 0], acc[ch].x); atomicAdd(&s_acc[ch][n4 * 4 + 1], acc[ch].y);
        atomicAdd(&s_acc[ch][n4 * 4 + 2], acc[ch].z); atomicAdd(&s_acc[ch][n4 * 4 + 3], acc[ch].w);
        if (n4 == 0) atomicAdd(&s_b[ch], accb[ch]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * EGN_HID; i += blockDim.x) atomicAdd(dW3 + i, (&s_acc[0][0])[i]);
    if (threadIdx.x < 3) atomicAdd(db3 + threadIdx.x, s_b[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------------
struct HidSmem {
    float W[EGN_HID][EGN_HID];      // renderModule.mlp.2.weight as stored: (out k, in n) = k-major for dz2 . W2
    float A[MLP_TM][MLP_AST];
};

__global__ void __launch_bounds__(MLP_THREADS, 1)
egn_mlp_bwd_hidden_kernel(const float* __restrict__ w2, long long M, const float* __restrict__ dz2,
                          const float* __restrict__ h1, float* __restrict__ dz1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HidSmem& sm = *reinterpret_cast<HidSmem*>(smem_raw);
    for (int i = threadIdx.x; i < EGN_HID * EGN_HID / 4; i += MLP_THREADS)
        reinterpret_cast<float4*>(&sm.W[0][0])[i] = reinterpret_cast<const float4*>(w2)[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = lane >> 3, cg = lane & 7;
    const int row0 = 32 * (warp >> 1) + rg, col0 = 64 * (warp & 1) + 4 * cg;
    const long long tiles = (M + MLP_TM - 1) / MLP_TM;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long m0 = tile * MLP_TM;
        __syncthreads();
        for (int i = threadIdx.x; i < MLP_TM * (EGN_HID / 4); i += MLP_THREADS) {
            const int s = i / (EGN_HID / 4), q = i % (EGN_HID / 4);
            float4 v = f4zero();
            if (m0 + s < M) v = *reinterpret_cast<const float4*>(dz2 + (m0 + s) * EGN_HID + q * 4);
            *reinterpret_cast<float4*>(&sm.A[s][q * 4]) = v;
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        egn_tile_gemm<EGN_HID>(acc, sm.A, &sm.W[0][0], EGN_HID, row0, col0);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const long long m = m0 + row0 + 4 * r;
            if (m < M) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int col = col0 + 32 * half;
                    const float4 h = *reinterpret_cast<const float4*>(h1 + m * EGN_HID + col);
                    float4 d;
                    d.x = h.x > 0.f ? acc[r][4 * half + 0] : 0.f; d.y = h.y > 0.f ? acc[r][4 * half + 1] : 0.f;
                    d.z = h.z > 0.f ? acc[r][4 * half + 2] : 0.f; d.w = h.w > 0.f ? acc[r][4 * half + 3] : 0.f;
                    *reinterpret_cast<float4*>(dz1 + m * EGN_HID + col) = d;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
#define BWD_NIN 160               // MLP input width padded to 16 column groups x 10
#define BWD_AST 164               // row stride of the 160-wide tiles (164 % 32 = 4: adjacent rows stay conflict-free)

struct InSmem {
    float W[EGN_HID][BWD_NIN];      // renderModule.mlp.0.weight (out k, in i), zero padded to 160 inputs
    float A[MLP_TM][BWD_AST];       // dz1 tile (128 wide), then the dx tile (160 wide)
};

__global__ void __launch_bounds__(MLP_THREADS, 1)
egn_mlp_bwd_input_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ w1, long long M,
                         const float* __restrict__ dz1, const float* __restrict__ feat, float* __restrict__ d_feat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    InSmem& sm = *reinterpret_cast<InSmem*>(smem_raw);
    const int in_dim = egn_mlp_in_dim(k.shading, k.app_dim, k.view_pe, k.fea_pe);
    for (int i = threadIdx.x; i < EGN_HID * BWD_NIN; i += MLP_THREADS) {
        const int kk = i / BWD_NIN, c = i % BWD_NIN;
        sm.W[kk][c] = (c < in_dim) ? w1[kk * in_dim + c] : 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rbase = 16 * warp + (lane >> 4), cg = lane & 15;     // rows rbase + 2r, cols 10cg .. 10cg+9
    const int AD = k.app_dim;
    const int F = (k.shading == EGN_SHADE_MLP_FEA) ? k.fea_pe : 0;
    const int off_fs = AD + 3, off_fc = off_fs + AD * F;
    const long long tiles = (M + MLP_TM - 1) / MLP_TM;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long m0 = tile * MLP_TM;
        __syncthreads();
        for (int i = threadIdx.x; i < MLP_TM * (EGN_HID / 4); i += MLP_THREADS) {
            const int s = i / (EGN_HID / 4), q = i % (EGN_HID / 4);
            float4 v = f4zero();
            if (m0 + s < M) v = *reinterpret_cast<const float4*>(dz1 + (m0 + s) * EGN_HID + q * 4);
            *reinterpret_cast<float4*>(&sm.A[s][q * 4]) = v;
        }
        __syncthreads();
        float acc[8][10];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 10; ++c) acc[r][c] = 0.f;
        for (int k4 = 0; k4 < EGN_HID; k4 += 4) {
            float4 a[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(&sm.A[rbase + 2 * r][k4]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float2 w[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) w[j] = *reinterpret_cast<const float2*>(&sm.W[k4 + kk][10 * cg + 2 * j]);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float x = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        acc[r][2 * j] = fmaf(x, w[j].x, acc[r][2 * j]);
                        acc[r][2 * j + 1] = fmaf(x, w[j].y, acc[r][2 * j + 1]);
                    }
                }
            }
        }
        __syncthreads();                                   // every warp is done reading the dz1 tile
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < 5; ++j)
                *reinterpret_cast<float2*>(&sm.A[rbase + 2 * r][10 * cg + 2 * j]) = make_float2(acc[r][2 * j], acc[r][2 * j + 1]);
        __syncthreads();
        // chain rule through the positional encoding of the features (tensorBase.py:14-19); viewdirs carry no gradient
        for (int idx = threadIdx.x; idx < MLP_TM * EGN_FEAT_STRIDE; idx += MLP_THREADS) {
            const int j = idx % EGN_FEAT_STRIDE, s = idx / EGN_FEAT_STRIDE;
            const long long m = m0 + s;
            if (m >= M) continue;
            float v = 0.f;
            if (j < AD) {
                v = sm.A[s][j];
                if (F > 0) {
                    const float x = feat[m * EGN_FEAT_STRIDE + j];
                    float freq = 1.f;
                    for (int f = 0; f < F; ++f) {
                        float sn, cs;
                        sincosf(x * freq, &sn, &cs);
                        v += freq * (cs * sm.A[s][off_fs + j * F + f] - sn * sm.A[s][off_fc + j * F + f]);
                        freq *= 2.f;
                    }
                }
            }
            d_feat[m * EGN_FEAT_STRIDE + j] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dW[k][n] += sum_m dz[m][k] * B[m][n]   (k < 128; n < N);   db[k] += sum_m dz[m][k]
// MODE 0: B = h1 (N = 128, thread cols {4cg..4cg+3, 64+4cg..}) ; MODE 1: B = MLP input x rebuilt from feat + dirs (N = 160)
#define WG_TM 64
template <int MODE>
struct WgSmem {
    float dz[WG_TM][EGN_HID];
    float B[WG_TM][BWD_AST];
};

template <int MODE>
__global__ void __launch_bounds__(MLP_THREADS, 2)
egn_mlp_wgrad_kernel(const __grid_constant__ EgnKernelCfg k, long long M, const float* __restrict__ dz,
                     const float* __restrict__ hsrc, const float* __restrict__ feat, const float* __restrict__ rays,
                     float* __restrict__ dW, float* __restrict__ db, int ldw, int ncols) {
    constexpr int TN = MODE == 0 ? 8 : 10;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WgSmem<MODE>& sm = *reinterpret_cast<WgSmem<MODE>*>(smem_raw);
    const int kg = threadIdx.x >> 4, cg = threadIdx.x & 15;
    const int in_dim = egn_mlp_in_dim(k.shading, k.app_dim, k.view_pe, k.fea_pe);
    float acc[8][TN];
    float accb[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        accb[r] = 0.f;
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;
    }
    const long long tiles = (M + WG_TM - 1) / WG_TM;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long m0 = tile * WG_TM;
        __syncthreads();
        for (int i = threadIdx.x; i < WG_TM * (EGN_HID / 4); i += MLP_THREADS) {
            const int s = i / (EGN_HID / 4), q = i % (EGN_HID / 4);
            float4 v = f4zero(), h = f4zero();
            if (m0 + s < M) {
                v = *reinterpret_cast<const float4*>(dz + (m0 + s) * EGN_HID + q * 4);
                if (MODE == 0) h = *reinterpret_cast<const float4*>(hsrc + (m0 + s) * EGN_HID + q * 4);
            }
            *reinterpret_cast<float4*>(&sm.dz[s][q * 4]) = v;
            if (MODE == 0) *reinterpret_cast<float4*>(&sm.B[s][q * 4]) = h;
        }
        if (MODE == 1) {
            egn_mlp_build_input<WG_TM, BWD_AST>(k, sm.B, feat, rays, m0, M, in_dim, in_dim);
            for (int idx = threadIdx.x; idx < WG_TM * (BWD_NIN - in_dim); idx += MLP_THREADS)
                sm.B[idx % WG_TM][in_dim + idx / WG_TM] = 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int s = 0; s < WG_TM; ++s) {
            const float4 d0 = *reinterpret_cast<const float4*>(&sm.dz[s][8 * kg]);
            const float4 d1 = *reinterpret_cast<const float4*>(&sm.dz[s][8 * kg + 4]);
            const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            float b[TN];
            if (MODE == 0) {
                const float4 b0 = *reinterpret_cast<const float4*>(&sm.B[s][4 * cg]);
                const float4 b1 = *reinterpret_cast<const float4*>(&sm.B[s][64 + 4 * cg]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN / 2; ++j) {
                    const float2 t = *reinterpret_cast<const float2*>(&sm.B[s][10 * cg + 2 * j]);
                    b[2 * j] = t.x; b[2 * j + 1] = t.y;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                accb[r] += d[r];
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(d[r], b[c], acc[r][c]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int kk = 8 * kg + r;
#pragma unroll
        for (int c = 0; c < TN; ++c) {
            const int n = MODE == 0 ? (c < 4 ? 4 * cg + c : 64 + 4 * cg + (c - 4)) : 10 * cg + c;
            if (n < ncols) atomicAdd(dW + (long long)kk * ldw + n, acc[r][c]);
        }
        if (cg == 0) atomicAdd(db + kk, accb[r]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
int egn_launch_mlp_bwd(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                       const float* rgbs, const float* d_rgbs, float* d_feat, float* h1, float* h2, float* dz1,
                       float* dz2, const EgnGrads* g, cudaStream_t st) {
    const long long M = n * k.S;
    if (M <= 0) return 0;
    int e;
    if ((e = egn_launch_mlp_save(k, p, rays, n, feat, h1, h2, st))) return e;
    {
        long long blocks = (M + 7) / 8;
        if (blocks > 148 * 8) blocks = 148 * 8;
        egn_mlp_bwd_out_kernel<<<(unsigned)blocks, 256, 0, st>>>(p->mlp_w[2], M, rgbs, d_rgbs, h2, dz2, g->mlp_w[2], g->mlp_b[2]);
        if ((e = (int)cudaGetLastError())) return e;
    }
    const long long tiles = (M + MLP_TM - 1) / MLP_TM;
    const int pblocks = (int)(tiles < 148 ? tiles : 148);
    cudaFuncSetAttribute(egn_mlp_bwd_hidden_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HidSmem));
    egn_mlp_bwd_hidden_kernel<<<pblocks, MLP_THREADS, sizeof(HidSmem), st>>>(p->mlp_w[1], M, dz2, h1, dz1);
    if ((e = (int)cudaGetLastError())) return e;
    cudaFuncSetAttribute(egn_mlp_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(InSmem));
    egn_mlp_bwd_input_kernel<<<pblocks, MLP_THREADS, sizeof(InSmem), st>>>(k, p->mlp_w[0], M, dz1, feat, d_feat);
    if ((e = (int)cudaGetLastError())) return e;
    const long long wt = (M + WG_TM - 1) / WG_TM;
    const int wblocks = (int)(wt < 148 * 2 ? wt : 148 * 2);
    const int in_dim = egn_mlp_in_dim(k.shading, k.app_dim, k.view_pe, k.fea_pe);
    cudaFuncSetAttribute(egn_mlp_wgrad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WgSmem<0>));
    egn_mlp_wgrad_kernel<0><<<wblocks, MLP_THREADS, sizeof(WgSmem<0>), st>>>(k, M, dz2, h1, nullptr, nullptr, g->mlp_w[1],
                                                                             g->mlp_b[1], EGN_HID, EGN_HID);
    if ((e = (int)cudaGetLastError())) return e;
    cudaFuncSetAttribute(egn_mlp_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WgSmem<1>));
    egn_mlp_wgrad_kernel<1><<<wblocks, MLP_THREADS, sizeof(WgSmem<1>), st>>>(k, M, dz1, nullptr, feat, rays, g->mlp_w[0],
                                                                             g->mlp_b[0], in_dim, in_dim);
    return (int)cudaGetLastError();
}
