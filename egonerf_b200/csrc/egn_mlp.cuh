// Shared pieces of the colour-decode MLP kernels (forward: egn_mlp.cu, backward: egn_mlp_bwd.cu), exact fp32 FFMA.
#pragma once
#include "egn_device.cuh"

#define MLP_TM 128
#define MLP_K1MAX 152             // padded input width (150 for MLP_Fea with fea_pe = view_pe = 2)
#define MLP_AST 156               // activation row stride in floats (156 % 32 = 28: rows rg..rg+3 hit distinct banks)
#define MLP_THREADS 256

struct MlpSmem {
    float W1[MLP_K1MAX][EGN_HID];   // k-major (transposed) renderModule.mlp.0.weight
    float W2[EGN_HID][EGN_HID];     // k-major renderModule.mlp.2.weight
    float W3[3][EGN_HID];
    float b1[EGN_HID], b2[EGN_HID], b3[4];
    float A[MLP_TM][MLP_AST];       // activations of the current tile (input, then h1, then h2)
};

__host__ __device__ inline int egn_mlp_in_dim(int shading, int app_dim, int view_pe, int fea_pe) {
    return app_dim + 3 + 2 * 3 * view_pe + (shading == EGN_SHADE_MLP_FEA ? 2 * fea_pe * app_dim : 0);
}

// Builds the MLP input rows of one tile in shared memory:
// [features, viewdirs, sin(PE(features)), cos(PE(features)), sin(PE(viewdirs)), cos(PE(viewdirs))]
// with PE index j*F + f (tensorBase.py:14-19,68-74).
template <int ROWS = MLP_TM, int AST = MLP_AST>
__device__ __forceinline__ void egn_mlp_build_input(const EgnKernelCfg& k, float (*A)[AST], const float* __restrict__ feat,
                                                    const float* __restrict__ rays, long long m0, long long M, int in_dim,
                                                    int k1p) {
    const int AD = k.app_dim;
    const int F = (k.shading == EGN_SHADE_MLP_FEA) ? k.fea_pe : 0, V = k.view_pe;
    const int nbase = AD + 3;
    const int off_fs = nbase, off_fc = off_fs + AD * F, off_vs = off_fc + AD * F, off_vc = off_vs + 3 * V;
    for (int idx = threadIdx.x; idx < ROWS * nbase; idx += MLP_THREADS) {
        const int s = idx % ROWS, j = idx / ROWS;
        const long long m = m0 + s;
        float x = 0.f;
        if (m < M) x = (j < AD) ? feat[m * EGN_FEAT_STRIDE + j] : rays[(m / k.S) * 6 + 3 + (j - AD)];
        A[s][j] = x;
        const int nf = (j < AD) ? F : V;
        const int os = (j < AD) ? off_fs + j * F : off_vs + (j - AD) * V;
        const int oc = (j < AD) ? off_fc + j * F : off_vc + (j - AD) * V;
        float freq = 1.f;
        for (int f = 0; f < nf; ++f) {
            float sn, cs;
            sincosf(x * freq, &sn, &cs);
            A[s][os + f] = sn;
            A[s][oc + f] = cs;
            freq *= 2.f;
        }
    }
    for (int idx = threadIdx.x; idx < ROWS * (k1p - in_dim); idx += MLP_THREADS)
        A[idx % ROWS][in_dim + idx / ROWS] = 0.f;
}

// acc[r][c] += sum_k A[row(r)][k] * W[k][col(c)]  for k < K (K % 4 == 0)
template <int WSTRIDE>
__device__ __forceinline__ void egn_tile_gemm(float acc[8][8], const float (*A)[MLP_AST], const float* __restrict__ W,
                                              int K, int row0, int col0) {
    for (int k4 = 0; k4 < K; k4 += 4) {
        float4 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(&A[row0 + 4 * r][k4]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w0 = *reinterpret_cast<const float4*>(W + (k4 + kk) * WSTRIDE + col0);
            const float4 w1 = *reinterpret_cast<const float4*>(W + (k4 + kk) * WSTRIDE + col0 + 32);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float x = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
                acc[r][0] = fmaf(x, w0.x, acc[r][0]); acc[r][1] = fmaf(x, w0.y, acc[r][1]);
                acc[r][2] = fmaf(x, w0.z, acc[r][2]); acc[r][3] = fmaf(x, w0.w, acc[r][3]);
                acc[r][4] = fmaf(x, w1.x, acc[r][4]); acc[r][5] = fmaf(x, w1.y, acc[r][5]);
                acc[r][6] = fmaf(x, w1.z, acc[r][6]); acc[r][7] = fmaf(x, w1.w, acc[r][7]);
            }
        }
    }
}

// writes relu(acc + bias) into A[row][col]; optionally mirrors the tile to global memory (training)
__device__ __forceinline__ void egn_tile_store_relu(const float acc[8][8], float (*A)[MLP_AST], const float* __restrict__ bias,
                                                    int row0, int col0, float* __restrict__ gsave, long long m0, long long M) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = row0 + 4 * r;
        float4 v0, v1;
        v0.x = fmaxf(acc[r][0] + bias[col0 + 0], 0.f); v0.y = fmaxf(acc[r][1] + bias[col0 + 1], 0.f);
        v0.z = fmaxf(acc[r][2] + bias[col0 + 2], 0.f); v0.w = fmaxf(acc[r][3] + bias[col0 + 3], 0.f);
        v1.x = fmaxf(acc[r][4] + bias[col0 + 32], 0.f); v1.y = fmaxf(acc[r][5] + bias[col0 + 33], 0.f);
        v1.z = fmaxf(acc[r][6] + bias[col0 + 34], 0.f); v1.w = fmaxf(acc[r][7] + bias[col0 + 35], 0.f);
        *reinterpret_cast<float4*>(&A[row][col0]) = v0;
        *reinterpret_cast<float4*>(&A[row][col0 + 32]) = v1;
        if (gsave && m0 + row < M) {
            *reinterpret_cast<float4*>(gsave + (m0 + row) * EGN_HID + col0) = v0;
            *reinterpret_cast<float4*>(gsave + (m0 + row) * EGN_HID + col0 + 32) = v1;
        }
    }
}

__device__ __forceinline__ void egn_mlp_load_weights(MlpSmem& sm, const float* __restrict__ w1, const float* __restrict__ b1,
                                                     const float* __restrict__ w2, const float* __restrict__ b2,
                                                     const float* __restrict__ w3, const float* __restrict__ b3, int in_dim,
                                                     int k1p) {
    for (int i = threadIdx.x; i < k1p * EGN_HID; i += MLP_THREADS) {
        const int kk = i / EGN_HID, nn = i % EGN_HID;
        sm.W1[kk][nn] = (kk < in_dim) ? w1[nn * in_dim + kk] : 0.f;
    }
    for (int i = threadIdx.x; i < EGN_HID * EGN_HID; i += MLP_THREADS) sm.W2[i / EGN_HID][i % EGN_HID] = w2[(i % EGN_HID) * EGN_HID + i / EGN_HID];
    for (int i = threadIdx.x; i < 3 * EGN_HID; i += MLP_THREADS) sm.W3[i / EGN_HID][i % EGN_HID] = w3[i];
    for (int i = threadIdx.x; i < EGN_HID; i += MLP_THREADS) { sm.b1[i] = b1[i]; sm.b2[i] = b2[i]; }
    if (threadIdx.x < 3) sm.b3[threadIdx.x] = b3[threadIdx.x];
}

