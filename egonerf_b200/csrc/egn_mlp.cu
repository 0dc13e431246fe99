// Colour decode: positional encoding + 3-layer MLP (MLPRender_Fea / MLPRender, models/tensorBase.py:54-129,
// positional_encoding :14-19), exact-fp32 FFMA version.  A persistent CTA keeps W1^T, W2^T, W3 and the biases in
// shared memory and walks 128-sample tiles:  PE -> [128 x K1] x [K1 x 128] -> ReLU -> [128 x 128] x [128 x 128]
// -> ReLU -> 128 -> 3 -> sigmoid.  Thread tile 8 x 8 (rows rg + 4r, cols 4cg + c and 32 + 4cg + c of a 32 x 64
// warp tile) so that every shared-memory access is a conflict-free 128-bit load.
#include "egn_device.cuh"
#include "egn_host.h"

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
__device__ __forceinline__ void egn_mlp_build_input(const EgnKernelCfg& k, float (*A)[MLP_AST], const float* __restrict__ feat,
                                                    const float* __restrict__ rays, long long m0, long long M, int in_dim,
                                                    int k1p) {
    const int AD = k.app_dim;
    const int F = (k.shading == EGN_SHADE_MLP_FEA) ? k.fea_pe : 0, V = k.view_pe;
    const int nbase = AD + 3;
    const int off_fs = nbase, off_fc = off_fs + AD * F, off_vs = off_fc + AD * F, off_vc = off_vs + 3 * V;
    for (int idx = threadIdx.x; idx < MLP_TM * nbase; idx += MLP_THREADS) {
        const int s = idx % MLP_TM, j = idx / MLP_TM;
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
    for (int idx = threadIdx.x; idx < MLP_TM * (k1p - in_dim); idx += MLP_THREADS)
        A[idx % MLP_TM][in_dim + idx / MLP_TM] = 0.f;
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

__global__ void __launch_bounds__(MLP_THREADS, 1)
egn_mlp_forward_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ w1, const float* __restrict__ b1,
                       const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                       const float* __restrict__ b3, const float* __restrict__ rays, long long M,
                       const float* __restrict__ feat, float* __restrict__ rgbs, float* __restrict__ h1_save,
                       float* __restrict__ h2_save) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MlpSmem& sm = *reinterpret_cast<MlpSmem*>(smem_raw);
    const int in_dim = egn_mlp_in_dim(k.shading, k.app_dim, k.view_pe, k.fea_pe);
    const int k1p = (in_dim + 3) & ~3;
    egn_mlp_load_weights(sm, w1, b1, w2, b2, w3, b3, in_dim, k1p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = lane >> 3, cg = lane & 7;
    const int row0 = 32 * (warp >> 1) + rg;           // rows row0 + 4r
    const int col0 = 64 * (warp & 1) + 4 * cg;         // cols col0 + c, col0 + 32 + c
    const long long tiles = (M + MLP_TM - 1) / MLP_TM;
    __syncthreads();
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long m0 = tile * MLP_TM;
        egn_mlp_build_input(k, sm.A, feat, rays, m0, M, in_dim, k1p);
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        egn_tile_gemm<EGN_HID>(acc, sm.A, &sm.W1[0][0], k1p, row0, col0);
        __syncthreads();
        egn_tile_store_relu(acc, sm.A, sm.b1, row0, col0, h1_save, m0, M);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        egn_tile_gemm<EGN_HID>(acc, sm.A, &sm.W2[0][0], EGN_HID, row0, col0);
        __syncthreads();
        egn_tile_store_relu(acc, sm.A, sm.b2, row0, col0, h2_save, m0, M);
        __syncthreads();
        // layer 3 + sigmoid: one (row, channel) pair per thread-iteration
        for (int idx = threadIdx.x; idx < MLP_TM * 3; idx += MLP_THREADS) {
            const int s = idx % MLP_TM, ch = idx / MLP_TM;
            float a = 0.f;
#pragma unroll 8
            for (int k4 = 0; k4 < EGN_HID; k4 += 4) {
                const float4 h = *reinterpret_cast<const float4*>(&sm.A[s][k4]);
                const float4 w = *reinterpret_cast<const float4*>(&sm.W3[ch][k4]);
                a = fmaf(h.x, w.x, a); a = fmaf(h.y, w.y, a); a = fmaf(h.z, w.z, a); a = fmaf(h.w, w.w, a);
            }
            if (m0 + s < M) rgbs[(m0 + s) * 3 + ch] = egn_sigmoid(a + sm.b3[ch]);
        }
        __syncthreads();
    }
}

int egn_launch_mlp(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                   float* rgbs, cudaStream_t st) {
    const long long M = n * k.S;
    long long tiles = (M + MLP_TM - 1) / MLP_TM;
    int blocks = (int)(tiles < 148 ? tiles : 148);
    cudaFuncSetAttribute(egn_mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpSmem));
    egn_mlp_forward_kernel<<<blocks, MLP_THREADS, sizeof(MlpSmem), st>>>(k, p->mlp_w[0], p->mlp_b[0], p->mlp_w[1], p->mlp_b[1],
                                                                          p->mlp_w[2], p->mlp_b[2], rays, M, feat, rgbs,
                                                                          nullptr, nullptr);
    return (int)cudaGetLastError();
}
