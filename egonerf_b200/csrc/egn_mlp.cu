// Colour decode: positional encoding + 3-layer MLP (MLPRender_Fea / MLPRender, models/tensorBase.py:54-129,
// positional_encoding :14-19), exact-fp32 FFMA version.  A persistent CTA keeps W1^T, W2^T, W3 and the biases in
// shared memory and walks 128-sample tiles:  PE -> [128 x K1] x [K1 x 128] -> ReLU -> [128 x 128] x [128 x 128]
// -> ReLU -> 128 -> 3 -> sigmoid.  Thread tile 8 x 8 (rows rg + 4r, cols 4cg + c and 32 + 4cg + c of a 32 x 64
// warp tile) so that every shared-memory access is a conflict-free 128-bit load.
#include "egn_device.cuh"
#include "egn_host.h"

#include "egn_mlp.cuh"

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
            if (rgbs && m0 + s < M) rgbs[(m0 + s) * 3 + ch] = egn_sigmoid(a + sm.b3[ch]);
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

// backward helper: recompute the hidden activations of a sub-chunk of rays (h1, h2: M x 128) without touching rgbs
int egn_launch_mlp_save(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                        float* h1, float* h2, cudaStream_t st) {
    const long long M = n * k.S;
    long long tiles = (M + MLP_TM - 1) / MLP_TM;
    int blocks = (int)(tiles < 148 ? tiles : 148);
    cudaFuncSetAttribute(egn_mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpSmem));
    egn_mlp_forward_kernel<<<blocks, MLP_THREADS, sizeof(MlpSmem), st>>>(k, p->mlp_w[0], p->mlp_b[0], p->mlp_w[1], p->mlp_b[1],
                                                                          p->mlp_w[2], p->mlp_b[2], rays, M, feat, nullptr,
                                                                          h1, h2);
    return (int)cudaGetLastError();
}
