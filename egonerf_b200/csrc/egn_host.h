// Internal launch prototypes shared between the translation units of libegn_b200.
#pragma once
#include <cuda_runtime.h>
#include "egn_device.cuh"

int egn_launch_pack(const EgnConfig* cfg, const EgnParams* params, float* tables, cudaStream_t st);
int egn_launch_unpack(const EgnConfig* cfg, const float* d_tables, const EgnGrads* grads, cudaStream_t st);

int egn_launch_coarse(const EgnKernelCfg& k, const float* rays, long long n, int is_train, const float* u_c,
                      const float* u_f, unsigned long long seed, long long ray0, float near_plane, float* z_out,
                      cudaStream_t st);
int egn_launch_gather(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                      float* fsig, float* feat, cudaStream_t st);
int egn_launch_gather_coords(const EgnKernelCfg& k, const EgnParams* p, const float* coords7, long long m, int coarse,
                             float* fsig, float* feat, cudaStream_t st);
int egn_launch_mlp(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                   float* rgbs, cudaStream_t st);
int egn_launch_composite(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                         const float* fsig, const float* feat, const float* rgbs, const EgnOutputs* out, float* wgt,
                         float* bgw, float* rgbpre, cudaStream_t st);
int egn_launch_coords(const EgnKernelCfg& k, const float* xyz, long long m, float* coords7, cudaStream_t st);
int egn_launch_envmap(int env_h, const float* emission, const float* dirs, long long n, float* out, cudaStream_t st);
int egn_launch_envmap_bwd(int env_h, const float* emission, const float* dirs, long long n, const float* d_out,
                          float* d_emission, cudaStream_t st);

// tensor-core colour decode (egn_mlp_tc.cu): split = 1 -> 3-term bf16 split (fp32-equivalent), 0 -> plain bf16
int egn_launch_mlp_tc(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                      float* rgbs, int split, cudaStream_t st);
// fused fine pass (egn_fused.cu): gather + basis + MLP in one warp-specialised tcgen05 kernel (bf16 operands)
int egn_launch_fused_fine(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                          float* fsig, float* feat_out, float* rgbs, const EgnOutputs* composite_out, void* image_buf,
                          cudaStream_t st);
long long egn_fused_image_bytes();
int egn_launch_adam_tables(const EgnConfig* cfg, const EgnGrads* params_out, const float* d_tables, float* m, float* v,
                           float* tables, void* tables_bf16, void* tables_h, float lr, float beta1, float beta2, float eps,
                           int step, cudaStream_t st);
int egn_launch_pack_grads(const EgnConfig* cfg, const EgnGrads* grads, float* d_tables, cudaStream_t st);
int egn_launch_regularize(const EgnConfig* cfg, const float* tables, float* d_tables, float tv_density, float tv_app,
                          float l1_density, float* losses, cudaStream_t st);
int egn_launch_pack_h(const EgnConfig* cfg, const float* tables, void* tables_h, cudaStream_t st);
int egn_launch_pack_bf16(const EgnConfig* cfg, const float* tables, void* tables_bf16, cudaStream_t st);
int egn_launch_resample_factor(const float* src, int C, int H, int W, const float* ypos, int H2, const float* xpos, int W2,
                               float* dst, cudaStream_t st);
int egn_launch_erp_rays(int H, int W, int row0, int n_rows, const float* c2w_host, float* rays, cudaStream_t st);
// backward
int egn_launch_composite_bwd(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                             const float* fsig, const float* feat, const float* rgbs, const float* rgbpre,
                             const float* d_rgb, const float* d_bg, const float* d_env, const float* d_alpha,
                             float* d_rgbs, float* d_fsig, float* d_feat, float* d_emission, float* d_env_rays,
                             unsigned* gmax_bits, cudaStream_t st);
// MLP backward of a sub-chunk of n rays: recomputes the hidden activations into scratch (h1, h2, dz1, dz2: n*S x 128 floats)
int egn_launch_mlp_bwd(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                       const float* rgbs, const float* d_rgbs, float* d_feat, float* h1, float* h2, float* dz1,
                       float* dz2, const EgnGrads* g, cudaStream_t st);
// tensor-core MLP backward of the whole chunk (throughput mode): no scratch, weight gradients accumulated in TMEM.
// gmax_bits: largest |d(sample colour)| of the launch as float bits (egn_launch_composite_bwd) -> tc_grad_scale()
int egn_launch_mlp_bwd_tc(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                          const float* rgbs, const float* d_rgbs, const unsigned* gmax_bits, float* d_feat, const EgnGrads* g,
                          cudaStream_t st);
int egn_launch_gather_bwd_tc(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                             const float* d_fsig, const float* d_feat, const unsigned* gmax_bits, float* d_tables, const EgnGrads* g,
                             cudaStream_t st);
int egn_launch_mlp_save(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* feat,
                        float* h1, float* h2, cudaStream_t st);
int egn_launch_gather_bwd(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                          const float* d_fsig, const float* d_feat, float* d_tables, const EgnGrads* g, cudaStream_t st);
