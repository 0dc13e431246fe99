/*
 * egn.h — C ABI of libegn_b200.so: the B200 (sm_100a) volume-rendering path of EgoNeRF.
 *
 * The reference (changwoonchoi/EgoNeRF) has no FFI layer: its boundary for this path is the Python
 * operator surface.  Each entry point below names the reference function(s) it
 * replaces (file:line under the reference tree).  The host-side mirror that binds them with ctypes
 * is egonerf_b200/_lib.py; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers and sizes only; every pointer marked "device" is a CUDA device pointer that
 *     the CALLER owns (PyTorch's caching allocator in the Python mirror); the library allocates
 *     nothing and keeps no global mutable state, so all entry points are re-entrant (autograd calls
 *     backward from its own thread).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - return value: 0 = ok, non-zero = error; egn_last_error() returns a thread-local message.
 *   - all floating-point data is IEEE fp32 unless a mode says otherwise.
 */
#ifndef EGN_H_
#define EGN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGN_ABI_VERSION 8

/* renderModule kinds, TensorBase.init_render_func (models/tensorBase.py:187-203) */
enum { EGN_SHADE_MLP_FEA = 0, EGN_SHADE_MLP = 1, EGN_SHADE_RGB = 2, EGN_SHADE_SH = 3 };
/* TensorBase.feature2density (models/tensorBase.py:415-419) */
enum { EGN_ACT_SOFTPLUS = 0, EGN_ACT_RELU = 1 };
/* MLP arithmetic: EGN_MLP_FP32 exact fp32 FFMA; EGN_MLP_TC_SPLIT tcgen05 tensor cores with a 3-term bf16 split
 * (fp32-equivalent, inside the 1e-4 parity bound); EGN_MLP_TC_BF16 throughput mode: gather + basis + MLP fused into one
 * warp-specialised tcgen05 kernel.  Since ABI 7 its forward pass computes with FP16 operands (appearance tables, packed
 * half2 interpolation, MMA operands; fp32 accumulate) and keeps the density channels, alpha and compositing in fp32: rgb stays
 * inside the 1e-4 bound (the bf16 operands of ABI 6 gave 1e-3).  The name EGN_MLP_TC_BF16 is kept as an alias; the tcgen05
 * backward kernels still use bf16 operands (gradients need the exponent range). */
enum { EGN_MLP_FP32 = 0, EGN_MLP_TC_SPLIT = 1, EGN_MLP_TC_BF16 = 2, EGN_MLP_TC_F16 = 2 };

/* Static description of one model / scene.  Scalars mirror the constructor arguments of
 * EgoNeRF / TensorBase (models/tensorBase.py:133-139) and YinYangSphericalCoords
 * (models/coordinates.py:432-520). */
typedef struct EgnConfig {
    int32_t grid[3];          /* [N_r, N_theta, N_phi] per hemisphere (coordinates.py:507-520) */
    int32_t c_sigma;          /* density components per plane, n_lamb_sigma[i] (all three equal) */
    int32_t c_app;            /* appearance components per plane, n_lamb_sh[i] (all three equal) */
    int32_t app_dim;          /* basis_mat output width (27) */
    int32_t shading;          /* EGN_SHADE_* */
    int32_t view_pe, fea_pe;  /* positional-encoding frequencies (tensorBase.py:14-19) */
    int32_t feature_c;        /* MLP hidden width (128) */
    int32_t fea2dense;        /* EGN_ACT_* */
    int32_t n_coarse;         /* coarse samples per ray */
    int32_t n_fine;           /* inverse-CDF draws per ray (0 when resampling is off) */
    int32_t use_coarse_sample;/* EgoNeRF.py:536-539 */
    int32_t resampling;       /* EgoNeRF.py:525 */
    int32_t env_h;            /* envmap is (3, 2*env_h, env_h); 0 = no envmap (models/envmap.py:17-23) */
    int32_t mlp_mode;         /* EGN_MLP_* */
    float   center[3];        /* coordinates.center = aabb.sum(0)/2 (coordinates.py:77) */
    float   near_plane;       /* near_far[0] */
    float   density_shift;
    float   distance_scale;
    float   ang_near[2];      /* fp32(pi/4), fp32(-3pi/4)            (coordinates.py:500-505) */
    float   ang_inv[2];       /* 1/(far-near) of theta, phi in fp32  (coordinates.py:505) */
    int32_t exp_sampling;     /* 1: exponential schedule of sample_ray_exp (every shipped config); 0: uniform march of
                                 TensorBase.sample_ray (models/tensorBase.py:308-327) from the AABB entry point */
    float   far_plane;        /* near_far[1] (uniform march only) */
    float   step_size;        /* TensorBase.stepSize = mean(aabbSize / (gridSize - 1)) * step_ratio (tensorBase.py:206-213) */
    float   aabb[6];          /* [min xyz, max xyz] (uniform march only) */
    int32_t bwd_tc;           /* 1: egn_render_backward uses the tcgen05 backward kernels (bf16 operands, fp32 accumulate) even
                                 when the forward ran in EGN_MLP_FP32 / EGN_MLP_TC_SPLIT; always on for EGN_MLP_TC_BF16 */
    int32_t plain_ladders;    /* 0: `interval_th` ladders (intervals shorter than r0 forced to r0; coordinates.py:112-131,
                                 EgoNeRF.py:68-82) -- every shipped config.  1: the plain exponential ladders of a run
                                 WITHOUT --interval_th (opt.py:190): normalize_r = 1 + k + lin on r0*ratio^k
                                 (coordinates.py:132-156), the coarse pass on the N_r/2 ladder (`downsample=2`, :137-139),
                                 train jitter in the exponent (EgoNeRF.py:59-67) */
    float   jitter_ratio;     /* plain_ladders, train: z_j = near + jitter_r0 * sum_{i<j} jitter_ratio^(i + u_i);  */
    float   jitter_r0;        /*   ratio = 1 + (pi/2)/n_coarse, r0 = (far-near)(ratio-1)/(ratio^n_coarse - 1)     */
    const float* r_knots;     /* device: reference r ladder of normalize_r.  interval_th: N_r+1 knots (coordinates.py:118-124).
                                 plain_ladders: N_r+3 knots [0, r0, r0*ratio, ..., r0*ratio^(N_r+1)] (two past the grid: the
                                 closed form of :141-155 extrapolates exponentially, the search here clamps to the last knot) */
    const float* r_knots_coarse; /* device, plain_ladders only: the same ladder for N_r/2 cells, N_r/2+3 knots (ratio recomputed,
                                 coordinates.py:137-139); ignored (may be NULL) otherwise: interval_th ignores `downsample` */
    const float* z_coarse;    /* device, n_coarse : r schedule of sample_ray_exp WITHOUT near (EgoNeRF.py:69-76);
                                 the kernels add near_plane and, in train mode, the interval jitter (:78-82) */
    const void*  tables_bf16; /* device, optional: bf16 copy of the fine render tables (egn_pack_tables_bf16), re-gathered by the
                                 tcgen05 backward kernels instead of the fp32 tables; NULL = re-gather from fp32 */
    const void*  tables_h;    /* device: "half" tables of the fused fine pass (egn_pack_tables_h): per fine texel 48 fp16
                                 appearance channels in one aligned 128-byte line + 16 fp32 density channels.  Required by
                                 EGN_MLP_TC_F16 forward passes */
} EgnConfig;

/* Parameters in the REFERENCE layout (contiguous NCHW fp32, shapes of EgoNeRF.init_one_svd,
 * models/EgoNeRF.py:102-122): hemisphere h (0 = yin, 1 = yang), factor i (matMode [[0,1],[0,2],[1,2]],
 * vecMode [2,1,0]).  plane[h][i] is (1,C,G[m1],G[m0]); line[h][i] is (1,C,G[v],1). */
typedef struct EgnParams {
    const float* density_plane[2][3];
    const float* density_line[2][3];
    const float* app_plane[2][3];
    const float* app_line[2][3];
    const float* basis[2];    /* basis_mat_{yin,yang}.weight (app_dim, 3*c_app) */
    const float* mlp_w[3];    /* renderModule.mlp.{0,2,4}.weight (out,in) ; NULL for RGB / SH */
    const float* mlp_b[3];    /* renderModule.mlp.{0,2,4}.bias */
    const float* emission;    /* envmap.emission (3, 2*env_h, env_h) or NULL */
} EgnParams;

/* Gradients, same layout as EgnParams.  basis / mlp / emission buffers are ACCUMULATED into (caller zeroes);
 * the factor plane/line buffers are written by egn_unpack_table_grads. */
typedef struct EgnGrads {
    float* density_plane[2][3];
    float* density_line[2][3];
    float* app_plane[2][3];
    float* app_line[2][3];
    float* basis[2];
    float* mlp_w[3];
    float* mlp_b[3];
    float* emission;
} EgnGrads;

/* Per-ray outputs of EgoNeRF.forward (models/EgoNeRF.py:602): all device, caller-allocated. */
typedef struct EgnOutputs {
    float* rgb;     /* (n,3)   rgb_map, clamped to [0,1] (EgoNeRF.py:593) */
    float* depth;   /* (n)     depth_map (EgoNeRF.py:595-598) */
    float* bg;      /* (n,3)   bg_map = bg_weight * env, NULL without envmap */
    float* env;     /* (n,3)   env_map, NULL without envmap */
    float* alpha;   /* (n, S [+1 with envmap])  per-sample alpha (EgoNeRF.py:587,602) */
} EgnOutputs;

const char* egn_last_error(void);
int32_t     egn_abi_version(void);

/* S = samples composited per ray (EgoNeRF.py:536-539). */
int32_t egn_samples_per_ray(const EgnConfig* cfg);

/* ---- render tables -------------------------------------------------------------------------
 * The kernels gather from "render tables": channels-last copies of the factor planes/lines with
 * density and appearance channels interleaved per texel ([H][W][c_sigma+c_app], one tap = one
 * contiguous run), plus the 2x average-pooled density tables of the coarse pass.
 * egn_pack_tables replaces EgoNeRF.update_coarse_sigma_grid (models/EgoNeRF.py:124-133) and is
 * re-run whenever the parameters change (once per optimiser step in training). */
int64_t egn_table_floats(const EgnConfig* cfg);
int32_t egn_pack_tables(const EgnConfig* cfg, const EgnParams* params, float* tables /*device*/, void* stream);
/* Adam step (torch.optim.Adam semantics without weight decay / amsgrad; train.py:172-186) of the 24 factor tensors in
 * render-table space: consumes the table-layout gradient of egn_render_backward directly (no
 * egn_unpack_table_grads), keeps exp_avg / exp_avg_sq in table layout (egn_table_floats floats each, caller-zeroed once),
 * and writes the updated values to the NCHW parameter tensors (`params_out`), the fp32 render tables, their bf16 copy
 * (optional) and the pooled coarse tables (EgoNeRF.update_coarse_sigma_grid, models/EgoNeRF.py:124-133).  step counts from 1. */
int32_t egn_adam_tables(const EgnConfig* cfg, const EgnGrads* params_out, const float* d_tables /*device*/,
                        float* exp_avg /*device*/, float* exp_avg_sq /*device*/, float* tables /*device*/,
                        void* tables_bf16 /*device, nullable*/, void* tables_h /*device, nullable*/, float lr, float beta1,
                        float beta2, float eps, int32_t step, void* stream);
/* bf16 copy of the fine sections (same element offsets; egn_table_bf16_elems elements of 2 bytes) for the throughput mode */
int64_t egn_table_bf16_elems(const EgnConfig* cfg);
int32_t egn_pack_tables_bf16(const EgnConfig* cfg, const float* tables /*device*/, void* tables_bf16 /*device*/, void* stream);
/* half tables of the fused fine pass (EgnConfig.tables_h): egn_table_h_bytes bytes, built from the fp32 render tables */
int64_t egn_table_h_bytes(const EgnConfig* cfg);
int32_t egn_pack_tables_h(const EgnConfig* cfg, const float* tables /*device*/, void* tables_h /*device*/, void* stream);
/* inverse scatter for training: writes d(tables) (table layout) into the reference-layout factor grads
 * (overwrites grads->{density,app}_{plane,line}; the other members are untouched) */
int32_t egn_unpack_table_grads(const EgnConfig* cfg, const float* d_tables /*device*/, const EgnGrads* grads,
                               void* stream);

/* the reverse direction: gradients that sit in reference-layout tensors (written by autograd for losses that do not go through
 * egn_render_backward) are ADDED to the table-layout gradient; NULL members are skipped */
int32_t egn_pack_table_grads(const EgnConfig* cfg, const EgnGrads* grads, float* d_tables /*device*/, void* stream);
/* Regularisers of the factor tensors in table space: adds the gradients of
 *   tv_density * EgoNeRF.TV_loss_density(TVLoss()) + tv_app * EgoNeRF.TV_loss_app(TVLoss()) + l1_density * EgoNeRF.density_L1()
 * (utils.py:155-171, models/EgoNeRF.py:204-229, as train.py:288-305 weights them) to d_tables, and ADDS the three unweighted
 * loss values to losses[0..2] (device, caller-zeroed, nullable). */
int32_t egn_regularize_tables(const EgnConfig* cfg, const float* tables /*device*/, float* d_tables /*device*/, float tv_density,
                              float tv_app, float l1_density, float* losses /*device, 3 floats*/, void* stream);

/* ---- whole path -------------------------------------------------------------------------------
 * egn_render_forward replaces EgoNeRF.forward (models/EgoNeRF.py:491-602) for one ray chunk, i.e.
 * sample_ray_exp (:56-87), YinYangSphericalCoords.from_cartesian / normalize_coord
 * (models/coordinates.py:442-498, 110-156), compute_coarse_densityfeature (:232-289), raw2alpha
 * (models/tensorBase.py:22-27), sample_pdf (dataLoader/ray_utils.py:156-187), the sort (:537),
 * compute_densityfeature (:291-347), compute_appfeature (:349-413), renderModule
 * (models/tensorBase.py:30-129), EnvironmentMap.get_radiance (models/envmap.py:25-34) and the
 * compositing of :579-598.
 *   rays      device (n,6) = [origin, direction]
 *   is_train  0: deterministic schedule + linspace u (EgoNeRF.py:515-516, ray_utils.py:165-167)
 *             1: jittered; the uniforms come from u_coarse (n,n_coarse) / u_fine (n,n_fine) when given,
 *                else from a counter-based generator keyed by (seed, ray_index0 + ray, sample)
 *   workspace device; keep_for_backward = 1: egn_workspace_bytes(cfg, n) bytes, and after the call it holds the
 *             per-sample state (z, sigma feature, app feature, sample rgb) that egn_render_backward consumes;
 *             keep_for_backward = 0: egn_workspace_bytes_eval(cfg, n) bytes suffice (24 B/sample in the fused mode). */
int64_t egn_workspace_bytes(const EgnConfig* cfg, int64_t n_rays);
int64_t egn_workspace_bytes_eval(const EgnConfig* cfg, int64_t n_rays);   /* forward-only (no backward scratch) */
int32_t egn_render_forward(const EgnConfig* cfg, const EgnParams* params, const float* tables,
                           const float* rays, int64_t n_rays, int32_t is_train,
                           const float* u_coarse, const float* u_fine, uint64_t seed, int64_t ray_index0,
                           const EgnOutputs* out, void* workspace, int32_t keep_for_backward, void* stream);

/* Same call with per-stage device times (CUDA events on `stream`; the call synchronises on the last one).
 * stage_ms (host, EGN_N_STAGES floats): 0 sampler (coarse pass + inverse CDF + sort), 1 fine gather + basis,
 * 2 colour decode (MLP), 3 compositing.  Used by bench.py for the roofline of the dominant kernel. */
#define EGN_N_STAGES 4
int32_t egn_render_forward_timed(const EgnConfig* cfg, const EgnParams* params, const float* tables,
                                 const float* rays, int64_t n_rays, int32_t is_train,
                                 const float* u_coarse, const float* u_fine, uint64_t seed, int64_t ray_index0,
                                 const EgnOutputs* out, void* workspace, void* stream, float* stage_ms /*host*/);

/* The two halves of egn_render_forward, for callers that keep the reference's sampler / renderer split:
 *   egn_sample_rays    = sample_ray_exp + coarse density + raw2alpha + sample_pdf + sort  -> z_out (n,S) sorted depths
 *                        (EgoNeRF.py:507-542)
 *   egn_render_samples = everything after the depths are known (EgoNeRF.py:544-602); z_vals (n,S) is copied into
 *                        the workspace (NULL: the workspace already holds the depths). */
int32_t egn_sample_rays(const EgnConfig* cfg, const float* tables, const float* rays, int64_t n_rays, int32_t is_train,
                        const float* u_coarse, const float* u_fine, uint64_t seed, int64_t ray_index0,
                        float* z_out, void* stream);
int32_t egn_render_samples(const EgnConfig* cfg, const EgnParams* params, const float* tables, const float* rays,
                           int64_t n_rays, const float* z_vals, const EgnOutputs* out, void* workspace,
                           int32_t keep_for_backward, void* stream);

/* Backward of the above w.r.t. every parameter that receives a gradient in the reference (SURVEY.md
 * Appendix A10): fine density/appearance planes+lines, both basis matrices, the MLP, the envmap —
 * through rgb, bg, env and alpha; not through depth, the coarse pass, coordinates or rays.
 * d_* may be NULL (treated as zero).  d_tables (egn_table_floats floats, caller-zeroed) receives the
 * factor-table gradients in table layout; run egn_unpack_table_grads afterwards. */
int32_t egn_render_backward(const EgnConfig* cfg, const EgnParams* params, const float* tables,
                            const float* rays, int64_t n_rays, const void* workspace,
                            const float* d_rgb, const float* d_bg, const float* d_env, const float* d_alpha,
                            float* d_tables, const EgnGrads* grads, void* stream);

/* Same, with the envmap gradient in SPARSE form: when d_env_rays (device, (n,3)) is given, the gradient w.r.t. each ray's env
 * radiance (the sigmoid output of EnvironmentMap.get_radiance) is written there and grads->emission is not touched; the caller
 * scatters it with egn_envmap_backward.  For ray-sharded training: ranks exchange 24 B per ray (direction +
 * this gradient) instead of all-reducing the dense (3, 2h, h) envmap gradient (88 MB at h = 1920). */
int32_t egn_render_backward_sparse_env(const EgnConfig* cfg, const EgnParams* params, const float* tables,
                                       const float* rays, int64_t n_rays, const void* workspace,
                                       const float* d_rgb, const float* d_bg, const float* d_env, const float* d_alpha,
                                       float* d_tables, const EgnGrads* grads, float* d_env_rays /*device, nullable*/, void* stream);

/* ---- stand-alone operators --------------------------------------------------------------------
 * coords: device (m,7) normalised Yin-Yang coordinates [r,theta,phi | r,theta,phi | Y] as produced by
 * YinYangSphericalCoords.normalize_coord (models/coordinates.py:442-466). */
int32_t egn_density_feature(const EgnConfig* cfg, const float* tables, const float* coords7, int64_t m,
                            int32_t coarse, float* out /*(m)*/, void* stream);        /* EgoNeRF.py:291-347 / 232-289 */
/* one gather serves both operators: sigma_out (m) = compute_densityfeature, feat_out (m,28) = compute_appfeature
 * (rows padded from app_dim to 28 floats) */
int32_t egn_app_feature(const EgnConfig* cfg, const EgnParams* params, const float* tables, const float* coords7,
                        int64_t m, float* sigma_out /*(m)*/, float* feat_out /*(m,28)*/, void* stream);  /* EgoNeRF.py:349-413 */
int32_t egn_yinyang_coords(const EgnConfig* cfg, const float* xyz /*(m,3)*/, int64_t m,
                           float* coords7 /*(m,7) normalised*/, void* stream);       /* coordinates.py:442-498 */
int32_t egn_envmap_radiance(const EgnConfig* cfg, const float* emission, const float* dirs /*(n,3)*/, int64_t n,
                            float* out /*(n,3)*/, void* stream);                     /* envmap.py:25-34 */
int32_t egn_envmap_backward(const EgnConfig* cfg, const float* emission, const float* dirs, int64_t n,
                            const float* d_out, float* d_emission, void* stream);

/* Rays of rows [row0, row0 + n_rows) of an H x W equirectangular frame with pose c2w (HOST pointer, 3 x 4 row-major):
 * get_ray_directions_360 (dataLoader/ray_utils.py:24-40) + normalisation (dataset_omniblender.py:43) + get_rays (:85-113).
 * rays: device (n_rows * W, 6). */
int32_t egn_erp_rays(int32_t H, int32_t W, int32_t row0, int32_t n_rows, const float* c2w /*host*/, float* rays /*device*/,
                     void* stream);

/* Coarse-to-fine resampling of ONE factor tensor (train.py:371-377 -> EgoNeRF.upsample_volume_grid models/EgoNeRF.py:427-436
 * -> coordinates.up_sampling_VM models/coordinates.py:27-39 (angular axes, F.interpolate) / :226-266 (r axis, F.grid_sample on
 * the exponential ladder)).  src: NCHW (channels, h, w); dst: (channels, h2, w2).  ypos[h2] / xpos[w2] (device) hold the
 * SOURCE position of every output row / column in texel units; bilinear, taps outside the source contribute zero.
 * The host computes the positions (egonerf_b200/models/coordinates.py up_sampling_positions). */
int32_t egn_resample_factor(const float* src /*device*/, int32_t channels, int32_t h, int32_t w, const float* ypos /*device*/,
                            int32_t h2, const float* xpos /*device*/, int32_t w2, float* dst /*device*/, void* stream);

/* ---- host helpers (no GPU): the two ladders, for callers that do not build them with torch ------ */
int32_t egn_host_sample_schedule(float near_plane, float far_plane, float r0, int32_t n, float* z_out);   /* EgoNeRF.py:69-76 */
int32_t egn_host_r_knots(float far_r, float r0, int32_t n_r, float* knots_out /*n_r+1*/);                  /* coordinates.py:118-124 */
/* ... and for EgnConfig.plain_ladders (a run without --interval_th): schedule of EgoNeRF.py:59-66 (eval; also returns the
 * ratio / r0 of the train jitter) and the N_r+3 knots r0*ratio^(i-1) behind coordinates.py:132-155 (call it with n_r = N_r/2
 * for r_knots_coarse). */
int32_t egn_host_plain_sample_schedule(float near_plane, float far_plane, int32_t n, float* z_out, float* ratio_out /*nullable*/,
                                       float* r0_out /*nullable*/);
int32_t egn_host_plain_r_knots(float far_r, float r0, int32_t n_r, float* knots_out /*n_r+3*/);

/* ---- gradient exchange of ray-sharded training over NVLink peer memory --------------------------------
 * The reference is single-process (train.py:20, no process group at train.py:409-422); the multi-GPU launcher sums the
 * gradients of the ray shards once per step.  These entry points replace the `ncclAllReduce` of the table-layout factor
 * gradient by ONE kernel over peer memory: rank r sums slice r of every rank's buffer (peer loads) and stores the sum into
 * slice r of every rank's buffer (peer stores); block b of every rank handshakes with block b of every peer before the first
 * load and after the last store (flags in peer memory, release / acquire at system scope; a missing peer traps after 4 s).
 *   egn_peer_alloc / _free      cudaMalloc'ed, zero-filled device memory (cudaIpc handles address whole allocations)
 *   egn_peer_export / _open / _close   cudaIpcGetMemHandle / cudaIpcOpenMemHandle (lazy peer access) / cudaIpcCloseMemHandle
 *   egn_peer_flag_bytes         size of the flag block every rank allocates with egn_peer_alloc and exports next to its buffer
 *   egn_peer_allreduce          bufs[world], flags[world]: HOST arrays of device pointers in rank order (own entries = the local
 *                               allocations); n_floats % 4 == 0; every rank passes the same n_floats, scale, epoch (1, 2, 3, ...
 *                               per call) and blocks (<= 256); result: buf_p[i] = scale * sum_q buf_q[i] on every rank p,
 *                               bit-identical across ranks. */
int64_t egn_peer_flag_bytes(void);
int32_t egn_peer_alloc(int64_t bytes, void** ptr_out);
int32_t egn_peer_free(void* ptr);
int32_t egn_peer_export(const void* ptr, unsigned char handle_out[64]);
int32_t egn_peer_open(const unsigned char handle[64], void** ptr_out);
int32_t egn_peer_close(void* ptr);
int32_t egn_peer_allreduce(void* const* bufs, void* const* flags, int32_t rank, int32_t world, int64_t n_floats, float scale,
                           uint32_t epoch, int32_t blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGN_H_ */
