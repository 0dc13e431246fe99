// C ABI of libegn_b200 (include/egn.h): argument validation, workspace carve-up, launch sequencing.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "egn_device.cuh"
#include "egn_host.h"

static thread_local char g_err[512] = "";
static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
// for the other translation units of the library (egn_peer.cu)
int egn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
static int cuda_fail(const char* what, int e) {
    return fail("%s: %s", what, cudaGetErrorString((cudaError_t)e));
}

extern "C" const char* egn_last_error(void) { return g_err; }
extern "C" int32_t egn_abi_version(void) { return EGN_ABI_VERSION; }

extern "C" int32_t egn_samples_per_ray(const EgnConfig* c) {
    if (!c->resampling) return c->n_coarse;
    return (c->use_coarse_sample ? c->n_coarse : 0) + c->n_fine;
}

static int validate(const EgnConfig* c, bool need_schedule) {
    if (!c) return fail("null config");
    if (c->c_sigma != EGN_CS || c->c_app != EGN_CA)
        return fail("this build supports n_lamb_sigma=[%d]*3, n_lamb_sh=[%d]*3 (got %d, %d)", EGN_CS, EGN_CA, c->c_sigma, c->c_app);
    for (int a = 0; a < 3; ++a)
        if (c->grid[a] < 4 || c->grid[a] > 65535) return fail("grid[%d]=%d outside [4, 65535]", a, c->grid[a]);
    if (c->grid[0] + 2 > EGN_MAX_KNOTS) return fail("N_r=%d exceeds %d", c->grid[0], EGN_MAX_KNOTS - 2);
    if (c->app_dim < 1 || c->app_dim > 27) return fail("app_dim=%d unsupported (1..27)", c->app_dim);
    if (c->shading < 0 || c->shading > 3) return fail("unknown shading %d", c->shading);
    if (c->shading == EGN_SHADE_SH && c->app_dim != 27) return fail("SH shading needs app_dim 27");
    if (c->shading == EGN_SHADE_RGB && c->app_dim != 3) return fail("RGB shading needs app_dim 3");
    if (c->shading <= EGN_SHADE_MLP) {
        if (c->feature_c != EGN_HID) return fail("featureC=%d unsupported (%d)", c->feature_c, EGN_HID);
        int in_dim = c->app_dim + 3 + 6 * c->view_pe + (c->shading == EGN_SHADE_MLP_FEA ? 2 * c->fea_pe * c->app_dim : 0);
        if (in_dim > 152) return fail("MLP input width %d exceeds 152", in_dim);
        if (c->mlp_mode < EGN_MLP_FP32 || c->mlp_mode > EGN_MLP_TC_BF16) return fail("unknown mlp_mode %d", c->mlp_mode);
        if (c->mlp_mode != EGN_MLP_FP32 && !(c->shading == EGN_SHADE_MLP_FEA && c->view_pe == 2 && c->fea_pe == 2))
            return fail("tensor-core MLP modes need shadingMode MLP_Fea with view_pe = fea_pe = 2 (use EGN_MLP_FP32)");
    }
    if (need_schedule) {
        if (c->n_coarse < 32 || c->n_coarse > 256 || c->n_coarse % 32) return fail("n_coarse=%d must be a multiple of 32 in [32,256]", c->n_coarse);
        if (c->resampling && (c->n_fine < 32 || c->n_fine > 256 || c->n_fine % 32)) return fail("n_fine=%d must be a multiple of 32 in [32,256]", c->n_fine);
        if (!c->r_knots || (c->exp_sampling && !c->z_coarse)) return fail("r_knots / z_coarse tables missing");
        if (!c->exp_sampling && !(c->step_size > 0.f)) return fail("uniform march needs step_size > 0");
        if (c->plain_ladders && !c->r_knots_coarse) return fail("plain_ladders needs r_knots_coarse (N_r/2 + 3 knots)");
        if (c->plain_ladders && c->exp_sampling && !(c->jitter_ratio > 1.f && c->jitter_r0 > 0.f))
            return fail("plain_ladders needs jitter_ratio > 1 and jitter_r0 > 0");
    }
    return 0;
}

static EgnKernelCfg make_kcfg(const EgnConfig* c, const float* tables, bool render = false) {
    EgnKernelCfg k;
    memset(&k, 0, sizeof(k));
    k.lay = egn_make_layout(c->grid);
    k.tables = tables;
    k.r_knots = c->r_knots;
    k.plain_ladders = c->plain_ladders ? 1 : 0;
    k.knots_last = c->grid[0] + (k.plain_ladders ? 2 : 0);
    k.r_knots_c = k.plain_ladders ? c->r_knots_coarse : c->r_knots;
    k.r_div_c = k.plain_ladders ? c->grid[0] / 2 : c->grid[0];
    k.knots_last_c = k.r_div_c + (k.plain_ladders ? 2 : 0);
    k.r_div = c->grid[0];
    if (render && k.plain_ladders && !c->resampling) {
        // reference quirk (EgoNeRF.py:523,565-572): without resampling the fine grid is read at the coordinates the coarse pass
        // normalised with downsample=2 -- with the plain ladders that is the N_r/2 ladder
        k.r_knots = k.r_knots_c; k.knots_last = k.knots_last_c; k.r_div = k.r_div_c;
    }
    k.jitter_ratio = c->jitter_ratio; k.jitter_r0 = c->jitter_r0;
    k.z_coarse = c->z_coarse;
    for (int a = 0; a < 3; ++a) k.center[a] = c->center[a];
    for (int a = 0; a < 2; ++a) { k.ang_near[a] = c->ang_near[a]; k.ang_inv[a] = c->ang_inv[a]; }
    k.density_shift = c->density_shift;
    k.distance_scale = c->distance_scale;
    k.n_coarse = c->n_coarse; k.n_fine = c->resampling ? c->n_fine : 0;
    k.S = egn_samples_per_ray(c);
    k.use_coarse_sample = c->use_coarse_sample; k.resampling = c->resampling;
    k.fea2dense = c->fea2dense; k.shading = c->shading; k.app_dim = c->app_dim;
    k.view_pe = c->view_pe; k.fea_pe = c->fea_pe; k.env_h = c->env_h;
    k.mlp_mode = c->mlp_mode;
    k.march = c->exp_sampling ? 0 : 1;
    k.step_size = c->step_size; k.far_plane = c->far_plane;
    for (int i = 0; i < 6; ++i) k.aabb[i] = c->aabb[i];
    k.tables_bf16 = c->tables_bf16;
    k.tables_h = c->tables_h;
    const EgnLayoutH LH = egn_make_layout_h(c->grid);
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) { k.texp[h][i] = LH.texp[h][i]; k.texl[h][i] = LH.texl[h][i]; }
    k.dens_byte_offset = LH.dens_byte_offset;
    return k;
}

extern "C" int64_t egn_table_floats(const EgnConfig* c) {
    if (validate(c, false)) return -1;
    return egn_make_layout(c->grid).total;
}

extern "C" int32_t egn_pack_tables(const EgnConfig* c, const EgnParams* p, float* tables, void* stream) {
    if (validate(c, false)) return 1;
    if (!p || !tables) return fail("null argument");
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i)
            if (!p->density_plane[h][i] || !p->density_line[h][i] || !p->app_plane[h][i] || !p->app_line[h][i])
                return fail("missing factor tensor h=%d i=%d", h, i);
    int e = egn_launch_pack(c, p, tables, (cudaStream_t)stream);
    return e ? cuda_fail("egn_pack_tables", e) : 0;
}

extern "C" int64_t egn_table_bf16_elems(const EgnConfig* c) {
    if (validate(c, false)) return -1;
    return egn_make_layout(c->grid).pc[0][0];
}

extern "C" int32_t egn_pack_tables_bf16(const EgnConfig* c, const float* tables, void* tables_bf16, void* stream) {
    if (validate(c, false)) return 1;
    if (!tables || !tables_bf16) return fail("null argument");
    int e = egn_launch_pack_bf16(c, tables, tables_bf16, (cudaStream_t)stream);
    return e ? cuda_fail("egn_pack_tables_bf16", e) : 0;
}

extern "C" int64_t egn_table_h_bytes(const EgnConfig* c) {
    if (validate(c, false)) return -1;
    return egn_make_layout_h(c->grid).total_bytes;
}

extern "C" int32_t egn_pack_tables_h(const EgnConfig* c, const float* tables, void* tables_h, void* stream) {
    if (validate(c, false)) return 1;
    if (!tables || !tables_h) return fail("null argument");
    int e = egn_launch_pack_h(c, tables, tables_h, (cudaStream_t)stream);
    return e ? cuda_fail("egn_pack_tables_h", e) : 0;
}

extern "C" int32_t egn_adam_tables(const EgnConfig* c, const EgnGrads* params_out, const float* d_tables, float* exp_avg,
                                   float* exp_avg_sq, float* tables, void* tables_bf16, void* tables_h, float lr, float beta1,
                                   float beta2, float eps, int32_t step, void* stream) {
    if (validate(c, false)) return 1;
    if (!params_out || !d_tables || !exp_avg || !exp_avg_sq || !tables) return fail("null argument");
    if (step < 1) return fail("Adam step counts from 1");
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i)
            if (!params_out->density_plane[h][i] || !params_out->density_line[h][i] || !params_out->app_plane[h][i] ||
                !params_out->app_line[h][i]) return fail("missing factor tensor h=%d i=%d", h, i);
    int e = egn_launch_adam_tables(c, params_out, d_tables, exp_avg, exp_avg_sq, tables, tables_bf16, tables_h, lr, beta1, beta2,
                                   eps, step, (cudaStream_t)stream);
    return e ? cuda_fail("egn_adam_tables", e) : 0;
}

extern "C" int32_t egn_unpack_table_grads(const EgnConfig* c, const float* d_tables, const EgnGrads* g, void* stream) {
    if (validate(c, false)) return 1;
    if (!d_tables || !g) return fail("null argument");
    int e = egn_launch_unpack(c, d_tables, g, (cudaStream_t)stream);
    return e ? cuda_fail("egn_unpack_table_grads", e) : 0;
}

extern "C" int32_t egn_pack_table_grads(const EgnConfig* c, const EgnGrads* g, float* d_tables, void* stream) {
    if (validate(c, false)) return 1;
    if (!d_tables || !g) return fail("null argument");
    int e = egn_launch_pack_grads(c, g, d_tables, (cudaStream_t)stream);
    return e ? cuda_fail("egn_pack_table_grads", e) : 0;
}

extern "C" int32_t egn_regularize_tables(const EgnConfig* c, const float* tables, float* d_tables, float tv_density, float tv_app,
                                         float l1_density, float* losses, void* stream) {
    if (validate(c, false)) return 1;
    if (!tables || !d_tables) return fail("null argument");
    int e = egn_launch_regularize(c, tables, d_tables, tv_density, tv_app, l1_density, losses, (cudaStream_t)stream);
    return e ? cuda_fail("egn_regularize_tables", e) : 0;
}

// ---- workspace ----------------------------------------------------------------------------------
static inline long long align256(long long b) { return (b + 255) / 256 * 256; }
// backward processes the MLP in sub-chunks of this many rays so that its M x 128 scratch stays bounded
#define EGN_BWD_SUB_RAYS 4096
struct WsPlan { long long z, image, fsig, rgbs, wgt, bgw, rgbpre, feat, coord, d_rgbs, d_fsig, d_feat, gmax, h1, h2, dz1, dz2, eval_total, total; };
// the fused fine pass keeps the r ladder in shared memory (EGN_FUSED_MAX_KNOTS entries); larger grids take the unfused kernels
static bool is_fused(const EgnConfig* c) {
    return c->shading == EGN_SHADE_MLP_FEA && c->mlp_mode == EGN_MLP_TC_F16 && c->grid[0] + 3 <= EGN_FUSED_MAX_KNOTS;
}
static bool tc_backward(const EgnConfig* c) {
    return c->shading == EGN_SHADE_MLP_FEA && c->view_pe == 2 && c->fea_pe == 2 && (c->mlp_mode == EGN_MLP_TC_BF16 || c->bwd_tc);
}
static WsPlan plan_ws(const EgnConfig* c, long long n) {
    const long long S = egn_samples_per_ray(c);
    const long long M = n * S;
    WsPlan w;
    long long off = 0;
    auto take = [&](long long floats) { long long o = off; off += align256(floats * 4); return o; };
    w.z = take(M);
    w.image = off; off += align256(is_fused(c) ? egn_fused_image_bytes() : 0);       // MLP operand image of the fused fine pass
    w.fsig = take(M); w.rgbs = take(M * 3); w.wgt = take(M); w.bgw = take(n); w.rgbpre = take(n * 3);
    // the fused fine pass never materialises the app feature in eval mode: 24 B/sample instead of 136
    const long long eval_fused = off;
    w.feat = take(M * EGN_FEAT_STRIDE);
    w.eval_total = is_fused(c) ? eval_fused : off;
    w.coord = take(is_fused(c) && tc_backward(c) ? M * 4 : 0);          // sample coordinates handed from the fused forward to the backward
    w.d_rgbs = take(M * 3); w.d_fsig = take(M); w.d_feat = take(M * EGN_FEAT_STRIDE);
    w.gmax = take(1);                                                    // launch-wide max |d(sample colour)| (tcgen05 backward)
    const bool mlp = c->shading <= EGN_SHADE_MLP && !tc_backward(c);   // the tcgen05 backward needs no scratch
    const long long Ms = (n < EGN_BWD_SUB_RAYS ? n : EGN_BWD_SUB_RAYS) * S;
    w.h1 = take(mlp ? Ms * EGN_HID : 0); w.h2 = take(mlp ? Ms * EGN_HID : 0);
    w.dz1 = take(mlp ? Ms * EGN_HID : 0); w.dz2 = take(mlp ? Ms * EGN_HID : 0);
    w.total = off;
    return w;
}
extern "C" int64_t egn_workspace_bytes(const EgnConfig* c, int64_t n) {
    if (validate(c, false)) return -1;
    return plan_ws(c, n).total;
}
// forward-only callers (eval) may pass a workspace of this many bytes instead: no backward scratch
extern "C" int64_t egn_workspace_bytes_eval(const EgnConfig* c, int64_t n) {
    if (validate(c, false)) return -1;
    return plan_ws(c, n).eval_total;
}

static int check_render_args(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays,
                             const EgnOutputs* out, const void* workspace) {
    if (validate(c, true)) return 1;
    if (!p || !tables || !rays || !out || !workspace) return fail("null argument");
    if (!out->rgb || !out->depth || !out->alpha) return fail("rgb/depth/alpha outputs are required");
    if (c->env_h > 0 && (!p->emission || !out->bg || !out->env)) return fail("envmap configured but emission/bg/env missing");
    if (c->shading <= EGN_SHADE_MLP)
        for (int l = 0; l < 3; ++l)
            if (!p->mlp_w[l] || !p->mlp_b[l]) return fail("MLP weights missing");
    if (!p->basis[0] || !p->basis[1]) return fail("basis matrices missing");
    return 0;
}

extern "C" int32_t egn_sample_rays(const EgnConfig* c, const float* tables, const float* rays, int64_t n, int32_t is_train,
                                   const float* u_c, const float* u_f, uint64_t seed, int64_t ray0, float* z_out,
                                   void* stream) {
    if (validate(c, true)) return 1;
    if (n <= 0) return 0;
    if (!tables || !rays || !z_out) return fail("null argument");
    EgnKernelCfg k = make_kcfg(c, tables, true);
    int e = egn_launch_coarse(k, rays, n, is_train, u_c, u_f, seed, ray0, c->near_plane, z_out, (cudaStream_t)stream);
    return e ? cuda_fail("egn_sample_rays", e) : 0;
}

// stage boundaries recorded when a caller asks for per-stage device times (egn_render_forward_timed)
struct StageEvents { cudaEvent_t ev[EGN_N_STAGES + 1]; bool on; };
static inline void mark(StageEvents* se, int i, cudaStream_t st) { if (se && se->on) cudaEventRecord(se->ev[i], st); }

static int render_samples_impl(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays, int64_t n,
                               const float* z_vals, const EgnOutputs* out, void* workspace, cudaStream_t st,
                               StageEvents* se, bool save_feat) {
    EgnKernelCfg k = make_kcfg(c, tables, true);
    WsPlan w = plan_ws(c, n);
    char* base = (char*)workspace;
    float* z = (float*)(base + w.z);
    float* fsig = (float*)(base + w.fsig);
    float* feat = (float*)(base + w.feat);
    float* rgbs = (float*)(base + w.rgbs);
    float* wgt = (float*)(base + w.wgt);
    float* bgw = (float*)(base + w.bgw);
    float* rgbpre = (float*)(base + w.rgbpre);
    int e;
    if (z_vals && z_vals != z)
        if ((e = (int)cudaMemcpyAsync(z, z_vals, sizeof(float) * n * k.S, cudaMemcpyDeviceToDevice, st))) return cuda_fail("z copy", e);
    mark(se, 1, st);
    const bool fused = is_fused(c);
    if (fused && !c->tables_h) return fail("EGN_MLP_TC_F16 needs EgnConfig.tables_h (egn_pack_tables_h)");
    if (fused && (long long)n * k.S >= 0xffffff00ll)
        return fail("fused fine pass: %lld samples in one call (32-bit sample indices inside the kernel): render in smaller chunks", (long long)n * k.S);
    if (fused) {
        // throughput mode: one warp-specialised kernel for gather + basis + MLP; the app feature is only written
        // when a backward pass will read it (full training workspace)
        // forward-only calls with whole 128-sample tiles per ray composite inside the kernel (no egn_composite_kernel launch)
        const bool comp = !save_feat && k.S % 128 == 0;
        if (save_feat && tc_backward(c)) k.coords = (float*)(base + w.coord);
        if ((e = egn_launch_fused_fine(k, p, rays, n, z, fsig, save_feat ? feat : nullptr, rgbs, comp ? out : nullptr, base + w.image, st)))
            return cuda_fail("fused fine pass", e);
        mark(se, 2, st);
        if (comp) { mark(se, 3, st); mark(se, 4, st); return 0; }
    } else {
        if ((e = egn_launch_gather(k, p, rays, n, z, fsig, feat, st))) return cuda_fail("gather", e);
        mark(se, 2, st);
    }
    if (!fused && c->shading <= EGN_SHADE_MLP) {
        if (c->mlp_mode == EGN_MLP_FP32) e = egn_launch_mlp(k, p, rays, n, feat, rgbs, st);
        else e = egn_launch_mlp_tc(k, p, rays, n, feat, rgbs, c->mlp_mode == EGN_MLP_TC_SPLIT, st);
        if (e) return cuda_fail("mlp", e);
    }
    mark(se, 3, st);
    if ((e = egn_launch_composite(k, p, rays, n, z, fsig, feat, rgbs, out, wgt, bgw, rgbpre, st))) return cuda_fail("composite", e);
    mark(se, 4, st);
    return 0;
}

extern "C" int32_t egn_render_samples(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays,
                                      int64_t n, const float* z_vals, const EgnOutputs* out, void* workspace,
                                      int32_t keep_for_backward, void* stream) {
    if (n <= 0) return validate(c, true);                 // an empty chunk is valid and launches nothing
    if (check_render_args(c, p, tables, rays, out, workspace)) return 1;
    return render_samples_impl(c, p, tables, rays, n, z_vals, out, workspace, (cudaStream_t)stream, nullptr, keep_for_backward != 0);
}

static int render_forward_impl(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays,
                               int64_t n, int32_t is_train, const float* u_c, const float* u_f, uint64_t seed,
                               int64_t ray0, const EgnOutputs* out, void* workspace, bool keep, cudaStream_t st, StageEvents* se) {
    float* z = (float*)((char*)workspace + plan_ws(c, n).z);
    EgnKernelCfg k = make_kcfg(c, tables, true);
    mark(se, 0, st);
    int e = egn_launch_coarse(k, rays, n, is_train, u_c, u_f, seed, ray0, c->near_plane, z, st);
    if (e) return cuda_fail("egn_sample_rays", e);
    return render_samples_impl(c, p, tables, rays, n, nullptr, out, workspace, st, se, keep);
}

extern "C" int32_t egn_render_forward(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays,
                                      int64_t n, int32_t is_train, const float* u_c, const float* u_f, uint64_t seed,
                                      int64_t ray0, const EgnOutputs* out, void* workspace, int32_t keep_for_backward,
                                      void* stream) {
    if (n <= 0) return validate(c, true);
    if (check_render_args(c, p, tables, rays, out, workspace)) return 1;
    return render_forward_impl(c, p, tables, rays, n, is_train, u_c, u_f, seed, ray0, out, workspace, keep_for_backward != 0,
                               (cudaStream_t)stream, nullptr);
}

extern "C" int32_t egn_render_forward_timed(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays,
                                            int64_t n, int32_t is_train, const float* u_c, const float* u_f,
                                            uint64_t seed, int64_t ray0, const EgnOutputs* out, void* workspace,
                                            void* stream, float* stage_ms) {
    if (check_render_args(c, p, tables, rays, out, workspace)) return 1;
    if (!stage_ms) return fail("stage_ms is required");
    for (int i = 0; i < EGN_N_STAGES; ++i) stage_ms[i] = 0.f;
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    StageEvents se;
    se.on = true;
    for (int i = 0; i <= EGN_N_STAGES; ++i) cudaEventCreate(&se.ev[i]);
    int rc = render_forward_impl(c, p, tables, rays, n, is_train, u_c, u_f, seed, ray0, out, workspace, false, st, &se);
    if (!rc) {
        int e = (int)cudaEventSynchronize(se.ev[EGN_N_STAGES]);
        if (e) rc = cuda_fail("egn_render_forward_timed", e);
        else for (int i = 0; i < EGN_N_STAGES; ++i) cudaEventElapsedTime(&stage_ms[i], se.ev[i], se.ev[i + 1]);
    }
    for (int i = 0; i <= EGN_N_STAGES; ++i) cudaEventDestroy(se.ev[i]);
    return rc;
}

extern "C" int32_t egn_render_backward(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays,
                                       int64_t n, const void* workspace, const float* d_rgb, const float* d_bg,
                                       const float* d_env, const float* d_alpha, float* d_tables, const EgnGrads* g,
                                       void* stream) {
    return egn_render_backward_sparse_env(c, p, tables, rays, n, workspace, d_rgb, d_bg, d_env, d_alpha, d_tables, g, nullptr, stream);
}

extern "C" int32_t egn_render_backward_sparse_env(const EgnConfig* c, const EgnParams* p, const float* tables, const float* rays,
                                                  int64_t n, const void* workspace, const float* d_rgb, const float* d_bg,
                                                  const float* d_env, const float* d_alpha, float* d_tables, const EgnGrads* g,
                                                  float* d_env_rays, void* stream) {
    if (validate(c, true)) return 1;
    if (n <= 0) return 0;
    if (!p || !tables || !rays || !workspace || !d_tables || !g) return fail("null argument");
    if (!g->basis[0] || !g->basis[1]) return fail("basis gradient buffers missing");
    const bool mlp = c->shading <= EGN_SHADE_MLP;
    if (mlp)
        for (int l = 0; l < 3; ++l)
            if (!p->mlp_w[l] || !p->mlp_b[l] || !g->mlp_w[l] || !g->mlp_b[l]) return fail("MLP weights / gradient buffers missing");
    if (c->env_h > 0 && (!p->emission || (!g->emission && !d_env_rays))) return fail("envmap configured but emission / its gradient missing");
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    EgnKernelCfg k = make_kcfg(c, tables, true);
    WsPlan w = plan_ws(c, n);
    char* base = (char*)workspace;
    const float* z = (const float*)(base + w.z);
    const float* fsig = (const float*)(base + w.fsig);
    const float* feat = (const float*)(base + w.feat);
    const float* rgbs = (const float*)(base + w.rgbs);
    const float* rgbpre = (const float*)(base + w.rgbpre);
    float* d_rgbs = (float*)(base + w.d_rgbs);
    float* d_fsig = (float*)(base + w.d_fsig);
    float* d_feat = (float*)(base + w.d_feat);
    unsigned* gmax = (mlp && tc_backward(c)) ? (unsigned*)(base + w.gmax) : nullptr;
    int e;
    if ((e = egn_launch_composite_bwd(k, p, rays, n, z, fsig, feat, rgbs, rgbpre, d_rgb, d_bg, d_env, d_alpha, d_rgbs,
                                      d_fsig, d_feat, g->emission, c->env_h > 0 ? d_env_rays : nullptr, gmax, st))) return cuda_fail("composite backward", e);
    if (mlp && tc_backward(c)) {
        if ((e = egn_launch_mlp_bwd_tc(k, p, rays, n, feat, rgbs, d_rgbs, gmax, d_feat, g, st))) return cuda_fail("mlp backward (tcgen05)", e);
    } else if (mlp) {
        float* h1 = (float*)(base + w.h1); float* h2 = (float*)(base + w.h2);
        float* dz1 = (float*)(base + w.dz1); float* dz2 = (float*)(base + w.dz2);
        for (long long r0 = 0; r0 < n; r0 += EGN_BWD_SUB_RAYS) {
            const long long ns = (n - r0 < EGN_BWD_SUB_RAYS) ? (n - r0) : EGN_BWD_SUB_RAYS;
            const long long m0 = r0 * k.S;
            if ((e = egn_launch_mlp_bwd(k, p, rays + r0 * 6, ns, feat + m0 * EGN_FEAT_STRIDE, rgbs + m0 * 3, d_rgbs + m0 * 3,
                                        d_feat + m0 * EGN_FEAT_STRIDE, h1, h2, dz1, dz2, g, st))) return cuda_fail("mlp backward", e);
        }
    }
    if (tc_backward(c)) {
        if (is_fused(c)) k.coords = (float*)(base + w.coord);          // written by the forward of this very workspace
        e = egn_launch_gather_bwd_tc(k, p, rays, n, z, d_fsig, d_feat, gmax, d_tables, g, st);
    }
    else
        e = egn_launch_gather_bwd(k, p, rays, n, z, d_fsig, d_feat, d_tables, g, st);
    if (e) return cuda_fail("gather backward", e);
    return 0;
}

// ---- stand-alone operators ---------------------------------------------------------------------------
extern "C" int32_t egn_density_feature(const EgnConfig* c, const float* tables, const float* coords7, int64_t m,
                                       int32_t coarse, float* out, void* stream) {
    if (validate(c, false)) return 1;
    if (!tables || !coords7 || !out) return fail("null argument");
    if (m <= 0) return 0;
    EgnKernelCfg k = make_kcfg(c, tables);
    int e = egn_launch_gather_coords(k, nullptr, coords7, m, coarse, out, nullptr, (cudaStream_t)stream);
    return e ? cuda_fail("egn_density_feature", e) : 0;
}

extern "C" int32_t egn_app_feature(const EgnConfig* c, const EgnParams* p, const float* tables, const float* coords7,
                                   int64_t m, float* fsig_scratch, float* out28, void* stream) {
    if (validate(c, false)) return 1;
    if (!p || !tables || !coords7 || !out28 || !fsig_scratch) return fail("null argument");
    if (!p->basis[0] || !p->basis[1]) return fail("basis matrices missing");
    if (m <= 0) return 0;
    EgnKernelCfg k = make_kcfg(c, tables);
    int e = egn_launch_gather_coords(k, p, coords7, m, 0, fsig_scratch, out28, (cudaStream_t)stream);
    return e ? cuda_fail("egn_app_feature", e) : 0;
}

extern "C" int32_t egn_yinyang_coords(const EgnConfig* c, const float* xyz, int64_t m, float* coords7, void* stream) {
    if (!c) return fail("null config");
    if (c->grid[0] < 4 || c->grid[0] > EGN_MAX_KNOTS) return fail("N_r=%d out of range", c->grid[0]);
    if (!xyz || !coords7 || !c->r_knots) return fail("null argument");
    if (m <= 0) return 0;
    EgnKernelCfg k = make_kcfg(c, nullptr);
    int e = egn_launch_coords(k, xyz, m, coords7, (cudaStream_t)stream);
    return e ? cuda_fail("egn_yinyang_coords", e) : 0;
}

extern "C" int32_t egn_envmap_radiance(const EgnConfig* c, const float* emission, const float* dirs, int64_t n, float* out,
                                       void* stream) {
    if (!c || c->env_h <= 0) return fail("no envmap configured");
    if (!emission || !dirs || !out) return fail("null argument");
    if (n <= 0) return 0;
    int e = egn_launch_envmap(c->env_h, emission, dirs, n, out, (cudaStream_t)stream);
    return e ? cuda_fail("egn_envmap_radiance", e) : 0;
}

extern "C" int32_t egn_envmap_backward(const EgnConfig* c, const float* emission, const float* dirs, int64_t n,
                                       const float* d_out, float* d_emission, void* stream) {
    if (!c || c->env_h <= 0) return fail("no envmap configured");
    if (!emission || !dirs || !d_out || !d_emission) return fail("null argument");
    if (n <= 0) return 0;
    int e = egn_launch_envmap_bwd(c->env_h, emission, dirs, n, d_out, d_emission, (cudaStream_t)stream);
    return e ? cuda_fail("egn_envmap_backward", e) : 0;
}

extern "C" int32_t egn_erp_rays(int32_t H, int32_t W, int32_t row0, int32_t n_rows, const float* c2w, float* rays, void* stream) {
    if (H <= 0 || W <= 0 || row0 < 0 || n_rows < 0 || row0 + n_rows > H) return fail("bad frame / row range");
    if (!c2w || !rays) return fail("null argument");
    int e = egn_launch_erp_rays(H, W, row0, n_rows, c2w, rays, (cudaStream_t)stream);
    return e ? cuda_fail("egn_erp_rays", e) : 0;
}

extern "C" int32_t egn_resample_factor(const float* src, int32_t channels, int32_t h, int32_t w, const float* ypos, int32_t h2,
                                       const float* xpos, int32_t w2, float* dst, void* stream) {
    if (channels <= 0 || h <= 0 || w <= 0 || h2 <= 0 || w2 <= 0) return fail("bad factor shape");
    if (!src || !ypos || !xpos || !dst) return fail("null argument");
    if (src == dst) return fail("egn_resample_factor cannot work in place");
    int e = egn_launch_resample_factor(src, channels, h, w, ypos, h2, xpos, w2, dst, (cudaStream_t)stream);
    return e ? cuda_fail("egn_resample_factor", e) : 0;
}

// ---- host helpers -------------------------------------------------------------------------------------
// "first K intervals forced to r0, the rest shifted" (EgoNeRF.py:72-76, coordinates.py:120-124), fp32 arithmetic
static void force_linear_prefix(float* r, int n, float r0) {
    int K = 0;
    float cum = 0.f, cumK = 0.f;
    for (int i = 0; i + 1 < n; ++i) {
        float iv = r[i + 1] - r[i];
        cum += iv;
        if (iv <= r0) { ++K; }
    }
    cum = 0.f;
    for (int i = 0; i < K; ++i) { cum += r[i + 1] - r[i]; }
    cumK = cum;
    for (int i = K + 1; i < n; ++i) r[i] = r[i] + r0 * (float)K - cumK;
    for (int i = 0; i <= K && i < n; ++i) r[i] = (float)i * r0;
}

extern "C" int32_t egn_host_sample_schedule(float near_plane, float far_plane, float r0, int32_t n, float* z_out) {
    if (n < 2 || !z_out) return fail("bad arguments");
    const double ratio = exp(log(((double)far_plane - (double)near_plane) / (double)r0) / (double)(n - 1));
    z_out[0] = 0.f;
    for (int i = 1; i < n; ++i) z_out[i] = r0 * powf((float)ratio, (float)(i - 1));
    force_linear_prefix(z_out, n, r0);
    return 0;
}

extern "C" int32_t egn_host_r_knots(float far_r, float r0, int32_t n_r, float* knots) {
    if (n_r < 2 || !knots) return fail("bad arguments");
    const float ratio = powf(far_r / r0, (float)(1.0 / (double)(n_r - 1)));
    knots[0] = 0.f;
    for (int i = 1; i <= n_r; ++i) knots[i] = r0 * powf(ratio, (float)(i - 1));
    force_linear_prefix(knots, n_r + 1, r0);
    return 0;
}

// ---- the same two ladders for a run without --interval_th (opt.py:190) ---------------------------------
extern "C" int32_t egn_host_plain_sample_schedule(float near_plane, float far_plane, int32_t n, float* z_out, float* ratio_out,
                                                  float* r0_out) {
    if (n < 2 || !z_out) return fail("bad arguments");
    const double ratio = 1.0 + (M_PI / 2.0) / (double)n;                                 // EgoNeRF.py:60
    const double r0 = ((double)far_plane - (double)near_plane) * (ratio - 1.0) / (pow(ratio, (double)n) - 1.0);
    float run = 0.f;
    for (int j = 0; j < n; ++j) {                                                        // exclusive running sum of ratio^i
        z_out[j] = run * (float)r0;
        run += powf((float)ratio, (float)j);
    }
    if (ratio_out) *ratio_out = (float)ratio;
    if (r0_out) *r0_out = (float)r0;
    return 0;
}

extern "C" int32_t egn_host_plain_r_knots(float far_r, float r0, int32_t n_r, float* knots) {
    if (n_r < 2 || !knots) return fail("bad arguments");
    const float ratio = powf(far_r / r0, 1.f / (float)(n_r - 1));                        // coordinates.py:139,215
    knots[0] = 0.f;
    for (int i = 1; i <= n_r + 2; ++i) knots[i] = r0 * powf(ratio, (float)(i - 1));
    return 0;
}
