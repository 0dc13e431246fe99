// Render-table packing: reference-layout (NCHW) factor tensors -> channels-last interleaved tables, plus
// the 2x average-pooled density tables of the coarse pass (replaces EgoNeRF.update_coarse_sigma_grid,
// models/EgoNeRF.py:124-133), and the inverse scatter of table gradients back to NCHW.
#include "egn_device.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include "egn_host.h"

struct PackJob {
    const float* src_d;   // density tensor (C=EGN_CS, H, W) or (C, L)
    const float* src_a;   // appearance tensor (C=EGN_CA, ...) ; unused for coarse jobs
    float* dst_d;         // unpack: destination grads
    float* dst_a;
    long long off;        // section offset in the table buffer
    int tex0;             // first global texel of the section in the half tables (EgnLayoutH); fine sections only
    int H, W;             // source extents (W = 1 for lines)
    int type;             // 0 fine plane/line (interleave), 1 coarse plane (2x2 mean), 2 coarse line (2 mean)
};
struct PackJobs { PackJob j[24]; };

__global__ void __launch_bounds__(256) egn_pack_kernel(const __grid_constant__ PackJobs jobs, float* __restrict__ tables) {
    const PackJob& J = jobs.j[blockIdx.y];
    const long long HW = (long long)J.H * J.W;
    float* dst = tables + J.off;
    if (J.type == 0) {
        // one thread = one float4 channel group of one texel: writes are fully coalesced 256-byte runs
        const long long n = HW * (EGN_CF / 4);
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
            const long long texel = t / (EGN_CF / 4);
            const int c = (int)(t % (EGN_CF / 4)) * 4;
            const float* s = (c < EGN_CS) ? J.src_d + (long long)c * HW : J.src_a + (long long)(c - EGN_CS) * HW;
            float4 v = make_float4(s[texel], s[texel + HW], s[texel + 2 * HW], s[texel + 3 * HW]);
            reinterpret_cast<float4*>(dst)[t] = v;
        }
    } else if (J.type == 1) {
        const int Hc = J.H / 2, Wc = J.W / 2;
        const long long n = (long long)Hc * Wc * (EGN_CS / 4);
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
            const long long texel = t / (EGN_CS / 4);
            const int c = (int)(t % (EGN_CS / 4)) * 4;
            const int y = (int)(texel / Wc), x = (int)(texel % Wc);
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* s = J.src_d + (long long)(c + q) * HW + (long long)(2 * y) * J.W + 2 * x;
                o[q] = (((s[0] + s[1]) + s[J.W]) + s[J.W + 1]) * 0.25f;
            }
            reinterpret_cast<float4*>(dst)[t] = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else {
        const int Lc = J.H / 2;
        const long long n = (long long)Lc * (EGN_CS / 4);
        for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
            const int l = (int)(t / (EGN_CS / 4));
            const int c = (int)(t % (EGN_CS / 4)) * 4;
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* s = J.src_d + (long long)(c + q) * J.H + 2 * l;
                o[q] = (s[0] + s[1]) * 0.5f;
            }
            reinterpret_cast<float4*>(dst)[t] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// d_tables (fine sections, [texel][EGN_CF]) -> NCHW gradients (overwrite: every element is produced once).  Thread = (channel, texel), texel fastest.
__global__ void __launch_bounds__(256) egn_unpack_kernel(const __grid_constant__ PackJobs jobs, const float* __restrict__ d_tables) {
    const PackJob& J = jobs.j[blockIdx.y];
    const long long HW = (long long)J.H * J.W;
    const float* src = d_tables + J.off;
    const long long n = HW * EGN_CF;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t / HW);
        const long long texel = t % HW;
        const float g = src[texel * EGN_CF + c];
        if (c < EGN_CS) { if (J.dst_d) J.dst_d[(long long)c * HW + texel] = g; }
        else            { if (J.dst_a) J.dst_a[(long long)(c - EGN_CS) * HW + texel] = g; }
    }
}

static int fill_jobs(const EgnConfig* cfg, const EgnParams* p, const EgnGrads* g, PackJobs& jobs, bool with_coarse) {
    EgnLayout L = egn_make_layout(cfg->grid);
    EgnLayoutH LH = egn_make_layout_h(cfg->grid);
    int n = 0;
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) {
            PackJob a{};   // fine plane
            a.src_d = p ? p->density_plane[h][i] : nullptr; a.src_a = p ? p->app_plane[h][i] : nullptr;
            a.dst_d = g ? g->density_plane[h][i] : nullptr; a.dst_a = g ? g->app_plane[h][i] : nullptr;
            a.off = L.pf[h][i]; a.H = L.G[egn_my(i)]; a.W = L.G[egn_mx(i)]; a.type = 0; a.tex0 = LH.texp[h][i];
            jobs.j[n++] = a;
            PackJob b{};   // fine line
            b.src_d = p ? p->density_line[h][i] : nullptr; b.src_a = p ? p->app_line[h][i] : nullptr;
            b.dst_d = g ? g->density_line[h][i] : nullptr; b.dst_a = g ? g->app_line[h][i] : nullptr;
            b.off = L.lf[h][i]; b.H = L.G[egn_vl(i)]; b.W = 1; b.type = 0; b.tex0 = LH.texl[h][i];
            jobs.j[n++] = b;
            if (with_coarse) {
                PackJob c = a; c.off = L.pc[h][i]; c.type = 1; jobs.j[n++] = c;
                PackJob d = b; d.off = L.lc[h][i]; d.type = 2; jobs.j[n++] = d;
            }
        }
    return n;
}

int egn_launch_pack(const EgnConfig* cfg, const EgnParams* params, float* tables, cudaStream_t st) {
    PackJobs jobs{};
    int n = fill_jobs(cfg, params, nullptr, jobs, true);
    dim3 grid(148 * 2, n);
    egn_pack_kernel<<<grid, 256, 0, st>>>(jobs, tables);
    return (int)cudaGetLastError();
}

int egn_launch_unpack(const EgnConfig* cfg, const float* d_tables, const EgnGrads* grads, cudaStream_t st) {
    PackJobs jobs{};
    int n = fill_jobs(cfg, nullptr, grads, jobs, false);
    dim3 grid(148 * 2, n);
    egn_unpack_kernel<<<grid, 256, 0, st>>>(jobs, d_tables);
    return (int)cudaGetLastError();
}

// bf16 copy of the fine sections (same element offsets) for the throughput-mode gather: one texel = 128 bytes
__global__ void __launch_bounds__(256) egn_pack_bf16_kernel(const float* __restrict__ src, __nv_bfloat162* __restrict__ dst, long long n4) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[t];
        dst[2 * t] = __floats2bfloat162_rn(v.x, v.y);
        dst[2 * t + 1] = __floats2bfloat162_rn(v.z, v.w);
    }
}
int egn_launch_pack_bf16(const EgnConfig* cfg, const float* tables, void* tables_bf16, cudaStream_t st) {
    const EgnLayout L = egn_make_layout(cfg->grid);
    const long long n4 = L.pc[0][0] / 4;                    // the fine sections come first
    egn_pack_bf16_kernel<<<148 * 4, 256, 0, st>>>(tables, reinterpret_cast<__nv_bfloat162*>(tables_bf16), n4);
    return (int)cudaGetLastError();
}

// Gradients that reached the NCHW parameter tensors by another road than egn_render_backward (plain-torch regularisers,
// anything autograd wrote into `.grad`) -> ADDED to the table-layout gradient: the transpose of egn_unpack_kernel.  A tensor
// without such a gradient is passed as NULL and skipped.
__global__ void __launch_bounds__(256) egn_pack_grads_kernel(const __grid_constant__ PackJobs jobs, float* __restrict__ d_tables) {
    const PackJob& J = jobs.j[blockIdx.y];
    if (J.src_d == nullptr && J.src_a == nullptr) return;
    const long long HW = (long long)J.H * J.W;
    float4* dst = reinterpret_cast<float4*>(d_tables + J.off);
    const long long n = HW * (EGN_CF / 4);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long texel = t / (EGN_CF / 4);
        const int c = (int)(t % (EGN_CF / 4)) * 4;
        const float* s = (c < EGN_CS) ? J.src_d : J.src_a;
        if (s == nullptr) continue;
        s += (long long)(c < EGN_CS ? c : c - EGN_CS) * HW;
        float4 v = dst[t];
        v.x += s[texel]; v.y += s[texel + HW]; v.z += s[texel + 2 * HW]; v.w += s[texel + 3 * HW];
        dst[t] = v;
    }
}
int egn_launch_pack_grads(const EgnConfig* cfg, const EgnGrads* grads, float* d_tables, cudaStream_t st) {
    PackJobs jobs{};
    EgnParams src{};
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) {
            src.density_plane[h][i] = grads->density_plane[h][i]; src.density_line[h][i] = grads->density_line[h][i];
            src.app_plane[h][i] = grads->app_plane[h][i]; src.app_line[h][i] = grads->app_line[h][i];
        }
    const int n = fill_jobs(cfg, &src, nullptr, jobs, false);
    dim3 grid(148 * 2, n);
    egn_pack_grads_kernel<<<grid, 256, 0, st>>>(jobs, d_tables);
    return (int)cudaGetLastError();
}

// =================================================================================================
// Regularisers of the factor tensors in table space: total variation of the 12 planes (utils.py:155-171 through
// EgoNeRF.TV_loss_density / TV_loss_app, models/EgoNeRF.py:213-229: sum over planes of 1e-2 * 2 * (h_tv / count_h + w_tv / count_w))
// and the L1 norm of the density planes and lines (EgoNeRF.density_L1, :204-211: sum of mean |x|).  One pass over the fp32 render
// tables: every thread owns one float4 channel group of one texel, reads its four plane neighbours, ADDS the weighted
// gradient to the table-layout gradient that egn_render_backward produced (so the regularisers need no NCHW gradient
// tensors, no transposes and no extra optimiser traffic) and contributes to the three loss values
// (losses[0] = TV_loss_density(reg), [1] = TV_loss_app(reg), [2] = density_L1(), unweighted, as train.py:293-305 logs them).
// =================================================================================================
struct RegWeights { float tv_density, tv_app, l1_density; };

__global__ void __launch_bounds__(256) egn_regularize_kernel(const __grid_constant__ PackJobs jobs, const float* __restrict__ tables,
                                                             float* __restrict__ d_tables, RegWeights rw, float* __restrict__ losses) {
    const PackJob& J = jobs.j[blockIdx.y];
    const int H = J.H, W = J.W;
    const bool plane = W > 1;
    const long long HW = (long long)H * W;
    const float4* src = reinterpret_cast<const float4*>(tables + J.off);
    float4* dst = reinterpret_cast<float4*>(d_tables + J.off);
    // TVLoss: count_h = C (H-1) W, count_w = C H (W-1) with C = channels of the TENSOR (16 density / 48 appearance)
    float acc_tv_d = 0.f, acc_tv_a = 0.f, acc_l1 = 0.f;
    const long long n = HW * (EGN_CF / 4);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long texel = t / (EGN_CF / 4);
        const int cg = (int)(t % (EGN_CF / 4));
        const bool dens = cg < EGN_CS / 4;
        const float C = dens ? (float)EGN_CS : (float)EGN_CA;
        const float4 x = src[t];
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        const float wtv = dens ? rw.tv_density : rw.tv_app;
        if (plane) {
            const int y = (int)(texel / W), xx = (int)(texel % W);
            const float inv_h = 1.f / (C * (float)(H - 1) * (float)W), inv_w = 1.f / (C * (float)H * (float)(W - 1));
            const float kh = wtv * 1e-2f * 2.f * 2.f * inv_h, kw = wtv * 1e-2f * 2.f * 2.f * inv_w;   // d/dx of 1e-2 * 2 * sum(diff^2) / count
            float sh = 0.f, sw = 0.f;
            if (y > 0) {
                const float4 o = src[t - (long long)W * (EGN_CF / 4)];
                const float4 d = make_float4(x.x - o.x, x.y - o.y, x.z - o.z, x.w - o.w);
                sh += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
                g.x += kh * d.x; g.y += kh * d.y; g.z += kh * d.z; g.w += kh * d.w;
            }
            if (y < H - 1) {
                const float4 o = src[t + (long long)W * (EGN_CF / 4)];
                g.x -= kh * (o.x - x.x); g.y -= kh * (o.y - x.y); g.z -= kh * (o.z - x.z); g.w -= kh * (o.w - x.w);
            }
            if (xx > 0) {
                const float4 o = src[t - (EGN_CF / 4)];
                const float4 d = make_float4(x.x - o.x, x.y - o.y, x.z - o.z, x.w - o.w);
                sw += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
                g.x += kw * d.x; g.y += kw * d.y; g.z += kw * d.z; g.w += kw * d.w;
            }
            if (xx < W - 1) {
                const float4 o = src[t + (EGN_CF / 4)];
                g.x -= kw * (o.x - x.x); g.y -= kw * (o.y - x.y); g.z -= kw * (o.z - x.z); g.w -= kw * (o.w - x.w);
            }
            const float tv = 1e-2f * 2.f * (sh * inv_h + sw * inv_w);
            if (dens) acc_tv_d += tv; else acc_tv_a += tv;
        }
        if (dens) {                                            // density_L1: mean |x| over the tensor (planes and lines)
            const float inv_n = 1.f / ((float)EGN_CS * (float)HW);
            acc_l1 += (fabsf(x.x) + fabsf(x.y) + fabsf(x.z) + fabsf(x.w)) * inv_n;
            const float kl = rw.l1_density * inv_n;
            auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
            g.x += kl * sgn(x.x); g.y += kl * sgn(x.y); g.z += kl * sgn(x.z); g.w += kl * sgn(x.w);
        }
        if ((plane && wtv != 0.f) || (dens && rw.l1_density != 0.f)) {
            float4 d = dst[t];
            d.x += g.x; d.y += g.y; d.z += g.z; d.w += g.w;
            dst[t] = d;
        }
    }
    // block reduction of the three loss values, one atomic per block and value
    __shared__ float red[3][8];
    float v[3] = {acc_tv_d, acc_tv_a, acc_l1};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 3 && losses != nullptr) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        if (s != 0.f) atomicAdd(losses + threadIdx.x, s);
    }
}
int egn_launch_regularize(const EgnConfig* cfg, const float* tables, float* d_tables, float tv_density, float tv_app,
                          float l1_density, float* losses, cudaStream_t st) {
    PackJobs jobs{};
    const int n = fill_jobs(cfg, nullptr, nullptr, jobs, false);
    RegWeights rw{tv_density, tv_app, l1_density};
    dim3 grid(148, n);
    egn_regularize_kernel<<<grid, 256, 0, st>>>(jobs, tables, d_tables, rw, losses);
    return (int)cudaGetLastError();
}

// half tables of the fused fine pass (EgnLayoutH): thread = one float4 channel group of one fine texel of the fp32 tables;
// density groups are copied as fp32, appearance groups converted to fp16 (round-to-nearest, saturating), and the 32 bytes of
// padding behind the 48 appearance halfs are zeroed.
__device__ __forceinline__ uint32_t egn_pack_h2(float a, float b) {             // a in the low half
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void egn_store_h_group(unsigned char* __restrict__ tab_h, long long dens_off, long long tex, int cg,
                                                  float4 v) {
    if (cg < EGN_CS / 4) {
        reinterpret_cast<float4*>(tab_h + dens_off)[tex * (EGN_CS / 4) + cg] = v;
    } else {
        uint2* dst = reinterpret_cast<uint2*>(tab_h + tex * (EGN_APP_TEXEL_HALFS * 2)) + (cg - EGN_CS / 4);
        *dst = make_uint2(egn_pack_h2(v.x, v.y), egn_pack_h2(v.z, v.w));
        if (cg >= EGN_CF / 4 - 4) dst[4] = make_uint2(0u, 0u);               // the last four groups also clear the padding
    }
}
__global__ void __launch_bounds__(256) egn_pack_h_kernel(const __grid_constant__ PackJobs jobs, const float* __restrict__ tables,
                                                         unsigned char* __restrict__ tab_h, long long dens_off) {
    const PackJob& J = jobs.j[blockIdx.y];
    const long long n = (long long)J.H * J.W * (EGN_CF / 4);
    const float4* src = reinterpret_cast<const float4*>(tables + J.off);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        egn_store_h_group(tab_h, dens_off, J.tex0 + t / (EGN_CF / 4), (int)(t % (EGN_CF / 4)), src[t]);
}
int egn_launch_pack_h(const EgnConfig* cfg, const float* tables, void* tables_h, cudaStream_t st) {
    PackJobs jobs{};
    const int n = fill_jobs(cfg, nullptr, nullptr, jobs, false);
    const EgnLayoutH LH = egn_make_layout_h(cfg->grid);
    dim3 grid(148 * 2, n);
    egn_pack_h_kernel<<<grid, 256, 0, st>>>(jobs, tables, reinterpret_cast<unsigned char*>(tables_h), LH.dens_byte_offset);
    return (int)cudaGetLastError();
}

// =================================================================================================
// Adam step of the factor tensors in render-table space (optimiser of train.py:172-186 = torch.optim.Adam
// without weight decay / amsgrad).  One pass reads the table-layout gradient egn_render_backward produced (no unpack), the
// moments (kept in table layout) and the current value (the fp32 table IS the parameter, texel-interleaved), and writes
// moments, the fp32 table, its bf16 copy, the half tables of the fused fine pass and — through a shared-memory transpose, coalesced along texels — the NCHW
// parameter tensors.  Replaces egn_unpack_kernel + the framework's Adam + egn_pack_kernel (fine) + egn_pack_bf16_kernel.
// =================================================================================================
struct AdamHp { float lr, beta1, beta2, eps, bc1, bc2_sqrt; };

__global__ void __launch_bounds__(256)
egn_adam_tables_kernel(const __grid_constant__ PackJobs jobs, const float* __restrict__ d_tab, float* __restrict__ m_tab,
                       float* __restrict__ v_tab, float* __restrict__ tab, __nv_bfloat162* __restrict__ tab16,
                       unsigned char* __restrict__ tab_h, long long dens_off, AdamHp hp) {
    __shared__ float tile[EGN_CF][33];
    const PackJob& J = jobs.j[blockIdx.y];
    const long long HW = (long long)J.H * J.W;
    const long long n_tiles = (HW + 31) / 32;
    const float step_size = hp.lr / hp.bc1;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long tex0 = t * 32;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int f = threadIdx.x + 256 * j;             // float4 index inside the 32-texel x 64-channel tile
            const int tx = f >> 4, cg = f & 15;
            if (tex0 + tx < HW) {
                const long long e4 = (J.off + (tex0 + tx) * EGN_CF) / 4 + cg;
                const float4 g = reinterpret_cast<const float4*>(d_tab)[e4];
                float4 m = reinterpret_cast<float4*>(m_tab)[e4], v = reinterpret_cast<float4*>(v_tab)[e4];
                float4 p = reinterpret_cast<float4*>(tab)[e4];
                const float gg[4] = {g.x, g.y, g.z, g.w};
                float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w}, pp[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    mm[q] = mm[q] + (gg[q] - mm[q]) * (1.f - hp.beta1);                 // lerp, as torch's fused kernel
                    vv[q] = hp.beta2 * vv[q] + (1.f - hp.beta2) * gg[q] * gg[q];
                    const float denom = sqrtf(vv[q]) / hp.bc2_sqrt + hp.eps;
                    pp[q] = pp[q] - step_size * (mm[q] / denom);
                    tile[cg * 4 + q][tx] = pp[q];
                }
                reinterpret_cast<float4*>(m_tab)[e4] = make_float4(mm[0], mm[1], mm[2], mm[3]);
                reinterpret_cast<float4*>(v_tab)[e4] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                reinterpret_cast<float4*>(tab)[e4] = make_float4(pp[0], pp[1], pp[2], pp[3]);
                if (tab16) {
                    tab16[2 * e4] = __floats2bfloat162_rn(pp[0], pp[1]);
                    tab16[2 * e4 + 1] = __floats2bfloat162_rn(pp[2], pp[3]);
                }
                if (tab_h) egn_store_h_group(tab_h, dens_off, J.tex0 + tex0 + tx, cg, make_float4(pp[0], pp[1], pp[2], pp[3]));
            }
        }
        __syncthreads();
        // NCHW parameters: warp w writes channels 8w .. 8w+7, lanes run along the 32 texels (128-byte stores)
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (tex0 + lane < HW) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = 8 * w + i;
                float* dst = (c < EGN_CS) ? J.dst_d + (long long)c * HW : J.dst_a + (long long)(c - EGN_CS) * HW;
                dst[tex0 + lane] = tile[c][lane];
            }
        }
        __syncthreads();
    }
}

int egn_launch_adam_tables(const EgnConfig* cfg, const EgnGrads* params_out, const float* d_tables, float* m, float* v,
                           float* tables, void* tables_bf16, void* tables_h, float lr, float beta1, float beta2, float eps,
                           int step, cudaStream_t st) {
    PackJobs jobs{};
    const int n = fill_jobs(cfg, nullptr, params_out, jobs, false);
    AdamHp hp;
    hp.lr = lr; hp.beta1 = beta1; hp.beta2 = beta2; hp.eps = eps;
    hp.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    hp.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    dim3 grid(148, n);
    egn_adam_tables_kernel<<<grid, 256, 0, st>>>(jobs, d_tables, m, v, tables, reinterpret_cast<__nv_bfloat162*>(tables_bf16),
                                                 reinterpret_cast<unsigned char*>(tables_h),
                                                 egn_make_layout_h(cfg->grid).dens_byte_offset, hp);
    int e = (int)cudaGetLastError();
    if (e) return e;
    // pooled coarse tables from the updated density parameters (EgoNeRF.update_coarse_sigma_grid, EgoNeRF.py:124-133)
    PackJobs cj{};
    EgnParams src{};
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) { src.density_plane[h][i] = params_out->density_plane[h][i]; src.density_line[h][i] = params_out->density_line[h][i]; }
    PackJobs all{};
    const int na = fill_jobs(cfg, &src, nullptr, all, true);
    int nc = 0;
    for (int i = 0; i < na; ++i) if (all.j[i].type != 0) cj.j[nc++] = all.j[i];
    dim3 cgrid(148, nc);
    egn_pack_kernel<<<cgrid, 256, 0, st>>>(cj, tables);
    return (int)cudaGetLastError();
}

// ---- coarse-to-fine resampling of one factor tensor (EgoNeRF.up_sampling_VM, models/EgoNeRF.py:415-425) ------------------
// One thread per output texel of an NCHW tensor.  The source position of every output row / column comes from the host
// (texel units): the exponential r ladder of coordinates.py:238-246 for an r axis, j*(L-1)/(L2-1) for an angular axis.
// Taps and weights follow F.grid_sample(bilinear, zeros, align_corners=True): out-of-range taps contribute zero.
__global__ void egn_resample_factor_kernel(const float* __restrict__ src, int C, int H, int W, const float* __restrict__ ypos,
                                           int H2, const float* __restrict__ xpos, int W2, float* __restrict__ dst) {
    const long long total = (long long)C * H2 * W2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x2 = (int)(idx % W2);
    const int y2 = (int)((idx / W2) % H2);
    const int c = (int)(idx / ((long long)W2 * H2));
    const float ix = __ldg(xpos + x2), iy = __ldg(ypos + y2);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx1 = ix - x0f, wx0 = (x0f + 1.f) - ix;
    const float wy1 = iy - y0f, wy0 = (y0f + 1.f) - iy;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float* img = src + (long long)c * H * W;
    auto tap = [&](int y, int x) { return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + (long long)y * W + x) : 0.f; };
    float acc = tap(y0, x0) * (wx0 * wy0);
    acc += tap(y0, x0 + 1) * (wx1 * wy0);
    acc += tap(y0 + 1, x0) * (wx0 * wy1);
    acc += tap(y0 + 1, x0 + 1) * (wx1 * wy1);
    dst[idx] = acc;
}

int egn_launch_resample_factor(const float* src, int C, int H, int W, const float* ypos, int H2, const float* xpos, int W2,
                               float* dst, cudaStream_t st) {
    const long long total = (long long)C * H2 * W2;
    if (total == 0) return 0;
    egn_resample_factor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, C, H, W, ypos, H2, xpos, W2, dst);
    return (int)cudaGetLastError();
}
