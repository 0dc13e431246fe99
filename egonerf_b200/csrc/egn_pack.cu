// Render-table packing: reference-layout (NCHW) factor tensors -> channels-last interleaved tables, plus
// the 2x average-pooled density tables of the coarse pass (replaces EgoNeRF.update_coarse_sigma_grid,
// models/EgoNeRF.py:124-133), and the inverse scatter of table gradients back to NCHW.
#include "egn_device.cuh"
#include <cuda_bf16.h>
#include "egn_host.h"

struct PackJob {
    const float* src_d;   // density tensor (C=EGN_CS, H, W) or (C, L)
    const float* src_a;   // appearance tensor (C=EGN_CA, ...) ; unused for coarse jobs
    float* dst_d;         // unpack: destination grads
    float* dst_a;
    long long off;        // section offset in the table buffer
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
    int n = 0;
    for (int h = 0; h < 2; ++h)
        for (int i = 0; i < 3; ++i) {
            PackJob a{};   // fine plane
            a.src_d = p ? p->density_plane[h][i] : nullptr; a.src_a = p ? p->app_plane[h][i] : nullptr;
            a.dst_d = g ? g->density_plane[h][i] : nullptr; a.dst_a = g ? g->app_plane[h][i] : nullptr;
            a.off = L.pf[h][i]; a.H = L.G[egn_my(i)]; a.W = L.G[egn_mx(i)]; a.type = 0;
            jobs.j[n++] = a;
            PackJob b{};   // fine line
            b.src_d = p ? p->density_line[h][i] : nullptr; b.src_a = p ? p->app_line[h][i] : nullptr;
            b.dst_d = g ? g->density_line[h][i] : nullptr; b.dst_a = g ? g->app_line[h][i] : nullptr;
            b.off = L.lf[h][i]; b.H = L.G[egn_vl(i)]; b.W = 1; b.type = 0;
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
