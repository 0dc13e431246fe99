// Forward kernels of the render path (sm_100a):
//   egn_coarse_kernel    sample schedule -> Yin-Yang coords -> pooled density gather -> raw2alpha -> inverse CDF -> sort
//   egn_gather_kernel    fine samples: coords -> 18-tap interleaved gather -> sigma feature + basis contraction
//   egn_composite_kernel feature2density -> warp-scan transmittance -> rgb / depth / bg / env / alpha
// Reference semantics: SURVEY.md Appendix A; citations are into changwoonchoi/EgoNeRF.
#include "egn_device.cuh"
#include "egn_host.h"
#include "egn_shared.cuh"

// -------------------------------------------------------------------------------------------------
// One lane's float4 slice of P_i * L_i (i < 3) for one sample: EgoNeRF.compute_densityfeature /
// compute_appfeature inner product terms (EgoNeRF.py:336-346, 394-412) with F.grid_sample(bilinear,
// zeros padding, align_corners=True) written out as taps.  C = channels per texel of this table set,
// `sub` = which float4 of the texel this lane owns.  All 18 loads are issued before any use.
// -------------------------------------------------------------------------------------------------
// per axis: clamped texel indices and tap weights with the zero padding folded in — an out-of-range tap gets
// weight 0 and reads a clamped in-range texel, so every load is unconditional (no predication, no divergence);
// the arithmetic on in-range taps is unchanged
struct EgnAxisTaps { unsigned j0[3], j1[3]; float wa0[3], wa1[3]; };
__device__ __forceinline__ void egn_axis_taps(const int G[3], const float c[3], EgnAxisTaps& A) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float ix = egn_unnorm(c[a], G[a]);
        const float fl = floorf(ix);
        const float fr = ix - fl;
        // clamp before the int conversion so far-out-of-range samples stay "invalid" instead of wrapping
        const int i0 = (int)fminf(fmaxf(fl, -2.f), (float)G[a] + 1.f);
        A.wa0[a] = ((i0 >= 0) & (i0 < G[a])) ? 1.f - fr : 0.f;
        A.wa1[a] = ((i0 + 1 >= 0) & (i0 + 1 < G[a])) ? fr : 0.f;
        A.j0[a] = (unsigned)min(max(i0, 0), G[a] - 1);
        A.j1[a] = (unsigned)min(max(i0 + 1, 0), G[a] - 1);
    }
}
template <int C>
__device__ __forceinline__ void egn_gather_taps(const float* __restrict__ tab, const long long pofs[3], const long long lofs[3],
                                                const int G[3], const EgnAxisTaps& A, int sub, float4 prod[3]) {
    float4 t[3][4], l[3][2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ax = egn_mx(i), ay = egn_my(i), al = egn_vl(i);
        const unsigned W = (unsigned)G[ax];
        const unsigned pbase = (unsigned)pofs[i] + sub * 4, lbase = (unsigned)lofs[i] + sub * 4;
        const unsigned ra = A.j0[ay] * W, rb = A.j1[ay] * W;
        t[i][0] = ldg4(tab + (size_t)(pbase + (ra + A.j0[ax]) * C));
        t[i][1] = ldg4(tab + (size_t)(pbase + (ra + A.j1[ax]) * C));
        t[i][2] = ldg4(tab + (size_t)(pbase + (rb + A.j0[ax]) * C));
        t[i][3] = ldg4(tab + (size_t)(pbase + (rb + A.j1[ax]) * C));
        l[i][0] = ldg4(tab + (size_t)(lbase + A.j0[al] * C));
        l[i][1] = ldg4(tab + (size_t)(lbase + A.j1[al] * C));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ax = egn_mx(i), ay = egn_my(i), al = egn_vl(i);
        float4 P = f4zero();
        P = f4fma(A.wa0[ax] * A.wa0[ay], t[i][0], P);
        P = f4fma(A.wa1[ax] * A.wa0[ay], t[i][1], P);
        P = f4fma(A.wa0[ax] * A.wa1[ay], t[i][2], P);
        P = f4fma(A.wa1[ax] * A.wa1[ay], t[i][3], P);
        float4 Lv = f4zero();
        Lv = f4fma(A.wa0[al], l[i][0], Lv);
        Lv = f4fma(A.wa1[al], l[i][1], Lv);
        prod[i] = f4mul(P, Lv);
    }
}
template <int C>
__device__ __forceinline__ void egn_gather_products(const float* __restrict__ tab, const long long pofs[3],
                                                    const long long lofs[3], const int G[3], const float c[3],
                                                    int sub, float4 prod[3]) {
    EgnAxisTaps A;
    egn_axis_taps(G, c, A);
    egn_gather_taps<C>(tab, pofs, lofs, G, A, sub, prod);
}

// Bitonic sort of 32 * EF floats held EF per lane in blocked order (element index = lane * EF + e), ascending.
// Compare-exchange partners further than EF apart sit in another lane (one shuffle), closer ones in the same lane.
template <int EF>
__device__ __forceinline__ void egn_bitonic_sort(float (&v)[EF], int lane) {
#pragma unroll
    for (int kk = 2; kk <= 32 * EF; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= EF) {
                const int lm = j / EF;
                const bool lower = (lane & lm) == 0;
#pragma unroll
                for (int e = 0; e < EF; ++e) {
                    const float o = __shfl_xor_sync(FULL, v[e], lm);
                    const bool up = ((lane * EF + e) & kk) == 0;
                    v[e] = (lower == up) ? fminf(v[e], o) : fmaxf(v[e], o);
                }
            } else {
#pragma unroll
                for (int e = 0; e < EF; ++e) {
                    if ((e & j) == 0) {
                        const float a = v[e], b = v[e ^ j];
                        const bool up = ((lane * EF + e) & kk) == 0;
                        v[e] = up ? fminf(a, b) : fmaxf(a, b);
                        v[e ^ j] = up ? fmaxf(a, b) : fminf(a, b);
                    }
                }
            }
        }
    }
}

// t_min of TensorBase.sample_ray (tensorBase.py:312-315): entry depth into the AABB, clamped to [near, far]
__device__ __forceinline__ float egn_aabb_entry(const EgnKernelCfg& k, float ox, float oy, float oz, float dx, float dy, float dz,
                                                float near_plane) {
    const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
    float t = -INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float vec = d[a] == 0.f ? 1e-6f : d[a];
        const float ra = (k.aabb[3 + a] - o[a]) / vec, rb = (k.aabb[a] - o[a]) / vec;
        t = fmaxf(t, fminf(ra, rb));
    }
    return fminf(fmaxf(t, near_plane), k.far_plane);
}

// =================================================================================================
// K1: coarse pass + resampling.  One warp per ray.
// =================================================================================================
#define K1_WARPS 8
#define K1_MAXC 256

#ifndef K1_SHARED_TAPS
#define K1_SHARED_TAPS 1             // per-axis texel indices / tap weights computed by the sample's owner lane and shuffled (coarse grids < 65 536 texels per axis)
#endif
#ifndef K1_MIN_BLOCKS
#define K1_MIN_BLOCKS 2
#endif
// MAXQ = per-lane slots of the scans / sorts: 4 for n_coarse, n_fine <= 128 (every shipped config), 8 up to 256
// Resident CTAs per SM: 4 (64 registers) for MAXQ = 4 -- measured 0.752 ms per 65 536 rays against 0.770 at 2 CTAs / 128 registers
// and 0.802 at 3 (the kernel is bound by ALU latency at few warps; 80 B of spills cost less than the extra warps bring)
#define K1_BLOCKS(MAXQ) ((MAXQ) == 4 ? 2 * K1_MIN_BLOCKS : K1_MIN_BLOCKS)
template <int MAXQ>
__global__ void __launch_bounds__(K1_WARPS * 32, K1_BLOCKS(MAXQ))
egn_coarse_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ rays, long long n, int is_train,
                  const float* __restrict__ u_c, const float* __restrict__ u_f, unsigned long long seed,
                  long long ray0, float near_plane, float* __restrict__ z_out) {
    __shared__ float s_knots[EGN_MAX_KNOTS + 1];
    __shared__ float s_r[K1_MAXC];
    __shared__ float s_zc[K1_WARPS][K1_MAXC];
    __shared__ float s_w[K1_WARPS][K1_MAXC];
    __shared__ float s_zn[K1_WARPS][K1_MAXC];
    const int nc = k.n_coarse, nf = k.n_fine;
    for (int i = threadIdx.x; i <= k.knots_last_c; i += blockDim.x) s_knots[i] = k.r_knots_c[i];
    if (!k.march) for (int i = threadIdx.x; i < nc; i += blockDim.x) s_r[i] = k.z_coarse[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* zc = s_zc[warp];
    float* sw = s_w[warp];
    float* zn = s_zn[warp];
    const int cnt = nc >> 5;                       // coarse samples per lane in the scans (nc % 32 == 0)

    for (long long ray = (long long)blockIdx.x * K1_WARPS + warp; ray < n; ray += (long long)gridDim.x * K1_WARPS) {
        const float ox = rays[ray * 6 + 0], oy = rays[ray * 6 + 1], oz = rays[ray * 6 + 2];
        const float dx = rays[ray * 6 + 3], dy = rays[ray * 6 + 4], dz = rays[ray * 6 + 5];
        // ---- 1. coarse depths ----
        if (!k.march && k.plain_ladders && is_train) {
            // without interval_th the jitter sits in the exponent and accumulates (EgoNeRF.py:59-67):
            // z_j = near + r0 * sum_{i<j} ratio^(i + u_i).  Each lane sums its cnt consecutive terms, a warp scan adds
            // the totals of the lanes before it.
            for (int j = lane; j < nc; j += 32) {
                const float u = u_c ? u_c[ray * nc + j] : egn_u01(egn_philox(seed, (unsigned long long)(ray0 + ray), (unsigned)j).x);
                zn[j] = powf(k.jitter_ratio, (float)j + u);
            }
            __syncwarp();
            float run = 0.f;
            for (int t = 0; t < cnt; ++t) run += zn[lane * cnt + t];
            float before = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_up_sync(FULL, before, o);
                if (lane >= o) before += v;
            }
            before -= run;                                   // exclusive: terms of the lanes before this one
            for (int t = 0; t < cnt; ++t) {
                zc[lane * cnt + t] = near_plane + before * k.jitter_r0;
                before += zn[lane * cnt + t];
            }
        } else if (!k.march) {
            // near + r, jittered by interval*U in train mode (EgoNeRF.py:69-82)
            for (int j = lane; j < nc; j += 32) {
                float r = s_r[j];
                if (is_train) {
                    float iv = (j + 1 < nc) ? (s_r[j + 1] - s_r[j]) : (s_r[nc - 1] - s_r[nc - 2]);
                    float u = u_c ? u_c[ray * nc + j] : egn_u01(egn_philox(seed, (unsigned long long)(ray0 + ray), (unsigned)j).x);
                    r = r + iv * u;
                }
                zc[j] = near_plane + r;
            }
        } else {
            // uniform march (TensorBase.sample_ray, tensorBase.py:308-327): z_j = t_min + stepSize * (j [+ U]).  The coarse
            // POINTS use the ray's own entry depth; in eval mode the depths handed on are those of the FIRST ray of the
            // chunk (EgoNeRF.py:515-516 `coarse_z_vals[0].repeat(N, 1)`) — reproduced as the reference does it.
            const float t_own = egn_aabb_entry(k, ox, oy, oz, dx, dy, dz, near_plane);
            const float t_use = is_train ? t_own
                                         : egn_aabb_entry(k, rays[0], rays[1], rays[2], rays[3], rays[4], rays[5], near_plane);
            for (int j = lane; j < nc; j += 32) {
                float rng = (float)j;
                if (is_train) rng += u_c ? u_c[ray * nc + j] : egn_u01(egn_philox(seed, (unsigned long long)(ray0 + ray), (unsigned)j).x);
                const float step = k.step_size * rng;
                zc[j] = t_use + step;
                zn[j] = t_own + step;             // zn is free until the inverse CDF: depths of the coarse points
            }
        }
        __syncwarp();
        if (!k.resampling) {                      // EgoNeRF.py:564-577: the coarse samples are the samples
            for (int j = lane; j < nc; j += 32) z_out[ray * k.S + j] = zc[j];
            __syncwarp();
            continue;
        }
        // ---- 2. pooled-grid density of every coarse sample (EgoNeRF.py:520-528) ----
        for (int t = 0; t < cnt; ++t) {
            const int j = t * 32 + lane;
            const float z = (k.march && !is_train) ? zn[j] : zc[j];
            YYCoord cc = egn_cart_to_yinyang(ox + dx * z, oy + dy * z, oz + dz * z, k, s_knots, k.knots_last_c, k.r_div_c);
#if K1_SHARED_TAPS
            // texel indices / tap weights of the sample once, in its owner lane (the four lanes of a sample used to repeat
            // this arithmetic); the two indices of an axis travel in one register
            EgnAxisTaps mine;
            egn_axis_taps(k.lay.Gc, cc.c, mine);
            unsigned jj[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) jj[a] = mine.j0[a] | (mine.j1[a] << 16);
#endif
            float myf = 0.f;
#pragma unroll
            for (int p = 0; p < 4; ++p) {         // 8 samples per pass, 4 lanes (one float4 each) per sample
                const int src = p * 8 + (lane >> 2);
                const int yang = __shfl_sync(FULL, cc.yang, src);
                float4 prod[3];
#if K1_SHARED_TAPS
                EgnAxisTaps A;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const unsigned j = __shfl_sync(FULL, jj[a], src);
                    A.j0[a] = j & 0xffffu; A.j1[a] = j >> 16;
                    A.wa0[a] = __shfl_sync(FULL, mine.wa0[a], src);
                    A.wa1[a] = __shfl_sync(FULL, mine.wa1[a], src);
                }
                egn_gather_taps<EGN_CS>(k.tables, k.lay.pc[yang], k.lay.lc[yang], k.lay.Gc, A, lane & 3, prod);
#else
                float c[3];
                c[0] = __shfl_sync(FULL, cc.c[0], src);
                c[1] = __shfl_sync(FULL, cc.c[1], src);
                c[2] = __shfl_sync(FULL, cc.c[2], src);
                egn_gather_products<EGN_CS>(k.tables, k.lay.pc[yang], k.lay.lc[yang], k.lay.Gc, c, lane & 3, prod);
#endif
                float f = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    float s = hsum4(prod[i]);
                    s += __shfl_xor_sync(FULL, s, 1);
                    s += __shfl_xor_sync(FULL, s, 2);
                    f += fmaxf(s, 0.f);
                }
                const float tf = __shfl_sync(FULL, f, (lane & 7) * 4);
                if ((lane >> 3) == p) myf = tf;
            }
            sw[j] = egn_density_act(myf, k.density_shift, k.fea2dense);
        }
        __syncwarp();
        // ---- 3. raw2alpha (tensorBase.py:22-27): lane owns `cnt` consecutive samples, warp product scan ----
        float a_loc[MAXQ], m_loc[MAXQ];
        float prodl = 1.f;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            if (q < cnt) {
                const int j = lane * cnt + q;
                const float dist = ((j + 1 < nc) ? (zc[j + 1] - zc[j]) : (zc[nc - 1] - zc[nc - 2])) * k.distance_scale;
                const float alpha = 1.f - expf(-sw[j] * dist);
                a_loc[q] = alpha;
                m_loc[q] = 1.f - alpha + 1e-10f;
                prodl *= m_loc[q];
            }
        }
        float incl = prodl;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            float o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl *= o;
        }
        float T = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) T = 1.f;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            if (q < cnt) {
                sw[lane * cnt + q] = a_loc[q] * T;     // weights
                T *= m_loc[q];
            }
        }
        __syncwarp();
        // ---- 4. pdf / cdf over weights[1:-1] + 1e-5 (ray_utils.py:159-162) ----
        const int nb = nc - 2;                      // number of pdf entries; cdf has nb + 1 = nc - 1 knots
        float wp[MAXQ];
        float suml = 0.f;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            if (q < cnt) {
                const int kk = lane * cnt + q;
                wp[q] = (kk < nb) ? (sw[kk + 1] + 1e-5f) : 0.f;
                suml += wp[q];
            }
        }
        float tot = suml;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(FULL, tot, d);
        float pl = 0.f;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
            if (q < cnt) { wp[q] = wp[q] / tot; pl += wp[q]; }
        float inc2 = pl;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            float o = __shfl_up_sync(FULL, inc2, d);
            if (lane >= d) inc2 += o;
        }
        float run = inc2 - pl;                      // exclusive prefix of this lane's chunk
        __syncwarp();
        if (lane == 0) sw[0] = 0.f;                 // cdf[0]
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            if (q < cnt) {
                const int kk = lane * cnt + q;
                run += wp[q];
                if (kk < nb) sw[kk + 1] = run;      // cdf[kk+1]
            }
        }
        __syncwarp();
        // ---- 5. inverse CDF (ray_utils.py:164-186) ----
        const int ncdf = nc - 1;
        const float step = 1.0f / (float)(nf - 1);
        for (int i = lane; i < nf; i += 32) {
            float u;
            if (!is_train) u = (i < nf / 2) ? (0.f + step * (float)i) : (1.f - step * (float)(nf - i - 1));   // torch.linspace
            else u = u_f ? u_f[ray * nf + i] : egn_u01(egn_philox(seed, (unsigned long long)(ray0 + ray), 0x10000u + (unsigned)i).x);
            int lo = 0, hi = ncdf;                  // searchsorted(cdf, u, right=True)
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (sw[mid] > u) hi = mid; else lo = mid + 1;
            }
            const int below = max(lo - 1, 0), above = min(lo, ncdf - 1);
            const float c0 = sw[below], c1 = sw[above];
            const float b0 = .5f * (zc[below + 1] + zc[below]), b1 = .5f * (zc[above + 1] + zc[above]);
            float den = c1 - c0;
            if (den < 1e-5f) den = 1.f;
            const float tt = (u - c0) / den;
            zn[i] = b0 + tt * (b1 - b0);
        }
        __syncwarp();
        // ---- 6. sort(cat(coarse, new)) (EgoNeRF.py:536-539).  Only the sorted VALUES are observable, so: the coarse
        // depths are already increasing (monotone schedule; the jitter stays inside its own interval), the fine draws are
        // rank-sorted among themselves (ties broken by position), and the two sorted lists are merged by binary search:
        // position = own rank + number of elements of the other list that come first. ----
        const int na = k.use_coarse_sample ? nc : 0;
        const int S = na + nf;
        const int EF = nf >> 5;
        float vf[MAXQ];
        int rf[MAXQ];
        if (EF == 4 || EF == 8) {
            // power-of-two draw counts (128: every shipped config; 256: the ERP-frame config): already-sorted fast path
            // (eval: u is a linspace and the inverse CDF is monotone), else a bitonic network in registers / shuffles
#pragma unroll
            for (int e = 0; e < MAXQ; ++e) vf[e] = (e < EF) ? zn[lane * EF + e] : 0.f;
            bool sorted_ok = true;
#pragma unroll
            for (int e = 0; e < MAXQ; ++e)
                if (e < EF) { const int idx = lane * EF + e; sorted_ok &= (idx + 1 >= nf) || (vf[e] <= zn[idx + 1]); }
            if (!__all_sync(FULL, sorted_ok)) {
                if (EF == 4) {
                    float w4[4] = {vf[0], vf[1], vf[2], vf[3]};
                    egn_bitonic_sort<4>(w4, lane);
                    vf[0] = w4[0]; vf[1] = w4[1]; vf[2] = w4[2]; vf[3] = w4[3];
                } else if constexpr (MAXQ == 8) {
                    egn_bitonic_sort<8>(vf, lane);
                }
            }
#pragma unroll
            for (int e = 0; e < MAXQ; ++e) rf[e] = lane * EF + e;
        } else {
#pragma unroll
            for (int e = 0; e < MAXQ; ++e) {
                rf[e] = 0;
                vf[e] = (e < EF) ? zn[e * 32 + lane] : 0.f;
            }
            for (int b = 0; b < nf; ++b) {
                const float vb = zn[b];
#pragma unroll
                for (int e = 0; e < MAXQ; ++e)
                    if (e < EF) rf[e] += (vb < vf[e]) || (vb == vf[e] && b < e * 32 + lane);
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < MAXQ; ++e)
            if (e < EF) sw[rf[e]] = vf[e];          // sw (the cdf) is dead: it now holds the fine draws in increasing order
        __syncwarp();
        float* zo = z_out + ray * S;
#pragma unroll
        for (int e = 0; e < MAXQ; ++e) {
            if (e < EF) {                            // fine element: coarse depths <= it come first
                int lo = 0, hi = na;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (zc[mid] <= vf[e]) lo = mid + 1; else hi = mid; }
                zo[rf[e] + lo] = vf[e];
            }
        }
        for (int i = lane; i < na; i += 32) {        // coarse element: fine draws strictly below it come first
            const float v = zc[i];
            int lo = 0, hi = nf;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sw[mid] < v) lo = mid + 1; else hi = mid; }
            zo[i + lo] = v;
        }
        __syncwarp();
    }
}

int egn_launch_coarse(const EgnKernelCfg& k, const float* rays, long long n, int is_train, const float* u_c,
                      const float* u_f, unsigned long long seed, long long ray0, float near_plane, float* z_out,
                      cudaStream_t st) {
    long long blocks = (n + K1_WARPS - 1) / K1_WARPS;
    const bool small = k.n_coarse <= 128 && k.n_fine <= 128;
    const long long cap = 148ll * 4 * (small ? K1_BLOCKS(4) : K1_BLOCKS(8));
    if (blocks > cap) blocks = cap;
    if (small)
        egn_coarse_kernel<4><<<(unsigned)blocks, K1_WARPS * 32, 0, st>>>(k, rays, n, is_train, u_c, u_f, seed, ray0, near_plane, z_out);
    else
        egn_coarse_kernel<8><<<(unsigned)blocks, K1_WARPS * 32, 0, st>>>(k, rays, n, is_train, u_c, u_f, seed, ray0, near_plane, z_out);
    return (int)cudaGetLastError();
}

__device__ __forceinline__ uint32_t egn_pack_bf16x2(float a, float b) {        // a in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// D (16 x 8, fp32) += A (16 x 16, bf16, row) * B (16 x 8, bf16, col)
__device__ __forceinline__ void egn_mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// =================================================================================================
// K2: fine gather.  A warp handles 16 samples per round: lanes 0-15 compute their sample's Yin-Yang
// coordinate; then 8 iterations gather 2 samples each (16 lanes x float4 = one 256-byte tap per
// half-warp); the appearance products go to a per-warp smem tile and are contracted with
// basis_mat_{yin,yang} (EgoNeRF.py:409,412) as a 16 x 28 x 144 register-tiled product.
// =================================================================================================
#define K2_WARPS 16
#define K2_VT 148                 // v-tile row stride (floats): conflict-free float4 rows
#define K2_K (3 * EGN_CA)         // 144

#define K2_BW 76                  // basis row stride in 32-bit words (72 bf16 pairs + pad: conflict-free fragment loads)
struct K2Smem {
    uint32_t Bhi[2][32][K2_BW];   // basis_mat_{yin,yang}.weight rows (output n), bf16x2 pairs along k: hi parts
    uint32_t Blo[2][32][K2_BW];   // lo parts of the split w = hi + lo (two bf16): products carry ~2^-17 relative error
    float vt[K2_WARPS][16][K2_VT];
    float knots[EGN_MAX_KNOTS + 1];
};

template <bool FROM_COORDS>
__global__ void __launch_bounds__(K2_WARPS * 32, 1)
egn_gather_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ basis0,
                  const float* __restrict__ basis1, const float* __restrict__ rays, long long M,
                  const float* __restrict__ zs, const float* __restrict__ coords7, float* __restrict__ fsig,
                  float* __restrict__ feat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K2Smem& sm = *reinterpret_cast<K2Smem*>(smem_raw);
    for (int i = threadIdx.x; i < 2 * 32 * (K2_K / 2); i += blockDim.x) {
        const int h = i / (32 * (K2_K / 2)), o = (i / (K2_K / 2)) % 32, pr = i % (K2_K / 2);
        const float* B = h ? basis1 : basis0;
        const float w0 = (o < k.app_dim && B) ? B[o * K2_K + 2 * pr] : 0.f, w1 = (o < k.app_dim && B) ? B[o * K2_K + 2 * pr + 1] : 0.f;
        const uint32_t hi = egn_pack_bf16x2(w0, w1);
        sm.Bhi[h][o][pr] = hi;
        sm.Blo[h][o][pr] = egn_pack_bf16x2(w0 - __uint_as_float(hi << 16), w1 - __uint_as_float(hi & 0xffff0000u));
    }
    if (!FROM_COORDS)
        for (int i = threadIdx.x; i <= k.knots_last; i += blockDim.x) sm.knots[i] = k.r_knots[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float(*vt)[K2_VT] = sm.vt[warp];
    const long long rounds = (M + 15) / 16;

    for (long long r = (long long)blockIdx.x * K2_WARPS + warp; r < rounds; r += (long long)gridDim.x * K2_WARPS) {
        // ---- a. coordinate of sample (lane & 15) ----
        const long long m = r * 16 + (lane & 15);
        YYCoord cc;
        cc.c[0] = cc.c[1] = cc.c[2] = -3.f;       // out of range -> every tap reads zero
        cc.yang = 0;
        if (m < M) {
            if (FROM_COORDS) {
                const float* c7 = coords7 + m * 7;
                cc.yang = (c7[6] != 0.f);           // EgoNeRF.py:292: last column == 0 selects Yin
                cc.c[0] = c7[cc.yang * 3 + 0]; cc.c[1] = c7[cc.yang * 3 + 1]; cc.c[2] = c7[cc.yang * 3 + 2];
            } else {
                const long long ray = egn_ray_of(m, k.S);
                const float z = zs[m];
                const float* ry = rays + ray * 6;
                cc = egn_cart_to_yinyang(ry[0] + ry[3] * z, ry[1] + ry[4] * z, ry[2] + ry[5] * z, k, sm.knots);
            }
        }
        // ---- b. gather: 2 samples per iteration ----
        float myf = 0.f;
        const int sub = lane & 15;
#pragma unroll 2
        for (int it = 0; it < 8; ++it) {
            const int src = 2 * it + (lane >> 4);
            float c[3];
            c[0] = __shfl_sync(FULL, cc.c[0], src);
            c[1] = __shfl_sync(FULL, cc.c[1], src);
            c[2] = __shfl_sync(FULL, cc.c[2], src);
            const int yang = __shfl_sync(FULL, cc.yang, src);
            float4 prod[3];
            egn_gather_products<EGN_CF>(k.tables, k.lay.pf[yang], k.lay.lf[yang], k.lay.G, c, sub, prod);
            float f = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float s = hsum4(prod[i]);          // only lanes sub < 4 carry density channels
                s += __shfl_xor_sync(FULL, s, 1);
                s += __shfl_xor_sync(FULL, s, 2);
                f += fmaxf(s, 0.f);
                if (sub >= EGN_CS / 4)
                    *reinterpret_cast<float4*>(&vt[src][i * EGN_CA + (sub - EGN_CS / 4) * 4]) = prod[i];
            }
            const float t0 = __shfl_sync(FULL, f, 0), t1 = __shfl_sync(FULL, f, 16);
            if (lane == 2 * it) myf = t0;
            if (lane == 2 * it + 1) myf = t1;
        }
        __syncwarp();
        if (lane < 16 && m < M) fsig[m] = myf;
        // ---- c. basis contraction feat = B_h v (EgoNeRF.py:409,412) on the warp-level tensor-core path: the 16 x 144 product
        // tile of this round times basis^T as m16n8k16 bf16 MMAs with a 3-term split (v = hi + lo, B = hi + lo,
        // D += v_lo B_hi + v_hi B_lo + v_hi B_hi, fp32 accumulate) — fp32-equivalent for the 1e-4 parity bound.  tcgen05
        // needs 128-row CTA-wide tiles; this GEMM is 16 rows per warp, so it uses mma.sync.  Hemispheres that do not
        // occur in the round are skipped; in a mixed round each row keeps the result of its own hemisphere. ----
        if (feat != nullptr) {
            const unsigned ymask = __ballot_sync(FULL, cc.yang != 0) & 0xffffu;
            const int g = lane >> 2, tq = lane & 3;
            for (int h = 0; h < 2; ++h) {
                if (h == 0 && ymask == 0xffffu) continue;     // no Yin sample in this round
                if (h == 1 && ymask == 0u) continue;          // no Yang sample
                float acc[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 3
                for (int kt = 0; kt < K2_K / 16; ++kt) {
                    const int k0 = kt * 16 + 2 * tq;
                    const float2 f0 = *reinterpret_cast<const float2*>(&vt[g][k0]);
                    const float2 f1 = *reinterpret_cast<const float2*>(&vt[g + 8][k0]);
                    const float2 f2 = *reinterpret_cast<const float2*>(&vt[g][k0 + 8]);
                    const float2 f3 = *reinterpret_cast<const float2*>(&vt[g + 8][k0 + 8]);
                    uint32_t ah[4], al[4];
                    ah[0] = egn_pack_bf16x2(f0.x, f0.y); ah[1] = egn_pack_bf16x2(f1.x, f1.y);
                    ah[2] = egn_pack_bf16x2(f2.x, f2.y); ah[3] = egn_pack_bf16x2(f3.x, f3.y);
                    al[0] = egn_pack_bf16x2(f0.x - __uint_as_float(ah[0] << 16), f0.y - __uint_as_float(ah[0] & 0xffff0000u));
                    al[1] = egn_pack_bf16x2(f1.x - __uint_as_float(ah[1] << 16), f1.y - __uint_as_float(ah[1] & 0xffff0000u));
                    al[2] = egn_pack_bf16x2(f2.x - __uint_as_float(ah[2] << 16), f2.y - __uint_as_float(ah[2] & 0xffff0000u));
                    al[3] = egn_pack_bf16x2(f3.x - __uint_as_float(ah[3] << 16), f3.y - __uint_as_float(ah[3] & 0xffff0000u));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = 8 * j + g;
                        const uint32_t bh0 = sm.Bhi[h][n][kt * 8 + tq], bh1 = sm.Bhi[h][n][kt * 8 + tq + 4];
                        const uint32_t bl0 = sm.Blo[h][n][kt * 8 + tq], bl1 = sm.Blo[h][n][kt * 8 + tq + 4];
                        egn_mma_bf16(acc[j], al, bh0, bh1);
                        egn_mma_bf16(acc[j], ah, bl0, bl1);
                        egn_mma_bf16(acc[j], ah, bh0, bh1);
                    }
                }
                // c0,c1: row g, outputs 8j + 2tq, +1;  c2,c3: row g + 8
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = g + 8 * half;
                    const long long mm = r * 16 + row;
                    if (mm < M && (int)((ymask >> row) & 1u) == h) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int n = 8 * j + 2 * tq;
                            if (n < EGN_FEAT_STRIDE)
                                *reinterpret_cast<float2*>(feat + mm * EGN_FEAT_STRIDE + n) = make_float2(acc[j][2 * half], acc[j][2 * half + 1]);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

static int k2_grid(long long M) {
    long long rounds = (M + 15) / 16;
    long long blocks = (rounds + K2_WARPS - 1) / K2_WARPS;
    if (blocks > 148) blocks = 148;               // 1 CTA / SM (187 KB smem), persistent over rounds
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int egn_launch_gather(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                      float* fsig, float* feat, cudaStream_t st) {
    const long long M = n * k.S;
    auto kern = egn_gather_kernel<false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2Smem));
    kern<<<k2_grid(M), K2_WARPS * 32, sizeof(K2Smem), st>>>(k, p->basis[0], p->basis[1], rays, M, z, nullptr, fsig, feat);
    return (int)cudaGetLastError();
}

// coarse stand-alone operator (compute_coarse_densityfeature, EgoNeRF.py:232-289): 4 lanes per sample
__global__ void __launch_bounds__(256)
egn_coarse_feature_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ coords7, long long M,
                          float* __restrict__ fsig) {
    const long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 2;
    const int sub = threadIdx.x & 3;
    float c[3] = {-3.f, -3.f, -3.f};
    int yang = 0;
    if (g < M) {
        const float* c7 = coords7 + g * 7;
        yang = (c7[6] != 0.f);
        c[0] = c7[yang * 3]; c[1] = c7[yang * 3 + 1]; c[2] = c7[yang * 3 + 2];
    }
    float4 prod[3];
    egn_gather_products<EGN_CS>(k.tables, k.lay.pc[yang], k.lay.lc[yang], k.lay.Gc, c, sub, prod);
    float f = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float s = hsum4(prod[i]);
        s += __shfl_xor_sync(FULL, s, 1);
        s += __shfl_xor_sync(FULL, s, 2);
        f += fmaxf(s, 0.f);
    }
    if (g < M && sub == 0) fsig[g] = f;
}

int egn_launch_gather_coords(const EgnKernelCfg& k, const EgnParams* p, const float* coords7, long long m, int coarse,
                             float* fsig, float* feat, cudaStream_t st) {
    if (coarse) {
        long long threads = m * 4;
        egn_coarse_feature_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(k, coords7, m, fsig);
        return (int)cudaGetLastError();
    }
    auto kern = egn_gather_kernel<true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K2Smem));
    kern<<<k2_grid(m), K2_WARPS * 32, sizeof(K2Smem), st>>>(k, p ? p->basis[0] : nullptr, p ? p->basis[1] : nullptr,
                                                              nullptr, m, nullptr, coords7, fsig, feat);
    return (int)cudaGetLastError();
}

// coordinates operator (from_cartesian + normalize_coord, coordinates.py:442-498): one thread per point
__global__ void __launch_bounds__(256)
egn_coords_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ xyz, long long M,
                  float* __restrict__ coords7) {
    const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= M) return;
    YYCoord cc = egn_cart_to_yinyang(xyz[m * 3], xyz[m * 3 + 1], xyz[m * 3 + 2], k, k.r_knots);
    // the inactive hemisphere's raw coordinates are all zero in the reference (coordinates.py:474,487-496);
    // normalising zeros gives the constants below (r=0 -> -1, angle 0 -> (0-near)*inv*2-1)
    const float z1 = (0.f - k.ang_near[0]) * k.ang_inv[0] * 2.f - 1.f;
    const float z2 = (0.f - k.ang_near[1]) * k.ang_inv[1] * 2.f - 1.f;
    float* o = coords7 + m * 7;
    const int a = cc.yang ? 3 : 0, b = cc.yang ? 0 : 3;
    o[a] = cc.c[0]; o[a + 1] = cc.c[1]; o[a + 2] = cc.c[2];
    o[b] = -1.f; o[b + 1] = z1; o[b + 2] = z2;
    o[6] = (float)cc.yang;
}

int egn_launch_coords(const EgnKernelCfg& k, const float* xyz, long long m, float* coords7, cudaStream_t st) {
    egn_coords_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(k, xyz, m, coords7);
    return (int)cudaGetLastError();
}

// =================================================================================================
// Envmap (models/envmap.py:6-34): equirect (3, 2h, h) bilinear + sigmoid.
// =================================================================================================
__global__ void __launch_bounds__(256)
egn_envmap_kernel(int h, const float* __restrict__ em, const float* __restrict__ dirs, long long n, float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o[3];
    egn_env_radiance(em, h, dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2], o);
    out[i * 3] = o[0]; out[i * 3 + 1] = o[1]; out[i * 3 + 2] = o[2];
}
int egn_launch_envmap(int env_h, const float* emission, const float* dirs, long long n, float* out, cudaStream_t st) {
    egn_envmap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(env_h, emission, dirs, n, out);
    return (int)cudaGetLastError();
}

// d_out is the gradient w.r.t. the sigmoid output
__global__ void __launch_bounds__(256)
egn_envmap_bwd_kernel(int h, const float* __restrict__ em, const float* __restrict__ dirs, long long n,
                      const float* __restrict__ d_out, float* __restrict__ d_em) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o[3];
    egn_env_radiance(em, h, dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2], o);
    const EnvTap t = egn_env_tap(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2], h);
    const int W = h, H = 2 * h;
    const float w[4] = {(1.f - t.fx) * (1.f - t.fy), t.fx * (1.f - t.fy), (1.f - t.fx) * t.fy, t.fx * t.fy};
    for (int ch = 0; ch < 3; ++ch) {
        const float g = d_out[i * 3 + ch] * o[ch] * (1.f - o[ch]);
        for (int q = 0; q < 4; ++q) {
            const int x = t.x0 + (q & 1), y = t.y0 + (q >> 1);
            if (x >= 0 && x < W && y >= 0 && y < H) atomicAdd(d_em + ((long long)ch * H + y) * W + x, w[q] * g);
        }
    }
}
int egn_launch_envmap_bwd(int env_h, const float* emission, const float* dirs, long long n, const float* d_out,
                          float* d_emission, cudaStream_t st) {
    egn_envmap_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(env_h, emission, dirs, n, d_out, d_emission);
    return (int)cudaGetLastError();
}

// =================================================================================================
// K4: compositing (tensorBase.py:22-27, EgoNeRF.py:579-598).  One warp per ray, lane owns S/32
// consecutive samples; transmittance = warp-shuffle product scan.
// =================================================================================================
__global__ void __launch_bounds__(256)
egn_composite_kernel(const __grid_constant__ EgnKernelCfg k, const float* __restrict__ emission,
                     const float* __restrict__ rays, long long n, const float* __restrict__ zs,
                     const float* __restrict__ fsig, const float* __restrict__ feat, const float* __restrict__ rgbs,
                     EgnOutputs out, float* __restrict__ wgt, float* __restrict__ bgw, float* __restrict__ rgbpre) {
    const int lane = threadIdx.x & 31;
    const long long ray = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (ray >= n) return;
    const int S = k.S, cnt = S >> 5;
    const float dx = rays[ray * 6 + 3], dy = rays[ray * 6 + 4], dz = rays[ray * 6 + 5];
    float sh[9];
    if (k.shading == EGN_SHADE_SH) egn_sh_basis(dx, dy, dz, sh);
    const long long base = ray * S;
    const int acols = S + (k.env_h > 0 ? 1 : 0);
    float a_loc[K4_MAXE], m_loc[K4_MAXE];
    float prodl = 1.f;
#pragma unroll
    for (int q = 0; q < K4_MAXE; ++q) {
        if (q < cnt) {
            const int j = lane * cnt + q;
            const float zj = zs[base + j];
            const float dist = ((j + 1 < S) ? (zs[base + j + 1] - zj) : (zj - zs[base + j - 1])) * k.distance_scale;
            const float sigma = egn_density_act(fsig[base + j], k.density_shift, k.fea2dense);
            const float alpha = 1.f - expf(-sigma * dist);
            a_loc[q] = alpha;
            m_loc[q] = 1.f - alpha + 1e-10f;
            prodl *= m_loc[q];
            out.alpha[ray * acols + j] = alpha;
        }
    }
    float incl = prodl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl *= o;
    }
    float T = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) T = 1.f;
    const float Tend = __shfl_sync(FULL, incl, 31);       // bg_weight = T[:, -1]
    float acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f;
#pragma unroll
    for (int q = 0; q < K4_MAXE; ++q) {
        if (q < cnt) {
            const int j = lane * cnt + q;
            const float w = a_loc[q] * T;
            T *= m_loc[q];
            float c[3];
            egn_sample_color(k, feat, rgbs, base + j, sh, c);
            acc += w; cr = fmaf(w, c[0], cr); cg = fmaf(w, c[1], cg); cb = fmaf(w, c[2], cb);
            dep = fmaf(w, zs[base + j], dep);
            if (wgt) wgt[base + j] = w;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc += __shfl_xor_sync(FULL, acc, d); cr += __shfl_xor_sync(FULL, cr, d); cg += __shfl_xor_sync(FULL, cg, d);
        cb += __shfl_xor_sync(FULL, cb, d); dep += __shfl_xor_sync(FULL, dep, d);
    }
    if (lane == 0) {
        if (k.env_h > 0) {
            float e[3];
            egn_env_radiance(emission, k.env_h, dx, dy, dz, e);
            const float b0 = Tend * e[0], b1 = Tend * e[1], b2 = Tend * e[2];
            out.env[ray * 3] = e[0]; out.env[ray * 3 + 1] = e[1]; out.env[ray * 3 + 2] = e[2];
            out.bg[ray * 3] = b0; out.bg[ray * 3 + 1] = b1; out.bg[ray * 3 + 2] = b2;
            cr += b0; cg += b1; cb += b2;
            out.alpha[ray * acols + S] = 1.f;       // EgoNeRF.py:587
        }
        if (rgbpre) { rgbpre[ray * 3] = cr; rgbpre[ray * 3 + 1] = cg; rgbpre[ray * 3 + 2] = cb; }   // clamp mask for backward
        out.rgb[ray * 3] = fminf(fmaxf(cr, 0.f), 1.f);
        out.rgb[ray * 3 + 1] = fminf(fmaxf(cg, 0.f), 1.f);
        out.rgb[ray * 3 + 2] = fminf(fmaxf(cb, 0.f), 1.f);
        out.depth[ray] = dep + (1.f - acc) * dz;     // EgoNeRF.py:598: rays_chunk[..., -1] is d_z
        if (bgw) bgw[ray] = Tend;
    }
}

int egn_launch_composite(const EgnKernelCfg& k, const EgnParams* p, const float* rays, long long n, const float* z,
                         const float* fsig, const float* feat, const float* rgbs, const EgnOutputs* out, float* wgt,
                         float* bgw, float* rgbpre, cudaStream_t st) {
    long long threads = n * 32;
    egn_composite_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(k, p->emission, rays, n, z, fsig, feat, rgbs,
                                                                           *out, wgt, bgw, rgbpre);
    return (int)cudaGetLastError();
}

// =================================================================================================
// Equirectangular ray generation on the device: get_ray_directions_360 (dataLoader/ray_utils.py:24-40),
// the normalisation of dataset_omniblender.py:43 and get_rays (ray_utils.py:85-113) for rows [row0, row0 + n_rows) of an
// H x W frame with camera-to-world pose c2w (3 x 4, row-major).  One thread per pixel; replaces the host-side
// meshgrid + matmul + H2D copy of a whole frame of rays (24 B/ray over PCIe) by 48 bytes of pose.
// =================================================================================================
struct EgnPose { float m[12]; };

__global__ void __launch_bounds__(256)
egn_erp_rays_kernel(int H, int W, int row0, long long n, const __grid_constant__ EgnPose pose, float* __restrict__ rays) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int col = (int)(idx % W), row = row0 + (int)(idx / W);
    const float i = (float)col + 0.5f, j = (float)row + 0.5f;
    const float phi = (1.f - 2.f * i / (float)W) * 3.14159265358979323846f;            // longitude (pi, -pi)
    const float theta = (1.f - 2.f * j / (float)H) * 3.14159265358979323846f / 2.f;    // latitude (pi/2, -pi/2)
    float dx = -cosf(theta) * sinf(phi), dy = sinf(theta), dz = -cosf(theta) * cosf(phi);
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= nrm; dy /= nrm; dz /= nrm;
    float* o = rays + idx * 6;
    o[0] = pose.m[3]; o[1] = pose.m[7]; o[2] = pose.m[11];
    o[3] = dx * pose.m[0] + dy * pose.m[1] + dz * pose.m[2];
    o[4] = dx * pose.m[4] + dy * pose.m[5] + dz * pose.m[6];
    o[5] = dx * pose.m[8] + dy * pose.m[9] + dz * pose.m[10];
}

int egn_launch_erp_rays(int H, int W, int row0, int n_rows, const float* c2w_host, float* rays, cudaStream_t st) {
    EgnPose pose;
    for (int i = 0; i < 12; ++i) pose.m[i] = c2w_host[i];
    const long long n = (long long)n_rows * W;
    if (n <= 0) return 0;
    egn_erp_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(H, W, row0, n, pose, rays);
    return (int)cudaGetLastError();
}
