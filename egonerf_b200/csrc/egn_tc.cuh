// tcgen05 / TMEM / mbarrier PTX wrappers and the shared-memory operand layout shared by the tensor-core kernels
// (egn_mlp_tc.cu, egn_fused.cu).  sm_100a only.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "egn_device.cuh"

#define TC_THREADS 256
#define TC_TM 128
#define TC_K1 160                      // 32 elements x 5 values
#define TC_CHUNK 2048                  // bytes of one 8-wide K chunk of a 128-row operand
#define TC_IDESC_128x128 0x08200490u   // kind::f16: D fp32, A/B bf16, both K-major, N = 128, M = 128
#define TC_IDESC_128x64  0x08100490u   // same, N = 64

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes = TC_CHUNK) {
    // start address >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | SWIZZLE_NONE
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}
// the same with an explicit stride between 8-row core matrices (SBO); 0 makes every 8-row group read the same rows
__device__ __forceinline__ uint64_t tc_desc_sbo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
// bounded wait: a lost arrive must not hang the GPU.  try_wait with a suspend-time hint parks the warp in hardware (no issue
// slots burnt, wake-up right after the arrive).  The bound is WALL-CLOCK (%globaltimer, 2 s), not a retry count: the parked
// wait may return early any number of times (other barrier traffic of the CTA, instrumented runs under a profiler), and a
// retry budget would then expire without a lost arrive.  On expiry the caller traps.
// BACKOFF_NS > 0: for waits that are known to be long and not on the critical path (the MLP group of the fused kernel waiting
// for the next gathered tile): the parked wait returns on every barrier event of the CTA (~50 times per tile there), and each
// failed poll costs issue slots the gather warps of the same scheduler need -- sleep between polls instead.
template <uint32_t BACKOFF_NS = 0>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1                                             // keep the wait sites small: their code sits inside the hot loops
    for (uint32_t spin = 0; spin < 400u; ++spin) {           // the common case: a few parked waits, no timer traffic
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        if (done) return true;
        if (BACKOFF_NS) __nanosleep(BACKOFF_NS);
    }
    uint32_t t0;                                             // low word of the ns timer: wraps at 4.29 s, the bound is 2 s
    asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t0));
#pragma unroll 1
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        if (done) return true;
        uint32_t now;
        asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(now));
        if (now - t0 > 2000000000u) return false;
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 4 registers of this thread -> 4 consecutive TMEM columns of its lane (scratch exchange between the warps of a lane quadrant)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(bar) : "memory");
}
// sub-CTA barrier: only the `count` threads of one warp-specialised group take part (ids 1.. ; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi); two values per 32-bit word, first value in the low half
__device__ __forceinline__ uint32_t pack_hi(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_lo(float a, float b, uint32_t hi) {
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    return pack_hi(a - ha, b - hb);
}
template <bool SPLIT>
__device__ __forceinline__ void store_chunk(unsigned char* hi_base, unsigned char* lo_base, int chunk, int row, const float* v) {
    uint4 h;
    h.x = pack_hi(v[0], v[1]); h.y = pack_hi(v[2], v[3]); h.z = pack_hi(v[4], v[5]); h.w = pack_hi(v[6], v[7]);
    *reinterpret_cast<uint4*>(hi_base + chunk * TC_CHUNK + row * 16) = h;
    if (SPLIT) {
        uint4 l;
        l.x = pack_lo(v[0], v[1], h.x); l.y = pack_lo(v[2], v[3], h.y); l.z = pack_lo(v[4], v[5], h.z); l.w = pack_lo(v[6], v[7], h.w);
        *reinterpret_cast<uint4*>(lo_base + chunk * TC_CHUNK + row * 16) = l;
    }
}
__device__ __forceinline__ void store_elem(unsigned char* hi_base, unsigned char* lo_base, bool split, int row, int kk, float x,
                                           int chunk_bytes = TC_CHUNK) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const int off = (kk >> 3) * chunk_bytes + row * 16 + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(hi_base + off) = h;
    if (split) *reinterpret_cast<__nv_bfloat16*>(lo_base + off) = __float2bfloat16_rn(x - __bfloat162float(h));
}

// ---- fp16 operand tiles (fused forward, tcgen05 backward) ------------------------------------------------------------
// two floats -> packed fp16 pair (first value in the low half), round to nearest, saturating instead of overflowing to inf
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// the same with max(x, 0) folded into the conversion (F2FP.RELU): relu(round(x)) == round(relu(x))
__device__ __forceinline__ uint32_t pack_h2_relu(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void store_chunk_h_relu(unsigned char* base, int chunk, int row, const float* v) {
    uint4 h;
    h.x = pack_h2_relu(v[0], v[1]); h.y = pack_h2_relu(v[2], v[3]); h.z = pack_h2_relu(v[4], v[5]); h.w = pack_h2_relu(v[6], v[7]);
    *reinterpret_cast<uint4*>(base + chunk * TC_CHUNK + row * 16) = h;
}
__device__ __forceinline__ void store_chunk_h(unsigned char* base, int chunk, int row, const float* v) {
    uint4 h;
    h.x = pack_h2(v[0], v[1]); h.y = pack_h2(v[2], v[3]); h.z = pack_h2(v[4], v[5]); h.w = pack_h2(v[6], v[7]);
    *reinterpret_cast<uint4*>(base + chunk * TC_CHUNK + row * 16) = h;
}
__device__ __forceinline__ void store_elem_h(unsigned char* base, int row, int kk, float x, int chunk_bytes = TC_CHUNK) {
    *reinterpret_cast<__half*>(base + (kk >> 3) * chunk_bytes + row * 16 + (kk & 7) * 2) = __float2half_rn(x);
}

// kind::f16 instruction descriptor with fp16 A / B instead of bf16 (format fields, bits 7-9 and 10-12, cleared)
#define TC_IDESC_F16(idesc) ((idesc) & ~0x480u)

// Power-of-two scale of the tcgen05 backward's GRADIENT operands (fp16 has 11 significand bits against bf16's 8, but only
// 2^-14 .. 2^16 of normal range): S brings the largest |d(sample colour)| of the launch -- written by
// egn_composite_bwd_kernel as the bit pattern of a non-negative float -- to (2^9, 2^10]; everything the kernels derive from
// it is linear in it, the fp32 results are multiplied by 1/S.  Headroom above: 64x for the MLP's gain (conversions saturate
// instead of overflowing); below: full precision down to 2^-24 of the launch maximum, gradual underflow down to 2^-34.
__device__ __forceinline__ float tc_grad_scale(const unsigned* __restrict__ gmax_bits, float& inv) {
    const float gmax = __uint_as_float(__ldg(gmax_bits));
    float e = 0.f;
    if (gmax > 0.f && gmax < 3.0e38f) e = fminf(fmaxf(10.f - ceilf(log2f(gmax)), -100.f), 100.f);
    inv = exp2f(-e);
    return exp2f(e);
}

// original input index (tensorBase.py:68-74 order) of our K position kk; -1 = padding; -2 = the constant-1 column that
// carries the layer-1 bias (first value of the first padding element)
__device__ __forceinline__ int tc_input_index(int kk, int AD) {
    const int e = kk / 5, r = kk % 5;
    if (e == AD + 3 && r == 0) return -2;
    if (e >= AD + 3) return -1;
    const int off_fs = AD + 3, off_fc = off_fs + 2 * AD, off_vs = off_fc + 2 * AD, off_vc = off_vs + 6;
    if (e < AD) return r == 0 ? e : ((r & 1) ? off_fs : off_fc) + e * 2 + ((r - 1) >> 1);
    const int d = e - AD;
    return r == 0 ? AD + d : ((r & 1) ? off_vs : off_vc) + d * 2 + ((r - 1) >> 1);
}


// ---- operand descriptors for canonical no-swizzle tiles --------------------------------------------------------------
// K-major canonical tile read as stored (LBO = chunk stride) or as its transpose (MN-major view of the same bytes:
// LBO = 128 B between 8-row groups, SBO = chunk stride between 8-column groups)
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr, uint32_t chunk = TC_CHUNK) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(chunk >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t chunk = TC_CHUNK) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(chunk >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
