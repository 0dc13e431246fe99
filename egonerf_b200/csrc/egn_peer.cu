// Gradient exchange of ray-sharded training over NVLink peer memory: ONE kernel per step sums the
// table-layout factor gradient (+ the small basis / MLP bucket behind it) of all ranks and leaves the sum in every rank's
// buffer -- the buffer egn_gather_bwd_* scattered into and egn_adam_tables reads next, so nothing is copied or staged.
//
//   rank r owns slice r of the buffer:  sum_p buf_p[slice r]  (peer loads, fixed order p = 0 .. world-1: every element is
//   computed once, by one rank, so all ranks end up with bit-identical sums)  ->  stored into slice r of EVERY rank's buffer
//   (posted peer stores).  Per GPU (world-1)/world of the buffer crosses NVLink in each direction -- the bytes of a ring
//   all-reduce, in one pass by all SMs instead of NCCL's channel-limited ring (measured: profiles/r02_scaling.md).
//
// Cross-GPU ordering: block b of every rank handshakes with block b of every peer through 4-byte flags in peer memory, once
// before the first peer load ("my backward has finished": the kernel runs behind it on the stream) and once after the last
// peer store (fence.sys + release store; the waiting side acquires).  A rank cannot leave the kernel before every peer has
// read its slice and written its sums, so the buffer may be reused right after.  Flags carry a call counter (epoch) and are
// never reset.  No block waits for another block of its own GPU: no co-residency requirement.  A peer that never arrives
// makes the wait time out (EGN_PEER_TIMEOUT_NS) and trap instead of hanging the GPU.
//
// Peer pointers come from cudaIpc handles (egn_peer_export / egn_peer_open): the buffers are plain cudaMalloc allocations
// made by this library (egn_peer_alloc), because IPC handles address whole allocations.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <cuda_runtime.h>
#include "../../include/egn.h"

#define PEER_MAX_WORLD 16
#define PEER_MAX_BLOCKS 256
#define PEER_THREADS 512
#ifndef PEER_UNROLL
#define PEER_UNROLL 2
#endif
#ifndef PEER_ACCESS
#define PEER_ACCESS 0                // 0: ld / st .relaxed.sys (SASS .STRONG.SYS); 1: ld.global.cg / st.global.cg (L2-only, weak) -- ordered by the handshakes
#endif
#ifndef EGN_PEER_TIMEOUT_NS
#define EGN_PEER_TIMEOUT_NS 4000000000ull
#endif

int egn_set_error(const char* fmt, ...);         // egn_abi.cu: thread-local message behind egn_last_error()

struct PeerArgs {
    float* buf[PEER_MAX_WORLD];
    unsigned* flag[PEER_MAX_WORLD];
};

__device__ __forceinline__ void peer_signal(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned peer_poll(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long peer_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// peer data changes between launches and is written by other GPUs: read it past L1, at system scope
__device__ __forceinline__ float4 peer_ld4(const float4* p) {
#if PEER_ACCESS == 1
    return __ldcg(p);
#else
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void peer_st4(float4* p, float4 v) {
#if PEER_ACCESS == 1
    __stcg(p, v);
#else
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}

// flags of one rank: [phase 2][source rank PEER_MAX_WORLD][block PEER_MAX_BLOCKS]
__device__ __forceinline__ void peer_handshake(const PeerArgs& a, int rank, int world, unsigned epoch, int phase) {
    __threadfence_system();                       // this thread's peer stores are performed before the block signals
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int p = threadIdx.x;
        peer_signal(a.flag[p] + (phase * PEER_MAX_WORLD + rank) * PEER_MAX_BLOCKS + blockIdx.x, epoch);
        const unsigned* mine = a.flag[rank] + (phase * PEER_MAX_WORLD + p) * PEER_MAX_BLOCKS + blockIdx.x;
        const unsigned long long t0 = peer_now();
        while ((int)(peer_poll(mine) - epoch) < 0) {
            if (peer_now() - t0 > EGN_PEER_TIMEOUT_NS) {
                printf("egn_peer_allreduce: rank %d block %d: peer %d did not arrive (phase %d, epoch %u)\n", rank, blockIdx.x, p, phase, epoch);
                __trap();
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS, 1)
egn_peer_allreduce_kernel(const __grid_constant__ PeerArgs a, int rank, int world_rt, long long n4, float scale, unsigned epoch) {
    const int world = WORLD > 0 ? WORLD : world_rt;
    peer_handshake(a, rank, world, epoch, 0);
    const long long per = (n4 + world - 1) / world;
    const long long lo = per * rank, hi = (lo + per < n4) ? lo + per : n4;
    const long long stride = (long long)gridDim.x * PEER_THREADS;
    for (long long i0 = lo + (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i0 < hi; i0 += PEER_UNROLL * stride) {
        float4 acc[PEER_UNROLL];
        float4 v[PEER_UNROLL][WORLD > 0 ? WORLD : 1];
        if constexpr (WORLD > 0) {
            // all world x PEER_UNROLL loads are requested before the first add (NVLink round trips overlap)
#pragma unroll
            for (int u = 0; u < PEER_UNROLL; ++u) {
                const long long i = i0 + u * stride;
#pragma unroll
                for (int p = 0; p < WORLD; ++p)
                    v[u][p] = (i < hi) ? peer_ld4(reinterpret_cast<const float4*>(a.buf[p]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < PEER_UNROLL; ++u) {
                acc[u] = v[u][0];
#pragma unroll
                for (int p = 1; p < WORLD; ++p) { acc[u].x += v[u][p].x; acc[u].y += v[u][p].y; acc[u].z += v[u][p].z; acc[u].w += v[u][p].w; }
            }
        } else {
#pragma unroll
            for (int u = 0; u < PEER_UNROLL; ++u) {
                const long long i = i0 + u * stride;
                acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < hi) {
                    acc[u] = peer_ld4(reinterpret_cast<const float4*>(a.buf[0]) + i);
                    for (int p = 1; p < world; ++p) {
                        const float4 t = peer_ld4(reinterpret_cast<const float4*>(a.buf[p]) + i);
                        acc[u].x += t.x; acc[u].y += t.y; acc[u].z += t.z; acc[u].w += t.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PEER_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                const float4 s = make_float4(acc[u].x * scale, acc[u].y * scale, acc[u].z * scale, acc[u].w * scale);
                for (int p = 0; p < world; ++p) peer_st4(reinterpret_cast<float4*>(a.buf[p]) + i, s);
            }
        }
    }
    peer_handshake(a, rank, world, epoch, 1);
}

extern "C" int64_t egn_peer_flag_bytes(void) { return (int64_t)2 * PEER_MAX_WORLD * PEER_MAX_BLOCKS * sizeof(unsigned); }

extern "C" int32_t egn_peer_alloc(int64_t bytes, void** ptr_out) {
    if (!ptr_out || bytes <= 0) return egn_set_error("egn_peer_alloc: bad arguments");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        return egn_set_error("egn_peer_alloc(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
    }
    *ptr_out = p;
    return 0;
}
extern "C" int32_t egn_peer_free(void* ptr) {
    const cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? 0 : egn_set_error("egn_peer_free: %s", cudaGetErrorString(e));
}
extern "C" int32_t egn_peer_export(const void* ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
    if (e != cudaSuccess) return egn_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    memcpy(handle_out, &h, 64);
    return 0;
}
extern "C" int32_t egn_peer_open(const unsigned char handle[64], void** ptr_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return egn_set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    *ptr_out = p;
    return 0;
}
extern "C" int32_t egn_peer_close(void* ptr) {
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? 0 : egn_set_error("cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
}

extern "C" int32_t egn_peer_allreduce(void* const* bufs, void* const* flags, int32_t rank, int32_t world, int64_t n_floats,
                                      float scale, uint32_t epoch, int32_t blocks, void* stream) {
    if (!bufs || !flags) return egn_set_error("egn_peer_allreduce: null pointer tables");
    if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world) return egn_set_error("egn_peer_allreduce: rank %d / world %d unsupported (world <= %d)", rank, world, PEER_MAX_WORLD);
    if (n_floats <= 0 || n_floats % 4) return egn_set_error("egn_peer_allreduce: n_floats=%lld must be a positive multiple of 4", (long long)n_floats);
    if (blocks < 1 || blocks > PEER_MAX_BLOCKS) return egn_set_error("egn_peer_allreduce: blocks=%d outside [1, %d]", blocks, PEER_MAX_BLOCKS);
    if (epoch == 0) return egn_set_error("egn_peer_allreduce: epoch counts calls from 1 (flags start at 0)");
    PeerArgs a;
    memset(&a, 0, sizeof(a));
    for (int p = 0; p < world; ++p) {
        if (!bufs[p] || !flags[p]) return egn_set_error("egn_peer_allreduce: null buffer / flags of rank %d", p);
        if (((uintptr_t)bufs[p]) % 16) return egn_set_error("egn_peer_allreduce: buffer of rank %d is not 16-byte aligned", p);
        a.buf[p] = static_cast<float*>(bufs[p]);
        a.flag[p] = static_cast<unsigned*>(flags[p]);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n4 = n_floats / 4;
    switch (world) {
        case 2: egn_peer_allreduce_kernel<2><<<blocks, PEER_THREADS, 0, st>>>(a, rank, world, n4, scale, epoch); break;
        case 4: egn_peer_allreduce_kernel<4><<<blocks, PEER_THREADS, 0, st>>>(a, rank, world, n4, scale, epoch); break;
        case 8: egn_peer_allreduce_kernel<8><<<blocks, PEER_THREADS, 0, st>>>(a, rank, world, n4, scale, epoch); break;
        default: egn_peer_allreduce_kernel<0><<<blocks, PEER_THREADS, 0, st>>>(a, rank, world, n4, scale, epoch); break;
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : egn_set_error("egn_peer_allreduce launch: %s", cudaGetErrorString(e));
}
