// Micro-benchmark: issue rate of the sm_100a mixed-precision FMA (fma.rn.f32.bf16 -> SASS FHFMA.BF16) against FFMA and
// against the unpack + FFMA pair it would replace.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fhfma_bench.cu -o fhfma_bench
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8

__device__ __forceinline__ float fhfma_lo(unsigned packed, unsigned w, float acc) {
    float r;
    asm volatile("{\n\t.reg .b16 lo, hi, wl, wh;\n\tmov.b32 {lo, hi}, %1;\n\tmov.b32 {wl, wh}, %2;\n\tfma.rn.f32.bf16 %0, lo, wl, %3;\n\t}\n"
                 : "=f"(r) : "r"(packed), "r"(w), "f"(acc));
    return r;
}

template <int MODE>
__global__ void bench(const unsigned* __restrict__ in, float* out) {
    unsigned v[ILP];
    float acc[ILP];
    for (int i = 0; i < ILP; ++i) { v[i] = in[threadIdx.x + 32 * i]; acc[i] = (float)i; }
    const unsigned wb = in[threadIdx.x] | 0x3f80u;
    const float wf = __uint_as_float(wb << 16);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) acc[i] = fmaf(__uint_as_float(v[i]), wf, acc[i]);                      // FFMA
            if (MODE == 1) acc[i] = fhfma_lo(v[i], wb, acc[i]);                                   // FHFMA.BF16
            if (MODE == 2) acc[i] = fmaf(__uint_as_float(v[i] << 16), wf, acc[i]);               // unpack + FFMA
        }
    }
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(const unsigned* in, float* out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    bench<MODE><<<148 * 4, 512>>>(in, out);
    cudaEventRecord(a);
    for (int r = 0; r < 10; ++r) bench<MODE><<<148 * 4, 512>>>(in, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 10.f;
}

int main() {
    unsigned* in; float* out;
    cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0x3f, 4096 * 4);
    cudaMalloc(&out, 148 * 4 * 512 * 4);
    const double ops = 148.0 * 4 * 512 * ITERS * ILP;
    const float t0 = run<0>(in, out), t1 = run<1>(in, out), t2 = run<2>(in, out);
    printf("FFMA            %.3f ms  %.1f Gop/s\n", t0, ops / t0 * 1e-6);
    printf("FHFMA.BF16      %.3f ms  %.1f Gop/s  (%.2fx the FFMA time)\n", t1, ops / t1 * 1e-6, t1 / t0);
    printf("unpack + FFMA   %.3f ms  %.1f Gop/s  (%.2fx the FFMA time)\n", t2, ops / t2 * 1e-6, t2 / t0);
    return cudaGetLastError() != cudaSuccess;
}
