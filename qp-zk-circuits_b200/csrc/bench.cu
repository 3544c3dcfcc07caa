// Microbenchmarks behind qpzk_bench_imad / qpzk_bench_poseidon: the measured integer-pipe denominators
// bench.py reports next to the HBM roofline ("IMAD_peak is measured by a microbenchmark").
#include "../../include/qpzk.h"
#include "internal.cuh"
#include "poseidon_rc.cuh"
#include "poseidon.cuh"

namespace qpzk {

// 16 independent 32x32+64 multiply-add chains per thread: enough ILP to saturate the pipe at 4 warps/SMSP.
__global__ void __launch_bounds__(256) imad_wide_kernel(uint32_t iters, uint32_t mul, uint64_t* sink) {
    uint64_t acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = threadIdx.x + 0x9E3779B97F4A7C15ull * k;
    uint32_t m = mul | 1u;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t lo = (uint32_t)acc[k];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(lo), "r"(m));
        }
    }
    uint64_t x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= acc[k];
    if (x == 0x1234567ull) sink[0] = x;  // practically never true; keeps the chains alive
}

__global__ void __launch_bounds__(HASH_THREADS) poseidon_bench_kernel(uint32_t perms, uint64_t* sink) {
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = (uint64_t)blockIdx.x * 4096 + threadIdx.x * 12 + i;
    for (uint32_t k = 0; k < perms; k++) poseidon_permute(s);
    uint64_t x = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) x ^= s[i];
    if (x == 0x1234567ull) sink[0] = x;
}

cudaError_t run_imad_bench(uint32_t iters, int num_sms, cudaStream_t st, float* ms, double* ops) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    uint64_t* sink;
    cudaError_t e = cudaMalloc((void**)&sink, 8);
    if (e != cudaSuccess) return e;
    int blocks = num_sms * 8;
    imad_wide_kernel<<<blocks, 256, 0, st>>>(16, 3, sink);  // warm-up
    cudaEventRecord(a, st);
    imad_wide_kernel<<<blocks, 256, 0, st>>>(iters, 3, sink);
    cudaEventRecord(b, st);
    e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, a, b);
    *ops = (double)blocks * 256.0 * iters * 16.0;
    cudaFree(sink); cudaEventDestroy(a); cudaEventDestroy(b);
    return e;
}

cudaError_t run_poseidon_bench(uint32_t perms, int num_sms, cudaStream_t st, float* ms, double* total) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    uint64_t* sink;
    cudaError_t e = cudaMalloc((void**)&sink, 8);
    if (e != cudaSuccess) return e;
    int blocks = num_sms * 16;
    poseidon_bench_kernel<<<blocks, HASH_THREADS, 0, st>>>(1, sink);
    cudaEventRecord(a, st);
    poseidon_bench_kernel<<<blocks, HASH_THREADS, 0, st>>>(perms, sink);
    cudaEventRecord(b, st);
    e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, a, b);
    *total = (double)blocks * HASH_THREADS * perms;
    cudaFree(sink); cudaEventDestroy(a); cudaEventDestroy(b);
    return e;
}

}  // namespace qpzk
