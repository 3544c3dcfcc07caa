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

// ---- field arithmetic, one operation per thread (qpzk_selftest_field_ops) -------------------------------------------------
// Every device formulation of the Goldilocks arithmetic against chosen operands: the tests drive the rare carry / borrow
// paths of the PTX chains (probability ~2^-32 on random data) with directed values and compare with big-integer results.
__global__ void __launch_bounds__(256) field_ops_kernel(uint32_t op, const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                                        const uint64_t* __restrict__ c, uint64_t count, uint64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t x = a[i], y = b[i], z = c[i];
    uint64_t r = 0;
    switch (op) {
        case 0: r = gl_add(x, y); break;                                   // canonical operands
        case 1: r = gl_sub(x, y); break;
        case 2: r = gl_mulc(x, y); break;                                  // any u64 operands from here on
        case 3: r = gl_canon(mulp(x, y)); break;
        case 4: r = gl_canon(sqrp(x)); break;
        case 5: r = gl_canon(mul_addp(x, y, z)); break;
        case 6: r = gl_canon(gl_mul_nc(x, y)); break;
        case 7: r = gl_canon(gl_reduce128_nc(x, y)); break;                // (lo, hi)
        case 8: { u128w w; w.w0 = (uint32_t)x; w.w1 = (uint32_t)(x >> 32); w.w2 = (uint32_t)y; w.w3 = (uint32_t)(y >> 32);
                  r = gl_canon(reduce_fma(w)); } break;
        case 9: { u128w w; w.w0 = (uint32_t)x; w.w1 = (uint32_t)(x >> 32); w.w2 = (uint32_t)y; w.w3 = (uint32_t)(y >> 32);
                  r = gl_canon(reduce_w(w)); } break;
        case 10: r = gl_canon(add_canon_p(x, y)); break;                   // x any, y canonical
        case 11: r = gl_inv(x); break;                                     // canonical, non-zero
        case 12: { acc160 acc = {0, 0, 0, 0, 0};                            // x*y + y*z + z*x + 29 more copies of x*y: five words in use
                   mac160(acc, x, y); mac160(acc, y, z); mac160(acc, z, x);
                   for (int k = 0; k < 29; k++) mac160(acc, x, y);
                   r = gl_canon(reduce160(acc)); } break;
        case 13: { const ext2 e = ext_mul(ext_make(x, y), ext_make(z, x)); r = (i & 1) ? e.b : e.a; } break;   // canonical operands
        case 14: r = gl_mul_pow2_6j(x, 6u * (unsigned)(y % 16)); break;    // canonical x times 2^(6 (y mod 16)): every in-register twiddle
        default: break;
    }
    out[i] = r;
}

cudaError_t launch_field_ops(uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t count, uint64_t* out, cudaStream_t st) {
    if (!count) return cudaSuccess;
    field_ops_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(op, a, b, c, count, out);
    return cudaGetLastError();
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
