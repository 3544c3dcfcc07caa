// dev aid: is the FP64 pipe of sm_100a a usable third pipe beside ALU and FMA for exact integer adds (< 2^53)?
// Prints cycles per warp-instruction per SM sub-partition for DADD / DFMA alone and interleaved with integer streams.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fp64pipe tools/ubench/fp64pipe.cu -lcudart
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 32768
#define CHAINS 8
template <int OP> __global__ void __launch_bounds__(256) k(double* sink, double m, uint32_t mi, long long* cyc) {
    double d[CHAINS], e[CHAINS]; uint32_t a[CHAINS], b[CHAINS]; uint64_t w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { d[i] = threadIdx.x * 7 + i; e[i] = threadIdx.x + 3 * i; a[i] = threadIdx.x * 5 + i; b[i] = threadIdx.x + i; w[i] = a[i]; }
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0)); long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));
            if (OP == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(e[i]));
            if (OP == 2) { asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));                     // DADD + IADD3
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
            if (OP == 3) { asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));                     // DADD + IADD3 + IMAD
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(mi), "r"(a[i])); }
            if (OP == 4) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                     // IADD3 + IMAD (no fp64)
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(mi), "r"(a[i])); }
            if (OP == 5) { asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));                     // DADD + IMAD.WIDE
                           asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(mi)); }
            if (OP == 6) { asm volatile("add.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));                     // 2 DADD + IADD3 + IMAD
                           asm volatile("add.f64 %0, %0, %1;" : "+d"(e[i]) : "d"(m));
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                           asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(mi), "r"(a[i])); }
            if (OP == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 10) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(mi), "r"(a[i]));
            if (OP == 11) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(mi));
            if (OP == 12) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(b[i]));
            if (OP == 7) asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[i]) : "r"(a[i] + it));               // I2F.F64
            if (OP == 8) { long long v; asm volatile("cvt.rzi.s64.f64 %0, %1;" : "=l"(v) : "d"(d[i])); w[i] += v; } // F2I.S64.F64 (+ add)
        }
    }
    long long t1 = clock64(); unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    double x = 0; uint32_t y = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { x += d[i] + e[i]; y ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32); }
    if (x == 0.12345 || y == 0x12345) sink[0] = x + y;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = (long long)(g1 - g0); }
}
template <int OP> void run(const char* name, int instr) {
    double* sink; long long* cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 16);
    int blocks = 148 * 2;  // 2 blocks x 8 warps per SM = 4 warps per sub-partition
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(sink, 3.0, 3, cyc);
    cudaEventRecord(e0); k<OP><<<blocks, 256>>>(sink, 3.0, 3, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); long long cc[2]; cudaMemcpy(cc, cyc, 16, cudaMemcpyDeviceToHost); long long c = cc[0];
    double steps = 4.0 * ITERS * CHAINS;  // chain steps per SMSP (4 warps)
    printf("%-36s %7.3f ms  %9.0f cycles (%.0f MHz by globaltimer) -> %.2f cycles per chain step per SMSP (%d instr: %.2f cycles per warp-instr)\n", name, ms, (double)c, 1e3 * c / (double)cc[1],
           c / steps, instr, c / steps / instr);
}
int main() {
    run<9>("IADD", 1); run<10>("IMAD", 1); run<11>("IMAD.WIDE", 1); run<12>("IADD dep chain x8 (latency)", 1); run<0>("DADD", 1); run<1>("DFMA", 1); run<4>("IADD + IMAD", 2); run<2>("DADD + IADD", 2); run<3>("DADD + IADD + IMAD", 3);
    run<5>("DADD + IMAD.WIDE", 2); run<6>("2 DADD + IADD + IMAD", 4); run<7>("I2F.F64.U32 (+iadd)", 2); run<8>("F2I.S64.F64 (+add64)", 3);
    return 0;
}
