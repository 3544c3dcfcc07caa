// Integer-pipe microbenchmarks v2 (sm_100a): MDS-like accumulate chains with loop-variant multiplicands,
// so nothing can be hoisted.  Reports SM-clock cycles per warp-instruction per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP> __global__ void __launch_bounds__(256) k(uint32_t* sink, uint32_t seed, long long* cyc) {
    uint32_t a[16]; uint64_t w[8]; uint32_t v[8], c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 7 + i + seed;
#pragma unroll
    for (int j = 0; j < 8; j++) { w[j] = j; v[j] = j + seed; c0[j] = seed + j; c1[j] = seed * 3 + j; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] ^= it;  // loop-variant multiplicands (16 LOP per 128 measured ops)
#pragma unroll
        for (int i = 0; i < 16; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(a[i]), "n"(29));
                if (OP == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(v[j]) : "r"(a[i]), "n"(29));
                if (OP == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(v[j]) : "r"(a[i]), "r"(c0[j]));
                if (OP == 3) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c0[j]), "+r"(c1[j]) : "r"(a[i]), "r"(a[(i + 1) & 15]));
                if (OP == 4) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(a[i]), "n"(29));
                               asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c0[j]), "+r"(c1[j]) : "r"(a[i]), "r"(a[(i + 1) & 15])); }
                if (OP == 5) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(a[i]), "n"(29));
                               asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c0[j]), "+r"(c1[j]) : "r"(a[i]), "r"(a[(i + 1) & 15]));
                               asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c1[j]), "+r"(c0[j]) : "r"(a[(i + 2) & 15]), "r"(a[(i + 3) & 15])); }
                if (OP == 6) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(v[j]) : "r"(a[i]), "n"(29));
                               asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(c0[j]), "+r"(c1[j]) : "r"(a[i]), "r"(a[(i + 1) & 15])); }
                if (OP == 7) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(a[i]), "r"(c0[j]));   // register x register
                if (OP == 8) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[j]) : "r"(a[i]), "r"(c0[j]));
            }
        }
    }
    long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) x ^= v[j] ^ c0[j] ^ c1[j] ^ (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32);
    if (x == 0x12345) sink[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name, int instr) {
    uint32_t* sink; long long* cyc; cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    for (int wps = 1; wps <= 4; wps *= 2) {   // warps per sub-partition
        int blocks = 148, threads = 32 * 4 * wps;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<OP><<<blocks, threads>>>(sink, 3, cyc);
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(sink, 3, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double wi = (double)wps * ITERS * 128 * instr;
        printf("%-34s warps/SMSP=%d  %7.3f ms  %.2f cycles per warp-instr per SMSP  (%.2f per op group)\n", name, wps, ms, c / wi, c / wi * instr);
    }
}
int main() {
    run<0>("mad.wide.u32 acc (imm)", 1); run<7>("mad.wide.u32 acc (reg)", 1); run<8>("mul.wide.u32", 1);
    run<1>("mad.lo.u32 acc (imm)", 1); run<2>("mad.hi.u32 acc", 1);
    run<3>("add.cc+addc", 2); run<4>("wide + add.cc/addc", 3); run<5>("wide + 2x(add.cc/addc)", 5); run<6>("mad.lo + add.cc/addc", 3);
    return 0;
}
