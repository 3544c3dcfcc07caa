// Integer-pipe microbenchmarks for sm_100a: warp-instruction issue cost of the ops the Goldilocks/Poseidon
// kernels are made of.  Prints cycles per warp-instruction per SM sub-partition (lower = faster).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define CHAINS 16
template <int OP> __global__ void __launch_bounds__(256) k(uint32_t* sink, uint32_t m, long long* cyc) {
    uint32_t a[CHAINS], b[CHAINS]; uint64_t w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = threadIdx.x * 7 + i; b[i] = threadIdx.x + 3 * i; w[i] = a[i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 4) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(m), "r"(m));
            if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m));
            if (OP == 7) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(m));   // WIDE + 2 adds
                           asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(m), "r"(m)); }
            if (OP == 8) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));      // LO + 1 add
                           asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m)); }
            if (OP == 9) asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(a[i]) : "h"((uint16_t)b[i]), "h"((uint16_t)m));
            if (OP == 10) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));     // LO + HI pair
                            asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(a[i])); }
            if (OP == 11) { uint32_t lo = (uint32_t)w[i];
                            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(lo), "r"(m)); }     // dependent wide chain
        }
    }
    long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) x ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (x == 0x12345) sink[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name, int instr_per_chain_step) {
    uint32_t* sink; long long* cyc; cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    int blocks = 148 * 2;  // 2 blocks x 8 warps per SM = 4 warps per sub-partition
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(sink, 3, cyc);
    cudaEventRecord(e0); k<OP><<<blocks, 256>>>(sink, 3, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double warp_instr_per_smsp = 4.0 * ITERS * CHAINS * instr_per_chain_step;  // 4 warps per SMSP
    printf("%-28s %7.3f ms  %9.0f cycles  -> %.2f cycles per warp-instr per SMSP   (%.3e lane-ops/s)\n", name, ms, (double)c,
           c / warp_instr_per_smsp, (double)blocks * 256 * ITERS * CHAINS * instr_per_chain_step / (ms * 1e-3));
}
int main() {
    run<0>("mad.lo.u32", 1); run<1>("mad.hi.u32", 1); run<2>("mad.wide.u32", 1); run<11>("mad.wide.u32 (dep chain)", 1);
    run<3>("add.u32", 1); run<4>("add.cc+addc (2 instr)", 2); run<5>("shf", 1); run<6>("lop3", 1);
    run<7>("wide + 2 add (3 instr)", 3); run<8>("mad.lo + add (2 instr)", 2); run<9>("mul.wide.u16", 1); run<10>("mad.lo + mad.hi (2 instr)", 2);
    return 0;
}
