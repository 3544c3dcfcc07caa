// dev aid: cycles of ONE warp-cooperative Poseidon permutation (single warp, clock64 around 64 permutations in a row).
// Round 2 measurements with it (B200, 1965 MHz): 546 cycles per round as shipped (24 shuffles + 24 IMAD.WIDE per round),
// 352 with the S-box replaced by x + 1 -- the linear layer, not the S-box, is most of a round; a shared-memory exchange
// (one store, six broadcast loads, four accumulators) measured 585 / 320 and was not kept; two accumulator chains of six in
// the shuffle version: 580.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I qp-zk-circuits_b200/csrc -o permlat tools/ubench/permlat.cu
#include <cstdio>
#include "internal.cuh"
#include "poseidon_rc.cuh"
#include "poseidon.cuh"
using namespace qpzk;
template <int W> __global__ void k(uint64_t* io, long long* cyc) {
    const int lane = threadIdx.x & (W - 1);
    uint64_t s = io[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < 64; i++) s = poseidon_permute_sub<W>(s, lane);
    long long t1 = clock64();
    io[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int W> void run(const char* name) {
    uint64_t h[32]; for (int i = 0; i < 32; i++) h[i] = 0x9E3779B97F4A7C15ull * (i + 1) % GL_P;
    uint64_t* d; long long* c; cudaMalloc(&d, sizeof h); cudaMalloc(&c, 8);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice); k<W><<<1, 32>>>(d, c);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice); k<W><<<1, 32>>>(d, c); cudaDeviceSynchronize();
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    printf("%-6s %8.0f cycles per permutation, %6.1f per round\n", name, cy / 64.0, cy / 64.0 / 30.0);
}
int main() { run<32>("W=32"); run<16>("W=16"); return 0; }
