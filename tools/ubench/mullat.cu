// dev aid: LATENCY (not throughput) of Goldilocks multiply formulations -- one warp, a dependent chain x <- x*x (+k),
// cycles per product by clock64.  The cooperative (one permutation per warp) Poseidon path is bound by exactly this.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I qp-zk-circuits_b200/csrc -o tools/ubench/mullat tools/ubench/mullat.cu
#include <cstdio>
#include "goldilocks.cuh"
using namespace qpzk;

// V1: column sums as 64-bit adds of zero-extended words (lets ptxas pick 3-input adds with two carries), IMAD reduce
__device__ __forceinline__ uint64_t mul_v1(uint64_t a, uint64_t b) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const uint64_t p00 = (uint64_t)a0 * b0, p01 = (uint64_t)a0 * b1, p10 = (uint64_t)a1 * b0, p11 = (uint64_t)a1 * b1;
    const uint64_t c1 = (p00 >> 32) + (uint32_t)p01 + (uint32_t)p10;
    const uint64_t c2 = (p01 >> 32) + (p10 >> 32) + (uint32_t)p11 + (c1 >> 32);
    const uint32_t w0 = (uint32_t)p00, w1 = (uint32_t)c1, w2 = (uint32_t)c2;
    const uint32_t w3 = (uint32_t)((p11 >> 32) + (c2 >> 32));
    // (w1:w0) + w2*E - w3
    const uint64_t u = (uint64_t)w2 * 0xFFFFFFFFu + w0;            // exact
    const uint64_t A = u + ((uint64_t)w1 << 32);
    const bool k = A < u;
    const uint64_t B = A - w3;
    const bool brw = A < (uint64_t)w3;
    uint64_t r = B;
    if (k && !brw) r = B + GL_EPS;
    if (brw && !k) r = B - GL_EPS;
    return r;
}
// V2: the library's chain product + fma reduction (gl_mulc without the canonicalisation)
__device__ __forceinline__ uint64_t mul_v2(uint64_t a, uint64_t b) { return reduce_fma(mul_wide_chain(a, b)); }
// V3: plain product words (PTX), reduction with speculative candidates
__device__ __forceinline__ uint64_t mul_v3(uint64_t a, uint64_t b) {
    const u128w x = mul_wide_w(a, b);
    const uint64_t u = (uint64_t)x.w2 * 0xFFFFFFFFu + x.w0;
    const uint64_t A = u + ((uint64_t)x.w1 << 32);
    const bool k = A < u;
    const uint64_t B = A - x.w3;
    const bool brw = A < (uint64_t)x.w3;
    const uint64_t Bp = B + GL_EPS, Bm = B - GL_EPS;
    return (k != brw) ? (k ? Bp : Bm) : B;
}
// V4: __umul64hi based
__device__ __forceinline__ uint64_t mul_v4(uint64_t a, uint64_t b) {
    const uint64_t lo = a * b, hi = __umul64hi(a, b);
    return gl_reduce128_nc(lo, hi);
}

template <int V> __global__ void k(uint64_t* io, int iters, long long* cyc) {
    uint64_t x = io[threadIdx.x], y = io[threadIdx.x + 32];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (V == 0) x = mulp(x, y);
            if (V == 1) x = mul_v1(x, y);
            if (V == 2) x = mul_v2(x, y);
            if (V == 3) x = mul_v3(x, y);
            if (V == 4) x = mul_v4(x, y);
            if (V == 5) x = sqrp(x);
            if (V == 6) x = gl_mul_nc(x, y);
            if (V == 7) x = gl_canon(mulp(x, y));
        }
    }
    long long t1 = clock64();
    io[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int V> void run(const char* name, uint64_t* ref) {
    uint64_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x9E3779B97F4A7C15ull * (i + 1) % GL_P;
    uint64_t* d; long long* c; cudaMalloc(&d, sizeof h); cudaMalloc(&c, 8);
    const int iters = 4096;
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice); k<V><<<1, 32>>>(d, iters, c);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice); k<V><<<1, 32>>>(d, iters, c); cudaDeviceSynchronize();
    long long cy; uint64_t out[32]; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost); cudaMemcpy(out, d, sizeof out, cudaMemcpyDeviceToHost);
    bool same = true;
    if (ref) { for (int i = 0; i < 32; i++) if (gl_canon(out[i]) != gl_canon(ref[i])) same = false; }
    printf("%-44s %7.1f cycles per dependent product %s\n", name, (double)cy / (iters * 8.0), ref ? (same ? "(matches mulp)" : "(MISMATCH)") : "");
    if (!ref) {}
}
int main() {
    // reference result from mulp
    uint64_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x9E3779B97F4A7C15ull * (i + 1) % GL_P;
    uint64_t* d; long long* c; cudaMalloc(&d, sizeof h); cudaMalloc(&c, 8);
    cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice); k<0><<<1, 32>>>(d, 4096, c); cudaDeviceSynchronize();
    uint64_t ref[32]; cudaMemcpy(ref, d, sizeof ref, cudaMemcpyDeviceToHost);
    run<0>("mulp (PTX carry chains, library)", ref); run<1>("v1 column sums + IMAD reduce + selects", ref);
    run<2>("v2 IMAD chain + reduce_fma", ref); run<3>("v3 PTX product + speculative reduce", ref);
    run<4>("v4 mul.lo/mul.hi 64 + C reduce", ref); run<6>("gl_mul_nc (C product + C reduce)", ref); run<7>("canon(mulp)", ref);
    run<5>("sqrp (different function)", nullptr);
    return 0;
}
