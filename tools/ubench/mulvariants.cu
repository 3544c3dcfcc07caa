// dev aid: SASS instruction counts of Goldilocks multiply variants (compile to cubin, read with cuobjdump -sass)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I qp-zk-circuits_b200/csrc -cubin -o /tmp/mv.cubin tools/ubench/mulvariants.cu
#include "goldilocks.cuh"
using namespace qpzk;

// wide product, IMAD chain (C)
__device__ __forceinline__ u128w wide_chain(uint64_t a, uint64_t b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint64_t p00 = (uint64_t)a0 * b0;
    uint64_t y = (uint64_t)a0 * b1 + (p00 >> 32);
    uint64_t z = (uint64_t)a1 * b0 + (uint32_t)y;
    uint64_t h = (uint64_t)a1 * b1 + (y >> 32) + (z >> 32);
    u128w r; r.w0 = (uint32_t)p00; r.w1 = (uint32_t)z; r.w2 = (uint32_t)h; r.w3 = (uint32_t)(h >> 32);
    return r;
}
// wide product, IMAD chain (PTX mad.wide, so that ptxas cannot turn it back into adds)
__device__ __forceinline__ u128w wide_chain_ptx(uint64_t a, uint64_t b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    u128w r;
    asm("{\n\t.reg .u64 p00,y,z,h,t;\n\t.reg .u32 x1,y0,y1,z1;\n\t"
        "mul.wide.u32 p00,%4,%6;\n\t"
        "mov.b64 {%0,x1},p00;\n\tcvt.u64.u32 t,x1;\n\t"
        "mad.wide.u32 y,%4,%7,t;\n\t"
        "mov.b64 {y0,y1},y;\n\tcvt.u64.u32 t,y0;\n\t"
        "mad.wide.u32 z,%5,%6,t;\n\t"
        "mov.b64 {%1,z1},z;\n\tcvt.u64.u32 t,y1;\n\t"
        "mad.wide.u32 h,%5,%7,t;\n\t"
        "cvt.u64.u32 t,z1;\n\tadd.u64 h,h,t;\n\t"
        "mov.b64 {%2,%3},h;\n\t}"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.w3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return r;
}
// reduction with w2 * EPS + w0 as one IMAD.WIDE
__device__ __forceinline__ uint64_t reduce_fma_local(u128w x) {
    uint32_t s0, s1;
    asm("{\n\t.reg .u64 u,t;\n\t.reg .u32 u0,u1,v1,k,t0,t1,brw,c,m0,m1;\n\t"
        "cvt.u64.u32 t,%2;\n\t"
        "mad.wide.u32 u,%4,0xffffffff,t;\n\t"
        "mov.b64 {u0,u1},u;\n\t"
        "add.cc.u32 v1,u1,%3;\n\taddc.u32 k,0,0;\n\t"
        "sub.cc.u32 t0,u0,%5;\n\tsubc.cc.u32 t1,v1,0;\n\tsubc.u32 brw,0,0;\n\t"
        "add.u32 c,k,brw;\n\t"
        "sub.u32 m0,0,c;\n\tshr.s32 m1,c,31;\n\t"
        "add.cc.u32 %0,t0,m0;\n\taddc.u32 %1,t1,m1;\n\t}"
        : "=r"(s0), "=r"(s1) : "r"(x.w0), "r"(x.w1), "r"(x.w2), "r"(x.w3));
    return ((uint64_t)s1 << 32) | s0;
}
// canonicalise: r + EPS carries out iff r >= p, and then r + EPS - 2^64 = r - p
__device__ __forceinline__ uint64_t canon_add(uint64_t r) {
    uint32_t r0 = (uint32_t)r, r1 = (uint32_t)(r >> 32), t0, t1, c;
    asm("add.cc.u32 %0,%3,0xffffffff;\n\taddc.cc.u32 %1,%4,0;\n\taddc.u32 %2,0,0;" : "=r"(t0), "=r"(t1), "=r"(c) : "r"(r0), "r"(r1));
    return c ? (((uint64_t)t1 << 32) | t0) : r;
}

#define K2(name, expr) extern "C" __global__ void name(uint64_t* p) { uint64_t a = p[threadIdx.x], b = p[threadIdx.x + 32]; p[threadIdx.x] = (expr); }
K2(k_mulc_old, gl_canon(mulp(a, b)))
K2(k_v1_chainC_oldred, gl_canon(reduce_w(wide_chain(a, b))))
K2(k_v2_chainP_oldred, gl_canon(reduce_w(wide_chain_ptx(a, b))))
K2(k_v3_old_fmared, gl_canon(reduce_fma_local(mul_wide_w(a, b))))
K2(k_v4_chainP_fmared, gl_canon(reduce_fma_local(wide_chain_ptx(a, b))))
K2(k_v5_chainP_fmared_canon2, canon_add(reduce_fma_local(wide_chain_ptx(a, b))))
K2(k_v6_nocanon, reduce_fma_local(wide_chain_ptx(a, b)))
extern "C" __global__ void k_addsub(uint64_t* p) {
    uint64_t a = p[threadIdx.x], b = p[threadIdx.x + 32];
    p[threadIdx.x] = gl_add(a, b);
    p[threadIdx.x + 32] = gl_sub(a, b);
}
// x * 2^K: general product by the constant vs the shift formulation (goldilocks.cuh: gl_mul_pow2c)
#define KP(name, K) extern "C" __global__ void name(uint64_t* p) { p[threadIdx.x] = gl_mul_pow2c<K>(p[threadIdx.x]); }
KP(k_pow2_12, 12) KP(k_pow2_48, 48) KP(k_pow2_72, 72)
extern "C" __global__ void k_pow2_12_general(uint64_t* p) { p[threadIdx.x] = gl_mulc(p[threadIdx.x], 1ull << 12); }
