// Goldilocks field (p = 2^64 - 2^32 + 1) and its quadratic extension F[X]/(X^2 - 7) for sm_100a.
//
// Replaces, on the device, qp-plonky2-field =1.5.5 `GoldilocksField` / `QuadraticExtension`
// (reference Cargo.toml:20, Cargo.lock:888-903; type aliases F, D=2 at common/src/circuit.rs:16-18;
// SURVEY.md 8a row a1, Appendix A.1).  The crate is not vendored in the reference tree, so the
// functions below implement the published arithmetic: 64x64->128 products built from four
// 32x32->64 IMAD.WIDE, and the special-form reduction 2^64 = 2^32 - 1, 2^96 = -1 (mod p).
//
// Representation: a field element is a uint64_t.  "canonical" means < p.  Functions whose name ends
// in _nc accept/return arbitrary u64 representatives (value mod p is what matters); everything
// that is stored to global memory or compared is canonicalised first.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// JIT-BEGIN  (the text from here to JIT-END is also the prelude of run-time compiled gate programs, gate_jit.cu)
namespace qpzk {

constexpr uint64_t GL_P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t GL_EPS = 0xFFFFFFFFULL;  // 2^64 mod p
constexpr uint64_t GL_GENERATOR = 14293326489335486720ULL;        // multiplicative generator == coset shift
constexpr uint64_t GL_POWER_OF_TWO_GENERATOR = 7277203076849721926ULL;  // order 2^32
constexpr uint64_t GL_EXT_W = 7;

__host__ __device__ __forceinline__ uint64_t gl_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }

// a + b for arbitrary u64 representatives; result is an arbitrary u64 representative.
__device__ __forceinline__ uint64_t gl_add_nc(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    if (s < a) {            // wrapped: add 2^64 mod p
        s += GL_EPS;
        if (s < GL_EPS) s += GL_EPS;  // second wrap is only possible when both inputs were >= p
    }
    return s;
}
// canonical in -> canonical out.  Device versions are carry chains without compares or selects (the NTT kernels are
// bound by ALU instruction issue): a - b wraps exactly when a < b, and adding p back is subtracting 2^32 - 1 mod 2^64;
// a + b is a - (p - b).  Only sub-flavoured carry ops are chained (see limb_combine in poseidon.cuh for why).
__host__ __device__ __forceinline__ uint64_t gl_sub(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32), d0, d1;
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0,%2,%4;\n\tsubc.cc.u32 %1,%3,%5;\n\tsubc.u32 m,0,0;\n\t"   // m = borrow ? 0xffffffff : 0
        "sub.cc.u32 %0,%0,m;\n\tsubc.u32 %1,%1,0;\n\t}"
        : "=&r"(d0), "=&r"(d1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((uint64_t)d1 << 32) | d0;
#else
    return a >= b ? a - b : a + (GL_P - b);
#endif
}
__host__ __device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32), d0, d1;
    asm("{\n\t.reg .u32 m,n0,n1;\n\t"
        "sub.cc.u32 n0,1,%4;\n\tsubc.u32 n1,0xffffffff,%5;\n\t"                  // p - b  (= p for b = 0: still fine)
        "sub.cc.u32 %0,%2,n0;\n\tsubc.cc.u32 %1,%3,n1;\n\tsubc.u32 m,0,0;\n\t"
        "sub.cc.u32 %0,%0,m;\n\tsubc.u32 %1,%1,0;\n\t}"
        : "=&r"(d0), "=&r"(d1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((uint64_t)d1 << 32) | d0;
#else
    uint64_t s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
#endif
}
__host__ __device__ __forceinline__ uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

// (lo + 2^64 * hi) mod p as an arbitrary u64 representative.
__device__ __forceinline__ uint64_t gl_reduce128_nc(uint64_t lo, uint64_t hi) {
    uint32_t hh = (uint32_t)(hi >> 32), hl = (uint32_t)hi;
    uint64_t t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;                       // borrowed 2^64 = EPS (mod p)
    uint64_t t1 = ((uint64_t)hl << 32) - hl;          // hl * (2^32 - 1), no overflow
    uint64_t r = t0 + t1;
    if (r < t0) r += GL_EPS;                         // cannot wrap twice (t1 <= 2^64 - 2^33 + 1)
    return r;
}

// Full 128-bit product from four 32x32->64 multiply-adds, none of which can overflow 64 bits.
__device__ __forceinline__ void gl_mul_wide(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint64_t p00 = (uint64_t)a0 * b0;
    uint64_t y = (uint64_t)a0 * b1 + (p00 >> 32);
    uint64_t z = (uint64_t)a1 * b0 + (uint32_t)y;
    hi = (uint64_t)a1 * b1 + (y >> 32) + (z >> 32);
    lo = (z << 32) | (uint32_t)p00;
}
__device__ __forceinline__ void gl_sqr_wide(uint64_t a, uint64_t& lo, uint64_t& hi) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    uint64_t p00 = (uint64_t)a0 * a0;
    uint64_t m = (uint64_t)a0 * a1;                   // cross term, counted twice
    uint64_t p11 = (uint64_t)a1 * a1;
    uint64_t y = m + (p00 >> 32);
    uint64_t z = m + (uint32_t)y;
    hi = p11 + (y >> 32) + (z >> 32);
    lo = (z << 32) | (uint32_t)p00;
}

__device__ __forceinline__ uint64_t gl_mul_nc(uint64_t a, uint64_t b) {
    uint64_t lo, hi;
    gl_mul_wide(a, b, lo, hi);
    return gl_reduce128_nc(lo, hi);
}
__device__ __forceinline__ uint64_t gl_sqr_nc(uint64_t a) {
    uint64_t lo, hi;
    gl_sqr_wide(a, lo, hi);
    return gl_reduce128_nc(lo, hi);
}
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b);

// ---- 128-bit product / reduction in PTX ---------------------------------------------------------
struct u128w { uint32_t w0, w1, w2, w3; };

__device__ __forceinline__ u128w mul_wide_w(uint64_t a, uint64_t b) {
    u128w r;
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    asm("{\n\t.reg .u64 p00,p01,p10,p11;\n\t.reg .u32 x1,y0,y1,z0,z1,v0,v1;\n\t"
        "mul.wide.u32 p00,%4,%6;\n\tmul.wide.u32 p01,%4,%7;\n\tmul.wide.u32 p10,%5,%6;\n\tmul.wide.u32 p11,%5,%7;\n\t"
        "mov.b64 {%0,x1},p00;\n\tmov.b64 {y0,y1},p01;\n\tmov.b64 {z0,z1},p10;\n\tmov.b64 {v0,v1},p11;\n\t"
        "add.cc.u32 %1,x1,y0;\n\taddc.cc.u32 %2,y1,v0;\n\taddc.u32 %3,v1,0;\n\t"
        "add.cc.u32 %1,%1,z0;\n\taddc.cc.u32 %2,%2,z1;\n\taddc.u32 %3,%3,0;\n\t}"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.w3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return r;
}

__device__ __forceinline__ u128w sqr_wide_w(uint64_t a) {
    u128w r;
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    asm("{\n\t.reg .u64 p00,pm,p11;\n\t.reg .u32 x1,y0,y1,v0,v1;\n\t"
        "mul.wide.u32 p00,%4,%4;\n\tmul.wide.u32 pm,%4,%5;\n\tmul.wide.u32 p11,%5,%5;\n\t"
        "mov.b64 {%0,x1},p00;\n\tmov.b64 {y0,y1},pm;\n\tmov.b64 {v0,v1},p11;\n\t"
        "add.cc.u32 %1,x1,y0;\n\taddc.cc.u32 %2,y1,v0;\n\taddc.u32 %3,v1,0;\n\t"
        "add.cc.u32 %1,%1,y0;\n\taddc.cc.u32 %2,%2,y1;\n\taddc.u32 %3,%3,0;\n\t}"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.w3) : "r"(a0), "r"(a1));
    return r;
}

// (w3:w2:w1:w0) mod p as an arbitrary u64 representative:  lo - w3 + w2*(2^32-1), with the borrow of the
// subtraction and the carry of the addition folded into one correction of (carry - borrow) * (2^32-1).
__device__ __forceinline__ uint64_t reduce_w(u128w x) {
    uint32_t s0, s1;
    asm("{\n\t.reg .u32 t0,t1,brw,u0,u1,k,m0,m1;\n\t"
        "sub.cc.u32 t0,%2,%5;\n\tsubc.cc.u32 t1,%3,0;\n\tsubc.u32 brw,0,0;\n\t"
        "sub.cc.u32 u0,0,%4;\n\tsubc.u32 u1,%4,0;\n\t"
        "add.cc.u32 t0,t0,u0;\n\taddc.cc.u32 t1,t1,u1;\n\taddc.u32 k,brw,0;\n\t"
        "sub.u32 m0,0,k;\n\tshr.s32 m1,k,31;\n\t"
        "add.cc.u32 %0,t0,m0;\n\taddc.u32 %1,t1,m1;\n\t}"
        : "=r"(s0), "=r"(s1) : "r"(x.w0), "r"(x.w1), "r"(x.w2), "r"(x.w3));
    return ((uint64_t)s1 << 32) | s0;
}

__device__ __forceinline__ uint64_t mulp(uint64_t a, uint64_t b) { return reduce_w(mul_wide_w(a, b)); }
__device__ __forceinline__ uint64_t sqrp(uint64_t a) { return reduce_w(sqr_wide_w(a)); }

// s + a*b  (mod p), s and the product arbitrary representatives: s is added into the 128-bit product
// (cannot overflow 128 bits) before the single reduction.
__device__ __forceinline__ uint64_t mul_addp(uint64_t a, uint64_t b, uint64_t s) {
    u128w r = mul_wide_w(a, b);
    uint32_t s0 = (uint32_t)s, s1 = (uint32_t)(s >> 32);
    asm("add.cc.u32 %0,%0,%4;\n\taddc.cc.u32 %1,%1,%5;\n\taddc.cc.u32 %2,%2,0;\n\taddc.u32 %3,%3,0;"
        : "+r"(r.w0), "+r"(r.w1), "+r"(r.w2), "+r"(r.w3) : "r"(s0), "r"(s1));
    return reduce_w(r);
}

// x + k with k canonical (< p): one conditional +EPS is enough
__device__ __forceinline__ uint64_t add_canon_p(uint64_t x, uint64_t k) {
    uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), k0 = (uint32_t)k, k1 = (uint32_t)(k >> 32), s0, s1;
    asm("{\n\t.reg .u32 c,m;\n\t"
        "add.cc.u32 %0,%2,%4;\n\taddc.cc.u32 %1,%3,%5;\n\taddc.u32 c,0,0;\n\t"
        "sub.u32 m,0,c;\n\t"
        "add.cc.u32 %0,%0,m;\n\taddc.u32 %1,%1,0;\n\t}"
        : "=&r"(s0), "=&r"(s1) : "r"(x0), "r"(x1), "r"(k0), "r"(k1));
    return ((uint64_t)s1 << 32) | s0;
}

// 160-bit accumulator for dot products of full-width operands
struct acc160 { uint32_t a0, a1, a2, a3, a4; };

__device__ __forceinline__ void mac160(acc160& acc, uint64_t a, uint64_t b) {
    u128w r = mul_wide_w(a, b);
    asm("add.cc.u32 %0,%0,%5;\n\taddc.cc.u32 %1,%1,%6;\n\taddc.cc.u32 %2,%2,%7;\n\taddc.cc.u32 %3,%3,%8;\n\taddc.u32 %4,%4,0;"
        : "+r"(acc.a0), "+r"(acc.a1), "+r"(acc.a2), "+r"(acc.a3), "+r"(acc.a4)
        : "r"(r.w0), "r"(r.w1), "r"(r.w2), "r"(r.w3));
}

// acc mod p;  a4 * 2^128 = -a4 * 2^32 (mod p), a4 < 2^5
__device__ __forceinline__ uint64_t reduce160(const acc160& acc) {
    u128w x; x.w0 = acc.a0; x.w1 = acc.a1; x.w2 = acc.a2; x.w3 = acc.a3;
    uint64_t r = reduce_w(x);
    uint64_t sub = (uint64_t)acc.a4 << 32;
    uint64_t d = r - sub;
    if (r < sub) d -= GL_EPS;  // borrowed 2^64 = EPS (mod p); d >= 2^64 - 2^37 here, no second borrow
    return d;
}

// ---- the multiply of the ALU-bound kernels (NTT / LDE / FRI / quotient) ------------------------------------------
// ncu, round 2: the LDE passes keep the ALU pipe 80 % busy with the FMA pipe at 12 %, and a product is two thirds of
// their ALU work.  mulp() above is tuned for the Poseidon kernels, where both pipes are loaded equally.  Here the
// partial products are chained through the 64-bit addend of IMAD.WIDE (a0*b1 + hi(a0*b0) cannot overflow) and
// w2 * (2^32 - 1) + w0 of the reduction is one more IMAD.WIDE: 16 ALU + 4 wide + 9 short FMA-pipe instructions
// instead of 19 + 4 + 3 (tools/ubench/mulvariants.cu counts them).
__device__ __forceinline__ u128w mul_wide_chain(uint64_t a, uint64_t b) {
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
// (w3:w2:w1:w0) mod p as an arbitrary u64 representative:  u = w2 * EPS + w0 (< 2^64 - 2^32, exact), then
// u + w1 * 2^32 - w3 with the carry k of the addition and the borrow of the subtraction folded into one correction
// of (k - borrow) * EPS, which cannot wrap a second time (k = 1: the wrapped sum is <= 2^64 - 2^33; borrow without k:
// the wrapped difference is > 2^64 - 2^32).
__device__ __forceinline__ uint64_t reduce_fma(u128w x) {
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
#ifndef QPZK_MULC_FMA
#define QPZK_MULC_FMA 1
#endif
// canonical product (used by the NTT / FRI / permutation-argument kernels)
__device__ __forceinline__ uint64_t gl_mulc(uint64_t a, uint64_t b) {
#if QPZK_MULC_FMA
    return gl_canon(reduce_fma(mul_wide_chain(a, b)));
#else
    return gl_canon(mulp(a, b));
#endif
}
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_mulc(a, b); }

// ---- x * 2^K for a compile-time K < 96 (the in-register twiddles of the NTT kernels: roots of unity of order <= 64 are
// powers of two) ----------------------------------------------------------------------------------------------------
// K = 32 q + r.  (v2 : v1 : v0) = x << r comes from two IMAD.WIDE by 2^r (the multiplier is read from constant memory:
// as an immediate ptxas turns the product into shifts on the ALU pipe, which is the pipe these kernels are bound by);
// then with 2^64 = EPS, 2^96 = -1, 2^128 = -2^32 (mod p):
//   q = 0:  (v1 : v0) + v2 * EPS          -> one IMAD.WIDE (v2 * EPS + v0, exact), one add with carry, one +EPS on wrap
//   q = 1:  v0 * 2^32 + v1 * EPS - v2     -> the general 128-bit reduction of (v2 : v1 : v0 : 0)
//   q = 2:  v0 * EPS - (v2 : v1)          -> one IMAD.WIDE, one 64-bit subtraction, -EPS on borrow
// 9 / 14 / 9 ALU-pipe instructions and 3 wide multiplies instead of 16 + 5 for the general product (canonical in and out).
__constant__ uint32_t c_pow2_shift[32] = {1u << 0, 1u << 1, 1u << 2, 1u << 3, 1u << 4, 1u << 5, 1u << 6, 1u << 7, 1u << 8, 1u << 9, 1u << 10,
                                          1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20,
                                          1u << 21, 1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30,
                                          1u << 31};
template <unsigned K>
__device__ __forceinline__ uint64_t gl_mul_pow2c(uint64_t x) {
    static_assert(K < 96, "2^96 = -1: fold the sign first");
    constexpr unsigned r = K % 32, q = K / 32;
    const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32);
    uint32_t v0, v1, v2;
    if (r == 0) {
        v0 = x0; v1 = x1; v2 = 0;
    } else {
        const uint32_t R = c_pow2_shift[r];
        asm("{\n\t.reg .u64 t,u,a;\n\t.reg .u32 t1;\n\t"
            "mul.wide.u32 t,%3,%5;\n\tmov.b64 {%0,t1},t;\n\tcvt.u64.u32 a,t1;\n\t"
            "mad.wide.u32 u,%4,%5,a;\n\tmov.b64 {%1,%2},u;\n\t}"
            : "=r"(v0), "=r"(v1), "=r"(v2) : "r"(x0), "r"(x1), "r"(R));
    }
    uint32_t s0, s1;
    if (q == 0) {
        asm("{\n\t.reg .u64 w,a;\n\t.reg .u32 w0,w1,k,m;\n\t"
            "cvt.u64.u32 a,%2;\n\tmad.wide.u32 w,%4,0xffffffff,a;\n\tmov.b64 {w0,w1},w;\n\t"
            "add.cc.u32 w1,w1,%3;\n\taddc.u32 k,0,0;\n\tsub.u32 m,0,k;\n\t"
            "add.cc.u32 %0,w0,m;\n\taddc.u32 %1,w1,0;\n\t}"
            : "=r"(s0), "=r"(s1) : "r"(v0), "r"(v1), "r"(v2));
        return gl_canon(((uint64_t)s1 << 32) | s0);
    } else if (q == 1) {
        u128w w; w.w0 = 0; w.w1 = v0; w.w2 = v1; w.w3 = v2;
        return gl_canon(reduce_fma(w));
    } else {
        asm("{\n\t.reg .u64 w;\n\t.reg .u32 w0,w1,b;\n\t"
            "mul.wide.u32 w,%2,0xffffffff;\n\tmov.b64 {w0,w1},w;\n\t"
            "sub.cc.u32 w0,w0,%3;\n\tsubc.cc.u32 w1,w1,%4;\n\tsubc.u32 b,0,0;\n\t"      // b = borrow ? 0xffffffff : 0
            "sub.cc.u32 %0,w0,b;\n\tsubc.u32 %1,w1,0;\n\t}"                           // - EPS on borrow (= + p mod 2^64)
            : "=r"(s0), "=r"(s1) : "r"(v0), "r"(v1), "r"(v2));
        return gl_canon(((uint64_t)s1 << 32) | s0);
    }
}
// the same for a K that is a constant only after loop unrolling (multiples of 6 below 96: every twiddle of an in-register
// transform of up to 32 points)
__device__ __forceinline__ uint64_t gl_mul_pow2_6j(uint64_t x, unsigned K) {
    switch (K) {
        case 0: return x;
        case 6: return gl_mul_pow2c<6>(x);   case 12: return gl_mul_pow2c<12>(x); case 18: return gl_mul_pow2c<18>(x);
        case 24: return gl_mul_pow2c<24>(x); case 30: return gl_mul_pow2c<30>(x); case 36: return gl_mul_pow2c<36>(x);
        case 42: return gl_mul_pow2c<42>(x); case 48: return gl_mul_pow2c<48>(x); case 54: return gl_mul_pow2c<54>(x);
        case 60: return gl_mul_pow2c<60>(x); case 66: return gl_mul_pow2c<66>(x); case 72: return gl_mul_pow2c<72>(x);
        case 78: return gl_mul_pow2c<78>(x); case 84: return gl_mul_pow2c<84>(x); case 90: return gl_mul_pow2c<90>(x);
        default: return gl_mulc(x, K < 64 ? (1ull << K) : ((1ull << (K - 32)) - (1ull << (K - 64))));
    }
}
// JIT-END
// host versions (table construction at context creation; never on the data path)
inline uint64_t gl_mul_host(uint64_t a, uint64_t b) {
    unsigned __int128 x = (unsigned __int128)a * b;
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;
    uint64_t t1 = hl * GL_EPS;
    uint64_t r = t0 + t1;
    if (r < t0) r += GL_EPS;
    return gl_canon(r);
}
inline uint64_t gl_exp_host(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = gl_mul_host(r, b);
        b = gl_mul_host(b, b);
        e >>= 1;
    }
    return r;
}
inline uint64_t gl_inv_host(uint64_t a) { return gl_exp_host(a, GL_P - 2); }
inline uint64_t gl_root_of_unity_host(unsigned log_n) {
    uint64_t r = GL_POWER_OF_TWO_GENERATOR;
    for (unsigned i = log_n; i < 32; i++) r = gl_mul_host(r, r);
    return r;
}

__device__ __forceinline__ uint64_t gl_exp(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_mul(b, b);
        e >>= 1;
    }
    return r;
}
__device__ __forceinline__ uint64_t gl_inv(uint64_t a) { return gl_exp(a, GL_P - 2); }

// ---- quadratic extension, canonical coordinates ----
struct ext2 {
    uint64_t a, b;
};
__device__ __forceinline__ ext2 ext_make(uint64_t a, uint64_t b) { ext2 r; r.a = a; r.b = b; return r; }
__device__ __forceinline__ ext2 ext_add(ext2 x, ext2 y) { return ext_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
__device__ __forceinline__ ext2 ext_sub(ext2 x, ext2 y) { return ext_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
__device__ __forceinline__ ext2 ext_mul(ext2 x, ext2 y) {
    uint64_t bb = gl_mul(x.b, y.b);
    uint64_t seven_bb = gl_mul(bb, GL_EXT_W);
    return ext_make(gl_add(gl_mul(x.a, y.a), seven_bb), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
__device__ __forceinline__ ext2 ext_scale(ext2 x, uint64_t s) { return ext_make(gl_mul(x.a, s), gl_mul(x.b, s)); }
__device__ __forceinline__ ext2 ext_inv(ext2 x) {
    uint64_t norm = gl_sub(gl_mul(x.a, x.a), gl_mul(GL_EXT_W, gl_mul(x.b, x.b)));
    uint64_t ni = gl_inv(norm);
    return ext_make(gl_mul(x.a, ni), gl_mul(gl_neg(x.b), ni));
}
__device__ __forceinline__ ext2 ext_exp(ext2 b, uint64_t e) {
    ext2 r = ext_make(1, 0);
    while (e) {
        if (e & 1) r = ext_mul(r, b);
        b = ext_mul(b, b);
        e >>= 1;
    }
    return r;
}

}  // namespace qpzk
