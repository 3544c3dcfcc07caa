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
// canonical in -> canonical out
__host__ __device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
}
__host__ __device__ __forceinline__ uint64_t gl_sub(uint64_t a, uint64_t b) {
    return a >= b ? a - b : a + (GL_P - b);
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
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_canon(gl_mul_nc(a, b)); }

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
