/* ORACLE (test infrastructure, NOT product code): Goldilocks field + quadratic extension.
 *
 * CPU restatement of qp-plonky2-field =1.5.5 (reference Cargo.toml:20, Cargo.lock:888-903; the
 * crate is NOT vendored under /root/reference -- see SURVEY.md 0.2).  The algorithm restated is
 * upstream plonky2 `field/src/goldilocks_field.rs` (reduce128, add/sub with one conditional
 * fix-up) and `field/src/extension/quadratic.rs` (F[X]/(X^2-7)), as specified in SURVEY.md A.1.
 * Reference-side anchors: p is spelled out at wormhole/inputs/src/lib.rs:23 and
 * wormhole/tests/tests/encoding_safety.rs:24; D=2, F=GoldilocksField at common/src/circuit.rs:16-18.
 *
 * All values crossing a function boundary here are canonical (< p).
 */
#ifndef ORACLE_GOLDILOCKS_H
#define ORACLE_GOLDILOCKS_H
#include <stdint.h>
#include <stddef.h>

typedef unsigned __int128 u128;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPSILON 0xFFFFFFFFULL           /* 2^64 mod p */
#define GL_GENERATOR 14293326489335486720ULL /* MULTIPLICATIVE_GROUP_GENERATOR == coset_shift() */
#define GL_POWER_OF_TWO_GENERATOR 7277203076849721926ULL /* order 2^32 */
#define GL_TWO_ADICITY 32
#define GL_EXT_W 7ULL                       /* X^2 = 7 */

static inline uint64_t gl_canon(uint64_t x) { return x - (GL_P & (0 - (uint64_t)(x >= GL_P))); }

static inline uint64_t gl_add(uint64_t a, uint64_t b) {
    /* branch-free: a,b < p so one subtraction suffices; (s < a) is the carry out of the 64-bit add */
    uint64_t s = a + b;
    uint64_t over = (uint64_t)(s < a) | (uint64_t)(s >= GL_P);
    return s - (GL_P & (0 - over));
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    return d + (GL_P & (0 - (uint64_t)(a < b)));
}
static inline uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

/* x = lo + 2^64*(hh*2^32 + hl)  ==  lo - hh + hl*(2^32-1)  (mod p), since 2^64 = 2^32-1, 2^96 = -1. */
static inline uint64_t gl_reduce128(u128 x) {
    /* branch-free form of upstream's reduce128 (the two fix-ups are data-dependent: a branchy version
     * mispredicts about once per multiplication on random data) */
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPSILON;
    uint64_t t0 = lo - hh;
    t0 -= GL_EPSILON & (0 - (uint64_t)(lo < hh));
    uint64_t t1 = hl * GL_EPSILON;
    uint64_t r = t0 + t1;
    r += GL_EPSILON & (0 - (uint64_t)(r < t0));
    return r - (GL_P & (0 - (uint64_t)(r >= GL_P)));
}
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
static inline uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }
static inline uint64_t gl_exp(uint64_t base, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, base);
        base = gl_sqr(base);
        e >>= 1;
    }
    return r;
}
static inline uint64_t gl_inv(uint64_t a) { return gl_exp(a, GL_P - 2); }
/* primitive_root_of_unity(k) = w^(2^(32-k)) */
static inline uint64_t gl_root_of_unity(unsigned log_n) {
    uint64_t r = GL_POWER_OF_TWO_GENERATOR;
    for (unsigned i = log_n; i < GL_TWO_ADICITY; i++) r = gl_sqr(r);
    return r;
}

typedef struct { uint64_t a, b; } ext2; /* a + b*X */

static inline ext2 ext_from(uint64_t a) { ext2 r = {a, 0}; return r; }
static inline ext2 ext_add(ext2 x, ext2 y) { ext2 r = {gl_add(x.a, y.a), gl_add(x.b, y.b)}; return r; }
static inline ext2 ext_sub(ext2 x, ext2 y) { ext2 r = {gl_sub(x.a, y.a), gl_sub(x.b, y.b)}; return r; }
static inline ext2 ext_mul(ext2 x, ext2 y) {
    ext2 r;
    r.a = gl_add(gl_mul(x.a, y.a), gl_mul(GL_EXT_W, gl_mul(x.b, y.b)));
    r.b = gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a));
    return r;
}
static inline ext2 ext_scale(ext2 x, uint64_t s) { ext2 r = {gl_mul(x.a, s), gl_mul(x.b, s)}; return r; }
static inline ext2 ext_inv(ext2 x) {
    /* (a + bX)^-1 = (a - bX) / (a^2 - 7 b^2) */
    uint64_t norm = gl_sub(gl_sqr(x.a), gl_mul(GL_EXT_W, gl_sqr(x.b)));
    uint64_t ni = gl_inv(norm);
    ext2 r = {gl_mul(x.a, ni), gl_mul(gl_neg(x.b), ni)};
    return r;
}
static inline int ext_eq(ext2 x, ext2 y) { return x.a == y.a && x.b == y.b; }
static inline ext2 ext_exp(ext2 base, uint64_t e) {
    ext2 r = {1, 0};
    while (e) {
        if (e & 1) r = ext_mul(r, base);
        base = ext_mul(base, base);
        e >>= 1;
    }
    return r;
}
#endif
