// Poseidon-12 permutation over Goldilocks with the whole state in registers (one thread = one
// permutation).  Replaces `PoseidonPermutation` / `Poseidon::poseidon` of qp-plonky2 =1.5.5
// (hasher `PoseidonGoldilocksConfig`, reference common/src/circuit.rs:16-18; SURVEY.md 8a rows
// a8/a9, Appendix A.3/A.4): width 12, rate 8, x^7, 4 full + 22 partial + 4 full rounds,
// MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...).
//
// B200 mapping.  The linear layer uses the literal small-constant MDS in every round: each output
// lane is two chains of twelve IMAD.WIDE.U32 (32-bit half x immediate + 64-bit accumulator), which
// is cheaper on the SM's integer pipes than the CPU-oriented sparse partial-round form (full-width
// field multiplies).  The next round's constants are pre-split into 32-bit halves and seed those
// accumulators, so adding round constants costs no instructions.  Rounds are a real loop (not
// unrolled) to stay inside the instruction cache.
#pragma once
#include "goldilocks.cuh"

namespace qpzk {

// RC split into zero-extended 32-bit halves: [0..360) low halves, [360..720) high halves,
// followed by 12+12 zeros used as the "next round" constants of the last round.
__constant__ uint64_t c_rc_split[2 * 372];
__constant__ uint64_t c_rc_first[12];  // round-0 constants, added to the input state directly

// MDS output lane from 32-bit half sums:  value = lo + 2^32 * hi  with lo, hi < 2^43.
__device__ __forceinline__ uint64_t mds_combine(uint64_t lo, uint64_t hi) {
    uint32_t h1 = (uint32_t)(hi >> 32), h0 = (uint32_t)hi;
    uint64_t t = lo + (uint64_t)h1 * 0xFFFFFFFFu;  // h1 * 2^64 = h1 * EPS (mod p); no overflow
    uint64_t u = (uint64_t)h0 << 32;
    uint64_t r = t + u;
    if (r < u) r += GL_EPS;                         // wrapped value is < 2^43, cannot wrap again
    return r;
}

__device__ __forceinline__ uint64_t sbox7_nc(uint64_t x) {
    uint64_t x2 = gl_sqr_nc(x);
    uint64_t x4 = gl_sqr_nc(x2);
    uint64_t x3 = gl_mul_nc(x, x2);
    return gl_mul_nc(x3, x4);
}

// s <- MDS * s + RC[next_round]   (constants folded into the accumulators)
__device__ __forceinline__ void mds_layer_rc(uint64_t (&s)[12], int next_round) {
    constexpr uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t l[12], h[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { l[i] = (uint32_t)s[i]; h[i] = (uint32_t)(s[i] >> 32); }
    const uint64_t* rc_lo = c_rc_split + 12 * next_round;
    const uint64_t* rc_hi = c_rc_split + 372 + 12 * next_round;
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint64_t lo = rc_lo[r], hi = rc_hi[r];
#pragma unroll
        for (int i = 0; i < 12; i++) {
            uint32_t c = CIRC[i] + ((r == 0 && i == 0) ? 8u : 0u);  // + DIAG
            lo += (uint64_t)l[(i + r) % 12] * c;
            hi += (uint64_t)h[(i + r) % 12] * c;
        }
        s[r] = mds_combine(lo, hi);
    }
}

// In/out: arbitrary u64 representatives (inputs from memory are canonical; outputs may be >= p).
__device__ __forceinline__ void poseidon_permute(uint64_t (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_nc(s[i], c_rc_first[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        s[0] = sbox7_nc(s[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) s[i] = sbox7_nc(s[i]);
        }
        mds_layer_rc(s, r + 1);  // round 30 "constants" are zeros
    }
}

}  // namespace qpzk
