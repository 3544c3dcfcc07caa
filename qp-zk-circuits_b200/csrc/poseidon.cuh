// Poseidon-12 permutation over Goldilocks with the whole state in registers (one thread = one
// permutation).  Replaces `PoseidonPermutation` / `Poseidon::poseidon` of qp-plonky2 =1.5.5
// (hasher `PoseidonGoldilocksConfig`, reference common/src/circuit.rs:16-18; SURVEY.md 8a rows
// a8/a9, Appendix A.3/A.4): width 12, rate 8, x^7, 4 full + 22 partial + 4 full rounds,
// MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...).
//
// B200 mapping (measured: IMAD.WIDE issues at half the 32-bit IMAD rate, and the ALU and FMA pipes of
// an SM sub-partition run side by side):
//  * field multiply = 4 IMAD.WIDE + an add-with-carry chain written in PTX (~14 IADD3): products are
//    assembled and reduced (2^64 = 2^32-1, 2^96 = -1) without any compare/select.
//  * the linear layer of FULL rounds is the literal small-constant MDS: per output lane two chains of
//    twelve IMAD.WIDE.U32 (32-bit half x immediate + 64-bit accumulator); the next round's constants
//    are pre-split into 32-bit halves and seed the accumulators, so adding them costs nothing.
//  * PARTIAL rounds are split: the first 22-m stay literal (FMA-pipe heavy), the last m use the sparse
//    factorisation (ALU-pipe heavy: 22 full-width multiplies, one 160-bit dot-product accumulator).  Mixing
//    the two forms balances the two pipes; m = QPZK_SPARSE_ROUNDS is set by the constants generator.
//  * rounds are real loops (not unrolled) so the body stays inside the instruction cache.
#pragma once
#include "goldilocks.cuh"

namespace qpzk {

constexpr int SPARSE_ROUNDS = QPZK_SPARSE_ROUNDS;
constexpr int LITERAL_PARTIAL = 22 - SPARSE_ROUNDS;

// Row `next` of this table = constants folded into the MDS layer that precedes round `next`:
// [0..372) zero-extended low halves, [372..744) high halves (31 rows of 12; row 30 is all zero).
__constant__ uint64_t c_rc_split[2 * 372];
__constant__ uint64_t c_rc_first[12];                  // round-0 constants, added to the input directly
// The two power-of-two MDS entries (16 and 2) are read from constant memory: as immediates ptxas rewrites
// `mad.wide x, 16, acc` into a shift + 64-bit add sequence (4 ALU instructions instead of one IMAD.WIDE).
__constant__ uint32_t c_mds_pow2[2];
__constant__ uint64_t c_sp_k[SPARSE_ROUNDS + 1];       // lane-0 scalar of sparse round i (entry 0 is folded into the table)
__constant__ uint64_t c_sp_init[121];
__constant__ uint64_t c_sp_w[11 * SPARSE_ROUNDS + 1];
__constant__ uint64_t c_sp_vhat[11 * SPARSE_ROUNDS + 1];
__constant__ uint64_t c_sp_v26[12];                    // POST + RC[26], added after the sparse segment

// ---- round pieces ---------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t sbox7_nc(uint64_t x) {
    uint64_t x2 = sqrp(x);
    uint64_t x4 = sqrp(x2);
    uint64_t x3 = mulp(x, x2);
    return mulp(x3, x4);
}

// acc += x * C  as one IMAD.WIDE.U32 with an immediate multiplier (written in PTX so that the compiler does
// not strength-reduce the small constants into shift/add sequences on the busier ALU pipe).
template <uint32_t C>
__device__ __forceinline__ void macw(uint64_t& acc, uint32_t x) {
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "n"(C));
}

// MDS output lane from 32-bit half sums:  value = lo + 2^32 * hi  with lo, hi < 2^43.
__device__ __forceinline__ uint64_t mds_combine(uint64_t lo, uint64_t hi) {
    uint32_t h1 = (uint32_t)(hi >> 32), h0 = (uint32_t)hi;
    uint64_t t = lo;
    macw<0xFFFFFFFFu>(t, h1);                       // h1 * 2^64 = h1 * EPS (mod p); no overflow
    uint32_t t0 = (uint32_t)t, t1 = (uint32_t)(t >> 32), s0, s1;
    asm("{\n\t.reg .u32 c,m;\n\t"
        "add.cc.u32 %1,%3,%4;\n\taddc.u32 c,0,0;\n\t"   // t + (h0 << 32): only the high word changes
        "sub.u32 m,0,c;\n\t"
        "add.cc.u32 %0,%2,m;\n\taddc.u32 %1,%1,0;\n\t}"  // + EPS on wrap; wrapped value < 2^43, cannot wrap again
        : "=&r"(s0), "=&r"(s1) : "r"(t0), "r"(t1), "r"(h0));
    return ((uint64_t)s1 << 32) | s0;
}

template <int R, int I>
struct MdsRow {
    static __device__ __forceinline__ void run(uint64_t& lo, uint64_t& hi, const uint32_t (&l)[12], const uint32_t (&h)[12]) {
        constexpr uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
        constexpr uint32_t c = CIRC[I] + ((R == 0 && I == 0) ? 8u : 0u);  // + DIAG
        if (c == 16 || c == 2) {
            const uint32_t cr = c_mds_pow2[c == 16 ? 0 : 1];
            asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(lo) : "r"(l[(I + R) % 12]), "r"(cr));
            asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(hi) : "r"(h[(I + R) % 12]), "r"(cr));
        } else {
            macw<c>(lo, l[(I + R) % 12]);
            macw<c>(hi, h[(I + R) % 12]);
        }
        MdsRow<R, I + 1>::run(lo, hi, l, h);
    }
};
template <int R>
struct MdsRow<R, 12> {
    static __device__ __forceinline__ void run(uint64_t&, uint64_t&, const uint32_t (&)[12], const uint32_t (&)[12]) {}
};

template <int R>
struct MdsLayer {
    static __device__ __forceinline__ void run(uint64_t (&s)[12], const uint32_t (&l)[12], const uint32_t (&h)[12],
                                               const uint64_t* rc_lo, const uint64_t* rc_hi) {
        uint64_t lo = rc_lo[R], hi = rc_hi[R];
        MdsRow<R, 0>::run(lo, hi, l, h);
        s[R] = mds_combine(lo, hi);
        MdsLayer<R + 1>::run(s, l, h, rc_lo, rc_hi);
    }
};
template <>
struct MdsLayer<12> {
    static __device__ __forceinline__ void run(uint64_t (&)[12], const uint32_t (&)[12], const uint32_t (&)[12],
                                               const uint64_t*, const uint64_t*) {}
};

#ifndef QPZK_MDS_FFT
#define QPZK_MDS_FFT 1
#endif
#ifndef QPZK_SBOX_PASSES
#define QPZK_SBOX_PASSES 1   // S-boxes of a full round: 1 pass over 12 lanes (1.6% faster than 3 passes over 4 with a lane rotation), 2 over 6, or 3 over 4 (smallest code)
#endif

#if QPZK_MDS_FFT
// ---- add/shift-only MDS on 22/22/20-bit limbs ----------------------------------------------------------------
// The circulant is a product in Z[x]/(x^12 - 1); a 4-point DFT over the three stride-3 subsequences splits it into
// Z[x]/(x^3-1) x Z[i][x]/(x^3-i) x Z[x]/(x^3+1), where the transformed matrix is [16,32,16], [(2,1),(-4,-1),(16,-1)],
// [-1,-8,2] (derivation and exhaustive self-check: tools/derive_mds_fft.py).  All arithmetic is wrapping 32-bit;
// the true outputs are < 2^31, so nothing is lost.  ~70 IADD3/LEA/IMAD per limb instead of 144 IMAD.WIDE per
// 32-bit half: the layer moves off the FMA-heavy pipe, the bottleneck of the literal form (profiles/r1_summary.md).
// Row `next` of c_rc_limb = the next round's constants split into the same limbs, added in the output stage.
__constant__ uint32_t c_rc_limb[31 * 36];

__device__ __forceinline__ void mds_limb(const uint32_t (&e)[12], const uint32_t* __restrict__ rc, uint32_t (&y)[12]) {
    uint32_t A[3], B[3], Cr[3], Ci[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        uint32_t t = e[j] + e[j + 6], u = e[j + 3] + e[j + 9];
        A[j] = t + u; B[j] = t - u;
        Cr[j] = e[j] - e[j + 6]; Ci[j] = e[j + 3] - e[j + 9];
    }
    const uint32_t T = A[0] + A[1] + A[2];
    const uint32_t TA[3] = {T + A[2], T + A[0], T + A[1]};
    const uint32_t P2[3] = {8u * B[2] - B[0] - 2u * B[1], 0u - 8u * B[0] - B[1] - 2u * B[2], 2u * B[0] - 8u * B[1] - B[2]};
    const uint32_t r0 = Cr[0], r1 = Cr[1], r2 = Cr[2], i0 = Ci[0], i1 = Ci[1], i2 = Ci[2];
    const uint32_t P1r[3] = {2u * r0 - i0 - 16u * i1 + r1 + 4u * i2 + r2,
                             i0 - 4u * r0 + 2u * r1 - i1 + r2 - 16u * i2,
                             16u * r0 + i0 - 4u * r1 + i1 + 2u * r2 - i2};
    const uint32_t P1i[3] = {2u * i0 + r0 + 16u * r1 + i1 - 4u * r2 + i2,
                             r1 - 4u * i0 - r0 + 2u * i1 + 16u * r2 + i2,
                             16u * i0 - r0 - 4u * i1 - r1 + 2u * i2 + r2};
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const uint32_t U = 16u * TA[j] + P2[j], V = 16u * TA[j] - P2[j];
        y[j] = U + P1r[j] + rc[j];
        y[j + 3] = V + P1i[j] + rc[j + 3];
        y[j + 6] = U - P1r[j] + rc[j + 6];
        y[j + 9] = V - P1i[j] + rc[j + 9];
    }
    y[0] += 8u * e[0];
}

// y0 + 2^22 y1 + 2^44 y2 - 4096 (each y < 2^31) as an arbitrary u64 representative mod p; the table constants
// carry the -4096 (c_rc_limb holds the limbs of c - 4096 mod p).  The top limb is pre-reduced in the limb domain:
// y2 = h * 2^20 + l and h * 2^64 = h * 2^32 - h (mod p), so limb 1 takes h << 10 and limb 0 takes 4096 - h >= 0.
// What remains can exceed 2^64 only through the carry of the last addition: one +EPS, which cannot wrap again.
__device__ __forceinline__ uint64_t limb_combine(uint32_t y0, uint32_t y1, uint32_t y2) {
    const uint32_t h = y2 >> 20;
    const uint32_t t0 = y0 + 4096u - h;
    const uint32_t t1 = y1 + (h << 10);
    const uint64_t a = (uint64_t)t0 + ((uint64_t)t1 << 22);           // < 2^54
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), l12 = y2 << 12, s0, s1;
    // Only add-flavoured carry ops: ptxas keeps the raw adder carry in CC.CF (after sub.cc it is NOT-borrow), so
    // chains that mix add.cc with subc do not compute what the PTX manual's prose suggests.
    asm("{\n\t.reg .u32 k,m;\n\t"
        "add.cc.u32 %1,%3,%4;\n\taddc.u32 k,0,0;\n\t"               // high word + l * 2^12;  k = carry
        "sub.u32 m,0,k;\n\t"                                         // carry * (2^32 - 1) = (0 : m)
        "add.cc.u32 %0,%2,m;\n\taddc.u32 %1,%1,0;\n\t}"
        : "=&r"(s0), "=&r"(s1) : "r"(a0), "r"(a1), "r"(l12));
    return ((uint64_t)s1 << 32) | s0;
}

// s <- MDS * s + table row `next`
__device__ __forceinline__ void mds_layer_rc(uint64_t (&s)[12], int next) {
    const uint32_t* rc = c_rc_limb + 36 * next;
    uint32_t e[12], y0[12], y1[12], y2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) e[i] = (uint32_t)s[i] & 0x3FFFFFu;
    mds_limb(e, rc, y0);
#pragma unroll
    for (int i = 0; i < 12; i++) e[i] = (uint32_t)(s[i] >> 22) & 0x3FFFFFu;
    mds_limb(e, rc + 12, y1);
#pragma unroll
    for (int i = 0; i < 12; i++) e[i] = (uint32_t)(s[i] >> 44);
    mds_limb(e, rc + 24, y2);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = limb_combine(y0[i], y1[i], y2[i]);
}

#ifndef QPZK_PARTIAL_LIMB
#define QPZK_PARTIAL_LIMB 1
#endif
#if QPZK_PARTIAL_LIMB
// ---- the 22 partial rounds with lanes 1..11 kept in the limb domain -------------------------------------------------
// Only lane 0 passes through an S-box in a partial round; the other eleven lanes see nothing but the linear layer, so
// they need not be recombined into 64-bit words and split again between rounds.  What they do need is a carry
// normalisation, because a layer grows a limb by up to 2^8.15 (sum of a matrix row = 284):
//     t1 = y1 + (y0 >> 22),  t2 = y2 + (t1 >> 22),  h = t2 >> 20   (h * 2^64 = h * 2^32 - h mod p)
//     n0 = (y0 & M22) + 4096 - h,   n1 = (t1 & M22) + (h << 10),   n2 = t2 & M20
// 10 instructions per lane instead of 12 (combine) + 5 (split).  The table constants carry -4096 (see limb_combine), so
// adding 4096 restores the true value and keeps n0 positive (h <= 2^11).  Bounds: n0 < 2^22 + 2^12, n1 < 2^22 + 2^21 +
// 2^10, n2 < 2^20, hence every output of the next layer is < 284 * (2^22 + 2^21 + 2^10) + 2^22 < 2^31: exact in the
// wrapping 32-bit arithmetic of mds_limb.
__device__ __forceinline__ void limb_normalize(uint32_t y0, uint32_t y1, uint32_t y2, uint32_t& n0, uint32_t& n1, uint32_t& n2) {
    const uint32_t t1 = y1 + (y0 >> 22);
    const uint32_t t2 = y2 + (t1 >> 22);
    const uint32_t h = t2 >> 20;
    n0 = (y0 & 0x3FFFFFu) + 4096u - h;
    n1 = (t1 & 0x3FFFFFu) + (h << 10);
    n2 = t2 & 0xFFFFFu;
}

// rounds 4..25: in s = state after the linear layer of round 3 (constants of round 4 added); out s = state after the
// linear layer of round 25 (constants of round 26 added)
__device__ __forceinline__ void partial_rounds_limb(uint64_t (&s)[12]) {
    uint32_t l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 1; i < 12; i++) {
        l0[i] = (uint32_t)s[i] & 0x3FFFFFu;
        l1[i] = (uint32_t)(s[i] >> 22) & 0x3FFFFFu;
        l2[i] = (uint32_t)(s[i] >> 44);
    }
    uint64_t s0 = s[0];
#pragma unroll 1
    for (int r = 4; r < 26; r++) {
        s0 = sbox7_nc(s0);
        l0[0] = (uint32_t)s0 & 0x3FFFFFu;
        l1[0] = (uint32_t)(s0 >> 22) & 0x3FFFFFu;
        l2[0] = (uint32_t)(s0 >> 44);
        const uint32_t* rc = c_rc_limb + 36 * (r + 1);
        uint32_t y0[12], y1[12], y2[12];
        mds_limb(l0, rc, y0);
        mds_limb(l1, rc + 12, y1);
        mds_limb(l2, rc + 24, y2);
        s0 = limb_combine(y0[0], y1[0], y2[0]);
#pragma unroll
        for (int i = 1; i < 12; i++) limb_normalize(y0[i], y1[i], y2[i], l0[i], l1[i], l2[i]);
    }
    s[0] = s0;
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = limb_combine(l0[i] - 4096u, l1[i], l2[i]);   // limb_combine adds the 4096 back
}
#endif  // QPZK_PARTIAL_LIMB
#else
// s <- MDS * s + table row `next`
__device__ __forceinline__ void mds_layer_rc(uint64_t (&s)[12], int next) {
    uint32_t l[12], h[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { l[i] = (uint32_t)s[i]; h[i] = (uint32_t)(s[i] >> 32); }
    MdsLayer<0>::run(s, l, h, c_rc_split + 12 * next, c_rc_split + 372 + 12 * next);
}

#endif  // QPZK_MDS_FFT

// The sparse segment: lanes 1..11 <- INIT * lanes 1..11, then SPARSE_ROUNDS rounds of
//   x0 <- (x0 + K[i])^7;  s0' = 25*x0 + <VHAT[i], s[1..]>;  s[1+j] += W[i][j]*x0;   and finally s += V26.
__device__ __forceinline__ void sparse_segment(uint64_t (&s)[12]) {
    {
        uint64_t t[11];
#pragma unroll 1
        for (int i = 0; i < 11; i++) {
            acc160 acc = {0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 11; j++) mac160(acc, c_sp_init[11 * i + j], s[1 + j]);
            uint64_t v = reduce160(acc);
#pragma unroll
            for (int k = 0; k < 11; k++) if (k == i) t[k] = v;
        }
#pragma unroll
        for (int j = 0; j < 11; j++) s[1 + j] = t[j];
    }
#pragma unroll 1
    for (int i = 0; i < SPARSE_ROUNDS; i++) {
        uint64_t x0 = s[0];
        if (i > 0) x0 = add_canon_p(x0, c_sp_k[i]);  // K[0] came in through the table row
        x0 = sbox7_nc(x0);
        acc160 acc = {0, 0, 0, 0, 0};
        mac160(acc, 25, x0);
        const uint64_t* vh = c_sp_vhat + 11 * i;
        const uint64_t* w = c_sp_w + 11 * i;
#pragma unroll
        for (int j = 0; j < 11; j++) mac160(acc, vh[j], s[1 + j]);
#pragma unroll
        for (int j = 0; j < 11; j++) s[1 + j] = mul_addp(w[j], x0, s[1 + j]);
        s[0] = reduce160(acc);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = add_canon_p(s[i], c_sp_v26[i]);
}

// In/out: arbitrary u64 representatives (inputs from memory are canonical; outputs may be >= p).
// ONE round loop shares a single copy of the S-box and MDS code between all full and literal partial rounds;
// the 12 S-boxes of a full round run as 3 passes over 4 lanes with a register rotation in between.  This keeps
// the whole permutation well inside the instruction cache (instruction-fetch stalls were the top stall reason
// of the straight-line version, profiles/r1_poseidon_notes.md).
__device__ __forceinline__ void poseidon_permute(uint64_t (&s)[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = add_canon_p(s[i], c_rc_first[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        if (SPARSE_ROUNDS > 0 && r == 4 + LITERAL_PARTIAL) {
            sparse_segment(s);
            r = 26;
        }
#if QPZK_MDS_FFT && QPZK_PARTIAL_LIMB
        if (SPARSE_ROUNDS == 0 && r == 4) {
            partial_rounds_limb(s);
            r = 26;
        }
#endif
        if (r < 4 || r >= 26) {
#if QPZK_SBOX_PASSES == 1
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = sbox7_nc(s[i]);
#elif QPZK_SBOX_PASSES == 2
#pragma unroll 1
            for (int k = 0; k < 2; k++) {
                uint64_t t[6];
#pragma unroll
                for (int i = 0; i < 6; i++) t[i] = sbox7_nc(s[i]);
#pragma unroll
                for (int i = 0; i < 6; i++) s[i] = s[i + 6];
#pragma unroll
                for (int i = 0; i < 6; i++) s[6 + i] = t[i];
            }
#else
#pragma unroll 1
            for (int k = 0; k < 3; k++) {
                uint64_t a = sbox7_nc(s[0]), b = sbox7_nc(s[1]), c = sbox7_nc(s[2]), d = sbox7_nc(s[3]);
#pragma unroll
                for (int i = 0; i < 8; i++) s[i] = s[i + 4];  // rotate by 4 lanes; three passes restore the order
                s[8] = a; s[9] = b; s[10] = c; s[11] = d;
            }
#endif
        } else {
            s[0] = sbox7_nc(s[0]);
        }
        mds_layer_rc(s, r + 1);  // row 30 is zero
    }
}

// ---- one permutation spread over a warp (latency path) ----------------------------------------------------------
// Lane i < 12 holds state word i.  The S-boxes of a full round run in parallel, the linear layer fetches the other
// lanes' words with shuffles and accumulates the literal small-constant matrix (one IMAD.WIDE per 32-bit half and
// term).  ~470 cycles per round instead of a 23 k-instruction serial chain: 3-5x lower latency for the places
// where permutations are sequential by construction (the Fiat-Shamir challenger, the top of a Merkle tree).  All 32
// lanes must call it; lanes >= 12 carry don't-care values.  Constants: d_rc_flat[12*r + i] (31 rows, row 30 zero).
__device__ uint64_t d_rc_flat[372];

// W = 32: one permutation per warp.  W = 16: TWO independent permutations per warp, one per half-warp (shuffles of
// width 16; `lane` = index inside the half): twice the throughput per warp instruction for levels / leaf sets that are
// wide enough to be throughput bound, the same latency per permutation.
template <int I, int W>
struct WarpMds {
    static __device__ __forceinline__ void run(uint64_t& lo, uint64_t& hi, uint32_t l, uint32_t h, int lane) {
        constexpr uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
        int src = lane + I;
        if (src >= 12) src -= 12;
        const uint32_t vl = __shfl_sync(0xffffffffu, l, src, W), vh = __shfl_sync(0xffffffffu, h, src, W);
        lo += (uint64_t)vl * CIRC[I];
        hi += (uint64_t)vh * CIRC[I];
        WarpMds<I + 1, W>::run(lo, hi, l, h, lane);
    }
};
template <int W>
struct WarpMds<12, W> {
    static __device__ __forceinline__ void run(uint64_t&, uint64_t&, uint32_t, uint32_t, int) {}
};

template <int W>
__device__ __forceinline__ uint64_t poseidon_permute_sub(uint64_t s, int lane) {
    const int li = lane < 12 ? lane : 0;
    // coefficient of state word 0 in output word `lane`: circ[(12 - lane) mod 12], + the diagonal 8 for lane 0
    constexpr uint32_t CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t c0 = 25;
#pragma unroll
    for (int k = 1; k < 12; k++) if (li == k) c0 = CIRC[12 - k];
    s = add_canon_p(s, d_rc_flat[li]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        const uint64_t rc = d_rc_flat[12 * (r + 1) + li];
        uint64_t lo = (uint32_t)rc, hi = rc >> 32;
        if (r < 4 || r >= 26) {
            s = sbox7_nc(s);
            const uint32_t l = (uint32_t)s, h = (uint32_t)(s >> 32);
            WarpMds<0, W>::run(lo, hi, l, h, lane);
            if (lane == 0) { lo += 8ull * l; hi += 8ull * h; }
        } else {
            // Partial round: only word 0 passes through the S-box, so the eleven other terms of every output do not
            // depend on it.  They are accumulated from the OLD words (lane 0 contributes zero) in the same basic block as
            // the S-box: the scheduler fills the stalls of the S-box's dependent multiplication chain with the shuffles
            // and multiply-adds of the linear layer, and only word 0's term is added afterwards.
            const uint32_t l = lane == 0 ? 0u : (uint32_t)s, h = lane == 0 ? 0u : (uint32_t)(s >> 32);
            WarpMds<0, W>::run(lo, hi, l, h, lane);
            const uint64_t x0 = sbox7_nc(s);                       // meaningful on lane 0 only
            const uint32_t x0l = __shfl_sync(0xffffffffu, (uint32_t)x0, 0, W);
            const uint32_t x0h = __shfl_sync(0xffffffffu, (uint32_t)(x0 >> 32), 0, W);
            lo += (uint64_t)x0l * c0;
            hi += (uint64_t)x0h * c0;
        }
        s = mds_combine(lo, hi);
    }
    return s;
}
__device__ __forceinline__ uint64_t poseidon_permute_warp(uint64_t s, int lane) { return poseidon_permute_sub<32>(s, lane); }

}  // namespace qpzk
