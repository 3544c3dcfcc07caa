// Strided passes of the coset low-degree extension, second generation: a 2^s-point transform (s = 5..8) over
// position bits [b_lo, b_lo+s) done as TWO register-resident sub-transforms with a single shared-memory exchange,
// and every position-dependent factor hoisted out of a loop over columns.
//
// Replaces the same upstream code as ntt.cu (`PolynomialCoeffs::{lde, coset_fft}` of qp-plonky2-field =1.5.5,
// SURVEY.md 8a rows a2/a5) and computes, bit for bit, what ntt_pass_kernel computes in its FWD modes:
//     out[hi, bitrev_s(f), low] = w_{2^b_hi}^(f*low) * sum_e x[hi, e, low] * w_S^(e*f),     S = 2^s,
//     x = in * s_j^p  in the first pass (coset scaling),  p = e * 2^b_lo + low.
//
// Why it is faster (profiles/r1_summary.md: the old pass is bound by ALU instruction issue, ~430 instructions per
// output element of a 7-bit pass, of which the butterflies are less than a third):
//   * S = A*B.  With e = eg + B*k and f = f1 + A*f2:  w_S^(e f) = w_S^(eg f1) * w_B^(eg f2) * w_A^(k f1), so a thread
//     first runs an A-point transform on A registers (inputs B apart: no exchange needed before it), multiplies by
//     T[eg][f1], exchanges through shared memory once, and runs a B-point transform on B registers.  The twiddles
//     inside the register transforms are roots of unity of order <= 16 = compile-time powers of two (2^12 has order
//     16 in Goldilocks: w_2 = 2^96, w_4 = 2^48, w_8 = 2^24, w_16 = 2^12 are exactly plonky2's roots).
//   * the coset scaling factors split as s^p = s^(B 2^b_lo k) * s^(eg 2^b_lo) * s^low: the first is the same for
//     every thread (A values per coset), the second merges into T, the third merges into the output factor.  The
//     output factors (one per output element of the tile) live in registers across the column loop, T and the
//     A per-coset values in shared memory: per element and pass there remain 2-3 general multiplications instead
//     of 5.5-7.5, and no table arithmetic at all inside the column loop.
#include <cstdlib>

#include "internal.cuh"
#include "goldilocks.cuh"
#include "ntt_local.cuh"

#ifndef QPZK_POW2_SHIFT
#define QPZK_POW2_SHIFT 1
#endif

namespace qpzk {

__host__ __device__ constexpr uint64_t pow2_mod_p(unsigned t) {   // 2^t mod p for t < 96, as a canonical constant
    return t < 64 ? (1ull << t) : ((1ull << (t - 32)) - (1ull << (t - 64)));  // 2^64 = 2^32 - 1
}

// In-register decimation-in-frequency transform of 2^LOG values: natural order in, bit-reversed order out
// (register r ends up holding frequency bitrev_LOG(r)).  Twiddle of the pair (k, k+half) is w_{2 half}^(k mod half)
// = 2^((96/half) * (k mod half)).
template <int LOG>
__device__ __forceinline__ void dft_regs(uint64_t (&x)[1 << LOG]) {
    constexpr int R = 1 << LOG;
#pragma unroll
    for (int t = 0; t < LOG; t++) {
        const int half = R >> (t + 1);
#pragma unroll
        for (int k = 0; k < R; k++) {
            if (k & half) continue;
            const int j = k & (half - 1);
            const uint64_t a = x[k], b = x[k + half];
            x[k] = gl_add(a, b);
            const uint64_t d = gl_sub(a, b);
#if QPZK_POW2_SHIFT
            x[k + half] = j == 0 ? d : gl_mul_pow2_6j(d, (96 / half) * j);     // shifts instead of a general product
#else
            x[k + half] = j == 0 ? d : gl_mulc(d, pow2_mod_p((96 / half) * j));
#endif
        }
    }
}

template <int BITS>
__host__ __device__ constexpr uint32_t brev_c(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < BITS; i++) r |= ((x >> i) & 1u) << (BITS - 1 - i);
    return r;
}

// Tile = 2^(SA+SB) positions of the pass x Wd = 4096 / 2^(SA+SB) consecutive low positions; 256 threads; every
// thread handles 16 elements of each column in both phases.
// LEAN = true: the 16 output factors of a thread live in shared memory instead of registers and the next column is not
// prefetched, so that the kernel fits 128 registers and TWO blocks share an SM (ncu, round 1: at 178-196 registers one
// block of 8 warps per SM issued on 54 % of the cycles -- nothing to switch to while a warp waits; with 16 warps the
// other block's arithmetic hides the loads the prefetch used to hide).
template <int SA, int SB, bool LEAN>
__global__ void __launch_bounds__(NTT_THREADS, LEAN ? 2 : 1) ntt_fwd_kernel(NttPassArgs a) {
    constexpr int S_LOG = SA + SB, A = 1 << SA, B = 1 << SB, S = 1 << S_LOG;
    constexpr int LWD = 12 - S_LOG, WD = 1 << LWD, G = NTT_THREADS / WD;   // lanes along `low`, thread groups
    constexpr int IT1 = B / G, IT2 = A / G;                                // items per thread in phase 1 / 2
    static_assert(IT1 >= 1 && IT2 >= 1 && IT1 * A == 16 && IT2 * B == 16, "tile shape");
    extern __shared__ uint64_t smem[];
    uint64_t* buf = smem;                 // [f1][eg][w]
    uint64_t* T = smem + S * WD;          // [eg][f1]
    uint64_t* Ck = T + S;                 // [k]
    uint64_t* Osm = Ck + A;               // LEAN: [it * B + r][thread]
    const bool coset_first = a.mode == NTT_FWD_COSET_FIRST;
    const uint32_t b_hi = a.b_lo + S_LOG;
    const uint32_t w = threadIdx.x & (WD - 1), g = threadIdx.x >> LWD;

    uint32_t tile = blockIdx.x, coset = 0;
    if (coset_first) { coset = tile & (a.num_cosets - 1); tile /= a.num_cosets; }
    const uint32_t llow = a.b_lo - LWD;
    const uint64_t hi = tile >> llow;
    const uint64_t low0 = (uint64_t)(tile & ((1u << llow) - 1)) << LWD;
    const uint64_t base = (hi << b_hi) + low0;
    const uint64_t* clo = a.cs_lo + (size_t)coset * TW_LO_SIZE;
    const uint64_t* chi = a.cs_hi + (size_t)coset * a.cs_hi_size;

    // ---- factors that do not depend on the column ----
    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) {
        const uint32_t eg = i >> SA, f1 = i & (A - 1);
        uint64_t t = tw_get(a.tw_lo, a.tw_hi, (uint64_t)(eg * f1) << (a.log_n - S_LOG));
        if (coset_first) t = gl_mulc(t, tw_get(clo, chi, (uint64_t)eg << a.b_lo));
        T[i] = t;
    }
    if (threadIdx.x < A) Ck[threadIdx.x] = coset_first ? tw_get(clo, chi, ((uint64_t)threadIdx.x << SB) << a.b_lo) : 1;
    uint64_t O[LEAN ? 1 : IT2][LEAN ? 1 : B];   // output factor of (f1 = g + G*it, f2 = bitrev(r)) at this thread's low position
    {
        const uint64_t pl = low0 + w;
        const uint64_t cw = coset_first ? tw_get(clo, chi, pl) : 1;
#pragma unroll
        for (int it = 0; it < IT2; it++) {
            const uint32_t f1 = g + G * it;
#pragma unroll
            for (int r = 0; r < B; r++) {
                const uint32_t f = f1 + A * brev_c<SB>(r);
                uint64_t o = tw_get(a.tw_lo, a.tw_hi, ((uint64_t)f * pl) << (a.log_n - b_hi));
                if (coset_first) o = gl_mulc(o, cw);
                if (LEAN) Osm[(it * B + r) * NTT_THREADS + threadIdx.x] = o; else O[LEAN ? 0 : it][LEAN ? 0 : r] = o;
            }
        }
    }
    __syncthreads();

    // ---- columns ----
    const uint32_t c_begin = blockIdx.y * a.cols_per_block;
    const uint32_t c_end = min((uint32_t)a.num_arrays, c_begin + a.cols_per_block);
    // The inputs of the next column are requested right after the first phase has consumed the current ones, so the
    // loads are in flight during the exchange and the second phase (one block per SM: nothing else would hide them).
    const uint64_t in_off = (uint64_t)blockIdx.z * a.group_stride + base + w;
    uint64_t x[IT1][A];
    if (!LEAN && c_begin < c_end) {
        const uint64_t* in = a.in + c_begin * a.in_stride + in_off;
#pragma unroll
        for (int it = 0; it < IT1; it++)
#pragma unroll
            for (int k = 0; k < A; k++) x[it][k] = in[(uint64_t)(g + G * it + B * k) << a.b_lo];
    }
    for (uint32_t col = c_begin; col < c_end; col++) {
        uint64_t* out = a.out + col * a.out_stride + ((uint64_t)coset + blockIdx.z) * a.group_stride + base + w;
        if (LEAN) {
            const uint64_t* in = a.in + col * a.in_stride + in_off;
#pragma unroll
            for (int it = 0; it < IT1; it++)
#pragma unroll
                for (int k = 0; k < A; k++) x[it][k] = in[(uint64_t)(g + G * it + B * k) << a.b_lo];
        }
#pragma unroll
        for (int it = 0; it < IT1; it++) {
            const uint32_t eg = g + G * it;
            if (coset_first) {
#pragma unroll
                for (int k = 1; k < A; k++) x[it][k] = gl_mulc(x[it][k], Ck[k]);
            }
            dft_regs<SA>(x[it]);
#pragma unroll
            for (int r = 0; r < A; r++) {
                const uint32_t f1 = brev_c<SA>(r);
                buf[(f1 * B + eg) * WD + w] = gl_mulc(x[it][r], T[eg * A + f1]);
            }
        }
        if (!LEAN && col + 1 < c_end) {
            const uint64_t* in = a.in + (col + 1) * a.in_stride + in_off;
#pragma unroll
            for (int it = 0; it < IT1; it++)
#pragma unroll
                for (int k = 0; k < A; k++) x[it][k] = in[(uint64_t)(g + G * it + B * k) << a.b_lo];
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < IT2; it++) {
            const uint32_t f1 = g + G * it;
            uint64_t y[B];
#pragma unroll
            for (int j = 0; j < B; j++) y[j] = buf[(f1 * B + j) * WD + w];
            dft_regs<SB>(y);
            const uint32_t e0 = (__brev(f1) >> (32 - SA)) * B;   // slot of frequency f = f1 + A*f2 is bitrev_s(f)
#pragma unroll
            for (int r = 0; r < B; r++)
                out[(uint64_t)(e0 + r) << a.b_lo] = gl_mulc(y[r], LEAN ? Osm[(it * B + r) * NTT_THREADS + threadIdx.x] : O[LEAN ? 0 : it][LEAN ? 0 : r]);
        }
        __syncthreads();
    }
}

// Last LDE pass, 32-point version: position bits [0, 5) of every column for CH chunks of 32 consecutive rows of one
// coset.  A thread owns one (column, chunk): 32 registers, a full 32-point transform whose twiddles are powers of
// two (w_32 = 2^6), no barrier inside the transform.  Shared memory is only the transposing buffer between the
// column-major loads and the whole-row stores into the leaf matrix (row stride RP odd: both sides conflict free).
__global__ void __launch_bounds__(LDE_LAST32_MAX_THREADS) lde_last32_kernel(LdeLastArgs a, uint32_t log_ch) {
    extern __shared__ uint64_t smem[];
    const uint32_t C = a.num_cols, CH = 1u << log_ch, R = 32u << log_ch, RP = R + 1;
    const uint64_t n = 1ull << a.log_n;
    const uint64_t blocks_per_coset = n >> (5 + log_ch);
    const uint32_t coset = blockIdx.x / blocks_per_coset;
    const uint64_t row0 = (blockIdx.x % blocks_per_coset) * R;

    const uint64_t* in = a.in + (uint64_t)coset * C * a.in_stride + row0;
#pragma unroll 8
    for (uint32_t i = threadIdx.x; i < C * (R / 2); i += blockDim.x) {     // 16-byte loads, many in flight per thread
        const uint32_t e = (i & (R / 2 - 1)) * 2, c = i >> (4 + log_ch);
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(in + (uint64_t)c * a.in_stride + e);
        smem[c * RP + e] = v.x;
        smem[c * RP + e + 1] = v.y;
    }
    __syncthreads();
    for (uint32_t item = threadIdx.x; item < C * CH; item += blockDim.x) {
        const uint32_t c = item % C, ch = item / C;
        uint64_t* col = smem + c * RP + ch * 32;
        uint64_t x[32];
#pragma unroll
        for (int k = 0; k < 32; k++) x[k] = col[k];
        dft_regs<5>(x);            // register r holds frequency bitrev_5(r): exactly the leaf order of the chunk
#pragma unroll
        for (int k = 0; k < 32; k++) col[k] = x[k];
    }
    __syncthreads();
    uint64_t* rows = a.leaves + ((uint64_t)coset * n + row0) * a.leaf_width;
    for (uint32_t e = threadIdx.x / 32; e < R; e += blockDim.x / 32)
        for (uint32_t c = threadIdx.x % 32; c < C; c += 32) rows[(uint64_t)e * a.leaf_width + c] = smem[c * RP + e];
}

// ---- the same pass as a persistent, TMA-fed pipeline --------------------------------------------------------------
// One block per SM walks over tiles.  A tile's input is C column segments of R*8 bytes: one warp requests them with
// bulk asynchronous copies (cp.async.bulk, completion counted on an mbarrier) into one of two input buffers while the
// block transforms the other; the transformed tile is written row-major into a third buffer, which for unsalted
// leaves is exactly the tile's R*C*8 contiguous bytes of the leaf matrix, and leaves the SM as ONE bulk store.  Loads
// of tile i+1, the 32-point transforms of tile i and the store of tile i-1 overlap; no thread issues a global
// load or store instruction.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// blockDim = consumer threads (a multiple of 32) + one producer warp (the last one).
// SPLIT: two threads per (column, chunk), each doing HALF of the 32-point transform -- the first DIF stage from all 32
// inputs (sums for one thread, twiddled differences for the other), then a 16-point transform on 16 registers.  Twice the
// consumer warps with fewer registers each: the transform phase of the unsplit kernel runs at 2.25 warps per scheduler
// (ncu, round 2: issue 43 %, ALU 56 %) because a tile has only 270 transforms to hand out.
constexpr int LDE_LAST32_SPLIT_MAX_THREADS = 576;
template <bool SPLIT>
__global__ void __launch_bounds__((SPLIT ? LDE_LAST32_SPLIT_MAX_THREADS : LDE_LAST32_MAX_THREADS) + 32)
lde_last32_tma_kernel(LdeLastArgs a, uint32_t log_ch, uint32_t total_tiles) {
    extern __shared__ __align__(128) uint64_t smem[];
    const uint32_t C = a.num_cols, CH = 1u << log_ch, R = 32u << log_ch, RP = R + 2;   // RP even: 16-byte aligned columns
    const size_t in_words = (size_t)C * RP;                 // input buffer b starts at smem + b * in_words
    uint64_t* out_buf = smem + 2 * in_words;                // [row][column]
    uint64_t* bars = out_buf + (size_t)R * C;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 2);   // barrier b at +8*b
    const uint64_t n = 1ull << a.log_n;
    const uint32_t tiles_per_coset = (uint32_t)(n >> (5 + log_ch));
    const uint32_t consumers = blockDim.x - 32;
    const uint32_t lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(full0 + 8 * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= consumers) {
        // ---- producer warp: C bulk copies of R*8 bytes per tile into the buffer the consumers have released ----
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const uint32_t b = it & 1;
            if (it >= 2) mbar_wait(empty0 + 8 * b, ((it >> 1) - 1) & 1);
            const uint32_t coset = tile / tiles_per_coset;
            const uint64_t row0 = (uint64_t)(tile % tiles_per_coset) * R;
            const uint64_t* src = a.in + (uint64_t)coset * C * a.in_stride + row0;
            if (lane == 0) {
                uint64_t st;
                asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 %0, [%1], %2;"
                             : "=l"(st) : "r"(full0 + 8 * b), "r"(C * R * 8u) : "memory");
                (void)st;
            }
            __syncwarp();
            for (uint32_t c = lane; c < C; c += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(smem + b * in_words + (size_t)c * RP)), "l"(src + (uint64_t)c * a.in_stride),
                                "r"(R * 8u), "r"(full0 + 8 * b) : "memory");
        }
        return;
    }

    // ---- consumers ----
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
        const uint32_t b = it & 1;
        mbar_wait(full0 + 8 * b, (it >> 1) & 1);
        // The previous tile's bulk store still reads out_buf.  Nobody writes out_buf before the first transform of this
        // tile has finished in registers, so the wait for that store (and the barrier that publishes it) sits between the
        // first transform and its write-back: the store of tile i-1 overlaps the arithmetic of tile i instead of preceding it.
        // (every consumer runs the same number of rounds, so the barrier is reached by whole warps: bar.sync is .aligned)
        const uint32_t items = C * CH * (SPLIT ? 2u : 1u);
        for (uint32_t base = 0; base < items; base += consumers) {
            const uint32_t item = base + threadIdx.x;
            const bool active = item < items;
            // SPLIT: items [0, C*CH) are the sum halves, [C*CH, 2*C*CH) the difference halves (warps stay uniform)
            const uint32_t half = SPLIT && active && item >= C * CH ? 1u : 0u, it2 = active ? item - half * C * CH : 0;
            const uint32_t c = it2 % C, ch = it2 / C;
            uint64_t x[SPLIT ? 16 : 32];
            if (active) {
                const uint64_t* col = smem + b * in_words + (size_t)c * RP + ch * 32;
                if (SPLIT) {
                    uint64_t y[16];
                    if (half == 0) {
#pragma unroll
                        for (int k = 0; k < 16; k++) y[k] = gl_add(col[k], col[k + 16]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            const uint64_t d = gl_sub(col[k], col[k + 16]);
#if QPZK_POW2_SHIFT
                            y[k] = k == 0 ? d : gl_mul_pow2_6j(d, 6 * k);          // w_32^k = 2^(6k)
#else
                            y[k] = k == 0 ? d : gl_mulc(d, pow2_mod_p(6 * k));     // w_32^k = 2^(6k)
#endif
                        }
                    }
                    dft_regs<4>(y);
#pragma unroll
                    for (int k = 0; k < 16; k++) x[k] = y[k];
                } else {
                    uint64_t y[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) y[k] = col[k];
                    dft_regs<5>(y);
#pragma unroll
                    for (int k = 0; k < (SPLIT ? 16 : 32); k++) x[k] = y[k];
                }
            }
            if (base == 0) {
                __syncwarp();   // lanes that had no item rejoin before the aligned barrier
                if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, %0;" :: "r"(consumers) : "memory");
            }
            if (active) {
                // register r of the full transform holds frequency bitrev_5(r) = leaf row r of the chunk; the sum half owns
                // registers 0..15 (even frequencies), the difference half 16..31
                uint64_t* o = out_buf + (size_t)(ch * 32 + (SPLIT ? 16 * half : 0)) * C + c;
#pragma unroll
                for (int k = 0; k < (SPLIT ? 16 : 32); k++) o[(size_t)k * C] = x[k];
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
        asm volatile("bar.sync 1, %0;" :: "r"(consumers) : "memory");
        if (threadIdx.x == 0) {
            uint64_t st;   // the input buffer may be refilled
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(empty0 + 8 * b) : "memory");
            (void)st;
            const uint32_t coset = tile / tiles_per_coset;
            const uint64_t row0 = (uint64_t)(tile % tiles_per_coset) * R;
            uint64_t* dst = a.leaves + ((uint64_t)coset * n + row0) * C;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"(smem_u32(out_buf)), "r"(R * C * 8u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static bool lde_last32_tma_ok(const LdeLastArgs& a, uint32_t* log_ch_out, size_t* smem_out) {
    if (a.leaf_width != a.num_cols) return false;          // salted leaves: rows are not contiguous
    static const int first = [] { const char* e = getenv("QPZK_TMA_LOGCH"); return e ? atoi(e) : 1; }();
    for (int log_ch = first; log_ch >= 0; log_ch--) {
        if (a.log_n < 5u + log_ch) continue;
        const size_t R = 32u << log_ch;
        const size_t smem = (2 * (size_t)a.num_cols * (R + 2) + R * a.num_cols + 4) * sizeof(uint64_t);
        if (smem > 220 * 1024) continue;
        if ((R * a.num_cols * 8) % 16) continue;
        *log_ch_out = log_ch; *smem_out = smem;
        return true;
    }
    return false;
}

bool lde_last32_supported(const LdeLastArgs& a) {
    return a.s == 5 && !a.from_coeffs && a.log_n >= 5 && a.num_cols >= 1 && (size_t)a.num_cols * 33 * 8 <= 160 * 1024;
}

cudaError_t launch_lde_last32(const LdeLastArgs& a, cudaStream_t st, int num_sms) {
    {
        uint32_t log_ch; size_t smem;
        static const bool no_tma = [] { const char* e = getenv("QPZK_NO_TMA"); return e && e[0] == '1'; }();
        const uint64_t tiles_min = (uint64_t)4 * num_sms;      // persistent pipeline only when every SM gets several tiles
        if (!no_tma && lde_last32_tma_ok(a, &log_ch, &smem) &&
            (((1ull << a.log_n) >> (5 + log_ch)) * a.num_cosets) >= tiles_min) {
            const uint64_t tiles = ((1ull << a.log_n) >> (5 + log_ch)) * a.num_cosets;
            static const bool split = [] { const char* e = getenv("QPZK_LAST32_SPLIT"); return !e || e[0] != '0'; }();
            const uint32_t items = (a.num_cols << log_ch) * (split ? 2u : 1u);
            const uint32_t max_threads = split ? LDE_LAST32_SPLIT_MAX_THREADS : LDE_LAST32_MAX_THREADS;
            uint32_t threads = (items + 31) / 32 * 32;
            if (threads > max_threads) threads = max_threads;
            if (threads < 64) threads = 64;
            static unsigned long long done = 0, done_split = 0;
            cudaError_t e = split ? raise_smem_limit((const void*)lde_last32_tma_kernel<true>, &done_split)
                                  : raise_smem_limit((const void*)lde_last32_tma_kernel<false>, &done);
            if (e != cudaSuccess) return e;
            const uint64_t resident = (uint64_t)num_sms * (smem <= 110 * 1024 ? 2 : 1);   // two pipelines per SM when they fit
            const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
            if (split) lde_last32_tma_kernel<true><<<grid, threads + 32, smem, st>>>(a, log_ch, (uint32_t)tiles);
            else lde_last32_tma_kernel<false><<<grid, threads + 32, smem, st>>>(a, log_ch, (uint32_t)tiles);
            return cudaGetLastError();
        }
    }
    // chunks per block: enough (column, chunk) items for ~256 threads, within ~96 KB of shared memory
    uint32_t log_ch = 0;
    static const size_t smem_cap = [] { const char* e = getenv("QPZK_LAST32_SMEM_KB"); return (size_t)(e ? atoi(e) : 96) * 1024; }();
    while (log_ch < 5 && ((uint64_t)a.num_cols << log_ch) < 224 && a.log_n > 5 + log_ch &&
           (size_t)a.num_cols * ((64u << log_ch) + 1) * 8 <= smem_cap) log_ch++;
    const uint32_t items = a.num_cols << log_ch;
    uint32_t threads = (items + 31) / 32 * 32;
    if (threads > LDE_LAST32_MAX_THREADS) threads = LDE_LAST32_MAX_THREADS;
    if (threads < 64) threads = 64;
    const size_t smem = (size_t)a.num_cols * ((32u << log_ch) + 1) * sizeof(uint64_t);
    static unsigned long long done32 = 0;
    cudaError_t e = raise_smem_limit((const void*)lde_last32_kernel, &done32);
    if (e != cudaSuccess) return e;
    const uint64_t blocks = ((1ull << a.log_n) >> (5 + log_ch)) * a.num_cosets;
    lde_last32_kernel<<<(unsigned)blocks, threads, smem, st>>>(a, log_ch);
    return cudaGetLastError();
}

// Short strided pass (1..4 bits): the whole 2^s-point transform of one low position fits one thread's registers, so
// there is no exchange at all.  Used for the leftover top bits when the pass plan is not made of 5..8-bit passes
// (n = 2^15, 2^16: 2 or 3 bits + 8 + 5).  256 consecutive low positions per block; same factor hoisting and column
// loop as ntt_fwd_kernel.
template <int S_LOG>
__global__ void __launch_bounds__(NTT_THREADS) ntt_fwd_small_kernel(NttPassArgs a) {
    constexpr int A = 1 << S_LOG;
    __shared__ uint64_t Ck[A];
    const bool coset_first = a.mode == NTT_FWD_COSET_FIRST;
    const uint32_t b_hi = a.b_lo + S_LOG;
    uint32_t tile = blockIdx.x, coset = 0;
    if (coset_first) { coset = tile & (a.num_cosets - 1); tile /= a.num_cosets; }
    const uint32_t llow = a.b_lo - 8;
    const uint64_t hi = tile >> llow;
    const uint64_t low = ((uint64_t)(tile & ((1u << llow) - 1)) << 8) + threadIdx.x;
    const uint64_t base = (hi << b_hi) + low;
    const uint64_t* clo = a.cs_lo + (size_t)coset * TW_LO_SIZE;
    const uint64_t* chi = a.cs_hi + (size_t)coset * a.cs_hi_size;
    if (threadIdx.x < A) Ck[threadIdx.x] = coset_first ? tw_get(clo, chi, (uint64_t)threadIdx.x << a.b_lo) : 1;
    uint64_t O[A];
    {
        const uint64_t cw = coset_first ? tw_get(clo, chi, low) : 1;
#pragma unroll
        for (int r = 0; r < A; r++) {
            uint64_t o = tw_get(a.tw_lo, a.tw_hi, ((uint64_t)brev_c<S_LOG>(r) * low) << (a.log_n - b_hi));
            if (coset_first) o = gl_mulc(o, cw);
            O[r] = o;
        }
    }
    __syncthreads();
    const uint32_t c_begin = blockIdx.y * a.cols_per_block;
    const uint32_t c_end = min((uint32_t)a.num_arrays, c_begin + a.cols_per_block);
    const uint64_t in_off = (uint64_t)blockIdx.z * a.group_stride + base;
    uint64_t x[A], y[A];
    if (c_begin < c_end) {
#pragma unroll
        for (int k = 0; k < A; k++) x[k] = a.in[c_begin * a.in_stride + in_off + ((uint64_t)k << a.b_lo)];
    }
    for (uint32_t col = c_begin; col < c_end; col++) {
#pragma unroll
        for (int k = 0; k < A; k++) y[k] = (coset_first && k) ? gl_mulc(x[k], Ck[k]) : x[k];
        if (col + 1 < c_end) {
#pragma unroll
            for (int k = 0; k < A; k++) x[k] = a.in[(col + 1) * a.in_stride + in_off + ((uint64_t)k << a.b_lo)];
        }
        dft_regs<S_LOG>(y);
        uint64_t* out = a.out + col * a.out_stride + ((uint64_t)coset + blockIdx.z) * a.group_stride + base;
#pragma unroll
        for (int r = 0; r < A; r++) out[(uint64_t)r << a.b_lo] = gl_mulc(y[r], O[r]);   // slot of frequency bitrev(r) is r
    }
}

bool ntt_fwd_supported(const NttPassArgs& a) {
    if ((a.mode == NTT_FWD_COSET_FIRST || a.mode == NTT_FWD_GENERIC) && a.s >= 1 && a.s <= 4 && a.scale == 1 &&
        a.b_lo >= 8 && a.log_n >= a.b_lo + a.s) return true;
    if (a.mode != NTT_FWD_COSET_FIRST && a.mode != NTT_FWD_GENERIC) return false;
    if (a.s < 5 || a.s > 8 || a.scale != 1) return false;
    return a.b_lo >= 12 - a.s && a.log_n >= a.b_lo + a.s;
}

template <int SA, int SB>
static cudaError_t launch_one(const NttPassArgs& a, dim3 grid, cudaStream_t st) {
    constexpr int S = 1 << (SA + SB);
    static const bool lean = [] { const char* e = getenv("QPZK_NTT_LEAN"); return !e || e[0] != '0'; }();
    const size_t smem = ((size_t)4096 + S + (1 << SA) + (lean ? 16 * NTT_THREADS : 0)) * sizeof(uint64_t);
    if (lean) {
        static unsigned long long done = 0;   // 65 KB of dynamic shared memory per block (two blocks per SM)
        cudaError_t e = raise_smem_limit((const void*)ntt_fwd_kernel<SA, SB, true>, &done);
        if (e != cudaSuccess) return e;
        ntt_fwd_kernel<SA, SB, true><<<grid, NTT_THREADS, smem, st>>>(a);
    } else {
        ntt_fwd_kernel<SA, SB, false><<<grid, NTT_THREADS, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_ntt_fwd(const NttPassArgs& a, cudaStream_t st) {
    if (a.s <= 4) {
        uint64_t t = (1ull << a.log_n) >> (8 + a.s);
        if (a.mode == NTT_FWD_COSET_FIRST) t *= a.num_cosets;
        const uint32_t cpb = a.cols_per_block ? a.cols_per_block : 1;
        dim3 grid((unsigned)t, (unsigned)((a.num_arrays + cpb - 1) / cpb), (unsigned)(a.num_groups ? a.num_groups : 1));
        switch (a.s) {
            case 1: ntt_fwd_small_kernel<1><<<grid, NTT_THREADS, 0, st>>>(a); break;
            case 2: ntt_fwd_small_kernel<2><<<grid, NTT_THREADS, 0, st>>>(a); break;
            case 3: ntt_fwd_small_kernel<3><<<grid, NTT_THREADS, 0, st>>>(a); break;
            default: ntt_fwd_small_kernel<4><<<grid, NTT_THREADS, 0, st>>>(a); break;
        }
        return cudaGetLastError();
    }
    uint64_t tiles = (1ull << a.log_n) >> 12;
    if (a.mode == NTT_FWD_COSET_FIRST) tiles *= a.num_cosets;
    const uint32_t cpb = a.cols_per_block ? a.cols_per_block : 1;
    dim3 grid((unsigned)tiles, (unsigned)((a.num_arrays + cpb - 1) / cpb), (unsigned)(a.num_groups ? a.num_groups : 1));
    switch (a.s) {
        case 5: return launch_one<3, 2>(a, grid, st);
        case 6: return launch_one<3, 3>(a, grid, st);
        case 7: return launch_one<4, 3>(a, grid, st);
        case 8: return launch_one<4, 4>(a, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace qpzk
