// Quotient polynomials of the permutation argument (plus caller-supplied gate constraint terms) on the device.
//
// Replaces the gate-independent part of qp-plonky2 =1.5.5 `compute_quotient_polys` (plonk/prover.rs) with
// `eval_vanishing_poly_base_batch` (plonk/vanishing_poly.rs), `check_partial_products` (util/partial_products.rs),
// `ZeroPolyOnCoset` (plonk/plonk_common.rs) and `reduce_with_powers_multi` -- SURVEY.md 8a row a17, a "next" row:
// the step between the zs/partial-products commitment and the quotient commitment of `prove()` (reference call
// site wormhole/prover/src/lib.rs:171-175).  What is NOT here: the gate constraint evaluators (the gate set lives in
// the un-vendored crate); their values enter as an optional matrix of "gate terms", one row per quotient-domain
// point, which take the places upstream gives `constraint_terms` in the alpha combination.  PARITY UNPINNED like the
// rest of the path: checked bit for bit against the CPU restatement, and through the verifier's identity
// vanishing(zeta) = Z_H(zeta) * sum_k zeta^(k n) chunk_k(zeta) at a random point, which holds only when the wires
// satisfy the copy constraints.
//
// Quotient domain: the coset g<w_{nQ}>, Q = quotient_degree_factor = 2^qbits <= 2^rate_bits; its point with natural
// index i is LDE row bitrev_{log nQ}(i) -- the FIRST nQ rows of every leaf matrix, in leaf order, so one thread per
// leaf row reads its operands from the three committed batches without any gather except the row of Z(g x)
// (natural index i + Q).  Per point and challenge c:
//     z1_c      = L_0(x) (Z_c(x) - 1),                     L_0(x) = Z_H(x) / (n (x - 1)),  Z_H(x) = g^n w_Q^(i mod Q) - 1
//     pp_{c,k}  = acc_{k-1} prod_{j in chunk k} (w_j + beta_c k_j x + gamma_c) - acc_k prod_{j in chunk k} (w_j + beta_c sigma_j + gamma_c)
//                 with acc_{-1} = Z_c(x), acc_k = partial product k, acc_last = Z_c(g x)
//     terms     = [z1_0.., pp_{0,0}.., pp_{1,0}.., gate terms..],   res_a = sum_t terms[t] alpha_a^t,   q_a(x) = res_a / Z_H(x)
// The values q_a (natural order) go through an inverse NTT of size nQ and the coset un-shift g^-k (capi.cu); the nQ
// coefficients of challenge a are, n at a time, the Q quotient chunks upstream commits.
#include "internal.cuh"
#include "goldilocks.cuh"
#include "ntt_local.cuh"

namespace qpzk {

constexpr int QUOT_THREADS = 128;

template <int NCH>
__global__ void __launch_bounds__(QUOT_THREADS) quotient_perm_kernel(QuotientArgs a) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t log_q = a.log_n + a.qbits;
    const uint64_t Lq = 1ull << log_q;
    if (j >= Lq) return;
    const uint64_t i = log_q ? __brevll(j) >> (64 - log_q) : 0;
    const uint64_t Q = 1ull << a.qbits;
    const uint64_t inext = (i + Q) & (Lq - 1);
    const uint64_t jn = log_q ? __brevll(inext) >> (64 - log_q) : 0;
    const uint64_t x = gl_mul(GL_GENERATOR, tw_get(a.tw_lo, a.tw_hi, i));
    const uint64_t zh = a.zh[i & (Q - 1)], zh_inv = a.zh[Q + (i & (Q - 1))];
    const uint64_t l0 = gl_mul(zh, gl_inv(gl_mul(a.n_field, gl_sub(x, 1))));
    const uint64_t* __restrict__ wrow = a.w_leaves + j * a.w_width;
    const uint64_t* __restrict__ srow = a.cs_leaves + j * a.cs_width + a.sigma_col0;
    const uint64_t* __restrict__ zrow = a.z_leaves + j * a.z_width;
    const uint64_t* __restrict__ znrow = a.z_leaves + jn * a.z_width;
    const uint32_t T = a.num_terms, np = a.nchunks - 1;
    uint64_t res[NCH], beta[NCH], gamma[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) { res[c] = 0; beta[c] = a.betas[c]; gamma[c] = a.gammas[c]; }
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const uint64_t term = gl_mul(l0, gl_sub(zrow[c], 1));
#pragma unroll
        for (int e = 0; e < NCH; e++) res[e] = gl_add(res[e], gl_mul(term, a.apow[(size_t)e * T + c]));
    }
    for (uint32_t k = 0; k < a.nchunks; k++) {
        uint64_t pn[NCH], pd[NCH];
#pragma unroll
        for (int c = 0; c < NCH; c++) pn[c] = pd[c] = 1;
        const uint32_t j0 = k * a.chunk, j1 = min(a.num_routed, j0 + a.chunk);
        // operands in groups of 8 words = 64 B of the thread's own rows, loaded back to back so that both sectors of
        // a group are fetched once (the access pattern of the leaf-hash kernel)
        for (uint32_t r0 = j0; r0 < j1; r0 += 8) {
            uint64_t w[8], sg[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const bool in = r0 + u < j1;
                w[u] = in ? __ldg(wrow + r0 + u) : 0;
                sg[u] = in ? __ldg(srow + r0 + u) : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (r0 + u >= j1) break;
#pragma unroll
                for (int c = 0; c < NCH; c++) {   // arbitrary u64 representatives inside the products
                    const uint64_t wg = add_canon_p(w[u], gamma[c]);
                    pn[c] = mulp(pn[c], mul_addp(a.beta_k[(size_t)c * a.num_routed + r0 + u], x, wg));
                    pd[c] = mulp(pd[c], mul_addp(beta[c], sg[u], wg));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const uint64_t prev = k == 0 ? zrow[c] : zrow[NCH + c * np + k - 1];
            const uint64_t next = k == np ? znrow[c] : zrow[NCH + c * np + k];
            const uint64_t term = gl_sub(gl_canon(mulp(prev, pn[c])), gl_canon(mulp(next, pd[c])));
            const uint32_t t = NCH + c * a.nchunks + k;
#pragma unroll
            for (int e = 0; e < NCH; e++) res[e] = gl_add(res[e], gl_mul(term, a.apow[(size_t)e * T + t]));
        }
    }
    if (a.num_gate_terms) {
        const uint64_t* __restrict__ grow = a.gate_terms + j * a.num_gate_terms;
        const uint32_t t0 = NCH + NCH * a.nchunks;
        for (uint32_t g = 0; g < a.num_gate_terms; g++) {
            const uint64_t term = __ldg(grow + g);
#pragma unroll
            for (int e = 0; e < NCH; e++) res[e] = gl_add(res[e], gl_mul(term, a.apow[(size_t)e * T + t0 + g]));
        }
    }
#pragma unroll
    for (int e = 0; e < NCH; e++) a.values[(size_t)e * Lq + i] = gl_mul(res[e], zh_inv);
}

// coefficient k of every array *= shift^-k   (second half of coset_ifft)
__global__ void __launch_bounds__(256) coset_unshift_kernel(uint64_t* __restrict__ coeffs, uint64_t len, uint64_t total,
                                                            const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    coeffs[t] = gl_mul(coeffs[t], tw_get(lo, hi, t & (len - 1)));
}

// ---- gate constraints as a straight-line program ---------------------------------------------------------------------
// The gate set of a circuit lives in the prover's crate, not here; what this library can do is EVALUATE it on the device
// when it is handed over as data: a register program over the base field (what upstream's `eval_unfiltered_base_batch`
// computes per point, selector filters included) that the host side obtains by running its gates' constraint code over
// a recording field type.  One thread per quotient-domain point interprets the program -- every thread executes the same
// instruction, so there is no divergence; the register file sits in local memory -- and writes one row of the gate-term
// matrix that quotient_perm_kernel consumes.
//   op  dst a b      0 LOAD_WIRE  r[dst] = wires[a]            3 ADD  r[dst] = r[a] + r[b]      6 EMIT  term = r[a]
//                    1 LOAD_CONST r[dst] = constants_sigmas[a] 4 SUB  r[dst] = r[a] - r[b]
//                    2 LOAD_IMM   r[dst] = imm[a]              5 MUL  r[dst] = r[a] * r[b]
//   fused forms (third source register c in bits 8..15 of the op word; the recorder's peephole pass produces them):
//                    7 MULADD  r[dst] = r[a] * r[b] + r[c]     8 MULI r[dst] = r[a] * imm[b]    9 ADDI r[dst] = r[a] + imm[b]
//                    10 MULIADD r[dst] = r[a] * imm[b] + r[c]
// A linear layer is multiply-adds by constants: fused, upstream's PoseidonGate shrinks from 5290 to about half as
// many steps, and each step is one 128-bit product with the addend folded in before the single reduction.
template <int NR>   // size of the register file: 64 (fits L1 comfortably) or GATE_MAX_REGS
__global__ void __launch_bounds__(QUOT_THREADS) gate_program_kernel(GateProgramArgs a) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.num_points) return;
    const uint64_t* __restrict__ wrow = a.w_leaves + j * a.w_width;
    const uint64_t* __restrict__ crow = a.cs_leaves + j * a.cs_width;
    uint64_t* __restrict__ out = a.terms + j * a.num_terms;
    uint64_t r[NR];
    uint32_t emitted = 0;
    for (uint32_t pc = 0; pc < a.num_ops; pc++) {
        const uint4 ins = __ldg(reinterpret_cast<const uint4*>(a.code) + pc);
        switch (ins.x & 0xffu) {
            case 0: r[ins.y] = __ldg(wrow + ins.z); break;
            case 1: r[ins.y] = __ldg(crow + ins.z); break;
            case 2: r[ins.y] = __ldg(a.imm + ins.z); break;
            case 3: r[ins.y] = gl_add(r[ins.z], r[ins.w]); break;
            case 4: r[ins.y] = gl_sub(r[ins.z], r[ins.w]); break;
            case 5: r[ins.y] = gl_mul(r[ins.z], r[ins.w]); break;
            case 7: r[ins.y] = gl_canon(mul_addp(r[ins.z], r[ins.w], r[(ins.x >> 8) & 0xffu])); break;
            case 8: r[ins.y] = gl_mul(r[ins.z], __ldg(a.imm + ins.w)); break;
            case 9: r[ins.y] = gl_add(r[ins.z], __ldg(a.imm + ins.w)); break;
            case 10: r[ins.y] = gl_canon(mul_addp(r[ins.z], __ldg(a.imm + ins.w), r[(ins.x >> 8) & 0xffu])); break;
            default: out[emitted++] = r[ins.z]; break;
        }
    }
}

cudaError_t launch_gate_program(const GateProgramArgs& a, cudaStream_t st) {
    if (!a.num_points) return cudaSuccess;
    const unsigned blocks = (unsigned)((a.num_points + QUOT_THREADS - 1) / QUOT_THREADS);
    if (a.num_regs <= 64) gate_program_kernel<64><<<blocks, QUOT_THREADS, 0, st>>>(a);
    else gate_program_kernel<GATE_MAX_REGS><<<blocks, QUOT_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_quotient_perm(const QuotientArgs& a, cudaStream_t st) {
    const uint64_t Lq = 1ull << (a.log_n + a.qbits);
    const unsigned blocks = (unsigned)((Lq + QUOT_THREADS - 1) / QUOT_THREADS);
    switch (a.num_challenges) {
        case 1: quotient_perm_kernel<1><<<blocks, QUOT_THREADS, 0, st>>>(a); break;
        case 2: quotient_perm_kernel<2><<<blocks, QUOT_THREADS, 0, st>>>(a); break;
        case 3: quotient_perm_kernel<3><<<blocks, QUOT_THREADS, 0, st>>>(a); break;
        case 4: quotient_perm_kernel<4><<<blocks, QUOT_THREADS, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_coset_unshift(uint64_t* coeffs, uint32_t log_len, uint32_t num_arrays, const uint64_t* lo,
                                 const uint64_t* hi, cudaStream_t st) {
    const uint64_t len = 1ull << log_len, total = len * num_arrays;
    coset_unshift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(coeffs, len, total, lo, hi);
    return cudaGetLastError();
}

}  // namespace qpzk
