// Internal declarations shared by the translation units of libqpzk.so (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qpzk {

// Opt-in dynamic shared memory limit of sm_100 (227 KB).  Kernels that may need more than 48 KB get THIS value set
// before every launch: the attribute is process-wide per function, and setting a per-launch size from several host
// threads races (one thread lowers it between another thread's set and launch); a constant is idempotent.
constexpr int MAX_DYN_SMEM = 227 * 1024;
// Raise a kernel's dynamic shared memory limit to MAX_DYN_SMEM once per (kernel, device).  `done` is a per-kernel
// bit mask of devices (callers pass a function-local static).  cudaFuncSetAttribute takes a device-wide lock; calling
// it before every launch made concurrent provers on several host threads several times slower.
cudaError_t raise_smem_limit(const void* fn, unsigned long long* done);
#ifndef QPZK_HASH_THREADS
#define QPZK_HASH_THREADS 128
#endif
constexpr int HASH_THREADS = QPZK_HASH_THREADS;          // threads per block for one-thread-per-permutation kernels

// Digest buffer of a tree with N leaves: level 0 (leaf digests, 4N words) first, then level 1 (2N words), ...
// Word offset of level l:  8N - (8N >> l).
__host__ __device__ __forceinline__ uint64_t digest_level_off(uint64_t num_leaves, uint32_t l) {
    return 8 * num_leaves - ((8 * num_leaves) >> l);
}

// Device-resident Fiat-Shamir challenger; same field order as qpzk_challenger_state in include/qpzk.h.
struct ChallengerDev {
    uint64_t state[12];
    uint64_t in[8];
    uint64_t out[8];
    uint32_t in_len;
    uint32_t out_len;
};

// ---- ntt.cu ----
constexpr int NTT_THREADS = 256;
constexpr uint32_t TW_LO_BITS = 10;             // two-level power tables: w^idx = lo[idx & 1023] * hi[idx >> 10]
constexpr uint32_t TW_LO_SIZE = 1u << TW_LO_BITS;

enum NttMode : uint32_t {
    NTT_INV_FIRST = 0,        // DIT, position bits [0, s); gathers the natural-order input at bit-reversed addresses
    NTT_INV_GENERIC = 1,      // DIT, position bits [b_lo, b_lo+s), pre-twiddle
    NTT_FWD_COSET_FIRST = 2,  // DIF, top bits [b_lo, log_n); scales by the coset powers on load; one output per coset
    NTT_FWD_GENERIC = 3,      // DIF, position bits [b_lo, b_lo+s), post-twiddle
};

struct NttPassArgs {
    const uint64_t* in;
    uint64_t* out;
    uint64_t in_stride, out_stride;  // words between consecutive arrays
    uint64_t num_arrays;             // grid.y: arrays per group
    uint64_t num_groups;             // grid.z (FWD_GENERIC: one group per coset; 1 otherwise)
    uint64_t group_stride;           // words between the same array of consecutive groups / cosets (in and out)
    uint32_t log_n, b_lo, s, W;
    uint32_t mode, num_cosets;
    const uint64_t *tw_lo, *tw_hi;   // powers of w_n (forward) or w_n^-1 (inverse)
    const uint64_t *cs_lo, *cs_hi;   // per-coset powers of s_j = g * w_L^bitrev(j)
    uint64_t cs_hi_size;
    uint64_t scale;                  // multiplied into every output (1/n on the last inverse pass), 1 = none
    uint32_t cols_per_block;         // ntt_fwd_kernel only: columns handled by one block (grid.y = ceil(num_arrays / it))
};

struct LdeLastArgs {
    const uint64_t* in;   // from_coeffs ? coefficients [C][n] : intermediate [coset][C][n]
    uint64_t in_stride;
    uint64_t* leaves;     // [2^rate_bits * n][leaf_width]
    uint32_t log_n, s, num_cols, leaf_width, num_cosets, from_coeffs;
    const uint64_t *tw_lo, *tw_hi, *cs_lo, *cs_hi;
    uint64_t cs_hi_size;
};

cudaError_t launch_ntt_pass(const NttPassArgs& a, cudaStream_t st);
bool ntt_fwd_supported(const NttPassArgs& a);                          // ntt_fwd.cu: register-resident forward passes
cudaError_t launch_ntt_fwd(const NttPassArgs& a, cudaStream_t st);
constexpr int LDE_LAST32_MAX_THREADS = 384;
bool lde_last32_supported(const LdeLastArgs& a);                       // ntt_fwd.cu: 32-point in-register last pass
cudaError_t launch_lde_last32(const LdeLastArgs& a, cudaStream_t st, int num_sms);
cudaError_t launch_lde_last(const LdeLastArgs& a, cudaStream_t st);
cudaError_t launch_fill_words(void* p, uint64_t words, uint64_t value, cudaStream_t st);   // replaces cudaMemsetAsync
cudaError_t launch_bitrev_copy(const uint64_t* in, uint64_t* out, uint32_t log_n, uint64_t num_arrays,
                               uint64_t in_stride, uint64_t out_stride, cudaStream_t st);
cudaError_t launch_salt_scatter(const uint64_t* salt, uint64_t* leaves, uint32_t log_L, uint32_t num_cols,
                                uint32_t leaf_width, uint32_t salt_cols, cudaStream_t st);

// ---- hash.cu ----
cudaError_t upload_poseidon_constants();
cudaError_t launch_hash_rows(const uint64_t* leaves, uint64_t num_rows, uint32_t leaf_len, uint64_t row_stride,
                             uint64_t* digests, cudaStream_t st);
cudaError_t launch_permute(const uint64_t* in, uint64_t count, int mds_only_row, uint64_t* out, cudaStream_t st);
cudaError_t launch_merkle_levels(uint64_t* digests, uint32_t num_levels, uint64_t num_leaves, cudaStream_t st);
uint32_t merkle_levels_launches(uint32_t num_levels, uint64_t num_leaves);
cudaError_t launch_challenger_observe(ChallengerDev* c, const uint64_t* elems_dev, uint32_t n, cudaStream_t st);
cudaError_t launch_challenger_get(ChallengerDev* c, uint64_t* out_dev, uint32_t n, cudaStream_t st);
// forced_witness == ~0: grind the MINIMUM witness below `max_candidates` in one launch (no host round trip), then
// observe it and draw the query indices; otherwise use forced_witness.  status[0]: 1 ok, 0 witness too weak, 2 none found.
cudaError_t launch_pow(ChallengerDev* c, uint32_t pow_bits, uint64_t forced_witness, uint32_t num_queries,
                       uint64_t lde_size, uint64_t max_candidates, int num_sms,
                       unsigned long long* found_dev, uint64_t* witness_out, uint64_t* indices, uint32_t* status,
                       cudaStream_t st);
// PoseidonHash::hash_no_pad of `len` device words (one warp; len == 0 gives four zeros): public_inputs_hash of prove()
cudaError_t launch_hash_no_pad(const uint64_t* in_dev, uint32_t len, uint64_t* out_dev, cudaStream_t st);

// ---- poseidon2.cu (the reference's application hash; not on the commitment path) ----
cudaError_t upload_poseidon2_constants();
cudaError_t launch_poseidon2_permute(const uint64_t* in, uint64_t count, uint64_t* out, cudaStream_t st);
cudaError_t launch_poseidon2_hash(const uint64_t* msgs, uint64_t count, uint32_t len, uint64_t stride, uint64_t* out,
                                  cudaStream_t st);

// ---- fri.cu ----
constexpr int EVAL_THREADS = 256;
constexpr int COMBINE_THREADS = 256;
constexpr uint32_t MAX_ORACLES = 8;
constexpr uint32_t MAX_FRI_ROUNDS = 16;
constexpr uint32_t MAX_FRI_BATCHES = 8;

struct FriCombineArgs {
    uint32_t log_n, rate_bits, num_batches, num_oracles;
    const uint64_t* leaves[MAX_ORACLES];
    uint32_t leaf_width[MAX_ORACLES];
    uint32_t batch_off[MAX_FRI_BATCHES + 1];  // prefix sums of batch sizes
    uint32_t eval_off[MAX_FRI_BATCHES * MAX_ORACLES];  // ext offset of evals[b][o] inside `evals`
    uint64_t points[2 * MAX_FRI_BATCHES];
    const uint64_t* points_dev;                // device, 2 * num_batches words; overrides `points` when not null
    const uint32_t *poly_oracle, *poly_index;  // device
    const uint64_t* apow;      // device: alpha^j, ext
    const uint64_t* evals;     // device: per (batch, oracle) evaluations of every column at the batch point
    uint64_t* reduced;         // device: 2 * num_batches
    uint64_t* openings_out;    // device: 2 * total polys, instance order
    uint64_t* values;          // device: 2 * L, leaf order
    const uint64_t *tw_lo, *tw_hi;  // powers of w_n
    const uint64_t* cs_lo;          // per-coset powers (entry 1 = the coset shift itself)
};

struct FriFoldArgs {
    const uint64_t* in;
    uint64_t* out;
    uint32_t lg, log_L;
    const uint64_t* beta;          // device, ext
    uint64_t shift_inv, arity_inv;
    uint64_t itw[8];               // w_arity^-j
    const uint64_t *itw_lo, *itw_hi;  // powers of w_L^-1
};

struct FriFinalArgs {
    const uint64_t* in;
    uint64_t* out;
    uint32_t lg, log_L, rate_bits;
    uint64_t shift_inv, len_inv;
    const uint64_t *itw_lo, *itw_hi;
};

struct FriTreeDesc {
    const uint64_t* leaves;
    const uint64_t* digests;
    uint32_t leaf_len, log_leaves, cap_height, arity_bits;
};

struct FriGatherArgs {
    FriTreeDesc trees[MAX_ORACLES + MAX_FRI_ROUNDS];
    uint32_t num_trees, num_oracles;
    const uint64_t* indices;  // device
    const uint32_t* status;   // device, may be null; the kernel does nothing unless status[0] == 1
    uint64_t* proof;          // device
    uint64_t query_base, query_words;
};

cudaError_t launch_eval_ext(const uint64_t* coeffs, uint64_t col_stride, uint32_t log_n, uint32_t num_cols,
                            const uint64_t z[2], uint64_t* out_dev, cudaStream_t st);
// several (coefficient matrix, point) jobs of the same degree in ONE launch (a proof evaluates four oracles at zeta
// and one at g*zeta: five dependent-latency launches otherwise)
constexpr uint32_t MAX_EVAL_JOBS = MAX_ORACLES * MAX_FRI_BATCHES;
struct EvalMultiArgs {
    uint32_t num_jobs, log_n;
    uint64_t col_stride;
    const uint64_t* coeffs[MAX_EVAL_JOBS];
    uint64_t* out[MAX_EVAL_JOBS];
    uint64_t z[2 * MAX_EVAL_JOBS];
    const uint64_t* z_dev[MAX_EVAL_JOBS];    // device point of job j (2 words); overrides z when not null
    uint32_t col_begin[MAX_EVAL_JOBS + 1];   // prefix sums of the jobs' column counts
};
cudaError_t launch_eval_ext_multi(const EvalMultiArgs& a, cudaStream_t st);
cudaError_t launch_alpha_powers(const uint64_t* alpha_dev, uint64_t* apow_dev, uint32_t count, cudaStream_t st);
cudaError_t launch_fri_combine(const FriCombineArgs& a, cudaStream_t st);
// openings_out[k] = evals[...] in FRI-instance order, without the alpha reduction (the prover observes the openings
// BEFORE alpha is drawn; launch_fri_combine writes the same values again)
cudaError_t launch_openings_gather(const FriCombineArgs& a, uint32_t total_polys, cudaStream_t st);
// out (ext) = in (ext) * s:  g * zeta for the openings of Z at the next row
cudaError_t launch_ext_scale(const uint64_t* in_dev, uint64_t s, uint64_t* out_dev, cudaStream_t st);
cudaError_t launch_fri_fold(const FriFoldArgs& a, uint32_t arity_bits, cudaStream_t st);
cudaError_t launch_fri_final_poly(const FriFinalArgs& a, cudaStream_t st);
cudaError_t launch_fri_gather(const FriGatherArgs& a, uint32_t num_queries, cudaStream_t st);

// ---- perm.cu ----
constexpr uint32_t PP_MAX_CHUNKS = 32;       // chunks per row: ceil(num_routed / chunk)
struct PartialProductsArgs {
    const uint64_t *wires, *sigmas;   // num_routed columns of n values, column-major
    uint64_t col_stride;              // words between consecutive columns of wires / sigmas
    uint32_t log_n, num_routed, chunk, nchunks, num_challenges;
    const uint64_t *betas, *gammas;   // device, num_challenges each
    const uint64_t* beta_k;           // device, [challenge][num_routed]: beta * g^j
    const uint64_t *tw_lo, *tw_hi;    // powers of w_n
    uint64_t* chunkprod;              // scratch: [challenge][chunk][row]
    uint64_t* rowprod;                // scratch: [challenge][row]
    uint64_t* blockprod;              // scratch: [challenge][pp_num_blocks(log_n)]
    uint64_t* out;                    // [Z of every challenge | partial products of challenge 0, 1, ...] x n
};
inline unsigned pp_num_blocks(uint32_t log_n) { return (unsigned)(((1ull << log_n) + 127) / 128); }
cudaError_t launch_partial_products(const PartialProductsArgs& a, cudaStream_t st);
// beta_k[c][j] = betas[c] * kpow[j] from device-resident challenges (kpow[j] = g^j, the coset shifts k_j)
cudaError_t launch_perm_params(const uint64_t* betas_dev, const uint64_t* kpow_dev, uint32_t num_challenges,
                               uint32_t num_routed, uint64_t* beta_k_dev, cudaStream_t st);
// apow[c][t] = alphas[c]^t, t < num_terms, from device-resident challenges
cudaError_t launch_base_powers(const uint64_t* alphas_dev, uint32_t num_challenges, uint32_t num_terms,
                               uint64_t* apow_dev, cudaStream_t st);

// ---- quotient.cu ----
constexpr uint32_t QUOT_MAX_CHALLENGES = 4;
struct QuotientArgs {
    const uint64_t *cs_leaves, *w_leaves, *z_leaves;   // leaf matrices of the three batches (row-major, leaf order)
    uint32_t cs_width, w_width, z_width, sigma_col0;
    uint32_t log_n, qbits, num_routed, chunk, nchunks, num_challenges, num_gate_terms, num_terms;
    const uint64_t* gate_terms;        // [n << qbits][num_gate_terms], row = leaf index, or null
    const uint64_t *betas, *gammas;    // device
    const uint64_t* beta_k;            // device, [challenge][num_routed]: beta * g^j
    const uint64_t* apow;              // device, [challenge][num_terms]: alpha^t
    const uint64_t* zh;                // device, [2 << qbits]: Z_H on the coset, then its inverses
    uint64_t n_field;                  // n as a field element
    const uint64_t *tw_lo, *tw_hi;     // powers of w_{n << qbits}
    uint64_t* values;                  // [challenge][n << qbits], natural order
};
cudaError_t launch_quotient_perm(const QuotientArgs& a, cudaStream_t st);
constexpr uint32_t GATE_MAX_REGS = 256;
struct GateProgramArgs {
    const uint64_t *cs_leaves, *w_leaves;   // leaf matrices (row = quotient-domain point in leaf order)
    uint32_t cs_width, w_width;
    const uint32_t* code;                   // device, num_ops x {op, dst, a, b}
    const uint64_t* imm;                    // device
    uint32_t num_ops, num_terms, num_regs;
    uint64_t num_points;
    uint64_t* terms;                        // device, [num_points][num_terms]
};
cudaError_t launch_gate_program(const GateProgramArgs& a, cudaStream_t st);
cudaError_t launch_coset_unshift(uint64_t* coeffs, uint32_t log_len, uint32_t num_arrays, const uint64_t* lo,
                                 const uint64_t* hi, cudaStream_t st);

}  // namespace qpzk

namespace qpzk {
// ---- bench.cu ----
cudaError_t run_imad_bench(uint32_t iters, int num_sms, cudaStream_t st, float* ms, double* ops);
cudaError_t launch_field_ops(uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t count, uint64_t* out, cudaStream_t st);
cudaError_t run_poseidon_bench(uint32_t perms, int num_sms, cudaStream_t st, float* ms, double* total);
}  // namespace qpzk
