/* ORACLE (test infrastructure, NOT product code) -- public C interface of the CPU restatement.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libqpzk.so) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" for LDE values / Merkle caps / FRI transcripts -- the reference
 * tree holds no golden vector for them and the reference cannot be compiled here
 * (no cargo/rustc, the algorithm lives in the un-vendored crate qp-plonky2 =1.5.5).  What IS pinned:
 * the Goldilocks generators, the complete Poseidon-12 permutation (three upstream KATs) and the
 * complete Poseidon2 hash of the application layer (seven vectors held by the reference's own tests), checked by
 * tests/test_oracle_kat.py.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- field helpers exported for tests ---- */
uint64_t orc_gl_mul(uint64_t a, uint64_t b);
uint64_t orc_gl_inv(uint64_t a);
uint64_t orc_gl_exp(uint64_t a, uint64_t e);
uint64_t orc_gl_root_of_unity(unsigned log_n);

/* ---- Poseidon-12 ---- */
void orc_poseidon_rc(uint64_t out[360]);   /* round constants */
/* partial products of the permutation argument (row a16, next tier, parity unpinned); see oracle.c */
int orc_partial_products(const uint64_t* wires, const uint64_t* sigmas, unsigned log_n, unsigned num_routed,
                         unsigned chunk, unsigned num_challenges, const uint64_t* betas, const uint64_t* gammas,
                         uint64_t* out);
void orc_poseidon_naive(uint64_t s[12]);  /* literal round structure */
void orc_poseidon(uint64_t s[12]);        /* sparse partial rounds; identical function */
void orc_hash_no_pad(const uint64_t* in, size_t n, uint64_t out[4]);
void orc_hash_or_noop(const uint64_t* in, size_t n, uint64_t out[4]);
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);

/* ---- Poseidon2 (the reference's application hash; PARITY PINNED by 7 reference-held vectors, see oracle.c) ---- */
void orc_poseidon2(uint64_t s[12]);
void orc_hash_no_pad_p2(const uint64_t* in, size_t n, uint64_t out[4]);

/* ---- FFT: natural order in and out ---- */
void orc_fft(uint64_t* a, unsigned log_n);
void orc_ifft(uint64_t* a, unsigned log_n);
void orc_coset_fft(uint64_t* a, unsigned log_n, uint64_t shift);

/* ---- MerkleTree ---- */
typedef struct orc_tree orc_tree;
orc_tree* orc_merkle_new(const uint64_t* leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height);
void orc_merkle_free(orc_tree* t);
size_t orc_merkle_cap(const orc_tree* t, uint64_t* out);                           /* returns #words */
size_t orc_merkle_prove(const orc_tree* t, size_t leaf_index, uint64_t* siblings); /* returns #siblings */
const uint64_t* orc_merkle_leaf(const orc_tree* t, size_t leaf_index);
int orc_merkle_verify(const uint64_t* leaf, size_t leaf_len, size_t leaf_index, const uint64_t* cap,
                      unsigned cap_height, const uint64_t* siblings, size_t num_siblings);

/* ---- PolynomialBatch ---- */
typedef struct orc_batch orc_batch;
/* cols: num_cols x 2^log_n, column-major canonical felts.  salt: NULL or 4 x 2^(log_n+rate_bits) column-major. */
orc_batch* orc_batch_from_values(const uint64_t* cols, unsigned log_n, size_t num_cols, unsigned rate_bits,
                                 unsigned cap_height, const uint64_t* salt);
orc_batch* orc_batch_from_coeffs(const uint64_t* cols, unsigned log_n, size_t num_cols, unsigned rate_bits,
                                 unsigned cap_height, const uint64_t* salt);
void orc_batch_free(orc_batch* b);
const orc_tree* orc_batch_tree(const orc_batch* b);
const uint64_t* orc_batch_coeffs(const orc_batch* b);   /* num_cols x n column-major */
size_t orc_batch_leaf_len(const orc_batch* b);
/* evaluate every (unsalted) polynomial at an extension point; out = num_cols x 2 words */
void orc_batch_eval_ext(const orc_batch* b, const uint64_t point[2], uint64_t* out);

/* ---- size-independent checks for the benchmarked sizes (definitions, not the FFT; see oracle.c) ---- */
int orc_eval_from_values(const uint64_t* values, size_t col_stride, size_t num_cols, unsigned log_n, uint64_t r, uint64_t* out);
int orc_eval_from_coeffs(const uint64_t* coeffs, size_t col_stride, size_t num_cols, unsigned log_n, uint64_t r, uint64_t* out);
int orc_lde_subgroup_rows(const uint64_t* coeffs, size_t col_stride, size_t num_cols, unsigned log_n, unsigned rate_bits,
                          uint64_t i0, unsigned log_m, uint64_t* out);

/* quotient polynomials of the permutation argument + caller-supplied gate terms (row a17, next tier, parity unpinned) */
int orc_quotient_polys(const orc_batch* cs, const orc_batch* wires, const orc_batch* zs, unsigned sigma_col0,
                       unsigned num_routed, unsigned chunk, unsigned num_challenges, const uint64_t* betas,
                       const uint64_t* gammas, const uint64_t* alphas, const uint64_t* gate_terms,
                       unsigned num_gate_terms, uint64_t* out);

/* gate constraints as a straight-line program (the format of qpzk_gate_program): [n Q][num_terms] gate-term matrix */
int orc_gate_terms(const orc_batch* cs, const orc_batch* wires, unsigned chunk, const uint32_t* code, unsigned num_ops,
                   const uint64_t* imm, uint64_t* out);

/* ---- Challenger ---- */
typedef struct {
    uint64_t state[12];
    uint64_t in[8];
    uint64_t out[8];
    uint32_t in_len;
    uint32_t out_len;
} orc_challenger;
void orc_challenger_init(orc_challenger* c);
void orc_challenger_observe(orc_challenger* c, const uint64_t* elems, size_t n);
void orc_challenger_get(orc_challenger* c, uint64_t* out, size_t n);

/* ---- FRI ---- */
#define ORC_MAX_FRI_ROUNDS 16
typedef struct {
    uint32_t rate_bits, cap_height, pow_bits, num_queries;
    uint32_t num_rounds;
    uint32_t arity_bits[ORC_MAX_FRI_ROUNDS];
} orc_fri_params;
/* ConstantArityBits(arity_bits, final_poly_bits) reduction strategy */
void orc_fri_params_constant_arity(orc_fri_params* p, unsigned log_n, unsigned rate_bits, unsigned cap_height,
                                   unsigned pow_bits, unsigned num_queries, unsigned arity_bits,
                                   unsigned final_poly_bits);
typedef struct {
    uint32_t num_batches;
    const uint64_t* points;        /* num_batches x 2 */
    const uint32_t* batch_sizes;   /* num_batches */
    const uint32_t* poly_oracle;   /* sum(batch_sizes) */
    const uint32_t* poly_index;    /* sum(batch_sizes) */
} orc_fri_instance;

size_t orc_fri_proof_words(const orc_fri_params* p, unsigned log_n, unsigned num_oracles, const uint32_t* leaf_lens);

/* pow_witness_in == UINT64_MAX => search the minimum witness.  Returns 0 on success. */
int orc_fri_prove(const orc_batch* const* oracles, unsigned num_oracles, const orc_fri_instance* inst,
                  const orc_fri_params* p, orc_challenger* ch, uint64_t pow_witness_in, uint64_t* proof_out);

/* openings: for each batch, batch_sizes[b] x 2 words in instance order.  caps: per oracle 2^h x 4 words.
 * ch must be the challenger state right before prove_openings (openings already observed). 0 = accept. */
int orc_fri_verify(unsigned log_n, unsigned num_oracles, const uint32_t* leaf_lens, const uint64_t* caps,
                   const orc_fri_instance* inst, const uint64_t* openings, const orc_fri_params* p,
                   orc_challenger* ch, const uint64_t* proof);

int orc_num_threads(void);
void orc_set_num_threads(int n);   /* OpenMP threads of every later call (overrides OMP_NUM_THREADS) */
#ifdef __cplusplus
}
#endif
#endif
