/* qpzk.h -- C ABI of libqpzk.so, the B200 (sm_100a) backend for the batched polynomial commitment
 * under plonky2's `prove()` as used by Quantus-Network/qp-zk-circuits.
 *
 * This is the drop-in boundary.  The reference reaches the hot path only through
 * `ProverCircuitData::prove` / `CircuitData::prove`:
 *     wormhole/prover/src/lib.rs:171-175
 *     wormhole/aggregator/src/private_batch/prover/lib.rs:326-330
 *     wormhole/aggregator/src/public_batch/prover/lib.rs:301-305
 * and `CircuitBuilder::build*` (wormhole/circuit/src/circuit.rs:206-216), which call, inside the
 * un-vendored crate qp-plonky2 =1.5.5 (Cargo.toml:20), the functions each entry point below replaces.
 * INTEGRATION.md shows the Rust `-sys` binding a maintainer adds to a `[patch.crates-io]` fork.
 *
 * Conventions
 *  - every function returns 0 on success, a negative QPZK_E* code on failure; qpzk_last_error() gives a
 *    thread-local message.  No C++ exception or abort crosses this boundary.
 *  - field elements are canonical Goldilocks values (< p = 2^64 - 2^32 + 1), little-endian uint64_t;
 *    extension elements are two consecutive words [a0, a1] (a0 + a1*X, X^2 = 7).
 *    Outputs are always canonical; opening points, observed elements and PoW witnesses are checked at the boundary.
 *  - host pointers are owned by the caller and only read/written during the call; device memory is
 *    owned by the library behind opaque handles.  Functions suffixed _dev take DEVICE pointers.
 *  - one qpzk_ctx = one device + one CUDA stream; a context must not be used from two threads at
 *    once, distinct contexts are independent (multi-GPU = one context per device).
 *  - randomness (salts, PoW witness policy) is always caller-controlled, so runs are reproducible.
 *  - there is NO CPU fallback: if the device or the kernels are unavailable the call fails.
 */
#ifndef QPZK_H
#define QPZK_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define QPZK_OK 0
#define QPZK_EINVAL (-1)   /* bad argument (shape, cap_height > log2(leaves), null pointer, ...) */
#define QPZK_ECUDA (-2)    /* CUDA runtime error (message has the CUDA error string) */
#define QPZK_ENOMEM (-3)   /* device allocation failed */
#define QPZK_EPOW (-4)     /* caller-supplied PoW witness does not satisfy the difficulty */
#define QPZK_EINTERNAL (-5)

#define QPZK_SALT_SIZE 4          /* plonky2 SALT_SIZE: blinding columns appended to each leaf */
#define QPZK_MAX_FRI_ROUNDS 16
#define QPZK_MAX_ORACLES 8
#define QPZK_MAX_FRI_BATCHES 8
#define QPZK_POW_SEARCH UINT64_MAX /* pow_witness_in value meaning "grind the minimum witness" */

typedef struct qpzk_ctx qpzk_ctx;
typedef struct qpzk_batch qpzk_batch;           /* PolynomialBatch and/or MerkleTree, device resident */
typedef struct qpzk_challenger qpzk_challenger; /* Challenger<F, PoseidonHash>, device resident */

const char* qpzk_last_error(void);
const char* qpzk_version(void);

/* ---- context ---------------------------------------------------------------------------------- */
int qpzk_ctx_create(int device, qpzk_ctx** out);
/* Same, but enqueue on an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream). */
int qpzk_ctx_create_on_stream(int device, void* cuda_stream, qpzk_ctx** out);
void qpzk_ctx_destroy(qpzk_ctx* ctx);
int qpzk_ctx_synchronize(qpzk_ctx* ctx);
/* Zero wire/LDE buffers that held secret-derived data (secret hygiene, reference wormhole/circuit/src/sensitive.rs:1-16).
 *   1 (default)  when the buffer is released (qpzk_batch_free, end of a call)
 *   2            deferred: released buffers stay in the context's private block cache as they are and are zeroed before
 *                they are returned to the driver (qpzk_ctx_trim, qpzk_ctx_destroy, switching the mode back).  For several
 *                prover contexts on one device: the zero fills at release were measured to trigger stalls of 0.2-2 s in
 *                about half of the runs of 4 concurrent provers, none without them.
 *   0            never */
int qpzk_ctx_set_zeroize(qpzk_ctx* ctx, int mode);
/* Stage timing with CUDA events on the context stream.  After a commit:
 *   ms[0]=iNTT  ms[1]=LDE passes  ms[2]=leaf hashing  ms[3]=Merkle levels  (0 when a stage did not run) */
/* Batches created by from_values keep their input VALUES (on the subgroup) on the device as well -- what the
 * permutation argument needs from the wires and the sigmas (qpzk_commit_zs_partial_products).  Default: off. */
int qpzk_ctx_set_keep_values(qpzk_ctx* ctx, int on);
int qpzk_ctx_set_timing(qpzk_ctx* ctx, int on);
int qpzk_ctx_last_commit_timing(qpzk_ctx* ctx, float ms[4]);
/* number of kernel launches issued through this context since creation */
uint64_t qpzk_ctx_launch_count(const qpzk_ctx* ctx);
/* Device buffers released by qpzk_batch_free / at the end of a call stay in a per-context cache and are reused by
 * the next request of the same shape (a prover commits the same shapes over and over, reference
 * wormhole/prover/src/lib.rs:137-175), so a steady-state commit performs no device allocation.
 * qpzk_ctx_trim returns the cached blocks to the driver; qpzk_ctx_cached_bytes reports their total size. */
int qpzk_ctx_trim(qpzk_ctx* ctx);
/* Page-locked host memory (cudaHostAlloc, portable): with pinned inputs the H2D copy of a host-buffer commit runs
 * asynchronously on the context's copy stream, overlapped with the per-column stages; pageable memory works too but
 * the driver stages it and blocks the calling thread.  NULL on failure. */
void* qpzk_pinned_alloc(size_t bytes);
void qpzk_pinned_free(void* p);
uint64_t qpzk_ctx_cached_bytes(const qpzk_ctx* ctx);
/* number of cudaMalloc calls the context's block cache has made since creation (does not grow in steady state) */
uint64_t qpzk_ctx_device_allocs(const qpzk_ctx* ctx);
/* High-water mark of the device memory this context holds (buffers in use + cached blocks), in bytes -- the role of the
 * reference's peak-RSS figure (wormhole/memprof/src/report.rs:18-54; common/src/circuit.rs:393-395 quotes 1.5 / 2.5 GB
 * for the 2^15 / 2^16 private-batch prover).  qpzk_ctx_reset_peak restarts the measurement from the current holding.
 * Every stage of the path is also an NVTX range ("qpzk:from_values", "qpzk:ifft", "qpzk:lde passes",
 * "qpzk:MerkleTree::new", "qpzk:prove_openings", ...) for nsys / ncu --nvtx. */
uint64_t qpzk_ctx_peak_bytes(const qpzk_ctx* ctx);
void qpzk_ctx_reset_peak(qpzk_ctx* ctx);

/* ---- PolynomialBatch::from_values / from_coeffs --------------------------------------------------
 * Replaces qp-plonky2 `PolynomialBatch::from_values(values, rate_bits, blinding, cap_height, timing,
 * fft_root_table)` and `::from_coeffs(...)`: per column iFFT, zero-padded
 * coset FFT with shift g, transpose, reverse_index_bits, MerkleTree::new.
 *   cols: num_cols x 2^log_n words, column-major (one polynomial after another).
 *   salt: NULL (blinding = false) or QPZK_SALT_SIZE x 2^(log_n+rate_bits) words, column-major, the
 *         blinding values for LDE index i (natural order) -- what upstream draws from its RNG.
 * The batch keeps: coefficients, the leaf matrix (row j = LDE point bitrev(j), num_cols+salt words),
 * all digests, the cap. */
int qpzk_commit_from_values(qpzk_ctx* ctx, const uint64_t* cols, uint32_t log_n, uint32_t num_cols,
                            uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, qpzk_batch** out);
int qpzk_commit_from_coeffs(qpzk_ctx* ctx, const uint64_t* cols, uint32_t log_n, uint32_t num_cols,
                            uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, qpzk_batch** out);
int qpzk_commit_from_values_dev(qpzk_ctx* ctx, const uint64_t* cols_dev, uint32_t log_n, uint32_t num_cols,
                                uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt_dev,
                                qpzk_batch** out);
int qpzk_commit_from_coeffs_dev(qpzk_ctx* ctx, const uint64_t* cols_dev, uint32_t log_n, uint32_t num_cols,
                                uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt_dev,
                                qpzk_batch** out);
/* Coset-split commitment for one commitment spread over several GPUs: compute only
 * the leaf blocks [coset_begin, coset_begin+coset_count) (each block = 2^log_n consecutive leaves =
 * one Merkle subtree) from the full coefficient matrix.  coset_count must be a power of two dividing
 * 2^rate_bits and cap_height >= rate_bits - log2(coset_count).  The resulting batch's "cap" is the slice
 * [coset_begin * 2^cap_height / 2^rate_bits, ...) of the full cap; leaf indices are local to the slice. */
int qpzk_commit_cosets_from_coeffs_dev(qpzk_ctx* ctx, const uint64_t* coeffs_dev, uint32_t log_n,
                                       uint32_t num_cols, uint32_t rate_bits, uint32_t cap_height,
                                       uint32_t coset_begin, uint32_t coset_count, qpzk_batch** out);
/* The same with the coefficient matrix spread over `num_slices` device buffers of cols_per_slice columns each (slice k
 * holds columns [k * cols_per_slice, ...)): the per-rank iNTT outputs of a multi-GPU commitment, the other ranks' buffers
 * mapped into this process with qpzk_peer_buffer_open.  The first LDE pass reads every slice where it lies -- over
 * NVLink / NVSwitch for peer slices -- so the exchange step of SURVEY.md 8e is fused into the compute kernel: there is
 * no all-gather and no assembled copy of the matrix.  Needs log_n >= 12.  The caller orders the ranks (every rank's
 * iNTT must have finished before any rank starts, no rank may overwrite its slice before all have finished: two
 * stream-ordered barriers, e.g. a 1-element NCCL all-reduce and the cap gather). */
int qpzk_commit_cosets_from_coeff_slices_dev(qpzk_ctx* ctx, const uint64_t* const* slices_dev, uint32_t num_slices,
                                             uint32_t cols_per_slice, uint32_t log_n, uint32_t num_cols,
                                             uint32_t rate_bits, uint32_t cap_height, uint32_t coset_begin,
                                             uint32_t coset_count, qpzk_batch** out);
/* Exchange buffers of that path: a device buffer exported as a CUDA IPC handle (one process per GPU), and the peers'
 * buffers mapped through their handles.  The handle bytes travel over the caller's own plumbing (torch.distributed). */
#define QPZK_IPC_HANDLE_BYTES 64
int qpzk_peer_buffer_create(qpzk_ctx* ctx, uint64_t words, uint64_t** dev_ptr, uint8_t handle[QPZK_IPC_HANDLE_BYTES]);
int qpzk_peer_buffer_open(qpzk_ctx* ctx, const uint8_t handle[QPZK_IPC_HANDLE_BYTES], uint64_t** peer_ptr);
int qpzk_peer_buffer_close(qpzk_ctx* ctx, uint64_t* peer_ptr);
int qpzk_peer_buffer_free(qpzk_ctx* ctx, uint64_t* dev_ptr);
/* The pass plan of the coset LDE for a shape (a pure host function; a test hook and a description of what a commit of
 * this shape launches): passes_out[0..*num_passes) = position bits taken by the strided passes, top down (1..4: the
 * register-only kernel, 5..8: the tile kernel, more: the generic kernel), *last_pass_bits = bits of the pass that writes
 * whole leaf rows.  peer_slices != 0: the plan of qpzk_commit_cosets_from_coeff_slices_dev (wide pass first). */
int qpzk_lde_pass_plan(uint32_t log_n, uint32_t num_cols, int peer_slices, uint32_t* passes_out, uint32_t* num_passes,
                       uint32_t* last_pass_bits);
/* column iNTT only (values -> coefficients), used by the coset-split path before the all-gather */
int qpzk_intt_columns_dev(qpzk_ctx* ctx, const uint64_t* values_dev, uint64_t* coeffs_dev, uint32_t log_n,
                          uint32_t num_cols);

/* ---- PoseidonPermutation::permute on `count` states of 12 words (host buffers) ---------------------
 * Replaces qp-plonky2 `PoseidonPermutation::permute` / `Poseidon::poseidon`.  Inputs may be
 * any u64 representatives of field elements; outputs are canonical. */
int qpzk_poseidon_permute(qpzk_ctx* ctx, const uint64_t* states, uint64_t count, uint64_t* out);
/* Self-test hook (tests only): out = MDS * state + round constants of round `next_round` (0 and 30: none),
 * canonical.  Lets a test drive the carry paths of the limb-domain linear layer with chosen states. */
int qpzk_selftest_mds_layer(qpzk_ctx* ctx, const uint64_t* states, uint64_t count, uint32_t next_round, uint64_t* out);
/* Device field arithmetic, one operation per element (a test hook: directed operands reach carry paths that random data
 * does not).  op: 0 add, 1 sub (canonical operands) | any u64 operands, canonical result: 2 gl_mulc(a,b) (the NTT kernels'
 * product), 3 mulp(a,b) (the Poseidon kernels' product), 4 a^2, 5 a*b + c, 6 the C-level product, 7 reduce(lo=a, hi=b),
 * 8 / 9 the two PTX reductions of the 128-bit value (lo=a, hi=b) | 10 a + b with b canonical | 11 1/a | 12 a 160-bit dot
 * product: 30 a*b + b*c + c*a | 13 extension product (a,b)*(c,a): even elements get the first coordinate, odd the second |
 * 14 canonical a times 2^(6 (b mod 16)) by shifts (the in-register twiddles of the NTT kernels). */
int qpzk_selftest_field_ops(qpzk_ctx* ctx, uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                            uint64_t count, uint64_t* out);

/* ---- Poseidon2: the reference's APPLICATION hash (not the FRI / Merkle hasher) -------------------------------------
 * Replaces qp-plonky2 `hash::poseidon2::{Poseidon2Permutation, Poseidon2Hash::hash_no_pad}` / qp-poseidon-core =3.1.0
 * (reference Cargo.toml:20-21): addresses H(H("wormhole" || secret)) (wormhole/circuit/src/unspendable_account.rs:63-95),
 * nullifiers, zk-tree nodes (wormhole/circuit/src/zk_merkle_proof.rs:504,606), block hashes
 * (wormhole/circuit/src/block_header/header.rs:132-141) -- SURVEY.md 8f-2.  Width 12, x^7, 4 + 22 + 4 rounds; the hash
 * is an additive sponge of rate 8 over `x || 1 || 0*`, four words out.  PARITY PINNED by the seven vectors the
 * reference's tests hold (tools/gen_poseidon2_constants.py; tests/golden/poseidon2_kats.json).
 *   qpzk_poseidon2_permute      : `count` states of 12 words (any u64 representatives in, canonical out)
 *   qpzk_poseidon2_hash_no_pad  : `count` messages of `len` canonical words each, row-major -> count x 4 words */
int qpzk_poseidon2_permute(qpzk_ctx* ctx, const uint64_t* states, uint64_t count, uint64_t* out);
int qpzk_poseidon2_hash_no_pad(qpzk_ctx* ctx, const uint64_t* msgs, uint64_t count, uint32_t len, uint64_t* out);
int qpzk_poseidon2_hash_no_pad_dev(qpzk_ctx* ctx, const uint64_t* msgs_dev, uint64_t count, uint32_t len, uint64_t* out_dev);

/* ---- MerkleTree::new(leaves, cap_height) / .cap / .prove / .leaves -------------------------------
 * leaves: num_leaves x leaf_len words, row-major.  num_leaves must be a power of two, cap_height <= log2. */
int qpzk_merkle_new(qpzk_ctx* ctx, const uint64_t* leaves, uint64_t num_leaves, uint32_t leaf_len,
                    uint32_t cap_height, qpzk_batch** out);
int qpzk_merkle_new_dev(qpzk_ctx* ctx, const uint64_t* leaves_dev, uint64_t num_leaves, uint32_t leaf_len,
                        uint32_t cap_height, qpzk_batch** out);

void qpzk_batch_free(qpzk_batch* b);
uint64_t qpzk_batch_num_leaves(const qpzk_batch* b);
uint32_t qpzk_batch_leaf_len(const qpzk_batch* b);   /* num_cols + salt */
uint32_t qpzk_batch_cap_len(const qpzk_batch* b);    /* number of cap digests */
/* cap: cap_len x 4 words, order = subtree index */
int qpzk_batch_cap(const qpzk_batch* b, uint64_t* out);
/* polynomials (coefficients): num_cols x 2^log_n, column-major, natural order */
int qpzk_batch_coeffs(const qpzk_batch* b, uint64_t* out);
/* leaf rows [first_row, first_row+num_rows): num_rows x leaf_len words.
 * get_lde_values(i, step) of upstream is row bitrev(i*step) without the salt words. */
int qpzk_batch_rows(const qpzk_batch* b, uint64_t first_row, uint64_t num_rows, uint64_t* out);
/* digests of tree level `level` (0 = leaf digests): (num_leaves >> level) x 4 words */
int qpzk_batch_digests(const qpzk_batch* b, uint32_t level, uint64_t* out);
/* MerkleTree::prove for k leaf indices: rows = k x leaf_len, siblings = k x (log2(num_leaves)-cap_height) x 4 */
int qpzk_batch_open(const qpzk_batch* b, const uint64_t* idx, uint32_t k, uint64_t* rows, uint64_t* siblings);
/* evaluate every (unsalted) polynomial at an extension point: out = num_cols x 2 words.
 * (OpeningSet::new / eval_commitment) */
int qpzk_batch_eval_ext(const qpzk_batch* b, const uint64_t point[2], uint64_t* out);
/* The same for several batches of equal degree at one point, in ONE kernel launch and one device->host copy:
 * out = (sum of num_cols) x 2 words, batches in the given order.  (OpeningSet::new evaluates all four oracles at zeta.) */
int qpzk_batches_eval_ext(qpzk_ctx* ctx, const qpzk_batch* const* batches, uint32_t num_batches,
                          const uint64_t point[2], uint64_t* out);
/* device views for zero-copy consumers (e.g. a caller's own gate-term kernel): valid until qpzk_batch_free */
const uint64_t* qpzk_batch_leaves_dev(const qpzk_batch* b);
const uint64_t* qpzk_batch_coeffs_dev(const qpzk_batch* b);
const uint64_t* qpzk_batch_values_dev(const qpzk_batch* b);   /* NULL unless committed with keep_values on */

/* ---- Challenger<F, PoseidonHash> ----------------------------------------------------------------- */
typedef struct {
    uint64_t sponge_state[12];
    uint64_t input_buffer[8];
    uint64_t output_buffer[8];
    uint32_t input_len;
    uint32_t output_len;
} qpzk_challenger_state;

int qpzk_challenger_create(qpzk_ctx* ctx, qpzk_challenger** out);
void qpzk_challenger_free(qpzk_challenger* ch);
int qpzk_challenger_observe(qpzk_challenger* ch, const uint64_t* elems, uint32_t n); /* observe_elements */
int qpzk_challenger_observe_cap(qpzk_challenger* ch, const qpzk_batch* b);           /* device to device */
int qpzk_challenger_get(qpzk_challenger* ch, uint64_t* out, uint32_t n);             /* get_n_challenges */
int qpzk_challenger_export(qpzk_challenger* ch, qpzk_challenger_state* out);
int qpzk_challenger_import(qpzk_challenger* ch, const qpzk_challenger_state* in);

/* ---- FRI: PolynomialBatch::prove_openings --------------------------------------------------------- */
typedef struct {
    uint32_t rate_bits, cap_height, proof_of_work_bits, num_query_rounds;
    uint32_t num_rounds;                            /* reduction_arity_bits.len() */
    uint32_t arity_bits[QPZK_MAX_FRI_ROUNDS];
} qpzk_fri_params;

/* FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits)
 * (the reference uses (4, 5): wormhole/circuit/src/profile.rs:168-174) */
int qpzk_fri_params_constant_arity(qpzk_fri_params* p, uint32_t log_n, uint32_t rate_bits, uint32_t cap_height,
                                   uint32_t proof_of_work_bits, uint32_t num_query_rounds, uint32_t arity_bits,
                                   uint32_t final_poly_bits);

/* FriInstanceInfo: batches[b] opens polynomials (poly_oracle[k], poly_index[k]) at points[b]. */
typedef struct {
    uint32_t num_batches;
    const uint64_t* points;       /* num_batches x 2 */
    const uint32_t* batch_sizes;  /* num_batches */
    const uint32_t* poly_oracle;  /* sum(batch_sizes) */
    const uint32_t* poly_index;   /* sum(batch_sizes) */
} qpzk_fri_instance;

/* Flat FriProof layout (uint64 words), R = num_rounds, H = 2^cap_height, L = 2^(log_n+rate_bits):
 *   commit_phase_merkle_caps : R x H x 4
 *   query_round_proofs       : num_query_rounds x {
 *        for each oracle o : leaf row (leaf_len[o]) , siblings (log2 L - cap_height) x 4
 *        for each round r  : evals 2^arity_bits[r] x 2 , siblings (log2(L_r) - arity_bits[r] - cap_height) x 4 }
 *   final_poly               : 2^(log_n - sum arity_bits) x 2
 *   pow_witness              : 1
 * (uncompressed proof, field order of upstream FriProof::to_bytes) */
size_t qpzk_fri_proof_words(const qpzk_fri_params* p, uint32_t log_n, uint32_t num_oracles,
                            const uint32_t* leaf_lens);

/* prove_openings(instance, oracles, challenger, fri_params): draws alpha from the challenger, builds the
 * batch-opening polynomial, runs the commit phase, proof of work and query phase, leaves the challenger
 * in the same state upstream would.  openings_out (optional, may be NULL) receives the opened values in
 * instance order (sum(batch_sizes) x 2 words).  pow_witness_in: QPZK_POW_SEARCH or a witness to use. */
int qpzk_fri_prove(qpzk_ctx* ctx, const qpzk_batch* const* oracles, uint32_t num_oracles,
                   const qpzk_fri_instance* instance, const qpzk_fri_params* params, qpzk_challenger* ch,
                   uint64_t pow_witness_in, uint64_t* proof_out, uint64_t* openings_out);

/* ---- permutation argument: all_wires_permutation_partial_products ----------------------------------
 * Replaces qp-plonky2 `all_wires_permutation_partial_products` / `wires_permutation_partial_products_and_zs`
 * (plonk/prover.rs; SURVEY.md 8a row a16, the step between the wires commitment and the zs/partial-products
 * commitment of `prove()`, reference call site wormhole/prover/src/lib.rs:171-175).
 *   wires, sigmas : num_routed x 2^log_n words, column-major: the routed wire values and the sigma polynomials'
 *                   values on the subgroup (sigma_j(w^i) = k_j' * w^i' of the position wire (i, j) is copied to)
 *   chunk         : quotient_degree_factor (8 for every reference circuit: max_quotient_degree_factor,
 *                   common/src/circuit.rs:378-414); 1 <= ceil(num_routed / chunk) <= 32
 *   betas, gammas : num_challenges (<= 8) canonical challenges each (host pointers in both variants)
 *   out           : (num_challenges * ceil(num_routed / chunk)) x 2^log_n words, column-major, in the order of the
 *                   committed zs_partial_products batch: Z of every challenge, then the partial products of
 *                   challenge 0, of challenge 1, ...  (Z(1) = 1; the row for x holds Z(x) and the running products
 *                   Z(x) * chunk_0 ... chunk_k for k < ceil(num_routed / chunk) - 1)
 * The _dev variant takes device pointers for wires / sigmas / out (out may be passed straight to
 * qpzk_commit_from_values_dev). */
int qpzk_partial_products(qpzk_ctx* ctx, const uint64_t* wires, const uint64_t* sigmas, uint32_t log_n,
                          uint32_t num_routed, uint32_t chunk, uint32_t num_challenges, const uint64_t* betas,
                          const uint64_t* gammas, uint64_t* out);
int qpzk_partial_products_dev(qpzk_ctx* ctx, const uint64_t* wires_dev, const uint64_t* sigmas_dev, uint32_t log_n,
                              uint32_t num_routed, uint32_t chunk, uint32_t num_challenges, const uint64_t* betas,
                              const uint64_t* gammas, uint64_t* out_dev);

/* The same from two committed batches that kept their values (qpzk_ctx_set_keep_values), followed by the
 * zs/partial-products commitment (from_values, rate_bits of `wires`) -- nothing leaves the device:
 * sigma_j = column sigma_col0 + j of constants_sigmas, routed wires = columns [0, num_routed) of wires.
 * salt_dev: NULL, or the DEVICE blinding columns of this oracle (QPZK_SALT_SIZE x 2^(log_n+rate_bits), as for
 * qpzk_commit_from_values_dev): upstream blinds WIRES, ZS_PARTIAL_PRODUCTS and QUOTIENT when the circuit is
 * zero-knowledge (PlonkOracle::*.blinding; the leaf and private-batch configurations, common/src/circuit.rs:378-402). */
int qpzk_commit_zs_partial_products(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                                    uint32_t sigma_col0, uint32_t num_routed, uint32_t chunk, uint32_t num_challenges,
                                    const uint64_t* betas, const uint64_t* gammas, uint32_t cap_height,
                                    const uint64_t* salt_dev, qpzk_batch** out);

/* ---- quotient polynomials: compute_quotient_polys for the permutation argument -----------------------
 * Replaces the gate-independent part of qp-plonky2 `compute_quotient_polys` (plonk/prover.rs) and
 * `eval_vanishing_poly_base_batch` (plonk/vanishing_poly.rs) -- SURVEY.md 8a row a17, the step between the
 * zs/partial-products commitment and the quotient commitment of `prove()` (wormhole/prover/src/lib.rs:171-175).
 * At every point x of the coset g<w_(n Q)> (Q = quotient_degree_factor, a power of two <= 2^rate_bits):
 *     terms = [ L_0(x)(Z_c(x) - 1) for c ] ++ [ partial-product checks, challenge-major ] ++ [ gate terms ]
 *     q_a(x) = (sum_t terms[t] alpha_a^t) / Z_H(x)
 * followed by coset_ifft and the split into Q chunks of n coefficients per challenge.
 *   constants_sigmas     : the circuit's constants/sigmas batch; sigma_j is column sigma_col0 + j
 *   wires                : wires batch; the routed wires are columns [0, num_routed)
 *   zs_partial_products  : the batch committed from qpzk_partial_products' output (same column order)
 *   betas, gammas, alphas: num_challenges (<= 4) canonical challenges each (host)
 *   gate_terms_dev       : NULL, or a DEVICE matrix [n*Q][num_gate_terms], row = leaf index (row j is the point with
 *                          natural index bitrev(j)): the values upstream's evaluate_gate_constraints would produce.
 *                          The gate set itself lives in the un-vendored crate and is not part of this library.
 *   out                  : (num_challenges * Q) x 2^log_n coefficients, column-major: chunk k of challenge a is
 *                          column a*Q + k -- the input of the quotient commitment (from_coeffs).
 * qpzk_commit_quotient does the same and commits the chunks without leaving the device (rate_bits of the inputs;
 * salt_dev as for qpzk_commit_zs_partial_products). */
int qpzk_quotient_polys(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                        const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                        uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                        const uint64_t* gammas, const uint64_t* alphas, const uint64_t* gate_terms_dev,
                        uint32_t num_gate_terms, uint64_t* out);
int qpzk_commit_quotient(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                         const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                         uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                         const uint64_t* gammas, const uint64_t* alphas, const uint64_t* gate_terms_dev,
                         uint32_t num_gate_terms, uint32_t cap_height, const uint64_t* salt_dev, qpzk_batch** out);

/* Gate constraints handed over as DATA: a straight-line register program over the base field that computes, for one
 * point, what upstream's `evaluate_gate_constraints` / `eval_unfiltered_base_batch` (plonk/vanishing_poly.rs, gates/*.rs)
 * computes -- selector filters included.  The fork obtains it by running its gates' constraint code once over a
 * recording field type; the gate set itself stays in its crate.  The library interprets the program on the device at
 * every point of the quotient domain and feeds the emitted terms to the alpha combination after the permutation terms.
 *   code : num_ops x 4 words {op, dst, a, b}
 *          0 LOAD_WIRE  r[dst] = wires column a          3 ADD r[dst] = r[a] + r[b]     6 EMIT constraint term = r[a]
 *          1 LOAD_CONST r[dst] = constants_sigmas col a  4 SUB r[dst] = r[a] - r[b]
 *          2 LOAD_IMM   r[dst] = imm[a]                  5 MUL r[dst] = r[a] * r[b]
 *          fused forms, third source register c in bits 8..15 of the op word:
 *          7 MULADD r[dst] = r[a]*r[b] + r[c]   8 MULI r[dst] = r[a]*imm[b]   9 ADDI r[dst] = r[a] + imm[b]
 *          10 MULIADD r[dst] = r[a]*imm[b] + r[c]
 *   registers 0..num_regs-1 (num_regs <= 256; up to 64 stay in L1); a register must be written before it is read (checked). */
typedef struct {
    const uint32_t* code;
    uint32_t num_ops;
    const uint64_t* imm;   /* canonical */
    uint32_t num_imm;
    uint32_t num_regs;
} qpzk_gate_program;
/* Programs of 512 steps or more are compiled at run time (NVRTC, sm_100a) into straight-line code with the program's
 * registers in hardware registers and cached per process; shorter ones, or all of them with QPZK_GATE_JIT=0 or when NVRTC
 * cannot be loaded, are interpreted on the device.  Both paths give the same bits.
 * qpzk_gate_jit_available: 1 if NVRTC was found.  qpzk_gate_jit_compiled: programs compiled so far in this process. */
int qpzk_gate_jit_available(void);
void qpzk_gate_jit_set_mode(int mode);   /* 0 interpret, 1 compile (default, or $QPZK_GATE_JIT); process wide, applies to later uploads */
/* compiles `prog` for sm_100a without loading it (no GPU needed; a build check of the run-time compiler path) */
int qpzk_gate_jit_compile_check(const qpzk_gate_program* prog, uint64_t* cubin_bytes);
uint64_t qpzk_gate_jit_compiled(void);
/* qpzk_quotient_polys / qpzk_commit_quotient with the gate terms computed from `prog` (NULL = no gates). */
int qpzk_quotient_polys_prog(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                             const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                             uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                             const uint64_t* gammas, const uint64_t* alphas, const qpzk_gate_program* prog, uint64_t* out);
int qpzk_commit_quotient_prog(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                              const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                              uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                              const uint64_t* gammas, const uint64_t* alphas, const qpzk_gate_program* prog,
                              uint32_t cap_height, const uint64_t* salt_dev, qpzk_batch** out);

/* ---- prove(): the whole proof after witness generation in one call, device resident -------------------------------
 * Replaces the body of qp-plonky2 `prove_with_partition_witness` (plonk/prover.rs) from the wires commitment to the
 * assembled Proof (transcript Appendix A.7); what the reference's three provers
 * call: wormhole/prover/src/lib.rs:171-175, wormhole/aggregator/src/private_batch/prover/lib.rs:326-330,
 * wormhole/aggregator/src/public_batch/prover/lib.rs:301-305.
 * A qpzk_prover is created once per circuit (per context) from what CircuitBuilder::build leaves behind: the committed
 * constants/sigmas batch (committed with from_values under qpzk_ctx_set_keep_values(ctx, 1)), the numbers of wires and
 * routed wires, the gate constraints as a qpzk_gate_program, the verifier-data digest and the FRI parameters.  Per proof
 * the caller hands over the witness (wire values), the public inputs and -- for a zero-knowledge circuit -- the salts.
 * The Fiat-Shamir challenges never visit the host: challenger, partial products, quotient, openings and FRI read them
 * from device memory; after two proofs the whole sequence is replayed as ONE CUDA graph (use_graph), so a proof is one
 * host->device copy, one graph launch, one device->host copy and one wait.  Bit-identical to the step-by-step calls.
 * Result layout (uint64 words, qpzk_prover_get_layout): wires cap | zs_partial_products cap | quotient cap |
 * openings (FRI-instance order: constants, sigmas, wires, zs, partial products, quotient chunks, zs at g*zeta; 2 words
 * each) | flat FriProof of qpzk_fri_prove for the four oracles -- the pieces qpzk_proof_to_bytes takes. */
typedef struct qpzk_prover qpzk_prover;
typedef struct {
    uint32_t num_wires, num_routed, sigma_col0;   /* routed wires are columns [0, num_routed); sigma_j = constants_sigmas column sigma_col0 + j */
    uint32_t quotient_degree_factor;              /* 8 for every reference circuit (common/src/circuit.rs:378-414) */
    uint32_t num_challenges;                      /* 2 */
    uint32_t num_public_inputs;
    uint32_t blinding;                            /* 1: zero-knowledge circuit, the wires / zs / quotient oracles are salted */
    uint32_t use_graph;                           /* 1: replay the proof as a CUDA graph from the third proof on */
    uint64_t circuit_digest[4];                   /* verifier_only.circuit_digest, observed first */
    qpzk_fri_params fri;
    const qpzk_gate_program* gates;               /* NULL: no gate constraints (copied to the device by qpzk_prover_create) */
} qpzk_prover_config;
typedef struct {
    uint64_t cap_words;            /* 4 << cap_height */
    uint64_t cap_offset[3];        /* wires, zs_partial_products, quotient */
    uint64_t openings_offset, num_openings;
    uint64_t fri_offset, fri_words;
    uint64_t total_words;
    uint32_t leaf_lens[4];         /* leaf lengths of the four oracles (incl. salt): the argument of qpzk_fri_proof_words */
} qpzk_prover_layout;
int qpzk_prover_create(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_prover_config* cfg, qpzk_prover** out);
void qpzk_prover_free(qpzk_prover* p);
int qpzk_prover_get_layout(const qpzk_prover* p, qpzk_prover_layout* out);
/* wires: num_wires x 2^log_n values, column-major (host; pinned memory makes the copy asynchronous); public_inputs:
 * num_public_inputs canonical words; salts: NULL or 3 x QPZK_SALT_SIZE x 2^(log_n+rate_bits) words (wires, zs, quotient
 * oracle; caller RNG); pow_witness_in: QPZK_POW_SEARCH (minimum witness) or a witness to use (runs without the graph).
 * qpzk_prover_prove_async returns when the caller's buffers have been read -- they may be zeroized at once -- and the
 * proof is still running; qpzk_prover_finish waits for it and copies the layout.total_words result words to out_words.
 * Several provers on several contexts may be in flight at once from one host thread (their graphs overlap on the GPU). */
int qpzk_prover_prove_async(qpzk_prover* p, const uint64_t* wires, const uint64_t* public_inputs, const uint64_t* salts,
                            uint64_t pow_witness_in);
int qpzk_prover_finish(qpzk_prover* p, uint64_t* out_words);
int qpzk_prover_prove(qpzk_prover* p, const uint64_t* wires, const uint64_t* public_inputs, const uint64_t* salts,
                      uint64_t pow_witness_in, uint64_t* out_words);
uint64_t qpzk_prover_launches_per_proof(const qpzk_prover* p);   /* kernels + copies enqueued (or replayed) per proof */
int qpzk_prover_uses_graph(const qpzk_prover* p);                 /* 1 once the graph has been captured */

/* ---- wire format: ProofWithPublicInputs::to_bytes / from_bytes ----------------------------------------
 * Replaces qp-plonky2 `util/serialization`: `write_proof_with_public_inputs` / `read_proof_with_public_inputs`
 * (what the reference passes between prover, aggregator and verifier:
 * wormhole/tests/src/prover/prover_tests.rs:118, field inventory wormhole/aggregator/src/common/utils.rs:349-537).
 * Host functions, no device work.  Byte layout in qp-zk-circuits_b200/csrc/wire.cu (upstream's order as recalled;
 * PARITY UNPINNED).
 *   openings : the (k, 2) array in FRI-instance order, as qpzk_fri_prove / the prover produce it:
 *              constants | sigmas | wires | zs | partial products | quotient chunks | zs at g*zeta
 *   fri_proof: the flat words of qpzk_fri_prove for the four oracles with leaf lengths leaf_lens
 *              (constants_sigmas, wires, zs_partial_products, quotient; + QPZK_SALT_SIZE where blinded). */
typedef struct {
    uint32_t log_n;
    uint32_t num_constants, num_sigmas, num_wires, num_zs, num_partial_products, num_quotient_polys;
    uint32_t leaf_lens[4];
    uint32_t num_public_inputs;
} qpzk_proof_shape;
size_t qpzk_proof_bytes_len(const qpzk_fri_params* p, const qpzk_proof_shape* s);   /* 0 = invalid shape */
int qpzk_proof_to_bytes(const qpzk_fri_params* p, const qpzk_proof_shape* s, const uint64_t* wires_cap,
                        const uint64_t* zs_partial_products_cap, const uint64_t* quotient_polys_cap,
                        const uint64_t* openings, const uint64_t* fri_proof, const uint64_t* public_inputs,
                        uint8_t* out, size_t out_len);
/* The inverse; fails (QPZK_EINVAL) on a wrong total length, a sibling count that does not match the circuit, or a
 * non-canonical field element. */
int qpzk_proof_from_bytes(const qpzk_fri_params* p, const qpzk_proof_shape* s, const uint8_t* bytes, size_t len,
                          uint64_t* wires_cap, uint64_t* zs_partial_products_cap, uint64_t* quotient_polys_cap,
                          uint64_t* openings, uint64_t* fri_proof, uint64_t* public_inputs);

/* ---- microbenchmarks used by bench.py for the roofline denominators ------------------------------- */
/* IMAD.WIDE throughput: returns 32x32+64 multiply-adds per second measured over `iters` dependent-chain
 * iterations on every SM. */
int qpzk_bench_imad(qpzk_ctx* ctx, uint32_t iters, double* imad_per_sec);
/* Poseidon permutations per second (state in registers, no memory traffic) */
int qpzk_bench_poseidon(qpzk_ctx* ctx, uint32_t perms_per_thread, double* perms_per_sec);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
