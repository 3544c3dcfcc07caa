//! Raw bindings to `include/qpzk.h`.  One item per C entry point; see the header for the contract
//! (return 0 = OK, negative = error, message via `qpzk_last_error()`; no panic crosses the boundary).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)] pub struct qpzk_ctx { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_batch { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_challenger { _p: [u8; 0] }
#[repr(C)] pub struct qpzk_prover { _p: [u8; 0] }

pub const QPZK_POW_SEARCH: u64 = u64::MAX;
pub const QPZK_MAX_FRI_ROUNDS: usize = 16;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qpzk_challenger_state {
    pub sponge_state: [u64; 12],
    pub input_buffer: [u64; 8],
    pub output_buffer: [u64; 8],
    pub input_len: u32,
    pub output_len: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qpzk_fri_params {
    pub rate_bits: u32,
    pub cap_height: u32,
    pub proof_of_work_bits: u32,
    pub num_query_rounds: u32,
    pub num_rounds: u32,
    pub arity_bits: [u32; QPZK_MAX_FRI_ROUNDS],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qpzk_proof_shape {
    pub log_n: u32,
    pub num_constants: u32,
    pub num_sigmas: u32,
    pub num_wires: u32,
    pub num_zs: u32,
    pub num_partial_products: u32,
    pub num_quotient_polys: u32,
    pub leaf_lens: [u32; 4],
    pub num_public_inputs: u32,
}

/// Gate constraints as a straight-line register program (see include/qpzk.h).
#[repr(C)]
pub struct qpzk_gate_program {
    pub code: *const u32,
    pub num_ops: u32,
    pub imm: *const u64,
    pub num_imm: u32,
    pub num_regs: u32,
}

/// What `CircuitBuilder::build` leaves behind for `qpzk_prover_create` (see include/qpzk.h).
#[repr(C)]
pub struct qpzk_prover_config {
    pub num_wires: u32,
    pub num_routed: u32,
    pub sigma_col0: u32,
    pub quotient_degree_factor: u32,
    pub num_challenges: u32,
    pub num_public_inputs: u32,
    pub blinding: u32,
    pub use_graph: u32,
    pub circuit_digest: [u64; 4],
    pub fri: qpzk_fri_params,
    pub gates: *const qpzk_gate_program,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qpzk_prover_layout {
    pub cap_words: u64,
    pub cap_offset: [u64; 3],
    pub openings_offset: u64,
    pub num_openings: u64,
    pub fri_offset: u64,
    pub fri_words: u64,
    pub total_words: u64,
    pub leaf_lens: [u32; 4],
}

#[repr(C)]
pub struct qpzk_fri_instance {
    pub num_batches: u32,
    pub points: *const u64,
    pub batch_sizes: *const u32,
    pub poly_oracle: *const u32,
    pub poly_index: *const u32,
}

extern "C" {
    pub fn qpzk_last_error() -> *const c_char;
    pub fn qpzk_ctx_create(device: c_int, out: *mut *mut qpzk_ctx) -> c_int;
    pub fn qpzk_ctx_create_on_stream(device: c_int, stream: *mut c_void, out: *mut *mut qpzk_ctx) -> c_int;
    pub fn qpzk_ctx_destroy(ctx: *mut qpzk_ctx);
    pub fn qpzk_ctx_synchronize(ctx: *mut qpzk_ctx) -> c_int;
    pub fn qpzk_ctx_trim(ctx: *mut qpzk_ctx) -> c_int;
    pub fn qpzk_batches_eval_ext(ctx: *mut qpzk_ctx, batches: *const *const qpzk_batch, num_batches: u32, point: *const u64,
                                 out: *mut u64) -> c_int;
    pub fn qpzk_pinned_alloc(bytes: usize) -> *mut c_void;
    pub fn qpzk_pinned_free(p: *mut c_void);
    pub fn qpzk_poseidon_permute(ctx: *mut qpzk_ctx, states: *const u64, count: u64, out: *mut u64) -> c_int;
    pub fn qpzk_ctx_cached_bytes(ctx: *const qpzk_ctx) -> u64;
    pub fn qpzk_ctx_device_allocs(ctx: *const qpzk_ctx) -> u64;
    pub fn qpzk_ctx_peak_bytes(ctx: *const qpzk_ctx) -> u64;
    pub fn qpzk_ctx_reset_peak(ctx: *mut qpzk_ctx);
    pub fn qpzk_commit_from_values(ctx: *mut qpzk_ctx, cols: *const u64, log_n: u32, num_cols: u32, rate_bits: u32,
                                   cap_height: u32, salt: *const u64, out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_commit_from_coeffs(ctx: *mut qpzk_ctx, cols: *const u64, log_n: u32, num_cols: u32, rate_bits: u32,
                                   cap_height: u32, salt: *const u64, out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_merkle_new(ctx: *mut qpzk_ctx, leaves: *const u64, num_leaves: u64, leaf_len: u32, cap_height: u32,
                           out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_free(b: *mut qpzk_batch);
    pub fn qpzk_batch_cap_len(b: *const qpzk_batch) -> u32;
    pub fn qpzk_batch_cap(b: *const qpzk_batch, out: *mut u64) -> c_int;
    pub fn qpzk_batch_coeffs(b: *const qpzk_batch, out: *mut u64) -> c_int;
    pub fn qpzk_batch_rows(b: *const qpzk_batch, first_row: u64, num_rows: u64, out: *mut u64) -> c_int;
    pub fn qpzk_batch_open(b: *const qpzk_batch, idx: *const u64, k: u32, rows: *mut u64, siblings: *mut u64) -> c_int;
    pub fn qpzk_batch_eval_ext(b: *const qpzk_batch, point: *const u64, out: *mut u64) -> c_int;
    pub fn qpzk_challenger_create(ctx: *mut qpzk_ctx, out: *mut *mut qpzk_challenger) -> c_int;
    pub fn qpzk_challenger_free(ch: *mut qpzk_challenger);
    pub fn qpzk_challenger_import(ch: *mut qpzk_challenger, state: *const qpzk_challenger_state) -> c_int;
    pub fn qpzk_challenger_export(ch: *mut qpzk_challenger, state: *mut qpzk_challenger_state) -> c_int;
    pub fn qpzk_fri_proof_words(p: *const qpzk_fri_params, log_n: u32, num_oracles: u32, leaf_lens: *const u32) -> usize;
    pub fn qpzk_fri_prove(ctx: *mut qpzk_ctx, oracles: *const *const qpzk_batch, num_oracles: u32,
                          instance: *const qpzk_fri_instance, params: *const qpzk_fri_params,
                          ch: *mut qpzk_challenger, pow_witness_in: u64, proof_out: *mut u64,
                          openings_out: *mut u64) -> c_int;
    // ---- remaining entry points of include/qpzk.h (device-pointer variants, accessors, challenger, permutation
    // argument, microbenchmarks): every symbol the header declares is bound here (tests/test_capi_cpu.py checks it)
    pub fn qpzk_version() -> *const c_char;
    pub fn qpzk_ctx_set_zeroize(ctx: *mut qpzk_ctx, on: c_int) -> c_int;
    pub fn qpzk_ctx_set_keep_values(ctx: *mut qpzk_ctx, on: c_int) -> c_int;
    pub fn qpzk_ctx_set_timing(ctx: *mut qpzk_ctx, on: c_int) -> c_int;
    pub fn qpzk_ctx_last_commit_timing(ctx: *mut qpzk_ctx, ms: *mut f32) -> c_int;
    pub fn qpzk_ctx_launch_count(ctx: *const qpzk_ctx) -> u64;
    pub fn qpzk_commit_from_values_dev(ctx: *mut qpzk_ctx, cols_dev: *const u64, log_n: u32, num_cols: u32, rate_bits: u32,
                                       cap_height: u32, salt_dev: *const u64, out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_commit_from_coeffs_dev(ctx: *mut qpzk_ctx, cols_dev: *const u64, log_n: u32, num_cols: u32, rate_bits: u32,
                                       cap_height: u32, salt_dev: *const u64, out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_commit_cosets_from_coeffs_dev(ctx: *mut qpzk_ctx, coeffs_dev: *const u64, log_n: u32, num_cols: u32,
                                              rate_bits: u32, cap_height: u32, coset_begin: u32, coset_count: u32,
                                              out: *mut *mut qpzk_batch) -> c_int;
    // multi-GPU coset split with the exchange fused into the first LDE pass (peer slices read over NVLink)
    pub fn qpzk_commit_cosets_from_coeff_slices_dev(ctx: *mut qpzk_ctx, slices_dev: *const *const u64, num_slices: u32,
                                                    cols_per_slice: u32, log_n: u32, num_cols: u32, rate_bits: u32,
                                                    cap_height: u32, coset_begin: u32, coset_count: u32,
                                                    out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_peer_buffer_create(ctx: *mut qpzk_ctx, words: u64, dev_ptr: *mut *mut u64, handle: *mut u8) -> c_int;
    pub fn qpzk_peer_buffer_open(ctx: *mut qpzk_ctx, handle: *const u8, peer_ptr: *mut *mut u64) -> c_int;
    pub fn qpzk_peer_buffer_close(ctx: *mut qpzk_ctx, peer_ptr: *mut u64) -> c_int;
    pub fn qpzk_peer_buffer_free(ctx: *mut qpzk_ctx, dev_ptr: *mut u64) -> c_int;
    pub fn qpzk_intt_columns_dev(ctx: *mut qpzk_ctx, values_dev: *const u64, coeffs_dev: *mut u64, log_n: u32,
                                 num_cols: u32) -> c_int;
    // Poseidon2: the reference's application hash (addresses, nullifiers, zk-tree nodes, block hashes)
    pub fn qpzk_poseidon2_permute(ctx: *mut qpzk_ctx, states: *const u64, count: u64, out: *mut u64) -> c_int;
    pub fn qpzk_poseidon2_hash_no_pad(ctx: *mut qpzk_ctx, msgs: *const u64, count: u64, len: u32, out: *mut u64) -> c_int;
    pub fn qpzk_poseidon2_hash_no_pad_dev(ctx: *mut qpzk_ctx, msgs_dev: *const u64, count: u64, len: u32, out_dev: *mut u64) -> c_int;
    pub fn qpzk_selftest_mds_layer(ctx: *mut qpzk_ctx, states: *const u64, count: u64, next_round: u32, out: *mut u64) -> c_int;
    pub fn qpzk_merkle_new_dev(ctx: *mut qpzk_ctx, leaves_dev: *const u64, num_leaves: u64, leaf_len: u32, cap_height: u32,
                               out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_batch_num_leaves(b: *const qpzk_batch) -> u64;
    pub fn qpzk_batch_leaf_len(b: *const qpzk_batch) -> u32;
    pub fn qpzk_batch_digests(b: *const qpzk_batch, level: u32, out: *mut u64) -> c_int;
    pub fn qpzk_batch_leaves_dev(b: *const qpzk_batch) -> *const u64;
    pub fn qpzk_batch_coeffs_dev(b: *const qpzk_batch) -> *const u64;
    pub fn qpzk_batch_values_dev(b: *const qpzk_batch) -> *const u64;
    pub fn qpzk_challenger_observe(ch: *mut qpzk_challenger, elems: *const u64, n: u32) -> c_int;
    pub fn qpzk_challenger_observe_cap(ch: *mut qpzk_challenger, b: *const qpzk_batch) -> c_int;
    pub fn qpzk_challenger_get(ch: *mut qpzk_challenger, out: *mut u64, n: u32) -> c_int;
    pub fn qpzk_fri_params_constant_arity(p: *mut qpzk_fri_params, log_n: u32, rate_bits: u32, cap_height: u32,
                                          proof_of_work_bits: u32, num_query_rounds: u32, arity_bits: u32,
                                          final_poly_bits: u32) -> c_int;
    // all_wires_permutation_partial_products (plonk/prover.rs) and the zs/partial-products commitment
    pub fn qpzk_partial_products(ctx: *mut qpzk_ctx, wires: *const u64, sigmas: *const u64, log_n: u32, num_routed: u32,
                                 chunk: u32, num_challenges: u32, betas: *const u64, gammas: *const u64, out: *mut u64) -> c_int;
    pub fn qpzk_partial_products_dev(ctx: *mut qpzk_ctx, wires_dev: *const u64, sigmas_dev: *const u64, log_n: u32,
                                     num_routed: u32, chunk: u32, num_challenges: u32, betas: *const u64,
                                     gammas: *const u64, out_dev: *mut u64) -> c_int;
    pub fn qpzk_commit_zs_partial_products(ctx: *mut qpzk_ctx, constants_sigmas: *const qpzk_batch, wires: *const qpzk_batch,
                                           sigma_col0: u32, num_routed: u32, chunk: u32, num_challenges: u32,
                                           betas: *const u64, gammas: *const u64, cap_height: u32,
                                           salt_dev: *const u64, out: *mut *mut qpzk_batch) -> c_int;
    // compute_quotient_polys for the permutation argument (+ caller-supplied gate terms) and the quotient commitment
    pub fn qpzk_quotient_polys(ctx: *mut qpzk_ctx, constants_sigmas: *const qpzk_batch, wires: *const qpzk_batch,
                               zs_partial_products: *const qpzk_batch, sigma_col0: u32, num_routed: u32,
                               quotient_degree_factor: u32, num_challenges: u32, betas: *const u64, gammas: *const u64,
                               alphas: *const u64, gate_terms_dev: *const u64, num_gate_terms: u32, out: *mut u64) -> c_int;
    pub fn qpzk_commit_quotient(ctx: *mut qpzk_ctx, constants_sigmas: *const qpzk_batch, wires: *const qpzk_batch,
                                zs_partial_products: *const qpzk_batch, sigma_col0: u32, num_routed: u32,
                                quotient_degree_factor: u32, num_challenges: u32, betas: *const u64, gammas: *const u64,
                                alphas: *const u64, gate_terms_dev: *const u64, num_gate_terms: u32, cap_height: u32,
                                salt_dev: *const u64, out: *mut *mut qpzk_batch) -> c_int;
    pub fn qpzk_quotient_polys_prog(ctx: *mut qpzk_ctx, constants_sigmas: *const qpzk_batch, wires: *const qpzk_batch,
                                    zs_partial_products: *const qpzk_batch, sigma_col0: u32, num_routed: u32,
                                    quotient_degree_factor: u32, num_challenges: u32, betas: *const u64, gammas: *const u64,
                                    alphas: *const u64, prog: *const qpzk_gate_program, out: *mut u64) -> c_int;
    pub fn qpzk_commit_quotient_prog(ctx: *mut qpzk_ctx, constants_sigmas: *const qpzk_batch, wires: *const qpzk_batch,
                                     zs_partial_products: *const qpzk_batch, sigma_col0: u32, num_routed: u32,
                                     quotient_degree_factor: u32, num_challenges: u32, betas: *const u64, gammas: *const u64,
                                     alphas: *const u64, prog: *const qpzk_gate_program, cap_height: u32,
                                     salt_dev: *const u64, out: *mut *mut qpzk_batch) -> c_int;
    // prove(): the whole proof after witness generation in one device-resident call (CUDA graph replay)
    pub fn qpzk_prover_create(ctx: *mut qpzk_ctx, constants_sigmas: *const qpzk_batch, cfg: *const qpzk_prover_config,
                              out: *mut *mut qpzk_prover) -> c_int;
    pub fn qpzk_prover_free(p: *mut qpzk_prover);
    pub fn qpzk_prover_get_layout(p: *const qpzk_prover, out: *mut qpzk_prover_layout) -> c_int;
    pub fn qpzk_prover_prove_async(p: *mut qpzk_prover, wires: *const u64, public_inputs: *const u64, salts: *const u64,
                                   pow_witness_in: u64) -> c_int;
    pub fn qpzk_prover_finish(p: *mut qpzk_prover, out_words: *mut u64) -> c_int;
    pub fn qpzk_prover_prove(p: *mut qpzk_prover, wires: *const u64, public_inputs: *const u64, salts: *const u64,
                             pow_witness_in: u64, out_words: *mut u64) -> c_int;
    pub fn qpzk_prover_launches_per_proof(p: *const qpzk_prover) -> u64;
    pub fn qpzk_prover_uses_graph(p: *const qpzk_prover) -> c_int;
    pub fn qpzk_lde_pass_plan(log_n: u32, num_cols: u32, peer_slices: c_int, passes_out: *mut u32, num_passes: *mut u32, last_pass_bits: *mut u32) -> c_int;
    pub fn qpzk_selftest_field_ops(ctx: *mut qpzk_ctx, op: u32, a: *const u64, b: *const u64, c: *const u64, count: u64, out: *mut u64) -> c_int;
    pub fn qpzk_gate_jit_available() -> c_int;
    pub fn qpzk_gate_jit_set_mode(mode: c_int);
    pub fn qpzk_gate_jit_compile_check(prog: *const qpzk_gate_program, cubin_bytes: *mut u64) -> c_int;
    pub fn qpzk_gate_jit_compiled() -> u64;
    // ProofWithPublicInputs::to_bytes / from_bytes (host functions)
    pub fn qpzk_proof_bytes_len(p: *const qpzk_fri_params, s: *const qpzk_proof_shape) -> usize;
    pub fn qpzk_proof_to_bytes(p: *const qpzk_fri_params, s: *const qpzk_proof_shape, wires_cap: *const u64,
                               zs_partial_products_cap: *const u64, quotient_polys_cap: *const u64, openings: *const u64,
                               fri_proof: *const u64, public_inputs: *const u64, out: *mut u8, out_len: usize) -> c_int;
    pub fn qpzk_proof_from_bytes(p: *const qpzk_fri_params, s: *const qpzk_proof_shape, bytes: *const u8, len: usize,
                                 wires_cap: *mut u64, zs_partial_products_cap: *mut u64, quotient_polys_cap: *mut u64,
                                 openings: *mut u64, fri_proof: *mut u64, public_inputs: *mut u64) -> c_int;
    pub fn qpzk_bench_imad(ctx: *mut qpzk_ctx, iters: u32, imad_per_sec: *mut f64) -> c_int;
    pub fn qpzk_bench_poseidon(ctx: *mut qpzk_ctx, perms_per_thread: u32, perms_per_sec: *mut f64) -> c_int;
}
