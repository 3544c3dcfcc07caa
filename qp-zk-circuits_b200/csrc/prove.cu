// qpzk_prover: plonky2's `prove()` after witness generation as ONE call whose work never leaves the device.
//
// Replaces the body of qp-plonky2 =1.5.5 `prove_with_partition_witness` (plonk/prover.rs) from the wires commitment to
// the assembled `Proof` -- SURVEY.md 3.1 steps 2-10, row a18 -- reached in the reference from
// wormhole/prover/src/lib.rs:171-175, wormhole/aggregator/src/private_batch/prover/lib.rs:326-330 and
// wormhole/aggregator/src/public_batch/prover/lib.rs:301-305.  Transcript: SURVEY.md A.7.
//
// The step-by-step entry points (qpzk_commit_*, qpzk_challenger_*, qpzk_fri_prove) bring every challenge to the host
// and take it back as an argument: about eight host synchronisations and sixty-odd launches per proof, which is most
// of a 2^13-row proof's latency.  Here the Fiat-Shamir challenger lives on the device and every consumer reads its
// challenges from device memory (partial products, quotient, opening points, FRI), the openings are observed device
// to device, the PoW search is one launch, and the three caps, the opened values and the FRI proof are written into
// one buffer that comes back with a single copy.  Nothing in the sequence depends on host-visible data, so after two
// eager proofs (tables, block cache) the whole sequence is captured into a CUDA graph and replayed: one copy in, one
// graph launch, one copy out, one wait.  Device buffers come from a block cache private to the prover, so the
// addresses baked into the graph stay the prover's for its lifetime.
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.cuh"
#include "goldilocks.cuh"

struct qpzk_prover {
    qpzk_ctx* ctx = nullptr;
    const qpzk_batch* cs = nullptr;     // constants/sigmas oracle of the circuit (borrowed; values kept on the device)
    qpzk_prover_config cfg{};
    uint32_t log_n = 0, rate_bits = 0, nchunks = 0, zs_cols = 0, quot_cols = 0, total_openings = 0;
    bool blinding = false;
    GateProgramDev gates;
    std::vector<uint32_t> refs;         // poly_oracle[total] then poly_index[total] (host copy for validation)
    uint32_t batch_sizes[2] = {0, 0};
    qpzk_prover_layout lay{};
    // device
    uint64_t *d_refs = nullptr, *d_wires = nullptr, *d_pi = nullptr, *d_salts = nullptr, *d_chal = nullptr,
             *d_out = nullptr, *d_misc = nullptr, *d_ch = nullptr;
    uint64_t* h_out = nullptr;          // pinned staging of the result
    // graph
    cudaGraphExec_t exec = nullptr;
    int eager_runs = 0;
    bool in_flight = false;
    uint64_t launches_per_proof = 0;
    cudaEvent_t done = nullptr;
    std::multimap<uint64_t, void*> cache;   // private block cache (swapped with the context's while the prover enqueues)
    uint64_t cached_bytes = 0;
};

namespace {

struct CacheSwap {   // the prover's blocks come from and return to its own cache
    qpzk_prover* p;
    explicit CacheSwap(qpzk_prover* pr) : p(pr) { std::swap(p->ctx->cache, p->cache); std::swap(p->ctx->cached_bytes, p->cached_bytes); }
    ~CacheSwap() { std::swap(p->ctx->cache, p->cache); std::swap(p->ctx->cached_bytes, p->cached_bytes); }
};

// the whole proof on the context stream; no host synchronisation, no pageable host memory touched
int enqueue_proof(qpzk_prover* pr, uint64_t pow_witness_in) {
    qpzk_ctx* ctx = pr->ctx;
    const qpzk_prover_config& c = pr->cfg;
    cudaStream_t st = ctx->stream;
    ChallengerDev* ch = reinterpret_cast<ChallengerDev*>(pr->d_ch);
    const uint32_t nc = c.num_challenges, H4 = 4u << c.fri.cap_height, cosets = 1u << pr->rate_bits;
    const uint64_t L = (1ull << pr->log_n) << pr->rate_bits;
    const uint64_t* salt[3] = {nullptr, nullptr, nullptr};
    if (pr->blinding)
        for (int k = 0; k < 3; k++) salt[k] = pr->d_salts + (uint64_t)k * QPZK_SALT_SIZE * L;
    uint64_t* d_pih = pr->d_misc;                    // 4 words
    uint64_t* d_indices = pr->d_misc + 8;            // num_query_rounds words
    int rc = 0;
    qpzk_batch *w_b = nullptr, *z_b = nullptr, *q_b = nullptr;
    uint64_t *terms = nullptr, *d_q = nullptr;
    uint32_t qbits = 0;
    const bool timing = ctx->timing, keep = ctx->keep_values;
    ctx->timing = false;
    auto cap_of = [](const qpzk_batch* b) { return b->digests + digest_level_off(b->num_leaves, b->num_levels - 1); };
    do {
        NvtxRange nvtx("qpzk:prove");
        cudaError_t e = launch_fill_words(ch, sizeof(ChallengerDev) / 8, 0, st);
        // observe circuit_digest, public_inputs_hash
        if (e == cudaSuccess) e = launch_challenger_observe(ch, pr->d_chal + 3 * nc + 4, 4, st);
        if (e == cudaSuccess) e = launch_hash_no_pad(pr->d_pi, c.num_public_inputs, d_pih, st);
        if (e == cudaSuccess) e = launch_challenger_observe(ch, d_pih, 4, st);
        ctx->launches += 4;
        if (e != cudaSuccess) { rc = fail(QPZK_ECUDA, std::string("prove: ") + cudaGetErrorString(e)); break; }
        // wires commitment -> betas, gammas
        ctx->keep_values = true;
        rc = commit_impl(ctx, dev_source(pr->d_wires), true, pr->log_n, c.num_wires, pr->rate_bits, c.fri.cap_height, salt[0], 0,
                         cosets, true, &w_b);
        ctx->keep_values = keep;
        if (rc) break;
        e = launch_challenger_observe(ch, cap_of(w_b), H4, st);
        if (e == cudaSuccess) e = launch_challenger_get(ch, pr->d_chal, 2 * nc, st);
        ctx->launches += 2;
        if (e != cudaSuccess) { rc = fail(QPZK_ECUDA, std::string("prove: ") + cudaGetErrorString(e)); break; }
        // partial products + Z -> commitment -> alphas
        if ((rc = commit_zs_impl(ctx, pr->cs, w_b, c.sigma_col0, c.num_routed, c.quotient_degree_factor, nc, nullptr, nullptr,
                                 pr->d_chal, c.fri.cap_height, salt[1], &z_b))) break;
        e = launch_challenger_observe(ch, cap_of(z_b), H4, st);
        if (e == cudaSuccess) e = launch_challenger_get(ch, pr->d_chal + 2 * nc, nc, st);
        ctx->launches += 2;
        if (e != cudaSuccess) { rc = fail(QPZK_ECUDA, std::string("prove: ") + cudaGetErrorString(e)); break; }
        // quotient (gate program + permutation argument) -> commitment -> zeta, g * zeta
        if ((rc = gate_terms_run(ctx, pr->cs, w_b, c.quotient_degree_factor, pr->gates, &terms))) break;
        if ((rc = quotient_impl(ctx, pr->cs, w_b, z_b, c.sigma_col0, c.num_routed, c.quotient_degree_factor, nc, nullptr, nullptr,
                                nullptr, pr->d_chal, terms, pr->gates.num_terms, &d_q, &qbits))) break;
        if ((rc = commit_impl(ctx, dev_source(d_q), false, pr->log_n, pr->quot_cols, pr->rate_bits, c.fri.cap_height, salt[2], 0,
                              cosets, true, &q_b))) break;
        e = launch_challenger_observe(ch, cap_of(q_b), H4, st);
        if (e == cudaSuccess) e = launch_challenger_get(ch, pr->d_chal + 3 * nc, 2, st);
        if (e == cudaSuccess) e = launch_ext_scale(pr->d_chal + 3 * nc, gl_root_of_unity_host(pr->log_n), pr->d_chal + 3 * nc + 2, st);
        ctx->launches += 3;
        // the three caps of the proof
        const qpzk_batch* committed[3] = {w_b, z_b, q_b};
        for (int k = 0; k < 3 && e == cudaSuccess; k++)
            e = cudaMemcpyAsync(pr->d_out + pr->lay.cap_offset[k], cap_of(committed[k]), H4 * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { rc = fail(QPZK_ECUDA, std::string("prove: ") + cudaGetErrorString(e)); break; }
        // openings at zeta and g * zeta, observed on the device; prove_openings
        const qpzk_batch* oracles[4] = {pr->cs, w_b, z_b, q_b};
        qpzk_fri_instance inst{2, nullptr, pr->batch_sizes, pr->refs.data(), pr->refs.data() + pr->total_openings};
        FriDeviceIO io;
        io.points = pr->d_chal + 3 * nc;
        io.poly_refs = reinterpret_cast<const uint32_t*>(pr->d_refs);
        io.observe_openings = true;
        io.proof = pr->d_out + pr->lay.fri_offset;
        io.openings = pr->d_out + pr->lay.openings_offset;
        io.indices = d_indices;
        io.status = reinterpret_cast<uint32_t*>(pr->d_out + pr->lay.total_words);
        if ((rc = fri_core(ctx, oracles, 4, &inst, &c.fri, ch, pow_witness_in, io))) break;
    } while (false);
    const uint64_t Lq = (1ull << pr->log_n) * c.quotient_degree_factor;
    if (terms) dev_free(ctx, terms, Lq * pr->gates.num_terms);
    if (d_q) dev_free(ctx, d_q, (uint64_t)pr->quot_cols << pr->log_n);
    if (q_b) batch_release(q_b);
    if (z_b) batch_release(z_b);
    if (w_b) batch_release(w_b);
    // the witness itself (secret hygiene, reference wormhole/circuit/src/sensitive.rs:1-16)
    if (ctx->zeroize == 1) { launch_fill_words(pr->d_wires, (uint64_t)c.num_wires << pr->log_n, 0, st); ctx->launches++; }
    ctx->timing = timing;
    return rc;
}

}  // namespace

extern "C" int qpzk_prover_create(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_prover_config* cfg,
                                  qpzk_prover** out) {
    GUARD_BEGIN
    if (!out) return fail(QPZK_EINVAL, "null argument");
    *out = nullptr;
    if (!ctx || !constants_sigmas || !cfg) return fail(QPZK_EINVAL, "null argument");
    const qpzk_batch* cs = constants_sigmas;
    if (cs->ctx != ctx) return fail(QPZK_EINVAL, "constants_sigmas belongs to another context");
    if (!cs->values || !cs->coeffs)
        return fail(QPZK_EINVAL, "constants_sigmas holds no values: commit it with from_values after qpzk_ctx_set_keep_values(ctx, 1)");
    const uint32_t log_n = cs->log_n, rate_bits = cs->rate_bits;
    if (cs->num_leaves != ((1ull << log_n) << rate_bits)) return fail(QPZK_EINVAL, "constants_sigmas is a coset slice");
    if (cfg->fri.rate_bits != rate_bits || cfg->fri.cap_height != cs->cap_height)
        return fail(QPZK_EINVAL, "fri parameters do not match the constants_sigmas commitment");
    if (!cfg->num_wires || cfg->num_wires > 65535) return fail(QPZK_EINVAL, "num_wires must be in 1..65535");
    if (!cfg->num_routed || cfg->num_routed > cfg->num_wires || (uint64_t)cfg->sigma_col0 + cfg->num_routed > cs->num_cols)
        return fail(QPZK_EINVAL, "routed wires / sigma columns out of range");
    const uint32_t qdf = cfg->quotient_degree_factor, nc = cfg->num_challenges;
    if (!qdf || (qdf & (qdf - 1)) || ilog2(qdf) > rate_bits) return fail(QPZK_EINVAL, "quotient_degree_factor must be a power of two <= 2^rate_bits");
    if (!nc || nc > QUOT_MAX_CHALLENGES) return fail(QPZK_EINVAL, "num_challenges must be in 1..4");
    const uint32_t nchunks = (cfg->num_routed + qdf - 1) / qdf;
    if (nchunks > PP_MAX_CHUNKS) return fail(QPZK_EINVAL, "more than 32 partial-product chunks per row");
    if (cfg->fri.num_query_rounds > 4096) return fail(QPZK_EINVAL, "too many query rounds");
    for (int k = 0; k < 4; k++) if (cfg->circuit_digest[k] >= GL_P) return fail(QPZK_EINVAL, "circuit_digest is not canonical");
    CK(cudaSetDevice(ctx->device));

    qpzk_prover* pr = new qpzk_prover();
    ctx->refs++;
    pr->ctx = ctx; pr->cs = cs; pr->cfg = *cfg; pr->cfg.gates = nullptr;
    pr->log_n = log_n; pr->rate_bits = rate_bits; pr->nchunks = nchunks; pr->blinding = cfg->blinding != 0;
    pr->zs_cols = nc * nchunks; pr->quot_cols = nc * qdf;
    const uint32_t salt_w = pr->blinding ? QPZK_SALT_SIZE : 0;
    const uint32_t cols[4] = {cs->num_cols, cfg->num_wires, pr->zs_cols, pr->quot_cols};
    const uint32_t leaf_lens[4] = {cs->leaf_width, cfg->num_wires + salt_w, pr->zs_cols + salt_w, pr->quot_cols + salt_w};
    pr->batch_sizes[0] = cols[0] + cols[1] + cols[2] + cols[3];
    pr->batch_sizes[1] = nc;
    pr->total_openings = pr->batch_sizes[0] + pr->batch_sizes[1];
    std::vector<uint32_t> po, pi;
    for (uint32_t o = 0; o < 4; o++) for (uint32_t i = 0; i < cols[o]; i++) { po.push_back(o); pi.push_back(i); }
    for (uint32_t i = 0; i < nc; i++) { po.push_back(2); pi.push_back(i); }   // Z polynomials at g * zeta
    pr->refs = po;
    pr->refs.insert(pr->refs.end(), pi.begin(), pi.end());
    const uint32_t H4 = 4u << cfg->fri.cap_height;
    qpzk_prover_layout& lay = pr->lay;
    lay.cap_words = H4;
    lay.cap_offset[0] = 0; lay.cap_offset[1] = H4; lay.cap_offset[2] = 2 * H4;
    lay.openings_offset = 3 * H4; lay.num_openings = pr->total_openings;
    lay.fri_offset = lay.openings_offset + 2ull * pr->total_openings;
    lay.fri_words = qpzk_fri_proof_words(&cfg->fri, log_n, 4, leaf_lens);
    lay.total_words = lay.fri_offset + lay.fri_words;
    for (int k = 0; k < 4; k++) lay.leaf_lens[k] = leaf_lens[k];

    const uint64_t n = 1ull << log_n, L = n << rate_bits;
    int rc = 0;
    auto fail_out = [&](int code) { qpzk_prover_free(pr); return code; };
    {
        CacheSwap swap(pr);
        if ((rc = gate_program_upload(ctx, cs->num_cols, cfg->num_wires, cfg->gates, &pr->gates))) {}
        else if ((rc = dev_alloc(ctx, &pr->d_refs, pr->total_openings + 1))) {}
        else if ((rc = dev_alloc(ctx, &pr->d_wires, (uint64_t)cfg->num_wires * n))) {}
        else if ((rc = dev_alloc(ctx, &pr->d_pi, cfg->num_public_inputs ? cfg->num_public_inputs : 1))) {}
        else if (pr->blinding && (rc = dev_alloc(ctx, &pr->d_salts, 3ull * QPZK_SALT_SIZE * L))) {}
        else if ((rc = dev_alloc(ctx, &pr->d_chal, 3 * nc + 8))) {}            // betas gammas alphas zeta g*zeta digest
        else if ((rc = dev_alloc(ctx, &pr->d_out, lay.total_words + 1))) {}    // + status word
        else if ((rc = dev_alloc(ctx, &pr->d_misc, 8 + cfg->fri.num_query_rounds))) {}
        else if ((rc = dev_alloc(ctx, &pr->d_ch, (sizeof(ChallengerDev) + 7) / 8))) {}
    }
    if (rc) return fail_out(rc);
    cudaError_t e = cudaHostAlloc((void**)&pr->h_out, (lay.total_words + 1) * sizeof(uint64_t), cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pr->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pr->d_refs, pr->refs.data(), pr->refs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pr->d_chal + 3 * nc + 4, cfg->circuit_digest, 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // both sources are caller / stack memory
    if (e != cudaSuccess) { fail_out(0); CK(e); }
    *out = pr;
    return 0;
    GUARD_END
}

extern "C" void qpzk_prover_free(qpzk_prover* pr) {
    if (!pr) return;
    qpzk_ctx* ctx = pr->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (pr->exec) cudaGraphExecDestroy(pr->exec);
    if (pr->done) cudaEventDestroy(pr->done);
    if (pr->h_out) cudaFreeHost(pr->h_out);
    {
        CacheSwap swap(pr);
        const uint64_t n = 1ull << pr->log_n;
        gate_program_release(ctx, &pr->gates);
        dev_free(ctx, pr->d_refs);
        dev_free(ctx, pr->d_wires, (uint64_t)pr->cfg.num_wires * n);
        dev_free(ctx, pr->d_pi);
        dev_free(ctx, pr->d_salts);
        dev_free(ctx, pr->d_chal, 3 * pr->cfg.num_challenges + 8);
        dev_free(ctx, pr->d_out, pr->lay.total_words + 1);
        dev_free(ctx, pr->d_misc);
        dev_free(ctx, pr->d_ch, (sizeof(ChallengerDev) + 7) / 8);
    }
    // the private blocks join the context's cache (mode 2 keeps them marked dirty until they leave the library)
    for (auto& kv : pr->cache) { ctx->cache.emplace(kv.first, kv.second); ctx->cached_bytes += kv.first; }
    delete pr;
    ctx_unref(ctx);
}

extern "C" int qpzk_prover_get_layout(const qpzk_prover* pr, qpzk_prover_layout* out) {
    if (!pr || !out) return fail(QPZK_EINVAL, "null argument");
    *out = pr->lay;
    return 0;
}
extern "C" uint64_t qpzk_prover_launches_per_proof(const qpzk_prover* pr) { return pr ? pr->launches_per_proof : 0; }
extern "C" int qpzk_prover_uses_graph(const qpzk_prover* pr) { return pr && pr->exec ? 1 : 0; }

extern "C" int qpzk_prover_prove_async(qpzk_prover* pr, const uint64_t* wires, const uint64_t* public_inputs,
                                       const uint64_t* salts, uint64_t pow_witness_in) {
    GUARD_BEGIN
    if (!pr || !wires) return fail(QPZK_EINVAL, "null argument");
    if (pr->in_flight) return fail(QPZK_EINVAL, "a proof is in flight on this prover: call qpzk_prover_finish first");
    qpzk_ctx* ctx = pr->ctx;
    const qpzk_prover_config& c = pr->cfg;
    if (c.num_public_inputs && !public_inputs) return fail(QPZK_EINVAL, "public_inputs is null");
    if (pr->blinding && !salts) return fail(QPZK_EINVAL, "the circuit is zero-knowledge: salts are required");
    for (uint32_t i = 0; i < c.num_public_inputs; i++) if (public_inputs[i] >= GL_P) return fail(QPZK_EINVAL, "public input is not canonical");
    if (pow_witness_in != QPZK_POW_SEARCH && pow_witness_in >= GL_P) return fail(QPZK_EINVAL, "pow witness is not canonical");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = 1ull << pr->log_n, L = n << pr->rate_bits;
    CK(cudaMemcpyAsync(pr->d_wires, wires, (uint64_t)c.num_wires * n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (c.num_public_inputs)
        CK(cudaMemcpyAsync(pr->d_pi, public_inputs, c.num_public_inputs * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (pr->blinding)
        CK(cudaMemcpyAsync(pr->d_salts, salts, 3ull * QPZK_SALT_SIZE * L * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->h2d_done, st));
    const bool graph_ok = c.use_graph && pow_witness_in == QPZK_POW_SEARCH && pr->eager_runs >= 2;
    int rc = 0;
    if (graph_ok && !pr->exec) {
        // capture the sequence once; the kernels are NOT run by the capture
        CacheSwap swap(pr);
        const uint64_t l0 = ctx->launches;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        rc = enqueue_proof(pr, pow_witness_in);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        pr->launches_per_proof = ctx->launches - l0;
        ctx->launches = l0;
        if (!rc && e == cudaSuccess) e = cudaGraphInstantiate(&pr->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc) return rc;
        if (e != cudaSuccess) { pr->exec = nullptr; cudaGetLastError(); pr->cfg.use_graph = 0; }   // fall back to eager enqueueing
    }
    if (pr->exec && graph_ok) {
        CK(cudaGraphLaunch(pr->exec, st));
        ctx->launches += pr->launches_per_proof;
    } else {
        CacheSwap swap(pr);
        const uint64_t l0 = ctx->launches;
        rc = enqueue_proof(pr, pow_witness_in);
        pr->launches_per_proof = ctx->launches - l0;
        pr->eager_runs++;
        if (rc) { cudaStreamSynchronize(st); return rc; }
    }
    CK(cudaMemcpyAsync(pr->h_out, pr->d_out, (pr->lay.total_words + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(pr->done, st));
    pr->in_flight = true;
    // the caller's buffers have been read when this returns (they may be zeroized at once), the proof is still running
    CK(cudaEventSynchronize(ctx->h2d_done));
    return 0;
    GUARD_END
}

extern "C" int qpzk_prover_finish(qpzk_prover* pr, uint64_t* out_words) {
    GUARD_BEGIN
    if (!pr || !out_words) return fail(QPZK_EINVAL, "null argument");
    if (!pr->in_flight) return fail(QPZK_EINVAL, "no proof in flight");
    CK(cudaSetDevice(pr->ctx->device));
    pr->in_flight = false;
    CK(cudaEventSynchronize(pr->done));
    const uint32_t status = (uint32_t)pr->h_out[pr->lay.total_words];
    if (status == 2) return fail(QPZK_EPOW, "proof of work failed: no witness found");
    if (status != 1) return fail(QPZK_EPOW, "proof-of-work witness does not reach the required leading zeros");
    memcpy(out_words, pr->h_out, pr->lay.total_words * sizeof(uint64_t));
    return 0;
    GUARD_END
}

extern "C" int qpzk_prover_prove(qpzk_prover* pr, const uint64_t* wires, const uint64_t* public_inputs,
                                 const uint64_t* salts, uint64_t pow_witness_in, uint64_t* out_words) {
    int rc = qpzk_prover_prove_async(pr, wires, public_inputs, salts, pow_witness_in);
    if (rc) return rc;
    return qpzk_prover_finish(pr, out_words);
}
