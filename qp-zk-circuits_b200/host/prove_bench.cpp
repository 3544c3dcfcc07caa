// Native host-side driver of the C ABI (include/qpzk.h): the commitment / opening half of plonky2's `prove()` for
// leaf-shaped synthetic witnesses, T host threads with one context (device + stream) each.
//
// This is the C++ stand-in for what a `[patch.crates-io]` fork of qp-plonky2 =1.5.5 does from Rust (INTEGRATION.md;
// the reference toolchain is absent in this image): `prove_with_partition_witness` steps 2-9,
// reached in the reference from wormhole/prover/src/lib.rs:171-175 and, one leaf proof per worker, from the
// aggregators (wormhole/aggregator/src/aggregator.rs:22-43 proves on detached worker threads without a lock, which is
// exactly the threading model exercised here).  Same transcript as qp-zk-circuits_b200/qpzk/prover.py.
//
//   prove_bench [--device D] [--devices G] [--threads T] [--proofs N] [--log-n 13] [--pinned 1] [--perm 0|1] [--zeroize 0|1|2] [--gates FILE]
//               [--prover 1 [--inflight K] [--graph 0|1] [--routed R]]
// --prover 1: the whole proof is ONE call (qpzk_prover_prove_async / _finish: challenges stay on the device, CUDA graph
// replay from the third proof; always with the permutation argument and, with --gates, the gate program).  --inflight K:
// every host thread drives K provers on K contexts and keeps K proofs in flight (their graphs overlap on the GPU) --
// the aggregator's dispatch of independent leaf proofs without one host thread per proof.
// --gates FILE (with --perm 1): the circuit's gate constraints as a qpzk_gate_program written by
// tools/export_gate_program.py (upstream's PoseidonGate on every row), interpreted on the device for the quotient.
// --devices G: one process drives G GPUs, thread t on device D + (t mod G) -- the aggregator's dispatch (independent leaf
// proofs, one context per worker thread and device, no collective; reference wormhole/aggregator/src/aggregator.rs:187-227).
// --perm 1: the zs/partial-products columns and the quotient chunks are not random stand-ins but computed on the device
// from the routed wires, the circuit's sigma columns and the transcript's challenges (qpzk_commit_zs_partial_products,
// qpzk_commit_quotient: SURVEY.md 8a rows a16/a17, permutation argument only -- the gate set is not part of this repo).
// prints one JSON line: proofs/s over all threads, mean latency per proof, and a checksum of every thread's first
// proof (identical inputs must give identical proofs on every thread).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qpzk.h"

static const uint64_t P = 0xFFFFFFFF00000001ULL;

static uint64_t mulmod(uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % P); }

static uint64_t root_of_unity(unsigned log_n) {
    uint64_t w = 7277203076849721926ULL;  // order 2^32
    for (unsigned i = log_n; i < 32; i++) w = mulmod(w, w);
    return w;
}

struct SplitMix {  // SURVEY.md 8d input recipe: SplitMix64, rejection-sampled below p
    uint64_t s;
    uint64_t next() {
        for (;;) {
            uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            z ^= z >> 31;
            if (z < P) return z;
        }
    }
};

static bool g_pinned = true, g_breakdown = false, g_perm = false;
static int g_zeroize = 1;      // qpzk_ctx_set_zeroize mode
static std::vector<uint32_t> g_gate_code;   // --gates: program shared by all threads (read only)
static std::vector<uint64_t> g_gate_imm;
static uint32_t g_gate_regs = 0;
static bool g_prover = false, g_graph = true;
static int g_inflight = 1;
static int g_stagger_us = 0;   // thread t starts t * g_stagger_us after the others (spreads the phases of concurrent provers)

struct HostBuf {   // pinned (default) or pageable host buffer
    uint64_t* p = nullptr;
    size_t words = 0;
    bool pinned = false;
    HostBuf() = default;
    explicit HostBuf(size_t w) : words(w), pinned(g_pinned) {
        p = pinned ? (uint64_t*)qpzk_pinned_alloc(w * 8) : (uint64_t*)malloc(w * 8);
        if (!p) { fprintf(stderr, "host allocation failed\n"); exit(2); }
    }
    HostBuf(HostBuf&& o) noexcept : p(o.p), words(o.words), pinned(o.pinned) { o.p = nullptr; }
    HostBuf(const HostBuf&) = delete;
    ~HostBuf() { if (p) { if (pinned) qpzk_pinned_free(p); else free(p); } }
    uint64_t* data() { return p; }
    size_t size() const { return words; }
};

static HostBuf random_columns(uint64_t seed, size_t cols, size_t n) {
    SplitMix g{seed};
    HostBuf v(cols * n);
    for (size_t i = 0; i < v.size(); i++) v.p[i] = g.next();
    return v;
}

#define QCHECK(call)                                                                          \
    do {                                                                                      \
        int rc_ = (call);                                                                     \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, qpzk_last_error()); exit(2); } \
    } while (0)

struct Shape { unsigned log_n = 13, constants_sigmas = 85, wires = 135, zs_pp = 20, quotient = 16, num_zs = 2, routed = 80, sigma_col0 = 5; };

struct Worker {
    int device;
    int index = 0;
    Shape sh;
    int proofs;
    double seconds = 0;
    uint64_t checksum = 0;
    uint64_t launches = 0, allocs = 0;
    double max_proof_s = 0;
    int graph = 0;
    uint64_t peak_bytes = 0;

    // one-call prover: K contexts, K provers, K proofs in flight from this thread
    void run_prover(std::atomic<int>* ready, std::atomic<bool>* go) {
        const size_t n = (size_t)1 << sh.log_n;
        const int K = g_inflight;
        auto cs = random_columns(1, sh.constants_sigmas, n);
        std::vector<HostBuf> wires;
        for (int s = 0; s < 4; s++) wires.push_back(random_columns(100 + s, sh.wires, n));
        uint64_t pi[21];
        for (int i = 0; i < 21; i++) pi[i] = 1000 + i;
        qpzk_prover_config cfg{};
        cfg.num_wires = sh.wires; cfg.num_routed = sh.routed; cfg.sigma_col0 = sh.sigma_col0; cfg.quotient_degree_factor = 8;
        cfg.num_challenges = sh.num_zs; cfg.num_public_inputs = 21; cfg.blinding = 0; cfg.use_graph = g_graph ? 1 : 0;
        for (int i = 0; i < 4; i++) cfg.circuit_digest[i] = 11 + i;
        QCHECK(qpzk_fri_params_constant_arity(&cfg.fri, sh.log_n, 3, 4, 16, 28, 4, 5));
        qpzk_gate_program prog{g_gate_code.data(), (uint32_t)(g_gate_code.size() / 4), g_gate_imm.data(), (uint32_t)g_gate_imm.size(), g_gate_regs};
        cfg.gates = g_gate_regs ? &prog : nullptr;
        std::vector<qpzk_ctx*> ctxs(K);
        std::vector<qpzk_batch*> csb(K);
        std::vector<qpzk_prover*> prs(K);
        qpzk_prover_layout lay{};
        for (int k = 0; k < K; k++) {
            QCHECK(qpzk_ctx_create(device, &ctxs[k]));
            QCHECK(qpzk_ctx_set_zeroize(ctxs[k], g_zeroize));
            QCHECK(qpzk_ctx_set_keep_values(ctxs[k], 1));
            QCHECK(qpzk_commit_from_values(ctxs[k], cs.data(), sh.log_n, sh.constants_sigmas, 3, 4, nullptr, &csb[k]));
            QCHECK(qpzk_ctx_set_keep_values(ctxs[k], 0));
            QCHECK(qpzk_prover_create(ctxs[k], csb[k], &cfg, &prs[k]));
            QCHECK(qpzk_prover_get_layout(prs[k], &lay));
        }
        std::vector<uint64_t> out(lay.total_words);
        for (int k = 0; k < K; k++)
            for (int w = 0; w < 4; w++) QCHECK(qpzk_prover_prove(prs[k], wires[w & 1].data(), pi, nullptr, QPZK_POW_SEARCH, out.data()));   // eager x2, capture, replay
        QCHECK(qpzk_prover_prove(prs[0], wires[0].data(), pi, nullptr, QPZK_POW_SEARCH, out.data()));
        for (size_t i = 0; i < out.size(); i++) checksum = (checksum ^ out[i]) * 0x100000001B3ULL;
        ready->fetch_add(1);
        while (!go->load()) std::this_thread::yield();
        uint64_t l0 = 0, a0 = 0;
        for (int k = 0; k < K; k++) { l0 += qpzk_ctx_launch_count(ctxs[k]); a0 += qpzk_ctx_device_allocs(ctxs[k]); }
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::chrono::steady_clock::time_point> started(K);
        for (int i = 0; i < proofs + K; i++) {
            const int k = i % K;
            if (i >= K) {
                QCHECK(qpzk_prover_finish(prs[k], out.data()));
                const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - started[k]).count();
                if (dt > max_proof_s) max_proof_s = dt;
            }
            if (i < proofs) {
                started[k] = std::chrono::steady_clock::now();
                QCHECK(qpzk_prover_prove_async(prs[k], wires[i & 3].data(), pi, nullptr, QPZK_POW_SEARCH));
            }
        }
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (int k = 0; k < K; k++) { launches += qpzk_ctx_launch_count(ctxs[k]); allocs += qpzk_ctx_device_allocs(ctxs[k]); }
        launches -= l0; allocs -= a0;
        graph = qpzk_prover_uses_graph(prs[0]);
        peak_bytes = qpzk_ctx_peak_bytes(ctxs[0]);
        for (int k = 0; k < K; k++) { qpzk_prover_free(prs[k]); qpzk_batch_free(csb[k]); qpzk_ctx_destroy(ctxs[k]); }
    }

    void run(std::atomic<int>* ready, std::atomic<bool>* go) {
        if (g_prover) return run_prover(ready, go);
        const size_t n = (size_t)1 << sh.log_n;
        qpzk_ctx* ctx;
        QCHECK(qpzk_ctx_create(device, &ctx));
        QCHECK(qpzk_ctx_set_zeroize(ctx, g_zeroize));
        // what CircuitBuilder::build leaves behind: the committed constants/sigmas oracle
        auto cs = random_columns(1, sh.constants_sigmas, n);
        qpzk_batch* cs_b;
        if (g_perm) QCHECK(qpzk_ctx_set_keep_values(ctx, 1));   // the permutation argument reads the sigma VALUES
        QCHECK(qpzk_commit_from_values(ctx, cs.data(), sh.log_n, sh.constants_sigmas, 3, 4, nullptr, &cs_b));
        QCHECK(qpzk_ctx_set_keep_values(ctx, 0));
        // four witnesses' worth of columns, reused round robin
        std::vector<HostBuf> wires, zs, quot;
        for (int s = 0; s < 4; s++) {
            wires.push_back(random_columns(100 + s, sh.wires, n));
            zs.push_back(random_columns(200 + s, sh.zs_pp, n));
            quot.push_back(random_columns(300 + s, sh.quotient, n));
        }
        qpzk_fri_params fp;
        QCHECK(qpzk_fri_params_constant_arity(&fp, sh.log_n, 3, 4, 16, 28, 4, 5));
        const uint32_t leaf_lens[4] = {sh.constants_sigmas, sh.wires, sh.zs_pp, sh.quotient};
        const size_t proof_words = qpzk_fri_proof_words(&fp, sh.log_n, 4, leaf_lens);
        const uint32_t total = sh.constants_sigmas + sh.wires + sh.zs_pp + sh.quotient;
        std::vector<uint32_t> po, pi;
        for (uint32_t o = 0; o < 4; o++) for (uint32_t i = 0; i < leaf_lens[o]; i++) { po.push_back(o); pi.push_back(i); }
        for (uint32_t i = 0; i < sh.num_zs; i++) { po.push_back(2); pi.push_back(i); }
        const uint32_t sizes[2] = {total, sh.num_zs};
        HostBuf proof(proof_words), openings(2 * (total + sh.num_zs)), znext(2 * sh.zs_pp);
        const uint64_t digest[4] = {11, 12, 13, 14}, pih[4] = {21, 22, 23, 24};
        const uint64_t g = root_of_unity(sh.log_n);

        double phase[8] = {0}, phase_max[8] = {0};
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto lap = [&](int k, std::chrono::steady_clock::time_point& t) {
            if (!g_breakdown) return;
            QCHECK(qpzk_ctx_synchronize(ctx));
            auto t1 = now();
            const double dt = std::chrono::duration<double>(t1 - t).count();
            phase[k] += dt;
            if (dt > phase_max[k]) phase_max[k] = dt;
            t = t1;
        };
        auto prove_one = [&](int s) {
            auto tp = now();
            qpzk_challenger* ch;
            QCHECK(qpzk_challenger_create(ctx, &ch));
            QCHECK(qpzk_challenger_observe(ch, digest, 4));
            QCHECK(qpzk_challenger_observe(ch, pih, 4));
            qpzk_batch *w_b, *z_b, *q_b;
            uint64_t chal[4], zeta[2];
            if (g_perm) QCHECK(qpzk_ctx_set_keep_values(ctx, 1));
            QCHECK(qpzk_commit_from_values(ctx, wires[s].data(), sh.log_n, sh.wires, 3, 4, nullptr, &w_b));
            QCHECK(qpzk_ctx_set_keep_values(ctx, 0));
            QCHECK(qpzk_challenger_observe_cap(ch, w_b));
            QCHECK(qpzk_challenger_get(ch, chal, 4));   // betas, gammas
            lap(0, tp);
            if (g_perm)
                QCHECK(qpzk_commit_zs_partial_products(ctx, cs_b, w_b, sh.sigma_col0, sh.routed, 8, 2, chal, chal + 2, 4, nullptr, &z_b));
            else
                QCHECK(qpzk_commit_from_values(ctx, zs[s].data(), sh.log_n, sh.zs_pp, 3, 4, nullptr, &z_b));
            QCHECK(qpzk_challenger_observe_cap(ch, z_b));
            uint64_t alphas[2];
            QCHECK(qpzk_challenger_get(ch, alphas, 2));
            lap(1, tp);
            if (g_perm && g_gate_regs) {
                qpzk_gate_program prog{g_gate_code.data(), (uint32_t)(g_gate_code.size() / 4), g_gate_imm.data(),
                                       (uint32_t)g_gate_imm.size(), g_gate_regs};
                QCHECK(qpzk_commit_quotient_prog(ctx, cs_b, w_b, z_b, sh.sigma_col0, sh.routed, 8, 2, chal, chal + 2, alphas, &prog, 4, nullptr, &q_b));
            } else if (g_perm)
                QCHECK(qpzk_commit_quotient(ctx, cs_b, w_b, z_b, sh.sigma_col0, sh.routed, 8, 2, chal, chal + 2, alphas, nullptr, 0, 4, nullptr, &q_b));
            else
                QCHECK(qpzk_commit_from_coeffs(ctx, quot[s].data(), sh.log_n, sh.quotient, 3, 4, nullptr, &q_b));
            QCHECK(qpzk_challenger_observe_cap(ch, q_b));
            QCHECK(qpzk_challenger_get(ch, zeta, 2));
            lap(2, tp);
            const uint64_t gz[2] = {mulmod(zeta[0], g), mulmod(zeta[1], g)};
            const qpzk_batch* oracles[4] = {cs_b, w_b, z_b, q_b};
            uint64_t* op = openings.data();
            QCHECK(qpzk_batches_eval_ext(ctx, oracles, 4, zeta, op));
            op += 2 * total;
            QCHECK(qpzk_batch_eval_ext(z_b, gz, znext.data()));
            memcpy(op, znext.data(), 2 * sh.num_zs * sizeof(uint64_t));
            lap(3, tp);
            QCHECK(qpzk_challenger_observe(ch, openings.data(), (uint32_t)openings.size()));
            lap(4, tp);
            const uint64_t points[4] = {zeta[0], zeta[1], gz[0], gz[1]};
            qpzk_fri_instance inst{2, points, sizes, po.data(), pi.data()};
            QCHECK(qpzk_fri_prove(ctx, oracles, 4, &inst, &fp, ch, QPZK_POW_SEARCH, proof.data(), nullptr));
            lap(5, tp);
            qpzk_batch_free(w_b); qpzk_batch_free(z_b); qpzk_batch_free(q_b);
            qpzk_challenger_free(ch);
            lap(6, tp);
        };
        prove_one(0); prove_one(1);   // warm-up: tables, block cache
        for (double& x : phase) x = 0;
        for (double& x : phase_max) x = 0;
        for (size_t i = 0; i < proof.size(); i++) checksum = (checksum ^ proof.p[i]) * 0x100000001B3ULL;
        QCHECK(qpzk_ctx_synchronize(ctx));
        ready->fetch_add(1);
        while (!go->load()) std::this_thread::yield();
        if (g_stagger_us) std::this_thread::sleep_for(std::chrono::microseconds((long)g_stagger_us * index));
        const uint64_t l0 = qpzk_ctx_launch_count(ctx), a0 = qpzk_ctx_device_allocs(ctx);
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < proofs; i++) {
            auto tp0 = std::chrono::steady_clock::now();
            prove_one(i & 3);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count();
            if (dt > max_proof_s) max_proof_s = dt;
        }
        QCHECK(qpzk_ctx_synchronize(ctx));
        seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        launches = qpzk_ctx_launch_count(ctx) - l0;
        allocs = qpzk_ctx_device_allocs(ctx) - a0;
        if (g_breakdown)
            fprintf(stderr, "thread %d longest single phase (ms): wires %.1f | zs %.1f | quotient %.1f | openings %.1f | observe %.1f | fri %.1f | free %.1f\n",
                    index, 1e3 * phase_max[0], 1e3 * phase_max[1], 1e3 * phase_max[2], 1e3 * phase_max[3], 1e3 * phase_max[4],
                    1e3 * phase_max[5], 1e3 * phase_max[6]);
        if (g_breakdown)
            fprintf(stderr, "per proof (ms): wires commit+challenges %.3f | zs commit %.3f | quotient commit %.3f | openings %.3f | "
                            "observe openings %.3f | fri_prove %.3f | free %.3f\n",
                    1e3 * phase[0] / proofs, 1e3 * phase[1] / proofs, 1e3 * phase[2] / proofs, 1e3 * phase[3] / proofs,
                    1e3 * phase[4] / proofs, 1e3 * phase[5] / proofs, 1e3 * phase[6] / proofs);
        qpzk_batch_free(cs_b);
        qpzk_ctx_destroy(ctx);
    }
};

int main(int argc, char** argv) {
    int device = 0, devices = 1, threads = 1, proofs = 16;
    Shape sh;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i];
        if (k == "--gates") {
            FILE* f = fopen(argv[i + 1], "rb");
            uint32_t hdr[4];
            if (!f || fread(hdr, 4, 4, f) != 4 || memcmp(hdr, "QPGP", 4) != 0) { fprintf(stderr, "cannot read gate program %s\n", argv[i + 1]); return 1; }
            g_gate_code.resize(4ull * hdr[1]); g_gate_imm.resize(hdr[2]); g_gate_regs = hdr[3];
            if (fread(g_gate_code.data(), 16, hdr[1], f) != hdr[1] || fread(g_gate_imm.data(), 8, hdr[2], f) != hdr[2]) { fprintf(stderr, "truncated gate program\n"); return 1; }
            fclose(f);
            continue;
        }
        int v = atoi(argv[i + 1]);
        if (k == "--device") device = v;
        else if (k == "--devices") devices = v > 0 ? v : 1;
        else if (k == "--threads") threads = v;
        else if (k == "--proofs") proofs = v;
        else if (k == "--log-n") sh.log_n = (unsigned)v;
        else if (k == "--pinned") g_pinned = v != 0;
        else if (k == "--breakdown") g_breakdown = v != 0;
        else if (k == "--perm") g_perm = v != 0;
        else if (k == "--stagger-us") g_stagger_us = v;
        else if (k == "--zeroize") g_zeroize = v;
        else if (k == "--prover") g_prover = v != 0;
        else if (k == "--inflight") g_inflight = v > 0 ? v : 1;
        else if (k == "--graph") g_graph = v != 0;
        else if (k == "--routed") sh.routed = (unsigned)v;
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 1; }
    }
    std::vector<Worker> ws(threads);
    std::vector<std::thread> th;
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    for (int t = 0; t < threads; t++) {
        ws[t].device = device + t % devices; ws[t].index = t; ws[t].sh = sh; ws[t].proofs = proofs;
        th.emplace_back([&, t] { ws[t].run(&ready, &go); });
    }
    while (ready.load() < threads) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    go.store(true);
    for (auto& t : th) t.join();
    (void)t0;
    double lat = 0, wall = 0, fastest = 1e30, maxp = 0;
    uint64_t allocs = 0;   // wall = the slowest thread's proving loop (threads start together; teardown excluded)
    bool same = true;
    for (auto& w : ws) {
        lat += w.seconds / proofs; same = same && w.checksum == ws[0].checksum;
        if (w.seconds > wall) wall = w.seconds;
        if (w.seconds < fastest) fastest = w.seconds;
        if (w.max_proof_s > maxp) maxp = w.max_proof_s;
        allocs += w.allocs;
    }
    printf("{\"prover\": %d, \"inflight\": %d, \"graph\": %d, \"routed\": %u, \"peak_device_MB\": %.1f, ", g_prover ? 1 : 0, g_prover ? g_inflight : 1, ws[0].graph, sh.routed, ws[0].peak_bytes / 1e6);
    printf("\"threads\": %d, \"devices\": %d, \"proofs_per_thread\": %d, \"log_n\": %u, \"perm\": %d, \"zeroize\": %d, \"gate_program_ops\": %u, \"proofs_per_s\": %.2f, \"latency_ms\": %.3f, "
           "\"launches_per_proof\": %llu, \"proofs_identical_across_threads\": %s, \"checksum\": \"%016llx\", \"slowest_thread_s\": %.4f, \"fastest_thread_s\": %.4f, \"max_proof_ms\": %.2f, \"device_allocs_in_loop\": %llu}\n",
           threads, devices, proofs, sh.log_n, (g_perm || g_prover) ? 1 : 0, g_zeroize, (unsigned)(g_gate_code.size() / 4), threads * proofs / wall, 1e3 * lat / threads,
           (unsigned long long)(ws[0].launches / proofs), same ? "true" : "false", (unsigned long long)ws[0].checksum, wall, fastest, 1e3 * maxp, (unsigned long long)allocs);
    return same ? 0 : 3;
}
