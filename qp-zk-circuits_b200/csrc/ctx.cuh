// Private declarations shared by the translation units that implement the C ABI (capi.cu, prove.cu): the context with
// its block cache, the batch / challenger handles and the stage functions behind the entry points.  Not part of the ABI.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/qpzk.h"
#include "internal.cuh"

using namespace qpzk;
int fail(int code, const std::string& msg);   // sets the thread-local message of qpzk_last_error(); returns code

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char buf_[512];                                                                            \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(e_ == cudaErrorMemoryAllocation ? QPZK_ENOMEM : QPZK_ECUDA, buf_);              \
        }                                                                                              \
    } while (0)

// NVTX range per stage of the path (the role of the reference's phase brackets, wormhole/memprof/src/report.rs:18-54,
// and of plonky2's TimingTree scopes): visible in nsys / ncu --nvtx, free when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define GUARD_BEGIN try {
#define GUARD_END                                                  \
    }                                                              \
    catch (const std::bad_alloc&) { return fail(QPZK_ENOMEM, "host allocation failed"); } \
    catch (const std::exception& ex) { return fail(QPZK_EINTERNAL, ex.what()); }          \
    catch (...) { return fail(QPZK_EINTERNAL, "unknown exception"); }

struct PowTables {
    uint64_t *lo = nullptr, *hi = nullptr;  // device
    uint64_t hi_size = 0;
};
struct CosetTables {
    uint64_t *lo = nullptr, *hi = nullptr;  // device, [cosets][TW_LO_SIZE] / [cosets][hi_size]
    uint64_t hi_size = 0;
};

struct qpzk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 0;
    int zeroize = 1;           // 0 off, 1 at release, 2 deferred until the block leaves the library (qpzk_ctx_set_zeroize)
    std::map<void*, bool> dirty;   // mode 2: cached blocks that held secrets and have not been zeroed yet
    bool legacy_ntt = false;   // QPZK_LEGACY_NTT=1: generic strided passes only (A/B testing of ntt_fwd.cu)
    bool timing = false;
    bool keep_values = false;  // from_values batches keep their input values on the device (qpzk_ctx_set_keep_values)
    cudaStream_t copy_stream = nullptr;   // H2D of host-buffer commits, overlapped with the per-column stages
    cudaEvent_t copy_done = nullptr;
    cudaEvent_t h2d_done = nullptr, h2d_done2 = nullptr;   // last read of the caller's column / salt buffers (host-buffer calls wait on them before returning)
    cudaEvent_t ev[5] = {};
    bool ev_valid[5] = {};
    uint64_t launches = 0;
    uint64_t device_allocs = 0;   // cudaMalloc calls of the block cache (0 in steady state)
    // The context stays alive until qpzk_ctx_destroy has been called AND every batch / challenger created
    // from it has been freed, whatever order a binding's finalizers run in.
    int refs = 1;
    std::multimap<uint64_t, void*> cache;                              // free device blocks by size in bytes
    std::map<void*, uint64_t> block_bytes;                             // every block this context allocated
    uint64_t cached_bytes = 0;
    uint64_t owned_bytes = 0, peak_owned_bytes = 0;   // device memory this context holds (in use + cached) and its high-water mark
    std::map<std::pair<uint32_t, int>, PowTables> pow_tables;        // (log_size, inverse)
    std::map<std::pair<uint32_t, uint32_t>, CosetTables> coset_tables;  // (log_n, rate_bits)
    std::map<uint32_t, PowTables> unshift_tables;                       // log_size -> powers of g^-1 (coset_ifft)
    uint64_t* kpow = nullptr;                                           // device: g^j (coset shifts k_j), kpow_size entries
    uint32_t kpow_size = 0;
    std::map<std::pair<uint32_t, uint32_t>, uint64_t*> zh_tables;       // (log_n, qbits) -> Z_H on the quotient coset and its inverses
};

struct qpzk_batch {
    qpzk_ctx* ctx = nullptr;
    uint32_t log_n = 0, rate_bits = 0, cap_height = 0, num_cols = 0, salt_cols = 0, leaf_width = 0;
    uint64_t num_leaves = 0;
    uint32_t log_leaves = 0, num_levels = 0;  // levels[0] = leaf digests ... levels[num_levels-1] = cap
    uint64_t* coeffs = nullptr;               // num_cols x n (null for a bare MerkleTree)
    uint64_t* values = nullptr;               // num_cols x n values on the subgroup (only with keep_values)
    uint64_t* leaves = nullptr;
    uint64_t* digests = nullptr;
    bool own_leaves = true;
};

struct qpzk_challenger {
    qpzk_ctx* ctx = nullptr;
    ChallengerDev* dev = nullptr;
    uint64_t* scratch = nullptr;  // device staging for observed elements / drawn challenges
    uint32_t scratch_words = 0;
};


// ---- capi.cu: shared stage functions -------------------------------------------------------------------------------
uint32_t ilog2(uint64_t x);
int dev_alloc(qpzk_ctx* ctx, uint64_t** p, uint64_t words);
void dev_free(qpzk_ctx* ctx, void* p, uint64_t zero_words = 0);
void scrub(qpzk_ctx* ctx, void* p, uint64_t words);
void ctx_unref(qpzk_ctx* ctx);
void batch_release(qpzk_batch* b);
int get_pow_tables(qpzk_ctx* ctx, uint32_t log_size, bool inverse, PowTables* out);
int get_coset_tables(qpzk_ctx* ctx, uint32_t log_n, uint32_t rate_bits, CosetTables* out);

// Device-side inputs / outputs of fri_core.  Pointers that are null fall back to the host-side instance.
struct FriDeviceIO {
    const uint64_t* points = nullptr;      // 2 * num_batches words: the opening points (else inst->points, host)
    const uint32_t* poly_refs = nullptr;   // poly_oracle[total] then poly_index[total] (else copied from inst, host)
    bool observe_openings = false;         // gather the opened values and observe them before alpha is drawn
    uint64_t* proof = nullptr;             // flat FriProof words (required)
    uint64_t* openings = nullptr;          // 2 * total polynomials, instance order (required)
    uint64_t* indices = nullptr;           // num_query_rounds words (required)
    uint32_t* status = nullptr;            // status[0]: 1 ok, 0 witness too weak, 2 search exhausted (required)
};
int fri_core(qpzk_ctx* ctx, const qpzk_batch* const* oracles, uint32_t num_oracles, const qpzk_fri_instance* inst,
             const qpzk_fri_params* p, ChallengerDev* ch, uint64_t pow_witness_in, const FriDeviceIO& io);

// ---- stage functions shared with the single-call prover (prove.cu) ---------------------------------------------------
struct ColumnSource {   // where a commit's input columns come from: device-resident, or a host buffer copied in column ranges
    const uint64_t* dev = nullptr;
    const uint64_t* host = nullptr;
    uint64_t* staging = nullptr;
    uint32_t chunks = 1;
    const uint64_t* const* slices = nullptr;   // coefficients spread over several device buffers (peers mapped over NVLink)
    uint32_t cols_per_slice = 0;
};
inline ColumnSource dev_source(const uint64_t* p) { ColumnSource s; s.dev = p; return s; }
int commit_impl(qpzk_ctx* ctx, ColumnSource src, bool from_values, uint32_t log_n, uint32_t num_cols, uint32_t rate_bits,
                uint32_t cap_height, const uint64_t* salt_dev, uint32_t coset_begin, uint32_t coset_count,
                bool keep_coeffs_copy, qpzk_batch** out);
int get_kpow_table(qpzk_ctx* ctx, uint32_t count, const uint64_t** out);
int commit_zs_impl(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires, uint32_t sigma_col0,
                   uint32_t num_routed, uint32_t chunk, uint32_t num_challenges, const uint64_t* betas,
                   const uint64_t* gammas, const uint64_t* chal_dev, uint32_t cap_height, const uint64_t* salt_dev,
                   qpzk_batch** out);
int quotient_impl(qpzk_ctx* ctx, const qpzk_batch* cs, const qpzk_batch* wires, const qpzk_batch* zs,
                  uint32_t sigma_col0, uint32_t num_routed, uint32_t chunk, uint32_t num_challenges,
                  const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas, const uint64_t* chal_dev,
                  const uint64_t* gate_terms_dev, uint32_t num_gate_terms, uint64_t** coeffs_out, uint32_t* qbits_out);
struct GateProgramDev {
    uint64_t* code = nullptr; uint64_t* imm = nullptr; uint32_t num_ops = 0, num_terms = 0, num_regs = 0;
    cudaKernel_t jit = nullptr;   // the program compiled at run time (gate_jit.cu); nullptr = interpret
};
namespace qpzk {
cudaKernel_t gate_jit_get(const uint32_t* code, uint32_t num_ops, const uint64_t* imm, uint32_t num_imm, uint32_t num_regs);
cudaError_t launch_gate_jit(cudaKernel_t k, const GateProgramArgs& a, cudaStream_t st);
}
int gate_program_upload(qpzk_ctx* ctx, uint32_t cs_cols, uint32_t wire_cols, const qpzk_gate_program* prog, GateProgramDev* out);
void gate_program_release(qpzk_ctx* ctx, GateProgramDev* g);
int gate_terms_run(qpzk_ctx* ctx, const qpzk_batch* cs, const qpzk_batch* wires, uint32_t chunk, const GateProgramDev& g,
                   uint64_t** terms_out);

struct DevBufs {  // frees everything it holds on scope exit
    qpzk_ctx* ctx;
    std::vector<std::pair<void*, uint64_t>> ptrs;   // block, words to scrub at release (0 = holds nothing secret)
    explicit DevBufs(qpzk_ctx* c) : ctx(c) {}
    ~DevBufs() { for (auto& p : ptrs) dev_free(ctx, p.first, p.second); }
    int alloc(uint64_t** p, uint64_t words, bool secret = false) {
        int rc = dev_alloc(ctx, p, words);
        if (!rc) ptrs.emplace_back(*p, secret ? (words ? words : 1) : 0);
        return rc;
    }
};
