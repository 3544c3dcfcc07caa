// C ABI of libqpzk.so (declared in include/qpzk.h): contexts, root-power tables, the commit pipeline
// (iNTT -> coset LDE -> leaf hashing -> Merkle levels), openings, challenger and the FRI prover driver.
// Host code here only plans passes, owns device buffers and enqueues kernels on the context stream; all
// field arithmetic on data runs in the kernels of ntt.cu / hash.cu / fri.cu.  There is no CPU fallback.
//
// Replaces the host-side orchestration of qp-plonky2 =1.5.5 `PolynomialBatch::{from_values,from_coeffs,
// prove_openings}`, `MerkleTree::new`, `fri_proof` (reference call sites
// wormhole/prover/src/lib.rs:171-175 and the two aggregator provers).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "ctx.cuh"
#include "goldilocks.cuh"

using namespace qpzk;

static thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
namespace qpzk { int wire_fail(int code, const std::string& msg) { return fail(code, msg); } }   // wire.cu

uint32_t ilog2(uint64_t x) { uint32_t l = 0; while ((1ull << l) < x) l++; return l; }

// ---------------------------------------------------------------------------------------------
// device memory helpers (stream-ordered pool)
// ---------------------------------------------------------------------------------------------
// Every buffer of the data path comes from a per-context cache of device blocks: a freed block is kept (size
// keyed) and handed to the next request of about the same size.  All work of a context is ordered on its one
// stream, so reuse needs no event and no driver call; in steady state (the prover commits the same shapes over
// and over) a commit performs no allocation at all.  cudaMalloc runs only while the cache warms up; when it fails,
// the cache is released once and the request retried.  (cudaMallocAsync was measured to stall the host for
// 0.1-3 s per commit when commits are issued back to back: its pool cannot reuse blocks whose frees are still
// pending and keeps mapping new memory instead.)
// zero `words` words of a block that held secret-derived data: now (mode 1) or before the block leaves the library (2)
void scrub(qpzk_ctx* ctx, void* p, uint64_t words) {
    if (!p || !words || !ctx->zeroize) return;
    if (ctx->zeroize == 1) launch_fill_words(p, words, 0, ctx->stream);
    else ctx->dirty[p] = true;
}
// mode 2: zero every dirty block that is about to be returned to the driver (whole block: the cache does not track
// which part the last owner used)
static void scrub_dirty(qpzk_ctx* ctx, bool only_cached) {
    for (auto it = ctx->dirty.begin(); it != ctx->dirty.end();) {
        auto bb = ctx->block_bytes.find(it->first);
        bool cached = false;
        if (only_cached) for (auto& kv : ctx->cache) if (kv.second == it->first) { cached = true; break; }
        if (bb != ctx->block_bytes.end() && (!only_cached || cached)) {
            launch_fill_words(it->first, bb->second / sizeof(uint64_t), 0, ctx->stream);
            it = ctx->dirty.erase(it);
        } else {
            ++it;
        }
    }
}
static void cache_trim(qpzk_ctx* ctx) {
    if (ctx->cache.empty()) return;
    scrub_dirty(ctx, true);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->cache) { cudaFree(kv.second); ctx->block_bytes.erase(kv.second); ctx->owned_bytes -= kv.first; }
    ctx->cache.clear();
    ctx->cached_bytes = 0;
}
int dev_alloc(qpzk_ctx* ctx, uint64_t** p, uint64_t words) {
    *p = nullptr;
    if (!words) words = 1;
    uint64_t bytes = (words * sizeof(uint64_t) + 511) & ~511ull;
    auto it = ctx->cache.lower_bound(bytes);
    if (it != ctx->cache.end() && it->first <= bytes + bytes / 8) {
        *p = (uint64_t*)it->second;
        ctx->cached_bytes -= it->first;
        ctx->cache.erase(it);
        return 0;
    }
    void* q = nullptr;
    ctx->device_allocs++;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        cache_trim(ctx);
        e = cudaMalloc(&q, bytes);
    }
    CK(e);
    ctx->block_bytes[q] = bytes;
    ctx->owned_bytes += bytes;
    if (ctx->owned_bytes > ctx->peak_owned_bytes) ctx->peak_owned_bytes = ctx->owned_bytes;
    *p = (uint64_t*)q;
    return 0;
}
void dev_free(qpzk_ctx* ctx, void* p, uint64_t zero_words) {
    if (!p) return;
    auto dirty = ctx->dirty.find(p);
    if (dirty != ctx->dirty.end() && ctx->zeroize == 1) {
        // a block that still carries deferred (mode 2) contents and is released under mode 1: zero all of it now
        auto bb = ctx->block_bytes.find(p);
        if (bb != ctx->block_bytes.end()) launch_fill_words(p, bb->second / sizeof(uint64_t), 0, ctx->stream);
        ctx->dirty.erase(dirty);
    } else {
        scrub(ctx, p, zero_words);
    }
    auto it = ctx->block_bytes.find(p);
    if (it == ctx->block_bytes.end()) { cudaFree(p); return; }  // not ours (cannot happen)
    ctx->cache.emplace(it->second, p);
    ctx->cached_bytes += it->second;
}

// ---------------------------------------------------------------------------------------------
// root-power tables (built once per size on the host, at most a few thousand elements)
// ---------------------------------------------------------------------------------------------
static int upload_words(qpzk_ctx* ctx, const std::vector<uint64_t>& h, uint64_t** d) {
    CK(cudaMalloc((void**)d, h.size() * sizeof(uint64_t)));
    CK(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static void fill_powers(uint64_t base, uint64_t hi_size, std::vector<uint64_t>& lo, std::vector<uint64_t>& hi) {
    lo.resize(TW_LO_SIZE);
    hi.resize(hi_size);
    uint64_t v = 1;
    for (uint32_t i = 0; i < TW_LO_SIZE; i++) { lo[i] = v; v = gl_mul_host(v, base); }
    uint64_t step = v;  // base^TW_LO_SIZE
    v = 1;
    for (uint64_t i = 0; i < hi_size; i++) { hi[i] = v; v = gl_mul_host(v, step); }
}

int get_pow_tables(qpzk_ctx* ctx, uint32_t log_size, bool inverse, PowTables* out) {
    auto key = std::make_pair(log_size, inverse ? 1 : 0);
    auto it = ctx->pow_tables.find(key);
    if (it == ctx->pow_tables.end()) {
        uint64_t w = gl_root_of_unity_host(log_size);
        if (inverse) w = gl_inv_host(w);
        PowTables t;
        t.hi_size = log_size > TW_LO_BITS ? (1ull << (log_size - TW_LO_BITS)) : 1;
        std::vector<uint64_t> lo, hi;
        fill_powers(w, t.hi_size, lo, hi);
        int rc;
        if ((rc = upload_words(ctx, lo, &t.lo))) return rc;
        if ((rc = upload_words(ctx, hi, &t.hi))) return rc;
        it = ctx->pow_tables.emplace(key, t).first;
    }
    *out = it->second;
    return 0;
}

// coset j' (leaf block index) uses shift s = g * w_L^bitrev_r(j')
int get_coset_tables(qpzk_ctx* ctx, uint32_t log_n, uint32_t rate_bits, CosetTables* out) {
    auto key = std::make_pair(log_n, rate_bits);
    auto it = ctx->coset_tables.find(key);
    if (it == ctx->coset_tables.end()) {
        uint32_t cosets = 1u << rate_bits;
        CosetTables t;
        t.hi_size = log_n > TW_LO_BITS ? (1ull << (log_n - TW_LO_BITS)) : 1;
        uint64_t wL = gl_root_of_unity_host(log_n + rate_bits);
        std::vector<uint64_t> lo_all, hi_all, lo, hi;
        for (uint32_t jp = 0; jp < cosets; jp++) {
            uint32_t j = 0;
            for (uint32_t b = 0; b < rate_bits; b++) j |= ((jp >> b) & 1) << (rate_bits - 1 - b);
            uint64_t s = gl_mul_host(GL_GENERATOR, gl_exp_host(wL, j));
            fill_powers(s, t.hi_size, lo, hi);
            lo_all.insert(lo_all.end(), lo.begin(), lo.end());
            hi_all.insert(hi_all.end(), hi.begin(), hi.end());
        }
        int rc;
        if ((rc = upload_words(ctx, lo_all, &t.lo))) return rc;
        if ((rc = upload_words(ctx, hi_all, &t.hi))) return rc;
        it = ctx->coset_tables.emplace(key, t).first;
    }
    *out = it->second;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" const char* qpzk_last_error(void) { return g_err.c_str(); }
extern "C" const char* qpzk_version(void) { return "qpzk-b200 0.1 (sm_100a)"; }

static int ctx_create_body(int device, void* stream, bool own, qpzk_ctx** out);
static int ctx_create_impl(int device, void* stream, bool own, qpzk_ctx** out) {
    if (!out) return fail(QPZK_EINVAL, "out is null");
    *out = nullptr;
    qpzk_ctx* ctx = nullptr;
    const int rc = ctx_create_body(device, stream, own, &ctx);
    if (rc) { if (ctx) ctx_unref(ctx); return rc; }   // a failed step leaves nothing behind (stream, events, the struct)
    *out = ctx;
    return 0;
}
static int ctx_create_body(int device, void* stream, bool own, qpzk_ctx** out) {
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(QPZK_EINVAL, "no such CUDA device");
    {   // QPZK_BLOCKING_SYNC=1: host threads sleep in cudaStreamSynchronize instead of spinning (many prover threads
        // on few or quota-limited cores); must precede the creation of the device's primary context
        const char* e = getenv("QPZK_BLOCKING_SYNC");
        const char* l = getenv("QPZK_LMEM_MAX");
        unsigned flags = 0;
        if (e && e[0] == '1') flags |= cudaDeviceScheduleBlockingSync;
        if (l && l[0] == '1') flags |= cudaDeviceLmemResizeToMax;
        if (flags) { cudaSetDeviceFlags(flags); cudaGetLastError(); }
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(QPZK_ECUDA, "libqpzk is built for sm_100a only; device is older than Blackwell");
    qpzk_ctx* ctx = new qpzk_ctx();
    *out = ctx;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    { const char* e = getenv("QPZK_LEGACY_NTT"); ctx->legacy_ntt = e && e[0] == '1'; }
    { const char* e = getenv("QPZK_ZEROIZE"); if (e && e[0] >= '0' && e[0] <= '2') ctx->zeroize = e[0] - '0'; }   // dev aid; qpzk_ctx_set_zeroize is the API
    if (own) {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    } else {
        ctx->stream = (cudaStream_t)stream;
    }
    CK(upload_poseidon_constants());
    CK(upload_poseidon2_constants());
    for (int i = 0; i < 5; i++) CK(cudaEventCreate(&ctx->ev[i]));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->h2d_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->h2d_done2, cudaEventDisableTiming));
    return 0;
}

extern "C" int qpzk_ctx_create(int device, qpzk_ctx** out) {
    GUARD_BEGIN return ctx_create_impl(device, nullptr, true, out); GUARD_END
}
extern "C" int qpzk_ctx_create_on_stream(int device, void* cuda_stream, qpzk_ctx** out) {
    GUARD_BEGIN return ctx_create_impl(device, cuda_stream, false, out); GUARD_END
}
void ctx_unref(qpzk_ctx* ctx) {
    if (--ctx->refs > 0) return;
    cudaSetDevice(ctx->device);
    scrub_dirty(ctx, false);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->block_bytes) cudaFree(kv.first);   // cached blocks (every live handle has been freed by now)
    for (auto& kv : ctx->pow_tables) { cudaFree(kv.second.lo); cudaFree(kv.second.hi); }
    for (auto& kv : ctx->coset_tables) { cudaFree(kv.second.lo); cudaFree(kv.second.hi); }
    for (auto& kv : ctx->unshift_tables) { cudaFree(kv.second.lo); cudaFree(kv.second.hi); }
    for (auto& kv : ctx->zh_tables) cudaFree(kv.second);
    if (ctx->kpow) cudaFree(ctx->kpow);
    if (ctx->h2d_done) cudaEventDestroy(ctx->h2d_done);
    if (ctx->h2d_done2) cudaEventDestroy(ctx->h2d_done2);
    for (int i = 0; i < 5; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" void qpzk_ctx_destroy(qpzk_ctx* ctx) {
    if (ctx) ctx_unref(ctx);
}
extern "C" int qpzk_ctx_synchronize(qpzk_ctx* ctx) {
    if (!ctx) return fail(QPZK_EINVAL, "ctx is null");
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int qpzk_ctx_set_zeroize(qpzk_ctx* ctx, int mode) {
    if (!ctx) return fail(QPZK_EINVAL, "ctx is null");
    if (mode < 0 || mode > 2) return fail(QPZK_EINVAL, "zeroize mode must be 0, 1 or 2");
    // nothing stays deferred: cached blocks are zeroed now, blocks that live handles still use when they are released
    if (ctx->zeroize == 2 && mode != 2) { cudaSetDevice(ctx->device); scrub_dirty(ctx, true); }
    ctx->zeroize = mode;
    return 0;
}
extern "C" int qpzk_ctx_set_keep_values(qpzk_ctx* ctx, int on) { if (!ctx) return fail(QPZK_EINVAL, "ctx is null"); ctx->keep_values = on != 0; return 0; }
extern "C" int qpzk_ctx_set_timing(qpzk_ctx* ctx, int on) { if (!ctx) return fail(QPZK_EINVAL, "ctx is null"); ctx->timing = on != 0; return 0; }
extern "C" uint64_t qpzk_ctx_launch_count(const qpzk_ctx* ctx) { return ctx ? ctx->launches : 0; }
// Page-locked host buffers for callers that want the host<->device copies of the commit / proof calls to be truly
// asynchronous (a pageable source is staged by the driver and blocks the calling thread).
extern "C" void* qpzk_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void qpzk_pinned_free(void* p) { if (p) cudaFreeHost(p); }
extern "C" int qpzk_ctx_trim(qpzk_ctx* ctx) {
    if (!ctx) return fail(QPZK_EINVAL, "ctx is null");
    CK(cudaSetDevice(ctx->device));
    cache_trim(ctx);
    return 0;
}
extern "C" uint64_t qpzk_ctx_cached_bytes(const qpzk_ctx* ctx) { return ctx ? ctx->cached_bytes : 0; }
extern "C" uint64_t qpzk_ctx_device_allocs(const qpzk_ctx* ctx) { return ctx ? ctx->device_allocs : 0; }
extern "C" uint64_t qpzk_ctx_peak_bytes(const qpzk_ctx* ctx) { return ctx ? ctx->peak_owned_bytes : 0; }
extern "C" void qpzk_ctx_reset_peak(qpzk_ctx* ctx) { if (ctx) ctx->peak_owned_bytes = ctx->owned_bytes; }

static void mark(qpzk_ctx* ctx, int i) {
    if (ctx->timing) { cudaEventRecord(ctx->ev[i], ctx->stream); ctx->ev_valid[i] = true; }
}
extern "C" int qpzk_ctx_last_commit_timing(qpzk_ctx* ctx, float ms[4]) {
    if (!ctx || !ms) return fail(QPZK_EINVAL, "null argument");
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; i++) {
        ms[i] = 0;
        if (ctx->ev_valid[i] && ctx->ev_valid[i + 1]) CK(cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// commit pipeline
// ---------------------------------------------------------------------------------------------
static void split_bits(uint32_t total, uint32_t max_bits, std::vector<uint32_t>& out) {
    out.clear();
    if (!total) return;
    uint32_t k = (total + max_bits - 1) / max_bits;
    uint32_t base = total / k, extra = total % k;
    for (uint32_t i = 0; i < k; i++) out.push_back(base + (i < extra ? 1 : 0));
}

// Tile width of a strided pass: at least 8 words (64 B segments), and wide enough that a block works on about
// 4096 elements (a 2^6 x 8 tile keeps only a quarter of the 256 threads busy in the butterfly phase).
static uint32_t tile_width(uint32_t s, uint32_t room_bits) {
    uint32_t lw = 3;
    while (s + lw < 12) lw++;
    if (lw > room_bits) lw = room_bits;
    return 1u << lw;
}

// values (natural order, column-major) -> coefficients (natural order, column-major)
static int run_intt(qpzk_ctx* ctx, const uint64_t* values, uint64_t* coeffs, uint32_t log_n, uint32_t num_cols) {
    if (!num_cols) return 0;
    NvtxRange nvtx_intt("qpzk:ifft");
    if (num_cols > 65535) return fail(QPZK_EINVAL, "num_cols > 65535");
    PowTables inv;
    int rc = get_pow_tables(ctx, log_n, true, &inv);
    if (rc) return rc;
    uint64_t n = 1ull << log_n;
    uint64_t ninv = gl_inv_host(n % GL_P);
    std::vector<uint32_t> passes;
    if (log_n <= 12) passes.push_back(log_n); else split_bits(log_n, 10, passes);
    if (passes.empty()) passes.push_back(0);
    uint32_t b_lo = 0;
    for (size_t i = 0; i < passes.size(); i++) {
        NttPassArgs a{};
        a.in = i == 0 ? values : coeffs;
        a.out = coeffs;
        a.in_stride = a.out_stride = n;
        a.num_arrays = num_cols;
        a.log_n = log_n; a.b_lo = b_lo; a.s = passes[i];
        uint32_t room = i == 0 ? log_n - passes[i] : b_lo;   // bits available for the tile width
        a.W = passes.size() == 1 ? 1 : tile_width(passes[i], room);
        a.mode = i == 0 ? NTT_INV_FIRST : NTT_INV_GENERIC;
        a.num_cosets = 1;
        a.tw_lo = inv.lo; a.tw_hi = inv.hi;
        a.scale = i + 1 == passes.size() ? ninv : 1;
        CK(launch_ntt_pass(a, ctx->stream));
        ctx->launches++;
        b_lo += passes[i];
    }
    return 0;
}

// coefficients -> leaf blocks [coset_begin, coset_begin + coset_count) of the LDE, row-major, leaf order.
// The strided passes work column by column, so they can run on a column range as soon as those columns are on the
// device (the host-buffer commit overlaps them with the H2D copy of the next range); the last pass needs every
// column of its rows and runs once.
// The pass plan of one coset LDE: which position bits every strided pass takes, top down, and how many bits the last
// pass (the one that writes whole leaf rows) keeps.  A pure function of the shape (qpzk_lde_pass_plan exposes it for
// tests); QPZK_LDE_PLAN can force the strided part for experiments.
static void lde_pass_plan(uint32_t log_n, uint32_t num_cols, bool legacy_ntt, bool remote_first, std::vector<uint32_t>& passes,
                          uint32_t* s_last_out) {
    passes.clear();
    // rows per block of the last pass: all columns resident in ~40 KB of shared memory (4-5 blocks per SM)
    uint32_t s_last = 0;
    while (s_last < 10 && s_last < log_n && ((uint64_t)num_cols << (s_last + 1)) <= 4608) s_last++;
    // preferred: the 32-point in-register last pass (ntt_fwd.cu) whenever the polynomial is long enough
    const bool fast = !legacy_ntt && log_n > 5 && (size_t)num_cols * 33 * 8 <= 160 * 1024;
    if (fast) s_last = 5;
    *s_last_out = s_last;
    if (log_n <= s_last) return;
    // 5..8-bit passes run on the register-resident tile kernel (ntt_fwd.cu), which needs b_lo >= 12 - s; 1..4-bit
    // passes on its register-only sibling (b_lo >= 8); fewer than 7 bits in all go to the generic kernel.
    const uint32_t rem = log_n - s_last;
    // QPZK_LDE_PLAN="4,5,8": the strided passes in order (experiments; used when the parts sum to the bits that
    // remain above the last pass and every part is something a kernel takes at its position)
    std::vector<uint32_t> forced;
    if (const char* e = getenv("QPZK_LDE_PLAN")) {
        uint32_t sum = 0, hi = log_n;
        bool good = fast;
        for (const char* q = e; *q && good;) {
            const uint32_t v = (uint32_t)strtoul(q, nullptr, 10);
            const uint32_t lo = hi >= v ? hi - v : 0;
            good = v >= 1 && v <= 8 && v <= hi && (v <= 4 ? lo >= 8 : lo + v >= 12);
            forced.push_back(v); sum += v; hi = lo;
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        if (!good || sum != rem) forced.clear();
    }
    if (!forced.empty()) {
        passes = forced;
    } else if (fast && rem >= 7) {
        // As many 8-bit passes as fit, the pass next to the last one always among them (it needs b_lo >= 12 - s), and
        // what is left over as a FIRST pass: 1..4 bits on the register-only kernel (no exchange, one general
        // product per element), 5..7 on the tile kernel.  Every pass costs two general products per element plus
        // ~8 ALU instructions per bit, so fewer, wider passes win: measured at 2^22 x 135 (17 bits above the last
        // pass) 5+5+7: 111.0 ms, 4+5+8: 107.7, 3+6+8: 105.5, 1+8+8: 105.2 (profiles/r2_summary.md).
        // When the first pass pulls its input over NVLink it should be a wide one, so that its arithmetic covers the
        // transfer: the short pass goes second (2 GPUs, 2^22: 8+1+8 311.3 ms per commit, 1+8+8 315.0; on one GPU
        // 1+8+8 is the better order, 101.6 against 107.1 ms of LDE).
        if (remote_first && rem % 8 && rem / 8 >= 2) {
            passes.push_back(8);
            passes.push_back(rem % 8);
            for (uint32_t k = 1; k < rem / 8; k++) passes.push_back(8);
        } else {
            if (rem % 8) passes.push_back(rem % 8);
            for (uint32_t k = 0; k < rem / 8; k++) passes.push_back(8);
        }
    } else {
        split_bits(rem, 10, passes);
    }
}
extern "C" int qpzk_lde_pass_plan(uint32_t log_n, uint32_t num_cols, int peer_slices, uint32_t* passes_out, uint32_t* num_passes,
                                  uint32_t* last_pass_bits) {
    if (!passes_out || !num_passes || !last_pass_bits) return fail(QPZK_EINVAL, "null argument");
    if (log_n > 26 || !num_cols || num_cols > 65535) return fail(QPZK_EINVAL, "log_n > 26 or num_cols not in 1..65535");
    std::vector<uint32_t> passes;
    static const bool legacy = [] { const char* e = getenv("QPZK_LEGACY_NTT"); return e && e[0] == '1'; }();
    lde_pass_plan(log_n, num_cols, legacy, peer_slices != 0, passes, last_pass_bits);
    if (passes.size() > 8) return fail(QPZK_EINTERNAL, "more than 8 strided passes");
    for (size_t i = 0; i < passes.size(); i++) passes_out[i] = passes[i];
    *num_passes = (uint32_t)passes.size();
    return 0;
}

struct LdePlan {
    qpzk_ctx* ctx = nullptr;
    const uint64_t* coeffs = nullptr;
    uint64_t* leaves = nullptr;
    uint64_t* inter = nullptr;
    uint32_t log_n = 0, num_cols = 0, leaf_width = 0, coset_count = 0, s_last = 0;
    PowTables fwd;
    CosetTables cs;
    const uint64_t *cs_lo = nullptr, *cs_hi = nullptr;
    std::vector<uint32_t> passes;

    // remote_first: the first pass reads coefficient slices of peer GPUs over NVLink (coset-split commitment)
    int prepare(qpzk_ctx* c, const uint64_t* coeffs_, uint64_t* leaves_, uint32_t log_n_, uint32_t num_cols_,
                uint32_t leaf_width_, uint32_t rate_bits, uint32_t coset_begin, uint32_t coset_count_, bool remote_first = false) {
        ctx = c; coeffs = coeffs_; leaves = leaves_; log_n = log_n_; num_cols = num_cols_; leaf_width = leaf_width_;
        coset_count = coset_count_;
        if (!num_cols) return 0;
        int rc = get_pow_tables(ctx, log_n, false, &fwd);
        if (rc) return rc;
        if ((rc = get_coset_tables(ctx, log_n, rate_bits, &cs))) return rc;
        if (num_cols > 65535) return fail(QPZK_EINVAL, "num_cols > 65535");
        lde_pass_plan(log_n, num_cols, ctx->legacy_ntt, remote_first, passes, &s_last);
        cs_lo = cs.lo + (size_t)coset_begin * TW_LO_SIZE;
        cs_hi = cs.hi + (size_t)coset_begin * cs.hi_size;
        if (log_n > s_last) {
            if ((rc = dev_alloc(ctx, &inter, (uint64_t)coset_count * num_cols << log_n))) return rc;
        }
        return 0;
    }
    // strided passes for columns [c0, c1); first_src (optional) = where column c0 of the coefficients starts when the
    // range does not live in `coeffs` (a peer GPU's slice, read over NVLink by the first pass itself)
    // which = 0: every pass, 1: only the first pass (the one that reads the coefficients), 2: every pass but the first
    int run_passes(uint32_t c0, uint32_t c1, const uint64_t* first_src = nullptr, int which = 0) {
        if (passes.empty() || c1 <= c0) return 0;
        NvtxRange nvtx_lde("qpzk:lde passes");
        const uint64_t n = 1ull << log_n;
        uint32_t b_hi = log_n;
        for (size_t i = 0; i < passes.size(); i++) {
            if ((which == 1 && i > 0) || (which == 2 && i == 0)) { b_hi -= passes[i]; continue; }
            NttPassArgs a{};
            a.log_n = log_n; a.s = passes[i]; a.b_lo = b_hi - passes[i];
            a.W = tile_width(a.s, a.b_lo);
            a.in_stride = a.out_stride = n;
            a.tw_lo = fwd.lo; a.tw_hi = fwd.hi; a.scale = 1;
            a.num_arrays = c1 - c0;
            a.group_stride = (uint64_t)num_cols * n;
            a.out = inter + (uint64_t)c0 * n;
            if (i == 0) {
                a.mode = NTT_FWD_COSET_FIRST; a.in = first_src ? first_src : coeffs + (uint64_t)c0 * n; a.num_groups = 1;
                a.num_cosets = coset_count; a.cs_lo = cs_lo; a.cs_hi = cs_hi; a.cs_hi_size = cs.hi_size;
            } else {
                a.mode = NTT_FWD_GENERIC; a.in = a.out; a.num_groups = coset_count; a.num_cosets = 1;
            }
            if (!ctx->legacy_ntt && ntt_fwd_supported(a)) {
                const uint64_t tile_cosets = (n >> 12) * coset_count;
                uint64_t cpb = tile_cosets * (c1 - c0) / 4096;
                // a narrow column range (a peer's slice) still amortises the per-tile factor tables over as many columns
                // as it has, as long as the tiles alone fill the machine
                if (which == 1 && tile_cosets >= 8ull * ctx->num_sms) cpb = c1 - c0;
                a.cols_per_block = (uint32_t)(cpb < 1 ? 1 : cpb > 32 ? 32 : cpb);
                CK(launch_ntt_fwd(a, ctx->stream));
            } else {
                CK(launch_ntt_pass(a, ctx->stream));
            }
            ctx->launches++;
            b_hi = a.b_lo;
        }
        return 0;
    }
    int run_last() {
        if (!num_cols) return 0;
        NvtxRange nvtx_last("qpzk:lde last pass + transpose");
        LdeLastArgs la{};
        la.leaves = leaves; la.log_n = log_n; la.s = s_last; la.num_cols = num_cols; la.leaf_width = leaf_width;
        la.num_cosets = coset_count; la.tw_lo = fwd.lo; la.tw_hi = fwd.hi; la.cs_lo = cs_lo; la.cs_hi = cs_hi;
        la.cs_hi_size = cs.hi_size; la.in_stride = 1ull << log_n;
        la.in = passes.empty() ? coeffs : inter;
        la.from_coeffs = passes.empty() ? 1 : 0;
        if (!ctx->legacy_ntt && lde_last32_supported(la)) CK(launch_lde_last32(la, ctx->stream, ctx->num_sms));
        else CK(launch_lde_last(la, ctx->stream));
        ctx->launches++;
        return 0;
    }
    void release() {
        if (inter) dev_free(ctx, inter, (uint64_t)coset_count * num_cols << log_n);
        inter = nullptr;
    }
};

static int build_tree(qpzk_ctx* ctx, qpzk_batch* b) {
    NvtxRange nvtx_tree("qpzk:MerkleTree::new");
    uint64_t words = 8 * b->num_leaves;  // all levels fit in 8N words
    int rc = dev_alloc(ctx, &b->digests, words);
    if (rc) return rc;
    CK(launch_hash_rows(b->leaves, b->num_leaves, b->leaf_width, b->leaf_width, b->digests, ctx->stream));
    ctx->launches++;
    mark(ctx, 3);
    CK(launch_merkle_levels(b->digests, b->num_levels, b->num_leaves, ctx->stream));
    ctx->launches += merkle_levels_launches(b->num_levels, b->num_leaves);
    mark(ctx, 4);
    return 0;
}

void batch_release(qpzk_batch* b) {
    if (!b) return;
    qpzk_ctx* ctx = b->ctx;
    cudaSetDevice(ctx->device);
    uint64_t n = 1ull << b->log_n;
    dev_free(ctx, b->coeffs, (uint64_t)b->num_cols * n);
    dev_free(ctx, b->values, (uint64_t)b->num_cols * n);
    if (b->own_leaves) dev_free(ctx, b->leaves, b->num_leaves * b->leaf_width);
    dev_free(ctx, b->digests);
    delete b;
    ctx_unref(ctx);
}

// Where the input columns come from.  Device-resident input: `dev` is set and fetch is a no-op.  Host input: the
// columns are copied range by range on the context's copy stream and `fetch(c0, c1)` makes the context stream wait
// for the range, so the per-column stages (iNTT, strided LDE passes) of one range overlap the copy of the next.

static int fetch_columns(qpzk_ctx* ctx, const ColumnSource& src, uint64_t n, uint32_t c0, uint32_t c1) {
    if (!src.host || c1 <= c0) return 0;
    if (src.chunks <= 1) {   // one piece: nothing to overlap, so no second stream and no cross-stream waits
        CK(cudaMemcpyAsync(src.staging + (uint64_t)c0 * n, src.host + (uint64_t)c0 * n, (uint64_t)(c1 - c0) * n * sizeof(uint64_t),
                           cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->h2d_done, ctx->stream));
        return 0;
    }
    CK(cudaMemcpyAsync(src.staging + (uint64_t)c0 * n, src.host + (uint64_t)c0 * n, (uint64_t)(c1 - c0) * n * sizeof(uint64_t),
                       cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->h2d_done, ctx->copy_stream));
    CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
    return 0;
}

int commit_impl(qpzk_ctx* ctx, ColumnSource src, bool from_values, uint32_t log_n, uint32_t num_cols,
                       uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt_dev, uint32_t coset_begin,
                       uint32_t coset_count, bool keep_coeffs_copy, qpzk_batch** out) {
    if (src.slices && (from_values || keep_coeffs_copy || src.host || src.dev || !src.cols_per_slice))
        return fail(QPZK_EINVAL, "coefficient slices: from_coeffs without a copy only");
    NvtxRange nvtx_commit(from_values ? "qpzk:from_values" : "qpzk:from_coeffs");
    if (!ctx || !out) return fail(QPZK_EINVAL, "null argument");
    *out = nullptr;
    if (!src.dev && !src.host && !src.slices && num_cols) return fail(QPZK_EINVAL, "cols is null");
    if (log_n > 26 || rate_bits > 8) return fail(QPZK_EINVAL, "log_n > 26 or rate_bits > 8");
    uint32_t cosets = 1u << rate_bits;
    if (!coset_count || (coset_count & (coset_count - 1)) || coset_begin % coset_count || coset_begin + coset_count > cosets)
        return fail(QPZK_EINVAL, "bad coset range");
    uint32_t lost = rate_bits - ilog2(coset_count);  // tree levels above this slice
    uint32_t log_L = log_n + rate_bits;
    if (cap_height > log_L) return fail(QPZK_EINVAL, "cap_height exceeds log2(number of leaves)");
    if (cap_height < lost) return fail(QPZK_EINVAL, "cap_height too small for this coset split");
    if (salt_dev && coset_count != cosets) return fail(QPZK_EINVAL, "salt is not supported with a coset split");
    CK(cudaSetDevice(ctx->device));

    qpzk_batch* b = new qpzk_batch();
    ctx->refs++;
    b->ctx = ctx; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height - lost;
    b->num_cols = num_cols; b->salt_cols = salt_dev ? QPZK_SALT_SIZE : 0; b->leaf_width = num_cols + b->salt_cols;
    uint64_t n = 1ull << log_n;
    b->num_leaves = n * coset_count;
    b->log_leaves = log_n + ilog2(coset_count);
    b->num_levels = b->log_leaves - b->cap_height + 1;
    int rc;
    for (int i = 0; i < 5; i++) ctx->ev_valid[i] = false;
    mark(ctx, 0);
    const bool own_coeffs = from_values || keep_coeffs_copy || src.host;
    if (own_coeffs) {
        if ((rc = dev_alloc(ctx, &b->coeffs, (uint64_t)num_cols * n))) { batch_release(b); return rc; }
    }
    if (src.host) {
        // from_coeffs copies straight into the batch's coefficient buffer; from_values needs a staging buffer
        // (the first iNTT pass is out of place).  The copy stream must not overtake earlier users of the block.
        if (from_values) {
            if ((rc = dev_alloc(ctx, &src.staging, (uint64_t)num_cols * n))) { batch_release(b); return rc; }
        } else {
            src.staging = b->coeffs;
        }
        cudaError_t e = cudaSuccess;
        if (src.chunks > 1) {
            e = cudaEventRecord(ctx->copy_done, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done, 0);
        }
        if (e != cudaSuccess) { if (from_values) dev_free(ctx, src.staging); batch_release(b); CK(e); }
    }
    const uint64_t* in_dev = src.host ? src.staging : src.dev;
    const uint64_t* coeffs = own_coeffs ? b->coeffs : src.dev;
    auto fail_out = [&](int code) {
        if (src.host && from_values) dev_free(ctx, src.staging, (uint64_t)num_cols * n);
        batch_release(b);
        return code;
    };
    if (!from_values && keep_coeffs_copy && !src.host) {
        cudaError_t e = cudaMemcpyAsync(b->coeffs, src.dev, (uint64_t)num_cols * n * sizeof(uint64_t),
                                        cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { batch_release(b); CK(e); }
    }
    if ((rc = dev_alloc(ctx, &b->leaves, b->num_leaves * b->leaf_width))) return fail_out(rc);
    LdePlan lde;
    if ((rc = lde.prepare(ctx, coeffs, b->leaves, log_n, num_cols, b->leaf_width, rate_bits, coset_begin, coset_count, src.slices != nullptr))) {
        lde.release(); return fail_out(rc);
    }
    if (src.slices) {
        // The coefficient matrix is spread over several buffers of cols_per_slice columns each (one per GPU of a coset-
        // split commitment, peers mapped over NVLink): the first strided pass reads every slice where it lies, so the
        // exchange IS the first pass -- no all-gather, no assembled copy of the matrix.
        if (lde.passes.empty()) { lde.release(); return fail_out(fail(QPZK_EINVAL, "coefficient slices need log_n > 11")); }
        mark(ctx, 1);
        // Every rank starts with a DIFFERENT slice (its own) and walks the ring from there: at any moment each GPU's
        // memory serves one reader.  With all ranks starting at slice 0 the eight readers shared one GPU's NVLink
        // egress at a time (measured at 8 GPUs: 99.8 ms per commit against 85.8 ms with an NCCL all-gather).
        const uint32_t num_slices = (num_cols + src.cols_per_slice - 1) / src.cols_per_slice;
        const uint32_t first = (coset_begin / coset_count) % num_slices;
        for (uint32_t t = 0; t < num_slices; t++) {
            const uint32_t k = (first + t) % num_slices;
            const uint32_t c0 = k * src.cols_per_slice;
            const uint32_t c1 = c0 + src.cols_per_slice < num_cols ? c0 + src.cols_per_slice : num_cols;
            if (!src.slices[k]) { lde.release(); return fail_out(fail(QPZK_EINVAL, "null coefficient slice")); }
            if ((rc = lde.run_passes(c0, c1, src.slices[k], 1))) { lde.release(); return fail_out(rc); }
        }
        if ((rc = lde.run_passes(0, num_cols, nullptr, 2))) { lde.release(); return fail_out(rc); }
    }
    const uint32_t chunks = src.host ? (src.chunks ? src.chunks : 1) : 1;
    const uint32_t per = (num_cols + chunks - 1) / chunks;
    for (uint32_t c0 = 0; c0 < num_cols && !src.slices; c0 += per) {
        const uint32_t c1 = c0 + per < num_cols ? c0 + per : num_cols;
        if ((rc = fetch_columns(ctx, src, n, c0, c1))) { lde.release(); return fail_out(rc); }
        if (from_values && (rc = run_intt(ctx, in_dev + (uint64_t)c0 * n, b->coeffs + (uint64_t)c0 * n, log_n, c1 - c0))) {
            lde.release(); return fail_out(rc);
        }
        if (c0 == 0) mark(ctx, 1);  // with one chunk: end of the iNTT stage (chunked runs interleave the stages)
        if ((rc = lde.run_passes(c0, c1))) { lde.release(); return fail_out(rc); }
    }
    if (!num_cols) mark(ctx, 1);
    rc = lde.run_last();
    lde.release();
    const bool keep = from_values && ctx->keep_values && coset_count == cosets && num_cols;
    if (src.host && from_values) {
        if (keep && !rc) b->values = src.staging;      // the staged input stays with the batch (no copy)
        else dev_free(ctx, src.staging, (uint64_t)num_cols * n);
    } else if (keep && !rc) {
        if ((rc = dev_alloc(ctx, &b->values, (uint64_t)num_cols * n)) == 0) {
            cudaError_t e = cudaMemcpyAsync(b->values, src.dev, (uint64_t)num_cols * n * sizeof(uint64_t),
                                            cudaMemcpyDeviceToDevice, ctx->stream);
            if (e != cudaSuccess) { batch_release(b); CK(e); }
        }
    }
    if (rc) { batch_release(b); return rc; }
    if (salt_dev) {
        cudaError_t e = launch_salt_scatter(salt_dev, b->leaves, log_L, num_cols, b->leaf_width, b->salt_cols, ctx->stream);
        ctx->launches++;
        if (e != cudaSuccess) { batch_release(b); CK(e); }
    }
    mark(ctx, 2);
    if ((rc = build_tree(ctx, b))) { batch_release(b); return rc; }
    *out = b;
    return 0;
}


extern "C" int qpzk_commit_from_values_dev(qpzk_ctx* ctx, const uint64_t* cols, uint32_t log_n, uint32_t num_cols,
                                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, qpzk_batch** out) {
    GUARD_BEGIN return commit_impl(ctx, dev_source(cols), true, log_n, num_cols, rate_bits, cap_height, salt, 0, 1u << (rate_bits & 15), true, out); GUARD_END
}
extern "C" int qpzk_commit_from_coeffs_dev(qpzk_ctx* ctx, const uint64_t* cols, uint32_t log_n, uint32_t num_cols,
                                           uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, qpzk_batch** out) {
    GUARD_BEGIN return commit_impl(ctx, dev_source(cols), false, log_n, num_cols, rate_bits, cap_height, salt, 0, 1u << (rate_bits & 15), true, out); GUARD_END
}
extern "C" int qpzk_commit_cosets_from_coeffs_dev(qpzk_ctx* ctx, const uint64_t* coeffs, uint32_t log_n, uint32_t num_cols,
                                                  uint32_t rate_bits, uint32_t cap_height, uint32_t coset_begin,
                                                  uint32_t coset_count, qpzk_batch** out) {
    GUARD_BEGIN return commit_impl(ctx, dev_source(coeffs), false, log_n, num_cols, rate_bits, cap_height, nullptr, coset_begin, coset_count, false, out); GUARD_END
}
extern "C" int qpzk_commit_cosets_from_coeff_slices_dev(qpzk_ctx* ctx, const uint64_t* const* slices_dev, uint32_t num_slices,
                                                        uint32_t cols_per_slice, uint32_t log_n, uint32_t num_cols,
                                                        uint32_t rate_bits, uint32_t cap_height, uint32_t coset_begin,
                                                        uint32_t coset_count, qpzk_batch** out) {
    GUARD_BEGIN
    if (!slices_dev || !cols_per_slice || (uint64_t)num_slices * cols_per_slice < num_cols)
        return fail(QPZK_EINVAL, "coefficient slices do not cover the columns");
    ColumnSource src;
    src.slices = slices_dev; src.cols_per_slice = cols_per_slice;
    return commit_impl(ctx, src, false, log_n, num_cols, rate_bits, cap_height, nullptr, coset_begin, coset_count, false, out);
    GUARD_END
}

// Exchange buffers for the coset split: plain cudaMalloc blocks (an IPC handle names a whole allocation, so they do not
// come from the block cache), exported to / mapped from the other ranks' processes.
static_assert(sizeof(cudaIpcMemHandle_t) == QPZK_IPC_HANDLE_BYTES, "IPC handle size");
extern "C" int qpzk_peer_buffer_create(qpzk_ctx* ctx, uint64_t words, uint64_t** dev_ptr, uint8_t handle[QPZK_IPC_HANDLE_BYTES]) {
    GUARD_BEGIN
    if (!ctx || !dev_ptr || !handle || !words) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CK(cudaMalloc(&p, words * sizeof(uint64_t)));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); CK(e); }
    memcpy(handle, &h, sizeof h);
    *dev_ptr = (uint64_t*)p;
    return 0;
    GUARD_END
}
extern "C" int qpzk_peer_buffer_open(qpzk_ctx* ctx, const uint8_t handle[QPZK_IPC_HANDLE_BYTES], uint64_t** peer_ptr) {
    GUARD_BEGIN
    if (!ctx || !handle || !peer_ptr) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = (uint64_t*)p;
    return 0;
    GUARD_END
}
extern "C" int qpzk_peer_buffer_close(qpzk_ctx* ctx, uint64_t* peer_ptr) {
    GUARD_BEGIN
    if (!ctx || !peer_ptr) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
    GUARD_END
}
extern "C" int qpzk_peer_buffer_free(qpzk_ctx* ctx, uint64_t* dev_ptr) {
    GUARD_BEGIN
    if (!ctx || !dev_ptr) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(dev_ptr));
    return 0;
    GUARD_END
}

extern "C" int qpzk_intt_columns_dev(qpzk_ctx* ctx, const uint64_t* values, uint64_t* coeffs, uint32_t log_n, uint32_t num_cols) {
    GUARD_BEGIN
    if (!ctx || !values || !coeffs) return fail(QPZK_EINVAL, "null argument");
    if (log_n > 26) return fail(QPZK_EINVAL, "log_n > 26");
    CK(cudaSetDevice(ctx->device));
    return run_intt(ctx, values, coeffs, log_n, num_cols);
    GUARD_END
}

// Host-buffer commit: the input is copied in column ranges of >= 32 MB (at most 16) so that the copy of one range
// overlaps the iNTT and the strided LDE passes of the previous one; small inputs go in one piece.
static int commit_host(qpzk_ctx* ctx, const uint64_t* cols, bool from_values, uint32_t log_n, uint32_t num_cols,
                       uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, qpzk_batch** out) {
    if (!ctx || !out) return fail(QPZK_EINVAL, "null argument");
    if (log_n > 26 || rate_bits > 8) return fail(QPZK_EINVAL, "log_n > 26 or rate_bits > 8");
    if (!cols && num_cols) return fail(QPZK_EINVAL, "cols is null");
    CK(cudaSetDevice(ctx->device));
    uint64_t n = 1ull << log_n;
    uint64_t salt_words = salt ? (uint64_t)QPZK_SALT_SIZE * (n << rate_bits) : 0;
    uint64_t* d_salt = nullptr;
    int rc;
    if (salt_words) {
        if ((rc = dev_alloc(ctx, &d_salt, salt_words))) return rc;
        cudaError_t e = cudaMemcpyAsync(d_salt, salt, salt_words * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->h2d_done2, ctx->stream);
        if (e != cudaSuccess) { dev_free(ctx, d_salt); CK(e); }
    }
    ColumnSource src;
    src.host = cols;
    uint64_t bytes = (uint64_t)num_cols * n * sizeof(uint64_t);
    src.chunks = (uint32_t)(bytes >> 25);
    static const uint32_t max_chunks = [] { const char* e = getenv("QPZK_H2D_CHUNKS"); const int v = e ? atoi(e) : 16; return (uint32_t)(v < 1 ? 1 : v > 64 ? 64 : v); }();   // measured at 2^22 x 135: 1: 696.6 ms, 2: 655.4, 4: 638.2, 8: 629.2, 16: 626.4, 32: 634.2 (device resident: 613)
    if (src.chunks > max_chunks) src.chunks = max_chunks;
    if (src.chunks < 1) src.chunks = 1;
    if (!num_cols) { src.host = nullptr; src.dev = nullptr; }
    rc = commit_impl(ctx, src, from_values, log_n, num_cols, rate_bits, cap_height, d_salt, 0, 1u << rate_bits, true, out);
    dev_free(ctx, d_salt, salt_words);
    // The caller's buffers are "only read during the call" (qpzk.h): wait for the host->device reads -- and only for
    // them; the kernels behind them stay asynchronous -- so that a caller may reuse or zeroize its witness buffer
    // (reference wormhole/circuit/src/sensitive.rs) as soon as the call returns, pinned or not.
    if (src.host) { cudaError_t e = cudaEventSynchronize(ctx->h2d_done); if (!rc) CK(e); }
    if (salt_words) { cudaError_t e = cudaEventSynchronize(ctx->h2d_done2); if (!rc) CK(e); }
    return rc;
}
extern "C" int qpzk_commit_from_values(qpzk_ctx* ctx, const uint64_t* cols, uint32_t log_n, uint32_t num_cols,
                                       uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, qpzk_batch** out) {
    GUARD_BEGIN return commit_host(ctx, cols, true, log_n, num_cols, rate_bits, cap_height, salt, out); GUARD_END
}
extern "C" int qpzk_commit_from_coeffs(qpzk_ctx* ctx, const uint64_t* cols, uint32_t log_n, uint32_t num_cols,
                                       uint32_t rate_bits, uint32_t cap_height, const uint64_t* salt, qpzk_batch** out) {
    GUARD_BEGIN return commit_host(ctx, cols, false, log_n, num_cols, rate_bits, cap_height, salt, out); GUARD_END
}

// ---------------------------------------------------------------------------------------------
// MerkleTree::new
// ---------------------------------------------------------------------------------------------
static int merkle_impl(qpzk_ctx* ctx, uint64_t* leaves_dev_owned, const uint64_t* leaves_dev_borrowed, uint64_t num_leaves,
                       uint32_t leaf_len, uint32_t cap_height, qpzk_batch** out) {
    qpzk_batch* b = new qpzk_batch();
    ctx->refs++;
    b->ctx = ctx; b->cap_height = cap_height; b->leaf_width = leaf_len; b->num_cols = leaf_len;
    b->num_leaves = num_leaves; b->log_leaves = ilog2(num_leaves); b->num_levels = b->log_leaves - cap_height + 1;
    b->leaves = leaves_dev_owned ? leaves_dev_owned : const_cast<uint64_t*>(leaves_dev_borrowed);
    b->own_leaves = leaves_dev_owned != nullptr;
    b->num_cols = 0;  // no polynomials behind a bare tree
    for (int i = 0; i < 5; i++) ctx->ev_valid[i] = false;
    mark(ctx, 2);
    int rc = build_tree(ctx, b);
    if (rc) { batch_release(b); return rc; }
    *out = b;
    return 0;
}

static int merkle_check(qpzk_ctx* ctx, const void* leaves, uint64_t num_leaves, uint32_t cap_height, qpzk_batch** out) {
    if (!ctx || !out) return fail(QPZK_EINVAL, "null argument");
    *out = nullptr;
    if (!num_leaves || (num_leaves & (num_leaves - 1))) return fail(QPZK_EINVAL, "num_leaves must be a power of two");
    if (num_leaves > (1ull << 34)) return fail(QPZK_EINVAL, "too many leaves");
    if (cap_height > ilog2(num_leaves)) return fail(QPZK_EINVAL, "cap_height exceeds log2(number of leaves)");
    if (!leaves) return fail(QPZK_EINVAL, "leaves is null");
    return 0;
}

extern "C" int qpzk_merkle_new_dev(qpzk_ctx* ctx, const uint64_t* leaves, uint64_t num_leaves, uint32_t leaf_len,
                                   uint32_t cap_height, qpzk_batch** out) {
    GUARD_BEGIN
    int rc = merkle_check(ctx, leaves, num_leaves, cap_height, out);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    uint64_t* copy;
    if ((rc = dev_alloc(ctx, &copy, num_leaves * leaf_len))) return rc;
    CK(cudaMemcpyAsync(copy, leaves, num_leaves * leaf_len * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    return merkle_impl(ctx, copy, nullptr, num_leaves, leaf_len, cap_height, out);
    GUARD_END
}
extern "C" int qpzk_merkle_new(qpzk_ctx* ctx, const uint64_t* leaves, uint64_t num_leaves, uint32_t leaf_len,
                               uint32_t cap_height, qpzk_batch** out) {
    GUARD_BEGIN
    int rc = merkle_check(ctx, leaves, num_leaves, cap_height, out);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    uint64_t* copy;
    if ((rc = dev_alloc(ctx, &copy, num_leaves * leaf_len))) return rc;
    CK(cudaMemcpyAsync(copy, leaves, num_leaves * leaf_len * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->h2d_done, ctx->stream));
    rc = merkle_impl(ctx, copy, nullptr, num_leaves, leaf_len, cap_height, out);
    { cudaError_t e = cudaEventSynchronize(ctx->h2d_done); if (!rc) CK(e); }   // the caller's leaves have been read
    return rc;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// batch accessors
// ---------------------------------------------------------------------------------------------
extern "C" void qpzk_batch_free(qpzk_batch* b) { batch_release(b); }
extern "C" uint64_t qpzk_batch_num_leaves(const qpzk_batch* b) { return b ? b->num_leaves : 0; }
extern "C" uint32_t qpzk_batch_leaf_len(const qpzk_batch* b) { return b ? b->leaf_width : 0; }
extern "C" uint32_t qpzk_batch_cap_len(const qpzk_batch* b) { return b ? 1u << b->cap_height : 0; }
extern "C" const uint64_t* qpzk_batch_leaves_dev(const qpzk_batch* b) { return b ? b->leaves : nullptr; }
extern "C" const uint64_t* qpzk_batch_coeffs_dev(const qpzk_batch* b) { return b ? b->coeffs : nullptr; }
extern "C" const uint64_t* qpzk_batch_values_dev(const qpzk_batch* b) { return b ? b->values : nullptr; }

static int d2h(const qpzk_batch* b, uint64_t* dst, const uint64_t* src, uint64_t words) {
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaMemcpyAsync(dst, src, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, b->ctx->stream));
    CK(cudaStreamSynchronize(b->ctx->stream));
    return 0;
}

extern "C" int qpzk_batch_cap(const qpzk_batch* b, uint64_t* out) {
    GUARD_BEGIN
    if (!b || !out) return fail(QPZK_EINVAL, "null argument");
    return d2h(b, out, b->digests + digest_level_off(b->num_leaves, b->num_levels - 1), 4ull << b->cap_height);
    GUARD_END
}
extern "C" int qpzk_batch_coeffs(const qpzk_batch* b, uint64_t* out) {
    GUARD_BEGIN
    if (!b || !out) return fail(QPZK_EINVAL, "null argument");
    if (!b->coeffs) return fail(QPZK_EINVAL, "batch holds no polynomials");
    return d2h(b, out, b->coeffs, (uint64_t)b->num_cols << b->log_n);
    GUARD_END
}
extern "C" int qpzk_batch_rows(const qpzk_batch* b, uint64_t first_row, uint64_t num_rows, uint64_t* out) {
    GUARD_BEGIN
    if (!b || !out) return fail(QPZK_EINVAL, "null argument");
    if (first_row + num_rows > b->num_leaves) return fail(QPZK_EINVAL, "row range out of bounds");
    return d2h(b, out, b->leaves + first_row * b->leaf_width, num_rows * b->leaf_width);
    GUARD_END
}
extern "C" int qpzk_batch_digests(const qpzk_batch* b, uint32_t level, uint64_t* out) {
    GUARD_BEGIN
    if (!b || !out) return fail(QPZK_EINVAL, "null argument");
    if (level >= b->num_levels) return fail(QPZK_EINVAL, "level above the cap");
    return d2h(b, out, b->digests + digest_level_off(b->num_leaves, level), 4 * (b->num_leaves >> level));
    GUARD_END
}

extern "C" int qpzk_batch_open(const qpzk_batch* b, const uint64_t* idx, uint32_t k, uint64_t* rows, uint64_t* siblings) {
    GUARD_BEGIN
    if (!b || !idx || !rows || !siblings) return fail(QPZK_EINVAL, "null argument");
    qpzk_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    for (uint32_t i = 0; i < k; i++) if (idx[i] >= b->num_leaves) return fail(QPZK_EINVAL, "leaf index out of bounds");
    uint32_t ns = b->log_leaves - b->cap_height;
    uint64_t per = b->leaf_width + 4ull * ns;
    uint64_t *d_idx, *d_out;
    int rc;
    if ((rc = dev_alloc(ctx, &d_idx, k))) return rc;
    if ((rc = dev_alloc(ctx, &d_out, per * k))) { dev_free(ctx, d_idx); return rc; }
    CK(cudaMemcpyAsync(d_idx, idx, k * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    FriGatherArgs g{};
    g.trees[0] = FriTreeDesc{b->leaves, b->digests, b->leaf_width, b->log_leaves, b->cap_height, 0};
    g.num_trees = 1; g.num_oracles = 1; g.indices = d_idx; g.proof = d_out; g.query_base = 0; g.query_words = per;
    CK(launch_fri_gather(g, k, ctx->stream));
    ctx->launches++;
    std::vector<uint64_t> h(per * k);
    CK(cudaMemcpyAsync(h.data(), d_out, per * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < k; i++) {
        memcpy(rows + (uint64_t)i * b->leaf_width, h.data() + i * per, b->leaf_width * sizeof(uint64_t));
        memcpy(siblings + (uint64_t)i * 4 * ns, h.data() + i * per + b->leaf_width, 4ull * ns * sizeof(uint64_t));
    }
    dev_free(ctx, d_idx); dev_free(ctx, d_out);
    return 0;
    GUARD_END
}

extern "C" int qpzk_batch_eval_ext(const qpzk_batch* b, const uint64_t point[2], uint64_t* out) {
    GUARD_BEGIN
    if (!b || !point || !out) return fail(QPZK_EINVAL, "null argument");
    if (!b->coeffs) return fail(QPZK_EINVAL, "batch holds no polynomials");
    if (point[0] >= GL_P || point[1] >= GL_P) return fail(QPZK_EINVAL, "point is not canonical");
    qpzk_ctx* ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    uint64_t* d;
    int rc = dev_alloc(ctx, &d, 2ull * b->num_cols);
    if (rc) return rc;
    CK(launch_eval_ext(b->coeffs, 1ull << b->log_n, b->log_n, b->num_cols, point, d, ctx->stream));
    ctx->launches++;
    rc = d2h(b, out, d, 2ull * b->num_cols);
    dev_free(ctx, d);
    return rc;
    GUARD_END
}

extern "C" int qpzk_batches_eval_ext(qpzk_ctx* ctx, const qpzk_batch* const* batches, uint32_t num_batches,
                                     const uint64_t point[2], uint64_t* out) {
    GUARD_BEGIN
    if (!ctx || !batches || !point || !out) return fail(QPZK_EINVAL, "null argument");
    if (!num_batches || num_batches > MAX_EVAL_JOBS) return fail(QPZK_EINVAL, "num_batches must be in 1..64");
    if (point[0] >= GL_P || point[1] >= GL_P) return fail(QPZK_EINVAL, "point is not canonical");
    EvalMultiArgs ea{};
    uint64_t total = 0;
    for (uint32_t j = 0; j < num_batches; j++) {
        const qpzk_batch* b = batches[j];
        if (!b || b->ctx != ctx || !b->coeffs) return fail(QPZK_EINVAL, "batch holds no polynomials or belongs to another context");
        if (b->log_n != batches[0]->log_n) return fail(QPZK_EINVAL, "batches must have the same degree");
        total += b->num_cols;
    }
    CK(cudaSetDevice(ctx->device));
    uint64_t* d;
    int rc = dev_alloc(ctx, &d, 2 * total);
    if (rc) return rc;
    ea.num_jobs = num_batches; ea.log_n = batches[0]->log_n; ea.col_stride = 1ull << ea.log_n;
    for (uint32_t j = 0; j < num_batches; j++) {
        ea.coeffs[j] = batches[j]->coeffs;
        ea.out[j] = d + 2ull * ea.col_begin[j];
        ea.z[2 * j] = point[0]; ea.z[2 * j + 1] = point[1];
        ea.col_begin[j + 1] = ea.col_begin[j] + batches[j]->num_cols;
    }
    cudaError_t e = launch_eval_ext_multi(ea, ctx->stream);
    ctx->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 2 * total * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d);
    CK(e);
    return 0;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// PoseidonPermutation::permute (batch) and the linear-layer self-test hook
// ---------------------------------------------------------------------------------------------
static int permute_impl(qpzk_ctx* ctx, const uint64_t* states, uint64_t count, int row, uint64_t* out) {
    if (!ctx || (!states && count) || (!out && count)) return fail(QPZK_EINVAL, "null argument");
    if (count > (1ull << 32)) return fail(QPZK_EINVAL, "too many states");
    CK(cudaSetDevice(ctx->device));
    uint64_t *d_in, *d_out;
    int rc;
    if ((rc = dev_alloc(ctx, &d_in, 12 * count))) return rc;
    if ((rc = dev_alloc(ctx, &d_out, 12 * count))) { dev_free(ctx, d_in); return rc; }
    cudaError_t e = cudaMemcpyAsync(d_in, states, 12 * count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) { e = launch_permute(d_in, count, row, d_out, ctx->stream); ctx->launches++; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, 12 * count * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d_in, 12 * count); dev_free(ctx, d_out, 12 * count);
    CK(e);
    return 0;
}
extern "C" int qpzk_poseidon_permute(qpzk_ctx* ctx, const uint64_t* states, uint64_t count, uint64_t* out) {
    GUARD_BEGIN return permute_impl(ctx, states, count, -1, out); GUARD_END
}
extern "C" int qpzk_selftest_mds_layer(qpzk_ctx* ctx, const uint64_t* states, uint64_t count, uint32_t next_round, uint64_t* out) {
    GUARD_BEGIN
    if (next_round > 30) return fail(QPZK_EINVAL, "next_round must be in 0..30");
    return permute_impl(ctx, states, count, (int)next_round, out);
    GUARD_END
}

extern "C" int qpzk_selftest_field_ops(qpzk_ctx* ctx, uint32_t op, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                                       uint64_t count, uint64_t* out) {
    GUARD_BEGIN
    if (!ctx || ((!a || !b || !c || !out) && count)) return fail(QPZK_EINVAL, "null argument");
    if (op > 14) return fail(QPZK_EINVAL, "unknown operation");
    if (count > (1ull << 28)) return fail(QPZK_EINVAL, "too many operands");
    CK(cudaSetDevice(ctx->device));
    DevBufs bufs(ctx);
    uint64_t* d[4];
    int rc;
    for (int k = 0; k < 4; k++) if ((rc = bufs.alloc(&d[k], count ? count : 1))) return rc;
    const uint64_t* src[3] = {a, b, c};
    for (int k = 0; k < 3 && count; k++) CK(cudaMemcpyAsync(d[k], src[k], count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_field_ops(op, d[0], d[1], d[2], count, d[3], ctx->stream)); ctx->launches++;
    if (count) CK(cudaMemcpyAsync(out, d[3], count * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// Poseidon2: the reference's application hash (poseidon2.cu)
// ---------------------------------------------------------------------------------------------
extern "C" int qpzk_poseidon2_permute(qpzk_ctx* ctx, const uint64_t* states, uint64_t count, uint64_t* out) {
    GUARD_BEGIN
    if (!ctx || (!states && count) || (!out && count)) return fail(QPZK_EINVAL, "null argument");
    if (count > (1ull << 32)) return fail(QPZK_EINVAL, "too many states");
    CK(cudaSetDevice(ctx->device));
    DevBufs bufs(ctx);
    uint64_t *d_in, *d_out;
    int rc;
    if ((rc = bufs.alloc(&d_in, 12 * count, true))) return rc;
    if ((rc = bufs.alloc(&d_out, 12 * count, true))) return rc;
    CK(cudaMemcpyAsync(d_in, states, 12 * count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_poseidon2_permute(d_in, count, d_out, ctx->stream)); ctx->launches++;
    CK(cudaMemcpyAsync(out, d_out, 12 * count * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
    GUARD_END
}
extern "C" int qpzk_poseidon2_hash_no_pad_dev(qpzk_ctx* ctx, const uint64_t* msgs_dev, uint64_t count, uint32_t len,
                                              uint64_t* out_dev) {
    GUARD_BEGIN
    if (!ctx || (!msgs_dev && count && len) || (!out_dev && count)) return fail(QPZK_EINVAL, "null argument");
    if (len > (1u << 30)) return fail(QPZK_EINVAL, "message too long");
    CK(cudaSetDevice(ctx->device));
    CK(launch_poseidon2_hash(msgs_dev, count, len, len, out_dev, ctx->stream)); ctx->launches++;
    return 0;
    GUARD_END
}
extern "C" int qpzk_poseidon2_hash_no_pad(qpzk_ctx* ctx, const uint64_t* msgs, uint64_t count, uint32_t len, uint64_t* out) {
    GUARD_BEGIN
    if (!ctx || (!msgs && count && len) || (!out && count)) return fail(QPZK_EINVAL, "null argument");
    if (len > (1u << 30) || count > (1ull << 32)) return fail(QPZK_EINVAL, "too many / too long messages");
    for (uint64_t i = 0; i < count * len; i++) if (msgs[i] >= GL_P) return fail(QPZK_EINVAL, "message word is not canonical");
    CK(cudaSetDevice(ctx->device));
    DevBufs bufs(ctx);
    uint64_t *d_in, *d_out;
    int rc;
    if ((rc = bufs.alloc(&d_in, count * len, true))) return rc;    // preimages may be secret (addresses, nullifiers)
    if ((rc = bufs.alloc(&d_out, 4 * count))) return rc;
    if (count * len) CK(cudaMemcpyAsync(d_in, msgs, count * len * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_poseidon2_hash(d_in, count, len, len, d_out, ctx->stream)); ctx->launches++;
    if (count) CK(cudaMemcpyAsync(out, d_out, 4 * count * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// Challenger
// ---------------------------------------------------------------------------------------------
extern "C" int qpzk_challenger_create(qpzk_ctx* ctx, qpzk_challenger** out) {
    GUARD_BEGIN
    if (!ctx || !out) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    // state and staging come from the context's block cache: cudaMalloc / cudaFree synchronise the whole device and
    // would serialise proofs that run concurrently on other contexts (one Challenger per proof)
    uint64_t *dev = nullptr, *scratch = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &dev, (sizeof(ChallengerDev) + 7) / 8))) return rc;
    if ((rc = dev_alloc(ctx, &scratch, 4096))) { dev_free(ctx, dev); return rc; }
    qpzk_challenger* ch = new qpzk_challenger();
    ctx->refs++;
    ch->ctx = ctx;
    ch->scratch_words = 4096;
    ch->dev = reinterpret_cast<ChallengerDev*>(dev);
    ch->scratch = scratch;
    static_assert(sizeof(ChallengerDev) % 8 == 0, "challenger state is a whole number of words");
    CK(launch_fill_words(ch->dev, sizeof(ChallengerDev) / 8, 0, ctx->stream));
    *out = ch;
    return 0;
    GUARD_END
}
extern "C" void qpzk_challenger_free(qpzk_challenger* ch) {
    if (!ch) return;
    cudaSetDevice(ch->ctx->device);
    qpzk_ctx* ctx = ch->ctx;
    dev_free(ctx, ch->dev, (sizeof(ChallengerDev) + 7) / 8);   // stream ordered: later users of the block run after us
    dev_free(ctx, ch->scratch);
    delete ch;
    ctx_unref(ctx);
}
extern "C" int qpzk_challenger_observe(qpzk_challenger* ch, const uint64_t* elems, uint32_t n) {
    GUARD_BEGIN
    if (!ch || (!elems && n)) return fail(QPZK_EINVAL, "null argument");
    qpzk_ctx* ctx = ch->ctx;
    CK(cudaSetDevice(ctx->device));
    for (uint32_t i = 0; i < n; i++) if (elems[i] >= GL_P) return fail(QPZK_EINVAL, "observed element is not canonical");
    for (uint32_t off = 0; off < n; off += ch->scratch_words) {
        uint32_t m = n - off < ch->scratch_words ? n - off : ch->scratch_words;
        CK(cudaMemcpyAsync(ch->scratch, elems + off, m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(launch_challenger_observe(ch->dev, ch->scratch, m, ctx->stream));
        ctx->launches++;
        CK(cudaStreamSynchronize(ctx->stream));  // scratch and the (pageable) source are reused
    }
    return 0;
    GUARD_END
}
extern "C" int qpzk_challenger_observe_cap(qpzk_challenger* ch, const qpzk_batch* b) {
    GUARD_BEGIN
    if (!ch || !b) return fail(QPZK_EINVAL, "null argument");
    if (ch->ctx != b->ctx) return fail(QPZK_EINVAL, "challenger and batch belong to different contexts");
    CK(cudaSetDevice(ch->ctx->device));
    CK(launch_challenger_observe(ch->dev, b->digests + digest_level_off(b->num_leaves, b->num_levels - 1),
                                 4u << b->cap_height, ch->ctx->stream));
    ch->ctx->launches++;
    return 0;
    GUARD_END
}
extern "C" int qpzk_challenger_get(qpzk_challenger* ch, uint64_t* out, uint32_t n) {
    GUARD_BEGIN
    if (!ch || (!out && n)) return fail(QPZK_EINVAL, "null argument");
    if (n > ch->scratch_words) return fail(QPZK_EINVAL, "too many challenges in one call");
    qpzk_ctx* ctx = ch->ctx;
    CK(cudaSetDevice(ctx->device));
    CK(launch_challenger_get(ch->dev, ch->scratch, n, ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ch->scratch, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
    GUARD_END
}
static_assert(sizeof(qpzk_challenger_state) == sizeof(ChallengerDev), "challenger state layouts must match");
extern "C" int qpzk_challenger_export(qpzk_challenger* ch, qpzk_challenger_state* out) {
    GUARD_BEGIN
    if (!ch || !out) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ch->ctx->device));
    CK(cudaMemcpyAsync(out, ch->dev, sizeof(ChallengerDev), cudaMemcpyDeviceToHost, ch->ctx->stream));
    CK(cudaStreamSynchronize(ch->ctx->stream));
    return 0;
    GUARD_END
}
extern "C" int qpzk_challenger_import(qpzk_challenger* ch, const qpzk_challenger_state* in) {
    GUARD_BEGIN
    if (!ch || !in) return fail(QPZK_EINVAL, "null argument");
    if (in->input_len >= 8 || in->output_len > 8) return fail(QPZK_EINVAL, "bad challenger buffer lengths");
    CK(cudaSetDevice(ch->ctx->device));
    CK(cudaMemcpyAsync(ch->dev, in, sizeof(ChallengerDev), cudaMemcpyHostToDevice, ch->ctx->stream));
    CK(cudaStreamSynchronize(ch->ctx->stream));
    return 0;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// FRI
// ---------------------------------------------------------------------------------------------
extern "C" int qpzk_fri_params_constant_arity(qpzk_fri_params* p, uint32_t log_n, uint32_t rate_bits, uint32_t cap_height,
                                              uint32_t pow_bits, uint32_t num_queries, uint32_t arity_bits,
                                              uint32_t final_poly_bits) {
    if (!p) return fail(QPZK_EINVAL, "null argument");
    if (arity_bits < 1 || arity_bits > 4) return fail(QPZK_EINVAL, "arity_bits must be in 1..4");
    memset(p, 0, sizeof *p);
    p->rate_bits = rate_bits; p->cap_height = cap_height; p->proof_of_work_bits = pow_bits; p->num_query_rounds = num_queries;
    uint32_t d = log_n;
    while (d > final_poly_bits && d + rate_bits - arity_bits >= cap_height) {
        if (p->num_rounds == QPZK_MAX_FRI_ROUNDS)    // upstream's ConstantArityBits would keep reducing: refuse, do not truncate
            return fail(QPZK_EINVAL, "reduction needs more than QPZK_MAX_FRI_ROUNDS rounds");
        p->arity_bits[p->num_rounds++] = arity_bits;
        d -= arity_bits;
    }
    return 0;
}

extern "C" size_t qpzk_fri_proof_words(const qpzk_fri_params* p, uint32_t log_n, uint32_t num_oracles, const uint32_t* leaf_lens) {
    uint32_t log_L = log_n + p->rate_bits;
    size_t w = (size_t)p->num_rounds * ((size_t)4 << p->cap_height);
    size_t per_q = 0;
    for (uint32_t o = 0; o < num_oracles; o++) per_q += leaf_lens[o] + 4 * (size_t)(log_L - p->cap_height);
    uint32_t lg = log_L, deg = log_n;
    for (uint32_t r = 0; r < p->num_rounds; r++) {
        lg -= p->arity_bits[r]; deg -= p->arity_bits[r];
        per_q += ((size_t)2 << p->arity_bits[r]) + 4 * (size_t)(lg - p->cap_height);
    }
    return w + per_q * p->num_query_rounds + ((size_t)2 << deg) + 1;
}


// prove_openings on the context stream, without any host synchronisation: everything it needs from the host is
// either a kernel argument or (FriDeviceIO) already on the device, so the single-call prover (prove.cu) can capture it
// in a CUDA graph.  Results land in device memory: the flat proof in io.proof, the opened values (instance order) in
// io.openings, the PoW status word in io.status[0] (1 ok, 0 supplied witness too weak, 2 search exhausted).
int fri_core(qpzk_ctx* ctx, const qpzk_batch* const* oracles, uint32_t num_oracles, const qpzk_fri_instance* inst,
             const qpzk_fri_params* p, ChallengerDev* ch, uint64_t pow_witness_in, const FriDeviceIO& io) {
    NvtxRange nvtx_fri("qpzk:prove_openings");
    if (!num_oracles || num_oracles > MAX_ORACLES) return fail(QPZK_EINVAL, "num_oracles must be in 1..8");
    if (!inst->num_batches || inst->num_batches > MAX_FRI_BATCHES) return fail(QPZK_EINVAL, "num_batches must be in 1..8");
    if (p->num_rounds > MAX_FRI_ROUNDS) return fail(QPZK_EINVAL, "too many FRI rounds");
    const uint32_t log_n = oracles[0]->log_n, rate_bits = p->rate_bits, log_L = log_n + rate_bits;
    for (uint32_t o = 0; o < num_oracles; o++) {
        const qpzk_batch* b = oracles[o];
        if (!b || b->ctx != ctx) return fail(QPZK_EINVAL, "oracle belongs to another context");
        if (!b->coeffs || b->log_n != log_n || b->rate_bits != rate_bits || b->num_leaves != (1ull << log_L))
            return fail(QPZK_EINVAL, "oracles must be full polynomial batches of equal degree and rate");
        if (b->cap_height != p->cap_height) return fail(QPZK_EINVAL, "oracle cap_height differs from fri params");
    }
    uint32_t total_arity = 0;
    for (uint32_t r = 0; r < p->num_rounds; r++) {
        if (p->arity_bits[r] < 1 || p->arity_bits[r] > 4) return fail(QPZK_EINVAL, "arity_bits must be in 1..4");
        total_arity += p->arity_bits[r];
        if (total_arity > log_n) return fail(QPZK_EINVAL, "reduction arities exceed the degree");
        if (log_L - total_arity < p->cap_height) return fail(QPZK_EINVAL, "cap_height exceeds a FRI layer's tree height");
    }
    if (pow_witness_in != QPZK_POW_SEARCH && pow_witness_in >= GL_P) return fail(QPZK_EINVAL, "pow witness is not canonical");
    if (p->proof_of_work_bits > 40) return fail(QPZK_EINVAL, "proof_of_work_bits too large");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    // ---- instance
    uint32_t total_polys = 0, max_k = 0;
    FriCombineArgs ca{};
    for (uint32_t b = 0; b < inst->num_batches; b++) {
        ca.batch_off[b] = total_polys;
        total_polys += inst->batch_sizes[b];
        if (inst->batch_sizes[b] > max_k) max_k = inst->batch_sizes[b];
        if (!io.points) {
            ca.points[2 * b] = inst->points[2 * b]; ca.points[2 * b + 1] = inst->points[2 * b + 1];
            if (ca.points[2 * b] >= GL_P || ca.points[2 * b + 1] >= GL_P) return fail(QPZK_EINVAL, "opening point is not canonical");
        }
    }
    ca.points_dev = io.points;
    ca.batch_off[inst->num_batches] = total_polys;
    std::vector<bool> used(inst->num_batches * num_oracles, false);
    {
        uint32_t k = 0;
        for (uint32_t b = 0; b < inst->num_batches; b++)
            for (uint32_t j = 0; j < inst->batch_sizes[b]; j++, k++) {
                uint32_t o = inst->poly_oracle[k];
                if (o >= num_oracles || inst->poly_index[k] >= oracles[o]->num_cols) return fail(QPZK_EINVAL, "polynomial reference out of range");
                used[b * num_oracles + o] = true;
            }
    }
    DevBufs bufs(ctx);
    int rc;
    uint64_t *d_refs = nullptr, *d_alpha, *d_apow, *d_evals, *d_reduced, *d_found;
    const uint32_t* d_po = io.poly_refs;
    if (!d_po) {   // the polynomial references come from caller memory: the wrapper synchronises before it returns
        if ((rc = bufs.alloc(&d_refs, total_polys + 1))) return rc;  // two u32 arrays packed
        uint32_t* w = reinterpret_cast<uint32_t*>(d_refs);
        CK(cudaMemcpyAsync(w, inst->poly_oracle, total_polys * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(w + total_polys, inst->poly_index, total_polys * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        d_po = w;
    }
    const uint32_t* d_pi = d_po + total_polys;
    if ((rc = bufs.alloc(&d_alpha, 8))) return rc;             // alpha, then beta per round reuses +2
    if ((rc = bufs.alloc(&d_apow, 2ull * (max_k + 1)))) return rc;
    uint64_t eval_words = 0;
    for (uint32_t b = 0; b < inst->num_batches; b++)
        for (uint32_t o = 0; o < num_oracles; o++) {
            ca.eval_off[b * num_oracles + o] = (uint32_t)eval_words;
            if (used[b * num_oracles + o]) eval_words += oracles[o]->num_cols;
        }
    if ((rc = bufs.alloc(&d_evals, 2 * eval_words, true))) return rc;
    if ((rc = bufs.alloc(&d_reduced, 2ull * inst->num_batches))) return rc;
    if ((rc = bufs.alloc(&d_found, 2))) return rc;
    uint64_t* d_open = io.openings;

    // ---- openings of every referenced oracle at every batch point; the prover observes them before alpha is drawn
    {
        EvalMultiArgs ea{};
        ea.log_n = log_n; ea.col_stride = 1ull << log_n;
        for (uint32_t b = 0; b < inst->num_batches; b++)
            for (uint32_t o = 0; o < num_oracles; o++)
                if (used[b * num_oracles + o]) {
                    const uint32_t j = ea.num_jobs++;
                    ea.coeffs[j] = oracles[o]->coeffs;
                    ea.out[j] = d_evals + 2ull * ca.eval_off[b * num_oracles + o];
                    ea.z[2 * j] = ca.points[2 * b]; ea.z[2 * j + 1] = ca.points[2 * b + 1];
                    ea.z_dev[j] = io.points ? io.points + 2 * b : nullptr;
                    ea.col_begin[j + 1] = ea.col_begin[j] + oracles[o]->num_cols;
                }
        CK(launch_eval_ext_multi(ea, st));
        ctx->launches++;
    }
    ca.log_n = log_n; ca.rate_bits = rate_bits; ca.num_batches = inst->num_batches; ca.num_oracles = num_oracles;
    for (uint32_t o = 0; o < num_oracles; o++) { ca.leaves[o] = oracles[o]->leaves; ca.leaf_width[o] = oracles[o]->leaf_width; }
    ca.poly_oracle = d_po; ca.poly_index = d_pi; ca.apow = d_apow; ca.evals = d_evals; ca.reduced = d_reduced;
    ca.openings_out = d_open;
    if (io.observe_openings) {
        CK(launch_openings_gather(ca, total_polys, st)); ctx->launches++;
        CK(launch_challenger_observe(ch, d_open, 2 * total_polys, st)); ctx->launches++;
    }
    CK(launch_challenger_get(ch, d_alpha, 2, st)); ctx->launches++;
    CK(launch_alpha_powers(d_alpha, d_apow, max_k + 1, st)); ctx->launches++;

    // ---- batch-opening polynomial, evaluated on the LDE domain in leaf order
    PowTables fwd_n, inv_L;
    CosetTables cs;
    if ((rc = get_pow_tables(ctx, log_n, false, &fwd_n))) return rc;
    if ((rc = get_pow_tables(ctx, log_L, true, &inv_L))) return rc;
    if ((rc = get_coset_tables(ctx, log_n, rate_bits, &cs))) return rc;
    const uint64_t L = 1ull << log_L;
    std::vector<uint64_t*> layer(p->num_rounds + 1);
    if ((rc = bufs.alloc(&layer[0], 2 * L, true))) return rc;
    ca.values = layer[0]; ca.tw_lo = fwd_n.lo; ca.tw_hi = fwd_n.hi; ca.cs_lo = cs.lo;
    CK(launch_fri_combine(ca, st)); ctx->launches += 2;

    // ---- proof buffer layout
    std::vector<uint32_t> leaf_lens(num_oracles);
    for (uint32_t o = 0; o < num_oracles; o++) leaf_lens[o] = oracles[o]->leaf_width;
    const size_t proof_words = qpzk_fri_proof_words(p, log_n, num_oracles, leaf_lens.data());
    const size_t cap_words = (size_t)4 << p->cap_height;
    const size_t final_len = (size_t)1 << (log_n - total_arity);
    const size_t query_base = p->num_rounds * cap_words;
    const size_t query_words = p->num_query_rounds ? (proof_words - 1 - 2 * final_len - query_base) / p->num_query_rounds : 0;
    uint64_t* d_proof = io.proof;

    // ---- commit phase
    std::vector<uint64_t*> round_digests(p->num_rounds);
    uint32_t lg = log_L;
    uint64_t shift = GL_GENERATOR;
    uint64_t* d_beta = d_alpha + 2;
    for (uint32_t r = 0; r < p->num_rounds; r++) {
        const uint32_t ab = p->arity_bits[r];
        const uint64_t rows = 1ull << (lg - ab);
        if ((rc = bufs.alloc(&round_digests[r], 8 * rows, true))) return rc;
        CK(launch_hash_rows(layer[r], rows, 2u << ab, 2u << ab, round_digests[r], st)); ctx->launches++;
        const uint32_t levels = (lg - ab) - p->cap_height + 1;
        CK(launch_merkle_levels(round_digests[r], levels, rows, st)); ctx->launches += merkle_levels_launches(levels, rows);
        const uint64_t* cap = round_digests[r] + digest_level_off(rows, levels - 1);
        CK(cudaMemcpyAsync(d_proof + r * cap_words, cap, cap_words * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        CK(launch_challenger_observe(ch, cap, (uint32_t)cap_words, st)); ctx->launches++;
        CK(launch_challenger_get(ch, d_beta, 2, st)); ctx->launches++;
        if ((rc = bufs.alloc(&layer[r + 1], 2 * rows, true))) return rc;
        FriFoldArgs fa{};
        fa.in = layer[r]; fa.out = layer[r + 1]; fa.lg = lg; fa.log_L = log_L; fa.beta = d_beta;
        fa.shift_inv = gl_inv_host(shift);
        fa.arity_inv = gl_inv_host(1ull << ab);
        uint64_t wa_inv = gl_inv_host(gl_root_of_unity_host(ab));
        for (uint32_t j = 0; j < 8; j++) fa.itw[j] = gl_exp_host(wa_inv, j);
        fa.itw_lo = inv_L.lo; fa.itw_hi = inv_L.hi;
        CK(launch_fri_fold(fa, ab, st)); ctx->launches++;
        shift = gl_exp_host(shift, 1ull << ab);
        lg -= ab;
    }
    // ---- final polynomial: coefficients of the last layer, truncated; observed by the challenger
    uint64_t* d_final = d_proof + proof_words - 1 - 2 * final_len;
    {
        FriFinalArgs fa{};
        fa.in = layer[p->num_rounds]; fa.out = d_final; fa.lg = lg; fa.log_L = log_L; fa.rate_bits = rate_bits;
        fa.shift_inv = gl_inv_host(shift); fa.len_inv = gl_inv_host(1ull << lg);
        fa.itw_lo = inv_L.lo; fa.itw_hi = inv_L.hi;
        cudaError_t e = launch_fri_final_poly(fa, st);
        if (e == cudaErrorInvalidValue) return fail(QPZK_EINVAL, "final FRI layer too large (reduce more rounds)");
        CK(e); ctx->launches++;
        CK(launch_challenger_observe(ch, d_final, (uint32_t)(2 * final_len), st)); ctx->launches++;
    }
    // ---- proof of work (minimum witness, one launch), query indices, query phase
    FriGatherArgs g{};
    for (uint32_t o = 0; o < num_oracles; o++)
        g.trees[o] = FriTreeDesc{oracles[o]->leaves, oracles[o]->digests, oracles[o]->leaf_width, log_L, p->cap_height, 0};
    lg = log_L;
    for (uint32_t r = 0; r < p->num_rounds; r++) {
        lg -= p->arity_bits[r];
        g.trees[num_oracles + r] = FriTreeDesc{layer[r], round_digests[r], 2u << p->arity_bits[r], lg, p->cap_height, p->arity_bits[r]};
    }
    g.num_trees = num_oracles + p->num_rounds; g.num_oracles = num_oracles; g.indices = io.indices; g.status = io.status;
    g.proof = d_proof; g.query_base = query_base; g.query_words = query_words;
    // the search gives up after 2^(bits + 10) candidates: without a hit with probability e^-1024
    const uint64_t limit = p->proof_of_work_bits + 10 < 63 ? (1ull << (p->proof_of_work_bits + 10)) : GL_P;
    CK(launch_pow(ch, p->proof_of_work_bits, pow_witness_in, p->num_query_rounds, L, limit, ctx->num_sms,
                  reinterpret_cast<unsigned long long*>(d_found), d_proof + proof_words - 1, io.indices, io.status, st));
    ctx->launches += pow_witness_in == QPZK_POW_SEARCH ? 3 : 1;
    CK(launch_fri_gather(g, p->num_query_rounds, st)); ctx->launches++;
    return 0;
}

extern "C" int qpzk_fri_prove(qpzk_ctx* ctx, const qpzk_batch* const* oracles, uint32_t num_oracles,
                              const qpzk_fri_instance* inst, const qpzk_fri_params* p, qpzk_challenger* ch,
                              uint64_t pow_witness_in, uint64_t* proof_out, uint64_t* openings_out) {
    GUARD_BEGIN
    if (!ctx || !oracles || !inst || !p || !ch || !proof_out) return fail(QPZK_EINVAL, "null argument");
    if (ch->ctx != ctx) return fail(QPZK_EINVAL, "challenger belongs to another context");
    if (!num_oracles || num_oracles > MAX_ORACLES || !oracles[0]) return fail(QPZK_EINVAL, "num_oracles must be in 1..8");
    if (!inst->num_batches || inst->num_batches > MAX_FRI_BATCHES) return fail(QPZK_EINVAL, "num_batches must be in 1..8");
    CK(cudaSetDevice(ctx->device));
    uint32_t total_polys = 0;
    for (uint32_t b = 0; b < inst->num_batches; b++) total_polys += inst->batch_sizes[b];
    std::vector<uint32_t> leaf_lens(num_oracles);
    for (uint32_t o = 0; o < num_oracles; o++) { if (!oracles[o]) return fail(QPZK_EINVAL, "null oracle"); leaf_lens[o] = oracles[o]->leaf_width; }
    const size_t proof_words = qpzk_fri_proof_words(p, oracles[0]->log_n, num_oracles, leaf_lens.data());
    DevBufs bufs(ctx);
    int rc;
    uint64_t *d_proof, *d_open, *d_misc;
    if ((rc = bufs.alloc(&d_proof, proof_words, true))) return rc;
    if ((rc = bufs.alloc(&d_open, 2ull * total_polys + 1))) return rc;
    if ((rc = bufs.alloc(&d_misc, 2 + p->num_query_rounds))) return rc;   // [0] status, [2..] query indices
    FriDeviceIO io{};
    io.proof = d_proof; io.openings = d_open; io.indices = d_misc + 2; io.status = reinterpret_cast<uint32_t*>(d_misc);
    CK(launch_fill_words(d_misc, 1, 0, ctx->stream));
    if ((rc = fri_core(ctx, oracles, num_oracles, inst, p, ch->dev, pow_witness_in, io))) return rc;
    uint32_t status[2] = {0, 0};
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(proof_out, d_proof, proof_words * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (openings_out) CK(cudaMemcpyAsync(openings_out, d_open, 2ull * total_polys * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status, d_misc, sizeof status, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (status[0] == 2) return fail(QPZK_EPOW, "proof of work failed: no witness found");
    if (!status[0]) return fail(QPZK_EPOW, "proof-of-work witness does not reach the required leading zeros");
    return 0;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// permutation argument: Z and partial products (perm.cu)
// ---------------------------------------------------------------------------------------------
// k_j = g^j (get_unique_coset_shifts) for j < count, on the device; one table per context, grown on demand
int get_kpow_table(qpzk_ctx* ctx, uint32_t count, const uint64_t** out) {
    if (ctx->kpow_size < count) {
        std::vector<uint64_t> h(count);
        uint64_t k = 1;
        for (uint32_t j = 0; j < count; j++) { h[j] = k; k = gl_mul_host(k, GL_GENERATOR); }
        uint64_t* d = nullptr;
        CK(cudaMalloc((void**)&d, count * sizeof(uint64_t)));
        cudaError_t e = cudaMemcpyAsync(d, h.data(), count * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // also: earlier kernels are done with the old table
        if (e != cudaSuccess) { cudaFree(d); CK(e); }
        if (ctx->kpow) cudaFree(ctx->kpow);
        ctx->kpow = d; ctx->kpow_size = count;
    }
    *out = ctx->kpow;
    return 0;
}

// betas / gammas: host challenges, or chal_dev = [betas | gammas] already on the device (the single-call prover: the
// challenges never visit the host; beta * k_j is then computed by a kernel instead of on the host)
int partial_products_impl(qpzk_ctx* ctx, const uint64_t* wires, const uint64_t* sigmas, bool host, uint32_t log_n,
                          uint32_t num_routed, uint32_t chunk, uint32_t num_challenges, const uint64_t* betas,
                          const uint64_t* gammas, const uint64_t* chal_dev, uint64_t* out) {
    NvtxRange nvtx_pp("qpzk:partial_products");
    if (!ctx || !wires || !sigmas || (!chal_dev && (!betas || !gammas)) || !out) return fail(QPZK_EINVAL, "null argument");
    if (log_n > 28) return fail(QPZK_EINVAL, "log_n must be <= 28");
    if (!num_routed || !chunk) return fail(QPZK_EINVAL, "num_routed and chunk must be positive");
    const uint32_t nchunks = (num_routed + chunk - 1) / chunk;
    if (nchunks > PP_MAX_CHUNKS) return fail(QPZK_EINVAL, "more than 32 chunks per row");
    if (!num_challenges || num_challenges > 8) return fail(QPZK_EINVAL, "num_challenges must be in 1..8");
    for (uint32_t c = 0; c < num_challenges && !chal_dev; c++)
        if (betas[c] >= GL_P || gammas[c] >= GL_P) return fail(QPZK_EINVAL, "challenge is not canonical");
    CK(cudaSetDevice(ctx->device));
    const uint64_t n = 1ull << log_n, in_words = (uint64_t)num_routed * n, out_words = (uint64_t)num_challenges * nchunks * n;
    PowTables tw;
    int rc = get_pow_tables(ctx, log_n, false, &tw);
    if (rc) return rc;
    const uint64_t* kpow = nullptr;
    if (chal_dev && (rc = get_kpow_table(ctx, num_routed, &kpow))) return rc;
    // small parameter block: betas | gammas | beta * k_j per challenge,  k_j = g^j (get_unique_coset_shifts)
    std::vector<uint64_t> par(2 * num_challenges + (size_t)num_challenges * num_routed);
    for (uint32_t c = 0; c < num_challenges && !chal_dev; c++) {
        par[c] = betas[c]; par[num_challenges + c] = gammas[c];
        uint64_t k = 1;
        for (uint32_t j = 0; j < num_routed; j++) {
            par[2 * num_challenges + (size_t)c * num_routed + j] = gl_mul_host(betas[c], k);
            k = gl_mul_host(k, GL_GENERATOR);
        }
    }
    DevBufs bufs(ctx);
    uint64_t *d_par, *d_chunk, *d_row, *d_blk, *d_w = nullptr, *d_s = nullptr, *d_out = nullptr;
    if ((rc = bufs.alloc(&d_par, par.size()))) return rc;
    if ((rc = bufs.alloc(&d_blk, (uint64_t)num_challenges * pp_num_blocks(log_n)))) return rc;
    if ((rc = bufs.alloc(&d_chunk, out_words))) return rc;
    if ((rc = bufs.alloc(&d_row, (uint64_t)num_challenges * n))) return rc;
    if (host) {
        if ((rc = bufs.alloc(&d_w, in_words))) return rc;
        if ((rc = bufs.alloc(&d_s, in_words))) return rc;
        if ((rc = bufs.alloc(&d_out, out_words))) return rc;
        CK(cudaMemcpyAsync(d_w, wires, in_words * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_s, sigmas, in_words * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!chal_dev) CK(cudaMemcpyAsync(d_par, par.data(), par.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    PartialProductsArgs a{};
    a.wires = host ? d_w : wires; a.sigmas = host ? d_s : sigmas; a.col_stride = n;
    a.log_n = log_n; a.num_routed = num_routed; a.chunk = chunk; a.nchunks = nchunks; a.num_challenges = num_challenges;
    a.betas = d_par; a.gammas = d_par + num_challenges; a.beta_k = d_par + 2 * num_challenges;
    if (chal_dev) {
        a.betas = chal_dev; a.gammas = chal_dev + num_challenges;
        CK(launch_perm_params(chal_dev, kpow, num_challenges, num_routed, d_par + 2 * num_challenges, ctx->stream));
        ctx->launches++;
    }
    a.tw_lo = tw.lo; a.tw_hi = tw.hi;
    a.chunkprod = d_chunk; a.rowprod = d_row; a.blockprod = d_blk; a.out = host ? d_out : out;
    CK(launch_partial_products(a, ctx->stream));
    ctx->launches += 3;
    if (host) CK(cudaMemcpyAsync(out, d_out, out_words * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    // everything here is a function of the (secret) wires
    scrub(ctx, d_chunk, out_words);
    scrub(ctx, d_row, (uint64_t)num_challenges * n);
    if (host) { scrub(ctx, d_w, in_words); scrub(ctx, d_out, out_words); }
    // host variant: the caller's buffers are released on return.  (`par` is pageable: its copy was staged by the
    // driver before cudaMemcpyAsync returned, so the device variant need not wait.)
    if (host) CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int qpzk_partial_products(qpzk_ctx* ctx, const uint64_t* wires, const uint64_t* sigmas, uint32_t log_n,
                                     uint32_t num_routed, uint32_t chunk, uint32_t num_challenges, const uint64_t* betas,
                                     const uint64_t* gammas, uint64_t* out) {
    GUARD_BEGIN return partial_products_impl(ctx, wires, sigmas, true, log_n, num_routed, chunk, num_challenges, betas, gammas, nullptr, out); GUARD_END
}
extern "C" int qpzk_partial_products_dev(qpzk_ctx* ctx, const uint64_t* wires, const uint64_t* sigmas, uint32_t log_n,
                                         uint32_t num_routed, uint32_t chunk, uint32_t num_challenges, const uint64_t* betas,
                                         const uint64_t* gammas, uint64_t* out) {
    GUARD_BEGIN return partial_products_impl(ctx, wires, sigmas, false, log_n, num_routed, chunk, num_challenges, betas, gammas, nullptr, out); GUARD_END
}

int commit_zs_impl(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires, uint32_t sigma_col0,
                   uint32_t num_routed, uint32_t chunk, uint32_t num_challenges, const uint64_t* betas,
                   const uint64_t* gammas, const uint64_t* chal_dev, uint32_t cap_height, const uint64_t* salt_dev,
                   qpzk_batch** out) {
    if (!out) return fail(QPZK_EINVAL, "null argument");
    *out = nullptr;
    if (!ctx || !constants_sigmas || !wires) return fail(QPZK_EINVAL, "null argument");
    if (constants_sigmas->ctx != ctx || wires->ctx != ctx) return fail(QPZK_EINVAL, "batch belongs to another context");
    if (!constants_sigmas->values || !wires->values)
        return fail(QPZK_EINVAL, "batches hold no values: commit them with from_values after qpzk_ctx_set_keep_values(ctx, 1)");
    if (constants_sigmas->log_n != wires->log_n) return fail(QPZK_EINVAL, "batches must have the same degree");
    if (!num_routed || num_routed > wires->num_cols || (uint64_t)sigma_col0 + num_routed > constants_sigmas->num_cols)
        return fail(QPZK_EINVAL, "routed wires / sigma columns out of range");
    if (!chunk || !num_challenges) return fail(QPZK_EINVAL, "chunk and num_challenges must be positive");
    const uint64_t n = 1ull << wires->log_n;
    const uint32_t cols = num_challenges * ((num_routed + chunk - 1) / chunk);
    uint64_t* d_out;
    int rc = dev_alloc(ctx, &d_out, (uint64_t)cols * n);
    if (rc) return rc;
    rc = partial_products_impl(ctx, wires->values, constants_sigmas->values + (uint64_t)sigma_col0 * n, false, wires->log_n,
                               num_routed, chunk, num_challenges, betas, gammas, chal_dev, d_out);
    if (!rc) rc = commit_impl(ctx, dev_source(d_out), true, wires->log_n, cols, wires->rate_bits, cap_height, salt_dev, 0,
                              1u << wires->rate_bits, true, out);
    dev_free(ctx, d_out, (uint64_t)cols * n);
    return rc;
}
extern "C" int qpzk_commit_zs_partial_products(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                                               uint32_t sigma_col0, uint32_t num_routed, uint32_t chunk,
                                               uint32_t num_challenges, const uint64_t* betas, const uint64_t* gammas,
                                               uint32_t cap_height, const uint64_t* salt_dev, qpzk_batch** out) {
    GUARD_BEGIN
    return commit_zs_impl(ctx, constants_sigmas, wires, sigma_col0, num_routed, chunk, num_challenges, betas, gammas, nullptr,
                          cap_height, salt_dev, out);
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// quotient polynomials of the permutation argument (+ caller-supplied gate terms) (quotient.cu)
// ---------------------------------------------------------------------------------------------
static int get_unshift_tables(qpzk_ctx* ctx, uint32_t log_size, PowTables* out) {
    auto it = ctx->unshift_tables.find(log_size);
    if (it == ctx->unshift_tables.end()) {
        PowTables t;
        t.hi_size = log_size > TW_LO_BITS ? (1ull << (log_size - TW_LO_BITS)) : 1;
        std::vector<uint64_t> lo, hi;
        fill_powers(gl_inv_host(GL_GENERATOR), t.hi_size, lo, hi);
        int rc;
        if ((rc = upload_words(ctx, lo, &t.lo))) return rc;
        if ((rc = upload_words(ctx, hi, &t.hi))) return rc;
        it = ctx->unshift_tables.emplace(log_size, t).first;
    }
    *out = it->second;
    return 0;
}

// Leaves *coeffs_out (device, (num_challenges << qbits) x n words = the quotient chunks, column-major) to the caller,
// who releases it with dev_free.
// Z_H(x) = g^n w_Q^k - 1 for the Q residues k of the quotient coset, then the Q inverses; cached per (log_n, qbits)
static int get_zh_table(qpzk_ctx* ctx, uint32_t log_n, uint32_t qbits, const uint64_t** out) {
    auto key = std::make_pair(log_n, qbits);
    auto it = ctx->zh_tables.find(key);
    if (it == ctx->zh_tables.end()) {
        const uint32_t Q = 1u << qbits;
        std::vector<uint64_t> h;
        const uint64_t g_pow_n = gl_exp_host(GL_GENERATOR, 1ull << log_n), wq = gl_root_of_unity_host(qbits);
        uint64_t r = 1;
        for (uint32_t k = 0; k < Q; k++) {
            const uint64_t v = gl_mul_host(g_pow_n, r);
            h.push_back(v ? v - 1 : GL_P - 1);
            r = gl_mul_host(r, wq);
        }
        for (uint32_t k = 0; k < Q; k++) h.push_back(gl_inv_host(h[k]));
        uint64_t* d = nullptr;
        int rc = upload_words(ctx, h, &d);
        if (rc) return rc;
        it = ctx->zh_tables.emplace(key, d).first;
    }
    *out = it->second;
    return 0;
}

// betas / gammas / alphas: host challenges, or chal_dev = [betas | gammas | alphas] already on the device
int quotient_impl(qpzk_ctx* ctx, const qpzk_batch* cs, const qpzk_batch* wires, const qpzk_batch* zs,
                  uint32_t sigma_col0, uint32_t num_routed, uint32_t chunk, uint32_t num_challenges,
                  const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas, const uint64_t* chal_dev,
                  const uint64_t* gate_terms_dev, uint32_t num_gate_terms, uint64_t** coeffs_out, uint32_t* qbits_out) {
    NvtxRange nvtx_q("qpzk:compute_quotient_polys");
    *coeffs_out = nullptr;
    if (!ctx || !cs || !wires || !zs || (!chal_dev && (!betas || !gammas || !alphas))) return fail(QPZK_EINVAL, "null argument");
    if (cs->ctx != ctx || wires->ctx != ctx || zs->ctx != ctx) return fail(QPZK_EINVAL, "batch belongs to another context");
    if (!num_challenges || num_challenges > QUOT_MAX_CHALLENGES) return fail(QPZK_EINVAL, "num_challenges must be in 1..4");
    if (!chunk || (chunk & (chunk - 1))) return fail(QPZK_EINVAL, "quotient_degree_factor must be a power of two");
    const uint32_t qbits = ilog2(chunk), log_n = wires->log_n;
    if (cs->log_n != log_n || zs->log_n != log_n || cs->rate_bits != wires->rate_bits || zs->rate_bits != wires->rate_bits)
        return fail(QPZK_EINVAL, "the three batches must have the same degree and rate");
    if (qbits > wires->rate_bits) return fail(QPZK_EINVAL, "quotient_degree_factor exceeds the blowup factor");
    const uint64_t full = (1ull << log_n) << wires->rate_bits;
    if (cs->num_leaves != full || wires->num_leaves != full || zs->num_leaves != full)
        return fail(QPZK_EINVAL, "coset-split batches cannot feed the quotient");
    if (!num_routed || num_routed > wires->num_cols || (uint64_t)sigma_col0 + num_routed > cs->num_cols)
        return fail(QPZK_EINVAL, "routed wires / sigma columns out of range");
    const uint32_t nchunks = (num_routed + chunk - 1) / chunk;
    if (zs->num_cols < num_challenges * nchunks) return fail(QPZK_EINVAL, "zs/partial-products batch has too few columns");
    if (num_gate_terms && !gate_terms_dev) return fail(QPZK_EINVAL, "gate_terms is null");
    for (uint32_t c = 0; c < num_challenges && !chal_dev; c++)
        if (betas[c] >= GL_P || gammas[c] >= GL_P || alphas[c] >= GL_P) return fail(QPZK_EINVAL, "challenge is not canonical");
    CK(cudaSetDevice(ctx->device));
    const uint32_t log_q = log_n + qbits;
    const uint64_t n = 1ull << log_n, Lq = 1ull << log_q;
    const uint32_t T = num_challenges + num_challenges * nchunks + num_gate_terms;
    PowTables tw, un;
    int rc;
    if ((rc = get_pow_tables(ctx, log_q, false, &tw))) return rc;
    if ((rc = get_unshift_tables(ctx, log_q, &un))) return rc;
    // parameter block: betas | gammas | beta * k_j | alpha powers (computed on the host from host challenges, by two
    // small kernels from device challenges); Z_H on the coset and its inverses come from a per-context table
    const uint64_t* d_zh = nullptr;
    if ((rc = get_zh_table(ctx, log_n, qbits, &d_zh))) return rc;
    const size_t off_bk = 2 * (size_t)num_challenges, off_apow = off_bk + (size_t)num_challenges * num_routed;
    const size_t par_words = off_apow + (size_t)num_challenges * T;
    std::vector<uint64_t> par;
    const uint64_t* kpow = nullptr;
    if (chal_dev) {
        if ((rc = get_kpow_table(ctx, num_routed, &kpow))) return rc;
    } else {
        par.insert(par.end(), betas, betas + num_challenges);
        par.insert(par.end(), gammas, gammas + num_challenges);
        for (uint32_t c = 0; c < num_challenges; c++) {
            uint64_t k = 1;
            for (uint32_t j = 0; j < num_routed; j++) { par.push_back(gl_mul_host(betas[c], k)); k = gl_mul_host(k, GL_GENERATOR); }
        }
        for (uint32_t c = 0; c < num_challenges; c++) {
            uint64_t v = 1;
            for (uint32_t t = 0; t < T; t++) { par.push_back(v); v = gl_mul_host(v, alphas[c]); }
        }
    }
    DevBufs bufs(ctx);
    uint64_t *d_par, *d_vals, *d_coeffs;
    if ((rc = bufs.alloc(&d_par, par_words))) return rc;
    if ((rc = bufs.alloc(&d_vals, (uint64_t)num_challenges * Lq))) return rc;
    if ((rc = dev_alloc(ctx, &d_coeffs, (uint64_t)num_challenges * Lq))) return rc;
    auto bail = [&](cudaError_t e, const char* what) {
        dev_free(ctx, d_coeffs);
        return fail(QPZK_ECUDA, std::string(what) + " failed: " + cudaGetErrorString(e));
    };
    cudaError_t e = cudaSuccess;
    if (chal_dev) {
        e = launch_perm_params(chal_dev, kpow, num_challenges, num_routed, d_par + off_bk, ctx->stream);
        if (e == cudaSuccess) e = launch_base_powers(chal_dev + 2 * num_challenges, num_challenges, T, d_par + off_apow, ctx->stream);
        ctx->launches += 2;
    } else {
        e = cudaMemcpyAsync(d_par, par.data(), par.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e != cudaSuccess) return bail(e, "quotient parameters");
    QuotientArgs a{};
    a.cs_leaves = cs->leaves; a.w_leaves = wires->leaves; a.z_leaves = zs->leaves;
    a.cs_width = cs->leaf_width; a.w_width = wires->leaf_width; a.z_width = zs->leaf_width; a.sigma_col0 = sigma_col0;
    a.log_n = log_n; a.qbits = qbits; a.num_routed = num_routed; a.chunk = chunk; a.nchunks = nchunks;
    a.num_challenges = num_challenges; a.num_gate_terms = num_gate_terms; a.num_terms = T;
    a.gate_terms = gate_terms_dev;
    a.betas = chal_dev ? chal_dev : d_par; a.gammas = a.betas + num_challenges; a.beta_k = d_par + off_bk;
    a.apow = d_par + off_apow; a.zh = d_zh;
    a.n_field = n % GL_P;
    a.tw_lo = tw.lo; a.tw_hi = tw.hi;
    a.values = d_vals;
    if ((e = launch_quotient_perm(a, ctx->stream)) != cudaSuccess) return bail(e, "quotient_perm_kernel");
    ctx->launches++;
    if ((rc = run_intt(ctx, d_vals, d_coeffs, log_q, num_challenges))) { dev_free(ctx, d_coeffs); return rc; }
    if ((e = launch_coset_unshift(d_coeffs, log_q, num_challenges, un.lo, un.hi, ctx->stream)) != cudaSuccess)
        return bail(e, "coset_unshift_kernel");
    ctx->launches++;
    scrub(ctx, d_vals, (uint64_t)num_challenges * Lq);
    // `par` is pageable host memory: the copy above was staged by the driver before cudaMemcpyAsync returned
    *coeffs_out = d_coeffs;
    *qbits_out = qbits;
    return 0;
}

// Validates a gate program against the column counts of the two batches it reads and copies it to the device (the copy
// has completed when this returns: the program lives in caller memory).  num_terms = number of EMITs.
int gate_program_upload(qpzk_ctx* ctx, uint32_t cs_cols, uint32_t wire_cols, const qpzk_gate_program* prog, GateProgramDev* out) {
    *out = GateProgramDev{};
    if (!prog) return 0;
    if (!prog->code || (!prog->imm && prog->num_imm)) return fail(QPZK_EINVAL, "gate program: null code / imm");
    if (!prog->num_regs || prog->num_regs > GATE_MAX_REGS) return fail(QPZK_EINVAL, "gate program: num_regs must be in 1..256");
    std::vector<bool> written(prog->num_regs, false);
    uint32_t emits = 0;
    for (uint32_t pc = 0; pc < prog->num_ops; pc++) {
        const uint32_t w0 = prog->code[4 * pc], op = w0 & 0xffu, c = (w0 >> 8) & 0xffu, dst = prog->code[4 * pc + 1],
                       a = prog->code[4 * pc + 2], b = prog->code[4 * pc + 3];
        auto bad = [&](const char* what) { return fail(QPZK_EINVAL, std::string("gate program, instruction ") + std::to_string(pc) + ": " + what); };
        auto reg_ok = [&](uint32_t r) { return r < prog->num_regs && written[r]; };
        if (op > 10 || (w0 >> 16)) return bad("unknown opcode");
        if (op != 6 && dst >= prog->num_regs) return bad("destination register out of range");
        if (op == 0 && a >= wire_cols) return bad("wire column out of range");
        if (op == 1 && a >= cs_cols) return bad("constants/sigmas column out of range");
        if (op == 2 && a >= prog->num_imm) return bad("immediate out of range");
        if ((op == 3 || op == 4 || op == 5 || op == 7) && (!reg_ok(a) || !reg_ok(b))) return bad("source register out of range or not yet written");
        if ((op == 8 || op == 9 || op == 10) && !reg_ok(a)) return bad("source register out of range or not yet written");
        if ((op == 8 || op == 9 || op == 10) && b >= prog->num_imm) return bad("immediate out of range");
        if ((op == 7 || op == 10) && !reg_ok(c)) return bad("source register out of range or not yet written");
        if (op == 6 && !reg_ok(a)) return bad("emitted register out of range or not yet written");
        if (op != 6) written[dst] = true; else emits++;
    }
    for (uint32_t i = 0; i < prog->num_imm; i++) if (prog->imm[i] >= GL_P) return fail(QPZK_EINVAL, "gate program: immediate is not canonical");
    if (!emits) return 0;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = dev_alloc(ctx, &out->code, 2ull * prog->num_ops + 2))) return rc;          // 4 x u32 per instruction, 16-byte aligned
    if ((rc = dev_alloc(ctx, &out->imm, prog->num_imm ? prog->num_imm : 1))) { dev_free(ctx, out->code); *out = GateProgramDev{}; return rc; }
    cudaError_t e = cudaMemcpyAsync(out->code, prog->code, 16ull * prog->num_ops, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && prog->num_imm)
        e = cudaMemcpyAsync(out->imm, prog->imm, 8ull * prog->num_imm, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { gate_program_release(ctx, out); CK(e); }
    out->num_ops = prog->num_ops; out->num_terms = emits; out->num_regs = prog->num_regs;
    out->jit = gate_jit_get(prog->code, prog->num_ops, prog->imm, prog->num_imm, prog->num_regs);   // nullptr: interpreted
    return 0;
}
void gate_program_release(qpzk_ctx* ctx, GateProgramDev* g) {
    dev_free(ctx, g->code); dev_free(ctx, g->imm);
    *g = GateProgramDev{};
}
// Interprets an uploaded program at every quotient-domain point: *terms_out (device, [n Q][num_terms]) is left to the
// caller (dev_free).  A program without EMITs: nothing allocated.
int gate_terms_run(qpzk_ctx* ctx, const qpzk_batch* cs, const qpzk_batch* wires, uint32_t chunk, const GateProgramDev& g,
                   uint64_t** terms_out) {
    *terms_out = nullptr;
    if (!g.num_terms) return 0;
    if (!chunk || (chunk & (chunk - 1))) return fail(QPZK_EINVAL, "quotient_degree_factor must be a power of two");
    if (ilog2(chunk) > wires->rate_bits || cs->log_n != wires->log_n || cs->rate_bits != wires->rate_bits)
        return fail(QPZK_EINVAL, "batches must have the same degree and rate, quotient_degree_factor <= blowup");
    const uint64_t points = (1ull << wires->log_n) * chunk;
    uint64_t* d_terms;
    int rc;
    if ((rc = dev_alloc(ctx, &d_terms, points * g.num_terms))) return rc;
    GateProgramArgs ga{};
    ga.cs_leaves = cs->leaves; ga.w_leaves = wires->leaves; ga.cs_width = cs->leaf_width; ga.w_width = wires->leaf_width;
    ga.code = reinterpret_cast<const uint32_t*>(g.code); ga.imm = g.imm; ga.num_ops = g.num_ops; ga.num_terms = g.num_terms; ga.num_regs = g.num_regs;
    ga.num_points = points; ga.terms = d_terms;
    cudaError_t e = g.jit ? launch_gate_jit(g.jit, ga, ctx->stream) : launch_gate_program(ga, ctx->stream);
    ctx->launches++;
    if (e != cudaSuccess) { dev_free(ctx, d_terms); CK(e); }
    *terms_out = d_terms;
    return 0;
}
static int gate_terms_from_program(qpzk_ctx* ctx, const qpzk_batch* cs, const qpzk_batch* wires, uint32_t chunk,
                                   const qpzk_gate_program* prog, uint64_t** terms_out, uint32_t* num_terms_out) {
    *terms_out = nullptr; *num_terms_out = 0;
    if (!prog) return 0;
    if (!ctx || !cs || !wires) return fail(QPZK_EINVAL, "null argument");
    GateProgramDev g;
    int rc = gate_program_upload(ctx, cs->num_cols, wires->num_cols, prog, &g);
    if (rc) return rc;
    rc = gate_terms_run(ctx, cs, wires, chunk, g, terms_out);
    *num_terms_out = rc ? 0 : g.num_terms;
    gate_program_release(ctx, &g);   // back to the block cache: stream order keeps it intact until the kernel has run
    return rc;
}

extern "C" int qpzk_quotient_polys_prog(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                                        const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                                        uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                                        const uint64_t* gammas, const uint64_t* alphas, const qpzk_gate_program* prog, uint64_t* out) {
    GUARD_BEGIN
    uint64_t* terms; uint32_t nt;
    int rc = gate_terms_from_program(ctx, constants_sigmas, wires, quotient_degree_factor, prog, &terms, &nt);
    if (rc) return rc;
    rc = qpzk_quotient_polys(ctx, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, quotient_degree_factor,
                             num_challenges, betas, gammas, alphas, terms, nt, out);
    if (terms) dev_free(ctx, terms, ((1ull << wires->log_n) * quotient_degree_factor) * nt);
    return rc;
    GUARD_END
}
extern "C" int qpzk_commit_quotient_prog(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                                         const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                                         uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                                         const uint64_t* gammas, const uint64_t* alphas, const qpzk_gate_program* prog,
                                         uint32_t cap_height, const uint64_t* salt_dev, qpzk_batch** out) {
    GUARD_BEGIN
    if (out) *out = nullptr;
    uint64_t* terms; uint32_t nt;
    int rc = gate_terms_from_program(ctx, constants_sigmas, wires, quotient_degree_factor, prog, &terms, &nt);
    if (rc) return rc;
    rc = qpzk_commit_quotient(ctx, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, quotient_degree_factor,
                              num_challenges, betas, gammas, alphas, terms, nt, cap_height, salt_dev, out);
    if (terms) dev_free(ctx, terms, ((1ull << wires->log_n) * quotient_degree_factor) * nt);
    return rc;
    GUARD_END
}

extern "C" int qpzk_quotient_polys(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                                   const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                                   uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                                   const uint64_t* gammas, const uint64_t* alphas, const uint64_t* gate_terms_dev,
                                   uint32_t num_gate_terms, uint64_t* out) {
    GUARD_BEGIN
    if (!out) return fail(QPZK_EINVAL, "null argument");
    uint64_t* d;
    uint32_t qbits;
    int rc = quotient_impl(ctx, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, quotient_degree_factor,
                           num_challenges, betas, gammas, alphas, nullptr, gate_terms_dev, num_gate_terms, &d, &qbits);
    if (rc) return rc;
    const uint64_t words = ((uint64_t)num_challenges << qbits) << wires->log_n;
    cudaError_t e = cudaMemcpyAsync(out, d, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, d, words);
    CK(e);
    return 0;
    GUARD_END
}

extern "C" int qpzk_commit_quotient(qpzk_ctx* ctx, const qpzk_batch* constants_sigmas, const qpzk_batch* wires,
                                    const qpzk_batch* zs_partial_products, uint32_t sigma_col0, uint32_t num_routed,
                                    uint32_t quotient_degree_factor, uint32_t num_challenges, const uint64_t* betas,
                                    const uint64_t* gammas, const uint64_t* alphas, const uint64_t* gate_terms_dev,
                                    uint32_t num_gate_terms, uint32_t cap_height, const uint64_t* salt_dev, qpzk_batch** out) {
    GUARD_BEGIN
    if (!out) return fail(QPZK_EINVAL, "null argument");
    *out = nullptr;
    uint64_t* d;
    uint32_t qbits;
    int rc = quotient_impl(ctx, constants_sigmas, wires, zs_partial_products, sigma_col0, num_routed, quotient_degree_factor,
                           num_challenges, betas, gammas, alphas, nullptr, gate_terms_dev, num_gate_terms, &d, &qbits);
    if (rc) return rc;
    const uint32_t cols = num_challenges << qbits;
    rc = commit_impl(ctx, dev_source(d), false, wires->log_n, cols, wires->rate_bits, cap_height, salt_dev, 0,
                     1u << wires->rate_bits, true, out);
    dev_free(ctx, d, (uint64_t)cols << wires->log_n);
    return rc;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------
// microbenchmarks (roofline denominators)
// ---------------------------------------------------------------------------------------------
extern "C" int qpzk_bench_imad(qpzk_ctx* ctx, uint32_t iters, double* imad_per_sec) {
    GUARD_BEGIN
    if (!ctx || !imad_per_sec) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    float ms = 0; double ops = 0;
    CK(run_imad_bench(iters, ctx->num_sms, ctx->stream, &ms, &ops));
    ctx->launches += 2;
    *imad_per_sec = ops / (ms * 1e-3);
    return 0;
    GUARD_END
}
extern "C" int qpzk_bench_poseidon(qpzk_ctx* ctx, uint32_t perms_per_thread, double* perms_per_sec) {
    GUARD_BEGIN
    if (!ctx || !perms_per_sec) return fail(QPZK_EINVAL, "null argument");
    CK(cudaSetDevice(ctx->device));
    float ms = 0; double total = 0;
    CK(run_poseidon_bench(perms_per_thread, ctx->num_sms, ctx->stream, &ms, &total));
    ctx->launches += 2;
    *perms_per_sec = total / (ms * 1e-3);
    return 0;
    GUARD_END
}
