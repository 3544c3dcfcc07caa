// Poseidon hashing kernels: Merkle leaf sponge straight from the row-major LDE buffer, 2-to-1
// compression levels down to the cap, the device-resident Fiat-Shamir challenger, and the FRI
// proof-of-work grinder.
//
// Replaces qp-plonky2 =1.5.5 `MerkleTree::new` (fill_digests_buf), `PoseidonHash::{hash_no_pad,
// hash_or_noop,two_to_one}`, `Challenger` and `fri_proof_of_work` (SURVEY.md 8a rows a7, a8, a10,
// a13; Appendix A.4-A.7).  Reached in the reference only through `prove()`
// (wormhole/prover/src/lib.rs:171-175) and `CircuitBuilder::build` (wormhole/circuit/src/circuit.rs:206-216).
#include <cstdio>
#include <cstdlib>
#include "internal.cuh"
#include "poseidon_rc.cuh"
#include "poseidon.cuh"

namespace qpzk {

// ---------------------------------------------------------------------------------------------
// constants upload (once per device)
// ---------------------------------------------------------------------------------------------
cudaError_t upload_poseidon_constants() {
    static uint64_t split[2 * 372];
    static uint32_t limb[31 * 36];
    for (int row = 0; row < 31; row++)
        for (int i = 0; i < 12; i++) {
            uint64_t c = (row >= 1 && row < 30) ? QPZK_POSEIDON_RC_HOST[12 * row + i] : 0;
            if (SPARSE_ROUNDS > 0 && row == 4 + LITERAL_PARTIAL) c = i == 0 ? QPZK_POSEIDON_DEV_K[0] : 0;
            split[12 * row + i] = c & 0xFFFFFFFFu;
            split[372 + 12 * row + i] = c >> 32;
            const uint64_t cl = c >= 4096 ? c - 4096 : c + (GL_P - 4096);  // limb_combine adds 4096 back
            limb[36 * row + i] = (uint32_t)cl & 0x3FFFFFu;
            limb[36 * row + 12 + i] = (uint32_t)(cl >> 22) & 0x3FFFFFu;
            limb[36 * row + 24 + i] = (uint32_t)(cl >> 44);
        }
    cudaError_t e;
    {
        static uint64_t flat[372];
        for (int row = 0; row < 31; row++)
            for (int i = 0; i < 12; i++) flat[12 * row + i] = row < 30 ? QPZK_POSEIDON_RC_HOST[12 * row + i] : 0;
        if ((e = cudaMemcpyToSymbol(d_rc_flat, flat, sizeof flat)) != cudaSuccess) return e;
    }
#if QPZK_MDS_FFT
    (void)split;
    e = cudaMemcpyToSymbol(c_rc_limb, limb, sizeof limb);
#else
    (void)limb;
    e = cudaMemcpyToSymbol(c_rc_split, split, sizeof split);
#endif
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_rc_first, QPZK_POSEIDON_RC_HOST, 12 * sizeof(uint64_t));
    if (e != cudaSuccess) return e;
#if !QPZK_MDS_FFT
    const uint32_t pow2[2] = {16, 2};
    if ((e = cudaMemcpyToSymbol(c_mds_pow2, pow2, sizeof pow2)) != cudaSuccess) return e;
#endif
    if (SPARSE_ROUNDS > 0) {
        if ((e = cudaMemcpyToSymbol(c_sp_k, QPZK_POSEIDON_DEV_K, SPARSE_ROUNDS * sizeof(uint64_t))) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_sp_init, QPZK_POSEIDON_DEV_INIT, 121 * sizeof(uint64_t))) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_sp_w, QPZK_POSEIDON_DEV_W, 11 * SPARSE_ROUNDS * sizeof(uint64_t))) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_sp_vhat, QPZK_POSEIDON_DEV_VHAT, 11 * SPARSE_ROUNDS * sizeof(uint64_t))) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_sp_v26, QPZK_POSEIDON_DEV_V26, 12 * sizeof(uint64_t))) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// leaf digests: digest[row] = hash_or_noop(leaves[row, 0..leaf_len))
// One thread per row; overwrite-mode sponge, rate 8.  The next chunk is loaded before the
// permutation of the current one so the (strided, L1-served) loads hide under ~20k integer ops.
// ---------------------------------------------------------------------------------------------
#ifndef QPZK_HASH_MINBLOCKS
#define QPZK_HASH_MINBLOCKS 6   // <= 80 registers: the limb-domain partial rounds would take 96 (5 blocks per SM) otherwise
#endif
__global__ void __launch_bounds__(HASH_THREADS, QPZK_HASH_MINBLOCKS)
hash_rows_kernel(const uint64_t* __restrict__ leaves, uint64_t num_rows, uint32_t leaf_len, uint64_t row_stride,
                 uint64_t* __restrict__ digests) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= num_rows) return;
    const uint64_t* p = leaves + row * row_stride;
    uint64_t* d = digests + 4 * row;
    if (leaf_len <= 4) {  // hash_or_noop: short rows are copied, zero-padded
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) d[i] = i < leaf_len ? p[i] : 0;
        return;
    }
    uint64_t s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    uint32_t full = leaf_len / 8, tail = leaf_len % 8;
    uint64_t nxt[8];
    if (full) {
#pragma unroll
        for (int i = 0; i < 8; i++) nxt[i] = __ldg(p + i);
    }
    for (uint32_t c = 0; c < full; c++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = nxt[i];
        if (c + 1 < full) {
#pragma unroll
            for (int i = 0; i < 8; i++) nxt[i] = __ldg(p + 8 * (c + 1) + i);
        }
        poseidon_permute(s);
    }
    if (tail) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if ((uint32_t)i < tail) s[i] = __ldg(p + 8 * full + i);
        poseidon_permute(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = gl_canon(s[i]);
}

// ---------------------------------------------------------------------------------------------
// one Merkle level: parent[i] = two_to_one(child[2i], child[2i+1])
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void compress_pair(const uint64_t* __restrict__ child, uint64_t* __restrict__ parent,
                                              uint64_t i) {
    const ulonglong2* c = reinterpret_cast<const ulonglong2*>(child + 8 * i);
    ulonglong2 v0 = c[0], v1 = c[1], v2 = c[2], v3 = c[3];
    uint64_t s[12] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, v3.x, v3.y, 0, 0, 0, 0};
    poseidon_permute(s);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(parent + 4 * i);
    o[0] = make_ulonglong2(gl_canon(s[0]), gl_canon(s[1]));
    o[1] = make_ulonglong2(gl_canon(s[2]), gl_canon(s[3]));
}

__global__ void __launch_bounds__(HASH_THREADS)
merkle_level_kernel(const uint64_t* __restrict__ child, uint64_t* __restrict__ parent, uint64_t num_parents) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < num_parents) compress_pair(child, parent, i);
}

// ---------------------------------------------------------------------------------------------
// Challenger (duplex sponge, overwrite mode; pop challenges from the back of the output buffer)
// ---------------------------------------------------------------------------------------------
// One warp runs the sponge: lane i < 12 keeps state word i in a register for the whole kernel, lane k < 8 keeps
// pending input k, and every duplexing is one warp-wide permutation (poseidon_permute_warp).  Observing the ~260
// opened values of a proof is 65 permutations in sequence; this is what sets the latency of that step.
struct WarpChallenger {
    uint64_t st, pend, outv;  // state word `lane`, pending input `lane`, output-buffer word `lane`
    uint32_t in_len, out_len; // warp-uniform
    int lane;
    __device__ void load(const ChallengerDev* c) {
        lane = threadIdx.x & 31;
        st = lane < 12 ? c->state[lane] : 0;
        pend = lane < 8 ? c->in[lane] : 0;
        outv = lane < 8 ? c->out[lane] : 0;
        in_len = c->in_len; out_len = c->out_len;
    }
    __device__ void store(ChallengerDev* c) const {
        if (lane < 12) c->state[lane] = st;
        if (lane < 8) { c->in[lane] = pend; c->out[lane] = outv; }
        if (lane == 0) { c->in_len = in_len; c->out_len = out_len; }
    }
    __device__ void duplex() {
        if ((uint32_t)lane < in_len) st = pend;
        in_len = 0;
        st = gl_canon(poseidon_permute_warp(st, lane));
        outv = st;
        out_len = 8;
    }
    __device__ void observe(uint64_t x) {   // x warp-uniform
        out_len = 0;
        if ((uint32_t)lane == in_len) pend = x;
        if (++in_len == 8) duplex();
    }
    __device__ uint64_t get() {             // returns the challenge on every lane
        if (in_len > 0 || out_len == 0) duplex();
        --out_len;
        return __shfl_sync(0xffffffffu, outv, out_len);
    }
};

__global__ void __launch_bounds__(32) challenger_observe_kernel(ChallengerDev* c, const uint64_t* __restrict__ elems, uint32_t n) {
    WarpChallenger w;
    w.load(c);
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {      // one coalesced load per 32 elements, then broadcast by shuffle
        const uint64_t mine = i0 + w.lane < n ? elems[i0 + w.lane] : 0;
        const uint32_t m = n - i0 < 32 ? n - i0 : 32;
        for (uint32_t j = 0; j < m; j++) w.observe(__shfl_sync(0xffffffffu, mine, j));
    }
    w.store(c);
}

__global__ void __launch_bounds__(32) challenger_get_kernel(ChallengerDev* c, uint64_t* __restrict__ out, uint32_t n) {
    WarpChallenger w;
    w.load(c);
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t v = w.get();
        if (w.lane == 0) out[i] = v;
    }
    w.store(c);
}

// ---------------------------------------------------------------------------------------------
// fri_proof_of_work: find the MINIMUM witness w such that, after observing w, the next challenge
// (state[7] of the duplex) has >= pow_bits leading zeros.  Upstream uses rayon find_any (any valid
// witness); the minimum is a valid choice and makes the result deterministic.
// ---------------------------------------------------------------------------------------------
// The same search in ONE launch: thread t tests the candidates t, t + T, t + 2T, ... (T = threads of the grid) in
// increasing order and stops as soon as its next candidate is not below the smallest hit so far.  Every candidate below
// the final minimum was therefore tested by its thread, so `found` ends as the minimum witness -- no grid barrier, no
// host round trip per window (a window without a hit used to cost one), and the whole proof stays one stream of work
// that a CUDA graph can replay.  `limit` bounds the search (status 2 in the finish kernel if nothing was found).
__global__ void __launch_bounds__(HASH_THREADS)
pow_grind_persistent_kernel(const ChallengerDev* __restrict__ c, uint32_t pow_bits, uint64_t limit,
                            unsigned long long* found) {
    const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
    uint64_t st0[12];
#pragma unroll
    for (int k = 0; k < 12; k++) st0[k] = c->state[k];
    const uint32_t pos = c->in_len;  // < 8 by the challenger invariant
#pragma unroll
    for (int k = 0; k < 8; k++)
        if ((uint32_t)k < pos) st0[k] = c->in[k];
    for (uint64_t cand = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; cand < limit; cand += T) {
        if (cand >= *reinterpret_cast<volatile unsigned long long*>(found)) break;
        const unsigned act = __activemask();
        uint64_t s[12];
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = st0[k];
#pragma unroll
        for (int k = 0; k < 8; k++)
            if ((uint32_t)k == pos) s[k] = cand;
        poseidon_permute(s);
        const uint64_t resp = gl_canon(s[7]);
        const bool hit = pow_bits == 0 || (resp >> (64 - pow_bits)) == 0;
        if (hit) atomicMin(found, (unsigned long long)cand);
        // The lanes of a warp must not run ahead of the lane that has a hit to publish: measured on B200, the diverged
        // rest of the warp otherwise keeps the scheduler for its whole remaining loop (590 permutations to the search
        // limit) before the hit lane gets to issue its atomic, and `found` stays empty for everyone until then.
        __syncwarp(act);
        if (hit) break;
    }
}

// PoseidonHash::hash_no_pad of a short device vector (public inputs): one half-warp runs the overwrite-mode sponge.
__global__ void __launch_bounds__(32) hash_no_pad_kernel(const uint64_t* __restrict__ in, uint32_t len, uint64_t* __restrict__ out) {
    const int lane = threadIdx.x & 15;
    uint64_t st = 0;
    for (uint32_t c = 0; c < len; c += 8) {
        const uint32_t m = len - c < 8 ? len - c : 8;
        if ((uint32_t)lane < m) st = in[c + lane];     // overwrite mode
        st = poseidon_permute_sub<16>(st, lane);
    }
    if (threadIdx.x < 4) out[threadIdx.x] = gl_canon(st);
}

// Observe the witness (the ground one or a caller-supplied one), draw the PoW response, then draw the
// query indices.  status[0]: 1 = ok, 0 = witness does not reach the difficulty, 2 = no hit in this window
// (challenger untouched; the host launches the next window).
__global__ void __launch_bounds__(32)
pow_finish_and_indices_kernel(ChallengerDev* c, const unsigned long long* found, uint64_t forced_witness,
                              uint32_t pow_bits, uint32_t num_queries, uint64_t lde_size,
                              uint64_t* __restrict__ witness_out, uint64_t* __restrict__ indices, uint32_t* status) {
    uint64_t w = forced_witness;
    if (forced_witness == ~0ULL) {
        w = (uint64_t)*found;
        if (w == ~0ULL) { if (threadIdx.x == 0) status[0] = 2; return; }
    }
    WarpChallenger ch;
    ch.load(c);
    ch.observe(w);
    const uint64_t resp = ch.get();
    if (ch.lane == 0) {
        status[0] = (pow_bits == 0 || (resp >> (64 - pow_bits)) == 0) ? 1u : 0u;
        *witness_out = w;
    }
    for (uint32_t q = 0; q < num_queries; q++) {
        const uint64_t v = ch.get();
        if (ch.lane == 0) indices[q] = v % lde_size;
    }
    ch.store(c);
}

// ---------------------------------------------------------------------------------------------
// Latency variants for small trees (a proof at 2^13 rows builds five trees of <= 2^16 leaves): one WARP per
// permutation chain (poseidon_permute_warp).  A tree level or a short leaf sponge then takes the ~8 us of a
// warp-wide permutation instead of the ~30 us of a one-thread permutation; used while the whole level fits the
// GPU's resident warps at once (<= WARP_COOP_MAX chains), where the thread-per-chain kernels would leave the
// machine idle anyway.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t WARP_COOP_MAX = 4096;
constexpr int WARP_COOP_THREADS = 128;

// Up to six consecutive levels in one launch: a block of 32 HALF-warps owns 32 consecutive parents of level l0 (its
// children are 64 digests of level l0-1), then 16 of level l0+1, ... -- a self-contained subtree, so the levels are
// separated by a block barrier instead of a kernel boundary (a proof builds five trees; every saved launch is
// ~10 us of dependent latency).  Two permutations per warp (poseidon_permute_sub<16>): the wide first levels of a
// subtree are throughput bound, and a half-warp already holds the 12 state words.
constexpr int SUBTREE_THREADS = 512;
__global__ void __launch_bounds__(SUBTREE_THREADS)
merkle_subtree_warp_kernel(uint64_t* __restrict__ digests, uint64_t num_leaves, uint32_t l0, uint32_t depth) {
    const uint32_t hw = threadIdx.x >> 4;
    const int lane = threadIdx.x & 15;
    for (uint32_t d = 0; d < depth; d++) {
        const uint32_t l = l0 + d, per_block = 32u >> d;
        const uint64_t parents = num_leaves >> l;
        const uint64_t i = (uint64_t)blockIdx.x * per_block + hw;
        const bool active = hw < per_block && i < parents;
        // both halves of a warp must run the shuffles together: a warp takes part when either of its halves is active
        const bool warp_active = (hw & ~1u) < per_block && (uint64_t)blockIdx.x * per_block + (hw & ~1u) < parents;
        if (warp_active) {
            const uint64_t* child = digests + digest_level_off(num_leaves, l - 1);
            uint64_t* parent = digests + digest_level_off(num_leaves, l);
            uint64_t st = active && lane < 8 ? child[8 * i + lane] : 0;
            st = poseidon_permute_sub<16>(st, lane);
            if (active && lane < 4) parent[4 * i + lane] = gl_canon(st);
        }
        __syncthreads();
    }
}

// one half-warp per row (two rows per warp)
__global__ void __launch_bounds__(WARP_COOP_THREADS)
hash_rows_warp_kernel(const uint64_t* __restrict__ leaves, uint64_t num_rows, uint32_t leaf_len, uint64_t row_stride,
                      uint64_t* __restrict__ digests) {
    const uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool active = row < num_rows;
    if ((row & ~1ull) >= num_rows) return;            // whole warp out of range (rows 2k and 2k+1 share a warp)
    const uint64_t* p = leaves + (active ? row : 0) * row_stride;
    uint64_t st = 0;
    if (leaf_len <= 4) {   // hash_or_noop
        if (active && lane < 4) digests[4 * row + lane] = (uint32_t)lane < leaf_len ? p[lane] : 0;
        return;
    }
    for (uint32_t c = 0; c < leaf_len; c += 8) {
        const uint32_t len = leaf_len - c < 8 ? leaf_len - c : 8;
        if ((uint32_t)lane < len) st = p[c + lane];     // overwrite mode
        st = poseidon_permute_sub<16>(st, lane);
    }
    if (active && lane < 4) digests[4 * row + lane] = gl_canon(st);
}

// ---------------------------------------------------------------------------------------------
// PoseidonPermutation::permute on a batch of states (12 words each, any u64 representatives), and a self-test
// hook that applies ONE linear layer (MDS + constants row `next`) so that tests can drive its carry paths with
// chosen states (random inputs reach them with probability ~2^-20).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HASH_THREADS)
permute_kernel(const uint64_t* __restrict__ in, uint64_t count, int mds_only_row, uint64_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[12 * i + k];
    if (mds_only_row >= 0) mds_layer_rc(s, mds_only_row); else poseidon_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[12 * i + k] = gl_canon(s[k]);
}

cudaError_t launch_permute(const uint64_t* in, uint64_t count, int mds_only_row, uint64_t* out, cudaStream_t st) {
    if (!count) return cudaSuccess;
    permute_kernel<<<(unsigned)((count + HASH_THREADS - 1) / HASH_THREADS), HASH_THREADS, 0, st>>>(in, count, mds_only_row, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_hash_rows(const uint64_t* leaves, uint64_t num_rows, uint32_t leaf_len, uint64_t row_stride,
                             uint64_t* digests, cudaStream_t st) {
    if (num_rows == 0) return cudaSuccess;
    if (num_rows <= WARP_COOP_MAX) {
        uint64_t blocks = (num_rows * 16 + WARP_COOP_THREADS - 1) / WARP_COOP_THREADS;
        hash_rows_warp_kernel<<<(unsigned)blocks, WARP_COOP_THREADS, 0, st>>>(leaves, num_rows, leaf_len, row_stride, digests);
        return cudaGetLastError();
    }
    // (smaller blocks for small leaf sets were measured: no gain, 5.08 vs 5.02 ms per proof)
    const uint32_t threads = HASH_THREADS;
    uint64_t blocks = (num_rows + threads - 1) / threads;
    hash_rows_kernel<<<(unsigned)blocks, threads, 0, st>>>(leaves, num_rows, leaf_len, row_stride, digests);
    return cudaGetLastError();
}

// digests: level 0 already filled.  Builds levels 1..num_levels-1.
// Plan shared by the launcher and the launch counter: thread-per-node levels while a level has more than
// WARP_COOP_MAX parents, then subtree kernels of up to 6 levels each.
static uint32_t merkle_plan(uint32_t num_levels, uint64_t num_leaves, uint64_t* digests, cudaStream_t st, bool launch) {
    uint32_t launches = 0;
    uint32_t l = 1;
    while (l < num_levels) {
        const uint64_t parents = num_leaves >> l;
        if (parents > WARP_COOP_MAX) {
            if (launch) {
                uint64_t blocks = (parents + HASH_THREADS - 1) / HASH_THREADS;
                merkle_level_kernel<<<(unsigned)blocks, HASH_THREADS, 0, st>>>(
                    digests + digest_level_off(num_leaves, l - 1), digests + digest_level_off(num_leaves, l), parents);
            }
            l++;
        } else {
            const uint32_t depth = num_levels - l < 6 ? num_levels - l : 6;
            if (launch)
                merkle_subtree_warp_kernel<<<(unsigned)((parents + 31) / 32), SUBTREE_THREADS, 0, st>>>(digests, num_leaves, l, depth);
            l += depth;
        }
        launches++;
    }
    return launches;
}

cudaError_t launch_merkle_levels(uint64_t* digests, uint32_t num_levels, uint64_t num_leaves, cudaStream_t st) {
    merkle_plan(num_levels, num_leaves, digests, st, true);
    return cudaGetLastError();
}

// number of kernel launches launch_merkle_levels issues (bookkeeping for qpzk_ctx_launch_count)
uint32_t merkle_levels_launches(uint32_t num_levels, uint64_t num_leaves) {
    return merkle_plan(num_levels, num_leaves, nullptr, nullptr, false);
}

cudaError_t launch_challenger_observe(ChallengerDev* c, const uint64_t* elems_dev, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    challenger_observe_kernel<<<1, 32, 0, st>>>(c, elems_dev, n);
    return cudaGetLastError();
}

cudaError_t launch_challenger_get(ChallengerDev* c, uint64_t* out_dev, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    challenger_get_kernel<<<1, 32, 0, st>>>(c, out_dev, n);
    return cudaGetLastError();
}

cudaError_t launch_hash_no_pad(const uint64_t* in_dev, uint32_t len, uint64_t* out_dev, cudaStream_t st) {
    hash_no_pad_kernel<<<1, 32, 0, st>>>(in_dev, len, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_pow(ChallengerDev* c, uint32_t pow_bits, uint64_t forced_witness, uint32_t num_queries,
                       uint64_t lde_size, uint64_t max_candidates, int num_sms,
                       unsigned long long* found_dev, uint64_t* witness_out, uint64_t* indices, uint32_t* status,
                       cudaStream_t st) {
    if (forced_witness == ~0ULL) {
        cudaError_t e = launch_fill_words(found_dev, 1, ~0ULL, st);
        if (e != cudaSuccess) return e;
        // one wave of the machine (about 2^17 candidates on a B200: a hit with probability 1 - e^-2 at 16 bits);
        // fewer blocks when the expected search is shorter than a wave
        uint64_t want = (pow_bits < 40 ? (2ull << pow_bits) : ~0ull) / HASH_THREADS + 1;
        // exactly the blocks that are resident at once (5 per SM at 89 registers): a sixth block per SM used to run as a
        // second wave after the first had finished its candidates (CUPTI trace, round 2: 294 us for a 16-bit search)
        static const int per_sm = [] {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pow_grind_persistent_kernel, HASH_THREADS, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 4; }
            return n;
        }();
        uint64_t blocks = (uint64_t)(num_sms > 0 ? num_sms : 148) * per_sm;
        if (want < blocks) blocks = want;
        pow_grind_persistent_kernel<<<(unsigned)blocks, HASH_THREADS, 0, st>>>(c, pow_bits, max_candidates < GL_P ? max_candidates : GL_P, found_dev);
    }
    pow_finish_and_indices_kernel<<<1, 32, 0, st>>>(c, found_dev, forced_witness, pow_bits, num_queries, lde_size,
                                                    witness_out, indices, status);
    return cudaGetLastError();
}

}  // namespace qpzk
