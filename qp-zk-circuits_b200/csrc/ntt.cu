// Batched Goldilocks NTT kernels: column iNTT and the coset low-degree extension whose last pass
// stores straight into the row-major, bit-reversed Merkle leaf buffer.
//
// Replaces qp-plonky2-field =1.5.5 `PolynomialValues::ifft`, `PolynomialCoeffs::{lde, coset_fft}`
// and the `transpose` + `reverse_index_bits_in_place` steps of `PolynomialBatch::from_coeffs`
// (Appendix A.2/A.5).  Reached from the reference through `prove()`
// (wormhole/prover/src/lib.rs:171-175) and circuit builds (wormhole/circuit/src/circuit.rs:206-216).
//
// Structure (all transforms are radix-2 stages in shared memory, grouped into passes of <= 10 bits):
//   iNTT  = decimation-in-time passes.  Pass 1 gathers the natural-order values at bit-reversed
//           addresses in W-wide coalesced segments; the last pass applies 1/n.  Output: natural-order
//           coefficients, column-major (this is `PolynomialBatch::polynomials`).
//   LDE   = for each of the 2^rate_bits cosets  s_j = g * w_L^bitrev(j):  decimation-in-frequency passes
//           on  c_i * s_j^i  (natural in, bit-reversed out).  In-place DIF leaves value k at position
//           bitrev(k), which is exactly plonky2's leaf order, and coset j is the contiguous block of n
//           leaves [j*n, (j+1)*n).  The LAST pass holds ALL columns of 2^s consecutive
//           rows in shared memory and writes whole rows, so the transpose and the bit reversal cost no
//           extra pass and every store is a full contiguous row range.
#include <cstdlib>
#include "internal.cuh"
#include "goldilocks.cuh"
#include "ntt_local.cuh"

namespace qpzk {

// One pass over position bits [b_lo, b_lo + s) of `num_arrays` column-major arrays of length 2^log_n.
// Tile = 2^s x W elements: positions  hi * 2^(b_lo+s) + e * 2^b_lo + low0 + w.
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(NttPassArgs a) {
    extern __shared__ uint64_t smem[];
    const uint32_t s = a.s, W = a.W, S = 1u << s;
    const TileMap map{W, 1u, W == 1 ? 1u : 0u};   // sm[e * W + w]; a single-column tile is padded
    uint64_t* sm = smem;
    uint64_t* ltw = smem + (W == 1 ? padded_len(S) : S * W);
    const uint32_t b_hi = a.b_lo + s;
    const uint32_t lW = 31 - __clz(W);

    uint32_t tile = blockIdx.x, coset = 0;
    if (a.mode == NTT_FWD_COSET_FIRST) { coset = tile & (a.num_cosets - 1); tile /= a.num_cosets; }
    const uint64_t col = blockIdx.y;
    const uint64_t* in = a.in + col * a.in_stride + (uint64_t)blockIdx.z * a.group_stride;

    // local twiddles w_S^j = w_n^(j * n / S)
    for (uint32_t j = threadIdx.x; j < (S >> 1); j += blockDim.x)
        ltw[j] = tw_get(a.tw_lo, a.tw_hi, (uint64_t)j << (a.log_n - s));

    uint64_t hi = 0, low0 = 0;
    if (a.mode == NTT_INV_FIRST) {
        // gather: slot (a_hi, w) <- in[a_hi * 2^(log_n - s) + q0 + w], placed at e = bitrev_s(a_hi)
        uint64_t q0 = (uint64_t)tile * W;
        for (uint32_t i = threadIdx.x; i < S * W; i += blockDim.x) {
            uint32_t w = i & (W - 1), ah = i >> lW;
            sm[map(brev(ah, s), w)] = in[((uint64_t)ah << (a.log_n - s)) + q0 + w];
        }
    } else {
        uint32_t llow = a.b_lo - lW;  // log2(number of W-wide tiles across the low bits)
        hi = tile >> llow;
        low0 = (uint64_t)(tile & ((1u << llow) - 1)) << lW;
        uint64_t base = (hi << b_hi) + low0;
        for (uint32_t i = threadIdx.x; i < S * W; i += blockDim.x) {
            uint32_t w = i & (W - 1), e = i >> lW;
            uint64_t p = base + ((uint64_t)e << a.b_lo) + w;
            uint64_t v = in[p];
            if (a.mode == NTT_FWD_COSET_FIRST) {  // c_p * s_j^p  (b_hi == log_n here, so p is the coefficient index)
                const uint64_t* clo = a.cs_lo + (size_t)coset * TW_LO_SIZE;
                const uint64_t* chi = a.cs_hi + (size_t)coset * a.cs_hi_size;
                v = gl_mulc(v, tw_get(clo, chi, p));
            } else if (a.mode == NTT_INV_GENERIC && a.b_lo) {  // DIT pre-twiddle w_{2^b_hi}^(bitrev_s(e) * k1)
                uint64_t k1 = low0 + w;
                v = gl_mulc(v, tw_get(a.tw_lo, a.tw_hi, ((uint64_t)brev(e, s) * k1) << (a.log_n - b_hi)));
            }
            sm[map(e, w)] = v;
        }
    }
    __syncthreads();

    if (a.mode == NTT_INV_FIRST || a.mode == NTT_INV_GENERIC) local_ntt<true>(sm, ltw, s, W, true, map);
    else local_ntt<false>(sm, ltw, s, W, true, map);

    uint64_t* out = a.out + col * a.out_stride + ((uint64_t)coset + blockIdx.z) * a.group_stride;
    if (a.mode == NTT_INV_FIRST) {
        uint64_t q0 = (uint64_t)tile * W;
        for (uint32_t i = threadIdx.x; i < S * W; i += blockDim.x) {
            uint32_t e = i & (S - 1), w = i >> s;  // contiguous chunk per w
            uint64_t c = brev((uint32_t)(q0 + w), a.log_n - s);
            uint64_t v = sm[map(e, w)];
            if (a.scale != 1) v = gl_mulc(v, a.scale);
            out[(c << s) + e] = v;
        }
    } else {
        uint64_t base = (hi << b_hi) + low0;
        for (uint32_t i = threadIdx.x; i < S * W; i += blockDim.x) {
            uint32_t w = i & (W - 1), e = i >> lW;
            uint64_t v = sm[map(e, w)];
            if ((a.mode == NTT_FWD_COSET_FIRST || a.mode == NTT_FWD_GENERIC) && a.b_lo) {
                // DIF post-twiddle w_{2^b_hi}^(k_local * p_low), k_local = bitrev_s(e)
                uint64_t pl = low0 + w;
                v = gl_mulc(v, tw_get(a.tw_lo, a.tw_hi, ((uint64_t)brev(e, s) * pl) << (a.log_n - b_hi)));
            }
            if (a.scale != 1) v = gl_mulc(v, a.scale);
            out[base + ((uint64_t)e << a.b_lo) + w] = v;
        }
    }
}

// Last LDE pass: position bits [0, s) of every column for one chunk of R = 2^s rows of one coset, then
// the transposed store of R complete rows into the leaf buffer.  sm[c * (R+1) + e] (padded: the row-major
// store reads it with stride R+1, conflict-free).
__global__ void __launch_bounds__(NTT_THREADS) lde_last_kernel(LdeLastArgs a) {
    extern __shared__ uint64_t smem[];
    const uint32_t s = a.s, R = 1u << s, RP = padded_len(R) | 1u, C = a.num_cols;
    const TileMap map{1u, RP, 1u};                 // sm[c * RP + pad(e)], RP odd: the row-major store is conflict free
    uint64_t* sm = smem;
    uint64_t* ltw = smem + (size_t)C * RP;
    const uint64_t n = 1ull << a.log_n;
    const uint64_t chunks = n >> s;
    const uint32_t coset = blockIdx.x / chunks;
    const uint64_t q = blockIdx.x % chunks;

    for (uint32_t j = threadIdx.x; j < (R >> 1); j += blockDim.x)
        ltw[j] = tw_get(a.tw_lo, a.tw_hi, (uint64_t)j << (a.log_n - s));

    const uint64_t* clo = a.cs_lo + (size_t)coset * TW_LO_SIZE;
    const uint64_t* chi = a.cs_hi + (size_t)coset * a.cs_hi_size;
    for (uint32_t i = threadIdx.x; i < C * R; i += blockDim.x) {
        uint32_t e = i & (R - 1), c = i >> s;
        uint64_t p = (q << s) + e;
        uint64_t v;
        if (a.from_coeffs) {  // single-pass LDE (log_n <= s): scale the coefficients here
            v = gl_mulc(a.in[(uint64_t)c * a.in_stride + p], tw_get(clo, chi, p));
        } else {
            v = a.in[((uint64_t)coset * C + c) * a.in_stride + p];
        }
        sm[map(e, c)] = v;
    }
    __syncthreads();

    local_ntt<false>(sm, ltw, s, C, false, map);

    uint64_t* rows = a.leaves + ((uint64_t)coset * n + (q << s)) * a.leaf_width;
    for (uint32_t e = threadIdx.x / 32; e < R; e += blockDim.x / 32)      // one warp per row: coalesced row stores
        for (uint32_t c = threadIdx.x % 32; c < C; c += 32) rows[(uint64_t)e * a.leaf_width + c] = sm[map(e, c)];
}

// out[c][bitrev(i)] = in[c][i]  (natural <-> bit-reversed order of one column-major matrix)
__global__ void bitrev_copy_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t log_n,
                                   uint64_t in_stride, uint64_t out_stride) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    uint64_t c = blockIdx.y;
    uint64_t r = log_n ? (__brevll(i) >> (64 - log_n)) : 0;
    out[c * out_stride + r] = in[c * in_stride + i];
}

// salt columns: leaves[bitrev(i)][num_cols + s] = salt[s][i]
__global__ void salt_scatter_kernel(const uint64_t* __restrict__ salt, uint64_t* __restrict__ leaves, uint32_t log_L,
                                    uint32_t num_cols, uint32_t leaf_width, uint32_t salt_cols) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_L) return;
    uint64_t r = log_L ? (__brevll(i) >> (64 - log_L)) : 0;
    for (uint32_t s = 0; s < salt_cols; s++) leaves[r * leaf_width + num_cols + s] = salt[((uint64_t)s << log_L) + i];
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
cudaError_t raise_smem_limit(const void* fn, unsigned long long* done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
    if (e == cudaSuccess) __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
    return e;
}

static cudaError_t ensure_smem(const void* fn, size_t bytes, unsigned long long* done) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    if (bytes > (size_t)MAX_DYN_SMEM) return cudaErrorInvalidValue;
    return raise_smem_limit(fn, done);
}

cudaError_t launch_ntt_pass(const NttPassArgs& a, cudaStream_t st) {
    uint64_t n = 1ull << a.log_n;
    uint64_t tiles = n / ((1ull << a.s) * a.W);
    if (a.mode == NTT_FWD_COSET_FIRST) tiles *= a.num_cosets;
    uint32_t S = 1u << a.s;
    size_t smem = ((size_t)(a.W == 1 ? padded_len(S) : S * a.W) + S / 2 + 1) * sizeof(uint64_t);
    static unsigned long long done = 0;
    cudaError_t e = ensure_smem((const void*)ntt_pass_kernel, smem, &done);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)tiles, (unsigned)a.num_arrays, (unsigned)(a.num_groups ? a.num_groups : 1));
    ntt_pass_kernel<<<grid, NTT_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_lde_last(const LdeLastArgs& a, cudaStream_t st) {
    uint64_t n = 1ull << a.log_n;
    uint64_t blocks = (n >> a.s) * a.num_cosets;
    uint32_t R = 1u << a.s;
    size_t smem = ((size_t)a.num_cols * (padded_len(R) | 1u) + R / 2 + 1) * sizeof(uint64_t);
    static unsigned long long done = 0;
    cudaError_t e = ensure_smem((const void*)lde_last_kernel, smem, &done);
    if (e != cudaSuccess) return e;
    lde_last_kernel<<<(unsigned)blocks, NTT_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_bitrev_copy(const uint64_t* in, uint64_t* out, uint32_t log_n, uint64_t num_arrays,
                               uint64_t in_stride, uint64_t out_stride, cudaStream_t st) {
    uint64_t n = 1ull << log_n;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)num_arrays);
    bitrev_copy_kernel<<<grid, 256, 0, st>>>(in, out, log_n, in_stride, out_stride);
    return cudaGetLastError();
}

cudaError_t launch_salt_scatter(const uint64_t* salt, uint64_t* leaves, uint32_t log_L, uint32_t num_cols,
                                uint32_t leaf_width, uint32_t salt_cols, cudaStream_t st) {
    uint64_t L = 1ull << log_L;
    salt_scatter_kernel<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(salt, leaves, log_L, num_cols, leaf_width,
                                                                    salt_cols);
    return cudaGetLastError();
}

// Fill `words` 64-bit words with `value`.  Used instead of cudaMemsetAsync everywhere: with several prover contexts
// (streams) active on one device, cudaMemsetAsync calls stalled EVERY stream for 0.2-2 s at random (about one run of
// 64 proofs x 4 threads in three; none in any run with the memsets removed or with a single hardware work queue) --
// a plain kernel on the context's stream has no such side effect.
__global__ void __launch_bounds__(256) fill_words_kernel(uint64_t* __restrict__ p, uint64_t words, uint64_t value) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {          // 16-byte stores for the bulk
        ulonglong2* q = reinterpret_cast<ulonglong2*>(p);
        const uint64_t pairs = words >> 1;
        for (uint64_t k = i; k < pairs; k += stride) q[k] = make_ulonglong2(value, value);
        if (i == 0 && (words & 1)) p[words - 1] = value;
    } else {
        for (; i < words; i += stride) p[i] = value;
    }
}

cudaError_t launch_fill_words(void* p, uint64_t words, uint64_t value, cudaStream_t st) {
    if (!words) return cudaSuccess;
    uint64_t blocks = (words / 2 + 255) / 256;
    if (blocks < 1) blocks = 1;
    static const uint64_t cap = [] { const char* e = getenv("QPZK_FILL_BLOCKS"); return (uint64_t)(e ? atoi(e) : 148 * 8); }();
    if (blocks > cap) blocks = cap;
    fill_words_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<uint64_t*>(p), words, value);
    return cudaGetLastError();
}

}  // namespace qpzk
