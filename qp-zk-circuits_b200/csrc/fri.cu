// FRI kernels: polynomial openings at an extension point, the batch-opening combine evaluated
// pointwise on the committed LDE leaves, value-domain folding, final-polynomial interpolation and the
// query-phase gather.
//
// Replaces qp-plonky2 =1.5.5 `PolynomialBatch::prove_openings`, `fri_proof` / `fri_committed_trees` /
// `fri_prover_query_rounds` and `OpeningSet::new`'s polynomial evaluations (SURVEY.md 8a rows a11, a12,
// a14, a15; Appendix A.7/A.8); reference call sites: `prove()` at wormhole/prover/src/lib.rs:171-175,
// wormhole/aggregator/src/private_batch/prover/lib.rs:326-330,
// wormhole/aggregator/src/public_batch/prover/lib.rs:301-305; FRI parameters at
// wormhole/circuit/src/profile.rs:168-174.
//
// Upstream works on COEFFICIENTS (reduce_polys_base, divide_by_linear, reduce_with_powers on chunks,
// coset_fft per round).  Field elements are canonical, so any algebraically equal computation gives
// identical bits; this file computes the same objects in the VALUE domain, where every step is an
// independent per-point (or per-chunk) kernel over data that is already resident and already in leaf
// (bit-reversed) order:
//   final_values[p] = sum_b alpha^(k_{b+1}+...) * (sum_j alpha^j f_bj(x_p) - sum_j alpha^j f_bj(z_b)) / (x_p - z_b)
//   fold:  16 consecutive values = one coset {x * w16^u};  size-16 inverse DFT gives a_t = x^t P_t(x^16),
//          folded value = sum_t a_t (beta/x)^t, and lands at the bit-reversed position of the next layer.
#include "internal.cuh"
#include "goldilocks.cuh"
#include "ntt_local.cuh"

namespace qpzk {

__device__ __forceinline__ ext2 ext_shfl_xor(ext2 v, int m) {
    ext2 r;
    r.a = __shfl_xor_sync(0xffffffffu, v.a, m);
    r.b = __shfl_xor_sync(0xffffffffu, v.b, m);
    return r;
}

// out[c] = poly_c(z) for every column; coefficients column-major, natural order.  One block per column;
// thread t runs Horner in z^T over coefficients t, t+T, ... (coalesced), then the block sums z^t * partial_t.
__global__ void __launch_bounds__(EVAL_THREADS)
eval_ext_kernel(const uint64_t* __restrict__ coeffs, uint64_t col_stride, uint32_t log_n, ext2 z,
                uint64_t* __restrict__ out) {
    __shared__ ext2 red[EVAL_THREADS];
    const uint64_t n = 1ull << log_n;
    const uint64_t* c = coeffs + (uint64_t)blockIdx.x * col_stride;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    ext2 zT = ext_exp(z, T);
    ext2 acc = ext_make(0, 0);
    if (t < n) {
        uint64_t last = t + ((n - 1 - t) / T) * T;
        for (int64_t i = (int64_t)last; i >= (int64_t)t; i -= T) {
            acc = ext_mul(acc, zT);
            acc.a = gl_add(acc.a, c[i]);
        }
        acc = ext_mul(acc, ext_exp(z, t));
    }
    red[t] = acc;
    __syncthreads();
    for (uint32_t h = T >> 1; h > 0; h >>= 1) {
        if (t < h) red[t] = ext_add(red[t], red[t + h]);
        __syncthreads();
    }
    if (t == 0) { out[2 * blockIdx.x] = red[0].a; out[2 * blockIdx.x + 1] = red[0].b; }
}

__global__ void __launch_bounds__(EVAL_THREADS) eval_ext_multi_kernel(EvalMultiArgs a) {
    uint32_t job = 0;
    while (job + 1 < a.num_jobs && blockIdx.x >= a.col_begin[job + 1]) job++;
    const uint32_t col = blockIdx.x - a.col_begin[job];
    __shared__ ext2 red[EVAL_THREADS];
    const uint64_t n = 1ull << a.log_n;
    const uint64_t* c = a.coeffs[job] + (uint64_t)col * a.col_stride;
    const ext2 z = a.z_dev[job] ? ext_make(a.z_dev[job][0], a.z_dev[job][1]) : ext_make(a.z[2 * job], a.z[2 * job + 1]);
    const uint32_t T = blockDim.x, t = threadIdx.x;
    ext2 zT = ext_exp(z, T);
    ext2 acc = ext_make(0, 0);
    if (t < n) {
        uint64_t last = t + ((n - 1 - t) / T) * T;
        for (int64_t i = (int64_t)last; i >= (int64_t)t; i -= T) {
            acc = ext_mul(acc, zT);
            acc.a = gl_add(acc.a, c[i]);
        }
        acc = ext_mul(acc, ext_exp(z, t));
    }
    red[t] = acc;
    __syncthreads();
    for (uint32_t h = T >> 1; h > 0; h >>= 1) {
        if (t < h) red[t] = ext_add(red[t], red[t + h]);
        __syncthreads();
    }
    if (t == 0) { a.out[job][2 * col] = red[0].a; a.out[job][2 * col + 1] = red[0].b; }
}

// apow[j] = alpha^j for j in [0, count): one thread per power (square-and-multiply, <= 2 log2(count) ext products)
__global__ void alpha_powers_kernel(const uint64_t* __restrict__ alpha, uint64_t* __restrict__ apow, uint32_t count) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    ext2 p = ext_exp(ext_make(alpha[0], alpha[1]), j);
    apow[2 * j] = p.a; apow[2 * j + 1] = p.b;
}

// reduced[b] = sum_j alpha^j * opening_bj  with opening_bj = evals[b][oracle_j][index_j]; one block per batch
__global__ void __launch_bounds__(EVAL_THREADS) reduce_openings_kernel(FriCombineArgs a) {
    __shared__ ext2 red[EVAL_THREADS];
    const uint32_t b = blockIdx.x, t = threadIdx.x;
    const uint32_t off = a.batch_off[b], k = a.batch_off[b + 1] - off;
    ext2 acc = ext_make(0, 0);
    for (uint32_t j = t; j < k; j += blockDim.x) {
        uint32_t o = a.poly_oracle[off + j], idx = a.poly_index[off + j];
        const uint64_t* ev = a.evals + 2 * ((uint64_t)a.eval_off[b * a.num_oracles + o] + idx);
        acc = ext_add(acc, ext_mul(ext_make(a.apow[2 * j], a.apow[2 * j + 1]), ext_make(ev[0], ev[1])));
        // also record the opening itself in instance order (the caller may want OpeningSet values)
        a.openings_out[2 * (off + j)] = ev[0];
        a.openings_out[2 * (off + j) + 1] = ev[1];
    }
    red[t] = acc;
    __syncthreads();
    for (uint32_t h = blockDim.x >> 1; h > 0; h >>= 1) {
        if (t < h) red[t] = ext_add(red[t], red[t + h]);
        __syncthreads();
    }
    if (t == 0) { a.reduced[2 * b] = red[0].a; a.reduced[2 * b + 1] = red[0].b; }
}

// the opened values in FRI-instance order (what OpeningSet holds and the challenger observes before alpha is drawn)
__global__ void __launch_bounds__(EVAL_THREADS) openings_gather_kernel(FriCombineArgs a, uint32_t total) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    uint32_t b = 0;
    while (b + 1 < a.num_batches && k >= a.batch_off[b + 1]) b++;
    const uint32_t o = a.poly_oracle[k], idx = a.poly_index[k];
    const uint64_t* ev = a.evals + 2 * ((uint64_t)a.eval_off[b * a.num_oracles + o] + idx);
    a.openings_out[2 * k] = ev[0];
    a.openings_out[2 * k + 1] = ev[1];
}

__global__ void ext_scale_kernel(const uint64_t* __restrict__ in, uint64_t s, uint64_t* __restrict__ out) {
    if (threadIdx.x < 2) out[threadIdx.x] = gl_mul(in[threadIdx.x], s);
}

// One warp per 32 consecutive LDE points (leaf order).  x_p = s_j' * w_n^bitrev(r) for p = j' * n + r.
// Phase 1: lane l inverts the denominators (x_p - z_b) of ITS point for every batch: 32 different inversions per
// warp-instruction stream (the first version inverted the same denominator on all 32 lanes: ~5 k of the ~5.5 k
// instructions a point cost).  Phase 2: point by point, the lanes share the sum over the opened polynomials and the
// owning lane's inverses arrive by shuffle.
__global__ void __launch_bounds__(COMBINE_THREADS) fri_combine_kernel(FriCombineArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n = 1ull << a.log_n, L = n << a.rate_bits;
    const uint64_t p0 = warp << 5;
    if (p0 >= L) return;
    ext2 inv[MAX_FRI_BATCHES];
    {
        const uint64_t p = p0 + lane < L ? p0 + lane : L - 1;
        const uint32_t jp = (uint32_t)(p >> a.log_n);
        const uint64_t r = p & (n - 1);
        const uint64_t m = a.log_n ? (__brevll(r) >> (64 - a.log_n)) : 0;
        const uint64_t x = gl_mul(a.cs_lo[(size_t)jp * TW_LO_SIZE + 1], tw_get(a.tw_lo, a.tw_hi, m));
#pragma unroll
        for (uint32_t b = 0; b < MAX_FRI_BATCHES; b++) {
            inv[b] = ext_make(0, 0);
            if (b < a.num_batches) {
                const uint64_t z0 = a.points_dev ? a.points_dev[2 * b] : a.points[2 * b];
                const uint64_t z1 = a.points_dev ? a.points_dev[2 * b + 1] : a.points[2 * b + 1];
                inv[b] = ext_inv(ext_make(gl_sub(x, z0), gl_neg(z1)));
            }
        }
    }
    const uint32_t count = L - p0 < 32 ? (uint32_t)(L - p0) : 32u;
    // Batch 0 is the big one (every polynomial of every oracle at zeta: ~276 for the leaf circuit).  The polynomials a
    // lane is responsible for do not change from point to point, so their column address, row stride and alpha power
    // are fetched ONCE per warp (they were four dependent loads per polynomial and point) and the point loop is left
    // with one independent load and one extension-by-base multiplication per polynomial.
    constexpr uint32_t PL = 10;                       // polynomials per lane kept in registers: 320 per batch
    const uint32_t k0 = a.batch_off[1] - a.batch_off[0];
    const bool hoist = k0 <= 32 * PL;
    const uint64_t* cptr[PL];
    uint32_t cstr[PL];
    ext2 cap[PL];
#pragma unroll
    for (uint32_t t = 0; t < PL; t++) {
        const uint32_t j = lane + 32 * t;
        cptr[t] = a.leaves[0]; cstr[t] = 0; cap[t] = ext_make(0, 0);     // harmless slot: value * 0
        if (hoist && j < k0) {
            const uint32_t o = a.poly_oracle[a.batch_off[0] + j];
            cptr[t] = a.leaves[o] + a.poly_index[a.batch_off[0] + j];
            cstr[t] = a.leaf_width[o];
            cap[t] = ext_make(a.apow[2 * j], a.apow[2 * j + 1]);
        }
    }
    // Points in groups of 8.  Every lane accumulates, for the 8 points of the group, the part of the sum that belongs
    // to ITS polynomials (8 independent loads per polynomial); a recursive-halving exchange (xor 16: keep 4 of the 8
    // partial sums and receive the partner's, xor 8: 2, xor 4: 1, then plain butterflies for xor 2 and 1) leaves the
    // total of point q = lane >> 2 on lanes 4q..4q+3: 9 exchange steps per group instead of 8 x 5, and the per-point
    // arithmetic that follows (two extension multiplications per batch) runs for 8 points at once instead of being
    // repeated by all 32 lanes for one point.  (ncu, before: 1940 instructions per point and lane, 3x the arithmetic.)
    (void)count;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    const uint32_t q = lane >> 2;
    for (uint32_t grp = 0; grp < 4; grp++) {
        const uint64_t pg = p0 + 8 * grp;                 // first point of the group
        if (pg >= L) break;
        ext2 sum = ext_make(0, 0);
#pragma unroll
        for (uint32_t b = 0; b < MAX_FRI_BATCHES; b++) {
            if (b >= a.num_batches) break;
            const uint32_t off = a.batch_off[b], k = a.batch_off[b + 1] - off;
            ext2 acc[8];
#pragma unroll
            for (int u = 0; u < 8; u++) acc[u] = ext_make(0, 0);
            if (b == 0 && hoist) {
#pragma unroll
                for (uint32_t t = 0; t < PL; t++) {
                    if (32 * t >= k0) break;
                    uint64_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) v[u] = cptr[t][min(pg + u, L - 1) * cstr[t]];
#pragma unroll
                    for (int u = 0; u < 8; u++) acc[u] = ext_add(acc[u], ext_scale(cap[t], v[u]));
                }
            } else {
                for (uint32_t jj = lane; jj < k; jj += 32) {
                    const uint32_t o = a.poly_oracle[off + jj], idx = a.poly_index[off + jj];
                    const ext2 ap = ext_make(a.apow[2 * jj], a.apow[2 * jj + 1]);
                    const uint64_t* col = a.leaves[o] + idx;
                    const uint64_t wdt = a.leaf_width[o];
                    uint64_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) v[u] = col[min(pg + u, L - 1) * wdt];
#pragma unroll
                    for (int u = 0; u < 8; u++) acc[u] = ext_add(acc[u], ext_scale(ap, v[u]));
                }
            }
            ext2 h4[4], h2[2], tot;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const ext2 keep = b4 ? acc[u + 4] : acc[u], give = b4 ? acc[u] : acc[u + 4];
                h4[u] = ext_add(keep, ext_shfl_xor(give, 16));
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const ext2 keep = b3 ? h4[u + 2] : h4[u], give = b3 ? h4[u] : h4[u + 2];
                h2[u] = ext_add(keep, ext_shfl_xor(give, 8));
            }
            {
                const ext2 keep = b2 ? h2[1] : h2[0], give = b2 ? h2[0] : h2[1];
                tot = ext_add(keep, ext_shfl_xor(give, 4));
            }
            tot = ext_add(tot, ext_shfl_xor(tot, 2));
            tot = ext_add(tot, ext_shfl_xor(tot, 1));
            // tot = sum over all polynomials of the batch at point pg + q  (q = 4*b4 + 2*b3 + b2 = lane >> 2)
            const ext2 numer = ext_sub(tot, ext_make(a.reduced[2 * b], a.reduced[2 * b + 1]));
            ext2 dinv;
            dinv.a = __shfl_sync(0xffffffffu, inv[b].a, 8 * grp + q);
            dinv.b = __shfl_sync(0xffffffffu, inv[b].b, 8 * grp + q);
            const ext2 quo = ext_mul(numer, dinv);
            sum = ext_add(ext_mul(sum, ext_make(a.apow[2 * k], a.apow[2 * k + 1])), quo);
        }
        const uint64_t p = pg + q;
        if ((lane & 3) == 0 && p < L) { a.values[2 * p] = sum.a; a.values[2 * p + 1] = sum.b; }
    }
}

// Fold one layer in the value domain.  in: 2^lg ext values in bit-reversed order on the coset
// shift * <w_{2^lg}>;  out: 2^(lg-AB) values, bit-reversed order, on the coset shift^(2^AB) * <w_{2^(lg-AB)}>.
template <int AB>
__global__ void __launch_bounds__(128) fri_fold_kernel(FriFoldArgs a) {
    constexpr int A = 1 << AB;
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lgo = a.lg - AB;
    if (c >> lgo) return;
    ext2 v[A];
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(a.in + 2 * (c << AB));
#pragma unroll
    for (int k = 0; k < A; k++) { ulonglong2 t = src[k]; v[k] = ext_make(t.x, t.y); }
    // inverse DFT, decimation in time: memory order is bit-reversed u, output natural t
#pragma unroll
    for (int lh = 0; lh < AB; lh++) {
        const int h = 1 << lh;
#pragma unroll
        for (int t = 0; t < A / 2; t++) {
            const int j = t & (h - 1), e0 = ((t >> lh) << (lh + 1)) + j, e1 = e0 + h;
            ext2 y = v[e1];
            if (j) y = ext_scale(y, a.itw[j << (AB - 1 - lh)]);
            ext2 x0 = v[e0];
            v[e0] = ext_add(x0, y);
            v[e1] = ext_sub(x0, y);
        }
    }
    // x_c^-1 = shift^-1 * w_{2^lg}^(-m),  m = bitrev_{lg-AB}(c);  table holds powers of w_L^-1
    const uint64_t m = lgo ? (__brevll(c) >> (64 - lgo)) : 0;
    const uint64_t xinv = gl_mul(a.shift_inv, tw_get(a.itw_lo, a.itw_hi, m << (a.log_L - a.lg)));
    ext2 beta = ext_make(a.beta[0], a.beta[1]);
    ext2 gam = ext_scale(beta, xinv);
    ext2 acc = v[A - 1];
#pragma unroll
    for (int t = A - 2; t >= 0; t--) acc = ext_add(ext_mul(acc, gam), v[t]);
    acc = ext_scale(acc, a.arity_inv);
    a.out[2 * c] = acc.a; a.out[2 * c + 1] = acc.b;
}

// Coefficients of the last layer: values (bit-reversed) on shift * <w_len>  ->  natural coefficients,
// truncated to len >> rate_bits.  Single block; the ext vector is a 2-wide tile for the shared transform.
__global__ void __launch_bounds__(NTT_THREADS) fri_final_poly_kernel(FriFinalArgs a) {
    extern __shared__ uint64_t smem[];
    const uint32_t len = 1u << a.lg;
    uint64_t* sm = smem;
    uint64_t* ltw = smem + 2 * (size_t)len;
    for (uint32_t j = threadIdx.x; j < (len >> 1); j += blockDim.x)
        ltw[j] = tw_get(a.itw_lo, a.itw_hi, (uint64_t)j << (a.log_L - a.lg));
    for (uint32_t i = threadIdx.x; i < 2 * len; i += blockDim.x) sm[i] = a.in[i];
    __syncthreads();
    local_ntt<true>(sm, ltw, a.lg, 2, true, TileMap{2u, 1u, 0u});
    const uint32_t keep = len >> a.rate_bits;
    for (uint32_t i = threadIdx.x; i < keep; i += blockDim.x) {
        uint64_t f = gl_mul(a.len_inv, gl_exp(a.shift_inv, i));
        a.out[2 * i] = gl_mul(sm[2 * i], f);
        a.out[2 * i + 1] = gl_mul(sm[2 * i + 1], f);
    }
}

// Query phase: one block per query copies leaf rows and Merkle paths into the flat proof buffer.
__global__ void __launch_bounds__(128) fri_query_gather_kernel(FriGatherArgs a) {
    const uint32_t q = blockIdx.x;
    if (a.status && a.status[0] != 1) return;  // no valid PoW witness yet: the indices are not defined
    uint64_t x = a.indices[q];
    uint64_t* out = a.proof + a.query_base + (uint64_t)q * a.query_words;
    for (uint32_t t = 0; t < a.num_trees; t++) {
        const FriTreeDesc& d = a.trees[t];
        if (t >= a.num_oracles) x >>= d.arity_bits;  // FRI round trees are indexed by the coset index
        const uint64_t* row = d.leaves + x * d.leaf_len;
        for (uint32_t i = threadIdx.x; i < d.leaf_len; i += blockDim.x) out[i] = row[i];
        out += d.leaf_len;
        const uint32_t ns = d.log_leaves - d.cap_height;
        const uint64_t N = 1ull << d.log_leaves;
        for (uint32_t i = threadIdx.x; i < 4 * ns; i += blockDim.x) {
            uint32_t l = i >> 2;
            uint64_t node = (x >> l) ^ 1;
            out[i] = d.digests[digest_level_off(N, l) + 4 * node + (i & 3)];
        }
        out += 4 * ns;
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_eval_ext(const uint64_t* coeffs, uint64_t col_stride, uint32_t log_n, uint32_t num_cols,
                            const uint64_t z[2], uint64_t* out_dev, cudaStream_t st) {
    if (!num_cols) return cudaSuccess;
    ext2 zz; zz.a = z[0]; zz.b = z[1];
    eval_ext_kernel<<<num_cols, EVAL_THREADS, 0, st>>>(coeffs, col_stride, log_n, zz, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_eval_ext_multi(const EvalMultiArgs& a, cudaStream_t st) {
    const uint32_t total = a.col_begin[a.num_jobs];
    if (!total) return cudaSuccess;
    eval_ext_multi_kernel<<<total, EVAL_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_alpha_powers(const uint64_t* alpha_dev, uint64_t* apow_dev, uint32_t count, cudaStream_t st) {
    alpha_powers_kernel<<<(count + 127) / 128, 128, 0, st>>>(alpha_dev, apow_dev, count);
    return cudaGetLastError();
}

cudaError_t launch_fri_combine(const FriCombineArgs& a, cudaStream_t st) {
    reduce_openings_kernel<<<a.num_batches, EVAL_THREADS, 0, st>>>(a);
    uint64_t L = 1ull << (a.log_n + a.rate_bits);
    uint64_t blocks = (L + COMBINE_THREADS - 1) / COMBINE_THREADS;   // one warp per 32 points
    fri_combine_kernel<<<(unsigned)blocks, COMBINE_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_openings_gather(const FriCombineArgs& a, uint32_t total_polys, cudaStream_t st) {
    if (!total_polys) return cudaSuccess;
    openings_gather_kernel<<<(total_polys + EVAL_THREADS - 1) / EVAL_THREADS, EVAL_THREADS, 0, st>>>(a, total_polys);
    return cudaGetLastError();
}

cudaError_t launch_ext_scale(const uint64_t* in_dev, uint64_t s, uint64_t* out_dev, cudaStream_t st) {
    ext_scale_kernel<<<1, 32, 0, st>>>(in_dev, s, out_dev);
    return cudaGetLastError();
}

cudaError_t launch_fri_fold(const FriFoldArgs& a, uint32_t arity_bits, cudaStream_t st) {
    uint64_t chunks = 1ull << (a.lg - arity_bits);
    unsigned blocks = (unsigned)((chunks + 127) / 128);
    switch (arity_bits) {
        case 1: fri_fold_kernel<1><<<blocks, 128, 0, st>>>(a); break;
        case 2: fri_fold_kernel<2><<<blocks, 128, 0, st>>>(a); break;
        case 3: fri_fold_kernel<3><<<blocks, 128, 0, st>>>(a); break;
        case 4: fri_fold_kernel<4><<<blocks, 128, 0, st>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_fri_final_poly(const FriFinalArgs& a, cudaStream_t st) {
    uint32_t len = 1u << a.lg;
    size_t smem = (2 * (size_t)len + len / 2 + 1) * sizeof(uint64_t);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        static unsigned long long done = 0;
        cudaError_t e = raise_smem_limit((const void*)fri_final_poly_kernel, &done);
        if (e != cudaSuccess) return e;
    }
    fri_final_poly_kernel<<<1, NTT_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fri_gather(const FriGatherArgs& a, uint32_t num_queries, cudaStream_t st) {
    if (!num_queries) return cudaSuccess;
    fri_query_gather_kernel<<<num_queries, 128, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace qpzk
