// Partial products and Z polynomials of the permutation argument on the device.
//
// Replaces qp-plonky2 =1.5.5 `all_wires_permutation_partial_products` / `wires_permutation_partial_products_and_zs`
// (plonk/prover.rs; SURVEY.md 8a row a16 -- a "next" row: the step between the wires commitment and the
// zs/partial-products commitment, reached from the reference through `prove()` at wormhole/prover/src/lib.rs:171-175).
// PARITY UNPINNED like the rest of the path (no golden vector in the reference tree): checked bit for bit against the
// CPU restatement (the oracle's partial_products, which is itself checked against big-integer arithmetic) and
// through the defining property of the argument -- for wires that satisfy the copy constraints Z closes to 1.
//
//   per row i (x = w^i) and challenge (beta, gamma):
//     chunk_k = prod_{j in chunk k} (wire_ij + beta k_j x + gamma) / (wire_ij + beta sigma_ij + gamma),   k_j = g^j
//     Z(w^0) = 1,  Z(w^(i+1)) = Z(w^i) * prod_k chunk_k(i),   pp_k(i) = Z(w^i) * chunk_0(i) ... chunk_k(i)
//   output columns (values on the subgroup, column-major): Z of every challenge, then pp_0..pp_{np-1} of challenge 0,
//   of challenge 1, ...  (np = number of chunks - 1) -- the order of upstream's zs_partial_products batch.
//
// Three kernels: (1) one thread per (row, challenge) computes the chunk quotients with ONE field inversion (Montgomery's
// trick over the chunk denominators; upstream inverts every wire's denominator) and the product of its 128-row block,
// (2) an exclusive prefix product over the block products (one 1024-thread block per challenge; n/128 elements),
// (3) one thread per (row, challenge) finishes the prefix product inside its block (shared-memory scan) = Z on the
// subgroup, and expands it into the partial products.  All values are exact field elements, so any order of
// multiplications gives upstream's bits.
#include "internal.cuh"
#include "goldilocks.cuh"
#include "ntt_local.cuh"

namespace qpzk {

constexpr int PP_THREADS = 128;
constexpr int PP_SCAN_THREADS = 1024;

// inclusive prefix product over the PP_THREADS values of a block; returns this thread's inclusive product and leaves
// the block total in sh[PP_THREADS - 1]
__device__ __forceinline__ uint64_t block_scan_mul(uint64_t v, uint64_t* sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
#pragma unroll
    for (int d = 1; d < PP_THREADS; d <<= 1) {
        const uint64_t t = threadIdx.x >= d ? gl_mul(sh[threadIdx.x - d], v) : v;
        __syncthreads();
        sh[threadIdx.x] = v = t;
        __syncthreads();
    }
    return v;
}

__global__ void __launch_bounds__(PP_THREADS) pp_chunk_kernel(PartialProductsArgs a) {
    __shared__ uint64_t sh[PP_THREADS];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n = 1ull << a.log_n;
    const uint32_t c = blockIdx.y;
    uint64_t row = 1;
    if (i < n) {
        const uint64_t x = tw_get(a.tw_lo, a.tw_hi, i);
        const uint64_t beta = a.betas[c], gamma = a.gammas[c];
        const uint64_t* bk = a.beta_k + (size_t)c * a.num_routed;       // beta * k_j
        uint64_t* cp = a.chunkprod + (size_t)c * a.nchunks * n + i;      // [challenge][chunk][row]
        uint64_t num[PP_MAX_CHUNKS], pre[PP_MAX_CHUNKS], den[PP_MAX_CHUNKS];
        uint64_t all_den = 1;
        for (uint32_t k = 0; k < a.nchunks; k++) {
            uint64_t pn = 1, pd = 1;
            const uint32_t j1 = min(a.num_routed, (k + 1) * a.chunk);
            for (uint32_t j = k * a.chunk; j < j1; j++) {
                const uint64_t wv = add_canon_p(a.wires[(uint64_t)j * a.col_stride + i], gamma);
                const uint64_t sg = a.sigmas[(uint64_t)j * a.col_stride + i];
                pn = mulp(pn, mul_addp(bk[j], x, wv));            // arbitrary u64 representatives inside the products
                pd = mulp(pd, mul_addp(beta, sg, wv));
            }
            pd = gl_canon(pd);                                    // a zero denominator must be the word 0 for gl_inv
            num[k] = pn; den[k] = pd; pre[k] = all_den;
            all_den = gl_mul(all_den, pd);
        }
        uint64_t inv = gl_inv(all_den);
        for (uint32_t k = a.nchunks; k-- > 0;) {          // 1/den[k] = inv(all) * prefix(k) * suffix(k)
            const uint64_t q = gl_mul(num[k], gl_mul(inv, pre[k]));
            inv = gl_mul(inv, den[k]);
            cp[(uint64_t)k * n] = q;
            row = gl_mul(row, q);
        }
        a.rowprod[(size_t)c * n + i] = row;
    }
    block_scan_mul(row, sh);
    if (threadIdx.x == 0) a.blockprod[(size_t)c * gridDim.x + blockIdx.x] = sh[PP_THREADS - 1];
}

// blockprod[c][b] <- prod_{b' < b} blockprod[c][b']   (exclusive prefix product over the row blocks)
__global__ void __launch_bounds__(PP_SCAN_THREADS) pp_scan_kernel(PartialProductsArgs a, uint32_t num_blocks) {
    __shared__ uint64_t part[PP_SCAN_THREADS];
    uint64_t* t = a.blockprod + (size_t)blockIdx.x * num_blocks;
    const uint32_t seg = (num_blocks + PP_SCAN_THREADS - 1) / PP_SCAN_THREADS;
    const uint32_t i0 = min(num_blocks, threadIdx.x * seg), i1 = min(num_blocks, i0 + seg);
    uint64_t p = 1;
    for (uint32_t i = i0; i < i1; i++) p = gl_mul(p, t[i]);
    part[threadIdx.x] = p;
    __syncthreads();
    for (uint32_t d = 1; d < PP_SCAN_THREADS; d <<= 1) {   // inclusive scan of the segment products
        const uint64_t v = threadIdx.x >= d ? gl_mul(part[threadIdx.x - d], part[threadIdx.x]) : part[threadIdx.x];
        __syncthreads();
        part[threadIdx.x] = v;
        __syncthreads();
    }
    uint64_t run = threadIdx.x ? part[threadIdx.x - 1] : 1;
    for (uint32_t i = i0; i < i1; i++) {
        const uint64_t v = t[i];
        t[i] = run;
        run = gl_mul(run, v);
    }
}

__global__ void __launch_bounds__(PP_THREADS) pp_finish_kernel(PartialProductsArgs a) {
    __shared__ uint64_t sh[PP_THREADS];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n = 1ull << a.log_n;
    const uint32_t c = blockIdx.y, np = a.nchunks - 1;
    const uint64_t mine = i < n ? a.rowprod[(size_t)c * n + i] : 1;
    const uint64_t incl = block_scan_mul(mine, sh);
    // exclusive product of the rows before this one: block prefix * (inclusive product of the previous thread)
    const uint64_t before = threadIdx.x ? sh[threadIdx.x - 1] : 1;
    (void)incl;
    if (i >= n) return;
    uint64_t acc = gl_mul(a.blockprod[(size_t)c * gridDim.x + blockIdx.x], before);
    a.out[(uint64_t)c * n + i] = acc;
    const uint64_t* cp = a.chunkprod + (size_t)c * a.nchunks * n + i;
    uint64_t* pp = a.out + ((uint64_t)a.num_challenges + (uint64_t)c * np) * n + i;
    for (uint32_t k = 0; k < np; k++) {
        acc = gl_mul(acc, cp[(uint64_t)k * n]);
        pp[(uint64_t)k * n] = acc;
    }
}

__global__ void perm_params_kernel(const uint64_t* __restrict__ betas, const uint64_t* __restrict__ kpow, uint32_t num_challenges,
                                   uint32_t num_routed, uint64_t* __restrict__ beta_k) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_challenges * num_routed) return;
    beta_k[t] = gl_mul(betas[t / num_routed], kpow[t % num_routed]);
}

__global__ void base_powers_kernel(const uint64_t* __restrict__ bases, uint32_t num_challenges, uint32_t num_terms,
                                   uint64_t* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_challenges * num_terms) return;
    out[t] = gl_exp(bases[t / num_terms], t % num_terms);
}

cudaError_t launch_perm_params(const uint64_t* betas_dev, const uint64_t* kpow_dev, uint32_t num_challenges,
                               uint32_t num_routed, uint64_t* beta_k_dev, cudaStream_t st) {
    const uint32_t total = num_challenges * num_routed;
    if (!total) return cudaSuccess;
    perm_params_kernel<<<(total + 127) / 128, 128, 0, st>>>(betas_dev, kpow_dev, num_challenges, num_routed, beta_k_dev);
    return cudaGetLastError();
}

cudaError_t launch_base_powers(const uint64_t* alphas_dev, uint32_t num_challenges, uint32_t num_terms,
                               uint64_t* apow_dev, cudaStream_t st) {
    const uint32_t total = num_challenges * num_terms;
    if (!total) return cudaSuccess;
    base_powers_kernel<<<(total + 127) / 128, 128, 0, st>>>(alphas_dev, num_challenges, num_terms, apow_dev);
    return cudaGetLastError();
}

cudaError_t launch_partial_products(const PartialProductsArgs& a, cudaStream_t st) {
    const uint64_t n = 1ull << a.log_n;
    const unsigned blocks = pp_num_blocks(a.log_n);
    (void)n;
    const dim3 grid(blocks, a.num_challenges);
    pp_chunk_kernel<<<grid, PP_THREADS, 0, st>>>(a);
    pp_scan_kernel<<<a.num_challenges, PP_SCAN_THREADS, 0, st>>>(a, blocks);
    pp_finish_kernel<<<grid, PP_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace qpzk
