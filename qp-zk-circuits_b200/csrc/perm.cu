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
// Three kernels: (1) one thread per row computes the chunk quotients with ONE field inversion per row and challenge
// (Montgomery's trick over the chunk denominators; upstream inverts every wire's denominator), (2) an exclusive
// prefix product over the rows (the only sequential dependency; one 1024-thread block per challenge), (3) one thread
// per row expands Z into the partial products.  All values are exact field elements, so any order of
// multiplications gives upstream's bits.
#include "internal.cuh"
#include "goldilocks.cuh"
#include "ntt_local.cuh"

namespace qpzk {

constexpr int PP_THREADS = 128;
constexpr int PP_SCAN_THREADS = 1024;

__global__ void __launch_bounds__(PP_THREADS) pp_chunk_kernel(PartialProductsArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n = 1ull << a.log_n;
    if (i >= n) return;
    const uint64_t x = tw_get(a.tw_lo, a.tw_hi, i);
    for (uint32_t c = 0; c < a.num_challenges; c++) {
        const uint64_t beta = a.betas[c], gamma = a.gammas[c];
        const uint64_t* bk = a.beta_k + (size_t)c * a.num_routed;       // beta * k_j
        uint64_t* cp = a.chunkprod + (size_t)c * a.nchunks * n + i;      // [challenge][chunk][row]
        uint64_t num[PP_MAX_CHUNKS], pre[PP_MAX_CHUNKS], den[PP_MAX_CHUNKS];
        uint64_t all_den = 1;
        for (uint32_t k = 0; k < a.nchunks; k++) {
            uint64_t pn = 1, pd = 1;
            const uint32_t j1 = min(a.num_routed, (k + 1) * a.chunk);
            for (uint32_t j = k * a.chunk; j < j1; j++) {
                const uint64_t wv = a.wires[(uint64_t)j * a.col_stride + i];
                const uint64_t sg = a.sigmas[(uint64_t)j * a.col_stride + i];
                pn = gl_mul(pn, gl_add(gl_add(wv, gl_mul(bk[j], x)), gamma));
                pd = gl_mul(pd, gl_add(gl_add(wv, gl_mul(beta, sg)), gamma));
            }
            num[k] = pn; den[k] = pd; pre[k] = all_den;
            all_den = gl_mul(all_den, pd);
        }
        uint64_t inv = gl_inv(all_den), row = 1;
        for (uint32_t k = a.nchunks; k-- > 0;) {          // 1/den[k] = inv(all) * prefix(k) * suffix(k)
            const uint64_t q = gl_mul(num[k], gl_mul(inv, pre[k]));
            inv = gl_mul(inv, den[k]);
            cp[(uint64_t)k * n] = q;
            row = gl_mul(row, q);
        }
        a.rowprod[(size_t)c * n + i] = row;
    }
}

// rowprod[c][i] <- prod_{i' < i} rowprod[c][i']   (exclusive prefix product = Z on the subgroup)
__global__ void __launch_bounds__(PP_SCAN_THREADS) pp_scan_kernel(PartialProductsArgs a) {
    __shared__ uint64_t part[PP_SCAN_THREADS];
    const uint64_t n = 1ull << a.log_n;
    uint64_t* t = a.rowprod + (size_t)blockIdx.x * n;
    const uint64_t seg = (n + PP_SCAN_THREADS - 1) / PP_SCAN_THREADS;
    const uint64_t i0 = min(n, (uint64_t)threadIdx.x * seg), i1 = min(n, i0 + seg);
    uint64_t p = 1;
    for (uint64_t i = i0; i < i1; i++) p = gl_mul(p, t[i]);
    part[threadIdx.x] = p;
    __syncthreads();
    for (uint32_t d = 1; d < PP_SCAN_THREADS; d <<= 1) {   // inclusive scan of the segment products
        const uint64_t v = threadIdx.x >= d ? gl_mul(part[threadIdx.x - d], part[threadIdx.x]) : part[threadIdx.x];
        __syncthreads();
        part[threadIdx.x] = v;
        __syncthreads();
    }
    uint64_t run = threadIdx.x ? part[threadIdx.x - 1] : 1;
    for (uint64_t i = i0; i < i1; i++) {
        const uint64_t v = t[i];
        t[i] = run;
        run = gl_mul(run, v);
    }
}

__global__ void __launch_bounds__(PP_THREADS) pp_finish_kernel(PartialProductsArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n = 1ull << a.log_n;
    if (i >= n) return;
    const uint32_t np = a.nchunks - 1;
    for (uint32_t c = 0; c < a.num_challenges; c++) {
        uint64_t acc = a.rowprod[(size_t)c * n + i];
        a.out[(uint64_t)c * n + i] = acc;
        const uint64_t* cp = a.chunkprod + (size_t)c * a.nchunks * n + i;
        uint64_t* pp = a.out + ((uint64_t)a.num_challenges + (uint64_t)c * np) * n + i;
        for (uint32_t k = 0; k < np; k++) {
            acc = gl_mul(acc, cp[(uint64_t)k * n]);
            pp[(uint64_t)k * n] = acc;
        }
    }
}

cudaError_t launch_partial_products(const PartialProductsArgs& a, cudaStream_t st) {
    const uint64_t n = 1ull << a.log_n;
    const unsigned blocks = (unsigned)((n + PP_THREADS - 1) / PP_THREADS);
    pp_chunk_kernel<<<blocks, PP_THREADS, 0, st>>>(a);
    pp_scan_kernel<<<a.num_challenges, PP_SCAN_THREADS, 0, st>>>(a);
    pp_finish_kernel<<<blocks, PP_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace qpzk
