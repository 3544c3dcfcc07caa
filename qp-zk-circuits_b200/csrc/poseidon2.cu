// Poseidon2 over Goldilocks (width 12, x^7, 4 + 22 + 4 rounds): the reference's APPLICATION hash `Poseidon2Hash`
// (qp-plonky2 =1.5.5 `hash::poseidon2`, qp-poseidon-core =3.1.0; reference Cargo.toml:20-21), used for addresses,
// nullifiers, zk-tree nodes and block hashes (wormhole/circuit/src/unspendable_account.rs:63-95,
// wormhole/circuit/src/nullifier.rs:298-299, wormhole/circuit/src/zk_merkle_proof.rs:504,606,
// wormhole/circuit/src/block_header/header.rs:132-141) -- SURVEY.md 8f-2 / 8c.  It is NOT the FRI / Merkle hasher of
// the commitment path (that is Poseidon, hash.cu).  PARITY PINNED: the parameters (poseidon2_rc.cuh, generated and
// checked by tools/gen_poseidon2_constants.py) reproduce the seven vectors the reference's own tests hold
// (wormhole/tests/src/circuit/unspendable_account_tests.rs:9-24, wormhole/tests/test-helpers/src/lib.rs:208-219).
//
// One thread = one permutation / one message, state in registers.  The external layer circ(2 M4, M4, M4) is additions
// only (M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]]: row r = sum + x_r + 2 x_{r+1}); the internal layer
// s <- diag * s + sum(s) is twelve full-width multiplications with the addend folded into the 128-bit product.
#include "internal.cuh"
#include "goldilocks.cuh"
#include "poseidon2_rc.cuh"

namespace qpzk {

__constant__ uint64_t c_p2_ext[96];
__constant__ uint64_t c_p2_int[22];
__constant__ uint64_t c_p2_diag[12];

cudaError_t upload_poseidon2_constants() {
    cudaError_t e = cudaMemcpyToSymbol(c_p2_ext, QPZK_POSEIDON2_EXT_RC_HOST, sizeof c_p2_ext);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_p2_int, QPZK_POSEIDON2_INT_RC_HOST, sizeof c_p2_int);
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_p2_diag, QPZK_POSEIDON2_DIAG_HOST, sizeof c_p2_diag);
    return e;
}

__device__ __forceinline__ uint64_t p2_sbox(uint64_t x) {   // canonical in, canonical out
    const uint64_t x2 = sqrp(x), x4 = sqrp(x2), x3 = mulp(x, x2);
    return gl_canon(mulp(x3, x4));
}

__device__ __forceinline__ void p2_external_layer(uint64_t (&s)[12]) {
    uint64_t t[12];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const uint64_t a = s[4 * c], b = s[4 * c + 1], d = s[4 * c + 2], e = s[4 * c + 3];
        const uint64_t sum = gl_add(gl_add(a, b), gl_add(d, e));
        t[4 * c + 0] = gl_add(gl_add(sum, a), gl_add(b, b));
        t[4 * c + 1] = gl_add(gl_add(sum, b), gl_add(d, d));
        t[4 * c + 2] = gl_add(gl_add(sum, d), gl_add(e, e));
        t[4 * c + 3] = gl_add(gl_add(sum, e), gl_add(a, a));
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint64_t col = gl_add(gl_add(t[k], t[4 + k]), t[8 + k]);
#pragma unroll
        for (int c = 0; c < 3; c++) s[4 * c + k] = gl_add(t[4 * c + k], col);
    }
}

// canonical in, canonical out
__device__ __forceinline__ void poseidon2_permute(uint64_t (&s)[12]) {
    p2_external_layer(s);
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        if (r == 4) {
#pragma unroll 1
            for (int q = 0; q < 22; q++) {
                s[0] = p2_sbox(gl_add(s[0], c_p2_int[q]));
                uint64_t sum = gl_add(s[0], s[1]);
#pragma unroll
                for (int i = 2; i < 12; i++) sum = gl_add(sum, s[i]);
#pragma unroll
                for (int i = 0; i < 12; i++) s[i] = gl_canon(mul_addp(s[i], c_p2_diag[i], sum));
            }
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = p2_sbox(gl_add(s[i], c_p2_ext[12 * r + i]));
        p2_external_layer(s);
    }
}

__global__ void __launch_bounds__(HASH_THREADS)
poseidon2_permute_kernel(const uint64_t* __restrict__ in, uint64_t count, uint64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = gl_canon(in[12 * i + k]);
    poseidon2_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[12 * i + k] = s[k];
}

// Poseidon2Hash::hash_no_pad of `count` messages of `len` words each (row stride `stride`): additive sponge of rate 8,
// padding x || 1 || 0* to a whole block, first four state words out.
__global__ void __launch_bounds__(HASH_THREADS)
poseidon2_hash_kernel(const uint64_t* __restrict__ msgs, uint64_t count, uint32_t len, uint64_t stride,
                      uint64_t* __restrict__ out) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count) return;
    const uint64_t* p = msgs + m * stride;
    uint64_t s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    const uint32_t padded = (len + 8) & ~7u;   // len + 1 rounded up to a multiple of 8
    for (uint32_t off = 0; off < padded; off += 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t k = off + i;
            const uint64_t v = k < len ? gl_canon(__ldg(p + k)) : (k == len ? 1ull : 0ull);
            s[i] = gl_add(s[i], v);
        }
        poseidon2_permute(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[4 * m + i] = s[i];
}

cudaError_t launch_poseidon2_permute(const uint64_t* in, uint64_t count, uint64_t* out, cudaStream_t st) {
    if (!count) return cudaSuccess;
    poseidon2_permute_kernel<<<(unsigned)((count + HASH_THREADS - 1) / HASH_THREADS), HASH_THREADS, 0, st>>>(in, count, out);
    return cudaGetLastError();
}

cudaError_t launch_poseidon2_hash(const uint64_t* msgs, uint64_t count, uint32_t len, uint64_t stride, uint64_t* out,
                                  cudaStream_t st) {
    if (!count) return cudaSuccess;
    poseidon2_hash_kernel<<<(unsigned)((count + HASH_THREADS - 1) / HASH_THREADS), HASH_THREADS, 0, st>>>(msgs, count, len, stride, out);
    return cudaGetLastError();
}

}  // namespace qpzk
