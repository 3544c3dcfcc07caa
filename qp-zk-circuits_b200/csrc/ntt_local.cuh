// Device helpers shared by ntt.cu and fri.cu: two-level root-power lookup and the shared-memory radix-2
// transform of a 2^s x W tile.
#pragma once
#include "internal.cuh"
#include "goldilocks.cuh"

namespace qpzk {

__device__ __forceinline__ uint64_t tw_get(const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi,
                                           uint64_t idx) {
    uint64_t v = lo[idx & (TW_LO_SIZE - 1)];
    uint64_t h = idx >> TW_LO_BITS;
    if (h) v = gl_mul(v, hi[h]);
    return v;
}

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? __brev(x) >> (32 - bits) : 0; }

// local radix-2 transform of size S = 2^s over index e of a tile sm[e * W + w]
template <bool DIT>
__device__ __forceinline__ void local_ntt(uint64_t* sm, const uint64_t* ltw, uint32_t s, uint32_t W) {
    uint32_t S = 1u << s;
    uint32_t nbf = (S >> 1) * W;
    if (DIT) {
        for (uint32_t lh = 0; lh < s; lh++) {
            uint32_t h = 1u << lh;
            for (uint32_t b = threadIdx.x; b < nbf; b += blockDim.x) {
                uint32_t w = b % W, t = b / W;
                uint32_t j = t & (h - 1), e0 = ((t >> lh) << (lh + 1)) + j, e1 = e0 + h;
                uint64_t tw = ltw[j << (s - 1 - lh)];
                uint64_t a = sm[e0 * W + w], x = gl_mul(sm[e1 * W + w], tw);
                sm[e0 * W + w] = gl_add(a, x);
                sm[e1 * W + w] = gl_sub(a, x);
            }
            __syncthreads();
        }
    } else {
        for (int lh = (int)s - 1; lh >= 0; lh--) {
            uint32_t h = 1u << lh;
            for (uint32_t b = threadIdx.x; b < nbf; b += blockDim.x) {
                uint32_t w = b % W, t = b / W;
                uint32_t j = t & (h - 1), e0 = ((t >> lh) << (lh + 1)) + j, e1 = e0 + h;
                uint64_t tw = ltw[j << (s - 1 - lh)];
                uint64_t a = sm[e0 * W + w], x = sm[e1 * W + w];
                sm[e0 * W + w] = gl_add(a, x);
                sm[e1 * W + w] = gl_mul(gl_sub(a, x), tw);
            }
            __syncthreads();
        }
    }
}

}  // namespace qpzk
