// Device helpers shared by ntt.cu and fri.cu: two-level root-power lookup and the shared-memory transform of a
// 2^s-point tile done in register groups of up to 3 radix-2 stages (radix-8 butterflies): one shared-memory
// round trip, one barrier and one index computation per 3 stages.
#pragma once
#include "internal.cuh"
#include "goldilocks.cuh"

namespace qpzk {

__device__ __forceinline__ uint64_t tw_get(const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi,
                                           uint64_t idx) {
    uint64_t v = lo[idx & (TW_LO_SIZE - 1)];
    uint64_t h = idx >> TW_LO_BITS;
    if (h) v = gl_mulc(v, hi[h]);
    return v;
}

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? __brev(x) >> (32 - bits) : 0; }

// Shared-memory addressing of tile element (e, w):  w * ws + pad(e) * es, where pad(e) = e + (e >> 3) when `padded`
// (one spare word every 8 keeps the stride-8 accesses of the smallest-stride group conflict free).
struct TileMap {
    uint32_t es, ws, padded;
    __device__ __forceinline__ uint32_t operator()(uint32_t e, uint32_t w) const {
        return w * ws + (padded ? e + (e >> 3) : e) * es;
    }
};
__host__ __device__ __forceinline__ uint32_t padded_len(uint32_t n) { return n + (n >> 3) + 1; }

// G consecutive radix-2 stages whose smallest stride is 2^lh, on 2^G elements held in registers.
// DIT: strides h, 2h, ...;  a' = a + tw*b, b' = a - tw*b.   DIF: strides ..., 2h, h;  a' = a + b, b' = (a - b)*tw.
// The twiddle of the pair at local position p with half-size H is w_{2H}^(p mod H) = ltw[(p mod H) << (s-1-log2 H)].
// W must be a power of two when w_fast (tile kernels); with !w_fast (all-columns kernel) W is arbitrary.
template <bool DIT, int G>
__device__ __forceinline__ void ntt_group(uint64_t* sm, const uint64_t* ltw, uint32_t s, uint32_t lh, uint32_t W,
                                          bool w_fast, TileMap map) {
    constexpr int R = 1 << G;
    const uint32_t h = 1u << lh;
    const uint32_t lgroups = s - G, groups = 1u << lgroups;
    const uint32_t total = groups * W;
    const uint32_t lW = 31 - __clz(W);
    for (uint32_t u = threadIdx.x; u < total; u += blockDim.x) {
        uint32_t w, q;
        if (w_fast) { w = u & (W - 1); q = u >> lW; } else { q = u & (groups - 1); w = u >> lgroups; }
        const uint32_t j = q & (h - 1), e0 = ((q >> lh) << (lh + G)) + j;
        uint64_t x[R];
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = sm[map(e0 + k * h, w)];
#pragma unroll
        for (int t = 0; t < G; t++) {
            const int half = DIT ? (1 << t) : (1 << (G - 1 - t));
            const uint32_t logH = lh + (DIT ? t : (G - 1 - t));
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (k & half) continue;
                const uint64_t tw = ltw[(j + (uint32_t)(k & (half - 1)) * h) << (s - 1 - logH)];
                if (DIT) {
                    uint64_t a = x[k], b = gl_mulc(x[k + half], tw);
                    x[k] = gl_add(a, b);
                    x[k + half] = gl_sub(a, b);
                } else {
                    uint64_t a = x[k], b = x[k + half];
                    x[k] = gl_add(a, b);
                    x[k + half] = gl_mulc(gl_sub(a, b), tw);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; k++) sm[map(e0 + k * h, w)] = x[k];
    }
    __syncthreads();
}

// local transform of size 2^s over index e for every w in [0, W).  DIT: bit-reversed in, natural out; DIF: natural
// in, bit-reversed out.  Values canonical in and out.
template <bool DIT>
__device__ __forceinline__ void local_ntt(uint64_t* sm, const uint64_t* ltw, uint32_t s, uint32_t W, bool w_fast,
                                          TileMap map) {
    if (s == 0) return;
    const uint32_t rem = s % 3;
    if (DIT) {
        uint32_t lh = 0;
        for (; lh + 3 <= s; lh += 3) ntt_group<true, 3>(sm, ltw, s, lh, W, w_fast, map);
        if (rem == 2) ntt_group<true, 2>(sm, ltw, s, lh, W, w_fast, map);
        if (rem == 1) ntt_group<true, 1>(sm, ltw, s, lh, W, w_fast, map);
    } else {
        int lh = (int)s;
        if (rem == 2) { lh -= 2; ntt_group<false, 2>(sm, ltw, s, lh, W, w_fast, map); }
        if (rem == 1) { lh -= 1; ntt_group<false, 1>(sm, ltw, s, lh, W, w_fast, map); }
        while (lh > 0) { lh -= 3; ntt_group<false, 3>(sm, ltw, s, lh, W, w_fast, map); }
    }
}

}  // namespace qpzk
