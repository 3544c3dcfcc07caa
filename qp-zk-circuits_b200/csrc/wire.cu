// Wire format: ProofWithPublicInputs::to_bytes / from_bytes for the flat proof pieces this library produces.
//
// Replaces qp-plonky2 =1.5.5 `util/serialization/mod.rs`: `write_proof_with_public_inputs` = `write_proof`
// (`write_merkle_cap` x3, `write_opening_set`, `write_fri_proof`) + `write_usize(len)` + `write_field_vec(public_inputs)`
// (the reference moves proofs around only as these bytes: wormhole/prover/src/lib.rs:171-175 returns
// the proof, wormhole/tests/src/prover/prover_tests.rs:118 and the aggregator CLI hand it on as `to_bytes()`; the field
// inventory is the one enumerated at wormhole/aggregator/src/common/utils.rs:349-537).  Host code only: no kernel, no
// device memory -- the inputs are the caps, openings and flat FRI proof words the other entry points return.
// PARITY UNPINNED: the byte order below is upstream's as recalled, no golden proof exists in
// the reference tree.
//
//   field element            u64 little endian (canonical)
//   extension element        2 field elements
//   hash                     4 field elements
//   merkle cap               2^cap_height hashes, no length prefix
//   opening set              constants | plonk_sigmas | wires | plonk_zs | plonk_zs_next | partial_products |
//                            quotient_polys | lookup_zs | next_lookup_zs      (extension elements, no length prefixes)
//   merkle proof             u8 number of siblings, then the siblings
//   fri proof                commit-phase caps; per query round: per oracle (leaf evals, merkle proof), per reduction
//                            step (coset evals, merkle proof); final polynomial coefficients; pow witness
//   public inputs            u64 little-endian count, then the field elements
#include <cstring>
#include <string>

#include "../../include/qpzk.h"

namespace {

constexpr uint64_t P = 0xFFFFFFFF00000001ULL;

struct Walker {      // walks the byte stream and the flat word arrays in lock step; `write` selects the direction
    bool write;
    uint8_t* bytes;
    size_t len, pos = 0;
    bool ok = true, canonical = true;
    void words(uint64_t* w, size_t n) {
        if (!ok || pos + 8 * n > len) { ok = false; return; }
        for (size_t i = 0; i < n; i++) {
            if (write) {
                uint64_t v = w[i];
                if (v >= P) canonical = false;
                for (int b = 0; b < 8; b++) bytes[pos + b] = (uint8_t)(v >> (8 * b));
            } else {
                uint64_t v = 0;
                for (int b = 0; b < 8; b++) v |= (uint64_t)bytes[pos + b] << (8 * b);
                if (v >= P) canonical = false;
                w[i] = v;
            }
            pos += 8;
        }
    }
    void u8(uint8_t expect) {
        if (!ok || pos + 1 > len) { ok = false; return; }
        if (write) bytes[pos] = expect; else if (bytes[pos] != expect) ok = false;
        pos++;
    }
    void u64len(uint64_t expect) {
        uint64_t v = expect;
        const bool c = canonical;
        words(&v, 1);
        canonical = c;
        if (!write && v != expect) ok = false;
    }
};

struct Layout {
    uint32_t H, R, log_L, final_len;
    uint32_t op_counts[7];   // constants, sigmas, wires, zs, partial products, quotient, (zs_next = zs)
};

bool make_layout(const qpzk_fri_params* p, const qpzk_proof_shape* s, Layout* l, std::string* why) {
    if (!p || !s) { *why = "null argument"; return false; }
    if (p->num_rounds > QPZK_MAX_FRI_ROUNDS || p->cap_height > 31 || s->log_n > 40) { *why = "bad parameters"; return false; }
    uint32_t sum = 0;
    for (uint32_t r = 0; r < p->num_rounds; r++) sum += p->arity_bits[r];
    if (sum > s->log_n) { *why = "reduction arities exceed the degree"; return false; }
    l->H = 1u << p->cap_height; l->R = p->num_rounds; l->log_L = s->log_n + p->rate_bits; l->final_len = 1u << (s->log_n - sum);
    if (p->cap_height > l->log_L) { *why = "cap_height exceeds log2(number of leaves)"; return false; }
    if (s->leaf_lens[0] < s->num_constants + s->num_sigmas || s->leaf_lens[1] < s->num_wires ||
        s->leaf_lens[2] < s->num_zs + s->num_partial_products || s->leaf_lens[3] < s->num_quotient_polys) {
        *why = "leaf_lens smaller than the opened polynomials of the oracle"; return false;
    }
    return true;
}

// one pass over the whole proof; the flat arrays are read (write) or filled (!write)
bool walk(Walker& w, const qpzk_fri_params* p, const qpzk_proof_shape* s, const Layout& l, uint64_t* wires_cap,
          uint64_t* zs_cap, uint64_t* quotient_cap, uint64_t* openings, uint64_t* fri, uint64_t* public_inputs) {
    w.words(wires_cap, 4ull * l.H);
    w.words(zs_cap, 4ull * l.H);
    w.words(quotient_cap, 4ull * l.H);
    // openings arrive in FRI-instance order: constants | sigmas | wires | zs | partial products | quotient | zs_next
    const uint64_t nc = s->num_constants, ns = s->num_sigmas, nw = s->num_wires, nz = s->num_zs, np = s->num_partial_products,
                   nq = s->num_quotient_polys;
    uint64_t* o = openings;
    w.words(o, 2 * (nc + ns + nw + nz));                 // constants, plonk_sigmas, wires, plonk_zs
    w.words(o + 2 * (nc + ns + nw + nz + np + nq), 2 * nz);   // plonk_zs_next
    w.words(o + 2 * (nc + ns + nw + nz), 2 * (np + nq));  // partial_products, quotient_polys; lookup vectors are empty
    // fri proof
    uint64_t* f = fri;
    w.words(f, 4ull * l.H * l.R); f += 4ull * l.H * l.R;
    for (uint32_t q = 0; q < p->num_query_rounds; q++) {
        for (int t = 0; t < 4; t++) {
            w.words(f, s->leaf_lens[t]); f += s->leaf_lens[t];
            const uint32_t sib = l.log_L - p->cap_height;
            w.u8((uint8_t)sib);
            w.words(f, 4ull * sib); f += 4ull * sib;
        }
        uint32_t lg = l.log_L;
        for (uint32_t r = 0; r < l.R; r++) {
            const uint32_t ab = p->arity_bits[r];
            w.words(f, 2ull << ab); f += 2ull << ab;
            const uint32_t leaves_log = lg - ab;
            const uint32_t sib = leaves_log > p->cap_height ? leaves_log - p->cap_height : 0;
            w.u8((uint8_t)sib);
            w.words(f, 4ull * sib); f += 4ull * sib;
            lg -= ab;
        }
    }
    w.words(f, 2ull * l.final_len); f += 2ull * l.final_len;
    w.words(f, 1);                                            // pow_witness
    w.u64len(s->num_public_inputs);
    w.words(public_inputs, s->num_public_inputs);
    return w.ok;
}

thread_local std::string g_wire_err;

}  // namespace

// defined in capi.cu
extern "C" const char* qpzk_last_error(void);
namespace qpzk { int wire_fail(int code, const std::string& msg); }

extern "C" size_t qpzk_proof_bytes_len(const qpzk_fri_params* p, const qpzk_proof_shape* s) {
    Layout l;
    std::string why;
    if (!make_layout(p, s, &l, &why)) { qpzk::wire_fail(QPZK_EINVAL, why); return 0; }
    size_t words = 3ull * 4 * l.H + 2ull * (s->num_constants + s->num_sigmas + s->num_wires + 2ull * s->num_zs +
                                            s->num_partial_products + s->num_quotient_polys);
    size_t bytes = 0;
    words += 4ull * l.H * l.R;
    size_t per_q = 0;
    for (int t = 0; t < 4; t++) { per_q += s->leaf_lens[t] + 4ull * (l.log_L - p->cap_height); bytes += p->num_query_rounds; }
    uint32_t lg = l.log_L;
    for (uint32_t r = 0; r < l.R; r++) {
        const uint32_t ab = p->arity_bits[r], leaves_log = lg - ab;
        per_q += (2ull << ab) + 4ull * (leaves_log > p->cap_height ? leaves_log - p->cap_height : 0);
        bytes += p->num_query_rounds;
        lg -= ab;
    }
    words += per_q * p->num_query_rounds + 2ull * l.final_len + 1 + 1 + s->num_public_inputs;
    return bytes + 8 * words;
}

extern "C" int qpzk_proof_to_bytes(const qpzk_fri_params* p, const qpzk_proof_shape* s, const uint64_t* wires_cap,
                                   const uint64_t* zs_partial_products_cap, const uint64_t* quotient_polys_cap,
                                   const uint64_t* openings, const uint64_t* fri_proof, const uint64_t* public_inputs,
                                   uint8_t* out, size_t out_len) {
    Layout l;
    std::string why;
    if (!make_layout(p, s, &l, &why)) return qpzk::wire_fail(QPZK_EINVAL, why);
    if (!wires_cap || !zs_partial_products_cap || !quotient_polys_cap || !openings || !fri_proof || !out ||
        (!public_inputs && s->num_public_inputs))
        return qpzk::wire_fail(QPZK_EINVAL, "null argument");
    if (out_len != qpzk_proof_bytes_len(p, s)) return qpzk::wire_fail(QPZK_EINVAL, "out_len is not qpzk_proof_bytes_len()");
    Walker w{true, out, out_len};
    const bool ok = walk(w, p, s, l, const_cast<uint64_t*>(wires_cap), const_cast<uint64_t*>(zs_partial_products_cap),
                         const_cast<uint64_t*>(quotient_polys_cap), const_cast<uint64_t*>(openings),
                         const_cast<uint64_t*>(fri_proof), const_cast<uint64_t*>(public_inputs));
    if (!ok || w.pos != out_len) return qpzk::wire_fail(QPZK_EINTERNAL, "serializer length mismatch");
    if (!w.canonical) return qpzk::wire_fail(QPZK_EINVAL, "a field element is not canonical");
    return 0;
}

extern "C" int qpzk_proof_from_bytes(const qpzk_fri_params* p, const qpzk_proof_shape* s, const uint8_t* bytes, size_t len,
                                     uint64_t* wires_cap, uint64_t* zs_partial_products_cap, uint64_t* quotient_polys_cap,
                                     uint64_t* openings, uint64_t* fri_proof, uint64_t* public_inputs) {
    Layout l;
    std::string why;
    if (!make_layout(p, s, &l, &why)) return qpzk::wire_fail(QPZK_EINVAL, why);
    if (!bytes || !wires_cap || !zs_partial_products_cap || !quotient_polys_cap || !openings || !fri_proof ||
        (!public_inputs && s->num_public_inputs))
        return qpzk::wire_fail(QPZK_EINVAL, "null argument");
    if (len != qpzk_proof_bytes_len(p, s)) return qpzk::wire_fail(QPZK_EINVAL, "proof has the wrong length for this circuit");
    Walker w{false, const_cast<uint8_t*>(bytes), len};
    const bool ok = walk(w, p, s, l, wires_cap, zs_partial_products_cap, quotient_polys_cap, openings, fri_proof, public_inputs);
    if (!ok || w.pos != len) return qpzk::wire_fail(QPZK_EINVAL, "malformed proof (a length field does not match the circuit)");
    if (!w.canonical) return qpzk::wire_fail(QPZK_EINVAL, "a field element is not canonical");
    return 0;
}
