"""plonky2's `prove()` after witness generation, driven through the C ABI.

`prove_with_partition_witness` (qp-plonky2 =1.5.5; reached from the reference at wormhole/prover/src/lib.rs:171-175)
does, per proof (transcript in Appendix A.7):

    wires        -> PolynomialBatch::from_values -> observe cap -> betas, gammas
    zs + partial products -> from_values -> observe cap -> alphas
    quotient chunks (coefficients) -> from_coeffs -> observe cap -> zeta
    openings of all four oracles at zeta, of zs at g*zeta -> observe -> PolynomialBatch::prove_openings (FRI)

Two ways to use this module:

* `prove_synthetic(circuit, wires, zs_partial_products, quotient_coeffs, ...)` with caller-supplied (random) columns for
  the zs/partial products and the quotient: exactly the commitment / opening half of `prove()`, the part measured as
  the *leaf-shaped synthetic prove* (BASELINE.json configs 1/2 stand-in).
* with `zs_partial_products=None` / `quotient_coeffs=None` those two steps run on the device as well: the partial
  products from the routed wires and the circuit's sigma columns (`PolynomialBatch.commit_zs_partial_products`, row a16),
  the quotient from the three committed batches (`commit_quotient_prog`, row a17) with the circuit's gate constraints
  handed over as a `GateProgram` (qpzk/gates.py records them; None = a circuit without gates).  `prove()` below wraps
  this into "circuit data + witness + public inputs -> ProofWithPublicInputs bytes".  What stays outside is what is
  specific to the reference's circuits: witness generation and the Rust-side recording of its gate set.

The constants/sigmas oracle is committed once per circuit (`CircuitBuilder::build`, reference
wormhole/circuit/src/circuit.rs:206-216), not per proof.

Shapes of the leaf circuit: degree 2^13, 85 constants+sigmas, 135 wires, 20 zs/partial
products, 16 quotient chunks; FRI rate_bits 3, cap_height 4, 16 PoW bits, 28 queries, arity 4, final poly <= 2^5
(reference wormhole/circuit/src/profile.rs:168-174).
"""
import numpy as np

from . import (GENERATOR, P, Challenger, FriInstanceInfo, PolynomialBatch, Prover, fri_params, hash_no_pad,  # noqa: F401
               proof_shape, proof_to_bytes)

LEAF_SHAPE = dict(log_n=13, constants_sigmas=85, wires=135, zs_partial_products=20, quotient=16, num_zs=2)
RATE_BITS, CAP_HEIGHT, POW_BITS, NUM_QUERIES, ARITY_BITS, FINAL_POLY_BITS = 3, 4, 16, 28, 4, 5


def _root_of_unity(log_n):
    w = 7277203076849721926  # order 2^32
    for _ in range(32 - log_n):
        w = w * w % P
    return w


class SyntheticCircuit:
    """What `CircuitBuilder::build` leaves behind for the prover: the committed constants/sigmas oracle."""

    def __init__(self, ctx, constants_sigmas, shape=LEAF_SHAPE, sigma_col0=None):
        self.ctx, self.shape = ctx, dict(shape)
        self.sigma_col0 = sigma_col0     # first sigma column inside constants_sigmas (needed for the device quotient)
        if sigma_col0 is not None:       # the permutation argument reads the sigma VALUES from this batch
            ctx.set_keep_values(True)
        self.log_n = shape["log_n"]
        self.constants_sigmas = PolynomialBatch.from_values(ctx, constants_sigmas, RATE_BITS, None, CAP_HEIGHT)
        ctx.set_keep_values(False)
        self.circuit_digest = np.arange(4, dtype=np.uint64) + 11   # stands for the verifier-data digest
        self.fri_params = fri_params(self.log_n, RATE_BITS, CAP_HEIGHT, POW_BITS, NUM_QUERIES, ARITY_BITS, FINAL_POLY_BITS)
        self._provers = {}

    def prover(self, num_wires, num_routed, num_public_inputs, gate_program=None, quotient_degree_factor=8, blinding=False,
               use_graph=True):
        """The circuit's qpzk.Prover (created once per shape and gate program, like a prover's circuit data)."""
        key = (num_wires, num_routed, num_public_inputs, id(gate_program), quotient_degree_factor, blinding, use_graph)
        if key not in self._provers:
            self._provers[key] = Prover(self.constants_sigmas, num_wires, num_routed, self.sigma_col0, self.fri_params,
                                        self.circuit_digest, num_public_inputs, gate_program, quotient_degree_factor,
                                        self.shape["num_zs"], blinding, use_graph)
        return self._provers[key]


def prove_synthetic(circuit, wires, zs_partial_products, quotient_coeffs, public_inputs_hash, pow_witness=None,
                    keep_batches=False, sigmas=None, quotient_degree_factor=8, gate_program=None):
    """One proof's worth of commitments + FRI.  Inputs are host arrays (values / coefficients, column-major).
    zs_partial_products may be None when `sigmas` (num_routed x n sigma values) is given: the columns are then
    computed from wires[:num_routed] with the transcript's betas and gammas (upstream
    `all_wires_permutation_partial_products`).
    Returns dict(caps=[wires, zs_pp, quotient], openings=(k, 2) array, fri_proof=flat words)."""
    ctx, sh = circuit.ctx, circuit.shape
    ch = Challenger(ctx)
    ch.observe_hash(circuit.circuit_digest)
    ch.observe_hash(public_inputs_hash)
    on_device = zs_partial_products is None and circuit.sigma_col0 is not None
    ctx.set_keep_values(on_device)                             # the wires batch keeps its values for the partial products
    wires_b = PolynomialBatch.from_values(ctx, wires, RATE_BITS, None, CAP_HEIGHT)
    ctx.set_keep_values(False)
    ch.observe_cap(wires_b)
    nz = sh["num_zs"]
    betas_gammas = ch.get_n_challenges(2 * nz)                 # betas, then gammas (num_challenges each)
    if on_device:                                              # wires and sigmas never leave the device
        zs_b = PolynomialBatch.commit_zs_partial_products(circuit.constants_sigmas, wires_b, circuit.sigma_col0,
                                                          np.asarray(sigmas).shape[0], betas_gammas[:nz], betas_gammas[nz:],
                                                          CAP_HEIGHT, quotient_degree_factor)
    else:
        if zs_partial_products is None:
            sg = np.asarray(sigmas, dtype=np.uint64)
            zs_partial_products = ctx.partial_products(np.asarray(wires)[:sg.shape[0]], sg, betas_gammas[:nz],
                                                       betas_gammas[nz:], quotient_degree_factor)
        zs_b = PolynomialBatch.from_values(ctx, zs_partial_products, RATE_BITS, None, CAP_HEIGHT)
    ch.observe_cap(zs_b)
    alphas = ch.get_n_challenges(nz)
    if quotient_coeffs is None:                                # permutation argument only (no gate terms)
        num_routed = np.asarray(sigmas).shape[0]
        # gate_program (a qpzk.GateProgram): the circuit's gate constraints as data, interpreted on the device
        quot_b = PolynomialBatch.commit_quotient_prog(circuit.constants_sigmas, wires_b, zs_b, circuit.sigma_col0, num_routed,
                                                      betas_gammas[:nz], betas_gammas[nz:], alphas, gate_program, CAP_HEIGHT,
                                                      quotient_degree_factor)
    else:
        quot_b = PolynomialBatch.from_coeffs(ctx, quotient_coeffs, RATE_BITS, None, CAP_HEIGHT)
    ch.observe_cap(quot_b)
    zeta = ch.get_extension_challenge()
    g = _root_of_unity(circuit.log_n)
    g_zeta = [int(zeta[0]) * g % P, int(zeta[1]) * g % P]
    oracles = [circuit.constants_sigmas, wires_b, zs_b, quot_b]
    all_polys = [(o, i) for o, b in enumerate(oracles) for i in range(b.num_cols)]
    zs_next = [(2, i) for i in range(sh["num_zs"])]
    inst = FriInstanceInfo([([int(zeta[0]), int(zeta[1])], all_polys), (g_zeta, zs_next)])
    openings = np.concatenate([PolynomialBatch.eval_ext_many(oracles, zeta), zs_b.eval_ext(g_zeta)[:sh["num_zs"]]])
    ch.observe_extension_elements(openings)
    proof = PolynomialBatch.prove_openings(inst, oracles, ch, circuit.fri_params, pow_witness=pow_witness)
    out = dict(caps=[wires_b.cap, zs_b.cap, quot_b.cap], openings=openings, fri_proof=proof,
               challenges=(betas_gammas, alphas, zeta))
    if keep_batches:
        out["batches"] = (wires_b, zs_b, quot_b)
    else:
        for b in (wires_b, zs_b, quot_b):
            b.free()
    ch.free()
    return out


def synthetic_inputs(seed, shape=LEAF_SHAPE):
    """Uniform canonical columns of the leaf shape (deterministic in `seed`)."""
    rng = np.random.default_rng(seed)
    n = 1 << shape["log_n"]
    mk = lambda c: rng.integers(0, P, size=(c, n), dtype=np.uint64)
    return dict(wires=mk(shape["wires"]), zs_partial_products=mk(shape["zs_partial_products"]),
                quotient_coeffs=mk(shape["quotient"]), public_inputs_hash=rng.integers(0, P, size=4, dtype=np.uint64))


def prove(circuit, wires, public_inputs, num_routed, gate_program=None, quotient_degree_factor=8, pow_witness=None,
          step_by_step=False):
    """`prove()` for a circuit given as data -- the committed constants/sigmas batch (sigma columns from
    circuit.sigma_col0), the number of routed wires and the gate program -- and a witness (the wire values):
    public-inputs hash, wires commitment, partial products, quotient, openings, FRI, then
    `ProofWithPublicInputs::to_bytes`.  Returns (proof bytes, the qpzk_proof_shape they were written with).
    Default: ONE device-resident call (qpzk.Prover = qpzk_prover_prove: challenges stay on the device, CUDA graph from
    the third proof of a circuit); step_by_step=True drives the individual entry points instead (same bytes)."""
    ctx, sh = circuit.ctx, circuit.shape
    if step_by_step:
        pih = hash_no_pad(ctx, public_inputs)
        sig_dummy = np.zeros((num_routed, 0), np.uint64)            # only its row count is read on the device path
        r = prove_synthetic(circuit, wires, None, None, pih, pow_witness=pow_witness, sigmas=sig_dummy,
                            quotient_degree_factor=quotient_degree_factor, gate_program=gate_program)
    else:
        r = circuit.prover(np.asarray(wires).shape[0], num_routed, np.asarray(public_inputs).size, gate_program,
                           quotient_degree_factor).prove(wires, public_inputs, pow_witness=pow_witness)
    nz = sh["num_zs"]
    nchunks = (num_routed + quotient_degree_factor - 1) // quotient_degree_factor
    shape = proof_shape(circuit.log_n, circuit.sigma_col0, circuit.constants_sigmas.num_cols - circuit.sigma_col0,
                        np.asarray(wires).shape[0], nz, nz * (nchunks - 1), nz * quotient_degree_factor,
                        np.asarray(public_inputs).size)
    data = proof_to_bytes(circuit.fri_params, shape, r["caps"][0], r["caps"][1], r["caps"][2], r["openings"], r["fri_proof"],
                          public_inputs)
    return data, shape
