"""GPU tests of the single-call, device-resident prover (qpzk_prover_*, csrc/prove.cu):
bit-exact with the step-by-step entry points and with the CPU oracle, eagerly and as a replayed CUDA graph."""
import numpy as np
import pytest

from conftest import P, splitmix_columns

pytestmark = pytest.mark.gpu


def _same(a, b):
    return all((x == y).all() for x, y in zip(a["caps"], b["caps"])) and (a["openings"] == b["openings"]).all() \
        and (a["fri_proof"] == b["fri_proof"]).all()


def test_prover_matches_oracle_and_step_by_step_on_a_circuit_with_gates(ctx, qpzk, orc):
    """Arithmetic gates chained by copy constraints (the circuit of test_plonk_proof_with_gates_end_to_end): the one-call
    prover, the step-by-step calls and the oracle give the same caps, openings and FRI proof -- for the two eager
    proofs, for the proof during which the graph is captured, and for replays of the graph with other witnesses."""
    from qpzk import prover
    from perm_util import arithmetic_circuit, arithmetic_program, verify_gateless_openings
    log_n = 8
    wires, consts, sigmas = arithmetic_circuit(log_n, 12)
    b = arithmetic_program()
    code, imm = b.arrays()
    prog = qpzk.GateProgram.make(code, imm, b.next)
    shape = dict(prover.LEAF_SHAPE, log_n=log_n, wires=16, constants_sigmas=19, zs_partial_products=4, quotient=16)
    cs = np.concatenate([consts, sigmas])
    circuit = prover.SyntheticCircuit(ctx, cs, shape, sigma_col0=3)
    ocs = orc.PolynomialBatch.from_values(cs, 3, 4)
    rng = np.random.default_rng(3)
    pr = circuit.prover(16, 16, 6, prog)
    for it in range(6):
        w = wires.copy()
        if it in (2, 4):                                      # an unsatisfied witness is proved all the same
            w[5, 17] = (int(w[5, 17]) + 1 + it) % P
        if it == 5:
            w = rng.integers(0, P, size=w.shape, dtype=np.uint64)
        pi = rng.integers(0, P, size=6, dtype=np.uint64)
        got = pr.prove(w, pi)
        assert pr.uses_graph == (it >= 2)
        pih = orc.hash_no_pad(pi)
        assert (qpzk.hash_no_pad(ctx, pi) == pih).all()
        step = prover.prove_synthetic(circuit, w, None, None, pih, sigmas=sigmas, gate_program=prog)
        assert _same(got, step), it
        if it in (0, 2, 3):
            want = orc.prove_synthetic(ocs, w, None, None, pih, log_n, sigmas=sigmas, sigma_col0=3, gate_program=(code, imm))
            assert _same(got, want), it
            assert orc.fri_verify(log_n, want["leaf_lens"], want["all_caps"], want["instance"], got["openings"], want["params"],
                                  want["verifier_challenger"], got["fri_proof"]) == 0
            bg, alphas, zeta = step["challenges"]
            ok = verify_gateless_openings(log_n, got["openings"], 3, 16, 16, 2, 8, bg[:2], bg[2:], alphas, zeta, program=(code, imm))
            assert ok == (it != 2)
    assert 30 < pr.launches_per_proof < 120
    pr.free()


def test_prover_leaf_shape_poseidon_gates_graph_replay(ctx, qpzk):
    """The leaf shape (2^13 rows, 135 wires of which 80 routed, 85 constants+sigmas, PoseidonGate program on every row):
    eight proofs through the graph equal the step-by-step path; the steady state performs no device allocation."""
    import gen_poseidon_constants as gen
    from qpzk import prover, gates
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    prog = rec.program()
    log_n = 13
    cs = splitmix_columns(1, 85, 1 << log_n)
    circuit = prover.SyntheticCircuit(ctx, cs, prover.LEAF_SHAPE, sigma_col0=5)
    pr = circuit.prover(135, 80, 21, prog)
    sig_dummy = np.zeros((80, 0), np.uint64)
    allocs = None
    for it in range(8):
        w = splitmix_columns(100 + it % 3, 135, 1 << log_n)
        pi = splitmix_columns(50 + it, 1, 21)[0]
        got = pr.prove(w, pi)
        if it == 3:
            allocs = ctx.device_allocs()
        if it in (0, 2, 5, 7):
            step = prover.prove_synthetic(circuit, w, None, None, qpzk.hash_no_pad(ctx, pi), sigmas=sig_dummy, gate_program=prog)
            assert _same(got, step), it
            if it == 5:
                allocs = ctx.device_allocs()             # the step-by-step path warms its own blocks
    assert pr.uses_graph
    assert ctx.device_allocs() == allocs
    pr.free()


def test_prover_bytes_accepted_by_the_independent_verifier(ctx, qpzk, orc):
    """prove() (one call) -> ProofWithPublicInputs bytes -> tests/perm_util.verify_proof_bytes; equal to the bytes of the
    step-by-step path; tampering is rejected."""
    import gen_poseidon_constants as gen
    from qpzk import prover, gates
    from perm_util import poseidon_chain_circuit, verify_proof_bytes
    log_n = 5
    rc = gen.round_constants()
    rec = gates.Recorder()
    gates.poseidon_gate(rec, rc, gen.derive_fast_partial(rc), selector=rec.const(0))
    code, imm, regs = rec.finish()
    prog = qpzk.GateProgram.make(code, imm, regs)
    wires, consts, sigmas = poseidon_chain_circuit(log_n, 21, rc)
    cs = np.concatenate([consts, sigmas])
    circuit = prover.SyntheticCircuit(ctx, cs, dict(prover.LEAF_SHAPE, log_n=log_n, constants_sigmas=81), sigma_col0=1)
    public_inputs = wires[12:16, (1 << log_n) - 3].copy()
    po = orc.fri_params(log_n, 3, 4, 16, 28, 4, 5)
    ref, _ = prover.prove(circuit, wires, public_inputs, 80, prog, step_by_step=True)
    for it in range(4):                                                     # eager, eager, capture, replay
        data, shape = prover.prove(circuit, wires, public_inputs, 80, prog)
        assert data == ref, it
    args = (circuit.fri_params, po, shape, circuit.constants_sigmas.cap, circuit.circuit_digest, 80)
    assert verify_proof_bytes(orc, qpzk, data, *args, program=(code, imm))
    bad = bytearray(data)
    bad[len(data) // 3] ^= 4
    try:
        ok = verify_proof_bytes(orc, qpzk, bytes(bad), *args, program=(code, imm))
    except qpzk.QpzkError:
        ok = False
    assert not ok


def test_prover_with_blinding_salts(ctx, qpzk, orc):
    """A zero-knowledge circuit: the wires, zs/partial-products and quotient oracles carry 4 salt columns each (caller's
    randomness).  One-call prover == step-by-step calls with the same salts; the oracle's FRI verifier accepts."""
    import torch
    from qpzk import prover
    log_n, W, R, K = 9, 24, 19, 3
    n, L = 1 << log_n, 1 << (log_n + 3)
    rng = np.random.default_rng(8)
    cs = rng.integers(0, P, size=(K + R, n), dtype=np.uint64)
    shape = dict(prover.LEAF_SHAPE, log_n=log_n, wires=W, constants_sigmas=K + R)
    circuit = prover.SyntheticCircuit(ctx, cs, shape, sigma_col0=K)
    pr = circuit.prover(W, R, 5, None, blinding=True)
    for it in range(4):
        wires = rng.integers(0, P, size=(W, n), dtype=np.uint64)
        salts = rng.integers(0, P, size=(3, 4, L), dtype=np.uint64)
        pi = rng.integers(0, P, size=5, dtype=np.uint64)
        got = pr.prove(wires, pi, salts=salts)
        # step by step with the same salts
        sd = torch.from_numpy(salts.view(np.int64)).cuda()
        torch.cuda.synchronize()
        ch = qpzk.Challenger(ctx)
        ch.observe_hash(circuit.circuit_digest)
        ch.observe_hash(orc.hash_no_pad(pi))
        ctx.set_keep_values(True)
        wb = qpzk.PolynomialBatch.from_values(ctx, wires, 3, salts[0], 4)
        ctx.set_keep_values(False)
        ch.observe_cap(wb)
        bg = ch.get_n_challenges(4)
        zb = qpzk.PolynomialBatch.commit_zs_partial_products(circuit.constants_sigmas, wb, K, R, bg[:2], bg[2:], 4, salt_ptr=sd[1].data_ptr())
        ch.observe_cap(zb)
        al = ch.get_n_challenges(2)
        qb = qpzk.PolynomialBatch.commit_quotient_prog(circuit.constants_sigmas, wb, zb, K, R, bg[:2], bg[2:], al, None, 4, salt_ptr=sd[2].data_ptr())
        ch.observe_cap(qb)
        zeta = ch.get_extension_challenge()
        g = prover._root_of_unity(log_n)
        gz = [int(zeta[0]) * g % P, int(zeta[1]) * g % P]
        oracles = [circuit.constants_sigmas, wb, zb, qb]
        assert [o.leaf_len for o in oracles] == [K + R, W + 4, 6 + 4, 16 + 4] == list(pr.layout.leaf_lens)
        inst = [([int(zeta[0]), int(zeta[1])], [(o, i) for o, b in enumerate(oracles) for i in range(b.num_cols)]), (gz, [(2, 0), (2, 1)])]
        op = np.concatenate([qpzk.PolynomialBatch.eval_ext_many(oracles, zeta), zb.eval_ext(gz)[:2]])
        ch.observe_extension_elements(op)
        proof = qpzk.PolynomialBatch.prove_openings(qpzk.FriInstanceInfo(inst), oracles, ch, circuit.fri_params)
        assert (got["caps"][0] == wb.cap).all() and (got["caps"][1] == zb.cap).all() and (got["caps"][2] == qb.cap).all()
        assert (got["openings"] == op).all() and (got["fri_proof"] == proof).all(), it
        if it == 0:
            vch = orc.Challenger()
            vch.observe(circuit.circuit_digest); vch.observe(orc.hash_no_pad(pi)); vch.observe(wb.cap); vch.get(4)
            vch.observe(zb.cap); vch.get(2); vch.observe(qb.cap); vch.get(2); vch.observe(op)
            rc = orc.fri_verify(log_n, [o.leaf_len for o in oracles], [o.cap for o in oracles], orc.Instance(inst), op,
                                orc.fri_params(log_n, 3, 4, 16, 28, 4, 5), vch, got["fri_proof"])
            assert rc == 0
        for b in (wb, zb, qb):
            b.free()
        ch.free()
    assert pr.uses_graph
    pr.free()


def test_provers_in_flight_on_several_contexts(qpzk):
    """Three provers on three contexts driven from one host thread, proofs overlapping on the GPU (the aggregator's
    dispatch of independent leaf proofs, reference wormhole/aggregator/src/aggregator.rs:187-227): every result equals
    the result of the same proof run alone."""
    from qpzk import prover
    log_n = 10
    shape = dict(prover.LEAF_SHAPE, log_n=log_n)
    cs = splitmix_columns(1, 85, 1 << log_n)
    ctxs = [qpzk.Context(0) for _ in range(3)]
    circuits = [prover.SyntheticCircuit(c, cs, shape, sigma_col0=5) for c in ctxs]
    provers = [c.prover(135, 80, 4) for c in circuits]
    ws = [splitmix_columns(200 + k, 135, 1 << log_n) for k in range(5)]
    pis = [splitmix_columns(300 + k, 1, 4)[0] for k in range(5)]
    alone = [provers[0].prove(ws[k], pis[k]) for k in range(5)]
    for rnd in range(4):
        for j, pr in enumerate(provers):
            pr.prove_async(ws[(rnd + j) % 5], pis[(rnd + j) % 5])
        for j, pr in enumerate(provers):
            assert _same(pr.finish(), alone[(rnd + j) % 5]), (rnd, j)
    with pytest.raises(qpzk.QpzkError):
        provers[0].finish()                                   # nothing in flight
    for pr in provers:
        pr.free()


def test_prover_pow_witness_policy_and_argument_errors(ctx, qpzk):
    from qpzk import prover
    log_n = 7
    shape = dict(prover.LEAF_SHAPE, log_n=log_n, wires=20, constants_sigmas=12)
    cs = splitmix_columns(1, 12, 1 << log_n)
    circuit = prover.SyntheticCircuit(ctx, cs, shape, sigma_col0=2)
    pr = circuit.prover(20, 10, 3)
    w, pi = splitmix_columns(2, 20, 1 << log_n), np.array([1, 2, 3], np.uint64)
    a = pr.prove(w, pi)
    witness = int(a["fri_proof"][-1])
    assert _same(pr.prove(w, pi, pow_witness=witness), a)                  # the caller's witness, if it is the minimum: same proof
    for k in range(3):
        assert _same(pr.prove(w, pi), a)                                    # graph replays
    rejected = 0
    for bad in range(witness + 1, witness + 9):                             # a weak witness is refused (each of these is valid with probability 2^-16)
        try:
            pr.prove(w, pi, pow_witness=bad)
        except qpzk.QpzkError as ex:
            assert ex.code == -4                                            # QPZK_EPOW
            rejected += 1
    assert rejected >= 7
    with pytest.raises(qpzk.QpzkError):
        pr.prove(w[:, :64], pi)
    with pytest.raises(qpzk.QpzkError):
        pr.prove(w, np.array([1, 2, P], np.uint64))                         # non-canonical public input
    with pytest.raises(qpzk.QpzkError):
        circuit.prover(20, 30, 3)                                           # more routed wires than wires
    plain = qpzk.PolynomialBatch.from_values(ctx, cs, 3, None, 4)           # committed without keep_values
    with pytest.raises(qpzk.QpzkError):
        qpzk.Prover(plain, 20, 10, 2, circuit.fri_params, circuit.circuit_digest, 3)
    pr.free()
