# dev aid: latency of the leaf-shaped synthetic prove (warm), with a per-phase breakdown
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
from qpzk import prover
ctx = qpzk.Context(0)
shape = prover.LEAF_SHAPE
rng = np.random.default_rng(1)
cs = rng.integers(0, qpzk.P, size=(shape["constants_sigmas"], 1 << shape["log_n"]), dtype=np.uint64)
circuit = prover.SyntheticCircuit(ctx, cs, shape)
inputs = [prover.synthetic_inputs(s, shape) for s in range(4)]
for i in range(3):
    prover.prove_synthetic(circuit, **inputs[i % 4])
ctx.synchronize()
N = 20
t0 = time.perf_counter()
for i in range(N):
    prover.prove_synthetic(circuit, **inputs[i % 4])
ctx.synchronize()
dt = (time.perf_counter() - t0) / N
print("prove_synthetic (leaf shape 2^13): %.2f ms per proof, %.1f proofs/s" % (1e3 * dt, 1 / dt))
# breakdown
def T(f):
    ctx.synchronize(); t = time.perf_counter(); r = f(); ctx.synchronize(); return r, 1e3 * (time.perf_counter() - t)
inp = inputs[0]
w, t_w = T(lambda: qpzk.PolynomialBatch.from_values(ctx, inp["wires"], 3, None, 4))
z, t_z = T(lambda: qpzk.PolynomialBatch.from_values(ctx, inp["zs_partial_products"], 3, None, 4))
q, t_q = T(lambda: qpzk.PolynomialBatch.from_coeffs(ctx, inp["quotient_coeffs"], 3, None, 4))
ch = qpzk.Challenger(ctx)
_, t_obs = T(lambda: [ch.observe_cap(b) for b in (w, z, q)])
zeta, t_get = T(lambda: ch.get_extension_challenge())
oracles = [circuit.constants_sigmas, w, z, q]
ops, t_ev = T(lambda: [b.eval_ext(zeta) for b in oracles])
allp = [(o, i) for o, b in enumerate(oracles) for i in range(b.num_cols)]
inst = qpzk.FriInstanceInfo([([int(zeta[0]), int(zeta[1])], allp)])
_, t_fri = T(lambda: qpzk.PolynomialBatch.prove_openings(inst, oracles, ch, circuit.fri_params))
print("commit wires %.2f, zs %.2f, quotient %.2f, observe caps %.2f, get %.2f, 4 eval_ext %.2f, fri %.2f ms" % (t_w, t_z, t_q, t_obs, t_get, t_ev, t_fri))
