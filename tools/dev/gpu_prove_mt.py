# dev aid: throughput of independent synthetic proofs with T host threads, one context (stream) per thread
import sys, time, threading, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
from qpzk import prover
shape = prover.LEAF_SHAPE
rng = np.random.default_rng(1)
cs = rng.integers(0, qpzk.P, size=(shape["constants_sigmas"], 1 << shape["log_n"]), dtype=np.uint64)
inputs = [prover.synthetic_inputs(s, shape) for s in range(4)]
def worker(ctx, circuit, n, out, k):
    for i in range(n):
        prover.prove_synthetic(circuit, **inputs[(i + k) % 4])
    ctx.synchronize()
for T in (1, 2, 4, 8, 16):
    ctxs = [qpzk.Context(0) for _ in range(T)]
    circuits = [prover.SyntheticCircuit(c, cs, shape) for c in ctxs]
    for c, ci in zip(ctxs, circuits):
        for i in range(2): prover.prove_synthetic(ci, **inputs[i])
        c.synchronize()
    N = 16
    th = [threading.Thread(target=worker, args=(ctxs[k], circuits[k], N, None, k)) for k in range(T)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print("threads %2d: %.1f proofs/s (%.2f ms per proof per thread)" % (T, T * N / dt, 1e3 * dt / N), flush=True)
    del circuits, ctxs
