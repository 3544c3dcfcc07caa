"""Helpers for the permutation-argument tests: a random copy-constraint system whose wires satisfy it."""
import numpy as np

P = 0xFFFFFFFF00000001
G = 14293326489335486720
W32 = 7277203076849721926


def root_of_unity(log_n):
    w = W32
    for _ in range(32 - log_n):
        w = w * w % P
    return w


def satisfied_permutation(log_n, num_routed, seed):
    """(wires, sigmas, ids): ids[j][i] = k_j * w^i with k_j = g^j; sigmas = ids permuted by a random permutation of all
    positions; wires constant on the cycles of that permutation (so the grand product of the argument is 1)."""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    w = root_of_unity(log_n)
    xs = [pow(w, i, P) for i in range(n)]
    ids = np.array([[pow(G, j, P) * x % P for x in xs] for j in range(num_routed)], dtype=np.uint64)
    perm = rng.permutation(num_routed * n)
    sigmas = ids.reshape(-1)[perm].reshape(num_routed, n)
    vals = np.zeros(num_routed * n, dtype=np.uint64)
    seen = np.zeros(num_routed * n, bool)
    for s0 in range(num_routed * n):
        if not seen[s0]:
            v = int(rng.integers(0, P, dtype=np.uint64))
            k = s0
            while not seen[k]:
                seen[k] = True
                vals[k] = v
                k = perm[k]
    return vals.reshape(num_routed, n), sigmas, ids
