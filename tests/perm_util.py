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


def horner(coeffs, z):
    acc = 0
    for c in coeffs[::-1]:
        acc = (acc * z + int(c)) % P
    return acc


def quotient_identity_holds(log_n, wires_coeffs, sigma_coeffs, zs_pp_coeffs, quotient_chunks, betas, gammas, alphas,
                            chunk, zeta, gate_term_values=()):
    """The verifier's check at a base-field point zeta (upstream plonk/verifier.rs + eval_vanishing_poly):
        sum_t terms[t](zeta) alpha_a^t  ==  Z_H(zeta) * sum_k zeta^(k n) chunk_{a,k}(zeta)      for every challenge a,
    terms = [L_0 (Z_c - 1)]_c ++ [partial-product checks]_c ++ gate terms.  All polynomials are given by coefficients
    and evaluated with big-integer Horner, independently of any NTT."""
    n = 1 << log_n
    R, nch = len(wires_coeffs), len(betas)
    nchunks = (R + chunk - 1) // chunk
    np_ = nchunks - 1
    w = root_of_unity(log_n)
    wv = [horner(c, zeta) for c in wires_coeffs]
    sv = [horner(c, zeta) for c in sigma_coeffs]
    zv = [horner(c, zeta) for c in zs_pp_coeffs]
    z_next = [horner(zs_pp_coeffs[c], zeta * w % P) for c in range(nch)]
    zh = (pow(zeta, n, P) - 1) % P
    l0 = zh * pow(n * (zeta - 1) % P, P - 2, P) % P
    terms = [l0 * (zv[c] - 1) % P for c in range(nch)]
    for c in range(nch):
        accs = [zv[c]] + zv[nch + c * np_: nch + (c + 1) * np_] + [z_next[c]]
        for k in range(nchunks):
            pn = pd = 1
            for j in range(k * chunk, min(R, (k + 1) * chunk)):
                pn = pn * ((wv[j] + int(betas[c]) * pow(G, j, P) * zeta + int(gammas[c])) % P) % P
                pd = pd * ((wv[j] + int(betas[c]) * sv[j] + int(gammas[c])) % P) % P
            terms.append((accs[k] * pn - accs[k + 1] * pd) % P)
    terms += [int(v) % P for v in gate_term_values]
    ok = True
    for a in range(nch):
        lhs = sum(t * pow(int(alphas[a]), i, P) for i, t in enumerate(terms)) % P
        q = sum(pow(zeta, k * n, P) * horner(quotient_chunks[a * chunk + k], zeta) for k in range(chunk)) % P
        ok = ok and lhs == zh * q % P
    return ok


# ---- quadratic extension F[X]/(X^2 - 7), elements as (a, b) tuples of Python ints ----------------------------------
def e_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def e_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def e_mul(x, y): return ((x[0] * y[0] + 7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def e_scale(x, s): return (x[0] * s % P, x[1] * s % P)


def e_inv(x):
    norm = (x[0] * x[0] - 7 * x[1] * x[1]) % P
    ni = pow(norm, P - 2, P)
    return (x[0] * ni % P, (-x[1]) * ni % P)


def e_pow(x, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = e_mul(r, x)
        x = e_mul(x, x)
        e >>= 1
    return r


def verify_gateless_openings(log_n, openings, num_consts, num_routed, num_wires, nch, chunk, betas, gammas, alphas, zeta):
    """The plonk verifier's vanishing check on OPENED values (upstream plonk/verifier.rs: eval_vanishing_poly at zeta
    against Z_H(zeta) * reduce_with_powers(quotient chunk openings, zeta^n)) for a circuit without gates.
    openings: (k, 2) array in the order constants | sigmas | wires | zs | partial products | quotient chunks | zs at g zeta."""
    n = 1 << log_n
    op = [(int(a), int(b)) for a, b in openings]
    nchunks = (num_routed + chunk - 1) // chunk
    np_ = nchunks - 1
    o = num_consts
    sig = op[o:o + num_routed]; o += num_routed
    wires = op[o:o + num_wires]; o += num_wires
    zs = op[o:o + nch]; o += nch
    pps = op[o:o + nch * np_]; o += nch * np_
    quot = op[o:o + nch * chunk]; o += nch * chunk
    zs_next = op[o:o + nch]; o += nch
    assert o == len(op)
    z = (int(zeta[0]), int(zeta[1]))
    zh = e_sub(e_pow(z, n), (1, 0))
    l0 = e_mul(zh, e_inv(e_scale(e_sub(z, (1, 0)), n % P)))
    terms = [e_mul(l0, e_sub(zs[c], (1, 0))) for c in range(nch)]
    for c in range(nch):
        accs = [zs[c]] + pps[c * np_:(c + 1) * np_] + [zs_next[c]]
        b, g = int(betas[c]), int(gammas[c])
        for k in range(nchunks):
            pn = pd = (1, 0)
            for j in range(k * chunk, min(num_routed, (k + 1) * chunk)):
                pn = e_mul(pn, e_add(e_add(wires[j], e_scale(z, b * pow(G, j, P) % P)), (g, 0)))
                pd = e_mul(pd, e_add(e_add(wires[j], e_scale(sig[j], b)), (g, 0)))
            terms.append(e_sub(e_mul(accs[k], pn), e_mul(accs[k + 1], pd)))
    zn = e_pow(z, n)
    for a in range(nch):
        lhs = (0, 0)
        for t in terms[::-1]:
            lhs = e_add(e_scale(lhs, int(alphas[a])), t)
        q = (0, 0)
        for k in range(chunk - 1, -1, -1):
            q = e_add(e_mul(q, zn), quot[a * chunk + k])
        if lhs != e_mul(zh, q):
            return False
    return True
