"""ORACLE (test infrastructure) -- Groth16 over BLS12-381 restating ark-groth16 0.4 ([recall];
no ark-groth16 / ark-relations pin exists anywhere in the reference -> PARITY UNPINNED, SURVEY.md
section 8c and Appendix B).  Setup with explicit toxic waste, the LibsnarkReduction witness map,
proof assembly with explicit (r, s), the pairing check, and the 192-byte compressed proof.
What pins it: every proof produced by any layer must satisfy the pairing equation here.
"""
from __future__ import annotations
from dataclasses import dataclass, field
from .bls12_381 import (R, G1, G2, finv, g1_compress, g2_compress, g1_decompress, g2_decompress,
                        pairing_product_is_one, FR_GENERATOR)
from .algos import Domain


@dataclass
class Toxic:
    alpha: int
    beta: int
    gamma: int
    delta: int
    tau: int


@dataclass
class SetupScalars:
    """Everything generate_parameters derives before touching the group (all in Fr)."""
    n: int                       # domain size
    num_inputs: int
    num_constraints: int
    a: list
    b: list
    c: list
    gamma_abc: list              # num_inputs entries
    l: list                      # num_aux entries
    h: list                      # n - 1 entries: tau^i * Z(tau) / delta


def setup_scalars(matrices, num_inputs: int, num_variables: int, tox: Toxic) -> SetupScalars:
    """LibsnarkReduction::instance_map_with_evaluation + the scalar side of generate_parameters."""
    A, B, C = matrices
    nc = len(A)
    dom = Domain(nc + num_inputs)
    n = dom.size
    t = tox.tau % R
    zt = (pow(t, n, R) - 1) % R
    # Lagrange coefficients L_i(tau) = Z(tau) * w^i / (n * (tau - w^i))
    u, wi = [], 1
    ninv = finv(n, R)
    for i in range(n):
        u.append(zt * wi % R * ninv % R * finv(t - wi, R) % R)
        wi = wi * dom.group_gen % R
    a = [0] * num_variables; b = [0] * num_variables; c = [0] * num_variables
    for i in range(num_inputs): a[i] = u[nc + i]
    for i in range(nc):
        for v, k in A[i]: a[v] = (a[v] + u[i] * k) % R
        for v, k in B[i]: b[v] = (b[v] + u[i] * k) % R
        for v, k in C[i]: c[v] = (c[v] + u[i] * k) % R
    ginv, dinv = finv(tox.gamma, R), finv(tox.delta, R)
    comb = [(tox.beta * a[i] + tox.alpha * b[i] + c[i]) % R for i in range(num_variables)]
    gamma_abc = [x * ginv % R for x in comb[:num_inputs]]
    l = [x * dinv % R for x in comb[num_inputs:]]
    h, ti = [], zt * dinv % R
    for i in range(n - 1):
        h.append(ti); ti = ti * t % R
    return SetupScalars(n, num_inputs, nc, a, b, c, gamma_abc, l, h)


@dataclass
class ProvingKey:
    n: int
    num_inputs: int
    alpha_g1: tuple
    beta_g1: tuple
    beta_g2: tuple
    gamma_g2: tuple
    delta_g1: tuple
    delta_g2: tuple
    gamma_abc_g1: list
    a_query: list
    b_g1_query: list
    b_g2_query: list
    l_query: list
    h_query: list


def _py_mul(curve):
    return lambda ks: [curve.mul(curve.gen, k) for k in ks]


def generate_parameters(matrices, num_inputs, num_variables, tox: Toxic, mul_g1=None, mul_g2=None) -> ProvingKey:
    """mul_g1 / mul_g2: batched fixed-base k -> k*G (default: Python; tests plug in the C++ oracle or
    the GPU fixed-base kernel for the full-size circuit)."""
    mul_g1 = mul_g1 or _py_mul(G1)
    mul_g2 = mul_g2 or _py_mul(G2)
    sc = setup_scalars(matrices, num_inputs, num_variables, tox)
    singles1 = mul_g1([tox.alpha, tox.beta, tox.delta])
    singles2 = mul_g2([tox.beta, tox.gamma, tox.delta])
    return ProvingKey(sc.n, num_inputs, singles1[0], singles1[1], singles2[0], singles2[1], singles1[2], singles2[2],
                      mul_g1(sc.gamma_abc), mul_g1(sc.a), mul_g1(sc.b), mul_g2(sc.b), mul_g1(sc.l), mul_g1(sc.h))


def witness_map(matrices, num_inputs: int, z, n: int):
    """LibsnarkReduction::witness_map_from_matrices: coefficients of H (n values)."""
    A, B, C = matrices
    nc = len(A)
    dot = lambda row: sum(k * z[v] for v, k in row) % R
    a = [dot(r) for r in A] + [0] * (n - nc)
    b = [dot(r) for r in B] + [0] * (n - nc)
    c = [dot(r) for r in C] + [0] * (n - nc)
    for i in range(num_inputs): a[nc + i] = z[i]
    dom = Domain(n); cos = dom.get_coset(FR_GENERATOR)
    a = cos.fft(dom.ifft(a)); b = cos.fft(dom.ifft(b)); c = cos.fft(dom.ifft(c))
    zinv = finv(cos.vanishing_on_coset(), R)
    ab = [(x * y - w) * zinv % R for x, y, w in zip(a, b, c)]
    return cos.ifft(ab)


def create_proof_with_reduction(matrices, pk: ProvingKey, z, r: int, s: int, msm=None):
    """Returns (A, B, C) affine.  msm(curve, bases, scalars) defaults to the naive oracle sum."""
    msm = msm or (lambda cv, bs, ss: cv.msm_naive(bs, ss))
    h = witness_map(matrices, pk.num_inputs, z, pk.n)
    aux = z[pk.num_inputs:]
    A = G1.add(G1.add(pk.alpha_g1, msm(G1, pk.a_query, z)), G1.mul(pk.delta_g1, r))
    B1 = G1.add(G1.add(pk.beta_g1, msm(G1, pk.b_g1_query, z)), G1.mul(pk.delta_g1, s))
    B2 = G2.add(G2.add(pk.beta_g2, msm(G2, pk.b_g2_query, z)), G2.mul(pk.delta_g2, s))
    Cc = G1.add(msm(G1, pk.l_query, aux), msm(G1, pk.h_query, h[:pk.n - 1]))
    Cc = G1.add(Cc, G1.mul(A, s))
    Cc = G1.add(Cc, G1.mul(B1, r))
    Cc = G1.add(Cc, G1.neg(G1.mul(pk.delta_g1, r * s % R)))
    return A, B2, Cc


def proof_to_bytes(proof) -> bytes:
    """ark-serialize compressed: A (48) || B (96) || C (48) = 192 B."""
    A, B, Cc = proof
    return g1_compress(A) + g2_compress(B) + g1_compress(Cc)


def proof_from_bytes(b: bytes):
    assert len(b) == 192
    return g1_decompress(b[:48]), g2_decompress(b[48:144]), g1_decompress(b[144:])


def verify(pk: ProvingKey, public_inputs, proof) -> bool:
    """e(A,B) = e(alpha,beta) e(sum x_i gamma_abc_i, gamma) e(C,delta); public_inputs excludes the 1."""
    A, B, Cc = proof
    if not (G1.on_curve(A) and G2.on_curve(B) and G1.on_curve(Cc)): return False
    acc = pk.gamma_abc_g1[0]
    for x, g in zip(public_inputs, pk.gamma_abc_g1[1:]):
        acc = G1.add(acc, G1.mul(g, x % R))
    return pairing_product_is_one([(G1.neg(A), B), (pk.alpha_g1, pk.beta_g2), (acc, pk.gamma_g2), (Cc, pk.delta_g2)])


def proof_via_scalars(matrices, sc: SetupScalars, tox: Toxic, z, r: int, s: int):
    """The same proof computed in the exponent: every key element is a known multiple of the
    generator (a_query[i] = a_i G, ...), so A, B, C are single fixed-base multiplications.  An
    independent, fast derivation used to diff full-size GPU proofs bit for bit."""
    h = witness_map(matrices, sc.num_inputs, z, sc.n)
    dot = lambda xs, ys: sum(x * y for x, y in zip(xs, ys)) % R
    a_s = (tox.alpha + dot(sc.a, z) + r * tox.delta) % R
    b_s = (tox.beta + dot(sc.b, z) + s * tox.delta) % R
    c_s = (dot(sc.l, z[sc.num_inputs:]) + dot(sc.h, h[:sc.n - 1]) + s * a_s + r * b_s - r * s % R * tox.delta) % R
    return G1.mul(G1.gen, a_s), G2.mul(G2.gen, b_s), G1.mul(G1.gen, c_s)


def verifying_key_from_toxic(sc: SetupScalars, tox: Toxic):
    """(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1) from the toxic waste."""
    return (G1.mul(G1.gen, tox.alpha), G2.mul(G2.gen, tox.beta), G2.mul(G2.gen, tox.gamma), G2.mul(G2.gen, tox.delta),
            [G1.mul(G1.gen, k) for k in sc.gamma_abc])


def verify_with_vk(vk, public_inputs, proof) -> bool:
    alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc = vk
    A, B, Cc = proof
    if not (G1.on_curve(A) and G2.on_curve(B) and G1.on_curve(Cc)): return False
    acc = gamma_abc[0]
    for x, g in zip(public_inputs, gamma_abc[1:]):
        acc = G1.add(acc, G1.mul(g, x % R))
    return pairing_product_is_one([(G1.neg(A), B), (alpha_g1, beta_g2), (acc, gamma_g2), (Cc, delta_g2)])
