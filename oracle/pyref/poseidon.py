"""ORACLE (test infrastructure) -- Poseidon as the shielder relations use it.

Reference call sites (the only first-party facts): parameters T=5, RATE=4, R_F=8, R_P=56
(/root/reference/shielder/relations/src/lib.rs:17-26), constructed as
`PoseidonHasher::<F,5,4>::new(OptimizedPoseidonSpec::new::<8,56,0>())`
(relations/src/relations/update_note.rs:115-117, update_account.rs:74-76) and used only
through `hash_fix_len_array` (update_note.rs:100,131; update_account.rs:62;
merkle_proof.rs:56).

PARITY UNPINNED against the reference: the permutation lives in the un-vendored halo2-base 0.4.1 @86376e7
(shielder/Cargo.lock:414-416); no hash vector exists in the tree.  Third-party pin (round 2): the Grain
generator and the round structure below, instantiated for the public BN254 instance x5_254_3, reproduce the
published round constants, MDS matrix and the circomlib vector poseidon([1, 2]) (tests/test_oracle_kat.py) --
so constants generation and permutation are not from recall any more; the sponge conventions of halo2-base
(initial 2^64, padding, output lane) still are.  What is restated here ([recall]):
  * constants from the Poseidon reference Grain LFSR (field tag 1, S-box tag 0 = x^alpha,
    n = 255 bits, t, R_F, R_P), round constants with rejection sampling, Cauchy MDS from
    2t elements without rejection (first t = xs, next t = ys), SECURE_MDS = 0 -> the
    first candidate matrix;
  * alpha = 5;
  * sponge: state starts [2^64, 0, 0, 0, 0]; each chunk of <= RATE inputs is added into
    state[1..], and a 1 is added at position len(chunk)+1 when that position exists;
    if len(inputs) % RATE == 0 one extra permutation with an empty chunk follows;
    the digest is state[1];
  * the "optimized" spec is numerically the plain permutation
    (ARK -> S-box -> MDS per round; partial rounds S-box lane 0 only).
The field is BLS12-381 Fr (BASELINE.json), where gcd(5, r-1) = 1.
"""
from __future__ import annotations
from functools import lru_cache
from .bls12_381 import R, finv

T_WIDTH, RATE, R_F, R_P = 5, 4, 8, 56
ALPHA = 5
FIELD_BITS = 255


class Grain:
    """Poseidon reference LFSR (80-bit, self-shrinking output)."""
    def __init__(self, field_bits: int, t: int, r_f: int, r_p: int, field_tag: int = 1, sbox_tag: int = 0):
        bits = []
        def push(v, n): bits.extend((v >> (n - 1 - i)) & 1 for i in range(n))
        push(field_tag, 2); push(sbox_tag, 4); push(field_bits, 12); push(t, 12)
        push(r_f, 10); push(r_p, 10); push((1 << 30) - 1, 30)
        assert len(bits) == 80
        self.state = bits
        for _ in range(160): self._clock()

    def _clock(self) -> int:
        s = self.state
        nb = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0); s.append(nb)
        return nb

    def next_bit(self) -> int:
        while True:
            if self._clock():
                return self._clock()
            self._clock()

    def random_bits(self, n: int) -> int:
        v = 0
        for _ in range(n): v = (v << 1) | self.next_bit()   # MSB first
        return v

    def field_element(self, modulus: int, bits: int) -> int:
        while True:
            v = self.random_bits(bits)
            if v < modulus: return v

    def field_element_no_reject(self, modulus: int, bits: int) -> int:
        return self.random_bits(bits) % modulus


@lru_cache(maxsize=None)
def constants(t: int = T_WIDTH, r_f: int = R_F, r_p: int = R_P, modulus: int = R, bits: int = FIELD_BITS):
    """(round_constants[(r_f + r_p)][t], mds[t][t])."""
    g = Grain(bits, t, r_f, r_p)
    rc = [[g.field_element(modulus, bits) for _ in range(t)] for _ in range(r_f + r_p)]
    while True:
        vals = [g.field_element_no_reject(modulus, bits) for _ in range(2 * t)]
        if len(set(vals)) != 2 * t: continue
        xs, ys = vals[:t], vals[t:]
        if any((x + y) % modulus == 0 for x in xs for y in ys): continue
        break
    mds = [[finv(xs[i] + ys[j], modulus) for j in range(t)] for i in range(t)]
    return rc, mds


def permute_params(state, t: int, r_f: int, r_p: int, modulus: int, bits: int, trace=None):
    """Plain Poseidon permutation (x^5) for any instance the Grain generator can parameterise.  Used with the
    shielder parameters below and -- as a third-party pin of generator AND round structure -- with the public
    BN254 instance x5_254_3 (t=3, R_F=8, R_P=57), whose constants and test vector are widely published
    (tests/test_oracle_kat.py::test_poseidon_generator_and_permutation_match_the_public_bn254_instance)."""
    rc, mds = constants(t, r_f, r_p, modulus, bits)
    s = [x % modulus for x in state]
    half = r_f // 2
    for rnd in range(r_f + r_p):
        s = [(s[i] + rc[rnd][i]) % modulus for i in range(t)]
        full = rnd < half or rnd >= half + r_p
        for i in range(t if full else 1):
            x2 = s[i] * s[i] % modulus; x4 = x2 * x2 % modulus; x5 = x4 * s[i] % modulus
            if trace is not None: trace.append((x2, x4, x5))
            s[i] = x5
        s = [sum(mds[i][j] * s[j] for j in range(t)) % modulus for i in range(t)]
    return s


def permute(state, trace=None):
    """The shielder instance (BLS12-381 Fr, T=5, R_F=8, R_P=56).  If `trace` is a list, appends (x^2, x^4, x^5) per
    S-box in evaluation order -- exactly the R1CS witness block of one permutation."""
    return permute_params(state, T_WIDTH, R_F, R_P, R, FIELD_BITS, trace)


def hash_fix_len_array(inputs, trace=None) -> int:
    """halo2-base PoseidonHasher::hash_fix_len_array semantics (see module docstring)."""
    state = [1 << 64, 0, 0, 0, 0]
    inputs = [x % R for x in inputs]
    chunks = [inputs[i:i + RATE] for i in range(0, len(inputs), RATE)]
    if len(inputs) % RATE == 0: chunks.append([])
    for ch in chunks:
        for i, x in enumerate(ch): state[1 + i] = (state[1 + i] + x) % R
        if len(ch) + 1 < T_WIDTH: state[len(ch) + 1] = (state[len(ch) + 1] + 1) % R
        state = permute(state, trace)
    return state[1]


def n_permutations(n_inputs: int) -> int:
    return n_inputs // RATE + 1
