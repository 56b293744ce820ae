"""CPU: the portable (host) path of zk-apps_b200/csrc/pairing.cuh -- the code the verifier kernels compile --
against the Python oracle's independent pairing (naive Fq[w]/(w^12 - 2w^6 + 2) arithmetic, affine lines,
exponentiation by the full (p^12-1)/r).  Bit-exact after mapping the tower basis onto the oracle's."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.pyref import bls12_381 as bls
from oracle.pyref.algos import SplitMix64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, R, X = bls.P, bls.R, bls.X_PARAM


@pytest.fixture(scope="module")
def hc(built):
    L = C.CDLL(os.path.join(ROOT, "tests", "host", "libhostcheck.so"))
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# tower slot k of the 576-byte layout holds the Fq2 coefficient of w^WPOW[k]
WPOW = [0, 2, 4, 1, 3, 5]


def f12_from_tower(buf) -> list:
    b = bytes(buf)
    out = [0] * 12
    for k in range(6):
        a0 = bls.fq_from_mont_bytes(b[96 * k:96 * k + 48])
        a1 = bls.fq_from_mont_bytes(b[96 * k + 48:96 * k + 96])
        e = WPOW[k]                                   # (a0 + a1 u) w^e with u = w^6 - 1
        out[e] = (out[e] + a0 - a1) % P
        out[e + 6] = (out[e + 6] + a1) % P
    return out


def f12_to_tower(v) -> np.ndarray:
    out = b""
    for k in range(6):
        e = WPOW[k]
        a1 = v[e + 6] % P
        a0 = (v[e] + a1) % P
        out += bls.fq_to_mont_bytes(a0) + bls.fq_to_mont_bytes(a1)
    return np.frombuffer(out, dtype=np.uint8).copy()


def rand_f12(g):
    return [g.next() * g.next() * g.next() * g.next() * g.next() * g.next() % P for _ in range(12)]


def test_hard_part_identity():
    """(x-1)^2 (x+p)(x^2+p^2-1) + 3 == 3 (p^4-p^2+1)/r: the exponent final_exponentiation implements."""
    assert (P ** 4 - P ** 2 + 1) % R == 0
    assert (X - 1) ** 2 * (X + P) * (X * X + P * P - 1) + 3 == 3 * ((P ** 4 - P ** 2 + 1) // R)


def test_fq12_tower_arithmetic(hc):
    g = SplitMix64(0xB2000801)
    a, b = rand_f12(g), rand_f12(g)
    ta, tb = f12_to_tower(a), f12_to_tower(b)
    assert f12_from_tower(ta) == a
    out = np.zeros(576, dtype=np.uint8)
    hc.hc_fq12_op(0, _p(ta), _p(tb), _p(out)); assert f12_from_tower(out) == bls.f12_mul(a, b)
    hc.hc_fq12_op(1, _p(ta), None, _p(out)); assert f12_from_tower(out) == bls.f12_mul(a, a)
    hc.hc_fq12_op(2, _p(ta), None, _p(out)); assert bls.f12_mul(f12_from_tower(out), a) == bls.F12_ONE
    hc.hc_fq12_op(3, _p(ta), None, _p(out)); assert f12_from_tower(out) == bls.f12_pow(a, P)
    hc.hc_fq12_op(4, _p(ta), None, _p(out)); assert f12_from_tower(out) == bls.f12_pow(a, P ** 6)
    # sparse line product: slots 0, 1, 4 of b
    sp = np.zeros(576, dtype=np.uint8)
    for k in (0, 1, 4):
        sp[96 * k:96 * k + 96] = tb[96 * k:96 * k + 96]
    hc.hc_fq12_op(6, _p(ta), _p(sp), _p(out)); assert f12_from_tower(out) == bls.f12_mul(a, f12_from_tower(sp))


def test_final_exponentiation(hc):
    g = SplitMix64(0xB2000802)
    a = rand_f12(g)
    out = np.zeros(576, dtype=np.uint8)
    hc.hc_fq12_op(5, _p(f12_to_tower(a)), None, _p(out))
    assert f12_from_tower(out) == bls.f12_pow(a, 3 * ((P ** 12 - 1) // R))


def test_pairing_vs_oracle_and_bilinearity(hc):
    """e(aG1, bG2) == oracle pairing cubed (the HHT exponent), == e(G1, G2)^(ab), and infinity -> 1."""
    a, b = 0x1234567, 0xABCDEF01
    p1, q2 = bls.G1.mul(bls.G1_GEN, a), bls.G2.mul(bls.G2_GEN, b)
    out = np.zeros(576, dtype=np.uint8)
    g1b = np.frombuffer(bls.g1_to_ffi(p1), dtype=np.uint8).copy()
    g2b = np.frombuffer(bls.g2_to_ffi(q2), dtype=np.uint8).copy()
    hc.hc_pairing(_p(g1b), _p(g2b), _p(out))
    got = f12_from_tower(out)
    want = bls.f12_pow(bls.pairing(p1, q2), 3)
    assert got == want
    base = np.zeros(576, dtype=np.uint8)
    hc.hc_pairing(_p(np.frombuffer(bls.g1_to_ffi(bls.G1_GEN), dtype=np.uint8).copy()),
                  _p(np.frombuffer(bls.g2_to_ffi(bls.G2_GEN), dtype=np.uint8).copy()), _p(base))
    assert bls.f12_pow(f12_from_tower(base), a * b % R) == got
    assert bls.f12_pow(got, R) == bls.F12_ONE and got != bls.F12_ONE
    hc.hc_pairing(_p(np.zeros(96, dtype=np.uint8)), _p(g2b), _p(out))
    assert f12_from_tower(out) == bls.F12_ONE


def test_wire_format_roundtrip(hc):
    g = SplitMix64(0xB2000803)
    for group, gen, curve, comp, to_ffi, n in ((1, bls.G1_GEN, bls.G1, bls.g1_compress, bls.g1_to_ffi, 48),
                                               (2, bls.G2_GEN, bls.G2, bls.g2_compress, bls.g2_to_ffi, 96)):
        pts = [None] + [curve.mul(gen, g.next()) for _ in range(6)]
        pts.append(curve.neg(pts[1]))
        for pt in pts:
            ffi = np.frombuffer(to_ffi(pt), dtype=np.uint8).copy()
            cb = np.zeros(n, dtype=np.uint8)
            hc.hc_compress(group, _p(ffi), _p(cb))
            assert cb.tobytes() == comp(pt)
            back = np.zeros(2 * n, dtype=np.uint8)
            assert hc.hc_decompress(group, _p(cb), _p(back)) == 0
            assert back.tobytes() == to_ffi(pt)
            assert hc.hc_in_subgroup(group, _p(ffi)) == 1
    # malformed: no compression flag; x >= p; x not on the curve
    bad = np.zeros(48, dtype=np.uint8)
    out = np.zeros(96, dtype=np.uint8)
    assert hc.hc_decompress(1, _p(bad), _p(out)) == 1
    bad = np.frombuffer(bytes([0x9F]) + bytes([0xFF] * 47), dtype=np.uint8).copy()
    assert hc.hc_decompress(1, _p(bad), _p(out)) == 1
    x = 1
    while bls.fq_sqrt((x ** 3 + 4) % P) is not None:
        x += 1
    bad = np.frombuffer(bytes([0x80]) + x.to_bytes(47, "big"), dtype=np.uint8).copy()
    assert hc.hc_decompress(1, _p(bad), _p(out)) == 2
    # on the curve but outside the r-torsion: a point of E(Fq) not multiplied by the cofactor
    x = 1
    while True:
        y = bls.fq_sqrt((x ** 3 + 4) % P)
        if y is not None and bls.G1.mul((x, y), R) is not None:
            break
        x += 1
    assert hc.hc_in_subgroup(1, _p(np.frombuffer(bls.g1_to_ffi((x, y)), dtype=np.uint8).copy())) == 0


def test_fast_subgroup_checks_agree_with_plain(hc):
    """ec_in_subgroup (endomorphism tests: phi(P) == [z^2 - 1]P on G1, psi(P) == [z]P on G2) against [r]P == O, the
    plain test, and against the oracle: subgroup points, random curve points (cofactor != 1, so almost surely
    outside), pure cofactor-torsion points [r]Q, and sums of a subgroup point with a torsion point."""
    import random
    rnd = random.Random(11)
    for group, cv, to_ffi in ((1, bls.G1, bls.g1_to_ffi), (2, bls.G2, bls.g2_to_ffi)):
        def curve_point():
            while True:
                if group == 1:
                    x = rnd.randrange(bls.P)
                    y = bls.fq_sqrt((x * x * x + 4) % bls.P)
                else:
                    x = (rnd.randrange(bls.P), rnd.randrange(bls.P))
                    y = bls.fq2_sqrt(bls.fq2_add(bls.fq2_mul(bls.fq2_sqr(x), x), (4, 4)))
                if y is not None:
                    return (x, y)
        pts = [cv.mul(cv.gen, rnd.randrange(1, bls.R)) for _ in range(4)]
        outside = [curve_point() for _ in range(4)]
        torsion = [cv.mul(q, bls.R) for q in outside[:2]]                      # order divides the cofactor
        mixed = [cv.add(pts[0], t) for t in torsion if t is not None]
        for pt in pts + outside + [t for t in torsion if t is not None] + mixed:
            want = 1 if cv.mul(pt, bls.R) is None else 0
            buf = np.frombuffer(to_ffi(pt), dtype=np.uint8).copy()
            assert hc.hc_in_subgroup_plain(group, _p(buf)) == want
            assert hc.hc_in_subgroup(group, _p(buf)) == want, (group, pt)
        assert sum(1 for pt in outside if cv.mul(pt, bls.R) is not None) >= 3    # the negative cases are real
