"""ORACLE (test infrastructure, never on the product path) -- BLS12-381 in Python big-ints.

Parity status: PARITY UNPINNED against the reference.  The reference tree
(/root/reference, Cardinal-Cryptography/zk-apps @967d180) contains no BLS12-381,
arkworks or Groth16 code and no golden vectors.  This
file restates the *published* curve (constants in SURVEY.md Appendix A, re-derived
below from the BLS parameter x) and arkworks-0.4 conventions (ark-ff 0.4.2 /
ark-ec 0.4.2 / ark-bls12-381 0.4.0, pinned only in
shielder/contract/Cargo.lock:195-281).  It is pinned by mathematical identities and
the known-answer vectors of Appendix A (tests/test_oracle_kat.py).

Everything here is slow and obviously-correct on purpose.
"""
from __future__ import annotations

# ----------------------------------------------------------------------------- parameters
X_PARAM = -0xd201000000010000
R = X_PARAM ** 4 - X_PARAM ** 2 + 1                      # scalar field modulus (Fr)
P = ((X_PARAM - 1) ** 2 * R) // 3 + X_PARAM              # base field modulus (Fq)
assert P == 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
assert R == 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
H1 = 0x396c8c005555e1568c00aaab0000aaab                  # G1 cofactor
B_G1 = 4

FR_BYTES, FQ_BYTES = 32, 48
FR_MONT_R = pow(2, 256, R)
FQ_MONT_R = pow(2, 384, P)
FR_GENERATOR = 7                                         # multiplicative generator of Fr (ark Fr::GENERATOR)
FR_TWO_ADICITY = 32
FR_ROOT_2_32 = pow(FR_GENERATOR, (R - 1) >> 32, R)       # ark TWO_ADIC_ROOT_OF_UNITY


def finv(a: int, m: int) -> int:
    return pow(a % m, -1, m)


# ----------------------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def fq2(a, b=0):
    return (a % P, b % P)

FQ2_ZERO, FQ2_ONE = (0, 0), (1, 0)

def fq2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def fq2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def fq2_neg(a): return ((-a[0]) % P, (-a[1]) % P)
def fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def fq2_sqr(a): return fq2_mul(a, a)
def fq2_muli(a, k: int): return ((a[0] * k) % P, (a[1] * k) % P)
def fq2_inv(a):
    d = finv(a[0] * a[0] + a[1] * a[1], P)
    return ((a[0] * d) % P, (-a[1] * d) % P)
def fq2_is_zero(a): return a[0] % P == 0 and a[1] % P == 0

B_G2 = (4, 4)                                            # 4(1+u)


class Field:
    """Tiny vtable so the curve code is written once for Fq (ints) and Fq2 (pairs)."""
    def __init__(self, zero, one, add, sub, neg, mul, inv, is_zero, muli):
        self.zero, self.one = zero, one
        self.add, self.sub, self.neg, self.mul, self.inv = add, sub, neg, mul, inv
        self.is_zero, self.muli = is_zero, muli
    def sqr(self, a): return self.mul(a, a)

FQ = Field(0, 1, lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a: (-a) % P,
           lambda a, b: (a * b) % P, lambda a: finv(a, P), lambda a: a % P == 0,
           lambda a, k: (a * k) % P)
FQ2 = Field(FQ2_ZERO, FQ2_ONE, fq2_add, fq2_sub, fq2_neg, fq2_mul, fq2_inv, fq2_is_zero, fq2_muli)


# ----------------------------------------------------------------------------- curves (short Weierstrass, a = 0)
# Affine points are (x, y) tuples or None for infinity.  Jacobian (X, Y, Z), Z == zero <=> infinity.
G1_GEN = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
          0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2_GEN = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
           0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
          (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
           0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))


class Curve:
    def __init__(self, F: Field, b, gen):
        self.F, self.b, self.gen = F, b, gen

    def on_curve(self, pt) -> bool:
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.is_zero(F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)))

    # -- affine group law (used by the naive reference paths)
    def neg(self, pt):
        return None if pt is None else (pt[0], self.F.neg(pt[1]))

    def add(self, p1, p2):
        F = self.F
        if p1 is None: return p2
        if p2 is None: return p1
        x1, y1 = p1; x2, y2 = p2
        if F.is_zero(F.sub(x1, x2)):
            if F.is_zero(F.add(y1, y2)):
                return None
            lam = F.mul(F.muli(F.sqr(x1), 3), F.inv(F.muli(y1, 2)))
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    # -- Jacobian (fast path for scalar multiplication)
    def to_jac(self, pt):
        F = self.F
        return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)

    def from_jac(self, j):
        F = self.F
        X, Y, Z = j
        if F.is_zero(Z): return None
        zi = F.inv(Z); zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jdbl(self, j):
        F = self.F
        X, Y, Z = j
        if F.is_zero(Z): return j
        A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
        D = F.muli(F.sub(F.sub(F.sqr(F.add(X, B)), A), C), 2)
        E = F.muli(A, 3); Fq_ = F.sqr(E)
        X3 = F.sub(Fq_, F.muli(D, 2))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muli(C, 8))
        Z3 = F.muli(F.mul(Y, Z), 2)
        return (X3, Y3, Z3)

    def jadd(self, j1, j2):
        F = self.F
        if F.is_zero(j1[2]): return j2
        if F.is_zero(j2[2]): return j1
        X1, Y1, Z1 = j1; X2, Y2, Z2 = j2
        Z1Z1 = F.sqr(Z1); Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2); S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        Hh = F.sub(U2, U1); Rr = F.sub(S2, S1)
        if F.is_zero(Hh):
            if F.is_zero(Rr): return self.jdbl(j1)
            return (F.one, F.one, F.zero)
        HH = F.sqr(Hh); HHH = F.mul(Hh, HH); V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.muli(V, 2))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), Hh)
        return (X3, Y3, Z3)

    def mul(self, pt, k: int):
        """k*pt, affine in / affine out, any integer k."""
        if pt is None: return None
        if k < 0: return self.mul(self.neg(pt), -k)
        acc = self.to_jac(None); base = self.to_jac(pt)
        while k:
            if k & 1: acc = self.jadd(acc, base)
            base = self.jdbl(base); k >>= 1
        return self.from_jac(acc)

    def msm_naive(self, bases, scalars):
        acc = self.to_jac(None)
        for b, s in zip(bases, scalars):
            s %= R
            if s and b is not None:
                acc = self.jadd(acc, self.to_jac(self.mul(b, s)))
        return self.from_jac(acc)


G1 = Curve(FQ, B_G1, G1_GEN)
G2 = Curve(FQ2, B_G2, G2_GEN)


# ----------------------------------------------------------------------------- byte layouts
def int_to_le(v: int, n: int) -> bytes: return int(v).to_bytes(n, "little")
def le_to_int(b: bytes) -> int: return int.from_bytes(b, "little")

def fr_to_mont_bytes(v: int) -> bytes: return int_to_le(v % R * FR_MONT_R % R, 32)
def fr_from_mont_bytes(b: bytes) -> int: return le_to_int(b) * finv(FR_MONT_R, R) % R
def fq_to_mont_bytes(v: int) -> bytes: return int_to_le(v % P * FQ_MONT_R % P, 48)
def fq_from_mont_bytes(b: bytes) -> int: return le_to_int(b) * finv(FQ_MONT_R, P) % P

def g1_to_ffi(pt) -> bytes:
    """96 B: x||y little-endian Montgomery limbs (include/b200zk.h).  Infinity = all zero."""
    if pt is None: return bytes(96)
    return fq_to_mont_bytes(pt[0]) + fq_to_mont_bytes(pt[1])

def g1_from_ffi(b: bytes):
    if b == bytes(96): return None
    return (fq_from_mont_bytes(b[:48]), fq_from_mont_bytes(b[48:96]))

def g2_to_ffi(pt) -> bytes:
    """192 B: x.c0||x.c1||y.c0||y.c1."""
    if pt is None: return bytes(192)
    (x0, x1), (y0, y1) = pt
    return b"".join(fq_to_mont_bytes(v) for v in (x0, x1, y0, y1))

def g2_from_ffi(b: bytes):
    if b == bytes(192): return None
    v = [fq_from_mont_bytes(b[48 * i:48 * i + 48]) for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))


# zcash / IETF compressed encoding as used by ark-bls12-381 0.4 CanonicalSerialize [recall, SURVEY App. B]
def _fq_lex_largest(y: int) -> bool: return y > (P - 1) // 2
def _fq2_lex_largest(y) -> bool:
    return _fq_lex_largest(y[1]) if y[1] != 0 else _fq_lex_largest(y[0])

def g1_compress(pt) -> bytes:
    if pt is None: return bytes([0xC0]) + bytes(47)
    b = bytearray(pt[0].to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if _fq_lex_largest(pt[1]) else 0)
    return bytes(b)

def g2_compress(pt) -> bytes:
    if pt is None: return bytes([0xC0]) + bytes(95)
    (x0, x1), y = pt
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if _fq2_lex_largest(y) else 0)
    return bytes(b)

def fq_sqrt(a: int):
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None

def fq2_sqrt(a):
    """Square root in Fq2 (p = 3 mod 4), via the norm: returns one root or None."""
    if fq2_is_zero(a): return FQ2_ZERO
    a0, a1 = a
    if a1 == 0:
        s = fq_sqrt(a0)
        if s is not None: return (s, 0)
        s = fq_sqrt((-a0) % P)
        return (0, s) if s is not None else None
    n = fq_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None: return None
    inv2 = finv(2, P)
    for nn in (n, (-n) % P):
        x0sq = (a0 + nn) * inv2 % P
        x0 = fq_sqrt(x0sq)
        if x0 is None or x0 == 0: continue
        x1 = a1 * finv(2 * x0, P) % P
        if fq2_sqr((x0, x1)) == (a0 % P, a1 % P): return (x0, x1)
    return None

def g1_decompress(b: bytes):
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40: return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = fq_sqrt((x * x * x + B_G1) % P)
    assert y is not None, "not on curve"
    if _fq_lex_largest(y) != bool(b[0] & 0x20): y = (-y) % P
    return (x, y)

def g2_decompress(b: bytes):
    assert len(b) == 96 and b[0] & 0x80
    if b[0] & 0x40: return None
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    x = (x0, x1)
    y = fq2_sqrt(fq2_add(fq2_mul(fq2_sqr(x), x), B_G2))
    assert y is not None, "not on curve"
    if _fq2_lex_largest(y) != bool(b[0] & 0x20): y = fq2_neg(y)
    return (x, y)


# ----------------------------------------------------------------------------- pairing (for the Groth16 verifier oracle)
# Fq12 as Fq[w]/(w^12 - 2 w^6 + 2): w^6 = 1 + u (the sextic non-residue xi) and u^2 = -1.
_FQ12_MOD = [2, 0, 0, 0, 0, 0, -2, 0, 0, 0, 0, 0]        # low-order coefficients of the monic modulus

def f12(c=None):
    v = [0] * 12
    if c:
        for i, x in enumerate(c): v[i] = x % P
    return v

F12_ONE = f12([1])

def f12_add(a, b): return [(x + y) % P for x, y in zip(a, b)]
def f12_sub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def f12_scal(a, k): return [(x * k) % P for x in a]

def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for i in range(22, 11, -1):                           # w^12 = 2 w^6 - 2
        top = t[i]
        if top:
            t[i - 6] += 2 * top
            t[i - 12] -= 2 * top
    return [x % P for x in t[:12]]

def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0: d -= 1
    return d

def f12_inv(a):
    """Extended Euclid over Fq[w]."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [x % P for x in _FQ12_MOD] + [1]
    while _poly_deg(low):
        # r = high / low (polynomial division, rounded)
        dega, degb = _poly_deg(high), _poly_deg(low)
        temp = list(high); o = [0] * 13
        binv = finv(low[degb], P)
        for i in range(dega - degb, -1, -1):
            o[i] = (o[i] + temp[degb + i] * binv) % P
            for c in range(degb + 1):
                temp[c + i] = (temp[c + i] - o[i] * low[c]) % P
        r = o
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] -= lm[i] * r[j]
                new[i + j] -= low[i] * r[j]
        nm = [x % P for x in nm]; new = [x % P for x in new]
        lm, low, hm, high = nm, new, lm, low
    c = finv(low[0], P)
    return [(x * c) % P for x in lm[:12]]

def f12_pow(a, e: int):
    out, base = F12_ONE, a
    while e:
        if e & 1: out = f12_mul(out, base)
        base = f12_mul(base, base); e >>= 1
    return out

def _fq2_to_f12(a):                                       # a0 + a1*u, u = w^6 - 1
    return f12([a[0] - a[1], 0, 0, 0, 0, 0, a[1]])

_W = f12([0, 1])
_W2_INV = f12_inv(f12_mul(_W, _W))
_W3_INV = f12_inv(f12_mul(f12_mul(_W, _W), _W))

def _untwist(q):
    """E'(Fq2) -> E(Fq12): (x', y') -> (x'/w^2, y'/w^3)  (M-type twist, y^2 = x^3 + 4 xi)."""
    return (f12_mul(_fq2_to_f12(q[0]), _W2_INV), f12_mul(_fq2_to_f12(q[1]), _W3_INV))

def _line(t, q, pt):
    """Line through t and q (points of E(Fq12)) evaluated at pt; also returns t+q."""
    xt, yt = t; xq, yq = q; xp, yp = pt
    if xt != xq:
        lam = f12_mul(f12_sub(yq, yt), f12_inv(f12_sub(xq, xt)))
    elif yt == yq:
        lam = f12_mul(f12_scal(f12_mul(xt, xt), 3), f12_inv(f12_scal(yt, 2)))
    else:
        return f12_sub(xp, xt), None                      # vertical line
    val = f12_sub(f12_sub(yp, yt), f12_mul(lam, f12_sub(xp, xt)))
    x3 = f12_sub(f12_sub(f12_mul(lam, lam), xt), xq)
    y3 = f12_sub(f12_mul(lam, f12_sub(xt, x3)), yt)
    return val, (x3, y3)

def miller_loop(p1, q2):
    """f_{|x|, Q}(P), conjugation for the negative BLS parameter folded in by inversion."""
    if p1 is None or q2 is None: return F12_ONE
    pt = (f12([p1[0]]), f12([p1[1]]))
    q = _untwist(q2)
    t = q; f = F12_ONE
    n = -X_PARAM
    for i in range(n.bit_length() - 2, -1, -1):
        l, t = _line(t, t, pt)
        f = f12_mul(f12_mul(f, f), l)
        if (n >> i) & 1:
            l, t = _line(t, q, pt)
            f = f12_mul(f, l)
    return f12_inv(f)                                     # x < 0

def final_exp(f): return f12_pow(f, (P ** 12 - 1) // R)

def pairing(p1, q2): return final_exp(miller_loop(p1, q2))

def pairing_product_is_one(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 with one shared final exponentiation."""
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(p1, q2))
    return final_exp(f) == F12_ONE
