"""ORACLE (test infrastructure) -- Fr NTT / coset NTT and bucketed MSM, restating arkworks 0.4.

PARITY UNPINNED: ark-poly 0.4.2 / ark-ec 0.4.2 are lockfile-only in the reference
(shielder/contract/Cargo.lock:207-281) -- sources absent; conventions restated from
SURVEY.md Appendix B.  Outputs are mathematically determined (NTT vector in natural
order, MSM point in affine form), so any correct algorithm yields the same bytes.
"""
from __future__ import annotations
from .bls12_381 import R, FR_ROOT_2_32, FR_GENERATOR, finv, G1, G2, Curve


# ----------------------------------------------------------------------------- Radix2EvaluationDomain
class Domain:
    """ark_poly::Radix2EvaluationDomain<Fr>::new(n): size = next_pow2(n), None if log > 32."""
    def __init__(self, n: int, offset: int = 1):
        size = 1
        while size < n: size <<= 1
        self.size = size
        self.log = size.bit_length() - 1
        if self.log > 32: raise ValueError("domain too large")
        self.group_gen = pow(FR_ROOT_2_32, 1 << (32 - self.log), R)
        self.group_gen_inv = finv(self.group_gen, R)
        self.size_inv = finv(size, R)
        self.offset = offset % R
        self.offset_inv = finv(self.offset, R)

    def get_coset(self, offset: int) -> "Domain":
        return Domain(self.size, offset)

    @staticmethod
    def _ntt(a, w):
        n = len(a)
        if n == 1: return list(a)
        lg = n.bit_length() - 1
        a = list(a)
        for i in range(n):                                # bit reversal
            j = int(format(i, "0%db" % lg)[::-1], 2)
            if i < j: a[i], a[j] = a[j], a[i]
        m = 1
        while m < n:
            wm = pow(w, n // (2 * m), R)
            for k in range(0, n, 2 * m):
                t = 1
                for j in range(m):
                    u, v = a[k + j], a[k + j + m] * t % R
                    a[k + j], a[k + j + m] = (u + v) % R, (u - v) % R
                    t = t * wm % R
            m *= 2
        return a

    def fft(self, coeffs):
        """Natural order in/out, X[k] = sum_j x[j] (offset^j) w^{jk}; zero-pads to size."""
        a = [c % R for c in coeffs] + [0] * (self.size - len(coeffs))
        assert len(a) == self.size
        if self.offset != 1:
            g = 1
            for j in range(self.size):
                a[j] = a[j] * g % R; g = g * self.offset % R
        return self._ntt(a, self.group_gen)

    def ifft(self, evals):
        a = [c % R for c in evals] + [0] * (self.size - len(evals))
        assert len(a) == self.size
        a = self._ntt(a, self.group_gen_inv)
        g = self.size_inv
        for j in range(self.size):
            a[j] = a[j] * g % R
            if self.offset != 1: g = g * self.offset_inv % R
        return a

    def dft_naive(self, coeffs):
        n = self.size
        return [sum(c * pow(self.offset, j, R) * pow(self.group_gen, j * k, R) for j, c in enumerate(coeffs)) % R
                for k in range(n)]

    def vanishing_on_coset(self) -> int:
        """Z(offset * w^k) = offset^n - 1 (constant over the coset)."""
        return (pow(self.offset, self.size, R) - 1) % R


# ----------------------------------------------------------------------------- VariableBaseMSM (ark-ec 0.4.2 msm_bigint)
def ln_without_floats(a: int) -> int:
    return (a.bit_length() - 1) * 69 // 100               # log2(a) * 69 / 100

def ark_window(n: int) -> int:
    return 3 if n < 32 else ln_without_floats(n) + 2

def msm_pippenger(curve: Curve, bases, scalars, c: int | None = None):
    """Bucket method as ark-ec 0.4.2 does it: unsigned c-bit digits, 2^c - 1 buckets per window,
    running-sum bucket reduction, windows combined high to low with c doublings."""
    n = min(len(bases), len(scalars))
    if c is None: c = ark_window(n)
    zero = curve.to_jac(None)
    nbits = 255
    window_sums = []
    for w_start in range(0, nbits, c):
        buckets = [zero] * ((1 << c) - 1)
        for b, s in zip(bases[:n], scalars[:n]):
            s %= R
            if s == 0 or b is None: continue
            d = (s >> w_start) & ((1 << c) - 1)
            if d: buckets[d - 1] = curve.jadd(buckets[d - 1], curve.to_jac(b))
        running, res = zero, zero
        for bkt in reversed(buckets):
            running = curve.jadd(running, bkt)
            res = curve.jadd(res, running)
        window_sums.append(res)
    total = window_sums[-1]
    for ws in reversed(window_sums[:-1]):
        for _ in range(c): total = curve.jdbl(total)
        total = curve.jadd(total, ws)
    return curve.from_jac(total)


# ----------------------------------------------------------------------------- deterministic synthetic data
MASK64 = (1 << 64) - 1

class SplitMix64:
    """The PRNG every layer (Python oracle, C++ oracle, CUDA host) uses for synthetic inputs."""
    def __init__(self, seed: int): self.s = seed & MASK64
    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)
    def fr(self) -> int:
        """4 x u64 little-endian limbs, reduced mod r (tiny modulo bias, irrelevant here)."""
        v = 0
        for i in range(4): v |= self.next() << (64 * i)
        return v % R
