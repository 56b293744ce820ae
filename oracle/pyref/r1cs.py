"""ORACLE (test infrastructure) -- a minimal R1CS constraint system and the gadgets the shielder
relation needs, in the shape of ark_relations::r1cs::ConstraintSystem ([recall]; ark-relations has
no pin anywhere in the reference, SURVEY.md section 8c -> PARITY UNPINNED).

Variable numbering follows arkworks: z = [1, instance variables..., witness variables...], i.e.
index 0 is the constant ONE, instance variables are 1..num_inputs-1 (num_inputs counts the ONE),
witness variables follow in allocation order.  Linear combinations are dicts {variable: coeff}.
"""
from __future__ import annotations
from .bls12_381 import R, finv
from . import poseidon as pos

ONE = 0


class LC(dict):
    """Linear combination sum coeff * var, coefficients mod r."""
    @staticmethod
    def var(v: int, c: int = 1) -> "LC":
        return LC({v: c % R}) if c % R else LC()

    @staticmethod
    def const(c: int) -> "LC":
        return LC.var(ONE, c)

    def __add__(self, o):
        if isinstance(o, int): o = LC.const(o)
        r = LC(self)
        for k, v in o.items():
            nv = (r.get(k, 0) + v) % R
            if nv: r[k] = nv
            else: r.pop(k, None)
        return r

    def __neg__(self): return self * (R - 1)
    def __sub__(self, o):
        if isinstance(o, int): o = LC.const(o)
        return self + (-o)

    def __mul__(self, c: int):
        c %= R
        return LC({k: v * c % R for k, v in self.items()}) if c else LC()

    def eval(self, z) -> int:
        return sum(c * z[v] for v, c in self.items()) % R


class ConstraintSystem:
    """Instance variables must all be allocated before the first witness variable (as the relation
    does), so indices are final at allocation time."""
    def __init__(self):
        self.num_inputs = 1                 # the constant ONE
        self.num_aux = 0
        self.z = [1]                        # full assignment
        self.A, self.B, self.C = [], [], []  # rows: LC

    def alloc_input(self, value: int) -> LC:
        assert self.num_aux == 0, "allocate instance variables first"
        self.z.append(value % R)
        self.num_inputs += 1
        return LC.var(self.num_inputs - 1)

    def alloc_witness(self, value: int) -> LC:
        self.z.append(value % R)
        self.num_aux += 1
        return LC.var(self.num_inputs + self.num_aux - 1)

    def enforce(self, a: LC, b: LC, c: LC):
        self.A.append(LC(a)); self.B.append(LC(b)); self.C.append(LC(c))

    @property
    def num_constraints(self): return len(self.A)
    @property
    def num_variables(self): return self.num_inputs + self.num_aux

    def val(self, lc: LC) -> int: return lc.eval(self.z)

    def is_satisfied(self) -> bool:
        return self.first_unsatisfied() is None

    def first_unsatisfied(self):
        for i, (a, b, c) in enumerate(zip(self.A, self.B, self.C)):
            if a.eval(self.z) * b.eval(self.z) % R != c.eval(self.z): return i
        return None

    def matrices(self):
        """CSR-like: per matrix a list of rows, each a sorted list of (variable, coeff)."""
        return tuple([sorted(r.items()) for r in M] for M in (self.A, self.B, self.C))


# ----------------------------------------------------------------------------- gadgets
def mul(cs: ConstraintSystem, a: LC, b: LC) -> LC:
    out = cs.alloc_witness(cs.val(a) * cs.val(b))
    cs.enforce(a, b, out)
    return out

def assert_equal(cs, a: LC, b: LC):
    cs.enforce(a - b, LC.const(1), LC())

def is_zero(cs, x: LC) -> LC:
    """out = 1 if x == 0 else 0.   x*inv = 1 - out ; x*out = 0.  (GateChip::is_zero)"""
    xv = cs.val(x)
    inv = cs.alloc_witness(finv(xv, R) if xv else 0)
    out = cs.alloc_witness(0 if xv else 1)
    cs.enforce(x, inv, LC.const(1) - out)
    cs.enforce(x, out, LC())
    return out

def is_equal(cs, a: LC, b: LC) -> LC:
    return is_zero(cs, a - b)

def select(cs, a: LC, b: LC, sel: LC) -> LC:
    """a if sel == 1 else b  (GateChip::select(a, b, sel)):  sel*(a-b) + b."""
    t = mul(cs, sel, a - b)
    return t + b

def range_bits(cs, x: LC, nbits: int):
    """Constrains x to [0, 2^nbits): boolean witnesses b_i with sum 2^i b_i = x."""
    xv = cs.val(x)
    acc = LC()
    for i in range(nbits):
        b = cs.alloc_witness((xv >> i) & 1)
        cs.enforce(b, b - 1, LC())
        acc = acc + b * (1 << i)
    assert_equal(cs, acc, x)

def poseidon_permute(cs, state):
    """Plain Poseidon permutation over LCs; every S-box allocates (x^2, x^4, x^5)."""
    rc, mds = pos.constants()
    t, half = pos.T_WIDTH, pos.R_F // 2
    s = list(state)
    for rnd in range(pos.R_F + pos.R_P):
        s = [s[i] + rc[rnd][i] for i in range(t)]
        full = rnd < half or rnd >= half + pos.R_P
        for i in range(t if full else 1):
            x2 = mul(cs, s[i], s[i]); x4 = mul(cs, x2, x2); s[i] = mul(cs, x4, s[i])
        ns = []
        for i in range(t):
            acc = LC()
            for j in range(t): acc = acc + s[j] * mds[i][j]
            ns.append(acc)
        s = ns
    return s

def poseidon_hash(cs, inputs) -> LC:
    """PoseidonHasher::hash_fix_len_array over LCs (see oracle/pyref/poseidon.py for the sponge)."""
    state = [LC.const(1 << 64)] + [LC() for _ in range(pos.T_WIDTH - 1)]
    chunks = [inputs[i:i + pos.RATE] for i in range(0, len(inputs), pos.RATE)]
    if len(inputs) % pos.RATE == 0: chunks.append([])
    for ch in chunks:
        for i, x in enumerate(ch): state[1 + i] = state[1 + i] + x
        if len(ch) + 1 < pos.T_WIDTH: state[len(ch) + 1] = state[len(ch) + 1] + 1
        state = poseidon_permute(cs, state)
    return state[1]
