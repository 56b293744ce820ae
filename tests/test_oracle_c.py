"""Pins the C++ oracle (oracle/c, the CPU baseline) against the Python big-int oracle and the
SURVEY.md Appendix A known answers.  CPU only."""
import random

import numpy as np

from oracle import corac
from oracle.pyref import bls12_381 as bls
from oracle.pyref import groth16 as og
from oracle.pyref import poseidon as pos
from oracle.pyref.algos import Domain, msm_pippenger
from oracle.pyref.r1cs import ConstraintSystem, LC, assert_equal, is_zero, mul, range_bits, select
from tests import util

R, P = bls.R, bls.P


def test_field_mul():
    rnd = random.Random(1)
    for field, mod, enc in ((0, R, bls.fr_to_mont_bytes), (1, P, bls.fq_to_mont_bytes)):
        A = [rnd.randrange(mod) for _ in range(300)] + [0, 1, mod - 1]
        B = [rnd.randrange(mod) for _ in range(300)] + [mod - 1, mod - 1, mod - 1]
        a = np.frombuffer(b"".join(enc(v) for v in A), dtype=np.uint8)
        b = np.frombuffer(b"".join(enc(v) for v in B), dtype=np.uint8)
        assert bytes(corac.field_mul(field, a, b)) == b"".join(enc(x * y % mod) for x, y in zip(A, B))


def test_ntt_vs_python():
    off = bls.fr_to_mont_bytes(7)
    for log_n in (0, 1, 2, 5, 10, 12):
        n = 1 << log_n
        x = util.rand_fr(log_n, n)
        buf = util.fr_mont_array(x)
        d, c = Domain(n), Domain(n, 7)
        assert util.fr_from_mont_array(corac.ntt(buf, log_n)) == d.fft(x)
        assert util.fr_from_mont_array(corac.ntt(buf, log_n, inverse=True)) == d.ifft(x)
        assert util.fr_from_mont_array(corac.ntt(buf, log_n, offset=off)) == c.fft(x)
        assert util.fr_from_mont_array(corac.ntt(buf, log_n, inverse=True, offset=off)) == c.ifft(x)
    assert util.fr_from_mont_array(corac.ntt(util.fr_mont_array([1, 2, 3, 4]), 2))[0] == 0xa


def test_msm_vs_python():
    for group, cv, enc, dec in ((1, bls.G1, util.g1_array, util.g1_list), (2, bls.G2, util.g2_array, util.g2_list)):
        bases = [cv.mul(cv.gen, k) for k in (1, 2, 3, 4)]
        assert dec(corac.msm(group, enc(bases), util.scalars_array([1, 2, 3, 4])))[0] == cv.mul(cv.gen, 30)
        assert dec(corac.msm(group, enc(bases), util.scalars_array([R - 1, 1, 0, 2])))[0] == cv.mul(cv.gen, 9)
        n = 70 if group == 1 else 40
        ks = util.rand_fr(3, n)
        pts = corac.fixed_base_mul(group, bytes(enc([cv.gen])), util.scalars_array(ks))
        assert dec(pts)[:3] == [cv.mul(cv.gen, k) for k in ks[:3]]
        ss = util.rand_fr(4, n)
        got = dec(corac.msm(group, pts, util.scalars_array(ss)))[0]
        assert got == cv.mul(cv.gen, sum(a * b for a, b in zip(ks, ss)) % R)
        assert got == msm_pippenger(cv, dec(pts), ss)


def test_poseidon_vs_python():
    for arity in (1, 2, 4, 5):
        vals = util.rand_fr(arity, 6 * arity)
        out = util.fr_from_mont_array(corac.poseidon_hash_batch(util.fr_mont_array(vals), arity))
        assert out == [pos.hash_fix_len_array(vals[i * arity:(i + 1) * arity]) for i in range(6)]


def test_groth16_prove_vs_python():
    cs = ConstraintSystem()
    out = cs.alloc_input(35); x = cs.alloc_witness(3)
    x2 = mul(cs, x, x); x3 = mul(cs, x2, x)
    assert_equal(cs, x3 + x + 5, out)
    iz = is_zero(cs, x - 3); sel = select(cs, x, x2, iz); assert_equal(cs, sel, LC.const(3))
    range_bits(cs, x, 4)
    assert cs.is_satisfied()
    tox = og.Toxic(11, 22, 33, 44, 55)
    M = cs.matrices()
    pk = og.generate_parameters(M, cs.num_inputs, cs.num_variables, tox)
    want = og.create_proof_with_reduction(M, pk, cs.z, 123, 456)
    key = dict(alpha_g1=util.g1_array([pk.alpha_g1]), beta_g1=util.g1_array([pk.beta_g1]), beta_g2=util.g2_array([pk.beta_g2]),
               delta_g1=util.g1_array([pk.delta_g1]), delta_g2=util.g2_array([pk.delta_g2]), a_query=util.g1_array(pk.a_query),
               b_g1_query=util.g1_array(pk.b_g1_query), b_g2_query=util.g2_array(pk.b_g2_query),
               l_query=util.g1_array(pk.l_query), h_query=util.g1_array(pk.h_query))
    got = corac.groth16_prove(corac.CsrMatrices.from_rows(M), cs.num_constraints, cs.num_inputs, cs.num_variables,
                              pk.n.bit_length() - 1, key, util.fr_mont_array(cs.z), 123, 456)
    assert (bls.g1_from_ffi(got[:96]), bls.g2_from_ffi(got[96:288]), bls.g1_from_ffi(got[288:])) == want
    assert og.verify(pk, [35], want)
