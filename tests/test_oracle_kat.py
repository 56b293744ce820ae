"""Pins the oracles (oracle/pyref and oracle/c) against the SURVEY.md Appendix A known answers
(written out literally below) and the committed fixtures in tests/golden/ (tests/golden/make_golden.py).
CPU only.  The reference holds no vectors for this path, so Appendix A --
mathematical facts about BLS12-381 -- is the anchor."""
import hashlib
import json
import os

import numpy as np

from oracle import corac
from oracle.pyref import bls12_381 as bls, groth16 as og, poseidon as pos, relations as rel
from oracle.pyref.algos import Domain, SplitMix64, ark_window, msm_pippenger
from tests import util

R, P = bls.R, bls.P
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return json.load(open(os.path.join(GOLD, name)))


def test_appendix_a_field_constants():
    assert P == 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    assert R == 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    assert bls.FQ_MONT_R == 0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd
    assert pow(2, 768, P) == 0x11988fe592cae3aa9a793e85b519952d67eb88a9939d83c08de5476c4c95b6d50a76e6a609d104f1f4df1f341c341746
    assert (-pow(P, -1, 1 << 64)) % (1 << 64) == 0x89f3fffcfffcfffd
    assert bls.FR_MONT_R == 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe
    assert pow(2, 512, R) == 0x0748d9d99f59ff1105d314967254398f2b6cedcb87925c23c999e990f3f29c6d
    assert (-pow(R, -1, 1 << 64)) % (1 << 64) == 0xfffffffeffffffff
    w = bls.FR_ROOT_2_32
    assert w == 10238227357739495823651030575849232062558860180284477541189508159991286009131
    assert pow(w, 1 << 32, R) == 1 and pow(w, 1 << 31, R) == R - 1
    assert pow(w, 1 << 30, R) == 0x8d51ccce760304d0ec030002760300000001000000000000
    assert pow(w, 1 << 29, R) == 0x345766f603fa66e78c0625cd70d77ce2b38b21c28713b7007228fd3397743f7a
    g = gold("field.json")
    assert int(g["fq_R2"], 16) == pow(2, 768, P) and int(g["omega_8"], 16) == pow(w, 1 << 29, R)
    assert bls.fr_to_mont_bytes(7).hex() == g["fr_mont_bytes_of_7"]


def test_appendix_a_points_and_compression():
    g1 = bls.G1
    assert g1.on_curve(g1.gen) and bls.G2.on_curve(bls.G2.gen)
    assert g1.mul(g1.gen, R) is None
    p2 = g1.mul(g1.gen, 2)
    assert p2[0] == 0x0572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e
    assert p2[1] == 0x166a9d8cabc673a322fda673779d8e3822ba3ecb8670e461f73bb9021d5fd76a4c56d9d4cd16bd1bba86881979749d28
    assert bls.g1_compress(g1.gen).hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                                              "6c55e83ff97a1aeffb3af00adb22c6bb")
    p30 = g1.mul(g1.gen, 30)
    assert bls.g1_compress(p30).hex() == ("ad84464b3966ec5bede84aa487facfca7823af383715078da03b387cc2f5d559"
                                           "7cdd7d025aa07db00a38b953bdeb6e3f")
    g = gold("curve.json")
    for k in (2, 9, 30):
        assert bls.g1_compress(g1.mul(g1.gen, k)).hex() == g["g1_%dG_compressed" % k]
        q = bls.G2.mul(bls.G2.gen, k)
        assert bls.g2_compress(q).hex() == g["g2_%dG_compressed" % k]
        assert bls.g2_decompress(bytes.fromhex(g["g2_%dG_compressed" % k])) == q
    assert bls.g1_decompress(bls.g1_compress(p30)) == p30
    assert bls.g1_compress(None)[0] == 0xC0


def test_appendix_a_msm_kats_both_oracles():
    for group, cv, enc, dec in ((1, bls.G1, util.g1_array, util.g1_list), (2, bls.G2, util.g2_array, util.g2_list)):
        bases = [cv.mul(cv.gen, k) for k in (1, 2, 3, 4)]
        for scalars, k in (([1, 2, 3, 4], 30), ([R - 1, 1, 0, 2], 9)):
            want = cv.mul(cv.gen, k)
            assert cv.msm_naive(bases, scalars) == want
            assert msm_pippenger(cv, bases, scalars) == want
            assert dec(corac.msm(group, enc(bases), util.scalars_array(scalars)))[0] == want
    p9 = bls.G1.mul(bls.G1.gen, 9)
    assert p9[0] == 0x19cdf3807146e68e041314ca93e1fee0991224ec2a74beb2866816fd0826ce7b6263ee31e953a86d1b72cc2215a57793
    assert p9[1] == 0x07481b1f261aabacf45c6e4fc278055441bfaf99f604d1f835c0752ac9742b4522c9f5c77db40989e7da608505d48616
    assert ark_window(1 << 24) == 18 and ark_window(1 << 16) == 13 and ark_window(31) == 3


def test_golden_msm_both_oracles():
    g = gold("msm.json")
    for group, name, cv, enc in ((1, "g1", bls.G1, util.g1_array), (2, "g2", bls.G2, util.g2_array)):
        bases = [cv.mul(cv.gen, k) for k in g[name]["base_multiples_of_generator"]]
        scalars = [int(s, 16) for s in g[name]["scalars"]]
        want = bytes.fromhex(g[name]["result_ffi"])
        assert bytes(corac.msm(group, enc(bases), util.scalars_array(scalars))) == want
        assert bytes(enc([msm_pippenger(cv, bases, scalars)])) == want


def test_appendix_a_ntt_kats_both_oracles():
    want = [0xa, 0x73eda753299d7d4718963e6b1d9bce637bb7a3fe13f85bfefffdfffeffffffff,
            0x73eda753299d7d483339d80809a1d80553bda402fffe5bfefffffffeffffffff,
            0x11aa3999cec0609a1d8060004ec0600000001fffffffffffe]
    cwant = [0x5fe, 0x73eda753299d7a5a8b4d68d2059e4bc15bd396f4fc145bfefab1fffeffffff6f,
             0x73eda753299d7d483339d80809a1d80553bda402fffe5bfefffffffefffffb2b,
             0x2eda7ec6f3604038c43f7ea0d0e03ea0000054dffffffffff6e]
    assert Domain(4).fft([1, 2, 3, 4]) == want and Domain(4).dft_naive([1, 2, 3, 4]) == want
    assert Domain(4, 7).fft([1, 2, 3, 4]) == cwant
    assert Domain(4, 7).vanishing_on_coset() == 0x960
    buf = util.fr_mont_array([1, 2, 3, 4])
    assert util.fr_from_mont_array(corac.ntt(buf, 2)) == want
    assert util.fr_from_mont_array(corac.ntt(buf, 2, offset=bls.fr_to_mont_bytes(7))) == cwant


def test_golden_ntt_both_oracles():
    g = gold("ntt.json")
    rng = SplitMix64(0xB2000002)
    off = bls.fr_to_mont_bytes(7)
    sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()
    for log_n in (6, 11):
        n = 1 << log_n
        x = [rng.fr() for _ in range(n)]
        e = g["log%d" % log_n]
        buf = util.fr_mont_array(x)
        assert sha(buf) == e["input_sha256"]
        d, c = Domain(n), Domain(n, 7)
        assert sha(util.fr_mont_array(d.fft(x))) == e["fft_sha256"]
        assert sha(util.fr_mont_array(c.ifft(x))) == e["coset7_ifft_sha256"]
        assert sha(corac.ntt(buf, log_n)) == e["fft_sha256"]
        assert sha(corac.ntt(buf, log_n, inverse=True)) == e["ifft_sha256"]
        assert sha(corac.ntt(buf, log_n, offset=off)) == e["coset7_fft_sha256"]
        assert sha(corac.ntt(buf, log_n, inverse=True, offset=off)) == e["coset7_ifft_sha256"]


def test_poseidon_parameters_and_golden():
    assert (R - 1) % 5 != 0 and (R - 1) % 3 == 0                                  # alpha = 5 is a permutation of Fr, alpha = 3 is not
    g = gold("poseidon.json")
    rc, mds = pos.constants()
    flat = [v for row in rc for v in row]
    assert len(rc) == 64 and len(rc[0]) == 5 and len(mds) == 5
    assert hashlib.sha256(b"".join(bls.fr_to_mont_bytes(v) for v in flat)).hexdigest() == g["round_constants_sha256"]
    assert int(g["mds_00"], 16) == mds[0][0]
    assert [int(v, 16) for v in g["permute_of_0_1_2_3_4"]] == pos.permute([0, 1, 2, 3, 4])
    for key, vals in (("hash_1", [1]), ("hash_1_2", [1, 2]), ("hash_1_2_3_4", [1, 2, 3, 4]), ("hash_1_2_3_4_5", [1, 2, 3, 4, 5])):
        assert pos.hash_fix_len_array(vals) == int(g[key], 16)
        got = util.fr_from_mont_array(corac.poseidon_hash_batch(util.fr_mont_array(vals), len(vals)))
        assert got == [int(g[key], 16)]
    assert pos.n_permutations(4) == 2 and pos.n_permutations(2) == 1


def test_golden_update_note_proofs():
    """Fixed toxic waste and r/s -> the committed 192-byte proofs; each satisfies the pairing equation."""
    g = gold("groth16_update_note.json")
    for kind, name in ((rel.DEPOSIT, "deposit"), (rel.WITHDRAW, "withdraw")):
        e = g[name]
        w = rel.make_witness(e["make_witness_seed"], kind)
        cs = rel.synthesize_update_note(w)
        assert cs.is_satisfied()
        assert (cs.num_constraints, cs.num_inputs, cs.num_variables) == (e["num_constraints"], e["num_inputs"], e["num_variables"])
        assert hashlib.sha256(b"".join(bls.fr_to_mont_bytes(v) for v in cs.z)).hexdigest() == e["assignment_sha256"]
        assert [int(v, 16) for v in e["public_inputs"]] == w.public_inputs()
        tox = og.Toxic(*e["toxic"])
        M = cs.matrices()
        sc = og.setup_scalars(M, cs.num_inputs, cs.num_variables, tox)
        proof = og.proof_via_scalars(M, sc, tox, cs.z, e["r"], e["s"])
        assert og.proof_to_bytes(proof).hex() == e["proof_hex"]
        assert og.verify_with_vk(og.verifying_key_from_toxic(sc, tox), w.public_inputs(), proof)
        bad = list(w.public_inputs())
        bad[0] = (bad[0] + 1) % R
        assert not og.verify_with_vk(og.verifying_key_from_toxic(sc, tox), bad, proof)


def test_relation_rejects_bad_witnesses():
    """Edge cases of the statement (update_note.rs:129-148): each broken witness leaves a constraint unsatisfied."""
    w = rel.make_witness(3, rel.WITHDRAW)
    assert rel.synthesize_update_note(w).is_satisfied()
    import copy
    for mutate in ("root", "nullifier", "user", "path"):
        b = copy.deepcopy(w)
        if mutate == "root":
            b.merkle_root = (b.merkle_root + 1) % R
        elif mutate == "nullifier":
            b.old_note.nullifier = (b.old_note.nullifier + 1) % R
        elif mutate == "user":
            b.op_priv_user = (b.op_priv_user + 1) % R
        else:
            b.path[0] = (b.path[0] + 1) % R
        assert not rel.synthesize_update_note(b).is_satisfied(), mutate


def test_poseidon_generator_and_permutation_match_the_public_bn254_instance():
    """Third-party pin of the Poseidon constants generator (Grain LFSR, rejection sampling, Cauchy MDS from xs + ys)
    and of the permutation's round structure: for the public instance x5_254_3 (BN254 scalar field, t = 3, R_F = 8,
    R_P = 57 -- the one circomlib / iden3 ship) the same code must reproduce the published numbers.  None of the
    literals below comes from this repository: they are the first round constants and MDS entries of
    poseidonperm_x5_254_3 / circomlib's poseidon_constants and circomlibjs' test vector poseidon([1, 2]).  The
    shielder instance (BLS12-381 Fr, t = 5, R_P = 56) runs through exactly the same generator and permutation code
    (oracle/pyref/poseidon.py: constants(), permute_params()), and the C++ generator of the product
    (csrc/host_r1cs.hpp) is diffed against this one in tests/test_host.py."""
    from oracle.pyref import poseidon as pos
    p_bn = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rc, mds = pos.constants(3, 8, 57, p_bn, 254)
    assert rc[0][0] == 0x0ee9a592ba9a9518d05986d656f40c2114c4993c11bb29938d21d47304cd8e6e
    assert rc[0][1] == 0x00f1445235f2148c5986587169fc1bcd887b08d4d00868df5696fff40956e864
    assert rc[0][2] == 0x08dff3487e8ac99e1f29a058d0fa80b930c728730b7ab36ce879f3890ecf73f5
    assert mds[0][0] == 0x109b7f411ba0e4c9b2b70caf5c36a7b194be7c11ad24378bfedb68592ba8118b
    assert mds[0][1] == 0x16ed41e13bb9c0c66ae119424fddbcbc9314dc9fdbdeea55d6c64543dc4903e0
    assert mds[0][2] == 0x2b90bba00fca0589f617e7dcbfe82e0df706ab640ceb247b791a93b74e36736d
    assert len(rc) == 65 and all(len(r) == 3 for r in rc)
    # circomlib: state = [0, inputs...], digest = state[0] after the permutation
    out = pos.permute_params([0, 1, 2], 3, 8, 57, p_bn, 254)
    assert out[0] == 7853200120776062878684798364095072458815029376092732009249414926327459813530
    assert out[0] == 0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a
