"""K6 + Groth16 parity (GPU): batched Poseidon, the update-note witness generator and the full
prover against the Python oracle, bit-exact, through the C ABI; every proof must also satisfy the
pairing equation (the stand-in for "the reference verifier accepts")."""
import numpy as np
import pytest

import zk_apps_b200 as z
from oracle.pyref import bls12_381 as bls
from oracle.pyref import groth16 as og
from oracle.pyref import poseidon as pos
from oracle.pyref import relations as rel
from tests import util

pytestmark = pytest.mark.gpu
R = bls.R
TOX = og.Toxic(alpha=0x1111_2222_3333, beta=0x4444_5555_6666_7777, gamma=0x8888_9999, delta=0xaaaa_bbbb_cccc, tau=0xdddd_eeee_ffff_0123)


@pytest.mark.parametrize("arity", [1, 2, 3, 4, 5, 8])
def test_poseidon_hash_batch(ctx, arity):
    n = 37
    vals = util.rand_fr(arity, n * arity)
    vals[:arity] = [0] * arity
    vals[arity:2 * arity] = [R - 1] * arity
    out = util.fr_from_mont_array(z.poseidon_hash_batch(ctx, util.fr_mont_array(vals), arity))
    assert out == [pos.hash_fix_len_array(vals[i * arity:(i + 1) * arity]) for i in range(n)]


@pytest.mark.parametrize("kind", [rel.WITHDRAW, rel.DEPOSIT])
def test_witness_matches_oracle(ctx, kind):
    relation = z.UpdateNoteRelation(kind, rel.TREE_HEIGHT)
    ws = [rel.make_witness(10 + i, kind) for i in range(6)]
    inputs = util.fr_mont_array([v for w in ws for v in rel.witness_to_inputs(w)])
    out, status = relation.witness_batch(ctx, inputs, len(ws))
    assert list(status) == [0] * len(ws)
    nv = relation.num_variables
    for i, w in enumerate(ws):
        cs = rel.synthesize_update_note(w)
        assert cs.is_satisfied() and cs.num_variables == nv
        assert util.fr_from_mont_array(out[i * nv * 32:(i + 1) * nv * 32]) == cs.z


def test_unsatisfied_witness_is_reported(ctx):
    relation = z.UpdateNoteRelation(rel.WITHDRAW, rel.TREE_HEIGHT)
    good = rel.make_witness(3, rel.WITHDRAW)
    rows = []
    bad_root = rel.witness_to_inputs(good); bad_root[4] = (bad_root[4] + 1) % R          # wrong merkle_root
    bad_user = rel.witness_to_inputs(good); bad_user[2] = (bad_user[2] + 1) % R          # op_pub.user != op_priv.user
    overdraw = rel.witness_to_inputs(good); overdraw[0] = (1 << 127)                     # amount > balance -> underflow
    for r_ in (rel.witness_to_inputs(good), bad_root, bad_user, overdraw):
        rows += r_
    out, status = relation.witness_batch(ctx, util.fr_mont_array(rows), 4)
    assert list(status) == [0, 1, 1, 1]
    rc = z.lib().b200zk_update_note_witness_batch(ctx.handle, relation.handle, util.fr_mont_array(rows).ctypes.data, 4,
                                                  None, None, None)
    assert rc == -6                                                                       # B200ZK_ERR_UNSATISFIED


def test_non_canonical_input_rows_are_unsatisfied_not_a_hang(ctx):
    """input rows are caller-supplied words; one holding the unreduced word r (or anything >= r) used to reach
    fp_inv and spin.  Such an instance must come back as `unsatisfied`, the other instances of the batch unaffected."""
    relation = z.UpdateNoteRelation(rel.WITHDRAW, rel.TREE_HEIGHT)
    good = rel.witness_to_inputs(rel.make_witness(5, rel.WITHDRAW))
    arr = util.fr_mont_array(good * 4).copy().reshape(4, len(good), 32)
    r_word = np.frombuffer(R.to_bytes(32, "little"), dtype=np.uint8)
    arr[1, 13] = r_word                                   # path_shape[0] = the word r  (fed to fp_inv before the fix)
    arr[2, 14 + 2 * rel.TREE_HEIGHT] = r_word             # old_account.token0 = r      (token difference inversion)
    arr[3, 0] = 0xFF                                      # amount = 2^256 - 1
    out, status = relation.witness_batch(ctx, arr.reshape(-1), 4)
    assert list(status) == [0, 1, 1, 1]


@pytest.fixture(scope="module")
def withdraw_key(ctx):
    relation = z.UpdateNoteRelation(rel.WITHDRAW, rel.TREE_HEIGHT)
    pk = z.Groth16.generate_parameters_with_toxic_waste(ctx, relation, (TOX.alpha, TOX.beta, TOX.gamma, TOX.delta, TOX.tau),
                                                        precompute=True)
    w0 = rel.make_witness(1, rel.WITHDRAW)
    cs0 = rel.synthesize_update_note(w0)
    M = cs0.matrices()
    sc = og.setup_scalars(M, cs0.num_inputs, cs0.num_variables, TOX)
    return relation, pk, M, sc


def test_setup_matches_oracle(ctx, withdraw_key):
    """Key generation: vk and sampled query points equal the oracle's scalars times the generator."""
    relation, pk, M, sc = withdraw_key
    vk = bytes(pk.vk)
    assert bls.g1_from_ffi(vk[:96]) == bls.G1.mul(bls.G1_GEN, TOX.alpha)
    assert bls.g2_from_ffi(vk[96:288]) == bls.G2.mul(bls.G2_GEN, TOX.beta)
    assert bls.g2_from_ffi(vk[288:480]) == bls.G2.mul(bls.G2_GEN, TOX.gamma)
    assert bls.g2_from_ffi(vk[480:672]) == bls.G2.mul(bls.G2_GEN, TOX.delta)
    assert util.g1_list(vk[672:]) == [bls.G1.mul(bls.G1_GEN, k) for k in sc.gamma_abc]
    for which, scal, grp in ((0, sc.a, 1), (1, sc.b, 1), (2, sc.b, 2), (3, sc.l, 1), (4, sc.h, 1)):
        q = pk.export_query(which)
        pts = util.g1_list(q) if grp == 1 else util.g2_list(q)
        assert len(pts) == len(scal)
        cv = bls.G1 if grp == 1 else bls.G2
        for i in (0, 1, 7, len(scal) // 2, len(scal) - 1):
            assert pts[i] == cv.mul(cv.gen, scal[i]), (which, i)


def test_proofs_bit_exact_and_verify(ctx, withdraw_key):
    """Fixed (r, s): GPU proof bytes == oracle proof bytes; the pairing equation holds; a wrong public
    input is rejected."""
    relation, pk, M, sc = withdraw_key
    ws = [rel.make_witness(40 + i, rel.WITHDRAW) for i in range(3)]
    zs = [rel.synthesize_update_note(w).z for w in ws]
    rs = [0x1234567890abcdef + i for i in range(3)]
    ss = [0xfedcba0987654321 * (i + 1) for i in range(3)]
    rs[2], ss[2] = 0, 0                                            # degenerate randomness still has to work
    zbuf = util.fr_mont_array([v for zz in zs for v in zz])
    proofs, points = z.Groth16.create_proof_with_reduction(pk, zbuf, rs, ss, batch=3, want_points=True)
    vk = og.verifying_key_from_toxic(sc, TOX)
    for i, w in enumerate(ws):
        want = og.proof_via_scalars(M, sc, TOX, zs[i], rs[i], ss[i])
        pb = bytes(proofs[i * 192:(i + 1) * 192])
        assert pb == og.proof_to_bytes(want)
        pt = bytes(points[i * 384:(i + 1) * 384])
        assert (bls.g1_from_ffi(pt[:96]), bls.g2_from_ffi(pt[96:288]), bls.g1_from_ffi(pt[288:])) == want
        assert og.proof_from_bytes(pb) == want
    assert og.verify_with_vk(vk, ws[0].public_inputs(), og.proof_from_bytes(bytes(proofs[:192])))
    wrong = ws[0].public_inputs(); wrong[0] = (wrong[0] + 1) % R
    assert not og.verify_with_vk(vk, wrong, og.proof_from_bytes(bytes(proofs[:192])))


def test_end_to_end_prove_update_note(ctx, withdraw_key):
    """The user-facing call: inputs -> witness (K6) -> proof, in one batch; deterministic for fixed r, s."""
    relation, pk, M, sc = withdraw_key
    batch = 9
    ws = [rel.make_witness(70 + i, rel.WITHDRAW) for i in range(batch)]
    inputs = util.fr_mont_array([v for w in ws for v in rel.witness_to_inputs(w)])
    rs = util.rand_fr(1, batch); ss = util.rand_fr(2, batch)
    proofs, status = z.Groth16.prove_update_note(pk, inputs, rs, ss, batch)
    assert list(status) == [0] * batch
    proofs2, _ = z.Groth16.prove_update_note(pk, inputs, rs, ss, batch)
    assert bytes(proofs) == bytes(proofs2)
    for i in (0, batch - 1):
        zz = rel.synthesize_update_note(ws[i]).z
        assert bytes(proofs[i * 192:(i + 1) * 192]) == og.proof_to_bytes(og.proof_via_scalars(M, sc, TOX, zz, rs[i], ss[i]))
    vk = og.verifying_key_from_toxic(sc, TOX)
    assert og.verify_with_vk(vk, ws[4].public_inputs(), og.proof_from_bytes(bytes(proofs[4 * 192:5 * 192])))
    bad = inputs.copy(); bad[4 * 32] ^= 1                          # corrupt merkle_root of instance 0
    with pytest.raises(z.B200zkError) as e:
        z.Groth16.prove_update_note(pk, bad, rs, ss, batch)
    assert e.value.code == -6


def test_deposit_relation_proves(ctx):
    relation = z.UpdateNoteRelation(rel.DEPOSIT, rel.TREE_HEIGHT)
    pk = z.Groth16.generate_parameters_with_toxic_waste(ctx, relation, (TOX.alpha, TOX.beta, TOX.gamma, TOX.delta, TOX.tau),
                                                        precompute=False)
    w = rel.make_witness(5, rel.DEPOSIT)
    cs = rel.synthesize_update_note(w)
    M = cs.matrices()
    sc = og.setup_scalars(M, cs.num_inputs, cs.num_variables, TOX)
    inputs = util.fr_mont_array(rel.witness_to_inputs(w))
    proofs, status = z.Groth16.prove_update_note(pk, inputs, [77], [99], 1)
    assert bytes(proofs) == og.proof_to_bytes(og.proof_via_scalars(M, sc, TOX, cs.z, 77, 99))
    assert og.verify_with_vk(og.verifying_key_from_toxic(sc, TOX), w.public_inputs(), og.proof_from_bytes(bytes(proofs)))
    pk.free()


# ----------------------------------------------------------------------------- update-account as a relation of its own
@pytest.mark.parametrize("kind", [rel.WITHDRAW, rel.DEPOSIT])
def test_update_account_witness_matches_oracle(ctx, kind):
    """K6 for update_account_circuit (update_account.rs:68-95): the assignment equals the oracle's synthesis; wrong
    hashes, an overdraw, a token the account does not hold and an unreduced word are reported per instance."""
    relation = z.UpdateAccountRelation(kind)
    ws = [rel.make_account_witness(20 + i, kind) for i in range(5)]
    rows = [rel.account_witness_to_inputs(w) for w in ws]
    out, status = relation.witness_batch(ctx, util.fr_mont_array([v for r_ in rows for v in r_]), len(ws))
    assert list(status) == [0] * len(ws)
    nv = relation.num_variables
    for i, w in enumerate(ws):
        cs = rel.synthesize_update_account(w)
        assert cs.is_satisfied() and cs.num_variables == nv
        assert util.fr_from_mont_array(out[i * nv * 32:(i + 1) * nv * 32]) == cs.z
    good = rows[0]
    bad_old = list(good); bad_old[0] = (bad_old[0] + 1) % R
    bad_new = list(good); bad_new[1] = (bad_new[1] + 1) % R
    bad_tok = list(good); bad_tok[3] = (bad_tok[3] + 1) % R
    big = list(good); big[2] = 1 << 128
    arr = util.fr_mont_array(good + bad_old + bad_new + bad_tok + big + good).copy().reshape(6, 9, 32)
    arr[5, 6] = np.frombuffer(R.to_bytes(32, "little"), dtype=np.uint8)       # balance0 = the word r
    _, status = relation.witness_batch(ctx, arr.reshape(-1), 6)
    assert list(status) == [0, 1, 1, 1, 1, 1]
    relation.free()


def test_update_account_proofs_bit_exact_and_verify(ctx):
    """Key generation, witness generation and proving for the standalone update-account relation: proof bytes equal
    the oracle's for fixed (r, s) and satisfy the pairing equation against the oracle's verifying key."""
    relation = z.UpdateAccountRelation(rel.WITHDRAW)
    pk = z.Groth16.generate_parameters_with_toxic_waste(ctx, relation, (TOX.alpha, TOX.beta, TOX.gamma, TOX.delta, TOX.tau),
                                                        precompute=True)
    ws = [rel.make_account_witness(50 + i, rel.WITHDRAW) for i in range(3)]
    cs0 = rel.synthesize_update_account(ws[0])
    M = cs0.matrices()
    sc = og.setup_scalars(M, cs0.num_inputs, cs0.num_variables, TOX)
    inputs = util.fr_mont_array([v for w in ws for v in rel.account_witness_to_inputs(w)])
    rs, ss = [11, 0, 0x1234567890abcdef], [13, 0, 0xfedcba0987654321]
    proofs, status = z.Groth16.prove_update_note(pk, inputs, rs, ss, 3)
    assert list(status) == [0, 0, 0]
    vk = og.verifying_key_from_toxic(sc, TOX)
    for i, w in enumerate(ws):
        zz = rel.synthesize_update_account(w).z
        pb = bytes(proofs[i * 192:(i + 1) * 192])
        assert pb == og.proof_to_bytes(og.proof_via_scalars(M, sc, TOX, zz, rs[i], ss[i]))
        assert og.verify_with_vk(vk, w.public_inputs(), og.proof_from_bytes(pb))
    wrong = ws[0].public_inputs(); wrong[1] = (wrong[1] + 1) % R
    assert not og.verify_with_vk(vk, wrong, og.proof_from_bytes(bytes(proofs[:192])))
    # a key of the other relation is refused, not misused
    import ctypes as C
    out = np.zeros(3 * 192, dtype=np.uint8)
    rb = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in rs), dtype=np.uint8)
    note_rows = np.zeros(3 * (18 + 2 * rel.TREE_HEIGHT) * 32, dtype=np.uint8)
    rc = z.lib().b200zk_update_note_prove_batch(ctx.handle, pk._h, note_rows.ctypes.data_as(C.c_void_p), 3,
                                                rb.ctypes.data_as(C.c_void_p), rb.ctypes.data_as(C.c_void_p),
                                                out.ctypes.data_as(C.c_void_p), None)
    assert rc == -1                                                 # B200ZK_ERR_BAD_ARG: the key is update-account's
    pk.free()


@pytest.mark.parametrize("height", [4, 20])
def test_update_note_proves_at_other_tree_heights(ctx, height):
    """TREE_HEIGHT is a const generic in the reference (merkle_proof.rs:11; the mock fixes 10): witness, key and proof
    at another depth, bit-exact against the oracle and pairing-verified."""
    relation = z.UpdateNoteRelation(rel.WITHDRAW, height)
    pk = z.Groth16.generate_parameters_with_toxic_waste(ctx, relation, (TOX.alpha, TOX.beta, TOX.gamma, TOX.delta, TOX.tau),
                                                        precompute=False)
    w = rel.make_witness(90 + height, rel.WITHDRAW, height)
    cs = rel.synthesize_update_note(w, height)
    assert cs.is_satisfied()
    M = cs.matrices()
    sc = og.setup_scalars(M, cs.num_inputs, cs.num_variables, TOX)
    inputs = util.fr_mont_array(rel.witness_to_inputs(w))
    out, status = relation.witness_batch(ctx, inputs, 1)
    assert list(status) == [0] and util.fr_from_mont_array(out) == cs.z
    proofs, status = z.Groth16.prove_update_note(pk, inputs, [5], [6], 1)
    assert bytes(proofs) == og.proof_to_bytes(og.proof_via_scalars(M, sc, TOX, cs.z, 5, 6))
    assert og.verify_with_vk(og.verifying_key_from_toxic(sc, TOX), w.public_inputs(), og.proof_from_bytes(bytes(proofs)))
    pk.free()


def test_create_random_proof_draws_r_then_s(ctx, withdraw_key):
    """ark_groth16 `create_random_proof_with_reduction(circuit, pk, rng)` draws r, then s, from the caller's RNG and calls
    create_proof_with_reduction: the host mirror must consume the RNG in that order and produce a valid proof."""
    relation, pk, M, sc = withdraw_key
    w = rel.make_witness(123, rel.WITHDRAW)
    inputs = util.fr_mont_array(rel.witness_to_inputs(w))

    class CountingRng:
        def __init__(self):
            self.drawn = []
        def fr(self):
            v = (0x9E3779B97F4A7C15 * (len(self.drawn) + 1)) % R
            self.drawn.append(v)
            return v

    rng = CountingRng()
    proof = z.Groth16.create_random_proof(pk, inputs, rng)
    assert len(rng.drawn) == 2
    zz = rel.synthesize_update_note(w).z
    assert bytes(proof) == og.proof_to_bytes(og.proof_via_scalars(M, sc, TOX, zz, rng.drawn[0], rng.drawn[1]))   # r first, s second
    assert og.verify_with_vk(og.verifying_key_from_toxic(sc, TOX), w.public_inputs(), og.proof_from_bytes(bytes(proof)))


def test_async_submit_wait_matches_the_synchronous_call(ctx, withdraw_key):
    """b200zk_update_note_prove_submit / b200zk_prove_wait: two batches in flight give the same proof bytes as the
    synchronous call; the caller's input buffer may be reused right after submit; a third submit without a wait and
    a wait on an unknown ticket are refused; an unsatisfied instance surfaces at the wait."""
    relation, pk, M, sc = withdraw_key
    B = 5
    sets, want = [], []
    rs = util.scalars_array(util.rand_fr(31, B)); ss = util.scalars_array(util.rand_fr(32, B))
    for k in range(3):
        ws = [rel.make_witness(200 + 10 * k + i, rel.WITHDRAW) for i in range(B)]
        inp = util.fr_mont_array([v for w in ws for v in rel.witness_to_inputs(w)])
        sets.append(inp)
        want.append(bytes(z.Groth16.prove_update_note(pk, inp, rs, ss, B)[0]))
    outs = [np.zeros(B * 192, dtype=np.uint8) for _ in range(3)]
    sts = [np.full(B, 9, dtype=np.uint8) for _ in range(3)]
    scratch_in = sets[0].copy()
    t0 = z.Groth16.prove_submit(pk, scratch_in, rs, ss, B, outs[0], sts[0])
    scratch_in[:] = sets[1]                                            # the input buffer is free again after submit
    t1 = z.Groth16.prove_submit(pk, scratch_in, rs, ss, B, outs[1], sts[1])
    with pytest.raises(z.B200zkError):
        z.Groth16.prove_submit(pk, sets[2], rs, ss, B, outs[2], sts[2])     # two already in flight
    with pytest.raises(z.B200zkError):
        z.Groth16.prove_wait(pk, t1 + 100)
    z.Groth16.prove_wait(pk, t0)
    t2 = z.Groth16.prove_submit(pk, sets[2], rs, ss, B, outs[2], sts[2])
    z.Groth16.prove_wait(pk, t1)
    z.Groth16.prove_wait(pk, t2)
    for k in range(3):
        assert bytes(outs[k]) == want[k] and list(sts[k]) == [0] * B
    bad = sets[0].copy(); bad[4 * 32] ^= 1                             # merkle_root of instance 0
    out = np.zeros(B * 192, dtype=np.uint8); st = np.zeros(B, dtype=np.uint8)
    t = z.Groth16.prove_submit(pk, bad, rs, ss, B, out, st)
    with pytest.raises(z.B200zkError) as e:
        z.Groth16.prove_wait(pk, t)
    assert e.value.code == -6 and list(st) == [1, 0, 0, 0, 0] and not out.any()
    # the ctx is usable afterwards, synchronous calls included
    assert bytes(z.Groth16.prove_update_note(pk, sets[1], rs, ss, B)[0]) == want[1]


def test_proofs_over_full_digit_tables(ctx, withdraw_key):
    """A key uploaded with precompute level 2 (full digit tables, here with small windows so that they are megabytes,
    not the 135 GB of the bench configuration) gives the same proof bytes as the bucket-method key."""
    relation, pk, M, sc = withdraw_key
    ctx.set_option("table_c_g1", 4); ctx.set_option("table_c_g2", 3)
    pk2 = z.Groth16.generate_parameters_with_toxic_waste(ctx, relation, (TOX.alpha, TOX.beta, TOX.gamma, TOX.delta, TOX.tau),
                                                         precompute=2)
    ctx.set_option("table_c_g1", 13); ctx.set_option("table_c_g2", 13)
    B = 6
    ws = [rel.make_witness(300 + i, rel.WITHDRAW) for i in range(B)]
    inputs = util.fr_mont_array([v for w in ws for v in rel.witness_to_inputs(w)])
    rs = util.rand_fr(41, B); ss = util.rand_fr(42, B)
    want, _ = z.Groth16.prove_update_note(pk, inputs, rs, ss, B)
    got, status = z.Groth16.prove_update_note(pk2, inputs, rs, ss, B)
    assert list(status) == [0] * B and bytes(got) == bytes(want)
    zz = rel.synthesize_update_note(ws[0]).z
    assert bytes(got[:192]) == og.proof_to_bytes(og.proof_via_scalars(M, sc, TOX, zz, rs[0], ss[0]))
    pk2.free()
