"""Batched Groth16 verifier and wire formats on the GPU, through the C ABI:
verdicts against the Python oracle's pairing check (oracle/pyref/groth16.py verify_with_vk), serialized keys
bit-exact against the Python oracle's compressed encoding (tests/golden/groth16_update_note.json)."""
import json
import os

import numpy as np
import pytest

import zk_apps_b200 as z
from oracle.pyref import bls12_381 as bls
from oracle.pyref import groth16 as og
from oracle.pyref import relations as rel
from tests import util

pytestmark = pytest.mark.gpu
R = bls.R
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_TOX = (11, 22, 33, 44, 55)


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "groth16_update_note.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def key(ctx):
    relation = z.UpdateNoteRelation(rel.WITHDRAW, rel.TREE_HEIGHT)
    pk = z.Groth16.generate_parameters_with_toxic_waste(ctx, relation, GOLDEN_TOX, precompute=True)
    vk = pk.verifying_key()
    yield relation, pk, vk
    vk.free()
    pk.free()


@pytest.fixture(scope="module")
def proven(ctx, key):
    """12 proofs of distinct withdraw instances + their public inputs (Montgomery bytes)."""
    relation, pk, vk = key
    batch = 12
    ws = [rel.make_witness(200 + i, rel.WITHDRAW) for i in range(batch)]
    inputs = util.fr_mont_array([v for w in ws for v in rel.witness_to_inputs(w)])
    proofs, status = z.Groth16.prove_update_note(pk, inputs, util.rand_fr(5, batch), util.rand_fr(6, batch), batch)
    assert list(status) == [0] * batch
    pub = [w.public_inputs() for w in ws]
    return ws, proofs, pub


def pub_bytes(pub):
    return util.fr_mont_array([v for row in pub for v in row])


def test_serialized_vk_matches_python_oracle(ctx, key, golden):
    """vk.serialize() == the bytes the Python oracle wrote; deserialize(serialize) round-trips."""
    relation, pk, vk = key
    want = bytes.fromhex(golden["withdraw"]["vk_compressed_hex"])
    got = vk.serialize()
    assert got == want
    vk2 = z.VerifyingKey.deserialize(ctx, got)
    assert bytes(vk2.export()) == bytes(vk.export()) == bytes(pk.vk)
    vk2.free()
    with pytest.raises(z.B200zkError) as e:
        z.VerifyingKey.deserialize(ctx, got[:-1])
    assert e.value.code == -11
    bad = bytearray(got); bad[0] &= 0x7F                           # alpha_g1 without the compression flag
    with pytest.raises(z.B200zkError) as e:
        z.VerifyingKey.deserialize(ctx, bytes(bad))
    assert e.value.code == -11


def test_golden_proof_verifies_on_gpu(ctx, key, golden):
    relation, pk, vk = key
    g = golden["withdraw"]
    proof = np.frombuffer(bytes.fromhex(g["proof_hex"]), dtype=np.uint8)
    xs = [int(x, 16) for x in g["public_inputs"]]
    assert list(z.Groth16.verify_proofs(vk, proof, util.fr_mont_array(xs))) == [0]
    xs[3] = (xs[3] + 1) % R
    assert list(z.Groth16.verify_proofs(vk, proof, util.fr_mont_array(xs))) == [1]


def test_verify_batch_verdicts_match_oracle(ctx, key, proven):
    relation, pk, vk = key
    ws, proofs, pub = proven
    batch = len(ws)
    st = z.Groth16.verify_proofs(vk, proofs, pub_bytes(pub))
    assert list(st) == [0] * batch
    ovk = (bls.g1_from_ffi(bytes(pk.vk[:96])), bls.g2_from_ffi(bytes(pk.vk[96:288])), bls.g2_from_ffi(bytes(pk.vk[288:480])),
           bls.g2_from_ffi(bytes(pk.vk[480:672])), util.g1_list(bytes(pk.vk[672:])))
    assert og.verify_with_vk(ovk, pub[0], og.proof_from_bytes(bytes(proofs[:192])))
    # a batch with every kind of failure mixed among valid proofs: verdicts are per proof
    pr = np.array(proofs, dtype=np.uint8).reshape(batch, 192).copy()
    px = [list(p) for p in pub]
    px[1][4] = (px[1][4] + 1) % R                                   # wrong merkle_root            -> rejected
    pr[2], pr[3] = pr[3].copy(), pr[2].copy()                       # proofs swapped               -> rejected x2
    pr[4][0] &= 0x7F                                                # A without compression flag   -> bad encoding
    pr[5][48] |= 0x40                                               # B "infinity" with x bits set -> bad encoding
    A, B, Cc = og.proof_from_bytes(bytes(pr[6]))
    pr[6] = np.frombuffer(og.proof_to_bytes((A, B, bls.G1.neg(Cc))), dtype=np.uint8)   # -C          -> rejected
    x = int.from_bytes(bytes([pr[7][144] & 0x1F]) + bytes(pr[7][145:192]), "big")
    while bls.fq_sqrt((x * x * x + 4) % bls.P) is not None:
        x += 1
    pr[7][144:192] = np.frombuffer(bytes([0x80 | (x >> 376)]) + (x & ((1 << 376) - 1)).to_bytes(47, "big"), dtype=np.uint8)  # C off curve
    xg = 5
    while True:
        yg = bls.fq_sqrt((xg ** 3 + 4) % bls.P)
        if yg is not None and bls.G1.mul((xg, yg), R) is not None:
            break
        xg += 1
    pr[8][:48] = np.frombuffer(bls.g1_compress((xg, yg)), dtype=np.uint8)               # A outside G1 -> not in subgroup
    xb = pub_bytes(px).reshape(batch, -1).copy()
    xb[9][:32] = 0xFF                                               # unreduced public input       -> bad input
    want = [0, 1, 1, 1, 2, 2, 1, 3, 4, 5, 0, 0]
    st = z.Groth16.verify_proofs(vk, pr.reshape(-1), xb.reshape(-1))
    assert list(st) == want
    # the oracle agrees on the well-formed ones
    for i in (1, 2, 6):
        assert not og.verify_with_vk(ovk, px[i], og.proof_from_bytes(bytes(pr[i])))
    # without the subgroup test the stray point decodes and the equation rejects it
    st = z.Groth16.verify_proofs(vk, pr.reshape(-1), xb.reshape(-1), check_subgroup=False)
    assert list(st) == [0, 1, 1, 1, 2, 2, 1, 3, 1, 5, 0, 0]


def test_verify_device_buffers_and_empty_batch(ctx, key, proven):
    relation, pk, vk = key
    ws, proofs, pub = proven
    xb = pub_bytes(pub)
    dp, dx = ctx.alloc(proofs.size), ctx.alloc(xb.size)
    ctx.upload(dp, proofs)
    ctx.upload(dx, xb)
    st = z.Groth16.verify_proofs(vk, dp, dx, batch=len(ws), device=True)
    assert list(st) == [0] * len(ws)
    ctx.free(dp)
    ctx.free(dx)
    assert list(z.Groth16.verify_proofs(vk, np.zeros(0, np.uint8), np.zeros(0, np.uint8), batch=0)) == []


def test_verify_aggregate(ctx, key, proven):
    relation, pk, vk = key
    ws, proofs, pub = proven
    batch = len(ws)
    rng = np.random.default_rng(7)
    coeffs = rng.integers(0, 256, size=batch * 16, dtype=np.uint8)
    xb = pub_bytes(pub)
    assert z.Groth16.verify_proofs_aggregate(vk, proofs, xb, coeffs) is True
    assert z.Groth16.verify_proofs_aggregate(vk, proofs[:192], xb[:6 * 32], coeffs[:16]) is True
    bad = xb.copy().reshape(batch, -1)
    bad[batch - 1] = bad[0]                                         # one instance with another one's inputs
    assert z.Groth16.verify_proofs_aggregate(vk, proofs, bad.reshape(-1), coeffs) is False
    pr = np.array(proofs, dtype=np.uint8).copy()
    pr[5 * 192] &= 0x7F                                             # one malformed proof
    assert z.Groth16.verify_proofs_aggregate(vk, pr, xb, coeffs) is False
    # two forgeries that cancel without randomisers (C_0 + D, C_1 - D) must not pass with them
    p0, p1 = og.proof_from_bytes(bytes(proofs[:192])), og.proof_from_bytes(bytes(proofs[192:384]))
    D = bls.G1.mul(bls.G1_GEN, 12345)
    f0 = og.proof_to_bytes((p0[0], p0[1], bls.G1.add(p0[2], D)))
    f1 = og.proof_to_bytes((p1[0], p1[1], bls.G1.add(p1[2], bls.G1.neg(D))))
    forged = np.frombuffer(f0 + f1, dtype=np.uint8)
    assert z.Groth16.verify_proofs_aggregate(vk, forged, xb[:2 * 6 * 32], coeffs[:32]) is False
    ones = np.zeros(32, dtype=np.uint8); ones[0] = 1; ones[16] = 1  # equal randomisers: the forgeries do cancel
    assert z.Groth16.verify_proofs_aggregate(vk, forged, xb[:2 * 6 * 32], ones) is True
    assert list(z.Groth16.verify_proofs(vk, forged, xb[:2 * 6 * 32])) == [1, 1]


def test_points_wire_format_matches_oracle(ctx, golden):
    with open(os.path.join(ROOT, "tests", "golden", "curve.json")) as f:
        cv = json.load(f)
    ks = [1, 2, 9, 30, R - 1, 0xdeadbeef]
    g1 = [bls.G1.mul(bls.G1_GEN, k) for k in ks] + [None]
    g2 = [bls.G2.mul(bls.G2_GEN, k) for k in ks] + [None]
    c1 = bytes(z.points_compress(ctx, 1, util.g1_array(g1)))
    c2 = bytes(z.points_compress(ctx, 2, util.g2_array(g2)))
    assert c1 == b"".join(bls.g1_compress(p) for p in g1)
    assert c2 == b"".join(bls.g2_compress(p) for p in g2)
    assert c1[:48].hex() == cv["g1_gen_compressed"] and c1[96:144].hex() == cv["g1_9G_compressed"]
    assert c2[:96].hex() == cv["g2_gen_compressed"] and c2[288:384].hex() == cv["g2_30G_compressed"]
    a1, s1 = z.points_decompress(ctx, 1, np.frombuffer(c1, dtype=np.uint8))
    a2, s2 = z.points_decompress(ctx, 2, np.frombuffer(c2, dtype=np.uint8))
    assert list(s1) == [0] * len(g1) and list(s2) == [0] * len(g2)
    assert util.g1_list(a1) == g1 and util.g2_list(a2) == g2


def test_pk_serialize_roundtrip_proves_identically(ctx, key, proven, golden):
    relation, pk, vk = key
    blob = pk.serialize(vk)
    want_vk = bytes.fromhex(golden["withdraw"]["vk_compressed_hex"])
    assert blob[:len(want_vk)] == want_vk
    nv, n, ni = relation.num_variables, 8192, relation.num_inputs
    assert len(blob) == len(want_vk) + 96 + 5 * 8 + 48 * (2 * nv + (n - 1) + (nv - ni)) + 96 * nv
    # a query in the blob equals the Python oracle's compression of the exported points
    off = len(want_vk) + 96
    assert int.from_bytes(blob[off:off + 8], "little") == nv
    a_pts = util.g1_list(pk.export_query(0))
    for i in (0, 1, nv - 1):
        assert blob[off + 8 + 48 * i:off + 56 + 48 * i] == bls.g1_compress(a_pts[i])
    pk2 = z.ProvingKey.deserialize(ctx, relation, blob, check_subgroup=False, precompute=True)
    assert bytes(pk2.vk) == bytes(pk.vk)
    ws = [rel.make_witness(300 + i, rel.WITHDRAW) for i in range(2)]
    inputs = util.fr_mont_array([v for w in ws for v in rel.witness_to_inputs(w)])
    p1, _ = z.Groth16.prove_update_note(pk, inputs, [3, 4], [5, 6], 2)
    p2, _ = z.Groth16.prove_update_note(pk2, inputs, [3, 4], [5, 6], 2)
    assert bytes(p1) == bytes(p2)
    assert list(z.Groth16.verify_proofs(vk, p2, pub_bytes([w.public_inputs() for w in ws]))) == [0, 0]
    pk2.free()
    with pytest.raises(z.B200zkError) as e:
        z.ProvingKey.deserialize(ctx, relation, blob[:-48])
    assert e.value.code == -11
