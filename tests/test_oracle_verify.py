"""CPU: the verifier's per-proof logic (zk-apps_b200/csrc/verify.cuh, the code the verify kernels compile) on the
portable host path, against the golden proofs and the Python oracle's verdicts (oracle/pyref/groth16.py
verify_with_vk -- an independent pairing on a different representation).  Also pins the serialized VerifyingKey
layout (tests/golden/groth16_update_note.json vk_compressed_hex, written by the Python oracle)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle.pyref import bls12_381 as bls, groth16 as og

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = bls.R


@pytest.fixture(scope="module")
def hc(built):
    L = C.CDLL(os.path.join(ROOT, "tests", "host", "libhostcheck.so"))
    L.hc_verify.restype = C.c_int
    L.hc_verify.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    return L


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "groth16_update_note.json")) as f:
        return json.load(f)


def vk_from_serialized(b: bytes):
    """(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc) decoded by the PYTHON oracle."""
    n = int.from_bytes(b[336:344], "little")
    assert len(b) == 344 + 48 * n
    return (bls.g1_decompress(b[:48]), bls.g2_decompress(b[48:144]), bls.g2_decompress(b[144:240]),
            bls.g2_decompress(b[240:336]), [bls.g1_decompress(b[344 + 48 * i:392 + 48 * i]) for i in range(n)])


def vk_raw(vk) -> np.ndarray:
    """b200zk_groth16_setup's vk_out layout."""
    return np.frombuffer(bls.g1_to_ffi(vk[0]) + bls.g2_to_ffi(vk[1]) + bls.g2_to_ffi(vk[2]) + bls.g2_to_ffi(vk[3]) +
                         b"".join(bls.g1_to_ffi(p) for p in vk[4]), dtype=np.uint8).copy()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run(hc, raw, n_inputs, proof: bytes, xs, subgroup=1):
    pr = np.frombuffer(proof, dtype=np.uint8).copy()
    xb = np.frombuffer(b"".join(bls.fr_to_mont_bytes(x) for x in xs), dtype=np.uint8).copy()
    return hc.hc_verify(_p(raw), n_inputs, _p(pr), _p(xb), subgroup)


@pytest.mark.parametrize("name", ["withdraw", "deposit"])
def test_golden_proof_accepted_and_forgeries_rejected(hc, golden, name):
    g = golden[name]
    vk = vk_from_serialized(bytes.fromhex(g["vk_compressed_hex"]))
    raw = vk_raw(vk)
    proof = bytes.fromhex(g["proof_hex"])
    xs = [int(x, 16) for x in g["public_inputs"]]
    n = g["num_inputs"]
    assert og.verify_with_vk(vk, xs, og.proof_from_bytes(proof))           # the oracle's verdict
    assert run(hc, raw, n, proof, xs) == 0                                  # accepted
    if name == "deposit":
        return
    # a wrong public input: rejected by both
    bad = list(xs)
    bad[0] = (bad[0] + 1) % R
    assert not og.verify_with_vk(vk, bad, og.proof_from_bytes(proof))
    assert run(hc, raw, n, proof, bad) == 1
    # a different valid curve point as C (2C): well-formed, equation fails
    A, B, Cc = og.proof_from_bytes(proof)
    forged = og.proof_to_bytes((A, B, bls.G1.add(Cc, Cc)))
    assert run(hc, raw, n, forged, xs) == 1
    # malformed encodings
    no_flag = bytes([proof[0] & 0x7F]) + proof[1:]
    assert run(hc, raw, n, no_flag, xs) == 2                                # compressed bit missing
    x_ge_p = bytes([0x9F]) + b"\xff" * 47 + proof[48:]
    assert run(hc, raw, n, x_ge_p, xs) == 2                                 # x >= p
    # an x with no point on the curve (search from the proof's own x)
    x = int.from_bytes(bytes([proof[0] & 0x1F]) + proof[1:48], "big")
    while bls.fq_sqrt((x * x * x + 4) % bls.P) is not None:
        x += 1
    off_curve = bytes([0x80 | (x >> 376)]) + (x & ((1 << 376) - 1)).to_bytes(47, "big") + proof[48:]
    assert run(hc, raw, n, off_curve, xs) == 3


def test_not_in_subgroup_and_unreduced_input(hc, golden):
    g = golden["withdraw"]
    vk = vk_from_serialized(bytes.fromhex(g["vk_compressed_hex"]))
    raw = vk_raw(vk)
    proof = bytes.fromhex(g["proof_hex"])
    xs = [int(x, 16) for x in g["public_inputs"]]
    # a curve point outside the r-torsion: E(Fq) has cofactor h != 1, so a random x on the curve is (w.h.p.) not in G1
    x = 5
    while True:
        y = bls.fq_sqrt((x * x * x + 4) % bls.P)
        if y is not None and bls.G1.mul((x, y), R) is not None:
            break
        x += 1
    outside = bls.g1_compress((x, y)) + proof[48:]
    assert hc.hc_verify is not None
    assert run(hc, raw, g["num_inputs"], outside, xs, subgroup=1) == 4
    assert run(hc, raw, g["num_inputs"], outside, xs, subgroup=0) == 1     # decodes, then the equation fails
    # a public input that is not a reduced Montgomery word
    pr = np.frombuffer(proof, dtype=np.uint8).copy()
    xb = np.frombuffer(b"".join(bls.fr_to_mont_bytes(v) for v in xs), dtype=np.uint8).copy()
    xb[:32] = 0xFF
    assert hc.hc_verify(_p(raw), g["num_inputs"], _p(pr), _p(xb), 1) == 5
