#!/usr/bin/env python3
"""Generates the committed golden fixtures in tests/golden/ from the pure-Python big-int oracle
(oracle/pyref).  Run from the repo root:  python tests/golden/make_golden.py

The reference tree has no vectors for this path ("parity unpinned"), so these
fixtures pin (1) mathematical facts about BLS12-381 (re-derived here) and
(2) the conventions this repo chose where the reference is silent (Poseidon parameter generation,
R1CS layout of the update-note relation, proof bytes for fixed toxic waste and r/s).  Group (2) is a
regression pin, not a reference output; every proof in it also satisfies the pairing equation."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyref import bls12_381 as bls, groth16 as og, poseidon as pos, relations as rel  # noqa: E402
from oracle.pyref.algos import Domain, SplitMix64  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R, P = bls.R, bls.P


def hx(v, n):
    return "0x%0*x" % (2 * n, v)


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.write("\n")


def field():
    w32 = bls.FR_ROOT_2_32
    return {
        "p": hx(P, 48), "r": hx(R, 32),
        "fq_R": hx(pow(2, 384, P), 48), "fq_R2": hx(pow(2, 768, P), 48), "fq_inv32": hx((-pow(P, -1, 1 << 32)) % (1 << 32), 4),
        "fq_inv64": hx((-pow(P, -1, 1 << 64)) % (1 << 64), 8),
        "fr_R": hx(pow(2, 256, R), 32), "fr_R2": hx(pow(2, 512, R), 32), "fr_inv32": hx((-pow(R, -1, 1 << 32)) % (1 << 32), 4),
        "fr_inv64": hx((-pow(R, -1, 1 << 64)) % (1 << 64), 8),
        "fr_root_2_32": str(w32), "omega_4": hx(pow(w32, 1 << 30, R), 32), "omega_8": hx(pow(w32, 1 << 29, R), 32),
        "fr_mont_bytes_of_7": bls.fr_to_mont_bytes(7).hex(), "fq_mont_bytes_of_4": bls.fq_to_mont_bytes(4).hex(),
    }


def curve():
    g1, g2 = bls.G1, bls.G2
    out = {"g1_gen_compressed": bls.g1_compress(g1.gen).hex(), "g2_gen_compressed": bls.g2_compress(g2.gen).hex()}
    for k in (2, 9, 30):
        p = g1.mul(g1.gen, k)
        out["g1_%dG" % k] = [hx(p[0], 48), hx(p[1], 48)]
        out["g1_%dG_compressed" % k] = bls.g1_compress(p).hex()
        q = g2.mul(g2.gen, k)
        out["g2_%dG_compressed" % k] = bls.g2_compress(q).hex()
        out["g2_%dG_ffi_sha256" % k] = hashlib.sha256(bls.g2_to_ffi(q)).hexdigest()
    out["g1_ffi_of_gen"] = bls.g1_to_ffi(g1.gen).hex()
    return out


def msm():
    out = {}
    for name, cv, to_ffi in (("g1", bls.G1, bls.g1_to_ffi), ("g2", bls.G2, bls.g2_to_ffi)):
        rng = SplitMix64(0xB2000001)
        n = 24
        ks = [rng.fr() % 5000 + 1 for _ in range(n)]
        ss = [rng.fr() for _ in range(n)]
        ss[3], ss[7], ss[11] = 0, 1, R - 1
        bases = [cv.mul(cv.gen, k) for k in ks]
        want = cv.mul(cv.gen, sum(a * b for a, b in zip(ks, ss)) % R)
        assert want == cv.msm_naive(bases, ss)
        out[name] = {"base_multiples_of_generator": ks, "scalars": [hx(s, 32) for s in ss], "result_ffi": to_ffi(want).hex()}
    return out


def ntt():
    out = {"ntt4_of_1234": [hx(v, 32) for v in Domain(4).fft([1, 2, 3, 4])],
           "coset7_ntt4_of_1234": [hx(v, 32) for v in Domain(4, 7).fft([1, 2, 3, 4])],
           "coset7_vanishing_n4": hx(Domain(4, 7).vanishing_on_coset(), 32)}
    rng = SplitMix64(0xB2000002)
    for log_n in (6, 11):
        n = 1 << log_n
        x = [rng.fr() for _ in range(n)]
        d, c = Domain(n), Domain(n, 7)
        if log_n == 6:
            assert d.fft(x) == d.dft_naive(x)
        enc = lambda v: hashlib.sha256(b"".join(bls.fr_to_mont_bytes(e) for e in v)).hexdigest()
        out["log%d" % log_n] = {"seed": "SplitMix64(0xB2000002), sequential", "input_sha256": enc(x), "fft_sha256": enc(d.fft(x)),
                                "ifft_sha256": enc(d.ifft(x)), "coset7_fft_sha256": enc(c.fft(x)),
                                "coset7_ifft_sha256": enc(c.ifft(x)), "fft_first": hx(d.fft(x)[0], 32), "fft_last": hx(d.fft(x)[-1], 32)}
    return out


def poseidon():
    rc, mds = pos.constants()
    flat = [v for row in rc for v in row]
    return {"params": {"t": 5, "rate": 4, "r_f": 8, "r_p": 56, "alpha": 5},
            "round_constants_sha256": hashlib.sha256(b"".join(bls.fr_to_mont_bytes(v) for v in flat)).hexdigest(),
            "round_constant_first": hx(flat[0], 32), "round_constant_last": hx(flat[-1], 32),
            "mds_sha256": hashlib.sha256(b"".join(bls.fr_to_mont_bytes(v) for row in mds for v in row)).hexdigest(),
            "mds_00": hx(mds[0][0], 32),
            "permute_of_0_1_2_3_4": [hx(v, 32) for v in pos.permute([0, 1, 2, 3, 4])],
            "hash_1_2": hx(pos.hash_fix_len_array([1, 2]), 32), "hash_1_2_3_4": hx(pos.hash_fix_len_array([1, 2, 3, 4]), 32),
            "hash_1": hx(pos.hash_fix_len_array([1]), 32), "hash_1_2_3_4_5": hx(pos.hash_fix_len_array([1, 2, 3, 4, 5]), 32)}


def groth16():
    out = {}
    tox = og.Toxic(11, 22, 33, 44, 55)
    for kind, name in ((rel.DEPOSIT, "deposit"), (rel.WITHDRAW, "withdraw")):
        w = rel.make_witness(1, kind)
        cs = rel.synthesize_update_note(w)
        assert cs.is_satisfied()
        M = cs.matrices()
        sc = og.setup_scalars(M, cs.num_inputs, cs.num_variables, tox)
        proof = og.proof_via_scalars(M, sc, tox, cs.z, 5, 7)
        vk = og.verifying_key_from_toxic(sc, tox)
        assert og.verify_with_vk(vk, w.public_inputs(), proof)
        # ark CanonicalSerialize (compressed) of VerifyingKey [recall]: alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | u64 LE len | gamma_abc_g1
        vk_ser = (bls.g1_compress(vk[0]) + bls.g2_compress(vk[1]) + bls.g2_compress(vk[2]) + bls.g2_compress(vk[3]) +
                  len(vk[4]).to_bytes(8, "little") + b"".join(bls.g1_compress(p) for p in vk[4]))
        rows = rel.witness_to_inputs(w)
        nnz = [sum(len(r) for r in m) for m in M]
        out[name] = {"make_witness_seed": 1, "toxic": [11, 22, 33, 44, 55], "r": 5, "s": 7,
                     "num_constraints": cs.num_constraints, "num_inputs": cs.num_inputs, "num_variables": cs.num_variables,
                     "nnz_abc": nnz, "domain": sc.n,
                     "instance_inputs_mont_hex": b"".join(bls.fr_to_mont_bytes(v) for v in rows).hex(),
                     "public_inputs": [hx(v, 32) for v in w.public_inputs()],
                     "assignment_sha256": hashlib.sha256(b"".join(bls.fr_to_mont_bytes(v) for v in cs.z)).hexdigest(),
                     "proof_hex": og.proof_to_bytes(proof).hex(), "vk_compressed_hex": vk_ser.hex()}
    return out


if __name__ == "__main__":
    dump("field.json", field())
    dump("curve.json", curve())
    dump("msm.json", msm())
    dump("ntt.json", ntt())
    dump("poseidon.json", poseidon())
    dump("groth16_update_note.json", groth16())
    print("golden fixtures written to", HERE)
