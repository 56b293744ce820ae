#!/usr/bin/env python3
"""the batched Groth16 verifier on one GPU.  Proves `batch` withdraw instances with the
library, then times b200zk_groth16_verify_batch (one verdict per proof: 3 Miller loops + 1 final exponentiation
each) and b200zk_groth16_verify_aggregate (one verdict per batch: batch + 3 Miller loops, 1 final exponentiation)
through the C ABI with host buffers, with and without subgroup checks, and checks that a tampered proof is
rejected by both.  Writes gpurun_out/verify.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_apps_b200 as z
from zk_apps_b200.workload import make_update_note_instances

TOXIC = (0x1f2e3d4c5b6a7988, 0x0123456789abcdef, 0x0fedcba987654321, 0x1122334455667788, 0x99aabbccddeeff00)
ctx = z.Context(0)
rel = z.UpdateNoteRelation(z.WITHDRAW, 10)
pk = z.Groth16.generate_parameters_with_toxic_waste(ctx, rel, TOXIC, precompute=True)
vk = pk.verifying_key()
rows = []
for B in (1, 16, 128, 1024):
    inst = make_update_note_instances(ctx, B, 5, z.WITHDRAW, 10)
    rng = np.random.default_rng(B)
    r = rng.integers(0, 256, size=(B, 32), dtype=np.uint8); r[:, 31] &= 0x3F
    s = rng.integers(0, 256, size=(B, 32), dtype=np.uint8); s[:, 31] &= 0x3F
    proofs = np.concatenate([z.Groth16.prove_update_note(pk, inst[i:i + 128].reshape(-1), r[i:i + 128].reshape(-1),
                                                         s[i:i + 128].reshape(-1), min(128, B - i))[0]
                             for i in range(0, B, 128)])
    pub = np.ascontiguousarray(inst[:, [0, 1, 2, 3, 4, 11], :]).reshape(-1)   # amount, token, user, new_note_hash, root, old nullifier
    coeffs = rng.integers(0, 256, size=B * 16, dtype=np.uint8)
    row = {"batch": B}
    for sub in (True, False):
        st = z.Groth16.verify_proofs(vk, proofs, pub, check_subgroup=sub)
        assert not st.any(), st
        assert z.Groth16.verify_proofs_aggregate(vk, proofs, pub, coeffs, check_subgroup=sub)
        for name, fn in (("per_proof", lambda: z.Groth16.verify_proofs(vk, proofs, pub, check_subgroup=sub)),
                         ("aggregate", lambda: z.Groth16.verify_proofs_aggregate(vk, proofs, pub, coeffs, check_subgroup=sub))):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
            key = "%s%s" % (name, "_subgroup" if sub else "")
            row[key + "_ms"] = round(best * 1e3, 3)
            row[key + "_proofs_per_s"] = round(B / best, 1)
    bad = proofs.copy()
    bad[192 * (B - 1) + 100] ^= 1                                   # flip a bit inside B of the last proof
    st = z.Groth16.verify_proofs(vk, bad, pub)
    assert st[B - 1] != 0 and not st[:B - 1].any()
    assert not z.Groth16.verify_proofs_aggregate(vk, bad, pub, coeffs)
    row["tampered_rejected"] = True
    print(json.dumps(row), flush=True)
    rows.append(row)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"rows": rows}, open("gpurun_out/verify.json", "w"), indent=1)
