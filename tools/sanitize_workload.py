#!/usr/bin/env python3
"""Small pass over every kernel family for compute-sanitizer (tools/sanitize.sh): 2^10-point G1/G2 MSMs (plain + GLV,
window multiples, full digit table, skewed scalars), NTTs in every mode, the note tree, witness generation and two
proof batches in flight (async submit / wait) for the update-account relation, verification.  Results are checked
against each other so that a silent corruption would also fail the run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_apps_b200 as z
from oracle.pyref import relations as rel
from tests import util

ctx = z.Context(0)
ctx.set_option("msm_affine_min_entries", 0)                   # the batched-affine levels at every size (msm_affine.cuh)
ctx.set_option("msm_affine_b", 8)
n = 1 << 10
for group in (1, 2):
    pts = ctx.fixed_base_mul(group, util.rand_fr_bytes_fast(1, n))
    ss = util.rand_fr_bytes_fast(2, n).copy()
    ss[:32 * 100] = np.tile(ss[:32], 100)                     # 100 equal scalars: one big bucket
    want, _ = z.VariableBaseMSM.msm_bigint(ctx, group, pts, ss)
    ctx.set_option("msm_glv", 0)
    assert z.VariableBaseMSM.msm_bigint(ctx, group, pts, ss)[0] == want
    ctx.set_option("msm_glv", 1)
    os.environ["B200ZK_MSM_C"] = "4"                          # few large buckets: the affine levels of the bucket path apply
    assert z.VariableBaseMSM.msm_bigint(ctx, group, pts, ss)[0] == want
    del os.environ["B200ZK_MSM_C"]
    for level in (1, 2):
        ctx.set_option("table_c_g1", 5); ctx.set_option("table_c_g2", 4)
        h = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=level)
        got, _ = h.msm(np.concatenate([ss, ss]), n=n, batch=2)
        assert bytes(got[:len(want)]) == want and bytes(got[len(want):]) == want, (group, level)
        h.free()
print("msm ok")
for lg in (3, 10, 13):
    x = util.rand_fr_bytes_fast(3 + lg, 1 << lg)
    d = z.Radix2EvaluationDomain(ctx, lg)
    assert bytes(d.ifft(d.fft(x))) == bytes(x)
    c = d.get_coset(z.fr_to_mont(7))
    assert bytes(c.ifft(c.fft(x))) == bytes(x)
print("ntt ok")
t = z.MerkleTree(ctx, 6)
t.add_leaves(util.rand_fr_bytes_fast(9, 20))
t.gen_proofs(list(range(5)))
print("tree ok")
relation = z.UpdateAccountRelation(rel.WITHDRAW)
ctx.set_option("table_c_g1", 4); ctx.set_option("table_c_g2", 3)
pk = z.Groth16.generate_parameters_with_toxic_waste(ctx, relation, (11, 22, 33, 44, 55), precompute=2)
B = 3
ws = [rel.make_account_witness(i, rel.WITHDRAW) for i in range(B)]
inputs = util.fr_mont_array([v for w in ws for v in rel.account_witness_to_inputs(w)])
rs = util.scalars_array([5, 6, 7]); ss = util.scalars_array([8, 9, 10])
sync, status = z.Groth16.prove_update_note(pk, inputs, rs, ss, B)
outs = [np.zeros(B * 192, dtype=np.uint8) for _ in range(2)]
t0 = z.Groth16.prove_submit(pk, inputs, rs, ss, B, outs[0])
t1 = z.Groth16.prove_submit(pk, inputs, rs, ss, B, outs[1])
z.Groth16.prove_wait(pk, t0); z.Groth16.prove_wait(pk, t1)
assert bytes(outs[0]) == bytes(sync) == bytes(outs[1])
vk = z.VerifyingKey(ctx, pk.vk)
pub = util.fr_mont_array([v for w in ws for v in w.public_inputs()])
assert list(z.Groth16.verify_proofs(vk, sync, pub)) == [0] * B
print("prove + verify ok")
