"""K1 parity (GPU): Fr / Fq / Fq2 Montgomery kernels vs the Python big-int oracle, bit-exact.
Through the C ABI (b200zk_dbg_field_op)."""
import random

import numpy as np
import pytest

from oracle.pyref import bls12_381 as bls
from tests import util

pytestmark = pytest.mark.gpu
R, P = bls.R, bls.P
ADD, SUB, MUL, SQR, INV, TO_MONT, FROM_MONT = range(7)


def _enc(field, vals):
    if field == 0:
        return np.frombuffer(b"".join(bls.fr_to_mont_bytes(v) for v in vals), dtype=np.uint8)
    if field == 1:
        return np.frombuffer(b"".join(bls.fq_to_mont_bytes(v) for v in vals), dtype=np.uint8)
    return np.frombuffer(b"".join(bls.fq_to_mont_bytes(v[0]) + bls.fq_to_mont_bytes(v[1]) for v in vals), dtype=np.uint8)


def _dec(field, buf):
    b = bytes(buf)
    if field == 0:
        return [bls.fr_from_mont_bytes(b[i:i + 32]) for i in range(0, len(b), 32)]
    if field == 1:
        return [bls.fq_from_mont_bytes(b[i:i + 48]) for i in range(0, len(b), 48)]
    return [(bls.fq_from_mont_bytes(b[i:i + 48]), bls.fq_from_mont_bytes(b[i + 48:i + 96])) for i in range(0, len(b), 96)]


@pytest.mark.parametrize("field,mod", [(0, R), (1, P)])
def test_prime_field_ops(ctx, field, mod):
    rnd = random.Random(field + 11)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, pow(2, 32 * (8 if field == 0 else 12), mod)]
    A = [rnd.randrange(mod) for _ in range(3000)] + edge + edge
    B = [rnd.randrange(mod) for _ in range(3000)] + edge + edge[::-1]
    a, b = _enc(field, A), _enc(field, B)
    assert _dec(field, ctx.field_op(field, ADD, a, b)) == [(x + y) % mod for x, y in zip(A, B)]
    assert _dec(field, ctx.field_op(field, SUB, a, b)) == [(x - y) % mod for x, y in zip(A, B)]
    assert _dec(field, ctx.field_op(field, MUL, a, b)) == [(x * y) % mod for x, y in zip(A, B)]
    assert _dec(field, ctx.field_op(field, SQR, a)) == [(x * x) % mod for x in A]
    small = A[:64] + edge
    assert _dec(field, ctx.field_op(field, INV, _enc(field, small))) == [pow(x, mod - 2, mod) for x in small]


@pytest.mark.parametrize("field,mod,width", [(0, R, 32), (1, P, 48)])
def test_montgomery_conversion(ctx, field, mod, width):
    rnd = random.Random(5)
    A = [rnd.randrange(mod) for _ in range(500)] + [0, 1, mod - 1]
    canon = np.frombuffer(b"".join(v.to_bytes(width, "little") for v in A), dtype=np.uint8)
    mont = ctx.field_op(field, TO_MONT, canon)
    assert bytes(mont) == bytes(_enc(field, A))                      # == arkworks' in-memory words
    assert bytes(ctx.field_op(field, FROM_MONT, mont)) == bytes(canon)


def test_fq2_ops(ctx):
    rnd = random.Random(7)
    edge = [(0, 0), (1, 0), (0, 1), (P - 1, P - 1), (P - 1, 0), (0, P - 1)]
    A = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(1500)] + edge
    B = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(1500)] + edge[::-1]
    a, b = _enc(2, A), _enc(2, B)
    assert _dec(2, ctx.field_op(2, ADD, a, b)) == [bls.fq2_add(x, y) for x, y in zip(A, B)]
    assert _dec(2, ctx.field_op(2, SUB, a, b)) == [bls.fq2_sub(x, y) for x, y in zip(A, B)]
    assert _dec(2, ctx.field_op(2, MUL, a, b)) == [bls.fq2_mul(x, y) for x, y in zip(A, B)]
    assert _dec(2, ctx.field_op(2, SQR, a)) == [bls.fq2_sqr(x) for x in A]
    small = A[:16] + [(1, 0), (0, 1)]
    assert _dec(2, ctx.field_op(2, INV, _enc(2, small))) == [bls.fq2_inv(x) for x in small]


def test_fixed_base_mul(ctx):
    """k*G on the GPU (used to make synthetic bases) vs the oracle's double-and-add."""
    ks = [0, 1, 2, 3, R - 1, R - 2, 0xdeadbeef, (1 << 255) % R] + util.rand_fr(99, 8)
    sc = util.scalars_array(ks)
    g1 = util.g1_list(ctx.fixed_base_mul(1, sc))
    assert g1 == [bls.G1.mul(bls.G1_GEN, k) for k in ks]
    g2 = util.g2_list(ctx.fixed_base_mul(2, sc[:32 * 10]))
    assert g2 == [bls.G2.mul(bls.G2_GEN, k) for k in ks[:10]]


def test_fq_mul_on_the_fp64_pipe(ctx):
    """Experiment (csrc/field_dfma.cuh): the Fq Montgomery product computed with DFMA on 48-bit limbs is bit-identical
    to the integer-pipe product, on the device (op 7 = B200ZK_OP_MUL_DFMA)."""
    import random
    rnd = random.Random(21)
    P = bls.P
    vals = [rnd.randrange(P) for _ in range(4096)] + [0, 1, P - 1, P - 2, (1 << 380), (1 << 48) - 1, 1 << 48]
    a = np.frombuffer(b"".join(bls.fq_to_mont_bytes(v) for v in vals), dtype=np.uint8).copy()
    b = np.frombuffer(b"".join(bls.fq_to_mont_bytes(v) for v in vals[::-1]), dtype=np.uint8).copy()
    ref = ctx.field_op(1, 2, a, b)
    got = ctx.field_op(1, 7, a, b)
    assert bytes(got) == bytes(ref)
    assert bytes(got) == b"".join(bls.fq_to_mont_bytes(x * y % P) for x, y in zip(vals, vals[::-1]))


@pytest.mark.parametrize("group", [1, 2])
def test_batch_affine_addition_kernel(ctx, group):
    """Groundwork (csrc/ec_batch_affine.cuh): chunks of affine additions sharing one inversion on the device, with the
    special cases inside a chunk, against the oracle."""
    cv, enc, dec = (bls.G1, util.g1_array, util.g1_list) if group == 1 else (bls.G2, util.g2_array, util.g2_list)
    rnd = random.Random(50 + group)
    pts = [cv.mul(cv.gen, rnd.randrange(1, 1 << 40)) for _ in range(40)]
    P = pts[:20] + [None, pts[0], None, pts[3], pts[4], pts[5]]
    Q = pts[20:] + [pts[1], None, None, pts[3], cv.neg(pts[4]), pts[6]]
    want = [cv.add(a, b) for a, b in zip(P, Q)]
    for chunk in (1, 3, 32, 64):
        out, _, _ = ctx.batch_add_affine(group, enc(P), enc(Q), chunk)
        assert dec(out) == want, chunk
