"""K2/K3 parity (GPU): NTT / iNTT / coset forms vs the Python oracle's restatement of
ark-poly Radix2EvaluationDomain, bit-exact, through b200zk_ntt_fr."""
import numpy as np
import pytest

import zk_apps_b200 as z
from oracle.pyref import bls12_381 as bls
from oracle.pyref.algos import Domain
from tests import util

pytestmark = pytest.mark.gpu
R = bls.R
OFFSET = bls.fr_to_mont_bytes(7)


def test_kat_n4(ctx):
    """NTT_4([1,2,3,4]) and the coset form with offset 7."""
    d = z.Radix2EvaluationDomain.new(ctx, 4)
    out = util.fr_from_mont_array(d.fft(util.fr_mont_array([1, 2, 3, 4])))
    assert out == [0xa, 0x73eda753299d7d4718963e6b1d9bce637bb7a3fe13f85bfefffdfffeffffffff,
                   0x73eda753299d7d483339d80809a1d80553bda402fffe5bfefffffffeffffffff,
                   0x11aa3999cec0609a1d8060004ec0600000001fffffffffffe]
    out = util.fr_from_mont_array(d.get_coset(OFFSET).fft(util.fr_mont_array([1, 2, 3, 4])))
    assert out == [0x5fe, 0x73eda753299d7a5a8b4d68d2059e4bc15bd396f4fc145bfefab1fffeffffff6f,
                   0x73eda753299d7d483339d80809a1d80553bda402fffe5bfefffffffefffffb2b,
                   0x2eda7ec6f3604038c43f7ea0d0e03ea0000054dffffffffff6e]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 10, 11, 12, 13, 14])
def test_forward_inverse_coset_vs_oracle(ctx, log_n):
    n = 1 << log_n
    x = util.rand_fr(1000 + log_n, n)
    dom = Domain(n)
    gd = z.Radix2EvaluationDomain(ctx, log_n)
    buf = util.fr_mont_array(x)
    assert util.fr_from_mont_array(gd.fft(buf)) == dom.fft(x)
    assert util.fr_from_mont_array(gd.ifft(buf)) == dom.ifft(x)
    cd, gcd = dom.get_coset(7), gd.get_coset(OFFSET)
    assert util.fr_from_mont_array(gcd.fft(buf)) == cd.fft(x)
    assert util.fr_from_mont_array(gcd.ifft(buf)) == cd.ifft(x)


def test_zero_padding_like_arkworks(ctx):
    """fft_in_place on fewer coefficients than the domain size zero-pads (ark-poly behaviour)."""
    x = util.rand_fr(5, 5)
    d = z.Radix2EvaluationDomain.new(ctx, 5)
    assert d.size == 8
    assert util.fr_from_mont_array(d.fft(util.fr_mont_array(x))) == Domain(5).fft(x)


def test_domain_too_large(ctx):
    assert z.Radix2EvaluationDomain.new(ctx, (1 << 32) + 1) is None
    buf = np.zeros(32, dtype=np.uint8)
    rc = z.lib().b200zk_ntt_fr(ctx.handle, buf.ctypes.data, 33, 0, None, 1)
    assert rc == -3


@pytest.mark.parametrize("log_n,batch", [(6, 7), (11, 3), (13, 5)])
def test_batched(ctx, log_n, batch):
    n = 1 << log_n
    xs = [util.rand_fr(50 + b, n) for b in range(batch)]
    buf = util.fr_mont_array([v for x in xs for v in x])
    gd = z.Radix2EvaluationDomain(ctx, log_n).get_coset(OFFSET)
    out = util.fr_from_mont_array(gd.fft(buf, batch=batch))
    cd = Domain(n, 7)
    for b in range(batch):
        assert out[b * n:(b + 1) * n] == cd.fft(xs[b])
    back = gd.ifft(gd.fft(buf, batch=batch), batch=batch)
    assert bytes(back) == bytes(buf)


@pytest.mark.parametrize("log_n", [16, 20, 22, 23])
def test_large_roundtrip_and_spot_values(ctx, log_n):
    """Full-size property checks: iNTT(NTT(x)) == x bit-exactly, coset round trip, and a few output
    coefficients recomputed directly from the definition by the oracle."""
    n = 1 << log_n
    buf = util.rand_fr_bytes_fast(log_n, n)
    gd = z.Radix2EvaluationDomain(ctx, log_n)
    fwd = gd.fft(buf)
    assert bytes(gd.ifft(fwd)) == bytes(buf)
    gcd = gd.get_coset(OFFSET)
    assert bytes(gcd.ifft(gcd.fft(buf))) == bytes(buf)
    if log_n <= 20:
        rinv = pow(bls.FR_MONT_R, -1, R)
        x = [v * rinv % R for v in util.le_ints(buf)]
        w = Domain(n).group_gen
        for k in (0, 1, n // 2 + 3, n - 1):
            wk = pow(w, k, R)
            acc, t = 0, 1
            for v in x:
                acc += v * t
                t = t * wk % R
            assert bls.fr_from_mont_bytes(bytes(fwd[32 * k:32 * k + 32])) == acc % R


def test_linearity_full_size(ctx):
    log_n = 18
    n = 1 << log_n
    a, b = util.rand_fr_bytes_fast(1, n), util.rand_fr_bytes_fast(2, n)
    s = ctx.field_op(0, 0, a, b)
    gd = z.Radix2EvaluationDomain(ctx, log_n)
    assert bytes(gd.fft(s)) == bytes(ctx.field_op(0, 0, gd.fft(a), gd.fft(b)))
