"""K4/K5 parity (GPU): bucketed MSM over G1 and G2 vs the oracle, bit-exact affine bytes,
through b200zk_msm_g1 / b200zk_msm_g2 / b200zk_msm_resident."""
import numpy as np
import pytest

import zk_apps_b200 as z
from oracle.pyref import bls12_381 as bls
from oracle.pyref.algos import msm_pippenger
from tests import util

pytestmark = pytest.mark.gpu
R = bls.R
CURVES = {1: (bls.G1, util.g1_array, util.g1_list, 96), 2: (bls.G2, util.g2_array, util.g2_list, 192)}


def _msm(ctx, group, bases, scalars, flags=None):
    cv, enc, dec, pt = CURVES[group]
    out, inf = z.VariableBaseMSM.msm_bigint(ctx, group, enc(bases), util.scalars_array(scalars), flags)
    res = dec(out)[0]
    assert inf == (res is None)
    return res


@pytest.mark.parametrize("group", [1, 2])
def test_kat(ctx, group):
    """SURVEY.md Appendix A KATs: (1G..4G)x(1,2,3,4) = 30G ; (r-1,1,0,2) = 9G."""
    cv = CURVES[group][0]
    bases = [cv.mul(cv.gen, k) for k in (1, 2, 3, 4)]
    assert _msm(ctx, group, bases, [1, 2, 3, 4]) == cv.mul(cv.gen, 30)
    assert _msm(ctx, group, bases, [R - 1, 1, 0, 2]) == cv.mul(cv.gen, 9)
    if group == 1:
        assert bls.g1_compress(_msm(ctx, 1, bases, [1, 2, 3, 4])).hex().startswith("ad84464b3966ec5b")


@pytest.mark.parametrize("group", [1, 2])
def test_edge_cases(ctx, group):
    cv = CURVES[group][0]
    g = cv.gen
    p5, p7 = cv.mul(g, 5), cv.mul(g, 7)
    assert _msm(ctx, group, [], []) is None                                   # empty
    assert _msm(ctx, group, [p5, p7], [0, 0]) is None                          # all-zero scalars
    assert _msm(ctx, group, [p5, cv.neg(p5)], [3, 3]) is None                  # P and -P cancel
    assert _msm(ctx, group, [p5] * 6, [1] * 6) == cv.mul(g, 30)                # duplicate bases, equal scalars
    assert _msm(ctx, group, [p5, p5, p7], [R - 1, R - 1, R - 1]) == cv.mul(g, (R - 1) * 17 % R)
    assert _msm(ctx, group, [p5, None, p7], [2, 12345, 3]) == cv.mul(g, 31)    # infinity among the bases
    flags = np.array([0, 1, 0], dtype=np.uint8)
    assert _msm(ctx, group, [p5, p7, p7], [2, 12345, 3], flags) == cv.mul(g, 31)   # flagged infinity
    with pytest.raises(z.B200zkError):                                          # ark msm: Err on length mismatch
        z.VariableBaseMSM.msm_bigint(ctx, group, CURVES[group][1]([p5, p7]), util.scalars_array([1]))


@pytest.mark.parametrize("group,n", [(1, 1), (1, 33), (1, 300), (2, 40)])
def test_vs_oracle_pippenger(ctx, group, n):
    """Same inputs into the oracle's arkworks-style bucket method (independent schedule)."""
    cv = CURVES[group][0]
    ks = util.rand_fr(7 + n, n)
    bases = [cv.mul(cv.gen, k % 1000 + 1) for k in ks]
    scalars = util.rand_fr(8 + n, n)
    assert _msm(ctx, group, bases, scalars) == msm_pippenger(cv, bases, scalars)


def _gpu_bases(ctx, group, seed, n):
    ks = util.rand_fr_bytes_fast(seed, n)
    pts = ctx.fixed_base_mul(group, ks)
    return util.le_ints(ks), pts


@pytest.mark.parametrize("group,log_n,precompute", [(1, 10, False), (1, 13, True), (1, 16, False), (1, 18, False), (1, 22, False),
                                                    (2, 10, False), (2, 13, True), (2, 15, False), (2, 20, False)])
def test_sum_identity(ctx, group, log_n, precompute):
    """Exact full-size check: MSM(s, k*G) == (sum s_i k_i mod r) * G  (SURVEY.md section 8d)."""
    cv, enc, dec, pt = CURVES[group]
    n = 1 << log_n
    ks, pts = _gpu_bases(ctx, group, 100 + log_n, n)
    sbuf = util.rand_fr_bytes_fast(200 + log_n, n)
    ss = util.le_ints(sbuf)
    want = cv.mul(cv.gen, sum(a * b for a, b in zip(ks, ss)) % R)
    h = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=precompute)
    out, inf = h.msm(sbuf)
    assert dec(out)[0] == want
    if not precompute:
        out2, _ = z.VariableBaseMSM.msm_bigint(ctx, group, pts, sbuf)
        assert bytes(out2) == bytes(out)
    h.free()


@pytest.mark.parametrize("group", [1, 2])
def test_adversarial_distributions(ctx, group):
    """Skewed scalars: all equal, all r-1, small (0/1-heavy like a Groth16 witness)."""
    cv, enc, dec, pt = CURVES[group]
    n = 1 << 11
    ks, pts = _gpu_bases(ctx, group, 77, n)
    ksum = sum(ks) % R
    h = z.VariableBaseMSM.Bases(ctx, group, pts)
    for s in (1, R - 1, 0x1234567, (1 << 254) + 12345):
        out, _ = h.msm(util.scalars_array([s] * n))
        assert dec(out)[0] == cv.mul(cv.gen, ksum * s % R)
    rng = np.random.default_rng(3)
    bits = [int(b) for b in rng.integers(0, 2, size=n)]
    out, _ = h.msm(util.scalars_array(bits))
    assert dec(out)[0] == cv.mul(cv.gen, sum(k for k, b in zip(ks, bits) if b) % R)
    h.free()


@pytest.mark.parametrize("group,precompute", [(1, False), (1, True), (2, True)])
def test_batched_shared_bases(ctx, group, precompute):
    """A batch of MSMs over the same bases (one per proof) in one pass."""
    cv, enc, dec, pt = CURVES[group]
    n, batch = 1 << 9, 5
    ks, pts = _gpu_bases(ctx, group, 31, n)
    h = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=precompute)
    sbuf = util.rand_fr_bytes_fast(32, n * batch)
    ss = util.le_ints(sbuf)
    out, inf = h.msm(sbuf, n=n, batch=batch)
    got = dec(out)
    for b in range(batch):
        assert got[b] == cv.mul(cv.gen, sum(k * s for k, s in zip(ks, ss[b * n:(b + 1) * n])) % R)
    h.free()


@pytest.mark.parametrize("group,log_n,parts", [(1, 11, 2), (1, 12, 3), (1, 12, 4), (2, 10, 4), (1, 20, 2)])
def test_window_group_parts(ctx, group, log_n, parts):
    """A single plain-bases MSM cut into window groups on several streams (b200zk_set_option "msm_parts";
    0 = the automatic choice, which splits from 2^23 points): same bytes as the unsplit pipeline, and the
    exact sum identity."""
    cv, enc, dec, pt = CURVES[group]
    n = 1 << log_n
    ks, pts = _gpu_bases(ctx, group, 300 + log_n, n)
    sbuf = util.rand_fr_bytes_fast(400 + log_n, n)
    want = cv.mul(cv.gen, sum(a * b for a, b in zip(ks, util.le_ints(sbuf))) % R)
    h = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=False)
    try:
        ctx.set_option("msm_parts", 1)
        one, _ = h.msm(sbuf)
        ctx.set_option("msm_parts", parts)
        out, inf = h.msm(sbuf)
        out_again, _ = h.msm(sbuf)                     # scratch buffers of the part slots are reused
    finally:
        ctx.set_option("msm_parts", 0)
        h.free()
    assert not inf and dec(out)[0] == want
    assert bytes(out) == bytes(one) == bytes(out_again)


@pytest.mark.parametrize("group", [1, 2])
def test_glv_path_equals_full_length_path(ctx, group):
    """b200zk_set_option "msm_glv": plain-bases MSMs split every scalar as k1 + k2*lambda (glv.cuh).  Same bytes as the
    full-length path and as the oracle, on random scalars and on the decomposition's corner cases (0, 1, lambda +- 1,
    multiples of lambda, r - 1), also batched."""
    cv, enc, dec, pt = CURVES[group]
    zz = -0xd201000000010000
    lam = zz * zz - 1
    corner = [0, 1, 2, lam - 1, lam, lam + 1, 2 * lam, lam * lam % R, (R - 1) // lam * lam, R - 1, R - 2, (1 << 128) - 1, 1 << 128,
              1 << 254]
    n = 256
    ks, pts = _gpu_bases(ctx, group, 500 + group, n)
    scal = corner + util.le_ints(util.rand_fr_bytes_fast(600 + group, n - len(corner)))
    sbuf = util.scalars_array(scal)
    want = cv.mul(cv.gen, sum(a * b for a, b in zip(ks, scal)) % R)
    h = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=False)
    try:
        outs = {}
        for glv in (0, 1):
            ctx.set_option("msm_glv", glv)
            outs[glv] = (h.msm(sbuf)[0], h.msm(np.concatenate([sbuf, util.scalars_array(scal[::-1])]), n=n, batch=2)[0],
                         z.VariableBaseMSM.msm_bigint(ctx, group, pts, sbuf)[0])
    finally:
        ctx.set_option("msm_glv", 1)
        h.free()
    assert dec(outs[1][0])[0] == want
    assert bytes(outs[0][0]) == bytes(outs[1][0]) and bytes(outs[0][1]) == bytes(outs[1][1])
    assert bytes(outs[0][2]) == bytes(outs[1][2]) == bytes(outs[1][0])
    # every scalar a corner case at once (all lanes take the same path)
    for s in (lam, lam - 1, R - 1):
        ctx.set_option("msm_glv", 1)
        out, _ = z.VariableBaseMSM.msm_bigint(ctx, group, pts, util.scalars_array([s] * n))
        assert dec(out)[0] == cv.mul(cv.gen, sum(ks) % R * s % R)


@pytest.mark.parametrize("group,n,c", [(1, 300, 5), (1, 77, 8), (2, 120, 4)])
def test_full_digit_table_msm(ctx, group, n, c):
    """precompute level 2: the full digit table (m + 1) 2^(c w) P_i resident in HBM; an MSM is then a plain sum of one
    table entry per non-zero signed digit (no buckets).  Same bytes as the bucket method, batched, with bases at
    infinity, zero / one / r - 1 scalars and a proof whose scalars are all zero."""
    ks = util.rand_fr_bytes_fast(300 + n, n)
    pts = ctx.fixed_base_mul(group, ks).copy()
    pt = 96 if group == 1 else 192
    pts[3 * pt:4 * pt] = 0                                           # a base at infinity
    batch = 4
    ss = util.rand_fr_bytes_fast(400 + n, batch * n).reshape(batch, n, 32).copy()
    ss[0, 0] = 0; ss[0, 1] = np.frombuffer((1).to_bytes(32, "little"), dtype=np.uint8)
    ss[0, 2] = np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint8)
    ss[2] = 0                                                        # an MSM of the batch with nothing to add
    plain = z.VariableBaseMSM.Bases(ctx, group, pts)
    want, want_inf = plain.msm(ss.reshape(-1), n=n, batch=batch)
    ctx.set_option("table_c_g1" if group == 1 else "table_c_g2", c)
    full = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=2)
    ctx.set_option("table_c_g1", 13); ctx.set_option("table_c_g2", 13)
    got, got_inf = full.msm(ss.reshape(-1), n=n, batch=batch)
    assert bytes(got) == bytes(want) and list(got_inf) == list(want_inf) and got_inf[2] == 1
    one, _ = full.msm(ss[1].reshape(-1), n=n)                        # batch of one, host scalars
    assert bytes(one) == bytes(want[pt:2 * pt])
    few, _ = full.msm(ss[3, :10].reshape(-1), n=10)                  # fewer scalars than bases
    ref, _ = plain.msm(ss[3, :10].reshape(-1), n=10)
    assert bytes(few) == bytes(ref)
    win = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=1)      # window multiples: the same with the bucket method
    few1, _ = win.msm(ss[3, :10].reshape(-1), n=10)
    assert bytes(few1) == bytes(ref)
    plain.free(); full.free(); win.free()


@pytest.mark.parametrize("group,n,c,levels,b", [(1, 300, 5, 4, 32), (1, 77, 8, 3, 8), (1, 1000, 6, 1, 200), (2, 120, 4, 4, 32),
                                                (2, 90, 5, 2, 16)])
def test_digit_table_affine_levels(ctx, group, n, c, levels, b):
    """Full digit tables with the batched-affine pairwise levels in front of the running sums (csrc/msm_affine.cuh):
    same bytes as the bucket method over plain bases -- random batch incl. bases at infinity, zero / one / r - 1
    scalars and an all-zero proof (entry lists padded to 2^levels per proof), then the special cases of the affine
    addition on purpose: repeated bases with equal one-digit scalars (every pair of every level is a doubling),
    P / -P pairs (sums to infinity at level 0, infinity + infinity above), and a mix."""
    ks = util.rand_fr_bytes_fast(900 + n, n)
    pts = ctx.fixed_base_mul(group, ks).copy()
    pt = 96 if group == 1 else 192
    pts[3 * pt:4 * pt] = 0
    batch = 5
    ss = util.rand_fr_bytes_fast(950 + n, batch * n).reshape(batch, n, 32).copy()
    ss[0, 0] = 0; ss[0, 1] = np.frombuffer((1).to_bytes(32, "little"), dtype=np.uint8)
    ss[0, 2] = np.frombuffer((R - 1).to_bytes(32, "little"), dtype=np.uint8)
    ss[2] = 0
    ss[4, 1:] = 0                                                    # a proof with a single non-zero scalar
    plain = z.VariableBaseMSM.Bases(ctx, group, pts)
    want, want_inf = plain.msm(ss.reshape(-1), n=n, batch=batch)
    opt = "table_c_g1" if group == 1 else "table_c_g2"
    ctx.set_option(opt, c)
    ctx.set_option("msm_affine_levels", levels); ctx.set_option("msm_affine_min_entries", 0); ctx.set_option("msm_affine_b", b)
    try:
        full = z.VariableBaseMSM.Bases(ctx, group, pts, precompute=2)
        got, got_inf = full.msm(ss.reshape(-1), n=n, batch=batch)
        assert bytes(got) == bytes(want) and list(got_inf) == list(want_inf) and got_inf[2] == 1
        full.free(); plain.free()
        # ---- special cases: m copies of P (and of -P) with the same small scalar
        cv, enc, dec, _ = CURVES[group]
        P = cv.mul(cv.gen, 0xabcdef12345)
        N = cv.neg(P)
        m = 64
        for name, seq, k_total in [("doublings", [P] * m, m), ("opposites", [P, N] * (m // 2), 0),
                                   ("mixed", [P, P, N, N] * (m // 4) + [P] * 3, 3),
                                   ("dbl-then-cancel", [P] * (m // 2) + [N] * (m // 2), 0)]:
            arr = enc(seq)
            h = z.VariableBaseMSM.Bases(ctx, group, arr, precompute=2)
            for s in (1, 3, (1 << (c - 1)) - 1):
                sc = util.scalars_array([s] * len(seq))
                out, inf = h.msm(sc, n=len(seq))
                res = dec(out)[0]
                expect = cv.mul(P, k_total * s % R) if k_total else None
                assert res == expect and bool(inf[0]) == (expect is None), (name, s)
            h.free()
    finally:
        ctx.set_option(opt, 13)
        ctx.set_option("msm_affine_levels", 4); ctx.set_option("msm_affine_min_entries", 1 << 22); ctx.set_option("msm_affine_b", 96)


@pytest.mark.parametrize("group,n,c,glv", [(1, 3000, 4, 1), (1, 2500, 5, 0), (2, 700, 3, 1)])
def test_plain_bases_affine_levels(ctx, group, n, c, glv, monkeypatch):
    """Affine levels in front of the bucket accumulation of a plain-bases MSM (buckets padded to 2^levels entries in the
    sorted list): same bytes as with the levels off, with and without GLV, incl. bases at infinity, repeated bases
    (doublings inside a bucket) and P / -P pairs."""
    monkeypatch.setenv("B200ZK_MSM_C", str(c))          # few, large buckets so that the levels apply at this size
    cv, enc, dec, pt = CURVES[group]
    ks = util.rand_fr_bytes_fast(1300 + n, n)
    pts = ctx.fixed_base_mul(group, ks).copy()
    pts[5 * pt:6 * pt] = 0
    pts[10 * pt:11 * pt] = pts[11 * pt:12 * pt]                    # a repeated base
    neg = dec(pts[20 * pt:21 * pt])[0]
    pts[21 * pt:22 * pt] = enc([cv.neg(neg)])                       # and an opposite pair
    ss = util.rand_fr_bytes_fast(1400 + n, n).reshape(n, 32).copy()
    ss[10] = ss[11]; ss[20] = ss[21]
    h = z.VariableBaseMSM.Bases(ctx, group, pts)
    try:
        ctx.set_option("msm_glv", glv)
        ctx.set_option("msm_affine_levels", 0)
        want, want_inf = h.msm(ss.reshape(-1), n=n)
        ctx.set_option("msm_affine_levels", 4); ctx.set_option("msm_affine_min_entries", 0); ctx.set_option("msm_affine_b", 16)
        got, got_inf = h.msm(ss.reshape(-1), n=n)
        assert bytes(got) == bytes(want) and list(got_inf) == list(want_inf)
        same = util.scalars_array([7] * n)                           # all-equal scalars: one bucket per window takes everything
        ctx.set_option("msm_affine_levels", 0)
        want2, _ = h.msm(same, n=n)
        ctx.set_option("msm_affine_levels", 4)
        got2, _ = h.msm(same, n=n)
        assert bytes(got2) == bytes(want2)
    finally:
        ctx.set_option("msm_glv", 1); ctx.set_option("msm_affine_levels", 4)
        ctx.set_option("msm_affine_min_entries", 1 << 22); ctx.set_option("msm_affine_b", 96)
        h.free()


def test_adversarial_distributions_2p20(ctx):
    """The skewed cases at n = 2^20: all-equal scalars put every point of a window into ONE bucket, 0/1
    scalars put half of them into bucket 1 of window 0.  Results against (sum s_i k_i) * G; the wall time of each case is
    printed next to the uniform one (pytest -s) -- the oversized buckets are summed by runs in parallel and folded by
    msm_fold_big, so they must stay within a small factor of the uniform time."""
    import time
    n = 1 << 20
    ks = util.rand_fr_bytes_fast(2020, n)
    pts = ctx.fixed_base_mul(1, ks)
    kv = ks.reshape(n, 32)
    # sum k_i mod r without Python big-int loops: limb-wise column sums
    cols = [int(kv[:, j].astype(np.uint64).sum()) for j in range(32)]
    ksum = sum(c << (8 * j) for j, c in enumerate(cols)) % R
    h = z.VariableBaseMSM.Bases(ctx, 1, pts)
    uni = util.rand_fr_bytes_fast(2021, n)
    h.msm(uni, n=n)
    t0 = time.perf_counter(); h.msm(uni, n=n); t_uni = time.perf_counter() - t0
    times = {}
    for name, s in (("all_equal_small", 0x1234567), ("all_equal_r_minus_1", R - 1), ("all_equal_254bit", (1 << 254) + 12345)):
        sc = np.tile(np.frombuffer((s % R).to_bytes(32, "little"), dtype=np.uint8), n)
        h.msm(sc, n=n)
        t0 = time.perf_counter(); out, _ = h.msm(sc, n=n); times[name] = time.perf_counter() - t0
        want = ctx.fixed_base_mul(1, np.frombuffer((ksum * s % R).to_bytes(32, "little"), dtype=np.uint8))
        assert bytes(out) == bytes(want), name
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2, size=n, dtype=np.uint8)
    sc = np.zeros((n, 32), dtype=np.uint8); sc[:, 0] = bits
    h.msm(sc.reshape(-1), n=n)
    t0 = time.perf_counter(); out, _ = h.msm(sc.reshape(-1), n=n); times["zero_one"] = time.perf_counter() - t0
    sel = kv[bits == 1]
    cols = [int(sel[:, j].astype(np.uint64).sum()) for j in range(32)]
    want = ctx.fixed_base_mul(1, np.frombuffer((sum(c << (8 * j) for j, c in enumerate(cols)) % R).to_bytes(32, "little"), dtype=np.uint8))
    assert bytes(out) == bytes(want)
    print("\nG1 MSM 2^20: uniform %.2f ms; " % (t_uni * 1e3) + ", ".join("%s %.2f ms" % (k, v * 1e3) for k, v in times.items()))
    h.free()
