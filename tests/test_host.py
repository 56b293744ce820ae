"""CPU-side checks of the product's host code (no GPU, no compute through the C ABI):
  * libb200zk.so loads and exports every symbol include/b200zk.h declares; b200zk_init fails
    loudly without a device (there is no CPU fallback);
  * the host-only entry points (R1CS of the update-note relation, Poseidon parameter generation)
    agree with the Python oracle;
  * the portable (host) instantiation of csrc/field.cuh + ec.cuh -- the same limb algorithms and EC
    formulas the kernels compile -- agrees with the Python oracle (tests/host/hostcheck.cpp)."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from oracle.pyref import bls12_381 as bls, poseidon as pos, relations as rel
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, P = bls.R, bls.P


@pytest.fixture(scope="module")
def z(built):
    import zk_apps_b200 as z
    return z


@pytest.fixture(scope="module")
def hc(built):
    L = C.CDLL(os.path.join(ROOT, "tests", "host", "libhostcheck.so"))
    vp = C.c_void_p
    L.hc_field_op.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_size_t]
    L.hc_msm_naive.argtypes = [C.c_int, vp, vp, C.c_size_t, vp]
    L.hc_madd_chain.argtypes = [C.c_int, vp, vp, C.c_size_t, vp]
    L.hc_glv_split.argtypes = [vp, C.c_size_t, vp, vp]
    L.hc_glv_phi.argtypes = [C.c_int, vp, C.c_size_t, vp]
    L.hc_fq_mul_dfma.argtypes = [vp, vp, vp, C.c_size_t]
    L.hc_batch_add_affine.argtypes = [C.c_int, vp, vp, C.c_size_t, C.c_int, vp]
    L.hc_bucket_sums_batch_affine.argtypes = [C.c_int, vp, C.c_size_t, vp, C.c_uint32, C.c_int, vp]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------- the C ABI
def test_abi_exports_every_declared_symbol(z):
    hdr = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40
    lib = C.CDLL(z.lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/b200zk.h but not exported: %s" % missing


def test_python_binding_covers_the_header(z):
    hdr = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", hdr))
    L = z.lib()
    unbound = [n for n in names if getattr(L, n).argtypes is None]
    assert not unbound, "no ctypes signature for: %s" % unbound


def test_no_device_means_error_not_fallback(z):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert z.lib().b200zk_init(0, C.byref(h)) == -5          # B200ZK_ERR_NO_DEVICE
    with pytest.raises(z.B200zkError):
        z.Context(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under zk-apps_b200/ may import or link it."""
    pkg = os.path.join(ROOT, "zk-apps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f == "build.py":          # build() also compiles the checker; building it is not using it
                continue
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and '#include "../../oracle' not in src, f


# ----------------------------------------------------------------------------- host-only entry points
@pytest.mark.parametrize("kind", [rel.DEPOSIT, rel.WITHDRAW])
def test_update_note_r1cs_matches_oracle(z, kind):
    """b200zk_update_note_r1cs (host_r1cs.hpp) vs the oracle's synthesis of update_note.rs:106-149."""
    cs = rel.synthesize_update_note(rel.make_witness(4, kind))
    r = z.UpdateNoteRelation(kind, rel.TREE_HEIGHT)
    assert (r.num_constraints, r.num_inputs, r.num_variables) == (cs.num_constraints, cs.num_inputs, cs.num_variables)
    for which, M in enumerate(cs.matrices()):
        rp, cols, vals = r.matrix(which)
        assert int(rp[-1]) == sum(len(row) for row in M) == r.nnz[which]
        vals = bytes(vals)
        k = 0
        for i, row in enumerate(M):
            assert int(rp[i]) == k
            got = {int(cols[k + j]): bls.fr_from_mont_bytes(vals[32 * (k + j):32 * (k + j + 1)]) for j in range(len(row))}
            want = {c: v % R for c, v in (row.items() if isinstance(row, dict) else row)}
            assert got == want, (which, i)
            k += len(row)
    r.free()


def test_update_note_r1cs_other_heights(z):
    for h in (1, 4, 16):
        r = z.UpdateNoteRelation(rel.WITHDRAW, h)
        cs = rel.synthesize_update_note(rel.make_witness(1, rel.WITHDRAW, h), h)
        assert (r.num_constraints, r.num_variables) == (cs.num_constraints, cs.num_variables)
        r.free()


def _diff_r1cs(r, cs):
    assert (r.num_constraints, r.num_inputs, r.num_variables) == (cs.num_constraints, cs.num_inputs, cs.num_variables)
    for which, M in enumerate(cs.matrices()):
        rp, cols, vals = r.matrix(which)
        assert int(rp[-1]) == sum(len(row) for row in M) == r.nnz[which]
        vals = bytes(vals)
        k = 0
        for i, row in enumerate(M):
            assert int(rp[i]) == k
            got = {int(cols[k + j]): bls.fr_from_mont_bytes(vals[32 * (k + j):32 * (k + j + 1)]) for j in range(len(row))}
            want = {c: v % R for c, v in (row.items() if isinstance(row, dict) else row)}
            assert got == want, (which, i)
            k += len(row)


@pytest.mark.parametrize("kind", [rel.DEPOSIT, rel.WITHDRAW])
def test_update_account_r1cs_matches_oracle(z, kind):
    """b200zk_update_account_r1cs: update_account_circuit (update_account.rs:68-95) as a relation of its own, every
    matrix entry against the oracle's synthesis; the oracle's own assignment satisfies it."""
    w = rel.make_account_witness(3, kind)
    cs = rel.synthesize_update_account(w)
    assert cs.is_satisfied() and cs.num_inputs == 6
    r = z.UpdateAccountRelation(kind)
    _diff_r1cs(r, cs)
    assert r.n_inputs_per_proof == 9 == len(rel.account_witness_to_inputs(w))
    r.free()


@pytest.mark.parametrize("height", [1, 4, 20])
def test_update_note_r1cs_matrices_at_other_heights(z, height):
    """TREE_HEIGHT is a const generic in the reference (merkle_proof.rs:11): full matrix diff away from the mock's 10."""
    w = rel.make_witness(2, rel.WITHDRAW, height)
    cs = rel.synthesize_update_note(w, height)
    assert cs.is_satisfied()
    r = z.UpdateNoteRelation(rel.WITHDRAW, height)
    _diff_r1cs(r, cs)
    r.free()


def test_poseidon_constants_match_oracle(z):
    rc, mds = z.poseidon_constants()
    want_rc, want_mds = pos.constants()
    assert util.fr_from_mont_array(rc) == [v for row in want_rc for v in row]
    assert util.fr_from_mont_array(mds) == [v for row in want_mds for v in row]


# ----------------------------------------------------------------------------- portable field / EC path
@pytest.mark.parametrize("field", [0, 1, 2])
def test_host_field_ops(hc, field):
    rnd = random.Random(field)
    mod, size = (R, 32) if field == 0 else (P, 48)
    n = 200
    def enc(vals):
        if field == 0:
            return np.frombuffer(b"".join(bls.fr_to_mont_bytes(v) for v in vals), dtype=np.uint8).copy()
        if field == 1:
            return np.frombuffer(b"".join(bls.fq_to_mont_bytes(v) for v in vals), dtype=np.uint8).copy()
        return np.frombuffer(b"".join(bls.fq_to_mont_bytes(v[0]) + bls.fq_to_mont_bytes(v[1]) for v in vals), dtype=np.uint8).copy()
    if field < 2:
        edge = [2, 3, 4, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 1 << 31, 1 << 32, 1 << 64, 1 << 200, (1 << 254) % mod,
                (1 << 128) - 1, 0xFFFFFFFF, pow(2, -1, mod), pow(3, -1, mod)]            # inversion: binary-GCD corner cases
        A = [rnd.randrange(mod) for _ in range(n)] + [0, 1, mod - 1, mod - 1] + edge
        B = [rnd.randrange(mod) for _ in range(n)] + [mod - 1, mod - 1, mod - 1, 1] + edge[::-1]
        ops = {0: lambda x, y: (x + y) % mod, 1: lambda x, y: (x - y) % mod, 2: lambda x, y: x * y % mod,
               3: lambda x, y: x * x % mod, 4: lambda x, y: pow(x, -1, mod) if x else 0,
               8: lambda x, y: pow(x, -1, mod) if x else 0,                  # op 8 = fp_inv_uniform (groundwork)
               9: lambda x, y: pow(x, -1, mod) if x else 0}                  # op 9 = fp_inv_binary (op 4 = fp_inv = safegcd)
    else:
        A = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(n)] + [(0, 0), (1, 0), (0, 1), (P - 1, P - 1)]
        B = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(n)] + [(P - 1, 1), (P - 1, P - 1), (0, P - 1), (P - 1, P - 1)]
        ops = {0: bls.fq2_add, 1: bls.fq2_sub, 2: bls.fq2_mul, 3: lambda x, y: bls.fq2_sqr(x),
               4: lambda x, y: bls.fq2_inv(x) if not bls.fq2_is_zero(x) else (0, 0)}
    a, b = enc(A), enc(B)
    for op, fn in ops.items():
        out = np.zeros_like(a)
        hc.hc_field_op(field, op, _p(a), _p(b), _p(out), len(A))
        assert bytes(out) == bytes(enc([fn(x, y) for x, y in zip(A, B)])), (field, op)


@pytest.mark.parametrize("field", [0, 1])
def test_host_inversion_is_total_on_unreduced_words(hc, field):
    """fp_inv used to spin forever on the raw word p (u becomes 0 after one subtraction).  Every N-limb
    word is now reduced first: multiples of p give 0, anything else the inverse of its residue."""
    mod, size = (R, 32) if field == 0 else (P, 48)
    Rm = 1 << (8 * size)
    words = [mod, 2 * mod if 2 * mod < Rm else mod, mod + 5, Rm - 1, mod + (mod - 1)]
    raw = np.frombuffer(b"".join(w.to_bytes(size, "little") for w in words), dtype=np.uint8).copy()
    out = np.zeros_like(raw)
    hc.hc_field_op(field, 4, _p(raw), _p(raw), _p(out), len(words))
    # the word w stands for the Montgomery form of (w mod p) * R^-1; its inverse in Montgomery form is (w mod p)^-1 * R^2
    for i, w in enumerate(words):
        got = int.from_bytes(bytes(out[i * size:(i + 1) * size]), "little")
        wm = w % mod
        want = pow(wm, -1, mod) * Rm * Rm % mod if wm else 0
        assert got == want, (field, i)


def test_host_mont_conversion(hc):
    vals = [0, 1, 7, R - 1, 1 << 200]
    canon = np.frombuffer(b"".join(bls.int_to_le(v, 32) for v in vals), dtype=np.uint8).copy()
    out = np.zeros_like(canon)
    hc.hc_field_op(0, 5, _p(canon), None, _p(out), len(vals))
    assert bytes(out) == b"".join(bls.fr_to_mont_bytes(v) for v in vals)
    back = np.zeros_like(canon)
    hc.hc_field_op(0, 6, _p(out), None, _p(back), len(vals))
    assert bytes(back) == bytes(canon)


@pytest.mark.parametrize("group", [1, 2])
def test_host_madd_special_cases(hc, group):
    """XYZZ += affine through every branch: first add, doubling (P + P), cancellation (P - P), infinity operand."""
    cv, enc, dec = (bls.G1, util.g1_array, util.g1_list) if group == 1 else (bls.G2, util.g2_array, util.g2_list)
    g = cv.gen
    p3, p5 = cv.mul(g, 3), cv.mul(g, 5)
    cases = [([p3, p3], [0, 0], 6), ([p3, p3], [0, 1], 0), ([p3, p5, p3, None, p5], [0, 0, 1, 0, 0], 10),
             ([None, p5], [0, 1], -5), ([p3, p3, p3, p3], [0, 0, 0, 0], 12), ([p5, p3, p3], [1, 0, 0], 1)]
    for pts, neg, k in cases:
        out = np.zeros(96 * group, dtype=np.uint8)
        hc.hc_madd_chain(group, _p(enc(pts)), _p(np.array(neg, dtype=np.uint8)), len(pts), _p(out))
        want = None if k == 0 else cv.mul(g, k % R)
        assert dec(out)[0] == want, (pts, neg, k)


@pytest.mark.parametrize("group", [1, 2])
def test_host_scalar_mul_and_sum(hc, group):
    cv, enc, dec = (bls.G1, util.g1_array, util.g1_list) if group == 1 else (bls.G2, util.g2_array, util.g2_list)
    ks = [1, 2, 3, 4]
    bases = [cv.mul(cv.gen, k) for k in ks]
    for scalars, want in (([1, 2, 3, 4], 30), ([R - 1, 1, 0, 2], 9)):
        out = np.zeros(96 * group, dtype=np.uint8)
        hc.hc_msm_naive(group, _p(enc(bases)), _p(util.scalars_array(scalars)), 4, _p(out))
        assert dec(out)[0] == cv.mul(cv.gen, want)


def test_rust_sys_bindings_match_header():
    """rust/b200zk-sys/src/lib.rs is generated from include/b200zk.h (tools/gen_rust_sys.py): it must not be
    stale, and every `sys::` item the arkworks-shaped shim uses must exist in it.  (Source only: no Rust
    toolchain in this image -- SURVEY.md section 8f rank 4.)"""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    sys_rs = open(os.path.join(root, "rust", "b200zk-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (b200zk_\w+)\(", sys_rs)) | set(re.findall(r"pub const (B200ZK_\w+):", sys_rs)) \
        | set(re.findall(r"pub struct (b200zk_\w+) ", sys_rs))
    header = open(os.path.join(root, "include", "b200zk.h")).read()
    exported = set(re.findall(r"^(?:const\s+)?\w+\s*\*?\s*(b200zk_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S), flags=re.M))
    assert exported and exported <= declared
    shim = open(os.path.join(root, "rust", "b200zk", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(\w+)", shim))
    assert used and used <= declared, sorted(used - declared)


def test_glv_split(hc):
    """glv.cuh: k = k1 + k2 * lambda as INTEGERS with k1 < lambda, k2 = floor(k / lambda), for every 256-bit k (also
    unreduced ones), and phi(P) = (beta x, y) = lambda * P on G1 (beta^2 on G2).  lambda = z^2 - 1, lambda^2 + lambda + 1 = r."""
    zz = -0xd201000000010000
    lam = zz * zz - 1
    assert lam * lam + lam + 1 == R
    rnd = random.Random(5)
    ks = [0, 1, 2, lam - 1, lam, lam + 1, 2 * lam - 1, 2 * lam, R - 1, R - 2, R, (R - 1) // lam * lam, (R - 1) // lam * lam - 1,
          (1 << 256) - 1, (1 << 255), (1 << 128) - 1, 1 << 128, lam * lam, lam * lam - 1, lam * (lam + 1)]
    ks += [rnd.randrange(R) for _ in range(3000)] + [rnd.randrange(1 << 256) for _ in range(500)]
    ks += [q * lam + d for q in (1, 7, lam // 3, lam, lam + 1) for d in (0, 1, lam - 1)]
    buf = np.frombuffer(b"".join(bls.int_to_le(k, 32) for k in ks), dtype=np.uint8).copy()
    k1 = np.zeros(len(ks) * 20, dtype=np.uint8)
    k2 = np.zeros(len(ks) * 20, dtype=np.uint8)
    hc.hc_glv_split(_p(buf), len(ks), _p(k1), _p(k2))
    for i, k in enumerate(ks):
        a = int.from_bytes(k1[i * 20:(i + 1) * 20].tobytes(), "little")
        b = int.from_bytes(k2[i * 20:(i + 1) * 20].tobytes(), "little")
        assert (a, b) == (k % lam, k // lam), hex(k)
        assert a < (1 << 128) and b < (1 << 129)
    for group, cv, enc_fn, dec_fn in ((1, bls.G1, util.g1_array, util.g1_list), (2, bls.G2, util.g2_array, util.g2_list)):
        pts = [cv.mul(cv.gen, s) for s in (1, 2, 12345, R - 1)]
        enc = enc_fn(pts)
        out = np.zeros_like(enc)
        hc.hc_glv_phi(group, _p(enc), len(pts), _p(out))
        assert dec_fn(out) == [cv.mul(p, lam) for p in pts]


def test_fq_mul_dfma(hc):
    """field_dfma.cuh (experiment): the Fq Montgomery product computed with double-precision FMAs (48-bit limbs, hi/lo
    halves from two round-toward-zero FMAs) gives the same bytes as the integer product and as the oracle."""
    rnd = random.Random(9)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 48) - 1, 1 << 48, (1 << 96) - 1, (1 << 380), (1 << 381) - 1 - ((1 << 381) - 1) // P * 0,
            sum(((1 << 48) - 1) << (48 * k) for k in range(8)) % P, sum(1 << (48 * k) for k in range(8)), P // 2, P // 3]
    A = [rnd.randrange(P) for _ in range(5000)] + edge + edge
    B = [rnd.randrange(P) for _ in range(5000)] + edge + edge[::-1]
    enc = lambda vals: np.frombuffer(b"".join(bls.fq_to_mont_bytes(v % P) for v in vals), dtype=np.uint8).copy()
    a, b = enc(A), enc(B)
    out, ref = np.zeros_like(a), np.zeros_like(a)
    hc.hc_fq_mul_dfma(_p(a), _p(b), _p(out), len(A))
    hc.hc_field_op(1, 2, _p(a), _p(b), _p(ref), len(A))
    assert bytes(out) == bytes(ref)
    assert bytes(out) == bytes(enc([x * y % P for x, y in zip(A, B)]))
    # Montgomery WORDS (not values) with every 48-bit limb saturated: the raw inputs the kernel would see
    raw = np.frombuffer(b"".join(bls.int_to_le(v, 48) for v in (P - 1, P - 2, (1 << 380) + 12345, ((1 << 381) - 1) % P)),
                        dtype=np.uint8).copy()
    out2, ref2 = np.zeros_like(raw), np.zeros_like(raw)
    hc.hc_fq_mul_dfma(_p(raw), _p(raw[::1].copy()), _p(out2), 4)
    hc.hc_field_op(1, 2, _p(raw), _p(raw), _p(ref2), 4)
    assert bytes(out2) == bytes(ref2)


@pytest.mark.parametrize("group", [1, 2])
def test_batch_affine_addition(hc, group):
    """ec_batch_affine.cuh (groundwork): chunks of affine additions sharing one inversion, with every special case inside
    a chunk (infinite operands, P + P, P + (-P)), and the round-by-round bucket sums built on it, against the oracle."""
    cv, enc, dec = (bls.G1, util.g1_array, util.g1_list) if group == 1 else (bls.G2, util.g2_array, util.g2_list)
    rnd = random.Random(40 + group)
    pts = [cv.mul(cv.gen, rnd.randrange(1, 1 << 40)) for _ in range(24)]
    neg = lambda p: cv.neg(p) if p is not None else None
    P = pts[:12] + [None, pts[0], None, pts[3], pts[4], pts[5]]
    Q = pts[12:] + [pts[1], None, None, pts[3], neg(pts[4]), pts[6]]
    want = [cv.add(a, b) for a, b in zip(P, Q)]
    for chunk in (1, 2, 5, 32):
        out = np.zeros(len(P) * (96 if group == 1 else 192), dtype=np.uint8)
        hc.hc_batch_add_affine(group, _p(enc(P)), _p(enc(Q)), len(P), chunk, _p(out))
        assert dec(out) == want, chunk
    # bucket sums: 7 buckets (one empty, one singleton), duplicates and opposite points inside a bucket
    members = {0: pts[:9], 1: [pts[9]], 3: [pts[10], pts[10], pts[11]], 4: [pts[12], neg(pts[12])],
               5: [pts[13], neg(pts[13]), pts[14]], 6: pts[15:] + [None, pts[15]]}
    flat, ids = [], []
    for b in sorted(members):
        flat += members[b]
        ids += [b] * len(members[b])
    want = []
    for b in range(7):
        acc = None
        for p in members.get(b, []):
            acc = cv.add(acc, p)
        want.append(acc)
    for chunk in (1, 3, 16):
        out = np.zeros(7 * (96 if group == 1 else 192), dtype=np.uint8)
        hc.hc_bucket_sums_batch_affine(group, _p(enc(flat)), len(flat), _p(np.array(ids, dtype=np.uint32)), 7, chunk, _p(out))
        assert dec(out) == want, chunk


def test_tools_and_bench_parse():
    """Every measurement script (bench.py, tools/*.py) at least compiles: they only run on a GPU box."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(root, "tools", "*.py")))
    assert len(files) >= 10
    for f in files:
        compile(open(f).read(), f, "exec")
