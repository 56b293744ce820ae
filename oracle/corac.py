"""ORACLE (test infrastructure) -- ctypes loader for oracle/c/liboracle.so (the C++ restatement that
doubles as the reported CPU baseline).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "c", "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c")], check=True)
        L = C.CDLL(path)
        L.orc_threads.restype = C.c_int
        _LIB = L
        from .pyref import poseidon as pos, bls12_381 as bls
        rc, mds = pos.constants()
        rcb = b"".join(bls.fr_to_mont_bytes(v) for row in rc for v in row)
        mdsb = b"".join(bls.fr_to_mont_bytes(v) for row in mds for v in row)
        L.orc_set_poseidon_constants(rcb, mdsb)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ntt(data: np.ndarray, log_n: int, inverse: bool = False, offset: bytes | None = None, batch: int = 1) -> np.ndarray:
    a = np.ascontiguousarray(data).copy()
    lib().orc_ntt(_p(a), C.c_uint32(log_n), C.c_int(1 if inverse else 0), offset, C.c_size_t(batch))
    return a


def msm(group: int, bases: np.ndarray, scalars: np.ndarray) -> bytes:
    pt = 96 if group == 1 else 192
    n = bases.nbytes // pt
    out = np.zeros(pt, dtype=np.uint8)
    fn = lib().orc_msm_g1 if group == 1 else lib().orc_msm_g2
    fn(_p(np.ascontiguousarray(bases)), _p(np.ascontiguousarray(scalars)), C.c_size_t(n), _p(out))
    return out.tobytes()


def fixed_base_mul(group: int, gen: bytes, scalars: np.ndarray) -> np.ndarray:
    pt = 96 if group == 1 else 192
    n = scalars.nbytes // 32
    out = np.zeros(n * pt, dtype=np.uint8)
    lib().orc_fixed_base_mul(C.c_int(group), gen, _p(np.ascontiguousarray(scalars)), C.c_size_t(n), _p(out))
    return out


def poseidon_hash_batch(inputs: np.ndarray, arity: int) -> np.ndarray:
    n = inputs.nbytes // (32 * arity)
    out = np.zeros(n * 32, dtype=np.uint8)
    lib().orc_poseidon_hash_batch(_p(np.ascontiguousarray(inputs)), C.c_size_t(n), C.c_uint32(arity), _p(out))
    return out


def field_mul(field: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    w = 32 if field == 0 else 48
    out = np.zeros_like(a)
    lib().orc_field_mul(C.c_int(field), _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out), C.c_size_t(a.nbytes // w))
    return out


class CsrMatrices:
    """CSR triple (row_ptr u64, cols u32, vals Montgomery bytes) x 3, built from oracle rows or taken
    from the product's b200zk_r1cs_matrix."""
    def __init__(self, mats):
        self.m = mats

    @classmethod
    def from_rows(cls, matrices):
        from .pyref import bls12_381 as bls
        out = []
        for M in matrices:
            rp = np.zeros(len(M) + 1, dtype=np.uint64)
            cols, vals = [], []
            for i, row in enumerate(M):
                rp[i] = len(cols)
                for v, c in row:
                    cols.append(v); vals.append(bls.fr_to_mont_bytes(c))
            rp[len(M)] = len(cols)
            out.append((rp, np.array(cols, dtype=np.uint32), np.frombuffer(b"".join(vals), dtype=np.uint8).copy()))
        return cls(out)


def groth16_prove(mats: CsrMatrices, nc: int, num_inputs: int, num_vars: int, log_n: int, key: dict, z: np.ndarray,
                  r: int, s: int) -> bytes:
    """key: alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query, b_g2_query, l_query, h_query
    (numpy uint8, FFI layout).  -> 384 B affine A | B | C."""
    out = np.zeros(384, dtype=np.uint8)
    args = [C.c_uint64(nc), C.c_uint64(num_inputs), C.c_uint64(num_vars), C.c_uint32(log_n)]
    keep = []
    for rp, cols, vals in mats.m:
        rp = np.ascontiguousarray(rp); cols = np.ascontiguousarray(cols); vals = np.ascontiguousarray(vals)
        keep += [rp, cols, vals]
        args += [_p(rp), _p(cols), _p(vals)]
    for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "a_query", "b_g1_query", "b_g2_query",
                 "l_query", "h_query"):
        a = np.ascontiguousarray(key[name]); keep.append(a); args.append(_p(a))
    zz = np.ascontiguousarray(z); keep.append(zz)
    args += [_p(zz), int(r).to_bytes(32, "little"), int(s).to_bytes(32, "little"), _p(out)]
    lib().orc_groth16_prove(*args)
    return out.tobytes()
