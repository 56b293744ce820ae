"""ctypes binding of include/b200zk.h plus thin arkworks-shaped wrappers.

Names follow the arkworks 0.4 items this backend replaces ([recall]):
  ark_ec::VariableBaseMSM::{msm, msm_bigint}            -> VariableBaseMSM.msm_bigint
  ark_poly::Radix2EvaluationDomain::{new, fft_in_place,
      ifft_in_place, get_coset}                         -> Radix2EvaluationDomain
Buffers are plain bytes / numpy uint8 arrays in the layouts include/b200zk.h documents.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

FR_BYTES, FQ_BYTES, G1_BYTES, G2_BYTES = 32, 48, 96, 192
R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
_FR_R = pow(2, 256, R_MOD)
_FR_RINV = pow(_FR_R, -1, R_MOD)

STATUS = {0: "OK", -1: "BAD_ARG", -2: "BAD_LEN", -3: "DOMAIN_TOO_LARGE", -4: "CUDA", -5: "NO_DEVICE",
          -6: "UNSATISFIED", -7: "NOT_IMPLEMENTED", -8: "MERKLE_LIMIT_EXCEEDED", -9: "MERKLE_PROOF_GEN_FAIL",
          -10: "MERKLE_NON_EXISTING_NODE", -11: "BAD_ENCODING", -12: "NCCL"}


class B200zkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("b200zk %s (%d): %s" % (STATUS.get(code, "?"), code, msg))
        self.code = code


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200zk.so")


_lib = None


def lib() -> C.CDLL:
    """Load libb200zk.so.  No fallback: a missing library is an error, not a reason to go slow."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise B200zkError(-5, "libb200zk.so not built (run `python __graft_entry__.py` / build())")
        L = C.CDLL(p)
        vp, sz, u8p, i32, u32 = C.c_void_p, C.c_size_t, C.c_char_p, C.c_int, C.c_uint32
        sig = {
            "b200zk_init": (i32, [i32, C.POINTER(vp)]),
            "b200zk_destroy": (None, [vp]),
            "b200zk_last_error": (C.c_char_p, [vp]),
            "b200zk_sync": (i32, [vp]),
            "b200zk_set_option": (i32, [vp, C.c_char_p, i32]),
            "b200zk_dev_alloc": (i32, [vp, sz, C.POINTER(vp)]),
            "b200zk_dev_free": (i32, [vp, vp]),
            "b200zk_dev_upload": (i32, [vp, vp, vp, sz]),
            "b200zk_dev_download": (i32, [vp, vp, vp, sz]),
            "b200zk_stream": (vp, [vp]),
            "b200zk_prof_enable": (i32, [vp, i32]),
            "b200zk_prof_reset": (i32, [vp]),
            "b200zk_prof_get": (i32, [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_long)]),
            "b200zk_prof_names": (i32, [vp, C.c_char_p, sz]),
            "b200zk_prof_timeline": (i32, [vp, C.c_char_p, sz]),
            "b200zk_launch_count": (C.c_long, [vp]),
            "b200zk_dbg_field_op": (i32, [vp, i32, i32, vp, vp, vp, sz]),
            "b200zk_dbg_int_peak": (i32, [vp, i32, C.POINTER(C.c_double)]),
            "b200zk_dbg_batch_add_affine": (i32, [vp, i32, vp, vp, C.c_size_t, i32, vp, C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double)]),
            "b200zk_fixed_base_mul": (i32, [vp, i32, vp, sz, vp]),
            "b200zk_fixed_base_mul_device": (i32, [vp, i32, vp, sz, vp]),
            "b200zk_ntt_fr": (i32, [vp, vp, u32, i32, vp, sz]),
            "b200zk_ntt_fr_device": (i32, [vp, vp, u32, i32, vp, sz]),
            "b200zk_ntt_twiddle_transpose_device": (i32, [vp, vp, vp, u32, C.c_uint64, C.c_uint64, C.c_uint64, i32]),
            "b200zk_copy2d_device": (i32, [vp, vp, sz, vp, sz, sz, sz]),
            "b200zk_msm_g1": (i32, [vp, vp, vp, vp, sz, vp, vp]),
            "b200zk_msm_g2": (i32, [vp, vp, vp, vp, sz, vp, vp]),
            "b200zk_bases_upload": (i32, [vp, i32, vp, vp, sz, i32, C.POINTER(vp)]),
            "b200zk_bases_from_device": (i32, [vp, i32, vp, sz, i32, C.POINTER(vp)]),
            "b200zk_bases_free": (None, [vp, vp]),
            "b200zk_msm_resident": (i32, [vp, vp, vp, i32, sz, sz, vp, vp]),
            "b200zk_msm_resident_device": (i32, [vp, vp, vp, i32, sz, sz, vp]),
            "b200zk_points_sum": (i32, [vp, i32, vp, sz, vp, vp]),
            "b200zk_points_sum_device": (i32, [vp, i32, vp, sz, vp]),
            "b200zk_update_note_r1cs": (i32, [i32, u32, C.POINTER(vp)]),
            "b200zk_update_account_r1cs": (i32, [i32, C.POINTER(vp)]),
            "b200zk_update_account_witness_batch": (i32, [vp, vp, vp, sz, vp, vp, vp]),
            "b200zk_update_account_prove_batch": (i32, [vp, vp, vp, sz, vp, vp, vp, vp]),
            "b200zk_update_note_prove_submit": (i32, [vp, vp, vp, i32, sz, vp, vp, vp, vp, C.POINTER(C.c_uint64)]),
            "b200zk_update_account_prove_submit": (i32, [vp, vp, vp, i32, sz, vp, vp, vp, vp, C.POINTER(C.c_uint64)]),
            "b200zk_prove_wait": (i32, [vp, C.c_uint64]),
            "b200zk_r1cs_free": (None, [vp]),
            "b200zk_r1cs_shape": (i32, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64 * 3)]),
            "b200zk_r1cs_matrix": (i32, [vp, i32, vp, vp, vp]),
            "b200zk_poseidon_constants": (i32, [vp, vp]),
            "b200zk_poseidon_hash_batch": (i32, [vp, vp, sz, u32, vp]),
            "b200zk_update_note_witness_batch": (i32, [vp, vp, vp, sz, vp, vp, vp]),
            "b200zk_pk_upload": (i32, [vp, vp] + [vp] * 10 + [i32, C.POINTER(vp)]),
            "b200zk_groth16_setup": (i32, [vp, vp, vp, i32, C.POINTER(vp), vp]),
            "b200zk_pk_free": (None, [vp, vp]),
            "b200zk_pk_export_query": (i32, [vp, vp, i32, vp, C.POINTER(sz)]),
            "b200zk_groth16_prove_batch": (i32, [vp, vp, vp, i32, sz, vp, vp, vp, vp]),
            "b200zk_update_note_prove_batch": (i32, [vp, vp, vp, sz, vp, vp, vp, vp]),
            "b200zk_update_note_prove_batch_device": (i32, [vp, vp, vp, sz, vp, vp, vp, vp]),
            "b200zk_merkle_new": (i32, [vp, u32, i32, C.POINTER(vp)]),
            "b200zk_merkle_free": (None, [vp, vp]),
            "b200zk_merkle_info": (i32, [vp, C.POINTER(u32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
            "b200zk_merkle_add_leaves": (i32, [vp, vp, vp, i32, sz, C.POINTER(C.c_uint64), vp]),
            "b200zk_merkle_root": (i32, [vp, vp, vp]),
            "b200zk_merkle_node": (i32, [vp, vp, C.c_uint64, vp]),
            "b200zk_merkle_is_historical_root": (i32, [vp, vp, vp, C.POINTER(i32)]),
            "b200zk_merkle_gen_proofs": (i32, [vp, vp, vp, sz, vp, vp]),
            "b200zk_merkle_fill_update_note_inputs_device": (i32, [vp, vp, vp, sz, vp]),
            "b200zk_vk_upload": (i32, [vp, vp, u32, C.POINTER(vp)]),
            "b200zk_vk_free": (None, [vp, vp]),
            "b200zk_vk_num_inputs": (i32, [vp, C.POINTER(u32)]),
            "b200zk_vk_export": (i32, [vp, vp]),
            "b200zk_groth16_verify_batch": (i32, [vp, vp, vp, vp, i32, sz, i32, vp]),
            "b200zk_groth16_verify_aggregate": (i32, [vp, vp, vp, vp, i32, sz, vp, i32, C.POINTER(i32)]),
            "b200zk_points_compress": (i32, [vp, i32, vp, sz, vp]),
            "b200zk_points_decompress": (i32, [vp, i32, vp, sz, i32, vp, vp]),
            "b200zk_vk_serialize": (i32, [vp, vp, vp, C.POINTER(sz)]),
            "b200zk_vk_deserialize": (i32, [vp, vp, sz, i32, C.POINTER(vp)]),
            "b200zk_pk_serialize": (i32, [vp, vp, vp, vp, C.POINTER(sz)]),
            "b200zk_pk_deserialize": (i32, [vp, vp, vp, sz, i32, i32, C.POINTER(vp), C.POINTER(vp)]),
            "b200zk_points_validate": (i32, [vp, i32, vp, sz, i32, vp]),
            "b200zk_init_multi": (i32, [i32, C.POINTER(vp)]),
            "b200zk_comm_unique_id": (i32, [vp]),
            "b200zk_comm_init": (i32, [vp, vp, i32, i32]),
            "b200zk_comm_destroy": (i32, [vp]),
            "b200zk_comm_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
            "b200zk_msm_sharded": (i32, [vp, vp, vp, i32, sz, vp, vp]),
            "b200zk_msm_sharded_device": (i32, [vp, vp, vp, i32, sz, vp]),
            "b200zk_msm_sharded_multi": (i32, [C.POINTER(vp), i32, C.POINTER(vp), C.POINTER(vp), i32, C.POINTER(sz), vp, vp]),
            "b200zk_ntt_sharded_device": (i32, [vp, vp, u32, u32, i32]),
            "b200zk_stat_get": (i32, [vp, C.c_char_p, C.POINTER(C.c_double)]),
            "b200zk_stat_reset": (i32, [vp]),
        }
        for name, (res, args) in sig.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _buf(x):
    """bytes / bytearray / numpy array -> (ctypes pointer, keepalive)"""
    if x is None:
        return None, None
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x)
        return a.ctypes.data_as(C.c_void_p), a
    if isinstance(x, (bytes, bytearray)):
        a = np.frombuffer(x, dtype=np.uint8)
        return a.ctypes.data_as(C.c_void_p), a
    raise TypeError(type(x))


def fr_to_mont(v: int) -> bytes:
    return (v % R_MOD * _FR_R % R_MOD).to_bytes(32, "little")


def fr_from_mont(b: bytes) -> int:
    return int.from_bytes(b, "little") * _FR_RINV % R_MOD


class Context:
    """One GPU, one stream (b200zk_ctx).  Single-threaded, like the C ABI says."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().b200zk_init(device, C.byref(self._h))
        if rc != 0:
            self._h = None
            raise B200zkError(rc, "b200zk_init failed (no CUDA device? there is no CPU fallback)")

    def close(self):
        if getattr(self, "_h", None):
            lib().b200zk_destroy(self._h)
            self._h = None

    __del__ = close

    def check(self, rc: int):
        if rc != 0:
            raise B200zkError(rc, lib().b200zk_last_error(self._h).decode(errors="replace"))

    @property
    def handle(self):
        return self._h

    def sync(self):
        self.check(lib().b200zk_sync(self._h))

    def set_option(self, name: str, value: int):
        self.check(lib().b200zk_set_option(self._h, name.encode(), int(value)))

    # ---- communicator (NCCL behind the C ABI, csrc/comm.cu)
    @staticmethod
    def comm_unique_id() -> bytes:
        """The 128-byte ncclUniqueId rank 0 hands to the other ranks."""
        buf = C.create_string_buffer(128)
        rc = lib().b200zk_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise B200zkError(rc, "b200zk_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, uid: bytes | None, rank: int, world: int):
        """Collective over all ranks of the job."""
        p, keep = _buf(uid) if uid is not None else (None, None)
        self.check(lib().b200zk_comm_init(self._h, p, rank, world))

    def comm_init_from_dist(self, dist=None):
        """Bootstrap through an existing torch.distributed process group (any backend): rank 0 creates the id and
        broadcasts it as an object.  world size 1 needs neither NCCL nor a process group."""
        if dist is None:
            self.comm_init(None, 0, 1)
            return
        box = [self.comm_unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(box[0], dist.get_rank(), dist.get_world_size())

    def comm_info(self):
        r, w = C.c_int(), C.c_int()
        self.check(lib().b200zk_comm_info(self._h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def comm_destroy(self):
        self.check(lib().b200zk_comm_destroy(self._h))

    def msm_sharded(self, bases_local, scalars=None, n: int | None = None, device_ptr: int | None = None):
        """VariableBaseMSM over the communicator: this rank's slice of bases / scalars -> (affine bytes, is_inf),
        the same on every rank."""
        pt = G1_BYTES if bases_local.group == 1 else G2_BYTES
        out = np.zeros(pt, dtype=np.uint8)
        inf = C.c_uint8()
        if device_ptr is not None:
            ps, on_dev = C.c_void_p(device_ptr), 1
        else:
            ps, ks = _buf(scalars)
            on_dev = 0
            if n is None:
                n = ks.nbytes // 32 if ks is not None else 0
        self.check(lib().b200zk_msm_sharded(self._h, bases_local._h, ps, on_dev, n, out.ctypes.data_as(C.c_void_p),
                                            C.byref(inf)))
        return out.tobytes(), bool(inf.value)

    def msm_sharded_device(self, bases_local, d_scalars: int, n: int, d_out: int):
        """Same, result left in device memory, nothing synchronised (one stream of work)."""
        self.check(lib().b200zk_msm_sharded_device(self._h, bases_local._h, C.c_void_p(d_scalars), 1, n, C.c_void_p(d_out)))

    def ntt_sharded_device(self, d_local: int, log_n: int, log_n1: int, inverse: bool = False):
        """Four-step NTT over the communicator, in place on this rank's n / world elements: layout L(n1, n2) in,
        L(n2, n1) out (include/b200zk.h)."""
        self.check(lib().b200zk_ntt_sharded_device(self._h, C.c_void_p(d_local), log_n, log_n1, 1 if inverse else 0))

    def points_validate(self, group: int, points, check_subgroup: bool = True) -> np.ndarray:
        pp, kp = _buf(points)
        n = kp.nbytes // (G1_BYTES if group == 1 else G2_BYTES)
        st = np.zeros(n, dtype=np.int32)
        self.check(lib().b200zk_points_validate(self._h, group, pp, n, 1 if check_subgroup else 0, st.ctypes.data_as(C.c_void_p)))
        return st

    # ---- device memory
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(lib().b200zk_dev_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, dptr: int):
        self.check(lib().b200zk_dev_free(self._h, C.c_void_p(dptr)))

    def upload(self, dptr: int, data):
        p, keep = _buf(data)
        self.check(lib().b200zk_dev_upload(self._h, C.c_void_p(dptr), p, keep.nbytes))

    def download(self, dptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        self.check(lib().b200zk_dev_download(self._h, out.ctypes.data_as(C.c_void_p), C.c_void_p(dptr), nbytes))
        return out

    # ---- profiling
    def prof_enable(self, on: bool = True):
        self.check(lib().b200zk_prof_enable(self._h, 1 if on else 0))

    def prof_reset(self):
        self.check(lib().b200zk_prof_reset(self._h))

    def prof_get(self, name: str):
        ms, n = C.c_double(), C.c_long()
        self.check(lib().b200zk_prof_get(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def prof_names(self):
        buf = C.create_string_buffer(4096)
        self.check(lib().b200zk_prof_names(self._h, buf, 4096))
        return [s for s in buf.value.decode().split(",") if s]

    def prof_timeline(self):
        """[(name, start_ms, end_ms)] of every profiled bracket since the last prof_reset."""
        buf = C.create_string_buffer(1 << 20)
        self.check(lib().b200zk_prof_timeline(self._h, buf, 1 << 20))
        rows = []
        for line in buf.value.decode().splitlines():
            name, a, b = line.rsplit(" ", 2)
            rows.append((name, float(a), float(b)))
        return rows

    def launch_count(self) -> int:
        return lib().b200zk_launch_count(self._h)

    def stat_get(self, name: str) -> float:
        v = C.c_double()
        self.check(lib().b200zk_stat_get(self._h, name.encode(), C.byref(v)))
        return v.value

    def stat_reset(self):
        self.check(lib().b200zk_stat_reset(self._h))

    def stream_ptr(self) -> int:
        return lib().b200zk_stream(self._h)

    # ---- K1 debug
    def field_op(self, field: int, op: int, a, b=None) -> np.ndarray:
        size = (32, 48, 96)[field]
        pa, ka = _buf(a)
        pb, kb = _buf(b)
        n = ka.nbytes // size
        out = np.empty(n * size, dtype=np.uint8)
        self.check(lib().b200zk_dbg_field_op(self._h, field, op, pa, pb, out.ctypes.data_as(C.c_void_p), n))
        return out

    def int_peak(self, kind: int) -> float:
        v = C.c_double()
        self.check(lib().b200zk_dbg_int_peak(self._h, kind, C.byref(v)))
        return v.value

    def batch_add_affine(self, group: int, p, q, chunk: int = 32):
        """Groundwork kernel (csrc/ec_batch_affine.cuh): p[i] + q[i], `chunk` pairs per thread sharing one inversion.
        -> (affine sums, ms of that kernel, ms of the same additions as XYZZ mixed additions)."""
        pp, kp = _buf(p)
        pq, kq = _buf(q)
        n = kp.nbytes // (G1_BYTES if group == 1 else G2_BYTES)
        out = np.empty(kp.nbytes, dtype=np.uint8)
        a, b = C.c_double(), C.c_double()
        self.check(lib().b200zk_dbg_batch_add_affine(self._h, group, pp, pq, n, chunk, out.ctypes.data_as(C.c_void_p),
                                                     C.byref(a), C.byref(b)))
        return out, a.value, b.value

    def fixed_base_mul(self, group: int, scalars) -> np.ndarray:
        ps, ks = _buf(scalars)
        n = ks.nbytes // 32
        out = np.empty(n * (G1_BYTES if group == 1 else G2_BYTES), dtype=np.uint8)
        self.check(lib().b200zk_fixed_base_mul(self._h, group, ps, n, out.ctypes.data_as(C.c_void_p)))
        return out


class VariableBaseMSM:
    """ark_ec::VariableBaseMSM for G1Projective (group=1) / G2Projective (group=2)."""

    @staticmethod
    def msm_bigint(ctx: Context, group: int, bases, scalars, inf_flags=None):
        """-> (affine bytes, is_infinity).  Like ark `msm`: a length mismatch is an error."""
        pt = G1_BYTES if group == 1 else G2_BYTES
        pb, kb = _buf(bases)
        ps, ks = _buf(scalars)
        pf, kf = _buf(inf_flags)
        n = kb.nbytes // pt if kb is not None else 0
        if ks is None or ks.nbytes // 32 != n:
            raise B200zkError(-2, "bases/scalars length mismatch")
        out = np.zeros(pt, dtype=np.uint8)
        inf = C.c_uint8()
        fn = lib().b200zk_msm_g1 if group == 1 else lib().b200zk_msm_g2
        ctx.check(fn(ctx.handle, pb, pf, ps, n, out.ctypes.data_as(C.c_void_p), C.byref(inf)))
        return out.tobytes(), bool(inf.value)

    class Bases:
        """Device-resident bases (a proving-key query); precompute=True stores all window multiples."""

        def __init__(self, ctx: Context, group: int, bases=None, inf_flags=None, precompute: bool = False,
                     device_ptr: int | None = None, n: int | None = None):
            self.ctx, self.group = ctx, group
            self._h = C.c_void_p()
            pt = G1_BYTES if group == 1 else G2_BYTES
            if device_ptr is not None:
                self.n = n
                ctx.check(lib().b200zk_bases_from_device(ctx.handle, group, C.c_void_p(device_ptr), n,
                                                         int(precompute), C.byref(self._h)))
            else:
                pb, kb = _buf(bases)
                pf, kf = _buf(inf_flags)
                self.n = kb.nbytes // pt
                ctx.check(lib().b200zk_bases_upload(ctx.handle, group, pb, pf, self.n, int(precompute),
                                                    C.byref(self._h)))

        def msm(self, scalars=None, n: int | None = None, batch: int = 1, device_ptr: int | None = None):
            pt = G1_BYTES if self.group == 1 else G2_BYTES
            out = np.zeros(batch * pt, dtype=np.uint8)
            inf = np.zeros(batch, dtype=np.uint8)
            if device_ptr is not None:
                ps, on_dev = C.c_void_p(device_ptr), 1
            else:
                ps, ks = _buf(scalars)
                on_dev = 0
                if n is None:
                    n = ks.nbytes // 32 // batch
            self.ctx.check(lib().b200zk_msm_resident(self.ctx.handle, self._h, ps, on_dev, n, batch,
                                                     out.ctypes.data_as(C.c_void_p), inf.ctypes.data_as(C.c_void_p)))
            return out, inf

        def msm_to_device(self, d_out: int, scalars=None, n: int | None = None, batch: int = 1,
                          device_ptr: int | None = None):
            """Same MSM, affine result(s) left at device address `d_out`; asynchronous on the ctx stream."""
            if device_ptr is not None:
                ps, on_dev = C.c_void_p(device_ptr), 1
            else:
                ps, ks = _buf(scalars)
                on_dev = 0
                if n is None:
                    n = ks.nbytes // 32 // batch
            self.ctx.check(lib().b200zk_msm_resident_device(self.ctx.handle, self._h, ps, on_dev, n, batch, C.c_void_p(d_out)))

        def free(self):
            if self._h:
                lib().b200zk_bases_free(self.ctx.handle, self._h)
                self._h = None

        __del__ = free


def points_sum(ctx: Context, group: int, points) -> tuple:
    """Sum of affine points (host buffer) -> (affine bytes, is_infinity)."""
    pt = G1_BYTES if group == 1 else G2_BYTES
    pp, kp = _buf(points)
    out = np.zeros(pt, dtype=np.uint8)
    inf = C.c_uint8()
    ctx.check(lib().b200zk_points_sum(ctx.handle, group, pp, kp.nbytes // pt, out.ctypes.data_as(C.c_void_p), C.byref(inf)))
    return out.tobytes(), bool(inf.value)


class Radix2EvaluationDomain:
    """ark_poly::Radix2EvaluationDomain<Fr>.  `new` returns None when log2(size) > 32, as arkworks does."""

    def __init__(self, ctx: Context, log_size: int, offset: bytes | None = None):
        self.ctx, self.log_size, self.size, self.offset = ctx, log_size, 1 << log_size, offset

    @classmethod
    def new(cls, ctx: Context, num_coeffs: int):
        log = max(0, (num_coeffs - 1).bit_length())
        if log > 32:
            return None
        return cls(ctx, log)

    def get_coset(self, offset_mont: bytes) -> "Radix2EvaluationDomain":
        return Radix2EvaluationDomain(self.ctx, self.log_size, bytes(offset_mont))

    def _run(self, data, inverse: bool, batch: int):
        a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data)
        a = a.copy()
        want = batch * self.size * FR_BYTES
        if a.nbytes < want:                                  # arkworks zero-pads to the domain size
            a = np.concatenate([a, np.zeros(want - a.nbytes, dtype=np.uint8)])
        off, keep = _buf(self.offset)
        self.ctx.check(lib().b200zk_ntt_fr(self.ctx.handle, a.ctypes.data_as(C.c_void_p), self.log_size,
                                           1 if inverse else 0, off, batch))
        return a

    def fft(self, coeffs, batch: int = 1) -> np.ndarray:
        return self._run(coeffs, False, batch)

    def ifft(self, evals, batch: int = 1) -> np.ndarray:
        return self._run(evals, True, batch)

    def fft_device(self, dptr: int, inverse: bool = False, batch: int = 1):
        off, keep = _buf(self.offset)
        self.ctx.check(lib().b200zk_ntt_fr_device(self.ctx.handle, C.c_void_p(dptr), self.log_size,
                                                  1 if inverse else 0, off, batch))


# ----------------------------------------------------------------------------- relation + Groth16
DEPOSIT, WITHDRAW = 0, 1


def poseidon_constants():
    """(round constants 64x5, MDS 5x5) as Montgomery bytes -- host-side Grain LFSR of the library."""
    rc = np.zeros(64 * 5 * 32, dtype=np.uint8)
    mds = np.zeros(25 * 32, dtype=np.uint8)
    rcode = lib().b200zk_poseidon_constants(rc.ctypes.data_as(C.c_void_p), mds.ctypes.data_as(C.c_void_p))
    if rcode != 0:
        raise B200zkError(rcode, "poseidon_constants")
    return rc, mds


class UpdateNoteRelation:
    """The R1CS of update_note_circuit (shielder/relations/src/relations/update_note.rs:106-149):
    the ConstraintSynthesizer of this backend.  Host-only; needs no GPU."""

    _witness_fn = "b200zk_update_note_witness_batch"

    def __init__(self, kind: int = WITHDRAW, tree_height: int = 10):
        self.kind, self.tree_height = kind, tree_height
        self._h = C.c_void_p()
        rc = self._make(kind, tree_height)
        if rc != 0:
            raise B200zkError(rc, "r1cs synthesis")
        self._shape()

    def _make(self, kind, tree_height):
        self.n_inputs_per_proof = 18 + 2 * tree_height
        return lib().b200zk_update_note_r1cs(kind, tree_height, C.byref(self._h))

    def _shape(self):
        nc, ni, na = C.c_uint64(), C.c_uint64(), C.c_uint64()
        nnz = (C.c_uint64 * 3)()
        lib().b200zk_r1cs_shape(self._h, C.byref(nc), C.byref(ni), C.byref(na), C.byref(nnz))
        self.num_constraints, self.num_inputs, self.num_aux = nc.value, ni.value, na.value
        self.num_variables = ni.value + na.value
        self.nnz = list(nnz)

    @property
    def handle(self):
        return self._h

    def matrix(self, which: int):
        """-> (row_ptr uint64[nc+1], cols uint32[nnz], vals uint8[nnz*32])"""
        rp = np.zeros(self.num_constraints + 1, dtype=np.uint64)
        cols = np.zeros(max(1, self.nnz[which]), dtype=np.uint32)
        vals = np.zeros(max(1, self.nnz[which]) * 32, dtype=np.uint8)
        rc = lib().b200zk_r1cs_matrix(self._h, which, rp.ctypes.data_as(C.c_void_p), cols.ctypes.data_as(C.c_void_p),
                                      vals.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise B200zkError(rc, "b200zk_r1cs_matrix")
        return rp, cols[:self.nnz[which]], vals[:self.nnz[which] * 32]

    def witness_batch(self, ctx: Context, inputs, batch: int, device_out: int | None = None, want_host: bool = True):
        """K6: full assignments z (batch * num_variables * 32 B) and per-instance status."""
        pi, ki = _buf(inputs)
        out = np.zeros(batch * self.num_variables * 32, dtype=np.uint8) if want_host else None
        status = np.zeros(batch, dtype=np.uint8)
        if ki.nbytes != batch * self.n_inputs_per_proof * 32:
            raise ValueError("inputs: expected %d bytes (batch x %d x 32), got %d" % (batch * self.n_inputs_per_proof * 32,
                                                                                     self.n_inputs_per_proof, ki.nbytes))
        rc = getattr(lib(), self._witness_fn)(
            ctx.handle, self._h, pi, batch, out.ctypes.data_as(C.c_void_p) if want_host else None,
            C.c_void_p(device_out) if device_out else None, status.ctypes.data_as(C.c_void_p))
        if rc not in (0, -6):
            ctx.check(rc)
        return out, status

    def free(self):
        if self._h:
            lib().b200zk_r1cs_free(self._h)
            self._h = None

    __del__ = free


class UpdateAccountRelation(UpdateNoteRelation):
    """The R1CS of update_account_circuit as a relation of its own (relations/src/relations/update_account.rs:68-95):
    instance (old_account_hash, new_account_hash, amount, token, user), witness old_account; rows of 9 Fr."""

    _witness_fn = "b200zk_update_account_witness_batch"

    def __init__(self, kind: int = WITHDRAW):
        super().__init__(kind, 0)

    def _make(self, kind, tree_height):
        self.n_inputs_per_proof = 9
        return lib().b200zk_update_account_r1cs(kind, C.byref(self._h))


def poseidon_hash_batch(ctx: Context, inputs, arity: int) -> np.ndarray:
    pi, ki = _buf(inputs)
    n = ki.nbytes // (32 * arity)
    out = np.zeros(n * 32, dtype=np.uint8)
    ctx.check(lib().b200zk_poseidon_hash_batch(ctx.handle, pi, n, arity, out.ctypes.data_as(C.c_void_p)))
    return out


class MerkleTree:
    """The note tree, device-resident: the reference's `MerkleTree<DEPTH>` (shielder/contract/merkle.rs:11-106)
    with the circuit's Poseidon-2 node hash (relations/src/merkle_proof.rs:49-57).  Method names, results and
    error behaviour are the reference's: add_leaf -> leaf id (MERKLE_LIMIT_EXCEEDED when full), root
    (MERKLE_NON_EXISTING_NODE when empty), is_historical_root, gen_proof (MERKLE_PROOF_GEN_FAIL once the tree is
    full).  add_leaves / gen_proofs are the batched forms the GPU is for."""

    def __init__(self, ctx: Context, depth: int = 10, log_roots: bool = True):
        self.ctx, self.depth, self.size = ctx, depth, 1 << depth
        self._h = C.c_void_p()
        ctx.check(lib().b200zk_merkle_new(ctx.handle, depth, 1 if log_roots else 0, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    @property
    def next_leaf_idx(self) -> int:
        v = C.c_uint64()
        lib().b200zk_merkle_info(self._h, None, None, C.byref(v))
        return v.value

    def add_leaves(self, leaves, want_roots: bool = False, device_ptr: int | None = None, n: int | None = None):
        """leaves: n x 32 B Montgomery Fr (host) or a device pointer.  -> first leaf id [, roots n x 32 B]"""
        if device_ptr is not None:
            pl, on_dev = C.c_void_p(device_ptr), 1
        else:
            pl, kl = _buf(leaves)
            on_dev, n = 0, kl.nbytes // 32
        first = C.c_uint64()
        roots = np.zeros(n * 32, dtype=np.uint8) if want_roots else None
        self.ctx.check(lib().b200zk_merkle_add_leaves(self.ctx.handle, self._h, pl, on_dev, n, C.byref(first),
                                                      roots.ctypes.data_as(C.c_void_p) if want_roots else None))
        return (first.value, roots) if want_roots else first.value

    def add_leaf(self, leaf_mont: bytes) -> int:
        return self.add_leaves(bytes(leaf_mont))

    def root(self) -> bytes:
        out = np.zeros(32, dtype=np.uint8)
        self.ctx.check(lib().b200zk_merkle_root(self.ctx.handle, self._h, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def node(self, node_id: int) -> bytes:
        out = np.zeros(32, dtype=np.uint8)
        self.ctx.check(lib().b200zk_merkle_node(self.ctx.handle, self._h, node_id, out.ctypes.data_as(C.c_void_p)))
        return out.tobytes()

    def is_historical_root(self, root_mont: bytes) -> bool:
        r = C.c_int()
        pr, kr = _buf(bytes(root_mont))
        self.ctx.check(lib().b200zk_merkle_is_historical_root(self.ctx.handle, self._h, pr, C.byref(r)))
        return bool(r.value)

    def gen_proofs(self, leaf_ids):
        """-> (path uint8[n, depth, 32], path_shape uint8[n, depth])"""
        ids = np.ascontiguousarray(np.asarray(leaf_ids, dtype=np.uint64))
        n = ids.size
        path = np.zeros((n, self.depth, 32), dtype=np.uint8)
        shape = np.zeros((n, self.depth), dtype=np.uint8)
        self.ctx.check(lib().b200zk_merkle_gen_proofs(self.ctx.handle, self._h, ids.ctypes.data_as(C.c_void_p), n,
                                                      path.ctypes.data_as(C.c_void_p), shape.ctypes.data_as(C.c_void_p)))
        return path, shape

    def gen_proof(self, leaf_id: int):
        """-> ([sibling bytes] * depth, [bool] * depth)"""
        path, shape = self.gen_proofs([leaf_id])
        return [path[0, i].tobytes() for i in range(self.depth)], [bool(b) for b in shape[0]]

    def fill_update_note_inputs_device(self, d_leaf_ids: int, n: int, d_inputs: int):
        self.ctx.check(lib().b200zk_merkle_fill_update_note_inputs_device(self.ctx.handle, self._h, C.c_void_p(d_leaf_ids),
                                                                          n, C.c_void_p(d_inputs)))

    def free(self):
        if self._h:
            lib().b200zk_merkle_free(self.ctx.handle, self._h)
            self._h = None

    __del__ = free


class ProvingKey:
    """Device-resident ark_groth16::ProvingKey + constraint matrices."""

    def __init__(self, ctx: Context, relation: UpdateNoteRelation, handle, vk: np.ndarray | None):
        self.ctx, self.relation, self._h, self.vk = ctx, relation, handle, vk

    def export_query(self, which: int) -> np.ndarray:
        cnt = C.c_size_t()
        self.ctx.check(lib().b200zk_pk_export_query(self.ctx.handle, self._h, which, None, C.byref(cnt)))
        out = np.zeros(cnt.value * (192 if which == 2 else 96), dtype=np.uint8)
        self.ctx.check(lib().b200zk_pk_export_query(self.ctx.handle, self._h, which, out.ctypes.data_as(C.c_void_p),
                                                    C.byref(cnt)))
        return out

    def verifying_key(self) -> "VerifyingKey":
        if self.vk is None:
            raise B200zkError(-1, "this proving key was uploaded without its verifying key")
        return VerifyingKey(self.ctx, self.vk, self.relation.num_inputs)

    def serialize(self, vk: "VerifyingKey | None" = None) -> bytes:
        """ark CanonicalSerialize (compressed) of ProvingKey."""
        own = vk is None
        vk = vk or self.verifying_key()
        n = C.c_size_t()
        self.ctx.check(lib().b200zk_pk_serialize(self.ctx.handle, self._h, vk._h, None, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint8)
        self.ctx.check(lib().b200zk_pk_serialize(self.ctx.handle, self._h, vk._h, out.ctypes.data_as(C.c_void_p),
                                                 C.byref(n)))
        if own:
            vk.free()
        return bytes(out[:n.value])

    @classmethod
    def deserialize(cls, ctx: Context, relation: "UpdateNoteRelation", data: bytes, check_subgroup: bool = False,
                    precompute: bool = True) -> "ProvingKey":
        pd, kd = _buf(data)
        h, hv = C.c_void_p(), C.c_void_p()
        ctx.check(lib().b200zk_pk_deserialize(ctx.handle, relation.handle, pd, len(data), 1 if check_subgroup else 0,
                                              int(precompute), C.byref(h), C.byref(hv)))
        vk = VerifyingKey(ctx, handle=hv)
        raw = vk.export()
        vk.free()
        return cls(ctx, relation, h, raw)

    def free(self):
        if self._h:
            lib().b200zk_pk_free(self.ctx.handle, self._h)
            self._h = None

    __del__ = free


PROOF_STATUS = {0: "accepted", 1: "rejected", 2: "bad encoding", 3: "not on curve", 4: "not in subgroup", 5: "bad input"}


def points_compress(ctx: Context, group: int, affine) -> np.ndarray:
    """Affine FFI points -> zcash / ark-serialize compressed bytes (48 B G1, 96 B G2)."""
    pa, ka = _buf(affine)
    n = ka.size // (96 if group == 1 else 192)
    out = np.zeros(n * (48 if group == 1 else 96), dtype=np.uint8)
    ctx.check(lib().b200zk_points_compress(ctx.handle, group, pa, n, out.ctypes.data_as(C.c_void_p)))
    return out


def points_decompress(ctx: Context, group: int, data, check_subgroup: bool = True):
    """-> (affine FFI points, int32 status per point: 0 ok, 1 bad encoding, 2 not on curve, 3 not in subgroup)."""
    pd, kd = _buf(data)
    n = kd.size // (48 if group == 1 else 96)
    out = np.zeros(n * (96 if group == 1 else 192), dtype=np.uint8)
    st = np.zeros(n, dtype=np.int32)
    ctx.check(lib().b200zk_points_decompress(ctx.handle, group, pd, n, 1 if check_subgroup else 0,
                                             out.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p)))
    return out, st


class VerifyingKey:
    """Device-resident ark_groth16::PreparedVerifyingKey (e(alpha, beta) evaluated once at upload)."""

    def __init__(self, ctx: Context, vk_bytes=None, num_inputs: int | None = None, handle=None):
        self.ctx = ctx
        if handle is not None:
            self._h = handle
        else:
            pv, kv = _buf(vk_bytes)
            if num_inputs is None:
                num_inputs = (kv.size - 672) // 96
            self._h = C.c_void_p()
            ctx.check(lib().b200zk_vk_upload(ctx.handle, pv, num_inputs, C.byref(self._h)))
        n = C.c_uint32()
        ctx.check(lib().b200zk_vk_num_inputs(self._h, C.byref(n)))
        self.num_inputs = n.value

    def export(self) -> np.ndarray:
        out = np.zeros(672 + 96 * self.num_inputs, dtype=np.uint8)
        self.ctx.check(lib().b200zk_vk_export(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def serialize(self) -> bytes:
        """ark CanonicalSerialize (compressed) of VerifyingKey."""
        n = C.c_size_t()
        self.ctx.check(lib().b200zk_vk_serialize(self.ctx.handle, self._h, None, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint8)
        self.ctx.check(lib().b200zk_vk_serialize(self.ctx.handle, self._h, out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return bytes(out[:n.value])

    @classmethod
    def deserialize(cls, ctx: Context, data: bytes, check_subgroup: bool = True) -> "VerifyingKey":
        pd, kd = _buf(data)
        h = C.c_void_p()
        ctx.check(lib().b200zk_vk_deserialize(ctx.handle, pd, len(data), 1 if check_subgroup else 0, C.byref(h)))
        return cls(ctx, handle=h)

    def free(self):
        if getattr(self, "_h", None):
            lib().b200zk_vk_free(self.ctx.handle, self._h)
            self._h = None

    __del__ = free


class Groth16:
    """ark_groth16::Groth16::<Bls12_381> call sites (names per arkworks 0.4 [recall])."""

    @staticmethod
    def _check_verify_buffers(vk: "VerifyingKey", kp: np.ndarray, ki: np.ndarray, batch: int):
        """The C side reads batch*192 proof bytes and batch*(num_inputs-1)*32 input bytes: a short host buffer
        would be an out-of-bounds read, so sizes are checked here."""
        if kp.size != batch * 192:
            raise ValueError("proofs: expected %d bytes (batch x 192), got %d" % (batch * 192, kp.size))
        want = batch * (vk.num_inputs - 1) * 32
        if ki.size != want:
            raise ValueError("public_inputs: expected %d bytes (batch x (num_inputs - 1) x 32), got %d" % (want, ki.size))

    @staticmethod
    def verify_proofs(vk: VerifyingKey, proofs, public_inputs, batch: int | None = None, check_subgroup: bool = True,
                      device: bool = False) -> np.ndarray:
        """Groth16::verify_proof for a batch: proofs = batch x 192 B compressed, public_inputs = batch x
        (num_inputs - 1) Montgomery Fr (host buffers, or device pointers with device=True and batch given).
        -> int32 status per proof (PROOF_STATUS; 0 = accepted)."""
        ctx = vk.ctx
        if device:
            pp, pi = C.c_void_p(proofs), C.c_void_p(public_inputs)
        else:
            pp, kp = _buf(proofs)
            pi, ki = _buf(public_inputs)
            if batch is None:
                batch = kp.size // 192
            Groth16._check_verify_buffers(vk, kp, ki, batch)
        st = np.zeros(batch, dtype=np.int32)
        ctx.check(lib().b200zk_groth16_verify_batch(ctx.handle, vk._h, pp, pi, 1 if device else 0, batch,
                                                    1 if check_subgroup else 0, st.ctypes.data_as(C.c_void_p)))
        return st

    @staticmethod
    def verify_proofs_aggregate(vk: VerifyingKey, proofs, public_inputs, coeffs, batch: int | None = None,
                                check_subgroup: bool = True, device: bool = False) -> bool:
        """One verdict for the whole batch by a random linear combination; coeffs = batch x 16 B randomisers
        chosen by the caller."""
        ctx = vk.ctx
        if device:
            pp, pi = C.c_void_p(proofs), C.c_void_p(public_inputs)
        else:
            pp, kp = _buf(proofs)
            pi, ki = _buf(public_inputs)
            if batch is None:
                batch = kp.size // 192
            Groth16._check_verify_buffers(vk, kp, ki, batch)
        pc, kc = _buf(coeffs)
        if kc.size != batch * 16:
            raise ValueError("coeffs: expected %d bytes (batch x 16), got %d" % (batch * 16, kc.size))
        ok = C.c_int()
        ctx.check(lib().b200zk_groth16_verify_aggregate(ctx.handle, vk._h, pp, pi, 1 if device else 0, batch, pc,
                                                        1 if check_subgroup else 0, C.byref(ok)))
        return bool(ok.value)

    @staticmethod
    def generate_parameters_with_toxic_waste(ctx: Context, relation: UpdateNoteRelation, toxic, precompute=True):
        """toxic = (alpha, beta, gamma, delta, tau) ints.  -> ProvingKey (vk bytes in .vk)."""
        tb = b"".join(int(t % R_MOD).to_bytes(32, "little") for t in toxic)
        h = C.c_void_p()
        vk = np.zeros(672 + relation.num_inputs * 96, dtype=np.uint8)
        pt, kt = _buf(tb)
        ctx.check(lib().b200zk_groth16_setup(ctx.handle, relation.handle, pt, int(precompute), C.byref(h),
                                             vk.ctypes.data_as(C.c_void_p)))
        return ProvingKey(ctx, relation, h, vk)

    @staticmethod
    def pk_upload(ctx: Context, relation: UpdateNoteRelation, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2,
                  a_query, b_g1_query, b_g2_query, l_query, h_query, precompute=True):
        bufs = [_buf(x) for x in (alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query, b_g2_query,
                                  l_query, h_query)]
        h = C.c_void_p()
        ctx.check(lib().b200zk_pk_upload(ctx.handle, relation.handle, *[b[0] for b in bufs], int(precompute),
                                         C.byref(h)))
        return ProvingKey(ctx, relation, h, None)

    @staticmethod
    def create_proof_with_reduction(pk: ProvingKey, assignments, r, s, batch: int = 1, device_ptr: int | None = None,
                                    want_points: bool = False):
        """assignments: batch full assignments z (Montgomery bytes) or a device pointer; r, s: lists of
        ints (one per proof).  -> proofs (batch x 192 B) [, affine points (batch x 384 B)]."""
        ctx = pk.ctx
        rb = np.frombuffer(b"".join(int(x % R_MOD).to_bytes(32, "little") for x in r), dtype=np.uint8)
        sb = np.frombuffer(b"".join(int(x % R_MOD).to_bytes(32, "little") for x in s), dtype=np.uint8)
        proofs = np.zeros(batch * 192, dtype=np.uint8)
        points = np.zeros(batch * 384, dtype=np.uint8) if want_points else None
        if device_ptr is not None:
            pa, on_dev = C.c_void_p(device_ptr), 1
        else:
            pa, ka = _buf(assignments)
            on_dev = 0
        ctx.check(lib().b200zk_groth16_prove_batch(ctx.handle, pk._h, pa, on_dev, batch, rb.ctypes.data_as(C.c_void_p),
                                                   sb.ctypes.data_as(C.c_void_p), proofs.ctypes.data_as(C.c_void_p),
                                                   points.ctypes.data_as(C.c_void_p) if want_points else None))
        return (proofs, points) if want_points else proofs

    @staticmethod
    def prove_update_note_device(pk: ProvingKey, d_inputs: int, rb: np.ndarray, sb: np.ndarray, batch: int,
                                 proofs: np.ndarray | None = None):
        """Same as prove_update_note with the instance inputs already in device memory."""
        ctx = pk.ctx
        if proofs is None:
            proofs = np.zeros(batch * 192, dtype=np.uint8)
        ctx.check(lib().b200zk_update_note_prove_batch_device(ctx.handle, pk._h, C.c_void_p(d_inputs), batch,
                                                              rb.ctypes.data_as(C.c_void_p), sb.ctypes.data_as(C.c_void_p),
                                                              proofs.ctypes.data_as(C.c_void_p), None))
        return proofs

    @staticmethod
    def prove_submit(pk: ProvingKey, inputs, rb: np.ndarray, sb: np.ndarray, batch: int, proofs: np.ndarray,
                     status: np.ndarray | None = None, device_ptr: int | None = None) -> int:
        """Asynchronous form (b200zk_*_prove_submit): enqueue witness generation + proving of one batch and return a
        ticket; `proofs` (batch x 192 B uint8) and `status` are filled by prove_wait(ticket) and must stay alive until
        then.  inputs: host rows (bytes / uint8 array) or, with device_ptr, rows already in device memory.
        Up to two batches in flight."""
        ctx = pk.ctx
        if device_ptr is not None:
            pi, on_dev = C.c_void_p(device_ptr), 1
        else:
            pi, ki = _buf(inputs)
            on_dev = 0
            if ki.nbytes != batch * pk.relation.n_inputs_per_proof * 32:
                raise ValueError("inputs do not match the batch size")
        if rb.nbytes != batch * 32 or sb.nbytes != batch * 32 or proofs.nbytes != batch * 192:
            raise ValueError("r / s / proofs do not match the batch size")
        fn = (lib().b200zk_update_account_prove_submit if isinstance(pk.relation, UpdateAccountRelation)
              else lib().b200zk_update_note_prove_submit)
        t = C.c_uint64()
        ctx.check(fn(ctx.handle, pk._h, pi, on_dev, batch, rb.ctypes.data_as(C.c_void_p), sb.ctypes.data_as(C.c_void_p),
                     proofs.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p) if status is not None else None,
                     C.byref(t)))
        return t.value

    @staticmethod
    def prove_wait(pk: ProvingKey, ticket: int):
        pk.ctx.check(lib().b200zk_prove_wait(pk.ctx.handle, ticket))

    @staticmethod
    def create_random_proof(pk: ProvingKey, inputs, rng):
        """ark_groth16 `create_random_proof_with_reduction(circuit, &pk, rng)`: draws r, THEN s from `rng` (an object
        with .fr() -> int in [0, r), the stand-in for `Fr::rand(rng)`) and proves one instance."""
        r = rng.fr()
        s = rng.fr()
        proofs, status = Groth16.prove_update_note(pk, inputs, [r], [s], 1)
        return proofs

    @staticmethod
    def prove_update_note(pk: ProvingKey, inputs, r, s, batch: int):
        """Witness generation (K6) + proving in one call: the user-facing path.  r, s: uint8 arrays
        (batch x 32 B canonical) or lists of ints."""
        ctx = pk.ctx
        def scal(x):
            if isinstance(x, np.ndarray):
                return x
            return np.frombuffer(b"".join(int(v % R_MOD).to_bytes(32, "little") for v in x), dtype=np.uint8)
        rb, sb = scal(r), scal(s)
        pi, ki = _buf(inputs)
        proofs = np.zeros(batch * 192, dtype=np.uint8)
        status = np.zeros(batch, dtype=np.uint8)
        fn = (lib().b200zk_update_account_prove_batch if isinstance(pk.relation, UpdateAccountRelation)
              else lib().b200zk_update_note_prove_batch)
        if ki.nbytes != batch * pk.relation.n_inputs_per_proof * 32 or rb.nbytes != batch * 32 or sb.nbytes != batch * 32:
            raise ValueError("inputs / r / s do not match the batch size")
        rc = fn(ctx.handle, pk._h, pi, batch, rb.ctypes.data_as(C.c_void_p), sb.ctypes.data_as(C.c_void_p),
                proofs.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p))
        ctx.check(rc)
        return proofs, status
