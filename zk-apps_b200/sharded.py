"""Multi-GPU host side: one process per GPU, `torch.distributed` for the plumbing.

* ShardedMSM      -- a large MSM split by contiguous point range.  Every rank runs the full bucket
                     pipeline on its slice (b200zk_msm_resident_device leaves the affine partial in this
                     rank's slot of the all_gather buffer, in device memory); one all_gather of
                     world * 96 B (G1) / 192 B (G2) follows and every rank adds the partials
                     (b200zk_points_sum_device) -- curve addition is not an NCCL reduction op.
* ShardedNTT      -- a large Fr NTT as a four-step transform with ONE exchange: n = n1 * n2, rank g owns
                     the columns j2 in [g*C, (g+1)*C) (C = n2 / G) as local[c][j1] = x[j1*n2 + g*C + c].
                     Local column transforms -> twiddle w_n^(j2*k1) fused into a transpose that leaves the
                     data chunk-major by destination (b200zk_ntt_twiddle_transpose_device) -> all_to_all
                     of (n1/G) x C chunks -> rows interleaved -> local row transforms.  The output has the
                     same kind of layout with n1 and n2 swapped (rank h: out[r][k2] = X[(h*R + r) + n1*k2]),
                     so an inverse transform consumes it directly.
* ProofSharder    -- independent proofs of a batch spread round-robin over the ranks; no collective on
                     the data path, only a gather of the 192-byte proofs to rank 0.

The arithmetic is delegated to a `backend`, so the partition / exchange logic runs unchanged with the
gloo backend on CPU in the tests (tests/test_sharded_cpu.py plugs the oracle in as the backend; the
product backend below is the only one this package ships and it needs the GPU library).
"""
from __future__ import annotations

import numpy as np

from .ffi import G1_BYTES, G2_BYTES


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous slice [lo, hi) of n items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def round_robin(n: int, rank: int, world: int) -> list:
    """Indices of the items of a batch that `rank` proves: i with i % world == rank."""
    return list(range(rank, n, world))


class _Dist:
    """Thin view of torch.distributed that also works for world size 1 without a process group."""

    def __init__(self, dist=None):
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0


class GpuBackend:
    """Product backend: libb200zk on this rank's GPU."""

    def __init__(self, ctx):
        import torch
        self.ctx, self.torch = ctx, torch
        self.device = torch.device("cuda", torch.cuda.current_device())

    def gather_buffer(self, nbytes: int):
        return self.torch.zeros(nbytes, dtype=self.torch.uint8, device=self.device)

    def msm_into(self, bases, scalars, n, buf, offset):
        bases.msm_to_device(buf.data_ptr() + offset, scalars=scalars if not isinstance(scalars, int) else None,
                            device_ptr=scalars if isinstance(scalars, int) else None, n=n)
        self.ctx.sync()                                   # the collective runs on torch's stream

    def sum_points(self, group, buf, count):
        from .ffi import lib
        import ctypes as C
        pt = G1_BYTES if group == 1 else G2_BYTES
        out = self.torch.zeros(pt, dtype=self.torch.uint8, device=self.device)
        self.torch.cuda.current_stream().synchronize()    # all_gather done before our stream reads the buffer
        self.ctx.check(lib().b200zk_points_sum_device(self.ctx.handle, group, C.c_void_p(buf.data_ptr()), count,
                                                      C.c_void_p(out.data_ptr())))
        self.ctx.sync()
        return out.cpu().numpy().tobytes()


    # ---- ShardedNTT primitives: buffers are torch uint8 tensors on this rank's GPU
    def buffer(self, nbytes: int):
        return self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)

    def ntt_batch(self, buf, log_len: int, inverse: bool, batch: int):
        from .ffi import lib
        import ctypes as C
        self.ctx.check(lib().b200zk_ntt_fr_device(self.ctx.handle, C.c_void_p(buf.data_ptr()), log_len,
                                                  1 if inverse else 0, None, batch))

    def twiddle_transpose(self, src, dst, log_n: int, rows: int, cols: int, row0: int, inverse: bool):
        from .ffi import lib
        import ctypes as C
        self.ctx.check(lib().b200zk_ntt_twiddle_transpose_device(self.ctx.handle, C.c_void_p(src.data_ptr()),
                                                                 C.c_void_p(dst.data_ptr()), log_n, rows, cols, row0,
                                                                 1 if inverse else 0))

    def copy2d(self, dst, dst_off: int, dpitch: int, src, src_off: int, spitch: int, width: int, height: int):
        from .ffi import lib
        import ctypes as C
        self.ctx.check(lib().b200zk_copy2d_device(self.ctx.handle, C.c_void_p(dst.data_ptr() + dst_off), dpitch,
                                                  C.c_void_p(src.data_ptr() + src_off), spitch, width, height))

    def before_collective(self):
        self.ctx.sync()                                   # our stream -> torch's (NCCL) stream

    def after_collective(self):
        self.torch.cuda.current_stream().synchronize()    # NCCL done before our stream touches the buffer


class ShardedMSM:
    """sum_i scalars[i] * bases[i] over `world` GPUs; each rank holds bases[lo:hi] resident.

    bases_local: this rank's `VariableBaseMSM.Bases` (or any object the backend's msm_into accepts)."""

    def __init__(self, backend, group: int, bases_local, n_total: int, dist=None):
        self.backend, self.group, self.bases, self.n_total = backend, group, bases_local, n_total
        self.d = _Dist(dist)
        self.lo, self.hi = shard_range(n_total, self.d.rank, self.d.world)
        self.pt = G1_BYTES if group == 1 else G2_BYTES
        self.buf = backend.gather_buffer(self.d.world * self.pt)

    def msm(self, scalars_local) -> bytes:
        """scalars_local: this rank's (hi - lo) canonical scalars (host uint8 array, or a device address).
        Returns the affine result, identical on every rank."""
        n = self.hi - self.lo
        self.backend.msm_into(self.bases, scalars_local, n, self.buf, self.d.rank * self.pt)
        if self.d.world > 1:
            mine = self.buf[self.d.rank * self.pt:(self.d.rank + 1) * self.pt]
            self.d.dist.all_gather_into_tensor(self.buf, mine.clone())
        return self.backend.sum_points(self.group, self.buf, self.d.world)


def ntt_split(log_n: int) -> tuple:
    """(log n1, log n2) of the four-step factorisation n = n1 * n2 (n1 >= n2)."""
    return (log_n + 1) // 2, log_n // 2


def ntt_local_from_natural(x: np.ndarray, log_n1: int, log_n2: int, rank: int, world: int) -> np.ndarray:
    """Rows of 32-byte elements x[j] (shape [n, 32]) -> this rank's local[c][j1] = x[j1*n2 + rank*C + c]."""
    n1, n2 = 1 << log_n1, 1 << log_n2
    C_ = n2 // world
    m = x.reshape(n1, n2, 32)[:, rank * C_:(rank + 1) * C_, :]
    return np.ascontiguousarray(m.transpose(1, 0, 2)).reshape(C_ * n1, 32)


def ntt_natural_from_locals(parts: list, log_n1: int, log_n2: int) -> np.ndarray:
    """Inverse of ntt_local_from_natural over all ranks' arrays (layout (n1, n2)) -> [n, 32]."""
    n1, n2 = 1 << log_n1, 1 << log_n2
    world = len(parts)
    C_ = n2 // world
    out = np.zeros((n1, n2, 32), dtype=np.uint8)
    for g, p in enumerate(parts):
        out[:, g * C_:(g + 1) * C_, :] = p.reshape(C_, n1, 32).transpose(1, 0, 2)
    return out.reshape(n1 * n2, 32)


class ShardedNTT:
    """Radix-2 Fr NTT of size 2^log_n over `world` GPUs ('large NTT').

    Layout L(a, b) of a length-a*b vector v: rank g holds local[c][j] = v[j*b + g*(b/world) + c].
    forward()/inverse() take this rank's part in layout L(n1, n2) and return the transform in layout
    L(n2, n1); apply the other direction with the factors swapped (`swapped()`) to come back."""

    def __init__(self, backend, log_n: int, dist=None, log_n1: int | None = None):
        self.backend, self.log_n, self.d = backend, log_n, _Dist(dist)
        self.log_n1, self.log_n2 = ntt_split(log_n) if log_n1 is None else (log_n1, log_n - log_n1)
        w = self.d.world
        if w & (w - 1) or (1 << self.log_n1) < w or (1 << self.log_n2) < w:
            raise ValueError("world size must be a power of two no larger than either factor of n")
        self.C = (1 << self.log_n2) // w          # columns this rank owns
        self.R = (1 << self.log_n1) // w          # rows this rank owns after the exchange
        nbytes = self.C * (1 << self.log_n1) * 32
        self.send = backend.buffer(nbytes)
        self.recv = backend.buffer(nbytes) if w > 1 else None

    def swapped(self) -> "ShardedNTT":
        return ShardedNTT(self.backend, self.log_n, self.d.dist, self.log_n2)

    def _run(self, local, inverse: bool):
        b, w, g = self.backend, self.d.world, self.d.rank
        n1, n2 = 1 << self.log_n1, 1 << self.log_n2
        b.ntt_batch(local, self.log_n1, inverse, self.C)                               # columns, in place
        b.twiddle_transpose(local, self.send, self.log_n, self.C, n1, g * self.C, inverse)  # -> [n1][C]
        if w == 1:
            out = local                                                                  # send is [n1][n2] already
            b.copy2d(out, 0, n2 * 32, self.send, 0, n2 * 32, n2 * 32, n1)
        else:
            b.before_collective()
            self.d.dist.all_to_all_single(self.recv, self.send)                          # recv[g'][r][c]
            b.after_collective()
            out = local                                                                  # R x n2, same byte count
            chunk = self.R * self.C * 32
            for src in range(w):                                                         # out[r][src*C + c]
                b.copy2d(out, src * self.C * 32, n2 * 32, self.recv, src * chunk, self.C * 32, self.C * 32, self.R)
        b.ntt_batch(out, self.log_n2, inverse, self.R)                                   # rows, in place
        return out

    def forward(self, local):
        """local: this rank's buffer in layout L(n1, n2), transformed in place; returns it in layout L(n2, n1)."""
        return self._run(local, False)

    def inverse(self, local):
        return self._run(local, True)


class ProofSharder:
    """Round-robin distribution of a batch of independent proofs (BASELINE.json config 5)."""

    def __init__(self, prove_fn, dist=None, torch_device=None):
        """prove_fn(inputs_rows: np.ndarray[k, row_bytes], r: np.ndarray[k,32], s: np.ndarray[k,32]) -> np.ndarray[k*192]"""
        self.prove_fn, self.d, self.device = prove_fn, _Dist(dist), torch_device

    def prove(self, inputs: np.ndarray, r: np.ndarray, s: np.ndarray):
        """inputs: [batch, row_bytes] uint8, r/s: [batch, 32].  Every rank passes the same arrays (or at least
        its own rows); rank 0 gets all proofs in batch order ([batch*192] uint8), other ranks get None."""
        import torch
        batch = inputs.shape[0]
        idx = round_robin(batch, self.d.rank, self.d.world)
        mine = self.prove_fn(inputs[idx], r[idx], s[idx]) if idx else np.zeros(0, dtype=np.uint8)
        if self.d.world == 1:
            return np.asarray(mine, dtype=np.uint8)
        per = (batch + self.d.world - 1) // self.d.world                      # pad so every rank sends the same size
        send = torch.zeros(per * 192, dtype=torch.uint8)
        send[:len(idx) * 192] = torch.from_numpy(np.ascontiguousarray(mine, dtype=np.uint8))
        if self.device is not None:
            send = send.to(self.device)
        recv = [torch.zeros_like(send) for _ in range(self.d.world)] if self.d.rank == 0 else None
        self.d.dist.gather(send, recv, dst=0)
        if self.d.rank != 0:
            return None
        out = np.zeros(batch * 192, dtype=np.uint8)
        for rk in range(self.d.world):
            got = recv[rk].cpu().numpy()
            for j, i in enumerate(round_robin(batch, rk, self.d.world)):
                out[i * 192:(i + 1) * 192] = got[j * 192:(j + 1) * 192]
        return out
