"""Multi-GPU host side: one process per GPU.

The product path is the C ABI (csrc/comm.cu): `b200zk_comm_init` attaches an NCCL communicator to the ctx and
`b200zk_msm_sharded[_device]` / `b200zk_ntt_sharded_device` run local kernels, collective and combine step as one
stream of work inside libb200zk.so -- no torch on the data path, no host synchronisation between the steps.  With
the GPU backend the classes below are thin callers of those entry points; `torch.distributed` is only the
bootstrap that carries the 128-byte NCCL id from rank 0 to the others.

* ShardedMSM      -- a large MSM split by contiguous point range.  Every rank runs the full bucket
                     pipeline on its slice; the last kernel leaves the affine partial in this rank's slot of
                     the all_gather buffer; one in-place all_gather of world * 96 B (G1) / 192 B (G2) follows and
                     every rank adds the partials -- curve addition is not an NCCL reduction op.
* ShardedNTT      -- a large Fr NTT as a four-step transform with ONE exchange: n = n1 * n2, rank g owns
                     the columns j2 in [g*C, (g+1)*C) (C = n2 / G) as local[c][j1] = x[j1*n2 + g*C + c].
                     Local column transforms -> twiddle w_n^(j2*k1) fused into a transpose that leaves the
                     data chunk-major by destination -> all_to_all of (n1/G) x C chunks -> rows interleaved ->
                     local row transforms.  The output has the same kind of layout with n1 and n2 swapped
                     (rank h: out[r][k2] = X[(h*R + r) + n1*k2]), so an inverse transform consumes it directly.
* ProofSharder    -- independent proofs of a batch spread round-robin over the ranks; no collective on
                     the data path, only a gather of the 192-byte proofs to rank 0.

The same partition / exchange steps are also spelled out here over a pluggable `backend`, so that they run with the
gloo backend on CPU in the tests (tests/test_sharded_cpu.py plugs the oracle in as the arithmetic): that generic
path is the executable description of what comm.cu does, not a second product path -- GpuBackend never takes it.
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
    """Product backend: libb200zk on this rank's GPU, communicator inside the library (csrc/comm.cu)."""

    native = True     # ShardedMSM / ShardedNTT call the C ABI's sharded entry points directly

    def __init__(self, ctx, dist=None):
        import torch
        self.ctx, self.torch = ctx, torch
        self.device = torch.device("cuda", torch.cuda.current_device())
        if ctx.comm_info()[1] == 1 and dist is not None and dist.get_world_size() > 1:
            ctx.comm_init_from_dist(dist)

    def buffer(self, nbytes: int):
        return self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)


class ShardedMSM:
    """sum_i scalars[i] * bases[i] over `world` GPUs; each rank holds bases[lo:hi] resident.

    bases_local: this rank's `VariableBaseMSM.Bases` (or any object the backend's msm_into accepts)."""

    def __init__(self, backend, group: int, bases_local, n_total: int, dist=None):
        self.backend, self.group, self.bases, self.n_total = backend, group, bases_local, n_total
        self.d = _Dist(dist)
        self.lo, self.hi = shard_range(n_total, self.d.rank, self.d.world)
        self.pt = G1_BYTES if group == 1 else G2_BYTES
        self.native = getattr(backend, "native", False)
        if not self.native:
            self.buf = backend.gather_buffer(self.d.world * self.pt)

    def msm(self, scalars_local) -> bytes:
        """scalars_local: this rank's (hi - lo) canonical scalars (host uint8 array, or a device address).
        Returns the affine result, identical on every rank."""
        n = self.hi - self.lo
        if self.native:      # b200zk_msm_sharded: local pipeline -> in-place all_gather -> sum, one stream
            if isinstance(scalars_local, int):
                return self.backend.ctx.msm_sharded(self.bases, device_ptr=scalars_local, n=n)[0]
            return self.backend.ctx.msm_sharded(self.bases, scalars=scalars_local, n=n)[0]
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
        self.native = getattr(backend, "native", False)
        if not self.native:
            self.send = backend.buffer(nbytes)
            self.recv = backend.buffer(nbytes) if w > 1 else None

    def swapped(self) -> "ShardedNTT":
        return ShardedNTT(self.backend, self.log_n, self.d.dist, self.log_n2)

    def _run(self, local, inverse: bool):
        if self.native:      # b200zk_ntt_sharded_device: asynchronous on the ctx stream, in place
            self.backend.ctx.ntt_sharded_device(local.data_ptr(), self.log_n, self.log_n1, inverse)
            return local
        b, w, g = self.backend, self.d.world, self.d.rank
        n1, n2 = 1 << self.log_n1, 1 << self.log_n2
        b.ntt_batch(local, self.log_n1, inverse, self.C)                               # columns, in place
        b.twiddle_transpose(local, self.send, self.log_n, self.C, n1, g * self.C, inverse)  # -> [n1][C]
        if w == 1:
            out = local                                                                  # send is [n1][n2] already
            b.copy2d(out, 0, n2 * 32, self.send, 0, n2 * 32, n2 * 32, n1)
        else:
            self.d.dist.all_to_all_single(self.recv, self.send)                          # recv[g'][r][c]
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
