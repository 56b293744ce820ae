"""Multi-GPU host side: one process per GPU, `torch.distributed` for the plumbing.

* ShardedMSM      -- a large MSM split by contiguous point range.  Every rank runs the full bucket
                     pipeline on its slice (b200zk_msm_resident_device leaves the affine partial in this
                     rank's slot of the all_gather buffer, in device memory); one all_gather of
                     world * 96 B (G1) / 192 B (G2) follows and every rank adds the partials
                     (b200zk_points_sum_device) -- curve addition is not an NCCL reduction op.
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
