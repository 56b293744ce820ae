"""Worker for tests/test_sharded_cpu.py: world_size-2 gloo run of the multi-GPU host logic
(zk-apps_b200/sharded.py) with the ORACLE plugged in as the arithmetic backend (test infrastructure:
the product backend needs the GPU library).  Launched with torch.multiprocessing.spawn."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleBackend:
    """CPU stand-in for GpuBackend: same interface, arithmetic by oracle/c."""

    def __init__(self):
        import torch
        from oracle import corac
        self.torch, self.corac = torch, corac

    def gather_buffer(self, nbytes):
        return self.torch.zeros(nbytes, dtype=self.torch.uint8)

    def msm_into(self, bases, scalars, n, buf, offset):
        group, pts = bases
        out = np.frombuffer(bytes(self.corac.msm(group, pts, scalars)), dtype=np.uint8)
        buf[offset:offset + len(out)] = self.torch.from_numpy(out.copy())

    def sum_points(self, group, buf, count):
        ones = np.frombuffer(b"".join((1).to_bytes(32, "little") for _ in range(count)), dtype=np.uint8)
        return bytes(self.corac.msm(group, buf.numpy().copy(), ones))


    # ---- ShardedNTT primitives on host memory (torch uint8 CPU tensors)
    def buffer(self, nbytes):
        return self.torch.zeros(nbytes, dtype=self.torch.uint8)

    def ntt_batch(self, buf, log_len, inverse, batch):
        out = self.corac.ntt(buf.numpy()[:batch * (32 << log_len)], log_len, inverse, None, batch)
        buf[:len(out)] = self.torch.from_numpy(out)

    def twiddle_transpose(self, src, dst, log_n, rows, cols, row0, inverse):
        from oracle.pyref import bls12_381 as bls
        from oracle.pyref.algos import Domain
        w = Domain(1 << log_n).group_gen
        if inverse:
            w = pow(w, -1, bls.R)
        a = src.numpy().reshape(rows, cols, 32)
        out = np.zeros((cols, rows, 32), dtype=np.uint8)
        for c in range(rows):
            for k in range(cols):
                v = bls.fr_from_mont_bytes(a[c, k].tobytes()) * pow(w, (row0 + c) * k, bls.R) % bls.R
                out[k, c] = np.frombuffer(bls.fr_to_mont_bytes(v), dtype=np.uint8)
        dst[:] = self.torch.from_numpy(out.reshape(-1))

    def copy2d(self, dst, dst_off, dpitch, src, src_off, spitch, width, height):
        d, s_ = dst.numpy(), src.numpy()
        for r in range(height):
            d[dst_off + r * dpitch:dst_off + r * dpitch + width] = s_[src_off + r * spitch:src_off + r * spitch + width]

    def before_collective(self):
        pass

    def after_collective(self):
        pass


def run(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import zk_apps_b200  # noqa: F401  (import shim)
    from zk_apps_b200 import sharded
    from oracle import corac
    from oracle.pyref import bls12_381 as bls
    from tests import util
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        # ---- sharded MSM: both groups, uneven split (n not divisible by world), one empty-ish case
        for group, cv, enc, n in ((1, bls.G1, util.g1_array, 37), (2, bls.G2, util.g2_array, 9), (1, bls.G1, util.g1_array, 1)):
            ks = [k % 997 + 1 for k in util.rand_fr(100 + n, n)]
            ss = util.rand_fr(200 + n, n)
            pts = corac.fixed_base_mul(group, bytes(enc([cv.gen])), util.scalars_array(ks))
            pt = 96 * group
            lo, hi = sharded.shard_range(n, rank, world)
            local = (group, pts[lo * pt:hi * pt].copy())
            m = sharded.ShardedMSM(OracleBackend(), group, local, n, dist)
            got = m.msm(util.scalars_array(ss[lo:hi]))
            want = bytes(enc([cv.mul(cv.gen, sum(a * b for a, b in zip(ks, ss)) % bls.R)]))
            assert got == want, (group, n, rank)
        # ---- four-step NTT with one all_to_all: forward vs the single-vector oracle, then inverse round trip
        import torch
        for log_n in (4, 7):
            n = 1 << log_n
            vals = util.rand_fr(300 + log_n, n)
            x = np.frombuffer(b"".join(bls.fr_to_mont_bytes(v) for v in vals), dtype=np.uint8).reshape(n, 32)
            want = corac.ntt(x.reshape(-1), log_n).reshape(n, 32)
            fwd = sharded.ShardedNTT(OracleBackend(), log_n, dist)
            l1, l2 = fwd.log_n1, fwd.log_n2
            local = torch.from_numpy(sharded.ntt_local_from_natural(x, l1, l2, rank, world).reshape(-1).copy())
            out = fwd.forward(local)
            mine = sharded.ntt_local_from_natural(want, l2, l1, rank, world).reshape(-1)      # layout L(n2, n1)
            assert bytes(out.numpy()) == bytes(mine), ("ntt forward", log_n, rank)
            back = fwd.swapped().inverse(out)
            assert bytes(back.numpy()) == bytes(sharded.ntt_local_from_natural(x, l1, l2, rank, world).reshape(-1)), \
                ("ntt round trip", log_n, rank)
        # ---- round-robin proofs: a fake prover that tags each proof with its row, so placement is checked
        batch = 7
        inputs = np.arange(batch * 4, dtype=np.uint8).reshape(batch, 4)
        r = np.zeros((batch, 32), dtype=np.uint8); s = np.zeros((batch, 32), dtype=np.uint8)
        r[:, 0] = np.arange(batch) + 100

        def fake_prove(rows, rr, ss_):
            out = np.zeros(len(rows) * 192, dtype=np.uint8)
            for j in range(len(rows)):
                out[j * 192] = rows[j][0]; out[j * 192 + 1] = rr[j][0]; out[j * 192 + 2] = rank
            return out
        res = sharded.ProofSharder(fake_prove, dist).prove(inputs, r, s)
        if rank == 0:
            for i in range(batch):
                assert res[i * 192] == inputs[i][0] and res[i * 192 + 1] == 100 + i and res[i * 192 + 2] == i % world
        else:
            assert res is None
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()
