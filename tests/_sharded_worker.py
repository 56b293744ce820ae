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
