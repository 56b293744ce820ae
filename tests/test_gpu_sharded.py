"""Multi-GPU building blocks on the GPU: b200zk_points_sum, b200zk_msm_resident_device and the
ShardedMSM / ProofSharder host classes with the product (GPU) backend.  world_size 1 always runs; the
NCCL world_size-2 case runs when the box has at least 2 GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

import zk_apps_b200 as z
from zk_apps_b200 import sharded
from zk_apps_b200.ffi import points_sum
from oracle.pyref import bls12_381 as bls
from tests import util

pytestmark = pytest.mark.gpu
R = bls.R
CURVES = {1: (bls.G1, util.g1_array, util.g1_list), 2: (bls.G2, util.g2_array, util.g2_list)}


@pytest.mark.parametrize("group", [1, 2])
def test_points_sum(ctx, group):
    cv, enc, dec = CURVES[group]
    g = cv.gen
    pts = [cv.mul(g, k) for k in (3, 5, 7, 11)]
    out, inf = points_sum(ctx, group, enc(pts))
    assert dec(out)[0] == cv.mul(g, 26) and not inf
    out, inf = points_sum(ctx, group, enc([pts[0], cv.neg(pts[0])]))
    assert inf and dec(out)[0] is None
    out, inf = points_sum(ctx, group, enc([pts[1], None, pts[1]]))                     # infinity operand + doubling
    assert dec(out)[0] == cv.mul(g, 10)
    many = [cv.mul(g, k + 1) for k in range(70)]                                        # more points than lanes
    out, _ = points_sum(ctx, group, enc(many))
    assert dec(out)[0] == cv.mul(g, 70 * 71 // 2)


@pytest.mark.parametrize("group,n", [(1, 500), (2, 60)])
def test_sharded_msm_world1(ctx, group, n):
    import torch
    cv, enc, dec = CURVES[group]
    ks = util.scalars_array([k % 100000 + 1 for k in util.rand_fr(n, n)])
    pts = ctx.fixed_base_mul(group, ks)
    ss = util.rand_fr(n + 1, n)
    want, _ = z.VariableBaseMSM.msm_bigint(ctx, group, pts, util.scalars_array(ss))
    torch.cuda.set_device(0)
    bases = z.VariableBaseMSM.Bases(ctx, group, pts)
    m = sharded.ShardedMSM(sharded.GpuBackend(ctx), group, bases, n, None)
    assert m.msm(util.scalars_array(ss)) == want
    kk = [int.from_bytes(bytes(ks[32 * i:32 * i + 32]), "little") for i in range(n)]
    assert dec(want)[0] == cv.mul(cv.gen, sum(a * b for a, b in zip(kk, ss)) % R)
    bases.free()


def _nccl_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = z.Context(rank)
        n = 3001
        ks = util.rand_fr_bytes_fast(5, n)
        pts = ctx.fixed_base_mul(1, ks)
        ss = util.rand_fr_bytes_fast(6, n)
        want, _ = z.VariableBaseMSM.msm_bigint(ctx, 1, pts, ss)
        lo, hi = sharded.shard_range(n, rank, world)
        bases = z.VariableBaseMSM.Bases(ctx, 1, pts[lo * 96:hi * 96].copy())
        m = sharded.ShardedMSM(sharded.GpuBackend(ctx), 1, bases, n, dist)
        assert m.msm(ss[lo * 32:hi * 32].copy()) == want
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
        bases.free()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_sharded_msm_world2_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
