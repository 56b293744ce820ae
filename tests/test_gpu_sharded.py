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


def _sharded_ntt_check(ctx, be, d, rank, world, log_n, seed):
    """forward over `world` ranks == the single-GPU transform of the whole vector; inverse brings it back"""
    import torch
    n = 1 << log_n
    x = util.rand_fr_bytes_fast(seed, n).reshape(n, 32)
    want = z.Radix2EvaluationDomain(ctx, log_n).fft(x.reshape(-1)).reshape(n, 32)
    fwd = sharded.ShardedNTT(be, log_n, d)
    l1, l2 = fwd.log_n1, fwd.log_n2
    mine = sharded.ntt_local_from_natural(x, l1, l2, rank, world).reshape(-1)
    local = torch.from_numpy(mine.copy()).to(be.device)
    out = fwd.forward(local)
    ctx.sync()
    assert bytes(out.cpu().numpy()) == bytes(sharded.ntt_local_from_natural(want, l2, l1, rank, world).reshape(-1)), log_n
    back = fwd.swapped().inverse(out)
    ctx.sync()
    assert bytes(back.cpu().numpy()) == bytes(mine), log_n


def test_twiddle_transpose(ctx):
    """out[k][c] = in[c][k] * w_n^((row0 + c) * k), both directions, ragged tile edges"""
    import ctypes as C
    from oracle.pyref.algos import Domain
    log_n, rows, cols, row0 = 7, 5, 40, 3                     # (row0 + c) * k wraps mod n
    vals = util.rand_fr(77, rows * cols)
    d_in = ctx.alloc(rows * cols * 32); d_out = ctx.alloc(rows * cols * 32)
    ctx.upload(d_in, util.fr_mont_array(vals))
    for inverse in (0, 1):
        w = Domain(1 << log_n).group_gen
        if inverse:
            w = pow(w, -1, R)
        ctx.check(z.lib().b200zk_ntt_twiddle_transpose_device(ctx.handle, C.c_void_p(d_in), C.c_void_p(d_out), log_n,
                                                              rows, cols, row0, inverse))
        got = util.fr_from_mont_array(ctx.download(d_out, rows * cols * 32))
        want = [vals[c * cols + k] * pow(w, (row0 + c) * k, R) % R for k in range(cols) for c in range(rows)]
        assert got == want
    ctx.free(d_in); ctx.free(d_out)


@pytest.mark.parametrize("log_n", [2, 5, 12, 17, 20])
def test_sharded_ntt_world1(ctx, log_n):
    import torch
    torch.cuda.set_device(0)
    _sharded_ntt_check(ctx, sharded.GpuBackend(ctx), None, 0, 1, log_n, 40 + log_n)


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
        be = sharded.GpuBackend(ctx, dist)                       # b200zk_comm_init: NCCL id from rank 0 through `dist`
        assert ctx.comm_info() == (rank, world)
        m = sharded.ShardedMSM(be, 1, bases, n, dist)
        assert m.msm(ss[lo * 32:hi * 32].copy()) == want           # b200zk_msm_sharded (host scalars)
        d_s = ctx.alloc((hi - lo) * 32); ctx.upload(d_s, ss[lo * 32:hi * 32].copy())
        d_o = ctx.alloc(96)
        ctx.msm_sharded_device(bases, d_s, hi - lo, d_o)           # asynchronous form, result stays on the device
        ctx.sync()
        assert bytes(ctx.download(d_o, 96)) == want
        ctx.free(d_s); ctx.free(d_o)
        bases.free()
        # uneven slices incl. an EMPTY one, G2
        n2 = 41
        ks2 = util.rand_fr_bytes_fast(7, n2); pts2 = ctx.fixed_base_mul(2, ks2); ss2 = util.rand_fr_bytes_fast(8, n2)
        want2, _ = z.VariableBaseMSM.msm_bigint(ctx, 2, pts2, ss2)
        lo2, hi2 = (0, n2) if rank == 0 else (n2, n2)
        b2 = z.VariableBaseMSM.Bases(ctx, 2, pts2[lo2 * 192:hi2 * 192].copy() if hi2 > lo2 else np.zeros(0, dtype=np.uint8))
        got2, _ = ctx.msm_sharded(b2, scalars=ss2[lo2 * 32:hi2 * 32].copy(), n=hi2 - lo2)
        assert got2 == want2
        b2.free()
        for log_n in (2, 9, 16, 21):
            _sharded_ntt_check(ctx, be, dist, rank, world, log_n, 60 + log_n)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_init_multi_and_msm_sharded_multi(ctx):
    """Single-process form (b200zk_init_multi + b200zk_msm_sharded_multi): one host thread drives every GPU of the box.
    Runs over however many GPUs are visible (1 on the default test box: no NCCL needed then)."""
    import ctypes as C
    import torch
    G = min(torch.cuda.device_count(), 2)
    L = z.lib()
    ctxs = (C.c_void_p * G)()
    assert L.b200zk_init_multi(G, ctxs) == 0
    try:
        n = 777
        ks = util.rand_fr_bytes_fast(15, n); pts = ctx.fixed_base_mul(1, ks); ss = util.rand_fr_bytes_fast(16, n)
        want, _ = z.VariableBaseMSM.msm_bigint(ctx, 1, pts, ss)
        hs, keep, ptrs, ns = (C.c_void_p * G)(), [], (C.c_void_p * G)(), (C.c_size_t * G)()
        for g in range(G):
            lo, hi = sharded.shard_range(n, g, G)
            h = C.c_void_p()
            pb = np.ascontiguousarray(pts[lo * 96:hi * 96]); sb = np.ascontiguousarray(ss[lo * 32:hi * 32])
            keep += [pb, sb]
            assert L.b200zk_bases_upload(ctxs[g], 1, pb.ctypes.data_as(C.c_void_p), None, hi - lo, 0, C.byref(h)) == 0
            hs[g], ptrs[g], ns[g] = h, sb.ctypes.data, hi - lo
        out = np.zeros(96, dtype=np.uint8)
        inf = C.c_uint8()
        rc = L.b200zk_msm_sharded_multi(ctxs, G, hs, ptrs, 0, ns, out.ctypes.data_as(C.c_void_p), C.byref(inf))
        assert rc == 0, L.b200zk_last_error(ctxs[0])
        assert out.tobytes() == want and not inf.value
        for g in range(G):
            L.b200zk_bases_free(ctxs[g], hs[g])
    finally:
        for g in range(G):
            L.b200zk_destroy(ctxs[g])


def test_sharded_entry_points_without_communicator(ctx):
    """world == 1 needs no NCCL: the sharded entry points degenerate to the local ones; a ctx that was told it is
    one rank of several but has no communicator refuses."""
    n = 300
    ks = util.rand_fr_bytes_fast(25, n); pts = ctx.fixed_base_mul(1, ks); ss = util.rand_fr_bytes_fast(26, n)
    want, _ = z.VariableBaseMSM.msm_bigint(ctx, 1, pts, ss)
    bases = z.VariableBaseMSM.Bases(ctx, 1, pts)
    assert ctx.comm_info() == (0, 1)
    assert ctx.msm_sharded(bases, scalars=ss)[0] == want
    ctx.comm_init(None, 0, 1)
    assert ctx.msm_sharded(bases, scalars=ss)[0] == want
    with pytest.raises(z.B200zkError):
        ctx.comm_init(None, 0, 2)                                   # world 2 without an id
    bases.free()


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
