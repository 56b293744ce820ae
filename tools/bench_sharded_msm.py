#!/usr/bin/env python3
"""G1 MSM sweep across GPUs (BASELINE.json config 4): torchrun --nproc-per-node N tools/bench_sharded_msm.py
Each rank holds a contiguous range of n / N synthetic bases (k_i * G made on the GPU) resident; the timed
region is local MSM -> all_gather of the affine partials -> local sum, max over ranks, CUDA-event-free
wall clock bracketed by barriers (latency of one call, not throughput).  Rank 0 prints one JSON line per
size and appends it to gpurun_out/sharded_msm_N<world>.jsonl.  Correctness at full size: the result must
equal (sum_i s_i k_i mod r) * G, computed on the host with Python integers and one fixed-base multiply."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import zk_apps_b200 as z
from zk_apps_b200 import sharded

R = z.ffi.R_MOD
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,18,20,22,24")
ap.add_argument("--group", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", type=int, default=1)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
d = None
if world > 1:
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    d = dist
ctx = z.Context(local)
be = sharded.GpuBackend(ctx, d)
pt = 96 if args.group == 1 else 192


def rand32(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x3F
    return a


def as_ints(a):
    return [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]


for lg in [int(x) for x in args.sizes.split(",") if x]:
    n = 1 << lg
    lo, hi = sharded.shard_range(n, rank, world)
    m = hi - lo
    ks = rand32(m, 1000 + lg * 16 + rank)            # per-rank seeds: every rank generates only its slice
    ss = rand32(m, 5000 + lg * 16 + rank)
    dks = ctx.alloc(m * 32); ctx.upload(dks, ks.reshape(-1))
    dpts = ctx.alloc(m * pt)
    ctx.check(z.lib().b200zk_fixed_base_mul_device(ctx.handle, args.group, dks, m, dpts))
    bases = z.VariableBaseMSM.Bases(ctx, args.group, device_ptr=dpts, n=m)
    ctx.free(dpts)
    ctx.upload(dks, ss.reshape(-1))
    sm = sharded.ShardedMSM(be, args.group, bases, n, d)
    res = sm.msm(dks)                                  # warm-up (allocations, NCCL channels)
    times = []
    for _ in range(args.reps):
        if d: d.barrier(device_ids=[local])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sm.msm(dks)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if d:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda"); d.all_reduce(t, op=d.ReduceOp.MAX); dt = float(t.item())
        times.append(dt)
    ok = None
    if args.check:
        part = 0
        step = 1 << 16
        kk, sv = ks.reshape(-1), ss.reshape(-1)
        for i in range(0, m, step):                    # exact sum_i s_i k_i mod r of this rank's slice
            a = as_ints(ks[i:i + step]); b = as_ints(ss[i:i + step])
            part = (part + sum(x * y for x, y in zip(a, b))) % R
        if d:
            parts = [None] * world
            d.all_gather_object(parts, part)
            total = sum(parts) % R
        else:
            total = part
        want = ctx.fixed_base_mul(args.group, np.frombuffer(total.to_bytes(32, "little"), dtype=np.uint8)).tobytes()
        ok = bool(want == res)
    if rank == 0:
        line = {"group": args.group, "log_n": lg, "n_gpus": world, "ms_best": min(times) * 1e3, "ms_all": [t * 1e3 for t in times],
                "points_per_gpu": m, "sum_identity_ok": ok}
        print(json.dumps(line), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        open("gpurun_out/sharded_msm_N%d.jsonl" % world, "a").write(json.dumps(line) + "\n")
    bases.free(); ctx.free(dks)
if d: dist.destroy_process_group()
