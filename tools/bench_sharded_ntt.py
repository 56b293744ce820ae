#!/usr/bin/env python3
"""Fr NTT sweep across GPUs (BASELINE.json config 4): torchrun --nproc-per-node N tools/bench_sharded_ntt.py
Four-step transform with one all_to_all (zk-apps_b200/sharded.py ShardedNTT): rank g holds n / N elements
(its block of columns) resident in HBM; the timed region is column NTTs -> twiddle + transpose ->
all_to_all -> interleave -> row NTTs, bracketed by barriers, max over ranks.  Correctness at full size:
inverse(forward(x)) == x bit for bit, and (log_n <= 22) the forward result equals the single-GPU transform.
Rank 0 prints one JSON line per size and appends it to gpurun_out/sharded_ntt_N<world>.jsonl."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import zk_apps_b200 as z
from zk_apps_b200 import sharded

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,18,20,22,24,26")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
d = None
if world > 1:
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    d = dist
ctx = z.Context(local)
be = sharded.GpuBackend(ctx, d)

for lg in [int(x) for x in args.sizes.split(",") if x]:
    n = 1 << lg
    fwd = sharded.ShardedNTT(be, lg, d)
    inv = fwd.swapped()
    m = n // world
    rng = np.random.default_rng(9000 + lg * 16 + rank)
    x = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
    x[:, 31] &= 0x3F
    buf = torch.from_numpy(x.reshape(-1)).to(be.device)
    keep = buf.clone()
    out = fwd.forward(buf); back = inv.inverse(out); ctx.sync()        # warm-up (tables, NCCL channels) + round trip
    ok = bool(torch.equal(back, keep))
    times = []
    for _ in range(args.reps):
        if d: d.barrier(device_ids=[local])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fwd.forward(buf)
        ctx.sync()
        dt = time.perf_counter() - t0
        if d:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda"); d.all_reduce(t, op=d.ReduceOp.MAX); dt = float(t.item())
        times.append(dt)
        buf = inv.inverse(out)
    if rank == 0:
        best = min(times)
        line = {"log_n": lg, "n_gpus": world, "ms_best": best * 1e3, "ms_all": [t * 1e3 for t in times],
                "elements_per_gpu": m, "alg_GBps": 64.0 * n / best / 1e9, "roundtrip_ok": ok,
                "exchange_bytes_per_gpu": m * 32 * (world - 1) // world}
        print(json.dumps(line), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        open("gpurun_out/sharded_ntt_N%d.jsonl" % world, "a").write(json.dumps(line) + "\n")
    del buf, keep, out, back, fwd, inv
    torch.cuda.empty_cache()
if d: dist.destroy_process_group()
