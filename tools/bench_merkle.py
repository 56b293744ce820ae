#!/usr/bin/env python3
"""the note tree on one GPU.  Bulk insertion of 2^k leaves into an empty tree (leaves
resident in HBM; CUDA events on the ctx stream), batched gen_proof, and the per-leaf historical roots, next to the
CPU port (oracle/c Poseidon, all host threads) hashing the same levels.  Roofline: the tree is bound by the
integer pipe -- one node = one Poseidon-t5 permutation = 64 x 25 (dense MDS) + 96 x 3 (S-boxes) = 1,888 Fr
multiplications for 96 B of traffic -- so the fraction reported is of the measured Fr-mul peak
(profiles/r01_int_peaks.json).  Writes gpurun_out/merkle.json."""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import zk_apps_b200 as z

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="10,16,20,22")
ap.add_argument("--cpu-max", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="gpurun_out/merkle.json")
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "profiles", "r01_int_peaks.json")))
FR_MUL_PER_HASH = 64 * 25 + 96 * 3
ctx = z.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream_ptr())
res = {"rows": [], "fr_mul_per_hash": FR_MUL_PER_HASH}


def timed(fn):
    best = 1e9
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.sync(); a.record(stream); fn(); b.record(stream); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for lg in [int(x) for x in args.sizes.split(",")]:
    n = 1 << lg
    rng = np.random.default_rng(lg)
    leaves = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); leaves[:, 31] &= 0x3F
    d_leaves = ctx.alloc(n * 32); ctx.upload(d_leaves, leaves.reshape(-1))
    trees = []

    def build():
        t = z.MerkleTree(ctx, lg + 1, log_roots=False)     # depth lg+1: half full, so gen_proof stays legal
        trees.append(t)
        return t

    # creation (cudaMalloc + memset) is outside the timed call: pre-create one tree per repetition
    pre = [build() for _ in range(args.reps + 1)]
    it = iter(pre)
    pre[0].add_leaves(None, device_ptr=d_leaves, n=n); next(it)   # warm-up
    ms = timed(lambda: next(it).add_leaves(None, device_ptr=d_leaves, n=n))
    t = pre[-1]
    root = t.root()
    row = {"log_leaves": lg, "hashes": n, "build_ms": ms, "hashes_per_s": n / ms * 1e3,
           "frac_of_fr_mul_peak": n * FR_MUL_PER_HASH / (ms * 1e-3) / peaks["fr_mul_per_s"]}
    # batched gen_proof, ids resident
    m = min(n, 1 << 16)
    ids = rng.integers(0, n, size=m, dtype=np.uint64)
    t0 = time.perf_counter(); t.gen_proofs(ids); row["gen_proofs_%d_ms_host" % m] = (time.perf_counter() - t0) * 1e3
    # historical roots of one appended batch (n x depth hashes)
    if lg <= 20:
        t2 = z.MerkleTree(ctx, lg + 1, log_roots=False)
        t0 = time.perf_counter(); _, roots = t2.add_leaves(None, want_roots=True, device_ptr=d_leaves, n=n)
        row["build_with_roots_ms_host"] = (time.perf_counter() - t0) * 1e3
        assert roots[-32:].tobytes() == root
        t2.free()
    if lg <= args.cpu_max:
        from oracle import corac
        t0 = time.perf_counter()
        level = leaves.reshape(-1)
        for _ in range(lg):
            level = corac.poseidon_hash_batch(level, 2)
        top = corac.poseidon_hash_batch(np.concatenate([level, np.zeros(32, dtype=np.uint8)]), 2)
        row["cpu_ms"] = (time.perf_counter() - t0) * 1e3
        row["cpu_threads"] = corac.lib().orc_threads()
        row["root_ok"] = top.tobytes() == root
    print(json.dumps(row), flush=True)
    res["rows"].append(row)
    for tr in trees:
        tr.free()
    ctx.free(d_leaves)

os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump(res, open(args.out, "w"), indent=1)
