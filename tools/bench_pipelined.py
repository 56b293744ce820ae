#!/usr/bin/env python3
"""Experiment: does proving with TWO contexts on one GPU (one host thread each, out of phase) hide the
latency-bound head (witness, sort) and tail (bucket reduction, assembly) of a batch behind the other
context's bucket accumulation?  include/b200zk.h says a ctx is single-threaded and that callers wanting
concurrency create one ctx per thread; the proving key is plain device memory and is shared.

  python tools/bench_pipelined.py [--batch 128] [--steps 6]

Prints one JSON line per mode: wall-clock proofs/s over `steps` batches of `batch` proofs.  Every mode's
proofs are compared byte for byte with the single-context result on the same inputs.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TOXIC = (0x1f2e3d4c5b6a7988, 0x0123456789abcdef, 0x0fedcba987654321, 0x1122334455667788, 0x99aabbccddeeff00)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    import zk_apps_b200 as z
    from zk_apps_b200.workload import make_update_note_instances
    B, K = args.batch, args.steps
    ctxs = [z.Context(0), z.Context(0)]
    relation = z.UpdateNoteRelation(z.WITHDRAW, 10)
    pk = z.Groth16.generate_parameters_with_toxic_waste(ctxs[0], relation, TOXIC, precompute=True)
    pks = [pk, z.ProvingKey(ctxs[1], relation, pk._h, pk.vk)]
    n_sets = 3
    sets = [make_update_note_instances(ctxs[0], B, 1000 + s, z.WITHDRAW, 10) for s in range(n_sets)]
    row_bytes = sets[0].nbytes // B
    d_sets = []
    for s in sets:
        d = ctxs[0].alloc(s.nbytes)
        ctxs[0].upload(d, s.reshape(-1))
        d_sets.append(d)
    rng = np.random.default_rng(0)
    rb = rng.integers(0, 256, size=(B, 32), dtype=np.uint8); rb[:, 31] &= 0x3F
    sb = rng.integers(0, 256, size=(B, 32), dtype=np.uint8); sb[:, 31] &= 0x3F

    def prove(which, step, lo, hi, out):
        z.Groth16.prove_update_note_device(pks[which], d_sets[step % n_sets] + lo * row_bytes,
                                           np.ascontiguousarray(rb[lo:hi]).reshape(-1),
                                           np.ascontiguousarray(sb[lo:hi]).reshape(-1), hi - lo, out[step][lo * 192:hi * 192])

    def fresh():
        return [np.zeros(B * 192, dtype=np.uint8) for _ in range(K)]

    # warm up both contexts at both batch sizes (scratch buffers grow, tables build)
    w = fresh()
    for c in (0, 1):
        prove(c, 0, 0, B, w)
        prove(c, 0, 0, B // 2, w)
    for c in ctxs:
        c.sync()

    # ---- mode A: one context, K batches back to back
    ref = fresh()
    t0 = time.perf_counter()
    for i in range(K):
        prove(0, i, 0, B, ref)
    ta = time.perf_counter() - t0
    print(json.dumps({"mode": "1 ctx, full batches", "proofs_per_s": B * K / ta, "ms_per_batch": ta / K * 1e3}), flush=True)

    # ---- mode B: two contexts, each proves half of every batch; thread 1 starts half a half-batch late
    def run_threads(work, delay_s):
        ths = [threading.Thread(target=work, args=(c,)) for c in (0, 1)]
        t0 = time.perf_counter()
        ths[0].start()
        time.sleep(delay_s)
        ths[1].start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    for delay in (0.0, 0.25 * ta / K):
        out = fresh()
        tb = run_threads(lambda c: [prove(c, i, c * (B // 2), (c + 1) * (B // 2) if c == 0 else B, out) for i in range(K)], delay)
        same = all(np.array_equal(out[i], ref[i]) for i in range(K))
        print(json.dumps({"mode": "2 ctx, half batches, stagger %.1f ms" % (delay * 1e3), "proofs_per_s": B * K / tb,
                          "ms_per_batch": tb / K * 1e3, "bit_exact_vs_1ctx": same}), flush=True)

    # ---- mode C: two contexts, alternate full batches
    for delay in (0.0, 0.5 * ta / K):
        out = fresh()
        tc = run_threads(lambda c: [prove(c, i, 0, B, out) for i in range(c, K, 2)], delay)
        same = all(np.array_equal(out[i], ref[i]) for i in range(K))
        print(json.dumps({"mode": "2 ctx, alternating full batches, stagger %.1f ms" % (delay * 1e3),
                          "proofs_per_s": B * K / tc, "ms_per_batch": tc / K * 1e3, "bit_exact_vs_1ctx": same}), flush=True)
    pks[1]._h = None            # the second wrapper only borrows the key: pk frees it


if __name__ == "__main__":
    main()
