#!/usr/bin/env python3
"""Experiment: latency of one plain-bases G1/G2 MSM vs the number of window groups it is cut into
(b200zk_set_option "msm_parts").  Prints one JSON line per (group, log_n)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_apps_b200 as z

ctx = z.Context(0)
sizes = [(1, int(x)) for x in (sys.argv[1] if len(sys.argv) > 1 else "18,20,22,24").split(",")]
sizes += [(2, int(x)) for x in (sys.argv[2] if len(sys.argv) > 2 else "20").split(",") if x]


def rand(n, seed):
    a = np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x3F
    return a.reshape(-1)


for group, lg in sizes:
    n = 1 << lg
    dks = ctx.alloc(n * 32)
    ctx.upload(dks, rand(n, 1))
    dpts = ctx.alloc(n * (96 if group == 1 else 192))
    ctx.check(z.lib().b200zk_fixed_base_mul_device(ctx.handle, group, dks, n, dpts))
    h = z.VariableBaseMSM.Bases(ctx, group, device_ptr=dpts, n=n, precompute=False)
    ctx.free(dpts)
    ctx.upload(dks, rand(n, 2))
    row, ref = {"group": group, "log_n": lg}, None
    for parts in (1, 2, 3, 4):
        ctx.set_option("msm_parts", parts)
        out, _ = h.msm(device_ptr=dks, n=n)
        ref = ref if ref is not None else bytes(out)
        assert bytes(out) == ref
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); h.msm(device_ptr=dks, n=n); best = min(best, time.perf_counter() - t0)
        row["parts%d_ms" % parts] = round(best * 1e3, 3)
    ctx.set_option("msm_parts", 0)
    print(json.dumps(row), flush=True)
    h.free(); ctx.free(dks)
