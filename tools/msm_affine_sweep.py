#!/usr/bin/env python3
"""Experiment: one plain-bases G1/G2 MSM with and without the batched-affine levels in front of the bucket accumulation,
over window sizes (B200ZK_MSM_C) and window-group counts.  One JSON line per configuration; every result is compared
with the first one of its size.  usage: msm_affine_sweep.py [g1 log sizes] [g2 log sizes] [windows]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_apps_b200 as z

ctx = z.Context(0)
sizes = [(1, int(x)) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,22,24").split(",") if x]
sizes += [(2, int(x)) for x in (sys.argv[2] if len(sys.argv) > 2 else "").split(",") if x]
cs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]


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
    ref = None
    for c in cs:
        if c:
            os.environ["B200ZK_MSM_C"] = str(c)
        else:
            os.environ.pop("B200ZK_MSM_C", None)
        for levels in (0, 4):
            for parts in (1, 4):
                ctx.set_option("msm_affine_levels", levels)
                ctx.set_option("msm_parts", parts)
                out, _ = h.msm(device_ptr=dks, n=n)
                ref = ref if ref is not None else bytes(out)
                ok = bytes(out) == ref
                best = 1e9
                for _ in range(3):
                    t0 = time.perf_counter(); h.msm(device_ptr=dks, n=n); best = min(best, time.perf_counter() - t0)
                print(json.dumps({"group": group, "log_n": lg, "c": c, "affine_levels": levels, "parts": parts,
                                  "ms": round(best * 1e3, 3), "same_result": ok}), flush=True)
    ctx.set_option("msm_parts", 0); ctx.set_option("msm_affine_levels", 4)
    h.free(); ctx.free(dks)
