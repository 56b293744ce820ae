#!/usr/bin/env python3
"""Groundwork measurement (lever 2): n independent affine additions on the GPU as chunks
sharing one inversion (csrc/ec_batch_affine.cuh, 5M + 1S + inversion/chunk) against the same additions as XYZZ mixed
additions (8M + 2S, what the bucket accumulation does today).  Writes gpurun_out/batch_affine.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_apps_b200 as z

ctx = z.Context(0)
rows = []
for group, lg in ((1, 21), (2, 19)):
    n = 1 << lg
    rng = np.random.default_rng(group)
    def scal():
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 31] &= 0x3F
        return a.reshape(-1)
    p, q = ctx.fixed_base_mul(group, scal()), ctx.fixed_base_mul(group, scal())
    for chunk in (8, 16, 32, 64):
        out, ms_b, ms_x = ctx.batch_add_affine(group, p, q, chunk)
        row = {"group": group, "log_n": lg, "chunk": chunk, "batch_affine_ms": round(ms_b, 3), "xyzz_madd_ms": round(ms_x, 3),
               "batch_affine_adds_per_s": n / (ms_b * 1e-3), "xyzz_madds_per_s": n / (ms_x * 1e-3), "ratio": ms_x / ms_b}
        print(json.dumps(row), flush=True)
        rows.append(row)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"rows": rows}, open("gpurun_out/batch_affine.json", "w"), indent=1)
