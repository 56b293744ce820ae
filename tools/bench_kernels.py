#!/usr/bin/env python3
"""Kernel-level sweep on one GPU: G1/G2 MSM and Fr NTT with operands resident in HBM, CUDA-event
timing per kernel family (b200zk_prof_*).  Writes gpurun_out/kernels.json."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_apps_b200 as z

ap = argparse.ArgumentParser()
ap.add_argument("--msm", default="16,20,22")
ap.add_argument("--msm-g2", default="16")
ap.add_argument("--ntt", default="16,20,22,24")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--precompute", type=int, default=0)
ap.add_argument("--parts", type=int, default=0, help="msm_parts option (0 = automatic)")
ap.add_argument("--glv", type=int, default=1, help="msm_glv option")
args = ap.parse_args()
ctx = z.Context(0)
ctx.set_option("msm_parts", args.parts)
ctx.set_option("msm_glv", args.glv)
res = {"msm": [], "ntt": []}

def rand_scalars(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 31] &= 0x3F
    return a.reshape(-1)

def run_msm(group, lg):
    n = 1 << lg
    pt = 96 if group == 1 else 192
    ks = rand_scalars(n, 1)
    dks = ctx.alloc(n * 32); ctx.upload(dks, ks)
    dpts = ctx.alloc(n * pt)
    t0 = time.time()
    ctx.check(z.lib().b200zk_fixed_base_mul_device(ctx.handle, group, dks, n, dpts)); ctx.sync()
    t_gen = time.time() - t0
    h = z.VariableBaseMSM.Bases(ctx, group, device_ptr=dpts, n=n, precompute=bool(args.precompute))
    ctx.free(dpts)
    ss = rand_scalars(n, 2)
    ctx.upload(dks, ss)
    h.msm(device_ptr=dks, n=n)          # warm-up (allocations)
    ctx.prof_enable(True); ctx.prof_reset()
    best = 1e9
    for _ in range(args.reps):
        t0 = time.time(); h.msm(device_ptr=dks, n=n); best = min(best, time.time() - t0)
    prof = {k: ctx.prof_get(k) for k in ctx.prof_names()}
    ctx.prof_enable(False)
    r = {"group": group, "log_n": lg, "ms": best * 1e3, "gen_s": t_gen,
         "kernels_ms": {k: v[0] / max(1, v[1]) * (v[1] / args.reps) for k, v in prof.items()}}
    print(json.dumps(r), flush=True)
    res["msm"].append(r)
    h.free(); ctx.free(dks)

def run_ntt(lg):
    n = 1 << lg
    d = ctx.alloc(n * 32); ctx.upload(d, rand_scalars(n, 3))
    dom = z.Radix2EvaluationDomain(ctx, lg)
    dom.fft_device(d); ctx.sync()
    ctx.prof_enable(True); ctx.prof_reset()
    for _ in range(args.reps): dom.fft_device(d)
    ms, cnt = ctx.prof_get("ntt_pass")
    ctx.prof_enable(False)
    per = ms / args.reps
    r = {"log_n": lg, "ms": per, "passes": cnt // args.reps, "GBps_alg": 64 * n / per / 1e6,
         "fr_mul_per_s": (n / 2) * lg / (per * 1e-3)}
    print(json.dumps(r), flush=True)
    res["ntt"].append(r)
    ctx.free(d)

for lg in [int(x) for x in args.ntt.split(",") if x]: run_ntt(lg)
for lg in [int(x) for x in args.msm.split(",") if x]: run_msm(1, lg)
for lg in [int(x) for x in args.msm_g2.split(",") if x]: run_msm(2, lg)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/kernels.json", "w"), indent=1)
