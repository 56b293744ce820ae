#!/usr/bin/env python3
"""BASELINE.json config 4 on one GPU: G1/G2 MSM and Fr NTT sweeps 2^16..2^26 with operands resident in HBM,
next to the CPU port (oracle/c, all host threads) at the sizes it finishes in seconds.  Every MSM result is
checked at full size with the identity MSM(s, k*G) = (sum s_i k_i mod r) * G; every NTT with iNTT(NTT(x)) = x
on a digest.  Writes gpurun_out/sweep.json (copied to profiles/ by hand)."""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_apps_b200 as z

R = z.ffi.R_MOD
ap = argparse.ArgumentParser()
ap.add_argument("--msm", default="16,18,20,22,24,26")
ap.add_argument("--msm-g2", default="16,18,20,22")
ap.add_argument("--ntt", default="16,18,20,22,24,26")
ap.add_argument("--cpu-msm-max", type=int, default=20)
ap.add_argument("--cpu-ntt-max", type=int, default=22)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="gpurun_out/sweep.json")
args = ap.parse_args()
ctx = z.Context(0)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_int_peaks.json")))
res = {"msm": [], "ntt": [], "host_threads": None}


def rand32(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x3F
    return a


def dot_mod_r(ks, ss):
    tot, step = 0, 1 << 16
    for i in range(0, ks.shape[0], step):
        a = [int.from_bytes(r.tobytes(), "little") for r in ks[i:i + step]]
        b = [int.from_bytes(r.tobytes(), "little") for r in ss[i:i + step]]
        tot = (tot + sum(x * y for x, y in zip(a, b))) % R
    return tot


def ark_ops(lg):                      # the arkworks operation count, in Fq multiplications
    n = 1 << lg
    c = 3 if n < 32 else lg * 69 // 100 + 2
    W = -(-255 // c)
    return (n * W * 11 + W * 2 * ((1 << c) - 1) * 16)


def run_msm(group, lg):
    n = 1 << lg
    pt = 96 if group == 1 else 192
    ks, ss = rand32(n, 100 + lg), rand32(n, 200 + lg)
    dks = ctx.alloc(n * 32); ctx.upload(dks, ks.reshape(-1))
    dpts = ctx.alloc(n * pt)
    ctx.check(z.lib().b200zk_fixed_base_mul_device(ctx.handle, group, dks, n, dpts)); ctx.sync()
    h = z.VariableBaseMSM.Bases(ctx, group, device_ptr=dpts, n=n, precompute=False)
    cpu_ms = None
    if lg <= args.cpu_msm_max:
        from oracle import corac
        pts_host = ctx.download(dpts, n * pt)
        t0 = time.perf_counter(); cpu_out = corac.msm(group, pts_host, ss.reshape(-1)); cpu_ms = (time.perf_counter() - t0) * 1e3
        res["host_threads"] = corac.lib().orc_threads()
    ctx.free(dpts)
    ctx.upload(dks, ss.reshape(-1))
    out, _ = h.msm(device_ptr=dks, n=n)
    best = 1e9
    for _ in range(args.reps):
        ctx.sync(); t0 = time.perf_counter(); out, _ = h.msm(device_ptr=dks, n=n); best = min(best, time.perf_counter() - t0)
    want = ctx.fixed_base_mul(group, np.frombuffer(dot_mod_r(ks, ss).to_bytes(32, "little"), dtype=np.uint8))
    ok = bytes(out) == bytes(want)
    if cpu_ms is not None:
        ok = ok and bytes(cpu_out) == bytes(out)
    ops = ark_ops(lg) * (3 if group == 2 else 1)
    r = {"group": group, "log_n": lg, "gpu_ms": best * 1e3, "cpu_ms": cpu_ms, "identity_ok": bool(ok),
         "ark_fq_mul": ops, "frac_of_fq_mul_peak_on_ark_count": ops / best / peaks["fq_mul_per_s"]}
    print(json.dumps(r), flush=True); res["msm"].append(r)
    h.free(); ctx.free(dks)


def run_ntt(lg):
    n = 1 << lg
    x = rand32(n, 300 + lg).reshape(-1)
    d = ctx.alloc(n * 32); ctx.upload(d, x)
    dom = z.Radix2EvaluationDomain(ctx, lg)
    dom.fft_device(d); dom.fft_device(d, inverse=True); ctx.sync()
    ok = hashlib.sha256(ctx.download(d, n * 32)).digest() == hashlib.sha256(x).digest()
    import torch
    st = torch.cuda.ExternalStream(ctx.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(args.reps):
        e0.record(st); dom.fft_device(d); e1.record(st); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    cpu_ms = None
    if lg <= args.cpu_ntt_max:
        from oracle import corac
        t0 = time.perf_counter(); corac.ntt(x, lg); cpu_ms = (time.perf_counter() - t0) * 1e3
    r = {"log_n": lg, "gpu_ms": best, "cpu_ms": cpu_ms, "roundtrip_ok": bool(ok), "alg_GBps": 64 * n / best / 1e6,
         "frac_of_hbm": 64 * n / best / 1e6 / 6551.4, "frac_of_fr_mul_peak": (n / 2) * lg / (best * 1e-3) / peaks["fr_mul_per_s"]}
    print(json.dumps(r), flush=True); res["ntt"].append(r)
    ctx.free(d)


for lg in [int(v) for v in args.ntt.split(",") if v]: run_ntt(lg)
for lg in [int(v) for v in args.msm.split(",") if v]: run_msm(1, lg)
for lg in [int(v) for v in args.msm_g2.split(",") if v]: run_msm(2, lg)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(res, open(args.out, "w"), indent=1)
