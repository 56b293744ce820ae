#!/usr/bin/env python3
"""Measures the integer-pipe peaks that MEASURED_PEAKS.json lacks and writes
gpurun_out/int_peaks.json.  kinds: see b200zk_dbg_int_peak in include/b200zk.h."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_apps_b200 as z

ctx = z.Context(0)
names = {0: "imad32_per_s", 1: "imad_wide_per_s", 2: "fr_mul_per_s", 3: "fq_mul_per_s", 4: "dfma_per_s",
         5: "fq_mul_dfma_per_s", 6: "fq_mul_int_plus_dfma_per_s", 7: "fq_mul_warp_split_per_s",
         8: "imad_wide_carry_chained_per_s", 9: "fq_mul_two_chains_per_thread_per_s", 10: "fq28_mul_per_s"}
out = {}
for k, name in names.items():
    out[name] = ctx.int_peak(k)
out["fr_mul_imadwide_equiv_per_s"] = out["fr_mul_per_s"] * 136
out["fq_mul_imadwide_equiv_per_s"] = out["fq_mul_per_s"] * 300
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/int_peaks.json", "w"), indent=1)
print(json.dumps(out))
