#!/usr/bin/env python3
"""Turns gpurun_out ncu artefacts into the small text summaries committed under profiles/.

  launches  <launches.csv> <out.txt>        per-kernel totals and shares of an
                                            `ncu --metrics gpu__time_duration.sum` launch list
  kernel    <report.ncu-rep> <out.txt> [json_out]   key counters of an `ncu --set full` capture
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    unit = "ns"
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row.get("Metric Unit", unit)
        k = row["Kernel Name"].replace("<unnamed>::", "").replace("b200zk::", "")
        k = k.split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)
    # key generation (digit tables, window multiples, fixed-base multiplications) runs once per key, outside the timed
    # region: listed apart so that the shares are those of the proof steps
    setup = ("table_multiples", "table_to_affine", "msm_precompute_step", "mark_infinity", "mont_conv_kernel", "fixed_base_kernel",
             "delta_table_kernel", "apply_inf_flags")
    step = {k: v for k, v in agg.items() if not any(t in k for t in setup)}
    once = {k: v for k, v in agg.items() if any(t in k for t in setup)}
    tot = sum(v[1] for v in step.values())
    with open(dst, "w") as f:
        f.write("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache,\n"
                "# serialised launches: compare SHARES with bench.py's kernel_ms_per_step, not absolutes)\n")
        f.write("# source: %s ; %d launches of the proof steps, %.3f ms total\n" % (src, sum(v[0] for v in step.values()), tot * scale))
        f.write("%-64s %8s %12s %7s\n" % ("kernel", "launches", "total_ms", "share"))
        for k, v in sorted(step.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s %8d %12.3f %6.1f%%\n" % (k[:64], v[0], v[1] * scale, 100 * v[1] / tot))
        f.write("\n# key setup, once per proving key, outside the timed region (%d launches, %.1f ms)\n"
                % (sum(v[0] for v in once.values()), sum(v[1] for v in once.values()) * scale))
        for k, v in sorted(once.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s %8d %12.3f\n" % (k[:64], v[0], v[1] * scale))


def kernel(rep, dst, json_out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# key counters from `ncu --set full --clock-control none --import-source on` (%s)\n" % rep)
        for r in data:
            f.write("\n## %s  grid %s block %s\n" % (r[idx["Kernel Name"]][:150], r[idx["Grid Size"]], r[idx["Block Size"]]))
            for k in KEYS:
                if k in idx:
                    f.write("%-84s %14s %s\n" % (k, r[idx[k]], units[idx[k]]))
    if json_out:
        def num(r, k):
            return float(r[idx[k]].replace(",", ""))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per = [num(r, "dram__bytes_read.sum") * mult[units[idx["dram__bytes_read.sum"]]] +
               num(r, "dram__bytes_write.sum") * mult[units[idx["dram__bytes_write.sum"]]] for r in data]
        # grouped by kernel instantiation; the headline figure is the G1 accumulation (the kernel bench.py's roofline is for)
        groups = collections.defaultdict(list)
        for r, b in zip(data, per):
            groups[r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")[:90]].append(b)
        # the dominant kernel: the G1 affine levels when the capture holds them, else the G1 accumulation
        g1 = ([b for k, v in groups.items() if "FqCfg" in k and "msm_affine_level" in k for b in v]
              or [b for k, v in groups.items() if "FqCfg" in k for b in v] or per)
        name = "msm_affine_level<Fq> (G1)" if any("msm_affine_level" in k for k in groups) else "msm_accumulate<Fq> (G1)"
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from bench import kernel_source_hash
        json.dump({"kernel": name, "launches_captured": len(g1),
                   "dram_bytes_per_launch": sum(g1) / len(g1), "per_launch": g1,
                   "by_kernel": {k: {"launches": len(v), "dram_bytes_per_launch": sum(v) / len(v)} for k, v in groups.items()},
                   "source_hash": kernel_source_hash(),
                   "source_hash_note": "hash of csrc/{msm.cu,msm_affine.cuh,ec.cuh,field.cuh,field_asm.cuh,glv.cuh} at summarising time: "
                                       "summarise right after the capture, before touching those files",
                   "capture": "%s (ncu --set full --clock-control none)" % rep}, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(*sys.argv[2:])
