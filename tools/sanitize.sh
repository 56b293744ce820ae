#!/bin/bash
# compute-sanitizer passes over a small workload touching every kernel family:
# memcheck (out-of-bounds / misaligned global, shared and local accesses, leaks of device memory) and racecheck
# (shared-memory hazards: the NTT tiles, the scan, the reduction trees, the cp.async point staging).
# Logs: gpurun_out/sanitize_memcheck.log, gpurun_out/sanitize_racecheck.log (copy the summaries under profiles/).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_workload.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  tail -4 gpurun_out/sanitize_$tool.log
done
