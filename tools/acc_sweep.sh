#!/bin/bash
# Sweep of the msm_accumulate build variants on one GPU (one bench.py process per configuration; the
# switches are read once per process).  Output: gpurun_out/acc_sweep.log, one summary line per configuration.
mkdir -p gpurun_out
out=gpurun_out/acc_sweep.log
: > $out
run() {
  echo "== $*" | tee -a $out
  env "$@" python bench.py --steps 6 --warmup 3 --no-cpu --no-extras $BENCH_ARGS 2>&1 | tail -1 | python -c '
import json,sys
d=json.loads(sys.stdin.read())
k=d["kernel_ms_per_step"]
print("step %.2f ms  %.0f proofs/s | acc_g1 %.2f acc_g2 %.2f fold %.2f reduce %.2f fin %.2f sort %.2f ntt %.2f wit %.2f matvec %.2f serial %.2f | int.frac %.3f" % (d["ms_per_step"], d["value"], k["msm_accumulate_g1"], k["msm_accumulate_g2"], k.get("msm_fold",0.0), k["msm_reduce"], k["finalize"], k["msm_sort"], k["ntt_pass"], k["witness"], k["r1cs_matvec"], k["_serialised_step_ms"], d["roofline"]["int"]["frac"]))
' | tee -a $out
}
for cfg in "$@"; do run $cfg; done
