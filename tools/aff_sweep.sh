#!/bin/bash
# affine-level experiments: correctness first, then the bench step with levels / B variants (gpurun_out/aff_sweep.log)
mkdir -p gpurun_out
L=gpurun_out/aff_sweep.log
: > $L
timeout 900 python -m pytest tests/test_gpu_msm.py -k "affine or digit_table" -x -q 2>&1 | tail -15 >> $L
B200ZK_AFFINE_MIN_ENTRIES=0 timeout 900 python -m pytest tests/test_gpu_groth16.py -x -q 2>&1 | tail -15 >> $L
run() {
  echo "== $*" >> $L
  env "$@" timeout 600 python bench.py --no-cpu --no-extras --steps 5 --warmup 3 2>>$L | python -c '
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(json.dumps({"value":d["value"],"ms":d["ms_per_step"],"e2e":d["e2e"]["value"],"k":d.get("kernel_ms_per_step")}))
' >> $L
}
run B200ZK_AFFINE_LEVELS=4
run B200ZK_AFFINE_LEVELS=0
for v in "$@"; do run $v; done
tail -40 $L
