#!/bin/bash
# experiment: CTA shape of the bucket reduction / width of the G2 accumulate CTAs vs step time
mkdir -p gpurun_out
for cfg in "256 128" "64 128" "32 128" "64 96" "32 96" "32 64"; do
  set -- $cfg
  B200ZK_RED_THREADS=$1 B200ZK_ACC_THREADS_G2=$2 python bench.py --no-cpu --no-extras --steps 6 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('red=$1 g2thr=$2', round(d['value'],1), 'proofs/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value'],1), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
done
