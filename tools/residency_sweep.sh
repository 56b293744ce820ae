#!/bin/bash
# experiment: registers left free beside the resident accumulate CTAs vs step time
# (168-register accumulate build capped at 2 CTAs/SM by a shared-memory request; bucket reduction CTA width)
for cfg in "2 0 256" "3 100 256" "3 100 64" "3 100 32" "3 70 256"; do
  set -- $cfg
  B200ZK_ACC_BLOCKS=$1 B200ZK_ACC_SMEM_KB=$2 B200ZK_RED_THREADS=$3 python bench.py --no-cpu --no-extras --steps 6 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('acc_build=$1 smem_kb=$2 red_threads=$3', round(d['value'],1), 'proofs/s', round(d['ms_per_step'],2), 'ms', {k: round(v,2) for k,v in d['kernel_ms_per_step'].items() if 'acc' in k or 'reduce' in k})"
done
