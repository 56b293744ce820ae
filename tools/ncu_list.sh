#!/bin/bash
# launch list (gpu__time_duration) of the MSM kernels of one bench step: gpurun_out/msm_launches.csv
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:"msm_affine|msm_accumulate|msm_sum_partials" -c 200 --csv \
  --log-file gpurun_out/msm_launches.csv python bench.py --no-cpu --no-extras --steps 1 --warmup 1 > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/msm_launches.csv') if not l.startswith('=='))]
h=rows[0]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); idx=h.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault(r[idx],{'k':r[kn][:60]})[r[mn]]=r[mv]
for i,(k,v) in enumerate(cur.items()):
    if i<60: print(k, v)
PY
