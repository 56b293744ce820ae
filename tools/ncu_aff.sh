#!/bin/bash
# ncu --set full capture of the affine-level kernels inside one bench step (gpurun_out/aff.ncu-rep)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:msm_affine_level --launch-skip ${2:-0} --launch-count ${1:-4} \
  -o gpurun_out/aff -f python bench.py --no-cpu --no-extras --steps 1 --warmup 1 > gpurun_out/ncu_aff.log 2>&1
tail -5 gpurun_out/ncu_aff.log
ls -la gpurun_out/aff.ncu-rep
