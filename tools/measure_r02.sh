#!/bin/bash
# round-2 measurement set on one GPU -> gpurun_out/ (summaries are copied under profiles/ by hand afterwards)
mkdir -p gpurun_out
[ "$1" = "notests" ] || timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_tests.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/r02_launches_bench.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"msm_affine_level|msm_accumulate" --launch-count 6 \
  -o gpurun_out/r02_affine -f python bench.py --no-cpu --no-extras --steps 1 --warmup 1 > gpurun_out/r02_ncu_affine.log 2>&1
# summarised here: the reports are too large to travel back (64 MiB limit on gpurun_out)
python tools/ncu_summarise.py kernel gpurun_out/r02_affine.ncu-rep gpurun_out/r02_msm_affine_ncu.txt gpurun_out/r02_msm_affine_traffic.json
ncu -i gpurun_out/r02_affine.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/r02_affine_l0_source.csv 2>/dev/null
python tools/ncu_opcodes.py gpurun_out/r02_affine_l0_source.csv > gpurun_out/r02_msm_affine_l0_opcodes.txt
rm -f gpurun_out/r02_affine.ncu-rep gpurun_out/r02_affine_l0_source.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_pass --launch-skip 4 --launch-count 2 \
  -o gpurun_out/r02_ntt -f python tools/bench_kernels.py --msm "" --msm-g2 "" --ntt 22 > gpurun_out/r02_ncu_ntt.log 2>&1
python tools/ncu_summarise.py kernel gpurun_out/r02_ntt.ncu-rep gpurun_out/r02_ntt_pass_ncu.txt gpurun_out/r02_ntt_pass_traffic.json
rm -f gpurun_out/r02_ntt.ncu-rep
gzip -f gpurun_out/r02_launches.csv
tools/sanitize.sh > gpurun_out/r02_sanitize.log 2>&1
cat gpurun_out/r02_tests.log; tail -3 gpurun_out/r02_bench.err; tail -3 gpurun_out/r02_sanitize.log; ls -la gpurun_out/r02_*
