for occ in 2 3 4; do
B200ZK_ACC_BLOCKS=$occ python bench.py --no-extras --no-cpu --steps 3 > gpurun_out/occ$occ.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/occ$occ.json')); k=d['kernel_ms_per_step']; print($occ, round(d['value'],1), round(k['msm_accumulate_g1'],2), round(k['msm_accumulate_g2'],2), round(d['roofline']['int']['frac'],3))"
done
