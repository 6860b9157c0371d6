mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r3k_tests.log 2>&1; tail -n 3 gpurun_out/r3k_tests.log
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; tail -c 300 gpurun_out/r3k_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3k_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['hbm_frac_step'], 'fused', d['pyramid_fused']['ms_per_step'], d['pyramid_fused']['mixed_levels'])
print([(o['op'],o['crop'],o['ms'],o['frac']) for o in d['roofline']['secondary']['per_op']])
print(d['roofline']['kernel'], d['roofline']['frac'], 'cfg4', d['roofline']['secondary']['cfg4'])
P
python profiles/prof_step.py 2 > gpurun_out/r3k_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'car3d_grad_image' -s 2 -c 2 -o gpurun_out/prof_r2_tma_bwd python profiles/prof_step.py 3 > gpurun_out/r3k_ncu_step.log 2>&1
tail -n 1 gpurun_out/r3k_ncu_step.log
