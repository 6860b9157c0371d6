mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r3l_tests.log 2>&1; tail -n 3 gpurun_out/r3l_tests.log
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err; tail -c 300 gpurun_out/r3l_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3l_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['hbm_frac_step'], 'fused', d['pyramid_fused']['ms_per_step'])
print([(o['op'],o['crop'],o['ms'],o['frac']) for o in d['roofline']['secondary']['per_op']])
print({k:v for k,v in d['roofline'].items() if k not in ('secondary','note')})
P
python profiles/bwd_tma_sweep.py 2>&1 | grep -E "production|full=1 stage   0" 
