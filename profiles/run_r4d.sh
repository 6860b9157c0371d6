mkdir -p gpurun_out
python bench.py > gpurun_out/r4d_bench.json 2> gpurun_out/r4d_bench.err; tail -c 200 gpurun_out/r4d_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4d_bench.json').read().strip().splitlines()[-1])
print('step', d['ms_per_step'], 'value', d['value'], 'frac_step', d['hbm_frac_step'], 'launches', d['gpu_launches'])
print({k:v for k,v in d['roofline'].items() if k not in ('secondary','note')})
print([(o['op'][6:],o['crop'],o['ms'],o['frac']) for o in d['roofline']['secondary']['per_op']])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'fused', d['pyramid_fused']['ms_per_step'], d['pyramid_fused']['mixed_levels']['per_op_ms'], d['pyramid_fused']['mixed_levels']['fused_ms'])
print('cfg4', d['roofline']['secondary']['cfg4'])
P
python -m pytest tests -x -q -m gpu 2>&1 | tail -n 1
