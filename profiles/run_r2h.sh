mkdir -p gpurun_out

python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 300 gpurun_out/r2h_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'fused',d['pyramid_fused']['ms_per_step'])
print('e2e',d['e2e'])
print({k:v for k,v in d['cpu_baseline'].items() if k!='sample'})
PY
