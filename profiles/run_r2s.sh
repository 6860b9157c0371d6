mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2s_tests.log 2>&1; tail -n 3 gpurun_out/r2s_tests.log
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -n 3 gpurun_out/r2s_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], json.dumps(d['pyramid_fused'])[:1500])
print([(o['op'],o['crop'],o['ms'],o['frac']) for o in d['roofline']['secondary']['per_op']])
P
