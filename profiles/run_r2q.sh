mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "car_ or cfg2 or pyramid" > gpurun_out/r2q_tests.log 2>&1; tail -n 3 gpurun_out/r2q_tests.log
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['pyramid_fused']['ms_per_step'], [(o['op'],o['crop'],o['ms']) for o in d['roofline']['secondary']['per_op']])
P
