mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-cfg4 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -c 300 gpurun_out/r2w_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
P
python -m pytest tests/test_bench_contract.py -q -m gpu 2>&1 | tail -2
