mkdir -p gpurun_out
for i in 1 2; do python -m pytest tests -x -q -m gpu 2>&1 | tail -n 1; done
python tests/stress_nms.py 400 2>&1 | tail -n 1
python bench.py --no-graph --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r4a_eager.json 2> gpurun_out/r4a_eager.err; tail -c 200 gpurun_out/r4a_eager.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4a_eager.json').read().strip().splitlines()[-1])
print('eager step', d['ms_per_step'], 'fused', d['pyramid_fused']['ms_per_step'], d['config']['launch'])
P
python bench.py --workload cfg4 --no-e2e --no-cpu-baseline > gpurun_out/r4a_cfg4.json 2> gpurun_out/r4a_cfg4.err; tail -c 200 gpurun_out/r4a_cfg4.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4a_cfg4.json').read().strip().splitlines()[-1])
print('cfg4 workload step', d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('frac_dram'))
P
