mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r3s_tests.log 2>&1; tail -n 3 gpurun_out/r3s_tests.log
( time python bench.py > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err ) 2> gpurun_out/r3s_time.txt; tail -n 3 gpurun_out/r3s_time.txt; tail -c 300 gpurun_out/r3s_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3s_ref.json 2> gpurun_out/r3s_ref.err; tail -c 300 gpurun_out/r3s_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3s_bench.json').read().strip().splitlines()[-1])
print('step', d['ms_per_step'], 'value', d['value'], 'frac_step', d['hbm_frac_step'], 'launches', d['gpu_launches'])
print({k:v for k,v in d['roofline'].items() if k not in ('secondary','note')})
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'fused', d['pyramid_fused']['ms_per_step'], d['pyramid_fused']['mixed_levels'])
print('nms', d['nms3d']['ms'], d['nms3d']['ms_graph_replay'], 'cfg4', d['roofline']['secondary']['cfg4'])
P
