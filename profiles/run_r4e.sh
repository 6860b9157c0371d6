mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "car_ or cfg2 or pyramid or fuzz or processing_order or autograd" 2>&1 | tail -n 2
for ex in 0 512 0 512; do
BENCH_CAR_EXPERIMENT=$ex timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r4e_bench_$ex.json 2> gpurun_out/r4e_bench.err; tail -c 200 gpurun_out/r4e_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r4e_bench_$ex.json').read().strip().splitlines()[-1])
print('experiment $ex: step', d['ms_per_step'], 'fused', d['pyramid_fused']['ms_per_step'], [(o['op'][6:],o['crop'],o['ms']) for o in d['roofline']['secondary']['per_op']], 'cfg4', d['roofline']['secondary']['cfg4']['ms_per_step'])
P
done
