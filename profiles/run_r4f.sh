mkdir -p gpurun_out
for n in 1 2 3 4; do
python bench.py --op-streams $n --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r4f_bench_$n.json 2> gpurun_out/r4f_bench.err; tail -c 200 gpurun_out/r4f_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r4f_bench_$n.json').read().strip().splitlines()[-1])
print('op-streams $n: step', d['ms_per_step'], d['value'], d['config']['launch'][:90])
P
done
