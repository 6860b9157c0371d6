mkdir -p gpurun_out
for mode in ready-first forward-first ready-first forward-first; do
BENCH_E2E_ORDER=$mode python bench.py --no-cpu-baseline --no-cfg4 > gpurun_out/r3a_bench_$mode.json 2> gpurun_out/r3a_bench.err; tail -c 200 gpurun_out/r3a_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r3a_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pcie_gbs_by_rank'])
P
done
python profiles/e2e_order.py 2>&1 | grep -E "^(B7 F7 F14 B14|F7 B7 F14 B14)"
