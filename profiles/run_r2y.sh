mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-cfg4 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -c 300 gpurun_out/r2y_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pcie_gbs_by_rank'])
P
for rep in 1 2; do
echo "== kept-row skip ON"; python profiles/nms_regimes.py 2>&1 | grep -v "^$" | cut -c1-120
echo "== kept-row skip OFF"; NMS_EXPERIMENT=8 python profiles/nms_regimes.py 2>&1 | grep -v "^$" | cut -c1-120
done > gpurun_out/r2y_nms_ab.txt 2>&1; cat gpurun_out/r2y_nms_ab.txt
