mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/pcie_probe.py > gpurun_out/r2g_pcie_probe_n$N.txt 2>&1
cat gpurun_out/r2g_pcie_probe_n$N.txt | grep -v "^$" | head -40
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err
tail -c 400 gpurun_out/r2g_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_bench_n$N.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])
print('cfg4',d['roofline']['secondary']['cfg4'])
PY
