mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r3x_bench_n$N.json 2> gpurun_out/r3x_bench_n$N.err
tail -c 300 gpurun_out/r3x_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3x_bench_n$N.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'by rank',d['ms_per_step_by_rank'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['pcie_gbs_by_rank'])
print('cfg4',d['roofline']['secondary']['cfg4'])
PY
