mkdir -p gpurun_out
N=${1:-8}
if [ "$N" = "1" ]; then python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_buffer_pipeline" 2>&1 | tail -30; exit 0; fi
for mode in uploads-first duplex; do
BENCH_E2E_PIPE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cfg4 > gpurun_out/r3c_bench_n${N}_$mode.json 2> gpurun_out/r3c_bench_n$N.err
tail -c 200 gpurun_out/r3c_bench_n$N.err | grep -v "^\*\|OMP_NUM"
python - <<PY
import json
d=json.loads(open('gpurun_out/r3c_bench_n${N}_$mode.json').read().strip().splitlines()[-1])
print('N=$N $mode: value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['pcie_gbs_by_rank'][0], d['e2e'].get('copy_policy'))
PY
done
