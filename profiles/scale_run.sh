set -x
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale2_n$n.json 2> gpurun_out/scale2_n$n.err
done
tail -c 200 gpurun_out/scale2_n8.err
