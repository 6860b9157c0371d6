mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r2f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r2f_ncu.log 2>&1
tail -n 2 gpurun_out/r2f_ncu.log
