mkdir -p gpurun_out
python profiles/prof_step.py 2 > gpurun_out/r3t_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'car3d_|zero_fill' -s 10 -c 8 -o gpurun_out/prof_r2_end_step python profiles/prof_step.py 3 > gpurun_out/r3t_ncu_step.log 2>&1
tail -n 1 gpurun_out/r3t_ncu_step.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r3t_plain.json 2> gpurun_out/r3t_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_end_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r3t_ncu.log 2>&1
tail -n 1 gpurun_out/r3t_ncu.log | cut -c1-120
