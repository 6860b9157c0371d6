mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r3d_plain.json 2> gpurun_out/r3d_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_final_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg4 > gpurun_out/r3d_ncu.log 2>&1
tail -n 2 gpurun_out/r3d_ncu.log | cut -c1-200
python profiles/prof_step.py 2 > gpurun_out/r3d_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'car3d_|zero_fill' -s 8 -c 6 -o gpurun_out/prof_r2_final_step python profiles/prof_step.py 3 > gpurun_out/r3d_ncu_step.log 2>&1
tail -n 2 gpurun_out/r3d_ncu_step.log
