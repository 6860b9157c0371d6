mkdir -p gpurun_out
python profiles/prof_step.py 2 > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'car3d_|zero_fill' -s 4 -c 5 -o gpurun_out/prof_r2_step python profiles/prof_step.py 2 > gpurun_out/ncu_step.log 2>&1
tail -n 2 gpurun_out/ncu_step.log
