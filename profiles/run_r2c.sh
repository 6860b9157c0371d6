mkdir -p gpurun_out
timeout 300 python profiles/split_experiment.py > gpurun_out/r2c_split.txt 2>&1
cat gpurun_out/r2c_split.txt
