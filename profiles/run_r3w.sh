mkdir -p gpurun_out
python profiles/locality_experiment3.py > gpurun_out/r3w_locality3.txt 2>&1; cat gpurun_out/r3w_locality3.txt
