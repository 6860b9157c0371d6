mkdir -p gpurun_out
python profiles/locality_experiment2.py > gpurun_out/r3u_locality2.txt 2>&1; cat gpurun_out/r3u_locality2.txt
