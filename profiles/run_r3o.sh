mkdir -p gpurun_out
python profiles/locality_experiment.py > gpurun_out/r3o_locality.txt 2>&1; cat gpurun_out/r3o_locality.txt
