mkdir -p gpurun_out
python profiles/fill_sweep.py > gpurun_out/r2o_fill.txt 2>&1; cat gpurun_out/r2o_fill.txt
