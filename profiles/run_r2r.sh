mkdir -p gpurun_out
python profiles/full_ab.py > gpurun_out/r2r_full_ab.txt 2>&1; cat gpurun_out/r2r_full_ab.txt
