mkdir -p gpurun_out
python profiles/l2hint_ab.py > gpurun_out/r3r_l2hint.txt 2>&1; cat gpurun_out/r3r_l2hint.txt
