mkdir -p gpurun_out
python profiles/fwd_threads_ab.py > gpurun_out/r4b_fwd_threads.txt 2>&1; cat gpurun_out/r4b_fwd_threads.txt
