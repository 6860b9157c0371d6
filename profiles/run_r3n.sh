mkdir -p gpurun_out
timeout 300 python profiles/bwd_tma_sweep.py > gpurun_out/r3n_bwd_tma.txt 2>&1; cat gpurun_out/r3n_bwd_tma.txt
