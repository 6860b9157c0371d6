mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grad_image_matches_oracle or out_of_range or fuzz" > gpurun_out/r3j_tests.log 2>&1; tail -n 4 gpurun_out/r3j_tests.log
timeout 300 python profiles/bwd_tma_sweep.py > gpurun_out/r3j_bwd_tma.txt 2>&1; cat gpurun_out/r3j_bwd_tma.txt
