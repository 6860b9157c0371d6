mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward_matches_oracle or out_of_range or fuzz" > gpurun_out/r3i_tests.log 2>&1; tail -n 4 gpurun_out/r3i_tests.log
timeout 300 python profiles/g4_sweep.py > gpurun_out/r3i_g4.txt 2>&1; cat gpurun_out/r3i_g4.txt
