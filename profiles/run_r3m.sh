mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grad_image_matches_oracle or out_of_range or fuzz or pyramid" > gpurun_out/r3m_tests.log 2>&1; tail -n 3 gpurun_out/r3m_tests.log
python profiles/bwd_tma_sweep.py 2>&1 | grep -E "production|full=1 stage   0" 
