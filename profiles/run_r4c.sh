mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward_matches_oracle or fuzz or pyramid or cfg2_full or golden or processing_order" 2>&1 | tail -n 2
python profiles/bwd_threads_ab.py > gpurun_out/r4c_bwd_threads.txt 2>&1; cat gpurun_out/r4c_bwd_threads.txt
