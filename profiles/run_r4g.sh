mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grad_image or cfg2_full or cfg4_full or fuzz or pyramid or autograd or processing_order or golden" 2>&1 | tail -n 2
python profiles/bwd_threads_ab.py 2>&1 | grep "256 threads"
