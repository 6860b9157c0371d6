mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grad_image_matches or full_size" 2>&1 | tail -15 > gpurun_out/r2a_tests.log
cat gpurun_out/r2a_tests.log
timeout 300 python profiles/os_sweep.py --cfg4 > gpurun_out/r2a_os_sweep.txt 2>&1
grep -E "os tz|scatter|failed|Error" gpurun_out/r2a_os_sweep.txt | tail -20
