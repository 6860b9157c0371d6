mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grad_image or full_size or cfg2" 2>&1 | tail -15 > gpurun_out/r2a_tests.log
python profiles/os_sweep.py --cfg4 > gpurun_out/r2a_os_sweep.txt 2>&1
grep -E "BEST|scatter|failed|Error" gpurun_out/r2a_os_sweep.txt | tail -20
cat gpurun_out/r2a_tests.log
