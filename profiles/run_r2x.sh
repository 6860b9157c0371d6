mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "nms or detect or proposal or refine" > gpurun_out/r2x_tests.log 2>&1; tail -n 3 gpurun_out/r2x_tests.log
python tests/stress_nms.py 400 > gpurun_out/r2x_stress.log 2>&1; tail -n 2 gpurun_out/r2x_stress.log
python profiles/nms_regimes.py > gpurun_out/r2x_nms_regimes.txt 2>&1; cat gpurun_out/r2x_nms_regimes.txt
