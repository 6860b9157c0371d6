mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fuzz" > gpurun_out/r2t_fuzz.log 2>&1; tail -n 5 gpurun_out/r2t_fuzz.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tests/stress_nms.py 24 > gpurun_out/r2t_racecheck.log 2>&1; tail -n 6 gpurun_out/r2t_racecheck.log
timeout 1200 python benchmarks/sweep.py --reps 10 > gpurun_out/sweep_r2.md 2> gpurun_out/sweep_r2.jsonl; tail -n 12 gpurun_out/sweep_r2.md
