mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward_matches_oracle or out_of_range" > gpurun_out/r2j_tests.log 2>&1; tail -n 3 gpurun_out/r2j_tests.log
python profiles/sep_sweep.py fwd > gpurun_out/r2j_sep_fwd.txt 2>&1; tail -n 5 gpurun_out/r2j_sep_fwd.txt
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 600 gpurun_out/r2j_bench.json
