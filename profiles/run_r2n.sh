mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward_matches_oracle or out_of_range" > gpurun_out/r2n_tests.log 2>&1; tail -n 3 gpurun_out/r2n_tests.log
python profiles/sep_sweep.py fwd > gpurun_out/r2n_sep_fwd.txt 2>&1; cat gpurun_out/r2n_sep_fwd.txt
python profiles/prof_sep.py fwd 14 2 6 7 32 > gpurun_out/r2n_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sep_kernel' -s 2 -c 1 -o gpurun_out/prof_r2n_sep_fwd python profiles/prof_sep.py fwd 14 2 6 7 32 > gpurun_out/r2n_ncu.log 2>&1
tail -n 2 gpurun_out/r2n_ncu.log
