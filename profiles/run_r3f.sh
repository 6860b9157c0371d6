mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward_matches_oracle or out_of_range or fuzz" > gpurun_out/r3f_tests.log 2>&1; tail -n 3 gpurun_out/r3f_tests.log
python profiles/prof_g4.py 14 24 > gpurun_out/r3f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'g4_kernel' -s 2 -c 1 -o gpurun_out/prof_r2_g4_fwd python profiles/prof_g4.py 14 24 > gpurun_out/r3f_ncu.log 2>&1
tail -n 2 gpurun_out/r3f_ncu.log
